#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2r_build.log 2>&1
for ra in 0 1 2; do
  timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --run-ahead $ra > gpurun_out/r2r_bench_ra$ra.json 2> gpurun_out/r2r_bench_ra$ra.log
  echo "== run-ahead $ra rc=$?"; grep -E "host ms per timed" gpurun_out/r2r_bench_ra$ra.log | cut -c1-330
  python -c "
import json
d=json.loads(open('gpurun_out/r2r_bench_ra$ra.json').read().strip().splitlines()[-1]); print('run-ahead $ra value', round(d['value'],2), 'ms', round(d['ms_per_step'],2))"
done
