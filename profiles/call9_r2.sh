#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2i_build.log 2>&1
for tag in base nosample nogc both; do
  case $tag in
    base) envs="";;
    nosample) envs="ESB_BENCH_NOSAMPLE=1";;
    nogc) envs="ESB_BENCH_NOGC=1";;
    both) envs="ESB_BENCH_NOGC=1 ESB_BENCH_NOSAMPLE=1";;
  esac
  env $envs timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2i_bench_$tag.json 2> gpurun_out/r2i_bench_$tag.log
  echo "== $tag rc=$?"; grep -E "host ms per timed|clock sample|gc collections" gpurun_out/r2i_bench_$tag.log | cut -c1-400
  python -c "
import json
d=json.loads(open('gpurun_out/r2i_bench_$tag.json').read().strip().splitlines()[-1]); print('$tag value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'launches/step', d['gpu_launches']/d['steps'], d['clocks'])"
done
