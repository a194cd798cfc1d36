#!/bin/bash
# round 2, call 19: fused head epilogue — full GPU suite, then default bench vs ESB200_HEAD_EPILOGUE=lib
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2u_build.log 2>&1
timeout 240 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2u_tests.log
timeout 120 python bench.py --steps 40 --warmup 5 > gpurun_out/r2u_bench_own.json 2> gpurun_out/r2u_bench_own.log; echo "own rc=$?"
ESB200_HEAD_EPILOGUE=lib timeout 120 python bench.py --steps 40 --warmup 5 > gpurun_out/r2u_bench_lib.json 2> gpurun_out/r2u_bench_lib.log; echo "lib rc=$?"
python - <<'PY'
import json
for k in ('own', 'lib'):
    try:
        d = json.loads(open(f'gpurun_out/r2u_bench_{k}.json').read().strip().splitlines()[-1])
        print(k, 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(k, 'unreadable', e)
PY
