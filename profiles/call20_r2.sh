#!/bin/bash
# round 2, call 20: head-weights shape fix — model-level tests, then a short default bench
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_model_gpu.py tests/test_a_golden_gpu.py -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2v_tests.log
timeout 70 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.log; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1]); print('value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'launches', d.get('gpu_launches'))"
