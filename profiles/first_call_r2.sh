#!/bin/bash
# First gpurun call of round 2 (run from the repo root: `gpurun --timeout 1500 -- 'bash profiles/first_call_r2.sh'`).
#   1. first-run parity of the never-executed tcgen05 kernels, each in a child process under a hard timeout:
#      conv_tma.cu (TMA + tcgen05 conv2d fwd / stride-1 dgrad) and conv2d_tc.cu (cp.async gather fwd / dgrad / wgrad)
#   2. the full -m gpu suite WITHOUT -x on the round-1 library 2D path (ESB200_CONV2D=cudnn), then again on the own kernels
#   3. bench A/B: cudnn vs own 2D backend, input vs morton row order
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 240 python tests/conv_tma_child.py --bench > gpurun_out/r2_conv_tma_cases.jsonl 2> gpurun_out/r2_conv_tma_cases.log
echo "conv_tma rc=$? ok=$(grep -c '"ok": true' gpurun_out/r2_conv_tma_cases.jsonl) bad=$(grep -c '"ok": false' gpurun_out/r2_conv_tma_cases.jsonl)"
grep '"ok": false' gpurun_out/r2_conv_tma_cases.jsonl | head -8
tail -3 gpurun_out/r2_conv_tma_cases.log
timeout 240 python tests/conv2d_tc_child.py > gpurun_out/r2_conv2d_cases.jsonl 2> gpurun_out/r2_conv2d_cases.log
echo "conv2d_tc rc=$? ok=$(grep -c '"ok": true' gpurun_out/r2_conv2d_cases.jsonl) bad=$(grep -c '"ok": false' gpurun_out/r2_conv2d_cases.jsonl)"
grep '"ok": false' gpurun_out/r2_conv2d_cases.jsonl | head -8
ESB200_CONV2D=cudnn timeout 900 python -m pytest tests -q -m gpu --junitxml gpurun_out/r2_gpu_tests.xml > gpurun_out/r2_gpu_tests.log 2>&1
echo "pytest(cudnn 2D) rc=$?"; tail -12 gpurun_out/r2_gpu_tests.log
ESB200_CONV2D=cudnn timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_input.json 2> gpurun_out/r2_bench_input.log
ESB200_CONV2D=cudnn timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --row-order morton > gpurun_out/r2_bench_morton.json 2> gpurun_out/r2_bench_morton.log
if ! grep -q '"ok": false' gpurun_out/r2_conv_tma_cases.jsonl && [ -s gpurun_out/r2_conv_tma_cases.jsonl ]; then
  timeout 600 python -m pytest tests -q -m gpu -k "detector or model or golden" > gpurun_out/r2_gpu_tests_own.log 2>&1
  echo "pytest(own 2D) rc=$?"; tail -12 gpurun_out/r2_gpu_tests_own.log
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_own.json 2> gpurun_out/r2_bench_own.log
fi
python - <<'PY'
import json
for tag in ('input', 'morton', 'own'):
    try:
        d = json.loads(open(f'gpurun_out/r2_bench_{tag}.json').read().strip().splitlines()[-1])
        r = d['roofline']
        print(tag, 'value', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'fwd GB/s', round(r['achieved']),
              'wgrad GB/s', round(r['wgrad_achieved_gbs']), 'conv share', round(r['share_of_step'], 3), 'launches', d['gpu_launches'])
    except Exception as e:
        print(tag, 'failed:', e)
PY
export ESB_CUDA_PROFILER_RANGE=1
ESB200_CONV2D=cudnn timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_input.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_input.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_input.csv 2>/dev/null | head -30
