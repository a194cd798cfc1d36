#!/bin/bash
# First gpurun call of round 2 (run from the repo root: `gpurun --timeout 1500 -- 'bash profiles/first_call_r2.sh'`).
# Everything written after round 1's GPU budget was spent gets its first run here, in ONE box acquisition:
#   1. the full -m gpu suite WITHOUT -x (so one red blind test does not hide the others), junit + log kept
#   2. the default bench line, then the same with Z-ordered rows (ESB200_ROW_ORDER=morton) for an A/B of `value` and
#      of roofline.achieved / wgrad_achieved_gbs
#   3. a launch list of one step for both row orders (shares, not absolutes)
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1
ESB200_RUN_EXPERIMENTAL=0 timeout 900 python -m pytest tests -q -m gpu --junitxml gpurun_out/r2_gpu_tests.xml > gpurun_out/r2_gpu_tests.log 2>&1
tail -15 gpurun_out/r2_gpu_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_input.json 2> gpurun_out/r2_bench_input.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --row-order morton > gpurun_out/r2_bench_morton.json 2> gpurun_out/r2_bench_morton.log
python - <<'PY'
import json
for tag in ('input', 'morton'):
    try:
        d = json.loads(open(f'gpurun_out/r2_bench_{tag}.json').read().strip().splitlines()[-1])
        r = d['roofline']
        print(tag, 'value', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'fwd GB/s', round(r['achieved']),
              'wgrad GB/s', round(r['wgrad_achieved_gbs']), 'conv share', round(r['share_of_step'], 3))
    except Exception as e:
        print(tag, 'failed:', e)
PY
export ESB_CUDA_PROFILER_RANGE=1
for order in input morton; do
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_${order}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
    --row-order ${order} > gpurun_out/r2_ncu_${order}.log 2>&1
  python profiles/summarize_launches.py gpurun_out/r2_launches_${order}.csv 2>/dev/null | head -25
done
# 4. LAST (a first-run tcgen05 kernel may hang: everything else is already on disk): the experimental conv2d family on
#    its own under a hard timeout (JSON line per case), then the bench with it switched in
timeout 300 python tests/conv2d_tc_child.py > gpurun_out/r2_conv2d_cases.jsonl 2> gpurun_out/r2_conv2d_cases.log
grep -c '"ok": true' gpurun_out/r2_conv2d_cases.jsonl; grep '"ok": false' gpurun_out/r2_conv2d_cases.jsonl | head -5
if ! grep -q '"ok": false' gpurun_out/r2_conv2d_cases.jsonl && [ -s gpurun_out/r2_conv2d_cases.jsonl ]; then
  ESB200_CONV2D=tc timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_conv2d_tc.json 2> gpurun_out/r2_bench_conv2d_tc.log
  tail -c 600 gpurun_out/r2_bench_conv2d_tc.json
fi
