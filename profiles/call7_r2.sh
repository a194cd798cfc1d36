#!/bin/bash
# 2 GPUs: NCCL gradient-equality test of the bucketed reducer + weak-scaling check N=1 vs N=2 on the same box
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2g_build.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_ddp_gpu.py -q > gpurun_out/r2g_ddp_test.log 2>&1
echo "ddp test rc=$?"; tail -5 gpurun_out/r2g_ddp_test.log | cut -c1-300
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.log
echo "n1 rc=$?"; grep "host ms" gpurun_out/r2g_bench_n1.log | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.log
echo "n2 rc=$?"; grep "host ms" gpurun_out/r2g_bench_n2.log | cut -c1-200; tail -3 gpurun_out/r2g_bench_n2.log | cut -c1-200
python - <<'PY'
import json
v = {}
for n in (1, 2):
    try:
        d = json.loads(open(f'gpurun_out/r2g_bench_n{n}.json').read().strip().splitlines()[-1])
        v[n] = d['value']
        print(n, 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2))
    except Exception as e:
        print(n, 'failed', e)
if 1 in v and 2 in v:
    print('scaling 1->2:', round(v[2] / v[1], 3))
PY
