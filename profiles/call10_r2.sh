#!/bin/bash
# 8 GPUs of one box: NCCL gradient test, weak scaling 1 / 2 / 4 / 8 with the per-rank torch profile at N = 8
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2j_build.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m pytest tests/test_ddp_gpu.py -q > gpurun_out/r2j_ddp_test.log 2>&1
echo "ddp test rc=$?"; tail -3 gpurun_out/r2j_ddp_test.log | cut -c1-300
timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.log
echo "n1 rc=$?"
for N in 2 4 8; do
  extra=""
  if [ $N -eq 8 ]; then extra="--torch-profile gpurun_out/r2_n8_torch_profile.txt"; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 30 --warmup 5 $extra > gpurun_out/r2j_bench_n$N.json 2> gpurun_out/r2j_bench_n$N.log
  echo "n$N rc=$?"; grep "host ms per timed" gpurun_out/r2j_bench_n$N.log | cut -c1-260
done
python - <<'PY'
import json
v = {}
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f'gpurun_out/r2j_bench_n{n}.json').read().strip().splitlines()[-1])
        v[n] = d['value']
        print(n, 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'clocks', d['clocks'])
    except Exception as e:
        print(n, 'failed', e)
for n in (2, 4, 8):
    if 1 in v and n in v:
        print(f'scaling 1->{n}: {v[n] / v[1]:.3f}  efficiency {v[n] / v[1] / n:.3f}')
PY
ls gpurun_out/r2_n8_torch_profile.txt.rank* 2>/dev/null | head -3
