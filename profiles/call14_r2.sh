#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2n_build.log 2>&1
timeout 300 python -m pytest tests/test_ddp_gpu.py -q > gpurun_out/r2n_ddp_test.log 2>&1
echo "ddp test rc=$?"; tail -2 gpurun_out/r2n_ddp_test.log | cut -c1-200
timeout 200 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.log
echo "n1 rc=$?"
for N in 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+N)) bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.log
  echo "n$N rc=$?"; grep "host ms per timed" gpurun_out/r2n_bench_n$N.log | cut -c1-300
done
python - <<'PY'
import json
v = {}
for n in (1, 4, 8):
    try:
        d = json.loads(open(f'gpurun_out/r2n_bench_n{n}.json').read().strip().splitlines()[-1])
        v[n] = d['value']
        print(n, 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'clocks', d['clocks'])
    except Exception as e:
        print(n, 'failed', e)
for n in (4, 8):
    if 1 in v and n in v:
        print(f'scaling 1->{n}: {v[n] / v[1]:.3f}  efficiency {v[n] / v[1] / n:.3f}')
PY
