#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2t_build.log 2>&1
timeout 200 python -m pytest tests/test_ddp_gpu.py -q > gpurun_out/r2t_ddp.log 2>&1; echo "ddp rc=$?"; tail -1 gpurun_out/r2t_ddp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.log
echo "n2 rc=$?"; grep "host ms per timed" gpurun_out/r2t_bench_n2.log | cut -c1-260
python -c "
import json
d=json.loads(open('gpurun_out/r2t_bench_n2.json').read().strip().splitlines()[-1]); print('N=2 value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), d['config'].get('run_ahead'))"
