#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2d_build.log 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --torch-profile gpurun_out/r2d_torch_profile.txt > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.log
echo rc=$?; tail -3 gpurun_out/r2d_bench.log | cut -c1-200
timeout 300 python - <<'PY' > gpurun_out/r2d_cprofile.txt 2>&1
import cProfile, pstats, sys, os, torch
sys.path.insert(0, os.getcwd())
from embodiedscan_b200 import MODELS
from embodiedscan_b200.engine import OptimWrapper
from embodiedscan_b200.synth import mv_det3d_config, synth_scan
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = MODELS.build(dict(mv_det3d_config('C2'), compute_dtype=torch.bfloat16)).to(dev).train()
optim = OptimWrapper(model, lr=1e-3, weight_decay=1e-4, max_norm=10.0)
scans = [synth_scan(i, augment=True, device=dev, n_views=20, H=480, W=640, n_points=100000) for i in range(4)]
data = dict(inputs=dict(points=[s['points'] for s in scans], img=[s['img'] for s in scans]), data_samples=[s['data_sample'] for s in scans])
for _ in range(5):
    model.train_step(data, optim)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    model.train_step(data, optim)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(45)
st.sort_stats('cumulative').print_stats(60)
PY
head -70 gpurun_out/r2d_cprofile.txt | cut -c1-170
