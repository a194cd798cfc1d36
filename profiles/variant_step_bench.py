"""Training-step timing of the sibling configurations on one GPU, secondary to bench.py (whose headline is C2):
  C3  occupancy: 1 scan x 20 views 480x640, 40x40x16 grid, full-width ResNet-50 + FPN + MinkResNet34 + dense 3D neck
  C4  grounding: 12 scans x 20 views 480x480, 100k points, ResNet-50/16 + MinkResNet34 + MinkNeck + 6-layer decoder
Prints one JSON line and, with --table PATH, a torch.profiler kernel table of 2 steps.
    python profiles/variant_step_bench.py --variant C4 --steps 5 --warmup 2 [--table profiles/r1_ground_kernels.txt]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--variant', default='C3')
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--table', default='')
    ap.add_argument('--batch', type=int, default=12, help='scans per step for C4 (cfg grounding :136)')
    args = ap.parse_args()
    from embodiedscan_b200 import MODELS, _ffi
    from embodiedscan_b200.engine import OptimWrapper
    from embodiedscan_b200.synth import (add_grounding_prompt, mv_grounding_config, mv_occ_config, synth_occupancy,
                                         synth_scan)
    dev = torch.device('cuda', 0)
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    torch.manual_seed(0)
    occ = args.variant.startswith('C3')
    small = args.variant.endswith('-small')
    cfg = mv_occ_config(args.variant) if occ else mv_grounding_config(args.variant)
    model = MODELS.build(dict(cfg, compute_dtype=dtype)).to(dev).train()
    if occ:
        optim = OptimWrapper(model, lr=1e-4, weight_decay=1e-2, max_norm=35.0)
        per_step = 1
        va = dict(n_views=2, H=240, W=320, n_points=4000) if small else dict(n_views=20, H=480, W=640, n_points=100000)
    else:
        optim = OptimWrapper(model, lr=5e-4, weight_decay=5e-4, max_norm=10.0)
        per_step = 2 if small else args.batch
        va = dict(n_views=2, H=240, W=320, n_points=2000) if small else dict(n_views=20, H=480, W=480, n_points=100000)
    scans = [synth_scan(10 + j, augment=not occ, device=dev, **va) for j in range(2 * per_step)]
    for j, s in enumerate(scans):
        if occ:
            s['data_sample'].gt_occupancy = synth_occupancy(s['data_sample'], cfg['point_cloud_range'],
                                                            cfg['n_voxels']).to(dev)
        else:
            add_grounding_prompt(s['data_sample'], 1 + j % 3, seed=j)

    def step(j):
        ss = scans[(j % 2) * per_step:(j % 2 + 1) * per_step]
        return model.train_step(dict(inputs=dict(points=[s['points'] for s in ss], img=[s['img'] for s in ss]),
                                     data_samples=[s['data_sample'] for s in ss]), optim)

    for j in range(2 + args.warmup):
        logs = step(j)
    torch.cuda.synchronize()
    _ffi.launch_counter.update(kernels=0, calls=0, by_name={})
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(args.steps):
        logs = step(j)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if args.table:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for j in range(2):
                step(j)
            torch.cuda.synchronize()
        with open(args.table, 'w') as f:
            f.write(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=80))
    print(json.dumps(dict(metric='train_scans_per_sec', workload=f'{args.variant}, {per_step} scan(s)/step', value=1000. * per_step / ms,
                          ms_per_step=ms, dtype=args.dtype, steps=args.steps, warmup=args.warmup,
                          params_m=sum(p.numel() for p in model.parameters()) / 1e6,
                          gpu_launches=_ffi.launch_counter['kernels'],
                          loss={k: float(v) for k, v in logs.items() if k in ('loss', 'loss_cls', 'loss_bbox', 'loss_occ_0')},
                          peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)))


if __name__ == '__main__':
    main()
