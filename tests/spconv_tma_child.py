"""Child process: the TMA-fed sparse-conv kernels (csrc/spconv_tma.cu: gather4 rows, tiled filter boxes, M = 256 tiles)
against the cp.async kernels of csrc/spconv_tc.cu on the same maps and bf16 operands (forward, dgrad, wgrad), one JSON line per
case; `--bench` adds timings at a C2-sized level. A first-run TMA kernel that hangs must not take the session with it."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def rand_coords(n, extent, batch, seed):
    g = np.random.default_rng(seed)
    c = np.concatenate([g.integers(0, batch, (n, 1)), g.integers(-extent, extent, (n, 3))], 1)
    return np.unique(c, axis=0)


def run(cin, cout, ksize, stride, n, extent, batch, bench=False):
    from embodiedscan_b200 import sparse as SP
    dev = torch.device('cuda:0')
    torch.manual_seed(cin + cout)
    c = rand_coords(n, extent, batch, cin * 7 + cout)
    mgr = SP.CoordinateManager(dev)
    key = mgr.insert_unique(torch.from_numpy(c).to(dev, torch.int32), 1)
    mgr.batch_size = batch
    conv = SP.MinkowskiConvolution(cin, cout, kernel_size=ksize, stride=stride).to(dev)
    x = torch.randn(c.shape[0], cin, device=dev).bfloat16()
    out = {}
    for backend in ('tc', 'tma'):
        os.environ['ESB200_SPCONV'] = backend
        conv.kernel.grad = None
        xd = x.clone().requires_grad_(True)
        y = conv(SP.SparseTensor(xd, coordinate_map_key=key, coordinate_manager=mgr))
        gy = torch.randn(y.F.shape, device=dev, generator=torch.Generator(dev).manual_seed(1)).bfloat16()
        y.F.backward(gy)
        torch.cuda.synchronize()
        out[backend] = (y.F.float(), xd.grad.float(), conv.kernel.grad.float().clone())
    res = dict(kind='parity', case=[cin, cout, ksize, stride, int(c.shape[0])])
    ok = True
    for name, a, b in zip(('fwd', 'dgrad', 'wgrad'), out['tc'], out['tma']):
        err = float((a - b).abs().max())
        tol = 2e-2 * max(float(a.abs().max()), 1e-6)
        res[name] = err
        ok = ok and err <= tol and bool(torch.isfinite(b).all())
    res['ok'] = ok
    print(json.dumps(res), flush=True)
    if not bench:
        return
    kmap = mgr.kernel_map(key, mgr.stride_key(key, stride) if stride > 1 else key, ksize)
    P = int(kmap.pairs[2][-1])
    byt = P * (cin + cout) * 2 + 8 * P + ksize ** 3 * cin * cout * 2
    for backend in ('tc', 'tma'):
        os.environ['ESB200_SPCONV'] = backend
        t = {}
        for phase in ('fwd', 'bwd'):
            ms = []
            for _ in range(6):
                xd = x.clone().requires_grad_(True)
                conv.kernel.grad = None
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if phase == 'fwd':
                    e0.record()
                    y = conv(SP.SparseTensor(xd, coordinate_map_key=key, coordinate_manager=mgr))
                    e1.record()
                else:
                    y = conv(SP.SparseTensor(xd, coordinate_map_key=key, coordinate_manager=mgr))
                    gy = torch.ones_like(y.F)
                    e0.record()
                    y.F.backward(gy)
                    e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            t[phase] = sorted(ms)[len(ms) // 2]
        print(json.dumps(dict(kind='bench', backend=backend, case=[cin, cout, ksize, stride, int(c.shape[0])], pairs=P,
                              fwd_us=1e3 * t['fwd'], bwd_us=1e3 * t['bwd'], fwd_gbs=byt / t['fwd'] / 1e6,
                              bwd_gbs=2 * byt / t['bwd'] / 1e6)), flush=True)


CASES = [
    (64, 64, 3, 1, 6000, 9, 2), (128, 128, 3, 1, 6000, 9, 2), (64, 128, 3, 2, 6000, 9, 2), (256, 512, 3, 2, 5000, 8, 2),
    (512, 512, 3, 1, 3000, 6, 2), (1024, 128, 3, 1, 3000, 6, 2), (64, 128, 1, 2, 6000, 9, 2), (192, 64, 3, 1, 6000, 9, 2),
    (64, 64, 3, 1, 60000, 24, 4),          # MT = 2 path (enough row tiles)
]
BENCH = [(64, 64, 3, 1, 400000, 60, 4), (128, 128, 3, 1, 120000, 40, 4), (256, 256, 3, 1, 40000, 28, 4),
         (512, 512, 3, 1, 12000, 18, 4)]

if __name__ == '__main__':
    for case in CASES:
        try:
            run(*case)
        except Exception as e:  # noqa
            print(json.dumps(dict(kind='error', case=list(case[:4]), ok=False, err=str(e)[:300])), flush=True)
    if '--bench' in sys.argv:
        for case in BENCH:
            try:
                run(*case, bench=True)
            except Exception as e:  # noqa
                print(json.dumps(dict(kind='error', case=list(case[:4]), ok=False, err=str(e)[:300])), flush=True)
