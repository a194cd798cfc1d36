"""Device-side data front-end: depth maps -> world points for all views of a scan in one
launch, replacing LoadDepthFromFile's `/ depth_shift`, ConvertRGBDToPoints + points_img2cam and AggregateMultiViewPoints
(embodiedscan/datasets/transforms/loading.py:70-73, points.py:30-81, structures/bbox_3d/utils.py:335-368,
multiview.py:139-169), followed by the reference's two PointSample stages (points.py:119-153) as seeded permutations.
"""
from typing import Optional, Sequence

import numpy as np
import torch

from ._ffi import call, ptr, query, stream
from .registry import TRANSFORMS


def unproject_multiview(depth_u16: torch.Tensor, intrinsics: Sequence, extrinsics: Sequence, depth_shift: float = 1000.0,
                        return_view: bool = False):
    """depth (V,H,W) integer millimetres on the GPU -> (n_valid, 3) fp32 world points, zero-depth pixels dropped, views
    concatenated in order and pixels row-major (the reference's `nonzero` order). The per-view matrix
    ``E^-1 @ K^-1`` is composed on the host in fp64 and applied as one fp32 4x4 per pixel."""
    assert depth_u16.is_cuda and depth_u16.dim() == 3
    V, H, W = depth_u16.shape
    dev = depth_u16.device
    d16 = depth_u16.to(torch.int32).clamp_(0, 65535).to(torch.int16).contiguous()       # uint16 bit pattern
    mats = []
    for v in range(V):
        K = np.eye(4, dtype=np.float64)
        Kin = np.asarray(intrinsics[v], dtype=np.float64)
        K[:Kin.shape[0], :Kin.shape[1]] = Kin
        E = np.asarray(extrinsics[v], dtype=np.float64).reshape(4, 4)
        mats.append((np.linalg.inv(E) @ np.linalg.inv(K)).astype(np.float32))
    md = torch.from_numpy(np.stack(mats)).to(dev).contiguous()
    out = torch.empty((V * H * W, 3), dtype=torch.float32, device=dev)
    view_of = torch.empty(V * H * W, dtype=torch.int32, device=dev) if return_view else None
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = query('esb_unproject_depth_workspace_bytes', V, H, W)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    call('esb_unproject_depth', ptr(d16), V, H, W, float(depth_shift), ptr(md), ptr(out), ptr(view_of), ptr(cnt), ptr(ws), wsb,
         stream())
    n = int(cnt.item())
    return (out[:n], view_of[:n]) if return_view else out[:n]


@TRANSFORMS.register_module()
class MultiViewDepthToPoints:
    """results['depth_imgs'] (V,H,W) + results['depth2img'] {'intrinsic': [...], 'extrinsic': [...]} -> results['points']
    (n_points, 3): per-view sample of `points_per_view`, then a global sample of `num_points` (seeded permutations; the
    reference draws with np.random.choice)."""

    def __init__(self, num_points: int = 100000, points_per_view: int = 10000, depth_shift: float = 1000.0,
                 seed: Optional[int] = None):
        self.num_points, self.points_per_view, self.depth_shift, self.seed = num_points, points_per_view, depth_shift, seed

    def __call__(self, results: dict) -> dict:
        depth = results['depth_imgs']
        pm = results['depth2img']
        pts, view = unproject_multiview(depth, pm['intrinsic'], pm['extrinsic'], self.depth_shift, return_view=True)
        gen = torch.Generator(device=pts.device)
        if self.seed is not None:
            gen.manual_seed(self.seed)
        keep = []
        for v in range(depth.shape[0]):
            idx = torch.nonzero(view == v).squeeze(1)
            perm = torch.randperm(idx.numel(), generator=gen, device=pts.device)[:self.points_per_view]
            keep.append(idx[perm])
        keep = torch.cat(keep)
        keep = keep[torch.randperm(keep.numel(), generator=gen, device=pts.device)[:self.num_points]]
        results['points'] = pts[keep].contiguous()
        return results
