"""Oracle (test infrastructure): the data front-end either side of the model — image collation and depth unprojection.

Restates
  * embodiedscan/models/data_preprocessors/data_preprocessor.py:249-264 (preprocess_img), :266-339 (collate_data,
    list-of-multi-view branch) and models/data_preprocessors/utils.py:9-63 (multiview_img_stack_batch)
  * embodiedscan/datasets/transforms/loading.py:70-73 (depth / depth_shift), transforms/points.py:30-81
    (ConvertRGBDToPoints), structures/bbox_3d/utils.py:335-368 (points_img2cam), transforms/multiview.py:139-169
    (AggregateMultiViewPoints: world = solve(extrinsic, p))
Pinned by tests/golden/frontend.npz (the reference's own code run by tests/golden/make_golden.py).
"""
import math
from typing import List, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def preprocess_multiview(imgs: List[torch.Tensor], mean, std, bgr_to_rgb=True, divisor=32) -> torch.Tensor:
    """list[B] of (V,3,H_b,W_b) uint8 -> (B,V,3,Hp,Wp) fp32: every scan normalised, then right/bottom zero padded to the
    batch maximum rounded up to `divisor`."""
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    norm = [((i[:, [2, 1, 0]] if bgr_to_rgb else i).float() - m) / s for i in imgs]
    Hp = int(math.ceil(max(i.shape[-2] for i in norm) / divisor) * divisor)
    Wp = int(math.ceil(max(i.shape[-1] for i in norm) / divisor) * divisor)
    return torch.stack([F.pad(i, (0, Wp - i.shape[-1], 0, Hp - i.shape[-2])) for i in norm])


def pad_shapes(imgs: List[torch.Tensor], divisor=32):
    """data_preprocessor.py `_get_pad_shape`: per scan, from its own first view."""
    return [(int(np.ceil(i.shape[-2] / divisor)) * divisor, int(np.ceil(i.shape[-1] / divisor)) * divisor) for i in imgs]


def unproject_depth(depth: torch.Tensor, intrinsics: Sequence, extrinsics: Sequence, depth_shift: float = 1000.0):
    """depth (V,H,W) integer millimetres -> (world points (n,3) fp32 in view order / row-major pixel order, counts (V,))."""
    pts, counts = [], []
    V, H, W = depth.shape
    for v in range(V):
        d = depth[v].numpy().astype(np.float32) / np.float32(depth_shift)
        us, vs = np.meshgrid(np.arange(W), np.arange(H))
        grid = torch.from_numpy(np.stack([us.astype(np.float32), vs.astype(np.float32), d], -1).reshape(-1, 3))
        nz = torch.from_numpy(d.reshape(-1).nonzero()[0])
        K = torch.eye(4, dtype=torch.float32)
        k = torch.as_tensor(np.asarray(intrinsics[v], dtype=np.float32))
        K[:k.shape[0], :k.shape[1]] = k
        unnormed = torch.cat([grid[:, :2] * grid[:, 2:3], grid[:, 2:3]], 1)
        homo = torch.cat([unnormed, torch.ones(unnormed.shape[0], 1)], 1)
        cam = torch.mm(homo, torch.inverse(K).t())[:, :3][nz]
        E = torch.from_numpy(np.asarray(extrinsics[v], dtype=np.float32).reshape(4, 4))
        p4 = torch.cat([cam, torch.ones(cam.shape[0], 1)], 1)
        pts.append(torch.linalg.solve(E, p4.t()).t()[:, :3])
        counts.append(cam.shape[0])
    return torch.cat(pts), np.asarray(counts)
