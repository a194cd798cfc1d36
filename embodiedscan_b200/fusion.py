"""Point painting host side: packs the per-scan image meta into the kernel's ``EsbPaintMeta``
records and exposes the fused project+gather+view-mean kernel as an autograd function.

Mirrors ``batch_point_sample`` / ``apply_3d_transformation``
(embodiedscan/models/layers/fusion_layers/point_fusion.py:208-311, :20-107) and the caller loop at
embodiedscan/models/detectors/sparse_featfusion_single_stage.py:142-207, for all scans of the batch and one
pyramid level per launch instead of one Python iteration per (scan, level).
"""
import struct
from typing import List, Sequence

import numpy as np
import torch

from . import _ffi
from ._ffi import call, ptr, stream

OP_CODE = {'T': 0, 'S': 1, 'R': 2, 'HF': 3, 'VF': 4}
MAX_OPS = 8


def compose_projection(intrinsic, extrinsic) -> np.ndarray:
    """fp32 ``intrinsic @ extrinsic`` with a fixed summation order (k = 0..3, no FMA), so host and oracle agree."""
    a = np.asarray(intrinsic, dtype=np.float32).reshape(4, 4)
    b = np.asarray(extrinsic, dtype=np.float32).reshape(4, 4)
    out = np.zeros((4, 4), dtype=np.float32)
    for i in range(4):
        for j in range(4):
            acc = np.float32(0.)
            for k in range(4):
                acc = np.float32(acc + np.float32(a[i, k] * b[k, j]))
            out[i, j] = acc
    return out


def reverse_flow_ops(img_meta: dict):
    """The reversed 3D augmentation flow as (op, 9 params) records (point_fusion.py:44-105)."""
    flow = list(img_meta.get('transformation_3d_flow', []))[::-1]
    ops = []
    for op in flow:
        p = [0.] * 9
        if op == 'T':
            t = -np.asarray(img_meta.get('pcd_trans', [0., 0., 0.]), dtype=np.float32).reshape(-1)
            p[:3] = [float(v) for v in t]
        elif op == 'S':
            p[0] = float(np.float32(1.0 / img_meta.get('pcd_scale_factor', 1.)))
        elif op == 'R':
            rot = torch.as_tensor(np.asarray(img_meta['pcd_rotation'], dtype=np.float32)) if 'pcd_rotation' in img_meta \
                else torch.eye(3)
            inv = rot.inverse().contiguous().view(-1)        # fp32 LU on the host, as the reference does
            p = [float(v) for v in inv]
        elif op == 'HF':
            if not img_meta.get('pcd_horizontal_flip', False):
                continue
        elif op == 'VF':
            if not img_meta.get('pcd_vertical_flip', False):
                continue
        else:
            raise AssertionError(f'This 3D data transformation op ({op}) is not supported')
        ops.append((OP_CODE[op], p))
    assert len(ops) <= MAX_OPS
    return ops


def pack_paint_metas(img_metas: Sequence[dict], device) -> torch.Tensor:
    """One EsbPaintMeta (csrc/paint.cu) per scan -> uint8 device tensor."""
    rec = bytearray()
    for m in img_metas:
        sf = m.get('scale_factor', (1., 1.))
        sx, sy = float(np.float32(sf[0])), float(np.float32(sf[1]))
        co = m.get('img_crop_offset', (0., 0.))
        ox, oy = float(np.float32(co[0])), float(np.float32(co[1]))
        flip = 1 if m.get('flip', False) else 0
        ori_w = float(m['img_shape'][1])
        ops = reverse_flow_ops(m)
        rec += struct.pack('<5fii', sx, sy, ox, oy, ori_w, flip, len(ops))
        rec += struct.pack(f'<{MAX_OPS}i', *([o[0] for o in ops] + [0] * (MAX_OPS - len(ops))))
        for i in range(MAX_OPS):
            rec += struct.pack('<9f', *(ops[i][1] if i < len(ops) else [0.] * 9))
    nbytes = _ffi.query('esb_paint_meta_bytes')
    assert len(rec) == nbytes * len(img_metas), (len(rec), nbytes)
    return torch.frombuffer(rec, dtype=torch.uint8).clone().to(device)


def pack_projections(img_metas: Sequence[dict], coord_type: str, device) -> torch.Tensor:
    """(B, V, 4, 4) fp32 ``intrinsic[v] @ extrinsic[v]`` (sparse_featfusion_single_stage.py:152-164)."""
    key = {'LIDAR': 'lidar2img', 'DEPTH': 'depth2img', 'CAMERA': 'cam2img'}[coord_type.upper()]
    mats = []
    for m in img_metas:
        pm = m[key]
        assert isinstance(pm, dict) and 'extrinsic' in pm and 'intrinsic' in pm
        intr = pm['intrinsic']
        if not isinstance(intr, (list, tuple)):
            intr = [intr] * len(pm['extrinsic'])
        a = np.stack([np.asarray(i, dtype=np.float32).reshape(4, 4) for i in intr])
        b = np.stack([np.asarray(e, dtype=np.float32).reshape(4, 4) for e in pm['extrinsic']])
        acc = np.zeros((a.shape[0], 4, 4), dtype=np.float32)
        for k in range(4):        # same order as compose_projection: acc = fl(acc + fl(a[i,k] * b[k,j])), k = 0..3
            acc = (acc + (a[:, :, k:k + 1] * b[:, k:k + 1, :]).astype(np.float32)).astype(np.float32)
        mats.append(acc)
    return torch.from_numpy(np.stack(mats)).to(device)


class _Paint(torch.autograd.Function):

    @staticmethod
    def forward(ctx, feat, coords, fpts, fbatch, metas, proj, voxel_size, pad_hw, V):
        """feat (B*V, C, Hf, Wf) in channels_last memory; coords (N,4) int32 voxel rows OR fpts (N,3) fp32 -> (N, C)."""
        assert feat.is_contiguous(memory_format=torch.channels_last) or feat.shape[1] == 1
        BV, C, Hf, Wf = feat.shape
        N = coords.shape[0] if coords is not None else fpts.shape[0]
        out = torch.empty((N, C), dtype=feat.dtype, device=feat.device)
        call('esb_paint_fwd', ptr(coords), ptr(fpts), ptr(fbatch), N, voxel_size, ptr(metas), ptr(proj), V, ptr(feat), Hf, Wf,
             C, float(pad_hw[0]), float(pad_hw[1]), ptr(out), None, _ffi.dtype_code(feat.dtype), stream())
        ctx.pts = (coords, fpts, fbatch, metas, proj)
        ctx.meta = (voxel_size, pad_hw, V, feat.shape, feat.dtype, N)
        return out

    @staticmethod
    def backward(ctx, dout):
        coords, fpts, fbatch, metas, proj = ctx.pts
        voxel_size, pad_hw, V, shape, dtype, N = ctx.meta
        if not ctx.needs_input_grad[0]:
            return (None, ) * 9
        BV, C, Hf, Wf = shape
        dout = dout.contiguous()
        dfeat = torch.zeros((BV, Hf, Wf, C), dtype=torch.float32, device=dout.device)
        call('esb_paint_bwd', ptr(coords), ptr(fpts), ptr(fbatch), N, voxel_size, ptr(metas), ptr(proj), V, ptr(dout), Hf, Wf,
             C, float(pad_hw[0]), float(pad_hw[1]), ptr(dfeat), _ffi.dtype_code(dout.dtype), stream())
        return (dfeat.permute(0, 3, 1, 2).to(dtype), ) + (None, ) * 8


def paint_points(feat: torch.Tensor, coords: torch.Tensor, metas: torch.Tensor, proj: torch.Tensor, voxel_size: float,
                 pad_hw, n_views: int) -> torch.Tensor:
    """Image features (B*V, C, Hf, Wf) sampled at the voxel centres ``coords[:,1:] * voxel_size`` of every scan."""
    if not feat.is_contiguous(memory_format=torch.channels_last):
        feat = feat.contiguous(memory_format=torch.channels_last)
    return _Paint.apply(feat, coords, None, None, metas, proj, float(np.float32(voxel_size)), pad_hw, n_views)


def paint_float_points(feat: torch.Tensor, points: torch.Tensor, batch: torch.Tensor, metas: torch.Tensor,
                       proj: torch.Tensor, pad_hw, n_views: int) -> torch.Tensor:
    """Same kernel on explicit fp32 locations ``points`` (N,3) of scans ``batch`` (N) int32 (None: scan 0)."""
    if not feat.is_contiguous(memory_format=torch.channels_last):
        feat = feat.contiguous(memory_format=torch.channels_last)
    return _Paint.apply(feat, None, points.float().contiguous(), batch, metas, proj, 1.0, pad_hw, n_views)
