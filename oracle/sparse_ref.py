"""Oracle (test infrastructure): sparse-tensor engine restating the MinkowskiEngine semantics the hot path relies on
(†upstream ME 0.5.x). Coordinates are numpy int64 (N,4) [b,x,y,z]; features torch CPU tensors so
``autograd`` provides the backward oracle. Slow and exact: sort/unique + explicit gathers, no hashing.

Call sites restated: ME.utils.batch_sparse_collate / ME.SparseTensor at
embodiedscan/models/detectors/sparse_featfusion_single_stage.py:109-118; MinkowskiConvolution / MaxPooling /
InstanceNorm / BatchNorm / GenerativeConvolutionTranspose at embodiedscan/models/backbones/mink_resnet.py:58-69,104-108 and
embodiedscan/models/dense_heads/fcaf3d_head.py:919-946; SparseTensor ``+`` at fcaf3d_head.py:1009.
"""
import numpy as np
import torch

def _pack(c: np.ndarray) -> np.ndarray:
    """Injective int64 key for |xyz| < 2^15, batch < 2^12 (same admissible range as the CUDA packing, wider bias)."""
    return ((c[:, 0].astype(np.int64) << 48) | ((c[:, 1] + 32768).astype(np.int64) << 32) |
            ((c[:, 2] + 32768).astype(np.int64) << 16) | (c[:, 3] + 32768).astype(np.int64))


def voxelize(points: torch.Tensor, voxel_size: float, batch: int) -> np.ndarray:
    """floor(p * fl32(1/voxel_size)): torch's CUDA ``tensor / python_scalar`` multiplies by the fp32 reciprocal, which
    is what the reference's GPU path evaluates for ``p[:, :3] / self.voxel_size`` (sparse_featfusion_single_stage.py:111)
    before ME floors the float coordinates (†upstream batched_coordinates)."""
    inv = np.float32(1.0) / np.float32(voxel_size)
    q = np.floor(points[:, :3].numpy().astype(np.float32) * inv).astype(np.int64)
    return np.concatenate([np.full((q.shape[0], 1), batch, dtype=np.int64), q], 1)


def unique_first(coords: np.ndarray, div: int = 1):
    """Deduplicate with ME-CPU semantics: first occurrence wins, first-occurrence order. Returns (out, in2out)."""
    c = coords.astype(np.int64).copy()
    if div > 1:
        c[:, 1:] = np.floor_divide(c[:, 1:], div) * div
    if c.shape[0] == 0:
        return c, np.zeros((0, ), dtype=np.int64)
    keys = _pack(c)
    _, first, inverse = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind='stable')          # unique ids sorted by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    return c[first[order]], rank[inverse.reshape(-1)]


def offsets(kernel_size: int, scale: int) -> np.ndarray:
    """x-fastest enumeration: k3 -> k = (dx+1)+3(dy+1)+9(dz+1); k2 -> k = dx+2dy+4dz."""
    if kernel_size == 1:
        return np.zeros((1, 3), dtype=np.int64)
    rng = (-1, 0, 1) if kernel_size == 3 else (0, 1)
    return np.array([[dx, dy, dz] for dz in rng for dy in rng for dx in rng], dtype=np.int64) * scale


def kernel_map(in_coords: np.ndarray, out_coords: np.ndarray, offs: np.ndarray) -> np.ndarray:
    """nbr[k, o] = row of (out_coords[o] + offs[k]) in in_coords, or -1."""
    K, n_out = offs.shape[0], out_coords.shape[0]
    nbr = np.full((K, n_out), -1, dtype=np.int64)
    if in_coords.shape[0] == 0 or n_out == 0:
        return nbr
    keys = _pack(in_coords)
    order = np.argsort(keys)
    sk = keys[order]
    for k in range(K):
        q = out_coords.copy()
        q[:, 1:] += offs[k]
        ok = (np.abs(q[:, 1:]) < 32768).all(1)
        qk = _pack(np.where(ok[:, None], q, 0))
        pos = np.searchsorted(sk, qk)
        pos = np.minimum(pos, sk.shape[0] - 1)
        hit = (sk[pos] == qk) & ok
        nbr[k, hit] = order[pos[hit]]
    return nbr


def conv(x: torch.Tensor, W: torch.Tensor, nbr: np.ndarray) -> torch.Tensor:
    """y[o] = sum_k x[nbr[k,o]] @ W[k]  (W: (K,Cin,Cout) or (Cin,Cout))."""
    if W.dim() == 2:
        W = W[None]
    y = x.new_zeros((nbr.shape[1], W.shape[2]))
    for k in range(nbr.shape[0]):
        sel = np.nonzero(nbr[k] >= 0)[0]
        if sel.size == 0:
            continue
        rows = torch.from_numpy(nbr[k, sel])
        y = y.index_add(0, torch.from_numpy(sel), x[rows] @ W[k])
    return y


def maxpool(x: torch.Tensor, nbr: np.ndarray) -> torch.Tensor:
    n_out = nbr.shape[1]
    idx = torch.from_numpy(np.where(nbr >= 0, nbr, 0))                 # (K, n_out)
    g = x[idx]                                                          # (K, n_out, C)
    mask = torch.from_numpy(nbr >= 0)[:, :, None]
    g = torch.where(mask, g, torch.full_like(g, float('-inf')))
    y = g.max(0).values
    return torch.where(torch.isinf(y), torch.zeros_like(y), y) if n_out else x.new_zeros((0, x.shape[1]))


def batch_norm(x, gamma, beta, eps=1e-5):
    """nn.BatchNorm1d over all rows, training statistics (biased variance)."""
    mean = x.mean(0, keepdim=True)
    var = ((x - mean) ** 2).mean(0, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma.view(1, -1) + beta.view(1, -1)


def instance_norm(x, batch_idx: np.ndarray, n_batch: int, gamma, beta, eps=1e-8):
    """MinkowskiInstanceNorm: per (scan, channel) over that scan's rows, eps 1e-8 (†upstream)."""
    out = torch.empty_like(x)
    for b in range(n_batch):
        sel = torch.from_numpy(np.nonzero(batch_idx == b)[0])
        xb = x[sel]
        mean = xb.mean(0, keepdim=True)
        var = ((xb - mean) ** 2).mean(0, keepdim=True)
        out[sel] = (xb - mean) * torch.rsqrt(var + eps) * gamma.view(1, -1) + beta.view(1, -1)
    return out


def generative_children(coords: np.ndarray, half: int) -> np.ndarray:
    """child row = 8*parent + k, k = dx + 2dy + 4dz, coord = parent + d*half."""
    offs = offsets(2, half)
    out = np.repeat(coords, 8, axis=0)
    out[:, 1:] += np.tile(offs, (coords.shape[0], 1))
    return out


def generative_conv(x: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """(N,Cin) x (8,Cin,Cout) -> (8N, Cout), parent-major."""
    return torch.einsum('nc,kcd->nkd', x, W).reshape(-1, W.shape[2])


def union(coords_a: np.ndarray, coords_b: np.ndarray):
    """Rows of A then rows of B absent from A (B order). Returns (coords, map_b)."""
    ka, kb = _pack(coords_a), _pack(coords_b)
    order = np.argsort(ka)
    sk = ka[order]
    pos = np.minimum(np.searchsorted(sk, kb), max(sk.shape[0] - 1, 0))
    hit = (sk[pos] == kb) if sk.shape[0] else np.zeros(kb.shape[0], dtype=bool)
    map_b = np.empty(kb.shape[0], dtype=np.int64)
    map_b[hit] = order[pos[hit]]
    new = ~hit
    map_b[new] = coords_a.shape[0] + np.arange(new.sum())
    return np.concatenate([coords_a, coords_b[new]], 0), map_b


def union_add(fa: torch.Tensor, fb: torch.Tensor, map_b: np.ndarray, n: int) -> torch.Tensor:
    out = torch.cat([fa, fa.new_zeros((n - fa.shape[0], fa.shape[1]))], 0)
    return out.index_add(0, torch.from_numpy(map_b), fb)


def features_at_coordinates(coords: np.ndarray, feats: torch.Tensor, ts: int, query: np.ndarray) -> torch.Tensor:
    """ME SparseTensor.features_at_coordinates (†upstream MinkowskiInterpolation): multilinear interpolation of `feats`
    (rows at `coords`, lattice spacing `ts`) at integer-valued query coordinates [b,x,y,z]; absent lattice points
    contribute 0. Corner order k = dx + 2dy + 4dz, accumulated in that order (weights are exact powers of two here)."""
    q = query.astype(np.int64)
    base = np.floor_divide(q[:, 1:], ts) * ts
    frac = torch.from_numpy((q[:, 1:] - base).astype(np.float32) / np.float32(ts))
    keys = _pack(coords)
    order = np.argsort(keys)
    sk = keys[order]
    out = torch.zeros((q.shape[0], feats.shape[1]), dtype=torch.float32)
    for k in range(8):
        d = np.array([k & 1, (k >> 1) & 1, (k >> 2) & 1], dtype=np.int64)
        w = torch.where(torch.from_numpy(d.astype(bool))[None], frac, 1 - frac).prod(1)
        nb = np.concatenate([q[:, :1], base + d * ts], 1)
        qk = _pack(nb)
        pos = np.minimum(np.searchsorted(sk, qk), sk.shape[0] - 1)
        hit = sk[pos] == qk
        rows = torch.from_numpy(np.where(hit, order[pos], 0))
        out = out + torch.where(torch.from_numpy(hit)[:, None], feats.float()[rows] * w[:, None], torch.zeros(()))
    return out


def prune_mask(scores: torch.Tensor, batch_idx: np.ndarray, n_batch: int, threshold: int) -> np.ndarray:
    """fcaf3d_head.py:1091-1114: per scan keep the top-`threshold` rows by score. torch.topk leaves ties unspecified;
    the frozen rule is: descending score, lowest row index first among equal scores."""
    keep = np.zeros(scores.shape[0], dtype=bool)
    for b in range(n_batch):
        sel = np.nonzero(batch_idx == b)[0]
        sc = scores[torch.from_numpy(sel)].view(-1)
        k = min(len(sel), threshold)
        idx = torch.sort(sc, descending=True, stable=True).indices[:k].numpy()
        keep[sel[idx]] = True
    return keep
