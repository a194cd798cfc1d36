"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Pinned by tests/golden/occupancy_g3.npz and continuous_occ.npz
(the reference's own DenseFusionOccPredictor / EmbodiedOccPredictor, neck, head and losses executed by
tests/golden/make_golden.py); the MinkowskiEngine / mmdet arithmetic underneath stays unpinned (not installable here).

CPU restatement of the occupancy path, evaluated functionally from the product's state_dict:
  embodiedscan/models/detectors/dense_fusion_occ.py:101-265 (extract_feat), :267-295 (loss)
  embodiedscan/models/necks/imvoxel_neck.py:8-143
  embodiedscan/models/dense_heads/imvoxel_occ_head.py:59-184
  embodiedscan/models/losses/occ_loss.py:7-141 (class loops kept literal)
  embodiedscan/models/task_modules/anchor/anchor_3d_generator.py:292-354
  mmdet FPN (†upstream): 1x1 laterals, nearest top-down, 3x3 output convs
"""
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import model_ref as M
from . import sparse_ref as S


def prior_points(ranges, n_voxels) -> torch.Tensor:
    """Voxel centres, z slowest / x fastest (anchors_single_range + the permute in grid_anchors)."""
    r = torch.tensor(ranges, dtype=torch.float32)
    nx, ny, nz = n_voxels
    axes = []
    for lo, hi, n in ((r[0], r[3], nx), (r[1], r[4], ny), (r[2], r[5], nz)):
        c = torch.linspace(lo, hi, n + 1)
        c = c + (c[1] - c[0]) / 2
        axes.append(c[:n])
    out = torch.zeros((nz, ny, nx, 3))
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                out[k, j, i, 0], out[k, j, i, 1], out[k, j, i, 2] = axes[0][i], axes[1][j], axes[2][k]
    return out.reshape(-1, 3)


def fpn(sd, prefix, feats: List[torch.Tensor]) -> List[torch.Tensor]:
    lat = [F.conv2d(x, sd[f'{prefix}lateral_convs.{i}.conv.weight'], sd[f'{prefix}lateral_convs.{i}.conv.bias'])
           for i, x in enumerate(feats)]
    for i in range(len(lat) - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
    return [F.conv2d(x, sd[f'{prefix}fpn_convs.{i}.conv.weight'], sd[f'{prefix}fpn_convs.{i}.conv.bias'], padding=1)
            for i, x in enumerate(lat)]


def _bn3(sd, p, x, training):
    return F.batch_norm(x, sd[p + '.running_mean'].clone(), sd[p + '.running_var'].clone(), sd[p + '.weight'],
                        sd[p + '.bias'], training, 0.1, 1e-5)


def _res_module(sd, p, x, stride, training):
    out = F.relu(_bn3(sd, p + 'norm1', F.conv3d(x, sd[p + 'conv1.weight'], None, stride, 1), training))
    out = _bn3(sd, p + 'norm2', F.conv3d(out, sd[p + 'conv2.weight'], None, 1, 1), training)
    idt = x
    if stride != 1:
        idt = _bn3(sd, p + 'downsample.1', F.conv3d(x, sd[p + 'downsample.0.weight'], None, stride), training)
    return F.relu(out + idt)


def imvoxel_neck(sd, prefix, x, n_blocks, training):
    downs = []
    for i, nb in enumerate(n_blocks):
        for j in range(nb):
            x = _res_module(sd, f'{prefix}down_layer_{i}.{j}.', x, 2 if (i > 0 and j == 0) else 1, training)
        downs.append(x)
    outs = []
    for i in range(len(n_blocks) - 1, -1, -1):
        if i < len(n_blocks) - 1:
            p = f'{prefix}up_block_{i + 1}.'
            x = F.relu(_bn3(sd, p + '1', F.conv_transpose3d(x, sd[p + '0.weight'], None, 2), training))
            x = F.relu(_bn3(sd, p + '4', F.conv3d(x, sd[p + '3.weight'], None, 1, 1), training))
            x = downs[i] + x
        p = f'{prefix}out_block_{i}.'
        outs.append(F.relu(_bn3(sd, p + '1', F.conv3d(x, sd[p + '0.weight'], None, 1, 1), training)))
    return outs[::-1]


def extract_feat(sd, cfg, points: List[torch.Tensor], imgs: torch.Tensor, img_metas: List[dict], training: bool,
                 continuous: bool = False):
    """dense_fusion_occ.py:120-259; continuous=True restates embodied_occ.py:118-247 (one scan, `points[b]` = frames
    0..b, sample b painted from views 0..b)."""
    B, V = imgs.shape[:2]
    if continuous:
        B = len(points)
    n_voxels = list(cfg['n_voxels'])
    pr = cfg['prior_generator']['ranges'][0]
    f2d = M.resnet2d(sd, 'backbone.', cfg['backbone']['depth'], imgs.reshape((-1, ) + tuple(imgs.shape[2:])))
    f0 = fpn(sd, 'neck.', f2d)[0]
    f0 = f0.reshape((imgs.shape[0], V) + tuple(f0.shape[1:]))
    prior = prior_points(pr, n_voxels)
    pad_hw = tuple(imgs.shape[-2:])
    vols = []
    for b in range(B):
        pm = img_metas[b]['depth2img']
        pts = prior + torch.as_tensor(np.asarray(pm['origin'], dtype=np.float32)) if 'origin' in pm else prior
        nv = b + 1 if continuous else V
        proj = torch.from_numpy(np.stack([M.compose_projection(pm['intrinsic'][v], pm['extrinsic'][v])
                                          for v in range(nv)]))
        vol, _ = M.batch_point_sample(img_metas[b], f0[0][:nv] if continuous else f0[b], pts, proj, pad_hw)
        vols.append(vol.reshape(n_voxels[::-1] + [-1]).permute(3, 2, 1, 0))
    img_volume = torch.stack(vols)

    stride = 64
    vs = torch.tensor([(pr[3] - pr[0]) / n_voxels[0] / stride, (pr[4] - pr[1]) / n_voxels[1] / stride,
                       (pr[5] - pr[2]) / n_voxels[2] / stride], dtype=torch.float32)
    lo = torch.tensor(cfg['point_cloud_range'][:3], dtype=torch.float32)
    coords = []
    for b, p in enumerate(points):
        q = torch.floor((p[:, :3] - lo) / vs).to(torch.int32)
        for d in range(3):
            q[:, d] = q[:, d].clamp(0, n_voxels[d] * stride - 1)
        coords.append(np.concatenate([np.full((q.shape[0], 1), b, np.int32), q.numpy()], 1))
    coords = np.concatenate(coords, 0)
    feats = torch.cat([p if cfg.get('use_xyz_feat', False) else p[:, 3:] for p in points])
    ucoords, in2out = S.unique_first(coords)
    first = np.full(ucoords.shape[0], coords.shape[0], dtype=np.int64)
    np.minimum.at(first, in2out, np.arange(coords.shape[0]))
    last = M.mink_resnet(sd, 'backbone_3d.', cfg['backbone_3d']['depth'], ucoords, feats[torch.from_numpy(first)], B,
                         training, {})[-1]
    C = last.F.shape[1]
    pv = torch.zeros((B, n_voxels[0], n_voxels[1], n_voxels[2], C))
    ijk = last.coords[:, 1:] // last.stride
    pv[last.coords[:, 0], ijk[:, 0], ijk[:, 1], ijk[:, 2]] = last.F
    fused = torch.cat([img_volume, pv.permute(0, 4, 1, 2, 3)], 1)
    return imvoxel_neck(sd, 'neck_3d.', fused, cfg['neck_3d']['n_blocks'], training)


def multiscale_gt(gt_occ: List[torch.Tensor], ratio, shape, masks=None) -> torch.Tensor:
    """occ_loss.py:7-37; `masks[i]` (X,Y,Z) bool at THIS scale: invisible voxels become 255 (ignored)."""
    gt = torch.zeros([shape[0], shape[2], shape[3], shape[4]], dtype=torch.long)
    for i in range(gt.shape[0]):
        for row in gt_occ[i].tolist():
            gt[i, row[0] // ratio, row[1] // ratio, row[2] // ratio] = row[3]
        if masks is not None:
            gt[i][~masks[i]] = 255
    return gt


def _bce1(x):
    return F.binary_cross_entropy(x, torch.ones_like(x))


def geo_scal_loss(pred, tgt):
    empty = F.softmax(pred, dim=1)[:, 0]
    nonempty = 1 - empty
    mask = tgt != 255
    nonempty_target = (tgt != 0)[mask].float()
    nonempty, empty = nonempty[mask], empty[mask]
    eps = 1e-6
    inter = (nonempty_target * nonempty).sum()
    precision = inter / (nonempty.sum() + eps)
    recall = inter / (nonempty_target.sum() + eps)
    spec = ((1 - nonempty_target) * empty).sum() / ((1 - nonempty_target).sum() + eps)
    return _bce1(precision) + _bce1(recall) + _bce1(spec)


def sem_scal_loss(pred, tgt):
    pred = F.softmax(pred, dim=1)
    loss, count = 0, 0
    mask = tgt != 255
    for i in range(pred.shape[1]):
        p = pred[:, i][mask]
        target = tgt[mask]
        ct = torch.ones_like(target).float()
        ct[target != i] = 0
        if torch.sum(ct) > 0:
            count += 1.0
            nominator = torch.sum(p * ct)
            lc = 0
            if torch.sum(p) > 0:
                lc = lc + _bce1(nominator / torch.sum(p))
            if torch.sum(ct) > 0:
                lc = lc + _bce1(nominator / torch.sum(ct))
            if torch.sum(1 - ct) > 0:
                lc = lc + _bce1(torch.sum((1 - p) * (1 - ct)) / torch.sum(1 - ct))
            loss = loss + lc
    return loss / count


def occ_loss(sd, cfg, points, imgs, data_samples, continuous=False) -> Dict[str, torch.Tensor]:
    feats = extract_feat(sd, cfg, points, imgs, [d.metainfo for d in data_samples], True, continuous)
    gt_occ = [d.gt_occupancy.cpu() for d in data_samples]
    masks = [torch.as_tensor(d.gt_occupancy_masks).cpu() for d in data_samples] \
        if 'gt_occupancy_masks' in data_samples[0] else None
    out = {}
    for i, f in enumerate(feats):
        pred = F.conv3d(f, sd[f'bbox_head.occ.{i}.weight'])
        pooled = None if masks is None else [F.max_pool3d(m.float()[None], 2 ** i, stride=2 ** i)[0].bool()
                                             for m in masks]           # imvoxel_occ_head.py:150-156
        gt = multiscale_gt(gt_occ, 2 ** i, pred.shape, pooled)
        li = F.cross_entropy(pred, gt, ignore_index=255) + sem_scal_loss(pred, gt) + geo_scal_loss(pred, gt)
        out[f'loss_occ_{i}'] = li * 0.5 ** i
    return out


def occ_predict(sd, cfg, points, imgs, data_samples) -> torch.Tensor:
    feats = extract_feat(sd, cfg, points, imgs, [d.metainfo for d in data_samples], False)
    pred = F.conv3d(feats[0], sd['bbox_head.occ.0.weight'])
    return torch.max(torch.softmax(pred, dim=1), dim=1)[1]
