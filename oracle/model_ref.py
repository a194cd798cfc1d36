"""Oracle (test infrastructure): functional CPU restatement of the mv-3ddet forward / loss / predict path, driven by the
product model's ``state_dict`` so both sides evaluate the same weights.

Follows, in order: data_preprocessor.py:249-264 + utils.py:9-63 -> sparse_featfusion_single_stage.py:86-221
-> mink_resnet.py:40-140 (+ ME BasicBlock/Bottleneck †upstream) -> point_fusion.py:20-107,208-311 +
bbox_3d/utils.py:289-332 -> fcaf3d_head.py:907-1020,1091-1149 -> :1578-1664,1527-1576 -> :1151-1294,1454-1525 ->
chamfer_distance.py -> :1352-1399,1666-1725. mmdet.ResNet (†upstream) is torch's own conv/BN(eval)/ReLU/maxpool.
"""
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import geometry_ref as G
from . import sparse_ref as S

ARCH3D = {14: ('basic', (1, 1, 1, 1)), 18: ('basic', (2, 2, 2, 2)), 34: ('basic', (3, 4, 6, 3)),
          50: ('bottleneck', (3, 4, 6, 3))}
ARCH2D = {18: ('basic', (2, 2, 2, 2)), 34: ('basic', (3, 4, 6, 3)), 50: ('bottleneck', (3, 4, 6, 3))}


# ------------------------------------------------------------------------------------------------ preprocessing
def preprocess_imgs(imgs_u8: torch.Tensor, mean, std, bgr_to_rgb=True, divisor=32) -> torch.Tensor:
    """(B,V,3,H,W) uint8 -> normalised fp32, right/bottom zero padded to a multiple of `divisor`."""
    x = imgs_u8[:, :, [2, 1, 0]] if bgr_to_rgb else imgs_u8
    x = (x.float() - torch.tensor(mean).view(1, 1, 3, 1, 1)) / torch.tensor(std).view(1, 1, 3, 1, 1)
    H, W = x.shape[-2:]
    Hp, Wp = -(-H // divisor) * divisor, -(-W // divisor) * divisor
    return F.pad(x, (0, Wp - W, 0, Hp - H))


# ------------------------------------------------------------------------------------------------ 2D backbone
def _cb(sd, conv, bn, x, stride, padding, relu):
    """conv -> eval BatchNorm (-> ReLU); `conv` / `bn` are the mmdet / torchvision module names of the state dict."""
    y = F.conv2d(x, sd[conv + '.weight'], None, stride, padding)
    y = F.batch_norm(y, sd[bn + '.running_mean'], sd[bn + '.running_var'], sd[bn + '.weight'], sd[bn + '.bias'],
                     False, 0., 1e-5)
    return F.relu(y) if relu else y


def resnet2d(sd, prefix, depth, x):
    kind, blocks = ARCH2D[depth]
    x = _cb(sd, prefix + 'conv1', prefix + 'bn1', x, 2, 3, True)
    x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for i, nb in enumerate(blocks):
        for j in range(nb):
            p = f'{prefix}layer{i + 1}.{j}.'
            stride = (1, 2, 2, 2)[i] if j == 0 else 1
            idt = _cb(sd, p + 'downsample.0', p + 'downsample.1', x, stride, 0, False) \
                if (p + 'downsample.0.weight') in sd else x
            if kind == 'bottleneck':
                o = _cb(sd, p + 'conv1', p + 'bn1', x, 1, 0, True)
                o = _cb(sd, p + 'conv2', p + 'bn2', o, stride, 1, True)
                o = _cb(sd, p + 'conv3', p + 'bn3', o, 1, 0, False)
            else:
                o = _cb(sd, p + 'conv1', p + 'bn1', x, stride, 1, True)
                o = _cb(sd, p + 'conv2', p + 'bn2', o, 1, 1, False)
            x = F.relu(o + idt)
        outs.append(x)
    return outs


# ------------------------------------------------------------------------------------------------ sparse backbone
class Lvl:
    def __init__(self, coords: np.ndarray, feats: torch.Tensor, stride: int):
        self.coords, self.F, self.stride = coords, feats, stride


def _bn(sd, p, x, training):
    if training:
        return S.batch_norm(x, sd[p + '.bn.weight'], sd[p + '.bn.bias'], 1e-5)
    return (x - sd[p + '.bn.running_mean']) * torch.rsqrt(sd[p + '.bn.running_var'] + 1e-5) * sd[p + '.bn.weight'] + \
        sd[p + '.bn.bias']


def _conv(sd, name, lv: Lvl, ksize, stride, cache):
    """Sparse conv with cached coordinate/kernel maps keyed by (id(coords), ksize, stride)."""
    key = (id(lv.coords), ksize, stride)
    if key not in cache:
        out_coords = S.unique_first(lv.coords, lv.stride * stride)[0] if stride > 1 else lv.coords
        ck = ('coords', id(lv.coords), stride)
        out_coords = cache.setdefault(ck, out_coords)
        cache[key] = (out_coords, S.kernel_map(lv.coords, out_coords, S.offsets(ksize, lv.stride)))
    out_coords, nbr = cache[key]
    return Lvl(out_coords, S.conv(lv.F, sd[name], nbr), lv.stride * stride)


def mink_resnet(sd, prefix, depth, coords, feats, n_batch, training, cache):
    kind, blocks = ARCH3D[depth]
    x = _conv(sd, prefix + 'conv1.kernel', Lvl(coords, feats, 1), 3, 2, cache)
    x.F = F.relu(S.instance_norm(x.F, x.coords[:, 0], n_batch, sd[prefix + 'norm1.weight'], sd[prefix + 'norm1.bias']))
    pooled = cache.setdefault(('coords', id(x.coords), 2), S.unique_first(x.coords, x.stride * 2)[0])
    nbr = S.kernel_map(x.coords, pooled, S.offsets(2, x.stride))
    x = Lvl(pooled, S.maxpool(x.F, nbr), x.stride * 2)
    outs = []
    for i, nb in enumerate(blocks):
        for j in range(nb):
            p = f'{prefix}layer{i + 1}.{j}.'
            stride = 2 if j == 0 else 1
            if (p + 'downsample.0.kernel') in sd:
                r = _conv(sd, p + 'downsample.0.kernel', x, 1, stride, cache)
                res = _bn(sd, p + 'downsample.1', r.F, training)
            else:
                res = x.F
            if kind == 'basic':
                o = _conv(sd, p + 'conv1.kernel', x, 3, stride, cache)
                o.F = F.relu(_bn(sd, p + 'norm1', o.F, training))
                o = _conv(sd, p + 'conv2.kernel', o, 3, 1, cache)
                o.F = F.relu(_bn(sd, p + 'norm2', o.F, training) + res)
            else:
                o = Lvl(x.coords, F.relu(_bn(sd, p + 'norm1', x.F @ sd[p + 'conv1.kernel'], training)), x.stride)
                o = _conv(sd, p + 'conv2.kernel', o, 3, stride, cache)
                o.F = F.relu(_bn(sd, p + 'norm2', o.F, training))
                o.F = F.relu(_bn(sd, p + 'norm3', o.F @ sd[p + 'conv3.kernel'], training) + res)
            x = o
        outs.append(x)
    return outs


# ------------------------------------------------------------------------------------------------ point painting
def compose_projection(intrinsic, extrinsic):
    """fp32 intrinsic @ extrinsic, sequential k = 0..3 sum of separately rounded products."""
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


def apply_3d_transformation_reverse(pcd: torch.Tensor, img_meta: dict) -> torch.Tensor:
    """point_fusion.py:20-107 with reverse=True for DEPTH points; every step a separately rounded fp32 op."""
    p = pcd.clone()
    flow = list(img_meta.get('transformation_3d_flow', []))[::-1]
    for op in flow:
        if op == 'T':
            t = -torch.as_tensor(np.asarray(img_meta.get('pcd_trans', [0., 0., 0.]), dtype=np.float32)).view(-1)
            p = p + t
        elif op == 'S':
            p = p * torch.tensor(np.float32(1.0 / img_meta.get('pcd_scale_factor', 1.)))
        elif op == 'R':
            rot = torch.as_tensor(np.asarray(img_meta['pcd_rotation'], dtype=np.float32)) if 'pcd_rotation' in img_meta \
                else torch.eye(3)
            r = rot.inverse()
            x, y, z = p[:, 0], p[:, 1], p[:, 2]
            p = torch.stack([(x * r[0, 0] + y * r[1, 0]) + z * r[2, 0], (x * r[0, 1] + y * r[1, 1]) + z * r[2, 1],
                             (x * r[0, 2] + y * r[1, 2]) + z * r[2, 2]], 1)
        elif op == 'HF':
            if img_meta.get('pcd_horizontal_flip', False):
                p = p * torch.tensor([-1., 1., 1.])
        elif op == 'VF':
            if img_meta.get('pcd_vertical_flip', False):
                p = p * torch.tensor([1., -1., 1.])
        else:
            raise AssertionError(op)
    return p


def batch_point_sample(img_meta, img_features, points, proj_mat, pad_hw):
    """point_fusion.py:208-311 with aligned=False, valid_flag=True. img_features (V,C,Hf,Wf); points (N,3);
    proj_mat (V,4,4). Returns (N,C) plus the per-view nearest pixel index / validity for integer parity checks."""
    pts = apply_3d_transformation_reverse(points, img_meta)
    V = proj_mat.shape[0]
    x, y, z = pts[:, 0][None], pts[:, 1][None], pts[:, 2][None]
    P = proj_mat
    X = ((x * P[:, 0, 0, None] + y * P[:, 0, 1, None]) + z * P[:, 0, 2, None]) + P[:, 0, 3, None]
    Y = ((x * P[:, 1, 0, None] + y * P[:, 1, 1, None]) + z * P[:, 1, 2, None]) + P[:, 1, 3, None]
    Z = ((x * P[:, 2, 0, None] + y * P[:, 2, 1, None]) + z * P[:, 2, 2, None]) + P[:, 2, 3, None]
    zc = Z.clamp(min=1e-3)
    sf = img_meta.get('scale_factor', (1., 1.))
    co = img_meta.get('img_crop_offset', (0., 0.))
    u = (X / zc) * torch.tensor(np.float32(sf[0])) - torch.tensor(np.float32(co[0]))
    v = (Y / zc) * torch.tensor(np.float32(sf[1])) - torch.tensor(np.float32(co[1]))
    if img_meta.get('flip', False):
        u = float(img_meta['img_shape'][1]) - u
    h, w = float(pad_hw[0]), float(pad_hw[1])
    grid = torch.stack([u / w * 2 - 1, v / h * 2 - 1], -1).unsqueeze(1)          # (V,1,N,2)
    feats = F.grid_sample(img_features, grid, mode='nearest', padding_mode='zeros', align_corners=True)   # (V,C,1,N)
    valid = (u < w) & (u > 0) & (v < h) & (v > 0) & (Z > 0)
    valid_num = valid.sum(0)
    out = feats.squeeze(2).sum(0).t()
    out = torch.where((valid_num > 0)[:, None], out, torch.zeros_like(out))
    out = out / torch.clamp(valid_num[:, None], min=1)
    return out, valid_num


# ------------------------------------------------------------------------------------------------ head
def _elu_bn(sd, p, x, training):
    return F.elu(_bn(sd, p, x, training))


def head_forward(sd, prefix, levels: List[Lvl], voxel_size, n_batch, training, cache, prune_threshold=100000):
    """fcaf3d_head.py:993-1020 + _forward_single :1116-1149. Returns per-level lists of per-scan tensors."""
    n_lv = len(levels)
    outs = [None] * n_lv
    x = levels[-1]
    prune_score = None          # (coords, (N,1) max-class logits, stride) of the previous (coarser) output
    for i in range(n_lv - 1, -1, -1):
        if i < n_lv - 1:
            p = f'{prefix}up_block_{i + 1}.'
            child = S.generative_children(x.coords, x.stride // 2)
            y = Lvl(child, _elu_bn(sd, p + '1', S.generative_conv(x.F, sd[p + '0.kernel']), training), x.stride // 2)
            y = _conv(sd, p + '3.kernel', y, 3, 1, cache)
            y.F = _elu_bn(sd, p + '4', y.F, training)
            ucoords, map_b = S.union(levels[i].coords, y.coords)
            x = Lvl(ucoords, S.union_add(levels[i].F, y.F, map_b, ucoords.shape[0]), y.stride)
            counts = np.bincount(x.coords[:, 0], minlength=n_batch)
            if counts.max() > prune_threshold:         # _prune (fcaf3d_head.py:1091-1114); identity otherwise
                with torch.no_grad():
                    interp = S.features_at_coordinates(prune_score[0], prune_score[1], prune_score[2], x.coords)
                    keep = S.prune_mask(interp, x.coords[:, 0], n_batch, prune_threshold)
                x = Lvl(x.coords[keep], x.F[torch.from_numpy(np.nonzero(keep)[0])], x.stride)
        p = f'{prefix}out_block_{i}.'
        o = _conv(sd, p + '0.kernel', x, 3, 1, cache)
        o.F = _elu_bn(sd, p + '1', o.F, training)
        center = o.F @ sd[prefix + 'conv_center.kernel']
        reg = o.F @ sd[prefix + 'conv_reg.kernel']
        cls = o.F @ sd[prefix + 'conv_cls.kernel'] + sd[prefix + 'conv_cls.bias']
        prune_score = (o.coords, cls.detach().max(dim=1, keepdim=True).values, o.stride)
        dist = torch.exp(reg[:, :6] * sd[f'{prefix}scales.{i}.scale']).clamp(min=1e-3)
        bbox = torch.cat((dist, reg[:, 6:]), 1)
        pts = torch.from_numpy(o.coords[:, 1:]).to(torch.int32) * voxel_size       # int32 * python float -> fp32
        per = []
        for b in range(n_batch):
            sel = torch.from_numpy(np.nonzero(o.coords[:, 0] == b)[0])
            per.append((center[sel], bbox[sel], cls[sel], pts[sel]))
        outs[i] = per
    return outs


def get_targets(points: List[torch.Tensor], boxes9: torch.Tensor, labels: torch.Tensor, assign_thr=27, center_thr=18):
    """fcaf3d_head.py:1578-1664 restated densely; rotation uses separately rounded fp32 products (left to right)."""
    float_max = 1e8
    n_levels = len(points)
    levels = torch.cat([torch.full((len(p), ), i, dtype=torch.long) for i, p in enumerate(points)])
    pts = torch.cat(points)
    n_points, n_boxes = len(pts), len(boxes9)
    if n_boxes == 0:
        return pts.new_zeros(n_points), pts.new_zeros((n_points, 9)), labels.new_full((n_points, ), -1)
    volumes = (boxes9[:, 3] * boxes9[:, 4] * boxes9[:, 5])[None].expand(n_points, n_boxes)
    R = G.euler_to_matrix(-boxes9[:, 6:9])                                  # (Nb,3,3)
    s = pts[:, None, :] - boxes9[None, :, :3]                               # (Np,Nb,3)
    rx = (s[..., 0] * R[None, :, 0, 0] + s[..., 1] * R[None, :, 0, 1]) + s[..., 2] * R[None, :, 0, 2]
    ry = (s[..., 0] * R[None, :, 1, 0] + s[..., 1] * R[None, :, 1, 1]) + s[..., 2] * R[None, :, 1, 2]
    rz = (s[..., 0] * R[None, :, 2, 0] + s[..., 1] * R[None, :, 2, 1]) + s[..., 2] * R[None, :, 2, 2]
    b = boxes9[None]
    cx, cy, cz = b[..., 0] + rx, b[..., 1] + ry, b[..., 2] + rz
    fd = torch.stack((cx - b[..., 0] + b[..., 3] / 2, b[..., 0] + b[..., 3] / 2 - cx, cy - b[..., 1] + b[..., 4] / 2,
                      b[..., 1] + b[..., 4] / 2 - cy, cz - b[..., 2] + b[..., 5] / 2, b[..., 2] + b[..., 5] / 2 - cz), -1)
    inside = fd.min(dim=-1).values > 0
    n_pos_per_level = torch.stack([inside[levels == i].sum(0) for i in range(n_levels)], 0)
    lower_mask = n_pos_per_level < assign_thr
    lower_index = torch.argmax(lower_mask.int(), dim=0) - 1
    lower_index = torch.where(lower_index < 0, 0, lower_index)
    all_upper = torch.all(torch.logical_not(lower_mask), dim=0)
    best_level = torch.where(all_upper, n_levels - 1, lower_index)
    level_cond = best_level[None].expand(n_points, n_boxes) == levels[:, None].expand(n_points, n_boxes)
    xd, yd, zd = fd[..., [0, 1]], fd[..., [2, 3]], fd[..., [4, 5]]
    cent = torch.sqrt(xd.min(-1)[0] / xd.max(-1)[0] * yd.min(-1)[0] / yd.max(-1)[0] * zd.min(-1)[0] / zd.max(-1)[0])
    cent = torch.where(inside, cent, torch.ones_like(cent) * -1)
    cent = torch.where(level_cond, cent, torch.ones_like(cent) * -1)
    top = torch.topk(cent, min(center_thr + 1, len(cent)), dim=0).values[-1]
    topk_cond = cent > top.unsqueeze(0)
    vol = torch.where(inside, volumes, torch.full_like(volumes, float_max))
    vol = torch.where(level_cond, vol, torch.full_like(vol, float_max))
    vol = torch.where(topk_cond, vol, torch.full_like(vol, float_max))
    min_vol, min_inds = vol.min(dim=1)
    ar = torch.arange(n_points)
    cls_t = torch.where(min_vol == float_max, -1, labels[min_inds])
    return cent[ar, min_inds], boxes9[min_inds], cls_t


def loss_single(center_preds, bbox_preds, cls_preds, points, boxes9, labels, n_pos_avg=None,
                decouple_weights=(0.2, 0.2, 0.2, 0.4)):
    """fcaf3d_head.py:1151-1294 for the configured head (BBoxCDLoss l1/g8, decoupled 4 groups)."""
    center_t, bbox_t, cls_t = get_targets(points, boxes9, labels)
    center_preds, bbox_preds, cls_preds, pts = (torch.cat(center_preds), torch.cat(bbox_preds), torch.cat(cls_preds),
                                                torch.cat(points))
    pos = torch.nonzero(cls_t >= 0).squeeze(1)
    n_pos = max(float(len(pos)) if n_pos_avg is None else n_pos_avg, 1.)
    cls_loss = G.sigmoid_focal_loss_sum(cls_preds, cls_t) / n_pos
    if len(pos) > 0:
        center_loss = F.binary_cross_entropy_with_logits(center_preds[pos], center_t[pos].unsqueeze(1),
                                                         reduction='none').sum() / n_pos
        dec = G.bbox_pred_to_bbox(pts[pos], bbox_preds[pos])
        tgt = bbox_t[pos]
        tc, ts, te, pc, ps, pe = tgt[:, :3], tgt[:, 3:6], tgt[:, 6:], dec[:, :3], dec[:, 3:6], dec[:, 6:]
        w = decouple_weights
        bbox_loss = w[0] * G.chamfer_l1_mean(torch.cat((pc, ts, te), -1), tgt)
        bbox_loss = bbox_loss + w[1] * G.chamfer_l1_mean(torch.cat((tc, ps, te), -1), tgt)
        bbox_loss = bbox_loss + w[2] * G.chamfer_l1_mean(torch.cat((tc, ts, pe), -1), tgt)
        bbox_loss = bbox_loss + w[3] * G.chamfer_l1_mean(dec, tgt)
    else:
        center_loss, bbox_loss = center_preds[pos].sum(), bbox_preds[pos].sum()
    return center_loss, bbox_loss, cls_loss, (center_t, bbox_t, cls_t)


def predict_single(center_preds, bbox_preds, cls_preds, points, nms_pre=1000, score_thr=.01, iou_thr=.5):
    """fcaf3d_head.py:1352-1399 + :1666-1725."""
    mb, ms = [], []
    for c, b, k, p in zip(center_preds, bbox_preds, cls_preds, points):
        scores = k.sigmoid() * c.sigmoid()
        mx = scores.max(dim=1).values
        if len(scores) > nms_pre > 0:
            ids = mx.topk(nms_pre).indices
            b, scores, p = b[ids], scores[ids], p[ids]
        mb.append(G.bbox_pred_to_bbox(p, b))
        ms.append(scores)
    return G.multiclass_nms(torch.cat(mb), torch.cat(ms), score_thr, iou_thr)


# ------------------------------------------------------------------------------------------------ whole detector
def extract_feat(sd, cfg, points: List[torch.Tensor], imgs: torch.Tensor, img_metas: List[dict], training: bool,
                 continuous: bool = False):
    """sparse_featfusion_single_stage.py:86-221. imgs (B,V,3,Hp,Wp) normalised fp32.
    continuous=True restates embodied_det3d.py:90-207 instead: imgs (1,V,...) of ONE scan, `points[b]` = its frames
    0..b, and sample b is painted from views 0..b only (:146-149)."""
    vs = cfg['bbox_head']['voxel_size']
    cache = {}
    coords = np.concatenate([S.voxelize(p, vs, b) for b, p in enumerate(points)], 0)
    feats = torch.cat([p if cfg.get('use_xyz_feat', False) else p[:, 3:] for p in points])
    ucoords, in2out = S.unique_first(coords)
    first = np.full(ucoords.shape[0], coords.shape[0], dtype=np.int64)
    np.minimum.at(first, in2out, np.arange(coords.shape[0]))
    B = len(points)
    levels = mink_resnet(sd, 'backbone_3d.', cfg['backbone_3d']['depth'], ucoords, feats[torch.from_numpy(first)], B,
                         training, cache)
    V = imgs.shape[1]
    f2d = resnet2d(sd, 'backbone.', cfg['backbone']['depth'], imgs.reshape((-1, ) + tuple(imgs.shape[2:])))
    pad_hw = tuple(imgs.shape[-2:])
    for li, lv in enumerate(levels):
        fl = f2d[li].reshape((imgs.shape[0], V) + tuple(f2d[li].shape[1:]))
        painted = torch.zeros((lv.coords.shape[0], fl.shape[2]))
        for b in range(B):
            sel = np.nonzero(lv.coords[:, 0] == b)[0]
            pm = img_metas[b]['depth2img']
            nv = b + 1 if continuous else V
            proj = torch.from_numpy(np.stack([compose_projection(pm['intrinsic'][v], pm['extrinsic'][v])
                                              for v in range(nv)]))
            pts = torch.from_numpy(lv.coords[sel, 1:]).to(torch.int32) * vs
            out, _ = batch_point_sample(img_metas[b], fl[0][:nv] if continuous else fl[b], pts, proj, pad_hw)
            painted = painted.index_copy(0, torch.from_numpy(sel), out)
        lv.F = torch.cat([lv.F, painted], 1)
    return levels, cache


def detector_loss(sd, cfg, points, imgs, data_samples, world_n_pos=None, continuous=False) -> Dict[str, torch.Tensor]:
    metas = [d.metainfo for d in data_samples]
    levels, cache = extract_feat(sd, cfg, points, imgs, metas, True, continuous)
    B = len(points)
    outs = head_forward(sd, 'bbox_head.', levels, cfg['bbox_head']['voxel_size'], B, True, cache,
                        cfg['bbox_head']['pts_prune_threshold'])
    cl, bl, kl = [], [], []
    for b in range(B):
        gt = data_samples[b].gt_instances_3d
        boxes9 = torch.cat((gt.bboxes_3d.gravity_center, gt.bboxes_3d.tensor[:, 3:]), 1).float()
        c, bb, k, _ = loss_single([outs[l][b][0] for l in range(len(outs))], [outs[l][b][1] for l in range(len(outs))],
                                  [outs[l][b][2] for l in range(len(outs))], [outs[l][b][3] for l in range(len(outs))],
                                  boxes9, gt.labels_3d, decouple_weights=cfg['bbox_head']['decouple_weights'])
        cl.append(c); bl.append(bb); kl.append(k)
    return dict(loss_center=torch.stack(cl).mean(), loss_bbox=torch.stack(bl).mean(), loss_cls=torch.stack(kl).mean())


def detector_predict(sd, cfg, points, imgs, data_samples, continuous=False):
    metas = [d.metainfo for d in data_samples]
    levels, cache = extract_feat(sd, cfg, points, imgs, metas, False, continuous)
    B = len(points)
    outs = head_forward(sd, 'bbox_head.', levels, cfg['bbox_head']['voxel_size'], B, False, cache,
                        cfg['bbox_head']['pts_prune_threshold'])
    t = cfg['test_cfg']
    return [predict_single([outs[l][b][0] for l in range(len(outs))], [outs[l][b][1] for l in range(len(outs))],
                           [outs[l][b][2] for l in range(len(outs))], [outs[l][b][3] for l in range(len(outs))],
                           t['nms_pre'], t['score_thr'], t['iou_thr']) for b in range(B)]
