"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Pinned by tests/golden/grounding_g4.npz (the reference's own
grounder, neck, decoder, head, assigner and match costs executed by tests/golden/make_golden.py); the mmcv / mmdet /
pytorch3d / MinkowskiEngine arithmetic underneath stays unpinned (not installable here).

CPU restatement of the grounding path, evaluated functionally from the product's state_dict:
  embodiedscan/models/necks/mink_neck.py:133-244                         (pruned sparse FPN, coarse -> fine concat)
  embodiedscan/models/detectors/sparse_featfusion_grounder.py:324-447    (pre_decoder / forward_decoder)
  embodiedscan/models/layers/ground_transformer/decoder.py:20-34,103-179,224-297
  embodiedscan/models/dense_heads/grounding_head.py:62-99,267-363,365-417,686-824
  embodiedscan/models/task_modules/assigners/hungarian_assigner.py:56-138  (scipy linear_sum_assignment, per sample)
  embodiedscan/models/losses/match_cost.py:49-76,95-113,214-265
  †upstream: mmcv MultiheadAttention / FFN (pos added to q/k only, identity residual), mmdet py_sigmoid_focal_loss,
  weight_reduce_loss(avg_factor) = sum / (avg_factor + eps_fp32).
Attention is written out explicitly (no nn.MultiheadAttention); the RoBERTa text encoder is a library model on both
sides, so its hidden states are an input here.
"""
import math
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from . import geometry_ref as G
from . import model_ref as M
from . import sparse_ref as S


# ------------------------------------------------------------------------------------------------ sparse neck
def mink_neck(sd, prefix, levels, voxel_size, n_batch, training, cache, prune_threshold):
    n_lv = len(levels)
    feats, scores, points = [], [], []
    x = levels[-1]
    prune_score = None
    for i in range(n_lv - 1, -1, -1):
        if i < n_lv - 1:
            p = f'{prefix}up_block_{i + 1}.'
            child = S.generative_children(x.coords, x.stride // 2)
            y = M.Lvl(child, M._elu_bn(sd, p + '1', S.generative_conv(x.F, sd[p + '0.kernel']), training), x.stride // 2)
            y = M._conv(sd, p + '3.kernel', y, 3, 1, cache)
            y.F = M._elu_bn(sd, p + '4', y.F, training)
            ucoords, map_b = S.union(levels[i].coords, y.coords)
            x = M.Lvl(ucoords, S.union_add(levels[i].F, y.F, map_b, ucoords.shape[0]), y.stride)
            counts = np.bincount(x.coords[:, 0], minlength=n_batch)
            if counts.max() > prune_threshold:
                with torch.no_grad():
                    interp = S.features_at_coordinates(prune_score[0], prune_score[1], prune_score[2], x.coords)
                    keep = S.prune_mask(interp, x.coords[:, 0], n_batch, prune_threshold)
                x = M.Lvl(x.coords[keep], x.F[torch.from_numpy(np.nonzero(keep)[0])], x.stride)
        p = f'{prefix}out_block_{i}.'
        o = M._conv(sd, p + '0.kernel', x, 3, 1, cache)
        o.F = M._elu_bn(sd, p + '1', o.F, training)
        cls = o.F @ sd[prefix + 'conv_cls.kernel'] + sd[prefix + 'conv_cls.bias']
        prune_score = (o.coords, cls.detach().max(dim=1, keepdim=True).values, o.stride)
        pts = torch.from_numpy(o.coords[:, 1:]).to(torch.int32) * voxel_size
        sel = [torch.from_numpy(np.nonzero(o.coords[:, 0] == b)[0]) for b in range(n_batch)]
        feats.append([o.F[s] for s in sel])
        scores.append([cls[s] for s in sel])
        points.append([pts[s] for s in sel])
    cat = lambda lv: [torch.cat([l[b] for l in lv], 0) for b in range(n_batch)]
    return cat(feats), cat(scores), cat(points)


# ------------------------------------------------------------------------------------------------ transformer
def posembed(sd, p, xyz, training):
    """Conv1d(k=1) - BatchNorm1d - ReLU - Conv1d(k=1) on (B, N, C): the norm runs over all B*N positions."""
    h = xyz @ sd[p + '0.weight'][:, :, 0].t() + sd[p + '0.bias']
    if training:
        flat = h.reshape(-1, h.shape[-1])
        mean, var = flat.mean(0), flat.var(0, unbiased=False)
    else:
        mean, var = sd[p + '1.running_mean'], sd[p + '1.running_var']
    h = (h - mean) / torch.sqrt(var + 1e-5) * sd[p + '1.weight'] + sd[p + '1.bias']
    return F.relu(h) @ sd[p + '3.weight'][:, :, 0].t() + sd[p + '3.bias']


def mha(sd, p, query, key, value, query_pos=None, key_pos=None, key_padding_mask=None, heads=8):
    identity = query
    if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
        key_pos = query_pos
    q = query + query_pos if query_pos is not None else query
    k = key + key_pos if key_pos is not None else key
    E = q.shape[-1]
    W, b = sd[p + 'attn.in_proj_weight'], sd[p + 'attn.in_proj_bias']
    q = q @ W[:E].t() + b[:E]
    k = k @ W[E:2 * E].t() + b[E:2 * E]
    v = value @ W[2 * E:].t() + b[2 * E:]
    B, nq, nk, d = q.shape[0], q.shape[1], k.shape[1], E // heads
    q = q.view(B, nq, heads, d).transpose(1, 2)
    k = k.view(B, nk, heads, d).transpose(1, 2)
    v = v.view(B, nk, heads, d).transpose(1, 2)
    att = q @ k.transpose(-1, -2) / math.sqrt(d)
    if key_padding_mask is not None:
        att = att.masked_fill(key_padding_mask[:, None, None, :], float('-inf'))
    out = (torch.softmax(att, -1) @ v).transpose(1, 2).reshape(B, nq, E)
    return identity + out @ sd[p + 'attn.out_proj.weight'].t() + sd[p + 'attn.out_proj.bias']


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1], ), sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def contrastive_embed(sd, p, visual, text, text_mask, visual_mask=None, max_text_len=256):
    res = visual @ text.transpose(-1, -2) / math.sqrt(visual.shape[-1]) + sd[p + 'bias']
    res = res.masked_fill(~text_mask[:, None, :], float('-inf'))
    if visual_mask is not None:
        res = res.masked_fill(~visual_mask[:, :, None], float('-inf'))
    out = torch.full(res.shape[:-1] + (max_text_len, ), float('-inf'))
    out[..., :res.shape[-1]] = res
    return out


def reg_branch(sd, p, x):
    h = F.relu(x @ sd[p + '0.weight'].t() + sd[p + '0.bias'])
    h = F.relu(h @ sd[p + '2.weight'].t() + sd[p + '2.bias'])
    return h @ sd[p + '4.weight'].t() + sd[p + '4.bias']


def decode_baseline(points, pred):
    return torch.cat((pred[..., :3] + points, torch.exp(pred[..., 3:6]).clamp(min=2e-2), pred[..., 6:]), -1)


def transformer(sd, cfg, feats_list, xyz_list, text_feats, text_mask, training):
    """pre_decoder + decoder. Returns (cls_scores (Ly,B,nq,T), boxes (Ly,B,nq,9))."""
    B = len(feats_list)
    lens = [f.shape[0] for f in feats_list]
    n_max, n_min = max(lens), min(lens)
    feats = torch.zeros((B, n_max, feats_list[0].shape[1]))
    coords = torch.zeros((B, n_max, 3))
    fmask = torch.zeros((B, n_max), dtype=torch.bool)
    for b in range(B):
        feats[b, :lens[b]], coords[b, :lens[b]], fmask[b, :lens[b]] = feats_list[b], xyz_list[b], True
    n_layers = cfg['decoder']['num_layers']
    T = cfg['bbox_head']['contrastive_cfg']['max_text_len']
    enc_cls = contrastive_embed(sd, f'bbox_head.cls_branches.{n_layers}.', feats, text_feats, text_mask, fmask, T)
    topk = min(cfg['num_queries'], n_min)
    order = torch.sort(enc_cls.max(-1)[0], dim=1, descending=True, stable=True).indices[:, :topk]
    boxes0 = decode_baseline(coords, reg_branch(sd, f'bbox_head.reg_branches.{n_layers}.', feats))
    gather = lambda t: torch.gather(t, 1, order.unsqueeze(-1).expand(-1, -1, t.shape[-1]))
    query, qcoords, pred_bboxes = gather(feats), gather(coords), gather(boxes0).detach().clone()
    inter, inter_boxes = [], []
    for lid in range(n_layers):
        p = f'decoder.layers.{lid}.'
        qpos = posembed(sd, 'decoder.self_posembed.position_embedding_head.', pred_bboxes, training)
        kpos = posembed(sd, 'decoder.cross_posembed.position_embedding_head.', coords, training)
        query = _ln(sd, p + 'norms.0', mha(sd, p + 'self_attn.', query, query, query, qpos, qpos))
        query = _ln(sd, p + 'norms.1', mha(sd, p + 'cross_attn_text.', query, text_feats, text_feats, qpos, None,
                                           ~text_mask))
        query = _ln(sd, p + 'norms.2', mha(sd, p + 'cross_attn.', query, feats, feats, qpos, kpos, ~fmask))
        h = F.relu(query @ sd[p + 'ffn.layers.0.0.weight'].t() + sd[p + 'ffn.layers.0.0.bias'])
        query = _ln(sd, p + 'norms.3', query + h @ sd[p + 'ffn.layers.1.weight'].t() + sd[p + 'ffn.layers.1.bias'])
        new_boxes = decode_baseline(qcoords, reg_branch(sd, f'bbox_head.reg_branches.{lid}.', query))
        pred_bboxes = new_boxes.detach().clone()
        inter.append(_ln(sd, 'decoder.norm', query))
        inter_boxes.append(new_boxes)
    cls = torch.stack([contrastive_embed(sd, f'bbox_head.cls_branches.{l}.', inter[l], text_feats, text_mask, None, T)
                       for l in range(n_layers)])
    return cls, torch.stack(inter_boxes)


# ------------------------------------------------------------------------------------------------ assignment + loss
def match_costs(cls_score, boxes, gt_boxes, pos_map, text_mask_row, weights=(1.0, 2.0, 2.0), alpha=0.25, gamma=2, eps=1e-12):
    tm = torch.nonzero(text_mask_row).squeeze(-1)
    p = cls_score[:, tm].sigmoid()
    gt = pos_map[:, tm].float()
    neg = -(1 - p + eps).log() * (1 - alpha) * p.pow(gamma)
    pos = -(p + eps).log() * alpha * (1 - p).pow(gamma)
    c_cls = torch.einsum('nc,mc->nm', pos, gt) + torch.einsum('nc,mc->nm', neg, 1 - gt)
    c_l1 = torch.cdist(boxes, gt_boxes, p=1)
    iou = G.box3d_overlap(G.container_corners(boxes).numpy(), G.container_corners(gt_boxes).numpy())[1]
    return c_cls * weights[0] + c_l1 * weights[1] - torch.from_numpy(iou).float() * weights[2]


def assign(cost: torch.Tensor) -> torch.Tensor:
    """HungarianAssigner3D.assign: 1-based gt index per prediction, 0 = background."""
    cost = torch.nan_to_num(cost.detach(), nan=100.0, posinf=100.0, neginf=-100.0)
    r, c = linear_sum_assignment(cost.numpy())
    out = torch.zeros(cost.shape[0], dtype=torch.long)
    out[torch.from_numpy(r)] = torch.from_numpy(c) + 1
    return out


def py_sigmoid_focal(pred, target, gamma=2.0, alpha=0.25):
    ps = pred.sigmoid()
    pt = (1 - ps) * target + ps * (1 - target)
    fw = (alpha * target + (1 - alpha) * (1 - target)) * pt.pow(gamma)
    return F.binary_cross_entropy_with_logits(pred, target, reduction='none') * fw


def loss_single_layer(cls_scores, boxes, gt_boxes_list, pos_maps, text_mask, T, decouple_weights=(0.2, 0.2, 0.2, 0.4)):
    """grounding_head.py:686-824 for one decoder layer. Returns (loss_cls, loss_bbox, gt_inds per sample)."""
    B, nq = cls_scores.shape[:2]
    labels = torch.zeros((B, nq, T))
    preds, tgts, inds, n_pos = [], [], [], 0
    for b in range(B):
        with torch.no_grad():
            gi = assign(match_costs(cls_scores[b], boxes[b], gt_boxes_list[b], pos_maps[b], text_mask[b]))
        inds.append(gi)
        pos = torch.nonzero(gi > 0).squeeze(-1)
        labels[b, pos] = pos_maps[b][gi[pos] - 1]
        preds.append(boxes[b][pos])
        tgts.append(gt_boxes_list[b][gi[pos] - 1])
        n_pos += len(pos)
    tm = torch.zeros((B, T), dtype=torch.bool)
    tm[:, :text_mask.shape[1]] = text_mask
    tm = tm[:, None, :].repeat(1, nq, 1)
    sel_scores, sel_labels = torch.masked_select(cls_scores, tm), torch.masked_select(labels, tm)
    avg = max(n_pos * 1.0, 1)
    loss_cls = py_sigmoid_focal(sel_scores, sel_labels).sum() / (avg + torch.finfo(torch.float32).eps)
    pred, tgt = torch.cat(preds), torch.cat(tgts)
    w = decouple_weights
    lb = w[0] * G.chamfer_l1_mean(torch.cat((pred[:, :3], tgt[:, 3:]), -1), tgt)
    lb = lb + w[1] * G.chamfer_l1_mean(torch.cat((tgt[:, :3], pred[:, 3:6], tgt[:, 6:]), -1), tgt)
    lb = lb + w[2] * G.chamfer_l1_mean(torch.cat((tgt[:, :6], pred[:, 6:]), -1), tgt)
    lb = lb + w[3] * G.chamfer_l1_mean(pred, tgt)
    return loss_cls, lb, inds


def extract_feat(sd, cfg, points, imgs, img_metas, training):
    det_cfg = dict(bbox_head=dict(voxel_size=cfg['voxel_size']), use_xyz_feat=cfg.get('use_xyz_feat', False),
                   backbone_3d=cfg['backbone_3d'], backbone=cfg['backbone'])
    levels, cache = M.extract_feat(sd, det_cfg, points, imgs, img_metas, training)
    return mink_neck(sd, 'neck_3d.', levels, cfg['voxel_size'], len(points), training, cache,
                     cfg['neck_3d']['pts_prune_threshold'])


def grounder_forward(sd, cfg, points, imgs, data_samples, text_hidden, text_mask, training):
    feats, scores, xyz = extract_feat(sd, cfg, points, imgs, [d.metainfo for d in data_samples], training)
    text_feats = text_hidden @ sd['text_feat_map.weight'].t() + sd['text_feat_map.bias']
    return transformer(sd, cfg, feats, xyz, text_feats, text_mask, training)


def grounder_loss(sd, cfg, points, imgs, data_samples, text_hidden, text_mask, pos_maps) -> Dict[str, torch.Tensor]:
    cls, boxes = grounder_forward(sd, cfg, points, imgs, data_samples, text_hidden, text_mask, True)
    gt_boxes = [d.gt_instances_3d.bboxes_3d.tensor.float().cpu() for d in data_samples]
    T = cfg['bbox_head']['contrastive_cfg']['max_text_len']
    out, Ly = {}, cls.shape[0]
    all_inds = []
    for l in range(Ly):
        lc, lb, inds = loss_single_layer(cls[l], boxes[l], gt_boxes, pos_maps, text_mask, T,
                                         cfg['bbox_head']['decouple_weights'])
        all_inds.append(inds)
        key = '' if l == Ly - 1 else f'd{l}.'
        out[key + 'loss_cls'], out[key + 'loss_bbox'] = lc, lb
    return out, all_inds
