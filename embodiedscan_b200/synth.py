"""Synthetic RGB-D scans of the named shapes (SURVEY §8d / BASELINE.md §3): a room with 24 oriented cuboids, V pinhole
cameras, analytic ray-cast depth in uint16 millimetres, points by the reference's unprojection semantics
(embodiedscan/datasets/transforms/points.py:30-81, multiview.py:139-169), and the ``img_meta`` / GT containers the
detector consumes (embodiedscan/datasets/transforms/formatting.py:67-78). Seeded; no dataset, no network.
"""
import math
from typing import Dict, Optional

import numpy as np
import torch

from .geometry import euler_angles_to_matrix, matrix_to_euler_angles_zxy
from .structures import Det3DDataSample, EulerDepthInstance3DBoxes, InstanceData

ROOM = (-3.0, 3.0, -3.0, 3.0, 0.0, 2.8)
DEPTH_SHIFT = 1000.0


def _intrinsic(H, W):
    f = 577.87 * W / 640.0
    K = np.eye(4, dtype=np.float32)
    K[0, 0] = K[1, 1] = f
    K[0, 2], K[1, 2] = W / 2.0, H / 2.0
    return K


def _camera(gen):
    pos = torch.stack([torch.empty(1).uniform_(ROOM[0] + .5, ROOM[1] - .5, generator=gen),
                       torch.empty(1).uniform_(ROOM[2] + .5, ROOM[3] - .5, generator=gen),
                       torch.empty(1).uniform_(1.2, 1.7, generator=gen)]).view(3).double()
    yaw = float(torch.empty(1).uniform_(-math.pi, math.pi, generator=gen))
    pitch = float(torch.empty(1).uniform_(math.radians(-20), math.radians(10), generator=gen))
    f = torch.tensor([math.cos(yaw) * math.cos(pitch), math.sin(yaw) * math.cos(pitch), math.sin(pitch)]).double()
    up = torch.tensor([0., 0., 1.]).double()
    right = torch.linalg.cross(f, up)
    right = right / right.norm()
    down = torch.linalg.cross(f, right)
    R = torch.stack([right, down, f], 1)                 # cam -> world
    E = torch.eye(4).double()
    E[:3, :3] = R.t()
    E[:3, 3] = -R.t() @ pos
    return E.float().numpy(), R, pos


def _raycast(R, pos, K, H, W, boxes, device):
    """depth (H, W) fp32 metres of the nearest hit among room planes and cuboids."""
    us, vs = torch.meshgrid(torch.arange(W, device=device), torch.arange(H, device=device), indexing='xy')
    d_cam = torch.stack([(us.float() - K[0, 2]) / K[0, 0], (vs.float() - K[1, 2]) / K[1, 1],
                         torch.ones_like(us, dtype=torch.float32)], -1).view(-1, 3)
    Rf, o = R.float().to(device), pos.float().to(device)
    d = d_cam @ Rf.t()
    t_best = torch.full((d.shape[0], ), float('inf'), device=device)
    for axis in range(3):
        for bound in (ROOM[2 * axis], ROOM[2 * axis + 1]):
            t = (bound - o[axis]) / d[:, axis]
            t = torch.where(t > 1e-4, t, torch.full_like(t, float('inf')))
            t_best = torch.minimum(t_best, t)
    Rb = euler_angles_to_matrix(boxes[:, 6:9].to(device), 'ZXY')          # (M,3,3) box -> world
    for m in range(boxes.shape[0]):
        oc = (o - boxes[m, :3].to(device)) @ Rb[m]                          # into the box frame (R^T . v)
        dc = d @ Rb[m]
        half = boxes[m, 3:6].to(device) / 2
        inv = 1.0 / torch.where(dc.abs() < 1e-9, torch.full_like(dc, 1e-9), dc)
        t1, t2 = (-half - oc) * inv, (half - oc) * inv
        tn = torch.minimum(t1, t2).max(1).values
        tf = torch.maximum(t1, t2).min(1).values
        hit = (tf >= tn) & (tf > 1e-4)
        t = torch.where(tn > 1e-4, tn, tf)
        t_best = torch.where(hit & (t < t_best), t, t_best)
    return t_best.view(H, W)


def synth_scan(scan_idx: int = 0, n_views: int = 20, H: int = 480, W: int = 640, n_points: int = 100000,
               n_boxes: int = 24, num_classes: int = 284, augment: bool = False, device='cpu') -> Dict:
    """Returns dict(points (n,3) fp32, img (V,3,H,W) uint8, depth (V,H,W) uint16 tensors on `device`,
    data_sample Det3DDataSample with metainfo + gt_instances_3d)."""
    gen = torch.Generator().manual_seed(1234 + scan_idx)
    ctr = torch.stack([torch.empty(n_boxes).uniform_(ROOM[0] + .3, ROOM[1] - .3, generator=gen),
                       torch.empty(n_boxes).uniform_(ROOM[2] + .3, ROOM[3] - .3, generator=gen),
                       torch.empty(n_boxes).uniform_(0.3, 2.2, generator=gen)], 1)
    size = torch.empty(n_boxes, 3).uniform_(0.2, 1.5, generator=gen)
    euler = torch.stack([torch.empty(n_boxes).uniform_(-math.pi, math.pi, generator=gen),
                         torch.empty(n_boxes).normal_(0, 0.05, generator=gen),
                         torch.empty(n_boxes).normal_(0, 0.05, generator=gen)], 1)
    boxes = torch.cat([ctr, size, euler], 1).float()
    labels = torch.randint(0, num_classes, (n_boxes, ), generator=gen)

    K = _intrinsic(H, W)
    Kt = torch.from_numpy(K)
    extr, intr, pts_all, depths = [], [], [], []
    per_view = max(-(-2 * n_points // n_views), 1)     # 20 views x 10k -> 200k -> sample 100k (cfg :140-152)
    for v in range(n_views):
        E, R, pos = _camera(gen)
        depth_m = _raycast(R, pos, Kt, H, W, boxes, device)
        depth_mm = torch.clamp(torch.round(depth_m * DEPTH_SHIFT), 0, 65535)
        drop = torch.rand(H * W, generator=gen) < 0.05
        depth_mm = torch.where(drop.view(H, W).to(device), torch.zeros_like(depth_mm), depth_mm)
        depth_u16 = depth_mm.to(torch.int32)
        depths.append(depth_u16)
        # unprojection with the reference's semantics, done here in fp32 torch (the CUDA data path has its own kernel)
        d = depth_u16.float() / DEPTH_SHIFT
        us, vs = torch.meshgrid(torch.arange(W, device=device), torch.arange(H, device=device), indexing='xy')
        grid = torch.stack([us.float() * d, vs.float() * d, d, torch.ones_like(d)], -1).view(-1, 4)
        nz = torch.nonzero(d.reshape(-1)).squeeze(1)
        Kinv_t = torch.inverse(Kt).t().to(device)
        cam = (grid @ Kinv_t)[nz]
        cam[:, 3] = 1
        Et = torch.from_numpy(E).to(device)
        world = torch.linalg.solve(Et, cam.t()).t()[:, :3]
        perm = torch.randperm(world.shape[0], generator=gen)[:per_view].to(device)
        pts_all.append(world[perm])
        extr.append(E)
        intr.append(K.copy())
    pts = torch.cat(pts_all)
    perm = torch.randperm(pts.shape[0], generator=gen)[:n_points].to(device)
    pts = pts[perm].contiguous()
    img = torch.randint(0, 256, (n_views, 3, H, W), generator=gen, dtype=torch.uint8).to(device)
    depth = torch.stack(depths).to(torch.int32)

    meta = dict(img_shape=(H, W), ori_shape=(H, W), scale_factor=(1.0, 1.0), flip=False,
                depth2img=dict(extrinsic=extr, intrinsic=intr, origin=np.zeros(3, dtype=np.float32)),
                box_type_3d=EulerDepthInstance3DBoxes, transformation_3d_flow=[], sample_idx=scan_idx)
    if augment:
        ang = float(torch.empty(1).uniform_(-0.087266, 0.087266, generator=gen))
        s = float(torch.empty(1).uniform_(0.9, 1.1, generator=gen))
        t = torch.empty(3).normal_(0, 0.1, generator=gen)
        Rz = torch.tensor([[math.cos(ang), -math.sin(ang), 0.], [math.sin(ang), math.cos(ang), 0.], [0., 0., 1.]])
        M = Rz.t().contiguous()                       # points @ M
        pts = ((pts @ M.to(device)) * s + t.to(device)).contiguous()
        Rb = euler_angles_to_matrix(boxes[:, 6:9], 'ZXY')
        new_e = matrix_to_euler_angles_zxy(Rz[None] @ Rb)
        boxes = torch.cat([(boxes[:, :3] @ M) * s + t, boxes[:, 3:6] * s, new_e], 1)
        meta.update(pcd_rotation=M.numpy(), pcd_scale_factor=s, pcd_trans=t.numpy(),
                    pcd_horizontal_flip=False, pcd_vertical_flip=False, transformation_3d_flow=['R', 'S', 'T'])
    ds = Det3DDataSample(metainfo=meta)
    gt = InstanceData()
    gt.bboxes_3d = EulerDepthInstance3DBoxes(boxes, box_dim=9)
    gt.labels_3d = labels
    ds.gt_instances_3d = gt
    return dict(points=pts, img=img, depth=depth, data_sample=ds)


def synth_batch(first_idx: int, batch_size: int, **kw) -> Dict:
    """The dict a dataloader hands to ``model.train_step``: {'inputs': {'points': [...], 'img': [...]}, 'data_samples'}."""
    scans = [synth_scan(first_idx + i, **kw) for i in range(batch_size)]
    return dict(inputs=dict(points=[s['points'] for s in scans], img=[s['img'] for s in scans]),
                data_samples=[s['data_sample'] for s in scans], depth=[s['depth'] for s in scans])


# model configs mirroring configs/detection/mv-det3d_8xb4_embodiedscan-3d-284class-9dof.py:16-59
def mv_det3d_config(variant: str = 'C2') -> dict:
    """C2: ResNet-50/16 + MinkResNet34 (the published config). C1: ResNet-18/16 + MinkResNet14 (BASELINE config 0)."""
    if variant == 'C2':
        depth2d, depth3d, in_ch = 50, 34, (128, 256, 512, 1024)
    elif variant == 'C1':
        depth2d, depth3d, in_ch = 18, 14, (80, 160, 320, 640)
    else:
        raise KeyError(variant)
    return dict(
        type='SparseFeatureFusionSingleStage3DDetector',
        data_preprocessor=dict(type='Det3DDataPreprocessor', mean=[123.675, 116.28, 103.53],
                               std=[58.395, 57.12, 57.375], bgr_to_rgb=True, pad_size_divisor=32),
        backbone=dict(type='mmdet.ResNet', depth=depth2d, base_channels=16, num_stages=4, out_indices=(0, 1, 2, 3),
                      frozen_stages=1, norm_cfg=dict(type='BN', requires_grad=False), norm_eval=True, style='pytorch'),
        backbone_3d=dict(type='MinkResNet', in_channels=3, depth=depth3d),
        use_xyz_feat=True,
        bbox_head=dict(type='FCAF3DHeadRotMat', in_channels=in_ch, out_channels=128, voxel_size=.01,
                       pts_prune_threshold=100000, pts_assign_threshold=27, pts_center_threshold=18, num_classes=284,
                       num_reg_outs=12, center_loss=dict(type='mmdet.CrossEntropyLoss', use_sigmoid=True),
                       bbox_loss=dict(type='BBoxCDLoss', mode='l1', loss_weight=1.0, group='g8'),
                       cls_loss=dict(type='mmdet.FocalLoss'), decouple_bbox_loss=True, decouple_groups=4,
                       decouple_weights=[0.2, 0.2, 0.2, 0.4]),
        coord_type='DEPTH', train_cfg=dict(), test_cfg=dict(nms_pre=1000, iou_thr=.5, score_thr=.01))


# occupancy (BASELINE config C3): configs/occupancy/mv-occ_8xb1_embodiedscan-occ-80class.py:10-60
def synth_occupancy(data_sample, point_cloud_range, n_voxels, num_classes: int = 81) -> torch.Tensor:
    """Ground-truth occupancy (M,4) int64 [ix, iy, iz, label] of the non-empty voxels: voxels whose centre lies in a
    ground-truth box get that box's class (1..num_classes-1), the floor layer gets class 1; empty voxels are omitted
    (label 0 by construction in ``occ_multiscale_supervision``)."""
    r = torch.tensor(point_cloud_range, dtype=torch.float32)
    n = torch.tensor(n_voxels)
    step = (r[3:] - r[:3]) / n
    ii = torch.stack(torch.meshgrid(*[torch.arange(k) for k in n_voxels], indexing='ij'), -1).view(-1, 3)
    ctr = r[:3] + (ii.float() + 0.5) * step
    gt = data_sample.gt_instances_3d
    boxes = gt.bboxes_3d.tensor.float().cpu()
    rot = euler_angles_to_matrix(boxes[:, 6:9], 'ZXY')
    local = torch.einsum('nbj,bjk->nbk', ctr[:, None] - boxes[None, :, :3], rot)       # R^T (p - c)
    inside = (local.abs() <= boxes[None, :, 3:6] / 2).all(-1)
    label = torch.zeros(ctr.shape[0], dtype=torch.long)
    label[ctr[:, 2] < r[2] + step[2]] = 1
    for b in range(boxes.shape[0]):
        label[inside[:, b]] = int(gt.labels_3d[b]) % (num_classes - 1) + 1
    nz = label > 0
    return torch.cat([ii[nz], label[nz, None]], 1)


def mv_occ_config(variant: str = 'C3') -> dict:
    """C3: full-width ResNet-50 + FPN + MinkResNet34 on a 40x40x16 grid (the published config). 'C3-small': ResNet-18/16,
    MinkResNet14 and an 8x8x4 grid for parity tests the CPU oracle finishes in seconds."""
    if variant == 'C3':
        depth2d, base, depth3d, n_vox = 50, 64, 34, [40, 40, 16]
        prior, pcr = [-3.2, -3.2, -1.28, 3.2, 3.2, 1.28], [-3.2, -3.2, -0.78, 3.2, 3.2, 1.78]
        c2d, c3d, fpn_out, neck_out = [256, 512, 1024, 2048], 512, 256, 128
    elif variant == 'C3-small':
        depth2d, base, depth3d, n_vox = 18, 16, 14, [8, 8, 4]
        prior, pcr = [-3.2, -3.2, -1.28, 3.2, 3.2, 1.28], [-3.2, -3.2, -0.78, 3.2, 3.2, 1.78]
        c2d, c3d, fpn_out, neck_out = [16, 32, 64, 128], 512, 32, 32
    else:
        raise KeyError(variant)
    return dict(
        type='DenseFusionOccPredictor', use_valid_mask=False, use_xyz_feat=True, point_cloud_range=pcr,
        data_preprocessor=dict(type='Det3DDataPreprocessor', mean=[123.675, 116.28, 103.53],
                               std=[58.395, 57.12, 57.375], bgr_to_rgb=True, pad_size_divisor=32),
        backbone=dict(type='mmdet.ResNet', depth=depth2d, base_channels=base, num_stages=4, out_indices=(0, 1, 2, 3),
                      frozen_stages=1, norm_cfg=dict(type='BN', requires_grad=False), norm_eval=True, style='pytorch'),
        backbone_3d=dict(type='MinkResNet', in_channels=3, depth=depth3d),
        neck=dict(type='mmdet.FPN', in_channels=c2d, out_channels=fpn_out, num_outs=4),
        neck_3d=dict(type='IndoorImVoxelNeck', in_channels=fpn_out + c3d, out_channels=neck_out, n_blocks=[1, 1, 1]),
        bbox_head=dict(type='ImVoxelOccHead', volume_h=[n_vox[0], n_vox[0] // 2, n_vox[0] // 4],
                       volume_w=[n_vox[1], n_vox[1] // 2, n_vox[1] // 4],
                       volume_z=[n_vox[2], n_vox[2] // 2, n_vox[2] // 4], num_classes=81,
                       in_channels=[neck_out] * 3, use_semantic=True),
        prior_generator=dict(type='AlignedAnchor3DRangeGenerator', ranges=[prior], rotations=[.0]),
        n_voxels=n_vox, coord_type='DEPTH')


# grounding (BASELINE config C4): configs/grounding/mv-grounding_8xb12_embodiedscan-vg-9dof.py:18-88
_NOUNS = ['chair', 'table', 'cabinet', 'lamp', 'sofa', 'shelf', 'monitor', 'plant', 'box', 'bed']


def add_grounding_prompt(data_sample, n_targets: int = 1, seed: int = 0):
    """Turn a detection sample into a visual-grounding sample: keep ``n_targets`` of its boxes as the targets of a
    synthetic prompt and attach ``text`` + ``tokens_positive`` (character spans of each target's phrase), the fields
    ``MultiView3DGroundingDataset`` provides (embodiedscan/datasets/mv_3dvg_dataset.py †)."""
    gen = torch.Generator().manual_seed(977 + seed)
    gt = data_sample.gt_instances_3d
    n = len(gt.bboxes_3d)
    pick = torch.randperm(n, generator=gen)[:n_targets]
    words = [_NOUNS[int(gt.labels_3d[i]) % len(_NOUNS)] for i in pick]
    text, spans = 'find the ', []
    for k, w in enumerate(words):
        if k:
            text += ' and the '
        spans.append([[len(text), len(text) + len(w)]])
        text += w
    text += ' that is close to the wall.'
    new = InstanceData()
    new.bboxes_3d = EulerDepthInstance3DBoxes(gt.bboxes_3d.tensor[pick], box_dim=9)
    new.labels_3d = gt.labels_3d[pick]
    data_sample.gt_instances_3d = new
    data_sample.text = text
    data_sample.tokens_positive = spans
    return data_sample


def mv_grounding_config(variant: str = 'C4') -> dict:
    """C4: the published grounding model. 'C4-small': ResNet-18/16 + MinkResNet14, 2 decoder layers, 32 queries and a
    prune threshold low enough to exercise pruning — for parity tests the CPU oracle finishes in seconds."""
    if variant == 'C4':
        depth2d, depth3d, in_ch, nq, n_layers, ffn, prune, T = 50, 34, [128, 256, 512, 1024], 256, 6, 2048, 1000, 256
    elif variant == 'C4-small':
        depth2d, depth3d, in_ch, nq, n_layers, ffn, prune, T = 18, 14, [80, 160, 320, 640], 32, 2, 256, 150, 32
    else:
        raise KeyError(variant)
    attn = dict(embed_dims=256, num_heads=8, dropout=0.0)
    return dict(
        type='SparseFeatureFusion3DGrounder', num_queries=nq, voxel_size=0.01,
        data_preprocessor=dict(type='Det3DDataPreprocessor', mean=[123.675, 116.28, 103.53],
                               std=[58.395, 57.12, 57.375], bgr_to_rgb=True, pad_size_divisor=32),
        backbone=dict(type='mmdet.ResNet', depth=depth2d, base_channels=16, num_stages=4, out_indices=(0, 1, 2, 3),
                      frozen_stages=1, norm_cfg=dict(type='BN', requires_grad=False), norm_eval=True, style='pytorch'),
        backbone_3d=dict(type='MinkResNet', in_channels=3, depth=depth3d), use_xyz_feat=True,
        neck_3d=dict(type='MinkNeck', num_classes=1, in_channels=in_ch, out_channels=256, voxel_size=0.01,
                     pts_prune_threshold=prune),
        decoder=dict(num_layers=n_layers, return_intermediate=True,
                     layer_cfg=dict(self_attn_cfg=dict(attn), cross_attn_text_cfg=dict(attn), cross_attn_cfg=dict(attn),
                                    ffn_cfg=dict(embed_dims=256, feedforward_channels=ffn, ffn_drop=0.0)),
                     post_norm_cfg=None),
        bbox_head=dict(type='GroundingHead', num_classes=256, num_pred_layer=n_layers + 1, sync_cls_avg_factor=True,
                       decouple_bbox_loss=True, decouple_groups=4, share_pred_layer=True,
                       decouple_weights=[0.2, 0.2, 0.2, 0.4],
                       contrastive_cfg=dict(max_text_len=T, log_scale='auto', bias=True),
                       loss_cls=dict(type='mmdet.FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
                       loss_bbox=dict(type='BBoxCDLoss', mode='l1', loss_weight=1.0, group='g8')),
        coord_type='DEPTH',
        train_cfg=dict(assigner=dict(type='HungarianAssigner3D', match_costs=[
            dict(type='BinaryFocalLossCost', weight=1.0), dict(type='BBox3DL1Cost', weight=2.0),
            dict(type='IoU3DCost', weight=2.0)])),
        test_cfg=None)
