"""Seeded inputs and configs of the golden cases, shared by `make_golden.py` (reference side) and the golden parity
tests (oracle / CUDA side). Test infrastructure; imports nothing from `/root/reference` or `oracle/`."""
import numpy as np
import torch


def division_safe(points, voxel_size=0.01):
    """Keep the points whose voxel index is the same under true fp32 division (what torch-CPU evaluates for
    `p / voxel_size`, the arithmetic the reference runs HERE) and under multiplication by the fp32 reciprocal (what
    torch-CUDA evaluates, the arithmetic frozen by the oracle and the product) - so the fixture does not depend on it."""
    p = points[:, :3].numpy().astype(np.float32)
    a = np.floor(p / np.float32(voxel_size))
    b = np.floor(p * (np.float32(1.0) / np.float32(voxel_size)))
    return points[torch.from_numpy((a == b).all(1))]


def det_inputs(n_scans, augment):
    from embodiedscan_b200.synth import synth_batch
    batch = synth_batch(1, n_scans, n_views=2, H=240, W=320, n_points=2000, augment=augment)
    batch['inputs']['points'] = [division_safe(p) for p in batch['inputs']['points']]
    return batch


def det_config():
    from embodiedscan_b200.synth import mv_det3d_config
    cfg = mv_det3d_config('C1')
    cfg['backbone_3d']['depth'] = 18                # the reference's MinkResNet has no depth 14 (mink_resnet.py:29-35)
    return cfg


def occ_config():
    from embodiedscan_b200.synth import mv_occ_config
    cfg = mv_occ_config('C3-small')
    cfg['backbone_3d']['depth'] = 18
    return cfg


def occ_inputs(seed_scan=1):
    from embodiedscan_b200.synth import synth_batch, synth_occupancy
    cfg = occ_config()
    batch = synth_batch(seed_scan, 1, n_views=2, H=240, W=320, n_points=4000)
    for ds in batch['data_samples']:
        ds.gt_occupancy = synth_occupancy(ds, cfg['point_cloud_range'], cfg['n_voxels'])
    return batch


def preprocess_inputs():
    """Two scans x two views of different sizes: exercises BGR->RGB, normalisation and the multi-view pad-to-max."""
    g = torch.Generator().manual_seed(77)
    return [torch.randint(0, 256, (2, 3, 30, 45), generator=g, dtype=torch.uint8),
            torch.randint(0, 256, (2, 3, 33, 40), generator=g, dtype=torch.uint8)]


def unproject_inputs():
    """V=3 small depth maps in uint16 millimetres (7 % zero pixels) with the synthetic cameras of scan 11."""
    from embodiedscan_b200.synth import synth_scan
    s = synth_scan(11, n_views=3, H=24, W=32, n_points=64)
    meta = s['data_sample'].metainfo
    return s['depth'], meta['depth2img']['intrinsic'], meta['depth2img']['extrinsic']


def fusion_inputs():
    """Point painting with every branch of the reversed augmentation flow switched on (HF, VF, R, S, T), an image flip,
    non-unit scale factors and a crop offset; features small enough for an exact element-wise comparison."""
    import math
    from embodiedscan_b200.synth import synth_scan
    s = synth_scan(21, n_views=3, H=60, W=90, n_points=400)
    g = torch.Generator().manual_seed(5)
    meta = dict(s['data_sample'].metainfo)
    ang = 0.07
    rot = torch.tensor([[math.cos(ang), -math.sin(ang), 0.], [math.sin(ang), math.cos(ang), 0.], [0., 0., 1.]]).t()
    meta.update(transformation_3d_flow=['HF', 'VF', 'R', 'S', 'T'], pcd_horizontal_flip=True, pcd_vertical_flip=True,
                pcd_rotation=rot.contiguous().numpy(), pcd_scale_factor=1.07,
                pcd_trans=np.array([0.11, -0.07, 0.03], dtype=np.float32), flip=True, scale_factor=(0.9, 1.1),
                img_crop_offset=(3.0, 5.0), img_shape=(60, 90))
    # points as the model sees them: augmented copies of the scan points (flip . rotate . scale . translate)
    # jitter: scan points are unprojected pixel centres, so they re-project onto exact half-pixel ties of the feature
    # grid (ix = 11.5), where nearest rounding is decided by the last ulp; real inputs are voxel centres, not pixel rays
    p = s['points'].clone() + 0.004 * torch.randn(s['points'].shape, generator=g)
    p[:, 0] = -p[:, 0]
    p[:, 1] = -p[:, 1]
    p = (p @ rot) * 1.07 + torch.tensor([0.11, -0.07, 0.03])
    feats = torch.randn(3, 8, 16, 24, generator=g)
    return meta, feats, p.contiguous(), (64, 96)


def target_cases():
    """(points per level, boxes (n,9), labels) triples for FCAF3D target assignment edge cases."""
    g = torch.Generator().manual_seed(9)
    lv = [torch.rand(n, 3, generator=g) * torch.tensor([4., 4., 2.]) - torch.tensor([2., 2., 0.])
          for n in (600, 150, 40, 10)]
    boxes = torch.tensor([[0.0, 0.0, 1.0, 2.0, 1.5, 1.2, 0.4, 0.05, -0.03],
                          [0.5, -0.5, 0.8, 0.8, 0.9, 0.7, -1.1, 0.0, 0.02],      # nested in the first: min-volume rule
                          [-1.2, 1.1, 0.5, 0.5, 0.4, 0.6, 2.0, 0.0, 0.0],
                          [5.0, 5.0, 5.0, 0.3, 0.3, 0.3, 0.0, 0.0, 0.0]])       # contains no point at all
    labels = torch.tensor([7, 200, 31, 5])
    few = [torch.rand(n, 3, generator=g) - 0.5 + torch.tensor([0., 0., 1.]) for n in (6, 3, 2, 1)]  # < 19 points
    return dict(regular=(lv, boxes, labels), empty_gt=(lv, boxes[:0], labels[:0]), few_points=(few, boxes[:2], labels[:2]))


def eval_inputs():
    """4 scans of ground truth + detections for indoor_eval: jittered copies of the GT (true positives at several IoU
    levels, duplicates), random false positives, a class that only occurs in the predictions (the nan filter), a class
    that only occurs in the GT (AP 0) and one paper-thin prediction (the edge clamp). Scores are all distinct."""
    g = torch.Generator().manual_seed(31)
    classes = [2, 5, 9, 11, 40]
    gts, dts = [], []
    for img in range(4):
        n = 6 + img
        ctr = torch.rand(n, 3, generator=g) * torch.tensor([5., 5., 2.]) - torch.tensor([2.5, 2.5, 0.])
        size = 0.3 + torch.rand(n, 3, generator=g)
        ang = torch.stack([torch.rand(n, generator=g) * 6.28 - 3.14, 0.1 * torch.randn(n, generator=g),
                           0.1 * torch.randn(n, generator=g)], 1)
        boxes = torch.cat([ctr, size, ang], 1)
        labels = torch.tensor([classes[int(i)] for i in torch.randint(0, len(classes), (n, ), generator=g)])
        if img == 0:
            labels[0] = 63                                          # GT-only class
        pb, ps, pl = [], [], []
        for i in range(n):
            for rep in range(int(torch.randint(0, 3, (1, ), generator=g))):   # 0, 1 or 2 detections per GT
                j = boxes[i].clone()
                mag = (0.05, 0.25)[rep]
                j[:3] += mag * size[i] * torch.randn(3, generator=g)
                j[3:6] *= 1 + mag * torch.randn(3, generator=g).clamp(-1.5, 1.5)
                j[6:] += 0.3 * mag * torch.randn(3, generator=g)
                pb.append(j); pl.append(int(labels[i])); ps.append(float(torch.rand(1, generator=g)) * 0.8 + 0.2)
        for _ in range(5):                                          # false positives, one of a prediction-only class
            j = torch.cat([torch.rand(3, generator=g) * 4 - 2, 0.3 + torch.rand(3, generator=g),
                           torch.rand(3, generator=g) - 0.5])
            pb.append(j); pl.append(int(torch.tensor(classes + [77])[torch.randint(0, 6, (1, ), generator=g)]))
            ps.append(float(torch.rand(1, generator=g)) * 0.5)
        thin = boxes[1].clone()
        thin[3:6] = torch.tensor([0.8, 0.004, 0.01])
        pb.append(thin); pl.append(int(labels[1])); ps.append(0.15 + 0.01 * img)
        keep = labels != 63 if img else torch.ones(n, dtype=torch.bool)
        gts.append(dict(gt_bboxes_3d=boxes.numpy(), gt_labels_3d=labels.numpy()))
        dts.append(dict(bboxes_3d=torch.stack(pb).numpy(), scores_3d=np.asarray(ps, np.float32),
                        labels_3d=np.asarray(pl, np.int64)))
    label2cat = {i: f'class{i}' for i in range(284)}
    return gts, dts, [0.25, 0.5], label2cat


class HashTextEncoder(torch.nn.Module):
    """Weight-free stand-in for RoBERTa on BOTH sides of the grounding fixture: hidden states are a fixed smooth function
    of (token id, position), evaluated in float64 on the CPU and rounded once, so every platform gets the same bits.
    The text encoder is a library model in the reference and in the product; the fixture pins what consumes its output."""

    def __init__(self, hidden_size=768):
        super().__init__()
        self.config = type('Config', (), dict(hidden_size=hidden_size))()
        self._anchor = torch.nn.Parameter(torch.zeros(1), requires_grad=False)      # gives the module a device

    def forward(self, input_ids, attention_mask=None, **kwargs):
        ids = input_ids.detach().cpu().double()[..., None]
        k = torch.arange(self.config.hidden_size, dtype=torch.float64)
        pos = torch.arange(input_ids.shape[1], dtype=torch.float64)[None, :, None]
        h = torch.sin(ids * (0.0137 + 0.00011 * k) + 0.5 * k) + 0.3 * torch.cos(pos * 0.21 * (k % 7 + 1))
        return type('Output', (), dict(last_hidden_state=h.float().to(input_ids.device)))()


def ground_config(prune=None):
    """`prune`: override of MinkNeck's pts_prune_threshold. The training case keeps the small threshold (150) of
    'C4-small' so that pruning is exercised; the prediction case switches pruning off, because interpolated parent
    scores tie exactly between sibling voxels and `torch.topk` (the reference) leaves the order of ties unspecified -
    a tie at the cut would make the fixture depend on the torch build instead of on the reference."""
    from embodiedscan_b200.synth import mv_grounding_config
    cfg = mv_grounding_config('C4-small')
    cfg['backbone_3d']['depth'] = 18
    if prune is not None:
        cfg['neck_3d']['pts_prune_threshold'] = prune
    return cfg


def ground_inputs(first_scan=1):
    from embodiedscan_b200.synth import add_grounding_prompt, synth_batch
    batch = synth_batch(first_scan, 2, n_views=2, H=240, W=320, n_points=2000)
    batch['inputs']['points'] = [division_safe(p) for p in batch['inputs']['points']]
    for i, ds in enumerate(batch['data_samples']):
        add_grounding_prompt(ds, 1 + 2 * i, seed=i)
    return batch


def augment_inputs():
    """Points + GT boxes of scan 5 for the 3D augmentations; `seed` makes numpy draw both flips."""
    from embodiedscan_b200.synth import synth_scan
    s = synth_scan(5, n_views=2, H=48, W=64, n_points=500)
    boxes = s['data_sample'].gt_instances_3d.bboxes_3d.tensor.clone()
    return s['points'].clone(), boxes, 2


def continuous_inputs(scan=7, n_views=3, per_view=500):
    """One scan as the continuous pipeline sees it before ConstructMultiSweeps: frame-ordered points with slice indices
    (AggregateMultiViewPoints(save_slices=True)), per-frame visibility of the GT instances, images, GT."""
    from embodiedscan_b200.synth import synth_scan
    from oracle import data_ref as D
    s = synth_scan(scan, n_views=n_views, H=240, W=320, n_points=200)
    meta = s['data_sample'].metainfo
    pm = meta['depth2img']
    g = torch.Generator().manual_seed(100 + scan)
    allp, counts = D.unproject_depth(s['depth'], pm['intrinsic'], pm['extrinsic'])
    chunks, off = [], 0
    for c in counts.tolist():
        p = division_safe(allp[off:off + c])
        chunks.append(p[torch.randperm(p.shape[0], generator=g)[:per_view]])
        off += c
    sl = [0]
    for c in chunks:
        sl.append(sl[-1] + c.shape[0])
    gt = s['data_sample'].gt_instances_3d
    n_box = len(gt.bboxes_3d)
    vis = torch.rand(n_views, n_box, generator=g) < 0.4
    vis[0, :3] = True
    return dict(points=torch.cat(chunks).contiguous(), points_slice_indices=sl, img=s['img'], meta=meta,
                boxes=gt.bboxes_3d.tensor.clone(), labels=gt.labels_3d.clone(),
                visible_instance_masks=[v.tolist() for v in vis])


def continuous_occ_inputs(scan=8, n_views=3, per_view=1200):
    """continuous_inputs + occupancy ground truth and per-frame visibility masks of the 8x8x4 grid."""
    from embodiedscan_b200.synth import synth_occupancy
    cfg = occ_config()
    ci = continuous_inputs(scan, n_views, per_view)
    ds = type('DS', (), {})()
    from embodiedscan_b200.structures import EulerDepthInstance3DBoxes, InstanceData
    gt = InstanceData()
    gt.bboxes_3d = EulerDepthInstance3DBoxes(ci['boxes'], box_dim=9)
    gt.labels_3d = ci['labels']
    ds.gt_instances_3d = gt
    ci['gt_occupancy'] = synth_occupancy(ds, cfg['point_cloud_range'], cfg['n_voxels'])
    g = torch.Generator().manual_seed(300 + scan)
    ci['visible_occupancy_masks'] = [(torch.rand(*cfg['n_voxels'], generator=g) < 0.45).numpy() for _ in range(n_views)]
    return ci


def grounding_metric_inputs():
    """8 prompts: 40 candidate boxes each (a few jittered copies of the target at different quality + noise), scores."""
    g = torch.Generator().manual_seed(51)
    dets, anns = [], []
    for i in range(8):
        tgt = torch.cat([torch.rand(3, generator=g) * 4 - 2, 0.4 + torch.rand(3, generator=g),
                         torch.rand(3, generator=g) - 0.5])[None]
        boxes = torch.cat([torch.rand(40, 3, generator=g) * 4 - 2, 0.4 + torch.rand(40, 3, generator=g),
                           torch.rand(40, 3, generator=g) - 0.5], 1)
        scores = torch.rand(40, generator=g)
        copies = {0: [], 1: [0.04], 2: [0.17], 3: [0.17, 0.04], 4: [0.3], 5: [0.04], 6: [0.17], 7: []}[i]
        for j, mag in enumerate(copies):                                  # near-copies of the target, varying quality
            b = tgt[0].clone()
            b[:3] += mag * b[3:6] * torch.randn(3, generator=g)
            b[3:6] *= 1 + mag * torch.randn(3, generator=g).clamp(-1, 1)
            boxes[5 * j] = b
            scores[5 * j] = 0.02 if i == 5 else 1.5 + j                   # prompt 5 buries its good box below the top 10
        dets.append(dict(bboxes_3d=boxes, target_scores_3d=scores))
        anns.append(dict(gt_bboxes_3d=tgt, is_view_dep=bool(i & 1), is_hard=bool(i & 2), is_unique=bool(i & 4)))
    return dets, anns


def occupancy_metric_inputs():
    g = torch.Generator().manual_seed(52)
    classes = [f'class{i}' for i in range(1, 9)]
    samples = []
    for i in range(3):
        gt_grid = torch.randint(0, 9, (8, 8, 4), generator=g) * (torch.rand(8, 8, 4, generator=g) < 0.5)
        gt_grid[gt_grid == 7] = 0                                            # class7 never in the ground truth
        pred = torch.where(torch.rand(8, 8, 4, generator=g) < 0.6, gt_grid, torch.randint(0, 9, (8, 8, 4), generator=g))
        pred[pred == 8] = 0                                                  # class8 never predicted
        idx = torch.nonzero(gt_grid)
        gt4 = torch.cat([idx, gt_grid[idx[:, 0], idx[:, 1], idx[:, 2]][:, None]], 1)
        d = dict(pred_occupancy=pred, gt_occupancy=gt4)
        if i:
            d['gt_occupancy_masks'] = torch.rand(8, 8, 4, generator=g) < 0.7
        samples.append(d)
    return classes, samples


def box_coder_inputs():
    g = torch.Generator().manual_seed(61)
    return torch.rand(2, 7, 3, generator=g) * 4 - 2, 0.7 * torch.randn(2, 7, 12, generator=g)
