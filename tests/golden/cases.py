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
