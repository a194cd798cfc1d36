"""Occupancy row, host-side pieces that need no GPU: the product's vectorised SurroundOcc losses, prior
grid, FPN and dense 3D neck against the oracle's literal restatement (oracle/occ_ref.py)."""
import pytest
import torch

from embodiedscan_b200 import occupancy as OCC
from oracle import occ_ref as R


@pytest.mark.parametrize('seed,ignore', [(0, False), (1, True), (2, True)])
def test_scal_losses_match_class_loop(seed, ignore):
    g = torch.Generator().manual_seed(seed)
    pred = torch.randn(2, 9, 4, 4, 2, generator=g) * 2
    tgt = torch.randint(0, 6, (2, 4, 4, 2), generator=g)        # classes 6..8 absent -> skipped by the loop
    if ignore:
        tgt[torch.rand(tgt.shape, generator=g) < 0.2] = 255
    a, b = OCC.sem_scal_loss(pred, tgt), R.sem_scal_loss(pred, tgt)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (a, b)
    a, b = OCC.geo_scal_loss(pred, tgt), R.geo_scal_loss(pred, tgt)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (a, b)


def test_sem_scal_gradient_matches():
    g = torch.Generator().manual_seed(3)
    pred = torch.randn(1, 5, 3, 3, 2, generator=g)
    tgt = torch.randint(0, 4, (1, 3, 3, 2), generator=g)
    p1, p2 = pred.clone().requires_grad_(), pred.clone().requires_grad_()
    OCC.sem_scal_loss(p1, tgt).backward()
    R.sem_scal_loss(p2, tgt).backward()
    assert torch.allclose(p1.grad, p2.grad, rtol=1e-4, atol=1e-6)


def test_multiscale_supervision():
    gt = [torch.tensor([[0, 0, 0, 3], [7, 7, 3, 5], [2, 5, 1, 9], [3, 5, 1, 4]])]
    for ratio, shape in ((1, (1, 81, 8, 8, 4)), (2, (1, 81, 4, 4, 2)), (4, (1, 81, 2, 2, 1))):
        a = OCC.occ_multiscale_supervision(gt, ratio, shape)
        assert torch.equal(a, R.multiscale_gt(gt, ratio, shape))
    m = torch.ones(8, 8, 4, dtype=torch.bool)
    m[0, 0, 0] = False
    a = OCC.occ_multiscale_supervision(gt, 1, (1, 81, 8, 8, 4), [m])
    assert a[0, 0, 0, 0] == 255 and a[0, 7, 7, 3] == 5


def test_prior_points_order_and_values():
    rng = [-3.2, -3.2, -1.28, 3.2, 3.2, 1.28]
    gen = OCC.AlignedAnchor3DRangeGenerator(ranges=[rng], rotations=[.0])
    for nv in ([8, 8, 4], [5, 3, 2]):
        a = gen.grid_anchors([nv[::-1]], device='cpu')[0]
        assert a.shape == (nv[0] * nv[1] * nv[2], 7)
        assert torch.equal(a[:, :3], R.prior_points(rng, nv))
    # x fastest: the second prior differs from the first in x only
    assert a[1, 0] > a[0, 0] and a[1, 1] == a[0, 1] and a[1, 2] == a[0, 2]


def test_fpn_and_neck_match_oracle():
    torch.manual_seed(0)
    fpn = OCC.FPN([4, 8, 16, 32], 6, 4)
    for p in fpn.parameters():
        torch.nn.init.normal_(p, 0, 0.2)
    feats = [torch.randn(2, c, 16 >> i, 24 >> i) for i, c in enumerate([4, 8, 16, 32])]
    sd = {'neck.' + k: v for k, v in fpn.state_dict().items()}
    for a, b in zip(fpn(feats), R.fpn(sd, 'neck.', feats)):
        assert torch.allclose(a, b, atol=1e-5)

    neck = OCC.IndoorImVoxelNeck(6, 5, [1, 1, 1]).train()
    for p in neck.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, 0, 0.1)
    sd = {'neck_3d.' + k: v.clone() for k, v in neck.state_dict().items()}
    x = torch.randn(1, 6, 8, 8, 4)
    outs, ref = neck(x), R.imvoxel_neck(sd, 'neck_3d.', x, [1, 1, 1], True)
    assert [tuple(o.shape[2:]) for o in outs] == [(8, 8, 4), (4, 4, 2), (2, 2, 1)]
    for a, b in zip(outs, ref):
        assert torch.allclose(a, b, atol=1e-4), (a - b).abs().max()


def test_scal_losses_full_grid_gradients_finite():
    """A coarse scale whose voxels are all occupied makes specificity exactly 0 (log clamp active): gradients must stay
    finite and equal to the oracle's."""
    g = torch.Generator().manual_seed(5)
    pred = torch.randn(1, 7, 2, 2, 1, generator=g)
    tgt = torch.tensor([1, 2, 3, 3]).view(1, 2, 2, 1)
    p1, p2 = pred.clone().requires_grad_(), pred.clone().requires_grad_()
    (OCC.geo_scal_loss(p1, tgt) + OCC.sem_scal_loss(p1, tgt)).backward()
    (R.geo_scal_loss(p2, tgt) + R.sem_scal_loss(p2, tgt)).backward()
    assert torch.isfinite(p1.grad).all()
    assert torch.allclose(p1.grad, p2.grad, rtol=1e-4, atol=1e-6)
