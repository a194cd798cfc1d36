"""Grounding row, host-side pieces that need no GPU: tokenizer / positive maps, the transformer decoder
and contrastive head against the oracle's explicit attention (oracle/ground_ref.py), and the batched loss (gathers,
inverse-map pairing, masks) with the two device calls — the Hungarian kernel and the 9-DoF IoU kernel — replaced by
their CPU references. The kernels themselves are covered by the `-m gpu` tests."""
import numpy as np
import pytest
import torch

from oracle import geometry_ref as G
from oracle import ground_ref as R


@pytest.fixture(scope='module')
def setup():
    import warnings
    from embodiedscan_b200 import MODELS
    from embodiedscan_b200.synth import add_grounding_prompt, mv_grounding_config, synth_batch
    torch.manual_seed(0)
    cfg = mv_grounding_config('C4-small')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = MODELS.build(cfg).train()
    with torch.no_grad():          # the reference initialises the last regression layer to zero: make it informative
        for p in model.bbox_head.reg_branches[0][-1].parameters():
            p.normal_(0, 0.05)
    batch = synth_batch(1, 2, n_views=2, H=240, W=320, n_points=2000)
    for i, ds in enumerate(batch['data_samples']):
        add_grounding_prompt(ds, 1 + 2 * i, seed=i)
    text = model.encode_text(batch['data_samples'], 'cpu')
    g = torch.Generator().manual_seed(1)
    feats = [torch.randn(n, 256, generator=g) for n in (50, 41)]
    xyz = [torch.rand(n, 3, generator=g) * 4 - 2 for n in (50, 41)]
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return cfg, model, batch, text, feats, xyz, sd


def test_tokenizer_and_positive_map(setup):
    cfg, model, batch, text, *_ = setup
    ds = batch['data_samples'][1]
    tok = model.tokenizer.batch_encode_plus([d.text for d in batch['data_samples']], padding='longest')
    assert tok.input_ids.shape == tok.attention_mask.shape and int(tok.input_ids[0, 0]) == 0
    assert int(tok.input_ids[0, tok.attention_mask[0].sum() - 1]) == 2 and int(tok.input_ids[0, -1]) == 1   # </s>, <pad>
    for (beg, end), row in zip([s[0] for s in ds.tokens_positive], ds.gt_instances_3d.positive_maps):
        t = tok.char_to_token(1, beg)
        assert ds.text[beg:end].isalpha() and row.nonzero().flatten().tolist() == [t]
    assert ds.gt_instances_3d.positive_maps.shape == (3, cfg['bbox_head']['contrastive_cfg']['max_text_len'])
    assert tok.char_to_token(0, 4) is None                      # the blank between words


def test_transformer_matches_oracle(setup):
    cfg, model, batch, text, feats, xyz, sd = setup
    head_in = model.forward_transformer(feats, None, xyz, text)
    cls = model.bbox_head(head_in['hidden_states'], head_in['text_feats'], head_in['text_token_mask'])[0]
    ref_cls, ref_boxes = R.transformer(sd, cfg, feats, xyz, text['text_feats'].detach(), text['text_token_mask'], True)
    assert cls.shape == ref_cls.shape == (2, 2, 32, 32)
    finite = torch.isfinite(ref_cls)
    assert torch.equal(finite, torch.isfinite(cls))
    assert torch.allclose(cls[finite], ref_cls[finite], atol=2e-4, rtol=1e-4), (cls[finite] - ref_cls[finite]).abs().max()
    assert torch.allclose(head_in['all_layers_pred_bboxes'], ref_boxes, atol=2e-4, rtol=1e-4)


def _cpu_hungarian(cost, n_gt):
    from scipy.optimize import linear_sum_assignment
    P, n_pred, Gm = cost.shape
    p2g = torch.full((P, n_pred), -1, dtype=torch.int32)
    g2p = torch.full((P, Gm), -1, dtype=torch.int32)
    for p in range(P):
        g = int(n_gt[p])
        if g:
            c = torch.nan_to_num(cost[p, :, :g], nan=100.0, posinf=100.0, neginf=-100.0).numpy()
            r, cidx = linear_sum_assignment(c)
            p2g[p, torch.from_numpy(r)] = torch.from_numpy(cidx).int()
            g2p[p, torch.from_numpy(cidx)] = torch.from_numpy(r).int()
    return p2g, g2p


def _cpu_overlap(c1, c2, eps=1e-4):
    vol, iou = G.box3d_overlap(c1.numpy(), c2.numpy())
    return torch.from_numpy(vol).float(), torch.from_numpy(iou).float()


def test_batched_loss_matches_per_sample_reference(setup, monkeypatch):
    from embodiedscan_b200 import grounding as GR
    cfg, model, batch, text, feats, xyz, sd = setup
    monkeypatch.setattr(GR, 'hungarian_batch', _cpu_hungarian)
    monkeypatch.setattr(GR, 'box3d_overlap', _cpu_overlap)
    head_in = model.forward_transformer(feats, None, xyz, text)
    losses = model.bbox_head.loss(**head_in, batch_data_samples=batch['data_samples'])
    cls = model.bbox_head(head_in['hidden_states'], head_in['text_feats'], head_in['text_token_mask'])[0].detach()
    boxes = head_in['all_layers_pred_bboxes'].detach()
    gt_boxes = [d.gt_instances_3d.bboxes_3d.tensor.float() for d in batch['data_samples']]
    pos_maps = [d.gt_instances_3d.positive_maps for d in batch['data_samples']]
    Ly = cls.shape[0]
    for l in range(Ly):
        lc, lb, inds = R.loss_single_layer(cls[l], boxes[l], gt_boxes, pos_maps, text['text_token_mask'], 32,
                                           cfg['bbox_head']['decouple_weights'])
        key = '' if l == Ly - 1 else f'd{l}.'
        assert abs(float(losses[key + 'loss_cls']) - float(lc)) <= 1e-4 * max(abs(float(lc)), 1e-3), (l, losses, lc)
        assert abs(float(losses[key + 'loss_bbox']) - float(lb)) <= 1e-4 * max(abs(float(lb)), 1e-3), (l, losses, lb)
        assert [int((i > 0).sum()) for i in inds] == [1, 3]
    sum(losses.values()).backward()
    g = model.bbox_head.reg_branches[0][-1].weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
    assert model.text_encoder.embeddings.word_embeddings.weight.grad is None     # frozen: no autograd through RoBERTa


def test_assigner_reference_interface(setup, monkeypatch):
    from embodiedscan_b200 import grounding as GR
    from embodiedscan_b200.structures import EulerDepthInstance3DBoxes, InstanceData
    cfg, model, batch, text, *_ = setup
    monkeypatch.setattr(GR, 'hungarian_batch', _cpu_hungarian)
    monkeypatch.setattr('embodiedscan_b200.geometry.box3d_overlap', _cpu_overlap)
    g = torch.Generator().manual_seed(3)
    ds = batch['data_samples'][1]
    pred = InstanceData()
    boxes = torch.cat([torch.rand(10, 3, generator=g) * 4 - 2, torch.rand(10, 3, generator=g) + 0.2,
                       torch.rand(10, 3, generator=g) - 0.5], 1)
    pred.bboxes_3d = EulerDepthInstance3DBoxes(boxes)
    pred.scores_3d = torch.randn(10, 32, generator=g)
    res = model.bbox_head.assigner.assign(pred, ds.gt_instances_3d)
    cost = R.match_costs(pred.scores_3d, boxes, ds.gt_instances_3d.bboxes_3d.tensor, ds.gt_instances_3d.positive_maps,
                         text['text_token_mask'][1])
    assert torch.equal(res.gt_inds, R.assign(cost))
    empty = InstanceData()
    empty.bboxes_3d = EulerDepthInstance3DBoxes(torch.zeros(0, 9))
    empty.labels_3d = torch.zeros(0, dtype=torch.long)
    assert torch.equal(model.bbox_head.assigner.assign(pred, empty).gt_inds, torch.zeros(10, dtype=torch.long))
