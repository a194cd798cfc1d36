"""The oracle pinned against golden vectors produced by the REFERENCE'S OWN Python (tests/golden/make_golden.py ran
`/root/reference/embodiedscan/**` in the dev container, third-party packages stood in; see tests/golden/README.md).

These tests need neither a GPU nor /root/reference: weights are rebuilt from the stored manifest by the name-keyed fill,
inputs by the seeded generators in tests/golden/cases.py.  Tolerances are fp32 round-off of a differently ordered
evaluation of the same formulas (the reference uses bmm / grid_sample / index ops where the oracle spells out sums)."""
import os
import sys

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
if GOLD not in sys.path:
    sys.path.insert(0, GOLD)

from cases import (box_coder_inputs, grounding_metric_inputs, occupancy_metric_inputs, HashTextEncoder, augment_inputs, continuous_inputs, continuous_occ_inputs, det_config, det_inputs, eval_inputs, fusion_inputs, ground_config, ground_inputs, occ_config, occ_inputs, preprocess_inputs,  # noqa: E402
                   target_cases, unproject_inputs)
from weights import adjust_fcaf3d_head, adjust_for_predict, adjust_grounder, fill_state_dict  # noqa: E402


def load(name):
    return np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=False)


def manifest(g):
    return [(str(k), tuple(int(x) for x in str(s).split(',') if x)) for k, s in
            zip(g['manifest_names'], g['manifest_shapes'])]


def product_state_dict(cfg, g, adjust):
    """Reference-named weights -> the product model (strict load = checkpoint compatibility) -> the
    product's own state_dict, which is what the oracle is driven by."""
    from embodiedscan_b200 import MODELS
    model = MODELS.build(cfg)
    ref_sd = adjust(fill_state_dict(manifest(g)))
    ref_sd.update({k[len('calib/'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('calib/')})
    missing, unexpected = model.load_state_dict(ref_sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith('num_batches_tracked') for k in missing), missing
    return model, {k: v.detach().clone() for k, v in model.state_dict().items()}


def rel(a, b):
    a, b = float(torch.as_tensor(a).detach()), float(torch.as_tensor(b).detach())
    return abs(a - b) / max(abs(b), 1e-6)


WATCH = {'bbox_head.conv_cls.kernel': None, 'bbox_head.conv_reg.kernel': None, 'bbox_head.out_block_0.0.kernel': None,
         'bbox_head.up_block_1.0.kernel': None, 'backbone_3d.conv1.kernel': None,
         'backbone_3d.layer2.0.conv1.kernel': None, 'backbone_3d.layer1.0.norm1.bn.weight': None,
         'backbone.layer2.0.conv1.weight': 'backbone.layer2.0.cb1.conv.weight', 'bbox_head.scales.1.scale': None}


def sampled(g):
    return g if g.numel() <= 4096 else g.flatten()[:: max(g.numel() // 4096, 1)][:4096]


@pytest.mark.parametrize('tag,n_scans,augment', [('a', 1, False), ('b', 2, True)])
def test_detector_loss_and_gradients_match_reference(tag, n_scans, augment):
    from oracle import model_ref as M
    g = load('detector_g1')
    cfg = det_config()
    _, sd = product_state_dict(cfg, g, adjust_fcaf3d_head)
    batch = det_inputs(n_scans, augment)
    assert [len(p) for p in batch['inputs']['points']] == g[f'{tag}_n_points'].tolist()
    imgs = M.preprocess_imgs(torch.stack(batch['inputs']['img']), cfg['data_preprocessor']['mean'],
                             cfg['data_preprocessor']['std'])
    for ref_name, own in WATCH.items():
        k = ref_name          # the oracle reads the reference-named state dict
        sd[k] = sd[k].clone().requires_grad_(True)
    out = M.detector_loss(sd, cfg, batch['inputs']['points'], imgs, batch['data_samples'])
    sum(out.values()).backward()
    for k in ('loss_center', 'loss_bbox', 'loss_cls'):
        assert rel(out[k], g[f'{tag}_{k}']) <= 2e-5, (k, float(out[k]), float(g[f'{tag}_{k}']))
    for ref_name, own in WATCH.items():
        grad = sd[ref_name].grad
        want = torch.from_numpy(g[f'{tag}_grad/{ref_name}'])
        got = sampled(grad).reshape(want.shape)
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 2e-4 * scale, (ref_name, float((got - want).abs().max()), scale)
        assert rel(grad.double().norm(), g[f'{tag}_gradnorm/{ref_name}']) <= 5e-4, ref_name   # rotation columns: atan2/asin chains


def test_detector_loss_with_active_pruning_matches_reference():
    from oracle import model_ref as M
    g = load('detector_g1')
    cfg = det_config()
    cfg['bbox_head']['pts_prune_threshold'] = int(g['c_prune'])
    _, sd = product_state_dict(cfg, g, adjust_fcaf3d_head)
    batch = det_inputs(1, False)
    imgs = M.preprocess_imgs(torch.stack(batch['inputs']['img']), cfg['data_preprocessor']['mean'],
                             cfg['data_preprocessor']['std'])
    with torch.no_grad():
        out = M.detector_loss(sd, cfg, batch['inputs']['points'], imgs, batch['data_samples'])
    assert rel(g['c_loss_cls'], g['a_loss_cls']) > 1e-3, 'pruning must change the result'
    for k in ('loss_center', 'loss_bbox', 'loss_cls'):
        assert rel(out[k], g[f'c_{k}']) <= 2e-5, (k, float(out[k]), float(g[f'c_{k}']))


def test_detector_predictions_match_reference():
    from oracle import model_ref as M
    g = load('detector_g1')
    cfg = det_config()
    cfg['test_cfg'] = dict(nms_pre=50, iou_thr=.5, score_thr=float(g['p_score_thr']))
    _, sd = product_state_dict(cfg, g, lambda s: adjust_for_predict(adjust_fcaf3d_head(s)))
    batch = det_inputs(1, False)
    imgs = M.preprocess_imgs(torch.stack(batch['inputs']['img']), cfg['data_preprocessor']['mean'],
                             cfg['data_preprocessor']['std'])
    with torch.no_grad():
        boxes, scores, labels = M.detector_predict(sd, cfg, batch['inputs']['points'], imgs, batch['data_samples'])[0]
    assert len(g['p_labels']) > 10, 'the fixture must exercise NMS'
    assert torch.equal(labels, torch.from_numpy(g['p_labels'])), 'selection order must be identical'
    assert float((scores - torch.from_numpy(g['p_scores'])).abs().max()) <= 1e-5
    want = torch.from_numpy(g['p_boxes'])
    assert want.shape[1] == 9 and float(want[:, 7:].abs().max()) == 0.0
    assert float((boxes - want[:, :7]).abs().max()) <= 1e-4 * float(want.abs().max())


# ------------------------------------------------------------------------------------------------ occupancy (a14)
OCC_WATCH = {'bbox_head.occ.0.weight': None, 'bbox_head.occ.2.weight': None, 'neck_3d.down_layer_1.0.conv1.weight': None,
             'neck_3d.up_block_1.0.weight': None, 'neck.lateral_convs.0.conv.weight': None,
             'neck.lateral_convs.3.conv.bias': None, 'backbone_3d.layer4.0.conv1.kernel': None,
             'backbone_3d.conv1.kernel': None, 'backbone.layer2.0.conv1.weight': 'backbone.layer2.0.cb1.conv.weight'}


def test_occupancy_loss_and_gradients_match_reference():
    from oracle import model_ref as M
    from oracle import occ_ref as R
    g = load('occupancy_g3')
    cfg = occ_config()
    _, sd = product_state_dict(cfg, g, lambda s: s)
    batch = occ_inputs(int(g['a_scan']))
    imgs = M.preprocess_imgs(torch.stack(batch['inputs']['img']), cfg['data_preprocessor']['mean'],
                             cfg['data_preprocessor']['std'])
    for ref_name, own in OCC_WATCH.items():
        k = ref_name          # the oracle reads the reference-named state dict
        sd[k] = sd[k].clone().requires_grad_(True)
    out = R.occ_loss(sd, cfg, batch['inputs']['points'], imgs, batch['data_samples'])
    sum(out.values()).backward()
    for k in ('loss_occ_0', 'loss_occ_1', 'loss_occ_2'):
        assert rel(out[k], g['a_' + k]) <= 2e-5, (k, float(out[k]), float(g['a_' + k]))
    for ref_name, own in OCC_WATCH.items():
        grad = sd[ref_name].grad
        want = torch.from_numpy(g[f'a_grad/{ref_name}'])
        got = sampled(grad).reshape(want.shape)
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 2e-4 * scale, (ref_name, float((got - want).abs().max()), scale)
        assert rel(grad.double().norm(), g[f'a_gradnorm/{ref_name}']) <= 2e-4, ref_name


def check_occupancy_prediction(pred, g, decisive_margin=1e-3):
    want, margin = torch.from_numpy(g['p_occupancy']), torch.from_numpy(g['p_margin'])
    assert pred.shape == want.shape
    decisive = margin > decisive_margin          # near-tied argmax voxels may flip under fp32 reordering
    assert int(decisive.sum()) > 0.5 * decisive.numel()
    assert torch.equal(pred[decisive], want[decisive])
    assert float((pred == want).float().mean()) >= 0.99


def test_occupancy_predictions_match_reference():
    from oracle import model_ref as M
    from oracle import occ_ref as R
    g = load('occupancy_g3')
    cfg = occ_config()
    _, sd = product_state_dict(cfg, g, lambda s: s)
    batch = occ_inputs(2)
    imgs = M.preprocess_imgs(torch.stack(batch['inputs']['img']), cfg['data_preprocessor']['mean'],
                             cfg['data_preprocessor']['std'])
    with torch.no_grad():
        pred = R.occ_predict(sd, cfg, batch['inputs']['points'], imgs, batch['data_samples'])[0]
    check_occupancy_prediction(pred, g)


# ------------------------------------------------------------------------------------------------ front-end (a1, a2)
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def test_image_collation_matches_reference():
    from oracle import data_ref as D
    g = load('frontend')
    imgs = preprocess_inputs()
    out = D.preprocess_multiview(imgs, MEAN, STD, True, 32)
    assert torch.equal(out, torch.from_numpy(g['pre_imgs'])), 'fp32 (x-mean)/std and zero padding are exact'
    assert D.pad_shapes(imgs, 32) == [tuple(r) for r in g['pre_pad_shape'].tolist()]
    # the uniform-shape helper used by the model-level oracle agrees with it
    from oracle import model_ref as M
    same = [imgs[0], imgs[0].flip(0)]
    assert torch.equal(M.preprocess_imgs(torch.stack(same), MEAN, STD), D.preprocess_multiview(same, MEAN, STD))


def test_unprojection_matches_reference():
    from oracle import data_ref as D
    g = load('frontend')
    depth, intr, extr = unproject_inputs()
    pts, counts = D.unproject_depth(depth, intr, extr)
    assert counts.tolist() == g['unproj_counts'].tolist()
    assert int((depth == 0).sum()) > 0, 'the fixture must contain dropped pixels'
    want = torch.from_numpy(g['unproj_points'])
    assert float((pts - want).abs().max()) <= 1e-6 * float(want.abs().max())


# ------------------------------------------------------------------------------------------------ function-level pins
def test_point_painting_all_branches_matches_reference():
    from oracle import model_ref as M
    g = load('functions')
    meta, feats, pts, pad_hw = fusion_inputs()
    pm = meta['depth2img']
    proj = torch.from_numpy(np.stack([M.compose_projection(pm['intrinsic'][v], pm['extrinsic'][v])
                                      for v in range(feats.shape[0])]))
    out, valid = M.batch_point_sample(meta, feats, pts, proj, pad_hw)
    want = torch.from_numpy(g['fusion_out'])
    assert int((valid > 0).sum()) > 100 and int((valid == 0).sum()) > 0
    # nearest-pixel selection must be identical: a wrong pixel changes a row by O(1), round-off by O(1e-7)
    assert float((out - want).abs().max()) <= 1e-5


@pytest.mark.parametrize('name', ['regular', 'empty_gt', 'few_points'])
def test_target_assignment_edge_cases_match_reference(name):
    from oracle import model_ref as M
    g = load('functions')
    lv, boxes, labels = target_cases()[name]
    c, b, k = M.get_targets(lv, boxes, labels)
    assert torch.equal(k, torch.from_numpy(g[f'targets_{name}_cls'])), 'assignment is integer-exact'
    assert float((c - torch.from_numpy(g[f'targets_{name}_center'])).abs().max()) <= 1e-5     # sqrt of a ratio product
    assert torch.equal(b, torch.from_numpy(g[f'targets_{name}_bbox']))


def test_box_corners_and_empty_nms_match_reference():
    from oracle import geometry_ref as G
    g = load('functions')
    boxes = target_cases()['regular'][1]
    assert float((G.container_corners(boxes) - torch.from_numpy(g['corners'])).abs().max()) <= 1e-6
    b, s_, l = G.multiclass_nms(torch.rand(5, 9)[:, :7], torch.full((5, 284), 0.001), 0.01, 0.5)
    assert [*b.shape, *s_.shape, *l.shape] == g['nms_empty_shapes'].tolist()


def test_product_box_container_matches_reference():
    """Host-side structures of the product (no CUDA involved): corners order / rotation of the 9-DoF box container."""
    from embodiedscan_b200.structures import EulerDepthInstance3DBoxes
    g = load('functions')
    boxes = target_cases()['regular'][1]
    c = EulerDepthInstance3DBoxes(boxes.clone(), box_dim=9, origin=(.5, .5, .5)).corners
    assert float((c - torch.from_numpy(g['corners'])).abs().max()) <= 1e-6


# ------------------------------------------------------------------------------------------------ evaluation (f3)
def _eval_golden():
    import json
    return json.loads(str(load('eval')['result_json']))


def test_oracle_indoor_eval_matches_reference():
    from oracle import eval_ref as E
    gts, dts, metric, label2cat = eval_inputs()
    want = _eval_golden()
    got = E.indoor_eval(gts, dts, metric, label2cat)
    assert set(got) == set(want)
    assert 'class77_AP_0.25' not in want and want['class63_AP_0.25'] == 0.0      # nan filter / GT-only class
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-6, (k, got[k], want[k])


def test_product_indoor_eval_host_logic_matches_reference():
    """The product's vectorised AP accumulation with the IoU matrix injected (the CUDA IoU is checked in -m gpu)."""
    from embodiedscan_b200.evaluation import indoor_eval
    from oracle import eval_ref as E
    gts, dts, metric, label2cat = eval_inputs()
    want = _eval_golden()
    got = indoor_eval(gts, dts, metric, label2cat,
                      iou_fn=lambda p, q: torch.from_numpy(E.iou_matrix(p.numpy(), q.numpy())))
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-6, (k, got[k], want[k])
    with pytest.raises(RuntimeError):
        if not torch.cuda.is_available():
            indoor_eval(gts, dts, metric, label2cat)         # no silent CPU fallback for the IoU
        else:
            raise RuntimeError('cuda present')


# ------------------------------------------------------------------------------------------------ grounding (a15)
GROUND_WATCH = {'bbox_head.reg_branches.0.4.weight': None, 'bbox_head.cls_branches.0.bias': None,
                'text_feat_map.weight': None, 'decoder.layers.0.cross_attn.attn.in_proj_weight': None,
                'decoder.layers.1.ffn.layers.1.weight': None,
                'decoder.cross_posembed.position_embedding_head.0.weight': None, 'neck_3d.out_block_0.0.kernel': None,
                'neck_3d.up_block_2.0.kernel': None, 'backbone_3d.conv1.kernel': None,
                'backbone.layer2.0.conv1.weight': 'backbone.layer2.0.cb1.conv.weight'}


def build_grounder(g, prune=None):
    """Product grounder with the fixture's weights and the weight-free text encoder of the fixture."""
    import warnings
    from embodiedscan_b200 import MODELS
    cfg = ground_config(prune)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = MODELS.build(cfg)
    model.text_encoder = HashTextEncoder()
    ref_sd = adjust_grounder(fill_state_dict(manifest(g)))
    ref_sd.update({k[len('calib/'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('calib/')})
    missing, unexpected = model.load_state_dict(ref_sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith('num_batches_tracked') for k in missing), missing
    return cfg, model


def grounder_text_side(model, batch):
    """Tokenisation, positive maps (host logic of the product) and the text encoder's hidden states."""
    tok = model.tokenizer.batch_encode_plus([d.text for d in batch['data_samples']], padding='longest')
    pos_maps = [m.bool().float() for m in model.get_positive_map(tok, [d.tokens_positive for d in batch['data_samples']])]
    hidden = model.text_encoder(**tok).last_hidden_state.float().cpu()
    return hidden, tok.attention_mask.bool(), pos_maps


def alias_shared_branches(sd):
    for k in list(sd):
        for kind in ('cls_branches', 'reg_branches'):
            tag = f'bbox_head.{kind}.'
            if k.startswith(tag) and not k.startswith(tag + '0.'):
                sd[k] = sd[tag + '0.' + k[len(tag):].split('.', 1)[1]]
    return sd


def test_grounder_loss_and_gradients_match_reference():
    from oracle import ground_ref as R
    from oracle import model_ref as M
    g = load('grounding_g4')
    cfg, model = build_grounder(g)
    batch = ground_inputs(1)
    hidden, tmask, pos_maps = grounder_text_side(model, batch)
    for i, pm in enumerate(pos_maps):
        assert torch.equal(pm, torch.from_numpy(g[f'a_positive_map_{i}'])), 'token spans of the targets'
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for ref_name, own in GROUND_WATCH.items():
        k = ref_name          # the oracle reads the reference-named state dict
        sd[k] = sd[k].clone().requires_grad_(True)
    sd = alias_shared_branches(sd)
    imgs = M.preprocess_imgs(torch.stack(batch['inputs']['img']), cfg['data_preprocessor']['mean'],
                             cfg['data_preprocessor']['std'])
    out, _ = R.grounder_loss(sd, cfg, batch['inputs']['points'], imgs, batch['data_samples'], hidden, tmask, pos_maps)
    sum(out.values()).backward()
    assert {'a_' + k for k in out} == {k for k in g.files if k.startswith('a_') and 'loss' in k}
    for k in out:
        assert rel(out[k], g['a_' + k]) <= 5e-5, (k, float(out[k]), float(g['a_' + k]))
    for ref_name, own in GROUND_WATCH.items():
        grad = sd[ref_name].grad
        want = torch.from_numpy(g[f'a_grad/{ref_name}'])
        got = sampled(grad).reshape(want.shape)
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 5e-4 * scale, (ref_name, float((got - want).abs().max()), scale)
        assert rel(grad.double().norm(), g[f'a_gradnorm/{ref_name}']) <= 5e-4, ref_name


def test_grounder_predictions_match_reference():
    from oracle import ground_ref as R
    from oracle import model_ref as M
    g = load('grounding_g4')
    cfg, model = build_grounder(g, prune=100000)
    batch = ground_inputs(3)
    hidden, tmask, _ = grounder_text_side(model, batch)
    sd = alias_shared_branches({k: v.detach().clone() for k, v in model.state_dict().items()})
    imgs = M.preprocess_imgs(torch.stack(batch['inputs']['img']), cfg['data_preprocessor']['mean'],
                             cfg['data_preprocessor']['std'])
    with torch.no_grad():
        cls, boxes = R.grounder_forward(sd, cfg, batch['inputs']['points'], imgs, batch['data_samples'], hidden, tmask,
                                        False)
    for b in range(2):
        scores = cls[-1][b].sigmoid().max(-1)[0]
        assert float((scores - torch.from_numpy(g[f'p_scores_{b}'])).abs().max()) <= 5e-5
        want = torch.from_numpy(g[f'p_boxes_{b}'])
        assert float((boxes[-1][b] - want).abs().max()) <= 1e-4 * float(want.abs().max())


# ------------------------------------------------------------------------------------------------ augmentation (f1)
def test_device_augmentations_match_reference():
    """RandomFlip3D + GlobalRotScaleTrans of the product are torch ops on whatever device holds the points: the same
    code runs here on CPU tensors. Same numpy seed -> same draws as the reference pipeline."""
    from embodiedscan_b200.structures import EulerDepthInstance3DBoxes
    from embodiedscan_b200.transforms import GlobalRotScaleTrans, RandomFlip3D
    g = load('augment')
    pts, boxes, seed = augment_inputs()
    np.random.seed(seed)
    d = dict(points=pts.clone(), gt_bboxes_3d=EulerDepthInstance3DBoxes(boxes.clone(), box_dim=9))
    d = RandomFlip3D(sync_2d=False, flip_2d=False, flip_ratio_bev_horizontal=0.5, flip_ratio_bev_vertical=0.5)(d)
    d = GlobalRotScaleTrans(rot_range=[-0.087266, 0.087266], scale_ratio_range=[.9, 1.1], translation_std=[.1, .1, .1],
                            shift_height=False)(d)
    assert d['transformation_3d_flow'] == g['flow'].tolist() == ['HF', 'VF', 'R', 'S', 'T']
    assert float(d['pcd_scale_factor']) == float(g['pcd_scale_factor'])
    assert np.array_equal(np.asarray(d['pcd_trans']), g['pcd_trans'])
    assert float((torch.as_tensor(d['pcd_rotation']) - torch.from_numpy(g['pcd_rotation'])).abs().max()) <= 1e-7
    assert float((d['points'] - torch.from_numpy(g['points'])).abs().max()) <= 1e-6
    got, want = d['gt_bboxes_3d'].tensor, torch.from_numpy(g['boxes'])
    assert float((got[:, :6] - want[:, :6]).abs().max()) <= 1e-6
    # Euler angles modulo 2*pi (both sides go through matrix -> ZXY angles)
    dang = torch.remainder(got[:, 6:] - want[:, 6:] + np.pi, 2 * np.pi) - np.pi
    assert float(dang.abs().max()) <= 1e-5
    # and the painting side reverses exactly this flow: augmented points map back onto the original ones
    from oracle import model_ref as M
    meta = {k: d[k] for k in ('transformation_3d_flow', 'pcd_horizontal_flip', 'pcd_vertical_flip', 'pcd_scale_factor',
                              'pcd_trans')}
    meta['pcd_rotation'] = torch.as_tensor(d['pcd_rotation']).numpy()
    back = M.apply_3d_transformation_reverse(d['points'], meta)
    assert float((back - pts).abs().max()) <= 1e-5


# ------------------------------------------------------------------------------------------------ continuous (f4)
CONT_WATCH = {'bbox_head.conv_cls.kernel': None, 'bbox_head.conv_reg.kernel': None, 'backbone_3d.conv1.kernel': None,
              'backbone.layer2.0.conv1.weight': 'backbone.layer2.0.cb1.conv.weight',
              'backbone.layer4.0.conv1.weight': 'backbone.layer4.0.cb1.conv.weight'}


def continuous_batch(scan=7):
    """The product's own continuous front-end: ConstructMultiSweeps -> one data sample with per-prefix lists, in the
    pseudo-collated layout a batch-size-1 dataloader hands to the model."""
    from embodiedscan_b200.structures import Det3DDataSample, EulerDepthInstance3DBoxes, InstanceData
    from embodiedscan_b200.transforms import ConstructMultiSweeps
    ci = continuous_inputs(scan)
    res = ConstructMultiSweeps()(dict(points=ci['points'].clone(), points_slice_indices=ci['points_slice_indices'],
                                      gt_bboxes_3d=EulerDepthInstance3DBoxes(ci['boxes'].clone(), box_dim=9),
                                      gt_labels_3d=ci['labels'].clone(),
                                      visible_instance_masks=ci['visible_instance_masks']))
    ds = Det3DDataSample(metainfo=dict(ci['meta']))
    gt = InstanceData()
    gt.bboxes_3d, gt.labels_3d = res['gt_bboxes_3d'], res['gt_labels_3d']
    ds.gt_instances_3d = gt
    ds.eval_ann_info = None
    return dict(inputs=dict(points=[[p] for p in res['points']], img=[ci['img']]), data_samples=[ds]), res


def continuous_config():
    cfg = det_config()
    cfg['type'] = 'Embodied3DDetector'
    cfg['data_preprocessor'] = dict(cfg['data_preprocessor'], batchwise_inputs=True)
    return cfg


def test_continuous_front_end_matches_reference():
    from embodiedscan_b200.detectors import Det3DDataPreprocessor
    g = load('continuous_det')
    data, res = continuous_batch()
    assert [len(p) for p in res['points']] == g['sweep_sizes'].tolist()
    assert [len(l) for l in res['gt_labels_3d']] == g['sweep_gt_counts'].tolist()
    assert np.asarray(res['gt_labels_3d'][-1]).tolist() == g['sweep_gt_labels_last'].tolist()
    assert res['points'][1].data_ptr() == res['points'][0].data_ptr(), 'prefixes are views into one buffer'
    split = Det3DDataPreprocessor.split_batchwise(data['data_samples'])
    assert len(split) == 3 and [len(s.gt_instances_3d.bboxes_3d) for s in split] == g['sweep_gt_counts'].tolist()
    assert all(s.metainfo['depth2img'] is data['data_samples'][0].metainfo['depth2img'] for s in split)


def test_continuous_detector_loss_and_gradients_match_reference():
    from oracle import model_ref as M
    g = load('continuous_det')
    cfg = continuous_config()
    _, sd = product_state_dict(cfg, g, adjust_fcaf3d_head)
    from embodiedscan_b200.detectors import Det3DDataPreprocessor
    data, res = continuous_batch()
    samples = Det3DDataPreprocessor.split_batchwise(data['data_samples'])
    imgs = M.preprocess_imgs(torch.stack(data['inputs']['img']), cfg['data_preprocessor']['mean'],
                             cfg['data_preprocessor']['std'])
    for ref_name, own in CONT_WATCH.items():
        k = ref_name          # the oracle reads the reference-named state dict
        sd[k] = sd[k].clone().requires_grad_(True)
    out = M.detector_loss(sd, cfg, list(res['points']), imgs, samples, continuous=True)
    sum(out.values()).backward()
    for k in ('loss_center', 'loss_bbox', 'loss_cls'):
        assert rel(out[k], g['a_' + k]) <= 2e-5, (k, float(out[k]), float(g['a_' + k]))
    for ref_name, own in CONT_WATCH.items():
        grad = sd[ref_name].grad
        want = torch.from_numpy(g[f'a_grad/{ref_name}'])
        got = sampled(grad).reshape(want.shape)
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 2e-4 * scale, (ref_name, float((got - want).abs().max()), scale)
        assert rel(grad.double().norm(), g[f'a_gradnorm/{ref_name}']) <= 5e-4, ref_name


def continuous_occ_batch(scan=8):
    from embodiedscan_b200.structures import Det3DDataSample, EulerDepthInstance3DBoxes, InstanceData
    from embodiedscan_b200.transforms import ConstructMultiSweeps
    ci = continuous_occ_inputs(scan)
    res = ConstructMultiSweeps()(dict(points=ci['points'].clone(), points_slice_indices=ci['points_slice_indices'],
                                      gt_bboxes_3d=EulerDepthInstance3DBoxes(ci['boxes'].clone(), box_dim=9),
                                      gt_labels_3d=ci['labels'].clone(),
                                      visible_instance_masks=ci['visible_instance_masks'],
                                      visible_occupancy_masks=ci['visible_occupancy_masks']))
    ds = Det3DDataSample(metainfo=dict(ci['meta']))
    gt = InstanceData()
    gt.bboxes_3d, gt.labels_3d = res['gt_bboxes_3d'], res['gt_labels_3d']
    ds.gt_instances_3d = gt
    ds.gt_occupancy = ci['gt_occupancy'].clone()
    ds.gt_occupancy_masks = [torch.from_numpy(np.asarray(m)) for m in res['gt_occupancy_masks']]
    ds.eval_ann_info = None
    return dict(inputs=dict(points=[[p] for p in res['points']], img=[ci['img']]), data_samples=[ds]), res


def continuous_occ_config():
    cfg = occ_config()
    cfg['type'] = 'EmbodiedOccPredictor'
    cfg['data_preprocessor'] = dict(cfg['data_preprocessor'], batchwise_inputs=True)
    return cfg


OCC_CONT_WATCH = ('bbox_head.occ.0.weight', 'bbox_head.occ.2.weight', 'neck.lateral_convs.0.conv.weight')


def test_continuous_occupancy_loss_and_gradients_match_reference():
    from embodiedscan_b200.detectors import Det3DDataPreprocessor
    from oracle import model_ref as M
    from oracle import occ_ref as R
    g = load('continuous_occ')
    cfg = continuous_occ_config()
    _, sd = product_state_dict(cfg, g, lambda s: s)
    data, res = continuous_occ_batch()
    assert [int(np.asarray(m).sum()) for m in res['gt_occupancy_masks']] == g['mask_counts'].tolist()
    samples = Det3DDataPreprocessor.split_batchwise(data['data_samples'])
    assert [int(s.gt_occupancy_masks.sum()) for s in samples] == g['mask_counts'].tolist()
    imgs = M.preprocess_imgs(torch.stack(data['inputs']['img']), cfg['data_preprocessor']['mean'],
                             cfg['data_preprocessor']['std'])
    for k in OCC_CONT_WATCH:
        sd[k] = sd[k].clone().requires_grad_(True)
    out = R.occ_loss(sd, cfg, list(res['points']), imgs, samples, continuous=True)
    sum(out.values()).backward()
    for k in ('loss_occ_0', 'loss_occ_1', 'loss_occ_2'):
        assert rel(out[k], g['a_' + k]) <= 2e-5, (k, float(out[k]), float(g['a_' + k]))
    for k in OCC_CONT_WATCH:      # parameters downstream of the 4-voxel BatchNorm layers (see the occupancy fixture)
        want = torch.from_numpy(g[f'a_grad/{k}'])
        got = sampled(sd[k].grad).reshape(want.shape)
        if k.startswith('bbox_head'):
            assert float((got - want).abs().max()) <= 2e-4 * float(want.abs().max()), k
        else:
            cos = float(torch.dot(got.flatten().double(), want.flatten().double()) / (got.norm() * want.norm()))
            assert cos >= 0.999, (k, cos)


def test_grounding_and_occupancy_metrics_match_reference():
    """Host logic of GroundingMetric (IoU injected; the CUDA IoU is checked in -m gpu) and OccupancyMetric (pure torch
    bincounts, same code on CPU and GPU tensors)."""
    import json
    from embodiedscan_b200.evaluation import GroundingMetric, OccupancyMetric
    from oracle import eval_ref as E
    g = load('metrics')
    dets, anns = grounding_metric_inputs()
    got = GroundingMetric(iou_thr=[0.25, 0.5]).ground_eval(
        anns, dets, iou_fn=lambda p, q: torch.from_numpy(E.iou_matrix(p.numpy(), q.numpy())))
    want = json.loads(str(g['grounding_json']))
    assert set(got) == set(want) and 0.0 < want['Overall@0.5'] < want['Overall@0.25'] < 1.0
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-12, (k, got[k], want[k])
    classes, samples = occupancy_metric_inputs()
    m = OccupancyMetric()
    m.dataset_meta = dict(classes=classes)
    m.process(None, samples)
    got = m.evaluate()
    want = json.loads(str(g['occupancy_json']))
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-12, (k, got[k], want[k])


@pytest.mark.parametrize('coder', ['baseline', 'FCAF'])
@pytest.mark.parametrize('nreg', [9, 12])
def test_grounding_box_coders_match_reference(coder, nreg):
    """GroundingHead._bbox_pred_to_bbox of the product (torch ops, device-agnostic) against the reference's."""
    from embodiedscan_b200.grounding import GroundingHead
    g = load('functions')
    pts, reg = box_coder_inputs()
    gh = GroundingHead.__new__(GroundingHead)
    gh.box_coder = coder
    got = gh._bbox_pred_to_bbox(pts.clone(), reg[..., :nreg].clone())
    want = torch.from_numpy(g[f'coder_{coder}_{nreg}'])
    assert got.shape == want.shape
    assert float((got[..., :6] - want[..., :6]).abs().max()) <= 1e-5
    dang = torch.remainder(got[..., 6:] - want[..., 6:] + np.pi, 2 * np.pi) - np.pi
    assert float(dang.abs().max()) <= 1e-5


def test_reference_checkpoint_loader(tmp_path):
    """An mmengine-style checkpoint file with the REFERENCE's parameter names (manifest read off its own modules), a DDP
    `module.` prefix and no num_batches_tracked buffers loads strictly into the product detector."""
    from embodiedscan_b200 import MODELS
    from embodiedscan_b200.checkpoint import load_reference_checkpoint
    g = load('detector_g1')
    ref_sd = adjust_fcaf3d_head(fill_state_dict(manifest(g)))
    ckpt = dict(meta=dict(epoch=12), state_dict={'module.' + k: v for k, v in ref_sd.items()
                                                 if not k.endswith('num_batches_tracked')}, optimizer=dict())
    path = str(tmp_path / 'mv-3ddet.pth')
    torch.save(ckpt, path)
    model = MODELS.build(det_config())
    missing, unexpected = load_reference_checkpoint(model, path)
    assert missing == [] and unexpected == []
    sd = model.state_dict()
    assert torch.equal(sd['backbone.layer1.0.conv1.weight'], ref_sd['backbone.layer1.0.conv1.weight'])
    assert torch.equal(sd['backbone_3d.layer2.0.conv1.kernel'], ref_sd['backbone_3d.layer2.0.conv1.kernel'])
    bad = dict(state_dict={k: v for k, v in ref_sd.items() if 'conv_cls' not in k})
    with pytest.raises(RuntimeError, match='missing'):
        load_reference_checkpoint(MODELS.build(det_config()), bad)
