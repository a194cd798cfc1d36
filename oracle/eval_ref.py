"""Oracle (test infrastructure): detection evaluation, the step right after the hot path.

Restates embodiedscan/eval/indoor_eval.py:7-54 (average_precision, 'area' mode), :57-182 (eval_det_cls),
:185-222 (eval_map_recall) and :225-310 (indoor_eval result dict) as plain loops over numpy arrays. Boxes are (n, 9)
arrays; the 9-DoF IoU is `geometry_ref.box3d_overlap` on the container's corners (euler_box3d.py:103-135).
Frozen where the reference is ambiguous: detections are ranked by a STABLE descending sort (np.argsort(-conf) with
quicksort leaves ties undefined). Pinned by tests/golden/eval.npz.
"""
import numpy as np
import torch

from . import geometry_ref as G


def average_precision(recalls, precisions):
    mrec = np.concatenate(([0.], recalls, [1.]))
    mpre = np.concatenate(([0.], precisions, [0.]))
    for i in range(mpre.shape[0] - 1, 0, -1):
        mpre[i - 1] = max(mpre[i - 1], mpre[i])
    ind = np.where(mrec[1:] != mrec[:-1])[0]
    return np.float32(np.sum((mrec[ind + 1] - mrec[ind]) * mpre[ind + 1]))


def clamp_thin(boxes):
    """indoor_eval.py:118-123: a predicted box with a face area below 2e-4 gets its edges clamped to >= 2e-2."""
    b = np.array(boxes, dtype=np.float32, copy=True).reshape(-1, 9)
    w, l, h = b[:, 3], b[:, 4], b[:, 5]
    thin = (w * l < 2e-4) | (w * h < 2e-4) | (h * l < 2e-4)
    b[thin, 3:6] = np.maximum(b[thin, 3:6], np.float32(2e-2))
    return b


def iou_matrix(pred9, gt9):
    if len(pred9) == 0 or len(gt9) == 0:
        return np.zeros((len(pred9), len(gt9)), np.float32)
    c1 = G.container_corners(torch.from_numpy(np.asarray(pred9, np.float32))).numpy().astype(np.float64)
    c2 = G.container_corners(torch.from_numpy(np.asarray(gt9, np.float32))).numpy().astype(np.float64)
    return G.box3d_overlap(c1, c2)[1].astype(np.float32)


def eval_det_cls(pred, gt, iou_thr):
    """pred: {img: [(box9, score), ...]}, gt: {img: [box9, ...]} of ONE class -> [(recall, precision, ap)] per thr."""
    npos = sum(len(v) for v in gt.values())
    det = {img: [[False] * len(v) for _ in iou_thr] for img, v in gt.items()}
    image_ids, conf, ious = [], [], []
    for img, lst in pred.items():
        if not lst:
            continue
        boxes = clamp_thin(np.stack([b for b, _ in lst]))
        m = iou_matrix(boxes, np.stack(gt[img]) if len(gt[img]) else np.zeros((0, 9), np.float32))
        for i, (_, s) in enumerate(lst):
            image_ids.append(img)
            conf.append(s)
            ious.append(m[i] if len(gt[img]) else np.zeros(1))
    order = np.argsort(-np.asarray(conf, dtype=np.float64), kind='stable')
    tp = [np.zeros(len(order)) for _ in iou_thr]
    fp = [np.zeros(len(order)) for _ in iou_thr]
    for d, x in enumerate(order):
        img, cur = image_ids[x], ious[x]
        iou_max, jmax = -np.inf, -1
        for j in range(len(gt[img])):
            if cur[j] > iou_max:
                iou_max, jmax = cur[j], j
        for t, thr in enumerate(iou_thr):
            if iou_max > thr and not det[img][t][jmax]:
                tp[t][d] = 1.
                det[img][t][jmax] = True
            else:
                fp[t][d] = 1.
    out = []
    for t in range(len(iou_thr)):
        ctp, cfp = np.cumsum(tp[t]), np.cumsum(fp[t])
        with np.errstate(divide='ignore', invalid='ignore'):
            recall = ctp / float(npos)
        precision = ctp / np.maximum(ctp + cfp, np.finfo(np.float64).eps)
        out.append((recall, precision, average_precision(recall, precision)))
    return out


def indoor_eval(gt_annos, dt_annos, metric, label2cat):
    """gt_annos[i] = dict(gt_bboxes_3d (n,9) array, gt_labels_3d (n,) ints); dt_annos[i] = dict(bboxes_3d (m,9),
    scores_3d (m,), labels_3d (m,)). Returns the reference's flat result dict (per-class AP / rec, mAP, mAR per thr)."""
    pred, gt = {}, {}
    for img, (ga, da) in enumerate(zip(gt_annos, dt_annos)):
        for b, s, l in zip(np.asarray(da['bboxes_3d']), np.asarray(da['scores_3d']), np.asarray(da['labels_3d'])):
            pred.setdefault(int(l), {}).setdefault(img, []).append((b, float(s)))
            gt.setdefault(int(l), {}).setdefault(img, [])
        for b, l in zip(np.asarray(ga['gt_bboxes_3d']), np.asarray(ga['gt_labels_3d'])):
            gt.setdefault(int(l), {}).setdefault(img, []).append(b)
    res = {}
    for label in gt:
        if label in pred:
            res[label] = eval_det_cls(pred[label], gt[label], metric)
        else:
            res[label] = [(np.zeros(1), np.zeros(1), np.float32(0.)) for _ in metric]
    keep = [l for l in gt if not np.isnan(res[l][0][2])]
    ret = {}
    for t, thr in enumerate(metric):
        for l in keep:
            ret[f'{label2cat[l]}_AP_{thr:.2f}'] = float(res[l][t][2])
            ret[f'{label2cat[l]}_rec_{thr:.2f}'] = float(res[l][t][0][-1])
        ret[f'mAP_{thr:.2f}'] = float(np.mean([res[l][t][2] for l in keep]))
        ret[f'mAR_{thr:.2f}'] = float(np.mean([res[l][t][0][-1] for l in keep]))
    return ret
