"""Detection evaluation on the device: the step right after the hot path.

Mirrors ``indoor_eval`` (embodiedscan/eval/indoor_eval.py:225-310) and its helpers ``eval_map_recall`` (:185-222),
``eval_det_cls`` (:57-182), ``average_precision`` (:7-54, 'area' mode) with the same arguments and the same flat
result dict, but not their shape: the reference walks ``classes x images`` in Python and calls pytorch3d's
``box3d_overlap`` once per (class, image). Here

* the 9-DoF IoU of EVERY prediction against EVERY ground-truth box of its scan comes from one launch of
  ``esb_box3d_overlap`` per scan (csrc/iou3d.cu, through ``EulerDepthInstance3DBoxes.overlaps``), masked to equal
  labels afterwards — the only heavy arithmetic, and it stays on the GPU;
* the greedy true-positive marking is a closed form: after a stable descending sort by score, a detection is a true
  positive at threshold t iff its best same-class IoU exceeds t and it is the FIRST such detection for that
  (scan, ground-truth box) — one ``np.unique`` per threshold instead of a Python loop over detections.

Frozen where the reference is ambiguous: ``np.argsort(-confidence)`` (quicksort) leaves equal scores unordered; the
rule here is a stable sort in input order (scans in order, detections in order).
"""
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from .registry import Registry

METRICS = Registry('metric')


def average_precision(recalls: np.ndarray, precisions: np.ndarray) -> np.float32:
    """Area under the monotone precision envelope (indoor_eval.py:27-39)."""
    mrec = np.concatenate(([0.], recalls, [1.]))
    mpre = np.concatenate(([0.], precisions, [0.]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    ind = np.where(mrec[1:] != mrec[:-1])[0]
    return np.float32(np.sum((mrec[ind + 1] - mrec[ind]) * mpre[ind + 1]))


def _device_iou(pred9: torch.Tensor, gt9: torch.Tensor) -> torch.Tensor:
    from .structures import EulerDepthInstance3DBoxes
    if not torch.cuda.is_available():
        raise RuntimeError('indoor_eval computes the 9-DoF IoU in libesb200.so: a CUDA device is required '
                           '(there is no CPU fallback)')
    dev = torch.device('cuda', torch.cuda.current_device())
    a = EulerDepthInstance3DBoxes(pred9.to(dev), box_dim=9, origin=(.5, .5, .5))
    b = EulerDepthInstance3DBoxes(gt9.to(dev), box_dim=9, origin=(.5, .5, .5))
    return EulerDepthInstance3DBoxes.overlaps(a, b)


def _as_boxes9(x) -> torch.Tensor:
    t = x.tensor if hasattr(x, 'tensor') else torch.as_tensor(np.asarray(x))
    return t.detach().float().reshape(-1, 9).cpu()


def _gt_boxes9(x) -> torch.Tensor:
    if isinstance(x, (list, tuple)):                       # the reference also accepts a list of single boxes
        return torch.cat([_as_boxes9(b) for b in x]) if len(x) else torch.zeros((0, 9))
    return _as_boxes9(x)


def best_same_class_iou(gt_annos: Sequence[dict], dt_annos: Sequence[dict], iou_fn: Callable):
    """Per detection: (scan id, label, score, best IoU against the same-class boxes of its scan, index of that box)."""
    scan, label, score, best, arg = [], [], [], [], []
    gt_labels_all = []
    for i, (ga, da) in enumerate(zip(gt_annos, dt_annos)):
        gl = torch.as_tensor(np.asarray(ga['gt_labels_3d'])).long().reshape(-1)
        gt_labels_all.append(gl.numpy())
        pl = torch.as_tensor(da['labels_3d']).long().reshape(-1).cpu()
        m = pl.numel()
        if m == 0:
            continue
        pb = _as_boxes9(da['bboxes_3d']).clone()
        # indoor_eval.py:118-123: a prediction with a face area below 2e-4 gets every edge clamped to >= 2e-2
        w, l, h = pb[:, 3], pb[:, 4], pb[:, 5]
        thin = (w * l < 2e-4) | (w * h < 2e-4) | (h * l < 2e-4)
        pb[thin, 3:6] = pb[thin, 3:6].clamp(min=2e-2)
        gb = _gt_boxes9(ga['gt_bboxes_3d'])
        if gb.shape[0]:
            iou = iou_fn(pb, gb).float().cpu()
            iou = torch.where(pl[:, None] == gl[None, :], iou, torch.full_like(iou, -1.0))
            # strict '>' scan over j (indoor_eval.py:158-163) = first maximum; no same-class box -> -inf
            b, a = iou.max(dim=1)
            has = (pl[:, None] == gl[None, :]).any(1)
            b = torch.where(has, b, torch.full_like(b, float('-inf')))
        else:
            b, a = torch.full((m, ), float('-inf')), torch.zeros(m, dtype=torch.long)
        scan.append(np.full(m, i))
        label.append(pl.numpy())
        score.append(torch.as_tensor(da['scores_3d']).double().reshape(-1).cpu().numpy())
        best.append(b.numpy())
        arg.append(a.numpy())
    cat = (lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt))
    return (cat(scan, np.int64), cat(label, np.int64), cat(score, np.float64), cat(best, np.float32),
            cat(arg, np.int64), gt_labels_all)


def eval_map_recall(gt_annos, dt_annos, metric: Sequence[float], iou_fn: Optional[Callable] = None):
    """-> (rec, prec, ap): per threshold a dict label -> recall array / precision array / AP, like
    embodiedscan/eval/indoor_eval.py:185-222."""
    iou_fn = iou_fn or _device_iou
    scan, label, score, best, arg, gt_labels = best_same_class_iou(gt_annos, dt_annos, iou_fn)
    npos: Dict[int, int] = {}
    for gl in gt_labels:
        for lb, c in zip(*np.unique(gl, return_counts=True)):
            npos[int(lb)] = npos.get(int(lb), 0) + int(c)
    order_of_classes: Dict[int, None] = {}                  # insertion order of the reference's `gt` dict (:254-281)
    for ga, da in zip(gt_annos, dt_annos):
        for x in torch.as_tensor(da['labels_3d']).reshape(-1).tolist() + \
                torch.as_tensor(np.asarray(ga['gt_labels_3d'])).reshape(-1).tolist():
            order_of_classes.setdefault(int(x))
    gt_classes = list(order_of_classes)
    rec = [dict() for _ in metric]
    prec = [dict() for _ in metric]
    ap = [dict() for _ in metric]
    max_gt = max([len(g) for g in gt_labels] + [1])
    for lb in gt_classes:
        sel = np.nonzero(label == lb)[0]
        if sel.size == 0:                                   # class without predictions (indoor_eval.py:217-220)
            for t in range(len(metric)):
                rec[t][lb], prec[t][lb], ap[t][lb] = np.zeros(1), np.zeros(1), np.zeros(1)
            continue
        order = sel[np.argsort(-score[sel], kind='stable')]
        key = scan[order] * max_gt + arg[order]             # (scan, matched ground-truth box)
        n = float(npos.get(lb, 0))
        for t, thr in enumerate(metric):
            hit = best[order] > thr
            tp = np.zeros(order.size)
            first = np.unique(key[hit], return_index=True)[1]
            tp[np.nonzero(hit)[0][first]] = 1.0
            ctp, cfp = np.cumsum(tp), np.cumsum(1.0 - tp)
            with np.errstate(divide='ignore', invalid='ignore'):
                recall = ctp / n
            precision = ctp / np.maximum(ctp + cfp, np.finfo(np.float64).eps)
            rec[t][lb], prec[t][lb] = recall, precision
            ap[t][lb] = np.array([average_precision(recall, precision)], dtype=np.float32)
    return rec, prec, ap


def indoor_eval(gt_annos: List[dict], dt_annos: List[dict], metric: Sequence[float], label2cat, logger=None,
                box_mode_3d=None, classes_split=None, iou_fn: Optional[Callable] = None) -> Dict[str, float]:
    """Same call and result keys as the reference's ``indoor_eval``: ``{cat}_AP_{thr}``, ``{cat}_rec_{thr}``,
    ``mAP_{thr}``, ``mAR_{thr}`` (+ ``{split}_mAP_/mAR_{thr}`` when ``classes_split`` is given; the reference only
    prints those). ``gt_annos[i]``: ``gt_bboxes_3d`` (box container or (n,9)), ``gt_labels_3d``; ``dt_annos[i]``:
    ``bboxes_3d``, ``scores_3d``, ``labels_3d``. ``iou_fn(pred (m,9), gt (n,9)) -> (m,n)`` defaults to the CUDA kernel."""
    assert len(dt_annos) == len(gt_annos)
    rec, prec, ap = eval_map_recall(gt_annos, dt_annos, metric, iou_fn)
    for key in list(ap[0].keys()):                          # classes without ground truth: recall = 0/0
        if np.isnan(ap[0][key][0]):
            for d in rec + prec + ap:
                del d[key]
    ret = {}
    for i, thr in enumerate(metric):
        for lb in ap[i]:
            ret[f'{label2cat[lb]}_AP_{thr:.2f}'] = float(ap[i][lb][0])
        ret[f'mAP_{thr:.2f}'] = float(np.mean(list(ap[i].values())))
        rec_list = []
        for lb in rec[i]:
            ret[f'{label2cat[lb]}_rec_{thr:.2f}'] = float(rec[i][lb][-1])
            rec_list.append(rec[i][lb][-1])
        ret[f'mAR_{thr:.2f}'] = float(np.mean(rec_list))
    if classes_split is not None:
        for name, members in zip(('head', 'common', 'tail'), classes_split):
            for i, thr in enumerate(metric):
                aps = [float(ap[i][lb][0]) for lb in members if lb in ap[i]]
                recs = [float(rec[i][lb][-1]) for lb in members if lb in rec[i]]
                if aps:
                    ret[f'{name}_mAP_{thr:.2f}'] = float(np.mean(aps))
                    ret[f'{name}_mAR_{thr:.2f}'] = float(np.mean(recs))
    return ret


@METRICS.register_module()
class IndoorDetMetric:
    """``embodiedscan/eval/metrics/det_metric.py`` IndoorDetMetric: collects ``(eval_ann_info, pred_instances_3d)``
    pairs from ``process`` and evaluates them with :func:`indoor_eval`."""

    def __init__(self, iou_thr=(0.25, 0.5), collect_device='cpu', prefix=None, batchwise_anns=False, **kwargs):
        self.iou_thr = [iou_thr] if isinstance(iou_thr, float) else list(iou_thr)
        self.prefix, self.results, self.dataset_meta = prefix, [], {}

    def process(self, data_batch, data_samples) -> None:
        for ds in data_samples:
            get = ds.get if hasattr(ds, 'get') else ds.__getitem__
            pred = get('pred_instances_3d')
            ann = get('eval_ann_info')
            self.results.append((ann, dict(bboxes_3d=pred['bboxes_3d'] if isinstance(pred, dict) else pred.bboxes_3d,
                                           scores_3d=pred['scores_3d'] if isinstance(pred, dict) else pred.scores_3d,
                                           labels_3d=pred['labels_3d'] if isinstance(pred, dict) else pred.labels_3d)))

    def compute_metrics(self, results=None) -> Dict[str, float]:
        results = self.results if results is None else results
        anns, preds = zip(*results) if results else ((), ())
        out = indoor_eval(list(anns), list(preds), self.iou_thr, self.dataset_meta['classes'],
                          classes_split=self.dataset_meta.get('classes_split'))
        return {'/'.join((self.prefix, k)): v for k, v in out.items()} if self.prefix else out

    def evaluate(self, size=None) -> Dict[str, float]:
        out = self.compute_metrics()
        self.results.clear()
        return out


@METRICS.register_module()
class GroundingMetric:
    """``embodiedscan/eval/metrics/grounding_metric.py``: a prompt counts as found at threshold t when one of its 10
    highest-scoring boxes overlaps a target box with 9-DoF IoU > t; accuracy overall and per Easy/Hard,
    View-Dep/View-Indep, Unique/Multi split. The IoU of the 10 candidates runs in ``esb_box3d_overlap``.
    Frozen: the top-10 come from a STABLE descending sort (torch's default argsort leaves ties unordered)."""

    TYPES = ('Easy', 'Hard', 'View-Dep', 'View-Indep', 'Unique', 'Multi', 'Overall')

    def __init__(self, iou_thr=(0.25, 0.5), collect_device='cpu', prefix=None, format_only=False, result_dir='', **kw):
        self.iou_thr = [iou_thr] if isinstance(iou_thr, float) else list(iou_thr)
        self.prefix, self.results = prefix, []

    def process(self, data_batch, data_samples) -> None:
        for ds in data_samples:
            get = ds.get if hasattr(ds, 'get') else ds.__getitem__
            pred = get('pred_instances_3d')
            pred = dict(pred.items()) if hasattr(pred, 'items') else {k: getattr(pred, k) for k in pred.keys()}
            self.results.append((get('eval_ann_info'), pred))

    def ground_eval(self, gt_annos, det_annos, iou_fn: Optional[Callable] = None) -> Dict[str, float]:
        assert len(det_annos) == len(gt_annos)
        iou_fn = iou_fn or _device_iou
        pred = {f'{o}@{t}': 0 for t in self.iou_thr for o in self.TYPES}
        gt = {f'{o}@{t}': 1e-14 for t in self.iou_thr for o in self.TYPES}
        for det, ann in zip(det_annos, gt_annos):
            scores = torch.as_tensor(det['target_scores_3d']).reshape(-1)
            top = torch.sort(scores, descending=True, stable=True).indices[:10].cpu()
            iou = iou_fn(_as_boxes9(det['bboxes_3d'])[top], _gt_boxes9(ann['gt_bboxes_3d'])).float().cpu()
            tags = ('View-Dep' if ann['is_view_dep'] else 'View-Indep', 'Hard' if ann['is_hard'] else 'Easy',
                    'Unique' if ann['is_unique'] else 'Multi', 'Overall')
            for t in self.iou_thr:
                found = int(bool((iou > t).any()))
                for tag in tags:
                    gt[f'{tag}@{t}'] += 1
                    pred[f'{tag}@{t}'] += found
        return {k: pred[k] / max(gt[k], 1) for t in self.iou_thr for k in (f'{o}@{t}' for o in self.TYPES)}

    def compute_metrics(self, results=None) -> Dict[str, float]:
        results = self.results if results is None else results
        anns, preds = zip(*results) if results else ((), ())
        out = self.ground_eval(list(anns), list(preds))
        return {'/'.join((self.prefix, k)): v for k, v in out.items()} if self.prefix else out

    def evaluate(self, size=None):
        out = self.compute_metrics()
        self.results.clear()
        return out


@METRICS.register_module()
class OccupancyMetric:
    """``embodiedscan/eval/metrics/occupancy_metric.py``: per-class IoU of the arg-max occupancy (class 0 = geometry:
    occupied vs empty), voxels with ground truth 255 (invisible) ignored. The per-scan counts are three ``bincount``s
    on whatever device holds the prediction instead of ``num_class`` masked passes."""

    def __init__(self, collect_device='cpu', prefix=None, batchwise_anns=False, **kw):
        self.prefix, self.results, self.dataset_meta = prefix, [], {}

    def process(self, data_batch, data_samples) -> None:
        for ds in data_samples:
            get = ds.get if hasattr(ds, 'get') else ds.__getitem__
            pred = get('pred_occupancy')
            gt4 = get('gt_occupancy').long().to(pred.device)
            gt = torch.zeros_like(pred)
            gt[gt4[:, 0], gt4[:, 1], gt4[:, 2]] = gt4[:, 3].to(pred.dtype)
            if 'gt_occupancy_masks' in ds:
                gt[~get('gt_occupancy_masks').to(pred.device)] = 255
            self.results.append((gt, pred))

    def compute_metrics(self, results=None) -> Dict[str, float]:
        results = self.results if results is None else results
        classes = self.dataset_meta['classes']
        n = len(classes) + 1
        score = torch.zeros((n, 3), dtype=torch.float64)
        for gt, pred in results:
            keep = gt != 255
            g, p = gt[keep].long().clamp(max=n - 1), pred[keep].long().clamp(max=n - 1)
            tp = torch.bincount(g[g == p], minlength=n)
            cg, cp = torch.bincount(g, minlength=n), torch.bincount(p, minlength=n)
            cnt = torch.stack([tp, cg, cp], 1).double().cpu()
            cnt[0] = torch.tensor([float(((g != 0) & (p != 0)).sum()), float((g != 0).sum()), float((p != 0).sum())])
            score += cnt
        ret = {}
        for i in range(n):
            tp, a, b = score[i].tolist()
            union = a + b - tp
            if union == 0:                       # the reference skips classes whose IoU is 0/0
                continue
            ret['empty' if i == 0 else classes[i - 1]] = tp / union
        return {'/'.join((self.prefix, k)): v for k, v in ret.items()} if self.prefix else ret

    def evaluate(self, size=None):
        out = self.compute_metrics()
        self.results.clear()
        return out
