"""Oracle (test infrastructure): rotations, box corners, chamfer, BEV rotated IoU and NMS, focal loss.

Restates
  * pytorch3d.transforms.euler_angles_to_matrix / matrix_to_euler_angles, convention 'ZXY' (†upstream pytorch3d 0.7.x)
  * embodiedscan/models/dense_heads/fcaf3d_head.py:1728-1750 (normalize_vector / cross_product / ortho_6d_2_Mat)
  * embodiedscan/models/losses/chamfer_distance.py:13-79,160-203,240-285
  * embodiedscan/structures/bbox_3d/euler_box3d.py:137-184 (corner order), bbox_3d/utils.py:32-86
  * mmcv.ops.nms3d / nms3d_normal (†upstream mmcv 2.0.0rc4 iou3d kernels: BEV polygon clipping, EPS 1e-8, margin 1e-2)
  * mmcv.ops.sigmoid_focal_loss CUDA forward as wrapped by mmdet.FocalLoss (†upstream)
"""
import math

import numpy as np
import torch


def axis_rot(axis, angle):
    c, s = torch.cos(angle), torch.sin(angle)
    o, z = torch.ones_like(angle), torch.zeros_like(angle)
    if axis == 'X':
        flat = (o, z, z, z, c, -s, z, s, c)
    elif axis == 'Y':
        flat = (c, z, s, z, o, z, -s, z, c)
    else:
        flat = (c, -s, z, s, c, z, z, z, o)
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_to_matrix(e, convention='ZXY'):
    m = [axis_rot(c, a) for c, a in zip(convention, torch.unbind(e, -1))]
    return m[0] @ m[1] @ m[2]


def matrix_to_euler_zxy(m):
    return torch.stack((torch.atan2(-m[..., 0, 1], m[..., 1, 1]), torch.asin(m[..., 2, 1]),
                        torch.atan2(-m[..., 2, 0], m[..., 2, 2])), -1)


def ortho_6d_2_mat(x_raw, y_raw):
    y = y_raw / (torch.norm(y_raw, dim=1, keepdim=True) + 1e-8)
    z = torch.cross(x_raw, y, dim=1)
    z = z / (torch.norm(z, dim=1, keepdim=True) + 1e-8)
    x = torch.cross(y, z, dim=1)
    return torch.cat((x.unsqueeze(2), y.unsqueeze(2), z.unsqueeze(2)), 2)


def rotation_3d_in_euler(points, angles):
    return torch.bmm(points, euler_to_matrix(angles).transpose(-2, -1)) if points.shape[0] else points


def bbox_to_corners(bbox):
    rot = euler_to_matrix(bbox[:, 6:9])
    centers = bbox[:, :3].unsqueeze(1).repeat(1, 8, 1)
    half = bbox[:, 3:6].unsqueeze(1).repeat(1, 8, 1) / 2
    sx = torch.tensor([1, 1, 1, 1, -1, -1, -1, -1.])
    sy = torch.tensor([1, 1, -1, -1, 1, 1, -1, -1.])
    sz = torch.tensor([1, -1, 1, -1, 1, -1, 1, -1.])
    eight = torch.stack((sx, sy, sz), -1)[None].repeat(bbox.shape[0], 1, 1) * half
    return centers + torch.matmul(eight, rot.transpose(1, 2))


def container_corners(boxes9):
    base = torch.from_numpy(np.stack(np.unravel_index(np.arange(8), [2] * 3), axis=1)).float()
    base = base[[0, 1, 3, 2, 4, 5, 7, 6]] - 0.5
    corners = boxes9[:, None, 3:6] * base[None]
    return rotation_3d_in_euler(corners, boxes9[:, 6:9]) + boxes9[:, None, :3]


def chamfer_l1_mean(src_boxes, dst_boxes, weight=1.0):
    """BBoxCDLoss(l1, g8, mean): mean over (N,8) of min_j sum_c |src_i - dst_j| (src->dst only)."""
    s, d = bbox_to_corners(src_boxes), bbox_to_corners(dst_boxes)
    dist = (s.unsqueeze(2) - d.unsqueeze(1)).abs().sum(-1)
    return dist.min(dim=2).values.mean() * weight


def bbox_pred_to_bbox(points, bbox_pred):
    """fcaf3d_head.py:1454-1525, 12-channel branch."""
    if bbox_pred.shape[0] == 0:
        return bbox_pred
    shift = torch.stack(((bbox_pred[:, 1] - bbox_pred[:, 0]) / 2, (bbox_pred[:, 3] - bbox_pred[:, 2]) / 2,
                         (bbox_pred[:, 5] - bbox_pred[:, 4]) / 2), -1).view(-1, 1, 3)
    euler = matrix_to_euler_zxy(ortho_6d_2_mat(bbox_pred[:, 6:9], bbox_pred[:, 9:]))
    shift = rotation_3d_in_euler(shift, euler)[:, 0, :]
    size = torch.stack((bbox_pred[:, 0] + bbox_pred[:, 1], bbox_pred[:, 2] + bbox_pred[:, 3],
                        bbox_pred[:, 4] + bbox_pred[:, 5]), -1)
    return torch.cat((points + shift, size, euler), -1)


def sigmoid_focal_loss_sum(logits, target, gamma=2.0, alpha=0.25):
    """Sum over all (row, class) of mmcv's CUDA sigmoid focal loss; target -1 => no positive class."""
    p = torch.sigmoid(logits)
    C = logits.shape[1]
    onehot = (target.view(-1, 1) == torch.arange(C).view(1, -1)).to(logits.dtype)
    tiny = torch.finfo(torch.float32).tiny
    term_p = (1 - p) ** gamma * torch.log(torch.clamp(p, min=tiny))
    term_n = p ** gamma * torch.log(torch.clamp(1 - p, min=tiny))
    return (-onehot * alpha * term_p - (1 - onehot) * (1 - alpha) * term_n).sum()


# ---------------------------------------------------------------------------------------------------------
# BEV rotated IoU / NMS in float32 scalar arithmetic (small inputs only)
# ---------------------------------------------------------------------------------------------------------
F = np.float32
EPS = F(1e-8)


def _cross3(p1, p2, p0):
    return F(F(F(p1[0] - p0[0]) * F(p2[1] - p0[1])) - F(F(p2[0] - p0[0]) * F(p1[1] - p0[1])))


def _in_box(box, p):
    margin = F(1e-2)
    ac, as_ = F(math.cos(-float(box[6]))), F(math.sin(-float(box[6])))
    rx = F(F(F(p[0] - box[0]) * ac) + F(F(p[1] - box[1]) * F(-as_)))
    ry = F(F(F(p[0] - box[0]) * as_) + F(F(p[1] - box[1]) * ac))
    return abs(rx) < F(F(box[3] / F(2)) + margin) and abs(ry) < F(F(box[4] / F(2)) + margin)


def _intersection(p1, p0, q1, q0):
    if not (min(p0[0], p1[0]) <= max(q0[0], q1[0]) and min(q0[0], q1[0]) <= max(p0[0], p1[0])
            and min(p0[1], p1[1]) <= max(q0[1], q1[1]) and min(q0[1], q1[1]) <= max(p0[1], p1[1])):
        return None
    s1, s2, s3, s4 = _cross3(q0, p1, p0), _cross3(p1, q1, p0), _cross3(p0, q1, q0), _cross3(q1, p1, q0)
    if not (F(s1 * s2) > 0 and F(s3 * s4) > 0):
        return None
    s5 = _cross3(q1, p1, p0)
    if abs(F(s5 - s1)) > EPS:
        return (F(F(F(s5 * q0[0]) - F(s1 * q1[0])) / F(s5 - s1)), F(F(F(s5 * q0[1]) - F(s1 * q1[1])) / F(s5 - s1)))
    a0, b0, c0 = F(p0[1] - p1[1]), F(p1[0] - p0[0]), F(F(p0[0] * p1[1]) - F(p1[0] * p0[1]))
    a1, b1, c1 = F(q0[1] - q1[1]), F(q1[0] - q0[0]), F(F(q0[0] * q1[1]) - F(q1[0] * q0[1]))
    D = F(F(a0 * b1) - F(a1 * b0))
    return (F(F(F(b0 * c1) - F(b1 * c0)) / D), F(F(F(a1 * c0) - F(a0 * c1)) / D))


def _corners(box):
    hx, hy = F(box[3] / F(2)), F(box[4] / F(2))
    pts = [(F(box[0] - hx), F(box[1] - hy)), (F(box[0] + hx), F(box[1] - hy)), (F(box[0] + hx), F(box[1] + hy)),
           (F(box[0] - hx), F(box[1] + hy))]
    ac, as_ = F(math.cos(float(box[6]))), F(math.sin(float(box[6])))
    out = []
    for p in pts:
        nx = F(F(F(F(p[0] - box[0]) * ac) + F(F(p[1] - box[1]) * F(-as_))) + box[0])
        ny = F(F(F(F(p[0] - box[0]) * as_) + F(F(p[1] - box[1]) * ac)) + box[1])
        out.append((nx, ny))
    return out + [out[0]]


def box_overlap_bev(a, b):
    a, b = a.astype(np.float32), b.astype(np.float32)
    A, B = _corners(a), _corners(b)
    cp, pcx, pcy = [], F(0), F(0)
    for i in range(4):
        for j in range(4):
            r = _intersection(A[i + 1], A[i], B[j + 1], B[j])
            if r is not None:
                cp.append(r)
                pcx, pcy = F(pcx + r[0]), F(pcy + r[1])
    for k in range(4):
        if _in_box(a, B[k]):
            cp.append(B[k]); pcx, pcy = F(pcx + B[k][0]), F(pcy + B[k][1])
        if _in_box(b, A[k]):
            cp.append(A[k]); pcx, pcy = F(pcx + A[k][0]), F(pcy + A[k][1])
    cnt = len(cp)
    if cnt == 0:
        return F(0)
    pcx, pcy = F(pcx / F(cnt)), F(pcy / F(cnt))
    ang = lambda p: math.atan2(float(F(p[1] - pcy)), float(F(p[0] - pcx)))
    for j in range(cnt - 1):
        for i in range(cnt - j - 1):
            if F(ang(cp[i])) > F(ang(cp[i + 1])):
                cp[i], cp[i + 1] = cp[i + 1], cp[i]
    area = F(0)
    for k in range(cnt - 1):
        ux, uy = F(cp[k][0] - cp[0][0]), F(cp[k][1] - cp[0][1])
        vx, vy = F(cp[k + 1][0] - cp[0][0]), F(cp[k + 1][1] - cp[0][1])
        area = F(area + F(F(ux * vy) - F(uy * vx)))
    return F(abs(area) / F(2))


def iou_bev(a, b):
    a, b = a.astype(np.float32), b.astype(np.float32)
    so = box_overlap_bev(a, b)
    return F(so / max(F(F(F(a[3] * a[4]) + F(b[3] * b[4])) - so), EPS))


def nms3d(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """Indices kept (into the input), in descending-score order; stable sort (ties keep input order)."""
    order = np.argsort(-scores.astype(np.float32), kind='stable')
    keep, sup = [], np.zeros(len(order), dtype=bool)
    for ii, i in enumerate(order):
        if sup[ii]:
            continue
        keep.append(i)
        for jj in range(ii + 1, len(order)):
            if not sup[jj] and iou_bev(boxes[i], boxes[order[jj]]) > F(thr):
                sup[jj] = True
    return np.array(keep, dtype=np.int64)


def multiclass_nms(bboxes: torch.Tensor, scores: torch.Tensor, score_thr: float, iou_thr: float):
    """fcaf3d_head.py:1666-1725 with_yaw branch: per class loop, 9->7 column truncation."""
    b7 = bboxes[:, :7].numpy()
    out_b, out_s, out_l = [], [], []
    for c in range(scores.shape[1]):
        ids = (scores[:, c] > score_thr).numpy()
        if not ids.any():
            continue
        cs, cb = scores[ids, c].numpy(), b7[ids]
        k = nms3d(cb, cs, iou_thr)
        out_b.append(cb[k]); out_s.append(cs[k]); out_l.append(np.full(len(k), c, dtype=np.int64))
    if not out_b:
        return torch.zeros((0, 7)), torch.zeros((0, )), torch.zeros((0, ), dtype=torch.long)
    return (torch.from_numpy(np.concatenate(out_b)), torch.from_numpy(np.concatenate(out_s)),
            torch.from_numpy(np.concatenate(out_l)))


# ---------------------------------------------------------------------------------------------------------
# exact 9-DoF box IoU (pytorch3d.ops.box3d_overlap contract, †upstream), float64 polygon clipping
# ---------------------------------------------------------------------------------------------------------
def _box_from_corners(k):
    c = k.mean(0)
    e = [k[4] - k[0], k[3] - k[0], k[1] - k[0]]
    h = [np.linalg.norm(v) / 2 for v in e]
    n = [v / (2 * hh) for v, hh in zip(e, h)]
    return c, n, h


def _clip(poly, n, d, strict):
    out = []
    for i in range(len(poly)):
        a, b = poly[i], poly[(i + 1) % len(poly)]
        da, db = n @ a - d, n @ b - d
        ina, inb = (da < -1e-12, db < -1e-12) if strict else (da <= 1e-12, db <= 1e-12)
        if ina:
            out.append(a)
        if ina != inb:
            out.append(a + (b - a) * (da / (da - db)))
    return out


def _faces_clipped(P, Q, strict):
    (pc, pn, ph), (qc, qn, qh) = P, Q
    acc = 0.0
    for a in range(3):
        for sgn in (-1.0, 1.0):
            n = pn[a] * sgn
            fc = pc + n * ph[a]
            u, v = pn[(a + 1) % 3] * ph[(a + 1) % 3], pn[(a + 2) % 3] * ph[(a + 2) % 3]
            poly = [fc - u - v, fc + u - v, fc + u + v, fc - u + v]
            for b in range(3):
                for s2 in (-1.0, 1.0):
                    if poly:
                        m = qn[b] * s2
                        poly = _clip(poly, m, m @ qc + qh[b], strict)
            if len(poly) < 3:
                continue
            av = sum((np.cross(poly[i] - poly[0], poly[i + 1] - poly[0]) for i in range(1, len(poly) - 1)), np.zeros(3))
            acc += (n @ fc) * 0.5 * abs(av @ n)
    return acc


def box3d_overlap(corners1: np.ndarray, corners2: np.ndarray):
    """(N,8,3),(M,8,3) in EulerInstance3DBoxes.corners order -> (vol, iou) float64. Faces of A∩B = faces of A clipped by B
    plus faces of B strictly clipped by A; V = 1/3 sum (n.p) area (divergence theorem), relative to A's centre."""
    n1, n2 = len(corners1), len(corners2)
    vol, iou = np.zeros((n1, n2)), np.zeros((n1, n2))
    for i in range(n1):
        o = corners1[i].astype(np.float64).mean(0)
        A = _box_from_corners(corners1[i].astype(np.float64) - o)
        va = 8 * A[2][0] * A[2][1] * A[2][2]
        for j in range(n2):
            B = _box_from_corners(corners2[j].astype(np.float64) - o)
            vb = 8 * B[2][0] * B[2][1] * B[2][2]
            v = (_faces_clipped(A, B, False) + _faces_clipped(B, A, True)) / 3.0
            v = min(max(v, 0.0), min(va, vb))
            vol[i, j], iou[i, j] = v, v / max(va + vb - v, 1e-12)
    return vol, iou
