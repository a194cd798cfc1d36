"""Box / rotation geometry of the hot path (device-agnostic torch; tiny tensors).

Restates the pytorch3d transforms the reference imports (†upstream pytorch3d 0.7.x, convention 'ZXY':
``R = Rz(a) @ Rx(b) @ Ry(c)``) and the helpers of ``embodiedscan/models/dense_heads/fcaf3d_head.py:1728-1750``
(``ortho_6d_2_Mat``), ``embodiedscan/models/losses/chamfer_distance.py:160-203`` (``bbox_to_corners``) and
``embodiedscan/structures/bbox_3d/euler_box3d.py:137-184`` (corner order of the box container).
"""
import torch


def _axis_rot(axis: str, angle: torch.Tensor) -> torch.Tensor:
    c, s = torch.cos(angle), torch.sin(angle)
    one, zero = torch.ones_like(angle), torch.zeros_like(angle)
    if axis == 'X':
        flat = (one, zero, zero, zero, c, -s, zero, s, c)
    elif axis == 'Y':
        flat = (c, zero, s, zero, one, zero, -s, zero, c)
    else:
        flat = (c, -s, zero, s, c, zero, zero, zero, one)
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(euler: torch.Tensor, convention: str = 'ZXY') -> torch.Tensor:
    mats = [_axis_rot(c, e) for c, e in zip(convention, torch.unbind(euler, -1))]
    return torch.matmul(torch.matmul(mats[0], mats[1]), mats[2])


def matrix_to_euler_angles_zxy(m: torch.Tensor) -> torch.Tensor:
    """pytorch3d.matrix_to_euler_angles(M, 'ZXY'): (alpha, beta, gamma) with
    beta = asin(M[2,1]); alpha = atan2(-M[0,1], M[1,1]); gamma = atan2(-M[2,0], M[2,2])."""
    beta = torch.asin(m[..., 2, 1])
    alpha = torch.atan2(-m[..., 0, 1], m[..., 1, 1])
    gamma = torch.atan2(-m[..., 2, 0], m[..., 2, 2])
    return torch.stack((alpha, beta, gamma), -1)


def ortho_6d_2_mat(x_raw: torch.Tensor, y_raw: torch.Tensor) -> torch.Tensor:
    """Gram-Schmidt with y first (fcaf3d_head.py:1739-1750); columns (x, y, z)."""
    y = y_raw / (torch.norm(y_raw, dim=1, keepdim=True) + 1e-8)
    z = torch.cross(x_raw, y, dim=1)
    z = z / (torch.norm(z, dim=1, keepdim=True) + 1e-8)
    x = torch.cross(y, z, dim=1)
    return torch.stack((x, y, z), 2)


def rotation_3d_in_euler(points: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    """points (N, M, 3) @ R(angles)^T  (embodiedscan/structures/bbox_3d/utils.py:32-86)."""
    rot_t = euler_angles_to_matrix(angles, 'ZXY').transpose(-2, -1)
    if points.shape[0] == 0:
        return points
    return torch.bmm(points, rot_t)


_CORNER_SIGNS = None


def bbox_to_corners(bbox: torch.Tensor) -> torch.Tensor:
    """(N, 9) -> (N, 8, 3) with the sign pattern of chamfer_distance.py:184-196."""
    n = bbox.shape[0]
    if bbox.shape[-1] == 9:
        rot = euler_angles_to_matrix(bbox[:, 6:], 'ZXY')
    elif bbox.shape[-1] == 7:
        ang = torch.cat((bbox[:, 6:], torch.zeros_like(bbox[:, 6:]).repeat(1, 2)), 1)
        rot = euler_angles_to_matrix(ang, 'ZXY')
    else:
        rot = torch.eye(3, device=bbox.device, dtype=bbox.dtype).expand(n, 3, 3)
    sx = bbox.new_tensor([1, 1, 1, 1, -1, -1, -1, -1])
    sy = bbox.new_tensor([1, 1, -1, -1, 1, 1, -1, -1])
    sz = bbox.new_tensor([1, -1, 1, -1, 1, -1, 1, -1])
    signs = torch.stack((sx, sy, sz), -1)[None]                      # (1, 8, 3)
    corners = signs * (bbox[:, None, 3:6] / 2)
    return bbox[:, None, :3] + torch.matmul(corners, rot.transpose(1, 2))


def box_corners_container(boxes9: torch.Tensor) -> torch.Tensor:
    """EulerInstance3DBoxes.corners order (euler_box3d.py:137-184): unravel(2,2,2)[[0,1,3,2,4,5,7,6]] - 0.5."""
    if boxes9.numel() == 0:
        return boxes9.new_zeros((0, 8, 3))
    base = torch.tensor([[0, 0, 0], [0, 0, 1], [0, 1, 1], [0, 1, 0], [1, 0, 0], [1, 0, 1], [1, 1, 1], [1, 1, 0]],
                        dtype=boxes9.dtype, device=boxes9.device) - 0.5
    corners = boxes9[:, None, 3:6] * base[None]
    corners = rotation_3d_in_euler(corners, boxes9[:, 6:9])
    return corners + boxes9[:, None, :3]


def chamfer_l1_src(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """(N,8,3),(N,8,3) -> (N,8) min over dst of the L1 distance (chamfer_distance.py:55-61, src->dst only)."""
    d = (src[:, :, None, :] - dst[:, None, :, :]).abs().sum(-1)
    return d.min(dim=2).values


def bbox_cd_loss(source: torch.Tensor, target: torch.Tensor, loss_weight: float = 1.0) -> torch.Tensor:
    """BBoxCDLoss(mode='l1', group='g8', reduction='mean') (chamfer_distance.py:240-285)."""
    return chamfer_l1_src(bbox_to_corners(source), bbox_to_corners(target)).mean() * loss_weight


def box3d_overlap(corners1: torch.Tensor, corners2: torch.Tensor, eps: float = 1e-4):
    """pytorch3d.ops.box3d_overlap contract: corners (N,8,3), (M,8,3) in the container's corner order -> (vol, iou),
    both (N,M). Exact convex clipping on the GPU (csrc/iou3d.cu); `eps` is accepted for signature parity."""
    from ._ffi import call, ptr, stream
    assert corners1.is_cuda, 'box3d_overlap runs in libesb200.so (no CPU fallback)'
    c1, c2 = corners1.float().contiguous(), corners2.float().contiguous()
    n1, n2 = c1.shape[0], c2.shape[0]
    vol = torch.empty((n1, n2), dtype=torch.float32, device=c1.device)
    iou = torch.empty((n1, n2), dtype=torch.float32, device=c1.device)
    call('esb_box3d_overlap', ptr(c1), n1, ptr(c2), n2, ptr(vol), ptr(iou), stream())
    return vol, iou
