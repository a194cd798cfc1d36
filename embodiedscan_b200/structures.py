"""I/O containers of the boundary: look-alikes of ``mmengine.structures.InstanceData``,
``Det3DDataElement`` (embodiedscan/utils/typing_config.py:11-35) and the 9-DoF box container
``EulerDepthInstance3DBoxes`` (embodiedscan/structures/bbox_3d/euler_box3d.py:24-58, euler_depth_box3d.py:41-47),
carrying exactly the fields the hot path reads and writes.
"""
import torch

from .geometry import box_corners_container


class _Bag:
    """Attribute bag with ``metainfo`` (BaseDataElement contract: set_metainfo / get / keys / in)."""

    def __init__(self, metainfo=None, **kwargs):
        object.__setattr__(self, '_metainfo', dict(metainfo or {}))
        object.__setattr__(self, '_data', {})
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        self._data[k] = v

    def __getattr__(self, k):
        data = object.__getattribute__(self, '_data')
        if k in data:
            return data[k]
        meta = object.__getattribute__(self, '_metainfo')
        if k in meta:
            return meta[k]
        raise AttributeError(k)

    def __contains__(self, k):
        return k in self._data or k in self._metainfo

    def get(self, k, default=None):
        return self._data.get(k, self._metainfo.get(k, default))

    def keys(self):
        return list(self._data.keys())

    @property
    def metainfo(self):
        return self._metainfo

    def set_metainfo(self, meta):
        self._metainfo.update(meta)

    def to(self, device):
        for k, v in list(self._data.items()):
            if hasattr(v, 'to'):
                self._data[k] = v.to(device)
        return self


class InstanceData(_Bag):

    def __len__(self):
        for v in self._data.values():
            if hasattr(v, '__len__'):
                return len(v)
        return 0


class Det3DDataSample(_Bag):
    """Fields used: metainfo, gt_instances_3d{bboxes_3d, labels_3d}, pred_instances_3d, pred_instances."""


Det3DDataElement = Det3DDataSample


class EulerDepthInstance3DBoxes:
    """(N, 9) = (x, y, z, dx, dy, dz, alpha, beta, gamma), gravity-centred (origin (.5,.5,.5)); 6- and 7-column inputs
    are padded with zero Euler angles exactly like the reference (euler_box3d.py:36-48)."""

    def __init__(self, tensor, box_dim=9, with_yaw=True, origin=(0.5, 0.5, 0.5)):
        tensor = torch.as_tensor(tensor, dtype=torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((0, box_dim))
        assert tensor.dim() == 2 and tensor.size(-1) == box_dim, tensor.size()
        if tensor.shape[-1] == 6:
            tensor = torch.cat((tensor, tensor.new_zeros(tensor.shape[0], 3)), -1)
        elif tensor.shape[-1] == 7:
            tensor = torch.cat((tensor, tensor.new_zeros(tensor.shape[0], 2)), -1)
        else:
            assert tensor.shape[-1] == 9
        self.box_dim = 9
        self.tensor = tensor.clone()
        self.with_yaw = with_yaw
        if tuple(origin) != (0.5, 0.5, 0.5):
            dst = self.tensor.new_tensor((0.5, 0.5, 0.5))
            src = self.tensor.new_tensor(origin)
            self.tensor[:, :3] += self.tensor[:, 3:6] * (dst - src)

    def __len__(self):
        return self.tensor.shape[0]

    @property
    def gravity_center(self):
        return self.tensor[:, :3]

    @property
    def dims(self):
        return self.tensor[:, 3:6]

    @property
    def volume(self):
        return self.tensor[:, 3] * self.tensor[:, 4] * self.tensor[:, 5]

    @property
    def corners(self):
        return box_corners_container(self.tensor)

    @property
    def device(self):
        return self.tensor.device

    @property
    def shape(self):
        return self.tensor.shape

    def to(self, device):
        out = EulerDepthInstance3DBoxes(self.tensor.to(device), box_dim=9, with_yaw=self.with_yaw)
        return out

    # ---- augmentation (euler_box3d.py:186-281, euler_depth_box3d.py:49-78, base_box3d.py:248-257); torch ops on
    # whatever device the box tensor lives on ----
    def transform(self, matrix):
        """Apply a 4x4 rigid transform: centres move, Euler angles become ZXY(matrix_R @ R(angles))."""
        from .geometry import euler_angles_to_matrix, matrix_to_euler_angles_zxy
        if self.tensor.shape[0] == 0:
            return
        matrix = torch.as_tensor(matrix, dtype=self.tensor.dtype, device=self.tensor.device)
        pts = torch.cat([self.tensor[:, :3], self.tensor.new_ones(self.tensor.shape[0], 1)], -1)
        ctr = torch.matmul(pts, matrix.transpose(-2, -1))[:, :3]
        ori = euler_angles_to_matrix(self.tensor[:, 6:], 'ZXY')
        ang = matrix_to_euler_angles_zxy(torch.bmm(matrix[:3, :3].expand_as(ori), ori))
        self.tensor = torch.cat([ctr, self.tensor[:, 3:6], ang], -1)

    def rotate(self, angle, points=None):
        """`angle`: yaw (scalar), 3 ZXY Euler angles or a 3x3 matrix. Rotates the boxes and, when given, `points`
        ((N, >=3) tensor, in place); returns (points, rot_mat_T) or rot_mat_T like the reference."""
        from .geometry import euler_angles_to_matrix
        angle = torch.as_tensor(angle, dtype=self.tensor.dtype, device=self.tensor.device)
        if angle.numel() == 1:
            rot = euler_angles_to_matrix(torch.stack([angle.reshape(()), angle.new_zeros(()), angle.new_zeros(())]), 'ZXY')
        elif angle.numel() == 3:
            rot = euler_angles_to_matrix(angle.reshape(3), 'ZXY')
        else:
            assert angle.shape == (3, 3)
            rot = angle
        m = torch.eye(4, dtype=self.tensor.dtype, device=self.tensor.device)
        m[:3, :3] = rot
        self.transform(m)
        rot_mat_T = rot.T
        if points is not None:
            points[:, :3] = points[:, :3] @ rot_mat_T.to(points.device)
            return points, rot_mat_T
        return rot_mat_T

    def flip(self, bev_direction='horizontal', points=None):
        """Depth coordinates: 'horizontal' mirrors x, 'vertical' mirrors y (boxes, and `points` in place when given)."""
        import math
        assert bev_direction in ('horizontal', 'vertical')
        if bev_direction == 'horizontal':
            self.tensor[:, 0] = -self.tensor[:, 0]
            self.tensor[:, 6] = -self.tensor[:, 6] + math.pi
            self.tensor[:, 8] = -self.tensor[:, 8]
        else:
            self.tensor[:, 1] = -self.tensor[:, 1]
            self.tensor[:, 6] = -self.tensor[:, 6]
            self.tensor[:, 7] = -self.tensor[:, 7] + math.pi
        if points is not None:
            points[:, 0 if bev_direction == 'horizontal' else 1] *= -1
            return points

    def scale(self, scale_factor: float):
        self.tensor[:, :6] *= scale_factor

    def translate(self, trans_vector):
        self.tensor[:, :3] += torch.as_tensor(trans_vector, dtype=self.tensor.dtype, device=self.tensor.device)

    @classmethod
    def overlaps(cls, boxes1, boxes2, mode='iou', eps=1e-4):
        """(N,M) 9-DoF 3D IoU (euler_box3d.py:103-135)."""
        from .geometry import box3d_overlap
        assert mode == 'iou'
        if len(boxes1) * len(boxes2) == 0:
            return boxes1.tensor.new_zeros((len(boxes1), len(boxes2)))
        return box3d_overlap(boxes1.corners, boxes2.corners, eps=eps)[1]

    def __getitem__(self, item):
        t = self.tensor[item]
        if t.dim() == 1:
            t = t.view(1, -1)
        return EulerDepthInstance3DBoxes(t, box_dim=9, with_yaw=self.with_yaw)
