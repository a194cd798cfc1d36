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
