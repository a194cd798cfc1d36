"""Minimal stand-ins for the reference's un-installable third-party imports (mmengine, mmcv, mmdet, pytorch3d,
MinkowskiEngine, open3d, cv2 ...), so that the reference's OWN Python (`/root/reference/embodiedscan/**`) can be imported
and executed in the dev container to produce golden vectors (`make_golden.py`).

TEST INFRASTRUCTURE ONLY.  Nothing here is imported by the product, by `-m gpu` tests, by `smoke()` or by `bench.py`;
it only runs where `/root/reference` exists.  Two kinds of stand-in:

* plumbing (registries, `BaseModule`, `InstanceData`, `multi_apply`, `reduce_mean`, `Scale` ...): behaviour-free glue,
  written from the public documentation of those packages;
* arithmetic that lives in an absent dependency (`pytorch3d.transforms` Euler conversions, `mmcv.ops.nms3d`,
  `MinkowskiEngine` sparse tensors): delegated to `oracle/` — so golden vectors pin the REFERENCE'S OWN CODE, while the
  third-party semantics stay "parity unpinned" exactly as DESIGN.md §3 says.
"""
import functools
import importlib.machinery
import sys
import types

import numpy as np
import torch
import torch.nn as nn


class _Dummy:
    """Permissive placeholder: usable as a base class, decorator, callable or attribute bag."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Dummy()


class _AutoModule(types.ModuleType):
    """A module whose unknown attributes resolve to placeholder classes and whose submodules exist on demand."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        cls = type(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


def _mod(name):
    if name in sys.modules:
        return sys.modules[name]
    m = _AutoModule(name)
    sys.modules[name] = m
    if '.' in name:
        parent, child = name.rsplit('.', 1)
        setattr(_mod(parent), child, m)
    return m


# ----------------------------------------------------------------------------------------------------- mmengine
class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class Registry:
    def __init__(self, name, parent=None, locations=None, scope=None, build_func=None):
        self.name, self.parent, self._m = name, parent, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self._m[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        key = key.split('.')[-1]
        r = self
        while r is not None:
            if key in r._m:
                return r._m[key]
            r = r.parent
        for reg in _ALL_REGISTRIES:
            if key in reg._m:
                return reg._m[key]
        raise KeyError(key)

    def build(self, cfg, **default_args):
        if cfg is None or isinstance(cfg, nn.Module):
            return cfg
        cfg = dict(cfg)
        cfg.update({k: v for k, v in default_args.items() if k not in cfg})
        t = cfg.pop('type')
        cls = self.get(t) if isinstance(t, str) else t
        return cls(**cfg)


_ALL_REGISTRIES = []


def _registry(*a, **k):
    r = Registry(*a, **k)
    _ALL_REGISTRIES.append(r)
    return r


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


class BaseModel(BaseModule):
    def __init__(self, data_preprocessor=None, init_cfg=None):
        super().__init__(init_cfg)
        self.data_preprocessor = data_preprocessor if isinstance(data_preprocessor, nn.Module) else nn.Identity()


class BaseDataElement:
    """attribute bag with `metainfo`; the subset of mmengine.structures.BaseDataElement the reference touches."""

    def __init__(self, *, metainfo=None, **kwargs):
        object.__setattr__(self, '_metainfo_fields', set())
        object.__setattr__(self, '_data_fields', set())
        if metainfo:
            self.set_metainfo(metainfo)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def set_metainfo(self, meta):
        for k, v in meta.items():
            self._metainfo_fields.add(k)
            object.__setattr__(self, k, v)

    def set_field(self, value, name, dtype=None, field_type='data'):
        (self._data_fields if field_type == 'data' else self._metainfo_fields).add(name)
        object.__setattr__(self, name, value)

    def __setattr__(self, k, v):
        if k in ('_metainfo_fields', '_data_fields'):
            object.__setattr__(self, k, v)
            return
        prop = getattr(type(self), k, None)
        if isinstance(prop, property) and prop.fset is not None:
            prop.fset(self, v)
            return
        self._data_fields.add(k)
        object.__setattr__(self, k, v)

    @property
    def metainfo(self):
        return {k: getattr(self, k) for k in self._metainfo_fields}

    def keys(self):
        return [k.lstrip('_') if isinstance(getattr(type(self), k.lstrip('_'), None), property) else k
                for k in self._data_fields]

    def metainfo_keys(self):
        return list(self._metainfo_fields)

    def get(self, k, default=None):
        return getattr(self, k, default)

    def __contains__(self, k):
        return k in self._data_fields or k in self._metainfo_fields or ('_' + k) in self._data_fields

    def to(self, *a, **k):
        return self

    def new(self, **k):
        return type(self)(metainfo=self.metainfo)

    def clone(self):
        out = type(self)(metainfo=dict(self.metainfo))
        for k in self._data_fields:
            object.__setattr__(out, k, getattr(self, k))
            out._data_fields.add(k)
        return out


class InstanceData(BaseDataElement):
    def __len__(self):
        for k in self._data_fields:
            return len(getattr(self, k))
        return 0

    def __getitem__(self, idx):
        out = type(self)(metainfo=self.metainfo)
        for k in self._data_fields:
            setattr(out, k, getattr(self, k)[idx])
        return out


def bias_init_with_prob(p):
    return float(-np.log((1 - p) / p))


def constant_init(module, val, bias=0):
    if getattr(module, 'weight', None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def _install_mmengine():
    m = _mod('mmengine')
    m.Registry = Registry
    m.ConfigDict = ConfigDict
    for n in ('DATASETS', 'METRICS', 'MODELS', 'TASK_UTILS', 'TRANSFORMS', 'VISBACKENDS', 'VISUALIZERS'):
        setattr(m, n, _registry(n.lower()))
    _mod('mmengine.config').ConfigDict = ConfigDict
    mm = _mod('mmengine.model')
    mm.BaseModule, mm.BaseModel = BaseModule, BaseModel
    mm.ModuleList, mm.Sequential = nn.ModuleList, nn.Sequential
    mm.bias_init_with_prob, mm.constant_init = bias_init_with_prob, constant_init
    ms = _mod('mmengine.structures')
    ms.BaseDataElement, ms.InstanceData = BaseDataElement, InstanceData
    ms.PixelData = type('PixelData', (BaseDataElement,), {})
    lg = _mod('mmengine.logging')
    lg.print_log = lambda *a, **k: None
    _mod('mmengine.evaluator').BaseMetric = type('BaseMetric', (), {'__init__': lambda self, *a, **k: None})
    _mod('mmengine.evaluator.metric')._to_cpu = lambda x: x
    _mod('mmengine.fileio')
    d = _mod('mmengine.dist')
    d.master_only = lambda f: f
    _mod('mmengine.dataset').BaseDataset = type('BaseDataset', (), {})
    u = _mod('mmengine.utils')
    u.is_seq_of = lambda seq, t, seq_type=None: isinstance(seq, (list, tuple)) and all(isinstance(s, t) for s in seq)
    _mod('mmengine.visualization')


# -------------------------------------------------------------------------------------------------------- mmcv
class Scale(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


class BaseTransform:
    def __call__(self, results):
        return self.transform(results)


class MultiheadAttention(nn.Module):
    """mmcv.cnn.bricks.transformer.MultiheadAttention (2.0.0rc4, restated from its published source): positional
    encodings added to query/key, `nn.MultiheadAttention` in (L, B, E) layout, residual `identity + proj(attn)`."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout_layer=None, init_cfg=None,
                 batch_first=False, dropout=None, **kwargs):
        super().__init__()
        assert not attn_drop and not proj_drop and not dropout
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, 0.0, **kwargs)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None,
                key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        if self.batch_first:
            query, key, value = query.transpose(0, 1), key.transpose(0, 1), value.transpose(0, 1)
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask, key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + out


class FFN(nn.Module):
    """mmcv.cnn.bricks.transformer.FFN: Sequential(Sequential(Linear, act, Dropout) x (num_fcs-1), Linear, Dropout),
    residual added."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=dict(type='ReLU', inplace=True),
                 ffn_drop=0., dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__()
        assert num_fcs == 2 and not ffn_drop and add_identity
        self.embed_dims = embed_dims
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True),
                                                  nn.Dropout(0.)),
                                    nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.))

    def forward(self, x, identity=None):
        return (x if identity is None else identity) + self.layers(x)


def _nms3d(boxes, scores, iou_threshold):
    from oracle import geometry_ref as G
    keep = G.nms3d(boxes.detach().cpu().numpy().astype(np.float32), scores.detach().cpu().numpy(), iou_threshold)
    return torch.as_tensor(np.asarray(keep), dtype=torch.long)


def _install_mmcv():
    _mod('mmcv')
    _mod('mmcv.transforms').BaseTransform = BaseTransform
    ops = _mod('mmcv.ops')
    ops.nms3d = _nms3d
    cnn = _mod('mmcv.cnn')
    cnn.Scale = Scale
    cnn.Linear = nn.Linear

    def build_conv_layer(cfg, *args, **kwargs):
        cfg = dict(cfg or dict(type='Conv2d'))
        layer = {'Conv1d': nn.Conv1d, 'Conv2d': nn.Conv2d, 'Conv3d': nn.Conv3d, 'Conv': nn.Conv2d}[cfg.pop('type')]
        return layer(*args, **kwargs, **cfg)
    cnn.build_conv_layer = build_conv_layer

    def build_norm_layer(cfg, num_features, postfix=''):
        cfg = dict(cfg)
        t = cfg.pop('type')
        cfg.pop('requires_grad', None)
        layer = {'BN': nn.BatchNorm2d, 'BN1d': nn.BatchNorm1d, 'BN3d': nn.BatchNorm3d, 'LN': nn.LayerNorm}[t]
        return t.lower() + str(postfix), layer(num_features, **cfg)
    cnn.build_norm_layer = build_norm_layer
    _mod('mmcv.cnn.bricks')
    tr = _mod('mmcv.cnn.bricks.transformer')
    tr.MultiheadAttention, tr.FFN = MultiheadAttention, FFN
    _mod('mmcv.utils').ext_loader = _Dummy()


# ------------------------------------------------------------------------------------------------------- mmdet
def multi_apply(func, *args, **kwargs):
    pfunc = functools.partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


def _install_mmdet():
    u = _mod('mmdet.utils')
    for n in ('ConfigType', 'InstanceList', 'OptMultiConfig', 'OptConfigType', 'OptInstanceList', 'MultiConfig'):
        setattr(u, n, object)
    u.reduce_mean = lambda t: t
    _mod('mmdet.models')
    _mod('mmdet.models.utils').multi_apply = multi_apply
    _mod('mmdet.models.utils.misc')
    tm = _mod('mmdet.models.task_modules')

    class AssignResult:
        def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
            self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels
    tm.AssignResult = AssignResult
    tm.BaseAssigner = type('BaseAssigner', (), {})
    _mod('mmdet.models.task_modules.samplers')
    _mod('mmdet.structures')
    _mod('mmdet.structures.bbox')
    _mod('mmdet.evaluation')
    _mod('mmdet.datasets')
    _mod('mmdet.datasets.transforms')



# ---- mmdet model zoo pieces the configs name (third-party: restated from the public mmdet 3.x sources) ----
class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)      # style='pytorch'
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        o = torch.relu(self.bn1(self.conv1(x)))
        o = torch.relu(self.bn2(self.conv2(o)))
        return torch.relu(self.bn3(self.conv3(o)) + idt)


class _Basic(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        o = torch.relu(self.bn1(self.conv1(x)))
        return torch.relu(self.bn2(self.conv2(o)) + idt)


class ResNet(nn.Module):
    """mmdet.ResNet (style='pytorch', no deep stem, no dcn/plugins): torchvision topology with `base_channels`."""
    arch = {18: (_Basic, (2, 2, 2, 2)), 34: (_Basic, (3, 4, 6, 3)), 50: (_Bottleneck, (3, 4, 6, 3))}

    def __init__(self, depth, in_channels=3, base_channels=64, num_stages=4, out_indices=(0, 1, 2, 3),
                 frozen_stages=-1, norm_cfg=None, norm_eval=True, style='pytorch', init_cfg=None, **kw):
        super().__init__()
        block, nblocks = self.arch[depth]
        self.out_indices, self.frozen_stages, self.norm_eval, self.num_stages = out_indices, frozen_stages, norm_eval, \
            num_stages
        self.conv1 = nn.Conv2d(in_channels, base_channels, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(base_channels)
        inplanes = base_channels
        for i in range(num_stages):
            planes, stride = base_channels * 2 ** i, (1, 2, 2, 2)[i]
            layers = []
            for j in range(nblocks[i]):
                s = stride if j == 0 else 1
                ds = None
                if j == 0 and (s != 1 or inplanes != planes * block.expansion):
                    ds = nn.Sequential(nn.Conv2d(inplanes, planes * block.expansion, 1, s, bias=False),
                                       nn.BatchNorm2d(planes * block.expansion))
                layers.append(block(inplanes, planes, s, ds))
                inplanes = planes * block.expansion
            setattr(self, f'layer{i + 1}', nn.Sequential(*layers))

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self

    def forward(self, x):
        x = torch.relu(self.bn1(self.conv1(x)))
        x = torch.nn.functional.max_pool2d(x, 3, 2, 1)
        outs = []
        for i in range(self.num_stages):
            x = getattr(self, f'layer{i + 1}')(x)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)


class _ConvModule(nn.Module):
    """mmcv ConvModule without norm/activation: holds `.conv` (that is the checkpoint name)."""

    def __init__(self, cin, cout, k, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding)

    def forward(self, x):
        return self.conv(x)


class FPN(nn.Module):
    """mmdet.FPN, default options (add_extra_convs=False, nearest upsampling, extra levels by stride-2 max pool)."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, **kw):
        super().__init__()
        self.num_outs = num_outs
        self.lateral_convs = nn.ModuleList(_ConvModule(c, out_channels, 1) for c in in_channels)
        self.fpn_convs = nn.ModuleList(_ConvModule(out_channels, out_channels, 3, 1) for _ in in_channels)

    def forward(self, inputs):
        lat = [l(x) for l, x in zip(self.lateral_convs, inputs)]
        for i in range(len(lat) - 1, 0, -1):
            lat[i - 1] = lat[i - 1] + torch.nn.functional.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
        outs = [c(x) for c, x in zip(self.fpn_convs, lat)]
        for _ in range(self.num_outs - len(outs)):
            outs.append(torch.nn.functional.max_pool2d(outs[-1], 1, stride=2))
        return tuple(outs)


def _weight_reduce(loss, weight, reduction, avg_factor):
    """mmdet.models.losses.utils.weight_reduce_loss (3.x: mean with avg_factor divides by avg_factor + fp32 eps)."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return {'none': loss, 'mean': loss.mean(), 'sum': loss.sum()}[reduction]
    if reduction == 'mean':
        return loss.sum() / (avg_factor + torch.finfo(torch.float32).eps)
    assert reduction == 'none'
    return loss


class FocalLoss(nn.Module):
    """mmdet.FocalLoss(use_sigmoid=True): integer labels, out-of-range label (−1 / num_classes) = background, the
    behaviour of the mmcv CUDA op the reference runs (†upstream); value formula = mmdet `py_sigmoid_focal_loss`."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean', loss_weight=1.0, activated=False):
        super().__init__()
        self.gamma, self.alpha, self.reduction, self.loss_weight = gamma, alpha, reduction, loss_weight
        self.use_sigmoid, self.activated = use_sigmoid, activated

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        if pred.dim() != target.dim():                 # integer labels; same rank = already a (soft) one-hot target
            target = (target[:, None] == torch.arange(pred.shape[1])[None]).to(pred.dtype)
        else:
            target = target.to(pred.dtype)
        p = pred.sigmoid()
        pt = (1 - p) * target + p * (1 - target)
        fw = (self.alpha * target + (1 - self.alpha) * (1 - target)) * pt.pow(self.gamma)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, target, reduction='none') * fw
        if weight is not None and weight.dim() == 1 and loss.dim() == 2:
            weight = weight.view(-1, 1)
        return self.loss_weight * _weight_reduce(loss, weight, reduction_override or self.reduction, avg_factor)


class CrossEntropyLoss(nn.Module):
    """mmdet.CrossEntropyLoss(use_sigmoid=True) -> binary_cross_entropy on same-shaped (pred, label)."""

    def __init__(self, use_sigmoid=False, reduction='mean', loss_weight=1.0, **kw):
        super().__init__()
        assert use_sigmoid
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, **kw):
        assert cls_score.dim() == label.dim()
        valid = ((label >= 0) & (label != -100)).float()
        weight = valid if weight is None else weight * valid
        loss = torch.nn.functional.binary_cross_entropy_with_logits(cls_score, label.float(), reduction='none')
        return self.loss_weight * _weight_reduce(loss, weight, reduction_override or self.reduction, avg_factor)


class DetDataPreprocessor(nn.Module):
    """mmdet.DetDataPreprocessor / mmengine.ImgDataPreprocessor: only the state the reference subclass reads
    (`mean`/`std` buffers of shape (3,1,1), `_channel_conversion`, `_enable_normalize`, pad settings, `cast_data`)."""

    def __init__(self, mean=None, std=None, pad_size_divisor=1, pad_value=0, pad_mask=False, mask_pad_value=0,
                 pad_seg=False, seg_pad_value=255, bgr_to_rgb=False, rgb_to_bgr=False, boxtype2tensor=True,
                 non_blocking=False, batch_augments=None):
        super().__init__()
        assert not (bgr_to_rgb and rgb_to_bgr)
        self._channel_conversion = rgb_to_bgr or bgr_to_rgb
        self._enable_normalize = mean is not None
        if mean is not None:
            self.register_buffer('mean', torch.tensor(mean, dtype=torch.float32).view(-1, 1, 1), False)
            self.register_buffer('std', torch.tensor(std, dtype=torch.float32).view(-1, 1, 1), False)
        self.pad_size_divisor, self.pad_value = pad_size_divisor, pad_value
        self.pad_mask, self.pad_seg, self.boxtype2tensor, self.batch_augments = pad_mask, pad_seg, False, batch_augments

    def cast_data(self, data):
        return data


def _install_mmdet_models():
    sys.modules['mmdet.models'].DetDataPreprocessor = DetDataPreprocessor
    reg = sys.modules['mmengine'].MODELS
    for cls in (ResNet, FPN, FocalLoss, CrossEntropyLoss):
        reg.register_module(module=cls)

# --------------------------------------------------------------------------------------------------- pytorch3d
def _install_pytorch3d():
    from oracle import geometry_ref as G
    t = _mod('pytorch3d.transforms')
    t.euler_angles_to_matrix = lambda e, convention: G.euler_to_matrix(e, convention)

    def m2e(m, convention):
        assert convention == 'ZXY'
        return G.matrix_to_euler_zxy(m)
    t.matrix_to_euler_angles = m2e
    o = _mod('pytorch3d.ops')

    def box3d_overlap(c1, c2, eps=1e-4):
        vol, iou = G.box3d_overlap(c1.detach().cpu().numpy().astype(np.float64),
                                   c2.detach().cpu().numpy().astype(np.float64))
        return torch.as_tensor(vol, dtype=c1.dtype), torch.as_tensor(iou, dtype=c1.dtype)
    o.box3d_overlap = box3d_overlap


def install(reference_root='/root/reference'):
    """Put the stand-ins into `sys.modules` and the reference on `sys.path`.  Idempotent."""
    if getattr(install, '_done', False):
        return
    _install_mmengine()
    _install_mmcv()
    _install_mmdet()
    _install_mmdet_models()
    _install_pytorch3d()
    for name in ('open3d', 'cv2', 'terminaltables', 'numba', 'mmdet3d'):
        _mod(name)
    import me_cpu
    me_cpu.install(sys.modules)
    if not hasattr(torch.cuda, 'LongTensor'):
        torch.cuda.LongTensor = torch.LongTensor
        torch.cuda.BoolTensor = torch.BoolTensor
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    install._done = True
