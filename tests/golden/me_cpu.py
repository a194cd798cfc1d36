"""CPU stand-in for the slice of the MinkowskiEngine Python API the reference touches (SURVEY.md §8b "Sparse-tensor
operator surface"), built on `oracle/sparse_ref.py`, so that the reference's OWN model code (MinkResNet, FCAF3DHeadRotMat,
SparseFeatureFusionSingleStage3DDetector, MinkNeck ...) can run in the dev container and produce golden vectors.

TEST INFRASTRUCTURE ONLY (see refstubs.py).  The MinkowskiEngine semantics themselves remain "parity unpinned"
(†upstream, frozen by the oracle: first-occurrence dedup, x-fastest offsets, child row 8*parent+k, union order); what
the golden vectors pin is everything the reference implements itself on top of this surface.
"""
import math
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import sparse_ref as S


class CoordinateManager:
    def __init__(self, n_batch):
        self.n_batch = n_batch
        self.cache = {}

    def strided(self, coords, ts, factor):
        key = ('coords', id(coords), factor)
        if key not in self.cache:
            self.cache[key] = (S.unique_first(coords, ts * factor)[0], coords)
        return self.cache[key][0]

    def kernel_map(self, cin, cout, ksize, ts):
        key = ('map', id(cin), id(cout), ksize)
        if key not in self.cache:
            self.cache[key] = (S.kernel_map(cin, cout, S.offsets(ksize, ts)), cin, cout)
        return self.cache[key][0]


class SparseTensor:
    def __init__(self, features=None, coordinates=None, coordinate_map_key=None, coordinate_manager=None, **kw):
        if coordinates is not None:
            c = coordinates.detach().cpu().numpy().astype(np.int64)
            uc, in2out = S.unique_first(c)
            first = np.full(uc.shape[0], c.shape[0], dtype=np.int64)
            np.minimum.at(first, in2out, np.arange(c.shape[0]))
            self._coords, self._stride = uc, 1
            self.F = features[torch.from_numpy(first)]
            self.coordinate_manager = CoordinateManager(int(c[:, 0].max()) + 1 if c.shape[0] else 1)
        else:
            self._coords, self._stride = coordinate_map_key
            self.F = features
            self.coordinate_manager = coordinate_manager
            assert features.shape[0] == self._coords.shape[0]

    @classmethod
    def wrap(cls, feats, coords, stride, cm):
        return cls(features=feats, coordinate_map_key=(coords, stride), coordinate_manager=cm)

    @property
    def coordinate_map_key(self):
        return (self._coords, self._stride)

    @property
    def features(self):
        return self.F

    @property
    def C(self):
        return torch.from_numpy(self._coords).to(torch.int32)

    coordinates = C

    @property
    def tensor_stride(self):
        return [self._stride] * 3

    @property
    def device(self):
        return self.F.device

    @property
    def decomposition_permutations(self):
        return [torch.from_numpy(np.nonzero(self._coords[:, 0] == b)[0]) for b in range(self.coordinate_manager.n_batch)]

    @property
    def decomposed_coordinates(self):
        C = self.C
        return [C[p][:, 1:] for p in self.decomposition_permutations]

    @property
    def decomposed_features(self):
        return [self.F[p] for p in self.decomposition_permutations]

    def __len__(self):
        return self._coords.shape[0]

    def __add__(self, other):
        if other._coords is self._coords:
            return SparseTensor.wrap(self.F + other.F, self._coords, self._stride, self.coordinate_manager)
        assert self._stride == other._stride
        uc, map_b = S.union(self._coords, other._coords)
        return SparseTensor.wrap(S.union_add(self.F, other.F, map_b, uc.shape[0]), uc, self._stride,
                                 self.coordinate_manager)

    def features_at_coordinates(self, query):
        q = query.detach().cpu().numpy()
        assert (q == np.floor(q)).all()
        return S.features_at_coordinates(self._coords, self.F, self._stride, q.astype(np.int64))

    def dense(self, shape=None, min_coordinate=None, contract_stride=True):
        """(B, C, X, Y, Z) dense tensor; coordinates are shifted by `min_coordinate` and divided by the tensor stride."""
        mn = min_coordinate.view(-1).numpy().astype(np.int64) if min_coordinate is not None else np.zeros(3, np.int64)
        idx = (self._coords[:, 1:] - mn[None]) // self._stride
        B, C, X, Y, Z = tuple(shape)
        ok = ((idx >= 0) & (idx < np.asarray([X, Y, Z])[None])).all(1)
        rows = torch.from_numpy(np.nonzero(ok)[0])
        b = torch.from_numpy(self._coords[ok, 0])
        i = torch.from_numpy(idx[ok])
        out = self.F.new_zeros((B, X, Y, Z, C)).index_put((b, i[:, 0], i[:, 1], i[:, 2]), self.F[rows])
        return out.permute(0, 4, 1, 2, 3), torch.from_numpy(mn), torch.tensor(self.tensor_stride)


def cat(a, b):
    assert a._coords is b._coords
    return SparseTensor.wrap(torch.cat([a.F, b.F], 1), a._coords, a._stride, a.coordinate_manager)


def batched_coordinates(coords_list):
    out = []
    for b, c in enumerate(coords_list):
        c = torch.floor(c.detach().float()).to(torch.int32) if c.is_floating_point() else c.to(torch.int32)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=torch.int32), c], 1))
    return torch.cat(out)


def batch_sparse_collate(data, dtype=torch.int32, device=None):
    coords, feats = zip(*data)
    return batched_coordinates(list(coords)), torch.cat(list(feats))


def kaiming_normal_(tensor, a=0, mode='fan_in', nonlinearity='leaky_relu'):
    """ME.utils.kaiming_normal_ (†upstream): fan of a (K, Cin, Cout) kernel = K*Cin (fan_in) / K*Cout (fan_out)."""
    K = tensor.shape[0] if tensor.dim() == 3 else 1
    fan = K * (tensor.shape[-2] if mode == 'fan_in' else tensor.shape[-1])
    std = nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)


class MinkowskiConvolution(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, dimension=3):
        super().__init__()
        self.kernel_size, self.stride = kernel_size, stride
        K = kernel_size ** 3
        self.kernel = nn.Parameter(torch.empty((K, in_channels, out_channels) if K > 1 else (in_channels, out_channels)))
        self.bias = nn.Parameter(torch.zeros(1, out_channels)) if bias else None
        kaiming_normal_(self.kernel, mode='fan_out', nonlinearity='relu')

    def forward(self, x):
        cm = x.coordinate_manager
        oc = cm.strided(x._coords, x._stride, self.stride) if self.stride > 1 else x._coords
        nbr = cm.kernel_map(x._coords, oc, self.kernel_size, x._stride)
        y = S.conv(x.F, self.kernel, nbr)
        if self.bias is not None:
            y = y + self.bias
        return SparseTensor.wrap(y, oc, x._stride * self.stride, cm)


class MinkowskiGenerativeConvolutionTranspose(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=2, stride=2, dilation=1, bias=False, dimension=3):
        super().__init__()
        assert kernel_size == 2 and stride == 2 and not bias
        self.kernel = nn.Parameter(torch.empty(8, in_channels, out_channels))
        kaiming_normal_(self.kernel, mode='fan_out', nonlinearity='relu')

    def forward(self, x):
        cm = x.coordinate_manager
        key = ('gen', id(x._coords))
        if key not in cm.cache:
            cm.cache[key] = (S.generative_children(x._coords, x._stride // 2), x._coords)
        return SparseTensor.wrap(S.generative_conv(x.F, self.kernel), cm.cache[key][0], x._stride // 2, cm)


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum)

    def forward(self, x):
        return SparseTensor.wrap(self.bn(x.F), x._coords, x._stride, x.coordinate_manager)


class MinkowskiInstanceNorm(nn.Module):
    def __init__(self, num_features):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(1, num_features))
        self.bias = nn.Parameter(torch.zeros(1, num_features))

    def forward(self, x):
        y = S.instance_norm(x.F, x._coords[:, 0], x.coordinate_manager.n_batch, self.weight, self.bias)
        return SparseTensor.wrap(y, x._coords, x._stride, x.coordinate_manager)


class _Act(nn.Module):
    fn = None

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x):
        return SparseTensor.wrap(type(self).fn(x.F), x._coords, x._stride, x.coordinate_manager)


class MinkowskiReLU(_Act):
    fn = staticmethod(F.relu)


class MinkowskiELU(_Act):
    fn = staticmethod(F.elu)


class MinkowskiMaxPooling(nn.Module):
    def __init__(self, kernel_size=2, stride=2, dilation=1, dimension=3):
        super().__init__()
        assert kernel_size == 2 and stride == 2

    def forward(self, x):
        cm = x.coordinate_manager
        oc = cm.strided(x._coords, x._stride, 2)
        nbr = cm.kernel_map(x._coords, oc, 2, x._stride)
        return SparseTensor.wrap(S.maxpool(x.F, nbr), oc, x._stride * 2, cm)


class MinkowskiPruning(nn.Module):
    def forward(self, x, mask):
        m = mask.detach().cpu().numpy().astype(bool)
        if m.all():
            return x                      # keeps the coordinate identity, like the oracle's no-prune fast path
        rows = torch.from_numpy(np.nonzero(m)[0])
        return SparseTensor.wrap(x.F[rows], x._coords[m], x._stride, x.coordinate_manager)


class BasicBlock(nn.Module):
    """ME.modules.resnet_block.BasicBlock (†upstream): conv3-BN-ReLU-conv3-BN-(+downsample)-ReLU."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=stride, dimension=dimension)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dimension=dimension)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        residual = x
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        return self.relu(out + residual)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=1, dimension=dimension)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=stride, dimension=dimension)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv3 = MinkowskiConvolution(planes, planes * 4, kernel_size=1, dimension=dimension)
        self.norm3 = MinkowskiBatchNorm(planes * 4, momentum=bn_momentum)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        residual = x
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.relu(self.norm2(self.conv2(out)))
        out = self.norm3(self.conv3(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        return self.relu(out + residual)


def install(sys_modules):
    """Register this file as `MinkowskiEngine` (+ `.utils`, `.modules.resnet_block`)."""
    me = types.ModuleType('MinkowskiEngine')
    for k, v in globals().items():
        if k.startswith('Minkowski') or k in ('SparseTensor', 'cat', 'CoordinateManager'):
            setattr(me, k, v)
    utils = types.ModuleType('MinkowskiEngine.utils')
    utils.batch_sparse_collate, utils.batched_coordinates = batch_sparse_collate, batched_coordinates
    utils.kaiming_normal_ = kaiming_normal_
    me.utils = utils
    modules = types.ModuleType('MinkowskiEngine.modules')
    rb = types.ModuleType('MinkowskiEngine.modules.resnet_block')
    rb.BasicBlock, rb.Bottleneck = BasicBlock, Bottleneck
    modules.resnet_block = rb
    me.modules = modules
    sys_modules['MinkowskiEngine'] = me
    sys_modules['MinkowskiEngine.utils'] = utils
    sys_modules['MinkowskiEngine.modules'] = modules
    sys_modules['MinkowskiEngine.modules.resnet_block'] = rb
