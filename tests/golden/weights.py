"""Deterministic, name-keyed parameter fill shared by `make_golden.py` (reference side, run in the dev container) and the
golden parity tests (oracle / CUDA side, run anywhere): the same (name, shape) always yields the same CPU tensor, so the
golden files only need to store the parameter manifest and the reference's outputs, not the weights themselves.
Test infrastructure; imports nothing from `/root/reference` or `oracle/`."""
import math
import zlib

import torch


def fill_tensor(name: str, shape) -> torch.Tensor:
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7fffffff)
    shape = tuple(shape)
    last = name.rsplit('.', 1)[-1]
    if last == 'num_batches_tracked':
        return torch.zeros(shape, dtype=torch.long)
    if last == 'running_mean':
        return 0.1 * torch.randn(shape, generator=g)
    if last == 'running_var':
        return 0.5 + torch.rand(shape, generator=g)
    if len(shape) >= 2 and not (len(shape) == 2 and shape[0] == 1):      # (1, C) = ME bias / instance-norm affine
        if last == 'kernel':                                   # sparse kernel (K, Cin, Cout) or (Cin, Cout)
            fan_in = math.prod(shape[:-1])
        else:                                                  # dense conv / linear weight (Cout, Cin, ...)
            fan_in = math.prod(shape[1:])
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
    if last in ('bias', ) or 'bias' in last:
        return 0.1 * torch.randn(shape, generator=g)
    return 1.0 + 0.1 * torch.randn(shape, generator=g)         # norm weights, learnable scales


def fill_state_dict(manifest) -> dict:
    """manifest: iterable of (name, shape)."""
    return {k: fill_tensor(k, s) for k, s in manifest}


def adjust_fcaf3d_head(sd: dict, prefix: str = 'bbox_head.') -> dict:
    """Bring the three prediction layers to the scale the reference initialises them at (std 0.01 kernels, negative
    class prior; fcaf3d_head.py:986-991) so losses and scores sit in their usual range. In place; returns `sd`."""
    for k in ('conv_center.kernel', 'conv_reg.kernel', 'conv_cls.kernel'):
        sd[prefix + k] = sd[prefix + k] * 0.08
    sd[prefix + 'conv_cls.bias'] = sd[prefix + 'conv_cls.bias'] - 2.0
    return sd


def adjust_for_predict(sd: dict, prefix: str = 'bbox_head.') -> dict:
    """Three classes clear the score threshold (the oracle NMS is a scalar python loop). In place; returns `sd`."""
    bias = torch.full((1, sd[prefix + 'conv_cls.bias'].shape[-1]), -9.0)
    bias[0, [3, 77, 200]] = -1.5
    sd[prefix + 'conv_cls.bias'] = bias
    sd[prefix + 'conv_cls.kernel'] = sd[prefix + 'conv_cls.kernel'] * 5.0
    sd[prefix + 'conv_center.kernel'] = sd[prefix + 'conv_center.kernel'] * 20.0
    return sd


def adjust_grounder(sd: dict) -> dict:
    """`share_pred_layer=True`: every `bbox_head.{cls,reg}_branches.i` IS entry 0 (grounding_head.py:205-214), so the
    name-keyed fill must give the aliases one value. In place; returns `sd`."""
    for k in list(sd):
        for kind in ('cls_branches', 'reg_branches'):
            tag = f'bbox_head.{kind}.'
            if k.startswith(tag) and not k.startswith(tag + '0.'):
                i, rest = k[len(tag):].split('.', 1)
                sd[k] = sd[tag + '0.' + rest]
    return sd
