"""Loading the reference's released checkpoints (`mv-3ddet.pth`, `mv-occ.pth`, `mv-grounding.pth`,
README.md:206 of the reference) into the esb200 modules.

mmengine saves ``{'meta': ..., 'state_dict': {...}, 'optimizer': ...}``; keys may carry a ``module.`` prefix from
DistributedDataParallel. The esb200 modules keep the reference's parameter names and shapes (ME ``kernel`` (K,Cin,Cout) /
(Cin,Cout), ``bias`` (1,Cout), ``.bn.*``; mmdet ``conv1 / bn1 / layerX.Y.convZ / downsample.N`` are mapped on load by
``backbones.ResNet._load_from_state_dict``), so no tensor is reshaped or transposed here. One assumption rides on
that (unpinned without MinkowskiEngine): slice k of a trained ``kernel`` belongs to the k-th offset of
ME's hypercube region iterator, taken to be x-fastest — the enumeration the kernels and the oracle use. If a real
checkpoint shows otherwise, the fix is one permutation of dim 0 of every 27- and 8-slice kernel here.
"""
from typing import Dict, Tuple

import torch


def reference_state_dict(ckpt) -> Dict[str, torch.Tensor]:
    """The flat name -> tensor dict of an mmengine / torch checkpoint object (or of an already flat dict)."""
    sd = ckpt.get('state_dict', ckpt) if isinstance(ckpt, dict) else ckpt
    out = {}
    for k, v in sd.items():
        if not torch.is_tensor(v):
            continue
        while k.startswith('module.'):
            k = k[len('module.'):]
        out[k] = v
    return out


def load_reference_checkpoint(model: torch.nn.Module, path_or_ckpt, strict: bool = True) -> Tuple[list, list]:
    """Load a released reference checkpoint (path or loaded object). Returns (missing, unexpected) after ignoring the
    bookkeeping buffers a checkpoint may lack (``num_batches_tracked``) and the text encoder's position ids; raises when
    `strict` and anything else does not line up."""
    ckpt = torch.load(path_or_ckpt, map_location='cpu', weights_only=False) if isinstance(path_or_ckpt, str) \
        else path_or_ckpt
    sd = reference_state_dict(ckpt)
    own = model.state_dict()
    for k, v in sd.items():
        if k in own and own[k].shape != v.shape:
            raise RuntimeError(f'checkpoint tensor {k} has shape {tuple(v.shape)}, the model expects {tuple(own[k].shape)}')
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [k for k in missing if not k.endswith('num_batches_tracked') and 'position_ids' not in k]
    unexpected = [k for k in unexpected if 'position_ids' not in k]
    if strict and (missing or unexpected):
        raise RuntimeError(f'checkpoint does not match the model: missing {missing[:8]} ... unexpected {unexpected[:8]} ...')
    return missing, unexpected
