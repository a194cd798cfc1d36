"""Host-side logic of the product exercised without a GPU: the C-ABI call is replaced by a numpy emulation of the ONE
kernel involved, so that batching / padding / bookkeeping code written while no B200 was available cannot hide a Python
error behind the GPU tests. The emulation lives only here; the product itself never falls back to the CPU."""
import ctypes
import os
import sys

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
if GOLD not in sys.path:
    sys.path.insert(0, GOLD)

MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def _emulated_img_normalize(calls):
    def fake_call(name, src_ptr, n_img, H, W, Hp, Wp, mean_p, std_p, conv, channels_last, dst_ptr, dcode, stream):
        assert name == 'esb_img_normalize' and channels_last == 1 and dcode == 0
        calls.append((n_img, H, W, Hp, Wp))
        src = np.ctypeslib.as_array(ctypes.cast(src_ptr, ctypes.POINTER(ctypes.c_ubyte)), shape=(n_img, 3, H, W))
        dst = np.ctypeslib.as_array(ctypes.cast(dst_ptr, ctypes.POINTER(ctypes.c_float)), shape=(n_img, Hp, Wp, 3))
        mean = np.ctypeslib.as_array(ctypes.cast(mean_p, ctypes.POINTER(ctypes.c_float)), shape=(3, ))
        std = np.ctypeslib.as_array(ctypes.cast(std_p, ctypes.POINTER(ctypes.c_float)), shape=(3, ))
        dst[:] = 0
        s = src[:, ::-1] if conv else src
        dst[:, :H, :W, :] = ((s.astype(np.float32) - mean[None, :, None, None]) / std[None, :, None, None]) \
            .transpose(0, 2, 3, 1)
    return fake_call


def test_preprocessor_batching_and_padding(monkeypatch):
    import embodiedscan_b200.detectors as D
    from cases import preprocess_inputs
    from embodiedscan_b200.structures import Det3DDataSample
    from oracle import data_ref as R
    calls = []
    monkeypatch.setattr(D, 'call', _emulated_img_normalize(calls))
    monkeypatch.setattr(D, 'stream', lambda: None)
    gold = np.load(os.path.join(GOLD, 'frontend.npz'))
    pre = D.Det3DDataPreprocessor(mean=MEAN, std=STD, bgr_to_rgb=True, pad_size_divisor=32)
    # mixed view sizes: one launch per scan into its slice of the batch buffer, padded to the batch maximum
    samples = [Det3DDataSample(metainfo={}), Det3DDataSample(metainfo={})]
    out = pre(dict(inputs=dict(img=preprocess_inputs()), data_samples=samples))
    assert torch.equal(out['inputs']['imgs'], torch.from_numpy(gold['pre_imgs']))
    assert calls == [(2, 30, 45, 64, 64), (2, 33, 40, 64, 64)]
    assert [s.metainfo['pad_shape'] for s in samples] == [tuple(r) for r in gold['pre_pad_shape'].tolist()]
    # the usual uniform batch stays ONE launch, whatever container the dataloader hands over
    g = torch.Generator().manual_seed(3)
    five = torch.randint(0, 256, (2, 3, 3, 40, 50), generator=g, dtype=torch.uint8)
    for inp, ref in ((five, list(five)), (list(five), list(five)), (five[:, 0], [x[None] for x in five[:, 0]])):
        calls.clear()
        out = pre(dict(inputs=dict(img=inp)))['inputs']['imgs']
        assert len(calls) == 1 and calls[0][1:] == (40, 50, 64, 64)
        assert torch.equal(out, R.preprocess_multiview(ref, MEAN, STD))
        assert out.shape[2] == 3 and out.stride(2) == 1, 'channels-last memory under a (B,V,3,H,W) view'
