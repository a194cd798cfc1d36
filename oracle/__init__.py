"""oracle/ — CPU restatement of the reference's algorithm for the EmbodiedScan multi-view 3D perception hot path.

TEST INFRASTRUCTURE ONLY. Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package, and only as the checker / the timed CPU baseline — never as part of the product path
(``embodiedscan_b200`` never imports it and fails loudly when ``libesb200.so`` is missing).

PARITY UNPINNED. The reference ships no tests, no golden vectors and no fixtures, and none of its
kernel dependencies (MinkowskiEngine, mmcv 2.0.0rc4, pytorch3d, mmdet, mmengine) can be imported or built in this
environment, so neither reference outputs nor reference fixtures exist to pin this oracle against. What pins it
instead (tests/test_oracle_cpu.py): closed forms — dense ``F.conv3d`` equivalence on a fully occupied cube, neighbour
counting with all-ones weights, identity-extrinsic pinhole projection, axis-aligned / 45-degree box IoU, Euler
round trips and the explicit ZXY formulas, focal loss at p = 0.5, chamfer of translated boxes, hand-countable
``get_targets`` — plus the only known-answer vector the reference holds for this path, the ``weighted_loss``
docstring example (embodiedscan/models/losses/reduce_loss.py:80-96).

Every function cites the reference file:line (or the †upstream operator) it restates. Arithmetic is numpy / torch CPU
fp32; where an integer selection depends on fp32 rounding (voxel indices, nearest-pixel indices, target assignment)
the operation ORDER is spelled out so the CUDA kernels can mirror it bit for bit.
"""
