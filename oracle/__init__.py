"""oracle/ — CPU restatement of the reference's algorithm for the EmbodiedScan multi-view 3D perception hot path.

TEST INFRASTRUCTURE ONLY. Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package, and only as the checker / the timed CPU baseline — never as part of the product path
(``embodiedscan_b200`` never imports it and fails loudly when ``libesb200.so`` is missing).

PINNED against the reference's own Python. The reference ships no tests, no golden vectors and no fixtures
(SURVEY.md §4, §8c), and none of its kernel dependencies (MinkowskiEngine, mmcv 2.0.0rc4, pytorch3d, mmdet, mmengine)
can be installed here. Its OWN code is plain Python, though: ``tests/golden/make_golden.py`` imports
``/root/reference/embodiedscan/**`` in place behind stand-ins for the missing packages, executes it on seeded inputs and
commits the outputs (``tests/golden/*.npz``); ``tests/test_golden_cpu.py`` holds this oracle to them at fp32 round-off —
whole detector / occupancy / grounding / continuous models (losses, gradients, predictions), front-end, augmentations,
evaluation, and function-level edge cases. STILL UNPINNED: the arithmetic that lives inside the absent packages (ME
coordinate / kernel-map conventions, mmcv BEV IoU and focal loss, pytorch3d Euler conversions and ``box3d_overlap``,
mmdet ResNet / FPN / loss reductions, tie orders of ``torch.topk`` / ``np.argsort``): frozen here from the published
algorithms and additionally pinned by closed forms (tests/test_oracle_cpu.py: dense ``F.conv3d`` equivalence, neighbour
counting, pinhole projection, axis-aligned / 45-degree box IoU, Euler round trips, focal loss at p = 0.5, chamfer of
translated boxes, hand-countable ``get_targets``) and the ``weighted_loss`` docstring vector
(embodiedscan/models/losses/reduce_loss.py:80-96).

Every function cites the reference file:line (or the †upstream operator) it restates. Arithmetic is numpy / torch CPU
fp32; where an integer selection depends on fp32 rounding (voxel indices, nearest-pixel indices, target assignment)
the operation ORDER is spelled out so the CUDA kernels can mirror it bit for bit.
"""
