"""One C2 predict call (top-k, decode, BEV rotated NMS), a depth unprojection and a 9-DoF IoU call: the hand-written kernels
that are not part of the training step, for profiles/capture_r2.sh."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodiedscan_b200 import MODELS  # noqa: E402
from embodiedscan_b200.synth import mv_det3d_config, synth_scan  # noqa: E402

dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = MODELS.build(dict(mv_det3d_config('C2'), compute_dtype=torch.bfloat16)).to(dev).eval()
scans = [synth_scan(i, augment=False, device=dev, n_views=20, H=480, W=640, n_points=100000) for i in range(2)]
data = dict(inputs=dict(points=[s['points'] for s in scans], img=[s['img'] for s in scans]), data_samples=[s['data_sample'] for s in scans])
with torch.no_grad():
    for _ in range(2):
        out = model.val_step(data)
torch.cuda.synchronize()
print('predicted', [len(o.pred_instances_3d.scores_3d) for o in out])
try:
    from embodiedscan_b200.transforms import unproject_multiview
    s = scans[0]
    d2i = s['data_sample'].metainfo['depth2img']
    pts = unproject_multiview(s['depth'].to(dev), [d2i['intrinsic']] * len(d2i['extrinsic']) if not isinstance(d2i['intrinsic'], (list, tuple)) else d2i['intrinsic'], d2i['extrinsic'])
    print('unprojected', tuple(pts.shape))
except Exception as e:  # noqa
    print('unproject skipped:', repr(e)[:200])
try:
    from embodiedscan_b200.geometry import box3d_overlap
    c = out[0].pred_instances_3d.bboxes_3d.corners[:200].float()
    if len(c):
        print('iou', box3d_overlap(c, c)[1].shape)
except Exception as e:  # noqa
    print('iou skipped:', repr(e)[:200])
torch.cuda.synchronize()
