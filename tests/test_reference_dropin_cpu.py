"""Drop-in check against the reference tree itself (dev container only: skipped where /root/reference is absent, e.g. on
the GPU box). With the third-party stand-ins of tests/golden/refstubs.py installed, `register_into_reference()` must make
the reference's OWN registry build the esb200 classes from the reference's UNMODIFIED config files."""
import os
import sys

import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'embodiedscan')), reason='needs /root/reference')
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_config(path):
    """The `model = dict(...)` of an mmengine python config (plain exec; `_base_` files only hold runtime settings)."""
    scope = {}
    exec(compile(open(path).read(), path, 'exec'), scope)
    return scope['model']


@pytest.mark.parametrize('cfg_file,cls_name', [
    ('configs/detection/mv-det3d_8xb4_embodiedscan-3d-284class-9dof.py', 'SparseFeatureFusionSingleStage3DDetector'),
    ('configs/detection/cont-det3d_8xb1_embodiedscan-3d-284class-9dof.py', 'Embodied3DDetector'),
    ('configs/occupancy/mv-occ_8xb1_embodiedscan-occ-80class.py', 'DenseFusionOccPredictor'),
    ('configs/occupancy/cont-occ_8xb1_embodiedscan-occ-80class.py', 'EmbodiedOccPredictor'),
    ('configs/grounding/mv-grounding_8xb12_embodiedscan-vg-9dof.py', 'SparseFeatureFusion3DGrounder'),
    ('configs/grounding/mv-grounding_8xb12_embodiedscan-vg-9dof_fcaf-coder.py', 'SparseFeatureFusion3DGrounder'),
])
def test_reference_registry_builds_esb200_from_unmodified_configs(cfg_file, cls_name):
    if GOLD not in sys.path:
        sys.path.insert(0, GOLD)
    import refstubs
    refstubs.install(REF)
    import embodiedscan.models  # noqa: F401  (the reference registers its own classes first)
    from embodiedscan.registry import MODELS as REF_MODELS

    import embodiedscan_b200
    from embodiedscan_b200.registry import register_into_reference
    register_into_reference()
    model_cfg = load_config(os.path.join(REF, cfg_file))
    assert model_cfg['type'] == cls_name
    for sub in ('backbone', 'neck'):                      # pretrained-checkpoint URLs need a network
        if isinstance(model_cfg.get(sub), dict):
            model_cfg[sub].pop('init_cfg', None)
    model = REF_MODELS.build(model_cfg)
    assert type(model).__module__.startswith('embodiedscan_b200'), type(model)
    assert type(model) is getattr(embodiedscan_b200, cls_name, None) or type(model).__name__ == cls_name
    names = dict(model.named_parameters())
    assert any(k.endswith('.kernel') for k in names) and 'backbone_3d.conv1.kernel' in names
