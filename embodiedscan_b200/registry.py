"""The drop-in boundary: the reference resolves ``type=`` strings of its python-dict configs through
mmengine registries (``embodiedscan/registry.py:10-31``). When mmengine is importable our modules register into child
registries of the same roots under the same names/scope; when it is absent (this image, and the GPU box) a minimal
look-alike with the same ``register_module`` / ``build`` contract is used so the same config dicts build the model.
"""
import inspect

try:  # pragma: no cover - mmengine is not installed in this image
    from mmengine.registry import MODELS as _MM_MODELS
    from mmengine.registry import TASK_UTILS as _MM_TASK_UTILS
    from mmengine.registry import Registry as _MMRegistry
    HAVE_MMENGINE = True
except Exception:  # noqa
    HAVE_MMENGINE = False


class Registry:
    """Minimal mmengine.Registry look-alike: name -> class, ``build(cfg)`` pops ``type`` and calls the class."""

    def __init__(self, name, scope='embodiedscan'):
        self.name, self.scope = name, scope
        self._module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            for k in ([key] if isinstance(key, str) else key):
                if k in self._module_dict and not force:
                    raise KeyError(f'{k} is already registered in {self.name}')
                self._module_dict[k] = cls
            return cls

        if module is not None:
            return _register(module)
        return _register

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        # scope-qualified names ("mmdet.ResNet") resolve to what we registered under that full name
        if '.' in key and key.split('.', 1)[1] in self._module_dict and key.startswith(self.scope + '.'):
            return self._module_dict[key.split('.', 1)[1]]
        return None

    def build(self, cfg, *args, **kwargs):
        if cfg is None:
            return None
        if not isinstance(cfg, dict):
            return cfg  # already built
        cfg = dict(cfg)
        typ = cfg.pop('type')
        cls = typ if inspect.isclass(typ) or callable(typ) else self.get(typ)
        if cls is None:
            raise KeyError(f'{typ} is not in the {self.name} registry')
        cfg.update(kwargs)
        return cls(*args, **cfg)

    def __contains__(self, key):
        return self.get(key) is not None


if HAVE_MMENGINE:  # pragma: no cover
    MODELS = _MMRegistry('model', parent=_MM_MODELS, scope='embodiedscan_b200')
    TASK_UTILS = _MMRegistry('task util', parent=_MM_TASK_UTILS, scope='embodiedscan_b200')
else:
    MODELS = Registry('model')
    TASK_UTILS = Registry('task util')
TRANSFORMS = Registry('transform') if not HAVE_MMENGINE else _MMRegistry('transform', scope='embodiedscan_b200')


def register_into_reference():  # pragma: no cover - needs the reference package importable
    """Register our implementations under the reference's own registry (embodiedscan.registry.MODELS) so that
    ``tools/train.py`` / ``tools/test.py`` build them from unmodified configs. See INTEGRATION.md."""
    from embodiedscan.registry import MODELS as REF_MODELS
    from embodiedscan.registry import TASK_UTILS as REF_TASK_UTILS
    for name, cls in MODELS._module_dict.items():
        REF_MODELS.register_module(name=name, force=True, module=cls)
    for name, cls in TASK_UTILS._module_dict.items():
        REF_TASK_UTILS.register_module(name=name, force=True, module=cls)
