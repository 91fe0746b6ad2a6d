"""mmcv-compatible registry shim (reference plugin mechanism: mmdet/models/builder.py:6-34).

When mmdet is importable the real BACKBONES / HEADS registries are used, so
``configs/mae/*.py`` build our classes by name.  Otherwise a minimal stand-in with the
same ``register_module`` / ``build`` surface keeps the drop-in testable here (this
container has no mmcv / mmdet)."""
import inspect


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def __contains__(self, key):
        return key in self._module_dict

    def _register(self, cls, name=None, force=False):
        if not inspect.isclass(cls):
            raise TypeError(f'module must be a class, but got {type(cls)}')
        names = [cls.__name__] if name is None else ([name] if isinstance(name, str) else list(name))
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f'{n} is already registered in {self._name}')
            self._module_dict[n] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def deco(cls):
            self._register(cls, name, force)
            return cls
        return deco

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    """cfg['type'] is the class-name string, the remaining keys are ctor kwargs (builder.py:15-34)."""
    if not isinstance(cfg, dict) or 'type' not in cfg:
        raise KeyError('cfg must be a dict containing the key "type"')
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop('type')
    if isinstance(t, str):
        cls = registry.get(t)
        if cls is None:
            raise KeyError(f'{t} is not in the {registry.name} registry')
    elif inspect.isclass(t):
        cls = t
    else:
        raise TypeError(f'type must be a str or class, got {type(t)}')
    return cls(**args)


try:  # pragma: no cover - mmdet is not installed in the build container
    from mmdet.models.builder import BACKBONES, HEADS
    USING_MMDET = True
except Exception:
    BACKBONES = Registry('backbone')
    HEADS = Registry('head')
    USING_MMDET = False


def _register_all():
    """importing the modules performs the registration (same side-effect mechanism as mmdet's __init__ files)."""
    from . import backbone  # noqa: F401
    try:
        from . import head  # noqa: F401
        from . import mil  # noqa: F401
    except ImportError:
        pass


def build_backbone(cfg):
    _register_all()
    return build_from_cfg(cfg, BACKBONES) if not USING_MMDET else BACKBONES.build(cfg)


def build_head(cfg):
    _register_all()
    return build_from_cfg(cfg, HEADS) if not USING_MMDET else HEADS.build(cfg)
