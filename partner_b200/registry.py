"""Registries mirroring det3d/utils/registry.py:6-78 and det3d/models/registry.py:3-4, so the
drop-in classes are selectable from det3d configs by their ``type=`` string."""
import inspect


class Registry(object):
    def __init__(self, name):
        self._name = name
        self._module_dict = dict()

    def __repr__(self):
        return "%s(name=%s, items=%s)" % (self.__class__.__name__, self._name, list(self._module_dict))

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key, None)

    def register_module(self, cls):
        if not inspect.isclass(cls):
            raise TypeError("module must be a class, but got %s" % type(cls))
        name = cls.__name__
        if name in self._module_dict:
            raise KeyError("%s is already registered in %s" % (name, self._name))
        self._module_dict[name] = cls
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    """det3d/utils/registry.py:49-78."""
    assert isinstance(cfg, dict) and "type" in cfg
    args = dict(cfg)
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError("%s is not in the %s registry" % (obj_type, registry.name))
    elif inspect.isclass(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError("type must be a str or class, got %s" % type(obj_type))
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    return obj_cls(**args)


READERS = Registry("reader")
BACKBONES = Registry("backbone")
