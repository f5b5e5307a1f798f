"""Python-dict config files with `_base_` inheritance (the mmcv.Config subset DeMF relies on).

The reference drives everything from configs like configs/demf/demf_votenet.py: a python
file whose top-level names become the config, a `_base_` list of files merged underneath
(configs/demf/demf_votenet.py:1-5), recursive dict merge with `_delete_=True`, attribute
access on nested dicts (`train_cfg.pts.sample_mod`, demfnet.py:166; `decoder.num_layers`,
class_agnostic_vote_head.py:386) and `--cfg-options a.b=c` overrides (train.py:24-26).
"""
import copy
import os
import types

BASE_KEY = "_base_"
DELETE_KEY = "_delete_"


class ConfigDict(dict):
    """dict with attribute access; nested dicts (also inside lists/tuples) are converted."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def setdefault(self, k, default=None):
        if k not in self:
            self[k] = default
        return self[k]

    def copy(self):
        return ConfigDict(super().copy())

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def to_dict(self):
        def plain(v):
            if isinstance(v, dict):
                return {k: plain(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(plain(x) for x in v)
            return v

        return plain(self)


def _merge_a_into_b(a, b):
    """Recursive merge of child `a` over base `b` (returns a new dict)."""
    b = dict(b)
    for k, v in a.items():
        if isinstance(v, dict) and k in b and not v.get(DELETE_KEY, False):
            if not isinstance(b[k], dict):
                raise TypeError(
                    f"{k}={v} in child config cannot inherit from base because {k} is a dict in "
                    f"the child config but is of type {type(b[k])} in base config. You may set "
                    f"`{DELETE_KEY}=True` to ignore the base config")
            b[k] = _merge_a_into_b(v, b[k])
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != DELETE_KEY}
            b[k] = v
    return b


def _file2dict(filename):
    filename = os.path.abspath(os.path.expanduser(filename))
    if not os.path.isfile(filename):
        raise FileNotFoundError(f"config file {filename} does not exist")
    if not filename.endswith(".py"):
        raise IOError("Only py type are supported now!")
    with open(filename, "r", encoding="utf-8") as f:
        text = f.read()
    scope = {"__file__": filename, "__name__": "_demf_config_"}
    exec(compile(text, filename, "exec"), scope)
    cfg = {k: v for k, v in scope.items()
           if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType))
           and not isinstance(v, type)}
    if BASE_KEY in cfg:
        bases = cfg.pop(BASE_KEY)
        bases = bases if isinstance(bases, (list, tuple)) else [bases]
        base_cfg = {}
        for rel in bases:
            c, _ = _file2dict(os.path.join(os.path.dirname(filename), rel))
            dup = base_cfg.keys() & c.keys()
            if dup:
                raise KeyError(f"Duplicate key is not allowed among bases. Duplicate keys: {dup}")
            base_cfg.update(c)
        cfg = _merge_a_into_b(cfg, base_cfg)
    return cfg, text


class Config:
    """`Config.fromfile(path)`, attribute / item access, `merge_from_dict({'a.b': 1})`."""

    def __init__(self, cfg_dict=None, cfg_text=None, filename=None):
        cfg_dict = {} if cfg_dict is None else cfg_dict
        if not isinstance(cfg_dict, dict):
            raise TypeError(f"cfg_dict must be a dict, but got {type(cfg_dict)}")
        object.__setattr__(self, "_cfg_dict", ConfigDict(cfg_dict))
        object.__setattr__(self, "_filename", filename)
        object.__setattr__(self, "_text", cfg_text or "")

    @staticmethod
    def fromfile(filename):
        cfg, text = _file2dict(filename)
        return Config(cfg, cfg_text=text, filename=filename)

    @property
    def filename(self):
        return self._filename

    @property
    def text(self):
        return self._text

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = value

    def __setitem__(self, name, value):
        self._cfg_dict[name] = value

    def __contains__(self, name):
        return name in self._cfg_dict

    def __iter__(self):
        return iter(self._cfg_dict)

    def __len__(self):
        return len(self._cfg_dict)

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def merge_from_dict(self, options):
        """`{'model.pts_bbox_head.num_classes': 5}` style overrides (train.py:25-26)."""
        nested = {}
        for full_key, v in options.items():
            d = nested
            keys = full_key.split(".")
            for sub in keys[:-1]:
                d = d.setdefault(sub, {})
            d[keys[-1]] = v
        merged = _merge_a_into_b(nested, self._cfg_dict.to_dict())
        object.__setattr__(self, "_cfg_dict", ConfigDict(merged))
