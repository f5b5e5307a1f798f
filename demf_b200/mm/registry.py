"""Registries with the mmcv.utils.Registry surface the reference uses.

The reference registers its classes by decorator at import time
(demf/modeling/detectors/demfnet.py:12 `@DETECTORS.register_module()`,
demf/modeling/heads/class_agnostic_vote_head.py:24,335 `@HEADS...`,
demf/modeling/layers/transformer.py:39 `@TRANSFORMER_LAYER...`,
demf/core/bbox/coders/class_agnostic_bbox_coder.py:8,140 `@BBOX_CODERS...`) and builds them
from config dicts by their `type=` string (configs/demf/demf_votenet.py:26-182). This module
keeps exactly that mechanism: `Registry.register_module()`, `Registry.build(cfg)`,
`build_from_cfg(cfg, registry, default_args)`.
"""
import inspect


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return self.get(key) is not None

    def __repr__(self):
        return f"Registry(name={self._name}, items={sorted(self._module_dict)})"

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        names = [name] if isinstance(name, str) else (name or [cls.__name__])
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f"{n} is already registered in {self._name}")
            self._module_dict[n] = cls

    def register_module(self, name=None, force=False, module=None):
        if not isinstance(force, bool):
            raise TypeError(f"force must be a boolean, but got {type(force)}")
        if module is not None:
            self._register(module, name, force)
            return module

        def decorator(cls):
            self._register(cls, name, force)
            return cls

        return decorator

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    """Instantiate `registry[cfg['type']](**rest_of_cfg)`; mirrors mmcv.utils.build_from_cfg."""
    if not isinstance(cfg, dict):
        raise TypeError(f"cfg must be a dict, but got {type(cfg)}")
    if "type" not in cfg and not (default_args and "type" in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}')
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f"{obj_type} is not in the {registry.name} registry")
    elif inspect.isclass(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError(f"type must be a str or valid type, but got {type(obj_type)}")
    try:
        return obj_cls(**args)
    except Exception as e:  # same re-raise style as mmcv: name the class that failed
        raise type(e)(f"{obj_cls.__name__}: {e}") from e


# mmdet / mmdet3d / mmcv registries the reference decorates or builds from
DETECTORS = Registry("detector")
BACKBONES = Registry("backbone")
NECKS = Registry("neck")
HEADS = Registry("head")
LOSSES = Registry("loss")
BBOX_CODERS = Registry("bbox_coder")
SA_MODULES = Registry("point_sa_module")
ATTENTION = Registry("attention")
FEEDFORWARD_NETWORK = Registry("feed-forward Network")
TRANSFORMER_LAYER = Registry("transformerLayer")
TRANSFORMER_LAYER_SEQUENCE = Registry("transformer-layers sequence")
POSITIONAL_ENCODING = Registry("position encoding")


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_neck(cfg):
    return NECKS.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_bbox_coder(cfg, **default_args):
    return build_from_cfg(cfg, BBOX_CODERS, default_args)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    return DETECTORS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg))


def build_model(cfg, train_cfg=None, test_cfg=None):
    """mmdet3d.models.build_model (reference: train.py:107-110)."""
    return build_detector(cfg, train_cfg=train_cfg, test_cfg=test_cfg)


def build_attention(cfg, default_args=None):
    return build_from_cfg(cfg, ATTENTION, default_args)


def build_feedforward_network(cfg, default_args=None):
    return build_from_cfg(cfg, FEEDFORWARD_NETWORK, default_args)


def build_transformer_layer(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER, default_args)


def build_transformer_layer_sequence(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER_SEQUENCE, default_args)


def build_positional_encoding(cfg, default_args=None):
    return build_from_cfg(cfg, POSITIONAL_ENCODING, default_args)


def build_sa_module(cfg, *args, **kwargs):
    """mmdet3d.ops.build_sa_module: default type PointSAModule, extra kwargs override cfg."""
    if cfg is None:
        cfg_ = dict(type="PointSAModule")
    else:
        if not isinstance(cfg, dict):
            raise TypeError("cfg must be a dict")
        if "type" not in cfg:
            raise KeyError('the cfg dict must contain the key "type"')
        cfg_ = dict(cfg)
    module_type = cfg_.pop("type")
    sa_module = SA_MODULES.get(module_type)
    if sa_module is None:
        raise KeyError(f"Unrecognized module type {module_type}")
    return sa_module(*args, **kwargs, **cfg_)
