"""``@configurable`` construction for the drop-in modules — the calling convention Detectron2's
``build_roi_heads`` (``ROI_HEADS_REGISTRY.get(name)(cfg, input_shape)``) and the reference's
``BOX_EMBEDDING_PREDICTORS[name](cfg, input_shape)`` (box_emb_head.py:239-249, roi_emb_heads.py:135,168-214)
rely on: a class whose ``__init__`` takes explicit arguments can ALSO be called with a config node as its first
argument, in which case ``cls.from_config(cfg, ...)`` supplies the explicit arguments (extra keyword arguments that
``from_config`` does not name are forwarded to ``__init__`` as overrides).

Own implementation (Detectron2 is not a dependency of this package); any node that exposes ``.MODEL`` counts as a
config: Detectron2 / yacs ``CfgNode``, omegaconf ``DictConfig`` and ``locov_b200.modeling.config.CfgNode`` alike.
"""
import functools
import inspect

import torch


def is_config(x) -> bool:
    return x is not None and hasattr(x, "MODEL") and not isinstance(x, (torch.Tensor, torch.nn.Module, type))


def called_with_cfg(*args, **kwargs) -> bool:
    return (len(args) > 0 and is_config(args[0])) or is_config(kwargs.get("cfg"))


def args_from_config(from_config, *args, **kwargs) -> dict:
    params = inspect.signature(from_config).parameters
    if next(iter(params)) != "cfg":
        raise TypeError(f"{from_config.__qualname__} must take 'cfg' as its first argument")
    if any(p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in params.values()):
        return from_config(*args, **kwargs)
    overrides = {k: kwargs.pop(k) for k in list(kwargs) if k not in params}
    explicit = from_config(*args, **kwargs)
    explicit.update(overrides)
    return explicit


def configurable(init):
    """Decorator for ``__init__``: ``Cls(cfg, *a, **kw)`` -> ``init(self, **Cls.from_config(cfg, *a, **kw))``."""
    if init.__name__ != "__init__":
        raise TypeError("@configurable decorates __init__")

    @functools.wraps(init)
    def wrapped(self, *args, **kwargs):
        if called_with_cfg(*args, **kwargs):
            if not hasattr(type(self), "from_config"):
                raise AttributeError(f"{type(self).__name__} needs a from_config classmethod to be built from a config")
            init(self, **args_from_config(type(self).from_config, *args, **kwargs))
        else:
            init(self, *args, **kwargs)

    return wrapped
