"""build_mmss_heads mirror (reference ovr/modeling/mmss_heads/mmss_heads.py:14-41): builds the
``nn.ModuleDict`` of heads named in ``cfg.MODEL.MMSS_HEAD.TYPES`` and ties ``v2l_projection`` to the
default head.  Only ``GroundingHead`` belongs to the B200 region-text path; other head types must be
registered by the caller (e.g. the reference's own TransformerHead) in ``MMSS_HEADS_REGISTRY``."""
from torch import nn

from .grounding_head import MMSS_HEADS_REGISTRY, build_grounding_head


def build_mmss_heads(cfg, *args, **kwargs):
    heads = {}
    for head_type in cfg.MODEL.MMSS_HEAD.TYPES:
        assert head_type in MMSS_HEADS_REGISTRY, \
            "cfg.MODEL.MMSS_HEAD.TYPE: {} is not registered in Registry".format(head_type)
        if head_type == "GroundingHead":
            heads[head_type] = build_grounding_head(head_type, cfg, *args, **kwargs)
        else:
            heads[head_type] = MMSS_HEADS_REGISTRY.get(head_type)(cfg, *args, **kwargs)
    if cfg.MODEL.MMSS_HEAD.TIE_VL_PROJECTION_WEIGHTS:
        default = heads[cfg.MODEL.MMSS_HEAD.DEFAULT_HEAD]
        weight, bias = default.v2l_projection.weight, default.v2l_projection.bias
        for head_type in cfg.MODEL.MMSS_HEAD.TYPES:
            if head_type == cfg.MODEL.MMSS_HEAD.DEFAULT_HEAD or not hasattr(heads[head_type], "v2l_projection"):
                continue
            assert weight.shape == heads[head_type].v2l_projection.weight.shape
            heads[head_type].v2l_projection.weight = weight
            heads[head_type].v2l_projection.bias = bias
    return nn.ModuleDict(heads)
