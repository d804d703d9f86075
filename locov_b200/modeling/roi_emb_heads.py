"""C4 embedding ROI heads — drop-ins for ``EmbeddingRes5ROIHeads`` and
``EmbeddingProposalsRes5ROIHeads`` (reference ovr/modeling/roi_heads/roi_emb_heads.py:121-360).

On the hot path: ``self.pooler`` (RoIAlign, :182-187/:243-245), ``box_features.mean(dim=[2,3])``
(:262) and ``self.box_predictor`` + ``.losses`` / ``.inference`` (:263-282, :345-357).
NOT on the hot path and therefore injected, not re-implemented: the ``res5``
stage (Detectron2 BottleneckBlocks, cuDNN) and ``label_and_sample_proposals`` (IoU matcher/sampler) —
pass Detectron2's own objects; with ``proposal_sampler=None`` the proposals must already carry
``gt_classes`` (which is what the LSM-stage mapper produces, coco_mappers.py:88-106).
"""
from typing import Callable, List, Optional

import torch
from torch import nn

from .. import functional as LF
from .box_emb_head import build_box_predictor
from .configurable import configurable
from .poolers import ROIPooler
from .registry import ROI_HEADS_REGISTRY
from .res5 import build_res5_block


@ROI_HEADS_REGISTRY.register()
class EmbeddingRes5ROIHeads(nn.Module):
    """Constructible both ways the reference's class is (roi_emb_heads.py:129-214): with explicit keyword arguments, or as
    ``EmbeddingRes5ROIHeads(cfg, input_shape)`` — the call Detectron2's ``build_roi_heads`` makes — which goes through
    ``from_config``.  Keyword arguments of Detectron2's ``ROIHeads`` base (``num_classes``, ``batch_size_per_image``,
    ``positive_fraction``, ``proposal_matcher``, ``proposal_append_gt``) are accepted and kept as attributes."""

    @configurable
    def __init__(self, *, in_features: List[str], pooler: ROIPooler, res5: nn.Module, box_predictor: nn.Module,
                 mask_head: Optional[nn.Module] = None, output_shape: Optional[int] = 0,
                 proposal_sampler: Optional[Callable] = None, num_classes: Optional[int] = None, batch_size_per_image: int = 512,
                 positive_fraction: float = 0.25, proposal_matcher=None, proposal_append_gt: bool = True):
        super().__init__()
        if mask_head is not None:
            raise NotImplementedError("mask heads are dead code in the reference (roi_emb_heads.py:206,268: un-imported helpers)")
        self.in_features = in_features
        self.pooler = pooler
        if isinstance(res5, (list, tuple)):
            res5 = nn.Sequential(*res5)
        self.res5 = res5
        self.output_shape = output_shape
        self.box_predictor = box_predictor
        self.mask_on = False
        self.num_classes = num_classes
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.proposal_matcher = proposal_matcher
        self.proposal_append_gt = proposal_append_gt
        self.proposal_sampler = proposal_sampler

    @classmethod
    def from_config(cls, cfg, input_shape, res5: Optional[nn.Module] = None, out_channels: Optional[int] = None,
                    proposal_sampler: Optional[Callable] = None):
        """roi_emb_heads.py:168-214.  ``res5`` may be injected (e.g. a channels-last / bf16 variant of the stage); by default it
        is built by ``_build_res5_block`` exactly as the reference does."""
        roi = cfg.MODEL.ROI_HEADS
        in_features = roi.IN_FEATURES
        assert not cfg.MODEL.KEYPOINT_ON
        assert len(in_features) == 1
        if getattr(cfg.MODEL, "MASK_ON", False):
            raise NotImplementedError("MODEL.MASK_ON: the reference's mask path calls un-imported helpers (roi_emb_heads.py:206)")
        stride = input_shape[in_features[0]].stride
        b200 = getattr(cfg.MODEL, "B200", None)                 # this library's optional keys; absent in the reference's cfg
        ret = {
            "in_features": in_features,
            "num_classes": roi.NUM_CLASSES,
            "batch_size_per_image": roi.BATCH_SIZE_PER_IMAGE,
            "positive_fraction": roi.POSITIVE_FRACTION,
            "proposal_append_gt": getattr(roi, "PROPOSAL_APPEND_GT", True),
            "proposal_sampler": proposal_sampler,
            "pooler": ROIPooler(output_size=cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION, scales=(1.0 / stride,),
                                sampling_ratio=cfg.MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO,
                                pooler_type=cfg.MODEL.ROI_BOX_HEAD.POOLER_TYPE,
                                channels_last=bool(getattr(b200, "POOLER_CHANNELS_LAST", False)),
                                out_dtype=torch.bfloat16 if getattr(b200, "POOLER_BF16", False) else torch.float32),
        }
        if res5 is None:
            res5, out_channels = cls._build_res5_block(cfg)
        elif out_channels is None:
            out_channels = cfg.MODEL.RESNETS.RES2_OUT_CHANNELS * 8
        ret["res5"] = res5
        ret["box_predictor"] = build_box_predictor(cfg, input_shape=out_channels)
        ret["output_shape"] = out_channels
        return ret

    @classmethod
    def from_cfg(cls, cfg, input_shape, res5: Optional[nn.Module] = None, out_channels: Optional[int] = None, proposal_sampler=None):
        """Explicit spelling of ``cls(cfg, input_shape, ...)`` (kept from round 1)."""
        return cls(cfg, input_shape, res5=res5, out_channels=out_channels, proposal_sampler=proposal_sampler)

    @classmethod
    def _build_res5_block(cls, cfg):
        return build_res5_block(cfg)

    def label_and_sample_proposals(self, proposals, targets):
        if self.proposal_sampler is None:
            assert all(p.has("gt_classes") for p in proposals), \
                "no proposal_sampler was injected: proposals must already carry gt_classes"
            return proposals
        return self.proposal_sampler(proposals, targets)

    def _shared_roi_transform(self, features, boxes):
        x = self.pooler(features, boxes)
        return self.res5(x)

    def _pooled_mean(self, box_features):
        """``box_features.mean(dim=[2, 3])`` (roi_emb_heads.py:262, :329, :351) as one pass that also writes the bf16 operand of
        the predictor's projection GEMM.  On the CPU (the CPU unit tests of the host logic) it is the torch expression."""
        if not box_features.is_cuda:
            return box_features.mean(dim=[2, 3])
        # fp32 means whatever the dtype of the res5 output (a bf16 stage included): the predictor computes in fp32 operands anyway, and the
        # tensor carries the bf16 operand written by the same pass
        return LF.spatial_mean(box_features, getattr(self.box_predictor, "precision", None))

    def forward(self, images, features, proposals, targets=None):
        del images
        if self.training:
            assert targets is not None
            proposals = self.label_and_sample_proposals(proposals, targets)
        del targets
        proposal_boxes = [x.proposal_boxes for x in proposals]
        box_features = self._shared_roi_transform([features[f] for f in self.in_features], proposal_boxes)
        predictions = self.box_predictor(self._pooled_mean(box_features))
        if self.training:
            del features
            return [], self.box_predictor.losses(predictions, proposals)
        pred_instances, _ = self.box_predictor.inference(predictions, proposals)
        return self.forward_with_given_boxes(features, pred_instances), {}

    def forward_with_given_boxes(self, features, instances):
        assert not self.training
        assert instances[0].has("pred_boxes") and instances[0].has("pred_classes")
        return instances


@ROI_HEADS_REGISTRY.register()
class EmbeddingProposalsRes5ROIHeads(EmbeddingRes5ROIHeads):
    def forward(self, images, features, proposals, targets=None):
        del images
        if targets is None:
            return self.inference_detection(features, proposals)
        proposals = self.label_and_sample_proposals(proposals, targets)
        del targets
        # grid features fed to the multimodal heads (roi_emb_heads.py:323)
        visual_grid_features = self.res5(features[self.in_features[0]])
        proposal_boxes = [x.proposal_boxes for x in proposals]
        boxes_per_image = [len(x) for x in proposals]
        box_features = self._shared_roi_transform([features[f] for f in self.in_features], proposal_boxes)
        del features
        losses = {}
        box_features = self._pooled_mean(box_features)
        predictions = self.box_predictor(box_features)
        box_features = list(box_features.split(boxes_per_image, dim=0))
        losses.update(self.box_predictor.losses(predictions, proposals))
        return visual_grid_features, box_features, proposals, losses

    def inference_detection(self, features, proposals):
        proposal_boxes = [x.proposal_boxes for x in proposals]
        box_features = self._shared_roi_transform([features[f] for f in self.in_features], proposal_boxes)
        predictions = self.box_predictor(self._pooled_mean(box_features))
        pred_instances, _ = self.box_predictor.inference(predictions, proposals)
        return self.forward_with_given_boxes(features, pred_instances), {}
