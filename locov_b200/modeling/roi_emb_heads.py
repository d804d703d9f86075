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

from .box_emb_head import build_box_predictor
from .poolers import ROIPooler
from .registry import ROI_HEADS_REGISTRY


@ROI_HEADS_REGISTRY.register()
class EmbeddingRes5ROIHeads(nn.Module):
    def __init__(self, *, in_features: List[str], pooler: ROIPooler, res5: nn.Module, box_predictor: nn.Module,
                 mask_head: Optional[nn.Module] = None, output_shape: Optional[int] = 0,
                 proposal_sampler: Optional[Callable] = None, num_classes: Optional[int] = None, **kwargs):
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
        self.proposal_sampler = proposal_sampler

    @classmethod
    def from_cfg(cls, cfg, input_shape, res5: nn.Module, out_channels: int = None, proposal_sampler=None):
        """Mirror of from_config (:168-214); ``res5`` is supplied by the caller (Detectron2's
        ``_build_res5_block`` result) because the conv stage is outside this package's scope."""
        in_features = cfg.MODEL.ROI_HEADS.IN_FEATURES
        assert not cfg.MODEL.KEYPOINT_ON
        assert len(in_features) == 1
        stride = input_shape[in_features[0]].stride
        pooler = ROIPooler(output_size=cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION, scales=(1.0 / stride,),
                           sampling_ratio=cfg.MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO,
                           pooler_type=cfg.MODEL.ROI_BOX_HEAD.POOLER_TYPE)
        if out_channels is None:
            out_channels = cfg.MODEL.RESNETS.RES2_OUT_CHANNELS * 8
        return cls(in_features=in_features, pooler=pooler, res5=res5,
                   box_predictor=build_box_predictor(cfg, input_shape=out_channels), output_shape=out_channels,
                   proposal_sampler=proposal_sampler, num_classes=cfg.MODEL.ROI_HEADS.NUM_CLASSES)

    def label_and_sample_proposals(self, proposals, targets):
        if self.proposal_sampler is None:
            assert all(p.has("gt_classes") for p in proposals), \
                "no proposal_sampler was injected: proposals must already carry gt_classes"
            return proposals
        return self.proposal_sampler(proposals, targets)

    def _shared_roi_transform(self, features, boxes):
        x = self.pooler(features, boxes)
        return self.res5(x)

    def forward(self, images, features, proposals, targets=None):
        del images
        if self.training:
            assert targets is not None
            proposals = self.label_and_sample_proposals(proposals, targets)
        del targets
        proposal_boxes = [x.proposal_boxes for x in proposals]
        box_features = self._shared_roi_transform([features[f] for f in self.in_features], proposal_boxes)
        predictions = self.box_predictor(box_features.mean(dim=[2, 3]))
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
        box_features = box_features.mean(dim=[2, 3])
        predictions = self.box_predictor(box_features)
        box_features = list(box_features.split(boxes_per_image, dim=0))
        losses.update(self.box_predictor.losses(predictions, proposals))
        return visual_grid_features, box_features, proposals, losses

    def inference_detection(self, features, proposals):
        proposal_boxes = [x.proposal_boxes for x in proposals]
        box_features = self._shared_roi_transform([features[f] for f in self.in_features], proposal_boxes)
        predictions = self.box_predictor(box_features.mean(dim=[2, 3]))
        pred_instances, _ = self.box_predictor.inference(predictions, proposals)
        return self.forward_with_given_boxes(features, pred_instances), {}
