"""Embedding box predictor — drop-in for the reference's ``EmbeddingFastRCNNOutputLayers``
(ovr/modeling/roi_heads/box_emb_head.py:60-236) including the behaviour it inherits from Detectron2's
``FastRCNNOutputLayers`` (.losses / .inference / .predict_probs / .predict_boxes).

Kept: constructor keywords, ``from_config`` keys, parameter names and layouts (``emb_pred.{weight,bias}``
[768,2048]/[768], ``bbox_pred.{weight,bias}`` [4,2048]/[4], ``cls_score.{weight,bias}`` [K+1,768]/[K+1]),
``set_class_embeddings`` (re-settable at run time), ``detach_cls_predictor`` semantics and loss keys.
New: the three GEMMs run on the tcgen05 kernels with softmax / log-sum-exp / argmax fused into the
scoring epilogue (locov_b200.functional.box_predict); ``losses`` re-uses the fused statistics.
"""
import math
from typing import Dict, List, Tuple, Union

import torch
from torch import nn
from torch.nn import functional as F

from .. import functional as LF
from .logged_module import normalize_vec, standardize_vec
from .registry import BOX_PREDICTORS
from .structures import Boxes, Instances, ShapeSpec

_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)


class Box2BoxTransform:
    """Detectron2 Box2BoxTransform (R-CNN box parameterisation)."""

    def __init__(self, weights: Tuple[float, float, float, float], scale_clamp: float = _DEFAULT_SCALE_CLAMP):
        self.weights = tuple(weights)
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        sw = src_boxes[:, 2] - src_boxes[:, 0]
        sh = src_boxes[:, 3] - src_boxes[:, 1]
        sx = src_boxes[:, 0] + 0.5 * sw
        sy = src_boxes[:, 1] + 0.5 * sh
        tw = target_boxes[:, 2] - target_boxes[:, 0]
        th = target_boxes[:, 3] - target_boxes[:, 1]
        tx = target_boxes[:, 0] + 0.5 * tw
        ty = target_boxes[:, 1] + 0.5 * th
        wx, wy, ww, wh = self.weights
        return torch.stack((wx * (tx - sx) / sw, wy * (ty - sy) / sh, ww * torch.log(tw / sw), wh * torch.log(th / sh)), 1)

    def apply_deltas(self, deltas, boxes):
        deltas = deltas.float()
        boxes = boxes.to(deltas.dtype)
        w = boxes[:, 2] - boxes[:, 0]
        h = boxes[:, 3] - boxes[:, 1]
        cx = boxes[:, 0] + 0.5 * w
        cy = boxes[:, 1] + 0.5 * h
        wx, wy, ww, wh = self.weights
        dx = deltas[:, 0::4] / wx
        dy = deltas[:, 1::4] / wy
        dw = torch.clamp(deltas[:, 2::4] / ww, max=self.scale_clamp)
        dh = torch.clamp(deltas[:, 3::4] / wh, max=self.scale_clamp)
        pcx = dx * w[:, None] + cx[:, None]
        pcy = dy * h[:, None] + cy[:, None]
        pw = torch.exp(dw) * w[:, None]
        ph = torch.exp(dh) * h[:, None]
        out = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), -1)
        return out.reshape(deltas.shape)


def _cat(tensors, dim=0):
    return tensors[0] if len(tensors) == 1 else torch.cat(tensors, dim)


def _box_tensor(b):
    return b.tensor if hasattr(b, "tensor") else b


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image):
    """Detectron2 fast_rcnn_inference_single_image (kept in PyTorch/torchvision: NMS is outside the
    named hot path)."""
    from torchvision.ops import batched_nms
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not bool(valid.all()):
        boxes, scores = boxes[valid], scores[valid]
    scores = scores[:, :-1]
    num_bbox_reg_classes = boxes.shape[1] // 4
    bx = Boxes(boxes.reshape(-1, 4))
    bx.clip(image_shape)
    boxes = bx.tensor.view(-1, num_bbox_reg_classes, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    boxes = boxes[filter_inds[:, 0], 0] if num_bbox_reg_classes == 1 else boxes[filter_mask]
    scores = scores[filter_mask]
    keep = batched_nms(boxes.float(), scores, filter_inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    boxes, scores, filter_inds = boxes[keep], scores[keep], filter_inds[keep]
    result = Instances(image_shape)
    result.pred_boxes = Boxes(boxes)
    result.scores = scores
    result.pred_classes = filter_inds[:, 1]
    return result, filter_inds[:, 0]


@BOX_PREDICTORS.register()
class EmbeddingFastRCNNOutputLayers(nn.Module):
    def __init__(self, input_shape, *, box2box_transform=None, num_classes: int = 80, test_score_thresh: float = 0.0,
                 test_nms_thresh: float = 0.5, test_topk_per_image: int = 100, cls_agnostic_bbox_reg: bool = False,
                 smooth_l1_beta: float = 0.0, box_reg_loss_type: str = "smooth_l1",
                 loss_weight: Union[float, Dict[str, float]] = 1.0, emb_dim: int = 768, embedding_based: bool = True,
                 freeze_emb_pred: bool = True, normalize_emb: bool = False, standardize_emb: bool = False,
                 detach_cls_predictor: bool = False, precision: str = "fp32"):
        super().__init__()
        if not isinstance(input_shape, int) and not hasattr(input_shape, "channels"):
            # (cfg, input_shape) calling convention of @configurable / build_box_predictor
            raise TypeError("use EmbeddingFastRCNNOutputLayers.from_cfg(cfg, input_shape) to build from a config")
        if isinstance(input_shape, int):
            input_shape = ShapeSpec(channels=input_shape)
        num_inputs = input_shape.channels * (input_shape.width or 1) * (input_shape.height or 1)
        if not embedding_based:
            raise NotImplementedError("EMBEDDING_BASED=False is Detectron2's stock FastRCNNOutputLayers, not the LocOV path")
        assert cls_agnostic_bbox_reg, "the embedding head requires CLS_AGNOSTIC_BBOX_REG (box_emb_head.py:137)"
        self.box2box_transform = box2box_transform or Box2BoxTransform((10.0, 10.0, 5.0, 5.0))
        self.smooth_l1_beta = smooth_l1_beta
        self.test_score_thresh = test_score_thresh
        self.test_nms_thresh = test_nms_thresh
        self.test_topk_per_image = test_topk_per_image
        self.box_reg_loss_type = box_reg_loss_type
        if isinstance(loss_weight, float):
            loss_weight = {"loss_cls": loss_weight, "loss_box_reg": loss_weight}
        self.loss_weight = dict(loss_weight)
        self.precision = precision

        box_dim = len(self.box2box_transform.weights)
        self.bbox_pred = nn.Linear(num_inputs, box_dim)                 # class-agnostic: 1 x 4
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        nn.init.constant_(self.bbox_pred.bias, 0)

        self.embedding_based = embedding_based
        self.normalize_emb = normalize_emb
        self.standardize_emb = standardize_emb
        self.emb_dim = emb_dim
        self.emb_pred = nn.Linear(num_inputs, self.emb_dim)
        nn.init.normal_(self.emb_pred.weight, mean=0, std=0.01)
        nn.init.constant_(self.emb_pred.bias, 0)
        # forward() can't be used until set_class_embeddings() is called (box_emb_head.py:138-140)
        self.num_classes = None
        self.cls_score = None
        if freeze_emb_pred:
            self.emb_pred.weight.requires_grad = False
            self.emb_pred.bias.requires_grad = False
        self.detach_cls_predictor = detach_cls_predictor
        if self.detach_cls_predictor:
            self.loss_weight.update({"loss_cls": 0.0})
        self._aux = None      # fused softmax statistics of the last forward (lse / probs / argmax)

    # ---- construction from a Detectron2-style config (box_emb_head.py:151-177) ------------------------
    @classmethod
    def from_config(cls, cfg, input_shape):
        return {
            "input_shape": input_shape,
            "box2box_transform": Box2BoxTransform(weights=cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS),
            "num_classes": cfg.MODEL.ROI_HEADS.NUM_CLASSES,
            "cls_agnostic_bbox_reg": cfg.MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG,
            "smooth_l1_beta": cfg.MODEL.ROI_BOX_HEAD.SMOOTH_L1_BETA,
            "test_score_thresh": cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST,
            "test_nms_thresh": cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST,
            "test_topk_per_image": cfg.TEST.DETECTIONS_PER_IMAGE,
            "box_reg_loss_type": cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_TYPE,
            "loss_weight": {"loss_box_reg": cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_WEIGHT},
            "emb_dim": cfg.MODEL.ROI_BOX_HEAD.EMB_DIM,
            "embedding_based": cfg.MODEL.ROI_BOX_HEAD.EMBEDDING_BASED,
            "freeze_emb_pred": cfg.MODEL.ROI_BOX_HEAD.FREEZE_EMB_PRED,
            "normalize_emb": cfg.MODEL.ROI_BOX_HEAD.NORMALIZE_EMB_PRED,
            "standardize_emb": cfg.MODEL.ROI_BOX_HEAD.STANDARDIZE_EMB_PRED,
            "detach_cls_predictor": cfg.MODEL.ROI_HEADS.DETACH_CLASS_PREDICTOR,
        }

    @classmethod
    def from_cfg(cls, cfg, input_shape):
        kw = cls.from_config(cfg, input_shape)
        shape = kw.pop("input_shape")
        lw = kw["loss_weight"]
        kw["loss_weight"] = {"loss_cls": 1.0, **lw}
        b200 = getattr(cfg.MODEL, "B200", None)
        if b200 is not None and hasattr(b200, "PRECISION"):
            kw["precision"] = b200.PRECISION
        return cls(shape, **kw)

    # ---- forward (box_emb_head.py:179-212) ------------------------------------------------------------
    def forward(self, x):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        if self.cls_score is None:
            raise RuntimeError("set_class_embeddings() must be called before forward (box_emb_head.py:138)")
        x = x.to(torch.float32).contiguous()
        if self.normalize_emb or self.standardize_emb:
            return self._forward_unfused(x)
        ep, bp, cs = self.emb_pred, self.bbox_pred, self.cls_score
        if self.detach_cls_predictor:
            # the classification branch sees x.detach() under no_grad (box_emb_head.py:197-199): scores
            # carry no graph; deltas keep theirs.
            with torch.no_grad():
                scores, deltas, aux = LF.box_predict(x.detach(), ep.weight, ep.bias, bp.weight, bp.bias, cs.weight,
                                                     cs.bias, self.precision, want_probs=not self.training)
            if torch.is_grad_enabled() and (x.requires_grad or bp.weight.requires_grad or bp.bias.requires_grad):
                deltas = LF.linear(x, bp.weight, bp.bias, self.precision)
        else:
            scores, deltas, aux = LF.box_predict(x, ep.weight, ep.bias, bp.weight, bp.bias, cs.weight, cs.bias,
                                                 self.precision, want_probs=not self.training)
        self._aux = (scores, aux)
        return scores, deltas

    def _forward_unfused(self, x):
        """NORMALIZE_EMB_PRED / STANDARDIZE_EMB_PRED (both off in the shipped configs, config.py:131,133):
        the row-wise normalisation sits between the two GEMMs, so they run as separate kernels."""
        deltas = LF.linear(x, self.bbox_pred.weight, self.bbox_pred.bias, self.precision)
        ctx = torch.no_grad() if self.detach_cls_predictor else torch.enable_grad()
        with ctx:
            xs = x.detach() if self.detach_cls_predictor else x
            scores = self.forward_cls_prediction(xs)
        self._aux = None
        return scores, deltas

    def forward_cls_prediction(self, x):
        e = LF.linear(x, self.emb_pred.weight, self.emb_pred.bias, self.precision)
        if self.normalize_emb:
            e = normalize_vec(e, dim=1)
        if self.standardize_emb:
            e = standardize_vec(e, dim=1)
        return LF.linear(e, self.cls_score.weight, self.cls_score.bias, self.precision)

    # ---- text matrix (box_emb_head.py:214-236) ----------------------------------------------------------
    def set_class_embeddings(self, embs):
        device = self.emb_pred.weight.device
        self.num_classes = embs.shape[0] - 1      # includes background
        self.cls_score = nn.Linear(self.emb_dim, self.num_classes + 1)
        self.cls_score.to(device)
        if torch.is_tensor(embs):
            embs = embs.clone().detach().to(device=device, dtype=torch.float32)
        else:
            embs = torch.tensor(embs, device=device, dtype=torch.float32)
        if self.normalize_emb or self.standardize_emb:
            assert embs.shape[1] == self.emb_dim, "The embedding dimension has to match the one saved in the model"
        if self.normalize_emb:
            embs = normalize_vec(embs, dim=1)
        if self.standardize_emb:
            embs = standardize_vec(embs, dim=1)
        self.cls_score.weight.data = embs
        self.cls_score.bias.data = torch.zeros_like(self.cls_score.bias.data)
        self.cls_score.weight.requires_grad = False
        self.cls_score.bias.requires_grad = False
        self._aux = None

    # ---- Detectron2 FastRCNNOutputLayers behaviour ------------------------------------------------------
    def _fused_aux(self, scores):
        if self._aux is not None and self._aux[0] is scores:
            return self._aux[1]
        return None

    def losses(self, predictions, proposals):
        scores, proposal_deltas = predictions
        gt_classes = _cat([p.gt_classes for p in proposals], 0) if len(proposals) else torch.empty(0, device=scores.device)
        if len(proposals):
            proposal_boxes = _cat([_box_tensor(p.proposal_boxes) for p in proposals], 0)
            assert not proposal_boxes.requires_grad, "Proposals should not require gradients!"
            gt_boxes = _cat([_box_tensor(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes) for p in proposals], 0)
        else:
            proposal_boxes = gt_boxes = torch.empty((0, 4), device=proposal_deltas.device)
        aux = self._fused_aux(scores)
        if scores.shape[0] == 0:
            loss_cls = scores.sum() * 0.0
        elif aux is not None and scores.is_cuda:
            loss_cls = LF.box_cross_entropy(scores, aux.lse, gt_classes)
        else:
            loss_cls = F.cross_entropy(scores, gt_classes, reduction="mean")
        losses = {"loss_cls": loss_cls,
                  "loss_box_reg": self.box_reg_loss(proposal_boxes, gt_boxes, proposal_deltas, gt_classes)}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    def box_reg_loss(self, proposal_boxes, gt_boxes, pred_deltas, gt_classes):
        box_dim = proposal_boxes.shape[1]
        fg_inds = torch.nonzero((gt_classes >= 0) & (gt_classes < self.num_classes), as_tuple=True)[0]
        fg_pred_deltas = pred_deltas[fg_inds] if pred_deltas.shape[1] == box_dim else \
            pred_deltas.view(-1, self.num_classes, box_dim)[fg_inds, gt_classes[fg_inds]]
        if self.box_reg_loss_type != "smooth_l1":
            raise NotImplementedError(f"box_reg_loss_type {self.box_reg_loss_type!r} (shipped configs use smooth_l1)")
        tgt = self.box2box_transform.get_deltas(proposal_boxes[fg_inds], gt_boxes[fg_inds])
        if self.smooth_l1_beta < 1e-5:
            loss = torch.abs(fg_pred_deltas - tgt).sum()
        else:
            n = torch.abs(fg_pred_deltas - tgt)
            loss = torch.where(n < self.smooth_l1_beta, 0.5 * n ** 2 / self.smooth_l1_beta, n - 0.5 * self.smooth_l1_beta).sum()
        return loss / max(gt_classes.numel(), 1.0)

    def predict_probs(self, predictions, proposals):
        scores, _ = predictions
        aux = self._fused_aux(scores)
        probs = aux.probs if (aux is not None and aux.probs is not None) else F.softmax(scores, dim=-1)
        return probs.split([len(p) for p in proposals], dim=0)

    def predict_classes(self, predictions):
        """Per-RoI argmax over the K foreground columns (int64), from the fused epilogue."""
        scores, _ = predictions
        aux = self._fused_aux(scores)
        return aux.argmax_fg if aux is not None else F.softmax(scores, -1)[:, :-1].argmax(1)

    def predict_boxes(self, predictions, proposals):
        if not len(proposals):
            return []
        _, proposal_deltas = predictions
        proposal_boxes = _cat([_box_tensor(p.proposal_boxes) for p in proposals], 0)
        boxes = self.box2box_transform.apply_deltas(proposal_deltas, proposal_boxes)
        return boxes.split([len(p) for p in proposals])

    def inference(self, predictions, proposals):
        boxes = self.predict_boxes(predictions, proposals)
        scores = self.predict_probs(predictions, proposals)
        image_shapes = [x.image_size for x in proposals]
        results = [fast_rcnn_inference_single_image(b, s, shp, self.test_score_thresh, self.test_nms_thresh,
                                                    self.test_topk_per_image)
                   for b, s, shp in zip(boxes, scores, image_shapes)]
        return [r[0] for r in results], [r[1] for r in results]


def build_box_predictor(cfg, input_shape):
    """box_emb_head.py:239-249 — resolves cfg.MODEL.ROI_BOX_HEAD.NAME."""
    name = cfg.MODEL.ROI_BOX_HEAD.NAME
    if name not in BOX_PREDICTORS:
        raise KeyError(f"box predictor {name!r} is not part of the B200 region-text path "
                       f"(available: {sorted(BOX_PREDICTORS)})")
    return BOX_PREDICTORS[name].from_cfg(cfg, input_shape)
