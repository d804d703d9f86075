"""Embedding box predictor — drop-in for the reference's ``EmbeddingFastRCNNOutputLayers``
(ovr/modeling/roi_heads/box_emb_head.py:60-236) including the behaviour it inherits from Detectron2's
``FastRCNNOutputLayers`` (.losses / .inference / .predict_probs / .predict_boxes).

Kept: constructor keywords, ``from_config`` keys, parameter names and layouts (``emb_pred.{weight,bias}``
[768,2048]/[768], ``bbox_pred.{weight,bias}`` [4,2048]/[4], ``cls_score.{weight,bias}`` [K+1,768]/[K+1]),
``set_class_embeddings`` (re-settable at run time), ``detach_cls_predictor`` semantics and loss keys.
New: the three GEMMs run on the tcgen05 kernels with softmax / log-sum-exp / argmax fused into the
scoring epilogue (locov_b200.functional.box_predict); ``losses`` re-uses the fused statistics.
"""
import contextlib
import math
from typing import Dict, List, Tuple, Union

import torch
from torch import nn

from .. import functional as LF
from .. import ops
from .configurable import configurable
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
    """Detectron2 fast_rcnn_inference_single_image with torch / torchvision operators: the path for CPU tensors (host-logic
    tests) and for settings outside loco_box_inference's range; CUDA inference goes through the kernel (see ``inference``)."""
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
    """Constructible both ways the reference's class is (box_emb_head.py:67-177): ``cls(input_shape, *, box2box_transform=...,
    ...)`` with explicit arguments, or ``cls(cfg, input_shape)`` — what ``build_box_predictor`` does (:249) — which goes
    through ``from_config``."""

    @configurable
    def __init__(self, input_shape, *, box2box_transform=None, num_classes: int = 80, test_score_thresh: float = 0.0,
                 test_nms_thresh: float = 0.5, test_topk_per_image: int = 100, cls_agnostic_bbox_reg: bool = False,
                 smooth_l1_beta: float = 0.0, box_reg_loss_type: str = "smooth_l1",
                 loss_weight: Union[float, Dict[str, float]] = 1.0, emb_dim: int = 768, embedding_based: bool = True,
                 freeze_emb_pred: bool = True, normalize_emb: bool = False, standardize_emb: bool = False,
                 detach_cls_predictor: bool = False, precision: str = "fp32"):
        super().__init__()
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
        if isinstance(loss_weight, (int, float)):
            loss_weight = {"loss_cls": float(loss_weight), "loss_box_reg": float(loss_weight)}
        self.loss_weight = dict(loss_weight)      # a missing key weighs 1.0 (Detectron2: loss_weight.get(k, 1.0))
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
            # the one key this package adds (optional): tensor-core operand mode, "fp32" (reference numerics) or "bf16"
            "precision": getattr(getattr(cfg.MODEL, "B200", None), "PRECISION", "fp32"),
        }

    @classmethod
    def from_cfg(cls, cfg, input_shape):
        """Explicit spelling of ``cls(cfg, input_shape)`` (kept from round 1)."""
        return cls(cfg, input_shape)

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
        """NORMALIZE_EMB_PRED / STANDARDIZE_EMB_PRED (both off in the shipped configs, config.py:131,133): the row-wise
        normalisation sits between the two GEMMs, so projection, normalisation and scoring are three launches — the
        scoring one still carries the fused softmax statistics."""
        deltas = LF.linear(x, self.bbox_pred.weight, self.bbox_pred.bias, self.precision)
        with (torch.no_grad() if self.detach_cls_predictor else contextlib.nullcontext()):
            xs = x.detach() if self.detach_cls_predictor else x
            scores, aux = self._scores_unfused(xs)
        self._aux = (scores, aux)
        return scores, deltas

    def _scores_unfused(self, x):
        e = LF.linear(x, self.emb_pred.weight, self.emb_pred.bias, self.precision)
        if self.normalize_emb:
            e = LF.normalize_rows(e)
        if self.standardize_emb:
            e = LF.standardize_rows(e)
        return LF.box_score(e, self.cls_score.weight, self.cls_score.bias, self.precision, want_probs=not self.training)

    def forward_cls_prediction(self, x):
        """box_emb_head.py:204-212 (scores only)."""
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        x = x.to(torch.float32).contiguous()
        if self.normalize_emb or self.standardize_emb:
            scores, aux = self._scores_unfused(x)
        else:
            ep, bp, cs = self.emb_pred, self.bbox_pred, self.cls_score
            scores, _, aux = LF.box_predict(x, ep.weight, ep.bias, bp.weight, bp.bias, cs.weight, cs.bias, self.precision,
                                            want_probs=not self.training)
        self._aux = (scores, aux)
        return scores

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
        # one-off preparation of the text matrix: on the device when the module already lives there (the reference calls this
        # after .to(device), trainer.py:365-407), with the host formulas of logged_module.py:55-72 while it is still on the CPU
        if self.normalize_emb:
            embs = LF.normalize_rows(embs) if embs.is_cuda else normalize_vec(embs, dim=1)
        if self.standardize_emb:
            embs = LF.standardize_rows(embs) if embs.is_cuda else standardize_vec(embs, dim=1)
        self.cls_score.weight.data = embs
        self.cls_score.bias.data = torch.zeros_like(self.cls_score.bias.data)
        self.cls_score.weight.requires_grad = False
        self.cls_score.bias.requires_grad = False
        self._aux = None

    # ---- Detectron2 FastRCNNOutputLayers behaviour ------------------------------------------------------
    def _fused_aux(self, scores, want_probs=False):
        """Softmax statistics of ``scores``: the ones the scoring epilogue produced when ``scores`` is the tensor the last
        forward returned, otherwise computed from the given matrix by one bandwidth kernel (loco_box_softmax)."""
        if self._aux is not None and self._aux[0] is scores:
            aux = self._aux[1]
            if not want_probs or aux.probs is not None:
                return aux
        lse, arg, probs = ops.box_softmax(scores.detach().to(torch.float32), want_probs=want_probs)
        aux = LF.BoxScoreAux(lse, probs, arg)
        self._aux = (scores, aux)
        return aux

    def losses(self, predictions, proposals):
        scores, proposal_deltas = predictions
        gt_classes = _cat([p.gt_classes for p in proposals], 0) if len(proposals) else torch.empty(0, device=scores.device)
        if len(proposals):
            proposal_boxes = _cat([_box_tensor(p.proposal_boxes) for p in proposals], 0)
            assert not proposal_boxes.requires_grad, "Proposals should not require gradients!"
            gt_boxes = _cat([_box_tensor(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes) for p in proposals], 0)
        else:
            proposal_boxes = gt_boxes = torch.empty((0, 4), device=proposal_deltas.device)
        if scores.shape[0] == 0:
            loss_cls = scores.sum() * 0.0
        else:
            loss_cls = LF.box_cross_entropy(scores, self._fused_aux(scores).lse, gt_classes)
        losses = {"loss_cls": loss_cls,
                  "loss_box_reg": self.box_reg_loss(proposal_boxes, gt_boxes, proposal_deltas, gt_classes)}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    def box_reg_loss(self, proposal_boxes, gt_boxes, pred_deltas, gt_classes):
        """Detectron2 box_reg_loss (smooth-L1 over the foreground rows, summed, / max(R, 1)): one kernel launch for the loss and
        its gradient — no ``nonzero`` (a host synchronisation per step in the stock implementation), no chain of small launches."""
        box_dim = proposal_boxes.shape[1]
        if self.box_reg_loss_type != "smooth_l1":
            raise NotImplementedError(f"box_reg_loss_type {self.box_reg_loss_type!r} (shipped configs use smooth_l1)")
        assert pred_deltas.shape[1] == box_dim == 4, "class-agnostic box regression only (box_emb_head.py:137)"
        return LF.box_reg_loss(pred_deltas, proposal_boxes, gt_boxes, gt_classes, self.num_classes, self.box2box_transform.weights,
                               self.smooth_l1_beta)

    def predict_probs(self, predictions, proposals):
        scores, _ = predictions
        return self._fused_aux(scores, want_probs=True).probs.split([len(p) for p in proposals], dim=0)

    def predict_classes(self, predictions):
        """Per-RoI argmax over the K foreground columns (int64), from the fused epilogue."""
        scores, _ = predictions
        return self._fused_aux(scores).argmax_fg

    def predict_boxes(self, predictions, proposals):
        if not len(proposals):
            return []
        _, proposal_deltas = predictions
        proposal_boxes = _cat([_box_tensor(p.proposal_boxes) for p in proposals], 0)
        boxes = self.box2box_transform.apply_deltas(proposal_deltas, proposal_boxes)
        return boxes.split([len(p) for p in proposals])

    def inference(self, predictions, proposals):
        """Detectron2 FastRCNNOutputLayers.inference -> fast_rcnn_inference.  On the device the whole tail (box decoding, score
        threshold, per-class NMS, top-k; all images) is four launches of loco_box_inference and ONE device->host copy of the per-image
        detection counts; the per-image torchvision path below remains for CPU tensors and for settings outside the kernel's range
        (no top-k limit, more than 2048 proposals in an image)."""
        scores_all, deltas_all = predictions
        rows = [len(p) for p in proposals]
        if (scores_all.is_cuda and len(proposals) and 1 <= self.test_topk_per_image <= 1024 and max(rows) <= 2048
                and scores_all.shape[0] > 0 and max(rows) * self.num_classes < 2 ** 32):
            probs = self._fused_aux(scores_all, want_probs=True).probs
            proposal_boxes = _cat([_box_tensor(p.proposal_boxes) for p in proposals], 0)
            b, s, c, r, n = ops.box_inference(probs, deltas_all.detach(), proposal_boxes, rows, [x.image_size for x in proposals],
                                              self.box2box_transform.weights, self.box2box_transform.scale_clamp, self.test_score_thresh,
                                              self.test_nms_thresh, self.test_topk_per_image)
            results, kept = [], []
            for i, (cnt, p) in enumerate(zip(n.tolist(), proposals)):
                inst = Instances(p.image_size)
                inst.pred_boxes = Boxes(b[i, :cnt])
                inst.scores = s[i, :cnt]
                inst.pred_classes = c[i, :cnt]
                results.append(inst)
                kept.append(r[i, :cnt])
            return results, kept
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
    return BOX_PREDICTORS[name](cfg, input_shape)
