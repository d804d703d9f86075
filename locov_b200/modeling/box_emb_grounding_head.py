"""Drop-in mirrors of ``GroundingModule`` and ``EmbeddingGroundingFastRCNNOutputLayers``
(reference ovr/modeling/roi_heads/box_emb_grounding_head.py:60-434): class scores from MULTI-TOKEN class embeddings — every class is
a list of token embeddings, a RoI's score for the class is the attention-pooled RoI x token similarity.

Unreachable from the reference's shipped configs (``MODEL.ROI_HEADS.MAX_TOKENS`` is read by ``from_config`` :352 but never defined by
ovr/config/config.py); here the key is optional (it only sizes a placeholder in the reference too).  Same constructor keywords,
``(cfg, input_shape)`` construction, parameter names (``emb_pred``, ``bbox_pred``, ``cls_score.token_score``) and return
structures.  The token GEMM runs on the tcgen05 core, the per-class pooling in ``loco_token_pool_fwd/bwd``; losses / inference are
inherited from the drop-in ``EmbeddingFastRCNNOutputLayers`` (fused softmax statistics, ``loco_box_inference``).
"""
from typing import Dict, Union

import torch
from torch import nn

from .. import functional as LF
from .. import ops
from .box_emb_head import Box2BoxTransform, EmbeddingFastRCNNOutputLayers
from .configurable import configurable
from .logged_module import LoggedModule, normalize_vec
from .registry import BOX_PREDICTORS
from .structures import ShapeSpec


class GroundingModule(LoggedModule):
    """box_emb_grounding_head.py:60-277.  ``forward(image_emb [R, D]) -> (scores [R, K(+1)], tok_attention)``.

    ``tok_attention`` ([R, K(+1), max_tok], zero beyond a class's tokens) is only materialised when ``return_attention`` is set —
    the box head discards it (:424) and at LVIS size it is 8 x the score matrix; otherwise the second element is None."""

    def __init__(self, emb_dim, num_classes, max_tokens=None, local_metric: str = "dot", global_metric: str = "aligned_local",
                 alignment: str = "softmax", temperature: float = 1.0, normalize_emb: bool = False, background_class: bool = True,
                 return_similarity: bool = False, precision: str = "fp32", return_attention: bool = False):
        super().__init__()
        if local_metric != "dot":
            raise NotImplementedError(f"LOCAL_METRIC {local_metric!r}: only 'dot' (the shipped value) runs on the B200 path")
        if global_metric != "aligned_local":
            raise NotImplementedError(f"GLOBAL_METRIC {global_metric!r} (box_emb_grounding_head.py:186: the reference raises too)")
        if alignment not in ("softmax", "hardmax"):
            raise NotImplementedError(f"ALIGNMENT {alignment!r}")
        if return_similarity:
            raise NotImplementedError("return_similarity: the per-class token similarity lists are never materialised on the B200 path")
        self.emb_dim, self.num_classes, self.max_tokens = emb_dim, num_classes, max_tokens
        self.local_metric, self.global_metric, self.alignment = local_metric, global_metric, alignment
        self.temperature = float(temperature)
        self.normalize_emb = normalize_emb
        self.background_class = background_class
        self.precision = precision
        self.return_attention = return_attention
        self.token_score = None            # nn.Linear(emb_dim, total tokens), built by set_class_embeddings (:243)
        self.num_tok = None
        self.seg_off = None

    def set_class_embeddings(self, embs, device):
        """embs: {class index: [n_tok, D] tensor / array}.  box_emb_grounding_head.py:223-262."""
        self.num_classes = len(embs)
        rows, off = [], [0]
        for cls_idx in range(self.num_classes):
            e = embs[cls_idx]
            e = e.clone().detach().to(device=device, dtype=torch.float32) if torch.is_tensor(e) else torch.tensor(e, device=device, dtype=torch.float32)
            e = e.reshape(-1, self.emb_dim)
            rows.append(e)
            off.append(off[-1] + e.shape[0])
        if self.background_class:          # one zero token: score exactly 0 (:239-241; num_tok 0 -> 1 by the in-place patch of :124-125)
            rows.append(torch.zeros(1, self.emb_dim, device=device))
            off.append(off[-1] + 1)
        class_emb = torch.cat(rows, 0)
        if self.normalize_emb:
            assert class_emb.shape[1] == self.emb_dim, "The embedding dimension has to match the one saved in the model"
            class_emb = LF.normalize_rows(class_emb) if class_emb.is_cuda else normalize_vec(class_emb, dim=1)
        self.token_score = nn.Linear(self.emb_dim, class_emb.shape[0]).to(device)
        self.token_score.weight.data = class_emb
        self.token_score.bias.data = torch.zeros_like(self.token_score.bias.data)
        self.token_score.weight.requires_grad = False
        self.token_score.bias.requires_grad = False
        self.seg_off = torch.tensor(off, dtype=torch.int32, device=device)
        self.num_tok = (self.seg_off[1:] - self.seg_off[:-1]).to(torch.int32)

    def forward(self, image_emb):
        if self.token_score is None:
            raise RuntimeError("set_class_embeddings() must be called before forward")
        self.log("image_emb", image_emb)
        scores = LF.grounding_scores(image_emb, self.token_score.weight, self.seg_off, self.temperature, self.alignment, self.precision)
        att = None
        if self.return_attention:
            with torch.no_grad():
                raw = LF.linear(image_emb.detach(), self.token_score.weight, None, self.precision).contiguous()
                _, flat = ops.token_pool(raw, self.seg_off, 1.0 / self.temperature, self.alignment == "hardmax", want_attention=True)
                k1, mt = self.seg_off.numel() - 1, int(self.num_tok.max())
                att = torch.zeros((raw.shape[0], k1, mt), dtype=torch.float32, device=raw.device)
                tok_class = torch.repeat_interleave(torch.arange(k1, device=raw.device), self.num_tok.long())
                tok_pos = torch.arange(raw.shape[1], device=raw.device) - self.seg_off[:-1].long()[tok_class]
                att[:, tok_class, tok_pos] = flat
                if self.background_class:
                    att[:, -1, :] = 0.0      # the reference's mask_emb row of the background class is all zero (:232-236)
        self.log("global_dist", scores)
        return scores, att


@BOX_PREDICTORS.register()
class EmbeddingGroundingFastRCNNOutputLayers(EmbeddingFastRCNNOutputLayers):
    """box_emb_grounding_head.py:280-434."""

    @configurable
    def __init__(self, input_shape, *, box2box_transform=None, num_classes: int = 80, test_score_thresh: float = 0.0,
                 test_nms_thresh: float = 0.5, test_topk_per_image: int = 100, cls_agnostic_bbox_reg: bool = False,
                 smooth_l1_beta: float = 0.0, box_reg_loss_type: str = "smooth_l1", loss_weight: Union[float, Dict[str, float]] = 1.0,
                 emb_dim: int = 768, embedding_based: bool = True, freeze_emb_pred: bool = True, normalize_emb: bool = False,
                 detach_cls_predictor: bool = False, grounding_module: GroundingModule = None, precision: str = "fp32"):
        # (the reference accepts freeze_emb_pred here but never applies it, :334-345: emb_pred stays trainable)
        super().__init__(input_shape, box2box_transform=box2box_transform, num_classes=num_classes, test_score_thresh=test_score_thresh,
                         test_nms_thresh=test_nms_thresh, test_topk_per_image=test_topk_per_image, cls_agnostic_bbox_reg=cls_agnostic_bbox_reg,
                         smooth_l1_beta=smooth_l1_beta, box_reg_loss_type=box_reg_loss_type, loss_weight=loss_weight, emb_dim=emb_dim,
                         embedding_based=embedding_based, freeze_emb_pred=False, normalize_emb=normalize_emb, standardize_emb=False,
                         detach_cls_predictor=detach_cls_predictor, precision=precision)
        if grounding_module is None:
            raise ValueError("grounding_module is required (box_emb_grounding_head.py:341)")
        grounding_module.precision = precision
        self.cls_score = grounding_module

    @classmethod
    def from_config(cls, cfg, input_shape):
        ret = super().from_config(cfg, input_shape)
        ret.pop("standardize_emb")
        g = cfg.MODEL.MMSS_HEAD.GROUNDING
        ret["grounding_module"] = GroundingModule(
            cfg.MODEL.ROI_BOX_HEAD.EMB_DIM, cfg.MODEL.ROI_HEADS.NUM_CLASSES, getattr(cfg.MODEL.ROI_HEADS, "MAX_TOKENS", None),
            local_metric=g.LOCAL_METRIC, global_metric=g.GLOBAL_METRIC, alignment=g.ALIGNMENT, temperature=g.ALIGNMENT_TEMPERATURE,
            normalize_emb=cfg.MODEL.ROI_BOX_HEAD.NORMALIZE_EMB_PRED, precision=ret["precision"])
        return ret

    def device(self):
        return self.emb_pred.weight.device

    def forward(self, x):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        x = x.to(torch.float32).contiguous()
        proposal_deltas = LF.linear(x, self.bbox_pred.weight, self.bbox_pred.bias, self.precision)
        if self.detach_cls_predictor:
            with torch.no_grad():
                scores = self.forward_cls_prediction(x.detach())
        else:
            scores = self.forward_cls_prediction(x)
        return scores, proposal_deltas

    def forward_cls_prediction(self, x):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        e = LF.linear(x.to(torch.float32).contiguous(), self.emb_pred.weight, self.emb_pred.bias, self.precision)
        if self.normalize_emb:
            e = LF.normalize_rows(e)
        scores, _ = self.cls_score(e)
        self._aux = None                     # softmax statistics are computed from `scores` on demand (loco_box_softmax)
        return scores

    def set_class_embeddings(self, embs):
        self.cls_score.set_class_embeddings(embs, self.device())
        self.num_classes = self.cls_score.num_classes
        self._aux = None
