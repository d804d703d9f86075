"""ORACLE (test infrastructure only): CPU restatement of the Detectron2 pieces the reference's ROI heads
inherit or call.  Detectron2 is a third-party dependency of lmb-freiburg/locov that is NOT vendored in
/root/reference and NOT pinned (README.md:28 "follow Detectron2 installation instructions"; torch 1.10 /
CUDA 11.2 in README.md:30 => the v0.6 release line).  Nothing here is shipped: these classes exist so that
the reference's OWN source files

    /root/reference/ovr/modeling/roi_heads/box_emb_head.py   (EmbeddingFastRCNNOutputLayers, :60-236)
    /root/reference/ovr/modeling/roi_heads/roi_emb_heads.py  (Embedding[Proposals]Res5ROIHeads, :121-360)

can be imported UNMODIFIED in the build container (oracle/ref_loader.py) and executed to generate the
golden vectors under tests/golden/ (tests/golden/make_golden.py).  Each class restates the published
v0.6 behaviour of the symbol the reference imports (call sites: box_emb_head.py:9-19,60,108-119;
roi_emb_heads.py:8-18,182-187,221-241):

  detectron2.config.configurable                              -> configurable
  detectron2.layers.{ShapeSpec,cat,cross_entropy,nonzero_tuple,batched_nms}
  detectron2.modeling.box_regression.Box2BoxTransform          -> Box2BoxTransform
  detectron2.modeling.roi_heads.fast_rcnn.FastRCNNOutputLayers -> FastRCNNOutputLayers (.losses/.inference)
  detectron2.modeling.roi_heads.fast_rcnn.fast_rcnn_inference[_single_image]
  detectron2.modeling.roi_heads.ROIHeads / ROI_HEADS_REGISTRY
  detectron2.modeling.poolers.ROIPooler                        -> ROIPooler (single level -> torchvision roi_align)
  detectron2.modeling.backbone.resnet.{BottleneckBlock,ResNet.make_stage}, FrozenBatchNorm2d
  detectron2.structures.{Boxes,Instances}
The arithmetic third-party op below Detectron2 is torchvision's compiled ``roi_align`` (present in this image).
"""
import functools
import inspect
import math
from collections import namedtuple

import torch
from torch import nn
from torch.nn import functional as F

ShapeSpec = namedtuple("ShapeSpec", ["channels", "height", "width", "stride"], defaults=[None, None, None, None])


# ---- detectron2.config.configurable ---------------------------------------------------------------------
def _is_cfg(x):
    return hasattr(x, "MODEL") and not isinstance(x, (torch.Tensor, nn.Module))


def _called_with_cfg(*args, **kwargs):
    if len(args) and _is_cfg(args[0]):
        return True
    return _is_cfg(kwargs.pop("cfg", None))


def _get_args_from_config(from_config_func, *args, **kwargs):
    sig = inspect.signature(from_config_func)
    assert list(sig.parameters)[0] == "cfg", "from_config must take 'cfg' as its first argument"
    var = any(p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in sig.parameters.values())
    if var:
        return from_config_func(*args, **kwargs)
    supported = set(sig.parameters)
    extra = {k: kwargs.pop(k) for k in list(kwargs) if k not in supported}
    ret = from_config_func(*args, **kwargs)
    ret.update(extra)
    return ret


def configurable(init_func=None, *, from_config=None):
    assert init_func is not None and from_config is None, "only the @configurable __init__ form is used by the reference"

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        fc = type(self).from_config
        if _called_with_cfg(*args, **kwargs):
            init_func(self, **_get_args_from_config(fc, *args, **kwargs))
        else:
            init_func(self, *args, **kwargs)

    return wrapped


# ---- detectron2.layers ----------------------------------------------------------------------------------
def cat(tensors, dim=0):
    assert isinstance(tensors, (list, tuple))
    return tensors[0] if len(tensors) == 1 else torch.cat(tensors, dim)


def cross_entropy(input, target, *, reduction="mean", **kwargs):
    if target.numel() == 0 and reduction == "mean":
        return input.sum() * 0.0
    return F.cross_entropy(input, target, reduction=reduction, **kwargs)


def nonzero_tuple(x):
    return x.nonzero(as_tuple=True) if x.dim() else x.unsqueeze(0).nonzero().unbind(1)


def batched_nms(boxes, scores, idxs, iou_threshold):
    from torchvision.ops import boxes as box_ops
    assert boxes.shape[-1] == 4
    return box_ops.batched_nms(boxes.float(), scores, idxs, iou_threshold)


class FrozenBatchNorm2d(nn.Module):
    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = num_features, eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def forward(self, x):
        scale = self.weight * (self.running_var + self.eps).rsqrt()
        bias = self.bias - self.running_mean * scale
        return x * scale.reshape(1, -1, 1, 1).to(x.dtype) + bias.reshape(1, -1, 1, 1).to(x.dtype)


def get_norm(norm, out_channels):
    if norm is None or norm == "":
        return None
    assert norm == "FrozenBN", "the shipped configs use FrozenBN (Detectron2 default RESNETS.NORM)"
    return FrozenBatchNorm2d(out_channels)


class Conv2d(nn.Conv2d):
    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


class BottleneckBlock(nn.Module):
    def __init__(self, in_channels, out_channels, *, bottleneck_channels, stride=1, num_groups=1, norm="BN",
                 stride_in_1x1=False, dilation=1):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        self.shortcut = None
        if in_channels != out_channels:
            self.shortcut = Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False, norm=get_norm(norm, out_channels))
        s1, s3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = Conv2d(in_channels, bottleneck_channels, kernel_size=1, stride=s1, bias=False, norm=get_norm(norm, bottleneck_channels))
        self.conv2 = Conv2d(bottleneck_channels, bottleneck_channels, kernel_size=3, stride=s3, padding=1 * dilation, bias=False,
                            groups=num_groups, dilation=dilation, norm=get_norm(norm, bottleneck_channels))
        self.conv3 = Conv2d(bottleneck_channels, out_channels, kernel_size=1, bias=False, norm=get_norm(norm, out_channels))
        for layer in [self.conv1, self.conv2, self.conv3, self.shortcut]:
            if layer is not None:
                nn.init.kaiming_normal_(layer.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        out = F.relu_(self.conv1(x))
        out = F.relu_(self.conv2(out))
        out = self.conv3(out)
        shortcut = self.shortcut(x) if self.shortcut is not None else x
        out = out + shortcut
        return F.relu_(out)


class ResNet:
    @staticmethod
    def make_stage(block_class, num_blocks, *, in_channels, out_channels, **kwargs):
        blocks = []
        for i in range(num_blocks):
            cur = {}
            for k, v in kwargs.items():
                if k.endswith("_per_block"):
                    assert len(v) == num_blocks
                    cur[k[: -len("_per_block")]] = v[i]
                else:
                    cur[k] = v
            blocks.append(block_class(in_channels=in_channels, out_channels=out_channels, **cur))
            in_channels = out_channels
        return blocks


# ---- detectron2.structures ---------------------------------------------------------------------------------
class Boxes:
    def __init__(self, tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        else:
            tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self):
        return Boxes(self.tensor.clone())

    def to(self, device):
        return Boxes(self.tensor.to(device=device))

    def clip(self, box_size):
        assert torch.isfinite(self.tensor).all(), "Box tensor contains infinite or NaN!"
        h, w = box_size
        x1 = self.tensor[:, 0].clamp(min=0, max=w)
        y1 = self.tensor[:, 1].clamp(min=0, max=h)
        x2 = self.tensor[:, 2].clamp(min=0, max=w)
        y2 = self.tensor[:, 3].clamp(min=0, max=h)
        self.tensor = torch.stack((x1, y1, x2, y2), dim=-1)

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        assert b.dim() == 2
        return Boxes(b)

    def __len__(self):
        return self.tensor.shape[0]

    @property
    def device(self):
        return self.tensor.device


class Instances:
    def __init__(self, image_size, **kwargs):
        self._image_size = image_size
        self._fields = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return self._fields[name]

    def set(self, name, value):
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def __len__(self):
        for v in self._fields.values():
            return v.__len__()
        raise NotImplementedError("Empty Instances does not support __len__!")

    def __getitem__(self, item):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret


# ---- detectron2.modeling.box_regression ------------------------------------------------------------------
_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)


class Box2BoxTransform:
    def __init__(self, weights, scale_clamp=_DEFAULT_SCALE_CLAMP):
        self.weights = weights
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        src_widths = src_boxes[:, 2] - src_boxes[:, 0]
        src_heights = src_boxes[:, 3] - src_boxes[:, 1]
        src_ctr_x = src_boxes[:, 0] + 0.5 * src_widths
        src_ctr_y = src_boxes[:, 1] + 0.5 * src_heights
        target_widths = target_boxes[:, 2] - target_boxes[:, 0]
        target_heights = target_boxes[:, 3] - target_boxes[:, 1]
        target_ctr_x = target_boxes[:, 0] + 0.5 * target_widths
        target_ctr_y = target_boxes[:, 1] + 0.5 * target_heights
        wx, wy, ww, wh = self.weights
        dx = wx * (target_ctr_x - src_ctr_x) / src_widths
        dy = wy * (target_ctr_y - src_ctr_y) / src_heights
        dw = ww * torch.log(target_widths / src_widths)
        dh = wh * torch.log(target_heights / src_heights)
        deltas = torch.stack((dx, dy, dw, dh), dim=1)
        assert (src_widths > 0).all().item(), "Input boxes to Box2BoxTransform are not valid!"
        return deltas

    def apply_deltas(self, deltas, boxes):
        deltas = deltas.float()
        boxes = boxes.to(deltas.dtype)
        widths = boxes[:, 2] - boxes[:, 0]
        heights = boxes[:, 3] - boxes[:, 1]
        ctr_x = boxes[:, 0] + 0.5 * widths
        ctr_y = boxes[:, 1] + 0.5 * heights
        wx, wy, ww, wh = self.weights
        dx = deltas[:, 0::4] / wx
        dy = deltas[:, 1::4] / wy
        dw = deltas[:, 2::4] / ww
        dh = deltas[:, 3::4] / wh
        dw = torch.clamp(dw, max=self.scale_clamp)
        dh = torch.clamp(dh, max=self.scale_clamp)
        pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
        pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
        pred_w = torch.exp(dw) * widths[:, None]
        pred_h = torch.exp(dh) * heights[:, None]
        x1 = pred_ctr_x - 0.5 * pred_w
        y1 = pred_ctr_y - 0.5 * pred_h
        x2 = pred_ctr_x + 0.5 * pred_w
        y2 = pred_ctr_y + 0.5 * pred_h
        pred_boxes = torch.stack((x1, y1, x2, y2), dim=-1)
        return pred_boxes.reshape(deltas.shape)


# ---- fvcore.nn ---------------------------------------------------------------------------------------------
def smooth_l1_loss(input, target, beta, reduction="none"):
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        cond = n < beta
        loss = torch.where(cond, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    elif reduction == "sum":
        loss = loss.sum()
    return loss


def giou_loss(*a, **k):
    raise NotImplementedError("giou is not used by the shipped configs (BBOX_REG_LOSS_TYPE = smooth_l1)")


# ---- detectron2.modeling.roi_heads.fast_rcnn --------------------------------------------------------------
def _log_classification_stats(pred_logits, gt_classes, prefix="fast_rcnn"):
    return      # event-storage logging only; no arithmetic on the path


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image):
    valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid_mask.all():
        boxes = boxes[valid_mask]
        scores = scores[valid_mask]
    scores = scores[:, :-1]
    num_bbox_reg_classes = boxes.shape[1] // 4
    boxes = Boxes(boxes.reshape(-1, 4))
    boxes.clip(image_shape)
    boxes = boxes.tensor.view(-1, num_bbox_reg_classes, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    if num_bbox_reg_classes == 1:
        boxes = boxes[filter_inds[:, 0], 0]
    else:
        boxes = boxes[filter_mask]
    scores = scores[filter_mask]
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    boxes, scores, filter_inds = boxes[keep], scores[keep], filter_inds[keep]
    result = Instances(image_shape)
    result.pred_boxes = Boxes(boxes)
    result.scores = scores
    result.pred_classes = filter_inds[:, 1]
    return result, filter_inds[:, 0]


def fast_rcnn_inference(boxes, scores, image_shapes, score_thresh, nms_thresh, topk_per_image):
    result_per_image = [
        fast_rcnn_inference_single_image(b, s, shp, score_thresh, nms_thresh, topk_per_image)
        for s, b, shp in zip(scores, boxes, image_shapes)
    ]
    return [x[0] for x in result_per_image], [x[1] for x in result_per_image]


class FastRCNNOutputLayers(nn.Module):
    @configurable
    def __init__(self, input_shape, *, box2box_transform, num_classes, test_score_thresh=0.0, test_nms_thresh=0.5,
                 test_topk_per_image=100, cls_agnostic_bbox_reg=False, smooth_l1_beta=0.0, box_reg_loss_type="smooth_l1",
                 loss_weight=1.0):
        super().__init__()
        if isinstance(input_shape, int):
            input_shape = ShapeSpec(channels=input_shape)
        self.num_classes = num_classes
        input_size = input_shape.channels * (input_shape.width or 1) * (input_shape.height or 1)
        self.cls_score = nn.Linear(input_size, num_classes + 1)
        num_bbox_reg_classes = 1 if cls_agnostic_bbox_reg else num_classes
        box_dim = len(box2box_transform.weights)
        self.bbox_pred = nn.Linear(input_size, num_bbox_reg_classes * box_dim)
        nn.init.normal_(self.cls_score.weight, std=0.01)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        for l in [self.cls_score, self.bbox_pred]:
            nn.init.constant_(l.bias, 0)
        self.box2box_transform = box2box_transform
        self.smooth_l1_beta = smooth_l1_beta
        self.test_score_thresh = test_score_thresh
        self.test_nms_thresh = test_nms_thresh
        self.test_topk_per_image = test_topk_per_image
        self.box_reg_loss_type = box_reg_loss_type
        if isinstance(loss_weight, float):
            loss_weight = {"loss_cls": loss_weight, "loss_box_reg": loss_weight}
        self.loss_weight = loss_weight

    def losses(self, predictions, proposals):
        scores, proposal_deltas = predictions
        gt_classes = cat([p.gt_classes for p in proposals], dim=0) if len(proposals) else torch.empty(0)
        _log_classification_stats(scores, gt_classes)
        if len(proposals):
            proposal_boxes = cat([p.proposal_boxes.tensor for p in proposals], dim=0)
            assert not proposal_boxes.requires_grad, "Proposals should not require gradients!"
            gt_boxes = cat([(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes).tensor for p in proposals], dim=0)
        else:
            proposal_boxes = gt_boxes = torch.empty((0, 4), device=proposal_deltas.device)
        losses = {
            "loss_cls": cross_entropy(scores, gt_classes, reduction="mean"),
            "loss_box_reg": self.box_reg_loss(proposal_boxes, gt_boxes, proposal_deltas, gt_classes),
        }
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    def box_reg_loss(self, proposal_boxes, gt_boxes, pred_deltas, gt_classes):
        box_dim = proposal_boxes.shape[1]
        fg_inds = nonzero_tuple((gt_classes >= 0) & (gt_classes < self.num_classes))[0]
        if pred_deltas.shape[1] == box_dim:
            fg_pred_deltas = pred_deltas[fg_inds]
        else:
            fg_pred_deltas = pred_deltas.view(-1, self.num_classes, box_dim)[fg_inds, gt_classes[fg_inds]]
        assert self.box_reg_loss_type == "smooth_l1"
        gt_pred_deltas = self.box2box_transform.get_deltas(proposal_boxes[fg_inds], gt_boxes[fg_inds])
        loss_box_reg = smooth_l1_loss(fg_pred_deltas, gt_pred_deltas, self.smooth_l1_beta, reduction="sum")
        return loss_box_reg / max(gt_classes.numel(), 1.0)

    def inference(self, predictions, proposals):
        boxes = self.predict_boxes(predictions, proposals)
        scores = self.predict_probs(predictions, proposals)
        image_shapes = [x.image_size for x in proposals]
        return fast_rcnn_inference(boxes, scores, image_shapes, self.test_score_thresh, self.test_nms_thresh, self.test_topk_per_image)

    def predict_boxes(self, predictions, proposals):
        if not len(proposals):
            return []
        _, proposal_deltas = predictions
        num_prop_per_image = [len(p) for p in proposals]
        proposal_boxes = cat([p.proposal_boxes.tensor for p in proposals], dim=0)
        predict_boxes = self.box2box_transform.apply_deltas(proposal_deltas, proposal_boxes)
        return predict_boxes.split(num_prop_per_image)

    def predict_probs(self, predictions, proposals):
        scores, _ = predictions
        num_inst_per_image = [len(p) for p in proposals]
        probs = F.softmax(scores, dim=-1)
        return probs.split(num_inst_per_image, dim=0)


# ---- detectron2.modeling.poolers ---------------------------------------------------------------------------
def convert_boxes_to_pooler_format(box_lists):
    boxes = torch.cat([x.tensor for x in box_lists], dim=0)
    sizes = torch.tensor([x.__len__() for x in box_lists], device=boxes.device)
    indices = torch.repeat_interleave(torch.arange(len(box_lists), dtype=boxes.dtype, device=boxes.device), sizes)
    return cat([indices[:, None], boxes], dim=1)


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        super().__init__()
        self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned = output_size, spatial_scale, sampling_ratio, aligned

    def forward(self, input, rois):
        from torchvision.ops import roi_align
        assert rois.dim() == 2 and rois.size(1) == 5
        return roi_align(input, rois.to(dtype=input.dtype), self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)


class ROIPooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        self.output_size = output_size
        assert len(scales) == 1, "the C4 heads of the reference pool one level (roi_emb_heads.py:180)"
        assert pooler_type in ("ROIAlign", "ROIAlignV2")
        self.level_poolers = nn.ModuleList(ROIAlign(output_size, spatial_scale=s, sampling_ratio=sampling_ratio, aligned=pooler_type == "ROIAlignV2")
                                           for s in scales)

    def forward(self, x, box_lists):
        assert isinstance(x, list) and isinstance(box_lists, list) and len(x) == 1
        assert len(box_lists) == x[0].size(0)
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        return self.level_poolers[0](x[0], convert_boxes_to_pooler_format(box_lists))


# ---- detectron2.modeling.roi_heads --------------------------------------------------------------------------
class _Registry(dict):
    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self[o.__name__] = o
                return o
            return deco
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        return self[name]


ROI_HEADS_REGISTRY = _Registry("ROI_HEADS")


class ROIHeads(nn.Module):
    """Base class: only the constructor plumbing the reference relies on (the proposal matcher / sampler are not
    on the region-text path; goldens are generated with proposals that already carry gt_classes)."""

    @configurable
    def __init__(self, *, num_classes, batch_size_per_image, positive_fraction, proposal_matcher=None, proposal_append_gt=True):
        super().__init__()
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.num_classes = num_classes
        self.proposal_matcher = proposal_matcher
        self.proposal_append_gt = proposal_append_gt

    @classmethod
    def from_config(cls, cfg):
        return {
            "batch_size_per_image": cfg.MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE,
            "positive_fraction": cfg.MODEL.ROI_HEADS.POSITIVE_FRACTION,
            "num_classes": cfg.MODEL.ROI_HEADS.NUM_CLASSES,
            "proposal_append_gt": cfg.MODEL.ROI_HEADS.PROPOSAL_APPEND_GT,
            "proposal_matcher": None,
        }
