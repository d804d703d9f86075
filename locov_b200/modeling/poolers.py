"""ROIAlign / ROIPooler mirrors of the Detectron2 classes the reference instantiates at
roi_emb_heads.py:182-187 and calls at :243-245 (``self.pooler(features, boxes)``).

Single-level pooling only, exactly as the reference asserts (``assert len(in_features) == 1``,
roi_emb_heads.py:180).  ``pooler_type`` "ROIAlignV2" = aligned=True, "ROIAlign" = aligned=False.
"""
from typing import List

import torch
from torch import nn

from .. import functional as LF
from .._lib import LocoError


class ROIAlign(nn.Module):
    """``channels_last`` / ``out_dtype`` are this library's additions: the pooled tensor keeps its logical
    [R,C,PH,PW] shape but is laid out ``torch.channels_last`` (fp32 or bf16), which is what cuDNN's tensor-core convolutions of
    the res5 stage consume without a transpose; the defaults are the reference's NCHW fp32."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True, channels_last=False, out_dtype=torch.float32):
        super().__init__()
        self.channels_last = bool(channels_last)
        self.out_dtype = out_dtype
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.aligned = bool(aligned)

    def forward(self, input, rois):
        assert rois.dim() == 2 and rois.size(1) == 5
        return LF.roi_align(input, rois.to(dtype=input.dtype), self.output_size, self.spatial_scale,
                            self.sampling_ratio, self.aligned, self.channels_last, self.out_dtype)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, aligned={self.aligned})")


def convert_boxes_to_pooler_format(box_lists: List):
    """List[Boxes] (or List[Tensor [Ri,4]]) -> [R,5] rows (batch_index, x1, y1, x2, y2)."""
    tensors = [b.tensor if hasattr(b, "tensor") else b for b in box_lists]
    sizes = torch.tensor([t.shape[0] for t in tensors], device=tensors[0].device if tensors else None)
    boxes = torch.cat(tensors, 0)
    idx = torch.repeat_interleave(torch.arange(len(tensors), dtype=boxes.dtype, device=boxes.device), sizes)
    return torch.cat([idx[:, None], boxes], 1)


class ROIPooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4,
                 channels_last=False, out_dtype=torch.float32):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2
        self.output_size = tuple(output_size)
        if len(scales) != 1:
            raise LocoError("ROIPooler: the C4 region-text path pools a single feature level (roi_emb_heads.py:180)")
        if pooler_type == "ROIAlignV2":
            aligned = True
        elif pooler_type == "ROIAlign":
            aligned = False
        else:
            raise NotImplementedError(f"pooler_type {pooler_type!r} is not part of the LocOV hot path")
        self.channels_last = bool(channels_last)
        self.out_dtype = out_dtype
        self.level_poolers = nn.ModuleList([ROIAlign(output_size, scales[0], sampling_ratio, aligned, channels_last, out_dtype)])

    def forward(self, x: List[torch.Tensor], box_lists: List):
        assert isinstance(x, list) and isinstance(box_lists, list) and len(x) == 1
        assert len(box_lists) == x[0].size(0), "unequal value, x[0] batch dim 0 is {}, but box_list has length {}".format(
            x[0].size(0), len(box_lists))
        if len(box_lists) == 0:
            return x[0].new_zeros((0, x[0].shape[1]) + self.output_size, dtype=self.out_dtype)
        rois = convert_boxes_to_pooler_format(box_lists)
        return self.level_poolers[0](x[0], rois)
