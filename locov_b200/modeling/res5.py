"""The ``res5`` stage between the two halves of the hot path (roi_emb_heads.py:216-245: three bottleneck blocks,
first one stride 2, on the pooled ``[R,1024,14,14]`` RoI features -> ``[R,2048,7,7]``).

The convolutions themselves stay cuDNN calls ("emit channels_last from RoIAlign straight into
cuDNN"); this module only provides the stage with Detectron2's parameter / buffer names (``res5.{i}.conv{1,2,3}.weight``,
``...norm.{weight,bias,running_mean,running_var}``, ``res5.0.shortcut.*``) so that reference checkpoints load with
``strict=True``, and is what ``EmbeddingRes5ROIHeads._build_res5_block`` returns when Detectron2 itself is not
importable.  FrozenBN only (Detectron2's default ``RESNETS.NORM``, not overridden by the shipped configs).
"""
import torch
from torch import nn
from torch.nn import functional as F


class FrozenBatchNorm2d(nn.Module):
    """Per-channel affine with fixed statistics: y = x * w / sqrt(var + eps) + (b - mean * w / sqrt(var + eps))."""
    _version = 3

    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = num_features, eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def scale_shift(self):
        scale = self.weight * (self.running_var + self.eps).rsqrt()
        return scale, self.bias - self.running_mean * scale

    def forward(self, x):
        scale, shift = self.scale_shift()
        return x * scale.reshape(1, -1, 1, 1).to(x.dtype) + shift.reshape(1, -1, 1, 1).to(x.dtype)


class NormConv2d(nn.Conv2d):
    """``nn.Conv2d`` with an attached norm sub-module named ``norm`` (Detectron2's ``layers.Conv2d`` key layout)."""

    def __init__(self, *args, norm=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm = norm

    def forward(self, x):
        x = F.conv2d(x, self.weight.to(x.dtype), None if self.bias is None else self.bias.to(x.dtype), self.stride, self.padding,
                     self.dilation, self.groups)
        return x if self.norm is None else self.norm(x)


def _norm(norm, channels):
    if norm in (None, ""):
        return None
    if norm != "FrozenBN":
        raise NotImplementedError(f"RESNETS.NORM {norm!r}: the shipped configs use FrozenBN; inject Detectron2's res5 for other norms")
    return FrozenBatchNorm2d(channels)


class BottleneckBlock(nn.Module):
    def __init__(self, in_channels, out_channels, *, bottleneck_channels, stride=1, num_groups=1, norm="FrozenBN", stride_in_1x1=True):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        self.shortcut = None
        if in_channels != out_channels:
            self.shortcut = NormConv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False, norm=_norm(norm, out_channels))
        s1, s3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = NormConv2d(in_channels, bottleneck_channels, kernel_size=1, stride=s1, bias=False, norm=_norm(norm, bottleneck_channels))
        self.conv2 = NormConv2d(bottleneck_channels, bottleneck_channels, kernel_size=3, stride=s3, padding=1, bias=False, groups=num_groups,
                                norm=_norm(norm, bottleneck_channels))
        self.conv3 = NormConv2d(bottleneck_channels, out_channels, kernel_size=1, bias=False, norm=_norm(norm, out_channels))
        for layer in (self.conv1, self.conv2, self.conv3, self.shortcut):
            if layer is not None:
                nn.init.kaiming_normal_(layer.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        out = F.relu_(self.conv1(x))
        out = F.relu_(self.conv2(out))
        out = self.conv3(out)
        out = out + (self.shortcut(x) if self.shortcut is not None else x)
        return F.relu_(out)


def build_res5_block(cfg):
    """roi_emb_heads.py:216-241 (``_build_res5_block``) -> (nn.Sequential of 3 blocks, out_channels).  Uses Detectron2's own
    ``ResNet.make_stage`` when it is importable (bit-identical modules for the caller), the blocks above otherwise."""
    stage_channel_factor = 2 ** 3
    num_groups = cfg.MODEL.RESNETS.NUM_GROUPS
    width_per_group = cfg.MODEL.RESNETS.WIDTH_PER_GROUP
    bottleneck_channels = num_groups * width_per_group * stage_channel_factor
    out_channels = cfg.MODEL.RESNETS.RES2_OUT_CHANNELS * stage_channel_factor
    stride_in_1x1 = cfg.MODEL.RESNETS.STRIDE_IN_1X1
    norm = cfg.MODEL.RESNETS.NORM
    assert not cfg.MODEL.RESNETS.DEFORM_ON_PER_STAGE[-1], "Deformable conv is not yet supported in res5 head."
    try:
        from detectron2.modeling.backbone.resnet import BottleneckBlock as D2Block, ResNet
        blocks = ResNet.make_stage(D2Block, 3, stride_per_block=[2, 1, 1], in_channels=out_channels // 2,
                                   bottleneck_channels=bottleneck_channels, out_channels=out_channels, num_groups=num_groups, norm=norm,
                                   stride_in_1x1=stride_in_1x1)
    except ImportError:
        blocks, cin = [], out_channels // 2
        for stride in (2, 1, 1):
            blocks.append(BottleneckBlock(cin, out_channels, bottleneck_channels=bottleneck_channels, stride=stride, num_groups=num_groups,
                                          norm=norm, stride_in_1x1=stride_in_1x1))
            cin = out_channels
    return nn.Sequential(*blocks), out_channels
