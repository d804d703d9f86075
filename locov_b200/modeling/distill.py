"""Distillation losses on the LSM pair matrices — drop-ins for the reference's ``MultiDistillLoss`` and ``MultiDistillLossL2``
(ovr/modeling/meta_arch/distill_mmss_gcnn.py:211-289, 381-433), the direct consumers of the ``{"w2r", "r2w"}`` matrices the
grounding head returns with ``DISTILLATION_LOSS`` (grounding_head.py:384-388); called three times per training step
(distill_prop_mmss_gcnn.py:424-442).  Same constructor arguments and call signature; loss and gradients come from two kernel launches
(``loco_pair_distill``) instead of ~40 ATen launches forward + backward.

``MultiDistillLossJS`` (:292-378; DISTILLATION_LOSS_TYPE "JS", not used by the shipped configuration) raises NotImplementedError at
construction — the error policy of SURVEY.md §8b: never a silent difference.
"""
from torch import nn

from .. import functional as LF
from .. import ops


class MultiDistillLoss(nn.Module):
    def __init__(self, temperature, loss_weight=1.0, detach_teacher=False, transformer_teacher=True):
        super().__init__()
        self.temp = temperature
        self.loss_weight = loss_weight
        self.detach_teacher = detach_teacher
        self.transformer_teacher = transformer_teacher

    def forward(self, trans_pw_cost, pw_cost_w2r, pw_cost_r2w):
        kind = ops.DISTILL_KD_TEACHER_TARGET if self.transformer_teacher else ops.DISTILL_KD_STUDENT_TARGET
        return LF.pair_distill(trans_pw_cost, pw_cost_w2r, pw_cost_r2w, self.temp, kind, self.loss_weight, self.detach_teacher)


class MultiDistillLossL2(nn.Module):
    def __init__(self, temperature, loss_weight=1.0, detach_teacher=False, transformer_teacher=True):
        super().__init__()
        self.temp = temperature
        self.loss_weight = loss_weight
        self.detach_teacher = detach_teacher
        self.transformer_teacher = transformer_teacher

    def forward(self, trans_pw_cost, pw_cost_w2r, pw_cost_r2w):
        # detach: the transformer teacher's matrix, or (LSM teacher) the two student matrices — same rule as the KD loss
        if self.detach_teacher and not self.transformer_teacher:
            pw_cost_w2r, pw_cost_r2w = pw_cost_w2r.detach(), pw_cost_r2w.detach()
        elif self.detach_teacher:
            trans_pw_cost = trans_pw_cost.detach()
        return LF.pair_distill(trans_pw_cost, pw_cost_w2r, pw_cost_r2w, max(float(self.temp), 1e-6), ops.DISTILL_MSE, self.loss_weight, False)


class MultiDistillLossJS(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("DISTILLATION_LOSS_TYPE 'JS' (distill_mmss_gcnn.py:292-378) is not part of the B200 path; the shipped "
                                  "configuration uses 'KD'")


def build_distill_loss(cfg):
    """distill_prop_mmss_gcnn.py:127-149: the loss object for cfg.MODEL.MMSS_HEAD.DISTILLATION_*, or None."""
    h = cfg.MODEL.MMSS_HEAD
    if not h.DISTILLATION_LOSS:
        return None
    kinds = {"KD": MultiDistillLoss, "JS": MultiDistillLossJS, "MSE": MultiDistillLossL2}
    kind = getattr(h, "DISTILLATION_LOSS_TYPE", "KD")
    if kind not in kinds:
        raise NotImplementedError(f"DISTILLATION_LOSS_TYPE {kind!r}")
    return kinds[kind](getattr(h, "DISTILLATION_TEMPERATURE", 1.0), getattr(h, "DISTILLATION_LOSS_WEIGHT", 1.0),
                       getattr(h, "DISTILLATION_DETACH_TEACHER", False), getattr(h, "DISTILLATION_TEACHER_TRANSFORMER", True))
