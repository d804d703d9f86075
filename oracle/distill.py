"""ORACLE (test infrastructure only): CPU restatement of the distillation losses on the LSM pair matrices.

Follows /root/reference/ovr/modeling/meta_arch/distill_mmss_gcnn.py:
  * MultiDistillLoss.forward    :226-289  (KL between the caption-wise / image-wise softmax distributions of a teacher pair matrix
                                           and of the w2r / r2w pair matrices, temperature-scaled; teacher = transformer or LSM head)
  * MultiDistillLossL2.forward  :381-433  (mean-squared error between the matrices, each counted twice: as is and transposed)
called three times per step from distill_prop_mmss_gcnn.py:424-442.  Pinned by tests/golden/distill_*.npz, the outputs (loss and
autograd gradients) of the REAL reference classes (oracle/ref_loader.load_reference_distill; generator tests/golden/make_golden_distill.py),
and against the live classes when /root/reference is present (tests/test_oracle_distill.py).
MultiDistillLossJS (:292-378) pairs image-wise log-probabilities with caption-wise mixtures (m_cap_* in the image terms, :361-370); it is
not used by the shipped configuration (DISTILLATION_LOSS_TYPE defaults to "KD") and is not part of the B200 path.
"""
import torch


def kd_loss(trans, w2r, r2w, temperature, loss_weight=1.0, detach_teacher=False, transformer_teacher=True):
    """MultiDistillLoss.forward.  kldiv(input = log q, target = p, 'batchmean') = sum p (log p - log q) / B."""
    b = trans.shape[0]
    t2 = temperature * temperature

    def kl(target_logits, input_logits, dim):
        lp = torch.log_softmax(target_logits, dim=dim)
        lq = torch.log_softmax(input_logits, dim=dim)
        return (lp.exp() * (lp - lq)).sum() / b * t2

    if transformer_teacher:
        if detach_teacher:
            trans = trans.detach()
        a_t = -trans / temperature
        total = 0.0
        for s in (w2r, r2w):
            a_s = -s / temperature
            total = total + kl(a_t, a_s, 0) + kl(a_t, a_s, 1)
    else:
        if detach_teacher:
            w2r, r2w = w2r.detach(), r2w.detach()
        a_t = -trans / temperature
        total = 0.0
        for s in (w2r, r2w):
            a_s = -s / temperature
            total = total + kl(a_s, a_t, 0) + kl(a_s, a_t, 1)
    return total * loss_weight


def mse_loss(trans, w2r, r2w, temperature=1.0, loss_weight=1.0, detach_teacher=False, transformer_teacher=True):
    """MultiDistillLossL2.forward: every pair of matrices is compared as is and transposed (the same value twice)."""
    if transformer_teacher:
        if detach_teacher:
            trans = trans.detach()
    elif detach_teacher:
        w2r, r2w = w2r.detach(), r2w.detach()
    return 2.0 * (((trans - w2r) ** 2).mean() + ((trans - r2w) ** 2).mean()) * loss_weight


DISTILL_CASES = {
    # name: (B, seed, kind, temperature, loss_weight, detach_teacher, transformer_teacher)
    "kd_lsm_teacher_b6": (6, 31, "KD", 10.0, 1.0, False, False),          # shipped: coco_lsm.yaml:60-63
    "kd_trans_teacher_b6": (6, 32, "KD", 2.0, 0.5, False, True),
    "kd_trans_detach_b7": (7, 33, "KD", 1.0, 1.0, True, True),
    "kd_lsm_detach_b7": (7, 34, "KD", 4.0, 2.0, True, False),
    "kd_lsm_teacher_b32": (32, 35, "KD", 10.0, 1.0, False, False),
    "kd_lsm_teacher_b256": (256, 36, "KD", 10.0, 1.0, False, False),      # the sharded head's global batch (BASELINE configs[3])
    "mse_b6": (6, 37, "MSE", 1.0, 1.0, False, True),
    "mse_detach_b32": (32, 38, "MSE", 1.0, 0.25, True, False),
}


def distill_inputs(b, seed):
    g = torch.Generator().manual_seed(seed)
    trans = torch.randn(b, b, generator=g) * 3.0 + 1.0
    w2r = trans * 0.5 + torch.randn(b, b, generator=g) * 2.0
    r2w = torch.randn(b, b, generator=g) * 4.0 - 2.0
    return trans, w2r, r2w
