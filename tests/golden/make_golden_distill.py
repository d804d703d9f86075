"""Generates tests/golden/distill_*.npz — run in the BUILD container only (needs /root/reference):
    python tests/golden/make_golden_distill.py
Loss values and autograd gradients of the REAL reference classes MultiDistillLoss / MultiDistillLossL2
(/root/reference/ovr/modeling/meta_arch/distill_mmss_gcnn.py:211-289, 381-433, imported unmodified through oracle/ref_loader.py) on the
seeded pair matrices of oracle/distill.py.  Inputs are regenerated from the seed (checksums stored); the b256 case stores a corner of
each gradient."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import distill, ref_loader  # noqa: E402


def main():
    mod = ref_loader.load_reference_distill()
    for name, (b, seed, kind, temp, lw, detach, trans_teacher) in distill.DISTILL_CASES.items():
        t, w, r = [x.clone().requires_grad_(True) for x in distill.distill_inputs(b, seed)]
        cls = {"KD": mod.MultiDistillLoss, "MSE": mod.MultiDistillLossL2}[kind]
        loss = cls(temp, lw, detach, trans_teacher)(t, w, r)
        loss.backward()
        rec = {"loss": np.float64(loss.item()), "checksum": np.array([float(x.double().sum()) for x in (t, w, r)])}
        k = b if b <= 32 else 16
        for nm, x in (("trans", t), ("w2r", w), ("r2w", r)):
            rec["grad_" + nm] = (x.grad[:k, :k] if x.grad is not None else torch.zeros(k, k)).numpy().astype(np.float32)
            rec["has_grad_" + nm] = np.bool_(x.grad is not None)
            rec["gradsum_" + nm] = np.float64(x.grad.double().abs().sum().item() if x.grad is not None else 0.0)
        np.savez_compressed(os.path.join(HERE, f"distill_{name}.npz"), **rec)
        print(name, float(loss))


if __name__ == "__main__":
    main()
