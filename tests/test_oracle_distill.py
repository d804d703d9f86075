"""oracle/distill.py (restated distill_mmss_gcnn.py:211-289, 381-433) against the golden outputs of the REAL reference classes
(tests/golden/distill_*.npz) and, when /root/reference is present, against the live classes."""
import numpy as np
import pytest
import torch

from oracle import distill, ref_loader
from util import load_golden, relerr


def _run(name):
    b, seed, kind, temp, lw, detach, tt = distill.DISTILL_CASES[name]
    t, w, r = [x.clone().requires_grad_(True) for x in distill.distill_inputs(b, seed)]
    fn = distill.kd_loss if kind == "KD" else distill.mse_loss
    loss = fn(t, w, r, temp, lw, detach, tt)
    loss.backward()
    return (t, w, r), loss


@pytest.mark.parametrize("name", sorted(distill.DISTILL_CASES))
def test_restatement_matches_reference_golden(name):
    z = load_golden("distill_" + name)
    (t, w, r), loss = _run(name)
    assert np.allclose([float(x.double().sum()) for x in (t, w, r)], z["checksum"], rtol=1e-9)
    assert relerr(loss.detach(), z["loss"]) < 2e-5
    k = z["grad_trans"].shape[0]
    for nm, x in (("trans", t), ("w2r", w), ("r2w", r)):
        if bool(z["has_grad_" + nm]):
            assert relerr(x.grad[:k, :k], z["grad_" + nm]) < 1e-4, nm
            assert abs(float(x.grad.double().abs().sum()) - float(z["gradsum_" + nm])) <= 1e-4 * float(z["gradsum_" + nm]) + 1e-9
        else:
            assert x.grad is None or float(x.grad.abs().max()) == 0.0, nm


@pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference (build container only)")
def test_restatement_matches_the_live_reference_classes():
    mod = ref_loader.load_reference_distill()
    t, w, r = distill.distill_inputs(9, 77)
    for tt in (True, False):
        a = mod.MultiDistillLoss(3.0, 0.7, False, tt)(t, w, r)
        assert relerr(distill.kd_loss(t, w, r, 3.0, 0.7, False, tt), a) < 1e-5
    assert relerr(distill.mse_loss(t, w, r, 1.0, 1.3), mod.MultiDistillLossL2(1.0, 1.3)(t, w, r)) < 1e-5
