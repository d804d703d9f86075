"""Distillation losses on the pair matrices on the B200 (loco_pair_distill through the drop-in MultiDistillLoss / MultiDistillLossL2)
against the outputs — loss and autograd gradients — of the REAL reference classes (tests/golden/distill_*.npz)."""
import numpy as np
import pytest
import torch

import locov_b200.modeling as M
from oracle import distill
from util import load_golden, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(distill.DISTILL_CASES))
def test_matches_reference_golden(cuda_device, name):
    z = load_golden("distill_" + name)
    b, seed, kind, temp, lw, detach, tt = distill.DISTILL_CASES[name]
    t, w, r = [x.to(cuda_device).requires_grad_(True) for x in distill.distill_inputs(b, seed)]
    mod = {"KD": M.MultiDistillLoss, "MSE": M.MultiDistillLossL2}[kind](temp, lw, detach, tt)
    loss = mod(t, w, r)
    loss.backward()
    assert relerr(loss.detach().cpu(), z["loss"]) < 1e-4
    k = z["grad_trans"].shape[0]
    for nm, x in (("trans", t), ("w2r", w), ("r2w", r)):
        if bool(z["has_grad_" + nm]):
            assert x.grad is not None, nm
            assert relerr(x.grad[:k, :k].cpu(), z["grad_" + nm]) < 1e-4, nm
            assert abs(float(x.grad.double().abs().sum()) - float(z["gradsum_" + nm])) <= 1e-4 * float(z["gradsum_" + nm]) + 1e-9
        else:
            assert x.grad is None or float(x.grad.abs().max()) == 0.0, nm


def test_on_the_grounding_head_outputs_and_config_builder(cuda_device):
    """The three calls of a training step (distill_prop_mmss_gcnn.py:424-442) on real head outputs; gradients reach the head."""
    from oracle import lsm_head
    cfg = M.get_cfg("lsm")
    loss_mod = M.build_distill_loss(cfg)
    assert isinstance(loss_mod, M.MultiDistillLoss) and loss_mod.temp == 10.0 and loss_mod.transformer_teacher is False
    ii, ic, w, b = lsm_head.make_lsm_inputs(B=8, Rg=20, T=9, V=64, D=64, seed=3, gain=8.0)
    head = M.GroundingHead(cfg, 64, 64).to(cuda_device)
    with torch.no_grad():
        head.v2l_projection.weight.copy_(w); head.v2l_projection.bias.copy_(b)
    _, _, d = head({k: v.to(cuda_device) for k, v in ii.items()}, {k: v.to(cuda_device) for k, v in ic.items()})
    # the head's distances are small numbers: at temperature 10 every softmax would be nearly uniform and the KL a second-order
    # residual of cancelling terms (fp32 noise ~1e-2 of it, in the reference's fp32 evaluation as well) — scale them to a spread of
    # ~30 so that the comparison is well conditioned; the teacher is an unrelated matrix of the same scale
    gain = 30.0 / float(d["w2r"].detach().std())
    w2r, r2w = d["w2r"] * gain, d["r2w"] * gain
    g = torch.Generator(device=cuda_device).manual_seed(5)
    teacher = (torch.randn(w2r.shape, generator=g, device=cuda_device) * 30.0 + w2r.detach().mean()).requires_grad_(True)
    loss = loss_mod(teacher, w2r, r2w)
    loss.backward()
    ref = distill.kd_loss(teacher.detach().cpu().double(), w2r.detach().cpu().double(), r2w.detach().cpu().double(), 10.0, 1.0, False, False)
    assert float(ref) > 0.1
    assert relerr(loss.detach().cpu(), ref) < 1e-4
    assert head.v2l_projection.weight.grad is not None and float(head.v2l_projection.weight.grad.abs().sum()) > 0
    with pytest.raises(NotImplementedError):
        M.MultiDistillLossJS(1.0)


def test_logged_module_statistics_on_the_device(cuda_device):
    """LoggedModule.log keeps a reference; reading log_info computes min / max / mean / std with one launch of loco_tensor_stats."""
    from locov_b200 import _lib
    m = M.LoggedModule()
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(32, 100, 2048, generator=g) * 3 + 7).to(cuda_device)
    n0 = _lib.load().loco_launch_count()
    m.log("region_features", x)
    assert _lib.load().loco_launch_count() == n0               # logging itself launches nothing
    s = m.log_info["region_features"]
    assert _lib.load().loco_launch_count() == n0 + 1
    xc = x.cpu().double()
    assert abs(s["min"] - float(xc.min())) < 1e-6 * abs(float(xc.min())) + 1e-6 and abs(s["max"] - float(xc.max())) < 1e-5
    assert abs(s["mean"] - float(xc.mean())) < 1e-5 and abs(s["std"] - float(xc.std())) < 1e-5
    assert tuple(s["shape"]) == (32, 100, 2048)
