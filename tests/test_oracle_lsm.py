"""Pins oracle/lsm_head.py against the REAL reference GroundingHead: the committed golden vectors
(tests/golden/lsm_*.npz, produced by tests/golden/make_golden.py from /root/reference) and, when the
reference tree is present (build container), the live module."""
import pytest
import torch

from oracle import lsm_head, ref_loader
from util import golden_cases, golden_lsm, relerr


@pytest.mark.parametrize("name", golden_cases("lsm_"))
def test_restatement_matches_golden(name):
    ii, ic, w, b, cfg_kw, exp = golden_lsm(name)
    info, losses, dists = lsm_head.grounding_head_forward(
        ii, ic, w, b, temperature=10.0, alignment=cfg_kw.get("alignment", "softmax"), loss=cfg_kw.get("loss", "cross_entropy"),
        align_words=cfg_kw.get("align_words", True), align_regions=cfg_kw.get("align_regions", True),
        negative_mining=cfg_kw.get("negative_mining", "hardest"))
    n = 0
    for k, v in exp.items():
        kind, key = k.split("::", 1)
        got = {"loss": losses, "info": info, "dist": dists}[kind][key]
        if kind == "info":
            assert float(got) == pytest.approx(float(v), abs=1e-6), k
        else:
            assert relerr(got, torch.from_numpy(v) if v.ndim else torch.tensor(float(v))) < 2e-5, k
        n += 1
    assert n >= 4


@pytest.mark.parametrize("name", ["small_softmax", "ragged_empty", "empty_mismatch"])
def test_literal_order_matches_golden(name):
    ii, ic, w, b, _, exp = golden_lsm(name)
    info, losses, dists = lsm_head.grounding_head_forward_literal(ii, ic, w, b, 10.0)
    for k, v in exp.items():
        kind, key = k.split("::", 1)
        got = {"loss": losses, "info": info, "dist": dists}[kind][key]
        assert relerr(got, torch.from_numpy(v) if v.ndim else torch.tensor(float(v))) < 2e-5, k


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree only exists in the build container")
def test_restatement_matches_live_reference():
    GH = ref_loader.load_reference_grounding_head()
    ii, ic, w, b = lsm_head.make_lsm_inputs(B=6, Rg=17, T=11, V=96, D=80, seed=5, ragged_regions=True, gain=10.0)
    head = GH(ref_loader.make_grounding_cfg(), 96, 80)
    with torch.no_grad():
        head.v2l_projection.weight.copy_(w)
        head.v2l_projection.bias.copy_(b)
    with ref_loader.cuda_to_cpu(), torch.no_grad():
        info_r, losses_r, dist_r = head(ii, ic)
    info, losses, dists = lsm_head.grounding_head_forward(ii, ic, w, b)
    for k in losses_r:
        assert relerr(losses[k], losses_r[k]) < 2e-5
    for k in dist_r:
        assert relerr(dists[k], dist_r[k]) < 2e-5
    for k in info_r:
        assert float(info[k]) == float(info_r[k])


def test_masks_and_padding_semantics():
    """An all-masked softmax row is uniform (finite fill), and the empty-pair guard only fires when
    BOTH the caption and the image are empty (grounding_head.py:240-251)."""
    ii, ic, w, b = lsm_head.make_lsm_inputs(B=4, Rg=6, T=5, V=16, D=16, seed=3, empty_caption=1, empty_image=2, gain=10.0)
    _, _, d = lsm_head.grounding_head_forward(ii, ic, w, b)
    for m in d.values():
        assert torch.isfinite(m).all()
        guard = m[1, 2]
        assert guard == m.max() and guard > 90          # max + 100
        assert (m[1, [0, 1, 3]].abs() < 50).all() and (m[[0, 2, 3], 2].abs() < 50).all()
