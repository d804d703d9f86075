"""LSM grounding head parity on the B200: GroundingHead module (projection GEMM + fused pair kernel +
pair-CE kernel) against the golden vectors of the REAL reference module and the oracle restatement.
Bars: fp32 mode 1e-4 robust-relative, bf16 mode 2e-2."""
import pytest
import torch

import locov_b200.modeling as M
from oracle import lsm_head
from util import frob_relerr, golden_cases, golden_lsm, relerr

pytestmark = pytest.mark.gpu


def _head(dev, V, D, w, b, precision, **g):
    cfg = M.get_cfg("lsm")
    cfg.MODEL.B200.PRECISION = precision
    dist = g.pop("distillation", True)
    cfg.MODEL.MMSS_HEAD.DISTILLATION_LOSS = dist
    m = {"alignment": "ALIGNMENT", "loss": "LOSS", "negative_mining": "NEGATIVE_MINING",
         "align_words": "ALIGN_WORDS_TO_REGIONS", "align_regions": "ALIGN_REGIONS_TO_WORDS"}
    for k, v in g.items():
        cfg.MODEL.MMSS_HEAD.GROUNDING[m[k]] = v
    head = M.GroundingHead(cfg, V, D).to(dev)
    with torch.no_grad():
        head.v2l_projection.weight.copy_(w)
        head.v2l_projection.bias.copy_(b)
    return head


def _to(d, dev):
    return {k: v.to(dev) for k, v in d.items()}


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("name", golden_cases("lsm_"))
def test_module_matches_reference_golden(cuda_device, name, precision, tol):
    ii, ic, w, b, cfg_kw, exp = golden_lsm(name)
    # (hardmax in bf16 is tested too: a flipped near-tie changes WHICH region / word is attended, not the attended value beyond the
    #  rounding error — the pair distances are max-type functions of the similarities, continuous in them)
    head = _head(cuda_device, w.shape[1], w.shape[0], w, b, precision, **cfg_kw)
    with torch.no_grad():
        out = head(_to(ii, cuda_device), _to(ic, cuda_device))
    info, losses = out[0], out[1]
    dists = out[2] if len(out) > 2 else {}
    assert len(out) == (3 if cfg_kw.get("distillation", True) else 2)
    n = 0
    for k, v in exp.items():
        kind, key = k.split("::", 1)
        got = {"loss": losses, "info": info, "dist": dists}[kind][key]
        ref = torch.from_numpy(v) if v.ndim else torch.tensor(float(v))
        if kind == "info":
            if precision == "fp32":
                assert abs(float(got) - float(ref)) <= 1.0 / ref.new_tensor(float(max(1, ii["region_mask"].shape[0]))) + 1e-6, k
        else:
            assert relerr(got.cpu(), ref) < tol, (k, relerr(got.cpu(), ref))
        n += 1
    assert n >= 4
    assert set(losses) == {k.split("::", 1)[1] for k in exp if k.startswith("loss::")}


@pytest.mark.parametrize("B,Rg,T,kw", [(3, 130, 9, dict(empty_caption=1, empty_image=1)), (7, 100, 70, dict(ragged_regions=True)),
                                        (33, 49, 20, dict(ragged_regions=True)), (2, 256, 128, {}), (1, 5, 3, dict(min_words=3))])
def test_shapes_and_edges_vs_oracle(cuda_device, B, Rg, T, kw):
    ii, ic, w, b = lsm_head.make_lsm_inputs(B=B, Rg=Rg, T=T, V=256, D=768, seed=B * 100 + T, gain=5.0, **kw)
    head = _head(cuda_device, 256, 768, w, b, "fp32")
    with torch.no_grad():
        info, losses, dists = head(_to(ii, cuda_device), _to(ic, cuda_device))
    rinfo, rlosses, rdists = lsm_head.grounding_head_forward(ii, ic, w, b, dtype=torch.float64)
    for k in rdists:
        assert torch.isfinite(dists[k]).all()
        assert relerr(dists[k].cpu(), rdists[k]) < 1e-4, k
    for k in rlosses:
        assert relerr(losses[k].cpu(), rlosses[k]) < 1e-4, k


def test_large_batch_properties(cuda_device):
    """BASELINE config 4 scale on one GPU (B=256): block consistency — any sub-block of the pair matrix
    equals the pair matrix of the sub-batch (the property the multi-GPU sharding relies on) — and
    permutation equivariance."""
    B, Rg, T = 256, 100, 20
    ii, ic, w, b = lsm_head.make_lsm_inputs(B=B, Rg=Rg, T=T, V=128, D=768, seed=4, gain=6.0, ragged_regions=True)
    head = _head(cuda_device, 128, 768, w, b, "fp32")
    with torch.no_grad():
        _, _, full = head(_to(ii, cuda_device), _to(ic, cuda_device))
        sl = slice(64, 96)
        sub_i = {k: v[sl] for k, v in ii.items()}
        sub_c = {k: v[sl] for k, v in ic.items()}
        _, _, sub = head(_to(sub_i, cuda_device), _to(sub_c, cuda_device))
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
        _, _, pm = head(_to({k: v[perm] for k, v in ii.items()}, cuda_device), _to({k: v[perm] for k, v in ic.items()}, cuda_device))
    for k in ("w2r", "r2w"):
        assert torch.equal(full[k][sl, sl], sub[k])
        assert torch.equal(full[k][perm][:, perm], pm[k])
    # oracle on a 16x16 corner
    c = slice(0, 16)
    _, _, ref = lsm_head.grounding_head_forward({k: v[c] for k, v in ii.items()}, {k: v[c] for k, v in ic.items()}, w, b, dtype=torch.float64)
    for k in ("w2r", "r2w"):
        assert relerr(full[k][c, c].cpu(), ref[k]) < 1e-4


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-4), ("bf16", 4e-2)])
@pytest.mark.parametrize("alignment", ["softmax", "hardmax"])
@pytest.mark.parametrize("B,Rg,T,kw", [(6, 37, 11, dict(ragged_regions=True)), (5, 100, 20, dict(empty_caption=1, empty_image=3)),
                                        (3, 100, 70, dict(ragged_regions=True))])
def test_backward_matches_autograd_of_the_oracle(cuda_device, precision, tol, alignment, B, Rg, T, kw):
    """d(sum of the 4 CE losses)/d(region_features, W, b, caption embeddings) vs torch autograd through the
    fp64 oracle — the gradients PyTorch derives for the reference module."""
    if alignment == "hardmax" and precision == "bf16":
        pytest.skip("hardmax is an argmax: bf16 rounding may legitimately flip near-ties")
    V, D = 192, 256
    ii, ic, w, b = lsm_head.make_lsm_inputs(B=B, Rg=Rg, T=T, V=V, D=D, seed=B * 7 + T, gain=8.0, **kw)
    head = _head(cuda_device, V, D, w, b, precision, alignment=alignment).train()
    feats = ii["region_features"].to(cuda_device).requires_grad_(True)
    cap = ic["input_embeddings"].to(cuda_device).requires_grad_(True)
    iid = dict(_to(ii, cuda_device), region_features=feats)
    icd = dict(_to(ic, cuda_device), input_embeddings=cap)
    info, losses, dists = head(iid, icd)
    weights = [1.0, 0.7, 1.3, 0.4]
    total = sum(wt * l for wt, l in zip(weights, losses.values()))
    total.backward()
    # oracle
    fr = ii["region_features"].double().requires_grad_(True)
    cr = ic["input_embeddings"].double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    _, rl, _ = lsm_head.grounding_head_forward(dict(ii, region_features=fr), dict(ic, input_embeddings=cr), wr, br,
                                               alignment=alignment, dtype=torch.float64)
    assert list(rl) == list(losses)
    rtotal = sum(wt * l for wt, l in zip(weights, rl.values()))
    rtotal.backward()
    assert relerr(total.detach().cpu(), rtotal.detach()) < tol
    err = relerr if precision == "fp32" else frob_relerr
    assert err(feats.grad.cpu(), fr.grad) < tol * 3
    assert err(head.v2l_projection.weight.grad.cpu(), wr.grad) < tol * 3
    assert err(head.v2l_projection.bias.grad.cpu(), br.grad) < tol * 3
    assert err(cap.grad.cpu(), cr.grad) < tol * 3


def test_masks_kernel(cuda_device):
    from locov_b200 import ops
    g = torch.Generator().manual_seed(0)
    att = torch.randint(0, 2, (9, 13), generator=g)
    spe = torch.randint(0, 2, (9, 13), generator=g)
    for reg in (torch.randint(0, 2, (9, 21), generator=g).to(torch.uint8), torch.randint(0, 2, (9, 21), generator=g).float(),
                torch.randint(0, 2, (9, 21), generator=g)):
        cm, rm = ops.lsm_masks(att.to(cuda_device), spe.to(cuda_device), reg.to(cuda_device))
        assert torch.equal(cm.cpu(), lsm_head.caption_mask_of(att, spe)) and torch.equal(rm.cpu(), reg.float())


@pytest.mark.parametrize("B", [5, 32, 33, 100, 256])
def test_pair_ce_kernel_all_batch_sizes(cuda_device, B):
    """Single-CTA staged path (B <= 32) and the multi-CTA ticketed path (B > 32): losses, accuracies, guard and the
    gradients of both cross-entropies vs torch autograd on the oracle."""
    from locov_b200 import ops
    g = torch.Generator().manual_seed(B)
    pw = torch.randn(2, B, B, generator=g) * 3
    mc = torch.ones(B, 7)
    mr = torch.ones(B, 9)
    mc[B // 2] = 0
    mr[B // 2] = 0
    mr[0] = 0
    pwd = pw.clone().to(cuda_device)
    for rep in range(2):                      # twice: the ticket counters must be left zeroed
        pwd = pw.clone().to(cuda_device)
        out, dcap, dimg = ops.pair_ce(pwd, mc.to(cuda_device), mr.to(cuda_device), 0, want_grad=True)
    ok = (mc.sum(1)[:, None] > 0) | (mr.sum(1)[None, :] > 0)
    for k in range(2):
        ref_in = pw[k].double().clone().requires_grad_(True)
        guarded = torch.where(ok, ref_in, (ref_in.max() + 100.0).detach())
        ce_cap, ce_img, acc_cap, acc_img = lsm_head.pair_losses(guarded)
        assert torch.allclose(pwd[k].cpu().double(), guarded.detach(), rtol=0, atol=1e-5)   # fp32 ulp at max+100 is 7.6e-6
        assert relerr(out[k].cpu(), torch.stack([ce_cap, ce_img, acc_cap.double(), acc_img.double()]).detach()) < 1e-5
        (gc,) = torch.autograd.grad(ce_cap, ref_in, retain_graph=True)
        (gi,) = torch.autograd.grad(ce_img, ref_in)
        assert relerr(dcap[k].cpu(), gc) < 1e-4 and relerr(dimg[k].cpu(), gi) < 1e-4


@pytest.mark.parametrize("accurate", [False, True])
@pytest.mark.parametrize("reg_dtype", [torch.uint8, torch.float32, torch.int64])
def test_lsm_prep_equals_masks_plus_split(cuda_device, accurate, reg_dtype):
    """The fused input-preparation launch is bit-identical to loco_lsm_masks + loco_split_bf16 (ragged D: pad columns zero)."""
    from locov_b200 import ops
    g = torch.Generator().manual_seed(11)
    b, t, rg, d = 5, 7, 9, 100                       # d % 8 != 0: padded operand row stride
    cap = (torch.randn(b * t, d, generator=g) * 0.05).to(cuda_device)
    att = (torch.rand(b, t, generator=g) > 0.3).long().to(cuda_device)
    spe = (torch.rand(b, t, generator=g) > 0.7).long().to(cuda_device)
    reg = (torch.rand(b, rg, generator=g) > 0.2).to(reg_dtype).to(cuda_device)
    op, cm, rm = ops.lsm_prep(cap, accurate, att, spe, reg)
    ref_op = ops.split_bf16(cap, accurate)
    cm2, rm2 = ops.lsm_masks(att, spe, reg)
    assert torch.equal(op.hi, ref_op.hi) and op.ld == ref_op.ld and (op.rows, op.cols) == (ref_op.rows, ref_op.cols)
    assert (op.lo is None) == (not accurate) and (op.lo is None or torch.equal(op.lo, ref_op.lo))
    assert torch.equal(cm, cm2) and torch.equal(rm, rm2)
    assert torch.equal(cm.cpu(), (att * (1 - spe)).float().cpu())
