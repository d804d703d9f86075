"""Box predictor parity on the B200: EmbeddingFastRCNNOutputLayers (tcgen05 projection + fused
scoring epilogue) vs oracle/box_head.py.  Bars (BASELINE.json north_star): fp32 mode 1e-4 relative and
identical per-RoI argmax; bf16 mode 2e-2 relative."""
import pytest
import torch
import torch.nn.functional as F

import locov_b200.modeling as M
from oracle import box_head
from util import frob_relerr, relerr

pytestmark = pytest.mark.gpu


def _predictor(dev, V, D, cls, w_emb, b_emb, w_box, b_box, precision, stage="stt", **over):
    cfg = M.get_cfg(stage)
    cfg.MODEL.ROI_BOX_HEAD.EMB_DIM = D
    cfg.MODEL.B200.PRECISION = precision
    for k, v in over.items():
        cfg.MODEL.ROI_BOX_HEAD[k] = v
    bp = M.build_box_predictor(cfg, V).to(dev)
    with torch.no_grad():
        bp.emb_pred.weight.copy_(w_emb); bp.emb_pred.bias.copy_(b_emb)
        bp.bbox_pred.weight.copy_(w_box); bp.bbox_pred.bias.copy_(b_box)
    bp.set_class_embeddings(cls)
    return bp


def _props(gt, dev):
    n = gt.shape[0]
    pb = torch.rand(n, 2) * 300
    prop = torch.cat([pb, pb + 16 + torch.rand(n, 2) * 100], 1)
    gtb = prop + torch.randn(n, 4) * 3
    half = n // 2
    mk = lambda s: M.Instances((800, 1216), proposal_boxes=M.Boxes(prop[s].to(dev)), gt_boxes=M.Boxes(gtb[s].to(dev)),
                               gt_classes=gt[s].to(dev))
    return [mk(slice(0, half)), mk(slice(half, n))], prop, gtb


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("R,K", [(1024, 65), (1000, 1203), (37, 17), (128, 48)])
def test_scores_match_oracle(cuda_device, precision, tol, R, K):
    x, we, be, wb, bb, cls, gt = box_head.make_box_inputs(R, K, seed=R + K)
    be = torch.randn(768) * 0.01
    bb = torch.randn(4) * 0.01
    bp = _predictor(cuda_device, 2048, 768, cls, we, be, wb, bb, precision).eval()
    with torch.no_grad():
        scores, deltas = bp(x.to(cuda_device))
    ref_s, ref_d, _ = box_head.box_predictor_forward(x, we, be, cls, torch.zeros(K + 1), wb, bb, dtype=torch.float64)
    assert scores.shape == (R, K + 1) and deltas.shape == (R, 4)
    assert relerr(scores.cpu(), ref_s) < tol
    assert relerr(deltas.cpu(), ref_d) < tol
    assert float(scores[:, -1].abs().max()) == 0.0          # zero background row -> logit exactly 0
    probs = torch.cat(bp.predict_probs((scores, deltas), [range(R)]), 0)
    ref_p, ref_arg = box_head.predict_probs(ref_s)
    assert relerr(probs.cpu(), ref_p) < tol
    arg = bp.predict_classes((scores, deltas)).cpu()
    if precision == "fp32":
        # identical argmax wherever the fp64 top-2 margin exceeds the fp32 tolerance
        top2 = ref_s[:, :-1].topk(2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * ref_s.abs().max()
        assert torch.equal(arg[clear], ref_arg[clear]) and clear.float().mean() > 0.99
        fp32_arg = box_head.predict_probs(box_head.box_predictor_forward(x, we, be, cls, torch.zeros(K + 1), wb, bb)[0])[1]
        assert (arg == fp32_arg).float().mean() >= 0.999
    else:
        assert (arg == ref_arg).float().mean() > 0.97


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_losses_and_gradients(cuda_device, precision, tol):
    R, K = 512, 48
    x, we, be, wb, bb, cls, gt = box_head.make_box_inputs(R, K, seed=77)
    bp = _predictor(cuda_device, 2048, 768, cls, we, be, wb, bb, precision, FREEZE_EMB_PRED=False).train()
    props, prop, gtb = _props(gt, cuda_device)
    xg = x.to(cuda_device).requires_grad_(True)
    pred = bp(xg)
    losses = bp.losses(pred, props)
    (losses["loss_cls"] + losses["loss_box_reg"]).backward()
    # oracle in fp64 with torch autograd
    xr = x.double().requires_grad_(True)
    wer, ber, wbr, bbr = [t.double().requires_grad_(True) for t in (we, be, wb, bb)]
    s, d, _ = box_head.box_predictor_forward(xr, wer, ber, cls.double(), torch.zeros(K + 1, dtype=torch.float64), wbr, bbr, dtype=torch.float64)
    ref = box_head.box_losses(s, d, gt, prop.double(), gtb.double())
    (ref["loss_cls"] + ref["loss_box_reg"]).backward()
    assert relerr(losses["loss_cls"].cpu(), ref["loss_cls"]) < tol
    assert relerr(losses["loss_box_reg"].cpu(), ref["loss_box_reg"]) < tol
    err = relerr if precision == "fp32" else frob_relerr
    assert err(xg.grad.cpu(), xr.grad) < tol * 5
    assert err(bp.emb_pred.weight.grad.cpu(), wer.grad) < tol * 5
    assert err(bp.emb_pred.bias.grad.cpu(), ber.grad) < tol * 5
    assert err(bp.bbox_pred.weight.grad.cpu(), wbr.grad) < tol * 5
    assert err(bp.bbox_pred.bias.grad.cpu(), bbr.grad) < tol * 5


def test_detached_classifier_and_inference(cuda_device):
    R, K = 200, 17
    x, we, be, wb, bb, cls, gt = box_head.make_box_inputs(R, K, seed=5)
    bp = _predictor(cuda_device, 2048, 768, cls * 4, we, be, wb, bb, "fp32", stage="lsm").train()
    xg = x.to(cuda_device).requires_grad_(True)
    scores, deltas = bp(xg)
    assert not scores.requires_grad and deltas.requires_grad       # box_emb_head.py:197-199
    props, _, _ = _props(gt, cuda_device)
    losses = bp.losses((scores, deltas), props)
    assert float(losses["loss_cls"]) == 0.0                        # loss weight 0 when detached (:147-149)
    losses["loss_box_reg"].backward()
    assert xg.grad is not None and bp.emb_pred.weight.grad is None
    bp.eval()
    with torch.no_grad():
        pred = bp(x.to(cuda_device))
        inst, kept = bp.inference(pred, props)
    assert len(inst) == 2 and all(len(i) <= 100 for i in inst)
    ref_s, ref_d, _ = box_head.box_predictor_forward(x, we, be, cls * 4, torch.zeros(K + 1), wb, bb)
    p = F.softmax(ref_s, -1)[:100, :-1]
    n_ref = int((p > 0.05).sum())
    assert n_ref == 0 or len(inst[0]) > 0


def test_reset_class_embeddings_refreshes_operands(cuda_device):
    x, we, be, wb, bb, cls, gt = box_head.make_box_inputs(64, 48, seed=9)
    bp = _predictor(cuda_device, 2048, 768, cls, we, be, wb, bb, "fp32").eval()
    with torch.no_grad():
        s1, _ = bp(x.to(cuda_device))
        cls2 = torch.cat([torch.randn(65, 768) * 0.05, torch.zeros(1, 768)])
        bp.set_class_embeddings(cls2)
        s2, _ = bp(x.to(cuda_device))
        bp.emb_pred.weight.mul_(2.0)                                # in-place update must invalidate the bf16 shadow
        s3, _ = bp(x.to(cuda_device))
    ref2 = box_head.box_predictor_forward(x, we, be, cls2, torch.zeros(66), wb, bb, dtype=torch.float64)[0]
    ref3 = box_head.box_predictor_forward(x, 2 * we, be, cls2, torch.zeros(66), wb, bb, dtype=torch.float64)[0]
    assert s1.shape == (64, 49) and s2.shape == (64, 66)
    assert relerr(s2.cpu(), ref2) < 1e-4 and relerr(s3.cpu(), ref3) < 1e-4
