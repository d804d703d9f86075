"""oracle/box_head.py (restated box_emb_head.py:179-236 + Detectron2 losses/inference semantics)
against torch's own primitives — the reference class itself needs Detectron2 (parity unpinned beyond
these primitives, see oracle/__init__.py)."""
import torch
import torch.nn.functional as F

from oracle import box_head


def test_forward_is_three_linears_and_background_logit_is_zero():
    x, we, be, wb, bb, cls, gt = box_head.make_box_inputs(64, 5, V=32, D=16, seed=2)
    w_cls, b_cls = box_head.prepare_class_embeddings(cls)
    scores, deltas, e = box_head.box_predictor_forward(x, we, be, w_cls, b_cls, wb, bb)
    lin_e = torch.nn.Linear(32, 16)
    lin_c = torch.nn.Linear(16, 6)
    with torch.no_grad():
        lin_e.weight.copy_(we); lin_e.bias.copy_(be); lin_c.weight.copy_(cls); lin_c.bias.zero_()
        assert torch.equal(scores, lin_c(lin_e(x)))
    assert torch.equal(scores[:, -1], torch.zeros(64))          # zero background row -> logit exactly 0
    assert deltas.shape == (64, 4)


def test_losses_and_probs():
    x, we, be, wb, bb, cls, gt = box_head.make_box_inputs(128, 7, V=32, D=16, seed=3)
    scores, deltas, _ = box_head.box_predictor_forward(x, we, be, cls, torch.zeros(8), wb, bb)
    out = box_head.box_losses(scores, deltas, gt)
    assert torch.allclose(out["loss_cls"], F.nll_loss(F.log_softmax(scores, 1), gt))
    probs, arg = box_head.predict_probs(scores)
    assert torch.allclose(probs.sum(1), torch.ones(128), atol=1e-6)
    assert torch.equal(arg, scores[:, :-1].argmax(1))
    pb = torch.rand(128, 2) * 100
    prop = torch.cat([pb, pb + 20 + torch.rand(128, 2) * 50], 1)
    gtb = prop + torch.randn(128, 4)
    out = box_head.box_losses(scores, deltas, gt, prop, gtb)
    fg = gt < 7
    tgt = box_head.get_deltas(prop[fg], gtb[fg], (10., 10., 5., 5.))
    assert torch.allclose(out["loss_box_reg"], (deltas[fg] - tgt).abs().sum() / 128)


def test_normalize_standardize_variants():
    e = torch.randn(10, 16)
    assert torch.allclose(box_head.normalize_vec(e).norm(dim=1), torch.ones(10), atol=1e-6)
    s = box_head.standardize_vec(e)
    assert torch.allclose(s.mean(1), torch.zeros(10), atol=1e-6) and torch.allclose(s.std(1), torch.ones(10), atol=1e-5)
