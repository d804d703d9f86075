"""The oracle restatement of the multi-token class scoring (oracle/grounding_module.py) against vectors produced by the reference's
own GroundingModule / EmbeddingGroundingFastRCNNOutputLayers (tests/golden/gbox_*.npz, box_emb_grounding_head.py:60-434)."""
import os

import numpy as np
import pytest
import torch

from oracle import grounding_module as gm
from util import relerr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _forward(c, d, dtype=torch.float64):
    tok, off = gm.class_token_matrix(d["embs"], c["D"], normalize=bool(c.get("normalize", False)))
    e = d["x"].to(dtype) @ d["w_emb"].to(dtype).t() + d["b_emb"].to(dtype)
    if c.get("normalize"):
        e = e / (e ** 2).sum(1, keepdim=True).sqrt()
    scores, att = gm.grounding_scores(e, tok, off, c["temperature"], c["alignment"], dtype)
    deltas = d["x"].to(dtype) @ d["w_box"].to(dtype).t() + d["b_box"].to(dtype)
    return scores, deltas, att, off


@pytest.mark.parametrize("name", sorted(gm.GBOX_CASES))
def test_restatement_matches_the_reference_class(name):
    c = gm.GBOX_CASES[name]
    d = gm.gbox_inputs(c)
    z = np.load(os.path.join(GOLDEN, f"gbox_{name}.npz"))
    assert abs(float(d["x"].double().sum()) - float(z["checksum_x"][0])) < 1e-9
    scores, deltas, att, off = _forward(c, d)
    assert scores.shape == z["scores"].shape == (c["R"], c["K"] + 1)
    assert int(z["num_classes"]) == c["K"]
    tol = 2e-5 if c["alignment"] == "softmax" else 2e-5
    assert relerr(scores, z["scores"]) < tol
    assert relerr(deltas, z["deltas"]) < 2e-5
    assert float(scores[:, -1].abs().max()) == 0.0 and float(np.abs(z["scores"][:, -1]).max()) == 0.0     # background logit exactly 0
    if "tok_attention" in z.files:
        ta = z["tok_attention"]                                     # [R, K1, max_tok], masked positions 0
        for k, a in enumerate(att[:-1]):
            assert relerr(a, ta[:, k, :a.shape[1]]) < 1e-4
            assert float(np.abs(ta[:, k, a.shape[1]:]).max(initial=0.0)) == 0.0
        assert list(z["num_tok"][:-1]) == [int(off[k + 1] - off[k]) for k in range(c["K"])]


def test_live_against_the_reference_when_present():
    from oracle import d2_stubs, ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference tree not present")
    mod = ref_loader.load_reference_grounding_box_head()
    c = dict(R=30, K=7, V=48, D=24, max_tok=4, seed=77, alignment="softmax", temperature=3.0, mode="eval")
    d = gm.gbox_inputs(c)
    cfg = ref_loader.make_roi_cfg("stt", **{"MODEL.ROI_HEADS.MAX_TOKENS": 4, "MODEL.ROI_BOX_HEAD.EMB_DIM": 24, "MODEL.ROI_HEADS.NUM_CLASSES": 7,
                                           "MODEL.MMSS_HEAD.GROUNDING.ALIGNMENT_TEMPERATURE": 3.0})
    bp = mod.EmbeddingGroundingFastRCNNOutputLayers(cfg, d2_stubs.ShapeSpec(channels=48))
    with torch.no_grad():
        bp.emb_pred.weight.copy_(d["w_emb"]); bp.emb_pred.bias.copy_(d["b_emb"])
        bp.bbox_pred.weight.copy_(d["w_box"]); bp.bbox_pred.bias.copy_(d["b_box"])
    bp.set_class_embeddings(d["embs"])
    with torch.no_grad():
        s_ref, d_ref = bp(d["x"])
    scores, deltas, _, _ = _forward(c, d)
    assert relerr(scores, s_ref) < 2e-5 and relerr(deltas, d_ref) < 2e-5
