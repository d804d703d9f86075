"""ORACLE (test infrastructure only; never imported by the product path): CPU restatement of the reference's multi-token class
scoring, ``GroundingModule`` (ovr/modeling/roi_heads/box_emb_grounding_head.py:60-277) as used by
``EmbeddingGroundingFastRCNNOutputLayers.forward_cls_prediction`` (:419-427).  PINNED: tests/golden/gbox_*.npz are outputs of the
reference's own classes (imported unmodified through oracle/ref_loader.py by tests/golden/make_golden_gbox.py); the restatement is
checked against them in tests/test_oracle_gbox.py.

Every class k is a list of n_k >= 1 token embeddings (the background class: one zero row, box_emb_grounding_head.py:239-241).
For a RoI embedding e (box_emb_grounding_head.py:89-127, :163-237):
    s_t   = <e, tok_t> / temperature                       token_score = Linear(D -> sum n_k), :93 / :116
    a     = softmax over the class's tokens of s           masked positions are filled with min(s) - 100 and get weight exp(-100 - ..) = 0
            (hardmax: one-hot of the first maximum)         :164-176
    score = sum_t a_t * s_t                                 -global_dist with local_distance = -s, :178-185, :218-221
The background class: its single token is a zero row -> s = 0 -> score 0 (num_tok is patched 0 -> 1 in place by :124-125 on the
first forward, so the ``num_tok > 0`` fill of :187-189 never applies).
"""
import torch


def class_token_matrix(embs: dict, emb_dim: int, normalize: bool = False, background: bool = True):
    """{class index: [n_k, D]} -> (tokens [sum n_k (+1), D], seg offsets [K(+1) + 1] int64), box_emb_grounding_head.py:223-262."""
    k = len(embs)
    rows, off = [], [0]
    for c in range(k):
        e = torch.as_tensor(embs[c], dtype=torch.float32).reshape(-1, emb_dim)
        rows.append(e)
        off.append(off[-1] + e.shape[0])
    if background:
        rows.append(torch.zeros(1, emb_dim))
        off.append(off[-1] + 1)
    tok = torch.cat(rows, 0)
    if normalize:                                   # logged_module.normalize_vec (:55-66): NaN (zero rows) -> 0
        n = (tok ** 2).sum(1, keepdim=True).sqrt()
        tok = torch.where(n > 0, tok / n, torch.zeros_like(tok))
    return tok, torch.tensor(off, dtype=torch.int64)


def grounding_scores(e, tok, seg_off, temperature=1.0, alignment="softmax", dtype=torch.float64):
    """e [R, D], tok [Ttot, D], seg_off [K1 + 1] -> (scores [R, K1], attention list of [R, n_k])."""
    e, tok = e.to(dtype), tok.to(dtype)
    s = (e @ tok.t()) / temperature
    scores, att = [], []
    for k in range(len(seg_off) - 1):
        sk = s[:, int(seg_off[k]):int(seg_off[k + 1])]
        if alignment == "softmax":
            a = torch.softmax(sk, 1)
        elif alignment == "hardmax":
            a = torch.nn.functional.one_hot(sk.argmax(1), sk.shape[1]).to(dtype)
        else:
            raise NotImplementedError(alignment)
        att.append(a)
        scores.append((a * sk).sum(1))
    return torch.stack(scores, 1), att


# ---- seeded cases behind tests/golden/gbox_*.npz ---------------------------------------------------------------------------------
GBOX_CASES = {
    # R rois, K foreground classes with 1..max_tok tokens each, V -> D projection
    "soft_t10": dict(R=96, K=23, V=128, D=64, max_tok=5, seed=301, alignment="softmax", temperature=10.0, mode="eval"),
    "soft_t1_train": dict(R=64, K=12, V=96, D=32, max_tok=4, seed=302, alignment="softmax", temperature=1.0, mode="train"),
    "hard_t10": dict(R=80, K=17, V=128, D=64, max_tok=6, seed=303, alignment="hardmax", temperature=10.0, mode="eval"),
    "hard_train": dict(R=48, K=9, V=64, D=32, max_tok=3, seed=306, alignment="hardmax", temperature=2.0, mode="train"),
    "single_tok": dict(R=40, K=30, V=64, D=32, max_tok=1, seed=304, alignment="softmax", temperature=10.0, mode="eval"),
    "normalize": dict(R=50, K=11, V=64, D=32, max_tok=4, seed=305, alignment="softmax", temperature=0.1, mode="eval", normalize=True),
}


def gbox_inputs(c):
    g = torch.Generator().manual_seed(c["seed"])
    x = torch.randn(c["R"], c["V"], generator=g)
    w_emb = torch.randn(c["D"], c["V"], generator=g) * (1.0 / c["V"] ** 0.5)
    b_emb = torch.randn(c["D"], generator=g) * 0.01
    w_box = torch.randn(4, c["V"], generator=g) * 0.02
    b_box = torch.randn(4, generator=g) * 0.01
    gain = 6.0 if c["temperature"] >= 1.0 else 1.0
    embs = {}
    for k in range(c["K"]):
        n = 1 + int(torch.randint(0, c["max_tok"], (1,), generator=g))
        embs[k] = torch.randn(n, c["D"], generator=g) * gain
    embs[0] = embs[0][:1]                                         # at least one single-token and one full-length class
    embs[c["K"] - 1] = torch.randn(c["max_tok"], c["D"], generator=g) * gain
    return dict(x=x, w_emb=w_emb, b_emb=b_emb, w_box=w_box, b_box=b_box, embs=embs)
