"""ORACLE (test infrastructure only): CPU restatement of the LSM grounding head.

Follows /root/reference/ovr/modeling/mmss_heads/grounding_head.py:92-388 (GroundingHead.forward),
closed form of SURVEY.md Appendix B.  Pinned against the real reference class (oracle/ref_loader.py)
by tests/golden/make_golden.py -> tests/golden/lsm_*.npz and tests/test_oracle_lsm.py.

Shapes: cap [Bc,T,D], attention_mask / special_tokens_mask [Bc,T], region_features [Bi,Rg,V],
region_mask [Bi,Rg], weight [D,V], bias [D].  The pair matrix has rows = captions, cols = images
(grounding_head.py:119-144: flat pair index = caption * B + image).
"""
import torch
import torch.nn.functional as F


def caption_mask_of(attention_mask, special_tokens_mask):
    # grounding_head.py:94-96,101
    return (attention_mask * (1 - special_tokens_mask)).to(torch.float32)


def project_regions(region_features, weight, bias):
    # grounding_head.py:111  (v2l_projection; the permute is layout only)
    return F.linear(region_features, weight, bias)


def pair_distances(cap, cap_mask, emb, reg_mask, temperature, alignment="softmax", dtype=torch.float32,
                   chunk=8):
    """Returns (d_w2r, d_r2w), each [Bc, Bi] — grounding_head.py:116-256 without B^2 replication.

    The masked fill uses ``min(S) - 100`` over the whole (unmasked) similarity tensor exactly as the
    reference does (grounding_head.py:156-160); the empty-pair guard (``:240-251``) is applied too.
    """
    cap = cap.to(dtype)
    emb = emb.to(dtype)
    mc = cap_mask.to(dtype)
    mr = reg_mask.to(dtype)
    Bc, T, _ = cap.shape
    Bi, Rg, _ = emb.shape
    n_words = mc.sum(1)
    n_regions = mr.sum(1)
    # global min of S over everything (computed in chunks to bound memory)
    gmin = None
    for c0 in range(0, Bc, chunk):
        S = torch.einsum("ctd,ird->citr", cap[c0:c0 + chunk], emb) / temperature
        m = S.min()
        gmin = m if gmin is None else torch.minimum(gmin, m)
    fill = gmin - 100.0
    d_w2r = torch.empty(Bc, Bi, dtype=dtype)
    d_r2w = torch.empty(Bc, Bi, dtype=dtype)
    for c0 in range(0, Bc, chunk):
        S = torch.einsum("ctd,ird->citr", cap[c0:c0 + chunk], emb) / temperature        # [c,i,T,Rg]
        valid = (mc[c0:c0 + chunk, None, :, None] * mr[None, :, None, :]) > 0
        Sm = torch.where(valid, S, fill)
        if alignment == "softmax":
            A_w2r = F.softmax(Sm, dim=3)
            A_r2w = F.softmax(Sm, dim=2)
        elif alignment == "hardmax":
            A_w2r = F.one_hot(Sm.argmax(dim=3), Rg).to(dtype)
            A_r2w = F.one_hot(Sm.argmax(dim=2), T).to(dtype).permute(0, 1, 3, 2)
        else:
            raise NotImplementedError(alignment)
        A_w2r = A_w2r * mc[c0:c0 + chunk, None, :, None]
        A_r2w = A_r2w * mr[None, :, None, :]
        D = -S
        d_w2r[c0:c0 + chunk] = (A_w2r * D).sum(3).sum(2) / torch.clamp(n_words[c0:c0 + chunk, None], min=1.0)
        d_r2w[c0:c0 + chunk] = (A_r2w * D).sum(3).sum(2) / torch.clamp(n_regions[None, :], min=1.0)
    ok = ((n_words[:, None] > 0) | (n_regions[None, :] > 0))
    d_w2r = torch.where(ok, d_w2r, d_w2r.max() + 100.0)
    d_r2w = torch.where(ok, d_r2w, d_r2w.max() + 100.0)
    return d_w2r, d_r2w


def pair_losses(pw):
    """pw [B,B] (rows captions, cols images) -> (CE choose caption, CE choose image, acc caption, acc image)
    grounding_head.py:272-290, 354-379."""
    ce_cap = torch.diag(-torch.log_softmax(-pw, dim=0)).mean()
    ce_img = torch.diag(-torch.log_softmax(-pw, dim=1)).mean()
    ar = torch.arange(pw.shape[0])
    acc_cap = (pw.argmin(dim=0) == ar).to(torch.float32).mean()
    acc_img = (pw.argmin(dim=1) == ar).to(torch.float32).mean()
    return ce_cap, ce_img, acc_cap, acc_img


def triplet_losses(pw, margin=1.0, negative_mining="hardest"):
    """grounding_head.py:292-350 for the deterministic mining modes."""
    n = pw.shape[0]
    pos = torch.diag(pw)
    if n < 2:
        neg_cap = neg_img = pos + margin
    else:
        off = ~torch.eye(n, dtype=torch.bool)
        if negative_mining == "hardest":
            neg_cap = pw.masked_fill(~off, float("inf")).min(dim=0).values
            neg_img = pw.masked_fill(~off, float("inf")).min(dim=1).values
        elif negative_mining == "easiest":
            neg_cap = pw.masked_fill(~off, float("-inf")).max(dim=0).values
            neg_img = pw.masked_fill(~off, float("-inf")).max(dim=1).values
        else:
            raise NotImplementedError(negative_mining)
    return F.relu(pos - neg_cap + margin).mean(), F.relu(pos - neg_img + margin).mean()


def grounding_head_forward(input_image, input_caption, weight, bias, temperature=10.0, alignment="softmax",
                           text_input="input_embeddings", dtype=torch.float32, loss="cross_entropy",
                           align_words=True, align_regions=True, margin=1.0, negative_mining="hardest"):
    """Restated GroundingHead.forward (aligned_local, dot).  Returns (other_info, losses, {"w2r","r2w"})
    with the reference's key strings."""
    cap = input_caption[text_input]
    mc = caption_mask_of(input_caption["attention_mask"], input_caption["special_tokens_mask"])
    emb = project_regions(input_image["region_features"].to(dtype), weight.to(dtype), bias.to(dtype))
    w2r, r2w = pair_distances(cap, mc, emb, input_image["region_mask"], temperature, alignment, dtype)
    losses, info, dists = {}, {}, {}
    for key, name, pw, on in (("w2r", "Words", w2r, align_words), ("r2w", "Regions", r2w, align_regions)):
        if not on:
            continue
        dists[key] = pw
        ce_cap, ce_img, acc_cap, acc_img = pair_losses(pw)
        if loss == "cross_entropy":
            losses[f"CE_loss (Align {name}, Choose Caption)"] = ce_cap
            losses[f"CE_loss (Align {name}, Choose Image)"] = ce_img
        elif loss == "triplet":
            t_cap, t_img = triplet_losses(pw, margin, negative_mining)
            losses[f"Triplet Loss (Align {name}, Choose Caption)"] = t_cap
            losses[f"Triplet Loss (Align {name}, Choose Image)"] = t_img
        else:
            raise NotImplementedError(loss)
        info[f"Batch Accuracy (Align {name}, Choose Caption)"] = acc_cap
        info[f"Batch Accuracy (Align {name}, Choose Image)"] = acc_img
    return info, losses, dists


def _log_stats(t):
    """What LoggedModule.log costs per call in the reference (logged_module.py:8-17): a host copy of the
    tensor plus four scalar reductions."""
    t.cpu().detach().numpy()
    return float(t.min()), float(t.max()), float(t.to(torch.float32).mean()), float(t.to(torch.float32).std())


def grounding_head_forward_literal(input_image, input_caption, weight, bias, temperature=10.0, log=True):
    """The reference's own evaluation order for the shipped configuration (softmax / aligned_local /
    cross_entropy), INCLUDING the B^2 replication and torch.bmm of grounding_head.py:116-147 — this is
    what bench.py times as the reference CPU arm; `grounding_head_forward` above is the memory-lean
    closed form used as the checker.  Returns the same triple."""
    cap = input_caption["input_embeddings"]
    mc = caption_mask_of(input_caption["attention_mask"], input_caption["special_tokens_mask"])
    mr = input_image["region_mask"].to(torch.float32)
    lg = _log_stats if log else (lambda t: None)
    for t in (input_caption["attention_mask"], input_caption["special_tokens_mask"], mc, cap):     # :97-100
        lg(t)
    n_words, n_regions = mc.sum(1), mr.sum(1)
    B, R, _ = input_image["region_features"].shape
    T = mc.shape[1]
    D = weight.shape[0]
    img = F.linear(input_image["region_features"], weight, bias).permute(0, 2, 1)          # [B, D, R]
    for t in (input_image["region_features"], mr, img):                                     # :112-114
        lg(t)
    img = img.unsqueeze(0).repeat(B, 1, 1, 1).reshape(B * B, D, R)                          # pair p = c*B + i
    capr = cap.unsqueeze(1).repeat(1, B, 1, 1).reshape(B * B, T, D)
    mr2 = mr.unsqueeze(0).repeat(B, 1, 1).reshape(B * B, R)
    mc2 = mc.unsqueeze(1).repeat(1, B, 1).reshape(B * B, T)
    nr2 = n_regions.unsqueeze(0).repeat(B, 1).reshape(B * B)
    nw2 = n_words.unsqueeze(1).repeat(1, B).reshape(B * B)
    sim = torch.bmm(capr, img) / temperature
    dist = -sim
    lg(sim)                                                                                 # :155
    sim = torch.where((mc2[:, :, None] * mr2[:, None, :]) > 0, sim, sim.min() - 100.0)
    a_w2r = F.softmax(sim, dim=2) * mc2[:, :, None]
    lg(F.softmax(sim, dim=2))                                                               # :208,210 (pre-mask attention)
    lg(F.softmax(sim, dim=1))
    a_r2w = F.softmax(sim, dim=1) * mr2[:, None, :]
    d_w2r = (a_w2r * dist).sum(2).sum(1) / torch.clamp(nw2, min=1.0)
    d_r2w = (a_r2w * dist).sum(2).sum(1) / torch.clamp(nr2, min=1.0)
    ok = (nw2 > 0) | (nr2 > 0)
    d_w2r = torch.where(ok, d_w2r, d_w2r.max() + 100.0).reshape(B, B)
    d_r2w = torch.where(ok, d_r2w, d_r2w.max() + 100.0).reshape(B, B)
    lg(d_w2r)                                                                               # :254,256
    lg(d_r2w)
    losses, info = {}, {}
    for name, pw in (("Words", d_w2r), ("Regions", d_r2w)):
        ce_cap, ce_img, acc_cap, acc_img = pair_losses(pw)
        losses[f"CE_loss (Align {name}, Choose Caption)"] = ce_cap
        losses[f"CE_loss (Align {name}, Choose Image)"] = ce_img
        info[f"Batch Accuracy (Align {name}, Choose Caption)"] = acc_cap
        info[f"Batch Accuracy (Align {name}, Choose Image)"] = acc_img
    return info, losses, {"w2r": d_w2r, "r2w": d_r2w}


def make_lsm_inputs(B, Rg, T, V=2048, D=768, seed=1992, ragged_regions=False, min_words=6, empty_caption=None,
                    empty_image=None, gain=1.0):
    """Synthetic inputs of SURVEY.md §8(d): randn*0.05 caption embeddings, caption lengths uniform in
    [min_words, T] with CLS/SEP/PAD flagged special, region mask all ones (boxes) or ragged (grid)."""
    g = torch.Generator().manual_seed(seed)
    cap = torch.randn(B, T, D, generator=g) * 0.05 * gain      # gain > 1: well-separated pair distances
    att = torch.zeros(B, T, dtype=torch.int64)
    spe = torch.zeros(B, T, dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(min(min_words, T), T + 1, (1,), generator=g))
        att[b, :n] = 1
        spe[b, 0] = 1                      # [CLS]
        spe[b, n - 1] = 1                  # [SEP]
        spe[b, n:] = 1                     # [PAD] (tokenizer flags pads as special)
    if empty_caption is not None:
        att[empty_caption] = 0
        spe[empty_caption] = 1
    feats = torch.randn(B, Rg, V, generator=g)
    rmask = torch.ones(B, Rg, dtype=torch.uint8)
    if ragged_regions:
        for b in range(B):
            n = int(torch.randint(max(1, Rg // 3), Rg + 1, (1,), generator=g))
            rmask[b, n:] = 0
            feats[b, n:] = 0               # pad_sequence zero-pads (distill_prop_mmss_gcnn.py:313-318)
    if empty_image is not None:
        rmask[empty_image] = 0
        feats[empty_image] = 0
    weight = torch.randn(D, V, generator=g) * 0.01 * gain
    bias = torch.randn(D, generator=g) * 0.01
    input_image = {"region_features": feats, "region_mask": rmask}
    input_caption = {"input_embeddings": cap, "attention_mask": att, "special_tokens_mask": spe}
    return input_image, input_caption, weight, bias
