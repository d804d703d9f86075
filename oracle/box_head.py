"""ORACLE (test infrastructure only): CPU restatement of the embedding box predictor.

Follows /root/reference/ovr/modeling/roi_heads/box_emb_head.py:
  * forward / forward_cls_prediction   :179-212
  * set_class_embeddings               :214-236
  * normalize_vec / standardize_vec    /root/reference/ovr/modeling/logged_module.py:55-72
and the Detectron2 behaviour the class inherits (un-vendored, version unpinned):
FastRCNNOutputLayers.losses (cross_entropy mean over all R; smooth-L1 over fg / R) and
.predict_probs / fast_rcnn_inference (softmax, drop background column, per-RoI argmax).
PINNED (round 2): the reference class IS importable once the handful of Detectron2 symbols it touches are stubbed
(oracle/d2_stubs.py restates them; oracle/ref_loader.load_reference_box_head executes the reference's own
box_emb_head.py unmodified).  tests/golden/box_*.npz hold the outputs of that real class — forward, probabilities,
losses, gradients, inference, normalise / standardise variants, re-set class embeddings, detached classifier — and
tests/test_oracle_box.py checks this restatement against every one of them (and against the live class when
/root/reference is present).  The reference itself ships no tests or golden vectors for this path.
"""
import torch
import torch.nn.functional as F


def normalize_vec(x, dim=1):
    return F.normalize(x, p=2, dim=dim)                                   # logged_module.py:55-65


def standardize_vec(x, dim=1):
    return (x - x.mean(dim, keepdim=True)) / (x.std(dim, keepdim=True) + 1e-12)   # logged_module.py:68-72


def prepare_class_embeddings(embs, normalize=False, standardize=False):
    """set_class_embeddings (box_emb_head.py:214-236): returns (cls_weight [K+1,D], cls_bias zeros)."""
    embs = torch.as_tensor(embs, dtype=torch.float32).clone()
    if normalize:
        embs = normalize_vec(embs, dim=1)
    if standardize:
        embs = standardize_vec(embs, dim=1)
    return embs, torch.zeros(embs.shape[0], dtype=torch.float32)


def box_predictor_forward(x, w_emb, b_emb, w_cls, b_cls, w_box, b_box, normalize=False, standardize=False,
                          dtype=torch.float32):
    """(scores [R,K+1], deltas [R,4], emb [R,D]) — box_emb_head.py:179-212."""
    x = x.to(dtype)
    if x.dim() > 2:
        x = torch.flatten(x, start_dim=1)
    deltas = F.linear(x, w_box.to(dtype), b_box.to(dtype))
    e = F.linear(x, w_emb.to(dtype), b_emb.to(dtype))
    if normalize:
        e = normalize_vec(e, dim=1)
    if standardize:
        e = standardize_vec(e, dim=1)
    scores = F.linear(e, w_cls.to(dtype), b_cls.to(dtype))
    return scores, deltas, e


def box_losses(scores, deltas, gt_classes, proposal_boxes=None, gt_boxes=None,
               bbox_reg_weights=(10.0, 10.0, 5.0, 5.0), smooth_l1_beta=0.0, loss_weight_cls=1.0,
               loss_weight_box=1.0):
    """Detectron2 FastRCNNOutputLayers.losses semantics."""
    K = scores.shape[1] - 1
    R = scores.shape[0]
    loss_cls = F.cross_entropy(scores, gt_classes, reduction="mean") if R > 0 else scores.sum() * 0.0
    out = {"loss_cls": loss_cls * loss_weight_cls}
    if proposal_boxes is not None:
        fg = (gt_classes >= 0) & (gt_classes < K)
        tgt = get_deltas(proposal_boxes[fg], gt_boxes[fg], bbox_reg_weights)
        d = deltas[fg]
        if smooth_l1_beta < 1e-5:
            l = (d - tgt).abs().sum()
        else:
            n = (d - tgt).abs()
            l = torch.where(n < smooth_l1_beta, 0.5 * n ** 2 / smooth_l1_beta, n - 0.5 * smooth_l1_beta).sum()
        out["loss_box_reg"] = l / max(R, 1) * loss_weight_box
    return out


def get_deltas(src, dst, weights):
    """Detectron2 Box2BoxTransform.get_deltas."""
    wx, wy, ww, wh = weights
    sw = src[:, 2] - src[:, 0]
    sh = src[:, 3] - src[:, 1]
    sx = src[:, 0] + 0.5 * sw
    sy = src[:, 1] + 0.5 * sh
    tw = dst[:, 2] - dst[:, 0]
    th = dst[:, 3] - dst[:, 1]
    tx = dst[:, 0] + 0.5 * tw
    ty = dst[:, 1] + 0.5 * th
    return torch.stack([wx * (tx - sx) / sw, wy * (ty - sy) / sh, ww * torch.log(tw / sw), wh * torch.log(th / sh)], 1)


def predict_probs(scores):
    """softmax over K+1 columns; per-RoI argmax over the K foreground columns (Appendix C)."""
    probs = F.softmax(scores, dim=-1)
    return probs, probs[:, :-1].argmax(dim=1)


def make_box_inputs(R, K, V=2048, D=768, seed=1992, bg_frac=0.25):
    """Synthetic inputs of SURVEY.md §8(d)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(R, V, generator=g)
    w_emb = torch.randn(D, V, generator=g) * 0.01
    b_emb = torch.zeros(D)
    w_box = torch.randn(4, V, generator=g) * 0.001
    b_box = torch.zeros(4)
    cls = torch.cat([torch.randn(K, D, generator=g) * 0.05, torch.zeros(1, D)], 0)   # zero background row
    gt = torch.randint(0, K, (R,), generator=g)
    gt[torch.rand(R, generator=g) < bg_frac] = K
    return x, w_emb, b_emb, w_box, b_box, cls, gt


def make_proposals(n_img, per_img, K, seed=1992, image_size=(320, 480), bg_frac=0.25):
    """Seeded proposals with labels for Detectron2's ``losses`` / ``inference``: per image a dict
    {proposal_boxes [Ri,4], gt_boxes [Ri,4], gt_classes [Ri] int64 in [0,K] (K = background)}, boxes inside the image,
    at least 8 px wide / high (SURVEY.md §8d box distribution, scaled to the image)."""
    g = torch.Generator().manual_seed(seed)
    h, w = image_size
    out = []
    for _ in range(n_img):
        cx = torch.rand(per_img, generator=g) * w
        cy = torch.rand(per_img, generator=g) * h
        s = 16.0 * (min(h, w) / 16.0) ** torch.rand(per_img, generator=g)
        a = 0.5 * 4.0 ** torch.rand(per_img, generator=g)
        bw, bh = s * a.sqrt(), s / a.sqrt()
        x1 = (cx - bw / 2).clamp(0, w - 9)
        y1 = (cy - bh / 2).clamp(0, h - 9)
        x2 = torch.minimum((cx + bw / 2).clamp(0, w), x1 + bw).clamp_min(0)
        y2 = torch.minimum((cy + bh / 2).clamp(0, h), y1 + bh).clamp_min(0)
        x2 = torch.maximum(x2, x1 + 8)
        y2 = torch.maximum(y2, y1 + 8)
        prop = torch.stack([x1, y1, x2, y2], 1)
        gtb = prop + torch.randn(per_img, 4, generator=g) * 2.0
        gtb[:, 2:] = torch.maximum(gtb[:, 2:], gtb[:, :2] + 4)
        gt = torch.randint(0, K, (per_img,), generator=g)
        gt[torch.rand(per_img, generator=g) < bg_frac] = K
        out.append({"proposal_boxes": prop, "gt_boxes": gtb, "gt_classes": gt})
    return out


def instances_from(props, image_size, Instances, Boxes, device=None):
    """Wrap ``make_proposals`` output in an Instances / Boxes implementation (Detectron2's, the oracle's restatement
    or the drop-in package's stand-ins)."""
    mv = (lambda t: t.to(device)) if device is not None else (lambda t: t)
    return [Instances(image_size, proposal_boxes=Boxes(mv(p["proposal_boxes"])), gt_boxes=Boxes(mv(p["gt_boxes"])),
                      gt_classes=mv(p["gt_classes"])) for p in props]
