"""ORACLE (test infrastructure only): CPU restatement of the embedding box predictor.

Follows /root/reference/ovr/modeling/roi_heads/box_emb_head.py:
  * forward / forward_cls_prediction   :179-212
  * set_class_embeddings               :214-236
  * normalize_vec / standardize_vec    /root/reference/ovr/modeling/logged_module.py:55-72
and the Detectron2 behaviour the class inherits (un-vendored, version unpinned):
FastRCNNOutputLayers.losses (cross_entropy mean over all R; smooth-L1 over fg / R) and
.predict_probs / fast_rcnn_inference (softmax, drop background column, per-RoI argmax).
The reference class cannot be imported without Detectron2 and the reference holds no golden vectors
for it: parity unpinned beyond torch's own F.linear / F.cross_entropy / F.softmax primitives.
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
