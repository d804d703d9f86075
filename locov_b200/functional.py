"""torch.autograd Functions over the C-ABI entry points (include/locov_b200.h).

Every forward AND backward computation is a call into liblocov_b200.so; torch is used for memory,
streams and the autograd graph only.  ``precision`` selects the tensor-core operand mode:
  "fp32" : fp32-accurate — operands split into bf16 hi+lo, three tcgen05 passes (≈1e-5 relative)
  "bf16" : single bf16 pass (≈4e-3 relative; the north-star tolerance for bf16 is 2e-2)
"""
import os
import threading
import weakref
from typing import Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops
from ._lib import LocoError

PRECISIONS = ("fp32", "bf16")

# How the reduced-precision ("bf16") mode multiplies fp32 activations by fp32 weights:
#   "tf32": kind::tf32 straight from the fp32 operands (no conversion pass; L2 -> SM operand traffic bound) — default
#   "bf16": one conversion pass (fp32 -> bf16, HBM bound) + a bf16 tcgen05 GEMM (half the operand bytes through L2 -> SM,
#           twice the tensor rate).  Measured on B200 (profiles/README.md): the GEMM gets faster but the extra launch and
#           pass cancel it at the BASELINE shapes (LSM step 71.8 vs 68.2 us, config-5 box chain 216 vs 225 us)
PROJECTION = os.environ.get("LOCOV_B200_PROJECTION", "tf32")


def _use_tf32(acc: bool, k: int) -> bool:
    return (not acc) and PROJECTION == "tf32" and k % 4 == 0


def _acc(precision: str) -> bool:
    if precision not in PRECISIONS:
        raise LocoError(f"precision must be one of {PRECISIONS}, got {precision!r}")
    return precision == "fp32"


# ------------------------------------------------------------------------------------------------
# weight operand cache: bf16 (hi, lo) shadows of parameters, refreshed when the parameter changes.
# Parameters are aliased between modules (v2l_projection <-> emb_pred weight tying), so the cache is
# keyed on storage identity + version counter, never on the module.
# ------------------------------------------------------------------------------------------------
# Freshness rule: a shadow is served only while (a) the SAME tensor object is alive, (b) its autograd version counter is
# unchanged and (c) the tensor does not require grad.  (c) closes the hole the version counter leaves: an in-place update
# through ``param.data`` (EMA helpers, hand-written SGD, weight surgery) does NOT bump ``param._version``; a trainable
# parameter changes every step anyway, so its shadow is simply re-converted on every call (one 3 us HBM pass for the
# 768 x 2048 projection).  Frozen tensors (requires_grad False: the class-embedding matrix, FREEZE_EMB_PRED weights) keep
# their shadow; after editing one of those through ``.data`` call ``clear_weight_cache()``, or set
# LOCOV_B200_WEIGHT_CACHE=0 to disable caching altogether.
_wcache = {}
WEIGHT_CACHE = os.environ.get("LOCOV_B200_WEIGHT_CACHE", "1") != "0"


def weight_operand(w: torch.Tensor, accurate: bool, transpose: bool = False, tag: str = "") -> ops.Bf16Operand:
    key = (w.data_ptr(), tuple(w.shape), tuple(w.stride()), accurate, transpose, tag, w.device.index)
    ent = _wcache.get(key)
    ver = w._version
    # identity check through a weak reference: a freed parameter's address (and version count) can be
    # re-used by a new tensor, which must never be served the old shadow.
    if WEIGHT_CACHE and not w.requires_grad and ent is not None and ent[0] == ver and ent[2]() is w:
        return ent[1]
    reuse = ent[1] if (ent is not None and ent[2]() is w and not torch.cuda.is_current_stream_capturing()) else None
    op = ops.split_bf16(w.detach(), accurate, transpose=transpose, out=reuse)
    _wcache[key] = (ver, op, weakref.ref(w))
    if len(_wcache) > 256:      # parameters re-created by set_class_embeddings: drop dead / oldest shadows
        for k in [k for k, v in _wcache.items() if v[2]() is None] or list(_wcache)[:128]:
            del _wcache[k]
    return op


def weight_operand_if_cached(w: torch.Tensor, accurate: bool, transpose: bool = False, tag: str = ""):
    """The cached operand of ``w`` when ``weight_operand`` would serve it without a conversion launch, else None."""
    ent = _wcache.get((w.data_ptr(), tuple(w.shape), tuple(w.stride()), accurate, transpose, tag, w.device.index))
    if WEIGHT_CACHE and not w.requires_grad and ent is not None and ent[0] == w._version and ent[2]() is w:
        return ent[1]
    return None


def adopt_weight_operand(w: torch.Tensor, accurate: bool, op: ops.Bf16Operand, transpose: bool = False, tag: str = ""):
    """Register an operand of ``w`` that another launch produced (ops.lsm_prep with ``extra``) under weight_operand's rules."""
    _wcache[(w.data_ptr(), tuple(w.shape), tuple(w.stride()), accurate, transpose, tag, w.device.index)] = (w._version, op, weakref.ref(w))


def clear_weight_cache():
    """Drop every cached bf16 weight shadow (call after modifying a FROZEN weight through ``.data``)."""
    _wcache.clear()
    _cat_cache.clear()


# ------------------------------------------------------------------------------------------------
# RoIAlign
# ------------------------------------------------------------------------------------------------
class _RoIAlign(Function):
    @staticmethod
    def forward(ctx, feat, rois, ph, pw, scale, sampling_ratio, aligned, channels_last, out_dtype):
        ctx.save_for_backward(rois)
        ctx.cfg = (tuple(feat.shape), scale, sampling_ratio, aligned)
        return ops.roi_align(feat, rois, (ph, pw), scale, sampling_ratio, aligned, channels_last, out_dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (rois,) = ctx.saved_tensors
        shape, scale, sampling_ratio, aligned = ctx.cfg
        dfeat = ops.roi_align_backward(dout, shape, rois, scale, sampling_ratio, aligned)
        return dfeat, None, None, None, None, None, None, None, None


def roi_align(feat, rois, output_size, spatial_scale, sampling_ratio=0, aligned=True, channels_last=False, out_dtype=torch.float32):
    """Differentiable RoIAlign (torchvision.ops.roi_align semantics; reference roi_emb_heads.py:243-245).

    ``channels_last`` / ``out_dtype``: the pooled [R,C,PH,PW] tensor in ``torch.channels_last`` memory format, fp32 or bf16 —
    the layout cuDNN's tensor-core convolutions of the res5 stage take without a transpose."""
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    return _RoIAlign.apply(feat, rois, int(ph), int(pw), float(spatial_scale), int(sampling_ratio), bool(aligned),
                           bool(channels_last), out_dtype)


class _TokenPool(Function):
    @staticmethod
    def forward(ctx, raw, seg_off, inv_temp, hardmax):
        scores, _ = ops.token_pool(raw, seg_off, inv_temp, hardmax)
        ctx.save_for_backward(raw, seg_off)
        ctx.meta = (inv_temp, hardmax)
        return scores

    @staticmethod
    @once_differentiable
    def backward(ctx, dscores):
        raw, seg_off = ctx.saved_tensors
        inv_temp, hardmax = ctx.meta
        return ops.token_pool_backward(raw, seg_off, inv_temp, hardmax, dscores), None, None, None


def grounding_scores(e, w_tok, seg_off, temperature: float, alignment: str = "softmax", precision: str = "fp32"):
    """Multi-token class scores (reference box_emb_grounding_head.py:89-221): token logits e . w_tok^T on the tensor-core GEMM core,
    then the per-class attention pooling; differentiable w.r.t. e."""
    if alignment not in ("softmax", "hardmax"):
        raise LocoError(f"grounding_scores: alignment {alignment!r}")
    raw = linear(e, w_tok, None, precision)
    return _TokenPool.apply(raw.contiguous(), seg_off, 1.0 / float(temperature), alignment == "hardmax")


_mean_tls = threading.local()     # forward's bf16 operand, handed to spatial_mean() below (a Function returns tensors only)


class _SpatialMean(Function):
    @staticmethod
    def forward(ctx, x, operand):
        ctx.meta = (tuple(x.shape), x.dtype, (not x.is_contiguous()) and x.is_contiguous(memory_format=torch.channels_last))
        out, _mean_tls.operand = ops.spatial_mean(x, operand)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        shape, dtype, cl = ctx.meta
        return ops.spatial_mean_backward(dy, shape, dtype, cl), None


def spatial_mean(x, operand_precision: Optional[str] = None):
    """``x.mean(dim=[2, 3])`` of the res5 output (reference roi_emb_heads.py:262, :329, :351) as one pass over x.

    With ``operand_precision`` ("fp32" / "bf16" — the precision of the box predictor that consumes the result) the same pass also
    writes the bf16 operand of the predictor's projection GEMM; it rides on the returned tensor as ``_loco_operand`` and
    ``box_predict`` picks it up instead of re-reading the means in a split kernel."""
    operand = None if operand_precision is None else _acc(operand_precision)
    if operand is not None and _use_tf32(operand, x.shape[1]):
        operand = None                       # the TF32 projection reads the fp32 means directly
    out = _SpatialMean.apply(x, operand)
    op = getattr(_mean_tls, "operand", None)
    _mean_tls.operand = None
    if op is not None:
        out._loco_operand = (op, operand, out._version)
    return out


# ------------------------------------------------------------------------------------------------
# y = x W^T + b on the tcgen05 core
# ------------------------------------------------------------------------------------------------
class _Linear(Function):
    @staticmethod
    def forward(ctx, x, w, b, precision):
        acc = _acc(precision)
        x2 = x.reshape(-1, x.shape[-1])
        if not _use_tf32(acc, x2.shape[1]):
            out, _ = ops.linear_fwd(ops.split_bf16(x2, acc), weight_operand(w, acc), b, want_f32=True)
        else:       # reduced-precision mode: fp32 operands in place as TF32, no split pass
            out, _ = ops.linear_tf32_fwd(x2, w.detach(), b, want_f32=True)
        ctx.save_for_backward(x2, w)
        ctx.meta = (precision, b is not None, x.shape)
        return out.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        precision, has_b, xshape = ctx.meta
        acc = _acc(precision)
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            g_op = ops.split_bf16(dy2, acc)
            wt_op = weight_operand(w, acc, transpose=True)
            dx, _ = ops.linear_fwd(g_op, wt_op, None, want_f32=True)
            dx = dx.reshape(xshape)
        if ctx.needs_input_grad[1]:
            gt_op = ops.split_bf16(dy2, acc, transpose=True)
            xt_op = ops.split_bf16(x2, acc, transpose=True)
            dw, _ = ops.linear_fwd(gt_op, xt_op, None, want_f32=True)
        if has_b and ctx.needs_input_grad[2]:
            db = dy2.sum(0)
        return dx, dw, db, None


def linear(x, w, b=None, precision="fp32"):
    return _Linear.apply(x, w, b, precision)


# ------------------------------------------------------------------------------------------------
# box predictor: x -> (scores [R,K+1], deltas [R,4]) with fused softmax statistics
# ------------------------------------------------------------------------------------------------
class BoxScoreAux:
    """Side outputs of the fused scoring epilogue (never differentiated)."""
    __slots__ = ("lse", "probs", "argmax_fg")

    def __init__(self, lse, probs, argmax_fg):
        self.lse, self.probs, self.argmax_fg = lse, probs, argmax_fg


class _BoxPredict(Function):
    """emb|deltas = x·[W_emb; W_box]^T + [b_emb; b_box];  scores = emb·W_cls^T + b_cls.
    reference box_emb_head.py:179-212 (bbox_pred, emb_pred, cls_score as three cuBLAS GEMMs)."""

    @staticmethod
    def forward(ctx, x, w_emb, b_emb, w_box, b_box, w_cls, b_cls, precision, want_probs, aux_box, x_op=None):
        acc = _acc(precision)
        d = w_emb.shape[0]
        nbox = w_box.shape[0]
        w_cat = _cat_weight(w_emb, w_box)
        b_cat = torch.cat([b_emb.detach(), b_box.detach()]).to(torch.float32)
        if not _use_tf32(acc, x.shape[1]):
            # the producer of x (spatial_mean) may already have written its bf16 operand in the same pass
            a_op = x_op if x_op is not None else ops.split_bf16(x, acc)
            wcat_op = weight_operand(w_cat, acc, tag="cat")
            out, e_op = ops.linear_fwd(a_op, wcat_op, b_cat, want_f32=True, n_bf16=d, accurate_out=acc)
        else:       # reduced-precision mode: TF32 projection straight from the fp32 activations (no split pass)
            out, e_op = ops.linear_tf32_fwd(x, w_cat, b_cat, want_f32=True, n_bf16=d)
        deltas = out[:, d:d + nbox]
        cls_op = weight_operand(w_cls, acc)
        logits, probs, lse, arg = ops.box_score(e_op, cls_op, b_cls, want_probs)
        aux_box.append(BoxScoreAux(lse, probs, arg))
        ctx.save_for_backward(x, w_emb, w_box, w_cls)
        ctx.meta = (precision, d, nbox)
        return logits, deltas.contiguous()

    @staticmethod
    @once_differentiable
    def backward(ctx, dscores, ddeltas):
        x, w_emb, w_box, w_cls = ctx.saved_tensors
        precision, d, nbox = ctx.meta
        acc = _acc(precision)
        r = x.shape[0]
        dev = x.device
        need_x, need_we, need_be, need_wb, need_bb = ctx.needs_input_grad[:5]
        dx = dwe = dbe = dwb = dbb = None
        # The upstream gradient of [emb | deltas] as ONE bf16 operand [R, pad8(D + 4)]: dE = dscores . W_cls lands in its first D columns
        # straight from the GEMM epilogue (no fp32 round trip, no separate split pass), the box gradient goes into the last 4.
        ldg = (d + nbox + 7) // 8 * 8
        have_de = dscores is not None and (need_x or need_we or need_be)
        mk = torch.empty if have_de else torch.zeros              # (the GEMM below overwrites the first D columns of every row)
        g_hi = mk((r, ldg), dtype=torch.bfloat16, device=dev)
        g_lo = mk((r, ldg), dtype=torch.bfloat16, device=dev) if acc else None
        de = None
        if have_de:
            gs_op = ops.split_bf16(dscores.contiguous(), acc)
            clst_op = weight_operand(w_cls, acc, transpose=True)
            de = ops.linear_fwd_into(gs_op, clst_op, None, want_f32=need_we or need_be, out_hi=g_hi, out_lo=g_lo, n_bf16=d)
            g_hi[:, d:].zero_()
            if acc:
                g_lo[:, d:].zero_()
        if ddeltas is not None:
            dd = ddeltas.to(torch.float32)
            g_hi[:, d:d + nbox] = dd
            if acc:
                g_lo[:, d:d + nbox] = dd - g_hi[:, d:d + nbox].to(torch.float32)
        g_op = ops.Bf16Operand(g_hi, g_lo, r, d + nbox)
        if need_x:
            w_cat = _cat_weight(w_emb, w_box)
            wt_op = weight_operand(w_cat, acc, transpose=True, tag="cat")
            dx, _ = ops.linear_fwd(g_op, wt_op, None, want_f32=True)
        if need_we:
            # trainable emb_pred (LSM stage): the full [D+4, V] weight gradient as a tensor-core GEMM over transposed operands
            xt_op = ops.split_bf16(x, acc, transpose=True)
            dw_cat, _ = ops.linear_fwd(ops.transpose_operand(g_op), xt_op, None, want_f32=True)
            dwe = dw_cat[:d]
            if need_wb:
                dwb = dw_cat[d:d + nbox]
            if need_bb and ddeltas is not None:
                dbb = ddeltas.to(torch.float32).sum(0)
        elif (need_wb or need_bb) and ddeltas is not None:
            # only the class-agnostic box regressor trains (FREEZE_EMB_PRED, coco_stt.yaml:36): 4 x V gradient, x streamed once
            dwb, dbb = ops.skinny_grad(ddeltas, x, want_bias=True)
        if need_be and de is not None:
            dbe = de.sum(0)
        if not need_wb:
            dwb = None
        if not need_bb:
            dbb = None
        return dx, dwe, dbe, dwb, dbb, None, None, None, None, None, None


_cat_cache = {}


def _cat_weight(w_emb, w_box):
    """[W_emb; W_box] as one [D+4, V] fp32 matrix.  Each part is re-copied only when ITS parameter may have changed (version counter
    moved, or the parameter is trainable — see the freshness rule above): with FREEZE_EMB_PRED the 768 x 2048 block is copied once and
    a training step refreshes only the 4 x 2048 rows of bbox_pred."""
    key = (w_emb.data_ptr(), w_box.data_ptr(), tuple(w_emb.shape), tuple(w_box.shape))
    ent = _cat_cache.get(key)
    alive = WEIGHT_CACHE and ent is not None and ent[2]() is w_emb and ent[3]() is w_box
    d = w_emb.shape[0]
    if alive:
        cat, (ve, vb) = ent[1], ent[0]
        if w_emb.requires_grad or ve != w_emb._version:
            cat[:d].copy_(w_emb.detach())       # in place: bumps cat._version -> the bf16 shadow of `cat` refreshes
        if w_box.requires_grad or vb != w_box._version:
            cat[d:].copy_(w_box.detach())
    else:
        cat = torch.cat([w_emb.detach(), w_box.detach()], 0).to(torch.float32).contiguous()
    _cat_cache[key] = ((w_emb._version, w_box._version), cat, weakref.ref(w_emb), weakref.ref(w_box))
    if len(_cat_cache) > 64:
        for k in [k for k, v in _cat_cache.items() if v[2]() is None] or list(_cat_cache)[:32]:
            del _cat_cache[k]
    return cat


def box_predict(x, w_emb, b_emb, w_box, b_box, w_cls, b_cls=None, precision="fp32", want_probs=True):
    """Returns (scores, deltas, BoxScoreAux)."""
    aux = []
    x_op = None
    carried = getattr(x, "_loco_operand", None)          # written by spatial_mean() in the pass that produced x
    if carried is not None and carried[1] == _acc(precision) and carried[2] == x._version and carried[0].rows == x.shape[0] and carried[0].cols == x.shape[1]:
        x_op = carried[0]
    scores, deltas = _BoxPredict.apply(x, w_emb, b_emb, w_box, b_box, w_cls, b_cls, precision, want_probs, aux, x_op)
    return scores, deltas, aux[0]


class _BoxScore(Function):
    """scores = e · W_cls^T + b_cls with the fused softmax statistics, for embeddings that are NOT the direct output of the
    projection GEMM (a row-wise normalisation sits in between: box_emb_head.py:207-211)."""

    @staticmethod
    def forward(ctx, e, w_cls, b_cls, precision, want_probs, aux_box):
        acc = _acc(precision)
        e_op = ops.split_bf16(e.contiguous(), acc)
        logits, probs, lse, arg = ops.box_score(e_op, weight_operand(w_cls, acc), b_cls, want_probs)
        aux_box.append(BoxScoreAux(lse, probs, arg))
        ctx.save_for_backward(w_cls)
        ctx.precision = precision
        return logits

    @staticmethod
    @once_differentiable
    def backward(ctx, dscores):
        (w_cls,) = ctx.saved_tensors
        acc = _acc(ctx.precision)
        de = None
        if ctx.needs_input_grad[0]:
            de, _ = ops.linear_fwd(ops.split_bf16(dscores.contiguous(), acc), weight_operand(w_cls, acc, transpose=True), None, want_f32=True)
        return de, None, None, None, None, None


def box_score(e, w_cls, b_cls=None, precision="fp32", want_probs=True):
    """Returns (scores, BoxScoreAux); the class matrix is a frozen input (box_emb_head.py:233-236)."""
    aux = []
    scores = _BoxScore.apply(e, w_cls, b_cls, precision, want_probs, aux)
    return scores, aux[0]


class _RowNormalize(Function):
    """normalize_vec / standardize_vec (logged_module.py:55-72) on the rows of a 2-D tensor, forward and backward on the device."""

    @staticmethod
    def forward(ctx, x, mode):
        ctx.save_for_backward(x)
        ctx.mode = mode
        return ops.row_normalize(x, mode)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.row_normalize(x, ctx.mode, dy=dy.contiguous()), None


def normalize_rows(x):
    return _RowNormalize.apply(x.to(torch.float32).contiguous(), ops.NORM_L2)


def standardize_rows(x):
    return _RowNormalize.apply(x.to(torch.float32).contiguous(), ops.NORM_STANDARDIZE)


class _BoxCE(Function):
    """mean cross-entropy from logits + the log-sum-exp the scoring epilogue already produced
    (Detectron2 FastRCNNOutputLayers.losses -> F.cross_entropy, reached from roi_emb_heads.py:266,347)."""

    @staticmethod
    def forward(ctx, logits, lse, labels):
        loss, dl, _ = ops.box_ce(logits, lse, labels, want_grad=True)
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return (dl * g if dl is not None else None), None, None


def box_cross_entropy(logits, lse, labels):
    return _BoxCE.apply(logits, lse, labels)


class _BoxRegLoss(Function):
    """Detectron2 FastRCNNOutputLayers.box_reg_loss (class-agnostic, smooth-L1 / L1, sum over foreground / max(R, 1)) with its gradient
    from the same launch (reached from roi_emb_heads.py:266,347)."""

    @staticmethod
    def forward(ctx, deltas, proposal_boxes, gt_boxes, labels, num_classes, reg_weights, beta):
        r = deltas.shape[0]
        loss, grad = ops.box_reg_loss(deltas, proposal_boxes, gt_boxes, labels, num_classes, reg_weights, beta, 1.0 / max(r, 1),
                                      want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g if grad is not None else None), None, None, None, None, None, None


def box_reg_loss(deltas, proposal_boxes, gt_boxes, labels, num_classes, reg_weights, beta):
    return _BoxRegLoss.apply(deltas, proposal_boxes, gt_boxes, labels, int(num_classes), tuple(float(w) for w in reg_weights), float(beta))


# ------------------------------------------------------------------------------------------------
# LSM grounding head: projection + pair distances in one autograd node
# ------------------------------------------------------------------------------------------------
def project_regions(x2d, w, b, d, acc, x_op=None, w_op=None):
    """v2l_projection (grounding_head.py:111) -> bf16 tensor-core operand of the region embeddings.
    fp32 mode: hi/lo split + three bf16 passes; reduced-precision mode: TF32 straight from the fp32 features.
    ``x_op`` / ``w_op``: operands of x2d / w that the preparation launch already produced (ops.lsm_prep with ``extra``)."""
    if not _use_tf32(acc, x2d.shape[1]):
        _, emb_op = ops.linear_fwd(x_op if x_op is not None else ops.split_bf16(x2d, acc), w_op if w_op is not None else weight_operand(w, acc), b,
                                   want_f32=False, n_bf16=d, accurate_out=acc)
    else:
        _, emb_op = ops.linear_tf32_fwd(x2d, w.detach(), b, want_f32=False, n_bf16=d)
    return emb_op


def new_pair_stack(bc, bi, device, both=True):
    """[2, Bc, Bi] buffer for (w2r, r2w): one allocation so the pair-CE kernel handles both in one launch."""
    return (torch.empty if both else torch.zeros)((2, bc, bi), dtype=torch.float32, device=device)


class _LsmHead(Function):
    """(region_features, W, b, cap) -> pw [2, Bc, Bi] = (d_w2r, d_r2w); reference grounding_head.py:111-256.
    A slice whose alignment is switched off is zero and carries no gradient."""

    @staticmethod
    def forward(ctx, feats, w, b, cap, cap_mask, reg_mask, inv_temp, alignment, precision, want_w2r, want_r2w, cap_op, x_op=None, w_op=None):
        acc = _acc(precision)
        bi, rg, v = feats.shape
        bc, t, d = cap.shape
        emb_op = project_regions(feats.reshape(bi * rg, v), w, b, d, acc, x_op, w_op)
        if cap_op is None or (cap_op.lo is None) == acc or cap_op.rows != bc * t or cap_op.cols != d:
            cap_op = ops.split_bf16(cap.reshape(bc * t, d), acc)    # (not prepared by ops.lsm_prep, or for another mode)
        stack = new_pair_stack(bc, bi, feats.device, want_w2r and want_r2w)
        ops.lsm_pair(cap_op, cap_mask, emb_op, reg_mask, inv_temp, alignment, want_w2r, want_r2w, stack[0], stack[1])
        ctx.ops_saved = (emb_op, cap_op)
        ctx.save_for_backward(feats, w, cap_mask, reg_mask)
        ctx.meta = (inv_temp, alignment, precision, b is not None, tuple(cap.shape), want_w2r, want_r2w)
        return stack

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        feats, w, cap_mask, reg_mask = ctx.saved_tensors
        emb_op, cap_op = ctx.ops_saved
        inv_temp, alignment, precision, has_b, cap_shape, want_w2r, want_r2w = ctx.meta
        acc = _acc(precision)
        bi, rg, v = feats.shape
        need_cap = ctx.needs_input_grad[3]
        g = g.contiguous()
        demb, dcap = ops.lsm_pair_bwd(cap_op, cap_mask, emb_op, reg_mask, inv_temp, alignment,
                                      g[0] if want_w2r else None, g[1] if want_r2w else None, need_cap)
        dx = dw = db = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            g_op = ops.split_bf16(demb, acc)
        if ctx.needs_input_grad[0]:
            wt_op = weight_operand(w, acc, transpose=True)
            dx, _ = ops.linear_fwd(g_op, wt_op, None, want_f32=True)
            dx = dx.reshape(bi, rg, v)
        if ctx.needs_input_grad[1]:
            x_op = ops.split_bf16(feats.reshape(bi * rg, v), acc)
            dw, _ = ops.linear_fwd(ops.transpose_operand(g_op), ops.transpose_operand(x_op), None, want_f32=True)
        if has_b and ctx.needs_input_grad[2]:
            db = demb.sum(0)
        if need_cap:
            dcap = dcap.reshape(cap_shape)
        return dx, dw, db, (dcap if need_cap else None), None, None, None, None, None, None, None, None, None, None


def lsm_head(feats, w, b, cap, cap_mask, reg_mask, temperature, alignment="softmax", precision="fp32",
             want_w2r=True, want_r2w=True, cap_op=None, x_op=None, w_op=None):
    """Stacked pair distance matrices [2, Bc, Bi] = (w2r, r2w) (rows = captions, cols = images) before the
    empty-pair guard; the slice of a switched-off alignment is zero.  ``cap_op`` = the caption operand when
    ``ops.lsm_prep`` already produced it together with the masks (same values as ``cap``)."""
    amode = {"softmax": ops.ALIGN_SOFTMAX, "hardmax": ops.ALIGN_HARDMAX}.get(alignment)
    if amode is None:
        raise NotImplementedError(f"alignment {alignment!r} is not implemented on the B200 path")
    return _LsmHead.apply(feats, w, b, cap, cap_mask, reg_mask, 1.0 / float(temperature), amode, precision,
                          bool(want_w2r), bool(want_r2w), cap_op, x_op, w_op)


class _PairCE(Function):
    """Empty-pair guard + the two cross-entropy losses and two accuracies of each stacked pair matrix
    (grounding_head.py:240-251, 272-290, 354-379), one launch.  Returns (guarded pw [n,Bc,Bi], out [n,4])."""

    @staticmethod
    def forward(ctx, pw, cap_mask, reg_mask, diag_offset):
        need = ctx.needs_input_grad[0]
        if need:
            pw = pw.clone()          # the guard writes in place; without a graph the fresh kernel output is reused

        res = ops.pair_ce(pw, cap_mask, reg_mask, diag_offset, want_grad=need)
        if need:
            out, dcap, dimg = res
            ctx.save_for_backward(dcap, dimg)
        else:
            out = res
        ctx.mark_non_differentiable(pw)
        return pw, out

    @staticmethod
    @once_differentiable
    def backward(ctx, _gpw, g):
        dcap, dimg = ctx.saved_tensors
        return dcap * g[:, 0, None, None] + dimg * g[:, 1, None, None], None, None, None


def pair_losses(pw, cap_mask, reg_mask, diag_offset=0):
    """pw [n,Bc,Bi] -> (guarded pw, out [n,4] = [CE caption, CE image, acc caption, acc image] per matrix)."""
    if not (torch.is_grad_enabled() and pw.requires_grad):
        # no graph: `pw` is the fresh output of the pair kernel, so the guard may write it in place
        return pw, ops.pair_ce(pw, cap_mask, reg_mask, int(diag_offset))
    return _PairCE.apply(pw, cap_mask, reg_mask, int(diag_offset))


# ------------------------------------------------------------------------------------------------
# distillation losses on the pair matrices
# ------------------------------------------------------------------------------------------------
class _PairDistill(Function):
    """MultiDistillLoss / MultiDistillLossL2 (distill_mmss_gcnn.py:211-289, 381-433): loss and all gradients from two launches."""

    @staticmethod
    def forward(ctx, trans, w2r, r2w, temperature, kind, loss_weight, detach_teacher):
        need_t, need_w, need_r = ctx.needs_input_grad[:3]
        # DISTILLATION_DETACH_TEACHER: the teacher side carries no gradient — `trans` when it is the target (kind 0 / MSE with a
        # transformer teacher), the two student matrices when THEY are the targets (kind 1)
        if detach_teacher:
            if kind == ops.DISTILL_KD_STUDENT_TARGET:
                need_w = need_r = False
            else:
                need_t = False
        loss, gt, gw, gr = ops.pair_distill(trans, w2r, r2w, temperature, kind, loss_weight, grad_trans=need_t, grad_students=need_w or need_r)
        ctx.save_for_backward(gt, gw if need_w else None, gr if need_r else None)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        gt, gw, gr = ctx.saved_tensors
        return (gt * g if gt is not None else None), (gw * g if gw is not None else None), (gr * g if gr is not None else None), None, None, None, None


def pair_distill(trans, w2r, r2w, temperature, kind, loss_weight=1.0, detach_teacher=False):
    return _PairDistill.apply(trans.to(torch.float32), w2r.to(torch.float32), r2w.to(torch.float32), float(temperature), int(kind), float(loss_weight),
                              bool(detach_teacher))
