"""Thin torch-tensor front-ends of the C ABI (include/locov_b200.h).

PyTorch is used for device memory and streams only; every computation below is a call into
liblocov_b200.so.  All functions require CUDA tensors and raise ``LocoError`` on failure — there is
no eager/CPU fallback.
"""
import ctypes
import threading
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import LocoError

ALIGN_SOFTMAX, ALIGN_HARDMAX = 0, 1
NCHW, NHWC = 0, 1
F32, BF16 = 0, 1          # LOCO_F32 / LOCO_BF16 (include/locov_b200.h)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise LocoError("locov_b200 ops need CUDA tensors (there is no CPU fallback)")


def _round_up(x, m):
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------
# RoIAlign
# ------------------------------------------------------------------------------------------------
_ws_cache = {}


def _ws_key(device, tag):
    # scratch is private to a (device, stream, host thread): two streams or the autograd thread never share tickets / scratch
    return (device.index, tag, torch.cuda.current_stream(device).cuda_stream, threading.get_ident())


def _workspace(device, nbytes, tag):
    key = _ws_key(device, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _zero_workspace(device, nbytes, tag):
    """Scratch that is zero-initialised once and that the kernels leave zeroed (ticket counters)."""
    key = _ws_key(device, tag) + ("zero",)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _drop_zero_workspace(device, tag):
    """After a failed launch the ticket counters may be non-zero: forget the buffer (a fresh zeroed one is made next call)."""
    _ws_cache.pop(_ws_key(device, tag) + ("zero",), None)


def roi_align(feat: torch.Tensor, rois: torch.Tensor, output_size, spatial_scale: float, sampling_ratio: int = 0,
              aligned: bool = True, channels_last: bool = False, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """feat [N,C,H,W] fp32 (contiguous NCHW, or channels_last memory format), rois [R,5] -> [R,C,PH,PW].

    ``channels_last=True`` returns the same logical [R,C,PH,PW] tensor in ``torch.channels_last`` memory format (what cuDNN's
    tensor-core convolutions of res5 consume without a transpose), in fp32 or bf16 (``out_dtype``)."""
    _need_cuda(feat, rois)
    if out_dtype not in (torch.float32, torch.bfloat16) or (out_dtype == torch.bfloat16 and not channels_last):
        raise LocoError("roi_align: output is NCHW fp32, or channels-last fp32 / bf16")
    if feat.dtype != torch.float32:
        raise LocoError("roi_align: fp32 features only")
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    n, c, h, w = feat.shape
    if feat.is_contiguous():
        layout = NCHW
    elif feat.is_contiguous(memory_format=torch.channels_last):
        layout = NHWC
    else:
        feat, layout = feat.contiguous(), NCHW
    rois = rois.to(torch.float32).contiguous()
    r = rois.shape[0]
    if channels_last:
        out = torch.empty((r, ph, pw, c), dtype=out_dtype, device=feat.device).permute(0, 3, 1, 2)
    else:
        out = torch.empty((r, c, ph, pw), dtype=torch.float32, device=feat.device)
    lib = _lib.load()
    ws = _workspace(feat.device, lib.loco_roi_align_workspace_bytes(n, c, h, w, layout, r), "roi_align")
    _lib.check(lib.loco_roi_align_fwd(_p(feat), n, c, h, w, layout, _p(rois), r, ph, pw, float(spatial_scale),
                                      int(sampling_ratio), int(bool(aligned)), _p(out), NHWC if channels_last else NCHW,
                                      BF16 if out_dtype == torch.bfloat16 else F32, _p(ws), _stream(feat)),
               "loco_roi_align_fwd")
    return out


def _mean_layout(x: torch.Tensor):
    """(tensor, layout, dtype code) of a [R,C,PH,PW] tensor as the spatial-mean kernels take it."""
    if x.dim() != 4:
        raise LocoError("spatial_mean: expects [R, C, PH, PW]")
    if x.dtype == torch.float32 and x.is_contiguous():
        return x, NCHW, F32
    if x.dtype in (torch.float32, torch.bfloat16) and x.shape[1] % 4 == 0:
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        return x, NHWC, F32 if x.dtype == torch.float32 else BF16
    return x.to(torch.float32).contiguous(), NCHW, F32


def spatial_mean(x: torch.Tensor, operand: Optional[bool] = None):
    """x [R,C,PH,PW] (NCHW fp32, or channels-last fp32 / bf16) -> (mean [R,C] fp32, Bf16Operand of it or None).

    ``operand``: None = no bf16 operand, False = hi only, True = hi + lo (the fp32-accurate three-pass operand)."""
    _need_cuda(x)
    x, layout, dt = _mean_layout(x)
    r, c, ph, pw = x.shape
    out = torch.empty((r, c), dtype=torch.float32, device=x.device)
    op = None
    if operand is not None:
        ld = _round_up(c, 8)
        mk = torch.empty if ld == c else torch.zeros
        hi = mk((r, ld), dtype=torch.bfloat16, device=x.device)
        lo = mk((r, ld), dtype=torch.bfloat16, device=x.device) if operand else None
        op = Bf16Operand(hi, lo, r, c)
    lib = _lib.load()
    _lib.check(lib.loco_spatial_mean(_p(x), r, c, ph * pw, layout, dt, _p(out), c, _p(op.hi) if op else None,
                                     _p(op.lo) if op is not None and op.lo is not None else None, op.ld if op else 0, _stream(x)),
               "loco_spatial_mean")
    return out, op


def spatial_mean_backward(dy: torch.Tensor, shape, dtype: torch.dtype = torch.float32, channels_last: bool = False) -> torch.Tensor:
    """dx [R,C,PH,PW] = dy [R,C] / (PH*PW) broadcast over the positions; NCHW fp32, or channels-last fp32 / bf16."""
    _need_cuda(dy)
    dy = dy.to(torch.float32)
    if dy.dim() != 2 or dy.stride(1) != 1:
        dy = dy.reshape(shape[0], shape[1]).contiguous()
    r, c, ph, pw = shape
    if channels_last and dtype in (torch.float32, torch.bfloat16):
        dx = torch.empty((r, ph, pw, c), dtype=dtype, device=dy.device).permute(0, 3, 1, 2)
        layout, dt = NHWC, F32 if dtype == torch.float32 else BF16
    else:
        dx, layout, dt = torch.empty((r, c, ph, pw), dtype=torch.float32, device=dy.device), NCHW, F32
    lib = _lib.load()
    _lib.check(lib.loco_spatial_mean_bwd(_p(dy), dy.stride(0) if r > 0 else c, r, c, ph * pw, layout, dt, _p(dx), _stream(dy)),
               "loco_spatial_mean_bwd")
    return dx if dx.dtype == dtype else dx.to(dtype)


def roi_align_backward(dout: torch.Tensor, feat_shape, rois: torch.Tensor, spatial_scale: float,
                       sampling_ratio: int = 0, aligned: bool = True) -> torch.Tensor:
    _need_cuda(dout, rois)
    n, c, h, w = feat_shape
    dout = dout.to(torch.float32).contiguous()
    rois = rois.to(torch.float32).contiguous()
    r, _, ph, pw = dout.shape
    lib = _lib.load()
    nbytes = int(lib.loco_roi_align_bwd_workspace_bytes(n, c, h, w, r))
    ws = _workspace(dout.device, nbytes, "roi_align_bwd") if nbytes > 0 else None
    # always zero-filled: the vectorised path overwrites it, but the library may decline that path (R == 0, pooled sizes whose
    # staging exceeds shared memory, LOCOV_B200_ROI_BWD_V4=0, huge maps) and accumulate with scalar atomics instead
    dfeat = torch.zeros(feat_shape, dtype=torch.float32, device=dout.device)
    _lib.check(lib.loco_roi_align_bwd(_p(dout), n, c, h, w, _p(rois), r, ph, pw, float(spatial_scale),
                                      int(sampling_ratio), int(bool(aligned)), _p(dfeat), _p(ws), _stream(dout)),
               "loco_roi_align_bwd")
    return dfeat


def roi_align_grid(rois: torch.Tensor, h: int, w: int, output_size, spatial_scale: float, sampling_ratio: int = 0,
                   aligned: bool = True, max_grid: int = 4):
    """Sampling grid as computed ON THE DEVICE: (grid_hw [R,2] i32, yx [R,PH,PW,G,G,2] f32, idx [...,4] i32)."""
    _need_cuda(rois)
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    rois = rois.to(torch.float32).contiguous()
    r = rois.shape[0]
    ghw = torch.zeros((r, 2), dtype=torch.int32, device=rois.device)
    yx = torch.full((r, ph, pw, max_grid, max_grid, 2), float("nan"), dtype=torch.float32, device=rois.device)
    idx = torch.full((r, ph, pw, max_grid, max_grid, 4), -2, dtype=torch.int32, device=rois.device)
    lib = _lib.load()
    _lib.check(lib.loco_roi_align_grid_dump(_p(rois), r, h, w, ph, pw, float(spatial_scale), int(sampling_ratio),
                                            int(bool(aligned)), max_grid, _p(ghw), _p(yx), _p(idx), _stream(rois)),
               "loco_roi_align_grid_dump")
    return ghw, yx, idx


# ------------------------------------------------------------------------------------------------
# bf16 operands
# ------------------------------------------------------------------------------------------------
@dataclass
class Bf16Operand:
    """A [rows, cols] matrix as bf16 ``hi`` (+ optional ``lo`` residual) with row stride ``ld`` elements."""
    hi: torch.Tensor
    lo: Optional[torch.Tensor]
    rows: int
    cols: int

    @property
    def ld(self):
        return self.hi.shape[1]


def split_bf16(x: torch.Tensor, accurate: bool, transpose: bool = False, out: Optional[Bf16Operand] = None) -> Bf16Operand:
    """fp32 2-D tensor -> Bf16Operand (ld padded to a multiple of 8, pad zero-filled).

    accurate=True also produces the ``lo`` residual (fp32-accurate three-pass mode)."""
    _need_cuda(x)
    if x.dim() != 2 or x.dtype != torch.float32:
        raise LocoError("split_bf16: expects a 2-D fp32 tensor")
    if x.stride(1) != 1:
        x = x.contiguous()
    rows, cols = x.shape
    drows, dcols = (cols, rows) if transpose else (rows, cols)
    ld = _round_up(max(dcols, 1), 8)
    if out is None:
        hi = torch.empty((drows, ld), dtype=torch.bfloat16, device=x.device)
        lo = torch.empty((drows, ld), dtype=torch.bfloat16, device=x.device) if accurate else None
        out = Bf16Operand(hi, lo, drows, dcols)
    lib = _lib.load()
    _lib.check(lib.loco_split_bf16(_p(x), rows, cols, x.stride(0), _p(out.hi), _p(out.lo) if accurate else None,
                                   out.ld, int(transpose), _stream(x)), "loco_split_bf16")
    return out


# ------------------------------------------------------------------------------------------------
# tensor-core contractions
# ------------------------------------------------------------------------------------------------
def linear_fwd(a: Bf16Operand, w: Bf16Operand, bias: Optional[torch.Tensor], want_f32: bool = True,
               n_bf16: int = 0, accurate_out: bool = False) -> Tuple[Optional[torch.Tensor], Optional[Bf16Operand]]:
    """out[M,N] = A[M,K] · W[N,K]^T + bias.  Returns (out_f32 or None, bf16 operand of the first n_bf16 columns or None)."""
    if a.cols != w.cols:
        raise LocoError(f"linear_fwd: K mismatch {a.cols} vs {w.cols}")
    if (a.lo is None) != (w.lo is None):
        raise LocoError("linear_fwd: both operands must be in the same precision mode")
    m, n, k = a.rows, w.rows, a.cols
    dev = a.hi.device
    out_f32 = torch.empty((m, n), dtype=torch.float32, device=dev) if want_f32 else None
    ob = None
    if n_bf16 > 0:
        ld = _round_up(n_bf16, 8)
        hi = torch.empty((m, ld), dtype=torch.bfloat16, device=dev)
        if ld != n_bf16:
            hi[:, n_bf16:].zero_()
        lo = None
        if accurate_out:
            lo = torch.empty((m, ld), dtype=torch.bfloat16, device=dev)
            if ld != n_bf16:
                lo[:, n_bf16:].zero_()
        ob = Bf16Operand(hi, lo, m, n_bf16)
    if bias is not None:
        bias = bias.to(torch.float32).contiguous()
    lib = _lib.load()
    _lib.check(lib.loco_linear_fwd(_p(a.hi), _p(a.lo), a.ld, _p(w.hi), _p(w.lo), w.ld, _p(bias), m, n, k,
                                   _p(out_f32), n if want_f32 else 0, _p(ob.hi) if ob else None,
                                   _p(ob.lo) if (ob and ob.lo is not None) else None, n_bf16, ob.ld if ob else 0,
                                   _stream(a.hi)), "loco_linear_fwd")
    return out_f32, ob


def linear_fwd_into(a: Bf16Operand, w: Bf16Operand, bias: Optional[torch.Tensor], want_f32: bool, out_hi: torch.Tensor,
                    out_lo: Optional[torch.Tensor], n_bf16: int) -> Optional[torch.Tensor]:
    """``linear_fwd`` whose bf16 output lands in the first ``n_bf16`` columns of caller-owned matrices (row stride = their width):
    lets a gradient GEMM write one slice of a wider operand.  Returns the fp32 output or None."""
    if a.cols != w.cols or (a.lo is None) != (w.lo is None):
        raise LocoError("linear_fwd_into: operand mismatch")
    m, n, k = a.rows, w.rows, a.cols
    if n_bf16 > n or out_hi.shape[0] != m or out_hi.stride(1) != 1 or out_hi.stride(0) % 8:
        raise LocoError("linear_fwd_into: bad destination")
    out_f32 = torch.empty((m, n), dtype=torch.float32, device=a.hi.device) if want_f32 else None
    if bias is not None:
        bias = bias.to(torch.float32).contiguous()
    _lib.check(_lib.load().loco_linear_fwd(_p(a.hi), _p(a.lo), a.ld, _p(w.hi), _p(w.lo), w.ld, _p(bias), m, n, k, _p(out_f32), n if want_f32 else 0,
                                           _p(out_hi), _p(out_lo), n_bf16, out_hi.stride(0), _stream(a.hi)), "loco_linear_fwd")
    return out_f32


def linear_tf32_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], want_f32: bool = True, n_bf16: int = 0):
    """out[M,N] = x[M,K] · w[N,K]^T + bias with the fp32 operands read in place as TF32 (no split pass).
    Returns (out_f32 or None, bf16 operand of the first n_bf16 columns or None)."""
    _need_cuda(x, w)
    if x.dtype != torch.float32 or w.dtype != torch.float32 or x.dim() != 2 or w.dim() != 2 or x.shape[1] != w.shape[1]:
        raise LocoError("linear_tf32_fwd: expects fp32 [M,K] and [N,K]")
    if x.stride(1) != 1 or x.stride(0) % 4 or x.data_ptr() % 16:
        x = x.contiguous()
    if w.stride(1) != 1 or w.stride(0) % 4 or w.data_ptr() % 16:
        w = w.contiguous()
    m, k = x.shape
    n = w.shape[0]
    if k % 4:
        raise LocoError("linear_tf32_fwd: K must be a multiple of 4 (16-byte TMA rows)")
    dev = x.device
    out_f32 = torch.empty((m, n), dtype=torch.float32, device=dev) if want_f32 else None
    ob = None
    if n_bf16 > 0:
        ld = _round_up(n_bf16, 8)
        hi = torch.empty((m, ld), dtype=torch.bfloat16, device=dev)
        if ld != n_bf16:
            hi[:, n_bf16:].zero_()
        ob = Bf16Operand(hi, None, m, n_bf16)
    if bias is not None:
        bias = bias.to(torch.float32).contiguous()
    lib = _lib.load()
    _lib.check(lib.loco_linear_tf32_fwd(_p(x), x.stride(0), _p(w), w.stride(0), _p(bias), m, n, k, _p(out_f32),
                                        n if want_f32 else 0, _p(ob.hi) if ob else None, None, n_bf16, ob.ld if ob else 0,
                                        _stream(x)), "loco_linear_tf32_fwd")
    return out_f32, ob


def box_score(e: Bf16Operand, cls: Bf16Operand, cls_bias: Optional[torch.Tensor] = None, want_probs: bool = True):
    """RoI x class logits with fused softmax: returns (logits [R,K1], probs or None, lse [R], argmax_fg [R] i64)."""
    if e.cols != cls.cols:
        raise LocoError(f"box_score: embedding dim mismatch {e.cols} vs {cls.cols}")
    if (e.lo is None) != (cls.lo is None):
        raise LocoError("box_score: both operands must be in the same precision mode")
    r, k1, d = e.rows, cls.rows, e.cols
    dev = e.hi.device
    logits = torch.empty((r, k1), dtype=torch.float32, device=dev)
    probs = torch.empty((r, k1), dtype=torch.float32, device=dev) if want_probs else None
    lse = torch.empty((r,), dtype=torch.float32, device=dev)
    arg = torch.empty((r,), dtype=torch.int64, device=dev)
    if cls_bias is not None:
        cls_bias = cls_bias.to(torch.float32).contiguous()
    lib = _lib.load()
    ws = _workspace(dev, lib.loco_box_score_workspace_bytes(r, k1), "box_score")
    _lib.check(lib.loco_box_score_fwd(_p(e.hi), _p(e.lo), e.ld, _p(cls.hi), _p(cls.lo), cls.ld, _p(cls_bias), r, k1, d,
                                      _p(logits), _p(probs), k1, _p(lse), _p(arg), _p(ws), _stream(e.hi)),
               "loco_box_score_fwd")
    return logits, probs, lse, arg


def box_ce(logits: torch.Tensor, lse: torch.Tensor, labels: torch.Tensor, want_grad: bool = False,
           grad_bf16: bool = False):
    """Mean cross-entropy (Detectron2 FastRCNNOutputLayers.losses) from the fused softmax statistics.
    Returns (loss scalar tensor, dlogits fp32 or None, dlogits bf16 [R, pad8(K1)] or None); gradients are of the MEAN loss."""
    _need_cuda(logits, lse, labels)
    r, k1 = logits.shape
    dev = logits.device
    loss = torch.zeros((), dtype=torch.float32, device=dev)
    if r == 0:
        return loss, None, None
    labels = labels.to(torch.int64).contiguous()
    dl = torch.empty_like(logits) if want_grad else None
    dlb = torch.empty((r, _round_up(k1, 8)), dtype=torch.bfloat16, device=dev) if grad_bf16 else None
    lib = _lib.load()
    _lib.check(lib.loco_box_ce_fwd_bwd(_p(logits), logits.stride(0), _p(lse), _p(labels), r, k1, 1.0 / r, _p(loss),
                                       1.0 / r, _p(dl), _p(dlb), dlb.shape[1] if dlb is not None else 0,
                                       _stream(logits)), "loco_box_ce_fwd_bwd")
    return loss, dl, dlb


def box_reg_loss(deltas: torch.Tensor, proposal_boxes: torch.Tensor, gt_boxes: torch.Tensor, labels: torch.Tensor, num_classes: int,
                 reg_weights, smooth_l1_beta: float, scale: float, want_grad: bool = False):
    """Class-agnostic box-regression loss (Detectron2 box_reg_loss) in one launch: returns (loss scalar tensor, d loss / d deltas or None)."""
    _need_cuda(deltas, proposal_boxes, gt_boxes, labels)
    r = deltas.shape[0]
    dev = deltas.device
    if deltas.dim() != 2 or deltas.shape[1] != 4:
        raise LocoError("box_reg_loss: class-agnostic deltas [R,4] expected")
    deltas = deltas.to(torch.float32)
    if deltas.stride(1) != 1:
        deltas = deltas.contiguous()
    prop = proposal_boxes.to(torch.float32).contiguous()
    gtb = gt_boxes.to(torch.float32).contiguous()
    labels = labels.to(torch.int64).contiguous()
    loss = torch.empty((), dtype=torch.float32, device=dev)
    grad = torch.empty((r, 4), dtype=torch.float32, device=dev) if want_grad else None
    lib = _lib.load()
    ws = _zero_workspace(dev, lib.loco_box_reg_loss_workspace_bytes(r), "box_reg_loss")
    w4 = (ctypes.c_float * 4)(*[float(v) for v in reg_weights])
    rc = lib.loco_box_reg_loss(_p(deltas), deltas.stride(0) if r else 4, _p(prop), _p(gtb), _p(labels), r, int(num_classes), w4, float(smooth_l1_beta),
                               float(scale), _p(loss), _p(grad), 4, _p(ws), _stream(deltas))
    if rc != 0:
        _drop_zero_workspace(dev, "box_reg_loss")
    _lib.check(rc, "loco_box_reg_loss")
    return loss, grad


def skinny_grad(dy: torch.Tensor, x: torch.Tensor, want_bias: bool = True):
    """dW [J,V] = dy^T . x and db [J] = column sums of dy for a linear layer with J <= 8 outputs (bbox_pred)."""
    _need_cuda(dy, x)
    dy = dy.to(torch.float32)
    if dy.stride(1) != 1:
        dy = dy.contiguous()
    if x.stride(1) != 1 or x.stride(0) % 4 or x.data_ptr() % 16:
        x = x.contiguous()
    r, j = dy.shape
    v = x.shape[1]
    dw = torch.empty((j, v), dtype=torch.float32, device=x.device)
    db = torch.empty((j,), dtype=torch.float32, device=x.device) if want_bias else None
    lib = _lib.load()
    ws = _workspace(x.device, lib.loco_skinny_grad_workspace_bytes(r, j, v), "skinny_grad")
    _lib.check(lib.loco_skinny_grad(_p(dy), dy.stride(0) if r else j, _p(x), x.stride(0) if r else v, r, j, v, _p(dw), _p(db), _p(ws), _stream(x)),
               "loco_skinny_grad")
    return dw, db


def box_softmax(logits: torch.Tensor, want_probs: bool = False):
    """Softmax statistics of an existing score matrix [R,K1]: returns (lse [R], argmax_fg [R] i64, probs or None)."""
    _need_cuda(logits)
    if logits.dim() != 2 or logits.dtype != torch.float32:
        raise LocoError("box_softmax: expects a 2-D fp32 score matrix")
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    r, k1 = logits.shape
    dev = logits.device
    lse = torch.empty((r,), dtype=torch.float32, device=dev)
    arg = torch.empty((r,), dtype=torch.int64, device=dev)
    probs = torch.empty((r, k1), dtype=torch.float32, device=dev) if want_probs else None
    _lib.check(_lib.load().loco_box_softmax(_p(logits), logits.stride(0) if r else k1, r, k1, _p(lse), _p(arg), _p(probs), k1, _stream(logits)),
               "loco_box_softmax")
    return lse, arg, probs


NORM_L2, NORM_STANDARDIZE = 0, 1


def row_normalize(x: torch.Tensor, mode: int, dy: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Row-wise L2 normalisation (mode 0) / standardisation (mode 1) of x [rows, cols]; with ``dy`` the gradient w.r.t. x."""
    _need_cuda(x, dy)
    if x.dim() != 2 or x.dtype != torch.float32:
        raise LocoError("row_normalize: expects a 2-D fp32 tensor")
    x = x if x.stride(1) == 1 else x.contiguous()
    if dy is not None:
        dy = dy.to(torch.float32)
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().loco_row_normalize(_p(x), x.stride(0) if rows else cols, rows, cols, int(mode), _p(dy),
                                              (dy.stride(0) if rows else cols) if dy is not None else 0, _p(out), cols, _stream(x)),
               "loco_row_normalize")
    return out


_REG_KIND = {torch.uint8: 0, torch.bool: 0, torch.float32: 1, torch.int64: 2}


def lsm_masks(attention_mask: torch.Tensor, special_tokens_mask: torch.Tensor, region_mask: torch.Tensor):
    """-> (caption_mask [B,T] fp32 = attention * (1 - special), region_mask [B,Rg] fp32), one launch."""
    _need_cuda(attention_mask, special_tokens_mask, region_mask)
    if attention_mask.dtype != torch.int64 or special_tokens_mask.dtype != torch.int64 or region_mask.dtype not in _REG_KIND:
        raise LocoError("lsm_masks: expects int64 token masks and a uint8/bool/fp32/int64 region mask")
    att, spe, reg = attention_mask.contiguous(), special_tokens_mask.contiguous(), region_mask.contiguous()
    cap_mask = torch.empty(att.shape, dtype=torch.float32, device=att.device)
    reg_mask = torch.empty(reg.shape, dtype=torch.float32, device=att.device)
    lib = _lib.load()
    _lib.check(lib.loco_lsm_masks(_p(att), _p(spe), att.numel(), _p(reg), _REG_KIND[reg.dtype], reg.numel(), _p(cap_mask),
                                  _p(reg_mask), _stream(att)), "loco_lsm_masks")
    return cap_mask, reg_mask


def lsm_prep(cap2d: torch.Tensor, accurate: bool, attention_mask: torch.Tensor, special_tokens_mask: torch.Tensor,
             region_mask: torch.Tensor, extra=()):
    """Caption embeddings [B*T, D] fp32 -> Bf16Operand, plus (caption_mask, region_mask) as in ``lsm_masks`` — one launch.

    ``extra``: up to two more 2-D fp32 matrices (the region features, the projection weight) converted to operands of the same
    precision IN THE SAME LAUNCH; their operands are returned as a fourth element (list)."""
    _need_cuda(cap2d, attention_mask, special_tokens_mask, region_mask)
    if cap2d.dim() != 2 or cap2d.dtype != torch.float32:
        raise LocoError("lsm_prep: expects 2-D fp32 caption embeddings")
    if attention_mask.dtype != torch.int64 or special_tokens_mask.dtype != torch.int64 or region_mask.dtype not in _REG_KIND:
        raise LocoError("lsm_prep: expects int64 token masks and a uint8/bool/fp32/int64 region mask")
    if cap2d.stride(1) != 1:
        cap2d = cap2d.contiguous()
    att, spe, reg = attention_mask.contiguous(), special_tokens_mask.contiguous(), region_mask.contiguous()
    cap_mask = torch.empty(att.shape, dtype=torch.float32, device=att.device)
    reg_mask = torch.empty(reg.shape, dtype=torch.float32, device=att.device)

    def operand(x):
        rows, cols = x.shape
        ld = _round_up(max(cols, 1), 8)
        hi = torch.empty((rows, ld), dtype=torch.bfloat16, device=x.device)
        lo = torch.empty((rows, ld), dtype=torch.bfloat16, device=x.device) if accurate else None
        return Bf16Operand(hi, lo, rows, cols)
    lib = _lib.load()
    cap_op = operand(cap2d)
    if not extra:
        _lib.check(lib.loco_lsm_prep(_p(cap2d), cap_op.rows, cap_op.cols, cap2d.stride(0), _p(cap_op.hi), _p(cap_op.lo) if accurate else None, cap_op.ld,
                                     _p(att), _p(spe), att.numel(), _p(reg), _REG_KIND[reg.dtype], reg.numel(), _p(cap_mask), _p(reg_mask),
                                     _stream(cap2d)), "loco_lsm_prep")
        return cap_op, cap_mask, reg_mask
    if len(extra) > 2:
        raise LocoError("lsm_prep: at most two extra matrices")
    srcs, outs = [], []
    for x in extra:
        _need_cuda(x)
        if x.dim() != 2 or x.dtype != torch.float32:
            raise LocoError("lsm_prep: extra matrices must be 2-D fp32")
        x = x if x.stride(1) == 1 else x.contiguous()
        srcs.append(x)
        outs.append(operand(x))
    jobs = sorted(zip(srcs + [cap2d], outs + [cap_op]), key=lambda so: -so[0].numel())         # largest job first
    n = len(jobs)
    vp, i64 = ctypes.c_void_p * n, ctypes.c_int64 * n
    _lib.check(lib.loco_lsm_prep_multi(
        n, vp(*[x.data_ptr() for x, _ in jobs]), i64(*[x.shape[0] for x, _ in jobs]), i64(*[x.shape[1] for x, _ in jobs]),
        i64(*[x.stride(0) for x, _ in jobs]), vp(*[o.hi.data_ptr() for _, o in jobs]), vp(*[(o.lo.data_ptr() if o.lo is not None else None) for _, o in jobs]),
        i64(*[o.ld for _, o in jobs]), _p(att), _p(spe), att.numel(), _p(reg), _REG_KIND[reg.dtype], reg.numel(), _p(cap_mask), _p(reg_mask),
        _stream(cap2d)), "loco_lsm_prep_multi")
    return cap_op, cap_mask, reg_mask, outs


def lsm_pair(cap: Bf16Operand, cap_mask: torch.Tensor, emb: Bf16Operand, reg_mask: torch.Tensor, inv_temperature: float,
             alignment: int = ALIGN_SOFTMAX, want_w2r: bool = True, want_r2w: bool = True,
             out_w2r: Optional[torch.Tensor] = None, out_r2w: Optional[torch.Tensor] = None):
    """Pair distances d_w2r, d_r2w [Bc, Bi] (rows captions, cols images) from caption word embeddings
    cap [Bc*T, D] and projected region embeddings emb [Bi*Rg, D]."""
    bc, t = cap_mask.shape
    bi, rg = reg_mask.shape
    if cap.rows != bc * t or emb.rows != bi * rg or cap.cols != emb.cols:
        raise LocoError("lsm_pair: operand shapes do not match the masks")
    if (cap.lo is None) != (emb.lo is None):
        raise LocoError("lsm_pair: both operands must be in the same precision mode")
    dev = cap.hi.device
    cap_mask = cap_mask.to(torch.float32).contiguous()
    reg_mask = reg_mask.to(torch.float32).contiguous()
    if out_w2r is None and want_w2r:
        out_w2r = torch.empty((bc, bi), dtype=torch.float32, device=dev)
    if out_r2w is None and want_r2w:
        out_r2w = torch.empty((bc, bi), dtype=torch.float32, device=dev)
    w2r = out_w2r if want_w2r else None
    r2w = out_r2w if want_r2w else None
    ld = (w2r if w2r is not None else r2w).stride(0)
    if w2r is not None and r2w is not None and w2r.stride(0) != r2w.stride(0):
        raise LocoError("lsm_pair: outputs must share a row stride")
    lib = _lib.load()
    ws = _workspace(dev, lib.loco_lsm_pair_workspace_bytes(bc, t, bi, rg), "lsm")
    _lib.check(lib.loco_lsm_pair_fwd(_p(cap.hi), _p(cap.lo), cap.ld, _p(cap_mask), _p(emb.hi), _p(emb.lo), emb.ld,
                                     _p(reg_mask), bc, t, bi, rg, cap.cols, float(inv_temperature), int(alignment),
                                     _p(w2r), _p(r2w), ld, _p(ws), _stream(cap.hi)), "loco_lsm_pair_fwd")
    return w2r, r2w


def transpose_operand(op: Bf16Operand) -> Bf16Operand:
    """bf16 operand [rows, cols] -> [cols, rows] (hi and lo)."""
    ld = _round_up(max(op.rows, 1), 8)
    dev = op.hi.device
    hi = torch.empty((op.cols, ld), dtype=torch.bfloat16, device=dev)
    lo = torch.empty((op.cols, ld), dtype=torch.bfloat16, device=dev) if op.lo is not None else None
    lib = _lib.load()
    for s_, d_ in ((op.hi, hi), (op.lo, lo)):
        if s_ is not None:
            _lib.check(lib.loco_transpose_bf16(_p(s_), op.rows, op.cols, op.ld, _p(d_), ld, _stream(op.hi)), "loco_transpose_bf16")
    return Bf16Operand(hi, lo, op.cols, op.rows)


def lsm_pair_bwd(cap: Bf16Operand, cap_mask: torch.Tensor, emb: Bf16Operand, reg_mask: torch.Tensor, inv_temperature: float,
                 alignment: int, g_w2r: Optional[torch.Tensor], g_r2w: Optional[torch.Tensor], need_cap: bool = False):
    """Gradients of the pair distances: returns (dEmb [Bi*Rg, D] fp32, dCap [Bc*T, D] fp32 or None).
    One recompute kernel (similarity GEMM + softmax Jacobians -> dS in bf16) and one tcgen05 GEMM per gradient."""
    bc, t = cap_mask.shape
    bi, rg = reg_mask.shape
    acc = cap.lo is not None
    dev = cap.hi.device
    d = cap.cols
    if g_w2r is None and g_r2w is None:
        return torch.zeros((bi * rg, d), dtype=torch.float32, device=dev), (torch.zeros((bc * t, d), dtype=torch.float32, device=dev) if need_cap else None)
    cap_mask = cap_mask.to(torch.float32).contiguous()
    reg_mask = reg_mask.to(torch.float32).contiguous()
    gs = [g.to(torch.float32).contiguous() if g is not None else None for g in (g_w2r, g_r2w)]
    ld_g = bi
    ld_t = _round_up(bc * t, 8)
    dst_hi = torch.empty((bi * rg, ld_t), dtype=torch.bfloat16, device=dev)
    dst_lo = torch.empty((bi * rg, ld_t), dtype=torch.bfloat16, device=dev) if acc else None
    ds_hi = ds_lo = None
    ld_s = _round_up(bi * rg, 8)
    if need_cap:
        ds_hi = torch.empty((bc * t, ld_s), dtype=torch.bfloat16, device=dev)
        ds_lo = torch.empty((bc * t, ld_s), dtype=torch.bfloat16, device=dev) if acc else None
    lib = _lib.load()
    _lib.check(lib.loco_lsm_pair_bwd(_p(cap.hi), _p(cap.lo), cap.ld, _p(cap_mask), _p(emb.hi), _p(emb.lo), emb.ld, _p(reg_mask),
                                     bc, t, bi, rg, d, float(inv_temperature), int(alignment), _p(gs[0]), _p(gs[1]), ld_g,
                                     _p(dst_hi), _p(dst_lo), ld_t, _p(ds_hi), _p(ds_lo), ld_s, _stream(cap.hi)),
               "loco_lsm_pair_bwd")
    # dEmb = dS^T [Bi*Rg, Bc*T] . cap [Bc*T, D]  ->  W operand = cap^T [D, Bc*T]
    demb, _ = linear_fwd(Bf16Operand(dst_hi, dst_lo, bi * rg, bc * t), transpose_operand(cap), None, want_f32=True)
    dcap = None
    if need_cap:
        dcap, _ = linear_fwd(Bf16Operand(ds_hi, ds_lo, bc * t, bi * rg), transpose_operand(emb), None, want_f32=True)
    return demb, dcap


DISTILL_KD_TEACHER_TARGET, DISTILL_KD_STUDENT_TARGET, DISTILL_MSE = 0, 1, 2


def pair_distill(trans: torch.Tensor, w2r: torch.Tensor, r2w: torch.Tensor, temperature: float, kind: int, loss_weight: float = 1.0,
                 grad_trans: bool = False, grad_students: bool = False):
    """Distillation loss on the [B,B] pair matrices (distill_mmss_gcnn.py:226-289, 381-433): returns (loss scalar tensor,
    g_trans or None, g_w2r or None, g_r2w or None)."""
    _need_cuda(trans, w2r, r2w)
    b = trans.shape[0]
    mats = []
    for m in (trans, w2r, r2w):
        if m.shape != (b, b) or m.dtype != torch.float32:
            raise LocoError("pair_distill: expects three fp32 [B,B] matrices")
        mats.append(m if m.stride(1) == 1 else m.contiguous())
    dev = trans.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    gt = torch.empty((b, b), dtype=torch.float32, device=dev) if grad_trans else None
    gw = torch.empty((b, b), dtype=torch.float32, device=dev) if grad_students else None
    gr = torch.empty((b, b), dtype=torch.float32, device=dev) if grad_students else None
    lib = _lib.load()
    ws = _zero_workspace(dev, lib.loco_pair_distill_workspace_bytes(b), "pair_distill")
    rc = lib.loco_pair_distill(_p(mats[0]), mats[0].stride(0), _p(mats[1]), mats[1].stride(0), _p(mats[2]), mats[2].stride(0), b, float(temperature), int(kind),
                               float(loss_weight), _p(loss), _p(gt), _p(gw), _p(gr), _p(ws), _stream(trans))
    if rc != 0:
        _drop_zero_workspace(dev, "pair_distill")
    _lib.check(rc, "loco_pair_distill")
    return loss, gt, gw, gr


def token_pool(raw: torch.Tensor, seg_off: torch.Tensor, inv_temp: float, hardmax: bool, want_attention: bool = False):
    """Multi-token class scoring reduction (box_emb_grounding_head.py:130-221): raw [R, Ttot] token logits, seg_off int32 [K1+1]
    (device) -> (scores [R, K1], attention [R, Ttot] or None)."""
    _need_cuda(raw, seg_off)
    if raw.dtype != torch.float32 or raw.dim() != 2 or raw.stride(1) != 1 or seg_off.dtype != torch.int32:
        raise LocoError("token_pool: expects fp32 [R, Ttot] logits with unit column stride and int32 segment offsets")
    r, k1 = raw.shape[0], seg_off.numel() - 1
    scores = torch.empty((r, k1), dtype=torch.float32, device=raw.device)
    att = torch.zeros((r, raw.stride(0) if r else raw.shape[1]), dtype=torch.float32, device=raw.device)[:, :raw.shape[1]] if want_attention else None
    _lib.check(_lib.load().loco_token_pool_fwd(_p(raw), raw.stride(0) if r else raw.shape[1], _p(seg_off), r, k1, float(inv_temp), int(bool(hardmax)),
                                               _p(scores), k1, _p(att), _stream(raw)), "loco_token_pool_fwd")
    return scores, att


def token_pool_backward(raw: torch.Tensor, seg_off: torch.Tensor, inv_temp: float, hardmax: bool, dscores: torch.Tensor) -> torch.Tensor:
    _need_cuda(raw, dscores)
    r, k1 = raw.shape[0], seg_off.numel() - 1
    dscores = dscores.to(torch.float32)
    if dscores.stride(1) != 1:
        dscores = dscores.contiguous()
    draw = torch.zeros_like(raw)          # (token columns outside every segment — there are none — would keep 0)
    _lib.check(_lib.load().loco_token_pool_bwd(_p(raw), raw.stride(0) if r else raw.shape[1], _p(seg_off), r, k1, float(inv_temp), int(bool(hardmax)),
                                               _p(dscores), dscores.stride(0) if r else k1, _p(draw), draw.stride(0) if r else raw.shape[1], _stream(raw)),
               "loco_token_pool_bwd")
    return draw


_small_dev = {}      # small host-built index tensors (image row offsets, image sizes) by (device, kind, values)


def _small_device_tensor(dev, kind, values, dtype):
    key = (dev, kind, values)
    t = _small_dev.get(key)
    if t is None:
        if len(_small_dev) > 256:
            _small_dev.clear()
        t = torch.tensor(values, dtype=dtype).reshape(-1).to(dev)
        _small_dev[key] = t
    return t


def box_inference(probs: torch.Tensor, deltas: torch.Tensor, proposals: torch.Tensor, rows_per_image, image_sizes, reg_weights,
                  scale_clamp: float, score_thresh: float, nms_thresh: float, topk: int):
    """Detectron2 ``fast_rcnn_inference`` for all images of a batch (reference roi_emb_heads.py:280 / :357) in four launches.

    probs [R, K+1] fp32 probabilities, deltas [R, 4], proposals [R, 4]; ``rows_per_image`` python ints, ``image_sizes`` (h, w) pairs.
    Returns (boxes [n,topk,4], scores [n,topk], classes [n,topk] int64, rows [n,topk] int64, counts [n] int32) on the device;
    entries past counts[i] are unspecified."""
    _need_cuda(probs, deltas, proposals)
    n_img = len(rows_per_image)
    r, k1 = probs.shape
    if sum(rows_per_image) != r or deltas.shape[0] != r or proposals.shape != (r, 4) or len(image_sizes) != n_img:
        raise LocoError("box_inference: inconsistent row counts")
    probs = probs if probs.dtype == torch.float32 and probs.stride(1) == 1 else probs.to(torch.float32).contiguous()
    deltas = deltas if deltas.dtype == torch.float32 and deltas.stride(1) == 1 else deltas.to(torch.float32).contiguous()
    proposals = proposals.to(torch.float32).contiguous()
    dev = probs.device
    offs = [0]
    for n in rows_per_image:
        offs.append(offs[-1] + int(n))
    img_off = _small_device_tensor(dev, "off", tuple(offs), torch.int32)
    img_hw = _small_device_tensor(dev, "hw", tuple(float(v) for hw in image_sizes for v in hw), torch.float32)
    max_rows = max([int(n) for n in rows_per_image] + [0])
    boxes = torch.empty((n_img, topk, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((n_img, topk), dtype=torch.float32, device=dev)
    classes = torch.empty((n_img, topk), dtype=torch.int64, device=dev)
    rows = torch.empty((n_img, topk), dtype=torch.int64, device=dev)
    counts = torch.empty((n_img,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    ws = _workspace(dev, lib.loco_box_inference_workspace_bytes(r, k1 - 1, n_img, max_rows, topk), "box_inference")
    w4 = (ctypes.c_float * 4)(*[float(v) for v in reg_weights])
    _lib.check(lib.loco_box_inference(_p(probs), probs.stride(0), _p(deltas), deltas.stride(0), _p(proposals), _p(img_off), _p(img_hw), n_img, max_rows,
                                      r, k1 - 1, w4, float(scale_clamp), float(score_thresh), float(nms_thresh), int(topk),
                                      _p(boxes), _p(scores), _p(classes), _p(rows), _p(counts), _p(ws), _stream(probs)), "loco_box_inference")
    return boxes, scores, classes, rows, counts


def tensor_stats(x: torch.Tensor) -> torch.Tensor:
    """[min, max, mean, std] of a CUDA tensor as a 4-element device tensor, one launch, no host synchronisation."""
    _need_cuda(x)
    xf = x.detach().to(torch.float32).contiguous().reshape(-1)
    out = torch.empty(4, dtype=torch.float32, device=x.device)
    if xf.numel() == 0:
        return out.fill_(float("nan"))
    lib = _lib.load()
    ws = _zero_workspace(x.device, lib.loco_tensor_stats_workspace_bytes(), "tensor_stats")
    rc = lib.loco_tensor_stats(_p(xf), xf.numel(), _p(out), _p(ws), _stream(xf))
    if rc != 0:
        _drop_zero_workspace(x.device, "tensor_stats")
    _lib.check(rc, "loco_tensor_stats")
    return out


PEER_STORE, PEER_SIGNAL, PEER_WAIT = 1, 2, 4


def peer_exchange(segments, peers_dev: int, n_peers: int, flag_peers_dev: int = 0, rank: int = 0, channel: int = 0, mode: int = PEER_STORE,
                  ticket: torch.Tensor = None, device: torch.device = None):
    """Store up to four local 2-D buffers at the same places of every rank's symmetric buffer and / or signal the peers and / or wait for
    their signals (``mode`` bits PEER_STORE | PEER_SIGNAL | PEER_WAIT), in one launch.  ``segments``: list of (src tensor [rows, cols]
    with unit column stride, dst_pitch_bytes, dst_offset_bytes); ``peers_dev`` / ``flag_peers_dev``: device arrays of base pointers
    (symmetric memory)."""
    n = len(segments)
    if n > 4:
        raise LocoError("peer_exchange: at most 4 segments")
    srcs = []
    for t, _, _ in segments:
        _need_cuda(t)
        t2 = t.reshape(1, -1) if t.dim() == 1 else t
        if t2.dim() != 2 or t2.stride(1) != 1:
            raise LocoError("peer_exchange: segments must be 2-D with unit column stride")
        srcs.append(t2)
    m = max(n, 1)
    vp = (ctypes.c_void_p * m)(*[t.data_ptr() for t in srcs])
    sp = (ctypes.c_int64 * m)(*[t.stride(0) * t.element_size() if t.shape[0] > 1 else t.shape[1] * t.element_size() for t in srcs])
    rows = (ctypes.c_int * m)(*[t.shape[0] for t in srcs])
    rb = (ctypes.c_int64 * m)(*[t.shape[1] * t.element_size() for t in srcs])
    dp = (ctypes.c_int64 * m)(*[int(s[1]) for s in segments])
    do = (ctypes.c_int64 * m)(*[int(s[2]) for s in segments])
    dev = srcs[0].device if srcs else device
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.load().loco_peer_exchange(n, vp, sp, rows, rb, dp, do, int(peers_dev), int(n_peers), int(flag_peers_dev) or None, int(rank), int(channel),
                                              int(mode), _p(ticket), stream), "loco_peer_exchange")


def pair_ce(pw: torch.Tensor, cap_mask: torch.Tensor, reg_mask: torch.Tensor, diag_offset: int = 0,
            want_grad: bool = False):
    """Empty-pair guard (in place on pw) + [CE choose caption, CE choose image, acc caption, acc image].
    pw is one matrix [Bc,Bi] -> out [4], or a stack [nmat,Bc,Bi] -> out [nmat,4] (one launch).
    want_grad=True returns (out, d out[...,0]/d pw, d out[...,1]/d pw), else out."""
    _need_cuda(pw, cap_mask, reg_mask)
    single = pw.dim() == 2
    pw3 = pw.unsqueeze(0) if single else pw
    nmat, bc, bi = pw3.shape
    if pw3.stride(2) != 1:
        raise LocoError("pair_ce: pw must have unit column stride")
    cap_mask = cap_mask.to(torch.float32).contiguous()
    reg_mask = reg_mask.to(torch.float32).contiguous()
    out = torch.empty((nmat, 4), dtype=torch.float32, device=pw.device)
    dcap = torch.empty((nmat, bc, bi), dtype=torch.float32, device=pw.device) if want_grad else None
    dimg = torch.empty((nmat, bc, bi), dtype=torch.float32, device=pw.device) if want_grad else None
    lib = _lib.load()
    ws = _zero_workspace(pw.device, lib.loco_pair_ce_workspace_bytes(nmat, bc, bi), "pair_ce") if max(bc, bi) > 32 else None
    rc = lib.loco_pair_ce(_p(pw3), nmat, pw3.stride(0), pw3.stride(1), bc, bi, int(diag_offset), _p(cap_mask),
                          cap_mask.shape[1], _p(reg_mask), reg_mask.shape[1], _p(out), _p(dcap), _p(dimg), _p(ws), _stream(pw))
    if rc != 0 and ws is not None:
        _drop_zero_workspace(pw.device, "pair_ce")
    _lib.check(rc, "loco_pair_ce")
    if single:
        out = out[0]
        dcap = dcap[0] if want_grad else None
        dimg = dimg[0] if want_grad else None
    return (out, dcap, dimg) if want_grad else out
