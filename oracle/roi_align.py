"""ORACLE (test infrastructure only): ctypes front-end of oracle/roi_align_oracle.c.

Follows /root/reference/ovr/modeling/roi_heads/roi_emb_heads.py:182-187,243-245 (ROIPooler ->
Detectron2 ROIAlign -> torchvision.ops.roi_align, aligned=True).  See the C file header.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_roi_align.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "roi_align_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_roi_align.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        _lib.oracle_roi_align_fwd.argtypes = [fp] + [ctypes.c_int] * 4 + [fp] + [ctypes.c_int] * 3 + [
            ctypes.c_float, ctypes.c_int, ctypes.c_int, fp]
        _lib.oracle_roi_align_bwd.argtypes = _lib.oracle_roi_align_fwd.argtypes
        _lib.oracle_roi_align_grid.argtypes = [fp] + [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_int,
                                                                            ctypes.c_int, ip, fp, ip, ctypes.c_int]
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def roi_align_fwd(feat, rois, output_size, spatial_scale, sampling_ratio=0, aligned=True):
    """feat [N,C,H,W] float32, rois [R,5] float32 -> [R,C,PH,PW] float32."""
    lib = _load()
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    n, c, h, w = feat.shape
    out = np.empty((rois.shape[0], c, ph, pw), dtype=np.float32)
    lib.oracle_roi_align_fwd(_fp(feat), n, c, h, w, _fp(rois), rois.shape[0], ph, pw,
                             np.float32(spatial_scale), int(sampling_ratio), int(bool(aligned)), _fp(out))
    return out


def roi_align_bwd(dout, feat_shape, rois, spatial_scale, sampling_ratio=0, aligned=True):
    lib = _load()
    dout = np.ascontiguousarray(dout, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    n, c, h, w = feat_shape
    r, c2, ph, pw = dout.shape
    assert c2 == c and r == rois.shape[0]
    dfeat = np.zeros(feat_shape, dtype=np.float32)
    lib.oracle_roi_align_bwd(_fp(dout), n, c, h, w, _fp(rois), r, ph, pw, np.float32(spatial_scale),
                             int(sampling_ratio), int(bool(aligned)), _fp(dfeat))
    return dfeat


def roi_align_grid(roi, h, w, output_size, spatial_scale, sampling_ratio=0, aligned=True, max_samples=1 << 16):
    """Sampling grid of ONE roi.  Returns (grid_hw int32[2], yx float32[S,2], idx int32[S,4]) with
    S = PH*PW*gh*gw samples in (ph, pw, iy, ix) nesting order; idx rows are
    (y_low, x_low, y_high, x_high), -1 for samples the reference skips (outside [-1,H]x[-1,W])."""
    lib = _load()
    roi = np.ascontiguousarray(roi, dtype=np.float32).reshape(5)
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    ghw = np.zeros(2, dtype=np.int32)
    yx = np.zeros((max_samples, 2), dtype=np.float32)
    idx = np.zeros((max_samples, 4), dtype=np.int32)
    n = lib.oracle_roi_align_grid(_fp(roi), h, w, ph, pw, np.float32(spatial_scale), int(sampling_ratio),
                                  int(bool(aligned)), _ip(ghw), _fp(yx), _ip(idx), max_samples)
    if n < 0:
        raise ValueError("roi needs more than max_samples samples")
    return ghw, yx[:n].copy(), idx[:n].copy()
