"""ctypes binding of liblocov_b200.so — the ONLY way the package reaches the GPU.

There is no CPU or PyTorch fallback: if the shared library is missing and cannot be built, or a call
fails, a ``LocoError`` is raised.  Signatures mirror include/locov_b200.h one to one.
"""
import ctypes
import os
import threading

from . import build as _build


class LocoError(RuntimeError):
    pass


_c = ctypes
_vp, _i, _i64, _f = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float

# name -> (restype, argtypes); kept in the order of include/locov_b200.h
SIGNATURES = {
    "loco_version": (_i, []),
    "loco_last_error": (_c.c_char_p, []),
    "loco_device_check": (_i, [_i]),
    "loco_sm_count": (_i, [_i]),
    "loco_launch_count": (_c.c_longlong, []),
    "loco_debug_timeline_read": (_i, [_vp, _i]),
    "loco_roi_align_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, _i]),
    "loco_roi_align_fwd": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _f, _i, _i, _vp, _i, _i, _vp, _vp]),
    "loco_roi_align_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "loco_roi_align_bwd": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _f, _i, _i, _vp, _vp, _vp]),
    "loco_spatial_mean": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _i64, _vp, _vp, _i64, _vp]),
    "loco_spatial_mean_bwd": (_i, [_vp, _i64, _i64, _i, _i, _i, _i, _vp, _vp]),
    "loco_box_inference_workspace_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "loco_box_inference": (_i, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _f, _f, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "loco_token_pool_fwd": (_i, [_vp, _i64, _vp, _i, _i, _f, _i, _vp, _i64, _vp, _vp]),
    "loco_token_pool_bwd": (_i, [_vp, _i64, _vp, _i, _i, _f, _i, _vp, _i64, _vp, _i64, _vp]),
    "loco_lsm_prep_multi": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i, _i64, _vp, _vp, _vp]),
    "loco_roi_align_grid_dump": (_i, [_vp, _i, _i, _i, _i, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "loco_split_bf16": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _i, _vp]),
    "loco_transpose_bf16": (_i, [_vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "loco_linear_fwd": (_i, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _i, _i, _i, _vp, _i64, _vp, _vp, _i, _i64, _vp]),
    "loco_linear_tf32_fwd": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _i, _i, _vp, _i64, _vp, _vp, _i, _i64, _vp]),
    "loco_box_score_workspace_bytes": (_i64, [_i, _i]),
    "loco_box_score_fwd": (_i, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _i, _i, _i, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "loco_box_softmax": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _i64, _vp]),
    "loco_row_normalize": (_i, [_vp, _i64, _i, _i, _i, _vp, _i64, _vp, _i64, _vp]),
    "loco_box_ce_fwd_bwd": (_i, [_vp, _i64, _vp, _vp, _i, _i, _f, _vp, _f, _vp, _vp, _i64, _vp]),
    "loco_box_reg_loss_workspace_bytes": (_i64, [_i]),
    "loco_box_reg_loss": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i, _c.POINTER(_f), _f, _f, _vp, _vp, _i64, _vp, _vp]),
    "loco_skinny_grad_workspace_bytes": (_i64, [_i, _i, _i]),
    "loco_skinny_grad": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "loco_lsm_masks": (_i, [_vp, _vp, _i64, _vp, _i, _i64, _vp, _vp, _vp]),
    "loco_lsm_prep": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _i, _i64, _vp, _vp, _vp]),
    "loco_lsm_pair_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "loco_lsm_pair_fwd": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _i64, _vp, _vp]),
    "loco_lsm_pair_bwd": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _i64, _vp, _vp, _i64,
                                _vp, _vp, _i64, _vp]),
    "loco_peer_exchange": (_i, [_i, _c.POINTER(_vp), _c.POINTER(_i64), _c.POINTER(_i), _c.POINTER(_i64), _c.POINTER(_i64), _c.POINTER(_i64), _vp, _i, _vp, _i, _i,
                                _i, _vp, _vp]),
    "loco_pair_distill_workspace_bytes": (_i64, [_i]),
    "loco_pair_distill": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i, _f, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "loco_tensor_stats_workspace_bytes": (_i64, []),
    "loco_tensor_stats": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "loco_pair_ce_workspace_bytes": (_i64, [_i, _i, _i]),
    "loco_pair_ce": (_i, [_vp, _i, _i64, _i64, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


def lib_path():
    return _build.lib_path()


def load():
    """Load (building first if needed) the C-ABI library.  Raises LocoError when impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.lib_path()
        # build() is a no-op while the stamp matches the sources; after a source / header edit the library is rebuilt
        # before it is bound, so the ctypes signatures below can never meet a stale binary (ABI drift)
        if not _build.is_current():
            try:
                _build.build()
            except Exception as e:  # no fallback: fail loudly
                if not os.path.exists(path):
                    raise LocoError(f"liblocov_b200.so is missing and could not be built: {e}") from e
                raise LocoError(f"liblocov_b200.so is older than its sources and could not be rebuilt: {e}") from e
        try:
            lib = ctypes.CDLL(path)
        except OSError as e:
            raise LocoError(f"cannot load {path}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise LocoError(f"{path} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().loco_last_error().decode("utf-8", "replace")
        raise LocoError(f"{what} failed with code {rc}: {msg}")
