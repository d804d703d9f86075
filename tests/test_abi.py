"""The C-ABI library loads without a GPU and exports exactly the entry points include/locov_b200.h
declares; argument validation answers before any CUDA call."""
import ctypes
import os
import re

from locov_b200 import _lib, build
from util import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "locov_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(loco_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in locov_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"


def test_version_and_error_plumbing():
    lib = _lib.load()
    assert lib.loco_version() >= 100
    rc = lib.loco_roi_align_fwd(None, 0, 0, 0, 0, 0, None, 0, 7, 7, 0.0625, 0, 1, None, 0, 0, None, None)
    assert rc == -1                                    # LOCO_E_BADARG
    assert b"roi_align_fwd" in lib.loco_last_error()
    rc = lib.loco_lsm_pair_fwd(None, None, 0, None, None, None, 0, None, 4, 200, 4, 10, 64, 0.1, 0, None, None, 4, None, None)
    assert rc == -2                                    # LOCO_E_UNSUPPORTED: T > 128
    assert lib.loco_roi_align_workspace_bytes(2, 8, 5, 6, 0, 0) == 2048            # the transposed map (1920 B), rounded to 256
    assert lib.loco_roi_align_workspace_bytes(2, 8, 5, 6, 1, 0) == 0
    assert lib.loco_roi_align_workspace_bytes(2, 8, 5, 6, 1, 100) == 512           # + the launch order of 100 rois


def test_sass_uses_blackwell_tensor_core_and_tma_instructions():
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG (B200_PROFILING.md)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.build()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    assert "sm_100a" in sass
