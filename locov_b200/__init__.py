"""locov_b200 — B200 (sm_100a) implementation of LocOV's region-text matching hot path.

Layers (bottom up):
  csrc/ + include/locov_b200.h : hand-written CUDA kernels behind a C ABI (liblocov_b200.so)
  _lib.py, ops.py              : ctypes binding and torch-tensor front-ends (no fallback paths)
  functional.py                : torch.autograd Functions over those entry points
  modeling/                    : drop-in mirrors of the reference's ovr/modeling modules for this path
  parallel.py                  : batch-sharded LSM pair matrix (peer-memory / NCCL caption exchange over NVLink)
  host_feed.py                 : pinned-host -> device staging ring (H2D of batch i+1 under the kernels of batch i)
"""
from ._lib import LocoError  # noqa: F401
from .host_feed import HostFeed  # noqa: F401

__version__ = "0.1.0"
