"""Image-sharded LSM head over NCCL on two real GPUs (skipped on a single-GPU box): each rank's block equals its
column slice of the single-device pair matrices, global losses / accuracies equal the single-device ones
(scripts/check_sharded_nccl.py holds the check; fp32 mode 1e-4, bf16 mode 2e-2)."""
import os
import subprocess
import sys

import pytest
import torch

from util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("symm,shape", [("1", "small"), ("0", "small"), ("1", "bench")])
def test_sharded_lsm_world2(symm, shape):
    """symm = 1: symmetric-memory peer stores (falls back to NCCL with a warning where unavailable); 0: NCCL all-gathers."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + (os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "scripts", "check_sharded_nccl.py"), shape],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, LOCOV_B200_SYMM=symm))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sharded NCCL parity: OK" in r.stdout
