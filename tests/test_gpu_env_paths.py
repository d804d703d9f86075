"""The A/B switches of the library select real alternative code paths (legacy write-out loop of RoIAlign, launches without
programmatic dependent launch, single-CTA GEMM tiles).  Each is run in a fresh process (the switches are read once) on a
small problem and must meet the same parity bars as the default paths."""
import os
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu

_CHECK = r"""
import sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import locov_b200.modeling as M
from locov_b200 import ops
from oracle import lsm_head, roi_align as ora
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(7)
feat = torch.randn(2, 160, 25, 38, generator=g)
rois = torch.tensor([[0, 10.3, 20.7, 200.1, 150.9], [1, 0, 0, 600, 400], [0, -20, -30, 90, 70], [1, 300, 100, 560, 380], [1, 5, 5, 30, 25]])
for ps in (14, 7):
    out = ops.roi_align(feat.to(dev), rois.to(dev), ps, 1 / 16).cpu()
    ref = torch.from_numpy(ora.roi_align_fwd(feat.numpy(), rois.numpy(), ps, 1 / 16, 0, True))
    err = float(((out - ref).abs() / ref.abs().clamp(min=ref.pow(2).mean().sqrt())).max())
    assert err < 1e-4, ("roi_align", ps, err)
fg = feat[:, :24].to(dev).requires_grad_(True)
out = M.ROIAlign(7, 1 / 16, 0, True)(fg, rois.to(dev))
dout = torch.randn(out.shape, generator=g).to(dev)
out.backward(dout)
ref = torch.from_numpy(ora.roi_align_bwd(dout.cpu().numpy(), (2, 24, 25, 38), rois.numpy(), 1 / 16, 0, True))
err = float(((fg.grad.cpu() - ref).abs() / ref.abs().clamp(min=ref.pow(2).mean().sqrt())).max())
assert err < 1e-4, ("roi_align_bwd", err)
ii, ic, w, b = lsm_head.make_lsm_inputs(B=5, Rg=40, T=9, V=256, D=768, seed=3, gain=5.0, ragged_regions=True)
for precision, tol in (("fp32", 1e-4), ("bf16", 2e-2)):
    cfg = M.get_cfg("lsm"); cfg.MODEL.B200.PRECISION = precision
    head = M.GroundingHead(cfg, 256, 768).to(dev)
    with torch.no_grad():
        head.v2l_projection.weight.copy_(w); head.v2l_projection.bias.copy_(b)
        info, losses, dists = head({k: v.to(dev) for k, v in ii.items()}, {k: v.to(dev) for k, v in ic.items()})
    _, rl, rd = lsm_head.grounding_head_forward(ii, ic, w, b, dtype=torch.float64)
    for k in rd:
        a, r = dists[k].double().cpu(), rd[k]
        err = float(((a - r).abs() / r.abs().clamp(min=r.pow(2).mean().sqrt())).max())
        assert err < tol, ("lsm", precision, k, err)
print("env path ok")
"""


@pytest.mark.parametrize("env", [{"LOCOV_B200_PDL": "0"}, {"LOCOV_B200_ROI_BULK": "0"}, {"LOCOV_B200_ROI_BWD_V4": "0"}, {"LOCOV_B200_2CTA": "0"},
                                 {"LOCOV_B200_2CTA": "0", "LOCOV_B200_CLUSTER": "1", "LOCOV_B200_CM": "1", "LOCOV_B200_CN": "2"}])
def test_alternative_paths_meet_the_parity_bars(cuda_device, env):
    r = subprocess.run([sys.executable, "-c", _CHECK % {"root": ROOT}], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "env path ok" in r.stdout, (env, r.stdout[-1500:], r.stderr[-1500:])
