"""Pins oracle/roi_align_oracle.c bit-for-bit against torchvision's compiled CPU op (the third-party
kernel the reference reaches through Detectron2 ROIPooler -> ROIAlign) — live and through the
committed golden outputs."""
import numpy as np
import pytest
import torch

from oracle import roi_align as ora
from util import load_golden

torchvision = pytest.importorskip("torchvision")

CFG = {"p7_s16_sr0_al1": (7, 1 / 16, 0, True), "p14_s16_sr0_al1": (14, 1 / 16, 0, True),
       "p7_s16_sr2_al1": (7, 1 / 16, 2, True), "p5_s16_sr0_al0": (5, 1 / 16, 0, False)}


@pytest.mark.parametrize("tag", sorted(CFG))
def test_oracle_bit_exact_vs_golden(tag):
    z = load_golden("roi_align")
    ps, scale, sr, al = CFG[tag]
    out = ora.roi_align_fwd(z["feat"], z["rois"], ps, scale, sr, al)
    assert np.array_equal(out.view(np.uint32), z["out_" + tag].view(np.uint32))


def _random_rois(n_img, r, h_img, w_img, seed):
    g = torch.Generator().manual_seed(seed)
    cx = torch.rand(r, generator=g) * w_img
    cy = torch.rand(r, generator=g) * h_img
    s = 16 * (600 / 16) ** torch.rand(r, generator=g)
    a = 0.5 * 4 ** torch.rand(r, generator=g)
    bw, bh = s * a.sqrt(), s / a.sqrt()
    boxes = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    boxes[: r // 8] += 0.75 * w_img               # partly / fully outside
    idx = torch.randint(0, n_img, (r, 1), generator=g).float()
    return torch.cat([idx, boxes], 1)


@pytest.mark.parametrize("ps,scale,sr,aligned", [(14, 1 / 16, 0, True), (7, 1 / 32, 0, True), (7, 1 / 16, 3, False)])
def test_oracle_bit_exact_vs_live_torchvision(ps, scale, sr, aligned):
    torch.manual_seed(0)
    feat = torch.randn(2, 5, 25, 38)
    rois = _random_rois(2, 96, 400, 608, seed=ps)
    ref = torch.ops.torchvision.roi_align(feat, rois, scale, ps, ps, sr, aligned).numpy()
    out = ora.roi_align_fwd(feat.numpy(), rois.numpy(), ps, scale, sr, aligned)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_oracle_backward_matches_torchvision():
    torch.manual_seed(1)
    feat = torch.randn(2, 3, 12, 17)
    rois = _random_rois(2, 20, 192, 272, seed=3)
    dout = torch.randn(20, 3, 7, 7)
    ref = torch.ops.torchvision._roi_align_backward(dout, rois, 1 / 16, 7, 7, 2, 3, 12, 17, 0, True)
    got = ora.roi_align_bwd(dout.numpy(), (2, 3, 12, 17), rois.numpy(), 1 / 16, 0, True)
    assert np.allclose(got, ref.numpy(), rtol=1e-5, atol=1e-6)


def test_grid_dump_consistency():
    roi = np.array([0, 10.3, 20.7, 200.1, 150.9], np.float32)
    ghw, yx, idx = ora.roi_align_grid(roi, 20, 30, 7, 1 / 16, 0, True)
    assert yx.shape[0] == 49 * ghw[0] * ghw[1] and idx.shape == (yx.shape[0], 4)
    ok = idx[:, 0] >= 0
    assert (idx[ok, 2] - idx[ok, 0]).max() <= 1 and (idx[ok, 3] - idx[ok, 1]).max() <= 1
    assert idx[ok, 2].max() <= 19 and idx[ok, 3].max() <= 29
