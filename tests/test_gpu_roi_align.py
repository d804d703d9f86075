"""RoIAlign parity on the B200 (through the C ABI via locov_b200.ops / the ROIPooler module).
Bar: sampling-grid coordinates and integer tap indices BIT-EXACT vs the float32 oracle; pooled values
within 1e-4 robust-relative (measured: ~1e-6)."""
import numpy as np
import pytest
import torch

from locov_b200 import ops
import locov_b200.modeling as M
from oracle import roi_align as ora
from util import load_golden, relerr

pytestmark = pytest.mark.gpu


def _rois(n_img, r, h_img, w_img, seed):
    g = torch.Generator().manual_seed(seed)
    cx = torch.rand(r, generator=g) * w_img
    cy = torch.rand(r, generator=g) * h_img
    s = 16 * (600 / 16) ** torch.rand(r, generator=g)
    a = 0.5 * 4 ** torch.rand(r, generator=g)
    bw, bh = s * a.sqrt(), s / a.sqrt()
    b = torch.stack([(cx - bw / 2).clamp(0, w_img), (cy - bh / 2).clamp(0, h_img), (cx + bw / 2).clamp(0, w_img),
                     (cy + bh / 2).clamp(0, h_img)], 1)
    return torch.cat([torch.randint(0, n_img, (r, 1), generator=g).float(), b], 1)


EDGE = torch.tensor([[0, 100, 100, 100, 100], [1, 100, 100, 101, 160], [0, -50, -40, 30, 20], [1, 2000, 2000, 2100, 2100],
                     [0, 0, 0, 1216, 800], [1, 1200, 780, 1216, 800], [0, 300, 200, 280, 180], [1, 5.5, 5.5, 6.5, 6.5],
                     [0, -500, -500, 1800, 1400]], dtype=torch.float32)


@pytest.mark.parametrize("tag,cfg", [("p7_s16_sr0_al1", (7, 1 / 16, 0, True)), ("p14_s16_sr0_al1", (14, 1 / 16, 0, True)),
                                     ("p7_s16_sr2_al1", (7, 1 / 16, 2, True)), ("p5_s16_sr0_al0", (5, 1 / 16, 0, False))])
def test_golden_vectors(cuda_device, tag, cfg):
    z = load_golden("roi_align")
    ps, scale, sr, al = cfg
    out = ops.roi_align(torch.from_numpy(z["feat"]).to(cuda_device), torch.from_numpy(z["rois"]).to(cuda_device), ps, scale, sr, al)
    assert relerr(out.cpu(), z["out_" + tag]) < 1e-5


@pytest.mark.parametrize("ps,scale,hw", [(14, 1 / 16, (50, 76)), (7, 1 / 32, (25, 38))])
def test_sampling_grid_bit_exact(cuda_device, ps, scale, hw):
    rois = torch.cat([_rois(2, 200, 800, 1216, seed=ps), EDGE])
    h, w = hw
    ghw, yx, idx = ops.roi_align_grid(rois.to(cuda_device), h, w, ps, scale, 0, True, max_grid=6)
    ghw, yx, idx = ghw.cpu().numpy(), yx.cpu().numpy(), idx.cpu().numpy()
    checked = 0
    for r in range(rois.shape[0]):
        g2, yx2, idx2 = ora.roi_align_grid(rois[r].numpy(), h, w, ps, scale, 0, True, max_samples=1 << 20)
        assert tuple(g2) == tuple(ghw[r])
        gh, gw = int(g2[0]), int(g2[1])
        if gh <= 0 or gw <= 0 or gh > 6 or gw > 6:
            continue
        a = np.ascontiguousarray(yx[r][:, :, :gh, :gw]).reshape(-1, 2)
        b = np.ascontiguousarray(idx[r][:, :, :gh, :gw]).reshape(-1, 4)
        assert np.array_equal(a.view(np.uint32), yx2.view(np.uint32)), f"roi {r}: coordinates differ"
        assert np.array_equal(b, idx2), f"roi {r}: tap indices differ"
        checked += a.shape[0]
    assert checked > 15000


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("c,ps,scale,hw,sr", [(96, 14, 1 / 16, (50, 76), 0), (70, 7, 1 / 32, (25, 38), 0), (33, 7, 1 / 16, (50, 76), 2)])
def test_values_vs_oracle(cuda_device, layout, c, ps, scale, hw, sr):
    torch.manual_seed(c)
    feat = torch.randn(2, c, *hw)
    rois = torch.cat([_rois(2, 64, 800, 1216, seed=c), EDGE])
    f = feat.to(cuda_device)
    if layout == "nhwc":
        f = f.contiguous(memory_format=torch.channels_last)
    out = ops.roi_align(f, rois.to(cuda_device), ps, scale, sr, True).cpu()
    ref = torch.from_numpy(ora.roi_align_fwd(feat.numpy(), rois.numpy(), ps, scale, sr, True))
    assert out.shape == ref.shape
    assert relerr(out, ref) < 1e-4
    assert (out - ref).abs().max() < 1e-5


def test_empty_and_module_interface(cuda_device):
    feat = torch.randn(2, 16, 20, 30, device=cuda_device)
    out = ops.roi_align(feat, torch.zeros(0, 5, device=cuda_device), 7, 1 / 16)
    assert out.shape == (0, 16, 7, 7)
    pool = M.ROIPooler(output_size=7, scales=(1 / 16,), sampling_ratio=0, pooler_type="ROIAlignV2")
    boxes = [M.Boxes(torch.tensor([[10., 20., 200., 150.], [0., 0., 480., 320.]], device=cuda_device)),
             M.Boxes(torch.tensor([[17.2, 33.3, 460.8, 300.1]], device=cuda_device))]
    got = pool([feat], boxes).cpu()
    rois = torch.tensor([[0, 10., 20., 200., 150.], [0, 0., 0., 480., 320.], [1, 17.2, 33.3, 460.8, 300.1]])
    ref = torch.from_numpy(ora.roi_align_fwd(feat.cpu().numpy(), rois.numpy(), 7, 1 / 16, 0, True))
    assert relerr(got, ref) < 1e-5


@pytest.mark.parametrize("c,ps,sr", [(24, 7, 0), (136, 14, 0), (22, 7, 0), (40, 6, 2)])
def test_backward_vs_oracle(cuda_device, c, ps, sr):
    """c % 4 == 0: channels-last vector-atomic kernel (adjoint column walk; c = 136 spans two 128-channel slabs);
    c = 22 and the fixed sampling ratio: per-element scatter (directly, resp. through the `handled` flags)."""
    torch.manual_seed(5)
    feat = torch.randn(2, c, 25, 38, device=cuda_device, requires_grad=True)
    rois = torch.cat([_rois(2, 40, 400, 608, seed=9), EDGE]).to(cuda_device)      # EDGE[-1] has a 6-wide sampling grid
    pool = M.ROIAlign(ps, 1 / 16, sr, True)
    out = pool(feat, rois)
    dout = torch.randn_like(out)
    out.backward(dout)
    ref = ora.roi_align_bwd(dout.cpu().numpy(), (2, c, 25, 38), rois.cpu().numpy(), 1 / 16, sr, True)
    assert relerr(feat.grad.cpu(), ref) < 1e-4


@pytest.mark.parametrize("ps", [16, 28])
def test_backward_large_pooled_sizes(cuda_device, ps):
    """Pooled sizes whose gradient staging exceeds shared memory (15 x 15 and above): the library declines the vectorised
    kernel and accumulates with scalar atomics into the (zero-filled) map — same result."""
    torch.manual_seed(ps)
    feat = torch.randn(1, 8, 25, 38, device=cuda_device, requires_grad=True)
    rois = torch.cat([_rois(1, 12, 400, 608, seed=ps), EDGE[:3] * torch.tensor([0, 1, 1, 1, 1.0])]).to(cuda_device)
    out = M.ROIAlign(ps, 1 / 16, 0, True)(feat, rois)
    dout = torch.randn_like(out)
    out.backward(dout)
    ref = ora.roi_align_bwd(dout.cpu().numpy(), (1, 8, 25, 38), rois.cpu().numpy(), 1 / 16, 0, True)
    assert relerr(feat.grad.cpu(), ref) < 1e-4


def test_backward_without_rois_is_zero(cuda_device):
    g = ops.roi_align_backward(torch.zeros(0, 8, 7, 7, device=cuda_device), (2, 8, 20, 30), torch.zeros(0, 5, device=cuda_device), 1 / 16)
    assert g.shape == (2, 8, 20, 30) and float(g.abs().max()) == 0.0
    feat = torch.randn(2, 8, 20, 30, device=cuda_device, requires_grad=True)
    out = M.ROIAlign(7, 1 / 16, 0, True)(feat, torch.zeros(0, 5, device=cuda_device))
    out.sum().backward()
    assert float(feat.grad.abs().max()) == 0.0


def test_full_size_properties(cuda_device):
    """BASELINE config 1 size (2 x 512 RoIs, [2,1024,50,76] -> [1024,1024,14,14]): linearity and
    agreement with the oracle on a channel subset (the oracle is scalar C: a subset keeps it in seconds)."""
    g = torch.Generator().manual_seed(1992)
    feat = torch.randn(2, 1024, 50, 76, generator=g)
    rois = _rois(2, 1024, 800, 1216, seed=1992)
    f, r = feat.to(cuda_device), rois.to(cuda_device)
    a = ops.roi_align(f, r, 14, 1 / 16)
    b = ops.roi_align(2.0 * f, r, 14, 1 / 16)
    assert torch.equal(b, 2.0 * a)                       # exact: scaling by 2 commutes with every rounding
    ones = ops.roi_align(torch.ones_like(f[:, :32]), r, 14, 1 / 16)
    assert float((ones - 1).abs().max()) < 1e-5           # in-image boxes: bilinear weights sum to one
    # every channel against the oracle for a spread of RoIs (all 1024 channels x 64 RoIs), a channel subset for all RoIs
    pick = torch.arange(0, 1024, 16)
    ref = torch.from_numpy(ora.roi_align_fwd(feat.numpy(), rois[pick].numpy(), 14, 1 / 16, 0, True))
    assert relerr(a[pick].cpu(), ref) < 1e-4
    sub = [0, 1, 127, 128, 511, 512, 1023]
    ref = torch.from_numpy(ora.roi_align_fwd(feat[:, sub].numpy(), rois.numpy(), 14, 1 / 16, 0, True))
    assert relerr(a[:, sub].cpu(), ref) < 1e-4


@pytest.mark.parametrize("feat_layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("c,ps,scale,hw,sr", [(96, 14, 1 / 16, (50, 76), 0), (1024, 7, 1 / 16, (25, 38), 0), (36, 7, 1 / 32, (25, 38), 2),
                                              (256, 14, 1 / 16, (50, 76), 0)])
def test_channels_last_output(cuda_device, feat_layout, c, ps, scale, hw, sr):
    """the pooled tensor in torch.channels_last memory format (fp32 / bf16) holds the numbers of the NCHW output."""
    torch.manual_seed(c + ps)
    feat = torch.randn(2, c, *hw)
    rois = torch.cat([_rois(2, 96, 800, 1216, seed=c), EDGE]).to(cuda_device)
    f = feat.to(cuda_device)
    if feat_layout == "nhwc":
        f = f.contiguous(memory_format=torch.channels_last)
    nchw = ops.roi_align(f, rois, ps, scale, sr, True)
    cl = ops.roi_align(f, rois, ps, scale, sr, True, channels_last=True)
    assert cl.shape == nchw.shape and cl.dtype == torch.float32
    assert cl.is_contiguous(memory_format=torch.channels_last)
    assert float((cl - nchw).abs().max()) <= 1e-6
    ref = torch.from_numpy(ora.roi_align_fwd(feat.numpy(), rois.cpu().numpy(), ps, scale, sr, True))
    assert relerr(cl.cpu(), ref) < 1e-4 and float((cl.cpu() - ref).abs().max()) < 1e-5
    b = ops.roi_align(f, rois, ps, scale, sr, True, channels_last=True, out_dtype=torch.bfloat16)
    assert b.dtype == torch.bfloat16 and b.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(b, cl.to(torch.bfloat16))              # round-to-nearest-even of the same fp32 value


def test_channels_last_output_edge_cases(cuda_device):
    from locov_b200._lib import LocoError
    feat = torch.randn(2, 16, 20, 30, device=cuda_device)
    empty = ops.roi_align(feat, torch.zeros(0, 5, device=cuda_device), 7, 1 / 16, 0, True, channels_last=True, out_dtype=torch.bfloat16)
    assert empty.shape == (0, 16, 7, 7) and empty.dtype == torch.bfloat16
    with pytest.raises(LocoError):                            # 4-channel lanes: C % 4 == 0
        ops.roi_align(torch.randn(1, 6, 20, 30, device=cuda_device), torch.tensor([[0, 1., 1., 90., 80.]], device=cuda_device), 7, 1 / 16, 0, True,
                      channels_last=True)
    with pytest.raises(LocoError):                            # bf16 only in the channels-last layout
        ops.roi_align(feat, torch.tensor([[0, 1., 1., 90., 80.]], device=cuda_device), 7, 1 / 16, 0, True, out_dtype=torch.bfloat16)
    # module interface + gradient through the channels-last output
    pooler = M.ROIPooler(7, (1 / 16,), 0, "ROIAlignV2", channels_last=True)
    f = feat.clone().requires_grad_(True)
    boxes = [torch.tensor([[10., 20., 200., 220.], [5., 5., 400., 300.]], device=cuda_device), torch.tensor([[50., 60., 300., 310.]], device=cuda_device)]
    y = pooler([f], boxes)
    assert y.shape == (3, 16, 7, 7) and y.is_contiguous(memory_format=torch.channels_last)
    f2 = feat.clone().requires_grad_(True)
    y2 = M.ROIPooler(7, (1 / 16,), 0, "ROIAlignV2")([f2], boxes)
    w = torch.randn_like(y2)
    (y * w).sum().backward()
    (y2 * w).sum().backward()
    assert float((f.grad - f2.grad).abs().max()) <= 1e-5
