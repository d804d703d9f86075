"""The spatial mean between res5 and the box predictor (reference roi_emb_heads.py:262, :329, :351: ``box_features.mean(dim=[2, 3])``)
as one liblocov_b200 pass that also writes the bf16 operand of the projection GEMM."""
import pytest
import torch

import locov_b200.functional as LF
from locov_b200 import _lib, ops
from util import relerr

pytestmark = pytest.mark.gpu


def _x(r, c, hw, layout, dtype, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(r, c, hw[0], hw[1], generator=g) + 0.3).to(dev)
    if dtype == torch.bfloat16:
        x = x.to(torch.bfloat16)
    if layout == "cl":
        x = x.contiguous(memory_format=torch.channels_last)
    return x


@pytest.mark.parametrize("layout,dtype", [("nchw", torch.float32), ("cl", torch.float32), ("cl", torch.bfloat16)])
@pytest.mark.parametrize("r,c,hw", [(37, 64, (7, 7)), (5, 2048, (7, 7)), (300, 20, (4, 5)), (2, 8, (14, 14)), (1, 4, (1, 1)), (3, 5, (3, 3)),
                                    (2, 8, (50, 76))])
def test_mean_matches_torch_in_double(cuda_device, layout, dtype, r, c, hw):
    x = _x(r, c, hw, layout, dtype, cuda_device, seed=r + c)
    out, op = ops.spatial_mean(x, operand=True)
    ref = x.double().mean(dim=[2, 3])
    assert out.shape == (r, c) and out.dtype == torch.float32
    assert relerr(out.cpu(), ref.cpu()) < (1e-6 if hw[0] * hw[1] <= 196 else 5e-6)       # (3800 sequential fp32 additions at 50 x 76)
    # the operand written by the same pass is exactly the split of the fp32 mean
    want = ops.split_bf16(out, True)
    assert torch.equal(op.hi[:, :c], want.hi[:, :c]) and torch.equal(op.lo[:, :c], want.lo[:, :c])
    assert op.hi.shape[1] % 8 == 0 and (op.hi.shape[1] == c or float(op.hi[:, c:].float().abs().sum()) == 0.0)
    out2, op2 = ops.spatial_mean(x, operand=False)
    assert torch.equal(out2, out) and op2.lo is None and torch.equal(op2.hi, op.hi)
    out3, op3 = ops.spatial_mean(x, operand=None)
    assert torch.equal(out3, out) and op3 is None


def test_non_contiguous_and_other_dtypes_fall_back_to_a_copy(cuda_device):
    x = torch.randn(6, 12, 7, 9, device=cuda_device)[:, :, :, ::2]              # strided view
    out, _ = ops.spatial_mean(x)
    assert relerr(out.cpu(), x.double().mean(dim=[2, 3]).cpu()) < 1e-6
    h = torch.randn(6, 10, 3, 3, device=cuda_device).half()                       # C % 4 != 0 and fp16
    out, _ = ops.spatial_mean(h)
    assert relerr(out.cpu(), h.double().mean(dim=[2, 3]).cpu()) < 1e-6
    e, op = ops.spatial_mean(torch.zeros(0, 8, 7, 7, device=cuda_device), operand=True)
    assert e.shape == (0, 8) and op.hi.shape == (0, 8)


@pytest.mark.parametrize("layout,dtype", [("nchw", torch.float32), ("cl", torch.float32), ("cl", torch.bfloat16)])
def test_gradient_layout_and_values(cuda_device, layout, dtype):
    x = _x(9, 24, (7, 7), layout, dtype, cuda_device).requires_grad_(True)
    n0 = _lib.load().loco_launch_count()
    y = LF.spatial_mean(x)
    w = torch.randn(9, 24, device=cuda_device)
    (y * w).sum().backward()
    assert _lib.load().loco_launch_count() - n0 == 2
    assert x.grad.shape == x.shape and x.grad.dtype == dtype
    want = (w / 49.0)[:, :, None, None].expand(9, 24, 7, 7)
    tol = 1e-7 if dtype == torch.float32 else 4e-3
    assert relerr(x.grad.float().cpu(), want.cpu()) <= tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_operand_is_picked_up_by_the_box_predictor(cuda_device, precision):
    """x = spatial_mean(res5 output) carries its bf16 operand; box_predict must use it (one split launch fewer) and produce the
    bits it produces from the plain fp32 x."""
    torch.manual_seed(3)
    r, v, d, k = 200, 256, 64, 17
    feats = torch.randn(r, v, 7, 7, device=cuda_device).contiguous(memory_format=torch.channels_last)
    we, be = torch.randn(d, v, device=cuda_device) * 0.05, torch.randn(d, device=cuda_device) * 0.1
    wb, bb = torch.randn(4, v, device=cuda_device) * 0.05, torch.zeros(4, device=cuda_device)
    wc, bc = torch.randn(k + 1, d, device=cuda_device), torch.zeros(k + 1, device=cuda_device)
    old = LF.PROJECTION
    LF.PROJECTION = "bf16"                                   # (the TF32 projection of reduced precision reads fp32 x directly)
    try:
        x = LF.spatial_mean(feats, precision)
        assert hasattr(x, "_loco_operand")
        lib = _lib.load()
        LF.box_predict(x.clone(), we, be, wb, bb, wc, bc, precision)      # (fills the weight-operand caches)
        n0 = lib.loco_launch_count()
        s1, d1, _ = LF.box_predict(x, we, be, wb, bb, wc, bc, precision)
        n1 = lib.loco_launch_count()
        s2, d2, _ = LF.box_predict(x.clone(), we, be, wb, bb, wc, bc, precision)
        n2 = lib.loco_launch_count()
        assert (n2 - n1) - (n1 - n0) == 1                    # the split pass
        assert torch.equal(s1, s2) and torch.equal(d1, d2)
        x.add_(1.0)                                          # in-place change: the carried operand is stale and must be ignored
        s3, _, _ = LF.box_predict(x, we, be, wb, bb, wc, bc, precision)
        s4, _, _ = LF.box_predict(x.clone(), we, be, wb, bb, wc, bc, precision)
        assert torch.equal(s3, s4)
    finally:
        LF.PROJECTION = old
