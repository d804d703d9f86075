"""Developer probe: per-CTA phase timeline of the tensor-core kernels (LOCOV_B200_TIMELINE=1, tc_gemm.cuh) — where the time of
one launch goes: CTA start skew, set-up, first operand stage, main loop, accumulator drain, epilogue.
usage (GPU box): LOCOV_B200_TIMELINE=1 python scripts/gemm_timeline.py"""
import ctypes
import os
import sys

os.environ.setdefault("LOCOV_B200_TIMELINE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from locov_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
NAMES = ["entry", "setup", "stage0", "mma_end", "acc_done", "epi_done", "exit", "-", "e8", "e9", "e10", "e11", "e12", "e13", "e14", "e15"]


def report(tag, fn, reps=5, cold=True):
    rows = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
        buf = np.zeros(8192 * 16, dtype=np.uint64)
        n = lib.loco_debug_timeline_read(buf.ctypes.data_as(ctypes.c_void_p), 8192)
        t = buf[: n * 16].reshape(n, 16).astype(np.int64)
        rows.append(t)
    t = rows[-1]
    t0 = t[:, 0].min()
    ok = t[:, 6] > 0
    print(f"## {tag}: {t.shape[0]} CTAs, span {1e-3 * (t[ok, 6].max() - t0):.1f} us")
    for j, nm in enumerate(NAMES):
        v = t[:, j]
        v = v[v > 0] - t0
        if len(v):
            print(f"   {nm:9s} n={len(v):4d}  min {1e-3 * v.min():7.2f}  median {1e-3 * np.median(v):7.2f}  max {1e-3 * v.max():7.2f} us")
    if (t[:, 15] > 0).any():
        print(f"   raw e14 (cycles inside fold) median {np.median(t[:, 14]):.0f}   e15 (cycles in the TMEM fence + copy) median {np.median(t[:, 15]):.0f}   "
              f"ns of the row pass median {np.median(t[:, 10] - t[:, 8]):.0f}")
    issuers = t[:, 3] > 0
    if issuers.any():
        ml = (t[issuers, 3] - t[issuers, 2]) * 1e-3
        print(f"   main loop (stage0 -> mma_end): median {np.median(ml):.2f} us, max {ml.max():.2f} us")
    ep = t[:, 5] > 0
    if ep.any():
        e = (t[ep, 5] - t[ep, 4]) * 1e-3
        print(f"   epilogue (acc_done -> epi_done): median {np.median(e):.2f} us, max {e.max():.2f} us")


def main():
    M, N, K = 3200, 768, 2048
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.01
    print("# LOCOV_B200_PF =", os.environ.get("LOCOV_B200_PF", "(default)"), " LOCOV_B200_DEBUG =", os.environ.get("LOCOV_B200_DEBUG", "0"))
    if "--lsm-only" in sys.argv:
        return lsm_case()
    report("projection tf32 3200x768x2048 (bf16 out), operands cold (HBM)", lambda: ops.linear_tf32_fwd(x, w, None, want_f32=False, n_bf16=N))
    report("projection tf32 3200x768x2048 (bf16 out), operands L2-resident", lambda: ops.linear_tf32_fwd(x, w, None, want_f32=False, n_bf16=N), cold=False)
    xa, wa = ops.split_bf16(x, False), ops.split_bf16(w, False)
    report("projection bf16 3200x768x2048 (bf16 out), cold", lambda: ops.linear_fwd(xa, wa, None, want_f32=False, n_bf16=N))
    report("projection bf16 3200x768x2048 (bf16 out), L2-resident", lambda: ops.linear_fwd(xa, wa, None, want_f32=False, n_bf16=N), cold=False)
    xh, wh = ops.split_bf16(x, True), ops.split_bf16(w, True)
    report("projection fp32-accurate (3 bf16 passes) 3200x768x2048 (bf16 hi+lo out), cold", lambda: ops.linear_fwd(xh, wh, None, want_f32=False, n_bf16=N, accurate_out=True))
    if "--proj-only" in sys.argv:
        return
    x8 = torch.randn(8192, K, device=dev)
    x8a = ops.split_bf16(x8, False)
    report("linear bf16 8192x768x2048", lambda: ops.linear_fwd(x8a, wa, None, want_f32=False, n_bf16=N))
    lsm_case()
    e5 = ops.split_bf16(torch.randn(8000, 768, device=dev) * 0.3, False)
    c5 = ops.split_bf16(torch.randn(1204, 768, device=dev) * 0.05, False)
    report("box_score 8000 x 1204", lambda: ops.box_score(e5, c5))


def lsm_case():
    B, T, RG, D = 32, 20, 100, 768
    cap = ops.split_bf16(torch.randn(B * T, D, device=dev) * 0.05, False)
    emb = ops.split_bf16(torch.randn(B * RG, D, device=dev) * 0.5, False)
    mc = torch.ones(B, T, device=dev)
    mr = torch.ones(B, RG, device=dev)
    w2r = torch.empty(B, B, device=dev)
    r2w = torch.empty(B, B, device=dev)
    report("lsm_pair B=32 T=20 Rg=100", lambda: ops.lsm_pair(cap, mc, emb, mr, 0.1, out_w2r=w2r, out_r2w=r2w))


if __name__ == "__main__":
    main()
