"""The tcgen05 GEMM core against cuBLAS at the shapes of the path (CUDA-graph replay, operands rotating through more than L2).
bf16 operands -> bf16 output (the projection's hi-only mode) and fp32 operands as TF32 -> bf16 output; cuBLAS: torch.matmul on the
same operands (bf16, and fp32 with allow_tf32).  LOCOV_B200_EPI_TMA=0/1 is read once per process, so each setting runs in a child.
usage (GPU box): python scripts/gemm_vs_cublas.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(3200, 768, 2048), (8192, 768, 2048), (1024, 768, 2048), (8192, 2048, 768)]


def child():
    sys.path.insert(0, ROOT)
    import torch
    from locov_b200 import ops
    from bench import graph_time
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = True
    out = {}
    for M, N, K in SHAPES:
        nrot = max(4, int(200e6 // (M * K * 4)) + 1)
        xf = [torch.randn(M, K, device=dev) for _ in range(nrot)]
        wf = torch.randn(N, K, device=dev) * 0.02
        xs = [ops.split_bf16(x, False) for x in xf]
        w = ops.split_bf16(wf, False)
        xb = [x.hi[:, :K] for x in xs]
        wb = w.hi[:, :K]
        res = {
            "ours_bf16": graph_time(torch, lambda i: ops.linear_fwd(xs[i % nrot], w, None, want_f32=False, n_bf16=N)),
            "cublas_bf16": graph_time(torch, lambda i: torch.matmul(xb[i % nrot], wb.t())),
            "ours_tf32": graph_time(torch, lambda i: ops.linear_tf32_fwd(xf[i % nrot], wf, None, want_f32=False, n_bf16=N)),
            "cublas_tf32": graph_time(torch, lambda i: torch.matmul(xf[i % nrot], wf.t())),
        }
        out[f"{M}x{N}x{K}"] = {k: round(v * 1e3, 2) for k, v in res.items()}
        del xf, xs, xb
        torch.cuda.empty_cache()
    print(json.dumps(out))


def main():
    for tma in ("1", "0"):
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, LOCOV_B200_EPI_TMA=tma), capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else r.stderr[-400:]
        print(f"EPI_TMA={tma} (us per call)", line, flush=True)


if __name__ == "__main__":
    child() if len(sys.argv) > 1 and sys.argv[1] == "child" else main()
