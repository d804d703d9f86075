"""Developer sweep of the tcgen05 GEMM core: tile width / cluster shape / pipeline depth vs time, with cuBLAS bf16
beside it.  Each configuration runs in a subprocess because the knobs are read once from the environment.
usage (GPU box): python scripts/gemm_sweep.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import torch
    from locov_b200 import ops
    M, N, K = [int(x) for x in sys.argv[2:5]]
    dev = torch.device("cuda:0")
    xs = [ops.split_bf16(torch.randn(M, K, device=dev), False) for _ in range(4)]
    w = ops.split_bf16(torch.randn(N, K, device=dev) * 0.01, False)

    def run(i):
        ops.linear_fwd(xs[i % 4], w, None, want_f32=False, n_bf16=N)
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(50):
        run(i)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    xa = [x.hi[:, :K] for x in xs]
    wb = w.hi[:, :K]
    for i in range(5):
        torch.matmul(xa[i % 4], wb.t())
    torch.cuda.synchronize()
    a.record()
    for i in range(50):
        torch.matmul(xa[i % 4], wb.t())
    b.record()
    torch.cuda.synchronize()
    ms_lib = a.elapsed_time(b) / 50
    print(json.dumps({"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "cublas_ms": ms_lib, "cublas_tflops": 2.0 * M * N * K / ms_lib / 1e9}))


def main():
    shapes = [(3200, 768, 2048), (8192, 768, 2048), (32768, 768, 2048), (8192, 8192, 2048)]
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        shapes = shapes[:2]
    cfgs = [dict(), dict(LOCOV_B200_2CTA="0")]
    for bn in (128, 192, 256):
        cfgs.append(dict(LOCOV_B200_BN=str(bn)))
    for cm, cn, bn in ((1, 2, 192), (1, 4, 192), (2, 2, 192), (2, 2, 256), (1, 2, 256), (2, 1, 256), (1, 3, 256)):
        cfgs.append(dict(LOCOV_B200_2CTA="0", LOCOV_B200_CLUSTER="1", LOCOV_B200_CM=str(cm), LOCOV_B200_CN=str(cn), LOCOV_B200_BN=str(bn)))
    for shp in shapes:
        for cfg in cfgs:
            env = dict(os.environ, **cfg)
            r = subprocess.run([sys.executable, __file__, "child"] + [str(x) for x in shp], env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else (r.stderr[-300:])
            print(shp, {k[11:]: v for k, v in cfg.items()}, line, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        main()
