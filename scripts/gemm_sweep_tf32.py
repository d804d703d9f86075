"""Developer sweep of the TF32 (fp32 operands in place) projection GEMM: tile width / cluster shape vs time, CUDA-graph
timed (20 launches per replay, rotating inputs larger than L2).  usage (GPU box): python scripts/gemm_sweep_tf32.py"""
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
    nrot = max(4, int(200e6 // (M * K * 4)) + 1)
    xs = [torch.randn(M, K, device=dev) for _ in range(nrot)]
    w = torch.randn(N, K, device=dev) * 0.01

    def run(i):
        ops.linear_tf32_fwd(xs[i % nrot], w, None, want_f32=False, n_bf16=N - N % 8)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(20):
            run(i)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 20)
    print(json.dumps({"us": round(best * 1e3, 1), "tflops": round(2.0 * M * N * K / best / 1e9, 1)}))


def main():
    shapes = [(3200, 768, 2048), (8000, 772, 2048)]
    cfgs = [dict(), dict(LOCOV_B200_2CTA="0")]
    for bn in (128, 192, 256):
        cfgs.append(dict(LOCOV_B200_BN=str(bn)))
    for cm, cn, bn in ((1, 2, 192), (1, 4, 192), (2, 2, 192), (2, 2, 256), (1, 2, 256), (2, 1, 256), (1, 3, 256)):
        cfgs.append(dict(LOCOV_B200_2CTA="0", LOCOV_B200_CLUSTER="1", LOCOV_B200_CM=str(cm), LOCOV_B200_CN=str(cn), LOCOV_B200_BN=str(bn)))
    for shp in shapes:
        for cfg in cfgs:
            env = dict(os.environ, **cfg)
            r = subprocess.run([sys.executable, __file__, "child"] + [str(x) for x in shp], env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else ("ERR " + r.stderr[-200:].replace("\n", " "))
            print(shp, {k[11:]: v for k, v in cfg.items()}, line, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        main()
