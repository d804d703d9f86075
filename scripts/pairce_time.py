import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
print("PCE_STAGED_MAX", os.environ.get("LOCOV_B200_PCE_STAGED_MAX", "-"))
from locov_b200 import ops
dev = torch.device("cuda:0")
for B in (32, 48, 64, 128, 256):
    pw = torch.randn(2, B, B, device=dev) * 3
    mc = torch.ones(B, 20, device=dev); mr = torch.ones(B, 1, device=dev)
    for _ in range(3): ops.pair_ce(pw.clone(), mc, mr)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    bufs = [pw.clone() for _ in range(8)]
    with torch.cuda.graph(g):
        for i in range(8): ops.pair_ce(bufs[i], mc, mr)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    print("pair_ce B=%d: %.1f us" % (B, a.elapsed_time(b) / 8 * 1e3))
