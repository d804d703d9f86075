"""Developer probe: pair-CE kernel alone (for ncu --sampling-interval 0)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locov_b200 import ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pw = torch.randn(2, B, B, device=dev)
mc = torch.ones(B, 20, device=dev); mr = torch.ones(B, 100, device=dev)
for _ in range(5):
    ops.pair_ce(pw, mc, mr)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200):
    ops.pair_ce(pw, mc, mr)
b.record(); torch.cuda.synchronize()
print("pair_ce B=%d: %.2f us per call (includes host launch overhead)" % (B, a.elapsed_time(b) / 200 * 1e3))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        ops.pair_ce(pw, mc, mr)
g.replay(); torch.cuda.synchronize()
a.record(); g.replay(); b.record(); torch.cuda.synchronize()
print("pair_ce B=%d: %.2f us per call inside a CUDA graph" % (B, a.elapsed_time(b) / 20 * 1e3))
