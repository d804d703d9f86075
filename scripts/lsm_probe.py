"""Developer probe: the LSM pair kernel alone at BASELINE config 2 (for ncu) + graph-timed variants."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locov_b200 import ops
dev = torch.device("cuda:0")
B, T, RG, D = 32, 20, 100, 768
if len(sys.argv) > 1:
    B = int(sys.argv[1])
bi = int(sys.argv[2]) if len(sys.argv) > 2 else B
D = int(sys.argv[3]) if len(sys.argv) > 3 else D
cap = ops.split_bf16(torch.randn(B * T, D, device=dev) * 0.05, False)
emb = ops.split_bf16(torch.randn(bi * RG, D, device=dev) * 0.5, False)
mc = torch.ones(B, T, device=dev); mr = torch.ones(bi, RG, device=dev)
w2r = torch.empty(B, bi, device=dev); r2w = torch.empty(B, bi, device=dev)
for _ in range(3):
    ops.lsm_pair(cap, mc, emb, mr, 0.1, out_w2r=w2r, out_r2w=r2w)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        ops.lsm_pair(cap, mc, emb, mr, 0.1, out_w2r=w2r, out_r2w=r2w)
g.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); g.replay(); b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / 20 * 1e3
print("lsm_pair Bc=%d Bi=%d: %.2f us per launch in a graph, %.1f TFLOP/s" % (B, bi, us, 2.0 * B * T * bi * RG * D / us / 1e6))
