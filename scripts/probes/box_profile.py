"""Developer probe: torch profiler view (op -> kernels) of one box-predictor forward (cfg5) to find glue kernels."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import locov_b200.modeling as M
from oracle import box_head
dev = torch.device("cuda:0")
r, k = 8000, 1203
x, we, be, wb, bb, cls, gt = box_head.make_box_inputs(r, k, seed=3)
cfg = M.get_cfg("stt"); cfg.MODEL.B200.PRECISION = sys.argv[1] if len(sys.argv) > 1 else "bf16"
bp = M.build_box_predictor(cfg, 2048).to(dev)
with torch.no_grad():
    bp.emb_pred.weight.copy_(we); bp.bbox_pred.weight.copy_(wb)
bp.set_class_embeddings(cls)
xd = x.to(dev)
bp.eval()
for _ in range(3):
    with torch.no_grad():
        bp(xd)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    with torch.no_grad():
        bp(xd)
    torch.cuda.synchronize()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA or "copy" in e.name or "contiguous" in e.name or "cat" in e.name:
        print(f"{e.name[:80]:80s} cpu={e.cpu_time_total:8.1f} dev={e.device_time_total:8.1f} shapes={e.input_shapes}")
