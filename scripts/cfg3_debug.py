"""Developer probe: CUDA-graph capture of the config-3 box-head training step (bf16 mode) with the full error."""
import os, sys, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import locov_b200.modeling as M
from locov_b200 import synthetic
dev = torch.device("cuda:0")
R, K = 8192, 48
x, we, be, wb, bb, cls, gt = synthetic.box_inputs(R, K, seed=1)
boxes = synthetic.coco_boxes(16, 512, seed=2)[:, 1:]
boxes[:, 2:] = torch.maximum(boxes[:, 2:], boxes[:, :2] + 8)
props = [M.Instances((800, 1216), proposal_boxes=M.Boxes(boxes[i * 512:(i + 1) * 512].to(dev)), gt_boxes=M.Boxes(boxes[i * 512:(i + 1) * 512].to(dev) + 1.0),
                     gt_classes=gt[i * 512:(i + 1) * 512].to(dev)) for i in range(16)]
for precision in ("bf16", "fp32"):
    cfg = M.get_cfg("stt"); cfg.MODEL.B200.PRECISION = precision
    bp = M.build_box_predictor(cfg, 2048).to(dev).train()
    with torch.no_grad():
        bp.emb_pred.weight.copy_(we); bp.bbox_pred.weight.copy_(wb)
    bp.set_class_embeddings(cls)
    xs = x.to(dev).requires_grad_(True)

    def step():
        pred = bp(xs)
        l = bp.losses(pred, props)
        (l["loss_cls"] + l["loss_box_reg"]).backward()
        xs.grad = None; bp.bbox_pred.weight.grad = None; bp.bbox_pred.bias.grad = None
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            step()
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): g.replay()
        b.record(); torch.cuda.synchronize()
        print(precision, "captured: %.1f us per step" % (a.elapsed_time(b) * 100))
    except Exception:
        traceback.print_exc()
        print(precision, "capture FAILED")
