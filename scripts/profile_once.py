"""One eager pass of a section of the hot path between cudaProfilerStart / Stop, for `ncu --profile-from-start off --set full`.
usage: python scripts/profile_once.py lsm|box5|box3|roi|mean|distill"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import locov_b200.modeling as M  # noqa: E402
from locov_b200 import functional as LF, ops, synthetic  # noqa: E402
import bench  # noqa: E402

what = sys.argv[1]
dev = torch.device("cuda:0")
torch.cuda.set_device(0)


def profiled(fn, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    fn()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if what == "lsm":
    ii, ic, w, b = bench.make_inputs(bench.SEED)
    ii = {k: v.to(dev) for k, v in ii.items()}
    ic = {k: v.to(dev) for k, v in ic.items()}
    heads = {}
    for precision in ("fp32", "bf16"):
        cfg = M.get_cfg("lsm")
        cfg.MODEL.B200.PRECISION = precision
        h = M.GroundingHead(cfg, bench.V, bench.D).to(dev)
        with torch.no_grad():
            h.v2l_projection.weight.copy_(w); h.v2l_projection.bias.copy_(b)
        heads[precision] = h

    def fn():
        with torch.no_grad():
            for precision in ("fp32", "bf16"):
                LF.clear_weight_cache()
                heads[precision](ii, ic)
    profiled(fn)
elif what in ("box5", "box3"):
    R, K = (8000, 1203) if what == "box5" else (8192, 48)
    x, we, be, wb, bb, cls, gt = synthetic.box_inputs(R, K, seed=bench.SEED + 7)
    x = x.to(dev)
    preds = {}
    for precision in ("fp32", "bf16"):
        cfg = M.get_cfg("stt")
        cfg.MODEL.B200.PRECISION = precision
        bp = M.build_box_predictor(cfg, 2048).to(dev)
        with torch.no_grad():
            bp.emb_pred.weight.copy_(we); bp.bbox_pred.weight.copy_(wb)
        bp.set_class_embeddings(cls)
        preds[precision] = bp.train(what == "box3")
    if what == "box5":
        def fn():
            with torch.no_grad():
                for bp in preds.values():
                    s, d = bp(x)
                    bp.predict_probs((s, d), [range(R)])
    else:
        boxes = synthetic.coco_boxes(16, 512, seed=bench.SEED + 1)[:, 1:]
        boxes[:, 2:] = torch.maximum(boxes[:, 2:], boxes[:, :2] + 8)
        gtb = boxes + 2.0
        props = [M.Instances((800, 1216), proposal_boxes=M.Boxes(boxes[i * 512:(i + 1) * 512].to(dev)), gt_boxes=M.Boxes(gtb[i * 512:(i + 1) * 512].to(dev)),
                             gt_classes=gt[i * 512:(i + 1) * 512].to(dev)) for i in range(16)]
        xg = x.clone().requires_grad_(True)

        def fn():
            for bp in preds.values():
                pred = bp(xg)
                l = bp.losses(pred, props)
                (l["loss_cls"] + l["loss_box_reg"]).backward()
                xg.grad = None
    profiled(fn)
elif what == "roi":
    feat = synthetic.res4_features(2, C=1024, stride=16, seed=1).to(dev)
    rois = synthetic.coco_boxes(2, 512, seed=0).to(dev)
    dout = torch.randn(rois.shape[0], 1024, 14, 14, device=dev)

    def fn():
        ops.roi_align(feat, rois, 14, 1 / 16)
        ops.roi_align(feat, rois, 14, 1 / 16, channels_last=True)
        ops.roi_align(feat, rois, 14, 1 / 16, channels_last=True, out_dtype=torch.bfloat16)
        ops.roi_align_backward(dout, feat.shape, rois, 1 / 16)
    profiled(fn)
elif what == "mean":
    x = torch.randn(8000, 2048, 7, 7, device=dev)
    xb = torch.randn(8000, 2048, 7, 7, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)

    def fn():
        ops.spatial_mean(x, True)
        ops.spatial_mean(xb, True)
    profiled(fn)
elif what == "distill":
    b = 256
    t, w, r = (torch.randn(b, b, device=dev) * 5 for _ in range(3))
    pw = torch.randn(2, b, b, device=dev)
    mc, mr = torch.ones(b, 20, device=dev), torch.ones(b, 100, device=dev)

    def fn():
        ops.pair_distill(t, w, r, 10.0, 0, 1.0, True, True)
        ops.pair_ce(pw, mc, mr)
        ops.tensor_stats(x_stats)
    x_stats = torch.randn(32, 100, 2048, device=dev)
    profiled(fn)
else:
    raise SystemExit("unknown section " + what)
print("profiled", what)
