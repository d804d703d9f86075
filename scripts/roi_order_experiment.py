import os, sys, torch
sys.path.insert(0, os.getcwd())
from locov_b200 import ops, synthetic
from bench import graph_time
dev = torch.device("cuda:0")
feats = [synthetic.res4_features(2, C=1024, stride=16, seed=i).to(dev) for i in range(2)]
rois = synthetic.coco_boxes(2, 512, seed=0).to(dev)
w = rois[:, 3] - rois[:, 1]; h = rois[:, 4] - rois[:, 2]
area = w * h
order_desc = torch.argsort(area, descending=True)
order_asc = torch.argsort(area)
print("w max", float(w.max()), "h max", float(h.max()), "frac w>672", float((w > 672).float().mean()), "frac h>672", float((h > 672).float().mean()), "frac max(w,h)>448", float((torch.maximum(w, h) > 448).float().mean()))
for tag, r in (("as given", rois), ("largest first", rois[order_desc]), ("smallest first", rois[order_asc]), ("only w,h<=448 (tiled)", rois[torch.maximum(w, h) <= 448].repeat(2, 1)[:1024]),
               ("only max(w,h)>448 (tiled)", rois[torch.maximum(w, h) > 448].repeat(20, 1)[:1024])):
    r = r.contiguous()
    res = {}
    for name, kw in (("nchw", {}), ("cl_fp32", {"channels_last": True})):
        res[name] = round(graph_time(torch, lambda i: ops.roi_align(feats[i % 2], r, 14, 1 / 16, **kw), iters=8) * 1e3, 1)
    print(f"{tag:28s} R={r.shape[0]}", res, flush=True)
