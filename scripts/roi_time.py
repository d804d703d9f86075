"""Times RoIAlign forward at BASELINE configs[1] (faithful: [2,1024,50,76] x 1024 RoIs -> 14x14) in the three output modes.
Env knobs are read by the library once per process: LOCOV_B200_ROI_PF, LOCOV_B200_ROI_SLABS."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locov_b200 import ops, synthetic  # noqa: E402
from bench import graph_time  # noqa: E402

dev = torch.device("cuda:0")
feats = [synthetic.res4_features(2, C=1024, stride=16, seed=i).to(dev) for i in range(2)]
rois = synthetic.coco_boxes(2, 512, seed=0).to(dev)
res = {}
for tag, kw in (("nchw", {}), ("cl_fp32", {"channels_last": True}), ("cl_bf16", {"channels_last": True, "out_dtype": torch.bfloat16})):
    res[tag] = round(graph_time(torch, lambda i: ops.roi_align(feats[i % 2], rois, 14, 1 / 16, **kw), iters=8) * 1e3, 1)
print("PF", os.environ.get("LOCOV_B200_ROI_PF", "-"), "SLABS", os.environ.get("LOCOV_B200_ROI_SLABS", "-"), res, flush=True)
