python -m pytest tests/test_gpu_roi_align.py tests/test_gpu_roi_heads.py tests/test_gpu_env_paths.py -m gpu -q 2>&1 | tail -2
for o in 1 0; do for sl in 0 1 2 4; do
  echo -n "ORDER=$o SLABS=$sl  "; LOCOV_B200_ROI_ORDER=$o LOCOV_B200_ROI_SLABS=$sl python scripts/roi_time.py 2>&1 | tail -1
done; done
python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from locov_b200 import ops, synthetic
from bench import graph_time
dev = torch.device("cuda:0")
feat = synthetic.res4_features(2, C=1024, stride=16, seed=1).to(dev)
rois = synthetic.coco_boxes(2, 512, seed=0).to(dev)
dout = torch.randn(rois.shape[0], 1024, 14, 14, device=dev)
print("bwd ms", round(graph_time(torch, lambda i: ops.roi_align_backward(dout, feat.shape, rois, 1 / 16), iters=4), 4))
PY
LOCOV_B200_ROI_ORDER=0 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from locov_b200 import ops, synthetic
from bench import graph_time
dev = torch.device("cuda:0")
feat = synthetic.res4_features(2, C=1024, stride=16, seed=1).to(dev)
rois = synthetic.coco_boxes(2, 512, seed=0).to(dev)
dout = torch.randn(rois.shape[0], 1024, 14, 14, device=dev)
print("bwd ms (no order)", round(graph_time(torch, lambda i: ops.roi_align_backward(dout, feat.shape, rois, 1 / 16), iters=4), 4))
PY
