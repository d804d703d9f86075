"""ORACLE (test infrastructure only): the seeded input sets behind tests/golden/box_*.npz and roiheads_*.npz.
Shared by the generator (tests/golden/make_golden_box.py, which runs the REAL reference classes on them) and by the
parity tests (which run the oracle restatement and the CUDA path on the same inputs)."""
import torch

from . import box_head

IMAGE = (320, 480)

BOX_CASES = {
    # R rois over n_img images, K foreground classes, V -> D projection; stage = which shipped YAML the config follows
    "k65_infer": dict(R=256, K=65, V=2048, D=768, n_img=2, seed=101, stage="stt", over={}, mode="eval"),
    "k17_infer_gain": dict(R=200, K=17, V=2048, D=768, n_img=2, seed=102, stage="stt", over={}, mode="eval", cls_gain=4.0),
    "k48_train": dict(R=192, K=48, V=512, D=128, n_img=2, seed=103, stage="stt", over={"MODEL.ROI_BOX_HEAD.FREEZE_EMB_PRED": False}, mode="train"),
    "k48_train_frozen": dict(R=192, K=48, V=512, D=128, n_img=2, seed=104, stage="stt", over={}, mode="train"),
    "k1203": dict(R=200, K=1203, V=2048, D=768, n_img=2, seed=105, stage="stt", over={}, mode="eval", rows=16, cls_gain=8.0),
    "normalize": dict(R=96, K=20, V=256, D=64, n_img=2, seed=106, stage="stt", over={"MODEL.ROI_BOX_HEAD.NORMALIZE_EMB_PRED": True}, mode="eval"),
    "standardize": dict(R=96, K=20, V=256, D=64, n_img=2, seed=107, stage="stt", over={"MODEL.ROI_BOX_HEAD.STANDARDIZE_EMB_PRED": True}, mode="eval"),
    "normalize_train": dict(R=96, K=20, V=256, D=64, n_img=2, seed=110, stage="stt",
                            over={"MODEL.ROI_BOX_HEAD.NORMALIZE_EMB_PRED": True, "MODEL.ROI_BOX_HEAD.FREEZE_EMB_PRED": False}, mode="train", cls_gain=40.0),
    "standardize_train": dict(R=96, K=20, V=256, D=64, n_img=2, seed=111, stage="stt",
                              over={"MODEL.ROI_BOX_HEAD.STANDARDIZE_EMB_PRED": True, "MODEL.ROI_BOX_HEAD.FREEZE_EMB_PRED": False}, mode="train", cls_gain=4.0),
    "detach_train": dict(R=128, K=30, V=256, D=64, n_img=2, seed=108, stage="lsm", over={}, mode="train"),
    # the inference tail at LVIS settings (coco/lvis evaluation: score threshold 1e-4, 300 detections per image): ~10^5 candidates per
    # image, every class has many -> torchvision's per-class ("vanilla") batched NMS in the reference
    "k1203_lvis": dict(R=600, K=1203, V=2048, D=768, n_img=2, seed=112, stage="stt", mode="eval", rows=8, cls_gain=8.0,
                       over={"MODEL.ROI_HEADS.SCORE_THRESH_TEST": 1e-4, "TEST.DETECTIONS_PER_IMAGE": 300}),
    "k65_dense": dict(R=900, K=65, V=256, D=64, n_img=3, seed=113, stage="stt", mode="eval", rows=8, cls_gain=2.0,
                      over={"MODEL.ROI_HEADS.SCORE_THRESH_TEST": 0.001, "MODEL.ROI_HEADS.NMS_THRESH_TEST": 0.3, "TEST.DETECTIONS_PER_IMAGE": 50}),
    "reset": dict(R=64, K=48, V=256, D=64, n_img=2, seed=109, stage="stt", over={}, mode="eval", K2=65),
}


def box_case_inputs(c):
    """-> dict(x, w_emb, b_emb, w_box, b_box, cls [K+1,D], props (list of per-image dicts), cls2 or None)."""
    R, K, V, D = c["R"], c["K"], c["V"], c["D"]
    x, we, be, wb, bb, cls, _ = box_head.make_box_inputs(R, K, V=V, D=D, seed=c["seed"])
    g = torch.Generator().manual_seed(c["seed"] + 7)
    be = torch.randn(D, generator=g) * 0.01
    bb = torch.randn(4, generator=g) * 0.01
    cls = cls * c.get("cls_gain", 1.0)
    props = box_head.make_proposals(c["n_img"], R // c["n_img"], K, seed=c["seed"] + 1, image_size=IMAGE)
    cls2 = None
    if "K2" in c:
        g2 = torch.Generator().manual_seed(c["seed"] + 3)
        cls2 = torch.cat([torch.randn(c["K2"], D, generator=g2) * 0.05, torch.zeros(1, D)], 0)
    return dict(x=x, w_emb=we, b_emb=be, w_box=wb, b_box=bb, cls=cls, props=props, cls2=cls2)


# ---- ROI-heads composite (pool -> res5 -> mean -> predictor -> losses / inference) at reduced channel counts --------
ROI_OVER = {"MODEL.RESNETS.RES2_OUT_CHANNELS": 8, "MODEL.RESNETS.WIDTH_PER_GROUP": 2, "MODEL.ROI_BOX_HEAD.EMB_DIM": 32}
ROI_SHAPE = dict(N=2, C=32, H=20, W=30, per_img=20, K=6, seed=211)
ROI_CASES = {
    "res5_train": ("EmbeddingRes5ROIHeads", "stt", "train"),
    "res5_eval": ("EmbeddingRes5ROIHeads", "stt", "eval"),
    "proposals_train": ("EmbeddingProposalsRes5ROIHeads", "lsm", "train"),
    "proposals_eval": ("EmbeddingProposalsRes5ROIHeads", "lsm", "eval"),
}


def roi_case_proposals():
    s = ROI_SHAPE
    return box_head.make_proposals(s["N"], s["per_img"], s["K"], seed=s["seed"] + 2, image_size=IMAGE)
