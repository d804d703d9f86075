"""Generates tests/golden/box_*.npz and tests/golden/roiheads_*.npz — run in the BUILD container only
(needs /root/reference):   python tests/golden/make_golden_box.py [case names ...]

Every vector is the output of the reference's OWN source, imported unmodified through oracle/ref_loader.py:
  box_*.npz      : ``build_box_predictor(cfg, input_shape)`` -> ``EmbeddingFastRCNNOutputLayers`` (box_emb_head.py:60-249):
                   forward / forward_cls_prediction / set_class_embeddings / from_config are the reference's code; the
                   inherited ``losses`` / ``inference`` are the restated Detectron2 base class (oracle/d2_stubs.py).
  roiheads_*.npz : ``EmbeddingRes5ROIHeads(cfg, input_shape)`` and ``EmbeddingProposalsRes5ROIHeads(cfg, input_shape)``
                   (roi_emb_heads.py:121-360) end to end: ROIPooler (torchvision roi_align) -> res5 -> mean -> predictor
                   -> losses / inference, at reduced channel counts, with the module's full state_dict stored so that the
                   drop-in must load it with identical keys.
Inputs are regenerated from seeds by oracle.box_head (checksums stored); outputs are stored as float32.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import box_head, d2_stubs, ref_loader  # noqa: E402
from oracle.box_cases import BOX_CASES, IMAGE, ROI_CASES, ROI_OVER, ROI_SHAPE, box_case_inputs, roi_case_proposals  # noqa: E402


def checksum(t):
    return np.array([float(t.double().sum()), float(t.double().abs().sum())])


def f32(t):
    return t.detach().to(torch.float32).numpy()


def build_predictor(mod, V, D, stage, over):
    over = dict(over)
    over["MODEL.ROI_BOX_HEAD.EMB_DIM"] = D
    cfg = ref_loader.make_roi_cfg(stage, **over)
    return mod.build_box_predictor(cfg, V)        # the reference's own (cfg, input_shape) construction path


def load_weights(bp, we, be, wb, bb):
    with torch.no_grad():
        bp.emb_pred.weight.copy_(we); bp.emb_pred.bias.copy_(be)
        bp.bbox_pred.weight.copy_(wb); bp.bbox_pred.bias.copy_(bb)


def box_case(mod, name, c):
    R, K, V, D = c["R"], c["K"], c["V"], c["D"]
    d = box_case_inputs(c)
    x, we, be, wb, bb, cls, props = d["x"], d["w_emb"], d["b_emb"], d["w_box"], d["b_box"], d["cls"], d["props"]
    inst = box_head.instances_from(props, IMAGE, d2_stubs.Instances, d2_stubs.Boxes)
    bp = build_predictor(mod, V, D, c["stage"], c["over"])
    load_weights(bp, we, be, wb, bb)
    bp.set_class_embeddings(cls)
    rec = {"meta_keys": np.array(list(c.keys())), "meta_vals": np.array([repr(v) for v in c.values()]),
           "checksum_x": checksum(x), "checksum_w_emb": checksum(we), "checksum_cls": checksum(cls),
           "cls_weight": f32(bp.cls_score.weight) if K <= 100 else checksum(bp.cls_score.weight)}
    if c["mode"] == "eval":
        bp.eval()
        with torch.no_grad():
            scores, deltas = bp(x)
            probs = torch.cat(bp.predict_probs((scores, deltas), inst), 0)
            results, kept = bp.inference((scores, deltas), inst)
        rows = c.get("rows")
        rec.update({"scores": f32(scores if rows is None else scores[:rows]), "deltas": f32(deltas),
                    "probs": f32(probs if rows is None else probs[:rows]),
                    "lse": f32(torch.logsumexp(scores.double(), 1)), "argmax_fg": probs[:, :-1].argmax(1).numpy(),
                    "top2_margin": f32((lambda t: t[:, 0] - t[:, 1])(scores[:, :-1].double().topk(2, 1).values))})
        for i, (r, k) in enumerate(zip(results, kept)):
            rec[f"inst{i}_boxes"] = f32(r.pred_boxes.tensor)
            rec[f"inst{i}_scores"] = f32(r.scores)
            rec[f"inst{i}_classes"] = r.pred_classes.numpy()
            rec[f"inst{i}_kept"] = k.numpy()
        if "K2" in c:       # set_class_embeddings again with another matrix (trainer.py:191,256-257)
            cls2 = d["cls2"]
            bp.set_class_embeddings(cls2)
            with torch.no_grad():
                s2, d2 = bp(x)
            rec.update({"cls2": f32(cls2), "scores2": f32(s2), "deltas2": f32(d2), "num_classes2": np.int64(bp.num_classes)})
    else:
        bp.train()
        xg = x.clone().requires_grad_(True)
        scores, deltas = bp(xg)
        losses = bp.losses((scores, deltas), inst)
        total = sum(losses.values())
        total.backward()
        rec.update({"scores": f32(scores), "deltas": f32(deltas), "scores_requires_grad": np.bool_(scores.requires_grad),
                    "loss_cls": np.float32(losses["loss_cls"].item()), "loss_box_reg": np.float32(losses["loss_box_reg"].item()),
                    "grad_x": f32(xg.grad)})
        for pname, p in bp.named_parameters():
            rec["requires_grad::" + pname] = np.bool_(p.requires_grad)
            if p.grad is not None:
                rec["grad::" + pname] = f32(p.grad)
    rec["state_keys"] = np.array(sorted(bp.state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, f"box_{name}.npz"), **rec)
    print("box", name, {k: (v.shape if hasattr(v, "shape") and v.ndim else v) for k, v in rec.items() if k.startswith(("loss", "scores", "inst0_sc"))})


# ---- ROI heads composite ------------------------------------------------------------------------------------
def randomize_frozen_bn(module, g):
    for m in module.modules():
        if isinstance(m, d2_stubs.FrozenBatchNorm2d):
            m.weight.copy_(1.0 + 0.2 * torch.randn(m.num_features, generator=g))
            m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))


def roi_case(mod, name, cls_name, stage, mode):
    s = ROI_SHAPE
    over = dict(ROI_OVER)
    over["MODEL.ROI_HEADS.NAME"] = cls_name
    over["MODEL.ROI_HEADS.NUM_CLASSES"] = s["K"]
    cfg = ref_loader.make_roi_cfg(stage, **over)
    torch.manual_seed(s["seed"])
    heads = getattr(mod, cls_name)(cfg, {"res4": d2_stubs.ShapeSpec(channels=s["C"], stride=16)})   # Detectron2's build_roi_heads call
    g = torch.Generator().manual_seed(s["seed"] + 1)
    with torch.no_grad():
        randomize_frozen_bn(heads, g)
        heads.box_predictor.emb_pred.weight.mul_(10.0)
        heads.box_predictor.bbox_pred.weight.mul_(30.0)
    cls = torch.cat([torch.randn(s["K"], 32, generator=g) * 0.5, torch.zeros(1, 32)], 0)
    heads.box_predictor.set_class_embeddings(cls)
    feat = torch.randn(s["N"], s["C"], s["H"], s["W"], generator=g)
    props = roi_case_proposals()
    inst = box_head.instances_from(props, IMAGE, d2_stubs.Instances, d2_stubs.Boxes)
    # the proposal matcher / sampler is not on the path: proposals already carry their labels (as the LSM-stage mapper
    # produces them, coco_mappers.py:88-106), so label_and_sample_proposals is the identity here
    heads.label_and_sample_proposals = lambda proposals, targets: proposals
    rec = {"feat": f32(feat), "cls": f32(cls), "meta": np.array([cls_name, stage, mode])}
    for k, v in heads.state_dict().items():
        rec["state::" + k] = v.numpy()
    if mode == "train":
        heads.train()
        out = heads(None, {"res4": feat}, inst, targets=inst)
        if cls_name == "EmbeddingProposalsRes5ROIHeads":
            grid, box_feats, props_out, losses = out
            rec["grid_features"] = f32(grid)
            for i, bf in enumerate(box_feats):
                rec[f"box_features{i}"] = f32(bf)
            assert props_out is inst
        else:
            empty, losses = out
            assert empty == []
        for k, v in losses.items():
            rec["loss::" + k] = np.float32(v.item())
    else:
        heads.eval()
        with torch.no_grad():
            results, extra = heads(None, {"res4": feat}, inst, targets=None)
        assert extra == {}
        for i, r in enumerate(results):
            rec[f"inst{i}_boxes"] = f32(r.pred_boxes.tensor)
            rec[f"inst{i}_scores"] = f32(r.scores)
            rec[f"inst{i}_classes"] = r.pred_classes.numpy()
    np.savez_compressed(os.path.join(HERE, f"roiheads_{name}.npz"), **rec)
    print("roiheads", name, {k: v for k, v in rec.items() if k.startswith("loss::")}, [len(rec[k]) for k in rec if k.endswith("_scores")])


def main():
    torch.set_num_threads(8)
    only = set(sys.argv[1:])                      # optional: names of the cases to (re)generate
    bmod = ref_loader.load_reference_box_head()
    for name, c in BOX_CASES.items():
        if not only or name in only:
            box_case(bmod, name, c)
    rmod = ref_loader.load_reference_roi_heads()
    for name, (cls_name, stage, mode) in ROI_CASES.items():
        if not only or name in only:
            roi_case(rmod, name, cls_name, stage, mode)


if __name__ == "__main__":
    main()
