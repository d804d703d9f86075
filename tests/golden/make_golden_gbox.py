"""Generates tests/golden/gbox_*.npz — run in the BUILD container only (needs /root/reference):
    python tests/golden/make_golden_gbox.py
Every vector is the output of the reference's OWN ``EmbeddingGroundingFastRCNNOutputLayers`` / ``GroundingModule``
(ovr/modeling/roi_heads/box_emb_grounding_head.py:60-434), imported unmodified through oracle/ref_loader.py and built through its
``(cfg, input_shape)`` path with the one config key the reference's config.py lacks (MODEL.ROI_HEADS.MAX_TOKENS).  Inputs are
regenerated from seeds by oracle.grounding_module (checksums stored)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import box_head, d2_stubs, ref_loader  # noqa: E402
from oracle.box_cases import IMAGE  # noqa: E402
from oracle.grounding_module import GBOX_CASES, gbox_inputs  # noqa: E402


def f32(t):
    return t.detach().to(torch.float32).numpy()


def main():
    torch.set_num_threads(4)
    mod = ref_loader.load_reference_grounding_box_head()
    for name, c in GBOX_CASES.items():
        d = gbox_inputs(c)
        over = {"MODEL.ROI_HEADS.MAX_TOKENS": c["max_tok"], "MODEL.ROI_BOX_HEAD.EMB_DIM": c["D"], "MODEL.ROI_HEADS.NUM_CLASSES": c["K"],
                "MODEL.MMSS_HEAD.GROUNDING.ALIGNMENT": c["alignment"], "MODEL.MMSS_HEAD.GROUNDING.ALIGNMENT_TEMPERATURE": c["temperature"],
                "MODEL.ROI_BOX_HEAD.NORMALIZE_EMB_PRED": bool(c.get("normalize", False))}
        cfg = ref_loader.make_roi_cfg("stt", **over)
        bp = mod.EmbeddingGroundingFastRCNNOutputLayers(cfg, d2_stubs.ShapeSpec(channels=c["V"]))
        with torch.no_grad():
            bp.emb_pred.weight.copy_(d["w_emb"]); bp.emb_pred.bias.copy_(d["b_emb"])
            bp.bbox_pred.weight.copy_(d["w_box"]); bp.bbox_pred.bias.copy_(d["b_box"])
        bp.set_class_embeddings(d["embs"])
        rec = {"checksum_x": np.array([float(d["x"].double().sum())]), "num_classes": np.int64(bp.num_classes),
               "state_keys": np.array(sorted(bp.state_dict().keys()))}
        props = box_head.make_proposals(2, c["R"] // 2, c["K"], seed=c["seed"] + 1, image_size=IMAGE)
        inst = box_head.instances_from(props, IMAGE, d2_stubs.Instances, d2_stubs.Boxes)
        if c["mode"] == "eval":
            bp.eval()
            with torch.no_grad():
                scores, deltas = bp(d["x"])
                e = bp.emb_pred(d["x"])
                if c.get("normalize"):
                    e = mod.normalize_vec(e, dim=1)
                s2, att = bp.cls_score(e)
                results, kept = bp.inference((scores, deltas), inst)
            assert torch.equal(s2, scores)
            rec.update({"scores": f32(scores), "deltas": f32(deltas), "tok_attention": f32(att), "num_tok": bp.cls_score.num_tok.numpy()})
            for i, (r, k) in enumerate(zip(results, kept)):
                rec[f"inst{i}_scores"] = f32(r.scores)
                rec[f"inst{i}_classes"] = r.pred_classes.numpy()
                rec[f"inst{i}_kept"] = k.numpy()
        else:
            bp.train()
            xg = d["x"].clone().requires_grad_(True)
            scores, deltas = bp(xg)
            losses = bp.losses((scores, deltas), inst)
            sum(losses.values()).backward()
            rec.update({"scores": f32(scores), "deltas": f32(deltas), "loss_cls": np.float32(losses["loss_cls"].item()),
                        "loss_box_reg": np.float32(losses["loss_box_reg"].item()), "grad_x": f32(xg.grad)})
            for pname, p in bp.named_parameters():
                rec["requires_grad::" + pname] = np.bool_(p.requires_grad)
                if p.grad is not None:
                    rec["grad::" + pname] = f32(p.grad)
        np.savez_compressed(os.path.join(HERE, f"gbox_{name}.npz"), **rec)
        print("gbox", name, rec["scores"].shape, {k: v for k, v in rec.items() if k.startswith("loss")})


if __name__ == "__main__":
    main()
