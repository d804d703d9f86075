"""Generates the committed golden vectors under tests/golden/ — run in the BUILD container only
(needs /root/reference and torchvision's CPU op):   python tests/golden/make_golden.py

  lsm_*.npz      : outputs of the REAL reference GroundingHead (imported from /root/reference through
                   oracle/ref_loader.py, unmodified source) on seeded inputs.  Small cases store the
                   full inputs; the BASELINE config-2 case stores outputs + input checksums (inputs are
                   regenerated from the seed by oracle.lsm_head.make_lsm_inputs).
  roi_align.npz  : torch.ops.torchvision.roi_align (compiled CPU op — the third-party kernel the
                   reference reaches through Detectron2 ROIPooler) on a small edge-case set.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import lsm_head, ref_loader  # noqa: E402

LSM_CASES = {
    # name: (make_lsm_inputs kwargs, cfg kwargs, store_inputs)
    "small_softmax": (dict(B=4, Rg=10, T=7, V=64, D=64, gain=12.0, seed=11), dict(), True),
    "ragged_empty": (dict(B=5, Rg=12, T=9, V=64, D=64, gain=12.0, seed=12, ragged_regions=True, empty_caption=2, empty_image=2), dict(), True),
    "empty_mismatch": (dict(B=5, Rg=12, T=9, V=64, D=64, gain=12.0, seed=13, empty_caption=1, empty_image=3), dict(), True),
    "small_hardmax": (dict(B=4, Rg=10, T=7, V=64, D=64, gain=12.0, seed=14), dict(alignment="hardmax"), True),
    "small_triplet": (dict(B=6, Rg=10, T=7, V=64, D=64, gain=12.0, seed=15), dict(loss="triplet", negative_mining="hardest"), True),
    "words_only": (dict(B=4, Rg=10, T=7, V=64, D=64, gain=12.0, seed=16), dict(align_regions=False, distillation=False), True),
    "train_shape": (dict(B=4, Rg=100, T=70, V=2048, D=768, seed=17, ragged_regions=True), dict(), False),
    "config2": (dict(B=32, Rg=100, T=20, V=2048, D=768, seed=1992), dict(), False),
    "config2_gain": (dict(B=32, Rg=100, T=20, V=2048, D=768, seed=1992, gain=4.0), dict(), False),
}


def checksum(t):
    return np.array([float(t.double().sum()), float(t.double().abs().sum())])


def run_reference(kw, cfg_kw):
    GH = ref_loader.load_reference_grounding_head()
    ii, ic, w, b = lsm_head.make_lsm_inputs(**kw)
    head = GH(ref_loader.make_grounding_cfg(**cfg_kw), kw["V"], kw["D"])
    with torch.no_grad():
        head.v2l_projection.weight.copy_(w)
        head.v2l_projection.bias.copy_(b)
    with ref_loader.cuda_to_cpu(), torch.no_grad():
        out = head(ii, ic)
    info, losses = out[0], out[1]
    dists = out[2] if len(out) > 2 else {}
    return ii, ic, w, b, info, losses, dists


def main():
    torch.set_num_threads(8)
    for name, (kw, cfg_kw, store) in LSM_CASES.items():
        ii, ic, w, b, info, losses, dists = run_reference(kw, cfg_kw)
        rec = {"kw_keys": np.array(list(kw.keys())), "kw_vals": np.array([str(v) for v in kw.values()]),
               "cfg_keys": np.array(list(cfg_kw.keys())), "cfg_vals": np.array([str(v) for v in cfg_kw.values()])}
        for k, v in losses.items():
            rec["loss::" + k] = np.float32(v.item())
        for k, v in info.items():
            rec["info::" + k] = np.float32(v.item())
        for k, v in dists.items():
            rec["dist::" + k] = v.numpy().astype(np.float32)
        if store:
            rec.update({"region_features": ii["region_features"].numpy(), "region_mask": ii["region_mask"].numpy(),
                        "input_embeddings": ic["input_embeddings"].numpy(), "attention_mask": ic["attention_mask"].numpy(),
                        "special_tokens_mask": ic["special_tokens_mask"].numpy(), "weight": w.numpy(), "bias": b.numpy()})
        else:
            rec["checksum_region_features"] = checksum(ii["region_features"])
            rec["checksum_input_embeddings"] = checksum(ic["input_embeddings"])
            rec["checksum_weight"] = checksum(w)
        np.savez_compressed(os.path.join(HERE, f"lsm_{name}.npz"), **rec)
        print("lsm", name, {k: float(v) for k, v in losses.items()})

    # ---- RoIAlign: compiled torchvision CPU op ------------------------------------------------------
    import torchvision  # noqa: F401
    g = torch.Generator().manual_seed(7)
    N, C, H, W = 2, 6, 20, 30
    feat = torch.randn(N, C, H, W, generator=g)
    boxes = [
        [0, 10.3, 20.7, 200.1, 150.9], [1, 0, 0, 480, 320], [0, -50, -40, 30, 20], [1, 400, 250, 600, 400],
        [0, 100, 100, 100, 100], [1, 100, 100, 101, 160], [0, 470, 300, 480, 320], [1, 900, 900, 950, 950],
        [0, 5.5, 5.5, 21.5, 21.5], [1, 17.2, 33.3, 460.8, 300.1], [0, 240, 10, 250, 310], [1, 300, 200, 280, 180],
    ]
    rnd = torch.rand(20, 4, generator=g)
    for k in range(20):
        cx, cy = rnd[k, 0] * 480, rnd[k, 1] * 320
        s = 16 * (600 / 16) ** float(rnd[k, 2])
        a = 0.5 * 4 ** float(rnd[k, 3])
        bw, bh = s * a ** 0.5, s / a ** 0.5
        boxes.append([k % 2, float(cx - bw / 2), float(cy - bh / 2), float(cx + bw / 2), float(cy + bh / 2)])
    rois = torch.tensor(boxes, dtype=torch.float32)
    rec = {"feat": feat.numpy(), "rois": rois.numpy()}
    for tag, (ps, scale, sr, al) in {"p7_s16_sr0_al1": (7, 1 / 16, 0, True), "p14_s16_sr0_al1": (14, 1 / 16, 0, True),
                                      "p7_s16_sr2_al1": (7, 1 / 16, 2, True), "p5_s16_sr0_al0": (5, 1 / 16, 0, False)}.items():
        out = torch.ops.torchvision.roi_align(feat, rois, scale, ps, ps, sr, al)
        rec["out_" + tag] = out.numpy()
    np.savez_compressed(os.path.join(HERE, "roi_align.npz"), **rec)
    print("roi_align golden written", {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
