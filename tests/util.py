"""Shared helpers of the test-suite."""
import ast
import glob
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def relerr(a, b):
    """max |a-b| / max(|b|, rms(b)) — the robust relative error of SURVEY.md §7 (scores cross zero and
    the background logit is exactly 0, so a plain element-wise relative error is undefined)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    if b.numel() == 0:
        return 0.0
    scale = torch.maximum(b.abs(), b.pow(2).mean().sqrt().clamp_min(1e-30))
    return float(((a - b).abs() / scale).max())


def golden_cases(prefix="lsm_"):
    return sorted(os.path.basename(p)[len(prefix):-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def _lit(s):
    try:
        return ast.literal_eval(s)
    except Exception:
        return s


def golden_lsm(name):
    """-> (input_image, input_caption, weight, bias, cfg_kw, expected dict) for a tests/golden/lsm_*.npz case."""
    from oracle import lsm_head
    z = load_golden("lsm_" + name)
    kw = {k: _lit(v) for k, v in zip(z["kw_keys"].tolist(), z["kw_vals"].tolist())}
    cfg_kw = {k: _lit(v) for k, v in zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist())}
    if "region_features" in z.files:
        ii = {"region_features": torch.from_numpy(z["region_features"]), "region_mask": torch.from_numpy(z["region_mask"])}
        ic = {"input_embeddings": torch.from_numpy(z["input_embeddings"]), "attention_mask": torch.from_numpy(z["attention_mask"]),
              "special_tokens_mask": torch.from_numpy(z["special_tokens_mask"])}
        w, b = torch.from_numpy(z["weight"]), torch.from_numpy(z["bias"])
    else:
        ii, ic, w, b = lsm_head.make_lsm_inputs(**kw)
        for key, t in (("region_features", ii["region_features"]), ("input_embeddings", ic["input_embeddings"]), ("weight", w)):
            cs = np.array([float(t.double().sum()), float(t.double().abs().sum())])
            assert np.allclose(cs, z["checksum_" + key], rtol=1e-9), f"seeded input {key} drifted from the golden run"
    exp = {k: z[k] for k in z.files if k.startswith(("loss::", "info::", "dist::"))}
    return ii, ic, w, b, cfg_kw, exp


def frob_relerr(a, b):
    """||a-b||_F / ||b||_F — used for bf16 GRADIENTS, where a max-norm is dominated by the legitimate sign
    flips of the L1 box loss at its kink (smooth_l1 beta = 0) and by argmax ties."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def golden_box(name):
    """-> (case dict, seeded inputs, golden arrays) for a tests/golden/box_*.npz case; checks the regenerated inputs against the
    checksums stored when the real reference class ran on them."""
    from oracle import box_cases
    z = load_golden("box_" + name)
    c = box_cases.BOX_CASES[name]
    d = box_cases.box_case_inputs(c)
    for key, t in (("x", d["x"]), ("w_emb", d["w_emb"]), ("cls", d["cls"])):
        cs = np.array([float(t.double().sum()), float(t.double().abs().sum())])
        assert np.allclose(cs, z["checksum_" + key], rtol=1e-9), f"seeded input {key} drifted from the golden run"
    return c, d, z
