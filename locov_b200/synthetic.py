"""Seeded synthetic workloads of the region-text path (SEED = 1992 in both shipped YAMLs) — what
``bench.py`` and ``scripts/`` feed the kernels: COCO-shaped R50-C4 features and boxes, box-head RoI features with
BERT-shaped class embeddings, LSM region features with caption word embeddings and token masks.  Host tensors only
(torch CPU generators, so every rank / box regenerates identical data); nothing here is arithmetic of the path.
tests/test_synthetic_cpu.py asserts these generators agree value for value with the oracle's own (oracle/lsm_head.py,
oracle/box_head.py), so golden vectors and benchmark inputs come from one definition."""
import torch


def lsm_inputs(B, Rg, T, V=2048, D=768, seed=1992, ragged_regions=False, min_words=6, gain=1.0):
    """-> (input_image dict, input_caption dict, v2l weight [D,V], bias [D]) in the caller contract of
    distill_prop_mmss_gcnn.py:273-399 (zero-padded region features + uint8 region mask; int64 token masks)."""
    g = torch.Generator().manual_seed(seed)
    cap = torch.randn(B, T, D, generator=g) * 0.05 * gain
    att = torch.zeros(B, T, dtype=torch.int64)
    spe = torch.zeros(B, T, dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(min(min_words, T), T + 1, (1,), generator=g))
        att[b, :n] = 1
        spe[b, 0] = 1
        spe[b, n - 1] = 1
        spe[b, n:] = 1
    feats = torch.randn(B, Rg, V, generator=g)
    rmask = torch.ones(B, Rg, dtype=torch.uint8)
    if ragged_regions:
        for b in range(B):
            n = int(torch.randint(max(1, Rg // 3), Rg + 1, (1,), generator=g))
            rmask[b, n:] = 0
            feats[b, n:] = 0
    weight = torch.randn(D, V, generator=g) * 0.01 * gain
    bias = torch.randn(D, generator=g) * 0.01
    return ({"region_features": feats, "region_mask": rmask},
            {"input_embeddings": cap, "attention_mask": att, "special_tokens_mask": spe}, weight, bias)


def box_inputs(R, K, V=2048, D=768, seed=1992, bg_frac=0.25):
    """-> (x [R,V], w_emb [D,V], b_emb, w_box [4,V], b_box, class matrix [K+1,D] with the zero background row, labels [R])."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(R, V, generator=g)
    w_emb = torch.randn(D, V, generator=g) * 0.01
    b_emb = torch.zeros(D)
    w_box = torch.randn(4, V, generator=g) * 0.001
    b_box = torch.zeros(4)
    cls = torch.cat([torch.randn(K, D, generator=g) * 0.05, torch.zeros(1, D)], 0)
    gt = torch.randint(0, K, (R,), generator=g)
    gt[torch.rand(R, generator=g) < bg_frac] = K
    return x, w_emb, b_emb, w_box, b_box, cls, gt


def coco_boxes(n_img, per_img, seed=1992, image_size=(800, 1216)):
    """[n_img * per_img, 5] rows (batch index, x1, y1, x2, y2): centres uniform in the image, scale log-uniform in
    [16, 600] px, aspect log-uniform in [1/2, 2], clipped to the image (1-4 adaptive samples per bin axis at stride 16)."""
    g = torch.Generator().manual_seed(seed)
    h, w = image_size
    rows = []
    for i in range(n_img):
        cx = torch.rand(per_img, generator=g) * w
        cy = torch.rand(per_img, generator=g) * h
        s = 16.0 * (600.0 / 16.0) ** torch.rand(per_img, generator=g)
        a = 0.5 * 4.0 ** torch.rand(per_img, generator=g)
        bw, bh = s * a.sqrt(), s / a.sqrt()
        box = torch.stack([(cx - bw / 2).clamp(0, w), (cy - bh / 2).clamp(0, h), (cx + bw / 2).clamp(0, w), (cy + bh / 2).clamp(0, h)], 1)
        rows.append(torch.cat([torch.full((per_img, 1), float(i)), box], 1))
    return torch.cat(rows, 0)


def res4_features(n_img, C=1024, image_size=(800, 1216), stride=16, seed=1992):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n_img, C, image_size[0] // stride, image_size[1] // stride, generator=g)
