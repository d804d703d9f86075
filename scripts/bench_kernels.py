"""Per-kernel roofline table for the parts of the path that bench.py's headline (LSM head) does not cover:
RoIAlign (HBM bound) and the box-predictor GEMM chain (tensor bound), at the BASELINE config sizes, with the
library kernels the reference lands on (torchvision roi_align CUDA, cuBLAS F.linear + ATen softmax) timed beside
them.  Prints one JSON object per line.  Run on the GPU box:  python scripts/bench_kernels.py [--quick]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import locov_b200.modeling as M  # noqa: E402
from locov_b200 import ops  # noqa: E402

PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, flush_l2=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush_l2:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def timeit_graph(fn, iters=20):
    """Device time of one call with the host launch overhead removed: the call is captured in a CUDA graph and replayed
    (these chains are a dozen short kernels; eager timing measures the Python / launch path, not the GPU)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def rois_for(n_img, per_img, h_img=800, w_img=1216, seed=1992):
    g = torch.Generator().manual_seed(seed)
    r = n_img * per_img
    cx, cy = torch.rand(r, generator=g) * w_img, torch.rand(r, generator=g) * h_img
    s = 16 * (float(os.environ.get('RA_SMAX', '600')) / 16) ** torch.rand(r, generator=g)
    a = 0.5 * 4 ** torch.rand(r, generator=g)
    bw, bh = s * a.sqrt(), s / a.sqrt()
    b = torch.stack([(cx - bw / 2).clamp(0, w_img), (cy - bh / 2).clamp(0, h_img), (cx + bw / 2).clamp(0, w_img), (cy + bh / 2).clamp(0, h_img)], 1)
    idx = torch.arange(n_img).repeat_interleave(per_img).float()[:, None]
    return torch.cat([idx, b], 1).to(dev)


def roi_case(name, n, c, h, w, per_img, ps, scale):
    import torchvision
    feat = torch.randn(n, c, h, w, device=dev)
    feat_cl = feat.contiguous(memory_format=torch.channels_last)
    rois = rois_for(n, per_img)
    r = rois.shape[0]
    nbytes = r * c * ps * ps * 4 + feat.numel() * 4 + rois.numel() * 4
    for tag, f in (("nchw-in", lambda: ops.roi_align(feat, rois, ps, scale)), ("nhwc-in", lambda: ops.roi_align(feat_cl, rois, ps, scale)),
                   ("torchvision-cuda", lambda: torchvision.ops.roi_align(feat, rois, ps, scale, 0, True))):
        ms = timeit(f, iters=10)
        print(json.dumps({"kernel": "roi_align", "case": name, "impl": tag, "R": r, "C": c, "out": ps, "ms": ms, "algorithmic_MB": nbytes / 1e6,
                          "GB/s": nbytes / ms / 1e6, "frac_of_hbm": nbytes / ms / 1e6 / PK["hbm_gbs"]}), flush=True)


def roi_bwd_case(name, n, c, h, w, per_img, ps, scale):
    """RoIAlign backward (gradient w.r.t. the feature map): ours vs torchvision's CUDA backward, same upstream gradient."""
    import torchvision
    feat = torch.randn(n, c, h, w, device=dev)
    rois = rois_for(n, per_img)
    r = rois.shape[0]
    dout = torch.randn(r, c, ps, ps, device=dev)
    nbytes = dout.numel() * 4 + feat.numel() * 4 * 2          # read dout once, read-modify-write the gradient map

    def tv():
        torch.ops.torchvision._roi_align_backward(dout, rois, scale, ps, ps, n, c, h, w, 0, True)

    for tag, f in (("b200", lambda: ops.roi_align_backward(dout, feat.shape, rois, scale)), ("torchvision-cuda", tv)):
        ms = timeit(f, iters=5)
        print(json.dumps({"kernel": "roi_align_bwd", "case": name, "impl": tag, "R": r, "C": c, "out": ps, "ms": ms, "algorithmic_MB": nbytes / 1e6,
                          "GB/s": nbytes / ms / 1e6, "frac_of_hbm": nbytes / ms / 1e6 / PK["hbm_gbs"]}), flush=True)


def box_case(name, r, k, precision, bwd=False):
    from oracle import box_head
    x, we, be, wb, bb, cls, gt = box_head.make_box_inputs(r, k, seed=3)
    cfg = M.get_cfg("stt")
    cfg.MODEL.B200.PRECISION = precision
    cfg.MODEL.ROI_BOX_HEAD.FREEZE_EMB_PRED = True
    bp = M.build_box_predictor(cfg, 2048).to(dev)
    with torch.no_grad():
        bp.emb_pred.weight.copy_(we); bp.bbox_pred.weight.copy_(wb)
    bp.set_class_embeddings(cls)
    xd = x.to(dev).requires_grad_(bwd)
    gtd = gt.to(dev)
    props = [M.Instances((800, 1216), proposal_boxes=M.Boxes(torch.rand(r, 4, device=dev) * 100 + torch.tensor([0, 0, 200, 200], device=dev)),
                         gt_boxes=M.Boxes(torch.rand(r, 4, device=dev) * 100 + torch.tensor([0, 0, 200, 200], device=dev)), gt_classes=gtd)]
    flops = 2.0 * r * 2048 * 772 + 2.0 * r * 768 * (k + 1)
    if bwd:
        flops += 2.0 * r * (k + 1) * 768 + 2.0 * r * 772 * 2048          # dE and dX (weights frozen, coco_stt.yaml:36)

    def ours():
        bp.train(bwd)
        with torch.set_grad_enabled(bwd):
            pred = bp(xd)
            if bwd:
                l = bp.losses(pred, props)
                (l["loss_cls"] + l["loss_box_reg"]).backward()
                xd.grad = None

    lin_e = torch.nn.Linear(2048, 768).to(dev); lin_b = torch.nn.Linear(2048, 4).to(dev); lin_c = torch.nn.Linear(768, k + 1).to(dev)
    with torch.no_grad():
        lin_c.weight.copy_(cls); lin_c.bias.zero_()
    for p in list(lin_e.parameters()) + list(lin_c.parameters()):
        p.requires_grad_(False)

    def lib(dtype):
        # the reference's arithmetic on library kernels: three F.linear (cuBLAS) + the SAME Detectron2-style loss code
        # (F.cross_entropy + smooth-L1 box regression through bp.losses) as our arm, so that the two differ only in the
        # GEMM / softmax / cross-entropy kernels
        def f():
            with torch.set_grad_enabled(bwd):
                with torch.autocast("cuda", dtype=dtype, enabled=dtype != torch.float32):
                    d = lin_b(xd)
                    s = lin_c(lin_e(xd))
                if bwd:
                    l = bp.losses((s.float(), d.float()), props)
                    (l["loss_cls"] + l["loss_box_reg"]).backward()
                    xd.grad = None
                else:
                    torch.softmax(s.float(), -1)
        return f

    rows = [(f"b200-{precision}", ours)]
    rows.append(("cublas-bf16-autocast" if precision == "bf16" else "cublas-fp32", lib(torch.bfloat16 if precision == "bf16" else torch.float32)))
    for tag, f in rows:
        for mode in (("eager",) if bwd else ("eager", "graph")):     # the training step (losses + autograd) syncs: not capturable
            try:
                ms = timeit(f, iters=10, flush_l2=False) if mode == "eager" else timeit_graph(f)
            except Exception as e:   # noqa: BLE001  (a chain that cannot be captured is reported, not hidden)
                print(json.dumps({"kernel": "box_predictor", "case": name, "impl": tag, "timing": mode, "error": str(e)[:200]}), flush=True)
                continue
            print(json.dumps({"kernel": "box_predictor" + ("_fwd_bwd" if bwd else "_fwd"), "case": name, "impl": tag, "timing": mode, "R": r, "K1": k + 1,
                              "ms": ms, "algorithmic_GFLOP": flops / 1e9, "TFLOP/s": flops / ms / 1e9,
                              "frac_of_bf16_peak": flops / ms / 1e9 / PK["bf16_tflops"], "scores/s": r * (k + 1) / ms * 1e3}), flush=True)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    if "--box-only" not in sys.argv:
      roi_case("cfg1 faithful [2,1024,50,76]->[1024,1024,14,14]", 2, 1024, 50, 76, 512, 14, 1 / 16)
      roi_case("cfg1 literal [2,2048,25,38]->[1024,2048,7,7]", 2, 2048, 25, 38, 512, 7, 1 / 32)
    if "--roi-bwd" in sys.argv:
        roi_bwd_case("cfg1 faithful bwd [1024,1024,14,14]->[2,1024,50,76]", 2, 1024, 50, 76, 512, 14, 1 / 16)
        sys.exit(0)
    if "--box-only" not in sys.argv:
        roi_bwd_case("cfg1 faithful bwd [1024,1024,14,14]->[2,1024,50,76]", 2, 1024, 50, 76, 512, 14, 1 / 16)
    if "--roi-only" in sys.argv:
        sys.exit(0)
    if "--box-only" not in sys.argv:
        pass
    if not quick and "--box-only" not in sys.argv:
        roi_case("cfg3 [16,1024,50,76]->[8192,1024,14,14]", 16, 1024, 50, 76, 512, 14, 1 / 16)
    only = os.environ.get("BK_ONLY", "")
    cases = [("cfg1 2x512 RoIs vs 65+1", 1024, 65, "fp32", False), ("cfg1 2x512 RoIs vs 65+1", 1024, 65, "bf16", False),
             ("cfg3 16x512 RoIs vs 48+1", 8192, 48, "bf16", True), ("cfg5 8x1000 RoIs vs 1203+1", 8000, 1203, "bf16", False),
             ("cfg5 8x1000 RoIs vs 1203+1", 8000, 1203, "fp32", False)]
    for name, r, k, prec, bwd in cases:
        if only and only not in f"{name.split()[0]}-{prec}":
            continue
        box_case(name, r, k, prec, bwd=bwd)
