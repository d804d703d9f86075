"""Developer probe (not part of the product or the test-suite): runs each kernel family once on the GPU
in its own subprocess (a trapped kernel kills only that stage) and prints error norms vs torch."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["roi", "split", "linear_small", "linear", "linear_acc", "score", "score_big", "lsm", "lsm_acc", "pair_ce"]


def relerr(a, b):
    import torch
    a = a.double(); b = b.double()
    scale = torch.maximum(b.abs(), b.pow(2).mean().sqrt())
    return ((a - b).abs() / scale).max().item()


def run(stage):
    import torch
    import numpy as np
    from locov_b200 import ops
    from oracle import box_head, lsm_head
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    if stage == "roi":
        from oracle import roi_align as ora
        feat = torch.randn(2, 64, 50, 76)
        g = torch.Generator().manual_seed(1)
        R = 64
        cx = torch.rand(R, generator=g) * 1216; cy = torch.rand(R, generator=g) * 800
        s = torch.exp(torch.rand(R, generator=g) * np.log(600 / 16)) * 16
        boxes = torch.stack([cx - s / 2, cy - s / 2, cx + s / 2, cy + s / 2], 1)
        boxes[:4] += 900
        rois = torch.cat([torch.randint(0, 2, (R, 1), generator=g).float(), boxes], 1)
        out = ops.roi_align(feat.to(dev), rois.to(dev), 14, 1 / 16, 0, True).cpu()
        ref = torch.from_numpy(ora.roi_align_fwd(feat.numpy(), rois.numpy(), 14, 1 / 16, 0, True))
        print("roi fwd relerr", relerr(out, ref), "absmax", (out - ref).abs().max().item())
        out2 = ops.roi_align(feat.to(dev).contiguous(memory_format=torch.channels_last), rois.to(dev), 14, 1 / 16, 0, True).cpu()
        print("roi fwd nhwc relerr", relerr(out2, ref))
        ghw, yx, idx = ops.roi_align_grid(rois.to(dev), 50, 76, 14, 1 / 16, 0, True, max_grid=4)
        ghw, yx, idx = ghw.cpu().numpy(), yx.cpu().numpy(), idx.cpu().numpy()
        bad = 0
        for r in range(R):
            g2, yx2, idx2 = ora.roi_align_grid(rois[r].numpy(), 50, 76, 14, 1 / 16, 0, True)
            assert tuple(g2) == tuple(ghw[r]), (g2, ghw[r])
            gh, gw = g2
            if gh > 4 or gw > 4 or gh <= 0 or gw <= 0:
                continue
            a = yx[r][:, :, :gh, :gw].reshape(-1, 2); b = idx[r][:, :, :gh, :gw].reshape(-1, 4)
            bad += int((a.view(np.uint32) != yx2.view(np.uint32)).sum()) + int((b != idx2).sum())
        print("roi grid mismatches", bad)
    elif stage == "split":
        x = torch.randn(37, 100, device=dev)
        o = ops.split_bf16(x, True)
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        print("split hi ok", torch.equal(o.hi[:, :100], hi), "lo ok", torch.equal(o.lo[:, :100], lo), "pad zero", float(o.hi[:, 100:].abs().max()))
        ot = ops.split_bf16(x, False, transpose=True)
        print("split T ok", torch.equal(ot.hi[:, :37], hi.t()), ot.hi.shape)
    elif stage in ("linear_small", "linear", "linear_acc"):
        M, N, K = (128, 64, 64) if stage == "linear_small" else (1000, 772, 2048)
        acc = stage == "linear_acc"
        x = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) * 0.01
        b = torch.randn(N, device=dev) * 0.01
        a_op = ops.split_bf16(x, acc)
        w_op = ops.split_bf16(w, acc)
        out, ob = ops.linear_fwd(a_op, w_op, b, want_f32=True, n_bf16=min(N, 768), accurate_out=acc)
        torch.cuda.synchronize()
        ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
        if not acc:
            ref_b = torch.nn.functional.linear(x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double(), b.double())
            print(stage, "relerr vs bf16-rounded inputs", relerr(out, ref_b))
        print(stage, "relerr vs fp64", relerr(out, ref))
        nb = min(N, 768)
        rec = ob.hi[:, :nb].float() + (ob.lo[:, :nb].float() if ob.lo is not None else 0)
        print(stage, "bf16 out relerr", relerr(rec, out[:, :nb]))
    elif stage in ("score", "score_big"):
        R, K = (1024, 65) if stage == "score" else (1000, 1203)
        x, we, be, wb, bb, cls, gt = box_head.make_box_inputs(R, K)
        e = torch.nn.functional.linear(x, we, be)
        for acc in (False, True):
            e_op = ops.split_bf16(e.to(dev), acc)
            c_op = ops.split_bf16(cls.to(dev), acc)
            logits, probs, lse, arg = ops.box_score(e_op, c_op, None, True)
            torch.cuda.synchronize()
            ref = torch.nn.functional.linear(e.double(), cls.double())
            rp = torch.softmax(ref, -1)
            print(stage, "acc" if acc else "bf16", "logits", relerr(logits.cpu(), ref), "probs", relerr(probs.cpu(), rp),
                  "lse", relerr(lse.cpu(), torch.logsumexp(ref, -1)), "argmax agree",
                  float((arg.cpu() == rp[:, :-1].argmax(1)).float().mean()))
            loss, dl, dlb = ops.box_ce(logits, lse, gt.to(dev), True, True)
            rl = torch.nn.functional.cross_entropy(ref, gt)
            print(stage, "ce", float(loss), float(rl))
    elif stage in ("lsm", "lsm_acc"):
        acc = stage == "lsm_acc"
        for (B, Rg, T, kw) in [(4, 10, 7, {}), (32, 100, 20, {}), (9, 100, 70, dict(ragged_regions=True)),
                               (5, 130, 9, dict(empty_caption=2, empty_image=3))]:
            ii, ic, w, b = lsm_head.make_lsm_inputs(B, Rg, T, V=256, D=768, **kw)
            emb = lsm_head.project_regions(ii["region_features"], w * 5, b)
            cap = ic["input_embeddings"]
            mc = lsm_head.caption_mask_of(ic["attention_mask"], ic["special_tokens_mask"])
            d1, d2 = lsm_head.pair_distances(cap, mc, emb, ii["region_mask"], 10.0, "softmax", torch.float64)
            cap_op = ops.split_bf16(cap.reshape(B * T, -1).to(dev), acc)
            emb_op = ops.split_bf16(emb.reshape(B * Rg, -1).to(dev), acc)
            w2r, r2w = ops.lsm_pair(cap_op, mc.to(dev), emb_op, ii["region_mask"].to(dev), 0.1)
            torch.cuda.synchronize()
            ok = (mc.sum(1)[:, None] > 0) | (ii["region_mask"].sum(1)[None, :] > 0)
            print(stage, (B, Rg, T), "w2r", relerr(w2r.cpu()[ok], d1[ok]), "r2w", relerr(r2w.cpu()[ok], d2[ok]))
    elif stage == "pair_ce":
        B = 32
        pw = torch.randn(B, B)
        mc = torch.ones(B, 5); mr = torch.ones(B, 7)
        mc[3] = 0; mr[3] = 0
        ref = torch.where(((mc.sum(1)[:, None] > 0) | (mr.sum(1)[None, :] > 0)), pw, pw.max() + 100)
        exp = lsm_head.pair_losses(ref)
        pwd = pw.to(dev)
        out = ops.pair_ce(pwd, mc.to(dev), mr.to(dev))
        print("pair_ce", out.cpu().tolist(), [float(v) for v in exp], "guard ok", torch.allclose(pwd.cpu(), ref))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] != "all":
        run(sys.argv[1])
    else:
        for s in STAGES:
            print(f"=== {s}", flush=True)
            try:
                r = subprocess.run([sys.executable, __file__, s], capture_output=True, text=True, timeout=300)
                print(r.stdout[-3000:], r.stderr[-2500:] if r.returncode else "", "rc", r.returncode, flush=True)
            except subprocess.TimeoutExpired:
                print("TIMEOUT", flush=True)
