"""Exchange logic of the batch-sharded LSM pair matrix (locov_b200/parallel.py) on CPU, world_size 2,
gloo backend: block gather / assembly, the global empty-pair guard and the four global losses must equal
the single-device oracle result, and the backward of the block gather must hand each rank its own
column slice.  (The CUDA kernels that produce the blocks are covered by the -m gpu tests; here the
blocks come from the oracle so that only the host-side N>1 logic is under test.)"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import locov_b200.modeling as M
        from locov_b200 import parallel
        from oracle import lsm_head
        B, Rg, T, V, D = 6, 9, 7, 32, 24
        ii, ic, w, b = lsm_head.make_lsm_inputs(B=B, Rg=Rg, T=T, V=V, D=D, seed=21, gain=10.0, ragged_regions=True,
                                                empty_caption=1, empty_image=1)
        bl = B // world
        sl = slice(rank * bl, (rank + 1) * bl)
        cap_mask = lsm_head.caption_mask_of(ic["attention_mask"], ic["special_tokens_mask"])
        # --- what the CUDA path does, with the oracle standing in for the kernels -------------------
        cap_all, mask_all = parallel.gather_packed([ic["input_embeddings"][sl].contiguous(), cap_mask[sl].contiguous()],
                                                   dist.group.WORLD)
        assert torch.equal(cap_all, ic["input_embeddings"]) and torch.equal(mask_all, cap_mask)
        emb_loc = lsm_head.project_regions(ii["region_features"][sl], w, b)
        # raw (unguarded) blocks of this rank's images against ALL captions
        big = 1e30
        S_blocks = []
        for align in ("w2r", "r2w"):
            pass
        d1, d2 = _raw_pair(lsm_head, cap_all, mask_all, emb_loc, ii["region_mask"][sl].float())
        blocks = torch.stack([d1, d2]).requires_grad_(True)

        head = M.GroundingHead(M.get_cfg("lsm"), V, D)
        head._pair_outputs = lambda full, cm, rm: _cpu_pair_outputs(lsm_head, full, cm, rm)
        losses, info, dists = parallel.global_pair_outputs(head, blocks, mask_all, ii["region_mask"][sl].float(), dist.group.WORLD)
        # --- single-device oracle ------------------------------------------------------------------------
        rinfo, rlosses, rdists = lsm_head.grounding_head_forward(ii, ic, w, b)
        ok = all(torch.allclose(dists[k], rdists[k], rtol=1e-5, atol=1e-6) for k in rdists)
        ok &= all(torch.allclose(losses[k], rlosses[k], rtol=1e-5, atol=1e-6) for k in rlosses)
        ok &= all(float(info[k]) == float(rinfo[k]) for k in rinfo)
        ok &= torch.equal(blocks.detach(), torch.stack([_unguard(rdists["w2r"], d1, sl), _unguard(rdists["r2w"], d2, sl)]))
        # backward: the gradient of the replicated global loss w.r.t. this rank's block = its column slice
        total = sum(losses.values())
        total.backward()
        full = torch.stack([rdists["w2r"], rdists["r2w"]]).clone().requires_grad_(True)
        rl = _cpu_pair_outputs(lsm_head, full, cap_mask, ii["region_mask"].float())[0]
        sum(rl.values()).backward()
        ok &= torch.allclose(blocks.grad, full.grad[:, :, sl], rtol=1e-5, atol=1e-7)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _raw_pair(lsm_head, cap, cap_mask, emb, reg_mask):
    """Unguarded distances (what loco_lsm_pair_fwd returns): oracle arithmetic without the max+100 guard."""
    import torch.nn.functional as F
    S = torch.einsum("ctd,ird->citr", cap, emb) / 10.0
    valid = (cap_mask[:, None, :, None] * reg_mask[None, :, None, :]) > 0
    Sm = torch.where(valid, S, torch.full_like(S, -1e30))
    a1 = F.softmax(Sm, 3) * cap_mask[:, None, :, None]
    a2 = F.softmax(Sm, 2) * reg_mask[None, :, None, :]
    d1 = (a1 * -S).sum(3).sum(2) / cap_mask.sum(1).clamp(min=1)[:, None]
    d2 = (a2 * -S).sum(3).sum(2) / reg_mask.sum(1).clamp(min=1)[None, :]
    return d1, d2


def _unguard(guarded, raw_block, sl):
    out = guarded[:, sl].clone()
    g = guarded[:, sl] == guarded.max()
    out[g] = raw_block[g]
    return out if bool(g.any()) else guarded[:, sl]


def _cpu_pair_outputs(lsm_head, full, cap_mask, reg_mask):
    """CPU stand-in for GroundingHead._pair_outputs (pair-CE kernel): guard + oracle losses."""
    ok = (cap_mask.sum(1)[:, None] > 0) | (reg_mask.sum(1)[None, :] > 0)
    losses, info, dists = {}, {}, {}
    for k, (key, name) in enumerate((("w2r", "Words"), ("r2w", "Regions"))):
        pw = torch.where(ok, full[k], (full[k].max() + 100.0).detach())
        dists[key] = pw
        ce_cap, ce_img, acc_cap, acc_img = lsm_head.pair_losses(pw)
        losses[f"CE_loss (Align {name}, Choose Caption)"] = ce_cap
        losses[f"CE_loss (Align {name}, Choose Image)"] = ce_img
        info[f"Batch Accuracy (Align {name}, Choose Caption)"] = acc_cap
        info[f"Batch Accuracy (Align {name}, Choose Image)"] = acc_img
    return losses, info, dists


def test_sharded_exchange_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(p.exitcode == 0 for p in procs)
    assert res == [(0, True), (1, True)]
