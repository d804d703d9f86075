"""Multi-GPU parity of the image-sharded LSM head (locov_b200/parallel.py) on real GPUs over NCCL:
every rank's [B, B_loc] block must equal its column slice of the single-device pair matrices and the four
global losses / accuracies must equal the single-device ones (SURVEY.md 8e parity definition).
usage: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/check_sharded_nccl.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import locov_b200.modeling as M  # noqa: E402
from locov_b200 import parallel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # default: a small ragged case; "bench" = the benchmarked shard shape (32 images x 100 regions x 20 tokens per rank)
    bl, rg, t, v, d = (32, 100, 20, 256, 768) if "bench" in sys.argv else (8, 37, 12, 256, 768)
    b = bl * world
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(b, rg, v, generator=g)
    caps = torch.randn(b, t, d, generator=g) * 0.05
    att = (torch.rand(b, t, generator=g) > 0.2).long()
    att[:, 0] = 1
    spe = torch.zeros(b, t, dtype=torch.long)
    spe[:, 0] = 1
    att[1] = 0                                              # an empty caption
    rmask = (torch.rand(b, rg, generator=g) > 0.1).to(torch.uint8)
    rmask[b - 1] = 0                                        # an image without valid regions
    ok = True
    for precision, tol in (("fp32", 1e-4), ("bf16", 2e-2)):
        cfg = M.get_cfg("lsm")
        cfg.MODEL.B200.PRECISION = precision
        cfg.MODEL.MMSS_HEAD.DISTILLATION_LOSS = True
        torch.manual_seed(3)
        head = M.GroundingHead(cfg, v, d).to(dev)
        ref = M.GroundingHead(cfg, v, d).to(dev)
        ref.load_state_dict(head.state_dict())
        parallel.shard_grounding_head(head)
        sl = slice(rank * bl, (rank + 1) * bl)
        with torch.no_grad():
            info, losses, dists = head({"region_features": feats[sl].to(dev), "region_mask": rmask[sl].to(dev)},
                                       {"input_embeddings": caps[sl].to(dev), "attention_mask": att[sl].to(dev),
                                        "special_tokens_mask": spe[sl].to(dev)})
            rinfo, rlosses, rdists = ref({"region_features": feats.to(dev), "region_mask": rmask.to(dev)},
                                         {"input_embeddings": caps.to(dev), "attention_mask": att.to(dev), "special_tokens_mask": spe.to(dev)})
        for k in rdists:
            a, r = dists[k], rdists[k]
            err = float((a - r).abs().max() / r.abs().max().clamp(min=1e-6))
            ok &= err < tol
            ok &= bool(torch.equal(a[:, sl], r[:, sl]))         # this rank's own block: bit-identical (fixed-order reductions)
            if rank == 0:
                print(f"{precision} dist {k}: max rel err vs single device {err:.2e}")
        for k in rlosses:
            err = abs(float(losses[k]) - float(rlosses[k])) / max(abs(float(rlosses[k])), 1e-6)
            ok &= err < tol
        for k in rinfo:
            ok &= abs(float(info[k]) - float(rinfo[k])) < (1e-6 if precision == "fp32" else 0.2)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        path = "NCCL all-gathers" if os.environ.get("LOCOV_B200_SYMM", "1") == "0" or not parallel._symm_states else "symmetric-memory peer stores"
        print("sharded NCCL parity:", "OK" if flag.item() > 0 else "FAILED", f"(world {world}, B_loc {bl}, exchange: {path})")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() > 0 else 1)


if __name__ == "__main__":
    main()
