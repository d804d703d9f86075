"""LSM grounding head — drop-in for the reference's ``GroundingHead``
(ovr/modeling/mmss_heads/grounding_head.py:50-388), registered under the same name in
``MMSS_HEADS_REGISTRY`` and exposing ``v2l_projection.{weight,bias}`` as ordinary Parameters so the
weight tying in mmss_heads.py:29-40 / distill_prop_mmss_gcnn.py:117-125 keeps working.

Forward = four kernel launches instead of the reference's ~60 ATen/cuBLAS launches and B^2
replication: bf16 operand split, tcgen05 projection GEMM, the fused pair kernel (similarity GEMM +
masked softmax both ways + attention pooling, grounding_head.py:116-256) and one small kernel for the
empty-pair guard, the four cross-entropy losses and the four batch accuracies (:240-251, :272-290,
:354-379).  Loss / metric key strings are the reference's.

Options outside the shipped configuration raise NotImplementedError at construction, exactly where
the reference itself raises for unknown values — never a silent difference.
"""
import torch
from torch import nn

from .. import functional as LF
from .. import ops
from .logged_module import LoggedModule
from .registry import MMSS_HEADS_REGISTRY

__all__ = ["GroundingHead", "build_grounding_head", "MMSS_HEADS_REGISTRY"]

_NAMES = {"w2r": "Words", "r2w": "Regions"}


@MMSS_HEADS_REGISTRY.register()
class GroundingHead(LoggedModule):
    def __init__(self, config, v_dim, l_dim, *args, **kwargs):
        super().__init__()
        self.config = config.MODEL.MMSS_HEAD.GROUNDING
        self.v_dim = v_dim
        self.l_dim = l_dim
        self.v2l_projection = nn.Linear(self.v_dim, self.l_dim)
        self.local_metric = self.config.LOCAL_METRIC
        self.global_metric = self.config.GLOBAL_METRIC
        self.alignment = self.config.ALIGNMENT
        self.temperature = self.config.ALIGNMENT_TEMPERATURE
        self.loss_type = self.config.LOSS
        self.negative_mining = self.config.NEGATIVE_MINING
        self.margin = self.config.TRIPLET_MARGIN
        self.align_words = self.config.ALIGN_WORDS_TO_REGIONS
        self.align_regions = self.config.ALIGN_REGIONS_TO_WORDS
        assert self.align_words or self.align_regions
        self.return_dist = config.MODEL.MMSS_HEAD.DISTILLATION_LOSS
        self.grounding_text_input = self.config.TEXT_INPUT
        b200 = getattr(config.MODEL, "B200", None)
        self.precision = getattr(b200, "PRECISION", "fp32") if b200 is not None else "fp32"
        # sharded pair matrix: set by locov_b200.parallel.shard_grounding_head
        self.process_group = None
        self.shard_captions = False

        if self.local_metric != "dot":
            raise NotImplementedError(f"LOCAL_METRIC {self.local_metric!r} (the reference implements 'dot' only, grounding_head.py:146-150)")
        if self.global_metric != "aligned_local":
            raise NotImplementedError(f"GLOBAL_METRIC {self.global_metric!r}: only 'aligned_local' (shipped, coco_lsm.yaml) runs on the B200 path")
        if self.alignment not in ("softmax", "hardmax"):
            raise NotImplementedError(f"ALIGNMENT {self.alignment!r}: 'softmax' (shipped) and 'hardmax' run on the B200 path")
        if self.loss_type not in ("cross_entropy", "triplet"):
            if self.loss_type == "matching":
                raise Exception("Matching loss is not defined for dot product because dot product is unbounded")
            raise NotImplementedError(f"LOSS {self.loss_type!r}")

    # ---------------------------------------------------------------------------------------------
    def forward(self, input_image, input_caption):
        caption_emb = input_caption[self.grounding_text_input]
        att, spe = input_caption["attention_mask"], input_caption["special_tokens_mask"]
        region_features = input_image["region_features"]
        region_mask = input_image["region_mask"]
        sharded = self.shard_captions and self.process_group is not None
        cap_op = x_op = w_op = None
        acc = self.precision == "fp32"
        masks_on_device = att.is_cuda and att.dtype == torch.int64 and spe.dtype == torch.int64 and region_mask.dtype in ops._REG_KIND
        if masks_on_device and caption_emb.is_cuda and caption_emb.dim() == 3:
            # grounding_head.py:94-106 and the caption operand of the pair GEMM in one launch; in the fp32-accurate mode the same launch
            # also converts the region features and (unless its operand is cached) the projection weight: the small jobs are
            # latency-bound on their own and free inside the HBM-bound feature split
            cap32 = caption_emb.to(torch.float32).contiguous()
            extra = []
            fuse = acc and not sharded and region_features.is_cuda and region_features.dim() == 3 and not LF._use_tf32(acc, region_features.shape[-1])
            if fuse:
                feats32 = region_features.to(torch.float32).contiguous()
                extra.append(feats32.reshape(-1, feats32.shape[-1]))
                wgt = self.v2l_projection.weight
                w_op = LF.weight_operand_if_cached(wgt, acc)
                if w_op is None:
                    extra.append(wgt.detach())
            got = ops.lsm_prep(cap32.reshape(-1, cap32.shape[-1]), acc, att, spe, region_mask, extra=extra)
            cap_op, caption_mask, region_mask = got[:3]
            if fuse:
                x_op = got[3][0]
                if w_op is None:
                    w_op = got[3][1]
                    LF.adopt_weight_operand(wgt, acc, w_op)
        elif masks_on_device:
            caption_mask, region_mask = ops.lsm_masks(att, spe, region_mask)       # grounding_head.py:94-106, one launch
        else:
            caption_mask = (att * (1 - spe)).to(torch.float32)
            region_mask = region_mask.to(torch.float32)
        self.log("attention_mask", att)
        self.log("special_tokens_mask", spe)
        self.log("caption_mask", caption_mask)
        self.log("caption_emb", caption_emb)
        self.log("region_features", region_features)
        self.log("region_mask", region_mask)
        batch_size = region_features.shape[0]

        if sharded:
            from .. import parallel
            return parallel.sharded_grounding_forward(self, region_features, region_mask, caption_emb, caption_mask, cap_op=cap_op)

        pw = LF.lsm_head(region_features.to(torch.float32).contiguous(), self.v2l_projection.weight,
                         self.v2l_projection.bias, caption_emb.to(torch.float32).contiguous(), caption_mask,
                         region_mask, self.temperature, self.alignment, self.precision,
                         want_w2r=self.align_words, want_r2w=self.align_regions, cap_op=cap_op, x_op=x_op, w_op=w_op)
        assert pw.shape == (2, batch_size, batch_size)
        losses, other_info, dists = self._pair_outputs(pw, caption_mask, region_mask)
        self.log_dict(losses)
        self.log_dict(other_info)
        if self.return_dist:
            return other_info, losses, dists
        return other_info, losses

    def _pair_outputs(self, pw, caption_mask, region_mask):
        """Stacked [2,B,B] distances -> (losses, other_info, dists) with the reference's key strings."""
        pw_g, out = LF.pair_losses(pw, caption_mask, region_mask, 0)
        # guarded matrix: value of pw_g, gradient path of pw (guard entries are constants)
        pw_cost = pw + (pw_g - pw).detach() if pw.requires_grad else pw_g
        losses, other_info, dists = {}, {}, {}
        for k, (key, on) in enumerate((("w2r", self.align_words), ("r2w", self.align_regions))):
            if not on:
                continue
            self.log(f"global_dist_{key}", pw_cost[k])
            dists[key] = pw_cost[k]
            name = _NAMES[key]
            if self.loss_type == "cross_entropy":
                losses[f"CE_loss (Align {name}, Choose Caption)"] = out[k, 0]
                losses[f"CE_loss (Align {name}, Choose Image)"] = out[k, 1]
            else:
                cap_l, img_l = self._triplet(pw_cost[k])
                losses[f"Triplet Loss (Align {name}, Choose Caption)"] = cap_l
                losses[f"Triplet Loss (Align {name}, Choose Image)"] = img_l
            other_info[f"Batch Accuracy (Align {name}, Choose Caption)"] = out[k, 2].detach()
            other_info[f"Batch Accuracy (Align {name}, Choose Image)"] = out[k, 3].detach()
        return losses, other_info, dists

    def _triplet(self, pw):
        """grounding_head.py:292-350 on the [B,B] matrix (a few hundred floats: PyTorch glue, not hot path)."""
        n = pw.shape[0]
        pos = torch.diag(pw)
        if n < 2:
            neg_cap = neg_img = pos + self.margin
        else:
            off = ~torch.eye(n, dtype=torch.bool, device=pw.device)
            big = torch.finfo(pw.dtype).max
            if self.negative_mining == "hardest":
                neg_cap = torch.where(off, pw, pw.new_full((), big)).min(dim=0).values
                neg_img = torch.where(off, pw, pw.new_full((), big)).min(dim=1).values
            elif self.negative_mining == "easiest":
                neg_cap = torch.where(off, pw, pw.new_full((), -big)).max(dim=0).values
                neg_img = torch.where(off, pw, pw.new_full((), -big)).max(dim=1).values
            elif self.negative_mining == "random":
                # a random off-diagonal row per column / column per row (remove_diag + randint gather)
                ar = torch.arange(n, device=pw.device)
                rc = torch.randint(n - 1, (n,), device=pw.device)
                rc = rc + (rc >= ar).to(rc.dtype)
                neg_cap = pw[rc, ar]
                ri = torch.randint(n - 1, (n,), device=pw.device)
                ri = ri + (ri >= ar).to(ri.dtype)
                neg_img = pw[ar, ri]
            else:
                raise NotImplementedError(self.negative_mining)
        relu = torch.nn.functional.relu
        return torch.mean(relu(pos - neg_cap + self.margin)), torch.mean(relu(pos - neg_img + self.margin))


def build_grounding_head(name, cfg, v_dim, l_dim, *args, **kwargs):
    return MMSS_HEADS_REGISTRY.get(name)(cfg, v_dim, l_dim)
