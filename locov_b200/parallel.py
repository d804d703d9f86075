"""Batch-sharded LSM pair matrix over the GPUs of one NVSwitch box (BASELINE config 4).

The reference computes the pair matrix per GPU over its LOCAL batch only (no collective anywhere in
its heads).  The cross-batch variant defined here shards the IMAGES (columns of the
pair matrix): rank g owns the regions of its B_loc images and scores them against ALL B = W*B_loc
captions, so the one exchange step is an all-gather of the caption word embeddings (already split to
bf16 tensor-core operands: half the bytes of fp32) and caption masks, issued with NCCL on a side
stream so that it overlaps the projection GEMM, which needs no captions.  The [B, B_loc] distance
blocks (B*B_loc*4 bytes) are then all-gathered so that every rank evaluates the four global
cross-entropy losses / accuracies and the reference's global ``max + 100`` empty-pair guard on the full
[B, B] matrix.

Parity definition: for identical global inputs, rank g's block equals columns [g*B_loc, (g+1)*B_loc)
of the single-device matrix and the four global losses equal the single-device losses
(tests/test_parallel_*.py).  The masked fill is a fixed finite constant, never a per-shard minimum.

Exchange: on GPUs of one NVSwitch box both steps are PEER STORES into CUDA symmetric memory (torch.distributed._symmetric_memory
buffers mapped into every process): the rank that produces a slice (caption operand rows from ``loco_lsm_prep``, its distance block from
the pair kernel) stores it straight into every rank's gathered buffer and the SAME launch ends with a cross-GPU flag barrier
(``loco_peer_exchange``) — two small kernels per step on ONE stream, no NCCL launch, no stream hand-off, no re-layout copy (the blocks
land in place in the [2, B, B] matrices).  The NCCL all-gather implementation below remains as the path for
process groups without symmetric memory (and is what the gloo / CPU test of the exchange logic exercises); LOCOV_B200_SYMM=0 forces it.

Gradient convention: every rank holds the SAME global loss; its autograd path covers only its own
column block, so summing parameter gradients over ranks gives the single-device gradient (under DDP's
mean-reduction multiply the loss by the world size).  Caption embeddings are frozen BERT inputs in the
shipped configuration (coco_lsm.yaml: LANGUAGE_BACKBONE.FREEZE) and receive no gradient here.
"""
import os
from typing import Optional

import torch
import torch.distributed as dist
from torch.autograd import Function

from . import functional as LF
from . import ops

_NAMES = {"w2r": "Words", "r2w": "Regions"}
_side_streams = {}


def shard_grounding_head(head, process_group: Optional[dist.ProcessGroup] = None):
    """Switch a GroundingHead to the sharded cross-batch pair matrix."""
    head.process_group = process_group if process_group is not None else dist.group.WORLD
    head.shard_captions = True
    return head


def _side_stream(device):
    s = _side_streams.get(device)
    if s is None:
        s = _side_streams[device] = torch.cuda.Stream(device=device)
    return s


def gather_rows(t: torch.Tensor, group) -> torch.Tensor:
    """all-gather along dim 0 (equal shapes on every rank)."""
    w = dist.get_world_size(group)
    out = t.new_empty((w * t.shape[0],) + tuple(t.shape[1:]))
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


class _GatherBlocks(Function):
    """[n, B, B_loc] column blocks per rank (+ the rank's per-image valid-region counts, packed into the same
    message) -> full [n, B, B] matrices and the [B] region counts on every rank: ONE collective.  Backward = this
    rank's column slice of the (replicated) full gradient: no communication."""

    @staticmethod
    def forward(ctx, block, nreg_loc, group):
        w, r = dist.get_world_size(group), dist.get_rank(group)
        n, b, bl = block.shape
        parts, nreg_all = gather_packed([block.reshape(1, n, b, bl), nreg_loc.reshape(-1).to(block.dtype)], group)
        ctx.cols = (r * bl, (r + 1) * bl)
        full = parts.permute(1, 2, 0, 3).reshape(n, b, w * bl)          # [w, n, b, bl] -> [n, b, w * bl]: one copy
        ctx.mark_non_differentiable(nreg_all)
        return full, nreg_all

    @staticmethod
    def backward(ctx, g, _gn):
        c0, c1 = ctx.cols
        return g[:, :, c0:c1].contiguous(), None, None


def assemble_blocks(block, nreg_loc, group):
    return _GatherBlocks.apply(block, nreg_loc, group)


def global_pair_outputs(head, blocks, cap_mask_all, reg_mask_loc, group):
    """[2, B, B_loc] blocks -> (losses, other_info, dists) on the full matrices; shared by the CUDA path
    and the gloo/CPU test of the exchange logic (``head._pair_outputs`` evaluates the losses).  The empty-pair
    guard only needs to know WHICH images have no valid region, so the per-image counts travel with the blocks
    and stand in for the region mask ([B, 1])."""
    full, nreg_all = assemble_blocks(blocks, reg_mask_loc.to(torch.float32).sum(1), group)
    return head._pair_outputs(full, cap_mask_all, nreg_all.reshape(-1, 1))     # a [B, 1] "mask" whose row sum is the count


def gather_packed(tensors, group):
    """all-gather several per-rank tensors (equal shapes on every rank) along dim 0 as ONE coalesced collective: on NCCL
    the all-gathers are grouped (one launch) and land directly in their dense destinations — no pack / unpack copies."""
    w = dist.get_world_size(group)
    ins = [t.contiguous() for t in tensors]
    outs = [t.new_empty((w * t.shape[0],) + tuple(t.shape[1:])) for t in ins]
    if len(ins) > 1 and ins[0].is_cuda:
        with dist._coalescing_manager(group=group):
            for o, i in zip(outs, ins):
                dist.all_gather_into_tensor(o, i, group=group)
    else:
        for o, i in zip(outs, ins):
            dist.all_gather_into_tensor(o, i, group=group)
    return outs


class _ShardedLsm(Function):
    """Local regions x ALL captions -> ([B, B_loc] w2r block, r2w block)."""

    @staticmethod
    def forward(ctx, feats, w, b, cap_loc, cap_mask_loc, reg_mask, inv_temp, alignment, precision, want_w2r, want_r2w, group, cap_op):
        acc = LF._acc(precision)
        bi, rg, v = feats.shape
        bl, t, d = cap_loc.shape
        dev = feats.device
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        # 1. caption operands: split locally (bf16 hi [+lo]; already done by ops.lsm_prep when the module could fuse it
        #    with the mask preparation), all-gather on the side stream
        if cap_op is None or (cap_op.lo is None) == acc or cap_op.rows != bl * t or cap_op.cols != d:
            cap_op = ops.split_bf16(cap_loc.reshape(bl * t, d), acc)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            parts = [cap_op.hi, cap_mask_loc] + ([cap_op.lo] if acc else [])
            got = gather_packed(parts, group)                 # one grouped NCCL all-gather for operands + masks
            hi_all, mask_all = got[0], got[1]
            lo_all = got[2] if acc else None
            for x in parts:
                x.record_stream(side)          # allocated on the main stream, read by the side stream
        for x in got:
            x.record_stream(main)              # allocated on the side stream, consumed on the main stream
        # 2. projection of the local regions overlaps the gather
        emb_op = LF.project_regions(feats.reshape(bi * rg, v), w, b, d, acc)
        main.wait_stream(side)
        cap_all = ops.Bf16Operand(hi_all, lo_all, hi_all.shape[0], d)
        stack = LF.new_pair_stack(hi_all.shape[0] // t, bi, dev, want_w2r and want_r2w)
        ops.lsm_pair(cap_all, mask_all, emb_op, reg_mask, inv_temp, alignment, want_w2r, want_r2w, stack[0], stack[1])
        ctx.ops_saved = (emb_op, cap_all)
        ctx.save_for_backward(feats, w, mask_all, reg_mask)
        ctx.meta = (inv_temp, alignment, precision, b is not None, want_w2r, want_r2w)
        ctx.mark_non_differentiable(mask_all)
        return stack, mask_all

    @staticmethod
    def backward(ctx, g, _gm):
        feats, w, mask_all, reg_mask = ctx.saved_tensors
        emb_op, cap_all = ctx.ops_saved
        inv_temp, alignment, precision, has_b, want_w2r, want_r2w = ctx.meta
        acc = LF._acc(precision)
        bi, rg, v = feats.shape
        if ctx.needs_input_grad[3]:
            raise NotImplementedError("sharded LSM: caption embeddings are frozen inputs (LANGUAGE_BACKBONE.FREEZE); "
                                      "a reduce-scatter of d(captions) is not implemented")
        g = g.contiguous()
        demb, _ = ops.lsm_pair_bwd(cap_all, mask_all, emb_op, reg_mask, inv_temp, alignment,
                                   g[0] if want_w2r else None, g[1] if want_r2w else None, False)
        dx = dw = db = None
        g_op = ops.split_bf16(demb, acc)
        if ctx.needs_input_grad[0]:
            dx, _ = ops.linear_fwd(g_op, LF.weight_operand(w, acc, transpose=True), None, want_f32=True)
            dx = dx.reshape(bi, rg, v)
        if ctx.needs_input_grad[1]:
            x_op = ops.split_bf16(feats.reshape(bi * rg, v), acc)
            dw, _ = ops.linear_fwd(ops.transpose_operand(g_op), ops.transpose_operand(x_op), None, want_f32=True)
        if has_b and ctx.needs_input_grad[2]:
            db = demb.sum(0)
        return dx, dw, db, None, None, None, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# symmetric-memory exchange (peer stores + barriers)
# ------------------------------------------------------------------------------------------------
_symm_states = {}
_symm_disabled = {}


class _SymmState:
    """Symmetric memory of one (process group, shape): ONE allocation per rank holding every gathered tensor, plus the flag words of
    the in-kernel barriers — allocated and exchanged once (collective), reused by every step.

    Regions.  "A side" (written by exchange A, read by the pair kernel only): caption operands hi [B*T, ld] (+ lo), caption masks
    [B, T].  "B side" (written by exchange B, read by the pair-CE kernel only): pair matrices [2, B, B], caption masks [B, T], region
    masks [B, Rg].  With the two exchanges alternating, a rank can only be overwritten on one side while it is reading the other:
    single buffering is race-free (a peer cannot start exchange A of step k+1 before this rank signalled in exchange B of step k,
    i.e. after its pair kernel; it cannot start exchange B of step k+1 before this rank signalled in exchange A of step k+1, i.e.
    after its pair-CE kernel of step k)."""

    def __init__(self, group, dev, bl, t, d, ld, rg, acc):
        import torch.distributed._symmetric_memory as symm
        self.w, self.rank = dist.get_world_size(group), dist.get_rank(group)
        b = self.w * bl
        self.bl, self.t, self.d, self.ld, self.b, self.rg, self.acc = bl, t, d, ld, b, rg, acc
        sizes = [("hi", b * t * ld * 2), ("lo", b * t * ld * 2 if acc else 0), ("maskA", b * t * 4), ("pw", 2 * b * b * 4), ("maskB", b * t * 4),
                 ("reg", b * rg * 4)]
        self.off, total = {}, 0
        for name, n in sizes:
            self.off[name] = total
            total += (n + 255) // 256 * 256
        self.buf = symm.empty((total,), dtype=torch.uint8, device=dev)
        self.h = symm.rendezvous(self.buf, group)
        self.flags = symm.empty((16 * self.w,), dtype=torch.int32, device=dev)
        self.h_flags = symm.rendezvous(self.flags, group)
        self.buf.zero_()              # (the pad columns [d, ld) of the operand rows stay zero: whole rows are re-written, pads included)
        self.flags.zero_()
        self.ticket = torch.zeros(4, dtype=torch.int32, device=dev)
        self.side = torch.cuda.Stream(dev)       # exchange A runs beside the projection GEMM
        torch.cuda.synchronize(dev)
        self.h.barrier(channel=0)     # one-off: every rank has zeroed its flags before any exchange kernel runs

        def view(name, dtype, shape):
            n = 1
            for k in shape:
                n *= k
            nbytes = n * torch.empty((), dtype=dtype).element_size()
            return self.buf[self.off[name]:self.off[name] + nbytes].view(dtype).view(shape)
        self.hi = view("hi", torch.bfloat16, (b * t, ld))
        self.lo = view("lo", torch.bfloat16, (b * t, ld)) if acc else None
        self.maskA = view("maskA", torch.float32, (b, t))
        self.pw = view("pw", torch.float32, (2, b, b))
        self.maskB = view("maskB", torch.float32, (b, t))
        self.reg = view("reg", torch.float32, (b, rg))

    def exchange(self, segments, channel, mode):
        """segments: [(region name, src tensor 2-D, dst pitch bytes, byte offset inside the region)]; mode: ops.PEER_* bits"""
        dbg = os.environ.get("LOCOV_B200_SYMM_DEBUG", "0")     # developer timing experiments: 1 = no exchange at all, 2 = stores without signals / waits
        if dbg == "1":
            return
        if dbg == "2":
            mode &= ops.PEER_STORE
            if mode == 0:
                return
        ops.peer_exchange([(src, pitch, self.off[name] + off) for name, src, pitch, off in segments], self.h.buffer_ptrs_dev, self.w,
                          self.h_flags.buffer_ptrs_dev, self.rank, channel, mode, self.ticket, device=self.buf.device)


def _symm_state(group, dev, bl, t, d, ld, rg, acc):
    """The symmetric buffers for this shape, or None when symmetric memory is unavailable for the group (then NCCL is used)."""
    if os.environ.get("LOCOV_B200_SYMM", "1") == "0" or dist.get_backend(group) != "nccl":
        return None
    gkey = id(group)
    if _symm_disabled.get(gkey):
        return None
    key = (gkey, dev.index, bl, t, d, ld, rg, acc)
    st = _symm_states.get(key)
    if st is None:
        try:
            st = _SymmState(group, dev, bl, t, d, ld, rg, acc)
        except Exception as e:      # noqa: BLE001  (no fabric / IPC support in this process group: fall back to NCCL, once, loudly)
            import warnings
            warnings.warn(f"locov_b200: CUDA symmetric memory is unavailable for this process group ({e}); the sharded LSM head uses NCCL all-gathers")
            _symm_disabled[gkey] = True
            return None
        _symm_states[key] = st
    return st


class _ShardedLsmSymm(Function):
    """Local regions x ALL captions with the symmetric-memory exchange -> (full [2, B, B] matrices, caption masks [B, T], region counts [B])."""

    @staticmethod
    def forward(ctx, feats, w, b, cap_loc, cap_mask_loc, reg_mask, inv_temp, alignment, precision, want_w2r, want_r2w, st, cap_op):
        acc = LF._acc(precision)
        bi, rg, v = feats.shape
        bl, t, d = cap_loc.shape
        dev = feats.device
        if cap_op is None or (cap_op.lo is None) == acc or cap_op.rows != bl * t or cap_op.cols != d or cap_op.ld != st.ld:
            cap_op = ops.split_bf16(cap_loc.reshape(bl * t, d), acc)
        # 1. exchange A: this rank's caption operand rows and caption masks -> every rank's gathered buffers, barrier in the same launch
        cm = cap_mask_loc.to(torch.float32).contiguous().reshape(1, -1)
        row0 = st.rank * bl * t
        segs = [("hi", cap_op.hi, st.ld * 2, row0 * st.ld * 2), ("maskA", cm, bl * t * 4, row0 * 4)]
        if acc:
            segs.append(("lo", cap_op.lo, st.ld * 2, row0 * st.ld * 2))
        # 2. the projection of the local regions needs no captions.  Exchange A (stores + fence + signal) runs on a side stream BESIDE
        #    the projection GEMM (130 of the 148 SMs), so the NVLink round trips and the system-scope fence are off the critical path;
        #    after the GEMM the main stream only polls flags that have long been set.  (LOCOV_B200_SYMM_SIDE=0: stores, GEMM, then
        #    signal + wait in stream order — the round-2a arrangement, kept for A/B measurements.)
        if os.environ.get("LOCOV_B200_SYMM_SIDE", "1") != "0":
            cur = torch.cuda.current_stream(dev)
            fork, join = torch.cuda.Event(), torch.cuda.Event()
            fork.record(cur)
            st.side.wait_event(fork)
            with torch.cuda.stream(st.side):
                st.exchange(segs, 0, ops.PEER_STORE | ops.PEER_SIGNAL)
                join.record(st.side)
            for t_ in (cap_op.hi, cap_op.lo, cm):
                if t_ is not None:
                    t_.record_stream(st.side)
            emb_op = LF.project_regions(feats.reshape(bi * rg, v), w, b, d, acc)
            cur.wait_event(join)
            st.exchange([], 0, ops.PEER_WAIT)           # every rank's slices have landed here
        else:
            st.exchange(segs, 0, ops.PEER_STORE)
            emb_op = LF.project_regions(feats.reshape(bi * rg, v), w, b, d, acc)
            st.exchange([], 0, ops.PEER_SIGNAL | ops.PEER_WAIT)
        cap_all = ops.Bf16Operand(st.hi, st.lo if acc else None, st.b * t, d)
        stack = LF.new_pair_stack(st.b, bi, dev, want_w2r and want_r2w)
        ops.lsm_pair(cap_all, st.maskA, emb_op, reg_mask, inv_temp, alignment, want_w2r, want_r2w, stack[0], stack[1])
        # 3. exchange B: this rank's [2, B, B_loc] block -> its columns of every rank's [2, B, B] matrices, in place (no re-layout), with
        #    the masks the pair-CE kernel's empty-pair guard reads
        rm = reg_mask.to(torch.float32).contiguous()
        st.exchange([("pw", stack.reshape(2 * st.b, bi), st.b * 4, st.rank * bi * 4), ("maskB", cm, bl * t * 4, row0 * 4),
                     ("reg", rm.reshape(1, -1), bi * rg * 4, st.rank * bi * rg * 4)], 1, ops.PEER_STORE | ops.PEER_SIGNAL | ops.PEER_WAIT)
        full = st.pw.clone()                            # (the caller keeps the matrices; the symmetric buffer is rewritten by the next step)
        mask_all, nreg_all = st.maskB, st.reg           # read by the pair-CE kernel of this step only
        need = any(ctx.needs_input_grad[:3])
        ctx.ops_saved = (emb_op, ops.Bf16Operand(st.hi.clone(), st.lo.clone() if acc else None, st.b * t, d) if need else None)
        ctx.save_for_backward(feats, w, st.maskA.clone() if need else mask_all, reg_mask)
        ctx.meta = (inv_temp, alignment, precision, b is not None, want_w2r, want_r2w, st.rank * bi, bi)
        ctx.mark_non_differentiable(mask_all, nreg_all)
        return full, mask_all, nreg_all

    @staticmethod
    def backward(ctx, g, _gm, _gn):
        feats, w, mask_all, reg_mask = ctx.saved_tensors
        emb_op, cap_all = ctx.ops_saved
        inv_temp, alignment, precision, has_b, want_w2r, want_r2w, c0, bl = ctx.meta
        acc = LF._acc(precision)
        bi, rg, v = feats.shape
        if ctx.needs_input_grad[3]:
            raise NotImplementedError("sharded LSM: caption embeddings are frozen inputs (LANGUAGE_BACKBONE.FREEZE); "
                                      "a reduce-scatter of d(captions) is not implemented")
        g = g[:, :, c0:c0 + bl].contiguous()            # this rank's column block of the (replicated) full gradient: no communication
        demb, _ = ops.lsm_pair_bwd(cap_all, mask_all, emb_op, reg_mask, inv_temp, alignment,
                                   g[0] if want_w2r else None, g[1] if want_r2w else None, False)
        dx = dw = db = None
        g_op = ops.split_bf16(demb, acc)
        if ctx.needs_input_grad[0]:
            dx, _ = ops.linear_fwd(g_op, LF.weight_operand(w, acc, transpose=True), None, want_f32=True)
            dx = dx.reshape(bi, rg, v)
        if ctx.needs_input_grad[1]:
            x_op = ops.split_bf16(feats.reshape(bi * rg, v), acc)
            dw, _ = ops.linear_fwd(ops.transpose_operand(g_op), ops.transpose_operand(x_op), None, want_f32=True)
        if has_b and ctx.needs_input_grad[2]:
            db = demb.sum(0)
        return dx, dw, db, None, None, None, None, None, None, None, None, None, None


def sharded_grounding_forward(head, region_features, region_mask, caption_emb, caption_mask, cap_op=None):
    group = head.process_group
    amode = {"softmax": ops.ALIGN_SOFTMAX, "hardmax": ops.ALIGN_HARDMAX}[head.alignment]
    st = None
    if region_features.is_cuda and caption_emb.dim() == 3:
        bl, t, d = caption_emb.shape
        st = _symm_state(group, region_features.device, bl, t, d, (d + 7) // 8 * 8, region_features.shape[1], head.precision == "fp32")
    if st is not None:
        full, mask_all, nreg_all = _ShardedLsmSymm.apply(
            region_features.to(torch.float32).contiguous(), head.v2l_projection.weight, head.v2l_projection.bias,
            caption_emb.to(torch.float32).contiguous(), caption_mask, region_mask, 1.0 / float(head.temperature), amode,
            head.precision, bool(head.align_words), bool(head.align_regions), st, cap_op)
        losses, info, dists = head._pair_outputs(full, mask_all, nreg_all)
        head.log_dict(losses)
        head.log_dict(info)
        if head.return_dist:
            return info, losses, dists
        return info, losses
    blocks, mask_all = _ShardedLsm.apply(
        region_features.to(torch.float32).contiguous(), head.v2l_projection.weight, head.v2l_projection.bias,
        caption_emb.to(torch.float32).contiguous(), caption_mask, region_mask, 1.0 / float(head.temperature), amode,
        head.precision, bool(head.align_words), bool(head.align_regions), group, cap_op)
    losses, info, dists = global_pair_outputs(head, blocks, mask_all, region_mask, group)
    head.log_dict(losses)
    head.log_dict(info)
    if head.return_dist:
        return info, losses, dists
    return info, losses
