#!/usr/bin/env python
"""bench.py — region-text scores/sec of LocOV's region-text matching path on B200.

Headline (`value`, `ms_per_step`): BASELINE.json configs[1], the LSM grounding head — 32 images x 100 regions x 20 caption tokens
per GPU, 2048 -> 768 projection, full image-caption pair matrix, both alignments, 4 CE losses + 4 accuracies — in the
fp32-ACCURATE mode (three bf16 tensor-core passes, 1e-4 of the fp32 reference: the reference's own precision); the bf16 mode
(north_star's 2e-2 bar) is measured in the same run and reported under `precisions`.  One "step" = one forward pass of the head
over one batch of synthetic inputs.  The other half of the metric — RoI x class scoring (BASELINE configs[0], [2], [4]) and RoIAlign —
is measured in the same run and reported under `workloads`, each with its own roofline.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (the product)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU arithmetic (oracle port) on the same workload

N > 1 (torchrun): each rank keeps 32 images; the pair matrix becomes the global [32N x 32N] matrix, sharded by image with the caption
operands all-gathered over NCCL/NVLink (locov_b200/parallel.py) — weak scaling in images per GPU; N = 8 is BASELINE configs[3]
(global batch 256).  The RoI x class stress case (configs[4]: 8 x 1000 RoIs vs 1203 classes) is sharded by image with no collective.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA-graph replay of the step, inputs already in HBM, a ring of
input sets larger than L2 so no step re-reads cached inputs); `e2e` = the same step called through the GroundingHead module with pinned
HOST buffers, H2D / D2H copies inside the timed region; `roofline` = the dominant kernel timed alone with CUDA events;
`cpu_baseline` = oracle port of the reference on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_LOC, RG, T, V, D = 32, 100, 20, 2048, 768
SEED = 1992
METRIC = "region-text scores/sec (LSM grounding head: region x word similarities scored per second)"
WORKLOAD = "BASELINE configs[1]: LSM grounding head, 32 images x 100 regions x 20 caption tokens per GPU, 2048->768 projection, full pair matrix, fwd"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"], help="mode of the headline value (the other one is reported under `precisions`)")
    ap.add_argument("--sets", type=int, default=6, help="ring of distinct input batches (6 x 28 MB > 126 MB L2)")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the RoIAlign / RoI x class workloads (headline only)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions run."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (ts, r) in self.rows if t0 - 0.05 <= ts <= t1 + 0.15] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(seed, b=B_LOC):
    from locov_b200 import synthetic      # seeded workload definitions; value-for-value the oracle's generator
    return synthetic.lsm_inputs(B=b, Rg=RG, T=T, V=V, D=D, seed=seed, ragged_regions=True)


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own arithmetic on the host cores, on the SAME workload as the b200 arm at this N
# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(ii, ic, w, b, literal=True):
    import torch
    from oracle import lsm_head
    with torch.no_grad():
        if literal:      # grounding_head.py:92-388 line by line, incl. the B^2 replication and the per-call logging copies
            return lsm_head.grounding_head_forward_literal(ii, ic, w, b, 10.0)
        return lsm_head.grounding_head_forward(ii, ic, w, b)     # closed form (the literal replication needs ~20 GB at B = 256)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    world = max(1, args.gpus)
    b_glob = B_LOC * world
    literal = b_glob <= 64
    ii, ic, w, b = make_inputs(SEED, b_glob)
    for _ in range(max(1, min(args.warmup, 3 if world == 1 else 1))):
        cpu_reference_pass(ii, ic, w, b, literal)
    steps = args.steps
    t0 = time.perf_counter()
    cpu_reference_pass(ii, ic, w, b, literal)
    one = time.perf_counter() - t0
    budget = 200.0
    timed = max(1, min(steps, int(budget / max(one, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(timed):
        cpu_reference_pass(ii, ic, w, b, literal)
    dt = (time.perf_counter() - t0) / timed
    scores = b_glob * T * b_glob * RG
    val = scores / dt
    cores = torch.get_num_threads()
    sample = (f"one global batch per step ({b_glob} images x {RG} regions x {T} tokens = {scores} scores, the workload of the b200 arm at {world} GPU(s)); "
              f"{'literal port of GroundingHead.forward' if literal else 'closed-form port (chunked einsum; the literal B^2 replication needs ~20 GB here)'}; "
              f"{timed} of the {steps} requested steps timed ({budget:.0f} s cap)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "scores/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "device": "host CPU", "images_per_gpu": B_LOC, "global_batch": b_glob, "regions": RG, "tokens": T, "scores_per_step": scores},
        "cpu_baseline": {"value": val, "unit": "scores/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def graph_time(torch, fn, iters=16, reps=5, warm=3):
    """Average device time (ms) of ONE call of fn(i): `iters` calls (rotating operands) captured in a CUDA graph so that host launch
    overhead (longer than most of these kernels) is not in the number; best of `reps` replays."""
    side = torch.cuda.Stream()                       # warm up on a side stream (as the capture runs on one): autograd's AccumulateGrad
    side.wait_stream(torch.cuda.current_stream())    # nodes must not be pinned to the legacy stream, or a captured backward fails
    with torch.cuda.stream(side):
        for i in range(warm):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b_.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b_) / iters)
    return best


def run_b200(args):
    import torch
    import torch.distributed as dist
    import locov_b200.modeling as M
    from locov_b200 import _lib, ops, parallel, synthetic
    from locov_b200 import functional as LF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.loco_device_check(local), "loco_device_check")
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs -------------------------------------------------------------------------------------
    host_sets, dev_sets = [], []
    for s in range(args.sets):
        ii, ic, w, b = make_inputs(SEED + rank + 1000 * s)
        hs = ({k: v.pin_memory() for k, v in ii.items()}, {k: v.pin_memory() for k, v in ic.items()})
        host_sets.append(hs)
        dev_sets.append(({k: v.to(dev) for k, v in hs[0].items()}, {k: v.to(dev) for k, v in hs[1].items()}))
    _, _, w0, b0 = make_inputs(SEED)        # identical weights on every rank
    b_glob = B_LOC * world
    scores_per_step = b_glob * T * b_glob * RG            # all ranks together: [B*T] x [B*Rg]
    h2d = sum(v.numel() * v.element_size() for d_ in host_sets[0] for v in d_.values())

    def make_head(precision):
        cfg = M.get_cfg("lsm")
        cfg.MODEL.B200.PRECISION = precision
        head = M.GroundingHead(cfg, V, D).to(dev)
        with torch.no_grad():
            head.v2l_projection.weight.copy_(w0)
            head.v2l_projection.bias.copy_(b0)
        if world > 1:
            parallel.shard_grounding_head(head)
        return head

    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()

    # ---- timed region: device-resident throughput, once per precision ---------------------------------------------
    def device_resident(precision, steps, warmup):
        head = make_head(precision)

        def step_eager(s):
            with torch.no_grad():
                return head(dev_sets[s][0], dev_sets[s][1])

        # eager warm-up (library caches, NCCL communicator), then one CUDA graph per input set.  The bf16 weight shadow is dropped
        # before each capture so that every step re-converts the projection weights (a training step sees new weights every iteration).
        for s in range(args.sets):
            step_eager(s)
        barrier()
        graphs, keep, launches = [], [], None
        for s in range(args.sets):
            LF.clear_weight_cache()
            g = torch.cuda.CUDAGraph()
            n0 = lib.loco_launch_count()
            with torch.cuda.graph(g):
                o = step_eager(s)
            launches = lib.loco_launch_count() - n0
            graphs.append(g)
            keep.append(o)
        barrier()
        for i in range(max(warmup, 3)):
            graphs[i % args.sets].replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            graphs[i % args.sets].replay()
        e1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1)) / steps
        barrier()
        return head, {"ms_per_step": ms, "value": scores_per_step / (ms * 1e-3), "unit": "scores/s", "gpu_launches_per_step": int(launches),
                      "dtype": "bf16x3 (fp32-accurate split: 1e-4 of the fp32 reference)" if precision == "fp32" else "bf16 (TF32 projection + bf16 pair GEMM: 2e-2 bar)"}

    other = "bf16" if args.precision == "fp32" else "fp32"
    _, res_other = device_resident(other, max(200, args.steps // 4), args.warmup)
    head, res_main = device_resident(args.precision, args.steps, args.warmup)
    precisions = {args.precision: res_main, other: res_other}

    # ---- per-kernel breakdown + roofline (each kernel timed alone, CUDA events, rotating operands > L2) ---
    def lsm_kernels(precision):
        acc = precision == "fp32"
        nrot = max(args.sets, 12)
        feats = [torch.randn(B_LOC * RG, V, device=dev) for _ in range(nrot)]
        x_ops = [ops.split_bf16(f, acc) for f in feats]
        wgt = head.v2l_projection.weight.detach()
        bias = head.v2l_projection.bias.detach()
        w_op = ops.split_bf16(wgt, acc)
        mask_c = torch.ones(b_glob, T, device=dev)
        mask_r = torch.ones(B_LOC, RG, device=dev)
        cap_ops = [ops.split_bf16(torch.randn(b_glob * T, D, device=dev) * 0.05, acc) for _ in range(4)]
        emb_ops = [ops.linear_fwd(x_ops[i], w_op, bias, want_f32=False, n_bf16=D, accurate_out=acc)[1] for i in range(4)]
        pw = torch.randn(2, b_glob, b_glob, device=dev)
        hi_buf = x_ops[0]
        # the step's preparation launch: token / region masks + caption operand, and in the fp32-accurate mode the operands of the region
        # features and of the projection weight as well (ops.lsm_prep with `extra`)
        ic0, ii0 = dev_sets[0][1], dev_sets[0][0]
        cap2d = ic0["input_embeddings"].reshape(-1, D).to(torch.float32)
        prep_extra = (lambda i: [feats[i % nrot], wgt]) if (acc and world == 1) else (lambda i: [])
        k_ms = {
            "lsm_prep (masks + operands, one launch)": graph_time(torch, lambda i: ops.lsm_prep(cap2d, acc, ic0["attention_mask"], ic0["special_tokens_mask"],
                                                                                              ii0["region_mask"], extra=prep_extra(i))),
            "split_bf16(features)": graph_time(torch, lambda i: ops.split_bf16(feats[i % nrot], acc, out=hi_buf)) if acc else 0.0,
            "tc_gemm<EpiLinear> (projection)": graph_time(
                torch, (lambda i: ops.linear_fwd(x_ops[i % nrot], w_op, None, want_f32=False, n_bf16=D, accurate_out=acc)) if acc else
                (lambda i: ops.linear_tf32_fwd(feats[i % nrot], wgt, None, want_f32=False, n_bf16=D))),
            "tc_gemm<EpiLsmFwd> (pair)": graph_time(torch, lambda i: ops.lsm_pair(cap_ops[i % 4], mask_c, emb_ops[i % 4], mask_r, 0.1)),
            "pair_ce": graph_time(torch, lambda i: ops.pair_ce(pw, mask_c, mask_r)),
        }
        m_rows = B_LOC * RG
        gemm_flops = 2.0 * m_rows * V * D
        pair_flops = 2.0 * (b_glob * T) * (B_LOC * RG) * D               # one similarity GEMM serves both alignments
        split_bytes = m_rows * V * (4 + 2 * (2 if acc else 1))
        prep_bytes = (cap2d.numel() * (4 + 2 * (2 if acc else 1)) + ((m_rows * V + V * D) * (4 + 2 * 2) if (acc and world == 1) else 0))
        kern = {
            "lsm_prep (masks + operands, one launch)": {"bound": "hbm", "ms": k_ms["lsm_prep (masks + operands, one launch)"],
                                                        "achieved": prep_bytes / (k_ms["lsm_prep (masks + operands, one launch)"] * 1e-3) / 1e9,
                                                        "peak": pk["hbm_gbs"], "unit": "GB/s",
                                                        "note": "fp32-accurate mode, one GPU: region features + projection weight + captions -> bf16 (hi, lo) and both masks; "
                                                                "otherwise captions + masks only (then latency, not bandwidth, bound)"},
            "split_bf16(features)": {"bound": "hbm", "ms": k_ms["split_bf16(features)"],
                                     "achieved": split_bytes / (max(k_ms["split_bf16(features)"], 1e-9) * 1e-3) / 1e9 if acc else 0.0,
                                     "peak": pk["hbm_gbs"], "unit": "GB/s", "note": "the feature split timed ALONE for reference: in the step it is part of the preparation launch above (sharded head: a launch of its own); "
                                             "the bf16 mode multiplies the fp32 features in place as TF32"},
            "tc_gemm<EpiLinear> (projection)": {"bound": "tensor", "ms": k_ms["tc_gemm<EpiLinear> (projection)"],
                                                "achieved": gemm_flops / (k_ms["tc_gemm<EpiLinear> (projection)"] * 1e-3) / 1e12,
                                                "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                                                "note": "3 bf16 passes for the same algorithmic flops (fp32-accurate)" if acc else
                                                "kind::tf32 from fp32 operands, reported against the measured bf16 peak (the TF32 pipe runs at half of it)"},
            "tc_gemm<EpiLsmFwd> (pair)": {"bound": "tensor", "ms": k_ms["tc_gemm<EpiLsmFwd> (pair)"],
                                          "achieved": pair_flops / (k_ms["tc_gemm<EpiLsmFwd> (pair)"] * 1e-3) / 1e12,
                                          "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                                          "note": "3 bf16 passes for the same algorithmic flops (fp32-accurate)" if acc else "single bf16 pass"},
            "pair_ce": {"bound": "latency", "ms": k_ms["pair_ce"]},
        }
        for v in kern.values():
            if "peak" in v:
                v["frac"] = v["achieved"] / v["peak"]
        return kern

    kern = lsm_kernels(args.precision)
    kern_other = lsm_kernels(other)
    precisions[args.precision]["kernels_ms"] = {k: v["ms"] for k, v in kern.items()}
    precisions[other]["kernels_ms"] = {k: v["ms"] for k, v in kern_other.items()}
    precisions[other]["pair_kernel_frac_of_bf16_peak"] = kern_other["tc_gemm<EpiLsmFwd> (pair)"]["frac"]
    dom = max((k for k in kern if "peak" in kern[k]), key=lambda k: kern[k]["ms"])
    traffic = tensor_pct = None     # DRAM bytes per launch / tensor-pipe activity of the dominant kernel, from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            ent = json.load(fh).get(f"{dom} [{args.precision}]")
        if ent is not None:
            traffic = float(ent["bytes"])
            tensor_pct = ent.get("tensor_pipe_active_pct")
    except (OSError, ValueError, KeyError):
        traffic = None
    roofline = {"kernel": dom, "bound": kern[dom]["bound"], "achieved": kern[dom]["achieved"], "peak": kern[dom]["peak"],
                "unit": kern[dom]["unit"], "frac": kern[dom]["frac"], "traffic": traffic,
                "tensor_pipe_active_pct_ncu": tensor_pct,      # sm__pipe_tensor_cycles_active of the same kernel in the committed capture (executed, not algorithmic, work)
                "peak_source": f"{pk['source']} (MEASURED_PEAKS.json burst figure: kernel timed alone)",
                "algorithmic": "2*M*N*K of the un-collapsed GEMM (tensor) / bytes read+written once (hbm); the fp32-accurate mode runs 3 bf16 passes for the same algorithmic flops",
                "kernels": kern}

    # ---- end to end through the module with HOST buffers (headline precision) -----------------------------------
    # The call a user makes: batches of PINNED host tensors go through locov_b200.HostFeed (the package's host->device staging ring: the
    # H2D copy of batch i + 1 runs on a copy stream under the kernels of batch i) into the drop-in head; the losses, accuracies and both
    # pair matrices come back to pinned host memory and the host waits for them EVERY step.  All `steps` H2D copies and D2H reads are
    # inside the timed region (the first submit comes after the start event).
    import locov_b200
    feed = locov_b200.HostFeed(dev)
    res_h = torch.empty(8 + 2 * b_glob * b_glob, dtype=torch.float32).pin_memory()
    d2h = res_h.numel() * 4

    def e2e_run(nsteps):
        feed.submit(host_sets[0])
        for i in range(nsteps):
            if i + 1 < nsteps:
                feed.submit(host_sets[(i + 1) % args.sets])
            ii_d, ic_d = feed.take()
            with torch.no_grad():
                info, losses, dists = head(ii_d, ic_d)
            flat = torch.cat([torch.stack(list(losses.values())), torch.stack(list(info.values())), dists["w2r"].reshape(-1), dists["r2w"].reshape(-1)])
            res_h.copy_(flat, non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the caller reads the losses every step

    e2e_run(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.loco_launch_count()
    b0_ = feed.bytes_copied
    e0.record()
    e2e_run(args.e2e_steps)
    e1.record()
    torch.cuda.synchronize()
    e2e_launches = lib.loco_launch_count() - n0
    assert feed.bytes_copied - b0_ == h2d * args.e2e_steps, "every timed step must copy its own inputs from the host"
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.e2e_steps
    barrier()

    # ---- the other half of the metric: RoIAlign and RoI x class scoring ----------------------------------------
    workloads = {}
    if not args.no_workloads:
        workloads = roi_workloads(torch, dist, M, ops, synthetic, dev, pk, world, rank, max_over_ranks, barrier, with_cpu=(rank == 0 and world == 1 and not args.no_cpu_baseline))
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        ii, ic, w, b = make_inputs(SEED)
        cpu_reference_pass(ii, ic, w, b)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            cpu_reference_pass(ii, ic, w, b)
            best = min(best, time.perf_counter() - t0)
        cpu = {"value": (B_LOC * T * B_LOC * RG) / best, "unit": "scores/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "one config-2 batch (32 img x 100 regions x 20 tokens = 2,048,000 scores), 1 warm-up + best of 3 passes of oracle/lsm_head.grounding_head_forward_literal (torch CPU fp32)",
               "ms": best * 1e3}

    if rank == 0:
        out = {
            "metric": METRIC, "value": res_main["value"], "unit": "scores/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": res_main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": res_main["dtype"], "data": "synthetic", "impl": "b200",
            "config": {"workload": WORKLOAD, "images_per_gpu": B_LOC, "global_batch": b_glob, "regions": RG, "tokens": T, "v_dim": V,
                       "l_dim": D, "precision": args.precision, "parallelism": f"image-sharded x{world}" if world > 1 else "single GPU",
                       "l2": f"ring of {args.sets} input batches ({args.sets * h2d / 1e6:.0f} MB > 126 MB L2), CUDA-graph replay per batch",
                       "scores_per_step": scores_per_step, "per_gpu_value": res_main["value"] / world},
            "precisions": precisions, "roofline": roofline, "workloads": workloads, "cpu_baseline": cpu,
            "e2e": {"value": scores_per_step / (e2e_ms * 1e-3), "unit": "scores/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_ms, "steps": args.e2e_steps, "precision": args.precision,
                    "gpu_launches_per_step": e2e_launches / args.e2e_steps,
                    "api": "locov_b200.HostFeed (pinned host batches, H2D of batch i+1 under the kernels of batch i) -> GroundingHead.forward -> pinned host result, host sync every step"},
            "gpu_launches": int(res_main["gpu_launches_per_step"] * args.steps), "gpu_launches_per_step": res_main["gpu_launches_per_step"],
            "clocks": clocks,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        # NCCL collectives live inside the captured CUDA graphs: tear down without destroy_process_group (which can
        # block on them) once every rank is past the timed regions
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def roi_workloads(torch, dist, M, ops, synthetic, dev, pk, world, rank, max_over_ranks, barrier, with_cpu):
    """RoIAlign (HBM roofline) and the RoI x class box-predictor chain (tensor roofline) at the BASELINE config sizes.
    N = 1: configs[0] (2 x 512 RoIs vs 65 classes), configs[2] (16 x 512 RoIs vs 48 classes, fwd + bwd), configs[4] (8 x 1000 RoIs vs
    1203 classes).  N > 1: configs[4] sharded by image (8 / N images per rank, no collective; strong scaling, max over ranks)."""
    out = {}

    def predictor(K, precision, cls, we, wb, train=False):
        cfg = M.get_cfg("stt")
        cfg.MODEL.B200.PRECISION = precision
        bp = M.build_box_predictor(cfg, 2048).to(dev)
        with torch.no_grad():
            bp.emb_pred.weight.copy_(we); bp.bbox_pred.weight.copy_(wb)
        bp.set_class_embeddings(cls)
        return bp.train(train)

    def box_fwd(tag, R, K, nrot):
        x, we, be, wb, bb, cls, gt = synthetic.box_inputs(R, K, seed=SEED + 7)
        xs = [(x + 0.01 * i).to(dev) for i in range(nrot)]
        flops = 2.0 * R * 2048 * 772 + 2.0 * R * 768 * (K + 1)
        res = {"R": R, "K1": K + 1, "scores_per_call": R * (K + 1), "algorithmic_GFLOP": flops / 1e9, "bound": "tensor", "peak": pk["bf16_tflops"], "unit": "TFLOP/s"}
        for precision in ("fp32", "bf16"):
            bp = predictor(K, precision, cls, we, wb)

            def fn(i):
                with torch.no_grad():
                    s, d = bp(xs[i % nrot])
                    return bp.predict_probs((s, d), [range(R)])
            ms = max_over_ranks(graph_time(torch, fn, iters=max(8, nrot)))
            res[precision] = {"ms": ms, "scores/s": R * (K + 1) * (world if tag.endswith("sharded") else 1) / (ms * 1e-3), "achieved": flops / (ms * 1e-3) / 1e12,
                              "frac": flops / (ms * 1e-3) / 1e12 / pk["bf16_tflops"]}
        out[tag] = res
        return x, we, wb, cls

    if world > 1:
        n_img = max(1, 8 // world)
        box_fwd(f"box_cfg5_fwd_sharded", n_img * 1000, 1203, 4)
        out["box_cfg5_fwd_sharded"]["note"] = f"BASELINE configs[4] sharded by image: {n_img} images x 1000 RoIs per rank x {world} ranks, no collective; time = max over ranks"
        barrier()
        return out

    # ---- RoIAlign ------------------------------------------------------------------------------------------
    for tag, (C, ps, stride) in {"roi_align_cfg1_faithful": (1024, 14, 16), "roi_align_cfg1_literal": (2048, 7, 32)}.items():
        feats = [synthetic.res4_features(2, C=C, stride=stride, seed=SEED + i).to(dev) for i in range(2)]
        rois = synthetic.coco_boxes(2, 512, seed=SEED).to(dev)
        R = rois.shape[0]
        nbytes = R * C * ps * ps * 4 + feats[0].numel() * 4 + rois.numel() * 4
        ms = graph_time(torch, lambda i: ops.roi_align(feats[i % 2], rois, ps, 1.0 / stride), iters=8)
        out[tag] = {"shape": f"[2,{C},{feats[0].shape[2]},{feats[0].shape[3]}] x {R} RoIs -> [{R},{C},{ps},{ps}] fp32", "ms": ms, "algorithmic_MB": nbytes / 1e6,
                    "bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]}
        if tag == "roi_align_cfg1_faithful":
            # the same pooling written channels-last (what cuDNN's res5 wants), fp32 and bf16 — algorithmic bytes follow the output dtype
            for sub, dt, esz in (("channels_last_fp32", torch.float32, 4), ("channels_last_bf16", torch.bfloat16, 2)):
                nb = R * C * ps * ps * esz + feats[0].numel() * 4 + rois.numel() * 4
                m = graph_time(torch, lambda i: ops.roi_align(feats[i % 2], rois, ps, 1.0 / stride, channels_last=True, out_dtype=dt), iters=8)
                out[f"roi_align_cfg1_{sub}"] = {"ms": m, "algorithmic_MB": nb / 1e6, "bound": "hbm", "achieved": nb / (m * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                                                "unit": "GB/s", "frac": nb / (m * 1e-3) / 1e9 / pk["hbm_gbs"]}
            dout = torch.randn(R, C, ps, ps, device=dev)
            msb = graph_time(torch, lambda i: ops.roi_align_backward(dout, feats[0].shape, rois, 1.0 / stride), iters=4)
            nb = dout.numel() * 4 + 2 * feats[0].numel() * 4
            out["roi_align_cfg1_faithful_bwd"] = {"ms": msb, "algorithmic_MB": nb / 1e6, "bound": "hbm", "achieved": nb / (msb * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                                                  "unit": "GB/s", "frac": nb / (msb * 1e-3) / 1e9 / pk["hbm_gbs"]}
            if with_cpu:
                import torchvision
                fc, rc = feats[0].cpu(), rois.cpu()
                torchvision.ops.roi_align(fc, rc, ps, 1.0 / stride, 0, True)
                t0 = time.perf_counter()
                torchvision.ops.roi_align(fc, rc, ps, 1.0 / stride, 0, True)
                out[tag]["cpu_baseline"] = {"ms": (time.perf_counter() - t0) * 1e3, "kind": "reference", "cores": torch.get_num_threads(),
                                            "sample": "one call of torchvision.ops.roi_align (the compiled CPU op the reference reaches) on the same inputs"}
            del dout
        del feats
    torch.cuda.empty_cache()

    # ---- spatial mean of the res5 output (+ the projection's bf16 operand), configs[1] and configs[4] row counts ----------------
    for tag, R, cl, dt in (("spatial_mean_cfg1_nchw_fp32", 1024, False, torch.float32), ("spatial_mean_cfg5_nchw_fp32", 8000, False, torch.float32),
                           ("spatial_mean_cfg5_channels_last_bf16", 8000, True, torch.bfloat16)):
        xs_ = [torch.randn(R, 2048, 7, 7, device=dev, dtype=dt) for _ in range(2)]
        if cl:
            xs_ = [t.contiguous(memory_format=torch.channels_last) for t in xs_]
        nb = xs_[0].numel() * xs_[0].element_size() + R * 2048 * (4 + 2 + 2)
        m = graph_time(torch, lambda i: ops.spatial_mean(xs_[i % 2], True), iters=8)
        out[tag] = {"shape": f"[{R},2048,7,7] -> [{R},2048] fp32 + bf16 (hi, lo) operand", "ms": m, "algorithmic_MB": nb / 1e6, "bound": "hbm",
                    "achieved": nb / (m * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": nb / (m * 1e-3) / 1e9 / pk["hbm_gbs"]}
        del xs_
    torch.cuda.empty_cache()

    # ---- RoI x class scoring (forward: scores + probabilities) ----------------------------------------------------
    x1, we1, wb1, cls1 = box_fwd("box_cfg1_fwd", 1024, 65, 20)
    if with_cpu:
        import torch.nn.functional as F
        xc = x1
        with torch.no_grad():
            for _ in range(2):
                t0 = time.perf_counter()
                F.softmax(F.linear(F.linear(xc, we1), cls1), -1); F.linear(xc, wb1)
                dt = time.perf_counter() - t0
        out["box_cfg1_fwd"]["cpu_baseline"] = {"ms": dt * 1e3, "kind": "port", "cores": torch.get_num_threads(),
                                               "sample": "one pass of the restated predictor (3 F.linear + softmax, torch CPU fp32) on the same inputs"}
    x5, we5, wb5, cls5 = box_fwd("box_cfg5_fwd", 8000, 1203, 3)

    # ---- configs[4] inference tail: box decoding + threshold + per-class NMS + top-k for 8 images x 1000 RoIs x 1203 classes ---------
    from locov_b200.modeling.box_emb_head import fast_rcnn_inference_single_image
    bp = predictor(1203, "fp32", cls5 * 8.0, we5, wb5)          # (class matrix scaled so that the softmax is peaked, as trained embeddings are)
    pb = synthetic.coco_boxes(8, 1000, seed=SEED + 3)[:, 1:]
    pb[:, 2:] = torch.maximum(pb[:, 2:], pb[:, :2] + 8)
    props = [M.Instances((800, 1216), proposal_boxes=M.Boxes(pb[i * 1000:(i + 1) * 1000].to(dev))) for i in range(8)]
    with torch.no_grad():
        pred = bp(x5.to(dev))
        probs = bp.predict_probs(pred, props)
        boxes = bp.predict_boxes(pred, props)
    tail = {}
    for name, (thr, topk) in {"coco_settings": (0.05, 100), "lvis_settings": (1e-4, 300)}.items():
        bp.test_score_thresh, bp.test_topk_per_image = thr, topk

        def ours():
            with torch.no_grad():
                return bp.inference(pred, props)

        def stock():
            return [fast_rcnn_inference_single_image(b, s, (800, 1216), thr, bp.test_nms_thresh, topk) for b, s in zip(boxes, probs)]
        res = {}
        for tag, fn, reps in (("ms", ours, 10), ("torchvision_cuda_ms", stock, 2)):
            fn()
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                out_ = fn()
            b_.record()
            torch.cuda.synchronize()
            res[tag] = a.elapsed_time(b_) / reps
        res["detections"] = int(sum(len(r) for r in ours()[0]))
        res["candidates"] = int(sum(int((p[:, :-1] > thr).sum()) for p in probs))
        tail[name] = res
    out["box_inference_cfg5"] = {"shape": "8 images x 1000 RoIs x 1203 classes, NMS 0.5", **tail,
                                 "note": "ms = EmbeddingFastRCNNOutputLayers.inference (4 launches of loco_box_inference + the device->host copy of the 8 detection counts); "
                                         "torchvision_cuda_ms = the same tail per image with torch.nonzero + torchvision.ops.batched_nms on the same device (what the reference reaches)"}
    del bp, pred, probs, boxes

    # ---- configs[2]: training step of the box head, fwd + bwd, weights frozen as in coco_stt.yaml:36 ---------------------
    R, K = 8192, 48
    x, we, be, wb, bb, cls, gt = synthetic.box_inputs(R, K, seed=SEED + 9)
    boxes = synthetic.coco_boxes(16, 512, seed=SEED + 1)[:, 1:]
    boxes[:, 2:] = torch.maximum(boxes[:, 2:], boxes[:, :2] + 8)
    gtb = boxes + torch.randn(boxes.shape, generator=torch.Generator().manual_seed(3)) * 2
    gtb[:, 2:] = torch.maximum(gtb[:, 2:], gtb[:, :2] + 4)
    props = [M.Instances((800, 1216), proposal_boxes=M.Boxes(boxes[i * 512:(i + 1) * 512].to(dev)), gt_boxes=M.Boxes(gtb[i * 512:(i + 1) * 512].to(dev)),
                         gt_classes=gt[i * 512:(i + 1) * 512].to(dev)) for i in range(16)]
    xs = [(x + 0.01 * i).to(dev).requires_grad_(True) for i in range(3)]
    flops = 2.0 * R * 2048 * 772 + 2.0 * R * 768 * (K + 1) + 2.0 * R * (K + 1) * 768 + 2.0 * R * 772 * 2048 + 2.0 * R * 4 * 2048
    res = {"R": R, "K1": K + 1, "scores_per_call": R * (K + 1), "algorithmic_GFLOP": flops / 1e9, "bound": "tensor", "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
           "note": "forward + losses + backward (d/dx, d/d bbox_pred); emb_pred and the class matrix frozen (coco_stt.yaml:36)"}
    for precision in ("bf16", "fp32"):
        bp = predictor(K, precision, cls, we, wb, train=True)

        def step(i):
            xi = xs[i % 3]
            pred = bp(xi)
            l = bp.losses(pred, props)
            (l["loss_cls"] + l["loss_box_reg"]).backward()
            xi.grad = None
            bp.bbox_pred.weight.grad = None
            bp.bbox_pred.bias.grad = None
        try:
            ms, how = graph_time(torch, step, iters=6), "CUDA-graph replay"
        except Exception as e:      # noqa: BLE001 (a step that cannot be captured is timed eagerly and says so)
            torch.cuda.synchronize()
            for i in range(3):
                step(i)
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(10):
                step(i)
            b_.record()
            torch.cuda.synchronize()
            ms, how = a.elapsed_time(b_) / 10, f"eager (graph capture failed: {str(e)[:80]})"
        res[precision] = {"ms": ms, "timing": how, "scores/s": R * (K + 1) / (ms * 1e-3), "achieved": flops / (ms * 1e-3) / 1e12, "frac": flops / (ms * 1e-3) / 1e12 / pk["bf16_tflops"]}
    out["box_cfg3_fwd_bwd"] = res
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
