#!/usr/bin/env python
"""bench.py — region-text scores/sec of the LSM grounding head (BASELINE.json configs[1]):
32 images x 100 regions x 20 caption tokens per GPU, 2048 -> 768 projection, full image-caption pair
matrix, both alignments, 4 CE losses + 4 accuracies.  One "step" = one forward pass of the head over
one batch of synthetic inputs (randn features, random 768-d BERT-shaped caption embeddings).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (the product)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU arithmetic (oracle port)

N > 1 (torchrun): each rank keeps 32 images; the pair matrix becomes the global [32N x 32N] matrix,
sharded by image with the caption operands all-gathered over NCCL/NVLink (locov_b200/parallel.py) —
weak scaling in images per GPU; N = 8 is BASELINE.json configs[3] (global batch 256).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA-graph replay of the step,
inputs already in HBM, a ring of input sets larger than L2 so no step re-reads cached inputs);
`e2e` = the same step called through the GroundingHead module with pinned HOST buffers, H2D / D2H
copies inside the timed region; `roofline` = the dominant kernel timed alone with CUDA events;
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
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--sets", type=int, default=6, help="ring of distinct input batches (6 x 28 MB > 126 MB L2)")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def make_inputs(seed):
    from oracle import lsm_head   # input generator only (SURVEY.md §8d shapes); nothing of the oracle is timed here
    ii, ic, w, b = lsm_head.make_lsm_inputs(B=B_LOC, Rg=RG, T=T, V=V, D=D, seed=seed, ragged_regions=True)
    return ii, ic, w, b


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own arithmetic on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(ii, ic, w, b):
    import torch
    from oracle import lsm_head
    with torch.no_grad():
        return lsm_head.grounding_head_forward_literal(ii, ic, w, b, 10.0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    ii, ic, w, b = make_inputs(SEED)
    for _ in range(max(1, min(args.warmup, 3))):
        cpu_reference_pass(ii, ic, w, b)
    steps = args.steps
    t0 = time.perf_counter()
    cpu_reference_pass(ii, ic, w, b)
    one = time.perf_counter() - t0
    budget = 240.0
    timed = max(1, min(steps, int(budget / max(one, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(timed):
        cpu_reference_pass(ii, ic, w, b)
    dt = (time.perf_counter() - t0) / timed
    scores = B_LOC * T * B_LOC * RG
    val = scores / dt
    cores = torch.get_num_threads()
    sample = f"one config-2 batch per step (32x100x20, {scores} scores); {timed} of the {steps} requested steps timed (240 s cap)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "scores/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "device": "host CPU"},
        "cpu_baseline": {"value": val, "unit": "scores/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import locov_b200.modeling as M
    from locov_b200 import _lib, ops, parallel
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

    # ---- model + inputs ---------------------------------------------------------------------------
    cfg = M.get_cfg("lsm")
    cfg.MODEL.B200.PRECISION = args.precision
    head = M.GroundingHead(cfg, V, D).to(dev)
    host_sets, dev_sets = [], []
    for s in range(args.sets):
        ii, ic, w, b = make_inputs(SEED + rank + 1000 * s)
        hs = ({k: v.pin_memory() for k, v in ii.items()}, {k: v.pin_memory() for k, v in ic.items()})
        host_sets.append(hs)
        dev_sets.append(({k: v.to(dev) for k, v in hs[0].items()}, {k: v.to(dev) for k, v in hs[1].items()}))
    _, _, w0, b0 = make_inputs(SEED)        # identical weights on every rank
    with torch.no_grad():
        head.v2l_projection.weight.copy_(w0)
        head.v2l_projection.bias.copy_(b0)
    if world > 1:
        parallel.shard_grounding_head(head)
    b_glob = B_LOC * world
    scores_per_step = b_glob * T * b_glob * RG            # all ranks together: [B*T] x [B*Rg]
    torch.cuda.synchronize()

    def step_eager(s):
        with torch.no_grad():
            return head(dev_sets[s][0], dev_sets[s][1])

    # eager warm-up (library caches, NCCL communicator), then one CUDA graph per input set.  The bf16
    # weight shadow is dropped before each capture so that every step re-converts the projection weights
    # (a training step sees new weights every iteration).
    for s in range(args.sets):
        step_eager(s)
    barrier()
    graphs, outs, launches_per_step = [], [], None
    for s in range(args.sets):
        LF.clear_weight_cache()
        g = torch.cuda.CUDAGraph()
        n0 = lib.loco_launch_count()
        with torch.cuda.graph(g):
            o = step_eager(s)
        launches_per_step = lib.loco_launch_count() - n0
        graphs.append(g)
        outs.append(o)
    barrier()

    # ---- timed region: device-resident throughput ---------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(max(args.warmup, 3)):
        graphs[i % args.sets].replay()
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        graphs[i % args.sets].replay()
    e1.record()
    torch.cuda.synchronize()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    barrier()
    ms_per_step = ms_total / args.steps
    value = scores_per_step / (ms_per_step * 1e-3)

    # ---- per-kernel breakdown + roofline (each kernel timed alone, CUDA events, rotating operands > L2) ---
    acc = args.precision == "fp32"
    nrot = max(args.sets, 12)
    feats = [torch.randn(B_LOC * RG, V, device=dev) for _ in range(nrot)]
    x_ops = [ops.split_bf16(f, acc) for f in feats]
    w_op = ops.split_bf16(head.v2l_projection.weight.detach(), acc)
    mask_c = torch.ones(b_glob, T, device=dev)
    mask_r = torch.ones(B_LOC, RG, device=dev)
    cap_ops = [ops.split_bf16(torch.randn(b_glob * T, D, device=dev) * 0.05, acc) for _ in range(4)]
    emb_ops = [ops.linear_fwd(x_ops[i], w_op, head.v2l_projection.bias.detach(), want_f32=False, n_bf16=D, accurate_out=acc)[1] for i in range(4)]
    pw = torch.randn(2, b_glob, b_glob, device=dev)

    def time_kernel(fn, iters=24, reps=5):
        """Average device time of ONE launch: `iters` launches (rotating operands) captured in a CUDA graph so that
        host launch overhead (longer than these kernels) is not in the number; best of `reps` replays."""
        for i in range(3):
            fn(i)
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

    hi_buf = x_ops[0]
    k_ms = {
        "split_bf16(features)": time_kernel(lambda i: ops.split_bf16(feats[i % nrot], acc, out=hi_buf)) if acc else 0.0,
        "tc_gemm<EpiLinear> (projection)": time_kernel(
            (lambda i: ops.linear_fwd(x_ops[i % nrot], w_op, None, want_f32=False, n_bf16=D, accurate_out=acc)) if acc else
            (lambda i: ops.linear_tf32_fwd(feats[i % nrot], head.v2l_projection.weight.detach(), None, want_f32=False, n_bf16=D))),
        "tc_gemm<EpiLsm> (pair)": time_kernel(lambda i: ops.lsm_pair(cap_ops[i % 4], mask_c, emb_ops[i % 4], mask_r, 0.1)),
        "pair_ce": time_kernel(lambda i: ops.pair_ce(pw, mask_c, mask_r)),
    }
    passes = 3 if acc else 1
    m_rows = B_LOC * RG
    gemm_flops = 2.0 * m_rows * V * D
    pair_flops = 2.0 * (b_glob * T) * (B_LOC * RG) * D               # one similarity GEMM serves both alignments
    split_bytes = m_rows * V * (4 + 2 * (2 if acc else 1))
    kern = {
        "split_bf16(features)": {"bound": "hbm", "ms": k_ms["split_bf16(features)"],
                                 "achieved": split_bytes / (max(k_ms["split_bf16(features)"], 1e-9) * 1e-3) / 1e9 if acc else 0.0,
                                 "peak": pk["hbm_gbs"], "unit": "GB/s", "note": "fp32 mode only; the reduced-precision mode multiplies the fp32 features in place as TF32"},
        "tc_gemm<EpiLinear> (projection)": {"bound": "tensor", "ms": k_ms["tc_gemm<EpiLinear> (projection)"],
                                            "achieved": gemm_flops / (k_ms["tc_gemm<EpiLinear> (projection)"] * 1e-3) / 1e12,
                                            "peak": pk["bf16_tflops"] * (1.0 if acc else 0.5), "unit": "TFLOP/s",
                                            "note": "bf16 x3 (fp32-accurate)" if acc else "kind::tf32 from fp32 operands: peak = half the measured bf16 peak"},
        "tc_gemm<EpiLsm> (pair)": {"bound": "tensor", "ms": k_ms["tc_gemm<EpiLsm> (pair)"],
                                            "achieved": pair_flops / (k_ms["tc_gemm<EpiLsm> (pair)"] * 1e-3) / 1e12,
                                            "peak": pk["bf16_tflops"], "unit": "TFLOP/s"},
        "pair_ce": {"bound": "latency", "ms": k_ms["pair_ce"]},
    }
    for v in kern.values():
        if "peak" in v:
            v["frac"] = v["achieved"] / v["peak"]
    dom = max((k for k in kern if "peak" in kern[k]), key=lambda k: kern[k]["ms"])
    traffic = None          # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            ent = json.load(fh).get(dom)
        if ent is not None and not acc:      # captured for the default (reduced-precision) mode only
            traffic = float(ent["bytes"])
    except (OSError, ValueError, KeyError):
        traffic = None
    roofline = {"kernel": dom, "bound": kern[dom]["bound"], "achieved": kern[dom]["achieved"], "peak": kern[dom]["peak"],
                "unit": kern[dom]["unit"], "frac": kern[dom]["frac"], "traffic": traffic,
                "peak_source": f"{pk['source']} (MEASURED_PEAKS.json burst figure: kernel timed alone)",
                "algorithmic": "2*M*N*K of the un-collapsed GEMM (tensor) / bytes read+written once (hbm); fp32 mode runs 3 bf16 passes for the same algorithmic flops",
                "kernels": kern}

    # ---- end to end through the module with HOST buffers -------------------------------------------
    ii_d = {k: torch.empty_like(v, device=dev) for k, v in host_sets[0][0].items()}
    ic_d = {k: torch.empty_like(v, device=dev) for k, v in host_sets[0][1].items()}
    res_h = torch.empty(8 + 2 * b_glob * b_glob, dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for d_ in host_sets[0] for v in d_.values())
    d2h = res_h.numel() * 4

    def e2e_step(s):
        hi_, hc_ = host_sets[s]
        for k in ii_d:
            ii_d[k].copy_(hi_[k], non_blocking=True)
        for k in ic_d:
            ic_d[k].copy_(hc_[k], non_blocking=True)
        with torch.no_grad():
            info, losses, dists = head(ii_d, ic_d)
        flat = torch.cat([torch.stack(list(losses.values())), torch.stack(list(info.values())), dists["w2r"].reshape(-1), dists["r2w"].reshape(-1)])
        res_h.copy_(flat, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller reads the losses every step

    for i in range(3):
        e2e_step(i % args.sets)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.loco_launch_count()
    e0.record()
    for i in range(args.e2e_steps):
        e2e_step(i % args.sets)
    e1.record()
    torch.cuda.synchronize()
    e2e_launches = lib.loco_launch_count() - n0
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.e2e_steps
    barrier()
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
            "metric": METRIC, "value": value, "unit": "scores/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (fp32-accurate split)", "data": "synthetic", "impl": "b200",
            "config": {"workload": WORKLOAD, "images_per_gpu": B_LOC, "global_batch": b_glob, "regions": RG, "tokens": T, "v_dim": V,
                       "l_dim": D, "precision": args.precision, "parallelism": f"image-sharded x{world}" if world > 1 else "single GPU",
                       "l2": f"ring of {args.sets} input batches ({args.sets * h2d / 1e6:.0f} MB > 126 MB L2), CUDA-graph replay per batch",
                       "scores_per_step": scores_per_step},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": scores_per_step / (e2e_ms * 1e-3), "unit": "scores/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_ms, "steps": args.e2e_steps,
                    "gpu_launches_per_step": e2e_launches / args.e2e_steps},
            "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
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


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
