#!/usr/bin/env python
"""Benchmark of the hot path: bi-temporal pairs/sec, X3D-L BCD train step (fwd + BCEDiceLoss + bwd + Adam),
synthetic LEVIR-shape 256x256, batch 32 per GPU (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W          # our arm (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle) on host cores

Prints ONE JSON line on rank 0.  `value` = whole-job pairs/s with the inputs resident in HBM; `e2e` = the same
step through the public API (change3d_b200.input_pipeline.DevicePrefetcher feeding change3d_b200.train_step.BCDTrainStep)
with pinned HOST inputs: every step's H2D copy (on a copy stream, one batch ahead) and a D2H read of its loss are inside
the timed region (--e2e-mode plain: copies on the compute stream; both: measure both).  `roofline` describes the kernel family that takes the most device
time, from CUDA events recorded around every launch of a separate profiled step; `cpu_baseline` times the
oracle (a port of the reference algorithm, oracle/change3d_oracle.py) on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bi-temporal pairs/sec (X3D-L BCD train step, 256x256)"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (reference --batch_size is per process)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of a captured CUDA graph")
    ap.add_argument("--cpu-batch", type=int, default=2, help="batch of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--e2e-mode", default="prefetch", choices=["plain", "prefetch", "both"],
                    help="how the e2e loop stages its pinned host inputs (both: measure both, report the better)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-roofline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on host cores
# ------------------------------------------------------------------------------------------------
def cpu_train_steps(batch: int, size: int, steps: int, warmup: int):
    """BCD train step exactly as scripts/train_BCD.py:179-216 (update_bcd, BCEDiceLoss, backward, Adam) with the
    oracle's functional model.  Returns (pairs/s, threads, seconds per step)."""
    import torch
    from oracle import change3d_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.synth_state_dict(O.trainer_schema("bcd", 1, size, size, 1), 16)
    sd = O.clone_sd(sd, requires_grad=True)
    params = [v for k, v in sd.items() if v.requires_grad]
    opt = torch.optim.Adam(params, 2e-4, (0.9, 0.99), eps=1e-8, weight_decay=1e-4)
    pre, post, target = O.synth_inputs(batch, size, size, 16)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = O.bce_dice_loss(O.trainer_forward(sd, "bcd", pre, post, True), target)
        loss.backward()
        opt.step()
        float(loss)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, torch.get_num_threads(), sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))          # bounded: each step is ~seconds of CPU work
    warm = 1
    t_start = time.perf_counter()
    v, threads, sec = cpu_train_steps(args.cpu_batch, args.size, steps, warm)
    sample = (f"BCD train step (fwd+BCEDice+bwd+Adam) batch {args.cpu_batch} at {args.size}x{args.size}, "
              f"{steps} timed steps after {warm} warm-up, oracle port of the reference on torch CPU ops")
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BCD X3D-L train, synthetic {args.size}x{args.size}, CPU sample batch {args.cpu_batch}",
                       "note": "reference CPU path = oracle port (reference needs pytorchvideo/fvcore, absent offline)"},
            "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(time.perf_counter() - t_start, 1)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the change3d_b200 engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import argparse as _ap
    import contextlib
    import io
    from change3d_b200 import _lib, ops
    from change3d_b200.model.trainer import Trainer
    from change3d_b200.train_step import BCDTrainStep
    _lib.load()

    B, S = args.batch, args.size
    margs = _ap.Namespace(num_perception_frame=1, num_class=1, in_height=S, in_width=S, dataset="LEVIR-CD",
                          pretrained="/nonexistent/X3D_L.pyth")
    torch.manual_seed(16)                                   # scripts/train_BCD.py:253 (identical init on every rank)
    with contextlib.redirect_stdout(io.StringIO()):
        model = Trainer(margs).to(dev).float()
    step = BCDTrainStep(model, lr=2e-4, use_graph=not args.no_graph)

    g = torch.Generator(device="cpu").manual_seed(16 + rank)
    h_pre = torch.randn(B, 3, S, S, generator=g).pin_memory()
    h_post = torch.randn(B, 3, S, S, generator=g).pin_memory()
    h_tgt = (torch.rand(B, 1, S, S, generator=g) < 0.05).float().pin_memory()
    pre, post, tgt = h_pre.to(dev), h_post.to(dev), h_tgt.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (>= 3) ----
    W = max(3, args.warmup)
    for _ in range(W):
        loss = step(pre, post, tgt)
    barrier()
    launches_before = _lib.LAUNCHES[0]
    if step.graph is None:
        pass

    # ---- timed region: K steps, device resident inputs ----
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(pre, post, tgt)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    loss_val = float(loss.item())
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)

    # launches per step: eager = counted; graph = launches recorded while capturing (one replay = same kernels)
    if args.no_graph:
        launches_per_step = (_lib.LAUNCHES[0] - launches_before) // max(1, args.steps)
    else:
        launches_per_step = getattr(step, "captured_launches", None)

    # ---- e2e: public API with pinned host buffers; every step's H2D copy + the step + the D2H read of its loss are
    # inside the timed region.  --e2e-mode plain: copies on the compute stream right before the step (what the
    # reference loop does); prefetch: through the package's DevicePrefetcher (copy stream, one batch ahead).
    from change3d_b200.input_pipeline import DevicePrefetcher
    k2 = max(3, args.steps)

    def host_batches(n):
        for _ in range(n):
            yield (h_pre, h_post, h_tgt)

    def e2e_plain(n):
        for _ in range(n):
            d_pre.copy_(h_pre, non_blocking=True); d_post.copy_(h_post, non_blocking=True); d_tgt.copy_(h_tgt, non_blocking=True)
            float(step(d_pre, d_post, d_tgt).item())

    def e2e_prefetch(n):
        pf = DevicePrefetcher(host_batches(n), dev)
        for a_, b_, c_ in pf:
            float(step(a_, b_, c_).item())
        assert pf.bytes_staged == n * (h_pre.numel() + h_post.numel() + h_tgt.numel()) * 4

    d_pre, d_post, d_tgt = torch.empty_like(pre), torch.empty_like(post), torch.empty_like(tgt)
    e2e_runs = {}
    for mode in (("plain", "prefetch") if args.e2e_mode == "both" else (args.e2e_mode,)):
        fn = e2e_plain if mode == "plain" else e2e_prefetch
        fn(2)
        barrier()
        t0 = time.perf_counter()
        fn(k2)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_runs[mode] = world * B * k2 / float(te.item())
    e2e_mode = args.e2e_mode if args.e2e_mode != "both" else max(e2e_runs, key=e2e_runs.get)
    e2e_value = e2e_runs[e2e_mode]
    h2d = (h_pre.numel() + h_post.numel() + h_tgt.numel()) * 4
    # training confusion matrix hist[target][output > 0.5], accumulated on the device by the loss kernel over every
    # step since the graph was captured (the reference copies a 16.8 MB mask to the host per step for this); one
    # 32-byte read-back here, outside both timed regions
    cm_host = step.cm.cpu().tolist()

    # ---- roofline probe: one eager step with CUDA events around every launch ----
    roofline = None
    families = {}
    if not args.skip_roofline and rank == 0:
        # rank-local: the probe must not enter a collective (only rank 0 runs it), so it sequences the iteration
        # and the Adam step itself instead of calling the DDP-aware BCDTrainStep.__call__
        def probe():
            step._iteration(pre, post, tgt)
            step.opt.step()
        # per-launch durations are only meaningful for kernels that run alone: the probe keeps the weight-gradient
        # GEMMs on the main stream (the timed steps above overlap them with the dgrad chain on a side stream)
        side_prev = os.environ.get("C3D_SIDE_STREAM")
        os.environ["C3D_SIDE_STREAM"] = "0"
        probe()                                   # eager warm-up
        torch.cuda.synchronize()
        ops.PROF = {}
        for _ in range(2):
            probe()
        torch.cuda.synchronize()
        if side_prev is None:
            os.environ.pop("C3D_SIDE_STREAM", None)
        else:
            os.environ["C3D_SIDE_STREAM"] = side_prev
        prof, ops.PROF = ops.PROF, None
        tot_ms = 0.0
        for fam, recs in prof.items():
            t_ms = sum(a.elapsed_time(b) for a, b, _ in recs)
            nb = sum(n for _, _, n in recs)
            families[fam] = {"ms_per_step": round(t_ms / 2, 3), "launches_per_step": len(recs) // 2,
                             "algorithmic_GB_per_step": round(nb / 2 / 1e9, 3),
                             "achieved_GBs": round(nb / (t_ms / 1e3) / 1e9, 1) if t_ms > 0 else None}
            tot_ms += t_ms / 2
        top = max(families, key=lambda k: families[k]["ms_per_step"])
        peak, peak_src = peaks()
        ach = families[top]["achieved_GBs"]
        # measured DRAM traffic per launch of this kernel family, from the committed ncu pass (never measured here:
        # a number taken under a profiler is not a bench value, the traffic of a launch does not depend on timing)
        traffic, traffic_src = None, None
        for tname in ("r01b_traffic.json", "r01_traffic.json"):          # latest committed launch list first
            tpath = os.path.join(ROOT, "profiles", tname)
            if not os.path.isfile(tpath):
                continue
            with open(tpath) as f:
                tj = json.load(f)
            if top in tj.get("families", {}):
                traffic = tj["families"][top]["dram_bytes_per_launch"]
                traffic_src = f"profiles/{tname} (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch)"
                break
        alg_per_launch = int(families[top]["algorithmic_GB_per_step"] * 1e9 / max(1, families[top]["launches_per_step"]))
        roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": alg_per_launch, "peak_source": peak_src,
                    "share_of_step": round(families[top]["ms_per_step"] / max(tot_ms, 1e-9), 3),
                    "families": families}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on host cores ----
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        v, threads, sec = cpu_train_steps(args.cpu_batch, S, args.cpu_steps, 1)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"BCD train step batch {args.cpu_batch} at {S}x{S}, {args.cpu_steps} timed steps after 1 warm-up "
                         f"({sec:.2f} s/step), oracle port on torch CPU ops"}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": W, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"BCD X3D-L train (fwd+BCEDiceLoss+bwd+Adam), synthetic LEVIR-shape {S}x{S}, "
                                       f"batch {B}/GPU, T=3", "global_batch": world * B, "parallelism": f"dp{world}",
                           "gemm_arithmetic": "3xTF32 split products on tcgen05, fp32 accumulate (fp32-class)",
                           "cuda_graph": not args.no_graph, "wgrad_side_stream": os.environ.get("C3D_SIDE_STREAM", "1") == "1",
                           "l2": "per-step working set (tens of GB of activations) >> 126 MB L2; no flush needed"},
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "steps": k2, "input_staging": e2e_mode,
                        "all_modes": {k: round(v, 2) for k, v in e2e_runs.items()}},
                "gpu_launches": launches_per_step, "loss": round(loss_val, 5), "clocks": sampler.summary(),
                "metrics_on_device": {"confusion_matrix": cm_host, "pixels": int(sum(map(sum, cm_host))),
                                      "note": "hist[target][output > 0.5] accumulated by the loss kernel, no per-step D2H"}}
        if roofline is not None:
            line["roofline"] = roofline
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier()                                 # the other ranks wait here while rank 0 runs its probe
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: self-launch one rank per GPU (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
