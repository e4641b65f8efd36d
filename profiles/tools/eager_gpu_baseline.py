"""Context number for the roofline discussion ("also time the reference eager GPU path"): the oracle's
functional restatement of the reference model (oracle/change3d_oracle.py — torch ops, i.e. cuDNN / ATen kernels in NCDHW)
run on the GPU for the same BCD train step that bench.py times.  PROFILING TOOL, not part of bench.py or the product:

    python profiles/tools/eager_gpu_baseline.py [--batch 32] [--size 256] [--steps 5] [--warmup 2] [--tf32]

Prints one JSON line (pairs/s, ms/step, peak memory).  Measured in round 2 on a B200 of this pool
(profiles/r02_eager_gpu_baseline.json): batch 32 fp32 472.4 ms/step = 67.7 pairs/s (50 GB peak), batch 16 62.8, batch 8 55.9;
batch 16 with TF32 allowed 93.2 pairs/s.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import change3d_oracle as O  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--tf32", action="store_true", help="allow TF32 in cuDNN / cuBLAS (default: fp32, like the reference)")
    ap.add_argument("--device", default="cuda:0", help="cpu runs the same loop with a wall clock (logic check only)")
    a = ap.parse_args()
    dev = torch.device(a.device)
    gpu = dev.type == "cuda"
    torch.backends.cudnn.benchmark = True                       # scripts/train_BCD.py:250-251
    torch.backends.cudnn.allow_tf32 = a.tf32
    torch.backends.cuda.matmul.allow_tf32 = a.tf32
    sd = O.synth_state_dict(O.trainer_schema("bcd", 1, a.size, a.size, 1), 16)
    sd = {k: v.to(dev) for k, v in O.clone_sd(sd).items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, 2e-4, (0.9, 0.99), eps=1e-8, weight_decay=1e-4)
    pre, post, target = (t.to(dev) for t in O.synth_inputs(a.batch, a.size, a.size, 16))
    import time
    if gpu:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = 0.0
    for it in range(a.warmup + a.steps):
        if it == a.warmup:
            if gpu:
                torch.cuda.synchronize()
                e0.record()
            t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = O.bce_dice_loss(O.trainer_forward(sd, "bcd", pre, post, True), target)
        loss.backward()
        opt.step()
    if gpu:
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
    else:
        ms = (time.perf_counter() - t0) * 1e3 / a.steps
    print(json.dumps({"what": "oracle (torch eager ops) BCD train step on the GPU", "batch": a.batch, "size": a.size,
                      "tf32": a.tf32, "ms_per_step": round(ms, 2), "pairs_per_s": round(a.batch / ms * 1e3, 1),
                      "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1) if gpu else None, "loss": round(loss.item(), 5)}))


if __name__ == "__main__":
    main()
