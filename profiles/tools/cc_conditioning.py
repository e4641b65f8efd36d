"""Conditioning of the change-captioning feature-path gradients (tests/test_gpu_cc_encoder.py) — CPU, oracle only.

Question: are the 3e-2 .. 2e-1 gradient differences between an fp32 implementation and the fp64
oracle on the `cc_enc_b2_32` case a defect of a kernel, or the conditioning of the test problem itself (55 residual
blocks in train-mode BatchNorm; res5 runs at 2x2 pixels, so every BatchNorm there normalises over B*T*H*W = 24
samples and one ReLU-mask flip moves a weight gradient by O(1/24))?

Method: stay entirely in fp64 (no rounding error to speak of) and perturb the INPUT IMAGES by a relative 2^-24
(one fp32 ulp: the least any fp32 implementation can differ by after its first operation).  The change of each
gradient between the two fp64 runs is the problem's own amplification of fp32-sized noise; an fp32 implementation
cannot be expected to land closer to the fp64 truth than that.  Several noise seeds show the spread (flips are a
lottery: a tensor is either hit by one or not).

    python profiles/tools/cc_conditioning.py [--size 32] [--batch 2] [--seeds 6] [--eps 5.96e-8]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import change3d_oracle as O                      # noqa: E402
from oracle.make_golden_cc import GRAD_KEYS, SEED, weights   # noqa: E402


def grads(full, pre, post, w, dtype):
    s = O.clone_sd(full, dtype=dtype, requires_grad=True)
    out = O.encoder_forward(s, pre.to(dtype), post.to(dtype), 1, True, output_final=True)
    (out * w.to(dtype)).sum().backward()
    return out.detach().double(), {k: s["encoder." + k].grad.detach().double() for k in GRAD_KEYS}


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--seeds", type=int, default=6)
    ap.add_argument("--eps", type=float, default=2.0 ** -24)
    a = ap.parse_args()
    H = W = a.size
    full = O.synth_state_dict(O.trainer_schema("bcd", 1, H, W, 1), SEED)
    pre, post, _ = O.synth_inputs(a.batch, H, W, SEED)
    w = weights(SEED) if (H == 32 and a.batch == 2) else torch.randn(a.batch, 192, H // 16, W // 16,
                                                                   generator=torch.Generator().manual_seed(SEED))
    o64, g64 = grads(full, pre, post, w, torch.float64)
    o32, g32 = grads(full, pre, post, w, torch.float32)
    rows = {k: {"torch_fp32_vs_fp64": rel(g32[k], g64[k]), "fp64_perturbed_vs_fp64": []} for k in GRAD_KEYS}
    fwd = {"torch_fp32_vs_fp64": rel(o32, o64), "fp64_perturbed_vs_fp64": []}
    for sd in range(a.seeds):
        g = torch.Generator().manual_seed(1000 + sd)
        p2 = pre.double() * (1 + a.eps * torch.randn(pre.shape, generator=g, dtype=torch.float64))
        q2 = post.double() * (1 + a.eps * torch.randn(post.shape, generator=g, dtype=torch.float64))
        o, gp = grads(full, p2, q2, w, torch.float64)
        fwd["fp64_perturbed_vs_fp64"].append(rel(o, o64))
        for k in GRAD_KEYS:
            rows[k]["fp64_perturbed_vs_fp64"].append(rel(gp[k], g64[k]))
    print(json.dumps({"case": f"cc encoder B{a.batch} {H}x{W} seed {SEED}", "input_perturbation_rel": a.eps,
                      "forward": fwd, "grads": rows}, indent=1))


if __name__ == "__main__":
    main()
