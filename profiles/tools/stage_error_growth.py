"""Per-block error growth through a whole ResStage, forward and backward, for three fp32 implementations against the
fp64 oracle (is the res5 / CC backward error chaos or a kernel defect?).

  torch_fp32   the oracle itself in fp32 on the host (ATen kernels)            -- the reference's arithmetic
  c3d_tc       this library, default path (tcgen05 3xTF32 GEMMs / weight gradients)
  c3d_ffma     this library with C3D_TC=0 (fp32 FFMA GEMMs, no tensor cores)

For every block b of the stage it prints
  fwd    max|out_b - out_b^64| / max|out_b^64|     and the number of ReLU-mask decisions (out_b > 0) that differ from fp64
  bwd    max|g_b - g_b^64| / max|g_b^64|           (g_b = gradient w.r.t. the input of block b)
plus the relative error of a few parameter gradients.  If all three implementations drift from fp64 alike, the growth
is the conditioning of the chain (train-mode BatchNorm over few samples + ReLU-mask flips); if only one of ours drifts,
that path has a defect.

    python profiles/tools/stage_error_growth.py --stage 4 --B 2 --H 4          # the CC golden case's res5 regime (2x2 out)
    python profiles/tools/stage_error_growth.py --stage 4 --B 8 --H 16
    python profiles/tools/stage_error_growth.py --stage 3 --B 2 --H 16
PROFILING / DIAGNOSTIC TOOL (imports the oracle as the checker); not part of the product or of bench.py.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import change3d_oracle as O   # noqa: E402


def ndhwc(x):
    return x.permute(0, 2, 3, 4, 1).contiguous()


def ncdhw(x):
    return x.permute(0, 4, 1, 2, 3)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


PARAM_KEYS = ["res_blocks.0.branch1_conv.weight", "res_blocks.0.branch2.conv_a.weight", "res_blocks.0.branch2.conv_b.weight",
              "res_blocks.0.branch2.norm_b.1.block.0.weight", "res_blocks.{mid}.branch2.norm_c.bias",
              "res_blocks.{last}.branch2.conv_c.weight", "res_blocks.{last}.branch2.conv_a.weight"]


def run_oracle(sd, s, depth, x, wgt, dtype):
    osd = O.clone_sd(sd, dtype=dtype, requires_grad=True)
    xr = x.detach().clone().to(dtype).requires_grad_(True)
    h, outs = xr, []
    for b in range(depth):
        h = O.res_block(osd, f"blocks.{s}.res_blocks.{b}.", h, b == 0, (b + 1) % 2 == 1, True)
        h.retain_grad()
        outs.append(h)
    (h * wgt.to(dtype)).sum().backward()
    gin = [xr.grad] + [o.grad for o in outs[:-1]]
    pg = {k: v.grad for k, v in osd.items() if v.requires_grad and v.grad is not None}
    return [o.detach() for o in outs], gin, pg


def run_c3d(net, s, x, wgt):
    from change3d_b200 import engine
    stage = net.blocks[s]
    xd = ndhwc(x).cuda()
    out, saved = engine.res_stage_forward(stage, xd, True, True)
    outs = [ncdhw(sv.out).clone() for sv in saved]
    N = xd.shape[0]
    params = stage.param_list()
    ga = engine.GradArena(params, xd.device)
    total = 0
    for blk in stage.res_blocks:
        ci = blk.branch2.conv_a.weight.shape[0]
        cout = blk.branch2.conv_c.weight.shape[0]
        total += engine.block_bwd_stat_doubles(N, ci, cout, blk.branch1_conv is not None and blk.branch1_norm is not None)
    arena = engine.StatArena(total, xd.device)
    counts = [len(blk.param_list()) for blk in stage.res_blocks]
    starts = [0]
    for c in counts[:-1]:
        starts.append(starts[-1] + c)
    g = ndhwc(wgt).cuda()
    gin = [None] * len(stage.res_blocks)
    for bi in range(len(stage.res_blocks) - 1, -1, -1):
        ga.i = starts[bi]
        g = engine.res_block_backward(stage.res_blocks[bi], saved[bi], g.clone(), arena, ga)
        gin[bi] = ncdhw(g).clone()
    torch.cuda.synchronize()
    names = [n for n, _ in stage.named_parameters()]
    byid = {id(p): n for n, p in stage.named_parameters()}
    pg = {f"blocks.{s}." + byid[id(p)]: v.clone() for p, v in zip(params, ga.views)}
    del names
    return outs, gin, pg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", type=int, default=4)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--H", type=int, default=4)
    ap.add_argument("--T", type=int, default=3)
    ap.add_argument("--seed", type=int, default=21)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    from change3d_b200.model.x3d import create_x3d
    s = a.stage
    cin, _, cout, depth = O.STAGES[s - 1]
    sd = O.synth_state_dict(O.x3d_schema(), a.seed)
    g = torch.Generator().manual_seed(a.seed + 1)
    x = torch.relu(torch.randn(a.B, cin, a.T, a.H, a.H, generator=g))
    wgt = torch.randn(a.B, cout, a.T, a.H // 2, a.H // 2, generator=g)

    o64, g64, p64 = run_oracle(sd, s, depth, x, wgt, torch.float64)
    res = {"torch_fp32": run_oracle(sd, s, depth, x, wgt, torch.float32)}
    for name, tc in (("c3d_tc", "1"), ("c3d_ffma", "0")):
        os.environ["C3D_TC"] = tc
        net = create_x3d(input_clip_length=3, depth_factor=5.0)
        net.load_state_dict(sd, strict=True)
        res[name] = run_c3d(net.cuda().train(), s, x, wgt)
    os.environ.pop("C3D_TC", None)

    keys = [k.format(mid=depth // 2, last=depth - 1) for k in PARAM_KEYS]
    keys = [f"blocks.{s}." + k for k in keys if f"blocks.{s}." + k in p64]
    report = {"case": f"stage {s} ({depth} blocks) B{a.B} T{a.T} {a.H}x{a.H} -> {a.H // 2}x{a.H // 2}, BN samples/channel "
                      f"{a.B * a.T * (a.H // 2) ** 2}", "impl": {}}
    print(report["case"])
    for name, (outs, gin, pg) in res.items():
        fwd = [rel(outs[b], o64[b]) for b in range(depth)]
        flips = [int(((outs[b].detach().cpu() > 0) != (o64[b] > 0)).sum()) for b in range(depth)]
        bwd = [rel(gin[b], g64[b]) for b in range(depth)]
        par = {k: rel(pg[k], p64[k]) for k in keys}
        report["impl"][name] = {"fwd": fwd, "relu_flips": flips, "bwd_input_grad": bwd, "param_grads": par}
        print(f"== {name}")
        print("  block  fwd_err    flips  bwd_err(grad wrt block input)")
        for b in range(depth):
            print(f"  {b:3d}   {fwd[b]:.3e}  {flips[b]:5d}  {bwd[b]:.3e}")
        for k, v in par.items():
            print(f"  grad {k}: {v:.3e}")
    if a.json:
        with open(a.json, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
