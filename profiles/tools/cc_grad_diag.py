"""Where does the error of a change-captioning gradient sit?  (tests/test_gpu_cc_encoder.py case: B=2, 32x32, seed 23.)

For the last res5 block's conv_c weight gradient (192 x 432) and the SE fc1 weight gradient of block 0: per-output-row
error against fp64, for this library (C3D_TC=1 and 0) and torch fp32, and the ReLU-mask decisions of the final stage
output (B,192,T,2,2) that differ from fp64.  One flipped decision zeroes / restores one d_pre element, which moves ONE row of
dW_c by O(1 / 24 samples): a flip shows as a single bad row whose channel matches the flipped element; a defective
kernel shows as errors spread over all rows.   DIAGNOSTIC TOOL (oracle = checker), GPU.
"""
import argparse, contextlib, io, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import change3d_oracle as O                               # noqa: E402
from oracle.make_golden_cc import B, GRAD_KEYS, H, SEED, W, weights   # noqa: E402

K14 = "x3d.blocks.4.res_blocks.14.branch2.conv_c.weight"
KSE = "x3d.blocks.4.res_blocks.0.branch2.norm_b.1.block.0.weight"


def oracle(full, pre, post, w, dtype):
    s = O.clone_sd(full, dtype=dtype, requires_grad=True)
    x = O.assemble_frames(pre.to(dtype), post.to(dtype), s["encoder.perception_frames"])
    for i in range(5):
        x = O.x3d_block(s, i, x, True, "encoder.x3d.")
    x.retain_grad()
    (x[:, :, 1] * w.to(dtype)).sum().backward()
    return x.detach(), {k: s["encoder." + k].grad.detach() for k in GRAD_KEYS}


def mine(full, pre, post, w, tc):
    from change3d_b200.model.trainer import Encoder
    os.environ["C3D_TC"] = tc
    args = argparse.Namespace(num_perception_frame=1, num_class=1, in_height=H, in_width=W, dataset="LEVIR-CC",
                              pretrained="/nonexistent")
    with contextlib.redirect_stdout(io.StringIO()):
        enc = Encoder(args, [24, 24, 48, 96])
    enc.load_state_dict({k[len("encoder."):]: v for k, v in full.items() if k.startswith("encoder.")}, strict=True)
    enc = enc.cuda().float().train()
    grabbed = {}
    enc.x3d.blocks[4].register_forward_hook(lambda m, i, o: grabbed.__setitem__("out", o.detach().clone()))
    out = enc(pre.cuda(), post.cuda(), True)
    (out * w.cuda()).sum().backward()
    torch.cuda.synchronize()
    named = dict(enc.named_parameters())
    return grabbed["out"].cpu(), {k: named[k].grad.detach().cpu() for k in GRAD_KEYS}


def rows(g, g64):
    d = (g.double() - g64).abs().reshape(g64.shape[0], -1).max(dim=1).values / g64.abs().max()
    return d


def main():
    full = O.synth_state_dict(O.trainer_schema("bcd", 1, H, W, 1), SEED)
    pre, post, _ = O.synth_inputs(B, H, W, SEED)
    w = weights(SEED)
    x64, g64 = oracle(full, pre, post, w, torch.float64)
    impls = {"torch_fp32": oracle(full, pre, post, w, torch.float32), "c3d_tc": mine(full, pre, post, w, "1"),
             "c3d_ffma": mine(full, pre, post, w, "0")}
    os.environ.pop("C3D_TC", None)
    rep = {}
    for name, (x, g) in impls.items():
        flips = ((x.double() > 0) != (x64 > 0)).nonzero()
        fl = [(int(i[1]), float(x64[tuple(i)] / x64.abs().max())) for i in flips]       # (channel, fp64 value / max)
        r14, rse = rows(g[K14], g64[K14]), rows(g[KSE], g64[KSE])
        rep[name] = {"final_out_err": float((x.double() - x64).abs().max() / x64.abs().max()),
                     "final_out_relu_flips(channel, fp64 value/max)": fl,
                     "conv_c14_rows_worst": [(int(i), float(r14[i])) for i in r14.argsort(descending=True)[:4]],
                     "conv_c14_rows_median": float(r14.median()),
                     "se_fc1_rows_worst": [(int(i), float(rse[i])) for i in rse.argsort(descending=True)[:4]],
                     "se_fc1_rows_median": float(rse.median()),
                     "all": {k: float((g[k].double() - g64[k]).abs().max() / g64[k].abs().max()) for k in GRAD_KEYS}}
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
