"""Generates tests/golden/cc_enc_b2_32.npz with the REFERENCE's own Encoder (model/trainer.py:20-167) on the
change-captioning feature path (`output_final=True`: x3d.blocks[0..4], frame P, no enhance) — train-mode
output and autograd gradients.  TEST INFRASTRUCTURE, authoring container only:  python -m oracle.make_golden_cc
"""
import contextlib
import io
import os

import numpy as np
import torch

from oracle import change3d_oracle as O
from oracle import reference_loader as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, H, W, SEED = 2, 32, 32, 23
GRAD_KEYS = ["perception_frames", "x3d.blocks.1.res_blocks.0.branch2.conv_a.weight",
             "x3d.blocks.4.res_blocks.0.branch1_conv.weight", "x3d.blocks.4.res_blocks.0.branch2.conv_b.weight",
             "x3d.blocks.4.res_blocks.0.branch2.norm_b.1.block.0.weight",
             "x3d.blocks.4.res_blocks.14.branch2.conv_c.weight", "x3d.blocks.4.res_blocks.7.branch2.norm_c.bias"]


def encoder_state(seed: int):
    sd = O.synth_state_dict(O.trainer_schema("bcd", 1, H, W, 1), seed)
    return {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}


def weights(seed: int) -> torch.Tensor:
    return torch.randn(B, 192, H // 16, W // 16, generator=torch.Generator().manual_seed(seed + 7))


def main() -> None:
    ref = R.load()
    torch.manual_seed(16)
    with contextlib.redirect_stdout(io.StringIO()):
        enc = ref.Encoder(R.make_args("bcd", H, W, 1), [24, 24, 48, 96]).float()
    esd = encoder_state(SEED)
    enc.load_state_dict(esd, strict=True)
    enc.train()
    pre, post, _ = O.synth_inputs(B, H, W, SEED)
    out = enc(pre, post, True)
    (out * weights(SEED)).sum().backward()
    named = dict(enc.named_parameters())
    res = {"out": out.detach().numpy()}
    for k in GRAD_KEYS:
        res["grad:" + k] = named[k].grad.numpy()
    res["stat:x3d.blocks.4.res_blocks.14.branch2.norm_c.running_mean"] = \
        enc.state_dict()["x3d.blocks.4.res_blocks.14.branch2.norm_c.running_mean"].numpy()
    path = os.path.join(ROOT, "tests", "golden", "cc_enc_b2_32.npz")
    np.savez_compressed(path, **res)
    print(path, f"{os.path.getsize(path) / 1024:.0f} KiB", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
