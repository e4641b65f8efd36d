"""Stem backward (c3d_stem_bwd: BN backward on the fly, temporal / spatial weight gradients, gradient of the shared
perception frames) against torch autograd over the fp64 oracle, for both kernels behind the entry point
(C3D_STEM_BWD=1: channel x pixel-pair kernel, the default; 0: the quad-per-thread kernel of round 1), with the ReLU
mask recomputed inside the kernels (C3D_STEM_MASKED=1, default) or applied by c3d_relu_bwd_stats beforehand.
Autograd of model/x3d.py:70-99 + the frame assembly of model/trainer.py:154-162."""
import os

import pytest
import torch

from oracle import change3d_oracle as O
from tests.gpu_util import check

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reference(sd, pre, post, perc, wgt, dtype):
    osd = O.clone_sd({k: v for k, v in sd.items() if k.startswith("blocks.0.")}, dtype=dtype, requires_grad=True)
    pr = perc.detach().clone().to(dtype).requires_grad_(True)
    B = pre.shape[0]
    x = torch.cat([pre.to(dtype).unsqueeze(2), pr.expand(B, -1, -1, -1, -1), post.to(dtype).unsqueeze(2)], dim=2)
    out = O.stem(osd, x, True)
    (out * wgt.to(dtype)).sum().backward()
    return out.detach(), pr.grad, osd


# odd widths / heights: partial tiles, pixel pairs cut by the image border, tiles in all corners
@pytest.mark.parametrize("P,H,W,B", [(1, 20, 36, 2), (1, 19, 37, 3), (2, 24, 41, 2), (3, 17, 33, 2), (1, 64, 96, 2)])
@pytest.mark.parametrize("impl,masked", [("1", "1"), ("1", "0"), ("0", "1"), ("0", "0")])
def test_stem_backward(P, H, W, B, impl, masked):
    _run(P, H, W, B, impl, masked)


@pytest.mark.parametrize("P,H,W,B", [(1, 256, 256, 4), (3, 128, 192, 4)])
def test_stem_backward_persistent(P, H, W, B):
    """More tiles than resident CTAs: every CTA of the (persistent) forward and backward kernels walks several tiles, so
    the shared-memory hand-over between consecutive tiles of a CTA is on the path (1024 / 768 tiles on 296 / 592 CTAs)."""
    _run(P, H, W, B, "1", "1")


def _run(P, H, W, B, impl, masked):
    from change3d_b200 import engine
    from change3d_b200.model.x3d import create_x3d
    sd = O.synth_state_dict(O.x3d_schema(), 11)
    net = create_x3d(input_clip_length=3, depth_factor=5.0)
    net.load_state_dict(sd, strict=True)
    stem = net.blocks[0].to(DEV).train()
    g = torch.Generator().manual_seed(100 * P + W)
    pre, post = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)
    perc = torch.randn(1, 3, P, H, W, generator=g)
    T = P + 2
    wgt = torch.randn(B, 24, T, H, W, generator=g)
    out64, dperc64, g64 = _reference(sd, pre, post, perc, wgt, torch.float64)

    hw = H * W
    pre_d, post_d, perc_d = pre.to(DEV).contiguous(), post.to(DEV).contiguous(), perc.to(DEV).contiguous()
    frames = [(pre_d, 3 * hw, hw)] + [(perc_d[0, :, f], 0, P * hw) for f in range(P)] + [(post_d, 3 * hw, hw)]
    old = {k: os.environ.get(k) for k in ("C3D_STEM_BWD", "C3D_STEM_MASKED")}
    os.environ["C3D_STEM_BWD"] = impl          # 1: channel x pixel-pair kernel, 0: quad-per-thread kernel
    os.environ["C3D_STEM_MASKED"] = masked     # 1: ReLU mask recomputed from y inside the kernels, 0: masked copy in memory
    try:
        out, (y, bnp, outs) = engine.stem_forward(stem, frames, B, H, W, True, True)
        gd = wgt.permute(0, 2, 3, 4, 1).contiguous().to(DEV)                 # (B, T, H, W, 24)
        dperc, dwxy, dwt, dgamma, dbeta = engine.stem_backward(stem, frames, y, bnp, outs, gd, P)
        torch.cuda.synchronize()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    tag = f"stem_bwd impl{impl} masked{masked} P{P} {H}x{W}"
    check(tag + " forward", out.permute(0, 4, 1, 2, 3), out64, 5e-6)
    check(tag + " dperception", dperc, dperc64, 2e-5)
    check(tag + " dw spatial", dwxy, g64["blocks.0.conv.conv_t.weight"].grad, 2e-5)
    check(tag + " dw temporal", dwt, g64["blocks.0.conv.conv_xy.weight"].grad, 2e-5)
    check(tag + " dgamma", dgamma, g64["blocks.0.norm.weight"].grad, 2e-5)
    check(tag + " dbeta", dbeta, g64["blocks.0.norm.bias"].grad, 2e-5)
