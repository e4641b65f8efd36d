"""B200-native ChangeDecoder — drop-in for the reference's `model/change_decoder.py`.

Same constructor (`ChangeDecoder(args, in_dim, has_sigmoid)`), module tree and state-dict keys
(`up_c{4,3,2}.0.weight`, `up_c{4,3,2}.1.{weight,bias}`, `up_c1.0.weight`), so `weight_init(decoder)`
(model/utils.py:20-82) and reference checkpoints apply unchanged.  `forward(f)` takes the list of four
(B,C,H,W) features and runs: 1x1 Conv2d (pointwise GEMM) -> ConvTranspose2d k4 s2 p1 as four
parity-class gather-GEMMs with bias + skip-add fused in the epilogue -> 3x3 head (+ sigmoid).
"""
from typing import List

import torch
import torch.nn as nn

from .. import engine


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c1, c2, c3, c4, dec, grad_on, *params):
        need = grad_on and any(ctx.needs_input_grad)
        pred, saved = engine.decoder_forward(dec, [c1, c2, c3, c4], need)
        ctx.dec = dec
        ctx.saved = saved
        return pred

    @staticmethod
    def backward(ctx, g):
        dfeats, grads = engine.decoder_backward(ctx.dec, ctx.saved, g.contiguous())
        ctx.saved = None
        return tuple(dfeats) + (None, None) + tuple(grads)


class ChangeDecoder(nn.Module):
    def __init__(self, args, in_dim: List[int] = [64, 128, 256, 384], has_sigmoid: bool = False) -> None:
        super().__init__()
        self.has_sigmoid = has_sigmoid
        c1, c2, c3, c4 = in_dim
        self.up_c4 = nn.Sequential(nn.Conv2d(c4, c3, kernel_size=1, bias=False),
                                   nn.ConvTranspose2d(c3, c3, kernel_size=4, stride=2, padding=1))
        self.up_c3 = nn.Sequential(nn.Conv2d(c3, c2, kernel_size=1, bias=False),
                                   nn.ConvTranspose2d(c2, c2, kernel_size=4, stride=2, padding=1))
        self.up_c2 = nn.Sequential(nn.Conv2d(c2, c1, kernel_size=1, bias=False),
                                   nn.ConvTranspose2d(c1, c1, kernel_size=4, stride=2, padding=1))
        num_class = 1 if has_sigmoid else args.num_class
        self.up_c1 = nn.Sequential(nn.Conv2d(c1, num_class, kernel_size=3, stride=1, padding=1, bias=False))

    def param_list(self):
        return [self.up_c4[0].weight, self.up_c4[1].weight, self.up_c4[1].bias,
                self.up_c3[0].weight, self.up_c3[1].weight, self.up_c3[1].bias,
                self.up_c2[0].weight, self.up_c2[1].weight, self.up_c2[1].bias,
                self.up_c1[0].weight]

    def forward(self, f: List[torch.Tensor]) -> torch.Tensor:
        c1, c2, c3, c4 = f
        if not c1.is_cuda:
            raise RuntimeError("change3d_b200.ChangeDecoder: CUDA tensors required (no CPU/eager fallback)")
        return _DecoderFn.apply(c1, c2, c3, c4, self, torch.is_grad_enabled(), *self.param_list())
