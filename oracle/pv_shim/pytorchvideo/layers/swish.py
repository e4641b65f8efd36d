"""Restated pytorchvideo.layers.swish (imported at /root/reference/model/x3d.py:15)."""
import torch
import torch.nn as nn


class _SwishFn(torch.autograd.Function):
    # memory-lean swish: only the input is kept for backward
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return x * torch.sigmoid(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        s = torch.sigmoid(x)
        return g * (s * (1 + x * (1 - s)))


class Swish(nn.Module):
    def forward(self, x):
        return _SwishFn.apply(x)
