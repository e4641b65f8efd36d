"""Restated pytorchvideo.models.stem.ResNetBasicStem (imported at /root/reference/model/x3d.py:20)."""
import torch.nn as nn
from pytorchvideo.layers.utils import set_attributes


class ResNetBasicStem(nn.Module):
    def __init__(self, *, conv=None, norm=None, activation=None, pool=None):
        super().__init__()
        set_attributes(self, locals())
        assert self.conv is not None

    def forward(self, x):
        x = self.conv(x)
        for m in (self.norm, self.activation, self.pool):
            if m is not None:
                x = m(x)
        return x
