"""Restated pytorchvideo.layers.convolutions.Conv2plus1d (imported at /root/reference/model/x3d.py:14)."""
import torch.nn as nn
from .utils import set_attributes


class Conv2plus1d(nn.Module):
    """Runs the module stored as `conv_t` FIRST unless conv_xy_first (reference stores the
    spatial conv there: /root/reference/model/x3d.py:87-92)."""

    def __init__(self, *, conv_t=None, norm=None, activation=None, conv_xy=None, conv_xy_first=False):
        super().__init__()
        set_attributes(self, locals())
        assert self.conv_t is not None and self.conv_xy is not None

    def forward(self, x):
        x = self.conv_xy(x) if self.conv_xy_first else self.conv_t(x)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return self.conv_t(x) if self.conv_xy_first else self.conv_xy(x)
