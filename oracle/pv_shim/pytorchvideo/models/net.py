"""Restated pytorchvideo.models.net.Net (imported at /root/reference/model/x3d.py:18)."""
import torch.nn as nn


class Net(nn.Module):
    def __init__(self, *, blocks):
        super().__init__()
        assert blocks is not None
        self.blocks = blocks

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        return x
