"""Restated pytorchvideo.models.head.ResNetBasicHead (imported at /root/reference/model/x3d.py:17).
Never executed on the Change3D path; exists so blocks.5.* state-dict keys are present."""
import torch.nn as nn
from pytorchvideo.layers.utils import set_attributes


class ResNetBasicHead(nn.Module):
    def __init__(self, pool=None, dropout=None, proj=None, activation=None, output_pool=None):
        super().__init__()
        set_attributes(self, locals())
        assert self.proj is not None

    def forward(self, x):
        if self.pool is not None:
            x = self.pool(x)
        if self.dropout is not None:
            x = self.dropout(x)
        if self.proj is not None:
            x = x.permute((0, 2, 3, 4, 1))
            x = self.proj(x)
            x = x.permute((0, 4, 1, 2, 3))
        if self.activation is not None:
            x = self.activation(x)
        if self.output_pool is not None:
            x = self.output_pool(x)
            x = x.view(x.shape[0], -1)
        return x
