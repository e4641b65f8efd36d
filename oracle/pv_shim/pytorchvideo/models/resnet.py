"""Restated pytorchvideo.models.resnet {BottleneckBlock, ResBlock, ResStage}
(imported at /root/reference/model/x3d.py:19)."""
import torch.nn as nn
from pytorchvideo.layers.utils import set_attributes


class BottleneckBlock(nn.Module):
    def __init__(self, *, conv_a=None, norm_a=None, act_a=None, conv_b=None, norm_b=None,
                 act_b=None, conv_c=None, norm_c=None):
        super().__init__()
        set_attributes(self, locals())
        assert all(m is not None for m in (self.conv_a, self.conv_b, self.conv_c))
        if self.norm_c is not None:
            self.norm_c.block_final_bn = True

    def forward(self, x):
        for m in (self.conv_a, self.norm_a, self.act_a, self.conv_b, self.norm_b, self.act_b,
                  self.conv_c, self.norm_c):
            if m is not None:
                x = m(x)
        return x


class ResBlock(nn.Module):
    def __init__(self, branch1_conv=None, branch1_norm=None, branch2=None, activation=None,
                 branch_fusion=None):
        super().__init__()
        set_attributes(self, locals())
        assert self.branch2 is not None

    def forward(self, x):
        if self.branch1_conv is None:
            x = self.branch_fusion(x, self.branch2(x))
        else:
            shortcut = self.branch1_conv(x)
            if self.branch1_norm is not None:
                shortcut = self.branch1_norm(shortcut)
            x = self.branch_fusion(shortcut, self.branch2(x))
        if self.activation is not None:
            x = self.activation(x)
        return x


class ResStage(nn.Module):
    def __init__(self, res_blocks):
        super().__init__()
        self.res_blocks = res_blocks

    def forward(self, x):
        for blk in self.res_blocks:
            x = blk(x)
        return x
