"""B200-native X3D builder — drop-in for the reference's `model/x3d.py`.

`create_x3d(**kw)` keeps the reference signature (model/x3d.py:543-584) and returns a module with
the same tree, parameter registration order and state-dict keys, so
`load_state_dict(strict=True)` of reference / pytorchvideo checkpoints works both ways.  The
nn.Conv3d / nn.BatchNorm3d objects are parameter containers only: `forward` of the stem and of
every ResStage runs the hand-written sm_100a kernels (change3d_b200.engine) through one
torch.autograd.Function per `blocks[i]`; there is no eager/cuDNN fallback.

Tensors at the module boundary are (B, C, T, H, W) like the reference; outputs are returned with
channels-last-3d strides (NDHWC in memory), inputs in any layout are accepted.
"""
from __future__ import annotations

import math
from typing import Callable, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import engine


# ---- pytorchvideo.layers.utils equivalents (call sites model/x3d.py:197,656-675) -----------------
def round_width(width, multiplier, min_width=8, divisor=8, ceil=False):
    if not multiplier:
        return width
    width *= multiplier
    min_width = min_width or divisor
    if ceil:
        width_out = max(min_width, int(math.ceil(width / divisor)) * divisor)
    else:
        width_out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if width_out < 0.9 * width:
        width_out += divisor
    return int(width_out)


def round_repeats(repeats, multiplier):
    if not multiplier:
        return repeats
    return int(math.ceil(multiplier * repeats))


class Swish(nn.Module):
    """Marker module (pytorchvideo.layers.swish.Swish): Swish is fused into conv_c's prologue."""

    def forward(self, x):  # pragma: no cover - never called on the product path
        raise RuntimeError("Swish is fused into the conv_c kernel; call the enclosing ResStage")


class SqueezeExcitation(nn.Module):
    """Parameter container with fvcore's layout: block = [Conv3d(C,r,1), ReLU, Conv3d(r,C,1), Sigmoid]
    (call site model/x3d.py:194-202).  Executed by c3d_bn_se_finalize."""

    def __init__(self, num_channels: int, num_channels_reduced: int, is_3d: bool = True):
        super().__init__()
        conv = nn.Conv3d if is_3d else nn.Conv2d
        self.is_3d = is_3d
        self.block = nn.Sequential(conv(num_channels, num_channels_reduced, kernel_size=1, stride=1, bias=True),
                                   nn.ReLU(),
                                   conv(num_channels_reduced, num_channels, kernel_size=1, stride=1, bias=True),
                                   nn.Sigmoid())


class Conv2plus1d(nn.Module):
    """Container mirroring pytorchvideo Conv2plus1d: `conv_t` holds the SPATIAL 1x3x3 conv and runs first,
    `conv_xy` holds the temporal depthwise 5x1x1 conv (model/x3d.py:70-92)."""

    def __init__(self, conv_t: nn.Module, conv_xy: nn.Module):
        super().__init__()
        self.conv_t = conv_t
        self.norm = None
        self.activation = None
        self.conv_xy = conv_xy


def _as_ndhwc(x: torch.Tensor) -> torch.Tensor:
    """(B,C,T,H,W) logical -> (B,T,H,W,C) dense view (copy only if not already channels-last-3d)."""
    y = x.permute(0, 2, 3, 4, 1)
    return y if y.is_contiguous() else y.contiguous()


def _as_ncdhw_view(y: torch.Tensor) -> torch.Tensor:
    return y.permute(0, 4, 1, 2, 3)


def _require_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"change3d_b200.{what}: CUDA tensor required — the B200 engine has no CPU/eager fallback")


class _StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, stem, grad_on, w_xy, w_t, gamma, beta):
        B, Cin, T, H, W = x.shape
        xc = x if (x.stride(4) == 1 and x.stride(3) == W) else x.contiguous()
        frames = [(xc[:, :, f], xc.stride(0), xc.stride(1)) for f in range(T)]
        need = grad_on and any(ctx.needs_input_grad)
        if need and ctx.needs_input_grad[0]:
            raise NotImplementedError("change3d_b200 stem: gradient w.r.t. a generic input clip is not implemented "
                                      "(the Change3D path only needs it for the shared perception frames — use "
                                      "change3d_b200.model.trainer.Encoder)")
        if need and not stem.training:
            raise RuntimeError("change3d_b200: backward through eval-mode BatchNorm is not implemented")
        out, saved = engine.stem_forward(stem, frames, B, H, W, stem.training, need)
        ctx.stem = stem
        ctx.saved = saved
        ctx.frames = frames
        ctx.keep = xc
        return _as_ncdhw_view(out)

    @staticmethod
    def backward(ctx, g):
        y, bnp, out = ctx.saved
        _, dwxy, dwt, dgamma, dbeta = engine.stem_backward(ctx.stem, ctx.frames, y, bnp, out, engine.owned_ndhwc(g),
                                                           len(ctx.frames) - 2, want_dperc=False)
        ctx.saved = None
        return None, None, None, dwxy, dwt, dgamma, dbeta


class ResNetBasicStem(nn.Module):
    """X3D stem (model/x3d.py:23-106): conv (Conv2plus1d) -> norm -> activation, one fused kernel chain."""

    def __init__(self, *, conv, norm, activation, pool=None):
        super().__init__()
        self.conv = conv
        self.norm = norm
        self.activation = activation
        self.pool = pool

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, "stem")
        return _StemFn.apply(x.float(), self, torch.is_grad_enabled(), self.conv.conv_t.weight, self.conv.conv_xy.weight, self.norm.weight,
                             self.norm.bias)


class BottleneckBlock(nn.Module):
    """Parameter container with pytorchvideo BottleneckBlock's attribute names (model/x3d.py:223-232)."""

    def __init__(self, *, conv_a, norm_a, act_a, conv_b, norm_b, act_b, conv_c, norm_c):
        super().__init__()
        self.conv_a, self.norm_a, self.act_a = conv_a, norm_a, act_a
        self.conv_b, self.norm_b, self.act_b = conv_b, norm_b, act_b
        self.conv_c, self.norm_c = conv_c, norm_c
        self.norm_c.block_final_bn = True


class ResBlock(nn.Module):
    """Parameter container with pytorchvideo ResBlock's attribute names (model/x3d.py:300-328)."""

    def __init__(self, *, branch1_conv, branch1_norm, branch2, activation):
        super().__init__()
        self.branch1_conv = branch1_conv
        self.branch1_norm = branch1_norm
        self.branch2 = branch2
        self.activation = activation

    def param_list(self):
        b2 = self.branch2
        ps = []
        if self.branch1_conv is not None:
            ps.append(self.branch1_conv.weight)
            if self.branch1_norm is not None:
                ps += [self.branch1_norm.weight, self.branch1_norm.bias]
        ps += [b2.conv_a.weight, b2.norm_a.weight, b2.norm_a.bias, b2.conv_b.weight, b2.norm_b[0].weight,
               b2.norm_b[0].bias]
        if hasattr(b2.norm_b[1], "block"):
            se = b2.norm_b[1].block
            ps += [se[0].weight, se[0].bias, se[2].weight, se[2].bias]
        ps += [b2.conv_c.weight, b2.norm_c.weight, b2.norm_c.bias]
        return ps


class _StageFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, stage, grad_on, *params):
        need = grad_on and any(ctx.needs_input_grad)
        if need and not stage.training:
            raise RuntimeError("change3d_b200: backward through eval-mode BatchNorm is not implemented")
        out, saved = engine.res_stage_forward(stage, _as_ndhwc(x), stage.training, need)
        ctx.stage = stage
        ctx.saved = saved
        return _as_ncdhw_view(out)

    @staticmethod
    def backward(ctx, g):
        dx, grads = engine.res_stage_backward(ctx.stage, ctx.saved, engine.owned_ndhwc(g))
        ctx.saved = None
        return (_as_ncdhw_view(dx), None, None) + tuple(grads)


class ResStage(nn.Module):
    """X3D residual stage (model/x3d.py:331-412); forward = engine.res_stage_forward."""

    def __init__(self, res_blocks: nn.ModuleList):
        super().__init__()
        self.res_blocks = res_blocks

    def param_list(self):
        ps = []
        for blk in self.res_blocks:
            ps += blk.param_list()
        return ps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, "ResStage")
        return _StageFn.apply(x.float(), self, torch.is_grad_enabled(), *self.param_list())


class ProjectedPool(nn.Module):
    """Parameter container for the X3D head pool (model/x3d.py:747-811); never executed by Change3D."""

    def __init__(self, *, pre_conv, pre_norm, pre_act, pool, post_conv, post_norm=None, post_act=None):
        super().__init__()
        self.pre_conv, self.pre_norm, self.pre_act = pre_conv, pre_norm, pre_act
        self.pool, self.post_conv, self.post_norm, self.post_act = pool, post_conv, post_norm, post_act


class ResNetBasicHead(nn.Module):
    """blocks[5]: exists so pretrained X3D-L checkpoints load strictly (model/trainer.py:43-48); the
    Change3D path never calls it (model/trainer.py:120-141 stops at blocks[3] / blocks[4])."""

    def __init__(self, *, pool, dropout, proj, activation, output_pool):
        super().__init__()
        self.pool, self.dropout, self.proj = pool, dropout, proj
        self.activation, self.output_pool = activation, output_pool

    def forward(self, x):
        raise NotImplementedError("the X3D classification head is outside the Change3D hot path "
                                  "(never executed by the reference either)")


class Net(nn.Module):
    def __init__(self, *, blocks: nn.ModuleList):
        super().__init__()
        self.blocks = blocks

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        return x


# ---- builders: same keyword surface as the reference ----------------------------------------------
def create_x3d_stem(*, in_channels: int, out_channels: int, conv_kernel_size=(5, 3, 3), conv_stride=(1, 2, 2),
                    conv_padding=(2, 1, 1), norm: Callable = nn.BatchNorm3d, norm_eps: float = 1e-5,
                    norm_momentum: float = 0.1, activation: Callable = nn.ReLU) -> nn.Module:
    if tuple(conv_kernel_size) != (5, 3, 3) or tuple(conv_stride) != (1, 1, 1) or in_channels != 3 or out_channels != 24:
        raise NotImplementedError("change3d_b200 stem kernel is specialised to the Change3D X3D-L stem "
                                  "(3->24, kernel (5,3,3), stride (1,1,1); model/x3d.py:563-564)")
    conv_xy_module = nn.Conv3d(in_channels, out_channels, kernel_size=(1, 3, 3), stride=(1, 1, 1), padding=(0, 1, 1),
                               bias=False)
    conv_t_module = nn.Conv3d(out_channels, out_channels, kernel_size=(5, 1, 1), stride=(1, 1, 1), padding=(2, 0, 0),
                              bias=False, groups=out_channels)
    return ResNetBasicStem(conv=Conv2plus1d(conv_t=conv_xy_module, conv_xy=conv_t_module),
                           norm=norm(num_features=out_channels, eps=norm_eps, momentum=norm_momentum),
                           activation=activation(), pool=None)


def create_x3d_bottleneck_block(*, dim_in: int, dim_inner: int, dim_out: int, conv_kernel_size=(3, 3, 3),
                                conv_stride=(1, 2, 2), norm: Callable = nn.BatchNorm3d, norm_eps: float = 1e-5,
                                norm_momentum: float = 0.1, se_ratio: float = 0.0625,
                                activation: Callable = nn.ReLU, inner_act: Callable = Swish) -> nn.Module:
    if tuple(conv_kernel_size) != (3, 3, 3) or conv_stride[0] != 1 or conv_stride[1] != conv_stride[2] \
            or conv_stride[1] not in (1, 2):
        raise NotImplementedError("depthwise kernel supports 3x3x3, stride (1,s,s), s in {1,2}")
    conv_a = nn.Conv3d(dim_in, dim_inner, kernel_size=(1, 1, 1), bias=False)
    norm_a = norm(num_features=dim_inner, eps=norm_eps, momentum=norm_momentum)
    conv_b = nn.Conv3d(dim_inner, dim_inner, kernel_size=conv_kernel_size, stride=conv_stride,
                       padding=[s // 2 for s in conv_kernel_size], bias=False, groups=dim_inner, dilation=(1, 1, 1))
    se = (SqueezeExcitation(num_channels=dim_inner, num_channels_reduced=round_width(dim_inner, se_ratio), is_3d=True)
          if se_ratio > 0.0 else nn.Identity())
    norm_b = nn.Sequential(norm(num_features=dim_inner, eps=norm_eps, momentum=norm_momentum), se)
    conv_c = nn.Conv3d(dim_inner, dim_out, kernel_size=(1, 1, 1), bias=False)
    norm_c = norm(num_features=dim_out, eps=norm_eps, momentum=norm_momentum)
    return BottleneckBlock(conv_a=conv_a, norm_a=norm_a, act_a=activation(), conv_b=conv_b, norm_b=norm_b,
                           act_b=inner_act(), conv_c=conv_c, norm_c=norm_c)


def create_x3d_res_block(*, dim_in: int, dim_inner: int, dim_out: int, bottleneck: Callable = create_x3d_bottleneck_block,
                         use_shortcut: bool = True, conv_kernel_size=(3, 3, 3), conv_stride=(1, 2, 2),
                         norm: Callable = nn.BatchNorm3d, norm_eps: float = 1e-5, norm_momentum: float = 0.1,
                         se_ratio: float = 0.0625, activation: Callable = nn.ReLU,
                         inner_act: Callable = Swish) -> nn.Module:
    has_conv = (dim_in != dim_out or np.prod(conv_stride) > 1) and use_shortcut
    if has_conv and tuple(conv_stride) != (1, 2, 2):
        raise NotImplementedError("shortcut conv kernel supports stride (1,2,2) only (the X3D configuration)")
    norm_model = norm(num_features=dim_out) if dim_in != dim_out else None   # default BN args (model/x3d.py:296-298)
    return ResBlock(
        branch1_conv=nn.Conv3d(dim_in, dim_out, kernel_size=(1, 1, 1), stride=conv_stride, bias=False) if has_conv else None,
        branch1_norm=norm_model if dim_in != dim_out and use_shortcut else None,
        branch2=bottleneck(dim_in=dim_in, dim_inner=dim_inner, dim_out=dim_out, conv_kernel_size=conv_kernel_size,
                           conv_stride=conv_stride, norm=norm, norm_eps=norm_eps, norm_momentum=norm_momentum,
                           se_ratio=se_ratio, activation=activation, inner_act=inner_act),
        activation=activation())


def create_x3d_res_stage(*, depth: int, dim_in: int, dim_inner: int, dim_out: int,
                         bottleneck: Callable = create_x3d_bottleneck_block, conv_kernel_size=(3, 3, 3),
                         conv_stride=(1, 2, 2), norm: Callable = nn.BatchNorm3d, norm_eps: float = 1e-5,
                         norm_momentum: float = 0.1, se_ratio: float = 0.0625, activation: Callable = nn.ReLU,
                         inner_act: Callable = Swish) -> nn.Module:
    blocks = []
    for idx in range(depth):
        blocks.append(create_x3d_res_block(
            dim_in=dim_in if idx == 0 else dim_out, dim_inner=dim_inner, dim_out=dim_out, bottleneck=bottleneck,
            conv_kernel_size=conv_kernel_size, conv_stride=conv_stride if idx == 0 else (1, 1, 1), norm=norm,
            norm_eps=norm_eps, norm_momentum=norm_momentum, se_ratio=(se_ratio if (idx + 1) % 2 else 0.0),
            activation=activation, inner_act=inner_act))
    return ResStage(res_blocks=nn.ModuleList(blocks))


def create_x3d_head(*, dim_in: int, dim_inner: int, dim_out: int, num_classes: int, pool_act: Callable = nn.ReLU,
                    pool_kernel_size=(13, 5, 5), norm: Callable = nn.BatchNorm3d, norm_eps: float = 1e-5,
                    norm_momentum: float = 0.1, bn_lin5_on=False, dropout_rate: float = 0.5,
                    activation: Callable = nn.Softmax, output_with_global_average: bool = True) -> nn.Module:
    pool = ProjectedPool(
        pre_conv=nn.Conv3d(dim_in, dim_inner, kernel_size=(1, 1, 1), bias=False),
        pre_norm=norm(num_features=dim_inner, eps=norm_eps, momentum=norm_momentum),
        pre_act=None if pool_act is None else pool_act(),
        pool=nn.AdaptiveAvgPool3d((1, 1, 1)) if pool_kernel_size is None else nn.AvgPool3d(pool_kernel_size, stride=1),
        post_conv=nn.Conv3d(dim_inner, dim_out, kernel_size=(1, 1, 1), bias=False),
        post_norm=norm(num_features=dim_out, eps=norm_eps, momentum=norm_momentum) if bn_lin5_on else None,
        post_act=None if pool_act is None else pool_act())
    if activation is None:
        act = None
    elif activation == nn.Softmax:
        act = activation(dim=1)
    elif activation == nn.Sigmoid:
        act = activation()
    else:
        raise NotImplementedError(f"{activation} is not supported as an activation function.")
    return ResNetBasicHead(proj=nn.Linear(dim_out, num_classes, bias=True), activation=act, pool=pool,
                           dropout=nn.Dropout(dropout_rate) if dropout_rate > 0 else None,
                           output_pool=nn.AdaptiveAvgPool3d(1) if output_with_global_average else None)


def create_x3d(*, input_channel: int = 3, input_clip_length: int = 13, input_crop_size: int = 160,
               model_num_class: int = 400, dropout_rate: float = 0.5, width_factor: float = 2.0,
               depth_factor: float = 2.2, norm: Callable = nn.BatchNorm3d, norm_eps: float = 1e-5,
               norm_momentum: float = 0.1, activation: Callable = nn.ReLU, stem_dim_in: int = 12,
               stem_conv_kernel_size: Tuple[int] = (5, 3, 3), stem_conv_stride: Tuple[int] = (1, 1, 1),
               stage_conv_kernel_size=((3, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
               stage_spatial_stride: Tuple[int] = (2, 2, 2, 2), stage_temporal_stride: Tuple[int] = (1, 1, 1, 1),
               bottleneck: Callable = create_x3d_bottleneck_block, bottleneck_factor: float = 2.25,
               se_ratio: float = 0.0625, inner_act: Callable = Swish, head_dim_out: int = 2048,
               head_pool_act: Callable = nn.ReLU, head_bn_lin5_on: bool = False, head_activation: Callable = None,
               head_output_with_global_average: bool = True) -> nn.Module:
    """Same keyword surface and defaults as the reference builder (model/x3d.py:543-584)."""
    if norm is not nn.BatchNorm3d or activation is not nn.ReLU or inner_act is not Swish:
        raise NotImplementedError("the fused kernels implement BatchNorm3d + ReLU + Swish (the Change3D configuration)")
    blocks = []
    stem_dim_out = round_width(stem_dim_in, width_factor)
    blocks.append(create_x3d_stem(in_channels=input_channel, out_channels=stem_dim_out,
                                  conv_kernel_size=stem_conv_kernel_size, conv_stride=stem_conv_stride,
                                  conv_padding=[s // 2 for s in stem_conv_kernel_size], norm=norm, norm_eps=norm_eps,
                                  norm_momentum=norm_momentum, activation=activation))
    stage_depths = [1, 2, 5, 3]
    exp_stage = 2.0
    d1 = stem_dim_in
    d2 = round_width(d1, exp_stage, divisor=8)
    d3 = round_width(d2, exp_stage, divisor=8)
    d4 = round_width(d3, exp_stage, divisor=8)
    dim_in = stem_dim_out
    dim_out = dim_inner = None
    for idx, sd in enumerate([d1, d2, d3, d4]):
        dim_out = round_width(sd, width_factor)
        dim_inner = int(bottleneck_factor * dim_out)
        depth = round_repeats(stage_depths[idx], depth_factor)
        blocks.append(create_x3d_res_stage(
            depth=depth, dim_in=dim_in, dim_inner=dim_inner, dim_out=dim_out, bottleneck=bottleneck,
            conv_kernel_size=stage_conv_kernel_size[idx],
            conv_stride=(stage_temporal_stride[idx], stage_spatial_stride[idx], stage_spatial_stride[idx]),
            norm=norm, norm_eps=norm_eps, norm_momentum=norm_momentum, se_ratio=se_ratio, activation=activation,
            inner_act=inner_act))
        dim_in = dim_out
    total_spatial_stride = stem_conv_stride[1] * np.prod(stage_spatial_stride)
    total_temporal_stride = stem_conv_stride[0] * np.prod(stage_temporal_stride)
    assert input_clip_length >= total_temporal_stride, "Clip length doesn't match temporal stride!"
    assert input_crop_size >= total_spatial_stride, "Crop size doesn't match spatial stride!"
    head_pool_kernel_size = (input_clip_length // total_temporal_stride,
                             int(math.ceil(input_crop_size / total_spatial_stride)),
                             int(math.ceil(input_crop_size / total_spatial_stride)))
    blocks.append(create_x3d_head(dim_in=dim_out, dim_inner=dim_inner, dim_out=head_dim_out,
                                  num_classes=model_num_class, pool_act=head_pool_act,
                                  pool_kernel_size=head_pool_kernel_size, norm=norm, norm_eps=norm_eps,
                                  norm_momentum=norm_momentum, bn_lin5_on=head_bn_lin5_on, dropout_rate=dropout_rate,
                                  activation=head_activation,
                                  output_with_global_average=head_output_with_global_average))
    return Net(blocks=nn.ModuleList(blocks))
