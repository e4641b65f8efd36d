"""Kernel sequencing for the X3D hot path (forward and backward), on NDHWC fp32 tensors.

One function per reference module on the path; each saves exactly what its backward needs.
Nothing here computes on the host or with torch math ops: torch only allocates device buffers.

Per bottleneck block (model/x3d.py:109-232, train mode = batch-statistics BN):
  conv_a      pw_gemm            x -> y_a (raw) + sum/sumsq          -> bn_finalize
  conv_b      dw_conv_fwd        relu(bn_a(y_a)) -> y_b (raw) + per-sample sums -> bn_se_finalize (BN_b + SE gate)
  conv_c      pw_gemm            swish(gate*bn_b(y_b)) -> y_c (raw) + sums       -> bn_finalize
  join        bn_add_relu        relu(shortcut + bn_c(y_c))
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from ._lib import (EPI_ABSDIFF_BWD, EPI_ADD2, EPI_RELU_ADD, EPI_STORE, EPI_SWISH_BWD, MAP_DENSE, MAP_SUB2,
                   PRO_ABSDIFF, PRO_BN_GATE_SWISH, PRO_BN_RELU, PRO_BNBWD, PRO_MASK_POS, PRO_NONE)


import os

_SIDE = {}      # device index -> side stream


def _side_stream():
    """Second stream for the weight-gradient GEMMs of a block's backward (default on; C3D_SIDE_STREAM=0 disables): they only depend on
    the BN-backward coefficients, not on the dgrad chain, so they can fill the SMs the tiny finalizer kernels and
    kernel tails leave idle.  Forked and joined inside res_block_backward (also under CUDA-graph capture)."""
    if os.environ.get("C3D_SIDE_STREAM", "1") != "1":
        return None
    dev = torch.cuda.current_device()
    if dev not in _SIDE:
        _SIDE[dev] = torch.cuda.Stream(device=dev)
    return _SIDE[dev]


class StatArena:
    """One zero-filled fp64 buffer per stage call; the BN/SE statistics of every layer are carved
    out of it (a single memset instead of one per layer)."""

    def __init__(self, n_doubles: int, device):
        self.buf = torch.zeros(max(n_doubles, 1), dtype=torch.float64, device=device)
        self.off = 0

    def take(self, n: int) -> torch.Tensor:
        if self.off + n > self.buf.numel():
            raise RuntimeError("StatArena overflow")
        t = self.buf[self.off:self.off + n]
        self.off += n
        return t


class BlockSaved:
    __slots__ = ("x", "y_a", "bnp_a", "y_b", "bnp_b", "zhat_mean", "hidden", "gate", "y_c", "bnp_c", "y_1", "bnp_1",
                 "out", "stride", "dims")


def block_stat_doubles(N: int, cin: int, ci: int, cout: int, first: bool, has_bn1: bool) -> int:
    cis = ops.pad8(ci)
    n = 2 * cis + 2 * N * cis + 2 * cout
    if first and has_bn1:
        n += 2 * cout
    return n


def res_block_forward(blk, x: torch.Tensor, training: bool, arena: StatArena, save: bool):
    """ResBlock forward (model/x3d.py:235-328).  x: (N,T,H,W,Cin) dense.  Returns (out, BlockSaved|None)."""
    N, T, H, W, Cin = x.shape
    b2 = blk.branch2
    Ci = b2.conv_a.weight.shape[0]
    Cis = ops.pad8(Ci)
    Cout = b2.conv_c.weight.shape[0]
    first = blk.branch1_conv is not None
    stride = 2 if first else 1
    dev = x.device
    M_in = N * T * H * W

    # conv_a (1x1x1) -> raw + stats
    y_a = torch.empty(N, T, H, W, Cis, device=dev, dtype=torch.float32)
    st_a = arena.take(2 * Cis) if training else None
    ops.pw_gemm(ops.operand(x, ld=Cin, OH=H, OW=W), b2.conv_a.weight, w_sr=1, w_so=Cin, Kred=Cin, N=Ci, Ns=Cis,
                M=M_in, Y=y_a, stats=st_a)
    bnp_a = ops.bn_finalize(st_a, 1, M_in, b2.norm_a, Ci, Cis, training)

    # conv_b (depthwise 3x3x3) on relu(bn_a(.)) -> raw + per-sample stats
    st_b = arena.take(2 * N * Cis)
    y_b = ops.dw_conv_fwd(y_a, bnp_a, b2.conv_b.weight, Ci, stride, st_b)
    OH, OW = y_b.shape[2], y_b.shape[3]
    se = b2.norm_b[1] if hasattr(b2.norm_b[1], "block") else None
    bnp_b, zhat_mean, hidden, gate = ops.bn_se_finalize(st_b, N, T * OH * OW, b2.norm_b[0], se, Ci, Cis, training)

    # conv_c (1x1x1) on swish(gate * bn_b(.)) -> raw + stats
    M_out = N * T * OH * OW
    y_c = torch.empty(N, T, OH, OW, Cout, device=dev, dtype=torch.float32)
    st_c = arena.take(2 * Cout) if training else None
    ops.pw_gemm(ops.operand(y_b, ld=Cis, OH=OH, OW=OW, mode=PRO_BN_GATE_SWISH, bnp=bnp_b, gate=gate,
                            frames_per_sample=T),
                b2.conv_c.weight, w_sr=1, w_so=Ci, Kred=Ci, N=Cout, Ns=Cout, M=M_out, Y=y_c, stats=st_c)
    bnp_c = ops.bn_finalize(st_c, 1, M_out, b2.norm_c, Cout, Cout, training)

    # shortcut + join
    y_1 = bnp_1 = None
    if first:
        y_1 = torch.empty(N, T, OH, OW, Cout, device=dev, dtype=torch.float32)
        has_bn1 = blk.branch1_norm is not None
        st_1 = arena.take(2 * Cout) if (training and has_bn1) else None
        ops.pw_gemm(ops.operand(x, ld=Cin, OH=OH, OW=OW, IH=H, IW=W, map_=MAP_SUB2), blk.branch1_conv.weight,
                    w_sr=1, w_so=Cin, Kred=Cin, N=Cout, Ns=Cout, M=M_out, Y=y_1, stats=st_1)
        if has_bn1:
            bnp_1 = ops.bn_finalize(st_1, 1, M_out, blk.branch1_norm, Cout, Cout, training)
            out = ops.bn_add_relu(y_c, bnp_c, y_1, bnp_1)
        else:
            out = ops.bn_add_relu(y_c, bnp_c, y_1, None)
    else:
        out = ops.bn_add_relu(y_c, bnp_c, x, None)

    if not save:
        return out, None
    s = BlockSaved()
    s.x, s.y_a, s.bnp_a, s.y_b, s.bnp_b = x, y_a, bnp_a, y_b, bnp_b
    s.zhat_mean, s.hidden, s.gate = zhat_mean, hidden, gate
    s.y_c, s.bnp_c, s.y_1, s.bnp_1, s.out = y_c, bnp_c, y_1, bnp_1, out
    s.stride = stride
    s.dims = (N, T, H, W, OH, OW, Cin, Ci, Cis, Cout)
    return out, s


def _bump_num_batches_tracked(module) -> None:
    """nn.BatchNorm's `num_batches_tracked += 1` for every BN under `module`, one fused launch."""
    cache = getattr(module, "_c3d_nbt", None)
    if cache is None:
        cache = [m.num_batches_tracked for m in module.modules()
                 if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None]
        object.__setattr__(module, "_c3d_nbt", cache)
    if cache and cache[0].device != next(module.parameters()).device:   # module was moved: refresh
        cache = [m.num_batches_tracked for m in module.modules()
                 if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None]
        object.__setattr__(module, "_c3d_nbt", cache)
    if cache:
        torch._foreach_add_(cache, 1)


def res_stage_forward(stage, x: torch.Tensor, training: bool, save: bool):
    """ResStage forward (model/x3d.py:331-412): sequential ResBlocks."""
    N = x.shape[0]
    if training:
        _bump_num_batches_tracked(stage)
    total = 0
    for i, blk in enumerate(stage.res_blocks):
        ci, cin = blk.branch2.conv_a.weight.shape[0], blk.branch2.conv_a.weight.shape[1]
        cout = blk.branch2.conv_c.weight.shape[0]
        total += block_stat_doubles(N, cin, ci, cout, blk.branch1_conv is not None, blk.branch1_norm is not None)
    arena = StatArena(total, x.device)
    saved: List[BlockSaved] = []
    for blk in stage.res_blocks:
        x, s = res_block_forward(blk, x, training, arena, save)
        if save:
            saved.append(s)
    return x, saved


def stem_forward(stem, frames, B: int, H: int, W: int, training: bool, save: bool):
    """Stem forward (model/x3d.py:23-106).  frames: list of (tensor, stride_n, stride_c)."""
    T = len(frames)
    dev = stem.norm.weight.device
    if training:
        _bump_num_batches_tracked(stem)
    st = torch.zeros(2 * 24, dtype=torch.float64, device=dev) if training else None
    y = ops.stem_fwd(frames, stem.conv.conv_t.weight, stem.conv.conv_xy.weight, B, H, W, st)
    bnp = ops.bn_finalize(st, 1, B * T * H * W, stem.norm, 24, 24, training)
    out = ops.bn_add_relu(y, bnp, None, None)
    return out, ((y, bnp, out) if save else None)


def enhance_forward(x: torch.Tensor, fc_weight: torch.Tensor, P: int, save_mid: bool):
    """Encoder.enhance (model/trainer.py:71-108), in place on frame T//2 of x (N,T,H,W,C):
    x[:, mid] += relu(conv1x1(|x[:, 0] - x[:, P+1]|)).  Returns the pre-enhance copy of the mid frame."""
    N, T, H, W, Cc = x.shape
    mid = T // 2
    fs = T * H * W * Cc
    mid_pre = torch.empty(N, H, W, Cc, device=x.device, dtype=torch.float32) if save_mid else None
    ops.pw_gemm(ops.operand(x[:, 0], ld=Cc, OH=H, OW=W, img_stride=fs, mode=PRO_ABSDIFF, A2=x[:, P + 1],
                            img_stride2=fs),
                fc_weight, w_sr=1, w_so=Cc, Kred=Cc, N=Cc, Ns=Cc, M=N * H * W, Y=x[:, mid], out_img_stride=fs,
                epi=EPI_RELU_ADD, Y2=mid_pre)
    return mid_pre


# ---------------------------------------------------------------------------------------------
# ChangeDecoder (model/change_decoder.py:57-81)
# ---------------------------------------------------------------------------------------------
def feature_view(f: torch.Tensor):
    """Accepts a (B,C,H,W) feature (typically x[:, :, k] of a channels-last-3d stage output) and returns
    (tensor, img_stride) with NHWC element order inside each image; copies only if the layout is not that."""
    B, Cc, H, W = f.shape
    st = f.stride()
    if st[1] == 1 and st[3] == Cc and st[2] == W * Cc:
        return f, st[0]
    g = f.permute(0, 2, 3, 1).contiguous()
    return g, H * W * Cc


def decoder_up_forward(up, c_hi, hi_stride: int, h: int, w: int, skip, skip_stride: int):
    """One up block: skip + ConvTranspose2d(k4,s2,p1)(Conv2d1x1(c_hi)) -> dense (B,2h,2w,Cout).
    The transposed conv is a dense GEMM t x W_all (all 16 kernel positions per input pixel) + a col2im gather."""
    conv1, convt = up[0], up[1]
    cmid, chi = conv1.weight.shape[0], conv1.weight.shape[1]
    cout = convt.weight.shape[1]
    B = c_hi.shape[0]
    dev = conv1.weight.device
    M = B * h * w
    t = torch.empty(B, h, w, cmid, device=dev, dtype=torch.float32)
    ops.pw_gemm(ops.operand(c_hi, ld=chi, OH=h, OW=w, img_stride=hi_stride), conv1.weight, w_sr=1, w_so=chi,
                Kred=chi, N=cmid, Ns=cmid, M=M, Y=t)
    w_all = convt.weight.detach().permute(0, 2, 3, 1).reshape(cmid, 16 * cout).contiguous()   # [ci][(ky,kx,co)]
    U = torch.empty(M, 16 * cout, device=dev, dtype=torch.float32)
    ops.pw_gemm(ops.operand(t, ld=cmid, OH=h, OW=w), w_all, w_sr=16 * cout, w_so=1, Kred=cmid, N=16 * cout,
                Ns=16 * cout, M=M, Y=U)
    out = torch.empty(B, 2 * h, 2 * w, cout, device=dev, dtype=torch.float32)
    ops.convt_col2im(U, skip, skip_stride, convt.bias, out, B, h, w, cout)
    return out, t


def decoder_forward(dec, feats, save: bool):
    """ChangeDecoder.forward (model/change_decoder.py:57-81).  feats = [c1, c2, c3, c4], (B,C,H,W)."""
    c1, s1 = feature_view(feats[0].float())
    c2, s2 = feature_view(feats[1].float())
    c3, s3 = feature_view(feats[2].float())
    c4, s4 = feature_view(feats[3].float())
    h4, w4 = feats[3].shape[2], feats[3].shape[3]
    c3f, t4 = decoder_up_forward(dec.up_c4, c4, s4, h4, w4, c3, s3)
    c2f, t3 = decoder_up_forward(dec.up_c3, c3f, c3f[0].numel(), 2 * h4, 2 * w4, c2, s2)
    c1f, t2 = decoder_up_forward(dec.up_c2, c2f, c2f[0].numel(), 4 * h4, 4 * w4, c1, s1)
    pred = ops.dec_head_fwd(c1f, dec.up_c1[0].weight, dec.has_sigmoid)
    saved = None
    if save:
        saved = dict(c4=(c4, s4), c3f=c3f, c2f=c2f, c1f=c1f, t4=t4, t3=t3, t2=t2, pred=pred, hw4=(h4, w4))
    return pred, saved


# =============================================================================================
# backward
# =============================================================================================
class GradArena:
    """One zero-filled fp32 buffer per backward call; every parameter gradient of the stage is a view
    into it (the wgrad kernels accumulate with atomics, so the memory must start at zero)."""

    def __init__(self, params, device, direct=None):
        self.i = 0
        self.direct = direct is not None
        if direct is not None:
            # views into a persistent flat gradient buffer (change3d_b200.train_step): the kernels accumulate
            # straight into it, autograd gets None for these parameters
            self.views = direct
            return
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]      # keep every view 16-byte aligned
        self.buf = torch.zeros(max(sum(sizes), 4), dtype=torch.float32, device=device)
        self.views = []
        off = 0
        for p, n in zip(params, sizes):
            self.views.append(self.buf[off:off + p.numel()].view(p.shape))
            off += n

    def returned(self):
        """What the autograd Function hands back for the parameters."""
        return [None] * len(self.views) if self.direct else list(self.views)

    def next(self) -> torch.Tensor:
        v = self.views[self.i]
        self.i += 1
        return v


def owned_ndhwc(g: torch.Tensor) -> torch.Tensor:
    """(B,C,T,H,W) gradient -> dense (B,T,H,W,C) tensor that this backward may modify in place."""
    y = g.permute(0, 2, 3, 4, 1)
    return y if y.is_contiguous() else y.contiguous()


def block_bwd_stat_doubles(N: int, ci: int, cout: int, has_bn1: bool) -> int:
    cis = ops.pad8(ci)
    return 2 * cout + (2 * cout if has_bn1 else 0) + 2 * N * cis + 2 * cis


def res_block_backward(blk, s: BlockSaved, dOut: torch.Tensor, arena: StatArena, ga: GradArena, pending: Optional[list] = None):
    """Backward of one ResBlock.  dOut: (N,T,OH,OW,Cout) dense.  Returns dx (N,T,H,W,Cin); parameter
    gradients are written into `ga` in ResBlock.param_list() order.  `pending`: when given, the side stream is NOT
    joined at the end of the block; the tensors its weight-gradient kernels still read are appended to the list and the
    caller joins later (res_stage_backward), so that those kernels may overlap the next blocks' dgrad chain too."""
    N, T, H, W, OH, OW, Cin, Ci, Cis, Cout = s.dims
    b2 = blk.branch2
    dev = dOut.device
    M_out, M_in = N * T * OH * OW, N * T * H * W
    first = blk.branch1_conv is not None
    has_bn1 = first and blk.branch1_norm is not None
    se = b2.norm_b[1] if hasattr(b2.norm_b[1], "block") else None

    # gradient slots in param_list order
    g_w1 = g_g1 = g_b1 = None
    if first:
        g_w1 = ga.next()
        if has_bn1:
            g_g1, g_b1 = ga.next(), ga.next()
    g_wa, g_ga, g_ba, g_wb, g_gb, g_bb = ga.next(), ga.next(), ga.next(), ga.next(), ga.next(), ga.next()
    se_grads = (ga.next(), ga.next(), ga.next(), ga.next()) if se is not None else None
    g_wc, g_gc, g_bc = ga.next(), ga.next(), ga.next()

    # 1. ReLU backward + BN_c (and shortcut BN) reductions
    st_c = arena.take(2 * Cout)
    st_1 = arena.take(2 * Cout) if has_bn1 else None
    d_pre = ops.relu_bwd_stats(dOut, s.out, s.y_c, s.bnp_c, s.y_1 if has_bn1 else None,
                               s.bnp_1 if has_bn1 else None, st_c, st_1)
    coef_c = ops.bn_bwd_finalize(st_c, 1, M_out, Cout, Cout, g_gc, g_bc)

    # 2. conv_c backward: dgrad fused with Swish/SE/BN_b backward prologue of the next stage; wgrad
    P_c = ops.operand(d_pre, ld=Cout, OH=OH, OW=OW, mode=PRO_BNBWD, A2=s.y_c, bnp=s.bnp_c, coef=coef_c)
    du = torch.empty(N, T, OH, OW, Cis, device=dev, dtype=torch.float32)
    st_du = arena.take(2 * N * Cis)
    ops.pw_gemm(P_c, b2.conv_c.weight, w_sr=Ci, w_so=1, Kred=Cout, N=Ci, Ns=Cis, M=M_out, Y=du, epi=EPI_SWISH_BWD,
                stats=st_du, E1=s.y_b, ebnp=s.bnp_b, egate=s.gate, rows_per_sample=T * OH * OW)
    Q_c = ops.operand(s.y_b, ld=Cis, OH=OH, OW=OW, mode=PRO_BN_GATE_SWISH, bnp=s.bnp_b, gate=s.gate,
                      frames_per_sample=T)
    side = _side_stream()
    main = torch.cuda.current_stream()
    if side is not None:
        side.wait_stream(main)                      # coef_c, d_pre are ready
        with torch.cuda.stream(side):
            ops.pw_wgrad(P_c, Q_c, M=M_out, dW=g_wc, dw_sn=Ci, dw_sk=1, N=Cout, K=Ci)
    else:
        ops.pw_wgrad(P_c, Q_c, M=M_out, dW=g_wc, dw_sn=Ci, dw_sk=1, N=Cout, K=Ci)

    # 3. SE backward + BN_b coefficients
    coef_b, dpool = ops.se_bn_bwd_finalize(st_du, N, T * OH * OW, s.bnp_b, b2.norm_b[0], se, s.gate, s.hidden,
                                           s.zhat_mean, Ci, Cis, g_gb, g_bb, se_grads)

    # 4. depthwise conv backward (+ ReLU mask + BN_a reductions)
    st_a = arena.take(2 * Cis)
    dr = ops.dw_conv_bwd(du, s.y_b, s.bnp_b, s.gate if se is not None else None, dpool, coef_b, s.y_a, s.bnp_a,
                         b2.conv_b.weight, Ci, s.stride, st_a, g_wb)
    coef_a = ops.bn_bwd_finalize(st_a, 1, M_in, Ci, Cis, g_ga, g_ba)

    # conv_a weight gradient: needs dr and coef_a only, so with a side stream it is queued before the dgrad kernels
    P_a = ops.operand(dr, ld=Cis, OH=H, OW=W, mode=PRO_BNBWD, A2=s.y_a, bnp=s.bnp_a, coef=coef_a)
    Q_a = ops.operand(s.x, ld=Cin, OH=H, OW=W)
    if side is not None:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.pw_wgrad(P_a, Q_a, M=M_in, dW=g_wa, dw_sn=Cin, dw_sk=1, N=Ci, K=Cin)

    # 5. shortcut conv backward (first block of a stage)
    dx1 = None
    if first:
        if has_bn1:
            coef_1 = ops.bn_bwd_finalize(st_1, 1, M_out, Cout, Cout, g_g1, g_b1)
            P_1 = ops.operand(d_pre, ld=Cout, OH=OH, OW=OW, mode=PRO_BNBWD, A2=s.y_1, bnp=s.bnp_1, coef=coef_1)
        else:
            P_1 = ops.operand(d_pre, ld=Cout, OH=OH, OW=OW)
        dx1 = torch.empty(N, T, OH, OW, Cin, device=dev, dtype=torch.float32)
        ops.pw_gemm(P_1, blk.branch1_conv.weight, w_sr=Cin, w_so=1, Kred=Cout, N=Cin, Ns=Cin, M=M_out, Y=dx1)
        Q_1 = ops.operand(s.x, ld=Cin, OH=OH, OW=OW, IH=H, IW=W, map_=MAP_SUB2)
        ops.pw_wgrad(P_1, Q_1, M=M_out, dW=g_w1, dw_sn=Cin, dw_sk=1, N=Cout, K=Cin)

    # 6. conv_a backward; the epilogue joins the shortcut gradient
    dx = torch.empty(N, T, H, W, Cin, device=dev, dtype=torch.float32)
    ops.pw_gemm(P_a, b2.conv_a.weight, w_sr=Cin, w_so=1, Kred=Ci, N=Cin, Ns=Cin, M=M_in, Y=dx, epi=EPI_ADD2,
                E1=None if first else d_pre, E2=dx1)
    if side is not None:
        if pending is not None:
            pending.append((s, d_pre, coef_c, dr, coef_a))     # keep what the side stream reads alive until the join
        else:
            main.wait_stream(side)                  # join before this block's tensors can be released
    else:
        ops.pw_wgrad(P_a, Q_a, M=M_in, dW=g_wa, dw_sn=Cin, dw_sk=1, N=Ci, K=Cin)
    return dx


def res_stage_backward(stage, saved: List[BlockSaved], g: torch.Tensor):
    """Backward of a ResStage.  g: (N,T,OH,OW,Cout) dense.  Returns (dx, [parameter grads in
    ResStage.param_list() order])."""
    N = g.shape[0]
    params = stage.param_list()
    ga = GradArena(params, g.device, getattr(stage, "_c3d_grad_views", None))
    total = 0
    for blk in stage.res_blocks:
        ci = blk.branch2.conv_a.weight.shape[0]
        cout = blk.branch2.conv_c.weight.shape[0]
        total += block_bwd_stat_doubles(N, ci, cout, blk.branch1_conv is not None and blk.branch1_norm is not None)
    arena = StatArena(total, g.device)
    # the gradient arena is laid out in forward (param_list) order; walk blocks in reverse with per-block cursors
    counts = [len(blk.param_list()) for blk in stage.res_blocks]
    starts = [0]
    for c in counts[:-1]:
        starts.append(starts[-1] + c)
    # side-stream join every C3D_JOIN_EVERY blocks (default 1 = after every block) and at the end of the stage
    join_every = int(os.environ.get("C3D_JOIN_EVERY", "1"))
    side = _side_stream()
    pending: list = []
    for bi in range(len(stage.res_blocks) - 1, -1, -1):
        ga.i = starts[bi]
        defer = side is not None and join_every > 1
        g = res_block_backward(stage.res_blocks[bi], saved[bi], g, arena, ga, pending if defer else None)
        saved[bi] = None          # release this block's activations (deferred joins keep them in `pending`)
        if defer and (len(pending) >= join_every or bi == 0):
            torch.cuda.current_stream().wait_stream(side)
            pending.clear()
    return g, ga.returned()


def stem_masked() -> bool:
    """C3D_STEM_MASKED=1 (default): the stem backward recomputes the ReLU mask from the raw stem output y instead of
    reading the published (enhanced in place) activation, so no masked gradient copy and no restore of the mid frame."""
    return os.environ.get("C3D_STEM_MASKED", "1") != "0"


def enhance_backward(out: torch.Tensor, mid_pre: torch.Tensor, fc_w: torch.Tensor, P: int, g: torch.Tensor,
                     restore: bool = True):
    """Backward of the in-place enhance (model/trainer.py:88-108).  `g` (N,T,H,W,C) is the gradient w.r.t. the
    enhanced tensor and is updated in place to the gradient w.r.t. the stage output (frames 0 and P+1 receive
    -/+ sign(x0-x1) * W^T (g_mid * relu')); finally the saved stage output gets its pre-enhance mid frame back
    so the stage's own ReLU mask is exact (`restore=False`: the caller's backward does not read `out`).  Returns dL/d fc_w."""
    N, T, H, W, Cc = g.shape
    mid = T // 2
    fs = T * H * W * Cc
    M = N * H * W
    x0, x1 = out[:, 0], out[:, P + 1]
    absd = ops.operand(x0, ld=Cc, OH=H, OW=W, img_stride=fs, mode=PRO_ABSDIFF, A2=x1, img_stride2=fs)
    e_pre = torch.empty(N, H, W, Cc, device=g.device, dtype=torch.float32)
    ops.pw_gemm(absd, fc_w, w_sr=1, w_so=Cc, Kred=Cc, N=Cc, Ns=Cc, M=M, Y=e_pre)     # recompute the pre-activation
    de = ops.operand(g[:, mid], ld=Cc, OH=H, OW=W, img_stride=fs, mode=PRO_MASK_POS, A2=e_pre,
                     img_stride2=H * W * Cc)
    ops.pw_gemm(de, fc_w, w_sr=Cc, w_so=1, Kred=Cc, N=Cc, Ns=Cc, M=M, Y=g[:, 0], Y2=g[:, P + 1], out_img_stride=fs,
                epi=EPI_ABSDIFF_BWD, E1=x0, E2=x1, e1_img_stride=fs)
    direct = getattr(fc_w, "_c3d_grad_view", None)
    dfc = direct if direct is not None else torch.zeros_like(fc_w)
    ops.pw_wgrad(de, absd, M=M, dW=dfc, dw_sn=Cc, dw_sk=1, N=Cc, K=Cc)
    if restore:
        out[:, mid].copy_(mid_pre)
    return None if direct is not None else dfc


def stem_backward(stem, frames, y, bnp, out, g, P: int, want_dperc: bool = True):
    """Backward of the stem (+ frame assembly).  g: (B,T,H,W,24) dense gradient w.r.t. the stem output.
    Returns (dperception (1,3,P,H,W) | None, dw_xy, dw_t, dgamma, dbeta)."""
    B, T, H, W, _ = y.shape
    dev = y.device
    w_xy, w_t = stem.conv.conv_t.weight, stem.conv.conv_xy.weight
    ga = GradArena([w_xy, w_t, stem.norm.weight, stem.norm.bias], dev, getattr(stem, "_c3d_grad_views", None))
    dwxy, dwt, dgamma, dbeta = ga.views
    st = torch.zeros(48, dtype=torch.float64, device=dev)
    # C3D_STEM_MASKED=1 (default): the masked gradient is never materialised -- the statistics pass and the stem
    # backward both recompute out > 0 from y (2 x 604 MB less traffic at batch 32, 256 x 256)
    masked = stem_masked()
    if masked:
        ops.relu_bwd_stats(g, None, y, bnp, None, None, st, None, store=False)
        d_pre = g
    else:
        d_pre = ops.relu_bwd_stats(g, out, y, bnp, None, None, st, None)
    coef = ops.bn_bwd_finalize(st, 1, B * T * H * W, 24, 24, dgamma, dbeta)
    dperc = None
    perc_direct = getattr(stem, "_c3d_perc_grad_view", None) if want_dperc else None
    if want_dperc:
        dperc = perc_direct if perc_direct is not None else torch.zeros(1, 3, P, H, W, device=dev, dtype=torch.float32)
    ops.stem_bwd(frames, d_pre, y, bnp, coef, w_xy, w_t, dwxy, dwt, dperc, relu_mask=masked)
    if ga.direct:
        return (None if perc_direct is not None else dperc), None, None, None, None
    return dperc, dwxy, dwt, dgamma, dbeta


def decoder_up_backward(up, d_out, t, c_hi, hi_stride: int, h: int, w: int, ga: GradArena):
    """Backward of one up block (Conv2d 1x1 -> ConvTranspose2d k4 s2 p1 [+ skip]).  d_out: (B,2h,2w,Cout) dense.
    Returns d c_hi (B,h,w,Chi) dense; writes (dW1, dWt, dbias) into `ga`."""
    conv1, convt = up[0], up[1]
    cmid, chi = conv1.weight.shape[0], conv1.weight.shape[1]
    cout = convt.weight.shape[1]
    B = d_out.shape[0]
    dev = d_out.device
    M = B * h * w
    g_w1, g_wt, g_bias = ga.next(), ga.next(), ga.next()
    ops.colsum(d_out, g_bias)
    # V[(j,i)][(ky,kx,co)] = d_out[2j-1+ky][2i-1+kx][co]; then d t = V x W_all^T and d W_all = t^T x V are dense
    V = torch.empty(M, 16 * cout, device=dev, dtype=torch.float32)
    ops.convt_im2col(d_out, V, B, h, w, cout)
    w_all = convt.weight.detach().permute(0, 2, 3, 1).reshape(cmid, 16 * cout).contiguous()   # [ci][(ky,kx,co)]
    d_t = torch.empty(B, h, w, cmid, device=dev, dtype=torch.float32)
    ops.pw_gemm(ops.operand(V, ld=16 * cout, OH=h, OW=w), w_all, w_sr=1, w_so=16 * cout, Kred=16 * cout, N=cmid,
                Ns=cmid, M=M, Y=d_t)
    dWp = torch.zeros(cmid, 16 * cout, device=dev, dtype=torch.float32)
    ops.pw_wgrad(ops.operand(t, ld=cmid, OH=h, OW=w), ops.operand(V, ld=16 * cout, OH=h, OW=w), M=M, dW=dWp,
                 dw_sn=16 * cout, dw_sk=1, N=cmid, K=16 * cout)
    g_wt.copy_(dWp.view(cmid, 4, 4, cout).permute(0, 3, 1, 2))
    # 1x1 conv backward
    d_chi = torch.empty(B, h, w, chi, device=dev, dtype=torch.float32)
    P_t = ops.operand(d_t, ld=cmid, OH=h, OW=w)
    ops.pw_gemm(P_t, conv1.weight, w_sr=chi, w_so=1, Kred=cmid, N=chi, Ns=chi, M=M, Y=d_chi)
    ops.pw_wgrad(P_t, ops.operand(c_hi, ld=chi, OH=h, OW=w, img_stride=hi_stride), M=M, dW=g_w1, dw_sn=chi, dw_sk=1,
                 N=cmid, K=chi)
    return d_chi


def decoder_backward(dec, saved, g: torch.Tensor):
    """Backward of ChangeDecoder.forward.  g: (B,ncls,H,W) contiguous.  Returns ([d c1, d c2, d c3, d c4] as
    (B,C,H,W) views of NHWC tensors, [parameter grads in ChangeDecoder.param_list() order])."""
    params = dec.param_list()
    ga = GradArena(params, g.device, getattr(dec, "_c3d_grad_views", None))
    h4, w4 = saved["hw4"]
    c1f, c2f, c3f = saved["c1f"], saved["c2f"], saved["c3f"]
    c4, s4 = saved["c4"]
    g_head = ga.views[9]
    d_c1f = ops.dec_head_bwd(g.float(), saved["pred"], c1f, dec.up_c1[0].weight, dec.has_sigmoid, g_head)
    ga.i = 6
    d_c2f = decoder_up_backward(dec.up_c2, d_c1f, saved["t2"], c2f, c2f[0].numel(), 4 * h4, 4 * w4, ga)
    ga.i = 3
    d_c3f = decoder_up_backward(dec.up_c3, d_c2f, saved["t3"], c3f, c3f[0].numel(), 2 * h4, 2 * w4, ga)
    ga.i = 0
    d_c4 = decoder_up_backward(dec.up_c4, d_c3f, saved["t4"], c4, s4, h4, w4, ga)
    dfeats = [t.permute(0, 3, 1, 2) for t in (d_c1f, d_c2f, d_c3f, d_c4)]
    return dfeats, ga.returned()
