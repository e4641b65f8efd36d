"""Op-level wrappers over the C ABI (include/change3d_b200.h).

torch is used for device memory and streams only; every function below launches hand-written
sm_100a kernels from libchange3d_b200.so and raises if the library or a CUDA device is missing.
Activations are fp32 NDHWC torch tensors of shape (N, T, H, W, Cs).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib as L
from ._lib import (EPI_ABSDIFF_BWD, EPI_ADD2, EPI_RELU_ADD, EPI_STORE, EPI_SWISH_BWD, MAP_DENSE, MAP_SUB2,
                   PRO_ABSDIFF, PRO_BN_GATE_SWISH, PRO_BN_RELU, PRO_BNBWD,
                   PRO_MASK_POS, PRO_NONE)

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# Optional per-launch profiler used by bench.py's roofline probe: when PROF is a dict, every wrapper brackets its
# launch with CUDA events on the launching stream and records (events, algorithmic bytes) under the kernel family.
PROF = None


class _Timed:
    """nbytes: every operand tensor of the fused design touched once; nbytes_model: the per-layer byte model of
    SURVEY.md section 8(d) (a conv reads its input once and writes its output once, + weights; BN / activation /
    residual operands are free) — what bench.py's `roofline.frac` is computed from.  Equal unless stated."""

    def __init__(self, family: str, nbytes: int, nbytes_model: Optional[int] = None):
        self.family, self.nbytes = family, nbytes
        self.nbytes_model = nbytes if nbytes_model is None else nbytes_model

    def __enter__(self):
        if PROF is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROF is not None:
            self.e1.record()
            PROF.setdefault(self.family, []).append((self.e0, self.e1, self.nbytes, self.nbytes_model))
        return False


def pad8(c: int) -> int:
    """Channel stride of the bottleneck's inner tensors: a multiple of 8, and of 16 floats (64 bytes, the DRAM access
    granularity) where that costs at most C3D_PAD16_PCT per cent (default 20: 54 -> 64, 216 -> 224; 108 -> 112 and 432
    already are).  With a pixel stride that is not a multiple of 64 bytes the depthwise kernels, which read one
    32-channel block per CTA, straddle DRAM atoms: measured 1.9-2.0x their input bytes from HBM at 56 / 216 channels
    against 1.08x at 64 (profiles/r02_summary.md).  C3D_PAD16=0 keeps the multiple of 8."""
    c8 = (c + 7) // 8 * 8
    c16 = (c + 15) // 16 * 16
    pct = int(os.environ.get("C3D_PAD16_PCT", "20"))
    if c16 != c8 and (c16 - c) * 100 <= pct * c and os.environ.get("C3D_PAD16", "1") == "1":
        return c16
    return c8


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*ts) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("change3d_b200 kernels need CUDA tensors (there is no CPU fallback)")


def operand(A: torch.Tensor, *, ld: int, OH: int, OW: int, IH: Optional[int] = None, IW: Optional[int] = None,
            img_stride: Optional[int] = None, mode: int = PRO_NONE, map_: int = MAP_DENSE,
            A2: Optional[torch.Tensor] = None, img_stride2: int = 0, bnp: Optional[torch.Tensor] = None,
            coef: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None,
            frames_per_sample: int = 1, seg0: int = 0, nseg: int = 0) -> L.Operand:
    IH = OH if IH is None else IH
    IW = OW if IW is None else IW
    if img_stride is None:
        img_stride = IH * IW * ld
    o = L.Operand()
    o.A = _ptr(A); o.A2 = _ptr(A2); o.bnp = _ptr(bnp); o.coef = _ptr(coef); o.gate = _ptr(gate)
    o.mode = mode; o.map = map_; o.ld = ld
    o.OH = OH; o.OW = OW; o.IH = IH; o.IW = IW
    o.img_stride = img_stride; o.img_stride2 = img_stride2
    o.frames_per_sample = frames_per_sample
    o.seg0 = seg0; o.nseg = nseg
    return o


def pw_gemm(a: L.Operand, W: torch.Tensor, *, w_sr: int, w_so: int, Kred: int, N: int, Ns: int, M: int,
            Y: torch.Tensor, epi: int = EPI_STORE, stats: Optional[torch.Tensor] = None, out_img_stride: int = 0,
            E1=None, e1_img_stride: int = 0, E2=None, ebnp=None, egate=None, Y2=None, rows_per_sample: int = 0) -> None:
    _require_cuda(W, Y)
    d = L.GemmDesc()
    d.a = a
    d.W = _ptr(W); d.w_sr = w_sr; d.w_so = w_so; d.w_cls_stride = 0
    d.Kred = Kred; d.N = N; d.Ns = Ns; d.M = M
    d.Y = _ptr(Y); d.out_img_stride = out_img_stride; d.epi = epi; d.stats = _ptr(stats)
    d.E1 = _ptr(E1); d.e1_img_stride = e1_img_stride; d.E2 = _ptr(E2)
    d.ebnp = _ptr(ebnp); d.egate = _ptr(egate); d.bias = None; d.Y2 = _ptr(Y2)
    d.rows_per_sample = rows_per_sample
    # module parameters are constant within a step; derived weight tensors (re-laid-out copies) are not
    d.flags = L.GEMM_W_CONSTANT if isinstance(W, torch.nn.Parameter) else 0
    nbytes = 4 * (M * a.ld * (2 if a.A2 else 1) + M * Ns * (1 + (1 if E1 is not None else 0)) + Kred * N)
    with _Timed("pw_gemm", nbytes, 4 * (M * Kred + M * N + Kred * N)):
        L.check(L.load().c3d_pw_gemm(C.byref(d), _stream()), "c3d_pw_gemm")


def pw_wgrad(p: L.Operand, q: L.Operand, *, M: int, dW: torch.Tensor, dw_sn: int, dw_sk: int, N: int, K: int) -> None:
    _require_cuda(dW)
    d = L.WgradDesc()
    d.p = p; d.q = q; d.M = M; d.dW = _ptr(dW); d.dw_sn = dw_sn; d.dw_sk = dw_sk; d.N = N; d.K = K
    nbytes = 4 * (M * p.ld * (2 if p.A2 else 1) + M * q.ld * (2 if q.A2 else 1) + N * K)
    with _Timed("pw_wgrad", nbytes, 4 * (M * N + M * K + N * K)):
        L.check(L.load().c3d_pw_wgrad(C.byref(d), _stream()), "c3d_pw_wgrad")


def bn_finalize(stats: Optional[torch.Tensor], groups: int, count: int, bn: torch.nn.Module, C_: int, Cs: int,
                training: bool) -> torch.Tensor:
    """stats -> bnp float[4][Cs]; updates bn.running_* in training (nn.BatchNorm3d semantics)."""
    bnp = torch.empty(4 * Cs, device=bn.weight.device, dtype=torch.float32)
    with _Timed("bn_finalize", 0):
        L.check(L.load().c3d_bn_finalize(_ptr(stats), groups, count, _ptr(bn.weight), _ptr(bn.bias),
                                         _ptr(bn.running_mean), _ptr(bn.running_var), C_, Cs,
                                         bn.momentum if bn.momentum is not None else BN_MOMENTUM, bn.eps,
                                         1 if training else 0, _ptr(bnp), _stream()), "c3d_bn_finalize")
    return bnp


def bn_se_finalize(stats: torch.Tensor, N: int, count_per_sample: int, bn: torch.nn.Module, se, C_: int, Cs: int,
                   training: bool):
    """BN_b finalize (+ SE gate when `se` is the SqueezeExcitation module).  Returns
    (bnp, zhat_mean[N,Cs], hidden[N,R] | None, gate[N,Cs] | None)."""
    dev = bn.weight.device
    bnp = torch.empty(4 * Cs, device=dev, dtype=torch.float32)
    zhat_mean = torch.empty(N, Cs, device=dev, dtype=torch.float32)
    if se is not None:
        w1, b1, w2, b2 = se.block[0].weight, se.block[0].bias, se.block[2].weight, se.block[2].bias
        R = w1.shape[0]
        hidden = torch.empty(N, R, device=dev, dtype=torch.float32)
        gate = torch.empty(N, Cs, device=dev, dtype=torch.float32)
    else:
        w1 = b1 = w2 = b2 = hidden = gate = None
        R = 0
    with _Timed("bn_se_finalize", 0):
        L.check(L.load().c3d_bn_se_finalize(_ptr(stats), N, count_per_sample, _ptr(bn.weight), _ptr(bn.bias),
                                            _ptr(bn.running_mean), _ptr(bn.running_var), C_, Cs,
                                            bn.momentum if bn.momentum is not None else BN_MOMENTUM, bn.eps,
                                            1 if training else 0, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), R,
                                            _ptr(bnp), _ptr(zhat_mean), _ptr(hidden), _ptr(gate), _stream()),
                "c3d_bn_se_finalize")
    return bnp, zhat_mean, hidden, gate


def bn_add_relu(A: torch.Tensor, bnpA: torch.Tensor, B: Optional[torch.Tensor], bnpB: Optional[torch.Tensor],
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    Cs = A.shape[-1]
    M = A.numel() // Cs
    Y = torch.empty_like(A) if out is None else out
    with _Timed("bn_add_relu", 4 * M * Cs * (3 if B is not None else 2)):
        L.check(L.load().c3d_bn_add_relu(_ptr(A), _ptr(bnpA), _ptr(B), _ptr(bnpB), _ptr(Y), M, Cs, _stream()),
                "c3d_bn_add_relu")
    return Y


def dw_conv_fwd(X: torch.Tensor, bnp_a: torch.Tensor, w: torch.Tensor, C_: int, stride: int,
                stats: Optional[torch.Tensor]) -> torch.Tensor:
    N, T, IH, IW, Cs = X.shape
    OH, OW = (IH - 1) // stride + 1, (IW - 1) // stride + 1
    Y = torch.empty(N, T, OH, OW, Cs, device=X.device, dtype=torch.float32)
    with _Timed("dw_conv_fwd", 4 * (X.numel() + Y.numel() + 27 * C_)):
        L.check(L.load().c3d_dw_conv_fwd(_ptr(X), _ptr(bnp_a), _ptr(w), _ptr(Y), _ptr(stats), N, T, IH, IW, C_, Cs,
                                         stride, _stream()), "c3d_dw_conv_fwd")
    return Y


def stem_fwd(frames, w_xy: torch.Tensor, w_t: torch.Tensor, B: int, H: int, W: int,
             stats: Optional[torch.Tensor]) -> torch.Tensor:
    """frames: list of T (tensor, stride_n, stride_c) — planes of H*W floats (see c3d_stem_fwd)."""
    T = len(frames)
    ptrs = (C.c_void_p * T)(*[f[0].data_ptr() for f in frames])
    sn = (C.c_longlong * T)(*[f[1] for f in frames])
    sc = (C.c_longlong * T)(*[f[2] for f in frames])
    Y = torch.empty(B, T, H, W, 24, device=w_xy.device, dtype=torch.float32)
    with _Timed("stem_fwd", 4 * (Y.numel() + B * T * 3 * H * W)):
        L.check(L.load().c3d_stem_fwd(ptrs, sn, sc, _ptr(w_xy), _ptr(w_t), _ptr(Y), _ptr(stats), B, T, H, W,
                                      _stream()), "c3d_stem_fwd")
    return Y


def dec_head_fwd(X: torch.Tensor, w: torch.Tensor, apply_sigmoid: bool) -> torch.Tensor:
    B, H, W, C_ = X.shape
    ncls = w.shape[0]
    Y = torch.empty(B, ncls, H, W, device=X.device, dtype=torch.float32)
    with _Timed("dec_head_fwd", 4 * (X.numel() + Y.numel())):
        L.check(L.load().c3d_dec_head_fwd(_ptr(X), _ptr(w), _ptr(Y), B, H, W, C_, ncls, 1 if apply_sigmoid else 0,
                                          _stream()), "c3d_dec_head_fwd")
    return Y


# ---------------------------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------------------------
def relu_bwd_stats(dOut, out, y_c, bnp_c, y_1, bnp_1, stats_c, stats_1, store: bool = True):
    """out=None: the ReLU mask is recomputed from y_c / bnp_c (no shortcut).  store=False: statistics only (returns None)."""
    Cs = dOut.shape[-1]
    M = dOut.numel() // Cs
    d_pre = torch.empty_like(dOut) if store else None
    nt = 2 + (1 if out is not None else 0) + (1 if store else 0) + (1 if y_1 is not None else 0)
    with _Timed("relu_bwd_stats", 4 * M * Cs * nt):
        L.check(L.load().c3d_relu_bwd_stats(_ptr(dOut), _ptr(out), _ptr(y_c), _ptr(bnp_c), _ptr(y_1), _ptr(bnp_1),
                                            _ptr(d_pre), _ptr(stats_c), _ptr(stats_1), M, Cs, _stream()),
                "c3d_relu_bwd_stats")
    return d_pre


def bn_bwd_finalize(stats, groups: int, count: int, C_: int, Cs: int, dgamma: torch.Tensor, dbeta: torch.Tensor):
    coef = torch.empty(2 * Cs, device=stats.device, dtype=torch.float32)
    with _Timed("bn_bwd_finalize", 0):
        L.check(L.load().c3d_bn_bwd_finalize(_ptr(stats), groups, count, C_, Cs, _ptr(coef), _ptr(dgamma), _ptr(dbeta),
                                             _stream()), "c3d_bn_bwd_finalize")
    return coef


def se_bn_bwd_finalize(stats, N: int, count_per_sample: int, bnp, bn, se, gate, hidden, zhat_mean, C_: int, Cs: int,
                       dgamma, dbeta, se_grads):
    dev = stats.device
    coef = torch.empty(2 * Cs, device=dev, dtype=torch.float32)
    if se is not None:
        w1, w2 = se.block[0].weight, se.block[2].weight
        R = w1.shape[0]
        dpool = torch.empty(N, Cs, device=dev, dtype=torch.float32)
        dw1, db1, dw2, db2 = se_grads
    else:
        w1 = w2 = dpool = dw1 = db1 = dw2 = db2 = None
        gate = hidden = zhat_mean = None
        R = 0
    with _Timed("se_bn_bwd_finalize", 0):
        L.check(L.load().c3d_se_bn_bwd_finalize(_ptr(stats), N, count_per_sample, _ptr(bnp), _ptr(bn.weight),
                                                _ptr(bn.bias), _ptr(gate), _ptr(hidden), _ptr(zhat_mean), _ptr(w1),
                                                _ptr(w2), C_, Cs, R, _ptr(coef), _ptr(dgamma), _ptr(dbeta), _ptr(dpool),
                                                _ptr(dw1), _ptr(db1), _ptr(dw2), _ptr(db2), _stream()),
                "c3d_se_bn_bwd_finalize")
    return coef, dpool


def dw_conv_bwd(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, C_: int, stride: int, stats_a, dW) -> torch.Tensor:
    N, T, IH, IW, Cs = y_a.shape
    dr = torch.empty_like(y_a)
    with _Timed("dw_conv_bwd", 4 * (du.numel() + y_b.numel() + y_a.numel() + dr.numel())):
      L.check(L.load().c3d_dw_conv_bwd(_ptr(du), _ptr(y_b), _ptr(bnp_b), _ptr(gate), _ptr(dpool), _ptr(coef_b),
                                     _ptr(y_a), _ptr(bnp_a), _ptr(w), _ptr(dr), _ptr(dW), _ptr(stats_a), N, T, IH, IW,
                                       C_, Cs, stride, _stream()), "c3d_dw_conv_bwd")
    return dr


def colsum(X: torch.Tensor, out: torch.Tensor) -> None:
    Cs = X.shape[-1]
    L.check(L.load().c3d_colsum(_ptr(X), X.numel() // Cs, Cs, _ptr(out), _stream()), "c3d_colsum")


def convt_col2im(U: torch.Tensor, skip, skip_img_stride: int, bias: torch.Tensor, out: torch.Tensor, B: int, h: int,
                 w: int, cout: int) -> None:
    with _Timed("convt_col2im", 4 * (U.numel() + 2 * out.numel())):
        L.check(L.load().c3d_convt_col2im(_ptr(U), _ptr(skip), skip_img_stride, _ptr(bias), _ptr(out), B, h, w, cout,
                                          _stream()), "c3d_convt_col2im")


def convt_im2col(d_out: torch.Tensor, V: torch.Tensor, B: int, h: int, w: int, cout: int) -> None:
    with _Timed("convt_im2col", 4 * (d_out.numel() + V.numel())):
        L.check(L.load().c3d_convt_im2col(_ptr(d_out), _ptr(V), B, h, w, cout, _stream()), "c3d_convt_im2col")


def stem_bwd(frames, d_pre, y, bnp, coef, w_xy, w_t, dwxy, dwt, dperc, relu_mask: bool = False) -> None:
    """relu_mask: `d_pre` is the gradient w.r.t. the ReLU output; the kernel recomputes the mask from y / bnp."""
    T = len(frames)
    B, _, H, W, _ = y.shape
    ptrs = (C.c_void_p * T)(*[f[0].data_ptr() for f in frames])
    sn = (C.c_longlong * T)(*[f[1] for f in frames])
    sc = (C.c_longlong * T)(*[f[2] for f in frames])
    with _Timed("stem_bwd", 4 * (2 * y.numel() + B * T * 3 * H * W)):
      L.check(L.load().c3d_stem_bwd(ptrs, sn, sc, _ptr(d_pre), _ptr(y), _ptr(bnp), _ptr(coef), _ptr(w_xy), _ptr(w_t),
                                  _ptr(dwxy), _ptr(dwt), _ptr(dperc), B, T, H, W, 1 if relu_mask else 0, _stream()), "c3d_stem_bwd")


def dec_head_bwd(dpred, pred, X, w, is_sigmoid: bool, dW) -> torch.Tensor:
    B, H, W, C_ = X.shape
    dX = torch.empty_like(X)
    with _Timed("dec_head_bwd", 4 * (2 * X.numel() + 2 * dpred.numel())):
      L.check(L.load().c3d_dec_head_bwd(_ptr(dpred), _ptr(pred), _ptr(X), _ptr(w), _ptr(dX), _ptr(dW), B, H, W, C_,
                                      w.shape[0], 1 if is_sigmoid else 0, _stream()), "c3d_dec_head_bwd")
    return dX


def adam_step(p, g, m, v, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float, step: int,
              grad_scale: float = 1.0, grad_clip: float = 0.0) -> None:
    """grad_clip > 0: every scaled gradient element is clamped to +-grad_clip first (model/utils.py:481-491)."""
    _require_cuda(p, g, m, v)
    L.check(L.load().c3d_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), lr, beta1, beta2, eps, weight_decay,
                                   step, grad_scale, grad_clip, _stream()), "c3d_adam_step")
