"""Attention of the captioning head on the sm_100a kernels (csrc/attention.cu).

`fused_mha(m, query, key, value, causal)` computes what `nn.MultiheadAttention.forward(query, key, value, attn_mask=
<causal mask or None>, need_weights=False)[0]` computes for the module `m` (model/caption_decoder.py:411-423 calls it
that way for `self_attn` and `multihead_attn2`): the three input projections and the output projection are plain
library GEMMs (F.linear on slices of `in_proj_weight`), the scaled-dot-product core — scores, causal mask, softmax,
attention dropout, weighted sum, and its backward — is one kernel launch per direction instead of six eager kernels.
CUDA tensors only: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib as L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _desc(q, k, v, o, P, keep, keep_scale, nh, causal) -> L.AttnDesc:
    Lq, B, E = q.shape
    d = L.AttnDesc()
    d.q, d.k, d.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    d.q_ls, d.q_bs = q.stride(0), q.stride(1)
    d.k_ls, d.k_bs = k.stride(0), k.stride(1)
    d.v_ls, d.v_bs = v.stride(0), v.stride(1)
    d.o = o.data_ptr() if o is not None else None
    if o is not None:
        d.o_ls, d.o_bs = o.stride(0), o.stride(1)
    d.P = P.data_ptr() if P is not None else None
    d.keep = keep.data_ptr() if keep is not None else None
    d.keep_scale = keep_scale
    d.scale = 1.0 / math.sqrt(E // nh)
    d.B, d.nh, d.hd, d.Lq, d.Lk, d.causal = B, nh, E // nh, Lq, k.shape[0], 1 if causal else 0
    return d


def _check(*ts) -> None:
    for t in ts:
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError("change3d_b200.attention: CUDA float32 tensors required (no CPU/eager fallback)")
        if t.stride(2) != 1:
            raise RuntimeError("change3d_b200.attention: channel stride must be 1")


def attention_forward(q, k, v, nh: int, causal: bool, P: Optional[torch.Tensor] = None,
                      keep: Optional[torch.Tensor] = None, keep_scale: float = 1.0) -> torch.Tensor:
    """q (Lq, B, E), k / v (Lk, B, E) (any position / batch strides) -> o (Lq, B, E)."""
    _check(q, k, v)
    o = torch.empty(q.shape[0], q.shape[1], q.shape[2], device=q.device, dtype=torch.float32)
    d = _desc(q, k, v, o, P, keep, keep_scale, nh, causal)
    L.check(L.load().c3d_attention_fwd(C.byref(d), _stream()), "c3d_attention_fwd")
    return o


class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, nh: int, causal: bool, p_drop: float):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        Lq, B, E = q.shape
        need = any(ctx.needs_input_grad[:3])
        P = torch.empty(B, nh, Lq, k.shape[0], device=q.device, dtype=torch.float32) if need else None
        keep, scale = None, 1.0
        if p_drop > 0.0:
            keep = (torch.rand(B, nh, Lq, k.shape[0], device=q.device) >= p_drop).to(torch.uint8)
            scale = 1.0 / (1.0 - p_drop)
        o = attention_forward(q, k, v, nh, causal, P, keep, scale)
        if need:
            ctx.save_for_backward(q, k, v, P, keep if keep is not None else torch.empty(0, device=q.device))
            ctx.meta = (nh, causal, scale, keep is not None)
        return o

    @staticmethod
    def backward(ctx, go):
        q, k, v, P, keep = ctx.saved_tensors
        nh, causal, scale, has_keep = ctx.meta
        go = go.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        d = _desc(q, k, v, None, P, keep if has_keep else None, scale, nh, causal)
        L.check(L.load().c3d_attention_bwd(C.byref(d), go.data_ptr(), go.stride(0), go.stride(1), dq.data_ptr(),
                                           dk.data_ptr(), dv.data_ptr(), _stream()), "c3d_attention_bwd")
        return dq, dk, dv, None, None, None


def fused_mha(m: torch.nn.MultiheadAttention, query, key, value, causal: bool) -> torch.Tensor:
    """nn.MultiheadAttention(query, key, value, attn_mask=causal mask | None, need_weights=False)[0] for a module with
    packed in-projection, biases and batch_first=False — the configuration of every attention in the captioning head."""
    if not query.is_cuda:
        raise RuntimeError("change3d_b200.attention.fused_mha: CUDA tensors required (no CPU/eager fallback)")
    if m.in_proj_weight is None or m.in_proj_bias is None or m.batch_first or m.bias_k is not None or m.add_zero_attn:
        raise RuntimeError("fused_mha: unsupported nn.MultiheadAttention configuration")
    E = m.embed_dim
    W, b = m.in_proj_weight, m.in_proj_bias
    q = F.linear(query, W[:E], b[:E])
    k = F.linear(key, W[E:2 * E], b[E:2 * E])
    v = F.linear(value, W[2 * E:], b[2 * E:])
    p = float(m.dropout) if m.training else 0.0
    o = _AttnFn.apply(q, k, v, m.num_heads, causal, p)
    return F.linear(o, m.out_proj.weight, m.out_proj.bias)
