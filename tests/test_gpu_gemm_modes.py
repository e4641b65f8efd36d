"""GPU parity of the pointwise-GEMM family through the C ABI, one fused prologue / epilogue at a time, for
every kernel variant the dispatcher can pick (17-warp and compact two-CTA-per-SM tcgen05 kernels, TMA feed or
gathered rows, FFMA fallback).  Truth is the same formula in fp64; the formulas restate what the fused stages
compute in the reference graph (BN apply + ReLU before conv_b' input, SE gate * Swish before conv_c, the BN
backward in front of the transposed convs, |a - b| of enhance, the ReLU mask: reference x3d.py:40-107 /
pytorchvideo ResBlock / BottleneckBlock as restated in oracle/change3d_oracle.py bottleneck()).
"""
import os

import pytest
import torch

from change3d_b200._lib import (EPI_ADD2, EPI_STORE, EPI_SWISH_BWD, PRO_ABSDIFF, PRO_BN_GATE_SWISH, PRO_BN_RELU,
                                PRO_BNBWD, PRO_MASK_POS, PRO_NONE)
from tests.gpu_util import check, pad_c

pytestmark = pytest.mark.gpu
DEV = "cuda"

VARIANTS = {
    "tma_auto": {},                                              # defaults: TMA feed, compact where it fits
    "tma_big": {"C3D_TC_COMPACT": "0"},
    "tma_compact": {"C3D_TC_COMPACT": "2"},
    "gather_big": {"C3D_TC_TMA": "0", "C3D_TC_COMPACT": "0"},
    "gather_compact": {"C3D_TC_TMA": "0", "C3D_TC_COMPACT": "2"},
    "ffma": {"C3D_TC": "0"},
}


class _Env:
    def __init__(self, kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _swish(u):
    return u * torch.sigmoid(u)


def _swish_grad(u):
    s = torch.sigmoid(u)
    return s * (1 + u * (1 - s))


def _bnp(C, Cs, g):
    mean = torch.randn(C, generator=g) * 0.3
    rstd = torch.rand(C, generator=g) + 0.5
    gamma = torch.randn(C, generator=g)
    beta = torch.randn(C, generator=g) * 0.2
    blk = torch.zeros(4, Cs)
    blk[0, :C], blk[1, :C], blk[2, :C], blk[3, :C] = mean, rstd, gamma * rstd, beta
    return blk, mean.double(), rstd.double(), (gamma * rstd).double(), beta.double()


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("mode", [PRO_NONE, PRO_BN_RELU, PRO_BN_GATE_SWISH, PRO_BNBWD, PRO_ABSDIFF, PRO_MASK_POS])
@pytest.mark.parametrize("K,N,M,S", [(24, 54, 40037, 4096), (54, 24, 39000, 3000), (216, 96, 3000, 1000), (96, 216, 41000, 8200),
                                     (108, 48, 38400, 128)])
def test_prologues(variant, mode, K, N, M, S):
    from change3d_b200 import ops
    g = torch.Generator().manual_seed(K * 131 + N * 7 + mode)
    Ks, Ns = ops.pad8(K), ops.pad8(N)
    nsamp = (M + S - 1) // S
    a = torch.randn(M, K, generator=g)
    a2 = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    blk, mean, rstd, scale, beta = _bnp(K, Ks, g)
    gate = torch.rand(nsamp, K, generator=g)
    coef = torch.zeros(2, Ks)
    coef[:, :K] = torch.randn(2, K, generator=g) * 0.1
    ad, a2d = a.double(), a2.double()
    samp = torch.arange(M) // S
    if mode == PRO_NONE:
        x = ad
    elif mode == PRO_BN_RELU:
        x = torch.relu((ad - mean) * scale + beta)
    elif mode == PRO_BN_GATE_SWISH:
        x = _swish(((ad - mean) * scale + beta) * gate.double()[samp])
    elif mode == PRO_BNBWD:
        x = scale * (ad - coef[0, :K].double() - (a2d - mean) * rstd * coef[1, :K].double())
    elif mode == PRO_ABSDIFF:
        x = (ad - a2d).abs()
    else:
        x = torch.where(a2d > 0, ad, torch.zeros_like(ad))
    ref = x @ w.double().t()
    A = pad_c(a, Ks).to(DEV)
    A2 = pad_c(a2, Ks).to(DEV)
    y = torch.full((M, Ns), float("nan"), device=DEV)
    stats = torch.zeros(2 * Ns, dtype=torch.float64, device=DEV)
    # operands carry raw pointers: every device tensor must stay referenced until the launch has run
    bnp_d, coef_d, gate_d, w_d = blk.to(DEV), coef.to(DEV), pad_c(gate, Ks).to(DEV), w.to(DEV)
    with _Env(VARIANTS[variant]):
        op = ops.operand(A, ld=Ks, OH=1, OW=S, mode=mode, A2=A2 if mode in (PRO_BNBWD, PRO_ABSDIFF, PRO_MASK_POS) else None,
                         bnp=bnp_d, coef=coef_d, gate=gate_d if mode == PRO_BN_GATE_SWISH else None, frames_per_sample=1)
        ops.pw_gemm(op, w_d, w_sr=1, w_so=K, Kred=K, N=N, Ns=Ns, M=M, Y=y, stats=stats)
        torch.cuda.synchronize()
    tag = f"{variant} mode{mode} {K}->{N} M{M}"
    check(f"gemm {tag}", y[:, :N], ref, 1e-5)
    assert torch.all(y[:, N:] == 0), "pad lanes must be zero"
    check(f"gemm {tag} sum", stats[:N], ref.sum(0), 1e-5)
    check(f"gemm {tag} sumsq", stats[Ns:Ns + N], (ref * ref).sum(0), 1e-5)


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("K,N,M,S", [(24, 54, 40037, 4096), (96, 216, 3000, 1000), (48, 108, 38400, 1280), (96, 216, 41000, 8200)])
def test_swish_backward_epilogue(variant, K, N, M, S):
    """dY (K channels) x W_c -> dA, then du = dA * Swish'(u), u = gate * BN_b(y): Y = du and the per-sample
    sums of du and du * zhat that the BN_b / SE backward finalizer needs."""
    from change3d_b200 import ops
    g = torch.Generator().manual_seed(K * 977 + N)
    Ks, Ns = ops.pad8(K), ops.pad8(N)
    nsamp = (M + S - 1) // S
    dy = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / K ** 0.5          # conv_c weight [Cout=K, Ci=N]; dA = dy @ w
    yb = torch.randn(M, N, generator=g)
    blk, mean, rstd, scale, beta = _bnp(N, Ns, g)
    gate = torch.rand(nsamp, N, generator=g)
    samp = torch.arange(M) // S
    dA = dy.double() @ w.double()
    u = ((yb.double() - mean) * scale + beta) * gate.double()[samp]
    du = dA * _swish_grad(u)
    dz = du * (yb.double() - mean) * rstd
    ref_stats = torch.zeros(nsamp, 2, N, dtype=torch.float64)
    ref_stats[:, 0].index_add_(0, samp, du)
    ref_stats[:, 1].index_add_(0, samp, dz)
    Y = torch.full((M, Ns), float("nan"), device=DEV)
    stats = torch.zeros(nsamp, 2, Ns, dtype=torch.float64, device=DEV)
    dy_d, w_d, yb_d, bnp_d, gate_d = pad_c(dy, Ks).to(DEV), w.to(DEV), pad_c(yb, Ns).to(DEV), blk.to(DEV), pad_c(gate, Ns).to(DEV)
    with _Env(VARIANTS[variant]):
        op = ops.operand(dy_d, ld=Ks, OH=1, OW=S, frames_per_sample=1)
        ops.pw_gemm(op, w_d, w_sr=N, w_so=1, Kred=K, N=N, Ns=Ns, M=M, Y=Y, epi=EPI_SWISH_BWD, stats=stats,
                    E1=yb_d, ebnp=bnp_d, egate=gate_d, rows_per_sample=S)
        torch.cuda.synchronize()
    tag = f"{variant} swish_bwd {K}->{N} M{M}"
    check(f"gemm {tag} du", Y[:, :N], du, 1e-5)
    check(f"gemm {tag} sum du", stats[:, 0, :N], ref_stats[:, 0], 2e-5)
    check(f"gemm {tag} sum du*zhat", stats[:, 1, :N], ref_stats[:, 1], 2e-5)


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("K,N,nimg,OH,OW", [(54, 24, 500, 8, 10), (216, 96, 40, 8, 8), (108, 48, 600, 8, 8)])
def test_add2_epilogue(variant, K, N, nimg, OH, OW):
    """Residual join of the block backward: dX = BN_a-backward(dy_a) x W_a + dX_skip (+ the stride-2 shortcut
    gradient scattered onto the even pixels)."""
    from change3d_b200 import ops
    g = torch.Generator().manual_seed(K * 31 + N + nimg)
    Ks, Ns = ops.pad8(K), ops.pad8(N)
    M = nimg * OH * OW
    a = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / K ** 0.5
    e1 = torch.randn(M, N, generator=g)
    e2 = torch.randn(nimg, OH // 2, OW // 2, N, generator=g)
    ref = (a.double() @ w.double() + e1.double()).view(nimg, OH, OW, N).clone()
    ref[:, ::2, ::2] += e2.double()
    Y = torch.full((M, Ns), float("nan"), device=DEV)
    a_d, w_d, e1_d, e2_d = pad_c(a, Ks).to(DEV), w.to(DEV), pad_c(e1, Ns).to(DEV), pad_c(e2.reshape(-1, N), Ns).to(DEV)
    with _Env(VARIANTS[variant]):
        op = ops.operand(a_d, ld=Ks, OH=OH, OW=OW, frames_per_sample=1)
        ops.pw_gemm(op, w_d, w_sr=N, w_so=1, Kred=K, N=N, Ns=Ns, M=M, Y=Y, epi=EPI_ADD2, E1=e1_d, E2=e2_d)
        torch.cuda.synchronize()
    check(f"gemm {variant} add2 {K}->{N} M{M}", Y[:, :N], ref.view(M, N), 1e-5)
