"""GPU parity tests, backward path + optimizer: hand-written backward kernels through the C ABI
against torch autograd over the CPU oracle, and against the golden gradients produced by the
unmodified reference (oracle/make_golden.py).

Tolerances are relative to the largest magnitude of each reference tensor.  Gradients pass through
~120 batch-statistics BatchNorm backward reductions in fp32 with a different summation order than
ATen, so the end-to-end bound is 2e-3 (the oracle-vs-golden test in tests/test_oracle.py uses the same).
"""
import argparse
import os

import numpy as np
import pytest
import torch

from oracle import change3d_oracle as O
from tests.gpu_util import build_trainer, check, log

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("n,k,M", [(24, 54, 4096), (54, 24, 5000), (96, 216, 3000), (216, 96, 4096), (48, 108, 1000),
                                    (24, 24, 2000)])
def test_pw_wgrad(n, k, M):
    """dW[n][k] = sum_rows P[row][n] * Q[row][k] (tcgen05 path: pixels are the MMA K dimension) vs fp64."""
    from change3d_b200 import ops
    from tests.gpu_util import pad_c
    g = torch.Generator().manual_seed(n * 1000 + k)
    ns, ks = ops.pad8(n), ops.pad8(k)
    P = torch.randn(M, n, generator=g)
    Q = torch.randn(M, k, generator=g)
    ref = P.double().t() @ Q.double()
    Pd, Qd = pad_c(P, ns).to(DEV), pad_c(Q, ks).to(DEV)
    dW = torch.zeros(n, k, device=DEV)
    ops.pw_wgrad(ops.operand(Pd, ld=ns, OH=1, OW=M), ops.operand(Qd, ld=ks, OH=1, OW=M), M=M, dW=dW, dw_sn=k, dw_sk=1,
                 N=n, K=k)
    torch.cuda.synchronize()
    check(f"pw_wgrad {n}x{k} M{M}", dW, ref, 3e-6)


@pytest.mark.parametrize("T,stride,C,H,W,N,se", [(3, 1, 216, 7, 32, 3, True), (3, 1, 108, 40, 64, 2, False), (3, 1, 54, 70, 128, 2, True),
                                                 (5, 1, 108, 9, 32, 2, True), (4, 1, 432, 6, 16, 2, False), (3, 1, 54, 6, 9, 1, True),
                                                 (3, 2, 54, 16, 12, 2, True), (3, 1, 54, 12, 20, 2, False),
                                                 (3, 2, 54, 40, 128, 2, True), (3, 2, 108, 22, 64, 3, False), (4, 2, 216, 12, 34, 2, True)])
def test_dw_conv_backward(T, stride, C, H, W, N, se):
    """conv_b backward through the C ABI (BN_b / SE backward transform of du, transposed depthwise conv, ReLU mask of
    BN_a, depthwise weight gradient, BN_a backward sums) against fp64 autograd of the same formulas
    (reference graph: model/x3d.py:184-211 conv_b -> norm_b (BN + SE) with norm_a/act_a in front)."""
    import torch.nn.functional as F
    from change3d_b200 import ops
    from tests.gpu_util import ndhwc, pad_c
    g = torch.Generator().manual_seed(T * 1000 + C + W)
    cs = ops.pad8(C)
    OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
    ya = torch.randn(N, C, T, H, W, generator=g)
    w = torch.randn(C, 1, 3, 3, 3, generator=g) / 5
    du = torch.randn(N, C, T, OH, OW, generator=g)
    yb = torch.randn(N, C, T, OH, OW, generator=g)

    def bn_block(gen):
        mean, rstd = torch.randn(C, generator=gen) * 0.2, torch.rand(C, generator=gen) + 0.5
        gamma, beta = torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen) * 0.3
        blk = torch.zeros(4, cs)
        blk[0, :C], blk[1, :C], blk[2, :C], blk[3, :C] = mean, rstd, gamma * rstd, beta
        return blk, mean.double(), rstd.double(), (gamma * rstd).double(), beta.double()

    blk_a, mean_a, rstd_a, scale_a, beta_a = bn_block(g)
    blk_b, mean_b, rstd_b, scale_b, _ = bn_block(g)
    coef = torch.zeros(2, cs)
    coef[:, :C] = torch.randn(2, C, generator=g) * 0.1
    gate = torch.rand(N, C, generator=g) + 0.2
    dpool = torch.randn(N, C, generator=g) * 0.05
    bc = lambda v: v.view(1, C, 1, 1, 1)                      # noqa: E731
    bs = lambda v: v.double().view(N, C, 1, 1, 1)             # noqa: E731
    zhat = (yb.double() - bc(mean_b)) * bc(rstd_b)
    inner = du.double() * bs(gate) + bs(dpool) if se else du.double()
    dy = bc(scale_b) * (inner - bc(coef[0, :C].double()) - zhat * bc(coef[1, :C].double()))
    z = ((ya.double() - bc(mean_a)) * bc(scale_a) + bc(beta_a)).requires_grad_(True)
    wd = w.double().requires_grad_(True)
    out = F.conv3d(F.relu(z), wd, None, stride=(1, stride, stride), padding=1, groups=C)
    (out * dy).sum().backward()
    dr_ref, dw_ref = z.grad, wd.grad.view(C, 27)
    yhat_a = (ya.double() - bc(mean_a)) * bc(rstd_a)
    st_ref = torch.stack([dr_ref.sum(dim=(0, 2, 3, 4)), (dr_ref * yhat_a).sum(dim=(0, 2, 3, 4))])

    dW = torch.zeros(C, 27, device=DEV)
    stats = torch.zeros(2 * cs, dtype=torch.float64, device=DEV)
    keep = [pad_c(ndhwc(du), cs).to(DEV), pad_c(ndhwc(yb), cs).to(DEV), blk_b.reshape(-1).to(DEV),
            pad_c(gate, cs).to(DEV) if se else None, pad_c(dpool, cs).to(DEV) if se else None, coef.reshape(-1).to(DEV),
            pad_c(ndhwc(ya), cs).to(DEV), blk_a.reshape(-1).to(DEV), w.to(DEV)]
    dr = ops.dw_conv_bwd(*keep, C, stride, stats, dW)
    torch.cuda.synchronize()
    tag = f"dw_conv_bwd T{T} s{stride} C{C} W{W} se{int(se)}"
    check(tag + " dr", dr[..., :C], ndhwc(dr_ref), 1e-5)
    assert torch.all(dr[..., C:] == 0)
    check(tag + " dW", dW, dw_ref, 2e-5)
    st = stats.view(2, cs)
    check(tag + " sum dr", st[0, :C], st_ref[0], 2e-5)
    check(tag + " sum dr*yhat", st[1, :C], st_ref[1], 2e-5)


def _stage_grads(stage, x, wgt, sd, dtype, depth):
    osd = O.clone_sd(sd, dtype=dtype, requires_grad=True)
    xr = x.detach().clone().to(dtype).requires_grad_(True)
    (O.res_stage(osd, stage, xr, True, depth=depth) * wgt.to(dtype)).sum().backward()
    return xr.grad, osd


@pytest.mark.parametrize("stage,depth,T,H,W,B", [
    (1, None, 3, 16, 16, 2), (2, None, 4, 8, 8, 1), (2, None, 5, 8, 12, 2),      # whole stages (5 / 10 blocks)
    (3, 3, 3, 8, 8, 2), (3, 3, 4, 16, 8, 1), (4, 3, 3, 8, 8, 2),                  # first 3 blocks: Ci = 216 / 432 kernels
    (3, None, 3, 16, 16, 2),                                                      # all 25 blocks (noise-floor bound)
])
def test_res_stage_backward(stage, depth, T, H, W, B):
    """Hand-written backward of a ResStage vs torch autograd over the oracle.  Truth = fp64 oracle; the
    yardstick is torch's own fp32 autograd of the same graph (CPU): each gradient must be within 4x of that
    fp32 noise (floor 2e-5 of the tensor's max).  Tensors whose magnitude is below 1e-6 of the largest
    gradient in the stage are pure cancellation residue and are only required to be small."""
    from change3d_b200.model.x3d import create_x3d
    from tests.gpu_util import check_vs_noise, rel_err
    sd = O.synth_state_dict(O.x3d_schema(), 21)
    cin, _, cout, full_depth = O.STAGES[stage - 1]
    g = torch.Generator().manual_seed(stage * 10 + T)
    x = torch.relu(torch.randn(B, cin, T, H, W, generator=g))
    wgt = torch.randn(B, cout, T, H // 2, W // 2, generator=g)
    dx64, g64 = _stage_grads(stage, x, wgt, sd, torch.float64, depth)
    dx32, g32 = _stage_grads(stage, x, wgt, sd, torch.float32, depth)

    net = create_x3d(input_clip_length=3, depth_factor=5.0)
    net.load_state_dict(sd, strict=True)
    if depth is not None:
        net.blocks[stage].res_blocks = net.blocks[stage].res_blocks[:depth]
    net = net.to(DEV).train()
    xg = x.detach().clone().to(DEV).requires_grad_(True)
    (net.blocks[stage](xg) * wgt.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    tag = f"stage{stage} depth{depth or full_depth} T{T}"
    # 25-block chain: one ReLU-mask flip of a near-zero activation moves dx by ~4e-3 of its maximum.  Whether torch's
    # CPU fp32 run has such a flip depends on the host (oneDNN kernel selection: 4.3e-3 on one box, 4.3e-6 on another
    # for this very seed), so its error cannot be the only yardstick: a flip is a legitimate fp32 outcome, floor 2e-2.
    deep = depth is None and full_depth > 10
    check_vs_noise(tag + " dx", xg.grad, dx64, dx32, floor=2e-2 if deep else 2e-5)
    # Full-depth stages: which near-zero activation flips its ReLU mask differs between any two fp32
    # implementations, so a tensor torch happens to get to 4e-6 can be 1e-3 off here and vice versa.  The
    # yardstick for those runs is the chain's own fp32 noise (torch fp32 vs fp64 on dx), not the per-tensor one.
    chain_noise = rel_err(dx32, dx64) if depth is None else 0.0
    if deep:
        chain_noise = max(chain_noise, 5e-3)
    gmax = max(v.grad.abs().max().item() for k, v in g64.items() if v.grad is not None)
    worst, bad = 0.0, []
    for name, p in net.blocks[stage].named_parameters():
        ref64, ref32 = g64[f"blocks.{stage}.{name}"].grad, g32[f"blocks.{stage}.{name}"].grad
        assert p.grad is not None, name
        scale = ref64.abs().max().item()
        if scale < 1e-6 * gmax:
            assert p.grad.abs().max().item() < 1e-5 * gmax, name
            continue
        e_mine, e_ref = rel_err(p.grad, ref64), rel_err(ref32, ref64)
        worst = max(worst, e_mine)
        if e_mine > max(4.0 * max(e_ref, chain_noise), 2e-5):
            bad.append((name, e_mine, e_ref))
            log(f"  BAD {tag} grad {name}: mine {e_mine:.3e} torch-fp32 {e_ref:.3e}")
    log(f"{tag}: worst parameter-gradient error {worst:.3e}, {len(bad)} outside 4x the fp32 noise")
    if depth is None and full_depth > 10:
        # 25 blocks: a single near-zero activation whose ReLU mask flips differently from torch's fp32 run moves
        # that one block's conv_a / norm_a gradients by percents while everything else stays at the noise floor.
        # Accept at most 2% of the tensors outside the band, none of them beyond 25x the chain noise.
        ntensors = sum(1 for _ in net.blocks[stage].named_parameters())
        assert len(bad) <= 0.02 * ntensors, bad[:5]
        assert all(e < 25.0 * chain_noise for _, e, _ in bad), bad[:5]
    else:
        assert not bad, bad[:5]


@pytest.mark.parametrize("ncls,sig", [(1, True), (7, False)])
def test_change_decoder_backward(ncls, sig):
    from change3d_b200.model.change_decoder import ChangeDecoder
    sd = O.synth_state_dict(O.decoder_schema("", ncls), 31)
    g = torch.Generator().manual_seed(ncls)
    B, h = 2, 3
    feats = [torch.randn(B, c, h * s, (h + 1) * s, generator=g) for c, s in ((24, 8), (24, 4), (48, 2), (96, 1))]
    wgt = torch.randn(B, ncls, h * 8, (h + 1) * 8, generator=g)
    # keep the sigmoid head out of saturation so its gradient is informative
    sd["up_c1.0.weight"] = sd["up_c1.0.weight"] * 0.05

    osd = O.clone_sd(sd, dtype=torch.float64, requires_grad=True)
    fr = [f.double().requires_grad_(True) for f in feats]
    (O.change_decoder(osd, "", fr, sig) * wgt.double()).sum().backward()

    dec = ChangeDecoder(argparse.Namespace(num_class=ncls), in_dim=[24, 24, 48, 96], has_sigmoid=sig)
    dec.load_state_dict(sd, strict=True)
    dec = dec.to(DEV)
    fg = [f.to(DEV).requires_grad_(True) for f in feats]
    (dec(fg) * wgt.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    for i in range(4):
        check(f"decoder ncls{ncls} d c{i + 1}", fg[i].grad, fr[i].grad, 2e-5)
    for name, p in dec.named_parameters():
        check(f"decoder ncls{ncls} grad {name}", p.grad, osd[name].grad, 2e-5)


def test_end_to_end_train_step_vs_golden(golden_dir):
    """update_bcd -> BCEDiceLoss -> backward -> Adam, against the reference's own gradients / updated weights."""
    from change3d_b200.model.utils import BCEDiceLoss
    task, B, H, W, ncls, seed = "bcd", 2, 64, 64, 1, 16
    gold = np.load(os.path.join(golden_dir, "bcd_b2_64.npz"))
    pre, post, target = O.synth_inputs(B, H, W, seed)
    sd = O.calibrate_running_stats(O.synth_state_dict(O.trainer_schema(task, 1, H, W, ncls), seed), task, pre, post)
    model = build_trainer(task, H, W, ncls, sd).train()
    pred = model.update_bcd(pre.to(DEV), post.to(DEV))
    loss = BCEDiceLoss(pred, target.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    log(f"train loss: got {loss.item():.6f} golden {float(gold['train_loss']):.6f}")
    assert abs(loss.item() - float(gold["train_loss"])) < 2e-3 * max(1.0, abs(float(gold["train_loss"])))
    named = dict(model.named_parameters())
    n_none = sum(p.grad is None for p in model.parameters())
    assert n_none == int(gold["grad_none_count"]), (n_none, int(gold["grad_none_count"]))
    # fp64 truth from the oracle (pinned to the reference in tests/test_oracle.py); the golden fp32 gradients
    # of the unmodified reference are the yardstick: mine must be within 8x of the reference's own fp32 error
    # (single-sample noise: encoder gradients differ by ReLU-mask flips of near-zero activations, ~1e-2 either way).
    from tests.gpu_util import rel_err
    sd64 = O.clone_sd(sd, dtype=torch.float64, requires_grad=True)
    l64 = O.bce_dice_loss(O.trainer_forward(sd64, task, pre.double(), post.double(), True), target.double())
    l64.backward()
    bad = []
    for key in gold.files:
        if key.startswith("grad:"):
            gr = named[key[5:]].grad.detach().cpu().numpy()
            t64 = sd64[key[5:]].grad.numpy()
            if gr.size >= 20000:
                gr, t64 = gr.reshape(-1)[::97], t64.reshape(-1)[::97]
            e_mine, e_ref = rel_err(gr, t64), rel_err(gold[key], t64)
            log(f"e2e {key}: |mine-fp64| {e_mine:.3e}  |reference_fp32-fp64| {e_ref:.3e}")
            if e_mine > max(8.0 * e_ref, 2e-5):
                bad.append((key, e_mine, e_ref))
    assert not bad, bad
    # Adam exactly as scripts/train_BCD.py:284-290, through torch.optim.Adam (drop-in) ...
    opt = torch.optim.Adam(model.parameters(), 2e-4, (0.9, 0.99), eps=1e-8, weight_decay=1e-4)
    before = {k: named[k].detach().clone() for k in named}
    grads = {k: (named[k].grad.detach().clone() if named[k].grad is not None else None) for k in named}
    opt.step()
    for key in gold.files:
        if key.startswith("adam:"):
            v = named[key[5:]].detach().cpu().numpy()
            if v.size >= 20000:
                v = v.reshape(-1)[::97]
            check("e2e " + key, v, gold[key], 2e-3)
    # ... and through the fused c3d_adam_step kernel on the same gradients
    from change3d_b200 import ops
    for k in ("encoder.x3d.blocks.1.res_blocks.0.branch2.conv_a.weight", "decoder.up_c4.1.weight",
              "encoder.perception_frames"):
        p0 = before[k].clone().reshape(-1)
        gk = grads[k].reshape(-1).contiguous()
        m = torch.zeros_like(p0)
        v = torch.zeros_like(p0)
        ops.adam_step(p0, gk, m, v, 2e-4, 0.9, 0.99, 1e-8, 1e-4, 1)
        torch.cuda.synchronize()
        check("fused adam " + k, p0, named[k].detach().reshape(-1), 1e-6)


def test_training_reduces_loss():
    """A few real optimisation steps on a fixed batch: the loss must go down (sanity of the whole chain)."""
    from change3d_b200.model.utils import BCEDiceLoss
    task, B, H, W = "bcd", 2, 64, 64
    pre, post, target = O.synth_inputs(B, H, W, 3)
    torch.manual_seed(16)
    model = build_trainer(task, H, W, 1).train()
    opt = torch.optim.Adam(model.parameters(), 2e-4, (0.9, 0.99), eps=1e-8, weight_decay=1e-4)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        loss = BCEDiceLoss(model.update_bcd(pre.to(DEV), post.to(DEV)), target.to(DEV))
        loss.backward()
        opt.step()
        losses.append(loss.item())
    log("loss trajectory: " + " ".join(f"{v:.4f}" for v in losses))
    assert losses[-1] < losses[0]
    assert all(np.isfinite(losses))
