"""GPU parity tests, forward path: every kernel through the C ABI against the CPU oracle
(oracle/change3d_oracle.py) and against the golden vectors produced by the unmodified reference.

Tolerances (fp32, different summation order than ATen/oneDNN): 1e-5 relative-to-max for single
kernels, 2e-4 end-to-end through 40 residual blocks; north_star's bar is 1e-3.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import change3d_oracle as O
from tests.gpu_util import build_trainer, check, log, ndhwc, pad_c

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from change3d_b200 import ops
    return ops


def test_library_loads_on_gpu():
    from change3d_b200 import _lib
    assert _lib.load().c3d_version() == 1


@pytest.mark.parametrize("cin,cout,M", [(24, 54, 1000), (54, 24, 777), (96, 216, 4096), (216, 96, 3000), (192, 432, 600)])
def test_pw_gemm_forward_with_stats(cin, cout, M):
    ops = _ops()
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    cins, couts = ops.pad8(cin), ops.pad8(cout)
    x = torch.randn(M, cin, generator=g)
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    ref = x.double() @ w.double().t()
    xs = pad_c(x, cins).to(DEV).view(1, 1, 1, M, cins)
    y = torch.full((M, couts), float("nan"), device=DEV)
    stats = torch.zeros(2 * couts, dtype=torch.float64, device=DEV)
    ops.pw_gemm(ops.operand(xs, ld=cins, OH=1, OW=M), w.to(DEV), w_sr=1, w_so=cin, Kred=cin, N=cout, Ns=couts, M=M,
                Y=y, stats=stats)
    torch.cuda.synchronize()
    # tcgen05 path: 3-term TF32 split with fp32 accumulation (fp32-class: ~2^-21 per product)
    check(f"pw_gemm {cin}->{cout}", y[:, :cout], ref, 3e-6)
    assert torch.all(y[:, cout:] == 0), "pad lanes must be zero"
    check(f"pw_gemm {cin}->{cout} sum", stats[:cout], ref.sum(0), 3e-6)
    check(f"pw_gemm {cin}->{cout} sumsq", stats[couts:couts + cout], (ref * ref).sum(0), 3e-6)


@pytest.mark.parametrize("T,stride,C,H,W,N", [(3, 1, 54, 12, 20, 2), (3, 2, 54, 16, 12, 2), (4, 1, 108, 8, 8, 1),
                                              (5, 2, 216, 8, 12, 2), (3, 2, 24, 9, 7, 1), (5, 1, 56, 5, 6, 1),
                                              # bench-shaped rows for the row-streaming kernel (32 / 64 / 128 wide: channel
                                              # blocks of 32 / 16 / 8, swizzled ring), enough rows that a CTA's span crosses
                                              # (sample, channel block) segments; odd width -> generic kernel
                                              (3, 1, 216, 7, 32, 3), (3, 1, 108, 40, 64, 2), (3, 1, 54, 70, 128, 2),
                                              (5, 1, 108, 9, 32, 2), (4, 1, 432, 6, 16, 2), (3, 1, 54, 6, 9, 1),
                                              # stride-2 ring kernel: strips of 16 output columns, spans crossing units, odd sizes
                                              (3, 2, 54, 40, 128, 2), (3, 2, 108, 22, 64, 3), (4, 2, 216, 11, 34, 2), (3, 2, 24, 256, 64, 1)])
def test_dw_conv_forward(T, stride, C, H, W, N):
    ops = _ops()
    g = torch.Generator().manual_seed(T * 100 + C)
    cs = ops.pad8(C)
    x = torch.randn(N, C, T, H, W, generator=g)
    w = torch.randn(C, 1, 3, 3, 3, generator=g) / 5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    mean, var = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5
    a = F.relu(F.batch_norm(x.double(), mean.double(), var.double(), gamma.double(), beta.double(), False, 0.0, 1e-5))
    ref = F.conv3d(a, w.double(), None, stride=(1, stride, stride), padding=1, groups=C)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    bnp = torch.zeros(4, cs)
    bnp[0, :C], bnp[1, :C], bnp[2, :C], bnp[3, :C] = mean, rstd, gamma * rstd, beta
    stats = torch.zeros(N * 2 * cs, dtype=torch.float64, device=DEV)
    y = ops.dw_conv_fwd(pad_c(ndhwc(x), cs).to(DEV), bnp.to(DEV).reshape(-1), w.to(DEV), C, stride, stats)
    torch.cuda.synchronize()
    check(f"dw_conv T{T} s{stride} C{C}", y[..., :C], ndhwc(ref), 5e-6)
    assert torch.all(y[..., C:] == 0)
    st = stats.view(N, 2, cs)
    check(f"dw_conv T{T} s{stride} C{C} per-sample sum", st[:, 0, :C], ref.sum(dim=(2, 3, 4)), 5e-6)
    check(f"dw_conv T{T} s{stride} C{C} per-sample sumsq", st[:, 1, :C], (ref * ref).sum(dim=(2, 3, 4)), 5e-6)


def _block_sd(prefix, sd):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


@pytest.mark.parametrize("P", [1, 2, 3])
@pytest.mark.parametrize("training", [False, True])
def test_stem_block(P, training):
    """x3d.blocks[0] on a generic (B,3,T,H,W) clip vs oracle.stem."""
    from change3d_b200.model.x3d import create_x3d
    sd = O.synth_state_dict(O.x3d_schema(), 5)
    net = create_x3d(input_clip_length=3, depth_factor=5.0)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).train(training)
    g = torch.Generator().manual_seed(P)
    x = torch.randn(2, 3, P + 2, 20, 40, generator=g)
    osd = O.clone_sd(sd)
    with torch.no_grad():
        ref = O.stem(osd, x, training)
        got = net.blocks[0](x.to(DEV))
    torch.cuda.synchronize()
    check(f"stem P{P} train{training}", got, ref, 5e-6)
    if training:
        check("stem running_mean", net.blocks[0].norm.running_mean, osd["blocks.0.norm.running_mean"], 1e-5)
        check("stem running_var", net.blocks[0].norm.running_var, osd["blocks.0.norm.running_var"], 1e-5)


@pytest.mark.parametrize("stage,T,H,W,B", [(1, 3, 16, 16, 2), (2, 4, 8, 8, 1), (3, 3, 8, 8, 2), (4, 3, 4, 4, 2)])
@pytest.mark.parametrize("training", [False, True])
def test_res_stage(stage, T, H, W, B, training):
    """x3d.blocks[s] (a whole ResStage: 5/10/25/15 blocks) vs oracle.res_stage, incl. BN buffers."""
    from change3d_b200.model.x3d import create_x3d
    sd = O.synth_state_dict(O.x3d_schema(), 7)
    cin = O.STAGES[stage - 1][0]
    g = torch.Generator().manual_seed(stage)
    x = torch.relu(torch.randn(B, cin, T, H, W, generator=g))
    if not training:
        # calibrate running stats so eval activations stay O(1) through the stage
        tmp = O.clone_sd(sd)
        O._momentum_override.append(1.0)
        with torch.no_grad():
            O.res_stage(tmp, stage, x, True)
        O._momentum_override.pop()
        sd = {k: (torch.zeros((), dtype=torch.int64) if k.endswith("num_batches_tracked") else v) for k, v in tmp.items()}
    net = create_x3d(input_clip_length=3, depth_factor=5.0)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).train(training)
    osd = O.clone_sd(sd)
    with torch.no_grad():
        ref = O.res_stage(osd, stage, x, training)
        got = net.blocks[stage](x.to(DEV))
    torch.cuda.synchronize()
    check(f"res_stage{stage} train{training}", got, ref, 1e-4 if training else 2e-5)
    if training:
        last = O.STAGES[stage - 1][3] - 1
        blk = net.blocks[stage].res_blocks[last].branch2
        p = f"blocks.{stage}.res_blocks.{last}.branch2."
        check("norm_c running_var", blk.norm_c.running_var, osd[p + "norm_c.running_var"], 1e-4)
        check("norm_b running_mean", blk.norm_b[0].running_mean, osd[p + "norm_b.0.running_mean"], 1e-4)
        check("norm_a running_var", blk.norm_a.running_var, osd[p + "norm_a.running_var"], 1e-4)


@pytest.mark.parametrize("P,C", [(1, 24), (2, 48), (3, 96)])
def test_enhance(P, C):
    from change3d_b200 import engine
    g = torch.Generator().manual_seed(P)
    x = torch.randn(2, C, P + 2, 6, 10, generator=g)
    w = torch.randn(C, C, 1, 1, generator=g) / C ** 0.5
    ref = O.enhance(x, w, P)
    xd = ndhwc(x).to(DEV)
    mid_pre = engine.enhance_forward(xd, w.to(DEV), P, True)
    torch.cuda.synchronize()
    check(f"enhance P{P}", xd, ndhwc(ref), 2e-6)
    check(f"enhance P{P} mid_pre", mid_pre, x[:, :, (P + 2) // 2].permute(0, 2, 3, 1), 1e-7)


@pytest.mark.parametrize("ncls,sig", [(1, True), (5, False), (7, False)])
def test_change_decoder(ncls, sig):
    import argparse
    from change3d_b200.model.change_decoder import ChangeDecoder
    sd = O.synth_state_dict(O.decoder_schema("", ncls), 11)
    dec = ChangeDecoder(argparse.Namespace(num_class=ncls), in_dim=[24, 24, 48, 96], has_sigmoid=sig)
    dec.load_state_dict(sd, strict=True)
    dec = dec.to(DEV)
    g = torch.Generator().manual_seed(ncls)
    B, h = 2, 3
    feats = [torch.randn(B, c, h * s, (h + 1) * s, generator=g) for c, s in ((24, 8), (24, 4), (48, 2), (96, 1))]
    with torch.no_grad():
        ref = O.change_decoder(sd, "", feats, sig)
        got = dec([f.to(DEV) for f in feats])                       # NCHW-contiguous inputs (copied to NHWC)
        # strided frame views of a channels-last-3d tensor, as the encoder hands them over
        views = []
        for f in feats:
            full = torch.randn(B, 3, f.shape[2], f.shape[3], f.shape[1], device=DEV)
            full[:, 1] = f.permute(0, 2, 3, 1).to(DEV)
            views.append(full.permute(0, 4, 1, 2, 3)[:, :, 1])
        got2 = dec(views)
    torch.cuda.synchronize()
    check(f"decoder ncls{ncls}", got, ref, 1e-5)
    check(f"decoder ncls{ncls} (frame views)", got2, ref, 1e-5)


def _vs_fp64(name, got, golden, truth64):
    """End-to-end criterion: these synthetic models are ill-conditioned enough that the fp32 REFERENCE sits
    ~2e-3 from the fp64 result at the sigmoid output (measured: 2.4e-3 eval / 2.1e-3 train on bcd_b2_64), so
    the product is required to be as close to the fp64 truth as the reference's own fp32 path, within 3x."""
    got = got.detach().double().cpu()
    golden = torch.as_tensor(np.asarray(golden)).double()
    scale = truth64.abs().max().item() + 1e-30
    e_mine = (got - truth64).abs().max().item() / scale
    e_ref = (golden - truth64).abs().max().item() / scale
    log(f"{name}: |mine-fp64| = {e_mine:.3e}, |reference_fp32-fp64| = {e_ref:.3e} (rel. to max |fp64|)")
    assert e_mine <= max(3.0 * e_ref, 1e-5), f"{name}: {e_mine:.3e} vs reference noise {e_ref:.3e}"


GOLDEN = {"bcd_b2_64": ("bcd", 2, 64, 64, 1, 16), "bda_b1_32": ("bda", 1, 32, 32, 5, 17),
          "scd_b1_32": ("scd", 1, 32, 32, 7, 18)}


@pytest.mark.parametrize("name", list(GOLDEN))
def test_end_to_end_eval_vs_golden(name, golden_dir):
    """Trainer.update_* in eval mode against the vectors produced by the UNMODIFIED reference."""
    task, B, H, W, ncls, seed = GOLDEN[name]
    P = {"bcd": 1, "bda": 2, "scd": 3}[task]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    pre, post, _ = O.synth_inputs(B, H, W, seed)
    sd = O.calibrate_running_stats(O.synth_state_dict(O.trainer_schema(task, P, H, W, ncls), seed), task, pre, post)
    model = build_trainer(task, H, W, ncls, sd).eval()
    with torch.no_grad():
        feats = model.encoder(pre.to(DEV), post.to(DEV))
        out = getattr(model, "update_" + task)(pre.to(DEV), post.to(DEV))
    torch.cuda.synchronize()
    outs = [out] if task == "bcd" else list(out)
    with torch.no_grad():
        t64 = O.trainer_forward(O.clone_sd(sd, dtype=torch.float64), task, pre.double(), post.double(), False)
    t64 = [t64] if task == "bcd" else list(t64)
    for i, o in enumerate(outs):
        _vs_fp64(f"{name} eval pred{i}", o, gold[f"eval_pred{i}"], t64[i])
    for lvl, fl in enumerate(feats):
        for k, f in enumerate(fl):
            # north_star bar: 1e-3 relative; measured margins are logged
            check(f"{name} eval feat l{lvl} p{k}", f[:, :, ::4, ::4], gold[f"eval_feat_l{lvl}_p{k}"], 1e-3)
    if task == "bcd":
        p = outs[0].cpu().numpy()
        ref = gold["eval_pred0"]
        margin = np.abs(ref - 0.5) > 1e-3
        flips = int(((p > 0.5) != (ref > 0.5))[margin].sum())
        log(f"{name}: mask flips outside |p-0.5|<=1e-3: {flips}; pixels inside the margin: {int((~margin).sum())}")
        assert flips == 0


def test_end_to_end_train_forward_vs_golden(golden_dir):
    task, B, H, W, ncls, seed = GOLDEN["bcd_b2_64"]
    gold = np.load(os.path.join(golden_dir, "bcd_b2_64.npz"))
    pre, post, _ = O.synth_inputs(B, H, W, seed)
    sd = O.calibrate_running_stats(O.synth_state_dict(O.trainer_schema(task, 1, H, W, ncls), seed), task, pre, post)
    model = build_trainer(task, H, W, ncls, sd).train()
    with torch.no_grad():
        out = model.update_bcd(pre.to(DEV), post.to(DEV))
    torch.cuda.synchronize()
    with torch.no_grad():
        t64 = O.trainer_forward(O.clone_sd(sd, dtype=torch.float64), task, pre.double(), post.double(), True)
    _vs_fp64("bcd train-mode pred", out, gold["train_pred0"], t64)
    new = model.state_dict()
    for key in gold.files:
        if key.startswith("stat:"):
            check("bcd train " + key, new[key[5:]], gold[key], 5e-4)
    k = "encoder.x3d.blocks.1.res_blocks.0.branch2.norm_a.num_batches_tracked"
    assert int(new[k]) == 1
