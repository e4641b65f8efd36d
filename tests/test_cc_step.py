"""Change-captioning training iteration (BASELINE.json config 5; scripts/train_CC.py:105-146, :440-463).

CPU: the graph-capturable loss `packed_caption_loss` equals the script's pack_padded_sequence + CrossEntropyLoss.
GPU: `CCTrainStep` — encoder feature path on the sm_100a kernels, captioning head, packed CE, +-grad_clip clamp and the
two Adam optimizers as fused launches — against the same iteration built from the oracle encoder (fp64 truth, fp32 as
the noise yardstick), the plain-torch restatement of the captioning head in fp64 (oracle/caption_oracle.py, pinned to the
reference's module by tests/test_caption_decoder.py), `clip_gradient` and torch.optim.Adam."""
import argparse
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import caption_oracle as CO
from oracle import change3d_oracle as O
from oracle.make_golden_caption import caption_loss

V, L = 50, 12


def _caps(B, seed):
    g = torch.Generator().manual_seed(seed)
    caps = torch.randint(1, V, (B, L), generator=g)
    lens = torch.randint(4, L + 1, (B, 1), generator=g)
    for b in range(B):
        caps[b, int(lens[b]):] = 0
    return caps, lens


def test_packed_caption_loss_equals_pack_padded_sequence():
    from change3d_b200.train_step import packed_caption_loss
    g = torch.Generator().manual_seed(5)
    for B in (1, 3, 6):
        caps, lens = _caps(B, 40 + B)
        scores = torch.randn(B, L, V, generator=g, dtype=torch.float64, requires_grad=True)
        dl, order = (lens.squeeze(1) - 1).sort(descending=True)
        caps_sorted = caps[order]
        a = caption_loss(scores, caps_sorted, dl.tolist())
        ga, = torch.autograd.grad(a, scores)
        b = packed_caption_loss(scores, caps_sorted, dl)
        gb, = torch.autograd.grad(b, scores)
        assert abs(a.item() - b.item()) < 1e-12 and torch.allclose(ga, gb, atol=1e-14)


def _decoder(args, dtype, device):
    from change3d_b200.model.caption_decoder import CaptionDecoder
    torch.manual_seed(16)
    with contextlib.redirect_stdout(io.StringIO()):
        dec = CaptionDecoder(args)
    dec.position_encoding.dropout.p = 0.0      # built with its own default p = 0.1 whatever args.dropout says
    return dec.to(device=device, dtype=dtype).train()


def _oracle_iteration(full, dec_sd, args, pre, post, caps, lens, dtype, clip):
    """The script's iteration on the host: oracle encoder + oracle captioning head + packed CE, clamp, two Adams."""
    s = O.clone_sd(full, dtype=dtype, requires_grad=True)
    dsd = {}
    for k, v in dec_sd.items():
        t = v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone()
        dsd[k] = t.requires_grad_(True) if (v.is_floating_point() and k != "position_encoding.pe") else t
    feat = O.encoder_forward(s, pre.to(dtype), post.to(dtype), 1, True, output_final=True)
    B, C, H, W = feat.shape
    memory = feat.permute(2, 3, 0, 1).reshape(H * W, B, C)
    scores, caps_sorted, dl, _ = CO.decoder_forward(dsd, memory, caps, lens, args.n_head)     # dropout 0: train == eval
    loss = caption_loss(scores, caps_sorted, dl)
    loss.backward()
    _oracle_iteration.last = (feat.detach().double(), scores.detach().double())
    enc_params = [v for k, v in s.items() if k.startswith("encoder.") and v.requires_grad and v.grad is not None]
    dec_params = [p for p in dsd.values() if p.requires_grad and p.grad is not None]
    grads = {k: v.grad.clone() for k, v in s.items() if v.requires_grad and v.grad is not None}
    grads.update({"decoder." + k: p.grad.clone() for k, p in dsd.items() if p.requires_grad and p.grad is not None})
    for ps in (enc_params, dec_params):
        for p in ps:
            p.grad.data.clamp_(-clip, clip)                                   # clip_gradient, model/utils.py:481-491
        torch.optim.Adam(ps, lr=1e-4, weight_decay=1e-5).step()               # scripts/train_CC.py:440-458
    after = {k: v.detach() for k, v in s.items()}
    after.update({"decoder." + k: p.detach() for k, p in dsd.items()})
    return loss.item(), grads, after


@pytest.mark.gpu
@pytest.mark.parametrize("clip", [5.0, 2e-4])
def test_cc_train_step_vs_oracle(clip):
    from change3d_b200.train_step import CCTrainStep
    from tests.gpu_util import log, rel_err
    from change3d_b200.model.trainer import Trainer
    B, H, W, seed = 2, 64, 64, 29
    args = argparse.Namespace(num_perception_frame=1, num_class=1, in_height=H, in_width=W, dataset="LEVIR-CC",
                              pretrained="/nonexistent", vocab_size=V, embed_dim=192, n_head=8, n_layer=2, dropout=0.0)
    full = O.synth_state_dict(O.trainer_schema("bcd", 1, H, W, 1), seed)
    full = {k: v for k, v in full.items() if k.startswith("encoder.")}
    pre, post, _ = O.synth_inputs(B, H, W, seed)
    caps, lens = _caps(B, seed)
    dec_sd = _decoder(args, torch.float32, "cpu").state_dict()
    l64, g64, a64 = _oracle_iteration(full, dec_sd, args, pre, post, caps, lens, torch.float64, clip)
    feat64, sc64 = _oracle_iteration.last
    l32, g32, _ = _oracle_iteration(full, dec_sd, args, pre, post, caps, lens, torch.float32, clip)
    feat32, sc32 = _oracle_iteration.last

    with contextlib.redirect_stdout(io.StringIO()):
        model = Trainer(args)
    model.encoder.load_state_dict({k[len("encoder."):]: v for k, v in full.items()}, strict=True)
    model.decoder.load_state_dict(dec_sd, strict=True)
    model.decoder.position_encoding.dropout.p = 0.0
    model = model.to("cuda").float()
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    step = CCTrainStep(model, grad_clip=clip)
    loss = step._iteration(pre.cuda(), post.cuda(), caps.cuda(), lens.cuda())
    torch.cuda.synchronize()
    with torch.no_grad():      # forward quantities of the same iteration, for the log
        f_m = model.update_cc(pre.cuda(), post.cuda())
        Bf, Cf, Hf, Wf = f_m.shape
        sc_m = model.decoder.forward_device(f_m.permute(2, 3, 0, 1).reshape(Hf * Wf, Bf, Cf), caps.cuda(), lens.cuda())[0]
    log(f"cc step feat: |mine-fp64| {rel_err(f_m, feat64):.3e} |torch_fp32-fp64| {rel_err(feat32, feat64):.3e};  scores: "
        f"|mine-fp64| {rel_err(sc_m, sc64):.3e} |torch_fp32-fp64| {rel_err(sc32, sc64):.3e}")
    log(f"cc step (clip {clip:g}) loss: mine {loss.item():.6f} fp64 {l64:.6f} torch-fp32 {l32:.6f}")
    assert abs(loss.item() - l64) < max(8 * abs(l32 - l64), 1e-3 * abs(l64))
    named = dict(model.named_parameters())
    keys = ["decoder.wdc.weight", "decoder.vocab_embedding.weight", "decoder.transformer.layers.1.self_attn.in_proj_weight",
            "decoder.transformer.layers.0.multihead_attn2.out_proj.weight", "encoder.perception_frames",
            "encoder.x3d.blocks.4.res_blocks.14.branch2.conv_c.weight", "encoder.x3d.blocks.4.res_blocks.0.branch1_conv.weight",
            "encoder.x3d.blocks.1.res_blocks.0.branch2.conv_a.weight", "encoder.x3d.blocks.0.conv.conv_t.weight"]
    noise = max(rel_err(g32[k], g64[k]) for k in keys)
    bad = []
    for k in keys:
        opt = step.dec_opt if k.startswith("decoder.") else step.opt
        e_mine, e_ref = rel_err(opt.grad_of(named[k]), g64[k]), rel_err(g32[k], g64[k])
        log(f"cc step grad {k}: |mine-fp64| {e_mine:.3e}  |torch_fp32-fp64| {e_ref:.3e}")
        if e_mine > max(8.0 * e_ref, 4.0 * noise, 2e-5):
            bad.append((k, e_mine, e_ref))
    assert not bad, (bad, noise)
    # the update: clamp + Adam, fused, both optimizers
    step._step()
    torch.cuda.synchronize()
    after = model.state_dict()
    lr = 1e-4
    for k in keys:
        # first Adam step moves every element by ~lr * sign(g): elements whose tiny gradient differs in sign between
        # fp32 and fp64 move the other way, so the bound is 2 * lr (+ weight decay), not a relative one
        d = (after[k].double().cpu() - a64[k]).abs().max().item()
        log(f"cc step update {k}: max |mine - fp64 oracle| {d:.3e} (lr {lr:g})")
        assert d <= 2.05 * lr
        frac = ((after[k].double().cpu() - a64[k]).abs() > 0.05 * lr).double().mean().item()
        assert frac < 0.02, (k, frac)
    # parameters without a gradient are untouched (torch.optim.Adam skips grad None: no weight decay either)
    for k, v in after.items():
        if (k.startswith("encoder.fc.") or ".blocks.5." in k or ".self_attn2." in k or ".linear1." in k or ".fc_alpha" in k
                or ".multihead_attn." in k or ".multihead_attn3." in k or ".norm3." in k):
            assert torch.equal(v, before[k]), k


@pytest.mark.gpu
def test_cc_train_step_graph_matches_eager():
    from change3d_b200.train_step import CCTrainStep
    from change3d_b200.model.trainer import Trainer
    from tests.gpu_util import log
    B, H, W, seed = 2, 64, 64, 30
    args = argparse.Namespace(num_perception_frame=1, num_class=1, in_height=H, in_width=W, dataset="LEVIR-CC",
                              pretrained="/nonexistent", vocab_size=V, embed_dim=192, n_head=8, n_layer=3, dropout=0.0)
    pre, post, _ = O.synth_inputs(B, H, W, seed)
    caps, lens = _caps(B, seed)
    traj, stats = [], []
    for use_graph in (False, True):
        torch.manual_seed(16)
        with contextlib.redirect_stdout(io.StringIO()):
            model = Trainer(args).to("cuda").float()
        model.decoder.position_encoding.dropout.p = 0.0
        step = CCTrainStep(model, use_graph=use_graph)
        ls = [step(pre.cuda(), post.cuda(), caps.cuda(), lens.cuda()).item() for _ in range(4)]
        assert all(np.isfinite(ls)) and ls[-1] < ls[0]
        traj.append(ls)
        sd = model.state_dict()
        stats.append((int(sd["encoder.x3d.blocks.4.res_blocks.3.branch2.norm_a.num_batches_tracked"]),
                      sd["encoder.x3d.blocks.1.res_blocks.0.branch2.norm_a.running_mean"].clone()))
    log(f"cc loss trajectory eager {traj[0]} graph {traj[1]}")
    assert abs(traj[0][0] - traj[1][0]) < 1e-4 * abs(traj[0][0])
    assert all(abs(a - b) < 1e-1 * abs(a) for a, b in zip(*traj))      # chaotic drift after step 1, see test_gpu_tasks.py
    # BatchNorm buffers advance once per batch under the graph too (the capture warm-up is rolled back): same update
    # count, and running means that agree to the drift of the trajectories (a second momentum update from the warm-up
    # batch would move them by ~10 % of the batch mean)
    assert stats[0][0] == stats[1][0] == 4
    assert torch.allclose(stats[0][1], stats[1][1], rtol=2e-2, atol=2e-3)
