"""CPU tests of the host-side mirror of the reference interface (no kernels run)."""
import argparse
import contextlib
import io

import pytest
import torch
import torch.nn.functional as F

from oracle import change3d_oracle as O
from oracle import reference_loader as R


def _trainer(task, H, W, ncls):
    from change3d_b200.model.trainer import Trainer
    P = {"bcd": 1, "bda": 2, "scd": 3}[task]
    a = argparse.Namespace(num_perception_frame=P, num_class=ncls, in_height=H, in_width=W,
                           dataset={"bcd": "LEVIR-CD", "bda": "xBD", "scd": "SECOND"}[task], pretrained="/nonexistent")
    with contextlib.redirect_stdout(io.StringIO()):
        return Trainer(a)


@pytest.mark.parametrize("task,ncls", [("bcd", 1), ("bda", 5), ("scd", 7)])
def test_state_dict_schema_matches_reference(task, ncls):
    """Same keys, order and shapes as the reference's Trainer -> checkpoints interchange."""
    P = {"bcd": 1, "bda": 2, "scd": 3}[task]
    m = _trainer(task, 32, 32, ncls)
    schema = O.trainer_schema(task, P, 32, 32, ncls)
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _ in schema]
    assert all(tuple(sd[k].shape) == tuple(s) for k, s in schema)
    m.load_state_dict(O.synth_state_dict(schema, 3), strict=True)
    # parameters() order is what torch.optim.Adam and checkpointed optimizer state index by
    assert [n for n, _ in m.named_parameters()] == [k for k, _ in schema if "running_" not in k and "num_batches" not in k]


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted")
def test_matches_reference_modules_directly():
    ref = R.build_trainer("bcd", 32, 32, 1)
    mine = _trainer("bcd", 32, 32, 1)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict(), strict=True)      # reference checkpoint -> product
    ref.load_state_dict(mine.state_dict(), strict=True)      # product checkpoint -> reference
    # weight_init: same traversal and initialisers => identical tensors from the same RNG state
    import model.utils as ref_utils   # the reference's file (path set up by reference_loader)
    from change3d_b200.model.change_decoder import ChangeDecoder
    from change3d_b200.model.utils import weight_init
    a = argparse.Namespace(num_class=5)
    torch.manual_seed(5)
    d1 = ChangeDecoder(a, in_dim=[24, 24, 48, 96])
    d2 = ChangeDecoder(a, in_dim=[24, 24, 48, 96])
    d2.load_state_dict(d1.state_dict())
    torch.manual_seed(7)
    ref_utils.weight_init(d1)
    torch.manual_seed(7)
    weight_init(d2)
    for (k1, v1), (k2, v2) in zip(d1.state_dict().items(), d2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1


def test_param_lists_cover_every_trained_parameter():
    m = _trainer("bcd", 32, 32, 1)
    enc = m.encoder
    ids = set()
    stem = enc.x3d.blocks[0]
    ids |= {id(p) for p in (stem.conv.conv_t.weight, stem.conv.conv_xy.weight, stem.norm.weight, stem.norm.bias)}
    for i in range(1, 4):
        pl = enc.x3d.blocks[i].param_list()
        assert len(pl) == len(list(enc.x3d.blocks[i].parameters()))
        assert [id(p) for p in pl] == [id(p) for p in enc.x3d.blocks[i].parameters()]   # registration order
        ids |= {id(p) for p in pl}
    ids |= {id(fc[0].weight) for fc in enc.fc} | {id(enc.perception_frames)} | {id(p) for p in m.decoder.param_list()}
    from change3d_b200.train_step import _trained_parameters
    trained = _trained_parameters(m)
    assert {id(p) for p in trained} == ids
    n = sum(p.numel() for p in trained) - enc.perception_frames.numel()
    assert n == 1_542_656                                     # published BCD parameter count


def test_flat_adam_buffers_on_cpu():
    """FlatAdam re-homes the trained parameters into one flat buffer and installs gradient views (no CUDA needed)."""
    from change3d_b200.train_step import FlatAdam
    m = _trainer("bcd", 32, 32, 1)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    opt = FlatAdam(m)
    after = m.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before)
    p = m.encoder.x3d.blocks[2].res_blocks[3].branch2.conv_a.weight
    assert p.data_ptr() >= opt.flat_p.data_ptr() and p.data_ptr() < opt.flat_p.data_ptr() + opt.flat_p.numel() * 4
    gv = m.encoder.x3d.blocks[2]._c3d_grad_views
    assert len(gv) == len(m.encoder.x3d.blocks[2].param_list())
    gv[5].fill_(1.0)
    assert opt.flat_g.sum().item() == gv[5].numel()
    opt.zero_grad()
    assert opt.flat_g.abs().sum().item() == 0
    # blocks.4 / blocks.5 (never trained by BCD) stay outside the flat buffers
    q = m.encoder.x3d.blocks[4].res_blocks[0].branch2.conv_a.weight
    assert not (opt.flat_p.data_ptr() <= q.data_ptr() < opt.flat_p.data_ptr() + opt.flat_p.numel() * 4)


def test_convtranspose_as_dense_gemm_plus_col2im_reproduces_conv_transpose2d():
    """The decomposition engine.decoder_up_forward uses (U = t x W_all with W_all[ci][(ky,kx,co)] =
    weight.permute(0, 2, 3, 1), then c3d_convt_col2im: out[y][x] = bias + the <= 4 entries of U with 2j-1+ky = y,
    2i-1+kx = x) equals F.conv_transpose2d(k=4, s=2, p=1) — host restatement of the index arithmetic."""
    cin, cout, h, w = 5, 3, 4, 6
    g = torch.Generator().manual_seed(0)
    wt = torch.randn(cin, cout, 4, 4, generator=g)
    bias = torch.randn(cout, generator=g)
    x = torch.randn(1, cin, h, w, generator=g)
    ref = F.conv_transpose2d(x, wt, bias, stride=2, padding=1)
    w_all = wt.permute(0, 2, 3, 1).reshape(cin, 16 * cout)
    U = (x[0].permute(1, 2, 0).reshape(h * w, cin) @ w_all).view(h, w, 4, 4, cout)
    out = bias.view(1, 1, cout).repeat(2 * h, 2 * w, 1)
    for y in range(2 * h):
        for xx in range(2 * w):
            for ky in range(4):
                for kx in range(4):
                    jn, in_ = y + 1 - ky, xx + 1 - kx
                    if jn % 2 == 0 and in_ % 2 == 0 and 0 <= jn // 2 < h and 0 <= in_ // 2 < w:
                        out[y, xx] += U[jn // 2, in_ // 2, ky, kx]
    assert torch.allclose(out.permute(2, 0, 1).unsqueeze(0), ref, atol=1e-5)


def test_convtranspose_as_gemm_col2im_im2col_matches_torch():
    """The decoder's ConvTranspose2d(k4, s2, p1) formulation (engine.decoder_up_forward / _backward):
    U = t x W_all over the input grid, out = bias + skip + col2im(U); backward V = im2col(d_out),
    d t = V x W_all^T, d W_all = t^T x V.  Restated here with torch index arithmetic exactly as the kernels in
    csrc/decoder.cu index (2j-1+ky = y, 2i-1+kx = x) and compared with F.conv_transpose2d and its autograd
    (reference: model/change_decoder.py:30-45,71-73)."""
    cmid, cout, B, h, w = 6, 4, 2, 3, 5
    g = torch.Generator().manual_seed(3)
    wt = torch.randn(cmid, cout, 4, 4, generator=g).double().requires_grad_(True)
    bias = torch.randn(cout, generator=g).double()
    t = torch.randn(B, cmid, h, w, generator=g).double().requires_grad_(True)
    skip = torch.randn(B, cout, 2 * h, 2 * w, generator=g).double()
    d_out = torch.randn(B, cout, 2 * h, 2 * w, generator=g).double()
    ref = F.conv_transpose2d(t, wt, bias, stride=2, padding=1) + skip
    ref.backward(d_out)

    w_all = wt.detach().permute(0, 2, 3, 1).reshape(cmid, 16 * cout)            # [ci][(ky,kx,co)]
    tn = t.detach().permute(0, 2, 3, 1)                                          # NHWC
    U = (tn.reshape(-1, cmid) @ w_all).view(B, h, w, 4, 4, cout)
    out = (bias.view(1, 1, 1, cout) + skip.permute(0, 2, 3, 1)).clone()
    for y in range(2 * h):
        for x in range(2 * w):
            for a in range(2):
                ky = ((y + 1) & 1) + 2 * a
                if y + 1 - ky < 0 or (y + 1 - ky) // 2 >= h:
                    continue
                for e in range(2):
                    kx = ((x + 1) & 1) + 2 * e
                    if x + 1 - kx < 0 or (x + 1 - kx) // 2 >= w:
                        continue
                    out[:, y, x] += U[:, (y + 1 - ky) // 2, (x + 1 - kx) // 2, ky, kx]
    assert torch.allclose(out.permute(0, 3, 1, 2), ref.detach(), atol=1e-10)

    dn = d_out.permute(0, 2, 3, 1)
    V = torch.zeros(B, h, w, 4, 4, cout, dtype=torch.float64)
    for ky in range(4):
        for kx in range(4):
            for j in range(h):
                for i in range(w):
                    y, x = 2 * j - 1 + ky, 2 * i - 1 + kx
                    if 0 <= y < 2 * h and 0 <= x < 2 * w:
                        V[:, j, i, ky, kx] = dn[:, y, x]
    Vm = V.reshape(-1, 16 * cout)
    d_t = (Vm @ w_all.t()).view(B, h, w, cmid).permute(0, 3, 1, 2)
    d_w = (tn.reshape(-1, cmid).t() @ Vm).view(cmid, 4, 4, cout).permute(0, 3, 1, 2)
    assert torch.allclose(d_t, t.grad, atol=1e-10)
    assert torch.allclose(d_w, wt.grad, atol=1e-10)


def test_lr_schedule_matches_oracle():
    from change3d_b200.model.utils import adjust_learning_rate
    args = argparse.Namespace(lr_mode="poly", lr=2e-4, max_epochs=10, step_loss=None)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    for epoch, it in ((0, 0), (0, 150), (0, 250), (3, 5000)):
        lr = adjust_learning_rate(args, opt, epoch, it, 800)
        assert abs(lr - O.poly_lr(2e-4, it, 8000, epoch)) < 1e-15
        assert opt.param_groups[0]["lr"] == lr


def test_round_helpers_and_depths():
    from change3d_b200.model.x3d import create_x3d, round_repeats, round_width
    assert [round_width(12, 2.0), round_width(24, 2.0, divisor=8), round_width(54, 0.0625)] == [24, 48, 8]
    assert [round_repeats(r, 5.0) for r in (1, 2, 5, 3)] == [5, 10, 25, 15]
    net = create_x3d(input_clip_length=3, depth_factor=5.0)
    assert len(net.blocks) == 6
    assert [len(net.blocks[i].res_blocks) for i in range(1, 5)] == [5, 10, 25, 15]
    assert sum(p.numel() for p in net.parameters()) == 6_153_384


def test_flat_adam_param_groups_state_dict_and_task_views():
    """FlatAdam looks like torch's Adam to `adjust_learning_rate` (param_groups), round-trips its state for the
    reference's checkpoint dict, and installs gradient views for every decoder head of the SCD / BDA models."""
    from argparse import Namespace
    from change3d_b200.model.utils import adjust_learning_rate
    from change3d_b200.train_step import FlatAdam
    m = _trainer("scd", 32, 32, 7)
    opt = FlatAdam(m, lr=2e-4)
    args = Namespace(lr=2e-4, lr_mode="poly", max_epochs=3, step_loss=100)
    lr = adjust_learning_rate(args, opt, 0, 10, 50)
    assert opt.param_groups[0]["lr"] == lr and abs(lr - (2e-4 * 0.9 * 11 / 200 + 0.1 * 2e-4)) < 1e-12
    for name in ("decoder_pre", "decoder_post", "decoder_change"):
        dec = getattr(m, name)
        assert len(dec._c3d_grad_views) == len(dec.param_list()) == 10
        assert all(v.shape == p.shape for v, p in zip(dec._c3d_grad_views, dec.param_list()))
    opt.step_count = 7
    opt.m.fill_(0.5)
    sd = opt.state_dict()
    assert sd["state"]["step"] == 7 and sd["param_groups"][0]["lr"] == lr and sd["param_groups"][0]["betas"] == (0.9, 0.99)
    opt2 = FlatAdam(_trainer("scd", 32, 32, 7))
    opt2.load_state_dict(sd)
    assert opt2.step_count == 7 and torch.equal(opt2.m, opt.m) and opt2.param_groups[0]["lr"] == lr
    with pytest.raises(ValueError):
        from change3d_b200.train_step import TrainStep
        TrainStep(m, task="cc")


def test_launch_summary_tool_reproduces_the_committed_traffic_table():
    """profiles/tools/launch_summary.py on the committed end-of-round ncu launch list: one step = 752 launches, and the
    per-family DRAM traffic equals profiles/r01b_traffic.json (what bench.py reports as roofline.traffic)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = os.path.join(root, "profiles", "r01b_launches_final.csv")
    out = subprocess.run([sys.executable, os.path.join(root, "profiles", "tools", "launch_summary.py"), csv_path,
                          "--one-step", "--json", os.path.join(str(os.environ.get("TMPDIR", "/tmp")), "c3d_traffic.json")],
                         capture_output=True, text=True, check=True).stdout
    assert "period = 752 launches per step" in out
    got = json.load(open(os.path.join(str(os.environ.get("TMPDIR", "/tmp")), "c3d_traffic.json")))["families"]
    want = json.load(open(os.path.join(root, "profiles", "r01b_traffic.json")))["families"]
    assert got == want
    assert want["pw_gemm"]["launches"] == 190 and 295e6 < want["pw_gemm"]["dram_bytes_per_launch"] < 330e6
