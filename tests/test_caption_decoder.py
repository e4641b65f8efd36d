"""Captioning head.

CPU: the module tree / state-dict schema of change3d_b200's CaptionDecoder against tests/golden/caption_decoder.npz (the
reference's own module: oracle/make_golden_caption.py), and the plain-torch restatement oracle/caption_oracle.py against
the same file — that pins the oracle the GPU tests check the kernels with.
GPU: CaptionDecoder forward + loss + gradients through the attention kernels (csrc/attention.cu) against the golden
file; the kernels alone against torch's scaled-dot-product arithmetic (with dropout masks); the cached batched caption
search against the restatement of scripts/train_CC.py's evaluate loop, token for token."""
import argparse
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from oracle import caption_oracle as CO
from oracle.make_golden_caption import ARGS, caption_loss, inputs


def _golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "caption_decoder.npz"))
    return g, {k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd:")}


def _module(sd):
    from change3d_b200.model.caption_decoder import CaptionDecoder
    with contextlib.redirect_stdout(io.StringIO()):
        dec = CaptionDecoder(argparse.Namespace(**ARGS))
    assert list(sd) == list(dec.state_dict())                       # same keys, same registration order
    dec.load_state_dict(sd, strict=True)
    return dec


def test_module_tree_matches_reference_and_has_no_cpu_path(golden_dir):
    g, sd = _golden(golden_dir)
    dec = _module(sd).eval()
    named = dict(dec.named_parameters())
    live = {id(p) for p in dec.live_parameters()}
    assert sorted(k for k, p in named.items() if id(p) not in live) == g["unused_grad_none"].tolist()
    memory, caps, lens = inputs()
    with pytest.raises(RuntimeError, match="no CPU"):
        dec(memory, caps, lens)


def test_oracle_restatement_matches_reference_module(golden_dir):
    g, sd = _golden(golden_dir)
    memory, caps, lens = inputs()
    pred, caps_sorted, decode_lengths, sort_ind = CO.decoder_forward(sd, memory, caps, lens, ARGS["n_head"])
    assert np.array_equal(sort_ind.numpy(), g["sort_ind"]) and decode_lengths == g["decode_lengths"].tolist()
    np.testing.assert_allclose(pred.numpy(), g["pred"], rtol=1e-4, atol=2e-5)
    assert abs(caption_loss(pred, caps_sorted, decode_lengths).item() - float(g["loss"])) < 1e-5 * float(g["loss"])


@pytest.mark.gpu
def test_caption_decoder_matches_reference_cuda(golden_dir):
    g, sd = _golden(golden_dir)
    dec = _module(sd).to("cuda").eval()
    memory, caps, lens = inputs()
    memory = memory.to("cuda").requires_grad_(True)
    pred, caps_sorted, decode_lengths, sort_ind = dec(memory, caps.to("cuda"), lens.to("cuda"))
    loss = caption_loss(pred, caps_sorted, decode_lengths)
    loss.backward()
    assert pred.shape == (caps.shape[0], caps.shape[1], ARGS["vocab_size"])
    assert np.array_equal(sort_ind.cpu().numpy(), g["sort_ind"]) and decode_lengths == g["decode_lengths"].tolist()
    assert np.array_equal(caps_sorted.cpu().numpy(), g["caps_sorted"])
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g["pred"], rtol=1e-4, atol=2e-5)
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * float(g["loss"])
    np.testing.assert_allclose(memory.grad.cpu().numpy(), g["grad_memory"], rtol=1e-3, atol=1e-6)
    named = dict(dec.named_parameters())
    for k in g.files:
        if k.startswith("grad:"):
            np.testing.assert_allclose(named[k[5:]].grad.cpu().numpy(), g[k], rtol=1e-3, atol=1e-6, err_msg=k)
    assert sorted(k for k, p in named.items() if p.grad is None) == g["unused_grad_none"].tolist()


def _sdpa_ref(q, k, v, nh, causal, keep, keep_scale):
    """fp64 restatement of the kernel's contract on (L, B, E) tensors."""
    Lq, B, E = q.shape
    Lk, hd = k.shape[0], E // nh
    qh = q.double().view(Lq, B, nh, hd).permute(1, 2, 0, 3)
    kh = k.double().view(Lk, B, nh, hd).permute(1, 2, 0, 3)
    vh = v.double().view(Lk, B, nh, hd).permute(1, 2, 0, 3)
    s = qh @ kh.transpose(2, 3) / np.sqrt(hd)
    if causal:
        i, j = torch.arange(Lq).view(-1, 1), torch.arange(Lk).view(1, -1)
        s = s.masked_fill((j > i + (Lk - Lq)).to(s.device), float("-inf"))
    p = torch.softmax(s, dim=-1)
    pd = p * keep.double() * keep_scale if keep is not None else p
    return (pd @ vh).permute(2, 0, 1, 3).reshape(Lq, B, E), p


@pytest.mark.gpu
@pytest.mark.parametrize("Lq,Lk,B,nh,hd,causal,p_drop", [(52, 52, 3, 8, 24, True, 0.0), (52, 256, 2, 8, 24, False, 0.1),
                                                         (1, 37, 5, 8, 24, False, 0.0), (12, 12, 3, 4, 8, True, 0.25),
                                                         (64, 256, 1, 2, 32, False, 0.0)])
def test_attention_kernels_forward_backward(Lq, Lk, B, nh, hd, causal, p_drop):
    from change3d_b200 import attention as A
    dev = "cuda"
    g = torch.Generator().manual_seed(Lq * 7 + Lk)
    E = nh * hd
    q = torch.randn(Lq, B, E, generator=g).to(dev)
    # k / v as slices of one packed (Lk, B, 2E) tensor: the strided addressing the decode step relies on
    kv = torch.randn(Lk, B, 2 * E, generator=g).to(dev)
    k, v = kv[..., :E], kv[..., E:]
    keep = (torch.rand(B, nh, Lq, Lk, generator=g) >= p_drop).to(torch.uint8).to(dev) if p_drop > 0 else None
    scale = 1.0 / (1.0 - p_drop)
    P = torch.empty(B, nh, Lq, Lk, device=dev)
    o = A.attention_forward(q, k, v, nh, causal, P, keep, scale)
    q64, k64, v64 = (t.detach().double().clone().requires_grad_(True) for t in (q, k, v))
    o_ref, p_ref = _sdpa_ref(q64, k64, v64, nh, causal, keep, scale)
    assert (o.double() - o_ref).abs().max().item() <= 2e-6 * max(1.0, o_ref.abs().max().item())
    assert (P.double() - p_ref).abs().max().item() <= 2e-6
    go = torch.randn(Lq, B, E, generator=g).to(dev)
    o_ref.backward(go.double())
    import ctypes as C
    from change3d_b200 import _lib as L
    kc, vc = k.contiguous(), v.contiguous()
    d = A._desc(q, kc, vc, None, P, keep, scale, nh, causal)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(kc), torch.empty_like(vc)
    L.check(L.load().c3d_attention_bwd(C.byref(d), go.data_ptr(), go.stride(0), go.stride(1), dq.data_ptr(), dk.data_ptr(),
                                       dv.data_ptr(), torch.cuda.current_stream().cuda_stream), "c3d_attention_bwd")
    for got, want in ((dq, q64.grad), (dk, k64.grad), (dv, v64.grad)):
        assert (got.double() - want).abs().max().item() <= 5e-6 * max(1.0, want.abs().max().item())


@pytest.mark.gpu
def test_training_dropout_path_runs_and_scales(golden_dir):
    """Train mode: attention dropout inside the kernel (keep mask drawn by torch): finite loss and gradients, and the
    loss moves away from the eval value."""
    g, sd = _golden(golden_dir)
    dec = _module(sd).to("cuda").train()
    memory, caps, lens = inputs()
    torch.manual_seed(0)
    pred, caps_sorted, decode_lengths, _ = dec(memory.to("cuda"), caps.to("cuda"), lens.to("cuda"))
    loss = caption_loss(pred, caps_sorted, decode_lengths)
    loss.backward()
    assert torch.isfinite(loss) and abs(loss.item() - float(g["loss"])) > 1e-6
    assert all(torch.isfinite(p.grad).all() for p in dec.live_parameters())


@pytest.mark.gpu
@pytest.mark.parametrize("beam", [1, 3])
def test_cached_batched_search_matches_evaluate_restatement(golden_dir, beam):
    from change3d_b200.caption_decode import CaptionSearch
    _g, sd = _golden(golden_dir)
    # a random-weight head either stops at once or never: make the <end> logit grow with the position (a ramp on channel 0
    # of the positional table that the <end> row of the vocabulary projection reads), so that captions end after 2..8
    # words and beam 3 picks different captions than beam 1 for some of the pairs
    sd = {k: v.clone() for k, v in sd.items()}
    start_id, end_id = 1, 2
    sd["position_encoding.pe"][:, 0, 0] += 0.3 * torch.arange(sd["position_encoding.pe"].shape[0])
    sd["wdc.weight"][end_id] = 0.0
    sd["wdc.weight"][end_id, 0] = 3.0
    dec = _module(sd).to("cuda").eval()
    gen = torch.Generator().manual_seed(5)
    B, S = 8, 16
    memory = torch.randn(S, B, ARGS["embed_dim"], generator=gen)
    got = CaptionSearch(dec).search(memory.to("cuda"), start_id, end_id, beam_size=beam, max_len=52)
    n_done = 0
    for b in range(B):
        want_seq, want_score = CO.search_one(sd, memory[:, b:b + 1], ARGS["n_head"], start_id, end_id, beam, 52)
        seq, score = got[b]
        assert seq == want_seq, (b, seq, want_seq)
        if want_seq is not None:
            n_done += 1
            assert abs(score - want_score) <= 1e-4 * max(1.0, abs(want_score))
    assert n_done == B and len({len(s_) for s_, _ in got}) >= 3          # captions of several lengths
