"""CaptionDecoder (torch ops with the reference's module tree) against the reference's own
module: tests/golden/caption_decoder.npz holds the reference state dict, outputs, loss (scripts/train_CC.py:118-133,
:463) and gradients (oracle/make_golden_caption.py).  Runs on CPU (the head is plain torch) and, marked gpu, on CUDA."""
import argparse
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from oracle.make_golden_caption import ARGS, caption_loss, inputs


def _run(device, golden_dir):
    from change3d_b200.model.caption_decoder import CaptionDecoder
    g = np.load(os.path.join(golden_dir, "caption_decoder.npz"))
    with contextlib.redirect_stdout(io.StringIO()):
        dec = CaptionDecoder(argparse.Namespace(**ARGS))
    sd = {k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd:")}
    assert list(sd) == list(dec.state_dict())                       # same keys, same registration order
    dec.load_state_dict(sd, strict=True)
    dec = dec.to(device).eval()
    memory, caps, lens = inputs()
    memory = memory.to(device).requires_grad_(True)
    pred, caps_sorted, decode_lengths, sort_ind = dec(memory, caps.to(device), lens.to(device))
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
    # parameters the reference registers but never uses stay without gradient here too
    assert sorted(k for k, p in named.items() if p.grad is None) == g["unused_grad_none"].tolist()


def test_caption_decoder_matches_reference_cpu(golden_dir):
    _run("cpu", golden_dir)


@pytest.mark.gpu
def test_caption_decoder_matches_reference_cuda(golden_dir):
    _run("cuda", golden_dir)
