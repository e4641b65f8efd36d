"""Generates tests/golden/caption_decoder.npz with the REFERENCE's own CaptionDecoder
(model/caption_decoder.py:528-612, Mesh_TransformerDecoderLayer :316-423) and the loss of scripts/train_CC.py:118-133
(pack_padded_sequence + CrossEntropyLoss) on seeded inputs, eval mode (dropout off), with autograd gradients.
Two test-only accommodations, neither changing arithmetic: the layer is wrapped to drop the tgt_is_causal /
memory_is_causal keywords torch >= 2.0 passes (SURVEY.md §8c CC caveat), and `Tensor.cuda` is a no-op while the
reference forward runs (it calls `mask.cuda()`, model/caption_decoder.py:593; this container has no GPU).
TEST INFRASTRUCTURE, authoring container only:  python -m oracle.make_golden_caption
"""
import argparse
import contextlib
import importlib
import io
import os

import numpy as np
import torch
from torch.nn.utils.rnn import pack_padded_sequence

from oracle import reference_loader as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = dict(vocab_size=60, embed_dim=32, n_head=4, n_layer=2, dropout=0.1)
S, B, L = 16, 3, 12


def inputs(seed: int = 31):
    g = torch.Generator().manual_seed(seed)
    memory = torch.randn(S, B, ARGS["embed_dim"], generator=g)
    caps = torch.randint(1, ARGS["vocab_size"], (B, L), generator=g)
    lens = torch.tensor([[7], [12], [5]])
    for b in range(B):
        caps[b, int(lens[b]):] = 0                                   # <pad>
    return memory, caps, lens


def caption_loss(scores, caps_sorted, decode_lengths):
    """scripts/train_CC.py:124-133 with the criterion of :463 (CrossEntropyLoss(ignore_index=0))."""
    targets = caps_sorted[:, 1:]
    s = pack_padded_sequence(scores, decode_lengths, batch_first=True).data
    t = pack_padded_sequence(targets, decode_lengths, batch_first=True).data
    return torch.nn.functional.cross_entropy(s, t, ignore_index=0)


def main() -> None:
    R.load()
    cd = importlib.import_module("model.caption_decoder")            # the reference's file, unchanged
    fwd = cd.Mesh_TransformerDecoderLayer.forward

    def fwd_compat(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                   memory_key_padding_mask=None, **_):
        return fwd(self, tgt, memory, tgt_mask, memory_mask, tgt_key_padding_mask, memory_key_padding_mask)

    cd.Mesh_TransformerDecoderLayer.forward = fwd_compat
    torch.manual_seed(16)
    with contextlib.redirect_stdout(io.StringIO()):
        dec = cd.CaptionDecoder(argparse.Namespace(**ARGS)).eval()
    memory, caps, lens = inputs()
    memory.requires_grad_(True)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        pred, caps_sorted, decode_lengths, sort_ind = dec(memory, caps, lens)
    finally:
        torch.Tensor.cuda = orig_cuda
    loss = caption_loss(pred, caps_sorted, decode_lengths)
    loss.backward()
    out = {"sd:" + k: v.numpy() for k, v in dec.state_dict().items()}
    named = dict(dec.named_parameters())
    out.update(pred=pred.detach().numpy(), caps_sorted=caps_sorted.numpy(), decode_lengths=np.array(decode_lengths),
               sort_ind=sort_ind.numpy(), loss=np.float64(loss.item()), grad_memory=memory.grad.numpy())
    for k in ("wdc.weight", "vocab_embedding.weight", "transformer.layers.0.self_attn.in_proj_weight",
              "transformer.layers.1.multihead_attn2.out_proj.weight", "transformer.layers.1.norm2.bias"):
        out["grad:" + k] = named[k].grad.numpy()
    out["unused_grad_none"] = np.array(sorted(k for k, p in named.items() if p.grad is None))
    path = os.path.join(ROOT, "tests", "golden", "caption_decoder.npz")
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1024:.0f} KiB", "loss", loss.item(), "unused", len(out["unused_grad_none"]))


if __name__ == "__main__":
    main()
