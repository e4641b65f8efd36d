"""CaptionDecoder — counterpart of the reference's `model/caption_decoder.py:272-612` (CC task).

The module tree is the reference's, so its checkpoints load: every parameter the reference registers is registered
here under the same name, including the ones its forward never uses (self_attn2, multihead_attn, multihead_attn3,
linear1/2, norm3, fc_alpha1-3, embedding_1D).  The nn.MultiheadAttention objects are parameter containers: the layer's
forward runs their scaled-dot-product core (scores, causal mask, softmax, attention dropout, weighted sum, and the
backward of all of it) on the sm_100a attention kernel (`change3d_b200.attention.fused_mha`, csrc/attention.cu); the
projections, LayerNorms and the vocabulary projection are library GEMMs / torch ops (< 1 % of
a CC step's FLOPs).  CUDA tensors only, like the rest of the package.

Two deliberate fixes relative to the reference, neither changing the arithmetic:
  * the decoder layer accepts (and ignores) the `tgt_is_causal` / `memory_is_causal` keywords that
    nn.TransformerDecoder passes on torch >= 2.0 (the reference layer raises TypeError there);
  * the causal mask is built on the input's device instead of `.cuda()`.
"""
import math
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from ..attention import fused_mha
from .utils import weight_init


class PositionalEncoding(nn.Module):
    """Fixed sinusoidal table added to the token embeddings (model/caption_decoder.py:272-313)."""

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        freq = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        table = torch.zeros(max_len, d_model)
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer('pe', table.unsqueeze(1))
        self.embedding_1D = nn.Embedding(52, int(d_model))    # registered by the reference, unused

    def forward(self, x):
        return self.dropout(x + self.pe[:x.size(0), :])


class Mesh_TransformerDecoderLayer(nn.Module):
    """Live path (model/caption_decoder.py:411-423): norm1(tgt + SA(tgt)) -> norm2(. + MHA2(., memory))."""
    __constants__ = ['batch_first', 'norm_first']

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, layer_norm_eps=1e-5,
                 batch_first=False, norm_first=False, device=None, dtype=None):
        super().__init__()
        mha = lambda: nn.MultiheadAttention(int(d_model), int(nhead), dropout=dropout)   # noqa: E731
        self.self_attn = mha()
        self.self_attn2 = mha()
        self.multihead_attn = mha()
        self.multihead_attn2 = mha()
        self.multihead_attn3 = mha()
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm_first = norm_first
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm3 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        for i in range(1, 6):
            setattr(self, f"dropout{i}", nn.Dropout(dropout))
        self.activation = nn.ReLU()
        self.activation2 = nn.Softmax(dim=-1)
        for i in range(1, 4):
            lin = nn.Linear(2 * d_model, d_model)
            nn.init.xavier_uniform_(lin.weight)
            nn.init.constant_(lin.bias, 0)
            setattr(self, f"fc_alpha{i}", lin)
        weight_init(self)

    def forward(self, tgt: Tensor, memory: Tensor, tgt_mask: Optional[Tensor] = None,
                memory_mask: Optional[Tensor] = None, tgt_key_padding_mask: Optional[Tensor] = None,
                memory_key_padding_mask: Optional[Tensor] = None, **_ignored) -> Tensor:
        if tgt_key_padding_mask is not None or memory_key_padding_mask is not None or memory_mask is not None:
            raise RuntimeError("Mesh_TransformerDecoderLayer: padding / memory masks are not used by the reference's "
                               "forward (model/caption_decoder.py:574-612) and are not implemented")
        # tgt_mask is the causal mask CaptionDecoder.forward builds (model/caption_decoder.py:589-593) or None
        sa = fused_mha(self.self_attn, tgt, tgt, tgt, causal=tgt_mask is not None)
        x = self.norm1(tgt + self.dropout1(sa))
        ca = fused_mha(self.multihead_attn2, x, memory, memory, causal=False)
        return self.norm2(x + self.dropout3(ca))


class CaptionDecoder(nn.Module):
    def __init__(self, args):
        super().__init__()
        print(f"decoder_n_layers={args.n_layer}")
        self.vocab_embedding = nn.Embedding(args.vocab_size, args.embed_dim)
        layer = Mesh_TransformerDecoderLayer(args.embed_dim, args.n_head, dim_feedforward=args.embed_dim * 4,
                                             dropout=args.dropout)
        self.transformer = nn.TransformerDecoder(layer, args.n_layer)
        self.position_encoding = PositionalEncoding(args.embed_dim)
        self.wdc = nn.Linear(args.embed_dim, args.vocab_size)
        self.dropout_layer = nn.Dropout(p=args.dropout)
        self.init_weights()

    def init_weights(self):
        self.vocab_embedding.weight.data.uniform_(-0.1, 0.1)
        self.wdc.bias.data.fill_(0)
        self.wdc.weight.data.uniform_(-0.1, 0.1)

    def forward_device(self, memory, encoded_captions, caption_lengths):
        """`forward` without the host read-back of the decode lengths (they stay a device tensor), so that a training
        iteration can be captured in a CUDA graph: returns (pred sorted, sorted captions, decode lengths (B,) int64,
        sort indices)."""
        tgt = encoded_captions.permute(1, 0)
        L = tgt.size(0)
        mask = torch.full((L, L), float('-inf'), device=tgt.device).triu(diagonal=1)
        emb = self.position_encoding(self.vocab_embedding(tgt))
        # tgt_is_causal=False: nn.TransformerDecoder otherwise inspects the mask on the host (`.all()` read-back per
        # call — a device synchronisation, and illegal while a CUDA graph is being captured); the layers receive the
        # explicit mask either way and ignore the hint
        pred = self.transformer(emb, memory, tgt_mask=mask, tgt_is_causal=False)
        pred = self.wdc(self.dropout_layer(pred)).permute(1, 0, 2)
        caption_lengths, sort_ind = caption_lengths.squeeze(1).sort(dim=0, descending=True)
        return pred[sort_ind], encoded_captions[sort_ind], caption_lengths - 1, sort_ind

    def forward(self, memory, encoded_captions, caption_lengths):
        """(memory (S,B,D), captions (B,L) int64, lengths (B,1)) -> (pred (B,L,V) sorted by length, sorted
        captions, decode lengths, sort indices) — model/caption_decoder.py:574-612."""
        pred, encoded_captions, decode_lengths, sort_ind = self.forward_device(memory, encoded_captions, caption_lengths)
        return pred, encoded_captions, decode_lengths.tolist(), sort_ind

    def live_parameters(self):
        """Parameters that `forward` uses, i.e. the ones that receive gradients (model/caption_decoder.py:411-423 runs
        self_attn, norm1, multihead_attn2, norm2 of every layer; torch.optim.Adam skips the others because their
        `.grad` stays None, so they are neither updated nor weight-decayed)."""
        ps = [self.vocab_embedding.weight]
        for layer in self.transformer.layers:
            for m in (layer.self_attn, layer.multihead_attn2, layer.norm1, layer.norm2):
                ps += list(m.parameters())
        return ps + [self.wdc.weight, self.wdc.bias]
