"""Batched caption search with cached keys / values.

The reference's `evaluate` (scripts/train_CC.py:209-322) handles one image pair at a time and, for every generated token,
re-runs the whole decoder over all 52 (zero-padded) positions of every live beam: O(L^2) decoder passes per caption, a
handful of tiny kernels each.  With the causal mask, position t of the decoder output depends on tokens 0..t only, so
the same scores come out of an incremental pass that feeds ONE new token per step, appends its self-attention key /
value to a per-layer cache and attends over the cache (one query row per sequence: `c3d_attention_fwd` with Lq = 1).
The cross-attention keys / values of the 256-token memory are projected once per image.  All image pairs of a batch
and all their beams advance together: N = B * beam sequences per step.

Search semantics are the script's: log-softmax scores accumulated per beam; step 1 expands beam 0 only; at every step
the best k_alive candidates over (alive beams x vocabulary) of an image become its new beams; a beam that emits <end>
is finished (recorded with its score) and k_alive shrinks; at most 51 steps; the finished sequence with the highest
score wins (first one on ties).  Images with no finished beam return `None` (the script records nothing for them).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from .attention import attention_forward


class CaptionSearch:
    def __init__(self, decoder):
        """decoder: change3d_b200.model.caption_decoder.CaptionDecoder (eval mode is enforced per call)."""
        self.dec = decoder

    # ------------------------------------------------------------------------------------------------------
    def _memory_kv(self, memory: torch.Tensor):
        """Per layer: keys / values of the memory for multihead_attn2, (S, B, E) each."""
        out = []
        for layer in self.dec.transformer.layers:
            m = layer.multihead_attn2
            E = m.embed_dim
            kv = F.linear(memory, m.in_proj_weight[E:], m.in_proj_bias[E:])          # (S, B, 2E), one GEMM
            out.append((kv[..., :E], kv[..., E:]))
        return out

    def _step(self, tok: torch.Tensor, pos: int, caches, mem_kv, nh: int) -> torch.Tensor:
        """tok (N,) int64: token at position `pos` of every sequence.  Appends to the caches, returns the decoder output
        at that position projected to the vocabulary, (N, V) log-probabilities."""
        dec = self.dec
        x = dec.vocab_embedding(tok).unsqueeze(0) + dec.position_encoding.pe[pos:pos + 1]      # (1, N, E)
        for li, layer in enumerate(dec.transformer.layers):
            m = layer.self_attn
            E = m.embed_dim
            qkv = F.linear(x, m.in_proj_weight, m.in_proj_bias)                                # (1, N, 3E)
            Kc, Vc = caches[li]
            Kc[pos] = qkv[0, :, E:2 * E]
            Vc[pos] = qkv[0, :, 2 * E:]
            sa = attention_forward(qkv[..., :E], Kc[:pos + 1], Vc[:pos + 1], nh, causal=False)
            x = layer.norm1(x + F.linear(sa, m.out_proj.weight, m.out_proj.bias))
            m2 = layer.multihead_attn2
            q2 = F.linear(x, m2.in_proj_weight[:E], m2.in_proj_bias[:E])
            mk, mv = mem_kv[li]
            ca = attention_forward(q2, mk, mv, nh, causal=False)
            x = layer.norm2(x + F.linear(ca, m2.out_proj.weight, m2.out_proj.bias))
        return F.log_softmax(dec.wdc(x[0]), dim=1)

    # ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def search(self, memory: torch.Tensor, start_id: int, end_id: int, beam_size: int = 1,
               max_len: int = 52) -> List[Tuple[Optional[List[int]], Optional[float]]]:
        """memory (S, B, D) on the GPU (`rearrange(encoder(A, B, output_final=True), 'b c h w -> (h w) b c')`).
        Returns, per image pair, (token ids incl. <start> / <end>, score) of the best finished beam, or (None, None)."""
        if not memory.is_cuda:
            raise RuntimeError("CaptionSearch: CUDA tensors required (no CPU/eager fallback)")
        dec = self.dec
        was_training = dec.training
        dec.eval()
        try:
            return self._search(memory.float(), start_id, end_id, beam_size, max_len)
        finally:
            dec.train(was_training)

    def _search(self, memory, start_id, end_id, k, max_len):
        dec = self.dec
        dev = memory.device
        S, B, D = memory.shape
        nh = dec.transformer.layers[0].self_attn.num_heads
        V = dec.wdc.out_features
        N = B * k
        # every beam of image b attends to the same memory: expand once (S, B, E) -> (S, B * k, E)
        mem_kv = [(mk.repeat_interleave(k, dim=1).contiguous(), mv.repeat_interleave(k, dim=1).contiguous())
                  for mk, mv in self._memory_kv(memory)]
        E = dec.wdc.in_features
        caches = [(torch.zeros(max_len, N, E, device=dev), torch.zeros(max_len, N, E, device=dev))
                  for _ in dec.transformer.layers]
        seqs = torch.full((B, k, max_len), 0, dtype=torch.int64, device=dev)
        seqs[:, :, 0] = start_id
        score = torch.zeros(B, k, device=dev)
        alive = torch.ones(B, k, dtype=torch.bool, device=dev)          # beam slot still being extended
        k_alive = torch.full((B,), k, dtype=torch.int64, device=dev)
        done_seq = [[] for _ in range(B)]
        done_score = [[] for _ in range(B)]
        slot = torch.arange(k, device=dev)
        img_base = (torch.arange(B, device=dev) * k).unsqueeze(1)
        neg = float("-inf")
        for step in range(1, max_len):                                  # step = number of tokens after <start>
            logp = self._step(seqs[:, :, step - 1].reshape(N), step - 1, caches, mem_kv, nh).view(B, k, V)
            cand = score.unsqueeze(2) + logp
            expand = alive.clone()
            if step == 1:
                expand[:, 1:] = False                                   # all beams are identical: expand beam 0 only
            cand = torch.where(expand.unsqueeze(2), cand, torch.full_like(cand, neg))
            top_s, top_i = cand.view(B, k * V).topk(k, dim=1, largest=True, sorted=True)
            parent, word = top_i // V, top_i % V
            chosen = slot.unsqueeze(0) < k_alive.unsqueeze(1)           # the best k_alive candidates become beams
            new_seqs = torch.gather(seqs, 1, parent.unsqueeze(2).expand(B, k, max_len)).clone()
            new_seqs[:, :, step] = word
            finished = chosen & (word == end_id)
            if bool(finished.any()):
                fb, fs = torch.nonzero(finished, as_tuple=True)
                fin_seq = new_seqs[fb, fs, :step + 1].tolist()
                fin_sc = top_s[fb, fs].tolist()
                for b_, sq, sc in zip(fb.tolist(), fin_seq, fin_sc):
                    done_seq[b_].append(sq)
                    done_score[b_].append(sc)
            # compact: surviving beams first (the script's `incomplete_inds` order), dead slots after them
            survive = chosen & ~finished
            order = torch.argsort((~survive).to(torch.int8), dim=1, stable=True)
            seqs = torch.gather(new_seqs, 1, order.unsqueeze(2).expand(B, k, max_len))
            score = torch.gather(top_s, 1, order)
            alive = torch.gather(survive, 1, order)
            k_alive = alive.sum(1)
            src = (img_base + torch.gather(parent, 1, order)).reshape(N)        # cache rows follow their beams
            for li in range(len(caches)):
                Kc, Vc = caches[li]
                caches[li] = (Kc[:, src].contiguous(), Vc[:, src].contiguous())
            if int(k_alive.sum()) == 0 or step > 50:
                break
        out = []
        for b_ in range(B):
            if not done_score[b_]:
                out.append((None, None))
            else:
                i = done_score[b_].index(max(done_score[b_]))
                out.append((done_seq[b_][i], done_score[b_][i]))
        return out
