"""CPU restatement (plain torch) of the reference's captioning head and caption search — TEST INFRASTRUCTURE.

Pinned against tests/golden/caption_decoder.npz (the reference's own CaptionDecoder, oracle/make_golden_caption.py) by
tests/test_caption_decoder.py.  Follows
  * model/caption_decoder.py:272-313 (PositionalEncoding), :393-423 (Mesh_TransformerDecoderLayer.forward: norm1(tgt +
    self_attn(tgt)) -> norm2(x + multihead_attn2(x, memory))), :574-612 (CaptionDecoder.forward), eval mode;
  * scripts/train_CC.py:209-322 (`evaluate`: per image pair, beam search that re-runs the whole decoder over all 52
    positions for every generated token, removes finished beams, stops after 50 steps, keeps the finished sequence with
    the highest score).
Functional on a state dict `sd` with the reference's keys (no module objects)."""
import math

import torch
import torch.nn.functional as F


def _mha(sd, p, query, key, value, nh, mask=None):
    E = query.shape[-1]
    out, _ = F.multi_head_attention_forward(
        query, key, value, E, nh, sd[p + "in_proj_weight"], sd[p + "in_proj_bias"], None, None, False, 0.0,
        sd[p + "out_proj.weight"], sd[p + "out_proj.bias"], training=False, need_weights=False, attn_mask=mask)
    return out


def n_layers(sd) -> int:
    return 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))


def transformer(sd, emb, memory, nh, mask):
    x = emb
    E = x.shape[-1]
    for i in range(n_layers(sd)):
        p = f"transformer.layers.{i}."
        x = F.layer_norm(x + _mha(sd, p + "self_attn.", x, x, x, nh, mask), (E,), sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        x = F.layer_norm(x + _mha(sd, p + "multihead_attn2.", x, memory, memory, nh), (E,), sd[p + "norm2.weight"],
                         sd[p + "norm2.bias"])
    return x


def causal_mask(L, dtype=torch.float32):
    return torch.full((L, L), float("-inf"), dtype=dtype).triu(diagonal=1)


def embed(sd, tgt):
    """vocab_embedding + positional encoding (eval: no dropout).  tgt (L, N) int64."""
    return F.embedding(tgt, sd["vocab_embedding.weight"]) + sd["position_encoding.pe"][:tgt.shape[0]]


def decoder_forward(sd, memory, encoded_captions, caption_lengths, nh):
    """CaptionDecoder.forward (model/caption_decoder.py:574-612), eval mode."""
    tgt = encoded_captions.permute(1, 0)
    pred = transformer(sd, embed(sd, tgt), memory, nh, causal_mask(tgt.shape[0], memory.dtype))
    pred = F.linear(pred, sd["wdc.weight"], sd["wdc.bias"]).permute(1, 0, 2)
    lens, sort_ind = caption_lengths.squeeze(1).sort(dim=0, descending=True)
    return pred[sort_ind], encoded_captions[sort_ind], (lens - 1).tolist(), sort_ind


def search_one(sd, memory_1, nh, start_id, end_id, beam_size, max_len=52):
    """scripts/train_CC.py:216-334 for ONE image pair.  memory_1 (S, 1, D).  Returns (best finished sequence as a list
    of token ids incl. <start> / <end>, its score) or (None, None) when no beam finished within 50 steps (the script
    then records nothing for the pair)."""
    vocab = sd["wdc.weight"].shape[0]
    k = beam_size
    tgt = torch.zeros(max_len, k, dtype=torch.int64)
    mask = causal_mask(max_len, memory_1.dtype)
    tgt[0, :] = start_id
    seqs = torch.full((k, 1), start_id, dtype=torch.int64)
    top_k_scores = torch.zeros(k, 1, dtype=memory_1.dtype)
    complete_seqs, complete_scores = [], []
    step = 1
    k_prev_words = tgt.permute(1, 0)
    enc = memory_1.expand(memory_1.shape[0], k, memory_1.shape[2]).permute(1, 0, 2)        # (k, S, D)
    while True:
        t = k_prev_words.permute(1, 0)
        pred = transformer(sd, embed(sd, t), enc.permute(1, 0, 2), nh, mask)
        scores = F.linear(pred, sd["wdc.weight"], sd["wdc.bias"]).permute(1, 0, 2)[:, step - 1, :]
        scores = F.log_softmax(scores, dim=1)
        scores = top_k_scores.expand_as(scores) + scores
        if step == 1:
            top_k_scores, top_k_words = scores[0].topk(k, 0, True, True)
        else:
            top_k_scores, top_k_words = scores.reshape(-1).topk(k, 0, True, True)
        prev_word_inds = top_k_words // vocab
        next_word_inds = top_k_words % vocab
        seqs = torch.cat([seqs[prev_word_inds], next_word_inds.unsqueeze(1)], dim=1)
        incomplete = [i for i, w in enumerate(next_word_inds.tolist()) if w != end_id]
        complete = sorted(set(range(len(next_word_inds))) - set(incomplete))
        if complete:
            complete_seqs.extend(seqs[complete].tolist())
            complete_scores.extend(top_k_scores[complete].tolist())
        k -= len(complete)
        if k == 0:
            break
        seqs = seqs[incomplete]
        enc = enc[prev_word_inds[incomplete]]
        top_k_scores = top_k_scores[incomplete].unsqueeze(1)
        k_prev_words = k_prev_words[incomplete]
        k_prev_words[:, :step + 1] = seqs
        if step > 50:
            break
        step += 1
    if not complete_scores:
        return None, None
    i = complete_scores.index(max(complete_scores))
    return complete_seqs[i], complete_scores[i]
