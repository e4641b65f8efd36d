"""Runner with the surface of the reference's `scripts/train_CC.py` (change captioning).

Same command-line flags and defaults (train_CC.py:533-684), same epoch structure (`trainValidate` :401-530: halve both
learning rates every 10 epochs, one epoch of training, `evaluate`, keep the best BLEU-4, `checkpoint_<dataset>.pth.tar`
with the script's keys), same iteration (`train_step.CCTrainStep`: encoder to res5 -> memory -> captioning head ->
packed cross-entropy -> +-grad_clip -> two Adams, train_CC.py:105-146).

What differs: one process per GPU under torchrun (DistributedSampler, NCCL all-reduce of the two flat gradient
buffers), batches staged by `DevicePrefetcher`, the iteration as a CUDA graph, top-1 accuracy reduced on the device, and
`evaluate` as a BATCHED cached search (`caption_decode.CaptionSearch`) instead of one image pair and one full decoder
pass per generated token.

Scoring: the script scores captions with pycocoevalcap (BLEU, METEOR [Java], ROUGE_L, CIDEr: `eval_func/**`, outside the
hot path).  `bleu_scores` here restates the corpus BLEU-1..4 of that package (clipped n-gram
counts, closest reference length, brevity penalty) because the loop selects its best checkpoint by BLEU-4; the other
three metrics are not computed.
"""
from __future__ import annotations

import math
import os
import time
from argparse import ArgumentParser
from collections import Counter
from os.path import join as osp
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist

from .caption_decode import CaptionSearch
from .input_pipeline import DevicePrefetcher
from .model.trainer import Trainer
from .train_step import CCTrainStep


def build_parser() -> ArgumentParser:
    """scripts/train_CC.py:533-684: same names / types / defaults, plus --synthetic, --vocab_size and --no_graph."""
    p = ArgumentParser()
    p.add_argument('--file_root', default="path/to/LEVIR-CC")
    p.add_argument('--dataset', default="LEVIR_CC_5_cap_per_img_5_min_word_freq")
    p.add_argument('--n_head', type=int, default=8, help='Multi-head attention in Transformer.')
    p.add_argument('--n_layer', type=int, default=3)
    p.add_argument('--decoder_n_layers', type=int, default=1)
    p.add_argument('--embed_dim', type=int, default=192)
    p.add_argument('--dropout', type=float, default=0.1, help='Dropout rate')
    p.add_argument('--num_perception_frame', type=int, default=1)
    p.add_argument('--in_height', type=int, default=256)
    p.add_argument('--in_width', type=int, default=256)
    p.add_argument('--epochs', type=int, default=200)
    p.add_argument('--batch_size', type=int, default=32)
    p.add_argument('--print_freq', type=int, default=100)
    p.add_argument('--workers', type=int, default=1)
    p.add_argument('--encoder_lr', type=float, default=1e-4)
    p.add_argument('--decoder_lr', type=float, default=1e-4)
    p.add_argument('--grad_clip', type=float, default=5.)
    p.add_argument('--fine_tune_encoder', type=bool, default=True)
    p.add_argument('--checkpoint', default=None)
    p.add_argument('--pretrained', default='model/X3D_L.pyth', type=str)
    p.add_argument('--gpu_id', default=0, type=int)
    p.add_argument('--Split', default="TEST", help='Validation split')
    p.add_argument('--beam_size', type=int, default=1, help='Beam size for beam search.')
    p.add_argument('--save_dir', default='./exp')
    # additions
    p.add_argument('--synthetic', type=int, default=0, help='use N seeded random pairs instead of --file_root')
    p.add_argument('--vocab_size', type=int, default=0, help='with --synthetic: vocabulary size (the script reads the word map)')
    p.add_argument('--no_graph', action='store_true')
    return p


# <pad> 0, <start> / <end> / <unk> as the word maps of the reference's preprocessing number them (last three ids)
def special_ids(vocab_size: int) -> Dict[str, int]:
    return {"<pad>": 0, "<unk>": vocab_size - 3, "<start>": vocab_size - 2, "<end>": vocab_size - 1}


class SyntheticCC(torch.utils.data.Dataset):
    """Items like CaptionDataset (data/dataset.py:343-440): TRAIN -> (img_pairs (2,3,H,W), caption (52,), caplen (1,));
    other splits add all captions of the pair (cpi, 52).  Every pair has `cpi` captions and is repeated cpi times, as in
    the reference's files; the caption depends on whether the pair "changed" so that there is something to learn."""
    L = 52

    def __init__(self, n_pairs: int, H: int, W: int, vocab: int, split: str, seed: int, cpi: int = 5):
        self.n, self.H, self.W, self.V, self.split, self.seed, self.cpi = n_pairs, H, W, vocab, split, seed, cpi

    def __len__(self) -> int:
        return self.n * self.cpi

    def _pair(self, i: int):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        a = torch.rand(3, self.H, self.W, generator=g) * 2 - 1
        changed = i % 2 == 0
        b = a + 0.05 * torch.randn(3, self.H, self.W, generator=g)
        if changed:
            s = self.H // 3
            b[:, s:2 * s, s:2 * s] = 1.0
        ids = special_ids(self.V)
        caps = torch.zeros(self.cpi, self.L, dtype=torch.int64)
        lens = torch.zeros(self.cpi, dtype=torch.int64)
        for c in range(self.cpi):
            body = [1 + (7 * c + j) % 5 + (5 if changed else 0) for j in range(3 + c % 3)]
            seq = [ids["<start>"]] + body + [ids["<end>"]]
            caps[c, :len(seq)] = torch.tensor(seq)
            lens[c] = len(seq)
        return torch.stack([a, b], 0), caps, lens

    def __getitem__(self, idx: int):
        pair, caps, lens = self._pair(idx // self.cpi)
        c = idx % self.cpi
        if self.split == "TRAIN":
            return pair, caps[c], lens[c:c + 1]
        return pair, caps[c], lens[c:c + 1], caps


# ---------------------------------------------------------------------------------------------
# BLEU (pycocoevalcap.bleu semantics: corpus level, clipped counts, closest reference length)
# ---------------------------------------------------------------------------------------------
def bleu_scores(references: Sequence[Sequence[Sequence[int]]], hypotheses: Sequence[Sequence[int]], n: int = 4) -> List[float]:
    tiny, small = 1e-15, 1e-9
    match, total = [0] * n, [0] * n
    hyp_len, ref_len = 0, 0
    for refs, hyp in zip(references, hypotheses):
        hyp = [str(t) for t in hyp]
        refs = [[str(t) for t in r] for r in refs]
        hyp_len += len(hyp)
        ref_len += min((abs(len(r) - len(hyp)), len(r)) for r in refs)[1]
        for k in range(1, n + 1):
            hc = Counter(tuple(hyp[i:i + k]) for i in range(len(hyp) - k + 1))
            mx: Counter = Counter()
            for r in refs:
                rc = Counter(tuple(r[i:i + k]) for i in range(len(r) - k + 1))
                for g_, c in rc.items():
                    mx[g_] = max(mx[g_], c)
            match[k - 1] += sum(min(c, mx[g_]) for g_, c in hc.items())
            total[k - 1] += max(0, len(hyp) - k + 1)
    out, bleu = [], 1.0
    ratio = (hyp_len + tiny) / (ref_len + small)
    bp = 1.0 if ratio >= 1 else math.exp(1 - 1 / ratio)
    for k in range(n):
        bleu *= (match[k] + tiny) / (total[k] + small)
        out.append(bleu ** (1.0 / (k + 1)) * bp)
    return out


# ---------------------------------------------------------------------------------------------
def _strip(seq, ids) -> List[int]:
    drop = {ids["<start>"], ids["<end>"], ids["<pad>"]}
    return [w for w in seq if w not in drop]


def train(args, loader, step: CCTrainStep, epoch: int, dev, verbose: bool = True):
    """scripts/train_CC.py:75-168.  Returns (average loss per decoded word, average top-1 accuracy in %)."""
    step.model.train()
    loss_sum = torch.zeros((), dtype=torch.float64, device=dev)
    acc_sum = torch.zeros((), dtype=torch.float64, device=dev)
    words = torch.zeros((), dtype=torch.float64, device=dev)
    full = args.batch_size
    t0 = time.time()
    for i, (pairs, caps, caplens) in enumerate(DevicePrefetcher(loader, dev)):
        a, b = pairs[:, 0].float().contiguous(), pairs[:, 1].float().contiguous()
        run = step.eager if (a.shape[0] != full and step.use_graph) else step
        loss = run(a, b, caps.long().contiguous(), caplens.long().contiguous())
        scores, caps_sorted, dl = step.outs
        Lc = caps_sorted.shape[1]
        keep = torch.arange(Lc - 1, device=dev).unsqueeze(0) < dl.unsqueeze(1)          # the packed positions
        hit = (scores[:, :Lc - 1].argmax(2) == caps_sorted[:, 1:]) & keep                # caption_accuracy(scores, targets, 1)
        n_tok = keep.sum().double()
        loss_sum += loss.double() * n_tok
        acc_sum += hit.sum().double() * 100.0
        words += n_tok
        if verbose and i % args.print_freq == 0:
            print(f"Epoch: {epoch}/{args.epochs} step: {i}/{len(loader)} Loss: {float(loss.item()):.4f} "
                  f"AVG_Loss: {float((loss_sum / words).item()):.4f} "
                  f"Top-5 Accuracy: {float((hit.sum().double() * 100.0 / n_tok).item()):.4f} "
                  f"Batch_time: {(time.time() - t0) / (i + 1):.4f}s")
    w = max(float(words.item()), 1.0)
    return float(loss_sum.item()) / w, float(acc_sum.item()) / w


@torch.no_grad()
def evaluate(args, loader, model, ids, dev, batch_pairs: int = 16, verbose: bool = True) -> dict:
    """scripts/train_CC.py:171-398 with the search batched: every `cpi`-th item of the loader is a new image pair (the
    script's `if (i + 1) % 5 != 0: continue`), `batch_pairs` of them are captioned per search call."""
    model.eval()
    search = CaptionSearch(model.decoder)
    refs, hyps = [], []
    pend_mem, pend_refs = [], []

    def flush():
        if not pend_mem:
            return
        out = search.search(torch.cat(pend_mem, 1), ids["<start>"], ids["<end>"], beam_size=args.beam_size, max_len=52)
        for (seq, _score), r in zip(out, pend_refs):
            if seq is None:                       # no finished beam within 50 steps: the script records nothing
                continue
            refs.append(r)
            hyps.append(_strip(seq, ids))
        pend_mem.clear()
        pend_refs.clear()

    cpi = getattr(loader.dataset, "cpi", 5)
    for i, (pairs, _caps, _lens, allcaps) in enumerate(loader):
        if (i + 1) % cpi != 0:
            continue
        pairs = pairs.to(dev, non_blocking=True).float()
        feat = model.update_cc(pairs[:, 0].contiguous(), pairs[:, 1].contiguous())
        B, C, H, W = feat.shape
        pend_mem.append(feat.permute(2, 3, 0, 1).reshape(H * W, B, C))
        pend_refs.append([_strip(c, ids) for c in allcaps[0].tolist()])
        if len(pend_mem) >= batch_pairs:
            flush()
    flush()
    b = bleu_scores(refs, hyps) if hyps else [0.0] * 4
    metrics = {"Bleu_1": b[0], "Bleu_2": b[1], "Bleu_3": b[2], "Bleu_4": b[3], "n_captions": len(hyps)}
    if verbose:
        print(f"evaluated {len(hyps)} pairs at beam size {args.beam_size}: " + " ".join(f"{k} {v:.4f}" for k, v in metrics.items()))
    return metrics


def train_validate(args, datasets=None) -> dict:
    """trainValidate (scripts/train_CC.py:401-530).  Returns the last epoch's metrics (rank 0) / {}."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(args.gpu_id)))
    if not torch.cuda.is_available():
        raise RuntimeError("change3d_b200.runner_cc needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if datasets is None:
        if args.synthetic <= 0 or args.vocab_size <= 8:
            raise RuntimeError("the reference's CaptionDataset (hdf5 + word map) is outside this package: pass "
                               "datasets=(train, eval) and args.vocab_size, or use --synthetic N --vocab_size V")
        datasets = (SyntheticCC(args.synthetic, args.in_height, args.in_width, args.vocab_size, "TRAIN", 1),
                    SyntheticCC(max(2, args.synthetic // 2), args.in_height, args.in_width, args.vocab_size, args.Split, 2))
    train_set, eval_set = datasets
    ids = special_ids(args.vocab_size)
    args.save_path = osp(args.save_dir, f"{args.dataset}_iter_{args.epochs}_lr_{args.encoder_lr}")
    if rank == 0:
        os.makedirs(args.save_path, exist_ok=True)
    torch.manual_seed(16)
    torch.cuda.manual_seed(16)
    model = Trainer(args).to(dev).float()
    step = CCTrainStep(model, encoder_lr=args.encoder_lr, decoder_lr=args.decoder_lr, grad_clip=args.grad_clip,
                       use_graph=not args.no_graph)
    sampler = None
    if world > 1:
        sampler = torch.utils.data.distributed.DistributedSampler(train_set, num_replicas=world, rank=rank, shuffle=True, seed=16)
    train_loader = torch.utils.data.DataLoader(train_set, batch_size=args.batch_size, shuffle=sampler is None, sampler=sampler,
                                               num_workers=args.workers, pin_memory=True)
    eval_loader = torch.utils.data.DataLoader(eval_set, batch_size=1, shuffle=False, num_workers=args.workers, pin_memory=True)
    best_bleu4, metrics = 0.0, {}
    for epoch in range(0, args.epochs):
        if sampler is not None:
            sampler.set_epoch(epoch)
        if epoch > 0 and epoch % 10 == 0:           # adjust_learning_rate(shrink_factor=0.5) on both optimizers (:477-479)
            for opt in (step.opt, step.dec_opt):
                for g in opt.param_groups:
                    g['lr'] = g['lr'] * 0.5
        loss_avg, acc_avg = train(args, train_loader, step, epoch, dev, verbose=rank == 0)
        if rank == 0:
            metrics = evaluate(args, eval_loader, model, ids, dev)
            metrics.update(train_loss=loss_avg, train_top1=acc_avg)
            recent = metrics["Bleu_4"]
            is_best = recent > best_bleu4
            best_bleu4 = max(recent, best_bleu4)
            state = {'epoch': epoch, 'bleu-4': recent,
                     'encoder_image': {k: v.detach().clone() for k, v in model.encoder.state_dict().items()},
                     'decoder': {k: v.detach().clone() for k, v in model.decoder.state_dict().items()},
                     'encoder_image_optimizer': step.opt.state_dict(), 'decoder_optimizer': step.dec_opt.state_dict()}
            torch.save(state, osp(args.save_path, 'checkpoint_' + args.dataset + '.pth.tar'))
            if is_best:
                torch.save(state, osp(args.save_path, 'BEST_checkpoint_' + args.dataset + '.pth.tar'))
            torch.save(state, osp(args.save_path, 'checkpoint_' + args.dataset + '_epoch_' + str(epoch) + '.pth.tar'))
        if world > 1:
            dist.barrier()
    return metrics


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    train_validate(args)
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
