"""Host -> device input staging for the train step (first piece): double-buffered copies of
pinned host batches on a dedicated copy stream, one batch ahead of the step that consumes them, so the 58.7 MB
per step of a batch-32 BCD pair (scripts/train_BCD.py:182-185: `img[:, 0:3].cuda()`, `img[:, 3:6].cuda()`,
`target.cuda()` — synchronous copies on the compute stream in the reference) overlap the previous step's kernels.

    for pre, post, target in DevicePrefetcher(host_batches, device):
        loss = step(pre, post, target)

Every yielded tuple lives in one of two device buffer sets; a set is overwritten only after the work the consumer
enqueued while holding it has been passed on the consumer's stream (event recorded when the next item is requested).
"""
from __future__ import annotations

import random
from typing import Iterable, Iterator, Optional, Sequence, Tuple

import torch

from . import _lib as L


class DevicePrefetcher:
    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device) -> None:
        self.batches = batches
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches onto a CUDA device")
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]      # never recorded yet: waiting on them is a no-op
        self.bytes_staged = 0

    def _stage(self, batch: Optional[Sequence[torch.Tensor]], slot: int) -> Optional[Tuple[torch.Tensor, ...]]:
        if batch is None:
            return None
        batch = tuple(batch)
        cur = self.bufs[slot]
        if cur is None or len(cur) != len(batch) or any(c.shape != b.shape or c.dtype != b.dtype for c, b in zip(cur, batch)):
            # allocated on the consumer's stream (the buffers outlive every use; a ragged last batch re-allocates)
            cur = tuple(torch.empty(b.shape, dtype=b.dtype, device=self.device) for b in batch)
            self.bufs[slot] = cur
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])
            for dst, src in zip(cur, batch):
                dst.copy_(src, non_blocking=True)                     # asynchronous when `src` is pinned
                self.bytes_staged += src.numel() * src.element_size()
            self.ready[slot].record(self.copy_stream)
        return cur

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        it = iter(self.batches)
        slot = 0
        nxt = self._stage(next(it, None), slot)
        while nxt is not None:
            cur, cur_slot = nxt, slot
            slot ^= 1
            nxt = self._stage(next(it, None), slot)                   # in flight while `cur` is being consumed
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self.ready[cur_slot])
            yield cur
            self.consumed[cur_slot].record(torch.cuda.current_stream(self.device))


# ---------------------------------------------------------------------------------------------------------------
# GPU augmentation (second piece): the reference's per-sample CPU transform chain
# (data/transforms.py: normalize -> scale -> random_crop_resize -> random_flip -> random_exchange -> to_tensor) as one
# kernel per batch on raw uint8 pairs (csrc/augment.cu).  The host copies 1 byte per value instead of 4 and does no
# image arithmetic; the random decisions are drawn here, from Python's `random`, in the reference's order.
# ---------------------------------------------------------------------------------------------------------------
LABEL_CHANNELS = {"bcd": 1, "scd": 3, "bda": 2}


def draw_params(batch: int, in_width: int, task: str = "bcd", train: bool = True, rng=random) -> torch.Tensor:
    """int32 (batch, 8) = {do_crop, x1, y1, flip_rows, flip_cols, exchange_images, exchange_labels01, 0} per sample.
    Training: the draws of random_crop_resize (data/transforms.py:82-99: random() < 0.5, then randint(0, crop_area) for
    x1 and for y1, crop_area = int(7 / 224 * in_width)), random_flip (:101-114: two random() < 0.5) and random_exchange
    (:116-125: one random() < 0.5; the SCD variant :300-312 also swaps the two class maps), in that order, sample after
    sample — a `random.seed(s)`-ed run makes the reference's choices.  Validation (train=False): all zeros."""
    if task not in LABEL_CHANNELS:
        raise ValueError(f"draw_params: unknown task {task!r}")
    crop_area = int(7.0 / 224.0 * in_width)
    rows = []
    for _ in range(batch):
        if not train:
            rows.append([0] * 8)
            continue
        do_crop, x1, y1 = 0, 0, 0
        if rng.random() < 0.5:
            do_crop, x1, y1 = 1, rng.randint(0, crop_area), rng.randint(0, crop_area)
        f0 = 1 if rng.random() < 0.5 else 0
        f1 = 1 if rng.random() < 0.5 else 0
        ex = 1 if rng.random() < 0.5 else 0
        rows.append([do_crop, x1, y1, f0, f1, ex, ex if task == "scd" else 0, 0])
    return torch.tensor(rows, dtype=torch.int32)


class GpuAugment:
    """`pre, post, label = GpuAugment(H, W, task)(img_u8, label_u8, params)` on the device.

    img_u8 (B, Hs, Ws, 6) uint8 HWC = [pre RGB | post RGB] as the datasets read them (data/dataset.py:77-84);
    label_u8 (B, Hs, Ws) or (B, Hs, Ws, L) uint8 (BCD: the 0 / 255 change mask; SCD: [class A, class B, change]; BDA:
    [building, damage]) or None; params from `draw_params` (host or device int32).  Returns pre / post (B, 3, H, W)
    float32 — what `Trainer.update_*` takes — and the label as the scripts' tensors: BCD (B, 1, H, W) float in {0, 1},
    SCD / BDA (B, L, H, W) int64.  Sources whose size differs from (H, W) are first rescaled by the same kernel with
    identity parameters (the reference's `scale` step), then augmented from that float intermediate."""

    def __init__(self, in_height: int, in_width: int, task: str = "bcd", mean: float = 0.5, std: float = 0.5):
        if task not in LABEL_CHANNELS:
            raise ValueError(f"GpuAugment: unknown task {task!r}")
        self.H, self.W, self.task, self.mean, self.std = in_height, in_width, task, float(mean), float(std)

    def _launch(self, img, label, params, Hs, Ws, Lc):
        B = img.shape[0]
        dev = img.device
        pre = torch.empty(B, 3, self.H, self.W, device=dev, dtype=torch.float32)
        post = torch.empty_like(pre)
        mode = 0 if self.task == "bcd" else 1
        lab = None
        if label is not None:
            lab = torch.empty(B, Lc, self.H, self.W, device=dev, dtype=torch.float32 if mode == 0 else torch.int64)
        L.check(L.load().c3d_augment_pairs(img.data_ptr(), 1 if img.dtype == torch.float32 else 0,
                                           label.data_ptr() if label is not None else None, params.data_ptr(), B, Hs, Ws,
                                           self.H, self.W, Lc, mode, self.mean, self.std, pre.data_ptr(), post.data_ptr(),
                                           lab.data_ptr() if lab is not None else None,
                                           torch.cuda.current_stream().cuda_stream), "c3d_augment_pairs")
        return pre, post, lab

    def __call__(self, img: torch.Tensor, label: Optional[torch.Tensor], params: Optional[torch.Tensor] = None):
        if not img.is_cuda or img.dtype != torch.uint8 or img.dim() != 4 or img.shape[3] != 6:
            raise RuntimeError("GpuAugment: img must be a CUDA uint8 tensor (B, Hs, Ws, 6); there is no CPU path")
        img = img.contiguous()
        B, Hs, Ws, _ = img.shape
        Lc = LABEL_CHANNELS[self.task]
        if label is not None:
            if label.dtype != torch.uint8 or not label.is_cuda:
                raise RuntimeError("GpuAugment: label must be a CUDA uint8 tensor")
            label = label.reshape(B, Hs, Ws, Lc).contiguous()
        if params is None:
            params = torch.zeros(B, 8, dtype=torch.int32)
        params = params.to(device=img.device, dtype=torch.int32).contiguous()
        if (Hs, Ws) == (self.H, self.W):
            return self._launch(img, label, params, Hs, Ws, Lc)
        # `scale` first (cv2.resize of the normalised image / nearest for the label), then the augmentation on the result
        ident = torch.zeros(B, 8, dtype=torch.int32, device=img.device)
        pre, post, lab = self._launch(img, label, ident, Hs, Ws, Lc)
        img_f = torch.cat([pre, post], 1).permute(0, 2, 3, 1).contiguous()
        lab_u8 = lab.permute(0, 2, 3, 1).to(torch.uint8).contiguous() if lab is not None else None
        return self._launch(img_f, lab_u8, params, self.H, self.W, Lc)


def is_raw_batch(batch) -> bool:
    """True for batches of raw pairs (img uint8 (B, Hs, Ws, 6), label uint8) -- what a dataset yields when it leaves the
    transform chain to the GPU -- as opposed to the reference's float (B, 6, H, W) tensors."""
    img = batch[0]
    return img.dtype == torch.uint8 and img.dim() == 4 and img.shape[-1] == 6


def augment_raw_batch(batch, task: str, in_height: int, in_width: int, train: bool):
    """(img_u8, label_u8) on the device -> (pre, post, label) through GpuAugment, decisions drawn for this batch."""
    img, label = batch[0], batch[1]
    params = draw_params(img.shape[0], in_width, task, train)
    return GpuAugment(in_height, in_width, task)(img, label, params)
