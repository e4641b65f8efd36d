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

from typing import Iterable, Iterator, Optional, Sequence, Tuple

import torch


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
