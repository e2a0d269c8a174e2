"""Double-buffered host -> device staging of (x, y) batches on a side stream, so the pinned-memory copy of
step i+1 overlaps the rollout of step i (the reference's DataLoader does the same with pin_memory +
non_blocking copies: satflow/data/datamodules.py:104-154)."""
from __future__ import annotations

from typing import Iterable, Iterator, Tuple

import torch


class DevicePrefetcher:
    def __init__(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], device, depth: int = 2):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.slots = []  # (x_dev, y_dev, ready_event, consumed_event)
        self.queue = []
        self.n = 0

    def _issue(self) -> bool:
        try:
            x, y = next(self.it)
        except StopIteration:
            return False
        i = self.n % self.depth
        if len(self.slots) <= i:
            self.slots.append([torch.empty(x.shape, dtype=x.dtype, device=self.device),
                               torch.empty(y.shape, dtype=y.dtype, device=self.device),
                               torch.cuda.Event(), torch.cuda.Event()])
            self.slots[i][3].record(torch.cuda.current_stream(self.device))
        xd, yd, ready, consumed = self.slots[i]
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(consumed)  # the step that last used this slot has finished
            xd.copy_(x, non_blocking=True)
            yd.copy_(y, non_blocking=True)
            ready.record(self.stream)
        self.queue.append(i)
        self.n += 1
        return True

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        return self

    def __next__(self):
        if not self.queue and not self._issue():
            raise StopIteration
        i = self.queue.pop(0)
        xd, yd, ready, consumed = self.slots[i]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ready)
        self._issue()  # start the next batch's copy behind this step's compute
        self._last = i
        return xd, yd

    def done_with_current(self) -> None:
        """Call after the step's kernels are enqueued: marks the slot reusable once they finish."""
        self.slots[self._last][3].record(torch.cuda.current_stream(self.device))
