"""Host-side serving loop around the hot path: pinned host pixels (or, with `input_shape`, any encoder input such as the
[4, 1001, 64] mel features of the CLAP tower) in, token ids in pinned host memory out, with the host->device copy of
batch i+1 running on a copy stream while batch i is on the compute stream (double-buffered device staging). This is what `bench.py`'s end-to-end number measures; the reference moves one image at a time with a blocking
`.to(device)` (docs/inference.md:22-27, preprocess/mapper.py:17).
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Optional, Tuple

import torch

from clipcap_b200.distributed import caption_step


class CaptionPipeline:
    def __init__(self, encode_fn: Callable, model, batch: int, image_size: int = 224, entry_length: int = 20,
                 stop_token: int = 50256, device="cuda", pixel_dtype: torch.dtype = torch.float32,
                 prefix_all: Optional[torch.Tensor] = None, input_shape: Optional[Tuple[int, ...]] = None):
        self.encode_fn, self.model = encode_fn, model
        self.entry_length, self.stop_token = entry_length, stop_token
        self.device = torch.device(device)
        self.prefix_all = prefix_all
        self.copy_stream = torch.cuda.Stream(device=self.device)
        shape = (batch, *input_shape) if input_shape is not None else (batch, 3, image_size, image_size)
        self._px = [torch.empty(shape, device=self.device, dtype=pixel_dtype) for _ in range(2)]
        self._copied = [torch.cuda.Event() for _ in range(2)]     # staging buffer i holds its batch
        self._consumed = [torch.cuda.Event() for _ in range(2)]   # the compute stream no longer reads buffer i
        self._tok = [torch.empty(batch, entry_length, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._len = [torch.empty(batch, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._done = [torch.cuda.Event() for _ in range(2)]       # results of slot i are in host memory
        self.h2d_bytes_per_batch = self._px[0].numel() * self._px[0].element_size()
        self.d2h_bytes_per_batch = self._tok[0].numel() * 4 + self._len[0].numel() * 4

    def _stage(self, slot: int, pixels_host: torch.Tensor, first_use: bool) -> None:
        if not pixels_host.is_pinned():
            raise ValueError("CaptionPipeline needs pinned host batches (tensor.pin_memory())")
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self._consumed[slot])
            self._px[slot][:pixels_host.shape[0]].copy_(pixels_host, non_blocking=True)
            self._copied[slot].record(self.copy_stream)

    def run(self, batches: Iterable[torch.Tensor]) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Yields (tokens [B, entry_length] int32, lengths [B] int32) in pinned host memory, one pair per input batch, in
        order. The tensors of a yielded pair are reused two batches later."""
        it = iter(batches)
        compute = torch.cuda.current_stream(self.device)
        try:
            nxt = next(it)
        except StopIteration:
            return
        self._stage(0, nxt, True)
        i = 0
        pending = None  # (slot, rows) whose results have been enqueued but not yet handed out
        while nxt is not None:
            slot, rows = i & 1, nxt.shape[0]
            try:
                upcoming = next(it)
            except StopIteration:
                upcoming = None
            if upcoming is not None:  # copy of batch i+1 overlaps the compute of batch i
                self._stage(slot ^ 1, upcoming, i == 0)
            compute.wait_event(self._copied[slot])
            toks, lens, _ = caption_step(self.encode_fn, self.model, self._px[slot][:rows], self.entry_length,
                                         self.stop_token, self.prefix_all)
            self._consumed[slot].record(compute)
            if pending is not None:  # hand out batch i-1 while batch i runs
                ps, pr = pending
                self._done[ps].synchronize()
                yield self._tok[ps][:pr], self._len[ps][:pr]
            self._tok[slot][:rows].copy_(toks, non_blocking=True)
            self._len[slot][:rows].copy_(lens, non_blocking=True)
            self._done[slot].record(compute)
            pending = (slot, rows)
            nxt = upcoming
            i += 1
        if pending is not None:
            ps, pr = pending
            self._done[ps].synchronize()
            yield self._tok[ps][:pr], self._len[ps][:pr]
