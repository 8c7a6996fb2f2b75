"""Host-side serving loop around the hot path: pinned host pixels (or, with `input_shape`, any encoder input such as the
[4, 1001, 64] mel features of the CLAP tower) in, token ids in pinned host memory out. This is what `bench.py` measures;
the reference moves one image at a time with a blocking `.to(device)` and decodes it before touching the next
(docs/inference.md:22-27, inference/demo.py:30-45, preprocess/mapper.py:17).

Two overlaps, both across consecutive batches:

* the host->device copy of batch i+1 runs on a copy stream while batch i computes (double-buffered device staging);
* with `partition_sms > 0` the GPU itself is split into two SM partitions (CUDA green contexts, engine.SmPartition): the
  image tower + mapper + GPT-2 prefill of batch i+1 — tensor-bound, and power-bound on a B200 — run on the large partition
  while the 19 decode steps of batch i — a chain of ~3200 dependent launches that leaves most of the machine idle — run
  on the small one. Two GPT-2 engines (same weights, separate KV caches) alternate between batches, so a prefill never
  touches the cache a decode loop is still reading. Token ids are the same function of the pixels either way.
  At the two ends of a stream of batches one half of that pipeline has nothing to do: the front of the FIRST batch (no
  decode loop is running yet) and the decode loop of the LAST batch (no front follows) are enqueued on a whole-device
  stream instead of their partition, sized for all SMs.
"""
from __future__ import annotations

import contextlib
from typing import Callable, Iterable, Iterator, Optional, Tuple

import torch

from clipcap_b200.distributed import all_gather_prefix


class CaptionPipeline:
    def __init__(self, encode_fn: Callable, model, batch: int, image_size: int = 224, entry_length: int = 20,
                 stop_token: int = 50256, device="cuda", pixel_dtype: torch.dtype = torch.float32,
                 prefix_all: Optional[torch.Tensor] = None, input_shape: Optional[Tuple[int, ...]] = None,
                 partition_sms: int = 0, mode: str = "greedy", beam: int = 1, comm=None,
                 prefix_dtype: Optional[torch.dtype] = None, prefill_defer: int = 0):
        """`comm` (distributed.PrefixComm, N > 1): the prefix all-gather runs through the C ABI, in place — the mapper
        writes this rank's slot of a [world * batch, K, d] buffer of `prefix_dtype` and cc_allgather_prefix completes it
        on a side stream (decode reads only the local slot, so nothing waits for the peers except the end of the step).
        `prefix_all` (legacy): torch.distributed all-gather into the given tensor. `prefix_dtype`: dtype of the prefix
        handed from the mapper to the language model (default: the encoder output's). `prefill_defer` (with a partition):
        that many trailing prefill blocks run at the head of the decode loop, on the small partition — the balance knob
        between the two partitions (cc_gpt2_set_prefill_defer); ids do not depend on it."""
        self.encode_fn, self.model = encode_fn, model
        self.entry_length, self.stop_token = entry_length, stop_token
        self.mode, self.beam = mode, (beam if mode == "beam" else 1)
        self.device = torch.device(device)
        self.prefix_all = prefix_all
        self.comm, self.prefix_dtype = comm, prefix_dtype
        self.batch = batch
        self.prefill_defer = prefill_defer if partition_sms > 0 else 0
        if comm is not None:
            K, d = model.transformer_mapper.prefix_length, model.transformer_mapper.lm_embedding_size
            self.prefix_all = torch.empty(comm.world * batch, K, d, device=self.device, dtype=prefix_dtype or pixel_dtype)
            self.comm_stream = torch.cuda.Stream(device=self.device)
            self._mapped = torch.cuda.Event()     # this rank's slot of prefix_all is written
            self._gathered = torch.cuda.Event()   # the all-gather of the current batch has completed
            self._gather_pending = False
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
        self.trace = None  # set to a list to collect per-batch stage events (bench.py: live stage times inside the step)
        self._lm = None
        self.partition = None
        if partition_sms > 0:
            from clipcap_b200.engine import SmPartition
            self.partition = SmPartition(partition_sms, self.device)
            self._prefilled = [torch.cuda.Event() for _ in range(2)]  # engine i holds the prefix + first token of its batch
            self._decoded = [torch.cuda.Event() for _ in range(2)]    # engine i is free for the next prefill
            self._whole_stream = torch.cuda.Stream(device=self.device)  # primary context: every SM
            self._front_done = torch.cuda.Event()  # the image tower / mapper workspaces are free for the next front
            self._front_pending = False

    # ------------------------------------------------------------------ staging
    def _stage(self, slot: int, pixels_host: torch.Tensor, first_use: bool) -> None:
        if not pixels_host.is_pinned():
            raise ValueError("CaptionPipeline needs pinned host batches (tensor.pin_memory())")
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self._consumed[slot])
            self._px[slot][:pixels_host.shape[0]].copy_(pixels_host, non_blocking=True)
            self._copied[slot].record(self.copy_stream)

    # ------------------------------------------------------------------ one batch, enqueued (no host sync)
    def _mark(self, rec, name: str, stream) -> None:
        if rec is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            rec[name] = ev

    def _engines(self):
        if self._lm is None:
            K = self.model.transformer_mapper.prefix_length
            n = 2 if self.partition is not None else 1
            self._lm = self.model.language_model.decode_engines(n, self.batch * self.beam, K + self.entry_length)
        return self._lm

    @contextlib.contextmanager
    def _whole(self):
        """The whole-device stream with launch heuristics sized for every SM (the ends of a run, see the module text)."""
        from clipcap_b200 import _ffi
        lib = _ffi.lib()
        before = lib.cc_get_sm_budget()
        lib.cc_set_sm_budget(0)
        try:
            with torch.cuda.stream(self._whole_stream):
                yield self._whole_stream
        finally:
            lib.cc_set_sm_budget(before)

    def _enqueue(self, pixels: torch.Tensor, slot: int, index: int, staged: bool, first: bool = False,
                 last: bool = False) -> None:
        """Image tower -> mapper -> [prefix all-gather] -> prefill + first token on the front stream, the remaining decode
        steps + the copy of the ids to the host on the back stream. Without a partition both are the caller's stream.
        `first` / `last`: no decode loop is running / no front follows — that half runs on the whole device."""
        part = self.partition
        rows = pixels.shape[0]
        K = self.model.transformer_mapper.prefix_length
        eng = self._engines()[slot if part is not None else 0]
        eng.set_prefill_defer(self.prefill_defer)
        kw = dict(mode=self.mode, beam=self.beam, entry_length=self.entry_length, stop_token=self.stop_token)
        rec = {} if self.trace is not None else None
        compute = torch.cuda.current_stream(self.device)
        front_ctx = contextlib.nullcontext(compute) if part is None else (self._whole() if first else part.on(0))
        back_ctx = contextlib.nullcontext(compute) if part is None else (self._whole() if last else part.on(1))
        with front_ctx as front:
            if part is not None and self._front_pending:
                front.wait_event(self._front_done)  # consecutive fronts may sit on different streams (whole device / partition)
            if staged:
                front.wait_event(self._copied[slot])
            self._mark(rec, "front0", front)
            emb = self.encode_fn(pixels)
            self._mark(rec, "vit", front)
            if self.comm is not None:
                if self._gather_pending:
                    front.wait_event(self._gathered)   # the previous batch's collective no longer touches the buffer
                prefix = self.model.transformer_mapper(emb, out=self.comm.slot(self.prefix_all)[:rows])
            elif self.prefix_dtype is not None:
                prefix = self.model.transformer_mapper(emb, out_dtype=self.prefix_dtype)
            else:
                prefix = self.model.transformer_mapper(emb)
            self._mark(rec, "mapper", front)
            if staged:
                self._consumed[slot].record(front)
            work = None
            if self.comm is not None:   # N > 1: every rank ends up with every rank's prefixes
                self._mapped.record(front)
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(self._mapped)
                    self.comm.all_gather_(self.prefix_all)
                    self._gathered.record(self.comm_stream)
                self._gather_pending = True
            elif self.prefix_all is not None:
                _, work = all_gather_prefix(prefix, None, self.prefix_all, async_op=True)
            if part is not None and index >= 2:
                front.wait_event(self._decoded[slot])  # this engine's previous batch has left the decode partition
            eng.prefill(prefix, **kw)
            self._mark(rec, "prefill", front)
            if work is not None:
                work.wait()
            if part is not None:
                self._prefilled[slot].record(front)
                self._front_done.record(front)
                self._front_pending = True
        with back_ctx as back:
            if part is not None:
                back.wait_event(self._prefilled[slot])
            self._mark(rec, "dec0", back)
            toks, lens, _ = eng.decode(rows, K, **kw)
            self._mark(rec, "dec1", back)
            if part is not None:
                self._decoded[slot].record(back)
            self._tok[slot][:rows].copy_(toks, non_blocking=True)
            self._len[slot][:rows].copy_(lens, non_blocking=True)
            if self.comm is not None:
                back.wait_event(self._gathered)   # a finished step means the gathered prefixes are complete, too
            self._done[slot].record(back)
        if rec is not None:
            self.trace.append(rec)

    # ------------------------------------------------------------------ the loop
    def run(self, batches: Iterable[torch.Tensor], resident: bool = False) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Yields (tokens [B, entry_length] int32, lengths [B] int32) in pinned host memory, one pair per input batch, in
        order. The tensors of a yielded pair are reused two batches later. `resident=True`: the batches are device tensors
        already (no staging copy)."""
        it = iter(batches)
        compute = torch.cuda.current_stream(self.device)
        try:
            nxt = next(it)
        except StopIteration:
            return
        if not resident:
            self._stage(0, nxt, True)
        elif self.partition is not None:
            # the caller produced the resident batches on its current stream: order both partitions after it
            ready = torch.cuda.Event()
            ready.record(compute)
            for st in (*self.partition.streams, self._whole_stream):
                st.wait_event(ready)
        i = 0
        pending = None  # (slot, rows) whose results have been enqueued but not yet handed out
        while nxt is not None:
            slot, rows = i & 1, nxt.shape[0]
            try:
                upcoming = next(it)
            except StopIteration:
                upcoming = None
            if upcoming is not None and not resident:  # copy of batch i+1 overlaps the compute of batch i
                self._stage(slot ^ 1, upcoming, i == 0)
            pixels = nxt if resident else self._px[slot][:rows]
            self._enqueue(pixels, slot, i, not resident, first=(i == 0), last=(upcoming is None))
            if pending is not None:  # hand out batch i-1 while batch i runs
                ps, pr = pending
                self._done[ps].synchronize()
                yield self._tok[ps][:pr], self._len[ps][:pr]
            pending = (slot, rows)
            nxt = upcoming
            i += 1
        if pending is not None:
            ps, pr = pending
            self._done[ps].synchronize()
            yield self._tok[ps][:pr], self._len[ps][:pr]
