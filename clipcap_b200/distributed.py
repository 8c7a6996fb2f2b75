"""Multi-GPU plumbing of the captioning path: one process per GPU, the batch is split contiguously by rank,
every rank encodes + maps its own images, the prefix embeddings are all-gathered so every rank holds the full
[B, K, d] tensor (the only exchange step of the path), decode stays data-parallel on the rank that owns the image, and
the token ids are gathered at the end. torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the transport; the
collectives are issued on the current stream, nothing here synchronises the host.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n units owned by `rank`: the first n % world ranks get one extra unit."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request n={n} rank={rank} world={world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_prefix(prefix_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                      out: Optional[torch.Tensor] = None, async_op: bool = False):
    """[B_local, K, d] on every rank -> [world * B_local, K, d] on every rank, rank r's block at rows r*B_local.. .
    Equal shard sizes (weak scaling: B per GPU fixed); a single in-place all-gather (ncclAllGather over NVLink).
    async_op=True returns (out, work): the collective runs on the communicator's stream behind the producer of
    `prefix_local`; `work.wait()` orders the caller's stream after it."""
    if not dist.is_available() or not dist.is_initialized():
        return (prefix_local, None) if async_op else prefix_local
    world = dist.get_world_size(group)
    if world == 1:
        return (prefix_local, None) if async_op else prefix_local
    prefix_local = prefix_local.contiguous()
    shape = (world * prefix_local.shape[0],) + tuple(prefix_local.shape[1:])
    if out is None:
        out = torch.empty(shape, dtype=prefix_local.dtype, device=prefix_local.device)
    elif tuple(out.shape) != shape or out.dtype != prefix_local.dtype:
        raise ValueError(f"all_gather_prefix: out must be {shape} {prefix_local.dtype}, got {tuple(out.shape)} {out.dtype}")
    work = dist.all_gather_into_tensor(out, prefix_local, group=group, async_op=async_op)
    return (out, work) if async_op else out


class PrefixComm:
    """cc_comm_* — the C-ABI collective of the path: one NCCL communicator over the ranks of the job and the in-place
    all-gather of the prefix buffer (each rank's mapper has written its own slot). The 128-byte NCCL id is made
    on rank 0 and handed to the others through the launcher's process group (any backend)."""

    def __init__(self, rank: int, world: int, unique_id: bytes, device="cuda"):
        import ctypes as C
        from clipcap_b200 import _ffi
        self.rank, self.world = rank, world
        self.device = torch.device(device)
        self._h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        with torch.cuda.device(self.device):
            _ffi.check(_ffi.lib().cc_comm_create(C.byref(self._h), buf, rank, world))

    @staticmethod
    def make_unique_id() -> bytes:
        import ctypes as C
        from clipcap_b200 import _ffi
        buf = C.create_string_buffer(128)
        _ffi.check(_ffi.lib().cc_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_process_group(cls, device="cuda", group: Optional[dist.ProcessGroup] = None) -> "PrefixComm":
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.make_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(rank, world, box[0], device)

    def slot(self, prefix_all: torch.Tensor) -> torch.Tensor:
        """This rank's block of the gathered buffer [world * B_local, K, d] (a view: hand it to the mapper as `out`)."""
        n = prefix_all.shape[0] // self.world
        return prefix_all[self.rank * n:(self.rank + 1) * n]

    def all_gather_(self, prefix_all: torch.Tensor) -> torch.Tensor:
        """In place on the current stream: slot `rank` already holds this rank's prefixes; afterwards every slot is filled."""
        from clipcap_b200 import _ffi
        if not prefix_all.is_contiguous() or prefix_all.shape[0] % self.world != 0:
            raise ValueError("prefix_all must be contiguous [world * B_local, K, d]")
        per_rank = prefix_all.numel() * prefix_all.element_size() // self.world
        with torch.cuda.device(prefix_all.device):
            _ffi.check(_ffi.lib().cc_allgather_prefix(self._h, prefix_all.data_ptr(), per_rank, _ffi.current_stream_ptr()))
        return prefix_all

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            from clipcap_b200 import _ffi
            _ffi.lib().cc_comm_destroy(self._h)
            self._h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def gather_tokens(tokens_local: torch.Tensor, lengths_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """Token ids [B_local, EL] + lengths [B_local] of every rank -> ([world*B_local, EL], [world*B_local]) on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tokens_local, lengths_local
    world = dist.get_world_size(group)
    toks = torch.empty((world * tokens_local.shape[0], tokens_local.shape[1]), dtype=tokens_local.dtype,
                       device=tokens_local.device)
    lens = torch.empty((world * lengths_local.shape[0],), dtype=lengths_local.dtype, device=lengths_local.device)
    dist.all_gather_into_tensor(toks, tokens_local.contiguous(), group=group)
    dist.all_gather_into_tensor(lens, lengths_local.contiguous(), group=group)
    return toks, lens


def caption_step(encode_fn, model, pixels_local: torch.Tensor, entry_length: int, stop_token: int,
                 prefix_all: Optional[torch.Tensor] = None, group: Optional[dist.ProcessGroup] = None,
                 mode: str = "greedy", beam: int = 1):
    """One pass of the hot path on this rank's shard: ViT -> mapper -> prefix all-gather -> greedy (or beam) decode of
    the local rows. Returns (tokens, lengths, prefix_all). Decode reads only this rank's own prefix, so the all-gather
    runs behind it on the communicator's stream and is joined at the end of the step: a rank never idles waiting for a
    slower peer's mapper before it may start decoding."""
    from clipcap_b200.inference.base import generate_beam_tokens, generate_greedy_tokens
    emb = encode_fn(pixels_local)
    prefix = model.transformer_mapper(emb)
    gathered, work = all_gather_prefix(prefix, group, prefix_all, async_op=True)
    if mode == "beam":
        tokens, lengths, _ = generate_beam_tokens(model, prefix, None, beam, entry_length, 1.0, stop_token)
    else:
        tokens, lengths, _ = generate_greedy_tokens(model, prefix, entry_length, stop_token)
    if work is not None:
        work.wait()  # the caller's stream now sees the complete [world * B, K, d] tensor
    return tokens, lengths, gathered
