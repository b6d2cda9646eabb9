"""Multi-GPU plumbing of the AMP path: one process per GPU, envs sharded, one gradient all-reduce.

The reference scales exactly this way through skrl (``train.py:53-58, 184-196``: ``--distributed`` +
``torch.distributed.run``; each rank owns ``num_envs`` envs on ``cuda:{local_rank}``); motion sampling, AMP observations
and the style reward need no communication.  The only collective on the path is skrl's ``Model.reduce_parameters``:
all-reduce(SUM) of the flattened gradients divided by the world size, once per model per mini-batch.  Here the three
models' gradients travel as ONE flat fp32 buffer per call (one NCCL launch over NVLink instead of three).
"""

from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib


def shard_envs(total_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """``[begin, end)`` of the envs owned by ``rank``; the remainder goes to the lowest ranks."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(total_envs, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class _DevicePointer:
    """Zero-copy view of library-owned device memory for ``torch.as_tensor`` (CUDA array interface v2)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class GradientBucket:
    """One rank's flat fp32 gradient bucket, averaged over the node's ranks by ONE peer-memory kernel per rank.

    Replaces the collective inside skrl ``Model.reduce_parameters`` (``train.py:53-58, 184-196`` switch it on): instead of
    copy-in / NCCL all-reduce / divide / copy-out, gradient producers (``AmpDiscriminatorUpdate(grad_weights=...)``) write
    straight into views of ``bucket.flat`` and ``all_reduce_mean()`` runs the two-shot NVLink kernel of
    ``csrc/amp_bucket.cu`` in place.  Construction is collective (every rank of ``group`` must create its bucket at the
    same point, with the same ``numel``): the CUDA IPC handles are exchanged through ``torch.distributed``.
    """

    def __init__(self, numel: int, device, group=None):
        self.device = _lib.require_cuda(device)
        self.group = group
        distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if distributed else 1
        self.rank = dist.get_rank(group) if distributed else 0
        lib, _stream = _lib.enter(self.device)
        h = C.c_void_p()
        _lib.check(lib.amp_bucket_create(int(numel), self.world, self.rank, C.byref(h)))
        self._h = h
        self.numel = int(numel)
        self.capacity = int(lib.amp_bucket_floats(h))
        self.flat = torch.as_tensor(_DevicePointer(int(lib.amp_bucket_data(h)), self.capacity), device=self.device)
        if self.world > 1:
            blob = (C.c_ubyte * 128)()
            lib, _stream = _lib.enter(self.device)
            _lib.check(lib.amp_bucket_export(h, blob))
            mine = torch.tensor(list(blob), dtype=torch.uint8, device=self.device)
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=group)
            everyone = bytes(torch.cat(gathered).cpu().numpy().tobytes())
            lib, _stream = _lib.enter(self.device)  # the peers' memory is mapped into THIS device's context
            _lib.check(lib.amp_bucket_connect(h, everyone))
            dist.barrier(group)  # every rank has mapped every peer before the first all-reduce touches peer memory

    def carve(self, shapes: Sequence[Sequence[int]]) -> List[torch.Tensor]:
        """Consecutive views of the bucket with the given shapes (e.g. the discriminator's six gradient tensors)."""
        out, offset = [], 0
        for shape in shapes:
            n = 1
            for d in shape:
                n *= int(d)
            if offset + n > self.numel:
                raise ValueError("shapes exceed the bucket")
            out.append(self.flat[offset : offset + n].view(*shape))
            offset += n
        return out

    def all_reduce_mean(self, offset: int = 0, count: Optional[int] = None) -> torch.Tensor:
        """``flat[offset : offset + count]`` becomes its mean over the ranks, in place, on torch's current stream."""
        count = self.numel - offset if count is None else int(count)
        lib, stream = _lib.enter(self.device)
        _lib.check(lib.amp_bucket_allreduce_mean(self._h, int(offset), count, stream))
        return self.flat[offset : offset + count]

    def poll_status(self) -> int:
        """Synchronising check: 0, or a bit mask (1: a peer never announced its data, 2: never announced its stores)."""
        lib, stream = _lib.enter(self.device)
        status = C.c_uint32(0)
        _lib.check(lib.amp_bucket_poll_status(self._h, stream, C.byref(status)))
        return int(status.value)

    def last_timing_us(self) -> dict:
        """Phases of the last all-reduce on this rank (device ``%globaltimer``): waiting for the peers' data, reducing and
        publishing the own slice, waiting for the peers' stores.  Synchronises."""
        lib, stream = _lib.enter(self.device)
        t = (C.c_uint64 * 4)()
        _lib.check(lib.amp_bucket_last_timing(self._h, stream, t))
        return {"barrier_a": (t[1] - t[0]) / 1e3, "reduce_publish": (t[2] - t[1]) / 1e3, "barrier_b": (t[3] - t[2]) / 1e3}

    def close(self):
        if getattr(self, "_h", None) is not None:
            self.flat = None
            try:
                _lib.load().amp_bucket_destroy(self._h)
            finally:
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals may already be gone; the driver reclaims the handle
            pass


def reduce_parameters(parameter_groups: Iterable[Sequence[torch.nn.Parameter]], group=None, flat: torch.Tensor | None = None,
                      bucket: Optional[GradientBucket] = None) -> torch.Tensor | None:
    """Average the gradients of every parameter in ``parameter_groups`` (e.g. policy, value, discriminator) over ranks.

    Semantics of skrl ``Model.reduce_parameters`` (missing grads count as zeros; SUM then divide by world size), but one
    flat buffer and one collective for all models.  ``flat`` may be a preallocated buffer of the right size (returned for
    reuse).  With a ``bucket`` (:class:`GradientBucket`, single node) the exchange is the peer-memory kernel instead of NCCL.
    No-op when torch.distributed is not initialised or the world has one rank.
    """
    if not dist.is_available() or not dist.is_initialized():
        return flat
    world = dist.get_world_size(group)
    params = [p for grp in parameter_groups for p in grp]
    if not params or world == 1:
        return flat
    total = sum(p.numel() for p in params)
    ref = params[0]
    if bucket is not None:
        if bucket.numel < total:
            raise ValueError(f"bucket of {bucket.numel} floats is smaller than the {total} gradient elements")
        flat = bucket.flat[:total]
    if flat is None or flat.numel() != total or flat.device != ref.device:
        flat = torch.empty(total, dtype=torch.float32, device=ref.device)
    offset = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            flat[offset : offset + n].zero_()
        else:
            flat[offset : offset + n].copy_(p.grad.reshape(-1))
        offset += n
    if bucket is not None:
        bucket.all_reduce_mean(0, total)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    offset = 0
    for p in params:
        n = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[offset : offset + n].view_as(p.grad))
        offset += n
    return flat
