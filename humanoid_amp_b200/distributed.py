"""Multi-GPU plumbing of the AMP path: one process per GPU, envs sharded, one gradient all-reduce.

The reference scales exactly this way through skrl (``train.py:53-58, 184-196``: ``--distributed`` +
``torch.distributed.run``; each rank owns ``num_envs`` envs on ``cuda:{local_rank}``); motion sampling, AMP observations
and the style reward need no communication.  The only collective on the path is skrl's ``Model.reduce_parameters``:
all-reduce(SUM) of the flattened gradients divided by the world size, once per model per mini-batch.  Here the three
models' gradients travel as ONE flat fp32 buffer per call (one NCCL launch over NVLink instead of three).
"""

from __future__ import annotations

import ctypes as C
import os
import socket
import struct
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


def swap_fds(own_fd: int, root_fd: int, rank: int, world: int, group=None, device=None, timeout: float = 60.0):
    """Every rank hands a file descriptor (``own_fd``) to every other rank, and rank 0 a second one (``root_fd``) on top:
    ``SCM_RIGHTS`` over unix sockets in the abstract namespace, named by a token rank 0 draws.  Returns ``(fds, root)``: ``fds[p]``
    is this process's descriptor for rank p's ``own_fd`` (-1 for itself), ``root`` the one for rank 0's ``root_fd`` (-1 on rank 0).
    Collective; any failure surfaces as ``OSError`` AFTER the rendezvous barrier, so the ranks stay in step.  The descriptors of
    :class:`GradientBucket`'s shared form travel this way (CUDA VMM allocations and the multicast object export as POSIX fds,
    which mean nothing in another process until the kernel has duplicated them into it)."""
    token = torch.randint(0, 2**62, (1,), dtype=torch.int64)
    if device is not None and dist.get_backend(group) == "nccl":
        token = token.to(device)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(token, src=src, group=group)
    tag = int(token.item())
    name = lambda r: f"\0amp_b200_bucket_{tag:x}_{r}"  # noqa: E731
    server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    got, root = [-1] * world, -1
    try:
        bind_error = None
        try:
            server.bind(name(rank))
            server.listen(world)
            server.settimeout(timeout)
        except OSError as e:  # still walk the barrier: the other ranks are waiting in it
            bind_error = e
        dist.barrier(group)  # everyone is listening
        if bind_error is not None:
            raise bind_error
        for p in range(world):
            if p == rank:
                continue
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                c.settimeout(timeout)
                c.connect(name(p))
                socket.send_fds(c, [struct.pack("i", rank)], [own_fd] + ([root_fd] if rank == 0 else []))
        for _ in range(world - 1):
            conn, _addr = server.accept()
            with conn:
                conn.settimeout(timeout)
                msg, received, _flags, _a = socket.recv_fds(conn, 4, 2)
                sender = struct.unpack("i", msg)[0]
                got[sender] = received[0]
                if sender == 0:
                    root = received[1]
    finally:
        server.close()
    return got, root


class _SharedUnavailable(Exception):
    pass


class GradientBucket:
    """One rank's flat fp32 gradient bucket, averaged over the node's ranks by ONE peer-memory kernel per rank.

    Replaces the collective inside skrl ``Model.reduce_parameters`` (``train.py:53-58, 184-196`` switch it on): instead of
    copy-in / NCCL all-reduce / divide / copy-out, gradient producers (``AmpDiscriminatorUpdate(grad_weights=...)``) write
    straight into views of ``bucket.flat`` and ``all_reduce_mean()`` runs one kernel of ``csrc/amp_bucket.cu`` per rank in
    place: in the NVSwitch (``multimem.ld_reduce`` / ``multimem.st`` on a multicast mapping, ``bucket.in_switch``) from 4
    ranks up when the node supports it, else two-shot over peer memory (``AMP_B200_BUCKET_IN_SWITCH`` = 0 / 1 overrides).  Construction is collective (every rank of ``group`` must create its bucket at the
    same point, with the same ``numel``): the CUDA IPC handles are exchanged through ``torch.distributed``.
    """

    def __init__(self, numel: int, device, group=None):
        self.device = _lib.require_cuda(device)
        self.group = group
        distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if distributed else 1
        self.rank = dist.get_rank(group) if distributed else 0
        self.numel = int(numel)
        self._h = None
        # Shared form (the all-reduce runs in the NVSwitch) from 4 ranks up -- measured on 2.65 M floats: 52.9 us against 69.4 us
        # over peer memory at 8 ranks, 54.0 against 55.6 at 4, but 56.4 against 40.3 at 2 (profiles/r02_allreduce_in_switch.md).
        # Every rank must manage it, or every rank falls back to the peer-memory form.  AMP_B200_BUCKET_IN_SWITCH (read here,
        # once): 0 never, 1 from 2 ranks up.
        knob = os.environ.get("AMP_B200_BUCKET_IN_SWITCH", "")
        if self.world > 1 and knob != "0" and (self.world >= 4 or knob == "1"):
            self._h = self._create_shared()
        if self._h is None:
            self._h = self._create_peer()
        lib, _stream = _lib.enter(self.device)
        self.capacity = int(lib.amp_bucket_floats(self._h))
        self.in_switch = bool(lib.amp_bucket_in_switch(self._h))
        self.flat = torch.as_tensor(_DevicePointer(int(lib.amp_bucket_data(self._h)), self.capacity), device=self.device)

    def _all_agree(self, ok: bool) -> bool:
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(flag.item())

    def _gather_bytes(self, blob: bytes) -> bytes:
        mine = torch.tensor(list(blob), dtype=torch.uint8, device=self.device)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=self.group)
        return bytes(torch.cat(gathered).cpu().numpy().tobytes())

    def _create_peer(self):
        """cudaMalloc + legacy CUDA IPC: peer loads / stores over NVLink (``amp_bucket_create``)."""
        lib, _stream = _lib.enter(self.device)
        h = C.c_void_p()
        _lib.check(lib.amp_bucket_create(self.numel, self.world, self.rank, C.byref(h)))
        if self.world > 1:
            blob = (C.c_ubyte * 128)()
            _lib.check(lib.amp_bucket_export(h, blob))
            everyone = self._gather_bytes(bytes(blob))
            lib, _stream = _lib.enter(self.device)  # the peers' memory is mapped into THIS device's context
            _lib.check(lib.amp_bucket_connect(h, everyone))
            dist.barrier(self.group)  # every rank has mapped every peer before the first all-reduce touches peer memory
        return h

    def _create_shared(self):
        """VMM allocation shared as POSIX fds + one NVSwitch multicast object (``amp_bucket_create_shared``); ``None`` when
        any rank cannot (no multicast support, fd passing refused): the caller then builds the peer-memory form."""
        lib, _stream = _lib.enter(self.device)
        h = C.c_void_p()
        rc = lib.amp_bucket_create_shared(self.numel, self.world, self.rank, C.byref(h))
        if not self._all_agree(rc == 0):
            if rc == 0:
                lib.amp_bucket_destroy(h)
            return None
        fds: List[int] = []
        try:
            blob = (C.c_ubyte * 64)()
            data_fd, mc_fd = C.c_int32(-1), C.c_int32(-1)
            rc = lib.amp_bucket_export_shared(h, blob, C.byref(data_fd), C.byref(mc_fd))
            fds += [fd for fd in (data_fd.value, mc_fd.value) if fd >= 0]
            if not self._all_agree(rc == 0):
                raise _SharedUnavailable
            everyone = self._gather_bytes(bytes(blob))
            try:
                peer_fds, peer_mc = swap_fds(data_fd.value, mc_fd.value, self.rank, self.world, self.group, self.device)
                ok = True
            except OSError:
                peer_fds, peer_mc, ok = [-1] * self.world, -1, False
            fds += [fd for fd in peer_fds if fd >= 0] + ([peer_mc] if peer_mc >= 0 else [])
            if not self._all_agree(ok):
                raise _SharedUnavailable
            lib, _stream = _lib.enter(self.device)
            rc = lib.amp_bucket_join_shared(h, peer_mc if self.rank else mc_fd.value)
            if not self._all_agree(rc == 0):  # binding blocks until the whole team has joined: go on only if everyone has
                raise _SharedUnavailable
            lib, _stream = _lib.enter(self.device)
            rc = lib.amp_bucket_connect_shared(h, everyone, (C.c_int32 * self.world)(*peer_fds))
            if not self._all_agree(rc == 0):
                raise _SharedUnavailable
            return h
        except _SharedUnavailable:
            lib.amp_bucket_destroy(h)
            return None
        finally:
            for fd in fds:  # imports hold their own references
                os.close(fd)

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
