"""Host <-> device staging for the AMP path: double-buffered input prefetch and deferred result read-back.

The reference hands ``collect_reference_motions`` host numpy ``times`` / ``motion_ids`` (``g1_amp_env.py:445-462``,
``motions/motion_loader.py:309-329``) and reads rewards back on the host once per rollout.  Done serially that costs a
PCIe round trip per step (16 B per sample in, 4 B per sample out) during which the SMs idle.  These two helpers keep the
copies of neighbouring steps underneath the kernels of the current one, on their own streams (B200 has separate H2D / D2H
copy engines); nothing here computes -- the kernels are still the ones of ``libamp_b200.so``.

    pre = InputPrefetcher(device, n)                  # depth 2
    out = ResultReader(device)
    slot = pre.submit(times_host, ids_host)           # H2D of step 0
    for i in range(steps):
        nxt = pre.submit(times_host, ids_host)        # H2D of step i+1 runs under step i
        t, ids = pre.acquire(slot)                    # current stream waits for step i's inputs
        obs = env.collect_reference_motions(n, t, ids, out=obs_buf)
        pre.release(slot)
        r = disc.style_reward(obs, out=reward_buf[i & 1])
        ticket = out.read_async(r)                    # D2H of step i on the read-back stream
        if prev is not None: host_rewards = prev.wait()   # host consumes step i-1 while step i runs
        prev, slot = ticket, nxt
"""

from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib


def _as_pinned(x, dtype: torch.dtype) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    if t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_pinned() else t.pin_memory()


class InputPrefetcher:
    """``depth`` device slots of (times float64[n], motion_ids int64[n]) filled by asynchronous copies on a private stream."""

    def __init__(self, device, num_samples: int, depth: int = 2):
        self.device = _lib.require_cuda(device)
        self.n = int(num_samples)
        self.depth = int(depth)
        if self.depth < 1:
            raise ValueError("depth must be >= 1")
        self._stream = torch.cuda.Stream(device=self.device)
        self._times = [torch.empty(self.n, dtype=torch.float64, device=self.device) for _ in range(self.depth)]
        self._ids = [torch.empty(self.n, dtype=torch.int64, device=self.device) for _ in range(self.depth)]
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]
        self._free: List[Optional[torch.cuda.Event]] = [None] * self.depth
        self._next = 0

    def submit(self, times_host, ids_host) -> int:
        """Start copying one step's inputs (pinned tensors are used as they are, anything else is staged through a pinned
        copy first).  Returns the slot to ``acquire`` later; at most ``depth`` submissions may be outstanding."""
        slot = self._next
        self._next = (slot + 1) % self.depth
        t = _as_pinned(times_host, torch.float64)
        i = _as_pinned(ids_host, torch.int64)
        if t.numel() != self.n or i.numel() != self.n:
            raise ValueError(f"expected {self.n} times / ids, got {t.numel()} / {i.numel()}")
        with torch.cuda.stream(self._stream):
            if self._free[slot] is not None:
                self._stream.wait_event(self._free[slot])  # the kernels that read this slot last have finished
            self._times[slot].copy_(t.view(-1), non_blocking=True)
            self._ids[slot].copy_(i.view(-1), non_blocking=True)
            self._ready[slot].record(self._stream)
        return slot

    def acquire(self, slot: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Make the current stream wait for the slot's copies; returns the device (times, ids) tensors."""
        torch.cuda.current_stream(self.device).wait_event(self._ready[slot])
        return self._times[slot], self._ids[slot]

    def release(self, slot: int) -> None:
        """Call after the kernels that read the slot have been enqueued on the current stream."""
        ev = self._free[slot] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[slot] = ev


class _Ticket:
    def __init__(self, host: torch.Tensor, done: torch.cuda.Event):
        self._host, self._done = host, done

    def wait(self) -> torch.Tensor:
        """Block until the copy has landed; returns the pinned host tensor (valid until its buffer is reused, i.e. for
        ``depth`` further ``read_async`` calls)."""
        self._done.synchronize()
        return self._host


class ResultReader:
    """Deferred device -> pinned-host reads on a private stream, ``depth`` rotating host buffers."""

    def __init__(self, device, depth: int = 2):
        self.device = _lib.require_cuda(device)
        self.depth = int(depth)
        self._stream = torch.cuda.Stream(device=self.device)
        self._host: List[Optional[torch.Tensor]] = [None] * self.depth
        self._next = 0

    def read_async(self, result: torch.Tensor) -> _Ticket:
        slot = self._next
        self._next = (slot + 1) % self.depth
        buf = self._host[slot]
        if buf is None or buf.shape != result.shape or buf.dtype != result.dtype:
            buf = self._host[slot] = torch.empty(result.shape, dtype=result.dtype).pin_memory()
        produced = torch.cuda.Event()
        produced.record(torch.cuda.current_stream(self.device))
        done = torch.cuda.Event()
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(produced)
            buf.copy_(result, non_blocking=True)
            done.record(self._stream)
        result.record_stream(self._stream)
        return _Ticket(buf, done)
