"""AMP state memories and the AMP state preprocessor.

Mirrors the part of upstream skrl (>= 1.4.3, third party, not vendored; configured by
``agents/skrl_g1_dance_amp_cfg.yaml:50-58, 76-98`` and driven from ``AMP._update``) that sits either side of the style
reward:

* ``AmpStateMemory``          <- ``skrl.memories.torch.RandomMemory`` holding one tensor ``"states"`` of width ``K*A``
  (the agent's ``motion_dataset`` and ``reply_buffer``): ``add_samples`` (ring write), ``sample`` /
  ``sample_by_index`` (row gather), ``__len__``.
* ``RunningStandardScaler``   <- ``skrl.resources.preprocessors.torch.RunningStandardScaler``: ``scaler(x, train=True)``
  merges the batch into the float64 running statistics, then returns the normalised, clipped batch.

Both run in ``libamp_b200.so`` (``csrc/amp_memory.cu``); there is no torch fallback.
"""

from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib


class AmpStateMemory:
    """Ring buffer of AMP observation rows ``(memory_size, width)`` fp32 on one device."""

    def __init__(self, memory_size: int, width: int, device):
        self.device = _lib.require_cuda(device)
        self.memory_size, self.width = int(memory_size), int(width)
        if self.memory_size < 1 or self.width < 1:
            raise ValueError("memory_size and width must be positive")
        self.states = torch.zeros((self.memory_size, self.width), dtype=torch.float32, device=self.device)
        self.memory_index = 0  # next row to write (skrl Memory.memory_index with num_envs = 1)
        self.filled = False
        self._flags = torch.zeros(1, dtype=torch.int32, device=self.device)

    def __len__(self) -> int:
        return self.memory_size if self.filled else self.memory_index

    def add_samples(self, states: torch.Tensor) -> None:
        """skrl ``Memory.add_samples(states=...)`` for a batch of rows: written at ``memory_index`` and wrapped around; a
        batch longer than the memory keeps its last ``memory_size`` rows, like successive overwrites would."""
        rows = states.to(self.device, torch.float32).reshape(-1, self.width)
        n = rows.shape[0]
        if n > self.memory_size:
            self.memory_index = (self.memory_index + n - self.memory_size) % self.memory_size
            rows, n = rows[-self.memory_size :], self.memory_size
        first = min(n, self.memory_size - self.memory_index)
        self.states[self.memory_index : self.memory_index + first].copy_(rows[:first])
        if n > first:
            self.states[: n - first].copy_(rows[first:])
        if self.memory_index + n >= self.memory_size:
            self.filled = True
        self.memory_index = (self.memory_index + n) % self.memory_size

    def write_cursor(self, n: int):
        """(start_row, capacity) for producers that write rows in place (``AmpEnvPath.collect_reference_motions_into``), and
        advance the cursor by ``n`` rows."""
        start = self.memory_index
        if self.memory_index + n >= self.memory_size:
            self.filled = True
        self.memory_index = (self.memory_index + n) % self.memory_size
        return start, self.memory_size

    def sample_indexes(self, batch_size: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """skrl ``RandomMemory.sample``: ``torch.randint(0, len(self), (batch_size,))`` (with replacement)."""
        if len(self) == 0:
            raise RuntimeError("cannot sample from an empty memory")
        return torch.randint(0, len(self), (int(batch_size),), device=self.device, generator=generator)

    def sample_by_index(self, indexes: torch.Tensor, mini_batches: int = 1, out: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """skrl ``Memory.sample_by_index``: the gathered rows, split into ``mini_batches`` chunks like ``np.array_split``."""
        idx = indexes.to(self.device, torch.int64).reshape(-1).contiguous()
        m = idx.numel()
        buf = out if out is not None else torch.empty((m, self.width), dtype=torch.float32, device=self.device)
        if buf.shape != (m, self.width) or buf.dtype != torch.float32 or buf.stride(1) != 1:
            raise ValueError("out must be a float32 (len(indexes), width) tensor with unit column stride")
        lib, stream = _lib.enter(self.device)
        _lib.check(lib.amp_gather_rows(_lib.ptr(self.states), self.states.stride(0), len(self), _lib.ptr(idx), m, self.width,
                                       _lib.ptr(buf), buf.stride(0), _lib.ptr(self._flags), stream))
        if mini_batches <= 1:
            return [buf]
        base, extra = divmod(m, mini_batches)
        sizes = [base + 1] * extra + [base] * (mini_batches - extra)
        return list(torch.split(buf, sizes))

    def sample(self, batch_size: int, mini_batches: int = 1, generator: Optional[torch.Generator] = None) -> List[torch.Tensor]:
        return self.sample_by_index(self.sample_indexes(batch_size, generator), mini_batches)

    def poll_flags(self) -> int:
        """Synchronising check of the sticky error word (bit 1: an index outside ``[0, len(self))`` was gathered)."""
        v = int(self._flags.item())
        if v:
            self._flags.zero_()
        return v


class RunningStandardScaler:
    """skrl ``RunningStandardScaler(size, epsilon=1e-8, clip_threshold=5.0)`` for ``(..., size)`` fp32 inputs."""

    def __init__(self, size: int, epsilon: float = 1e-8, clip_threshold: float = 5.0, device=None):
        self.device = _lib.require_cuda(device if device is not None else "cuda")
        self.size = int(size)
        self.epsilon, self.clip_threshold = float(epsilon), float(clip_threshold)
        self.running_mean = torch.zeros(self.size, dtype=torch.float64, device=self.device)
        self.running_variance = torch.ones(self.size, dtype=torch.float64, device=self.device)
        self.current_count = torch.ones((), dtype=torch.float64, device=self.device)
        nbytes = int(_lib.load().amp_scaler_scratch_bytes(self.size))
        self._scratch = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=self.device)

    def _rows(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.device, torch.float32)
        if x.shape[-1] != self.size:
            raise RuntimeError(f"expected last dimension {self.size}, got {x.shape[-1]}")
        x = x.reshape(-1, self.size)
        return x if x.stride(1) == 1 else x.contiguous()

    def update(self, x: torch.Tensor) -> None:
        """``_parallel_variance(mean(x, 0), var(x, 0), n)``: merge a batch into the running statistics, in place."""
        rows = self._rows(x)
        lib, stream = _lib.enter(self.device)
        _lib.check(lib.amp_scaler_update(_lib.ptr(rows), rows.stride(0), rows.shape[0], self.size, _lib.ptr(self.running_mean),
                                         _lib.ptr(self.running_variance), _lib.ptr(self.current_count), _lib.ptr(self._scratch),
                                         self._scratch.numel(), stream))

    def __call__(self, x: torch.Tensor, train: bool = False, inverse: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if inverse:
            raise NotImplementedError("inverse scaling is not on the AMP path (skrl only uses it for value preprocessors)")
        if train:
            self.update(x)
        rows = self._rows(x)
        res = out if out is not None else torch.empty_like(rows)
        lib, stream = _lib.enter(self.device)
        _lib.check(lib.amp_scaler_apply(_lib.ptr(rows), rows.stride(0), rows.shape[0], self.size, _lib.ptr(self.running_mean),
                                        _lib.ptr(self.running_variance), self.epsilon, self.clip_threshold, _lib.ptr(res),
                                        res.stride(0), stream))
        return res.view(x.shape)
