"""skrl AMP discriminator forward + style reward on the tcgen05 tensor cores.

Replaces, inside skrl's ``AMP._update`` (upstream skrl >= 1.4.3, configured by the reference at
``agents/skrl_g1_dance_amp_cfg.yaml:31-39, 80, 94-95``):

    amp_logits = discriminator.act({"states": amp_state_preprocessor(amp_states)})[0]
    style_reward = -log(maximum(1 - 1/(1 + exp(-amp_logits)), 1e-4)) * discriminator_reward_scale

The fp32 master weights and the scaler's float64 statistics stay owned by torch (the trainer updates them); this class
keeps a bf16 shadow for the tensor cores, refreshed with :meth:`load`.  There is no CPU path.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib


class AmpDiscriminator:
    def __init__(self, in_features: int, hidden: Sequence[int] = (1024, 512), reward_scale: float = 2.0, device="cuda", max_rows: int = 65536):
        if len(hidden) != 2:
            raise ValueError("the reference discriminator has exactly two hidden layers (1024, 512)")
        self.device = _lib.require_cuda(device)
        self.in_features, self.hidden, self.reward_scale = int(in_features), tuple(int(h) for h in hidden), float(reward_scale)
        lib, stream = _lib.enter(self.device)
        h = C.c_void_p()
        _lib.check(lib.amp_disc_create(self.in_features, self.hidden[0], self.hidden[1], int(max_rows), stream, C.byref(h)))
        self._h = h
        self._masters = None
        self.chunk_rows = int(lib.amp_disc_chunk_rows(h))  # rows one persistent wave covers (148 CTAs x 128 rows)

    def launch_count(self, rows: int) -> int:
        """Kernel launches one ``style_reward`` call over ``rows`` rows issues: ONE (scaler + cast + both layers + reward)."""
        return int(_lib.load().amp_disc_launch_count(self._h, int(rows)))

    def load(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], running_mean: torch.Tensor, running_variance: torch.Tensor) -> None:
        """Refresh the staged copies from ``[W1 (h1,in), W2 (h2,h1), W3 (1,h2)]``, ``[b1, b2, b3]`` (torch.nn.Linear layout)
        and the ``RunningStandardScaler`` buffers (float64).  Call after every optimiser step that changed them."""
        dev = self.device
        W = [w.detach().to(dev, torch.float32).contiguous() for w in weights]
        b = [x.detach().to(dev, torch.float32).contiguous() for x in biases]
        mean = running_mean.detach().to(dev, torch.float64).contiguous()
        var = running_variance.detach().to(dev, torch.float64).contiguous()
        h1, h2 = self.hidden
        expect = [(h1, self.in_features), (h2, h1), (1, h2)]
        for w, shape in zip(W, expect):
            if tuple(w.shape) != shape:
                raise RuntimeError(f"weight shape {tuple(w.shape)} != {shape}")
        if mean.numel() != self.in_features or var.numel() != self.in_features:
            raise RuntimeError("scaler statistics must have in_features entries")
        lib, stream = _lib.enter(dev)
        _lib.check(
            lib.amp_disc_load(self._h, _lib.ptr(W[0]), _lib.ptr(b[0]), _lib.ptr(W[1]), _lib.ptr(b[1]), _lib.ptr(W[2]), _lib.ptr(b[2]), _lib.ptr(mean), _lib.ptr(var), stream)
        )
        self._masters = (W, b, mean, var)  # keep the sources alive until the async copies have run

    def style_reward(self, amp_states: torch.Tensor, return_logits: bool = False, out: Optional[torch.Tensor] = None):
        """``amp_states (..., K*A)`` fp32 -> style reward ``(..., 1)`` (and the logits if asked)."""
        x = amp_states.to(self.device, torch.float32)
        lead = x.shape[:-1]
        if x.shape[-1] != self.in_features:
            raise RuntimeError(f"expected last dimension {self.in_features}, got {x.shape[-1]}")
        x = x.reshape(-1, self.in_features)
        if x.stride(-1) != 1:
            x = x.contiguous()
        M = x.shape[0]
        reward = out if out is not None else torch.empty(M, dtype=torch.float32, device=self.device)
        logits = torch.empty(M, dtype=torch.float32, device=self.device) if return_logits else None
        lib, stream = _lib.enter(self.device)
        _lib.check(lib.amp_disc_style_reward(self._h, _lib.ptr(x), x.stride(0), M, self.reward_scale, _lib.ptr(reward), _lib.ptr(logits), stream))
        reward = reward.view(*lead, 1)
        return (reward, logits.view(*lead, 1)) if return_logits else reward

    def style_reward_sampled(self, memory_states: torch.Tensor, indexes: torch.Tensor, return_logits: bool = False,
                             out: Optional[torch.Tensor] = None, flags: Optional[torch.Tensor] = None):
        """Style reward of ``memory_states[indexes]`` without materialising the gathered batch: skrl
        ``Memory.sample_by_index`` fused into the preprocessor + discriminator.  ``memory_states`` is the ``(capacity, K*A)``
        fp32 tensor of a memory (``AmpStateMemory.states``), ``indexes`` int64 ``(M,)``; ``flags`` an optional int32
        device word whose bit 1 is raised by an out-of-range index."""
        mem = memory_states
        if mem.device != self.device or mem.dtype != torch.float32 or mem.dim() != 2 or mem.stride(1) != 1:
            raise RuntimeError("memory_states must be a float32 (capacity, K*A) tensor on the discriminator's device")
        if mem.shape[1] != self.in_features:
            raise RuntimeError(f"expected memory rows of width {self.in_features}, got {mem.shape[1]}")
        idx = indexes.to(self.device, torch.int64).reshape(-1).contiguous()
        M = idx.numel()
        reward = out if out is not None else torch.empty(M, dtype=torch.float32, device=self.device)
        logits = torch.empty(M, dtype=torch.float32, device=self.device) if return_logits else None
        lib, stream = _lib.enter(self.device)
        _lib.check(lib.amp_disc_style_reward_indexed(self._h, _lib.ptr(mem), mem.stride(0), mem.shape[0], _lib.ptr(idx), M,
                                                     self.reward_scale, _lib.ptr(reward), _lib.ptr(logits), _lib.ptr(flags), stream))
        reward = reward.view(M, 1)
        return (reward, logits.view(M, 1)) if return_logits else reward

    def close(self):
        if getattr(self, "_h", None) is not None:
            try:
                _lib.load().amp_disc_destroy(self._h)
            finally:
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals may already be gone; the driver reclaims the handle
            pass


def style_reward_from_logits(logits: torch.Tensor, reward_scale: float = 2.0) -> torch.Tensor:
    """Only the reward expression, for callers that keep their own discriminator forward."""
    dev = _lib.require_cuda(logits.device)
    flat = logits.to(dev, torch.float32).contiguous().view(-1)
    out = torch.empty_like(flat)
    lib, stream = _lib.enter(dev)
    _lib.check(lib.amp_style_reward_from_logits(_lib.ptr(flat), flat.numel(), float(reward_scale), _lib.ptr(out), stream))
    return out.view(logits.shape)
