"""skrl AMP discriminator loss + gradients on the tcgen05 tensor cores.

Replaces, inside skrl's ``AMP._update`` (upstream skrl >= 1.4.3; configured by the reference at
``agents/skrl_g1_dance_amp_cfg.yaml:89, 94, 96-98``), the "compute discriminator loss" block and the part of
``(policy_loss + entropy_loss + value_loss + discriminator_loss).backward()`` that flows into the discriminator:

    sampled_amp_states        = amp_state_preprocessor(sampled_amp_states[:discriminator_batch_size], train=True)
    sampled_amp_replay_states = amp_state_preprocessor(replay_batch[:discriminator_batch_size], train=True)
    sampled_amp_motion_states = amp_state_preprocessor(motion_batch[:discriminator_batch_size], train=True)
    ... BCE(cat(agent, replay) -> 0) / BCE(motion -> 1), logit regularisation, gradient penalty, weight decay ...
    discriminator_loss *= discriminator_loss_scale

The fp32 master parameters stay owned by the trainer (torch); the step reads them, and writes ``d loss / d parameter`` into
caller-provided fp32 tensors -- normally views into the flat bucket that :func:`humanoid_amp_b200.distributed.reduce_parameters`
all-reduces.  bf16 tensor-core operands with fp32 accumulation; there is no CPU path.
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .memory import RunningStandardScaler

TERM_NAMES = ("bce_agent_replay", "bce_motion", "logit_regularization", "gradient_penalty", "weight_decay", "loss")


class AmpDiscriminatorUpdate:
    """Field names of the four scales follow skrl's ``AMP_DEFAULT_CONFIG``; defaults are the reference's yaml values."""

    def __init__(self, in_features: int, hidden: Sequence[int] = (1024, 512), max_batch_rows: int = 4096, device="cuda",
                 discriminator_loss_scale: float = 5.0, discriminator_logit_regularization_scale: float = 0.05,
                 discriminator_gradient_penalty_scale: float = 5.0, discriminator_weight_decay_scale: float = 1.0e-4):
        if len(hidden) != 2:
            raise ValueError("the reference discriminator has exactly two hidden layers (1024, 512)")
        self.device = _lib.require_cuda(device)
        self.in_features, self.hidden = int(in_features), tuple(int(h) for h in hidden)
        self.max_batch_rows = int(max_batch_rows)
        self.discriminator_loss_scale = float(discriminator_loss_scale)
        self.discriminator_logit_regularization_scale = float(discriminator_logit_regularization_scale)
        self.discriminator_gradient_penalty_scale = float(discriminator_gradient_penalty_scale)
        self.discriminator_weight_decay_scale = float(discriminator_weight_decay_scale)
        lib, stream = _lib.enter(self.device)
        h = C.c_void_p()
        _lib.check(lib.amp_disc_train_create(self.in_features, self.hidden[0], self.hidden[1], self.max_batch_rows, stream, C.byref(h)))
        self._h = h
        self._batch_rows: Optional[int] = None
        self._keep = []  # staged sources stay alive until the step has been enqueued

    # ---- staging ---------------------------------------------------------------------------------------------------
    def stage(self, source: int, states: torch.Tensor, scaler: Optional[RunningStandardScaler] = None, train: bool = True) -> None:
        """Stage batch ``source`` (0 agent, 1 replay, 2 motion).  With a ``scaler`` the rows go through
        ``scaler(states, train=train)`` (statistics update first when ``train``, as skrl does, then normalise + clip);
        without one they are taken as already normalised."""
        x = states.to(self.device, torch.float32).reshape(-1, self.in_features)
        if x.stride(1) != 1:
            x = x.contiguous()
        rows = x.shape[0]
        if self._batch_rows is None:
            self._batch_rows = rows
        elif rows != self._batch_rows:
            raise RuntimeError(f"the three batches of a step must have the same length (got {rows}, expected {self._batch_rows})")
        mean = var = None
        if scaler is not None:
            if scaler.size != self.in_features:
                raise RuntimeError("scaler size does not match in_features")
            if scaler.epsilon != 1e-8 or scaler.clip_threshold != 5.0:
                # the staging kernel evaluates skrl's DEFAULT preprocessor (epsilon 1e-8, clip +-5, the reference's yaml sets
                # neither): refuse anything else instead of normalising the batch differently from scaler() itself
                raise RuntimeError(
                    f"the staging kernel implements RunningStandardScaler(epsilon=1e-8, clip_threshold=5.0); got epsilon={scaler.epsilon}, "
                    f"clip_threshold={scaler.clip_threshold} -- normalise with scaler(states) first and stage without a scaler"
                )
            if train:
                scaler.update(x)
            mean, var = scaler.running_mean, scaler.running_variance
        lib, stream = _lib.enter(self.device)
        _lib.check(lib.amp_disc_train_stage(self._h, int(source), _lib.ptr(x), x.stride(0), rows, _lib.ptr(mean), _lib.ptr(var), stream))
        self._keep.append(x)

    # ---- the step --------------------------------------------------------------------------------------------------
    def loss_and_grads(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                       grad_weights: Optional[Sequence[torch.Tensor]] = None, grad_biases: Optional[Sequence[torch.Tensor]] = None,
                       return_logits: bool = False, bucket=None):
        """Returns ``(terms, grad_weights, grad_biases[, logits])``: ``terms`` is a float32 device tensor of 6 values named by
        ``TERM_NAMES`` (``terms[5]`` = the scaled discriminator loss); the gradient lists mirror ``weights`` / ``biases``
        (``[W1 (h1,in), W2 (h2,h1), W3 (1,h2)]``, ``[b1, b2, b3]``) and are overwritten in place when given.

        With ``bucket`` (a :class:`humanoid_amp_b200.distributed.GradientBucket`; the gradient tensors must be six consecutive
        views of it, e.g. ``bucket.carve``) the step's last kernel also runs the gradient exchange: the tensors come back holding
        the MEAN over the ranks, exactly what ``bucket.all_reduce_mean`` over that range would have made of them, without the
        second launch.  Collective: every rank calls it at the same point."""
        if self._batch_rows is None:
            raise RuntimeError("stage() the three batches before loss_and_grads()")
        dev = self.device
        W = [w.detach().to(dev, torch.float32).contiguous() for w in weights]
        b = [x.detach().to(dev, torch.float32).contiguous() for x in biases]
        h1, h2 = self.hidden
        for w, shape in zip(W, [(h1, self.in_features), (h2, h1), (1, h2)]):
            if tuple(w.shape) != shape:
                raise RuntimeError(f"weight shape {tuple(w.shape)} != {shape}")
        gW = list(grad_weights) if grad_weights is not None else [torch.empty_like(w) for w in W]
        gb = list(grad_biases) if grad_biases is not None else [torch.empty_like(x) for x in b]
        for g, ref in zip(gW + gb, W + b):
            if g.shape != ref.shape or g.dtype != torch.float32 or g.device != dev or not g.is_contiguous():
                raise RuntimeError("gradient tensors must be contiguous float32 tensors of the parameter shapes on the same device")
        B = self._batch_rows
        Bp = (B + 127) // 128 * 128
        terms = torch.empty(6, dtype=torch.float32, device=dev)
        logits = torch.empty(3 * Bp, dtype=torch.float32, device=dev) if return_logits else None
        lib, stream = _lib.enter(dev)
        self._batch_rows = None  # whatever happens next, the following step starts from fresh staging
        args = (self._h, _lib.ptr(W[0]), _lib.ptr(b[0]), _lib.ptr(W[1]), _lib.ptr(b[1]), _lib.ptr(W[2]), _lib.ptr(b[2]), B,
                self.discriminator_loss_scale, self.discriminator_logit_regularization_scale,
                self.discriminator_gradient_penalty_scale, self.discriminator_weight_decay_scale,
                _lib.ptr(gW[0]), _lib.ptr(gb[0]), _lib.ptr(gW[1]), _lib.ptr(gb[1]), _lib.ptr(gW[2]), _lib.ptr(gb[2]),
                _lib.ptr(terms), _lib.ptr(logits))
        if bucket is not None:
            if bucket.device != dev:
                raise RuntimeError("the bucket lives on another device")
            _lib.check(lib.amp_disc_train_step_exchange(*args, bucket._h, stream))
        else:
            _lib.check(lib.amp_disc_train_step(*args, stream))
        self._keep = [W, b]
        if return_logits:
            return terms, gW, gb, logits.view(3, Bp)[:, :B]
        return terms, gW, gb

    def __call__(self, weights, biases, amp_states, replay_states, motion_states, scaler: Optional[RunningStandardScaler] = None,
                 train: bool = True, grad_weights=None, grad_biases=None, return_logits: bool = False, bucket=None):
        """One discriminator update's loss + gradients from the three raw batches, in skrl's order."""
        self.stage(0, amp_states, scaler, train)
        self.stage(1, replay_states, scaler, train)
        self.stage(2, motion_states, scaler, train)
        return self.loss_and_grads(weights, biases, grad_weights, grad_biases, return_logits, bucket)

    def close(self):
        if getattr(self, "_h", None) is not None:
            try:
                _lib.load().amp_disc_train_destroy(self._h)
            finally:
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals may already be gone; the driver reclaims the handle
            pass
