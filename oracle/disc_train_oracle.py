"""Oracle restatement of the skrl AMP discriminator LOSS and its gradients (TEST INFRASTRUCTURE -- see
``oracle/__init__.py``).  SURVEY.md section 8f item 2.

The algorithm lives in a third-party dependency that is NOT in ``/root/reference``: **skrl >= 1.4.3** (required at
``train.py:122-129``; no lock file, not vendored, not installed here).  PARITY UNPINNED: literal restatement of the
published upstream code (``skrl/agents/torch/amp/amp.py::AMP._update``, "compute discriminator loss" block), anchored on
the reference's configuration of it (``agents/skrl_g1_dance_amp_cfg.yaml``):

* ``discriminator_batch_size: 4096`` ..................... yaml ``:94``  (first 4096 rows of each of the three batches)
* ``discriminator_loss_scale: 5.0`` ...................... yaml ``:89``
* ``discriminator_logit_regularization_scale: 0.05`` ..... yaml ``:96``
* ``discriminator_gradient_penalty_scale: 5.0`` .......... yaml ``:97``
* ``discriminator_weight_decay_scale: 1.0e-04`` .......... yaml ``:98``

Upstream expressions restated (the three batches have already been through ``amp_state_preprocessor(x, train=True)``):

    motion.requires_grad_(True)
    logits        = D(agent); replay_logits = D(replay); motion_logits = D(motion)
    cat_logits    = cat([logits, replay_logits], 0)
    loss  = 0.5 * (BCEWithLogits(cat_logits, zeros) + BCEWithLogits(motion_logits, ones))
    loss += logit_reg * sum(square(flatten(last_linear.weight)))
    grad  = autograd.grad(motion_logits, motion, grad_outputs=ones, create_graph=True, retain_graph=True)[0]
    loss += grad_penalty * sum(square(grad), -1).mean()
    loss += weight_decay * sum(square(cat([flatten(m.weight) for m in linears])))
    loss *= loss_scale

Two implementations:

* :func:`discriminator_loss_autograd` -- the expression above evaluated by torch autograd (the reference semantics);
* :func:`discriminator_loss_manual`   -- the same gradients in closed form (the formulas the CUDA path implements), with
  an optional emulation of the bf16 roundings of the tensor-core operands.  ``tests/test_oracle_pins.py`` checks it
  against the autograd version in float64, so a mismatch of the CUDA path can be attributed to rounding or to a bug.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch


@dataclass
class DiscLossCfg:
    """Field names follow skrl's AMP_DEFAULT_CONFIG; defaults are the reference's yaml values."""

    discriminator_loss_scale: float = 5.0
    discriminator_logit_regularization_scale: float = 0.05
    discriminator_gradient_penalty_scale: float = 5.0
    discriminator_weight_decay_scale: float = 1.0e-4


def _mlp(x, W, b):
    h = x
    for i in range(len(W)):
        h = h @ W[i].t() + b[i]
        if i < len(W) - 1:
            h = torch.relu(h)
    return h


def discriminator_loss_autograd(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], agent: torch.Tensor,
                                replay: torch.Tensor, motion: torch.Tensor, cfg: DiscLossCfg = DiscLossCfg(),
                                dtype=torch.float32):
    """Returns ``(loss, terms, grads_W, grads_b)``; ``terms`` = dict of the unscaled loss components."""
    W = [w.detach().to(dtype).clone().requires_grad_(True) for w in weights]
    b = [x.detach().to(dtype).clone().requires_grad_(True) for x in biases]
    agent, replay = agent.detach().to(dtype), replay.detach().to(dtype)
    motion = motion.detach().to(dtype).clone().requires_grad_(True)
    bce = torch.nn.BCEWithLogitsLoss()

    amp_logits = _mlp(agent, W, b)
    amp_replay_logits = _mlp(replay, W, b)
    amp_motion_logits = _mlp(motion, W, b)
    amp_cat_logits = torch.cat([amp_logits, amp_replay_logits], dim=0)

    bce_cat = bce(amp_cat_logits, torch.zeros_like(amp_cat_logits))
    bce_motion = bce(amp_motion_logits, torch.ones_like(amp_motion_logits))
    loss = 0.5 * (bce_cat + bce_motion)
    terms = {"bce_agent_replay": bce_cat.detach(), "bce_motion": bce_motion.detach()}

    logit_reg = torch.sum(torch.square(torch.flatten(W[-1])))
    terms["logit_regularization"] = logit_reg.detach()
    if cfg.discriminator_logit_regularization_scale:
        loss = loss + cfg.discriminator_logit_regularization_scale * logit_reg

    grad = torch.autograd.grad(amp_motion_logits, motion, grad_outputs=torch.ones_like(amp_motion_logits), create_graph=True,
                               retain_graph=True, only_inputs=True)[0]
    gradient_penalty = torch.sum(torch.square(grad), dim=-1).mean()
    terms["gradient_penalty"] = gradient_penalty.detach()
    if cfg.discriminator_gradient_penalty_scale:
        loss = loss + cfg.discriminator_gradient_penalty_scale * gradient_penalty

    weight_decay = torch.sum(torch.square(torch.cat([torch.flatten(w) for w in W], dim=-1)))
    terms["weight_decay"] = weight_decay.detach()
    if cfg.discriminator_weight_decay_scale:
        loss = loss + cfg.discriminator_weight_decay_scale * weight_decay

    loss = loss * cfg.discriminator_loss_scale
    loss.backward()
    return loss.detach(), terms, [w.grad for w in W], [x.grad for x in b]


def _r(t: torch.Tensor, on: bool) -> torch.Tensor:
    """bf16 rounding of a tensor-core operand (round to nearest even), kept in the working dtype."""
    return t.to(torch.bfloat16).to(t.dtype) if on else t


def discriminator_loss_manual(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], agent: torch.Tensor,
                              replay: torch.Tensor, motion: torch.Tensor, cfg: DiscLossCfg = DiscLossCfg(),
                              dtype=torch.float64, emulate_bf16: bool = False
                              ) -> Tuple[torch.Tensor, dict, List[torch.Tensor], List[torch.Tensor]]:
    """Closed-form loss + gradients of a 2-hidden-layer ReLU discriminator (the shape the reference configures).

    With ``z1 = x W1^T + b1, a1 = relu(z1), z2 = a1 W2^T + b2, a2 = relu(z2), d = a2 w3 + b3`` and the ReLU masks
    ``m1 = [z1 > 0], m2 = [z2 > 0]`` (piecewise constant: no gradient flows through them, as in autograd):

        dL/dd    = s/2 * sigmoid(d) / (2B)        agent + replay rows (target 0)
                 = s/2 * (sigmoid(d) - 1) / B     motion rows (target 1)
        gradient of d w.r.t. the input, motion rows:   u2 = w3 * m2,  v1 = (u2 W2) * m1,  g = v1 W1
        GP = mean_r |g_r|^2;   G = dL/dg = s * c_gp * 2 g / B
        dL/dW1 += v1^T G;   q1 = (G W1^T) * m1;   dL/dW2 += u2^T q1;   dL/dw3 += sum_r (q1 W2^T) * m2

    ``emulate_bf16`` rounds every tensor-core operand to bf16 where the CUDA path does (x, W1, W2, a1, a2, dz2, dz1, u2,
    v1, G, q1); accumulation stays in ``dtype``.
    """
    e = emulate_bf16
    W1, W2, W3 = [w.detach().to(dtype) for w in weights]
    b1, b2, b3 = [x.detach().to(dtype) for x in biases]
    w3 = W3.reshape(-1)
    s = cfg.discriminator_loss_scale
    B = motion.shape[0]
    n_cat = agent.shape[0] + replay.shape[0]
    x = _r(torch.cat([agent, replay, motion], 0).detach().to(dtype), e)
    W1r, W2r = _r(W1, e), _r(W2, e)

    a1 = _r(torch.relu(x @ W1r.t() + b1), e)
    a2_full = torch.relu(a1 @ W2r.t() + b2)
    a2 = _r(a2_full, e)
    m1, m2 = (a1 > 0).to(dtype), (a2 > 0).to(dtype)
    d = a2_full @ w3 + b3  # the CUDA path folds the last layer into the layer-2 read-out, before a2 is rounded
    sig = torch.sigmoid(d)
    softplus = torch.nn.functional.softplus
    bce_cat = softplus(d[:n_cat]).mean()
    bce_motion = softplus(-d[n_cat:]).mean()

    dd = torch.empty_like(d)
    dd[:n_cat] = 0.5 * s * sig[:n_cat] / n_cat
    dd[n_cat:] = 0.5 * s * (sig[n_cat:] - 1.0) / B

    gW3 = (dd[:, None] * a2).sum(0)
    gb3 = dd.sum().reshape(1)
    dz2_full = dd[:, None] * w3[None, :] * m2
    gb2 = dz2_full.sum(0)
    dz2 = _r(dz2_full, e)
    gW2 = dz2.t() @ a1
    dz1_full = (dz2 @ W2r) * m1
    gb1 = dz1_full.sum(0) if not e else _r(dz1_full, e).sum(0)
    dz1 = _r(dz1_full, e)
    gW1 = dz1.t() @ x

    # gradient penalty on the motion rows
    m1m, m2m = m1[n_cat:], m2[n_cat:]
    u2 = _r(w3[None, :] * m2m, e)
    v1 = _r((u2 @ W2r) * m1m, e)
    g = v1 @ W1r
    gp = (g * g).sum(-1).mean()
    c = s * cfg.discriminator_gradient_penalty_scale
    G = _r(c * 2.0 * g / B, e)
    gW1 = gW1 + v1.t() @ G
    q1 = _r((G @ W1r.t()) * m1m, e)
    gW2 = gW2 + u2.t() @ q1
    sfull = (q1 @ W2r.t()) * m2m
    gW3 = gW3 + (_r(sfull, e)).sum(0)

    logit_reg = (w3 * w3).sum()
    wd = (W1 * W1).sum() + (W2 * W2).sum() + (w3 * w3).sum()
    cw = 2.0 * s * cfg.discriminator_weight_decay_scale
    gW1 = gW1 + cw * W1
    gW2 = gW2 + cw * W2
    gW3 = gW3 + (cw + 2.0 * s * cfg.discriminator_logit_regularization_scale) * w3

    loss = s * (0.5 * (bce_cat + bce_motion) + cfg.discriminator_logit_regularization_scale * logit_reg
                + cfg.discriminator_gradient_penalty_scale * gp + cfg.discriminator_weight_decay_scale * wd)
    terms = {"bce_agent_replay": bce_cat, "bce_motion": bce_motion, "logit_regularization": logit_reg,
             "gradient_penalty": gp, "weight_decay": wd}
    return loss, terms, [gW1, gW2, gW3.reshape(1, -1)], [gb1, gb2, gb3]
