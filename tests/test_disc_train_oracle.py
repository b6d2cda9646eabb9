"""CPU checks of the discriminator-loss oracle (upstream skrl, parity unpinned): the closed-form
gradients the CUDA path implements must equal torch autograd on the literal skrl expression."""

from __future__ import annotations

import pytest
import torch

from oracle.disc_train_oracle import DiscLossCfg, discriminator_loss_autograd, discriminator_loss_manual


def _problem(in_features, h1, h2, B, seed):
    g = torch.Generator().manual_seed(seed)
    W = [torch.randn(h1, in_features, generator=g) * 0.2, torch.randn(h2, h1, generator=g) * 0.2, torch.randn(1, h2, generator=g) * 0.2]
    b = [torch.randn(h1, generator=g) * 0.1, torch.randn(h2, generator=g) * 0.1, torch.randn(1, generator=g) * 0.1]
    batches = [torch.randn(B, in_features, generator=g).clamp(-5, 5) for _ in range(3)]
    return W, b, batches


@pytest.mark.parametrize("shape", [(70, 64, 48, 33), (166, 128, 64, 100), (9, 16, 8, 1)])
def test_closed_form_gradients_equal_autograd_in_float64(shape):
    W, b, (agent, replay, motion) = _problem(*shape, seed=sum(shape))
    la, ta, gWa, gba = discriminator_loss_autograd(W, b, agent, replay, motion, dtype=torch.float64)
    lm, tm, gWm, gbm = discriminator_loss_manual(W, b, agent, replay, motion, dtype=torch.float64)
    assert abs(float(la) - float(lm)) <= 1e-12 * max(1.0, abs(float(la)))
    for k in ta:
        assert abs(float(ta[k]) - float(tm[k])) <= 1e-12 * max(1.0, abs(float(ta[k]))), k
    for a, m in zip(gWa + gba, gWm + gbm):
        assert a.shape == m.shape
        assert float((a - m).abs().max()) <= 1e-12 * max(1.0, float(a.abs().max()))


def test_each_scale_switches_its_term():
    """Zero scales drop their terms from the loss and the gradients (skrl guards each block with ``if scale:``)."""
    W, b, (agent, replay, motion) = _problem(20, 16, 8, 12, seed=5)
    base = DiscLossCfg(5.0, 0.0, 0.0, 0.0)
    l0, _, gW0, _ = discriminator_loss_autograd(W, b, agent, replay, motion, base, dtype=torch.float64)
    for field in ("discriminator_logit_regularization_scale", "discriminator_gradient_penalty_scale", "discriminator_weight_decay_scale"):
        cfg = DiscLossCfg(5.0, 0.0, 0.0, 0.0)
        setattr(cfg, field, 0.5)
        la, ta, gWa, gba = discriminator_loss_autograd(W, b, agent, replay, motion, cfg, dtype=torch.float64)
        lm, tm, gWm, gbm = discriminator_loss_manual(W, b, agent, replay, motion, cfg, dtype=torch.float64)
        assert float(la) > float(l0)
        assert abs(float(la) - float(lm)) <= 1e-12 * abs(float(la))
        for a, m in zip(gWa + gba, gWm + gbm):
            assert float((a - m).abs().max()) <= 1e-12 * max(1.0, float(a.abs().max()))


def test_bf16_emulation_stays_close_to_fp32():
    W, b, (agent, replay, motion) = _problem(166, 256, 128, 256, seed=9)
    _, _, gWa, gba = discriminator_loss_autograd(W, b, agent, replay, motion, dtype=torch.float32)
    _, _, gWe, gbe = discriminator_loss_manual(W, b, agent, replay, motion, dtype=torch.float64, emulate_bf16=True)
    for a, e in zip(gWa + gba, gWe + gbe):
        assert float((a.double() - e).abs().max()) <= 3e-2 * float(a.abs().max())
