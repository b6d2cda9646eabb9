"""Oracle restatement of the skrl AMP discriminator style reward (TEST INFRASTRUCTURE -- see ``oracle/__init__.py``).

The algorithm lives in a third-party dependency that is NOT in ``/root/reference``: **skrl >= 1.4.3** (required at
``train.py:122-129``; no lock file, not vendored, not installed here).  PARITY UNPINNED: this is a literal restatement
of the published upstream code, anchored on the reference's own configuration of it:

* network: ``agents/skrl_g1_dance_amp_cfg.yaml:31-39`` -- ``Linear(K*A,1024)-ReLU-Linear(1024,512)-ReLU-Linear(512,1)``
* ``amp_state_preprocessor: RunningStandardScaler`` .... yaml ``:80``
* ``discriminator_reward_scale: 2.0`` .................. yaml ``:95``

Upstream expressions restated (skrl ``resources/preprocessors/torch/running_standard_scaler.py`` and
``agents/torch/amp/amp.py::_update``):

    x_hat  = clamp((x - mean.float()) / (sqrt(var.float()) + 1e-8), -5, 5)          # scaler, train=False
    logits = W3 @ relu(W2 @ relu(W1 @ x_hat + b1) + b2) + b3
    style  = -log(maximum(1 - 1 / (1 + exp(-logits)), 1e-4)) * discriminator_reward_scale
"""

from __future__ import annotations

import math

import torch


def running_standard_scaler_eval(x, running_mean, running_variance, epsilon=1e-8, clip_threshold=5.0):
    """skrl ``RunningStandardScaler._compute(train=False, inverse=False)``; statistics are float64 buffers."""
    return torch.clamp(
        (x - running_mean.float()) / (torch.sqrt(running_variance.float()) + epsilon),
        min=-clip_threshold,
        max=clip_threshold,
    )


def style_reward_from_logits(logits, reward_scale=2.0):
    """skrl ``AMP._update``: ``-log(max(1 - sigmoid(d), 1e-4)) * scale`` written exactly as upstream writes it."""
    prob_fake = 1 - 1 / (1 + torch.exp(-logits))
    return -torch.log(torch.maximum(prob_fake, torch.tensor(0.0001, device=logits.device))) * reward_scale


class OracleDiscriminator:
    """fp32 CPU discriminator with skrl-style running statistics."""

    def __init__(self, in_features, hidden=(1024, 512), reward_scale=2.0, seed=42, device="cpu", weights=None, biases=None):
        if weights is not None:
            self.weights = [w.detach().to(device, torch.float32).clone() for w in weights]
            self.biases = [b.detach().to(device, torch.float32).clone() for b in biases]
        else:
            g = torch.Generator(device="cpu").manual_seed(seed)
            dims = [in_features, *hidden, 1]
            self.weights, self.biases = [], []
            for fan_in, fan_out in zip(dims[:-1], dims[1:]):
                bound = 1.0 / math.sqrt(fan_in)  # torch.nn.Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for W and b
                self.weights.append(((torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound).to(device))
                self.biases.append(((torch.rand(fan_out, generator=g) * 2 - 1) * bound).to(device))
        self.in_features = in_features
        self.reward_scale = reward_scale
        self.running_mean = torch.zeros(in_features, dtype=torch.float64, device=device)
        self.running_variance = torch.ones(in_features, dtype=torch.float64, device=device)
        self.current_count = torch.ones((), dtype=torch.float64, device=device)

    def update_statistics(self, x):
        """skrl ``RunningStandardScaler._parallel_variance`` with batch mean / unbiased variance over dim 0."""
        mean = torch.mean(x, dim=0)
        var = torch.var(x, dim=0)
        count = x.shape[0]
        delta = mean - self.running_mean
        total = self.current_count + count
        m2 = (
            self.running_variance * self.current_count
            + var * count
            + delta**2 * self.current_count * count / total
        )
        self.running_mean = self.running_mean + delta * count / total
        self.running_variance = m2 / total
        self.current_count = total

    def normalise(self, x):
        return running_standard_scaler_eval(x, self.running_mean, self.running_variance)

    def logits(self, amp_states, emulate_bf16=False):
        """``emulate_bf16=True`` rounds the layer-1/2 operands to bf16 (fp32 accumulate), the arithmetic the tensor-core
        kernel performs; used only to separate quantisation error from kernel bugs in the tests."""
        h = self.normalise(amp_states)
        n_layers = len(self.weights)
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            if emulate_bf16 and i < n_layers - 1:
                h = h.to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t() + b
            else:
                h = h @ w.t() + b
            if i < n_layers - 1:
                h = torch.relu(h)
        return h

    def style_reward(self, amp_states, emulate_bf16=False):
        return style_reward_from_logits(self.logits(amp_states, emulate_bf16), self.reward_scale)
