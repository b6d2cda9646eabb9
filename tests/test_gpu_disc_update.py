"""Discriminator loss + gradients on tcgen05 (``-m gpu``) against the oracle.

Stated bf16 tolerance.  Every tensor-core operand (states, weights, activations, back-propagated errors) is rounded to bf16
and accumulated in fp32, so per gradient tensor ``g`` (``ref`` = the oracle's gradient):

    vs the fp32 autograd oracle (the reference semantics):   max|g - ref| <= 3e-2 * max(1, sqrt(1024 / B)) * max|ref|
                                                             and   cos(g, ref) >= 0.999      (B = rows per batch: the
                                                             rounding noise of a gradient entry averages over the batch)
    vs the closed-form oracle with the SAME bf16 roundings:  max|g - ref| <= 4e-3 * max|ref|
    loss terms: |t - ref| <= 1e-2 * max(1, |ref|)  (fp32 oracle),  <= 2e-3 * max(1, |ref|)  (same roundings)
    logits:     |d - ref| <= 1e-2 * max(1, max|ref|)

(The second line separates rounding from bugs: only accumulation order and the rare mask flip of a value that rounds to
zero differ.  The fp32 comparison is made for batches of at least 64 rows.)
"""

from __future__ import annotations

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def make_problem(in_features, hidden, B, seed, gain=1.0):
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params

    W, b = skrl_style_discriminator_params(in_features, hidden=hidden, seed=seed, logit_gain=gain)
    g = torch.Generator().manual_seed(seed + 7)
    batches = [torch.randn(B, in_features, generator=g).clamp_(-5, 5) for _ in range(3)]
    batches[2] = batches[2] * 0.7 + 0.3  # the motion batch comes from a different distribution
    return W, b, batches


def compare(got, ref, rel, what, cos_min=None):
    got, ref = got.detach().double().cpu().reshape(-1), ref.detach().double().cpu().reshape(-1)
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    assert err <= rel * max(scale, 1e-30), f"{what}: max err {err:.3e} vs scale {scale:.3e} (allowed {rel:.0e} relative)"
    if cos_min is not None and got.numel() > 1:
        cos = float(torch.dot(got, ref) / (got.norm() * ref.norm()).clamp_min(1e-300))
        assert cos >= cos_min, f"{what}: cosine {cos:.6f} < {cos_min}"


def run_and_check(in_features, hidden, B, seed=3, gain=1.0):
    import humanoid_amp_b200 as amp
    from oracle.disc_train_oracle import DiscLossCfg, discriminator_loss_autograd, discriminator_loss_manual

    W, b, (agent, replay, motion) = make_problem(in_features, hidden, B, seed, gain)
    upd = amp.AmpDiscriminatorUpdate(in_features, hidden, max_batch_rows=max(B, 128), device=DEV)
    terms, gW, gb, logits = upd(W, b, agent.to(DEV), replay.to(DEV), motion.to(DEV), return_logits=True)
    torch.cuda.synchronize()
    cfg = DiscLossCfg()
    names = ["gW1", "gW2", "gW3", "gb1", "gb2", "gb3"]

    loss_e, terms_e, gW_e, gb_e = discriminator_loss_manual(W, b, agent, replay, motion, cfg, dtype=torch.float64, emulate_bf16=True)
    for name, got, ref in zip(names, gW + gb, gW_e + gb_e):
        compare(got, ref, 4e-3, f"{name} vs same-rounding oracle")
    loss_a, terms_a, gW_a, gb_a = discriminator_loss_autograd(W, b, agent, replay, motion, cfg, dtype=torch.float32)
    if B >= 64:  # with a handful of rows one ReLU unit whose pre-activation changes sign under bf16 rounding moves a whole
        # gradient row by O(1); the fp32 comparison is only meaningful once the batch averages over such flips
        for name, got, ref in zip(names, gW + gb, gW_a + gb_a):
            compare(got, ref, 3e-2 * max(1.0, (1024.0 / B) ** 0.5), f"{name} vs fp32 autograd oracle", cos_min=0.999)

    t = terms.cpu().double()
    for i, key in enumerate(["bce_agent_replay", "bce_motion", "logit_regularization", "gradient_penalty", "weight_decay"]):
        for ref, rel in ((float(terms_e[key]), 2e-3), (float(terms_a[key]), 1e-2)):
            assert abs(float(t[i]) - ref) <= rel * max(1.0, abs(ref)), f"{key}: {float(t[i])} vs {ref}"
    assert abs(float(t[5]) - float(loss_a)) <= 1e-2 * max(1.0, abs(float(loss_a)))

    # logits of the three batches (fp32 oracle forward)
    from oracle.disc_train_oracle import _mlp

    ref_logits = torch.stack([_mlp(x, W, b).reshape(-1) for x in (agent, replay, motion)])
    span = max(1.0, float(ref_logits.abs().max()))
    assert float((logits.cpu() - ref_logits).abs().max()) <= 1e-2 * span
    return upd


@pytest.mark.parametrize("in_features", [166, 830, 162])
def test_reference_shape_batch_4096(in_features):
    """The reference configuration: 1024-512 discriminator, discriminator_batch_size 4096, K*A = 166 / 830 / 162."""
    run_and_check(in_features, (1024, 512), 4096)


@pytest.mark.parametrize("B", [1, 100, 128, 129, 1000])
def test_ragged_batches(B):
    """Batch lengths that are not multiples of the 128-row tile (padding rows must contribute nothing)."""
    run_and_check(166, (1024, 512), B, seed=B)


def test_other_hidden_sizes_and_wide_logits():
    run_and_check(64, (256, 256), 512, seed=11)        # in_features a multiple of 64: no padding column at all
    run_and_check(830, (1024, 512), 512, seed=12, gain=20.0)  # saturated sigmoids


def test_reuse_with_shrinking_batch_and_gradient_bucket():
    """One handle, a large step then a small one (stale rows of the first must not leak), gradients written in place into
    views of one flat bucket as ``reduce_parameters`` expects them."""
    import humanoid_amp_b200 as amp
    from oracle.disc_train_oracle import discriminator_loss_manual

    in_features, hidden = 166, (1024, 512)
    upd = amp.AmpDiscriminatorUpdate(in_features, hidden, max_batch_rows=1024, device=DEV)
    for B, seed in ((1024, 21), (200, 22)):
        W, b, (agent, replay, motion) = make_problem(in_features, hidden, B, seed)
        sizes = [w.numel() for w in W] + [x.numel() for x in b]
        bucket = torch.full((sum(sizes),), float("nan"), device=DEV)
        views = list(torch.split(bucket, sizes))
        gW = [v.view_as(w) for v, w in zip(views[:3], W)]
        gb = [v.view_as(x) for v, x in zip(views[3:], b)]
        upd(W, b, agent.to(DEV), replay.to(DEV), motion.to(DEV), grad_weights=gW, grad_biases=gb)
        assert bool(torch.isfinite(bucket).all())
        _, _, gW_e, gb_e = discriminator_loss_manual(W, b, agent, replay, motion, dtype=torch.float64, emulate_bf16=True)
        for name, got, ref in zip(["gW1", "gW2", "gW3", "gb1", "gb2", "gb3"], gW + gb, gW_e + gb_e):
            compare(got, ref, 4e-3, f"{name} (B={B})")


def test_with_running_standard_scaler_in_train_mode():
    """skrl feeds the three RAW batches through amp_state_preprocessor(train=True) one after the other: the statistics move
    between the batches."""
    import humanoid_amp_b200 as amp
    from oracle import OracleDiscriminator
    from oracle.disc_train_oracle import discriminator_loss_manual

    in_features, hidden, B = 166, (1024, 512), 512
    W, b, raw = make_problem(in_features, hidden, B, seed=31)
    g = torch.Generator().manual_seed(5)
    scale, shift = torch.rand(in_features, generator=g) * 3 + 0.1, torch.randn(in_features, generator=g)
    raw = [x * scale + shift for x in raw]
    scaler = amp.RunningStandardScaler(in_features, device=DEV)
    upd = amp.AmpDiscriminatorUpdate(in_features, hidden, max_batch_rows=B, device=DEV)
    terms, gW, gb = upd(W, b, *[x.to(DEV) for x in raw], scaler=scaler, train=True)

    ora = OracleDiscriminator(in_features, weights=W, biases=b)
    normed = []
    for x in raw:
        ora.update_statistics(x)
        normed.append(ora.normalise(x))
    assert torch.allclose(scaler.running_mean.cpu(), ora.running_mean, rtol=1e-6, atol=1e-7)
    _, _, gW_e, gb_e = discriminator_loss_manual(W, b, *normed, dtype=torch.float64, emulate_bf16=True)
    for name, got, ref in zip(["gW1", "gW2", "gW3", "gb1", "gb2", "gb3"], gW + gb, gW_e + gb_e):
        compare(got, ref, 6e-3, name)  # a normalised value on a bf16 rounding boundary may round the other way


def test_full_minibatch_size_through_a_size_independent_property():
    """skrl's full mini-batch (rollouts 16 x 4096 envs / 2 mini-batches = 32 768 rows per source, what the update sees when
    ``discriminator_batch_size`` is 0) is too large for the CPU oracle to finish in seconds.  Property: every loss term is a
    MEAN over rows or does not depend on the rows, so the gradient of the whole batch equals the average of the gradients
    of its two halves (each half paired across the three sources)."""
    import humanoid_amp_b200 as amp

    in_features, hidden, B = 830, (1024, 512), 32768
    W, b, batches = make_problem(in_features, hidden, B, seed=77)
    W, b = [w.to(DEV) for w in W], [x.to(DEV) for x in b]
    batches = [x.to(DEV) for x in batches]
    upd = amp.AmpDiscriminatorUpdate(in_features, hidden, max_batch_rows=B, device=DEV)
    t_full, gW, gb = upd(W, b, *batches)
    full = [g.clone() for g in gW + gb]
    halves = []
    for sl in (slice(0, B // 2), slice(B // 2, B)):
        t_half, hW, hb = upd(W, b, *[x[sl] for x in batches])
        halves.append(([g.clone() for g in hW + hb], t_half.clone()))
    for i, name in enumerate(["gW1", "gW2", "gW3", "gb1", "gb2", "gb3"]):
        compare(full[i], 0.5 * (halves[0][0][i] + halves[1][0][i]), 2e-3, f"{name}: whole batch vs mean of halves")
    t = 0.5 * (halves[0][1] + halves[1][1])
    assert torch.allclose(t_full, t, rtol=2e-3, atol=1e-5)
    assert bool(torch.isfinite(t_full).all())


def test_whole_update_replays_as_a_cuda_graph():
    """Staging + loss + gradients enqueue only stream work (no host synchronisation, no allocation in the library), so one
    mini-batch update with static buffers is capturable; the replay must reproduce the eager result and track new inputs
    written into the same buffers."""
    import humanoid_amp_b200 as amp

    in_features, hidden, B = 166, (1024, 512), 512
    W, b, batches = make_problem(in_features, hidden, B, seed=41)
    W, b = [w.to(DEV) for w in W], [x.to(DEV) for x in b]
    bufs = [x.to(DEV).clone() for x in batches]
    gW, gb = [torch.zeros_like(w) for w in W], [torch.zeros_like(x) for x in b]
    scaler = amp.RunningStandardScaler(in_features, device=DEV)
    scaler.update(torch.cat(bufs))
    upd = amp.AmpDiscriminatorUpdate(in_features, hidden, max_batch_rows=B, device=DEV)
    terms_box = {}

    def step():
        terms_box["t"] = upd(W, b, *bufs, scaler=scaler, train=False, grad_weights=gW, grad_biases=gb)[0]

    step()
    eager = [g.clone() for g in gW + gb]
    graph = amp.capture_step(step, DEV)
    for g in gW + gb:
        g.zero_()
    graph.replay()
    torch.cuda.synchronize()
    for got, ref in zip(gW + gb, eager):
        compare(got, ref, 1e-5, "graph replay vs eager")  # atomics order only
    # new data in the same buffers: the replay follows
    _, _, fresh = make_problem(in_features, hidden, B, seed=42)
    for dst, src in zip(bufs, fresh):
        dst.copy_(src)
    graph.replay()
    torch.cuda.synchronize()
    replayed = [g.clone() for g in gW + gb]
    step()
    torch.cuda.synchronize()
    for got, ref in zip(replayed, gW + gb):
        compare(got, ref, 1e-5, "graph replay on new inputs vs eager")
    assert not torch.allclose(replayed[0], eager[0])


def test_errors():
    import humanoid_amp_b200 as amp

    upd = amp.AmpDiscriminatorUpdate(166, (1024, 512), max_batch_rows=256, device=DEV)
    W, b, (agent, replay, motion) = make_problem(166, (1024, 512), 64, 1)
    with pytest.raises(RuntimeError):
        upd.loss_and_grads(W, b)  # nothing staged
    upd.stage(0, agent.to(DEV))
    with pytest.raises(RuntimeError):
        upd.stage(1, replay[:32].to(DEV))  # lengths differ
    with pytest.raises(amp.AmpB200Error):
        upd.loss_and_grads(W, b)  # sources 1 and 2 missing
    with pytest.raises(amp.AmpB200Error):
        amp.AmpDiscriminatorUpdate(166, (1000, 512), device=DEV)  # hidden sizes must be multiples of 256
    big = torch.zeros(512, 166, device=DEV)
    upd2 = amp.AmpDiscriminatorUpdate(166, (1024, 512), max_batch_rows=256, device=DEV)
    with pytest.raises(amp.AmpB200Error):
        upd2.stage(0, big)  # exceeds max_batch_rows
    # a scaler with non-default settings is refused instead of being normalised with the defaults
    odd = amp.RunningStandardScaler(166, epsilon=1e-3, clip_threshold=2.0, device=DEV)
    upd3 = amp.AmpDiscriminatorUpdate(166, (1024, 512), max_batch_rows=256, device=DEV)
    with pytest.raises(RuntimeError, match="epsilon"):
        upd3.stage(0, agent.to(DEV), scaler=odd)
