"""The discriminator side of one skrl ``AMP._update`` end to end (``-m gpu``), every stage through the C ABI, against the same
flow restated with the oracle (SURVEY.md sections 3.4 and 8f-2; upstream skrl, parity unpinned):

    refill the motion dataset with collect_reference_motions  ->  style reward of the rollout's AMP states (eval-mode scaler)
    ->  per mini-batch: sample replay / motion rows by index, amp_state_preprocessor(train=True) on the three batches in
    turn, discriminator loss + gradients  ->  push the rollout's AMP states into the replay buffer

Tolerances: memories bit-exact, scaler statistics rtol 2e-6, style reward and gradients the bf16 bars of
test_gpu_discriminator.py / test_gpu_disc_update.py.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import clip_path

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_discriminator_side_of_one_amp_update():
    import humanoid_amp_b200 as amp
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params
    from oracle import OracleDiscriminator, OracleMotionLoader, OracleRandomMemory, env_oracle
    from oracle.disc_train_oracle import discriminator_loss_manual

    K, n_envs, rollouts, mini_batches, disc_batch = 2, 256, 4, 2, 384
    clip = clip_path("G1_walk")
    env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file=clip, num_envs=n_envs, num_amp_observations=K, robot=amp.G1), DEV)
    ora_loader = OracleMotionLoader([clip])
    width = K * 83
    dof_idx, key_idx = ora_loader.get_dof_index(amp.G1.joint_names), ora_loader.get_body_index(amp.G1.key_body_names)

    def oracle_collect(times, ids):
        return env_oracle.collect_reference_motions(ora_loader, len(times), K, dof_idx, 0, key_idx, current_times=times, motion_ids=ids)

    rng = np.random.default_rng(11)
    W, b = skrl_style_discriminator_params(width, seed=42, logit_gain=3.0)

    # ---- memories: motion dataset refilled in place by the fused collect kernel, replay buffer pre-filled ----
    motion_dataset, reply_buffer = amp.AmpStateMemory(1000, width, DEV), amp.AmpStateMemory(1500, width, DEV)
    o_motion, o_reply = OracleRandomMemory(1000, width), OracleRandomMemory(1500, width)
    for n in (700, 500):  # the second refill wraps around the ring
        times = rng.uniform(0, ora_loader.durations[0], n)
        ids = np.zeros(n, dtype=np.int64)
        motion_dataset.memory_index = env.collect_reference_motions_into(motion_dataset.states, motion_dataset.memory_index, n, times, ids)
        o_motion.add_samples(oracle_collect(times, ids))
    assert torch.allclose(motion_dataset.states.cpu(), o_motion.states, rtol=1e-5, atol=1e-6)
    assert motion_dataset.memory_index == o_motion.memory_index
    motion_dataset.filled = o_motion.filled  # rows were written in place: the caller keeps the memory's bookkeeping
    seed_rows = oracle_collect(rng.uniform(0, ora_loader.durations[0], 900), np.zeros(900, dtype=np.int64)) * 1.1
    reply_buffer.add_samples(seed_rows.to(DEV))
    o_reply.add_samples(seed_rows)

    # ---- the rollout's AMP states (what env.extras["amp_obs"] delivered) ----
    amp_states = oracle_collect(rng.uniform(0, ora_loader.durations[0], n_envs * rollouts), np.zeros(n_envs * rollouts, dtype=np.int64)) + 0.05
    amp_states_d = amp_states.to(DEV)

    # ---- scaler with some history; style reward with the eval-mode scaler ----
    scaler = amp.RunningStandardScaler(width, device=DEV)
    o_disc = OracleDiscriminator(width, weights=W, biases=b, reward_scale=2.0)
    warm = oracle_collect(rng.uniform(0, ora_loader.durations[0], 512), np.zeros(512, dtype=np.int64))
    scaler.update(warm.to(DEV))
    o_disc.update_statistics(warm)
    disc = amp.AmpDiscriminator(width, reward_scale=2.0, device=DEV, max_rows=n_envs * rollouts)
    disc.load(W, b, scaler.running_mean, scaler.running_variance)
    style, logits = disc.style_reward(amp_states_d, return_logits=True)
    want_logits = o_disc.logits(amp_states)
    span = max(1.0, float(want_logits.abs().max()))
    assert float((logits.cpu() - want_logits).abs().max()) <= 1e-2 * span
    assert float((style.cpu() - o_disc.style_reward(amp_states)).abs().max()) <= 2e-2 * span

    # ---- mini-batches: sampled indexes are shared by both sides (the RNG stream itself is torch's) ----
    upd = amp.AmpDiscriminatorUpdate(width, (1024, 512), max_batch_rows=disc_batch, device=DEV)
    g = torch.Generator().manual_seed(5)
    batch = n_envs * rollouts // mini_batches
    agent_idx = torch.randperm(n_envs * rollouts, generator=g)
    for mb in range(mini_batches):
        idx_a = agent_idx[mb * batch : (mb + 1) * batch][:disc_batch]
        idx_r = torch.randint(0, len(o_reply), (batch,), generator=g)[:disc_batch]
        idx_m = torch.randint(0, len(o_motion), (batch,), generator=g)[:disc_batch]
        rows_a = amp_states_d[idx_a.to(DEV)]
        rows_r = reply_buffer.sample_by_index(idx_r.to(DEV))[0]
        rows_m = motion_dataset.sample_by_index(idx_m.to(DEV))[0]
        assert torch.equal(rows_r.cpu(), o_reply.sample_by_index(idx_r)[0])
        terms, gW, gb = upd(W, b, rows_a, rows_r, rows_m, scaler=scaler, train=True)

        normed = []
        for rows in (amp_states[idx_a], o_reply.sample_by_index(idx_r)[0], o_motion.sample_by_index(idx_m)[0]):
            o_disc.update_statistics(rows)
            normed.append(o_disc.normalise(rows))
        assert torch.allclose(scaler.running_mean.cpu(), o_disc.running_mean, rtol=2e-6, atol=1e-7)
        assert torch.allclose(scaler.running_variance.cpu(), o_disc.running_variance, rtol=2e-5, atol=1e-9)
        loss, terms_o, gW_o, gb_o = discriminator_loss_manual(W, b, *normed, dtype=torch.float64, emulate_bf16=True)
        for name, got, ref in zip(["gW1", "gW2", "gW3", "gb1", "gb2", "gb3"], gW + gb, gW_o + gb_o):
            err = float((got.cpu().double() - ref).abs().max())
            assert err <= 8e-3 * float(ref.abs().max()), f"mini-batch {mb} {name}: {err:.3e}"
        assert abs(float(terms[5]) - float(loss)) <= 5e-3 * max(1.0, abs(float(loss)))

    # ---- the rollout's AMP states go into the replay buffer ----
    reply_buffer.add_samples(amp_states_d)
    o_reply.add_samples(amp_states)
    assert torch.equal(reply_buffer.states.cpu(), o_reply.states) and reply_buffer.memory_index == o_reply.memory_index
