"""Parity of the per-step env rows (``-m gpu``) against fixtures written by the REFERENCE'S OWN TEXT.

``tests/golden/vectors.npz::env/*`` holds simulator-state sequences and what ``G1AmpEnv._get_observations``
(g1_amp_env.py:175-242), ``HumanoidAmpEnv._get_observations`` (humanoid_amp_env.py:105-126), ``G1AmpEnv._get_rewards``
(:246-319 with compute_rewards / exp_reward_with_floor) and ``G1AmpEnv._reset_strategy_random`` (:371-441) made of them --
those methods were cut out of the reference with ``ast`` (``oracle/build_ref.py``) and executed unmodified
(``oracle/ref_harness.py``, ``tests/golden/make_golden.py``).  The CUDA path is called through the C ABI.

Bar: copies bit-identical; the six tangent/normal columns within 1e-6 + 1e-5 |ref| (closed-form rotation columns vs
torch's quat_apply: <= 1 ulp); per-joint reward sums within 2e-5 relative (fixed shuffle-tree order vs torch.sum).
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import clip_path

pytestmark = pytest.mark.gpu

SIM_KEYS = ("joint_pos", "joint_vel", "body_pos_w", "body_quat_w", "body_lin_vel_w", "body_ang_vel_w")
G1_OBS_CASES = [  # (tag, K, num_actor_observations, rew_track_vel, history_include_last_actions, history_include_command)
    ("k2_a1", 2, 1, 0.0, True, True), ("k10_a1", 10, 1, 0.0, True, True), ("k1_a1_cmd", 1, 1, 1.0, True, True),
    ("k3_a3_cmd", 3, 3, 1.0, True, True), ("k2_a4_noact", 2, 4, 1.0, False, True), ("k2_a3_nocmd", 2, 3, 1.0, True, False),
    ("k2_a5", 2, 5, 0.0, True, True),
]  # fmt: skip
REWARD_SCALES = dict(rew_termination=-1.0, rew_action_l2=-0.1, rew_joint_pos_limits=-10.0, rew_joint_acc_l2=-1.0e-06, rew_joint_vel_l2=-0.001)


def close(actual, expected, rtol=1e-5, atol=1e-6):
    a = actual.detach().cpu().numpy().astype(np.float64) if isinstance(actual, torch.Tensor) else np.asarray(actual, dtype=np.float64)
    e = np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, (a.shape, e.shape)
    err = np.abs(a - e)
    bad = ~(err <= atol + rtol * np.abs(e))
    assert not bad.any(), f"{bad.sum()} / {bad.size} out of tolerance, max err {np.nanmax(err):.3e}"


def obs_rows_match(got, want, D, A):
    """(..., k*A) AMP rows: everything but the six tangent/normal columns bit-identical, those within the bar."""
    g = got.detach().cpu().numpy().reshape(-1, A)
    w = np.asarray(want).reshape(-1, A)
    tn = slice(2 * D + 1, 2 * D + 7)
    cols = np.r_[0 : 2 * D + 1, 2 * D + 7 : A]
    assert np.array_equal(g[:, cols], w[:, cols])
    close(g[:, tn], w[:, tn])


@pytest.fixture(scope="module")
def amp():
    import humanoid_amp_b200 as amp

    return amp


def sim_state(golden, robot_name, step):
    return [torch.from_numpy(golden[f"env/{robot_name}/{k}"][step]).cuda() for k in SIM_KEYS]


@pytest.mark.parametrize("fused", [False, True], ids=["three_launches", "one_launch"])
@pytest.mark.parametrize("case", G1_OBS_CASES, ids=[c[0] for c in G1_OBS_CASES])
def test_g1_get_observations_vs_reference_text(golden, amp, case, fused):
    tag, K, n_actor, track, inc_act, inc_cmd = case
    n_steps, N = golden["env/g1/joint_pos"].shape[:2]
    loader = amp.MotionLoader(clip_path("G1_dance"), "cuda:0")
    cfg = amp.AmpEnvCfg(motion_file="<preloaded>", num_envs=N, num_amp_observations=K, robot=amp.G1, num_actor_observations=n_actor,
                        rew_track_vel=track, history_include_last_actions=inc_act, history_include_command=inc_cmd)  # fmt: skip
    env = amp.AmpEnvPath(cfg, "cuda:0", motion_loader=loader)
    kept = golden[f"env/g1/{tag}/steps"].tolist()
    A, D = amp.G1.amp_observation_space, 29
    base = A - 12
    for s in range(n_steps):
        env.last_actions.copy_(torch.from_numpy(golden["env/g1/last_actions"][s]))
        env.command_target_speed.copy_(torch.from_numpy(golden["env/g1/command"][s]))
        if n_actor > 1:
            env._just_reset_mask |= torch.from_numpy(golden["env/g1/reset_mask"][s]).cuda()
        if fused:
            policy = env.step_observations(*sim_state(golden, "g1", s))["policy"]
        else:
            policy = env.get_observations(*sim_state(golden, "g1", s))["policy"]
        if s not in kept:
            continue
        j = kept.index(s)
        obs_rows_match(env.extras["amp_obs"], golden[f"env/g1/{tag}/amp_obs"][j], D, A)
        want = golden[f"env/g1/{tag}/policy"][j]
        assert policy.shape == want.shape == (N, cfg.observation_space)
        g = policy.cpu().numpy()
        # the base block repeats in every history frame: tangent/normal columns at the same offsets inside each block
        cur = base + 29 + (2 if track > 0 else 0)
        P = cfg.hist_frame_size
        starts = [0] + [cur + i * P for i in range(n_actor - 1)]
        tn_cols = np.concatenate([np.arange(st + 2 * D + 1, st + 2 * D + 7) for st in starts])
        other = np.setdiff1d(np.arange(want.shape[1]), tn_cols)
        assert np.array_equal(g[:, other], want[:, other])
        close(g[:, tn_cols], want[:, tn_cols])


@pytest.mark.parametrize("K", [2, 10])
def test_humanoid_get_observations_vs_reference_text(golden, amp, K):
    n_steps, N = golden["env/humanoid28/joint_pos"].shape[:2]
    loader = amp.MotionLoader(clip_path("humanoid_walk"), "cuda:0")
    cfg = amp.AmpEnvCfg(motion_file="<preloaded>", num_envs=N, num_amp_observations=K, robot=amp.HUMANOID28)
    env = amp.AmpEnvPath(cfg, "cuda:0", motion_loader=loader)
    kept = list(range(n_steps)) if K == 2 else [2, n_steps - 1]
    A, D = 81, 28
    for s in range(n_steps):
        policy = torch.empty(N, A - 12, device="cuda")
        view = env.update_amp_observations(*sim_state(golden, "humanoid28", s), policy_obs=policy)
        if s in kept:
            j = kept.index(s)
            obs_rows_match(view, golden[f"env/humanoid28/k{K}/amp_obs"][j], D, A)
            # HumanoidAmpEnv returns the whole row as the policy observation (humanoid_amp_env.py:126)
            obs_rows_match(view[:, :A], golden[f"env/humanoid28/k{K}/policy"][j], D, A)
            assert torch.equal(policy, view[:, : A - 12])


@pytest.mark.parametrize("track", [0.0, 1.0])
def test_g1_task_reward_vs_reference_text(golden, amp, track):
    N = golden["env/g1/joint_pos"].shape[1]
    loader = amp.MotionLoader(clip_path("G1_dance"), "cuda:0")
    cfg = amp.AmpEnvCfg(motion_file="<preloaded>", num_envs=N, num_amp_observations=2, robot=amp.G1, rew_track_vel=track, **REWARD_SCALES)
    env = amp.AmpEnvPath(cfg, "cuda:0", motion_loader=loader)
    jp, jv, bp, bq, _, ba = sim_state(golden, "g1", 0)
    bl = torch.from_numpy(golden["env/g1/reward/body_lin_vel_w"]).cuda()
    env.command_target_speed.copy_(torch.from_numpy(golden["env/g1/command"][0]))
    total, terms, err = env.get_rewards(
        torch.from_numpy(golden["env/g1/reward/terminated"]).cuda(), torch.from_numpy(golden["env/g1/last_actions"][0]).cuda(), jp,
        torch.from_numpy(golden["env/g1/reward/soft_limits"]).cuda(), torch.from_numpy(golden["env/g1/reward/joint_acc"]).cuda(), jv, bl, bq,
        return_terms=True,
    )  # fmt: skip
    t = f"env/g1/reward/track{int(track)}"
    close(total, golden[f"{t}/total"], rtol=2e-5, atol=2e-6)
    log = dict(zip(golden[f"{t}/log_keys"].tolist(), golden[f"{t}/log_values"].tolist()))
    means = terms.double().mean(dim=0).cpu().numpy()
    for i, key in enumerate(("pub_termination", "pub_action_l2", "pub_joint_pos_limits", "pub_joint_acc_l2", "pub_joint_vel_l2")):
        assert abs(means[i] - log[key]) <= 2e-5 * abs(log[key]) + 1e-7, key
    assert abs(float(total.double().mean()) - log["total_reward"]) <= 2e-5 * abs(log["total_reward"]) + 1e-6
    if track > 0:
        assert abs(means[5] - log["rew_track_vel"]) <= 2e-5 * abs(log["rew_track_vel"]) + 1e-7
        assert abs(float(err.double().mean()) - log["error_track_vel"]) <= 1e-5 * abs(log["error_track_vel"]) + 1e-7


@pytest.mark.parametrize("K", [2, 10])
def test_g1_reset_strategy_random_vs_reference_text(golden, amp, K):
    r = f"env/g1/reset_k{K}"
    N = golden[f"{r}/default_root_state"].shape[0]
    loader = amp.MotionLoader(clip_path("G1_dance"), "cuda:0")
    cfg = amp.AmpEnvCfg(motion_file="<preloaded>", num_envs=N, num_amp_observations=K, robot=amp.G1)
    env = amp.AmpEnvPath(cfg, "cuda:0", motion_loader=loader)
    env.amp_observation_buffer.fill_(3.0)
    env_ids = torch.from_numpy(golden[f"{r}/env_ids"])
    np.random.seed(99)
    root, dof_p, dof_v, mids, times = env.reset_strategy_random(
        env_ids, torch.from_numpy(golden[f"{r}/default_root_state"])[env_ids].cuda(), torch.from_numpy(golden[f"{r}/env_origins"])[env_ids].cuda())
    assert np.array_equal(mids, golden[f"{r}/motion_ids"][env_ids.numpy()])
    assert np.array_equal(times.astype(np.float32), golden[f"{r}/motion_start_times"][env_ids.numpy()])
    close(root, golden[f"{r}/root_state"])  # columns 3:7 are a slerp (<= 1 ulp), the rest lerps
    keep = np.r_[0:3, 7:13]
    assert np.array_equal(root.cpu().numpy()[:, keep], golden[f"{r}/root_state"][:, keep])
    assert np.array_equal(dof_p.cpu().numpy(), golden[f"{r}/dof_pos"]) and np.array_equal(dof_v.cpu().numpy(), golden[f"{r}/dof_vel"])
    obs_rows_match(env.amp_observation_buffer, golden[f"{r}/amp_observation_buffer"], 29, 83)


def test_humanoid_reset_uses_torso_and_its_own_lift(amp):
    """``HumanoidAmpEnv._reset_strategy_random`` takes the root from ``torso`` and lifts it by 0.15
    (humanoid_amp_env.py:194-201), not ``pelvis`` + 0.05 as the G1 env does -- checked against the oracle restatement (the
    reference method itself no longer runs against the current ``MotionLoader.sample_times`` signature)."""
    from oracle import OracleMotionLoader, env_oracle

    loader = amp.MotionLoader(clip_path("humanoid_walk"), "cuda:0")
    ora = OracleMotionLoader([clip_path("humanoid_walk")])
    assert "pelvis" in loader.body_names and "torso" in loader.body_names
    N, K = 40, 2
    env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file="<preloaded>", num_envs=N, num_amp_observations=K, robot=amp.HUMANOID28), "cuda:0", motion_loader=loader)
    g = torch.Generator().manual_seed(4)
    default_root, origins = torch.randn(N, 13, generator=g), torch.randn(N, 3, generator=g)
    env_ids = torch.arange(0, N, 2)
    np.random.seed(5)
    root, dof_p, dof_v, mids, times = env.reset_strategy_random(env_ids, default_root[env_ids].cuda(), origins[env_ids].cuda())
    w_root, w_dp, w_dv = env_oracle.reset_root_and_dof_state(ora, times, mids, default_root[env_ids], origins[env_ids],
                                                            ora.get_dof_index(amp.HUMANOID28.joint_names), ora.get_body_index(["torso"])[0], lift=0.15)  # fmt: skip
    close(root, w_root.numpy())
    assert np.array_equal(dof_p.cpu().numpy(), w_dp.numpy()) and np.array_equal(dof_v.cpu().numpy(), w_dv.numpy())
    wrong, _, _ = env_oracle.reset_root_and_dof_state(ora, times, mids, default_root[env_ids], origins[env_ids],
                                                     ora.get_dof_index(amp.HUMANOID28.joint_names), ora.get_body_index(["pelvis"])[0])  # fmt: skip
    assert not np.allclose(root.cpu().numpy(), wrong.numpy(), atol=1e-3)


@pytest.mark.parametrize("N", [24, 20011])
def test_fused_step_equals_the_three_entry_points(golden, amp, N):
    """``amp_env_step`` (one launch) == ``amp_obs_step`` + ``amp_actor_obs_step`` + ``amp_task_reward`` bit for bit, including
    the in-place histories, the warm start and the cleared reset mask; N = 20011 exceeds the resident warps (grid-stride)."""
    from humanoid_amp_b200.synthetic import synthetic_sim_state

    loader = amp.MotionLoader(clip_path("G1_dance"), "cuda:0")
    cfg = amp.AmpEnvCfg(motion_file="<preloaded>", num_envs=N, num_amp_observations=3, robot=amp.G1, num_actor_observations=4,
                        rew_track_vel=1.0, **REWARD_SCALES)  # fmt: skip
    a = amp.AmpEnvPath(cfg, "cuda:0", motion_loader=loader)
    b = amp.AmpEnvPath(cfg, "cuda:0", motion_loader=loader)
    g = torch.Generator().manual_seed(N)
    for step in range(5):
        state = synthetic_sim_state(N, amp.G1, "cuda", seed=40 + step)
        actions = torch.randn(N, 29, generator=g).cuda()
        command = (torch.rand(N, 2, generator=g) * 2 - 1).cuda()
        mask = (torch.rand(N, generator=g) < 0.25).cuda()
        limits = torch.stack([torch.full((N, 29), -1.0), torch.full((N, 29), 1.5)], dim=-1).cuda()
        acc = (torch.randn(N, 29, generator=g) * 30).cuda()
        term = (torch.rand(N, generator=g) < 0.1).cuda()
        for env in (a, b):
            env.last_actions.copy_(actions)
            env.command_target_speed.copy_(command)
            env._just_reset_mask |= mask
        want_policy = a.get_observations(*state)["policy"]
        want_total, want_terms, want_err = a.get_rewards(term, actions, state[0], limits, acc, state[1], state[4], state[3], return_terms=True)
        got = b.step_observations(*state, reward_inputs=dict(reset_terminated=term, actions=actions, soft_joint_pos_limits=limits,
                                                             joint_acc=acc, return_terms=True))  # fmt: skip
        assert torch.equal(got["policy"], want_policy)
        assert torch.equal(b.amp_observation_buffer, a.amp_observation_buffer)
        assert torch.equal(b.actor_obs_history_buffer, a.actor_obs_history_buffer)
        assert not b._just_reset_mask.any() and not a._just_reset_mask.any()
        assert torch.equal(got["reward"], want_total) and torch.equal(got["reward_terms"], want_terms) and torch.equal(got["track_err"], want_err)
