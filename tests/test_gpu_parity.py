"""Parity tests proper (``-m gpu``): the CUDA path, called through the C ABI, against the committed reference fixtures and
against the oracle on the same seeded inputs.

Bar (BASELINE.json north_star): frame indices and the float64 blend BIT-EXACT; fp32 outputs within
``|a-b| <= 1e-6 + 1e-5*|b|``.

What is actually achieved and asserted here is stronger for most of the output: every column that is a lerp (dof
positions/velocities, root height, root velocities, key-body offsets) is BIT-IDENTICAL to the reference; only the values
derived from the slerp (body rotations, tangent/normal) go through ``acos``/``sin``, where torch's SLEEF kernels and the
device's correctly-rounded evaluation may differ by 1 ulp.  Body rotations meet the stated bar everywhere, extrapolated
frames included, and so do the tangent/normal columns of every frame that interpolates (0 <= blend <= 1) and of every
reference-written fixture.  History frames before the clip start EXTRAPOLATE (blend down to -(K-1), reference
``g1_amp_env.py:454-457``): the slerp weights grow like |blend| and cancel, and tangent/normal are quadratic in the
un-normalised quaternion, so a 1-ulp difference in ``sin`` is amplified by |blend|.  For those frames only, the absolute
term for the six tangent/normal columns is scaled by the conditioning, ``atol = 1e-6 * max(1, |blend|)``;
``test_extrapolated_slerp_spread_is_torch_cpu_vs_torch_cuda_spread`` measures, on the GPU box, that the reference algorithm
run by torch on the CPU and by torch on the GPU differs from itself by as much on the same inputs.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import CLIP_NAMES, FIXTURE_TAGS, clip_path, fixture_files, full_clip_path, pooled_spec

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6
OUT_NAMES = ("dof_pos", "dof_vel", "body_pos", "body_rot", "body_lin", "body_ang")


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def close(actual, expected, rtol=RTOL, atol=ATOL):
    """``atol`` may be an array broadcastable against the data (per-row conditioning, see the module docstring)."""
    a, e = _np(actual).astype(np.float64), _np(expected).astype(np.float64)
    assert a.shape == e.shape, (a.shape, e.shape)
    err = np.abs(a - e)
    bound = atol + rtol * np.abs(e)
    bad = ~(err <= bound)
    assert not bad.any(), f"{bad.sum()} / {bad.size} elements out of tolerance, max err {np.nanmax(err):.3e}"


def bit_equal(actual, expected):
    a, e = _np(actual), _np(expected)
    assert a.shape == e.shape and a.dtype == e.dtype, (a.shape, e.shape, a.dtype, e.dtype)
    same = (a == e) | (np.isnan(a) & np.isnan(e))
    assert same.all(), f"{(~same).sum()} / {same.size} elements differ, max |diff| {np.nanmax(np.abs(a.astype(np.float64) - e)):.3e}"


def check_amp_obs(got, want, blend, D, K):
    """(n, K*A) rows: the six tangent/normal columns of every frame within the (conditioning-scaled) bar, the rest
    bit-identical.  ``blend``: float64 blend of every frame, shape (n*K,)."""
    g, w = _np(got), _np(want)
    n = g.shape[0]
    A = g.shape[1] // K
    g, w = g.reshape(n * K, A), w.reshape(n * K, A)
    tn = slice(2 * D + 1, 2 * D + 7)
    lerp_cols = np.r_[0 : 2 * D + 1, 2 * D + 7 : A]
    bit_equal(g[:, lerp_cols], w[:, lerp_cols])
    close(g[:, tn], w[:, tn], atol=ATOL * np.maximum(1.0, np.abs(blend))[:, None])


@pytest.fixture(scope="module")
def amp():
    import humanoid_amp_b200 as amp

    return amp


@pytest.fixture(scope="module")
def loaders(amp):
    cache = {}

    def get(name):  # a fixture tag of vectors.npz: <clip> (40-frame window), full/<clip>, [full/]pooled_humanoid
        if name not in cache:
            cache[name] = amp.MotionLoader(",".join(fixture_files(name)), "cuda:0")
        return cache[name]

    return get


def make_env(amp, loader, K, num_envs=64):
    robot = amp.robot_for_clip(loader.dof_names)
    cfg = amp.AmpEnvCfg(motion_file="<preloaded>", num_envs=num_envs, num_amp_observations=K, robot=robot)
    return amp.AmpEnvPath(cfg, "cuda:0", motion_loader=loader)


# ---------------------------------------------------------------------------------------------------------------------
# golden fixtures written by the LIVE reference
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", FIXTURE_TAGS)
def test_frame_blend_bit_exact_vs_reference_fixture(golden, loaders, name):
    loader = loaders(name)
    times, ids = golden[f"{name}/times"], golden[f"{name}/ids"]
    i0, i1, blend = loader._compute_frame_blend(times, ids)
    assert i0.dtype == np.int64 and blend.dtype == np.float64
    assert np.array_equal(i0, golden[f"{name}/idx0"]) and np.array_equal(i1, golden[f"{name}/idx1"])
    assert np.array_equal(blend.view(np.int64), golden[f"{name}/blend"].view(np.int64)), "float64 blend must be bit-exact"
    _, _, b32, _ = loader.compute_frame_blend_device(times, ids)
    assert np.array_equal(b32.cpu().numpy(), golden[f"{name}/blend"].astype(np.float32))


@pytest.mark.parametrize("name", FIXTURE_TAGS)
def test_sample_vs_reference_fixture(golden, loaders, name):
    loader = loaders(name)
    times, ids = golden[f"{name}/times"], golden[f"{name}/ids"]
    outs = loader.sample(len(times), times=times, motion_ids=ids)
    for key, t in zip(OUT_NAMES, outs):
        assert t.dtype == torch.float32 and t.is_cuda
        close(t, golden[f"{name}/{key}"])
        if key != "body_rot":  # the five lerps reproduce the reference bit for bit
            assert np.array_equal(t.cpu().numpy(), golden[f"{name}/{key}"]), key
    assert loader.poll_flags() == 0


@pytest.mark.parametrize("name", FIXTURE_TAGS)
@pytest.mark.parametrize("K", [2, 10])
def test_collect_reference_vs_reference_fixture(golden, loaders, amp, name, K):
    """Rows written by the reference's own ``collect_reference_motions`` text (tests/golden/make_golden.py), on the
    40-frame windows and on the full shipped clips; the stated bar, unscaled."""
    loader = loaders(name)
    env = make_env(amp, loader, K)
    times, ids = golden[f"{name}/times"], golden[f"{name}/ids"]
    obs = env.collect_reference_motions(len(times), times, ids)
    assert obs.shape == (len(times), K * env.cfg.amp_observation_space)
    close(obs, golden[f"{name}/amp_obs_k{K}"])
    D = env.cfg.robot.num_joints  # everything but the six slerp-derived columns is bit-identical to the reference
    A = env.cfg.amp_observation_space
    lerp_cols = np.r_[0 : 2 * D + 1, 2 * D + 7 : A]
    bit_equal(obs.cpu().numpy().reshape(-1, A)[:, lerp_cols], golden[f"{name}/amp_obs_k{K}"].reshape(-1, A)[:, lerp_cols])
    # same thing with device-resident inputs
    obs2 = env.collect_reference_motions(len(times), torch.from_numpy(times).cuda(), torch.from_numpy(ids).cuda())
    assert torch.equal(obs, obs2)
    assert env.poll_flags() == 0


def test_reference_defaults_and_argument_handling(golden, loaders, amp):
    loader = loaders("G1_walk")
    times = golden["G1_walk/times"]
    # motion_ids=None with explicit times -> clip 0 (reference :365-366)
    a = loader.sample(len(times), times=times)
    b = loader.sample(len(times), times=times, motion_ids=np.zeros(len(times), dtype=np.int32))
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    # times=None -> numpy global RNG stream identical to the reference's sample_times
    np.random.seed(11)
    ids_ref, t_ref = loader.sample_times(33)
    np.random.seed(11)
    drawn = loader.sample(33)
    again = loader.sample(33, times=t_ref, motion_ids=ids_ref)
    for x, y in zip(drawn, again):
        assert torch.equal(x, y)
    assert loader.get_dof_index(["left_knee_joint"]) == [loader.dof_names.index("left_knee_joint")]
    with pytest.raises(AssertionError):
        loader.get_body_index(["no_such_body"])
    with pytest.raises(AssertionError):
        loader.get_dof_index(["no_such_dof"])


def test_pooled_ids_numpy_indexing_semantics(golden, loaders, amp):
    loader = loaders("pooled_humanoid")
    assert loader.num_trajectories == 3 and loader.traj_starts.tolist() == [0, 40, 80]
    t = np.array([0.1, 0.2, 0.3])
    with pytest.raises(IndexError):
        loader.sample(3, times=t, motion_ids=np.array([0, 3, 1]))
    with pytest.raises(IndexError):
        loader.sample(3, times=t, motion_ids=np.array([0, -4, 1]))
    neg = loader.sample(3, times=t, motion_ids=np.array([-1, -2, -3]))  # numpy wraps negative indices
    pos = loader.sample(3, times=t, motion_ids=np.array([2, 1, 0]))
    for x, y in zip(neg, pos):
        assert torch.equal(x, y)
    # device ids cannot raise synchronously: clamped + sticky flag
    loader.poll_flags()
    loader.sample(3, times=torch.tensor(t, device="cuda"), motion_ids=torch.tensor([0, 7, 1], device="cuda"))
    assert loader.poll_flags() & 1
    loader.sample(2, times=torch.tensor([float("nan"), 0.1], dtype=torch.float64, device="cuda"))
    assert loader.poll_flags() & 2
    assert loader.poll_flags() == 0


# ---------------------------------------------------------------------------------------------------------------------
# seeded comparisons against the oracle (larger, random + edge times)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["G1_walk", "G1_dance", "humanoid_dance", "G1_walk_lafan1", "pooled_humanoid"])
def test_sample_and_collect_vs_oracle_seeded(loaders, amp, name):
    from oracle import OracleMotionLoader, env_oracle

    loader = loaders(name)
    files = pooled_spec().split(",") if name == "pooled_humanoid" else [clip_path(name)]
    ora = OracleMotionLoader(files)
    rng = np.random.default_rng(1234)
    n = 6000
    ids = rng.integers(0, ora.num_trajectories, n)
    times = rng.uniform(-0.3, 1.3, n) * ora.durations[ids]
    times[:200] = (rng.integers(0, 39, 200) + 0.5) * ora.dt  # half-frame ties
    i0, i1, blend = loader._compute_frame_blend(times, ids)
    r0, r1, rb = ora.compute_frame_blend(times, ids)
    assert np.array_equal(i0, r0) and np.array_equal(i1, r1) and np.array_equal(blend.view(np.int64), rb.view(np.int64))
    for key, got, want in zip(OUT_NAMES, loader.sample(n, times=times, motion_ids=ids), ora.sample(n, times=times, motion_ids=ids)):
        if key == "body_rot":
            close(got, want)  # the stated bar, unscaled, extrapolated frames (blend outside [0, 1]) included
        else:
            bit_equal(got, want)
    robot = amp.robot_for_clip(loader.dof_names)
    for K in (1, 3, 10):
        env = make_env(amp, loader, K)
        want = env_oracle.collect_reference_motions(
            ora, n, K, ora.get_dof_index(robot.joint_names), ora.get_body_index([robot.reference_body])[0],
            ora.get_body_index(robot.key_body_names), current_times=times, motion_ids=ids,
        )  # fmt: skip
        _, _, frame_blend = ora.compute_frame_blend(env_oracle.history_times(times, ora.dt, K), np.repeat(ids, K))
        check_amp_obs(env.collect_reference_motions(n, times, ids), want, frame_blend, robot.num_joints, K)


def test_interpolate_and_slerp_helpers_vs_oracle(loaders):
    from oracle import lerp_f32, slerp_f32

    loader = loaders("G1_dance_old")
    g = torch.Generator().manual_seed(3)
    n = 513
    blend = torch.rand(n, generator=g) * 3 - 1
    a, b = torch.randn(n, 7, 3, generator=g), torch.randn(n, 7, 3, generator=g)
    got = loader._interpolate(a.cuda(), b=b.cuda(), blend=blend.cuda())
    assert torch.equal(got.cpu(), lerp_f32(a, b, blend))
    a2, b2 = torch.randn(n, 5, generator=g), torch.randn(n, 5, generator=g)
    assert torch.equal(loader._interpolate(a2.cuda(), b=b2.cuda(), blend=blend.cuda()).cpu(), lerp_f32(a2, b2, blend))
    q0 = torch.nn.functional.normalize(torch.randn(n, 9, 4, generator=g), dim=-1)
    q1 = torch.nn.functional.normalize(q0 + 0.3 * torch.randn(n, 9, 4, generator=g), dim=-1)
    q1[::7] = q0[::7]  # dot == 1 -> identity fallback (and NaN masking)
    q1[1::7] = -q1[1::7]  # shortest-arc flip
    q1[2::7] = torch.nn.functional.normalize(q0[2::7] + 1e-4 * torch.randn(q0[2::7].shape, generator=g), dim=-1)  # midpoint
    close(loader._slerp(q0.cuda(), q1=q1.cuda(), blend=blend.cuda()), slerp_f32(q0, q1, blend))
    # start/end form gathers from dimension 0
    idx0, idx1 = np.array([0, 5, 39]), np.array([1, 6, 39])
    w = torch.tensor([0.25, 0.5, 2.0])
    got = loader._slerp(loader.body_rotations, blend=w.cuda(), start=idx0, end=idx1)
    rot = loader.body_rotations.cpu()
    close(got, slerp_f32(rot[idx0], rot[idx1], w))


def test_free_functions_vs_oracle(amp):
    from oracle import env_oracle

    g = torch.Generator().manual_seed(5)
    n, D, Kb = 1000, 29, 4
    args = [torch.randn(n, D, generator=g), torch.randn(n, D, generator=g), torch.randn(n, 3, generator=g),
            torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=-1), torch.randn(n, 3, generator=g),
            torch.randn(n, 3, generator=g), torch.randn(n, Kb, 3, generator=g)]  # fmt: skip
    want = env_oracle.compute_obs(*args)
    got = amp.compute_obs(*[a.cuda() for a in args])
    assert got.shape == (n, 83)
    close(got, want)
    q = torch.randn(4, 6, 4, generator=g)  # not normalised on purpose; arbitrary leading dims
    close(amp.quaternion_to_tangent_and_normal(q.cuda()), env_oracle.quaternion_to_tangent_and_normal(q))
    # humanoid widths
    args28 = [torch.randn(7, 28, generator=g), torch.randn(7, 28, generator=g), *[a[:7] for a in args[2:]]]
    close(amp.compute_obs(*[a.cuda() for a in args28]), env_oracle.compute_obs(*args28))


@pytest.mark.parametrize("K", [1, 2, 10, 17])
@pytest.mark.parametrize("robot_name", ["g1", "humanoid28"])
def test_env_step_history_vs_oracle(loaders, amp, K, robot_name):
    from oracle import env_oracle
    from humanoid_amp_b200.synthetic import synthetic_sim_state

    loader = loaders("G1_dance" if robot_name == "g1" else "humanoid_walk")
    N = 300
    env = make_env(amp, loader, K, num_envs=N)
    robot = env.cfg.robot
    ref_buf = torch.zeros(N, K, robot.amp_observation_space)
    for step in range(K + 3):
        state = synthetic_sim_state(N, robot, "cpu", seed=100 + step)
        jp, jv, bp, bq, bl, ba = state
        obs = env_oracle.compute_obs(jp, jv, bp[:, env.ref_body_index], bq[:, env.ref_body_index], bl[:, env.ref_body_index],
                                     ba[:, env.ref_body_index], bp[:, env.key_body_indexes])  # fmt: skip
        want_view = env_oracle.shift_and_write_history(ref_buf, obs)
        policy = torch.full((N, robot.amp_observation_space - 12 + 5), -7.0, device="cuda")
        got_view = env.update_amp_observations(*[t.cuda() for t in state], policy_obs=policy)
        assert got_view.data_ptr() == env.amp_observation_buffer.data_ptr()  # a view, in place, as in the reference
        close(got_view, want_view)
        close(policy[:, : robot.amp_observation_space - 12], obs[:, :-12])
        assert (policy[:, robot.amp_observation_space - 12 :] == -7.0).all()
    assert env.extras["amp_obs"].shape == (N, K * robot.amp_observation_space)


@pytest.mark.parametrize("K", [2, 7])
def test_env_step_with_several_envs_per_warp(loaders, amp, K):
    """More envs than resident warps (the grid is capped at 8 CTAs per SM = 9472 warps): every warp walks several envs, two
    in flight per trip, and an odd count leaves unpaired tails -- the path the 300-env test above never enters."""
    from oracle import env_oracle
    from humanoid_amp_b200.synthetic import synthetic_sim_state

    loader = loaders("G1_dance")
    N = 30001
    env = make_env(amp, loader, K, num_envs=N)
    robot = env.cfg.robot
    ref_buf = torch.zeros(N, K, robot.amp_observation_space)
    for step in range(K + 1):
        state = synthetic_sim_state(N, robot, "cpu", seed=500 + step)
        jp, jv, bp, bq, bl, ba = state
        obs = env_oracle.compute_obs(jp, jv, bp[:, env.ref_body_index], bq[:, env.ref_body_index], bl[:, env.ref_body_index],
                                     ba[:, env.ref_body_index], bp[:, env.key_body_indexes])  # fmt: skip
        want_view = env_oracle.shift_and_write_history(ref_buf, obs)
        policy = torch.empty((N, robot.amp_observation_space - 12), device="cuda")
        got_view = env.update_amp_observations(*[t.cuda() for t in state], policy_obs=policy)
        close(got_view, want_view)
        close(policy, obs[:, :-12])


@pytest.mark.parametrize("n_actor,track,inc_act,inc_cmd", [(1, 0.0, True, True), (1, 1.0, True, True), (2, 1.0, True, True),
                                                         (4, 1.0, False, True), (3, 1.0, True, False), (5, 0.0, True, True)])
def test_actor_observation_history_vs_oracle(loaders, amp, n_actor, track, inc_act, inc_cmd, N=200):
    """the policy observation of ``_get_observations`` (g1_amp_env.py:195-242) incl. warm start."""
    from oracle import env_oracle
    from humanoid_amp_b200.synthetic import synthetic_sim_state

    loader = loaders("G1_dance")
    K = 3
    cfg = amp.AmpEnvCfg(motion_file="<preloaded>", num_envs=N, num_amp_observations=K, robot=amp.G1, num_actor_observations=n_actor,
                        rew_track_vel=track, history_include_last_actions=inc_act, history_include_command=inc_cmd)
    env = amp.AmpEnvPath(cfg, "cuda:0", motion_loader=loader)
    P = cfg.hist_frame_size
    ref_hist = torch.zeros(N, max(n_actor - 1, 1), P)
    ref_mask = torch.zeros(N, dtype=torch.bool)
    g = torch.Generator().manual_seed(n_actor * 7 + int(track))
    for step in range(n_actor + 3):
        state = synthetic_sim_state(N, amp.G1, "cpu", seed=300 + step)
        jp, jv, bp, bq, bl, ba = state
        actions = torch.randn(N, 29, generator=g)
        command = torch.rand(N, 2, generator=g) * 2 - 1
        if step in (0, 2):  # some envs were reset before this step
            picked = torch.rand(N, generator=g) < 0.3
            ref_mask |= picked
            if n_actor > 1:
                env._just_reset_mask |= picked.cuda()
        env.last_actions.copy_(actions)
        env.command_target_speed.copy_(command)
        obs = env_oracle.compute_obs(jp, jv, bp[:, env.ref_body_index], bq[:, env.ref_body_index], bl[:, env.ref_body_index],
                                     ba[:, env.ref_body_index], bp[:, env.key_body_indexes])  # fmt: skip
        want = env_oracle.actor_observations(obs, actions, command if track > 0 else None, ref_hist, ref_mask, n_actor,
                                             history_include_last_actions=inc_act, history_include_command=inc_cmd)  # fmt: skip
        got = env.get_observations(*[t.cuda() for t in state])["policy"]
        assert got.shape == want.shape == (N, cfg.observation_space)
        close(got, want)  # the six tangent/normal columns differ from torch by <= 1 ulp, everything else is a copy
        if n_actor > 1:
            close(env.actor_obs_history_buffer, ref_hist)
            assert not env._just_reset_mask.any() and not ref_mask.any()


def test_actor_observation_with_several_envs_per_warp(loaders, amp):
    """More envs than resident warps (grid capped at 9472 warps): the grid-stride loops of both observation kernels."""
    test_actor_observation_history_vs_oracle(loaders, amp, 3, 1.0, True, True, N=21001)


@pytest.mark.parametrize("track", [0.0, 1.0])
def test_task_reward_vs_oracle(loaders, amp, track, N=777):
    """``_get_rewards`` (g1_amp_env.py:246-288) with compute_rewards / exp_reward_with_floor."""
    from oracle import env_oracle
    from humanoid_amp_b200.synthetic import synthetic_sim_state

    loader = loaders("G1_dance")
    scales = dict(rew_termination=-1.0, rew_action_l2=-0.1, rew_joint_pos_limits=-10.0, rew_joint_acc_l2=-1.0e-06,
                  rew_joint_vel_l2=-0.001, rew_track_vel=track)  # the _CUSTOM cfg values (g1_amp_env_cfg.py:86-91)
    cfg = amp.AmpEnvCfg(motion_file="<preloaded>", num_envs=N, num_amp_observations=2, robot=amp.G1, **scales)
    env = amp.AmpEnvPath(cfg, "cuda:0", motion_loader=loader)
    g = torch.Generator().manual_seed(11)
    jp, jv, bp, bq, bl, ba = synthetic_sim_state(N, amp.G1, "cpu", seed=77)
    actions = torch.randn(N, 29, generator=g)
    acc = torch.randn(N, 29, generator=g) * 50
    lo = torch.rand(N, 29, generator=g) * -2.0
    limits = torch.stack([lo, lo + torch.rand(N, 29, generator=g) * 3.0], dim=-1)  # some joints outside their soft limits
    terminated = torch.rand(N, generator=g) < 0.2
    command = torch.rand(N, 2, generator=g) * 2 - 1
    bl = bl * 0.5
    bl[::5] *= 4  # a share of envs beyond the exp/linear threshold of the tracking reward
    env.command_target_speed.copy_(command)
    want_total, want_terms, want_err = env_oracle.task_rewards(scales, terminated, actions, jp, limits, acc, jv,
                                                               bl[:, env.ref_body_index], bq[:, env.ref_body_index], command)  # fmt: skip
    total, terms, err = env.get_rewards(terminated.cuda(), actions.cuda(), jp.cuda(), limits.cuda(), acc.cuda(), jv.cuda(),
                                        bl.cuda(), bq.cuda(), return_terms=True)  # fmt: skip
    # per-joint sums are accumulated in a different (fixed shuffle-tree) order than torch.sum: 29 fp32 addends
    close(terms, want_terms, rtol=2e-5, atol=1e-6)
    close(total, want_total, rtol=2e-5, atol=2e-6)
    if track > 0:
        close(err, want_err, rtol=1e-5, atol=1e-6)
        assert (want_err**2 > 1.0).any() and (want_err**2 < 1.0).any()  # both branches of exp_reward_with_floor exercised
    only_total = env.get_rewards(terminated.cuda(), actions.cuda(), jp.cuda(), limits.cuda(), acc.cuda(), jv.cuda(), bl.cuda(), bq.cuda())
    assert torch.equal(only_total, total)


def test_task_reward_with_several_envs_per_warp(loaders, amp):
    """More envs than resident warps (grid capped at 9472 warps): the grid-stride loop of the reward kernel."""
    test_task_reward_vs_oracle(loaders, amp, 1.0, N=25003)


def test_reset_strategy_random_state_vs_oracle(loaders, amp):
    """SURVEY 8f item 3 (reset-state write, g1_amp_env.py:371-419): same host RNG stream, same root / dof state, and the
    reset envs' AMP history rows."""
    from oracle import OracleMotionLoader, env_oracle

    loader = loaders("G1_walk")
    ora = OracleMotionLoader([clip_path("G1_walk")])
    N, K = 96, 10
    env = make_env(amp, loader, K, num_envs=N)
    robot = env.cfg.robot
    g = torch.Generator().manual_seed(2)
    default_root = torch.randn(N, 13, generator=g)
    origins = torch.randn(N, 3, generator=g) * 5
    env_ids = torch.arange(0, N, 3)
    np.random.seed(99)
    root, dof_p, dof_v, mids, times = env.reset_strategy_random(env_ids, default_root[env_ids].cuda(), origins[env_ids].cuda())
    np.random.seed(99)
    want_ids, want_times = ora.sample_times(len(env_ids))
    assert np.array_equal(mids, want_ids) and np.array_equal(times, want_times)
    w_root, w_dp, w_dv = env_oracle.reset_root_and_dof_state(ora, want_times, want_ids, default_root[env_ids], origins[env_ids],
                                                            ora.get_dof_index(robot.joint_names), ora.get_body_index(["pelvis"])[0])  # fmt: skip
    close(root, w_root)
    bit_equal(dof_p, w_dp)
    bit_equal(dof_v, w_dv)
    rows = env_oracle.collect_reference_motions(ora, len(env_ids), K, ora.get_dof_index(robot.joint_names), 0,
                                                ora.get_body_index(robot.key_body_names), current_times=want_times, motion_ids=want_ids)  # fmt: skip
    close(env.amp_observation_buffer[env_ids.cuda()].view(len(env_ids), -1), rows)
    untouched = torch.ones(N, dtype=torch.bool)
    untouched[env_ids] = False
    assert not env.amp_observation_buffer[untouched.cuda()].any()


def test_reset_fill_scatter_and_ring_memory(loaders, amp):
    from oracle import OracleMotionLoader, env_oracle

    loader = loaders("G1_walk")
    ora = OracleMotionLoader([clip_path("G1_walk")])
    K, N = 10, 128
    env = make_env(amp, loader, K, num_envs=N)
    robot = env.cfg.robot
    rng = np.random.default_rng(9)
    env.amp_observation_buffer.fill_(3.0)
    env_ids = torch.tensor(rng.permutation(N)[:37])
    times = rng.uniform(0, ora.durations[0], len(env_ids))
    ids = np.zeros(len(env_ids), dtype=np.int64)
    env.reset_amp_history(env_ids, times, ids)
    want = torch.full((N, K, 83), 3.0)
    rows = env_oracle.collect_reference_motions(ora, len(env_ids), K, ora.get_dof_index(robot.joint_names), 0,
                                                ora.get_body_index(robot.key_body_names), current_times=times, motion_ids=ids)  # fmt: skip
    env_oracle.reset_fill(want, env_ids, rows)
    close(env.amp_observation_buffer, want)
    # skrl RandomMemory-like ring (memory_size, 1, K*A): write 50 rows starting at 40 into a 64-row ring -> wraps
    mem = torch.zeros(64, 1, K * 83, device="cuda")
    t50 = rng.uniform(0, ora.durations[0], 50)
    nxt = env.collect_reference_motions_into(mem, 40, 50, t50, np.zeros(50, dtype=np.int64))
    assert nxt == (40 + 50) % 64
    rows50 = env_oracle.collect_reference_motions(ora, 50, K, ora.get_dof_index(robot.joint_names), 0,
                                                  ora.get_body_index(robot.key_body_names), current_times=t50,
                                                  motion_ids=np.zeros(50, dtype=np.int64))  # fmt: skip
    want_mem = torch.zeros(64, K * 83)
    want_mem[(40 + np.arange(50)) % 64] = rows50
    close(mem.view(64, -1), want_mem)


def test_empty_and_maximum_history(loaders, amp):
    loader = loaders("humanoid_run")
    env = make_env(amp, loader, 64)  # K = 64 is the maximum the fused kernel accepts
    out = env.collect_reference_motions(0, np.zeros(0), np.zeros(0, dtype=np.int64))
    assert out.shape == (0, 64 * 81)
    obs = env.collect_reference_motions(5, np.linspace(0, 0.6, 5), np.zeros(5, dtype=np.int64))
    assert obs.shape == (5, 64 * 81) and torch.isfinite(obs).all()
    for t in loader.sample(0, times=np.zeros(0)):
        assert t.shape[0] == 0
    cfg = amp.AmpEnvCfg(motion_file="x", num_envs=4, num_amp_observations=65, robot=amp.HUMANOID28)
    too_long = amp.AmpEnvPath(cfg, "cuda:0", motion_loader=loader)
    with pytest.raises(amp.AmpB200Error):
        too_long.collect_reference_motions(3, np.zeros(3))


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE sizes: size-independent properties
# ---------------------------------------------------------------------------------------------------------------------
def test_one_million_sample_refill_properties(tmp_path, amp):
    """configs[3]: 1M samples x 2 history on the G1_walk shape.  Checked through properties that do not need the oracle at
    full size, plus an oracle spot check on a random subset of rows."""
    from oracle import OracleMotionLoader, env_oracle
    from humanoid_amp_b200.synthetic import write_synthetic_clip

    path = write_synthetic_clip(str(tmp_path / "g1_walk_syn.npz"), "G1_walk", seed=4)
    loader = amp.MotionLoader(path, "cuda:0")
    env = make_env(amp, loader, 2)
    n = 1_000_000
    g = torch.Generator(device="cuda").manual_seed(77)
    ids, times = loader.sample_times_device(n, generator=g)
    a = env.collect_reference_motions(n, times, ids)
    b = env.collect_reference_motions(n, times, ids)
    assert a.shape == (n, 166) and torch.equal(a, b), "idempotent / deterministic"
    assert torch.isfinite(a).all()
    # history structure: slot k of sample i == a K=1 collect at time t_i - k*dt (float64 subtraction on the host)
    env1 = make_env(amp, loader, 1)
    t_host = times.cpu().numpy()
    older = env1.collect_reference_motions(n, t_host - loader.dt * 1, ids)
    assert torch.equal(a[:, 83:], older) and torch.equal(a[:, :83], env1.collect_reference_motions(n, times, ids))
    # fused path == unfused path (sample_full kernel -> gathers -> compute_obs kernel) on a 200k slice, bit for bit
    m = 200_000
    dp, dv, bp, br, bl, ba = loader.sample(m, times=times[:m], motion_ids=ids[:m])
    r = env.motion_ref_body_index
    unfused = amp.compute_obs(dp[:, env.motion_dof_indexes], dv[:, env.motion_dof_indexes], bp[:, r], br[:, r], bl[:, r],
                              ba[:, r], bp[:, env.motion_key_body_indexes])  # fmt: skip
    assert torch.equal(a[:m, :83], unfused)
    # oracle spot check
    pick = np.random.default_rng(0).choice(n, 4096, replace=False)
    ora = OracleMotionLoader([path])
    robot = env.cfg.robot
    want = env_oracle.collect_reference_motions(ora, len(pick), 2, ora.get_dof_index(robot.joint_names), 0,
                                                ora.get_body_index(robot.key_body_names), current_times=t_host[pick],
                                                motion_ids=ids.cpu().numpy()[pick])  # fmt: skip
    close(a[torch.from_numpy(pick).cuda()], want)
    assert env.poll_flags() == 0


# ---------------------------------------------------------------------------------------------------------------------
# every variant of the fused collect kernel: table in global / shared memory x history length x destination form,
# on the FULL shipped clips (G1_walk 134 KB, G1_dance 202 KB packed tables; the pooled humanoid table, 382 KB, does not
# fit in shared memory and must fall back to the global-table variant)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dest", ["contiguous", "ring", "scatter"])
@pytest.mark.parametrize("K", [1, 2, 3, 10])
@pytest.mark.parametrize("table", ["global", "smem"])
@pytest.mark.parametrize("name", ["full/G1_walk", "full/G1_dance", "full/pooled_humanoid"])
def test_collect_variants_vs_oracle(loaders, amp, name, table, K, dest):
    from oracle import OracleMotionLoader, env_oracle

    loader = loaders(name)
    ora = OracleMotionLoader(fixture_files(name))
    env = make_env(amp, loader, K)
    env._handle.set_collect_table(table)
    robot = env.cfg.robot
    A = robot.amp_observation_space
    rng = np.random.default_rng(hash((name, K)) % 2**31)
    n = 3001
    ids = rng.integers(0, ora.num_trajectories, n)
    times = rng.uniform(-0.05, 1.05, n) * ora.durations[ids]
    times[:64] = (rng.integers(0, 30, 64) + 0.5) * ora.dt  # half-frame ties
    want = env_oracle.collect_reference_motions(
        ora, n, K, ora.get_dof_index(robot.joint_names), ora.get_body_index([robot.reference_body])[0],
        ora.get_body_index(robot.key_body_names), current_times=times, motion_ids=ids,
    )  # fmt: skip
    _, _, frame_blend = ora.compute_frame_blend(env_oracle.history_times(times, ora.dt, K), np.repeat(ids, K))
    t_d, i_d = torch.from_numpy(times).cuda(), torch.from_numpy(ids).cuda()
    if dest == "contiguous":
        got = env.collect_reference_motions(n, t_d, i_d)
    elif dest == "ring":  # a skrl RandomMemory-like ring with padded rows, written from row 2900 on: wraps at 3500
        mem = torch.full((3500, 1, K * A + 5), -3.0, device="cuda")
        rows = mem.view(3500, -1)
        env._launch_collect(t_d, i_d, n, rows, None, start_row=2900, capacity_rows=3500)
        order = (2900 + np.arange(n)) % 3500
        got = rows[torch.from_numpy(order).cuda(), : K * A]
        assert (rows[:, K * A :] == -3.0).all(), "padding columns of the destination rows must not be touched"
        untouched = np.setdiff1d(np.arange(3500), order)
        assert (rows[torch.from_numpy(untouched).cuda()] == -3.0).all()
    else:  # scatter into amp_observation_buffer[env_ids] (reset fill)
        big = make_env(amp, loader, K, num_envs=2 * n)
        big._handle.set_collect_table(table)
        big.amp_observation_buffer.fill_(-3.0)
        env_ids = torch.from_numpy(rng.permutation(2 * n)[:n]).cuda()
        big.reset_amp_history(env_ids, t_d, i_d)
        got = big.amp_observation_buffer[env_ids].view(n, -1)
        mask = torch.ones(2 * n, dtype=torch.bool, device="cuda")
        mask[env_ids] = False
        assert (big.amp_observation_buffer[mask] == -3.0).all()
    check_amp_obs(got, want, frame_blend, robot.num_joints, K)
    inside = np.abs(frame_blend.reshape(n, K)).max(axis=1) <= 1.0  # samples whose every history frame interpolates:
    close(got[torch.from_numpy(inside).cuda()], want[torch.from_numpy(inside)])  # the stated bar, unscaled
    assert env.poll_flags() == 0


def test_collect_smem_table_large_batch_matches_global_table(loaders, amp):
    """The shared-memory-table variant at the size that selects it by itself (>= 2 full tiles per resident warp), K = 10 on
    the 202 KB G1_dance table -- what a 200 000 x 10 motion_dataset fill (skrl_g1_dance_amp_cfg.yaml:50-58) runs -- must equal
    the global-table variant bit for bit (same arithmetic, different staging) and the oracle on a subset of the rows."""
    from oracle import OracleMotionLoader, env_oracle

    name, K, n = "full/G1_dance", 10, 200_000
    loader = loaders(name)
    env = make_env(amp, loader, K)
    g = torch.Generator(device="cuda").manual_seed(5)
    ids, times = loader.sample_times_device(n, generator=g)
    outs = {}
    for table in ("auto", "global", "smem"):
        env._handle.set_collect_table(table)
        outs[table] = env.collect_reference_motions(n, times, ids).clone()
    assert torch.equal(outs["global"], outs["smem"]) and torch.equal(outs["auto"], outs["smem"])
    ora = OracleMotionLoader(fixture_files(name))
    pick = np.random.default_rng(1).choice(n, 2048, replace=False)
    robot = env.cfg.robot
    want = env_oracle.collect_reference_motions(ora, len(pick), K, ora.get_dof_index(robot.joint_names), 0, ora.get_body_index(robot.key_body_names),
                                                current_times=times.cpu().numpy()[pick], motion_ids=ids.cpu().numpy()[pick])  # fmt: skip
    _, _, fb = ora.compute_frame_blend(env_oracle.history_times(times.cpu().numpy()[pick], ora.dt, K), np.repeat(ids.cpu().numpy()[pick], K))
    check_amp_obs(outs["smem"][torch.from_numpy(pick).cuda()], want, fb, robot.num_joints, K)


def test_extrapolated_slerp_spread_is_torch_cpu_vs_torch_cuda_spread(loaders):
    """Justifies the conditioning term of ``check_amp_obs`` (``atol = 1e-6 * max(1, |blend|)`` on the tangent/normal columns
    of EXTRAPOLATED history frames): the reference algorithm itself (oracle restatement, bit-pinned to the reference), run by
    torch on the CPU (SLEEF ``acos`` / ``sin``) and by torch on this GPU (libdevice), is compared on the same extrapolated
    inputs, slerp -> quaternion_to_tangent_and_normal.  Asserted: (a) this kernel's body rotations meet the UNSCALED bar;
    (b) its tangent/normal stay within the scaled bar; (c) it is no further from torch-CPU than twice what torch-CUDA is
    -- i.e. the spread the scaled term allows is the spread two builds of the reference show between themselves."""
    from oracle import env_oracle, slerp_f32

    loader = loaders("full/G1_dance")
    rot = loader.body_rotations.cpu()
    rng = np.random.default_rng(3)
    n = 20000
    f0 = rng.integers(0, rot.shape[0] - 1, n)
    blend = torch.from_numpy(rng.uniform(-12.0, -1.0, n).astype(np.float32))
    q0, q1 = rot[f0, 0], rot[f0 + 1, 0]  # the root body
    ref_cpu = slerp_f32(q0, q1, blend)
    ref_cuda = slerp_f32(q0.cuda(), q1.cuda(), blend.cuda())
    ours = loader._slerp(q0.cuda().unsqueeze(1), q1=q1.cuda().unsqueeze(1), blend=blend.cuda()).squeeze(1)
    tn_cpu = env_oracle.quaternion_to_tangent_and_normal(ref_cpu)
    tn_cuda = env_oracle.quaternion_to_tangent_and_normal(ref_cuda).cpu()
    import humanoid_amp_b200 as amp

    tn_ours = amp.quaternion_to_tangent_and_normal(ours).cpu()
    ours, ref_cuda = ours.cpu(), ref_cuda.cpu()

    def ratio(a, b):  # max error in units of the UNSCALED bound
        return float(((a - b).abs() / (ATOL + RTOL * b.abs())).max())

    r = dict(q_torch=ratio(ref_cuda, ref_cpu), q_ours=ratio(ours, ref_cpu), tn_torch=ratio(tn_cuda, tn_cpu), tn_ours=ratio(tn_ours, tn_cpu))
    print("extrapolated frames, |blend| in [1, 12], max err / unscaled bound: " + ", ".join(f"{k} {v:.2f}" for k, v in r.items()))
    assert r["q_ours"] <= 1.0, r
    scaled = (ATOL * blend.abs().clamp(min=1.0))[:, None] + RTOL * tn_cpu.abs()
    assert bool(((tn_ours - tn_cpu).abs() <= scaled).all()), r
    assert r["tn_ours"] <= max(1.0, 2.0 * r["tn_torch"]), r
