"""Pin the oracle: it must reproduce, BIT FOR BIT, what the live reference produced for the committed fixtures.

The fixtures (``tests/golden``) were written by ``tests/golden/make_golden.py`` from the reference ``MotionLoader``
imported by path.  When ``/root/reference`` is present (build container) the oracle is additionally compared with the
live reference on the full shipped clips.
"""

from __future__ import annotations

import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import CLIP_NAMES, FIXTURE_TAGS, REFERENCE_ROOT, clip_path, fixture_files, pooled_spec
from oracle import OracleMotionLoader, env_oracle, ref_harness
from humanoid_amp_b200.robots import G1, HUMANOID28, robot_for_clip

OUT_NAMES = ("dof_pos", "dof_vel", "body_pos", "body_rot", "body_lin", "body_ang")


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", FIXTURE_TAGS)
def test_oracle_matches_reference_fixture_bitwise(golden, name):
    """Windows AND the full shipped clips; the ``amp_obs_k*`` rows were produced by the reference's own
    ``collect_reference_motions`` text (``oracle/_ref`` via ``oracle/ref_harness.py``), not by ``env_oracle``."""
    loader = OracleMotionLoader(fixture_files(name))
    times, ids = golden[f"{name}/times"], golden[f"{name}/ids"]
    i0, i1, blend = loader.compute_frame_blend(times, ids)
    assert _same(i0, golden[f"{name}/idx0"]) and _same(i1, golden[f"{name}/idx1"])
    assert _same(blend, golden[f"{name}/blend"])
    outs = loader.sample(len(times), times=times, motion_ids=ids)
    for key, t in zip(OUT_NAMES, outs):
        assert _same(t.numpy(), golden[f"{name}/{key}"]), key

    robot = robot_for_clip(loader.dof_names)
    for k in (2, 10):
        obs = env_oracle.collect_reference_motions(
            loader, len(times), k,
            loader.get_dof_index(robot.joint_names), loader.get_body_index([robot.reference_body])[0],
            loader.get_body_index(robot.key_body_names), current_times=times, motion_ids=ids,
        )  # fmt: skip
        assert _same(obs.numpy(), golden[f"{name}/amp_obs_k{k}"])


def test_pooled_sample_times_stream(golden):
    loader = OracleMotionLoader(pooled_spec().split(","))
    np.random.seed(123)
    ids, times = loader.sample_times(16)
    assert _same(ids, golden["pooled_humanoid/seed123_ids"])
    assert _same(times, golden["pooled_humanoid/seed123_times"])
    np.random.seed(5)
    ids, times = loader.sample_times(7, start=True)
    assert times.dtype == np.float64 and not times.any() and ids.shape == (7,)


def test_survey_known_answers_index_math(kat):
    """SURVEY.md 8c pins: frame indices / blends for G1_walk (F=399, 60 fps) need only the clip's frame count."""
    from oracle import frame_blend_f64

    g = kat["g1_walk"]
    dt = 1.0 / np.int64(60)
    assert float(dt).hex() == g["dt_hex"] == "0x1.1111111111111p-6"
    durations = np.array([dt * (g["num_frames"] - 1)])
    assert float(durations[0]).hex() == g["duration_hex"]
    cur = np.array(g["current_times"])
    t = env_oracle.history_times(cur, dt, 2)
    i0, i1, b = frame_blend_f64(t, np.zeros(len(t), dtype=np.int32), durations, np.array([0]), np.array([g["num_frames"] - 1]), dt)
    assert i0.tolist() == g["idx0"] == [0, 0, 0, 0, 1, 0, 60, 59, 200, 199, 398, 397, 398, 397, 398, 398]
    assert i1.tolist() == g["idx1"] == [1, 1, 1, 1, 2, 1, 61, 60, 201, 200, 398, 398, 398, 398, 398, 398]
    assert b.tolist() == g["blend"] == [0, -1, 0.24, -0.76, -0.25, -0.25, 0, 0, 0, 0, -0.2, -0.2, 0, 0, 22, 21]
    assert g["dof_perm"] == [0, 6, 12, 1, 7, 13, 2, 8, 14, 3, 9, 15, 22, 4, 10, 16, 23, 5, 11, 17, 24, 18, 25, 19, 26, 20, 27, 21, 28]  # fmt: skip


def test_env_oracle_known_answers_shapes():
    q = torch.tensor([[1.0, 0, 0, 0], [0.70710678, 0, 0, 0.70710678]])
    tn = env_oracle.quaternion_to_tangent_and_normal(q)
    assert torch.allclose(tn[0], torch.tensor([1.0, 0, 0, 0, 0, 1]))
    assert torch.allclose(tn[1], torch.tensor([0.0, 1, 0, 0, 0, 1]), atol=1e-6)  # yaw 90 deg: x-hat -> y-hat
    buf = torch.zeros(3, 4, 5)
    for step in range(1, 7):
        view = env_oracle.shift_and_write_history(buf, torch.full((3, 5), float(step)))
    assert view.shape == (3, 20) and buf[0, :, 0].tolist() == [6.0, 5.0, 4.0, 3.0]  # slot 0 = newest


needs_reference = pytest.mark.skipif(
    not os.path.exists(f"{REFERENCE_ROOT}/motions/motion_loader.py"), reason="live reference only exists in the build container"
)


@needs_reference
@pytest.mark.parametrize("name", ["G1_walk", "G1_dance", "humanoid_walk", "G1_walk_lafan1"])
def test_oracle_vs_live_reference_full_clip(name, kat):
    spec = importlib.util.spec_from_file_location("ref_motion_loader", f"{REFERENCE_ROOT}/motions/motion_loader.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    path = f"{REFERENCE_ROOT}/motions/{name}.npz"
    ref, ora = mod.MotionLoader(path, "cpu"), OracleMotionLoader([path])
    rng = np.random.default_rng(7)
    times = np.concatenate([rng.uniform(-0.2, ref.durations[0] + 0.2, 4000), (np.arange(300) + 0.5) * ref.dt])
    ids = np.zeros(len(times), dtype=np.int64)
    for a, b in zip(ref._compute_frame_blend(times, ids), ora.compute_frame_blend(times, ids)):
        assert _same(a, b)
    for a, b in zip(ref.sample(len(times), times=times, motion_ids=ids), ora.sample(len(times), times=times, motion_ids=ids)):
        assert _same(a.numpy(), b.numpy())
    if name == "G1_walk":
        g = kat["g1_walk"]
        obs = env_oracle.collect_reference_motions(
            ora, 8, 2, ora.get_dof_index(G1.joint_names), ora.get_body_index(["pelvis"])[0],
            ora.get_body_index(G1.key_body_names), current_times=np.array(g["current_times"]),
            motion_ids=np.zeros(8, dtype=np.int64),
        )  # fmt: skip
        assert np.allclose(obs.double().sum(dim=1).numpy(), g["obs_row_sums"], rtol=0, atol=1e-9)
        assert np.allclose(obs[3, 58:71].numpy(), g["obs3_58_71"], rtol=0, atol=0)
        # values quoted (rounded) in SURVEY.md 8c
        assert np.allclose(g["obs_row_sums"][:3], [2.52549723, 2.555956552, 2.592755598], atol=1e-8)
        assert HUMANOID28.reference_body == "torso"


def test_upstream_quaternion_helpers_against_an_independent_library():
    """``quat_apply`` / ``quat_rotate_inverse`` are upstream Isaac Lab functions that are not vendored with the reference
    (parity unpinned).  As a known-answer check from an INDEPENDENT implementation, the restatements must agree with
    scipy's ``Rotation`` (float64) to fp32 rounding on random unit quaternions -- this pins the convention (wxyz, active
    rotation for ``quat_apply``, the inverse rotation for ``quat_rotate_inverse``), which is what a restatement can get wrong."""
    import torch
    from scipy.spatial.transform import Rotation

    from oracle import env_oracle

    rng = np.random.default_rng(7)
    q = rng.normal(size=(4096, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    v = rng.uniform(-3, 3, size=(4096, 3))
    rot = Rotation.from_quat(q[:, [1, 2, 3, 0]])  # scipy is xyzw
    qt, vt = torch.from_numpy(q.astype(np.float32)), torch.from_numpy(v.astype(np.float32))
    fwd = env_oracle.quat_apply(qt, vt).double().numpy()
    inv = env_oracle.quat_rotate_inverse(qt, vt).double().numpy()
    assert np.abs(fwd - rot.apply(v)).max() < 5e-6
    assert np.abs(inv - rot.inv().apply(v)).max() < 5e-6
    # tangent / normal = first and third column of the rotation matrix (g1_amp_env.py:489-497)
    tn = env_oracle.quaternion_to_tangent_and_normal(qt).double().numpy()
    R = rot.as_matrix()
    assert np.abs(tn[:, :3] - R[:, :, 0]).max() < 5e-6 and np.abs(tn[:, 3:] - R[:, :, 2]).max() < 5e-6


# ---------------------------------------------------------------------------------------------------------------------
# env rows: env_oracle must reproduce, BIT FOR BIT, what the reference's own method text (oracle/_ref, cut by
# oracle/build_ref.py from g1_amp_env.py / humanoid_amp_env.py) made of the stored simulator states
# ---------------------------------------------------------------------------------------------------------------------
SIM_KEYS = ("joint_pos", "joint_vel", "body_pos_w", "body_quat_w", "body_lin_vel_w", "body_ang_vel_w")
G1_OBS_CASES = [  # (tag, K, num_actor_observations, rew_track_vel, history_include_last_actions, history_include_command)
    ("k2_a1", 2, 1, 0.0, True, True), ("k10_a1", 10, 1, 0.0, True, True), ("k1_a1_cmd", 1, 1, 1.0, True, True),
    ("k3_a3_cmd", 3, 3, 1.0, True, True), ("k2_a4_noact", 2, 4, 1.0, False, True), ("k2_a3_nocmd", 2, 3, 1.0, True, False),
    ("k2_a5", 2, 5, 0.0, True, True),
]  # fmt: skip


def _sim_state(golden, robot_name, step):
    return [torch.from_numpy(golden[f"env/{robot_name}/{k}"][step]) for k in SIM_KEYS]


def _oracle_obs(robot, state):
    jp, jv, bp, bq, bl, ba = state
    ref = robot.body_names.index(robot.reference_body)
    keys = [robot.body_names.index(n) for n in robot.key_body_names]
    return env_oracle.compute_obs(jp, jv, bp[:, ref], bq[:, ref], bl[:, ref], ba[:, ref], bp[:, keys])


@pytest.mark.parametrize("case", G1_OBS_CASES, ids=[c[0] for c in G1_OBS_CASES])
def test_env_oracle_get_observations_matches_reference_text(golden, case):
    """``G1AmpEnv._get_observations`` (g1_amp_env.py:175-242): AMP history view and the actor observation."""
    tag, K, n_actor, track, inc_act, inc_cmd = case
    steps_kept = golden[f"env/g1/{tag}/steps"].tolist()
    n_steps, N = golden["env/g1/joint_pos"].shape[:2]
    buf = torch.zeros(N, K, G1.amp_observation_space)
    P = (G1.amp_observation_space - 12) + (29 if inc_act else 0) + (2 if (inc_cmd and track > 0) else 0)
    hist = torch.zeros(N, max(n_actor - 1, 1), P)
    mask = torch.zeros(N, dtype=torch.bool)
    for s in range(n_steps):
        obs = _oracle_obs(G1, _sim_state(golden, "g1", s))
        view = env_oracle.shift_and_write_history(buf, obs)
        actions = torch.from_numpy(golden["env/g1/last_actions"][s])
        command = torch.from_numpy(golden["env/g1/command"][s])
        if n_actor > 1:
            mask |= torch.from_numpy(golden["env/g1/reset_mask"][s])
        policy = env_oracle.actor_observations(obs, actions, command if track > 0 else None, hist, mask, n_actor,
                                               history_include_last_actions=inc_act, history_include_command=inc_cmd)  # fmt: skip
        if s in steps_kept:
            j = steps_kept.index(s)
            assert _same(view.numpy(), golden[f"env/g1/{tag}/amp_obs"][j]), (tag, s)
            assert _same(policy.numpy(), golden[f"env/g1/{tag}/policy"][j]), (tag, s)


@pytest.mark.parametrize("K", [2, 10])
def test_env_oracle_humanoid_get_observations_matches_reference_text(golden, K):
    """``HumanoidAmpEnv._get_observations`` (humanoid_amp_env.py:105-126)."""
    n_steps, N = golden["env/humanoid28/joint_pos"].shape[:2]
    buf = torch.zeros(N, K, HUMANOID28.amp_observation_space)
    kept = list(range(n_steps)) if K == 2 else [2, n_steps - 1]
    for s in range(n_steps):
        obs = _oracle_obs(HUMANOID28, _sim_state(golden, "humanoid28", s))
        view = env_oracle.shift_and_write_history(buf, obs)
        if s in kept:
            j = kept.index(s)
            assert _same(view.numpy(), golden[f"env/humanoid28/k{K}/amp_obs"][j])
            assert _same(obs.numpy(), golden[f"env/humanoid28/k{K}/policy"][j])


REWARD_SCALES = dict(rew_termination=-1.0, rew_action_l2=-0.1, rew_joint_pos_limits=-10.0, rew_joint_acc_l2=-1.0e-06, rew_joint_vel_l2=-0.001)


@pytest.mark.parametrize("track", [0.0, 1.0])
def test_env_oracle_task_reward_matches_reference_text(golden, track):
    """``G1AmpEnv._get_rewards`` + ``compute_rewards`` + ``exp_reward_with_floor`` (g1_amp_env.py:246-319, 500-532, 564-606)."""
    jp, jv, bp, bq, _, ba = _sim_state(golden, "g1", 0)
    bl = torch.from_numpy(golden["env/g1/reward/body_lin_vel_w"])
    ref = G1.body_names.index("pelvis")
    scales = dict(REWARD_SCALES, rew_track_vel=track)
    total, terms, err = env_oracle.task_rewards(
        scales, torch.from_numpy(golden["env/g1/reward/terminated"]), torch.from_numpy(golden["env/g1/last_actions"][0]), jp,
        torch.from_numpy(golden["env/g1/reward/soft_limits"]), torch.from_numpy(golden["env/g1/reward/joint_acc"]), jv,
        bl[:, ref], bq[:, ref], torch.from_numpy(golden["env/g1/command"][0]),
    )  # fmt: skip
    t = f"env/g1/reward/track{int(track)}"
    assert _same(total.numpy(), golden[f"{t}/total"])
    log = dict(zip(golden[f"{t}/log_keys"].tolist(), golden[f"{t}/log_values"].tolist()))
    assert log["total_reward"] == total.mean().item()
    for i, key in enumerate(("pub_termination", "pub_action_l2", "pub_joint_pos_limits", "pub_joint_acc_l2", "pub_joint_vel_l2")):
        assert log[key] == terms[:, i].contiguous().mean().item(), key
    if track > 0:
        assert log["rew_track_vel"] == terms[:, 5].contiguous().mean().item() and log["error_track_vel"] == err.mean().item()
        assert (err**2 > 1.0).any() and (err**2 < 1.0).any()  # both branches of exp_reward_with_floor


@pytest.mark.parametrize("K", [2, 10])
def test_env_oracle_reset_strategy_random_matches_reference_text(golden, K):
    """``G1AmpEnv._reset_strategy_random`` (g1_amp_env.py:371-441): host RNG stream, root / dof state, history fill."""
    r = f"env/g1/reset_k{K}"
    ora = OracleMotionLoader([clip_path("G1_dance")])
    env_ids = torch.from_numpy(golden[f"{r}/env_ids"])
    np.random.seed(99)
    ids, times = ora.sample_times(len(env_ids))
    assert _same(ids, golden[f"{r}/motion_ids"][env_ids.numpy()])
    assert _same(times.astype(np.float32), golden[f"{r}/motion_start_times"][env_ids.numpy()])
    root, dp, dv = env_oracle.reset_root_and_dof_state(
        ora, times, ids, torch.from_numpy(golden[f"{r}/default_root_state"])[env_ids], torch.from_numpy(golden[f"{r}/env_origins"])[env_ids],
        ora.get_dof_index(G1.joint_names), ora.get_body_index(["pelvis"])[0],
    )  # fmt: skip
    assert _same(root.numpy(), golden[f"{r}/root_state"]) and _same(dp.numpy(), golden[f"{r}/dof_pos"]) and _same(dv.numpy(), golden[f"{r}/dof_vel"])
    rows = env_oracle.collect_reference_motions(ora, len(env_ids), K, ora.get_dof_index(G1.joint_names), ora.get_body_index(["pelvis"])[0],
                                                ora.get_body_index(G1.key_body_names), current_times=times, motion_ids=ids)  # fmt: skip
    buf = torch.full((golden[f"{r}/amp_observation_buffer"].shape[0], K, G1.amp_observation_space), 3.0)
    env_oracle.reset_fill(buf, env_ids, rows)
    assert _same(buf.numpy(), golden[f"{r}/amp_observation_buffer"])


needs_ref_text = pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref (installed by oracle/build_ref.py) is not present")


@needs_ref_text
def test_reference_text_is_what_the_manifest_says():
    """``oracle/_ref`` holds the unmodified reference: in the build container its hashes are re-checked against the source."""
    import hashlib
    import json

    with open(os.path.join(ref_harness.REF_DIR, "MANIFEST.json")) as f:
        man = json.load(f)
    with open(os.path.join(ref_harness.REF_DIR, "motion_loader.py"), "rb") as f:
        assert hashlib.sha256(f.read()).hexdigest() == man["files"]["motions/motion_loader.py"]["sha256"]
    if os.path.exists(f"{REFERENCE_ROOT}/g1_amp_env.py"):
        for rel, rec in man["files"].items():
            with open(f"{REFERENCE_ROOT}/{rel}", "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == rec["sha256"], rel


@needs_ref_text
@pytest.mark.parametrize("name", ["G1_walk", "G1_dance", "humanoid_dance"])
def test_env_oracle_vs_reference_text_live_seeded(name):
    """Larger seeded comparison of ``env_oracle.collect_reference_motions`` with the reference's method text, full clips."""
    from conftest import full_clip_path

    path = full_clip_path(name)
    ref = ref_harness.reference_motion_loader_class()(path, "cpu")
    ora = OracleMotionLoader([path])
    robot = robot_for_clip(ora.dof_names)
    rng = np.random.default_rng(21)
    times = rng.uniform(-0.2, ref.durations[0] + 0.2, 3000)
    ids = np.zeros(len(times), dtype=np.int64)
    for K in (1, 2, 10):
        want = ref_harness.reference_collect(ref, robot, K, times, ids)
        got = env_oracle.collect_reference_motions(ora, len(times), K, ora.get_dof_index(robot.joint_names),
                                                   ora.get_body_index([robot.reference_body])[0], ora.get_body_index(robot.key_body_names),
                                                   current_times=times, motion_ids=ids)  # fmt: skip
        assert _same(got.numpy(), want.numpy())
