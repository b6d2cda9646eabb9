#!/usr/bin/env python
"""Generate the committed golden fixtures by RUNNING THE REFERENCE (build container only; needs ``/root/reference``).

    python tests/golden/make_golden.py

What executes here is the reference's own text, installed under the git-ignored ``oracle/_ref/`` by the committed recipe
``oracle/build_ref.py``: the unmodified ``motions/motion_loader.py`` and the ``ast``-cut bodies of ``G1AmpEnv`` /
``HumanoidAmpEnv`` methods and scripted free functions (``g1_amp_env.py:175-319, 371-606``), driven through
``oracle/ref_harness.py`` (a bare object carrying the attributes Isaac Lab's ``DirectRLEnv`` would provide).  The only
non-reference code on that path is ``quat_apply`` / ``quat_rotate_inverse`` (upstream Isaac Lab, not vendored: restated
in ``oracle/env_oracle.py`` and checked against scipy).  ``oracle/env_oracle.py`` itself is NOT used to produce any
fixture; ``tests/test_oracle_pins.py`` asserts that it reproduces them bit for bit.

The fixtures hold

* ``clips/<name>.npz``       -- a 40-frame window of each shipped motion clip in the reference's on-disk format, chosen
                                around the frame with the most negative consecutive-frame quaternion dot so the
                                shortest-arc flip and both slerp fall-backs are exercised;
* ``clips_full/<name>.npz``  -- the eight shipped clips themselves (data files, 5.4 MB): BASELINE's real tables
                                (134 / 202 / 382 KB packed) under test on the GPU box;
* ``vectors.npz``            -- for every window, every full clip and a pooled 3-clip humanoid loader: seeded + edge-case
                                ``times`` / ``motion_ids`` and the reference's ``_compute_frame_blend``, ``sample`` and
                                ``collect_reference_motions`` (K = 2, 10) outputs; ``env/*``: simulator-state sequences
                                and what ``_get_observations`` (AMP history + actor observation), ``_get_rewards`` and
                                ``_reset_strategy_random`` made of them;
* ``kat.json``               -- known answers on the FULL shipped clips recorded in SURVEY.md section 8c, re-derived here.

``/root/reference`` does not exist on the GPU box, so tests only ever read these files.
"""

from __future__ import annotations

import json
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from humanoid_amp_b200.robots import G1, HUMANOID28  # noqa: E402  (simulator joint/body order table)
from humanoid_amp_b200.synthetic import synthetic_sim_state  # noqa: E402
from oracle import build_ref, ref_harness  # noqa: E402

WINDOW = 40
CLIPS = [
    "G1_walk",
    "G1_dance",
    "G1_dance_old",
    "G1_walk_lafan1",
    "custom_motion",
    "humanoid_walk",
    "humanoid_run",
    "humanoid_dance",
]


def pick_window(rot: np.ndarray, width: int) -> int:
    """Start frame of the window centred on the most negative consecutive-frame quaternion dot (any body)."""
    dots = np.sum(rot[:-1] * rot[1:], axis=-1).min(axis=1)
    centre = int(np.argmin(dots))
    return int(np.clip(centre - width // 2, 0, max(rot.shape[0] - width, 0)))


def edge_times(duration: float, dt: float, n_frames: int, rng: np.random.Generator, n_random: int = 24) -> np.ndarray:
    mid = n_frames // 2
    fixed = [
        0.0, -0.3 * dt, -dt, -2.5 * dt, -9.0 * dt,  # negative history times extrapolate (not clamped)
        dt, 2 * dt, mid * dt, (mid + 1) * dt,  # exact frame times
        0.5 * dt, 1.5 * dt, 2.5 * dt, (mid + 0.5) * dt,  # half-frame ties (round-half-even)
        0.25 * dt, 0.75 * dt, (mid + 0.49999) * dt, (mid + 0.50001) * dt,
        duration, duration - 0.5 * dt, duration - 1e-9, duration + 0.01, duration + 1.0,  # clip end / past the end
        np.nextafter(duration, 0.0), np.nextafter(duration, 10.0),
    ]  # fmt: skip
    return np.concatenate([np.array(fixed, dtype=np.float64), rng.uniform(0.0, duration, n_random)])


def robot_of(name: str):
    return HUMANOID28 if name.startswith(("humanoid", "pooled_humanoid")) else G1


def main() -> None:
    assert build_ref.build_ref(REF), "needs /root/reference"
    RefLoader = ref_harness.reference_motion_loader_class()
    os.makedirs(f"{HERE}/clips", exist_ok=True)
    os.makedirs(f"{HERE}/clips_full", exist_ok=True)
    rng = np.random.default_rng(20261018)
    vec: dict[str, np.ndarray] = {}

    def record(tag: str, loader, times, ids, robot):
        i0, i1, blend = loader._compute_frame_blend(times, ids)
        outs = loader.sample(len(times), times=times, motion_ids=ids)
        vec[f"{tag}/times"] = times
        vec[f"{tag}/ids"] = np.asarray(ids, dtype=np.int64)
        vec[f"{tag}/idx0"], vec[f"{tag}/idx1"], vec[f"{tag}/blend"] = i0.astype(np.int64), i1.astype(np.int64), blend
        for name, t in zip(("dof_pos", "dof_vel", "body_pos", "body_rot", "body_lin", "body_ang"), outs):
            vec[f"{tag}/{name}"] = t.numpy()
        # collect_reference_motions: the reference method's own text (g1_amp_env.py:445-486), K = 2 and 10
        for k in (2, 10):
            vec[f"{tag}/amp_obs_k{k}"] = ref_harness.reference_collect(loader, robot, k, times, np.asarray(ids)).numpy()

    for name in CLIPS:
        with np.load(f"{REF}/motions/{name}.npz") as d:
            full = {k: d[k] for k in d.files}
        start = pick_window(full["body_rotations"], WINDOW)
        window = {}
        for k, v in full.items():
            window[k] = v[start : start + WINDOW] if (v.ndim >= 1 and v.shape[0] == full["dof_positions"].shape[0] and k not in ("dof_names", "body_names")) else v
        window.pop("record_time", None)
        path = f"{HERE}/clips/{name}.npz"
        np.savez_compressed(path, **window)

        loader = RefLoader(path, "cpu")
        times = edge_times(float(loader.durations[0]), float(loader.dt), WINDOW, rng)
        record(name, loader, times, np.zeros(len(times), dtype=np.int64), robot_of(name))
        vec[f"{name}/window_start"] = np.array(start)

        # the full shipped clip (data file) + reference outputs on it
        shutil.copyfile(f"{REF}/motions/{name}.npz", f"{HERE}/clips_full/{name}.npz")
        os.chmod(f"{HERE}/clips_full/{name}.npz", 0o644)
        floader = RefLoader(f"{HERE}/clips_full/{name}.npz", "cpu")
        ftimes = edge_times(float(floader.durations[0]), float(floader.dt), int(floader.num_frames), rng, n_random=40)
        record(f"full/{name}", floader, ftimes, np.zeros(len(ftimes), dtype=np.int64), robot_of(name))

    # pooled multi-clip loaders (comma list), ids spread over the three trajectories: windows and full clips
    for tag, folder in (("pooled_humanoid", "clips"), ("full/pooled_humanoid", "clips_full")):
        pooled = ",".join(f"{HERE}/{folder}/{n}.npz" for n in ("humanoid_walk", "humanoid_run", "humanoid_dance"))
        loader = RefLoader(pooled, "cpu")
        per = [edge_times(float(loader.durations[j]), float(loader.dt), WINDOW, rng, n_random=8) for j in range(3)]
        times = np.concatenate(per)
        ids = np.concatenate([np.full(len(p), j, dtype=np.int64) for j, p in enumerate(per)])
        perm = rng.permutation(len(times))
        record(tag, loader, times[perm], ids[perm], HUMANOID28)
        np.random.seed(123)
        st_ids, st_times = loader.sample_times(16)
        vec[f"{tag}/seed123_ids"], vec[f"{tag}/seed123_times"] = st_ids.astype(np.int64), st_times

    record_env_fixtures(vec, RefLoader)
    np.savez_compressed(f"{HERE}/vectors.npz", **vec)

    # ---- known answers on the FULL clips, re-derived from the live reference ----------------
    kat = {}
    g1 = RefLoader(f"{REF}/motions/G1_walk.npz", "cpu")
    cur = np.array([0, 0.004, 0.0125, 1.0, 10 / 3, 6.63, 6.6333333333333333, 7.0])
    t = (np.expand_dims(cur, axis=-1) - g1.dt * np.arange(0, 2)).flatten()
    i0, i1, b = g1._compute_frame_blend(t, np.zeros_like(t, dtype=np.int32))
    obs = ref_harness.reference_collect(g1, G1, 2, cur, np.zeros(len(cur), dtype=np.int64))
    kat["g1_walk"] = {
        "num_frames": int(g1.num_frames), "dt_hex": float(g1.dt).hex(), "duration_hex": float(g1.durations[0]).hex(),
        "current_times": cur.tolist(), "idx0": i0.tolist(), "idx1": i1.tolist(), "blend": b.tolist(),
        "obs_row_sums": obs.double().sum(dim=1).tolist(), "obs3_58_71": obs[3, 58:71].tolist(),
        "dof_perm": g1.get_dof_index(list(G1.joint_names)),
    }  # fmt: skip
    pool = RefLoader(",".join(f"{REF}/motions/humanoid_{n}.npz" for n in ("walk", "run", "dance")), "cpu")
    np.random.seed(123)
    pid, ptm = pool.sample_times(5)
    kat["pooled_full"] = {
        "traj_starts": pool.traj_starts.tolist(), "traj_ends": pool.traj_ends.tolist(),
        "durations_hex": [float(x).hex() for x in pool.durations], "num_frames": int(pool.num_frames),
        "seed123_ids": pid.tolist(), "seed123_times": ptm.tolist(),
    }  # fmt: skip
    with open(f"{HERE}/kat.json", "w") as f:
        json.dump(kat, f, indent=1)
    total = sum(os.path.getsize(os.path.join(dp, fn)) for dp, _, fns in os.walk(HERE) for fn in fns)
    print(f"wrote fixtures under {HERE}: {total / 1e6:.2f} MB, torch {torch.__version__}, numpy {np.__version__}")


# ---------------------------------------------------------------------------------------------------------------------
# env rows: the reference's _get_observations / _get_rewards / _reset_strategy_random text on stored simulator states
# ---------------------------------------------------------------------------------------------------------------------
ENV_N = 24  # envs per fixture
ENV_STEPS = 12  # simulator states stored per robot (K = 10 needs > 10 steps to fill and shift its history)
# (tag, K, num_actor_observations, rew_track_vel, history_include_last_actions, history_include_command)
G1_OBS_CASES = [
    ("k2_a1", 2, 1, 0.0, True, True),
    ("k10_a1", 10, 1, 0.0, True, True),
    ("k1_a1_cmd", 1, 1, 1.0, True, True),
    ("k3_a3_cmd", 3, 3, 1.0, True, True),
    ("k2_a4_noact", 2, 4, 1.0, False, True),
    ("k2_a3_nocmd", 2, 3, 1.0, True, False),
    ("k2_a5", 2, 5, 0.0, True, True),
]
REWARD_SCALES = dict(rew_termination=-1.0, rew_action_l2=-0.1, rew_joint_pos_limits=-10.0, rew_joint_acc_l2=-1.0e-06,
                     rew_joint_vel_l2=-0.001)  # fmt: skip  (the _CUSTOM cfg values, g1_amp_env_cfg.py:86-91)


def record_env_fixtures(vec, RefLoader) -> None:
    g = torch.Generator().manual_seed(20261019)
    for robot, clip in ((G1, "G1_dance"), (HUMANOID28, "humanoid_walk")):
        loader = RefLoader(f"{HERE}/clips/{clip}.npz", "cpu")
        tag = f"env/{robot.name}"
        N, D = ENV_N, robot.num_joints
        states = [synthetic_sim_state(N, robot, "cpu", seed=9000 + s) for s in range(ENV_STEPS)]
        for i, nm in enumerate(("joint_pos", "joint_vel", "body_pos_w", "body_quat_w", "body_lin_vel_w", "body_ang_vel_w")):
            vec[f"{tag}/{nm}"] = np.stack([s[i].numpy() for s in states])
        actions = torch.randn(ENV_STEPS, N, D, generator=g)
        command = torch.rand(ENV_STEPS, N, 2, generator=g) * 2 - 1
        reset_mask = torch.rand(ENV_STEPS, N, generator=g) < 0.3
        reset_mask[1] = False
        vec[f"{tag}/last_actions"], vec[f"{tag}/command"], vec[f"{tag}/reset_mask"] = actions.numpy(), command.numpy(), reset_mask.numpy()

        if robot is HUMANOID28:
            # HumanoidAmpEnv._get_observations (humanoid_amp_env.py:105-126): AMP history only, policy obs = the row
            for K in (2, 10):
                env = ref_harness.make_ref_env(loader, robot, N, K, humanoid=True)
                views, pol = [], []
                for s in range(ENV_STEPS):
                    ref_harness.set_sim_state(env, *states[s])
                    pol.append(env._get_observations()["policy"].clone().numpy())
                    views.append(env.extras["amp_obs"].clone().numpy())
                vec[f"{tag}/k{K}/amp_obs"] = np.stack(views if K == 2 else [views[2], views[-1]])
                vec[f"{tag}/k{K}/policy"] = np.stack(pol if K == 2 else [pol[2], pol[-1]])
            continue

        for case, K, n_actor, track, inc_act, inc_cmd in G1_OBS_CASES:
            env = ref_harness.make_ref_env(loader, robot, N, K, num_actor_observations=n_actor, rew_track_vel=track,
                                           history_include_last_actions=inc_act, history_include_command=inc_cmd)  # fmt: skip
            views, pol = [], []
            for s in range(ENV_STEPS):
                ref_harness.set_sim_state(env, *states[s])
                env.last_actions = actions[s].clone()
                env.command_target_speed = command[s].clone()
                if n_actor > 1:
                    env._just_reset_mask |= reset_mask[s]
                pol.append(env._get_observations()["policy"].clone().numpy())
                views.append(env.extras["amp_obs"].clone().numpy())
            keep = range(ENV_STEPS) if K <= 3 else (2, ENV_STEPS - 1)  # long histories: an early and the final step only
            vec[f"{tag}/{case}/amp_obs"] = np.stack([views[s] for s in keep])
            vec[f"{tag}/{case}/policy"] = np.stack([pol[s] for s in keep])
            vec[f"{tag}/{case}/steps"] = np.array(list(keep))

        # ---- _get_rewards (g1_amp_env.py:246-319 with compute_rewards / exp_reward_with_floor) ----
        acc = torch.randn(N, D, generator=g) * 50
        lo = torch.rand(N, D, generator=g) * -2.0
        limits = torch.stack([lo, lo + torch.rand(N, D, generator=g) * 3.0], dim=-1)
        terminated = torch.rand(N, generator=g) < 0.2
        vec[f"{tag}/reward/joint_acc"], vec[f"{tag}/reward/soft_limits"], vec[f"{tag}/reward/terminated"] = acc.numpy(), limits.numpy(), terminated.numpy()
        jp, jv, bp, bq, bl, ba = states[0]
        bl = bl * 0.5
        bl[::5] *= 4  # a share of envs beyond the exp / linear threshold of the tracking reward
        vec[f"{tag}/reward/body_lin_vel_w"] = bl.numpy()
        for track in (0.0, 1.0):
            env = ref_harness.make_ref_env(loader, robot, N, 2, rew_track_vel=track, **REWARD_SCALES)
            ref_harness.set_sim_state(env, jp, jv, bp, bq, bl, ba, joint_acc=acc, soft_joint_pos_limits=limits)
            env.actions = actions[0].clone()
            env.command_target_speed = command[0].clone()
            env.reset_terminated = terminated.clone()
            total = env._get_rewards()
            vec[f"{tag}/reward/track{int(track)}/total"] = total.numpy()
            log = env.extras["log"]
            vec[f"{tag}/reward/track{int(track)}/log_keys"] = np.array(sorted(log))
            vec[f"{tag}/reward/track{int(track)}/log_values"] = np.array([log[k] for k in sorted(log)], dtype=np.float64)

        # ---- _reset_strategy_random (g1_amp_env.py:371-441): host RNG stream, root / dof state, history fill ----
        for K in (2, 10):
            env = ref_harness.make_ref_env(loader, robot, N, K, track_vel_range=(0.5, 0.5))
            default_root = torch.randn(N, 13, generator=g)
            origins = torch.randn(N, 3, generator=g) * 5
            env.robot.data.default_root_state = default_root
            env.scene.env_origins = origins
            env.amp_observation_buffer.fill_(3.0)
            env_ids = torch.arange(0, N, 3)
            np.random.seed(99)
            root, dof_p, dof_v = env._reset_strategy_random(env_ids)
            r = f"{tag}/reset_k{K}"
            vec[f"{r}/default_root_state"], vec[f"{r}/env_origins"], vec[f"{r}/env_ids"] = default_root.numpy(), origins.numpy(), env_ids.numpy()
            vec[f"{r}/root_state"], vec[f"{r}/dof_pos"], vec[f"{r}/dof_vel"] = root.numpy(), dof_p.numpy(), dof_v.numpy()
            vec[f"{r}/amp_observation_buffer"] = env.amp_observation_buffer.numpy().copy()
            vec[f"{r}/motion_ids"] = env.motion_ids.numpy().copy()
            vec[f"{r}/motion_start_times"] = env.motion_start_times.numpy().copy()
            vec[f"{r}/command_target_speed"] = env.command_target_speed.numpy().copy()


if __name__ == "__main__":
    main()
