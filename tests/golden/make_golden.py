#!/usr/bin/env python
"""Generate the committed golden fixtures from the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py

The reference ``MotionLoader`` is imported by file path from ``/root/reference/motions/motion_loader.py`` (it needs only
numpy, torch, yaml; the package ``__init__`` would pull matplotlib, so the package itself is not imported).  Nothing
from the reference's source is copied: the fixtures hold

* ``clips/<name>.npz``   -- a short window (``WINDOW`` frames) of each shipped motion clip in the reference's own on-disk
                            format, chosen around the frame with the most negative consecutive-frame quaternion dot so
                            the shortest-arc flip and both slerp fall-backs are exercised;
* ``vectors.npz``        -- for every window (and for a pooled 3-clip humanoid loader): seeded + edge-case ``times`` /
                            ``motion_ids`` and the outputs of the live reference ``_compute_frame_blend`` and ``sample``;
                            plus ``collect_reference_motions`` rows built from the live loader's ``sample`` output and
                            the restated ``compute_obs`` (the env module needs isaaclab and cannot be imported);
* ``kat.json``           -- known answers on the FULL shipped clips recorded in SURVEY.md section 8c, re-derived here.

``/root/reference`` does not exist on the GPU box, so tests only ever read these files.
"""

from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import env_oracle  # noqa: E402

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
G1_KEYS = ["right_rubber_hand", "left_rubber_hand", "right_ankle_roll_link", "left_ankle_roll_link"]
HUM_KEYS = ["right_hand", "left_hand", "right_foot", "left_foot"]
# robot joint order printed by Isaac Sim, recorded as a comment in the reference (motions/test/get_joint_name.py:231)
G1_ROBOT_JOINTS = [
    "left_hip_pitch_joint", "right_hip_pitch_joint", "waist_yaw_joint", "left_hip_roll_joint", "right_hip_roll_joint",
    "waist_roll_joint", "left_hip_yaw_joint", "right_hip_yaw_joint", "waist_pitch_joint", "left_knee_joint",
    "right_knee_joint", "left_shoulder_pitch_joint", "right_shoulder_pitch_joint", "left_ankle_pitch_joint",
    "right_ankle_pitch_joint", "left_shoulder_roll_joint", "right_shoulder_roll_joint", "left_ankle_roll_joint",
    "right_ankle_roll_joint", "left_shoulder_yaw_joint", "right_shoulder_yaw_joint", "left_elbow_joint",
    "right_elbow_joint", "left_wrist_roll_joint", "right_wrist_roll_joint", "left_wrist_pitch_joint",
    "right_wrist_pitch_joint", "left_wrist_yaw_joint", "right_wrist_yaw_joint",
]  # fmt: skip


def load_reference_loader_class():
    spec = importlib.util.spec_from_file_location("ref_motion_loader", f"{REF}/motions/motion_loader.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MotionLoader


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


def main() -> None:
    RefLoader = load_reference_loader_class()
    os.makedirs(f"{HERE}/clips", exist_ok=True)
    rng = np.random.default_rng(20261018)
    vec: dict[str, np.ndarray] = {}

    def record(tag: str, loader, times, ids, dof_names_robot, ref_body, key_names):
        i0, i1, blend = loader._compute_frame_blend(times, ids)
        outs = loader.sample(len(times), times=times, motion_ids=ids)
        vec[f"{tag}/times"] = times
        vec[f"{tag}/ids"] = np.asarray(ids, dtype=np.int64)
        vec[f"{tag}/idx0"], vec[f"{tag}/idx1"], vec[f"{tag}/blend"] = i0.astype(np.int64), i1.astype(np.int64), blend
        for name, t in zip(("dof_pos", "dof_vel", "body_pos", "body_rot", "body_lin", "body_ang"), outs):
            vec[f"{tag}/{name}"] = t.numpy()
        # collect_reference_motions rows: live loader.sample + restated compute_obs, K = 2 and 10
        dof_idx = loader.get_dof_index(dof_names_robot)
        ref_idx = loader.get_body_index([ref_body])[0]
        key_idx = loader.get_body_index(key_names)
        for k in (2, 10):
            obs = env_oracle.collect_reference_motions(
                loader, len(times), k, dof_idx, ref_idx, key_idx, current_times=times, motion_ids=np.asarray(ids)
            )
            vec[f"{tag}/amp_obs_k{k}"] = obs.numpy()

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
        humanoid = name.startswith("humanoid")
        times = edge_times(float(loader.durations[0]), float(loader.dt), WINDOW, rng)
        ids = np.zeros(len(times), dtype=np.int64)
        record(
            name, loader, times, ids,
            loader.dof_names if humanoid else G1_ROBOT_JOINTS,
            "torso" if humanoid else "pelvis",
            HUM_KEYS if humanoid else G1_KEYS,
        )  # fmt: skip
        vec[f"{name}/window_start"] = np.array(start)

    # pooled multi-clip loader (comma list), ids spread over the three trajectories
    pooled = ",".join(f"{HERE}/clips/{n}.npz" for n in ("humanoid_walk", "humanoid_run", "humanoid_dance"))
    loader = RefLoader(pooled, "cpu")
    per = [edge_times(float(loader.durations[j]), float(loader.dt), WINDOW, rng, n_random=8) for j in range(3)]
    times = np.concatenate(per)
    ids = np.concatenate([np.full(len(p), j, dtype=np.int64) for j, p in enumerate(per)])
    perm = rng.permutation(len(times))
    record("pooled_humanoid", loader, times[perm], ids[perm], loader.dof_names, "torso", HUM_KEYS)
    np.random.seed(123)
    st_ids, st_times = loader.sample_times(16)
    vec["pooled_humanoid/seed123_ids"], vec["pooled_humanoid/seed123_times"] = st_ids.astype(np.int64), st_times

    np.savez_compressed(f"{HERE}/vectors.npz", **vec)

    # ---- known answers on the FULL clips, re-derived from the live reference ----------------
    kat = {}
    g1 = RefLoader(f"{REF}/motions/G1_walk.npz", "cpu")
    cur = np.array([0, 0.004, 0.0125, 1.0, 10 / 3, 6.63, 6.6333333333333333, 7.0])
    t = env_oracle.history_times(cur, g1.dt, 2)
    i0, i1, b = g1._compute_frame_blend(t, np.zeros_like(t, dtype=np.int32))
    obs = env_oracle.collect_reference_motions(
        g1, len(cur), 2, g1.get_dof_index(G1_ROBOT_JOINTS), g1.get_body_index(["pelvis"])[0], g1.get_body_index(G1_KEYS),
        current_times=cur, motion_ids=np.zeros(len(cur), dtype=np.int64),
    )  # fmt: skip
    kat["g1_walk"] = {
        "num_frames": int(g1.num_frames), "dt_hex": float(g1.dt).hex(), "duration_hex": float(g1.durations[0]).hex(),
        "current_times": cur.tolist(), "idx0": i0.tolist(), "idx1": i1.tolist(), "blend": b.tolist(),
        "obs_row_sums": obs.double().sum(dim=1).tolist(), "obs3_58_71": obs[3, 58:71].tolist(),
        "dof_perm": g1.get_dof_index(G1_ROBOT_JOINTS),
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


if __name__ == "__main__":
    main()
