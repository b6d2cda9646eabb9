"""Synthetic inputs of the shipped clips' shapes (the reference's ``.npz`` clips and PhysX do not travel to the GPU box).

* :func:`write_synthetic_clip` -- a clip in the reference on-disk format (``motions/README.md:11-21``) with the frame /
  dof / body counts of ``G1_walk.npz`` (399, 29, 11), ``G1_dance.npz`` (601, 29, 39), ``humanoid_walk.npz`` (154, 28, 15)
  ...; smooth joint trajectories, unit quaternions from integrated angular velocity, plus the features that make the
  real clips hard: repeated frames (quaternion dot == 1), sign-flipped quaternions (dot < 0) and float64 dof arrays.
* :func:`synthetic_sim_state` -- Isaac-Lab-shaped simulator state for the per-step path; value ranges follow the clips
  (SURVEY.md section 8d).
* :func:`skrl_style_discriminator_params` -- torch.nn.Linear default initialisation of the 1024-512-1 MLP from seed 42
  (``agents/*.yaml``: ``seed: 42``) and scaler statistics.
"""

from __future__ import annotations

import math
import os

import numpy as np
import torch

from .robots import G1, HUMANOID28, RobotSpec

# clip name -> (frames, robot, clip body names, clip dof order, dof dtype)
_G1_WALK_BODIES = (
    "pelvis", "left_shoulder_pitch_link", "right_shoulder_pitch_link", "left_elbow_link", "right_elbow_link",
    "right_hip_yaw_link", "left_hip_yaw_link", "right_rubber_hand", "left_rubber_hand", "right_ankle_roll_link",
    "left_ankle_roll_link",
)  # fmt: skip
# URDF depth-first joint order used by the converted LAFAN1 clips (differs from the simulator order -> real permutation)
_G1_URDF_JOINTS = (
    "left_hip_pitch_joint", "left_hip_roll_joint", "left_hip_yaw_joint", "left_knee_joint", "left_ankle_pitch_joint",
    "left_ankle_roll_joint", "right_hip_pitch_joint", "right_hip_roll_joint", "right_hip_yaw_joint", "right_knee_joint",
    "right_ankle_pitch_joint", "right_ankle_roll_joint", "waist_yaw_joint", "waist_roll_joint", "waist_pitch_joint",
    "left_shoulder_pitch_joint", "left_shoulder_roll_joint", "left_shoulder_yaw_joint", "left_elbow_joint",
    "left_wrist_roll_joint", "left_wrist_pitch_joint", "left_wrist_yaw_joint", "right_shoulder_pitch_joint",
    "right_shoulder_roll_joint", "right_shoulder_yaw_joint", "right_elbow_joint", "right_wrist_roll_joint",
    "right_wrist_pitch_joint", "right_wrist_yaw_joint",
)  # fmt: skip

CLIP_SHAPES = {
    "G1_walk": dict(frames=399, robot=G1, bodies=_G1_WALK_BODIES, dofs=_G1_URDF_JOINTS, dof_dtype=np.float64),
    "G1_dance": dict(frames=601, robot=G1, bodies=G1.body_names, dofs=G1.joint_names, dof_dtype=np.float32),
    "G1_walk_lafan1": dict(frames=1199, robot=G1, bodies=_G1_WALK_BODIES, dofs=_G1_URDF_JOINTS, dof_dtype=np.float64),
    "humanoid_walk": dict(frames=154, robot=HUMANOID28, bodies=HUMANOID28.body_names, dofs=HUMANOID28.joint_names, dof_dtype=np.float32),
    "humanoid_run": dict(frames=82, robot=HUMANOID28, bodies=HUMANOID28.body_names, dofs=HUMANOID28.joint_names, dof_dtype=np.float32),
    "humanoid_dance": dict(frames=902, robot=HUMANOID28, bodies=HUMANOID28.body_names, dofs=HUMANOID28.joint_names, dof_dtype=np.float32),
}


def _quat_mul(a, b):
    w1, x1, y1, z1 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    w2, x2, y2, z2 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack(
        [
            w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
            w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
            w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
        ],
        axis=-1,
    )


def _smooth(rng, frames, channels, scale, harmonics=4):
    t = np.linspace(0.0, 1.0, frames)[:, None, None]
    k = np.arange(1, harmonics + 1)[None, None, :]
    amp = rng.normal(0, 1, (1, channels, harmonics)) / k
    phase = rng.uniform(0, 2 * math.pi, (1, channels, harmonics))
    return scale * np.sum(amp * np.sin(2 * math.pi * k * t + phase), axis=-1)


def synthetic_clip_arrays(name: str = "G1_walk", seed: int = 0, frames: int | None = None) -> dict:
    spec = CLIP_SHAPES[name]
    rng = np.random.default_rng(seed)
    F = int(frames or spec["frames"])
    D, B = len(spec["dofs"]), len(spec["bodies"])
    fps = 60
    dt = 1.0 / fps
    dof_pos = _smooth(rng, F, D, 0.8) + rng.uniform(-0.5, 0.5, (1, D))
    dof_vel = np.gradient(dof_pos, dt, axis=0)
    root = np.stack([np.linspace(0, 3.0, F), 0.2 * np.sin(np.linspace(0, 6, F)), 0.75 + 0.03 * np.sin(np.linspace(0, 40, F))], axis=-1)
    offsets = rng.uniform(-0.6, 0.6, (1, B, 3))
    offsets[:, 0] = 0.0
    body_pos = root[:, None, :] + offsets + _smooth(rng, F, B * 3, 0.1).reshape(F, B, 3)
    body_lin = np.gradient(body_pos, dt, axis=0)
    ang = _smooth(rng, F, B * 3, 2.0).reshape(F, B, 3)
    # integrate angular velocity into unit quaternions
    q = np.zeros((F, B, 4))
    q0 = rng.normal(0, 1, (B, 4))
    q[0] = q0 / np.linalg.norm(q0, axis=-1, keepdims=True)
    for f in range(1, F):
        half = 0.5 * dt * ang[f - 1]
        dq = np.concatenate([np.ones((B, 1)), half], axis=-1)
        nq = _quat_mul(dq, q[f - 1])
        q[f] = nq / np.linalg.norm(nq, axis=-1, keepdims=True)
    # features of the real clips: held poses (dot == 1 exactly after fp32 rounding) and sign flips (dot < 0)
    hold = rng.choice(np.arange(1, F), size=max(F // 10, 1), replace=False)
    q[hold] = q[hold - 1]
    flip = rng.random((F, B)) < 0.02
    q[flip] *= -1.0
    return {
        "fps": np.array(fps, dtype=np.int64),
        "dof_names": np.array(spec["dofs"]),
        "body_names": np.array(spec["bodies"]),
        "dof_positions": dof_pos.astype(spec["dof_dtype"]),
        "dof_velocities": dof_vel.astype(spec["dof_dtype"]),
        "body_positions": body_pos.astype(np.float32),
        "body_rotations": q.astype(np.float32),
        "body_linear_velocities": body_lin.astype(np.float32),
        "body_angular_velocities": ang.astype(np.float32),
    }


def write_synthetic_clip(path: str, name: str = "G1_walk", seed: int = 0, frames: int | None = None) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path, **synthetic_clip_arrays(name, seed, frames))
    return path


def synthetic_sim_state(num_envs: int, robot: RobotSpec = G1, device="cuda", seed: int = 0):
    """``(joint_pos, joint_vel, body_pos_w, body_quat_w, body_lin_vel_w, body_ang_vel_w)`` as Isaac Lab lays them out."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    D, B = robot.num_joints, len(robot.body_names)
    joint_pos = torch.rand(num_envs, D, generator=g) * 4.3 - 1.8
    joint_vel = torch.randn(num_envs, D, generator=g) * 3.0
    body_pos = torch.rand(num_envs, B, 3, generator=g) * 5.0 - 1.0
    quat = torch.randn(num_envs, B, 4, generator=g)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    lin = torch.randn(num_envs, B, 3, generator=g) * 2.0
    ang = torch.randn(num_envs, B, 3, generator=g) * 2.0
    return tuple(t.to(device) for t in (joint_pos, joint_vel, body_pos, quat, lin, ang))


def skrl_style_discriminator_params(in_features: int, hidden=(1024, 512), seed: int = 42, logit_gain: float = 1.0):
    """``([W1, W2, W3], [b1, b2, b3])`` with torch.nn.Linear's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)) init.

    ``logit_gain`` rescales the last layer to emulate a trained discriminator with a wider logit range."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    dims = [in_features, *hidden, 1]
    weights, biases = [], []
    for fan_in, fan_out in zip(dims[:-1], dims[1:]):
        bound = 1.0 / math.sqrt(fan_in)
        weights.append((torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound)
        biases.append((torch.rand(fan_out, generator=g) * 2 - 1) * bound)
    weights[-1] = weights[-1] * logit_gain
    biases[-1] = biases[-1] * logit_gain
    return weights, biases
