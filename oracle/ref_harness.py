"""Harness that RUNS the reference's own text for the env rows of the path (TEST INFRASTRUCTURE -- see ``oracle/__init__.py``).

``oracle/build_ref.py`` installs, under the git-ignored ``oracle/_ref/``, the unmodified ``MotionLoader`` module and the
source text of the ``G1AmpEnv`` / ``HumanoidAmpEnv`` methods on the path (``_get_observations``, ``_get_rewards``,
``_reset_strategy_random``, ``collect_reference_motions``) plus the scripted free functions.  Those methods read
``self.robot.data.*``, ``self.cfg.*``, ``self.scene.env_origins`` ... which Isaac Lab's ``DirectRLEnv`` would provide; this
module builds a bare object with exactly those attributes (simulator state = caller tensors, PhysX is out of scope) so
the reference's statements execute verbatim on CPU torch.

Used by ``tests/golden/make_golden.py`` (fixtures), by the CPU tests when ``oracle/_ref`` is present (live comparison with
``oracle/env_oracle.py``), and by ``bench.py``'s reference arm / ``cpu_baseline`` (kind "reference").
"""

from __future__ import annotations

import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("motion_loader.py", "g1_amp_env_ref.py", "humanoid_amp_env_ref.py"))


def _load(name: str):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    key = f"oracle._ref.{name}"
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(REF_DIR, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod  # torch.jit.script resolves sources through the module of the function
    spec.loader.exec_module(mod)
    return mod


def reference_motion_loader_class():
    """The reference's ``MotionLoader`` (``motions/motion_loader.py:87-430``), unmodified."""
    return _load("motion_loader").MotionLoader


def g1_module():
    return _load("g1_amp_env_ref")


def humanoid_module():
    return _load("humanoid_amp_env_ref")


def make_ref_env(loader, robot, num_envs: int, num_amp_observations: int, *, num_actor_observations: int = 1, rew_track_vel: float = 0.0,
                 rew_termination: float = 0.0, rew_action_l2: float = 0.0, rew_joint_pos_limits: float = 0.0,
                 rew_joint_acc_l2: float = 0.0, rew_joint_vel_l2: float = 0.0, history_include_last_actions: bool = True,
                 history_include_command: bool = True, track_vel_range=(0.0, 0.0), command_resampling_time_range=(1.0, 2.0),
                 humanoid: bool = False):
    """A bare ``RefG1AmpEnv`` / ``RefHumanoidAmpEnv`` carrying the attributes ``G1AmpEnv.__init__`` (``g1_amp_env.py:25-119``)
    sets, derived the same way (``robot.data.body_names.index`` ..., ``MotionLoader.get_dof_index`` ...).  ``robot`` is a
    ``humanoid_amp_b200.robots.RobotSpec`` (simulator joint / body order); ``loader`` the reference ``MotionLoader``."""
    mod = humanoid_module() if humanoid else g1_module()
    cls = mod.RefHumanoidAmpEnv if humanoid else mod.RefG1AmpEnv
    env = object.__new__(cls)
    A = robot.amp_observation_space
    key_body_names = list(robot.key_body_names)
    env.device = torch.device("cpu")
    env.num_envs = num_envs
    env.cfg = SimpleNamespace(
        num_amp_observations=num_amp_observations, amp_observation_space=A, num_actor_observations=num_actor_observations,
        rew_track_vel=rew_track_vel, rew_termination=rew_termination, rew_action_l2=rew_action_l2,
        rew_joint_pos_limits=rew_joint_pos_limits, rew_joint_acc_l2=rew_joint_acc_l2, rew_joint_vel_l2=rew_joint_vel_l2,
        history_include_last_actions=history_include_last_actions, history_include_command=history_include_command,
        reference_body=robot.reference_body, track_vel_range=tuple(track_vel_range),
        command_resampling_time_range=tuple(command_resampling_time_range),
    )  # fmt: skip
    env.robot = SimpleNamespace(data=SimpleNamespace(joint_names=list(robot.joint_names), body_names=list(robot.body_names)))
    env.scene = SimpleNamespace(env_origins=torch.zeros(num_envs, 3))
    env._motion_loader = loader
    # g1_amp_env.py:47-62
    env.ref_body_index = env.robot.data.body_names.index(env.cfg.reference_body)
    env.key_body_indexes = [env.robot.data.body_names.index(n) for n in key_body_names]
    env.motion_dof_indexes = loader.get_dof_index(env.robot.data.joint_names)
    env.motion_ref_body_index = loader.get_body_index([env.cfg.reference_body])[0]
    env.motion_key_body_indexes = loader.get_body_index(key_body_names)
    env.amp_observation_size = num_amp_observations * A
    env.amp_observation_buffer = torch.zeros((num_envs, num_amp_observations, A))
    # g1_amp_env.py:75-119
    env.command_target_speed = torch.zeros((num_envs, 2), dtype=torch.float32)
    env.command_time_left = torch.zeros(num_envs, dtype=torch.float32)
    env.motion_ids = torch.zeros(num_envs, dtype=torch.long)
    env.motion_start_times = torch.zeros(num_envs, dtype=torch.float32)
    env.key_body_obs_size = len(key_body_names) * 3
    env.last_actions = torch.zeros((num_envs, len(robot.joint_names)))
    env.actions = torch.zeros((num_envs, len(robot.joint_names)))
    env.reset_terminated = torch.zeros(num_envs, dtype=torch.bool)
    env.extras = {}
    if num_actor_observations > 1:
        base = A - env.key_body_obs_size
        cmd = 2 if rew_track_vel > 0.0 else 0
        per = base + (len(robot.joint_names) if history_include_last_actions else 0) + (cmd if history_include_command else 0)
        env.actor_obs_hist_per_frame = per
        env.actor_obs_history_buffer = torch.zeros((num_envs, num_actor_observations - 1, per))
        env._just_reset_mask = torch.zeros(num_envs, dtype=torch.bool)
    return env


def set_sim_state(env, joint_pos, joint_vel, body_pos_w, body_quat_w, body_lin_vel_w, body_ang_vel_w, *, joint_acc=None,
                  soft_joint_pos_limits=None, default_root_state=None) -> None:
    """What PhysX would have written into ``robot.data`` this step."""
    d = env.robot.data
    d.joint_pos, d.joint_vel = joint_pos, joint_vel
    d.body_pos_w, d.body_quat_w, d.body_lin_vel_w, d.body_ang_vel_w = body_pos_w, body_quat_w, body_lin_vel_w, body_ang_vel_w
    if joint_acc is not None:
        d.joint_acc = joint_acc
    if soft_joint_pos_limits is not None:
        d.soft_joint_pos_limits = soft_joint_pos_limits
    if default_root_state is not None:
        d.default_root_state = default_root_state


def reference_collect(loader, robot, K: int, current_times, motion_ids=None) -> torch.Tensor:
    """``G1AmpEnv.collect_reference_motions`` (``g1_amp_env.py:445-486``) executed from the reference's text."""
    env = make_ref_env(loader, robot, 1, K)
    n = len(current_times)
    return env.collect_reference_motions(n, np.asarray(current_times), None if motion_ids is None else np.asarray(motion_ids))
