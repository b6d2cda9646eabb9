"""Env-side AMP observation path with the reference's call surface, on hand-written sm_100a kernels.

Mirrors the AMP part of ``G1AmpEnv`` / ``HumanoidAmpEnv`` (``/root/reference/g1_amp_env.py``):

* free functions ``compute_obs`` (``:535-561``) and ``quaternion_to_tangent_and_normal`` (``:489-497``);
* :class:`AmpEnvPath` -- the state and methods the skrl AMP agent and the env's step/reset touch:
  ``collect_reference_motions`` (``:445-486``), the history update of ``_get_observations`` (``:176-193``) and the
  reset-time history fill of ``_reset_strategy_random`` (``:371-419``).

PhysX stepping is out of scope: the simulator state (``robot.data.joint_pos`` ... ``body_ang_vel_w``) is passed in as
tensors.  There is no CPU path.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .motion_loader import MotionLoader
from .robots import G1, RobotSpec


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


def quaternion_to_tangent_and_normal(q: torch.Tensor) -> torch.Tensor:
    """``(..., 4)`` wxyz -> ``(..., 6)``: the rotated x axis (tangent) and z axis (normal).  Reference ``:489-497``."""
    dev = _lib.require_cuda(q.device)
    qc = _f32c(q, dev).view(-1, 4)
    out = torch.empty((qc.shape[0], 6), dtype=torch.float32, device=dev)
    lib, stream = _lib.enter(dev)
    _lib.check(lib.amp_tangent_normal(_lib.ptr(qc), qc.shape[0], _lib.ptr(out), stream))
    return out.view(*q.shape[:-1], 6)


def compute_obs(
    dof_positions: torch.Tensor,
    dof_velocities: torch.Tensor,
    root_positions: torch.Tensor,
    root_rotations: torch.Tensor,
    root_linear_velocities: torch.Tensor,
    root_angular_velocities: torch.Tensor,
    key_body_positions: torch.Tensor,
) -> torch.Tensor:
    """One AMP observation row per sample (reference ``:535-561``):

    ``[dof_pos D | dof_vel D | root_z | tangent 3 | normal 3 | root_lin_vel 3 | root_ang_vel 3 | key_pos - root_pos 3*Kb]``
    """
    dev = _lib.require_cuda(dof_positions.device)
    n, D = dof_positions.shape
    Kb = key_body_positions.shape[1]
    args = [
        _f32c(t, dev)
        for t in (dof_positions, dof_velocities, root_positions, root_rotations, root_linear_velocities, root_angular_velocities, key_body_positions)
    ]
    for t, width in zip(args[2:6], (3, 4, 3, 3)):
        if t.shape != (n, width):
            raise RuntimeError(f"expected a ({n}, {width}) tensor, got {tuple(t.shape)}")
    out = torch.empty((n, 2 * D + 13 + 3 * Kb), dtype=torch.float32, device=dev)
    lib, stream = _lib.enter(dev)
    _lib.check(lib.amp_compute_obs(*[_lib.ptr(t) for t in args], n, D, Kb, _lib.ptr(out), stream))
    return out


@dataclass
class AmpEnvCfg:
    """The cfg fields the AMP path reads (``g1_amp_env_cfg.py:45-53, 108-109``; ``humanoid_amp_env_cfg.py:35-42``)."""

    motion_file: str
    num_envs: int = 4096
    num_amp_observations: int = 2
    robot: RobotSpec = G1
    # policy-observation fields (g1_amp_env_cfg.py:32, 43, 47, 173-184); only used by get_observations()
    num_actor_observations: int = 1
    rew_track_vel: float = 0.0  # > 0: the 2-d velocity command is part of the actor observation
    # task-reward scales (g1_amp_env_cfg.py:27-31, 86-90); only used by get_rewards()
    rew_termination: float = 0.0
    rew_action_l2: float = 0.0
    rew_joint_pos_limits: float = 0.0
    rew_joint_acc_l2: float = 0.0
    rew_joint_vel_l2: float = 0.0
    history_include_last_actions: bool = True
    history_include_command: bool = True

    @property
    def amp_observation_space(self) -> int:
        return self.robot.amp_observation_space

    @property
    def reference_body(self) -> str:
        return self.robot.reference_body

    @property
    def action_space(self) -> int:
        return self.robot.num_joints

    @property
    def command_size(self) -> int:
        return 2 if self.rew_track_vel > 0.0 else 0

    @property
    def hist_frame_size(self) -> int:
        base = self.amp_observation_space - 3 * len(self.robot.key_body_names)
        return base + (self.action_space if self.history_include_last_actions else 0) + (self.command_size if self.history_include_command else 0)

    @property
    def observation_space(self) -> int:
        """Actor observation width, as ``__post_init__`` sizes it (g1_amp_env_cfg.py:186-206)."""
        cur = self.amp_observation_space - 3 * len(self.robot.key_body_names) + self.action_space + self.command_size
        return cur + max(self.num_actor_observations - 1, 0) * self.hist_frame_size


class AmpEnvPath:
    """State + methods of the env that lie on the AMP hot path.

    Attribute names follow the reference env (``g1_amp_env.py:35-74``): ``_motion_loader``, ``ref_body_index``,
    ``key_body_indexes``, ``motion_dof_indexes``, ``motion_ref_body_index``, ``motion_key_body_indexes``,
    ``amp_observation_size``, ``amp_observation_buffer``, ``extras``.
    """

    def __init__(self, cfg: AmpEnvCfg, device, motion_loader: Optional[MotionLoader] = None):
        self.cfg = cfg
        self.device = _lib.require_cuda(device)
        self.num_envs = cfg.num_envs
        self._motion_loader = motion_loader or MotionLoader(motion_file=cfg.motion_file, device=self.device)
        robot = cfg.robot
        key_body_names = list(robot.key_body_names)
        self.ref_body_index = robot.body_names.index(cfg.reference_body)
        self.key_body_indexes = [robot.body_names.index(n) for n in key_body_names]
        self.motion_dof_indexes = self._motion_loader.get_dof_index(robot.joint_names)
        self.motion_ref_body_index = self._motion_loader.get_body_index([cfg.reference_body])[0]
        self.motion_key_body_indexes = self._motion_loader.get_body_index(key_body_names)
        self.amp_observation_size = cfg.num_amp_observations * cfg.amp_observation_space
        self.amp_observation_buffer = torch.zeros(
            (self.num_envs, cfg.num_amp_observations, cfg.amp_observation_space), device=self.device
        )
        self.extras: dict = {}
        self._handle = self._motion_loader.make_env_handle(
            self.motion_dof_indexes, self.motion_ref_body_index, self.motion_key_body_indexes
        )
        assert self._handle.obs_width == cfg.amp_observation_space
        self._key_idx_host = np.ascontiguousarray(self.key_body_indexes, dtype=np.int32)
        # policy-observation state (g1_amp_env.py:75-119)
        self.key_body_obs_size = len(key_body_names) * 3
        self.last_actions = torch.zeros((self.num_envs, cfg.action_space), device=self.device)
        self.command_target_speed = torch.zeros((self.num_envs, 2), device=self.device, dtype=torch.float32)
        if cfg.num_actor_observations > 1:
            self.actor_obs_hist_per_frame = cfg.hist_frame_size
            self.actor_obs_history_buffer = torch.zeros(
                (self.num_envs, cfg.num_actor_observations - 1, self.actor_obs_hist_per_frame), device=self.device
            )
            self._just_reset_mask = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)

    # ---- reference motions (skrl calls this with one argument) ------------------------------------------------------
    def collect_reference_motions(self, num_samples: int, current_times=None, motion_ids=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``(num_samples, K*A)`` history-stacked reference observations, slot 0 = newest (reference ``:445-486``).

        ``current_times`` / ``motion_ids``: host numpy (reference behaviour) or CUDA tensors.  ``motion_ids=None`` with
        given times means clip 0 for every sample (reference ``:462``).  ``out``: optional destination with rows of
        ``K*A`` contiguous floats (e.g. a slice of a replay memory); a fresh tensor is returned otherwise.
        """
        loader = self._motion_loader
        if current_times is None:
            motion_ids, current_times = loader.sample_times(num_samples)
        t = loader._times_to_device(current_times)
        n = t.numel()
        ids = loader._ids_to_device(motion_ids, n)
        K, width = self.cfg.num_amp_observations, self.amp_observation_size
        if out is None:
            out = torch.empty((n, width), dtype=torch.float32, device=self.device)
        self._launch_collect(t, ids, n, out, row_index=None)
        return out.view(-1, width)

    def _launch_collect(self, t, ids, n, out: torch.Tensor, row_index: Optional[torch.Tensor], start_row: int = 0, capacity_rows: int = 0):
        width = self.amp_observation_size
        if out.dtype != torch.float32 or out.device != self.device:
            raise RuntimeError("destination must be a float32 tensor on the env device")
        rows = out.reshape(-1, out.shape[-1]) if out.dim() != 2 else out
        if rows.data_ptr() != out.data_ptr() or rows.stride(-1) != 1 or rows.shape[-1] < width:
            raise RuntimeError("destination rows must hold K*A contiguous floats")
        lib, stream = _lib.enter(self.device)
        _lib.check(
            lib.amp_collect_reference(
                self._handle.raw, _lib.ptr(t), _lib.ptr(ids), n, self.cfg.num_amp_observations, _lib.ptr(rows),
                rows.stride(0), capacity_rows, start_row, _lib.ptr(row_index), stream,
            )  # fmt: skip
        )

    def collect_reference_motions_into(self, memory: torch.Tensor, start_row: int, num_samples: int, current_times=None, motion_ids=None) -> int:
        """Write the rows straight into a ring memory ``(capacity, ..., K*A)`` starting at ``start_row`` and wrapping at
        the capacity (how skrl's ``RandomMemory.add_samples`` lays rows out); returns the next write index."""
        loader = self._motion_loader
        if current_times is None:
            motion_ids, current_times = loader.sample_times(num_samples)
        t = loader._times_to_device(current_times)
        n = t.numel()
        ids = loader._ids_to_device(motion_ids, n)
        rows = memory.view(memory.shape[0], -1)
        self._launch_collect(t, ids, n, rows, None, start_row=start_row, capacity_rows=rows.shape[0])
        return (start_row + n) % rows.shape[0]

    # ---- per-step update (reference _get_observations :176-193) -----------------------------------------------------
    def update_amp_observations(self, joint_pos, joint_vel, body_pos_w, body_quat_w, body_lin_vel_w, body_ang_vel_w, policy_obs: Optional[torch.Tensor] = None) -> torch.Tensor:
        """compute_obs from simulator state, shift the history, write slot 0 -- one kernel, in place.

        Returns ``extras["amp_obs"]``: the ``(num_envs, K*A)`` view of ``amp_observation_buffer``.  ``policy_obs``
        optionally receives ``obs[:, :-3*Kb]`` (the base actor observation, reference ``:196``).
        """
        dev = self.device
        tensors = [_f32c(t, dev) for t in (joint_pos, joint_vel, body_pos_w, body_quat_w, body_lin_vel_w, body_ang_vel_w)]
        N, D = tensors[0].shape
        Bsim = tensors[2].shape[1]
        if N != self.num_envs or D != self.cfg.robot.num_joints:
            raise RuntimeError(f"expected joint_pos ({self.num_envs}, {self.cfg.robot.num_joints}), got {(N, D)}")
        stride = 0
        if policy_obs is not None:
            if policy_obs.dtype != torch.float32 or policy_obs.stride(-1) != 1:
                raise RuntimeError("policy_obs must be float32 with unit inner stride")
            stride = policy_obs.stride(0)
        lib, stream = _lib.enter(dev)
        _lib.check(
            lib.amp_obs_step(
                *[_lib.ptr(t) for t in tensors], N, D, Bsim, self.ref_body_index, _lib.ptr(self._key_idx_host),
                len(self.key_body_indexes), self.cfg.num_amp_observations, _lib.ptr(self.amp_observation_buffer),
                _lib.ptr(policy_obs), stride, stream,
            )  # fmt: skip
        )
        self.extras = {"amp_obs": self.amp_observation_buffer.view(-1, self.amp_observation_size)}
        return self.extras["amp_obs"]

    # ---- policy observation (reference _get_observations :195-242) ---------------------------------
    def get_observations(self, joint_pos, joint_vel, body_pos_w, body_quat_w, body_lin_vel_w, body_ang_vel_w, out: Optional[torch.Tensor] = None) -> dict:
        """The whole ``_get_observations``: AMP history update (``extras["amp_obs"]``) and ``{"policy": actor_obs}``.

        Uses ``self.last_actions``, ``self.command_target_speed`` (when ``cfg.rew_track_vel > 0``), the actor history
        buffer and ``self._just_reset_mask`` exactly as the reference does; two kernel launches in total."""
        self.update_amp_observations(joint_pos, joint_vel, body_pos_w, body_quat_w, body_lin_vel_w, body_ang_vel_w)
        cfg = self.cfg
        n = cfg.num_actor_observations
        width = cfg.observation_space
        if out is None:
            out = torch.empty((self.num_envs, width), dtype=torch.float32, device=self.device)
        if out.dtype != torch.float32 or out.stride(-1) != 1 or out.shape[-1] < width:
            raise RuntimeError("actor observation output must be float32 rows of at least observation_space columns")
        cmd = self.command_target_speed if cfg.command_size else None
        hist = self.actor_obs_history_buffer if n > 1 else None
        mask = self._just_reset_mask if n > 1 else None
        actions = _f32c(self.last_actions, self.device)
        lib, stream = _lib.enter(self.device)
        _lib.check(
            lib.amp_actor_obs_step(
                _lib.ptr(self.amp_observation_buffer), self.num_envs, cfg.num_amp_observations, cfg.amp_observation_space,
                cfg.amp_observation_space - self.key_body_obs_size, _lib.ptr(actions), cfg.action_space, _lib.ptr(cmd),
                cfg.command_size, n, int(cfg.history_include_last_actions), int(cfg.history_include_command), _lib.ptr(hist),
                _lib.ptr(mask), _lib.ptr(out), out.stride(0), stream,
            )  # fmt: skip
        )
        return {"policy": out[:, :width]}

    # ---- the whole per-step path in one launch -------------------------------------------------------
    def step_observations(self, joint_pos, joint_vel, body_pos_w, body_quat_w, body_lin_vel_w, body_ang_vel_w, out: Optional[torch.Tensor] = None,
                          reward_inputs: Optional[dict] = None) -> dict:
        """``_get_observations`` (AMP history + actor observation, reference ``:175-242``) as ONE kernel launch
        (``amp_env_step``); same results as :meth:`get_observations`.  ``reward_inputs`` (optional dict with
        ``reset_terminated, actions, soft_joint_pos_limits, joint_acc``, optionally ``return_terms``) adds
        ``_get_rewards`` (``:246-319``) on the SAME state to the launch -- valid when no reset happens between the two in
        the caller's step; the result dict then carries ``"reward"`` (and ``"reward_terms"``, ``"track_err"``)."""
        cfg, dev = self.cfg, self.device
        tensors = [_f32c(t, dev) for t in (joint_pos, joint_vel, body_pos_w, body_quat_w, body_lin_vel_w, body_ang_vel_w)]
        N, D = tensors[0].shape
        if N != self.num_envs or D != cfg.robot.num_joints:
            raise RuntimeError(f"expected joint_pos ({self.num_envs}, {cfg.robot.num_joints}), got {(N, D)}")
        n, width = cfg.num_actor_observations, cfg.observation_space
        if out is None:
            out = torch.empty((N, width), dtype=torch.float32, device=dev)
        if out.dtype != torch.float32 or out.stride(-1) != 1 or out.shape[-1] < width:
            raise RuntimeError("actor observation output must be float32 rows of at least observation_space columns")
        a = _lib.EnvStepArgs()
        keep = []  # tensors the argument block points at must outlive the call

        def dptr(t):
            keep.append(t)
            return None if t is None else (t.data_ptr() if isinstance(t, torch.Tensor) else t.ctypes.data)

        (a.joint_pos, a.joint_vel, a.body_pos_w, a.body_quat_w, a.body_lin_vel_w, a.body_ang_vel_w) = [dptr(t) for t in tensors]
        a.num_envs, a.num_dofs, a.num_sim_bodies = N, D, tensors[2].shape[1]
        a.ref_body, a.num_key_bodies, a.key_bodies = self.ref_body_index, len(self.key_body_indexes), dptr(self._key_idx_host)
        a.num_amp_observations, a.amp_buf = cfg.num_amp_observations, dptr(self.amp_observation_buffer)
        a.last_actions = dptr(_f32c(self.last_actions, dev))
        a.command = dptr(self.command_target_speed if cfg.command_size else None)
        a.action_size, a.command_size, a.num_actor_observations = cfg.action_space, cfg.command_size, n
        a.hist_include_actions, a.hist_include_command = int(cfg.history_include_last_actions), int(cfg.history_include_command)
        a.hist_buf = dptr(self.actor_obs_history_buffer if n > 1 else None)
        a.just_reset = dptr(self._just_reset_mask if n > 1 else None)
        a.actor_obs, a.actor_stride = dptr(out), out.stride(0)
        result = {}
        if reward_inputs is not None:
            scales = np.array([cfg.rew_termination, cfg.rew_action_l2, cfg.rew_joint_pos_limits, cfg.rew_joint_acc_l2,
                               cfg.rew_joint_vel_l2, cfg.rew_track_vel], dtype=np.float32)  # fmt: skip
            a.reward_scales = dptr(scales)
            a.reset_terminated = dptr(reward_inputs["reset_terminated"].to(dev).to(torch.uint8).contiguous())
            a.actions = dptr(_f32c(reward_inputs["actions"], dev))
            a.soft_limits = dptr(_f32c(reward_inputs["soft_joint_pos_limits"], dev))
            a.joint_acc = dptr(_f32c(reward_inputs["joint_acc"], dev))
            if cfg.rew_track_vel > 0.0:
                a.command = dptr(_f32c(self.command_target_speed, dev))
            result["reward"] = torch.empty(N, dtype=torch.float32, device=dev)
            a.reward_total = dptr(result["reward"])
            if reward_inputs.get("return_terms"):
                result["reward_terms"] = torch.empty((N, 6), dtype=torch.float32, device=dev)
                result["track_err"] = torch.empty(N, dtype=torch.float32, device=dev)
                a.reward_terms, a.track_err = dptr(result["reward_terms"]), dptr(result["track_err"])
        lib, stream = _lib.enter(dev)
        _lib.check(lib.amp_env_step(C.byref(a), stream))
        self.extras = {"amp_obs": self.amp_observation_buffer.view(-1, self.amp_observation_size)}
        result["policy"] = out[:, :width]
        return result

    # ---- task reward (reference _get_rewards :246-319) ----------------------------------------------
    def get_rewards(self, reset_terminated, actions, joint_pos, soft_joint_pos_limits, joint_acc, joint_vel, body_lin_vel_w=None,
                    body_quat_w=None, return_terms: bool = False):
        """``basic_reward + rew_track_vel`` in one kernel.  ``return_terms=True`` also returns the ``(N, 6)`` per-term tensor
        (termination, action_l2, joint_pos_limits, joint_acc_l2, joint_vel_l2, track_vel) and the tracking error, from
        which the reference's ``extras["log"]`` means follow without a host sync per term."""
        cfg, dev = self.cfg, self.device
        term = reset_terminated.to(dev).to(torch.uint8).contiguous()
        args = [_f32c(t, dev) for t in (actions, joint_pos, soft_joint_pos_limits, joint_acc, joint_vel)]
        N, D = args[1].shape
        scales = np.array([cfg.rew_termination, cfg.rew_action_l2, cfg.rew_joint_pos_limits, cfg.rew_joint_acc_l2,
                           cfg.rew_joint_vel_l2, cfg.rew_track_vel], dtype=np.float32)  # fmt: skip
        track = cfg.rew_track_vel > 0.0
        lin = _f32c(body_lin_vel_w, dev) if track else None
        quat = _f32c(body_quat_w, dev) if track else None
        cmd = _f32c(self.command_target_speed, dev) if track else None
        total = torch.empty(N, dtype=torch.float32, device=dev)
        terms = torch.empty((N, 6), dtype=torch.float32, device=dev) if return_terms else None
        err = torch.empty(N, dtype=torch.float32, device=dev) if return_terms else None
        lib, stream = _lib.enter(dev)
        _lib.check(
            lib.amp_task_reward(
                _lib.ptr(scales), _lib.ptr(term), _lib.ptr(args[0]), args[0].shape[1], _lib.ptr(args[1]), _lib.ptr(args[2]),
                _lib.ptr(args[3]), _lib.ptr(args[4]), D, _lib.ptr(lin), _lib.ptr(quat), quat.shape[1] if track else 0,
                self.ref_body_index, _lib.ptr(cmd), N, _lib.ptr(total), _lib.ptr(terms), _lib.ptr(err), stream,
            )  # fmt: skip
        )
        return (total, terms, err) if return_terms else total

    # ---- reset (reference _reset_strategy_random :371-419, AMP part) -------------------------------------------------
    def reset_amp_history(self, env_ids, times, motion_ids) -> None:
        """Reset envs receive their reference history: fused collect + scatter into ``amp_observation_buffer[env_ids]``."""
        loader = self._motion_loader
        t = loader._times_to_device(times)
        n = t.numel()
        ids = loader._ids_to_device(motion_ids, n)
        rows = torch.as_tensor(env_ids, device=self.device, dtype=torch.int64).contiguous().view(-1)
        if rows.numel() != n:
            raise RuntimeError("env_ids and times must have the same length")
        self._launch_collect(t, ids, n, self.amp_observation_buffer.view(self.num_envs, -1), row_index=rows)

    def reset_strategy_random(self, env_ids, default_root_state: torch.Tensor, env_origins: torch.Tensor, start: bool = False):
        """Reference ``_reset_strategy_random`` (``g1_amp_env.py:371-419``) without the simulator writes and the command
        resampling: draws (motion id, time) per env with the reference's host RNG stream, samples the frame, builds
        ``root_state (n, 13)`` (position + env origin, z lifted by 0.05, rotation, linear / angular velocity of the
        robot's reset root body: ``pelvis`` for G1, ``torso`` lifted by 0.15 for the 28-DoF humanoid,
        ``humanoid_amp_env.py:194-201``), the robot-order dof state, and fills ``amp_observation_buffer[env_ids]`` with the
        reference history (fused collect + scatter).  Returns ``(root_state, dof_pos, dof_vel, motion_ids, times)``.
        A clip without that body raises ``AssertionError`` like the reference's ``get_body_index``.

        The root/dof assembly is a handful of torch index ops on the sampled tensors (reset-time glue);
        sampling and the history fill run in ``libamp_b200.so``."""
        loader = self._motion_loader
        ids_dev = torch.as_tensor(env_ids, device=self.device, dtype=torch.int64).view(-1)
        n = ids_dev.numel()
        motion_ids, times = loader.sample_times(n, start=start)
        dof_p, dof_v, body_p, body_r, body_lv, body_av = loader.sample(num_samples=n, times=times, motion_ids=motion_ids)
        robot = self.cfg.robot
        torso = loader.get_body_index([robot.reset_root_body])[0]
        root_state = default_root_state.to(self.device, torch.float32).clone()
        root_state[:, 0:3] = body_p[:, torso] + env_origins.to(self.device, torch.float32)
        root_state[:, 2] += robot.reset_root_lift
        root_state[:, 3:7] = body_r[:, torso]
        root_state[:, 7:10] = body_lv[:, torso]
        root_state[:, 10:13] = body_av[:, torso]
        self.reset_amp_history(ids_dev, times, motion_ids)
        self.last_actions[ids_dev] = 0.0
        if self.cfg.num_actor_observations > 1:
            self._just_reset_mask[ids_dev] = True
        return root_state, dof_p[:, self.motion_dof_indexes], dof_v[:, self.motion_dof_indexes], motion_ids, times

    def poll_flags(self) -> int:
        lib, stream = _lib.enter(self.device)
        flags = C.c_uint32(0)
        _lib.check(lib.amp_lib_poll_flags(self._handle.raw, stream, C.byref(flags)))
        return flags.value
