"""Oracle restatement of the env-side AMP observation path (TEST INFRASTRUCTURE -- see ``oracle/__init__.py``).

Follows ``/root/reference/g1_amp_env.py`` (the 28-DoF humanoid env ``humanoid_amp_env.py:219-286`` is the same math):

* ``quaternion_to_tangent_and_normal`` ... ``:489-497``
* ``compute_obs`` ........................ ``:535-561``
* ``collect_reference_motions`` .......... ``:445-486``
* history shift + slot-0 write ........... ``:187-193``
* reset-time history fill ................ ``:414-419``
* actor observation + its history ........ ``:195-242`` (the first "next" row)
* task reward ............................ ``:246-288, 500-532, 564-606`` (same row)

``quat_apply`` is upstream ``isaaclab.utils.math.quat_apply`` (Isaac Lab 2.2.0, not vendored; call sites
``g1_amp_env.py:16, 495-496``): restated from its published form, PARITY UNPINNED for that function alone.
"""

from __future__ import annotations

import numpy as np
import torch


def quat_apply(quat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """Rotate ``vec`` by the wxyz quaternion ``quat``: ``v + w*t + xyz x t`` with ``t = 2 * (xyz x v)``."""
    shape = vec.shape
    quat = quat.reshape(-1, 4)
    vec = vec.reshape(-1, 3)
    xyz = quat[:, 1:]
    t = xyz.cross(vec, dim=-1) * 2
    return (vec + quat[:, 0:1] * t + xyz.cross(t, dim=-1)).view(shape)


def quaternion_to_tangent_and_normal(q: torch.Tensor) -> torch.Tensor:
    """Reference ``g1_amp_env.py:489-497``: rotate x-hat (tangent) and z-hat (normal) by ``q``, concatenate."""
    x_hat = torch.zeros_like(q[..., :3])
    z_hat = torch.zeros_like(q[..., :3])
    x_hat[..., 0] = 1
    z_hat[..., -1] = 1
    return torch.cat([quat_apply(q, x_hat), quat_apply(q, z_hat)], dim=q.ndim - 1)


def compute_obs(dof_pos, dof_vel, root_pos, root_rot, root_lin_vel, root_ang_vel, key_body_pos) -> torch.Tensor:
    """Reference ``g1_amp_env.py:545-561``.

    Row layout (A = 2D + 13 + 3*Kb):
      [dof_pos D | dof_vel D | root_z 1 | tangent 3 | normal 3 | root_lin_vel 3 (world) | root_ang_vel 3 (world) |
       (key_pos - root_pos) 3*Kb (world axes)]
    """
    rel = (key_body_pos - root_pos.unsqueeze(-2)).view(key_body_pos.shape[0], -1)
    return torch.cat(
        [dof_pos, dof_vel, root_pos[:, 2:3], quaternion_to_tangent_and_normal(root_rot), root_lin_vel, root_ang_vel, rel],
        dim=-1,
    )


def history_times(current_times, dt, num_amp_observations):
    """Reference ``g1_amp_env.py:454-457``: ``t - dt*k`` for k = 0..K-1, flattened sample-major (slot 0 = newest).

    Times are NOT clamped: negative values extrapolate, values past the clip end hold the last frame.
    """
    return (np.expand_dims(current_times, axis=-1) - dt * np.arange(0, num_amp_observations)).flatten()


def collect_reference_motions(
    loader,
    num_samples,
    num_amp_observations,
    dof_indexes,
    ref_body_index,
    key_body_indexes,
    current_times=None,
    motion_ids=None,
):
    """Reference ``g1_amp_env.py:445-486``.  ``loader`` is any object with the reference ``MotionLoader`` surface."""
    if current_times is None:
        motion_ids, current_times = loader.sample_times(num_samples)
    times = history_times(current_times, loader.dt, num_amp_observations)
    if motion_ids is not None:
        ids = np.repeat(motion_ids, num_amp_observations)
    else:
        ids = np.zeros_like(times, dtype=np.int32)
    dof_p, dof_v, body_p, body_r, body_lv, body_av = loader.sample(num_samples=num_samples, times=times, motion_ids=ids)
    obs = compute_obs(
        dof_p[:, dof_indexes],
        dof_v[:, dof_indexes],
        body_p[:, ref_body_index],
        body_r[:, ref_body_index],
        body_lv[:, ref_body_index],
        body_av[:, ref_body_index],
        body_p[:, key_body_indexes],
    )
    return obs.view(-1, num_amp_observations * obs.shape[-1])


def shift_and_write_history(buffer: torch.Tensor, obs: torch.Tensor) -> torch.Tensor:
    """Reference ``g1_amp_env.py:187-193``: slot i -> i+1 from the oldest down, then slot 0 = obs.  In place.

    Returns the ``(N, K*A)`` view the reference exposes as ``extras["amp_obs"]``.
    """
    k = buffer.shape[1]
    for i in reversed(range(k - 1)):
        buffer[:, i + 1] = buffer[:, i]
    buffer[:, 0] = obs.clone()
    return buffer.view(-1, k * buffer.shape[2])


def reset_fill(buffer: torch.Tensor, env_ids: torch.Tensor, amp_observations: torch.Tensor) -> None:
    """Reference ``g1_amp_env.py:417-419``: reset envs receive the reference history rows.  In place."""
    buffer[env_ids] = amp_observations.view(env_ids.shape[0], buffer.shape[1], -1)


def actor_observations(obs, last_actions, command, history_buffer, just_reset_mask, num_actor_observations, key_body_obs_size=12,
                       history_include_last_actions=True, history_include_command=True):
    """Reference ``g1_amp_env.py:195-242`` (policy observation of ``_get_observations``).

    ``obs`` is this step's ``compute_obs`` row; ``command`` is ``None`` when ``cfg.rew_track_vel <= 0``.  With history
    (``num_actor_observations > 1``) ``history_buffer`` is ``(N, n-1, P)`` and ``just_reset_mask`` ``(N,) bool``; both are
    updated in place exactly like the reference: warm start of just-reset envs, shift, slot 0 = the new history frame.
    """
    base = obs[:, :-key_body_obs_size]
    if num_actor_observations > 1:
        current_parts = [base, last_actions]
        if command is not None:
            current_parts.append(command)
        current_frame = torch.cat(current_parts, dim=-1)
        hist_parts = [base]
        if history_include_last_actions:
            hist_parts.append(last_actions)
        if history_include_command and command is not None:
            hist_parts.append(command)
        hist_frame = torch.cat(hist_parts, dim=-1)
        if just_reset_mask.any():
            for i in range(num_actor_observations - 1):
                history_buffer[just_reset_mask, i] = hist_frame[just_reset_mask]
            just_reset_mask[:] = False
        for i in reversed(range(num_actor_observations - 2)):
            history_buffer[:, i + 1] = history_buffer[:, i]
        history_buffer[:, 0] = hist_frame
        return torch.cat([current_frame, history_buffer.view(obs.shape[0], -1)], dim=-1)
    actor_obs = torch.cat([base, last_actions], dim=-1)
    if command is not None:
        actor_obs = torch.cat([actor_obs, command], dim=-1)
    return actor_obs


def reset_root_and_dof_state(loader, times, motion_ids, default_root_state, env_origins, motion_dof_indexes, torso_index, lift=0.05):
    """Reference ``g1_amp_env.py:386-411`` (state part of ``_reset_strategy_random``): root pose / velocity and dof state of
    the sampled frame; the root is lifted to avoid ground collisions (G1: body ``pelvis``, 0.05, ``:398-405``; the 28-DoF
    humanoid: body ``torso``, 0.15, ``humanoid_amp_env.py:194-201``)."""
    dof_p, dof_v, body_p, body_r, body_lv, body_av = loader.sample(num_samples=len(times), times=times, motion_ids=motion_ids)
    root_state = default_root_state.clone()
    root_state[:, 0:3] = body_p[:, torso_index] + env_origins
    root_state[:, 2] += lift
    root_state[:, 3:7] = body_r[:, torso_index]
    root_state[:, 7:10] = body_lv[:, torso_index]
    root_state[:, 10:13] = body_av[:, torso_index]
    return root_state, dof_p[:, motion_dof_indexes], dof_v[:, motion_dof_indexes]


def quat_rotate_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Upstream ``isaaclab.utils.math.quat_rotate_inverse`` (Isaac Lab 2.2.0, not vendored; call site ``g1_amp_env.py:253``),
    restated from its published form -- PARITY UNPINNED: ``v*(2w^2-1) - 2w (xyz x v) + 2 xyz (xyz . v)``."""
    q_w = q[..., 0]
    q_vec = q[..., 1:]
    a = v * (2.0 * q_w**2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(q.shape[0], 1, 3), v.view(q.shape[0], 3, 1)).squeeze(-1) * 2.0
    return a - b + c


def exp_reward_with_floor(error: torch.Tensor, weight: float, sigma: float, floor: float = 3.0) -> torch.Tensor:
    """Reference ``g1_amp_env.py:500-532``: exponential reward that continues linearly (C1) beyond ``floor * sigma^2``.
    The threshold constants are Python floats (double) in the scripted reference and enter the tensor ops as scalars."""
    import math

    sigma_sq = sigma * sigma
    threshold = floor * sigma_sq
    exp_val_at_threshold = weight * math.exp(-floor)
    linear_slope = weight / sigma_sq * math.exp(-floor)
    linear_reward = exp_val_at_threshold - linear_slope * (error - threshold)
    exp_reward = weight * torch.exp(-error / sigma_sq)
    return torch.where(error > threshold, linear_reward, exp_reward)


def task_rewards(scales, reset_terminated, actions, joint_pos, soft_joint_pos_limits, joint_acc, joint_vel, root_lin_vel_w=None,
                 root_quat_w=None, command=None):
    """Reference ``_get_rewards`` (``g1_amp_env.py:246-288``) + ``compute_rewards`` (``:564-606``).

    ``scales`` = dict(rew_termination, rew_action_l2, rew_joint_pos_limits, rew_joint_acc_l2, rew_joint_vel_l2, rew_track_vel).
    Returns ``(total, terms)`` with ``terms`` an ``(N, 6)`` tensor in that order plus the tracking error ``(N,)``.
    """
    n = actions.shape[0]
    if scales["rew_track_vel"] > 0.0:
        speed = quat_rotate_inverse(root_quat_w, root_lin_vel_w)[:, :2]
        track_err = torch.norm(speed - command, dim=-1)
        rew_track = exp_reward_with_floor(torch.square(track_err), scales["rew_track_vel"], 0.5, floor=4.0)
    else:
        track_err = torch.zeros(n)
        rew_track = torch.zeros(n, dtype=torch.float)
    r_term = scales["rew_termination"] * reset_terminated.float()
    r_act = scales["rew_action_l2"] * torch.sum(torch.square(actions), dim=1)
    out_of_limits = -(joint_pos - soft_joint_pos_limits[:, :, 0]).clip(max=0.0)
    out_of_limits += (joint_pos - soft_joint_pos_limits[:, :, 1]).clip(min=0.0)
    r_lim = scales["rew_joint_pos_limits"] * torch.sum(out_of_limits, dim=1)
    r_acc = scales["rew_joint_acc_l2"] * torch.sum(torch.square(joint_acc), dim=1)
    r_vel = scales["rew_joint_vel_l2"] * torch.sum(torch.square(joint_vel), dim=1)
    basic = r_term + r_act + r_lim + r_acc + r_vel
    total = basic + rew_track
    return total, torch.stack([r_term, r_act, r_lim, r_acc, r_vel, rew_track], dim=1), track_err
