#!/usr/bin/env python
"""Per-entry-point micro-benchmark: achieved algorithmic GB/s (or TFLOP/s) of every kernel of the path at sizes larger than
the 126 MB L2, CUDA events on the launching stream, 3 warm-ups + median of 10.  Output goes to profiles/."""

from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import humanoid_amp_b200 as amp  # noqa: E402
from humanoid_amp_b200.synthetic import skrl_style_discriminator_params, synthetic_sim_state  # noqa: E402

PEAKS = bench.measured_peaks()
DEV = torch.device("cuda", 0)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def row(name, ms, nbytes=None, flops=None, note=""):
    out = {"kernel": name, "ms": round(ms, 4), "note": note}
    if nbytes is not None:
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.update(algorithmic_MB=round(nbytes / 1e6, 1), GBps=round(gbs, 1), hbm_frac=round(gbs / PEAKS["hbm_gbs"], 3))
    if flops is not None:
        tf = flops / (ms * 1e-3) / 1e12
        out.update(GFLOP=round(flops / 1e9, 1), TFLOPs=round(tf, 1), tensor_frac_sustained=round(tf / PEAKS["bf16_sustained"], 3))
    print(json.dumps(out), flush=True)


def main():
    tmp = tempfile.TemporaryDirectory()
    loaders = {c: amp.MotionLoader(bench.make_clip_files(tmp.name, c), DEV) for c in ("G1_walk", "G1_dance", "pooled_humanoid")}

    # ---- frame_blend ------------------------------------------------------------------------------------------------
    ld = loaders["G1_walk"]
    S = 8_000_000
    ids_h, t_h = bench.host_inputs(ld.durations, S, 1)
    t_d, i_d = torch.from_numpy(t_h).to(DEV), torch.from_numpy(ids_h).to(DEV)
    ms = timed(lambda: ld.compute_frame_blend_device(t_d, i_d))
    row("frame_blend_kernel", ms, S * (16 + 8 + 8 + 4), note=f"S={S}, f64 index math, i64 idx0/idx1 + f32 blend out")

    # ---- sample_full (all bodies) -------------------------------------------------------------------------------------
    for clip, S in (("G1_dance", 400_000), ("G1_walk", 1_500_000)):
        ld = loaders[clip]
        ids_h, t_h = bench.host_inputs(ld.durations, S, 2)
        t_d, i_d = torch.from_numpy(t_h).to(DEV), torch.from_numpy(ids_h).to(DEV)
        per = (2 * ld.num_dofs + 13 * ld.num_bodies) * 4
        ms = timed(lambda: ld.sample(S, times=t_d, motion_ids=i_d))
        row("sample_full_kernel", ms, S * (per + 16), note=f"{clip} S={S} ({per} B/frame out, incl. torch.empty of 6 outputs)")

    # ---- fused collect -------------------------------------------------------------------------------------------------
    for clip, n, K in (("G1_walk", 1_000_000, 2), ("G1_dance", 100_000, 10), ("pooled_humanoid", 1_000_000, 2), ("G1_walk", 4096, 2), ("G1_dance", 4096, 10)):
        ld = loaders[clip]
        robot = amp.robot_for_clip(ld.dof_names)
        env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file="", num_envs=1, num_amp_observations=K, robot=robot), DEV, motion_loader=ld)
        ids_h, t_h = bench.host_inputs(ld.durations, n, 3)
        t_d, i_d = torch.from_numpy(t_h).to(DEV), torch.from_numpy(ids_h).to(DEV)
        A = robot.amp_observation_space
        out = torch.empty((n, K * A), device=DEV)
        ms = timed(lambda: env.collect_reference_motions(n, t_d, i_d, out=out))
        nbytes = n * K * A * 4 + n * 16 + ld.num_frames * ((A + 3) // 4 * 4) * 4
        row("collect_reference_kernel", ms, nbytes, note=f"{clip} n={n} K={K} (table {ld.num_frames * ((A + 3) // 4 * 4) * 4 / 1e3:.0f} KB)")

    # ---- env step --------------------------------------------------------------------------------------------------------
    for n, K in ((1_000_000, 2), (65_536, 10), (4096, 10)):
        ld = loaders["G1_dance"]
        env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file="", num_envs=n, num_amp_observations=K, robot=amp.G1), DEV, motion_loader=ld)
        state = synthetic_sim_state(n, amp.G1, DEV, seed=5)
        ms = timed(lambda: env.update_amp_observations(*state))
        per_env = (2 * 29 + 25) * 4 + (K - 1) * 83 * 4 + K * 83 * 4
        row("obs_step_kernel", ms, n * per_env, note=f"G1 N={n} K={K} ({per_env} B/env)")
        del env, state

    # ---- actor observation + task reward ------------------------------------------------------------------
    for n, n_actor in ((65_536, 2), (1_000_000, 1)):
        ld = loaders["G1_dance"]
        cfg = amp.AmpEnvCfg(motion_file="", num_envs=n, num_amp_observations=2, robot=amp.G1, num_actor_observations=n_actor, rew_track_vel=1.0,
                            rew_termination=-1.0, rew_action_l2=-0.1, rew_joint_pos_limits=-10.0, rew_joint_acc_l2=-1e-6, rew_joint_vel_l2=-1e-3)
        env = amp.AmpEnvPath(cfg, DEV, motion_loader=ld)
        state = synthetic_sim_state(n, amp.G1, DEV, seed=6)
        out = torch.empty((n, cfg.observation_space), device=DEV)
        env.update_amp_observations(*state)
        lib_call = lambda: env.get_observations(*state, out=out)  # noqa: E731
        ms_both = timed(lib_call)
        ms_amp = timed(lambda: env.update_amp_observations(*state))
        P, cur = cfg.hist_frame_size, 71 + 29 + 2
        nbytes = n * ((cur + max(n_actor - 2, 0) * P) * 4 + ((n_actor - 1) * P + cfg.observation_space) * 4)
        row("actor_obs_kernel", max(ms_both - ms_amp, 1e-4), nbytes, note=f"G1 N={n} n_actor={n_actor} (get_observations minus the AMP-history launch)")
        g = torch.Generator(device="cuda").manual_seed(1)
        acts, acc = torch.randn(n, 29, device=DEV, generator=g), torch.randn(n, 29, device=DEV, generator=g)
        lim = torch.randn(n, 29, 2, device=DEV, generator=g)
        term = torch.zeros(n, dtype=torch.bool, device=DEV)
        ms = timed(lambda: env.get_rewards(term, acts, state[0], lim, acc, state[1], state[4], state[3]))
        row("task_reward_kernel", ms, n * ((29 * 6 + 9) * 4 + 1 + 4), note=f"G1 N={n}, velocity tracking on")
        # the three per-step kernels in ONE launch (amp_env_step): every simulator tensor read once
        K = cfg.num_amp_observations
        rin = dict(reset_terminated=term, actions=acts, soft_joint_pos_limits=lim, joint_acc=acc)
        ms_three = ms_both + ms
        ms = timed(lambda: env.step_observations(*state, out=out, reward_inputs=rin))
        step_bytes = n * (((2 * 29 + 25) + (K - 1) * 83 + K * 83) * 4) + nbytes + n * ((29 * 4 + 9) * 4 + 1 + 4)
        row("env_step_kernel (obs_step + actor_obs + task_reward fused)", ms, step_bytes,
            note=f"G1 N={n} K={K} n_actor={n_actor}; the three separate launches take {ms_three:.4f} ms")
        del env, state, out

    # ---- the per-step env path at the reference's real scale (4096 envs): one fused launch vs three, graph replay, cold L2 ------
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

    def timed_graph(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = amp.capture_step(fn, DEV)
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    for n, K, n_actor in ((4096, 2, 2), (4096, 10, 2)):
        ld = loaders["G1_dance"]
        cfg = amp.AmpEnvCfg(motion_file="", num_envs=n, num_amp_observations=K, robot=amp.G1, num_actor_observations=n_actor, rew_track_vel=1.0,
                            rew_termination=-1.0, rew_action_l2=-0.1, rew_joint_pos_limits=-10.0, rew_joint_acc_l2=-1e-6, rew_joint_vel_l2=-1e-3)
        env = amp.AmpEnvPath(cfg, DEV, motion_loader=ld)
        state = synthetic_sim_state(n, amp.G1, DEV, seed=6)
        out = torch.empty((n, cfg.observation_space), device=DEV)
        g = torch.Generator(device="cuda").manual_seed(1)
        acts, acc = torch.randn(n, 29, device=DEV, generator=g), torch.randn(n, 29, device=DEV, generator=g)
        lim = torch.randn(n, 29, 2, device=DEV, generator=g)
        term = torch.zeros(n, dtype=torch.uint8, device=DEV)
        rin = dict(reset_terminated=term, actions=acts, soft_joint_pos_limits=lim, joint_acc=acc)

        def three():
            env.get_observations(*state, out=out)
            env.get_rewards(term, acts, state[0], lim, acc, state[1], state[4], state[3])

        ms3 = timed_graph(three)
        ms1 = timed_graph(lambda: env.step_observations(*state, out=out, reward_inputs=rin))
        P = cfg.hist_frame_size
        nbytes = n * (((2 * 29 + 25) + (K - 1) * 83 + K * 83) * 4 + (71 + 29 + 2 + max(n_actor - 2, 0) * P) * 4 + ((n_actor - 1) * P + cfg.observation_space) * 4 + (29 * 4 + 9) * 4 + 5)
        row("env step at 4096 envs: ONE launch (amp_env_step), graph replay, L2 flushed", ms1, nbytes, note=f"G1 N={n} K={K} n_actor={n_actor}")
        row("env step at 4096 envs: three launches (obs_step + actor_obs + task_reward), graph replay, L2 flushed", ms3, nbytes, note=f"G1 N={n} K={K} n_actor={n_actor}")
        del env, state, out

    # ---- compute_obs free function ---------------------------------------------------------------------------------------
    n = 2_000_000
    g = torch.Generator(device="cuda").manual_seed(0)
    args = [torch.randn(n, 29, device=DEV, generator=g), torch.randn(n, 29, device=DEV, generator=g), torch.randn(n, 3, device=DEV, generator=g),
            torch.nn.functional.normalize(torch.randn(n, 4, device=DEV, generator=g), dim=-1), torch.randn(n, 3, device=DEV, generator=g),
            torch.randn(n, 3, device=DEV, generator=g), torch.randn(n, 4, 3, device=DEV, generator=g)]  # fmt: skip
    ms = timed(lambda: amp.compute_obs(*args))
    row("compute_obs_kernel", ms, n * (83 * 4 + 83 * 4 + 4), note=f"n={n}, incl. torch.empty")
    del args

    # ---- discriminator ---------------------------------------------------------------------------------------------------
    for width, M in ((166, 1_000_000), (830, 262_144), (166, 65_536), (830, 65_536), (166, 4096)):
        W, b = skrl_style_discriminator_params(width, seed=42, logit_gain=5.0)
        disc = amp.AmpDiscriminator(width, device=DEV, max_rows=M)
        disc.load(W, b, torch.zeros(width, dtype=torch.float64), torch.ones(width, dtype=torch.float64))
        x = torch.randn(M, width, device=DEV)
        r = torch.empty(M, device=DEV)
        ms = timed(lambda: disc.style_reward(x, out=r))
        row("disc style reward (cast + fused tcgen05)", ms, flops=M * bench.flops_per_row(width), note=f"K*A={width} M={M}")
        del disc, x

    # ---- AMP memories + state preprocessor ----------------------------------------------------------------
    for width, cap, M in ((166, 2_000_000, 1_000_000), (830, 400_000, 262_144), (166, 200_000, 65_536)):
        mem = amp.AmpStateMemory(cap, width, DEV)
        mem.add_samples(torch.randn(cap, width, device=DEV))
        idx = mem.sample_indexes(M, generator=torch.Generator(device="cuda").manual_seed(3))
        out = torch.empty((M, width), device=DEV)
        ms = timed(lambda: mem.sample_by_index(idx, out=out))
        row("gather_rows_kernel", ms, M * (2 * width * 4 + 8), note=f"RandomMemory.sample_by_index W={width} capacity={cap} M={M}")
        sc = amp.RunningStandardScaler(width, device=DEV)
        ms = timed(lambda: sc.update(out))
        row("scaler_partial+merge_kernel", ms, M * width * 4, note=f"RunningStandardScaler train update W={width} M={M}")
        res = torch.empty_like(out)
        ms = timed(lambda: sc(out, out=res))
        row("scaler_apply_kernel", ms, 2 * M * width * 4, note=f"RunningStandardScaler eval W={width} M={M}")
        W, b = skrl_style_discriminator_params(width, seed=42, logit_gain=5.0)
        disc = amp.AmpDiscriminator(width, device=DEV, max_rows=M)
        disc.load(W, b, sc.running_mean, sc.running_variance)
        r = torch.empty(M, device=DEV)
        ms = timed(lambda: disc.style_reward_sampled(mem.states, idx, out=r))
        row("disc style reward on sampled rows (gather fused into the cast)", ms, flops=M * bench.flops_per_row(width),
            note=f"K*A={width} M={M} from capacity {cap}")
        del mem, out, res, disc, sc
    tmp.cleanup()


if __name__ == "__main__":
    main()
