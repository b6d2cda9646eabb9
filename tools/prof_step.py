#!/usr/bin/env python
"""Tiny driver for ncu: builds the default bench workload (G1_walk shape, n samples x K history) and runs a few steps of
collect_reference_motions + style reward.  Used only for profiling captures (profiles/), never for reported numbers."""

from __future__ import annotations

import argparse
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import humanoid_amp_b200 as amp  # noqa: E402
from humanoid_amp_b200.synthetic import skrl_style_discriminator_params  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="refill_1m")
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    spec = bench.WORKLOADS[a.workload]
    n, K = a.samples or spec["n"], spec["K"]
    dev = torch.device("cuda", 0)
    with tempfile.TemporaryDirectory() as tmp:
        files = bench.make_clip_files(tmp, spec["clip"])
        loader = amp.MotionLoader(files, dev)
    robot = amp.robot_for_clip(loader.dof_names)
    env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file=files, num_envs=1, num_amp_observations=K, robot=robot), dev, motion_loader=loader)
    width = K * robot.amp_observation_space
    ids_h, times_h = bench.host_inputs(loader.durations, n, 1234)
    stats = env.collect_reference_motions(4096, times_h[:4096], ids_h[:4096])
    W, b = skrl_style_discriminator_params(width, seed=42, logit_gain=5.0)
    disc = amp.AmpDiscriminator(width, device=dev, max_rows=n * spec["reward_mult"])
    disc.load(W, b, stats.double().mean(0), stats.double().var(0) + 1e-4)
    times_d, ids_d = torch.from_numpy(times_h).to(dev), torch.from_numpy(ids_h).to(dev)
    obs = torch.empty((n, width), device=dev)
    rows = obs if spec["reward_mult"] == 1 else torch.randn((n * spec["reward_mult"], width), device=dev)
    reward = torch.empty(rows.shape[0], device=dev)
    for _ in range(a.steps):
        env.collect_reference_motions(n, times_d, ids_d, out=obs)
        disc.style_reward(rows, out=reward)
    torch.cuda.synchronize()
    print("ok", float(reward.mean()))


if __name__ == "__main__":
    main()
