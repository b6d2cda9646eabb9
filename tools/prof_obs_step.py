#!/usr/bin/env python
"""Tiny driver for ncu: the per-step AMP observation kernel (obs_step_kernel) on synthetic simulator state.
Used only for profiling captures (profiles/), never for reported numbers."""
import argparse
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import humanoid_amp_b200 as amp  # noqa: E402
from humanoid_amp_b200.synthetic import synthetic_sim_state  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1_000_000)
ap.add_argument("--K", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
with tempfile.TemporaryDirectory() as tmp:
    loader = amp.MotionLoader(bench.make_clip_files(tmp, "G1_walk"), dev)
env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file="", num_envs=a.envs, num_amp_observations=a.K, robot=amp.G1), dev, motion_loader=loader)
state = synthetic_sim_state(a.envs, amp.G1, dev, seed=5)
start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    env.update_amp_observations(*state)
start.record()
for _ in range(a.steps):
    env.update_amp_observations(*state)
end.record()
torch.cuda.synchronize()
per_env = (2 * 29 + 25) * 4 + (a.K - 1) * 83 * 4 + a.K * 83 * 4
ms = start.elapsed_time(end) / a.steps
print(f"obs_step N={a.envs} K={a.K}: {ms:.4f} ms, {a.envs * per_env / ms / 1e6:.1f} GB/s algorithmic")
