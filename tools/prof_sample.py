#!/usr/bin/env python
"""ncu driver: MotionLoader.sample on the G1_dance / G1_walk shapes and the env-step kernel (profiling only)."""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import humanoid_amp_b200 as amp  # noqa: E402
from humanoid_amp_b200.synthetic import synthetic_sim_state  # noqa: E402

dev = torch.device("cuda", 0)
with tempfile.TemporaryDirectory() as tmp:
    for clip, S in (("G1_dance", 400_000), ("G1_walk", 1_500_000)):
        ld = amp.MotionLoader(bench.make_clip_files(tmp, clip), dev)
        ids_h, t_h = bench.host_inputs(ld.durations, S, 2)
        t_d, i_d = torch.from_numpy(t_h).to(dev), torch.from_numpy(ids_h).to(dev)
        for _ in range(2):
            ld.sample(S, times=t_d, motion_ids=i_d)
    env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file="", num_envs=65536, num_amp_observations=10, robot=amp.G1), dev, motion_loader=ld)
    state = synthetic_sim_state(65536, amp.G1, dev, seed=5)
    for _ in range(2):
        env.update_amp_observations(*state)
torch.cuda.synchronize()
print("ok")
