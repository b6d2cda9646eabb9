#!/usr/bin/env python
"""ncu driver: the fused collect kernel on the default bench shape (G1_walk, 1 M samples x K = 2) -- profiling only."""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import humanoid_amp_b200 as amp  # noqa: E402

dev = torch.device("cuda", 0)
n, K = 1_000_000, 2
with tempfile.TemporaryDirectory() as tmp:
    ld = amp.MotionLoader(bench.make_clip_files(tmp, "G1_walk"), dev)
    env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file="", num_envs=1, num_amp_observations=K, robot=amp.G1), dev, motion_loader=ld)
    ids_h, t_h = bench.host_inputs(ld.durations, n, 2)
    t_d, i_d = torch.from_numpy(t_h).to(dev), torch.from_numpy(ids_h).to(dev)
    out = torch.empty((n, K * 83), device=dev)
    for _ in range(4):
        env.collect_reference_motions(n, t_d, i_d, out=out)
torch.cuda.synchronize()
print("ok")
