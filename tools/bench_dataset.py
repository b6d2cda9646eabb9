#!/usr/bin/env python
"""Developer benchmark of the offline dataset pipeline: GPU conversion (humanoid_amp_b200.dataset) against the
oracle restatement of motions/data_convert.py on the host, synthetic CSV of the G1 layout.

    python tools/bench_dataset.py [--frames 7840] [--cpu-frames 400]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
URDF = os.path.join(ROOT, "tests", "golden", "dataset", "g1_29dof_kinematic.urdf")


def synthetic_rows(n, seed=0):
    rng = np.random.default_rng(seed)
    t = np.linspace(0, n / 30, n)
    rows = np.zeros((n, 36))
    rows[:, 0:3] = np.stack([0.5 * t, 0.2 * np.sin(t), 0.8 + 0.05 * np.cos(3 * t)], axis=1)
    ang = 0.8 * np.sin(0.7 * t)
    rows[:, 5] = np.sin(ang / 2)
    rows[:, 6] = np.cos(ang / 2)
    rows[:, 7:] = 0.6 * np.sin(t[:, None] * rng.uniform(0.5, 3.0, 29)[None, :] + rng.uniform(0, 6, 29)[None, :])
    return rows.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=7840, help="input frames (the shipped LAFAN1 walk file has 7840)")
    ap.add_argument("--cpu-frames", type=int, default=400)
    args = ap.parse_args()
    from humanoid_amp_b200 import dataset
    from oracle import dataset_oracle as do

    tree = dataset.UrdfTree(URDF, dataset.JOINT_NAMES)
    rows = synthetic_rows(args.frames)
    dataset.convert_rows(rows[:64], tree, device="cuda:0")  # warm-up (module load)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = dataset.convert_rows(rows, tree, device="cuda:0")
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    otree = do.load_urdf_tree(URDF, do.G1_JOINT_NAMES)
    t0 = time.perf_counter()
    do.convert(rows[: args.cpu_frames], otree)
    cpu_s = time.perf_counter() - t0
    n_out = out["dof_positions"].shape[0]
    print(json.dumps({"input_frames": args.frames, "output_frames": n_out, "gpu_s_end_to_end_incl_copies": gpu_s, "gpu_frames_per_s": n_out / gpu_s,
                      "cpu_oracle_frames_per_s": (2 * args.cpu_frames - 1) / cpu_s, "cpu_sample_frames": args.cpu_frames,
                      "speedup": (n_out / gpu_s) / ((2 * args.cpu_frames - 1) / cpu_s)}))


if __name__ == "__main__":
    main()
