#!/usr/bin/env python
"""Developer tool: a few launches of the fused env-step kernel (amp_env_step) at N envs, for `ncu -k regex:env_step_kernel`.

    python tools/prof_env_step.py [--envs 1000000] [--K 2] [--actor 1] [--iters 5]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1_000_000)
    ap.add_argument("--K", type=int, default=2)
    ap.add_argument("--actor", type=int, default=1)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    import bench
    import humanoid_amp_b200 as amp
    from humanoid_amp_b200.synthetic import synthetic_sim_state

    dev = torch.device("cuda", 0)
    import tempfile

    tmp = tempfile.TemporaryDirectory()
    ld = amp.MotionLoader(bench.make_clip_files(tmp.name, "G1_dance"), dev)
    n = a.envs
    cfg = amp.AmpEnvCfg(motion_file="", num_envs=n, num_amp_observations=a.K, robot=amp.G1, num_actor_observations=a.actor, rew_track_vel=1.0,
                        rew_termination=-1.0, rew_action_l2=-0.1, rew_joint_pos_limits=-10.0, rew_joint_acc_l2=-1e-6, rew_joint_vel_l2=-1e-3)
    env = amp.AmpEnvPath(cfg, dev, motion_loader=ld)
    state = synthetic_sim_state(n, amp.G1, dev, seed=6)
    out = torch.empty((n, cfg.observation_space), device=dev)
    g = torch.Generator(device="cuda").manual_seed(1)
    acts, acc = torch.randn(n, 29, device=dev, generator=g), torch.randn(n, 29, device=dev, generator=g)
    lim = torch.randn(n, 29, 2, device=dev, generator=g)
    term = torch.zeros(n, dtype=torch.bool, device=dev)
    rin = dict(reset_terminated=term, actions=acts, soft_joint_pos_limits=lim, joint_acc=acc)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
    for i in range(a.iters):
        ev[i].record()
        env.step_observations(*state, out=out, reward_inputs=rin)
    ev[a.iters].record()
    torch.cuda.synchronize()
    print("ms per launch:", [round(ev[i].elapsed_time(ev[i + 1]), 4) for i in range(a.iters)])


if __name__ == "__main__":
    main()
