import os, sys, tempfile, torch
sys.path.insert(0, '/root/repo')
import bench, humanoid_amp_b200 as amp
dev = torch.device("cuda", 0)
with tempfile.TemporaryDirectory() as tmp:
    for clip, n, K in (("G1_walk", 1_000_000, 2), ("G1_dance", 100_000, 10)):
        ld = amp.MotionLoader(bench.make_clip_files(tmp, clip), dev)
        env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file="", num_envs=1, num_amp_observations=K, robot=amp.G1), dev, motion_loader=ld)
        ids_h, t_h = bench.host_inputs(ld.durations, n, 2)
        t_d, i_d = torch.from_numpy(t_h).to(dev), torch.from_numpy(ids_h).to(dev)
        out = torch.empty((n, K * 83), device=dev)
        for _ in range(3): env.collect_reference_motions(n, t_d, i_d, out=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); env.collect_reference_motions(n, t_d, i_d, out=out); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ts.sort(); print(clip, n, K, "ms", round(ts[len(ts)//2], 4), "checksum", float(out.double().sum()))
