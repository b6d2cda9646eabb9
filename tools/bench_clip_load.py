"""Host-side load time of a large clip pool: the reference's .npz path against the packed .ampclip cache.

    python tools/bench_clip_load.py [--clips 60] [--frames 2850]     # ~171 k frames of G1 shape, like the deploy pools

Runs without a GPU (times the host work: unzip + concatenate + narrow, versus one readinto); with a GPU it also times the
MotionLoader constructor end to end.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=60)
    ap.add_argument("--frames", type=int, default=2850)
    a = ap.parse_args()
    import torch

    from humanoid_amp_b200 import clip_cache as cc
    from humanoid_amp_b200.synthetic import write_synthetic_clip

    with tempfile.TemporaryDirectory() as tmp:
        files = []
        for i in range(a.clips):
            p = os.path.join(tmp, f"clip_{i:03d}.npz")
            write_synthetic_clip(p, "G1_dance", frames=a.frames, seed=i)
            files.append(p)
        cache_dir = os.path.join(tmp, "cache")
        t0 = time.perf_counter(); clip = cc.load_npz_clips(files); t_npz = time.perf_counter() - t0
        t0 = time.perf_counter(); path = cc.write_clip_cache(cc.cache_path_for(files, cache_dir), clip, files); t_write = time.perf_counter() - t0
        t0 = time.perf_counter(); back = cc.read_clip_cache(path); t_read = time.perf_counter() - t0
        t0 = time.perf_counter(); cc.read_clip_cache(path, verify=False); t_read_nocheck = time.perf_counter() - t0
        assert np.array_equal(back.arena, clip.arena)
        out = {"frames": clip.num_frames, "clips": a.clips, "arena_mb": clip.arena.size / 1e6, "npz_load_s": t_npz,
               "cache_write_s": t_write, "cache_read_s": t_read, "cache_read_noverify_s": t_read_nocheck,
               "speedup_host": t_npz / t_read}
        if torch.cuda.is_available():
            import humanoid_amp_b200 as amp

            spec = os.path.join(tmp, "clip_*.npz")
            amp.MotionLoader(files[0], "cuda:0")  # context + library warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter(); amp.MotionLoader(spec, "cuda:0"); torch.cuda.synchronize(); out["loader_npz_s"] = time.perf_counter() - t0
            t0 = time.perf_counter(); amp.MotionLoader(path, "cuda:0"); torch.cuda.synchronize(); out["loader_ampclip_s"] = time.perf_counter() - t0
        print(json.dumps(out))


if __name__ == "__main__":
    main()
