"""``MotionLoader`` fed from the packed clip cache (``-m gpu``): identical tensors and identical
samples whichever way the clips arrived."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import clip_path, pooled_spec

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("spec", ["G1_dance", "pooled"])
def test_loader_from_cache_equals_loader_from_npz(spec, tmp_path, monkeypatch):
    import humanoid_amp_b200 as amp
    from humanoid_amp_b200 import clip_cache as cc
    from humanoid_amp_b200.motion_loader import _resolve_motion_files

    motion_file = pooled_spec() if spec == "pooled" else clip_path(spec)
    plain = amp.MotionLoader(motion_file, DEV)
    assert plain.load_path == "npz"
    cache_dir = str(tmp_path / "cache")
    miss = amp.MotionLoader(motion_file, DEV, cache_dir=cache_dir)
    hit = amp.MotionLoader(motion_file, DEV, cache_dir=cache_dir)
    monkeypatch.setenv("AMP_B200_CLIP_CACHE", cache_dir)
    env_hit = amp.MotionLoader(motion_file, DEV)
    packed = amp.MotionLoader(cc.cache_path_for(_resolve_motion_files(motion_file), cache_dir), DEV)
    assert (miss.load_path, hit.load_path, env_hit.load_path, packed.load_path) == ("cache-miss", "cache-hit", "cache-hit", "ampclip")

    rng = np.random.default_rng(3)
    ids = rng.integers(0, plain.num_trajectories, 4096)
    times = rng.uniform(0, 1, 4096) * plain.durations[ids]
    want = plain.sample(4096, times=times, motion_ids=ids)
    for other in (miss, hit, env_hit, packed):
        assert other.dof_names == plain.dof_names and other.body_names == plain.body_names
        assert float(other.dt) == float(plain.dt) and other.num_frames == plain.num_frames
        assert np.array_equal(other.durations, plain.durations) and np.array_equal(other.traj_ends, plain.traj_ends)
        for key in ("dof_positions", "dof_velocities", "body_positions", "body_rotations", "body_linear_velocities", "body_angular_velocities"):
            a, b = getattr(other, key), getattr(plain, key)
            assert a.shape == b.shape and a.is_contiguous() and torch.equal(a, b), key
        got = other.sample(4096, times=times, motion_ids=ids)
        for g, w in zip(got, want):
            assert torch.equal(g, w)
