"""The packed clip cache (host code, no GPU): a ``.ampclip`` file must give back exactly what
the reference's loader leaves behind for the same ``.npz`` sources (``motions/motion_loader.py:98-164``)."""

from __future__ import annotations

import os
import shutil

import numpy as np
import pytest

from conftest import CLIP_NAMES, clip_path
from humanoid_amp_b200 import clip_cache as cc
from oracle import OracleMotionLoader


def _same(a: cc.ClipArrays, b: cc.ClipArrays):
    assert a.dof_names == b.dof_names and a.body_names == b.body_names
    assert float(a.dt) == float(b.dt)
    for k in ("traj_starts", "traj_ends", "durations"):
        assert np.array_equal(getattr(a, k), getattr(b, k)) and getattr(a, k).dtype == getattr(b, k).dtype
    assert a.shapes == b.shapes and a.offsets == b.offsets
    for k in cc.TENSOR_KEYS:
        assert np.array_equal(a.array(k).view(np.uint32), b.array(k).view(np.uint32)), k


@pytest.mark.parametrize("name", CLIP_NAMES)
def test_round_trip_is_bit_exact(name, tmp_path):
    clip = cc.load_npz_clips([clip_path(name)])
    path = cc.write_clip_cache(str(tmp_path / f"{name}.ampclip"), clip, [clip_path(name)])
    back = cc.read_clip_cache(path)
    _same(clip, back)
    assert cc.cache_sources(path)[0][0] == f"{name}.npz"
    # and it is what the oracle (pinned to the live reference loader) holds
    ora = OracleMotionLoader([clip_path(name)])
    assert np.array_equal(back.array("body_rotations"), np.asarray(ora.body_rotations))
    assert np.array_equal(back.durations, ora.durations) and float(back.dt) == float(ora.dt)


def test_pooled_clips_and_cache_directory(tmp_path):
    files = [clip_path(n) for n in ("humanoid_walk", "humanoid_run", "humanoid_dance")]
    cache_dir = str(tmp_path / "cache")
    first, how1 = cc.load_clips(files, cache_dir)
    second, how2 = cc.load_clips(files, cache_dir)
    assert (how1, how2) == ("cache-miss", "cache-hit")
    _same(first, second)
    assert first.num_trajectories == 3 and list(first.traj_starts) == [0] + list(np.cumsum([e - s + 1 for s, e in zip(first.traj_starts, first.traj_ends)])[:-1])
    # a different order is a different clip set (names / dt come from the first file)
    other, how3 = cc.load_clips(files[::-1], cache_dir)
    assert how3 == "cache-miss" and len(os.listdir(cache_dir)) == 2
    # a single .ampclip path is read directly
    direct, how4 = cc.load_clips([cc.cache_path_for(files, cache_dir)])
    assert how4 == "ampclip"
    _same(first, direct)
    # no cache directory: plain npz load
    assert cc.load_clips(files)[1] == "npz"


def test_changed_source_invalidates(tmp_path):
    src = str(tmp_path / "clip.npz")
    shutil.copy(clip_path("humanoid_run"), src)
    cache_dir = str(tmp_path / "cache")
    assert cc.load_clips([src], cache_dir)[1] == "cache-miss"
    assert cc.load_clips([src], cache_dir)[1] == "cache-hit"
    shutil.copy(clip_path("humanoid_walk"), src)  # same path, different content (size / mtime change)
    os.utime(src, ns=(1, 1))
    clip, how = cc.load_clips([src], cache_dir)
    assert how == "cache-miss"
    assert clip.num_frames == cc.load_npz_clips([clip_path("humanoid_walk")]).num_frames


def test_corruption_is_detected(tmp_path):
    clip = cc.load_npz_clips([clip_path("humanoid_run")])
    path = cc.write_clip_cache(str(tmp_path / "c.ampclip"), clip)
    raw = bytearray(open(path, "rb").read())

    def write(b, name):
        p = str(tmp_path / name)
        open(p, "wb").write(b)
        return p

    bad = bytearray(raw); bad[0:8] = b"NOTACLIP"
    with pytest.raises(cc.ClipCacheError, match="magic"):
        cc.read_clip_cache(write(bad, "magic.ampclip"))
    bad = bytearray(raw); bad[8] = 9
    with pytest.raises(cc.ClipCacheError, match="version"):
        cc.read_clip_cache(write(bad, "version.ampclip"))
    with pytest.raises(cc.ClipCacheError):
        cc.read_clip_cache(write(raw[:-100], "short.ampclip"))
    bad = bytearray(raw); bad[-5] ^= 0x40
    with pytest.raises(cc.ClipCacheError, match="checksum"):
        cc.read_clip_cache(write(bad, "flip.ampclip"))
    assert cc.read_clip_cache(write(bad, "flip2.ampclip"), verify=False).num_frames == clip.num_frames
    # an unusable cache file in the cache directory is ignored and rewritten
    cache_dir = str(tmp_path / "cache")
    files = [clip_path("humanoid_run")]
    cc.load_clips(files, cache_dir)
    open(cc.cache_path_for(files, cache_dir), "wb").write(b"garbage")
    assert cc.load_clips(files, cache_dir)[1] == "cache-miss"
    assert cc.load_clips(files, cache_dir)[1] == "cache-hit"
