"""Host-side logic that needs no GPU: motion-file resolution (reference ``motion_loader.py:14-84``), robot tables, env
sharding, the synthetic clip generator and the CPU-only pieces of the bench harness."""

from __future__ import annotations

import os

import numpy as np
import pytest
import yaml

from conftest import GOLDEN, clip_path
from humanoid_amp_b200 import G1, HUMANOID28, robot_for_clip, shard_envs
from humanoid_amp_b200.motion_loader import _resolve_motion_files
from humanoid_amp_b200.synthetic import CLIP_SHAPES, synthetic_clip_arrays, write_synthetic_clip


def test_resolve_single_comma_glob_dir(tmp_path):
    a, b = clip_path("humanoid_walk"), clip_path("humanoid_run")
    assert _resolve_motion_files(a) == [a]
    assert _resolve_motion_files(f"{a}, {b}") == [a, b]
    assert _resolve_motion_files(f"{a},/nonexistent.npz") == [a]  # missing entries of a comma list are dropped
    clips = os.path.join(GOLDEN, "clips")
    assert _resolve_motion_files(os.path.join(clips, "humanoid_*.npz")) == sorted(
        os.path.join(clips, f"humanoid_{n}.npz") for n in ("dance", "run", "walk")
    )
    assert len(_resolve_motion_files(clips)) == 8
    with pytest.raises(ValueError):
        _resolve_motion_files(str(tmp_path / "nothing*.npz"))
    with pytest.raises(ValueError):
        _resolve_motion_files(str(tmp_path / "missing.npz"))


def test_resolve_yaml(tmp_path):
    clips = os.path.join(GOLDEN, "clips")
    cfg = tmp_path / "motions.yaml"
    cfg.write_text(yaml.safe_dump({"motion_files": [clip_path("G1_walk"), "rel_missing.npz"]}))
    assert _resolve_motion_files(str(cfg)) == [clip_path("G1_walk")]
    cfg.write_text(yaml.safe_dump({"glob_pattern": os.path.join(clips, "G1_*.npz")}))
    assert len(_resolve_motion_files(str(cfg))) == 4
    cfg.write_text(yaml.safe_dump({"motion_files": []}))
    with pytest.raises(ValueError):
        _resolve_motion_files(str(cfg))
    # relative entries resolve against the config's directory
    write_synthetic_clip(str(tmp_path / "local.npz"), "humanoid_run", frames=8)
    cfg.write_text(yaml.safe_dump({"motion_files": ["local.npz"]}))
    assert _resolve_motion_files(str(cfg)) == [str(tmp_path / "local.npz")]


def test_robot_tables(kat):
    assert G1.amp_observation_space == 83 and HUMANOID28.amp_observation_space == 81
    assert len(G1.body_names) == 39 and len(G1.joint_names) == 29
    with np.load(clip_path("G1_walk")) as d:
        names = d["dof_names"].tolist()
    assert robot_for_clip(names) is G1
    assert [names.index(j) for j in G1.joint_names] == kat["g1_walk"]["dof_perm"]
    with np.load(clip_path("humanoid_walk")) as d:
        assert robot_for_clip(d["dof_names"].tolist()) is HUMANOID28
        assert tuple(d["body_names"].tolist()) == HUMANOID28.body_names
    with np.load(clip_path("G1_dance")) as d:
        assert tuple(d["body_names"].tolist()) == G1.body_names


def test_shard_envs_partitions_exactly():
    for total, world in ((4096, 8), (65536, 8), (10, 3), (5, 8)):
        spans = [shard_envs(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_envs(8, 8, 8)


@pytest.mark.parametrize("name", sorted(CLIP_SHAPES))
def test_synthetic_clip_has_reference_format(name):
    arr = synthetic_clip_arrays(name, seed=1, frames=24)
    spec = CLIP_SHAPES[name]
    D, B = len(spec["dofs"]), len(spec["bodies"])
    assert arr["dof_positions"].shape == (24, D) and arr["dof_positions"].dtype == spec["dof_dtype"]
    assert arr["body_rotations"].shape == (24, B, 4) and arr["body_rotations"].dtype == np.float32
    assert np.allclose(np.linalg.norm(arr["body_rotations"], axis=-1), 1.0, atol=1e-5)
    assert int(arr["fps"]) == 60
    # the oracle loads it like a shipped clip
    robot = spec["robot"]
    assert set(arr["dof_names"].tolist()) == set(robot.joint_names)
    for key in robot.key_body_names + (robot.reference_body,):
        assert key in arr["body_names"].tolist()


def test_oracle_random_memory_ring_and_split():
    """The memory oracle itself: wrap-around writes, ``filled``, and ``np.array_split`` mini-batches (skrl semantics)."""
    import torch

    from oracle import OracleRandomMemory

    mem = OracleRandomMemory(5, 2)
    mem.add_samples(torch.arange(6, dtype=torch.float32).view(3, 2))
    assert len(mem) == 3 and not mem.filled and mem.memory_index == 3
    mem.add_samples(torch.arange(6, 14, dtype=torch.float32).view(4, 2))
    assert len(mem) == 5 and mem.filled and mem.memory_index == 2
    assert mem.states[:, 0].tolist() == [10.0, 12.0, 4.0, 6.0, 8.0]
    parts = mem.sample_by_index([0, 1, 2, 3, 4, 0, 1], mini_batches=3)
    assert [p.shape[0] for p in parts] == [3, 2, 2]
    assert parts[0][:, 0].tolist() == [10.0, 12.0, 4.0]


def test_memory_and_scaler_need_cuda():
    import torch

    import humanoid_amp_b200 as amp

    if torch.cuda.is_available():
        import pytest

        pytest.skip("CPU-only check")
    for make in (lambda: amp.AmpStateMemory(8, 4, "cpu"), lambda: amp.RunningStandardScaler(4, device="cpu"),
                 lambda: amp.InputPrefetcher("cpu", 8)):
        try:
            make()
        except amp.AmpB200Error as e:
            assert "no CPU" in str(e)
        else:
            raise AssertionError("constructed without a CUDA device")
