"""Offline dataset pipeline on the GPU (``-m gpu``; reference motions/data_convert.py:161-379) against the oracle, through the
C ABI (amp_dataset_interp_fk / amp_dataset_velocities).

Tolerances (the reference's dtypes: joint arrays float64, body arrays float32):
  * joint positions:            bit-exact (scipy interp1d's expression evaluated in the same order, float64, no FMA)
  * joint velocities:           1e-12 relative (float64 differences + the 9-tap gaussian in scipy's summation order)
  * body positions / rotations: north_star's fp32 bar, 1e-6 + 1e-5 |ref| (float64 FK rounded once; cos/sin differ from
                                numpy's by an ulp of float64)
  * body linear velocities:     stage-wise (same positions in) bit-exact; end to end 1 ulp of position / (2 dt) = 1e-5
  * body angular velocities:    the kernel evaluates compute_angular_velocity in float64 on the float32 rotations; against the
                                oracle doing the same it is held to 1e-4 rad/s (the raw values pass through float32), and the
                                reference's float32 evaluation must lie within its own conditioning noise (5e-2 rad/s, see
                                tests/test_dataset_oracle.py) of it.
"""

from __future__ import annotations

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
DS = os.path.join(HERE, "golden", "dataset")
URDF = os.path.join(DS, "g1_29dof_kinematic.urdf")
CSV = os.path.join(DS, "walk1_rows_110_265.csv")


def close32(got, want):
    return bool((np.abs(got.astype(np.float64) - want) <= 1e-6 + 1e-5 * np.abs(want)).all())


def check_against_oracle(rows):
    from humanoid_amp_b200 import dataset
    from oracle import dataset_oracle as do

    tree = dataset.UrdfTree(URDF, dataset.JOINT_NAMES)
    got = dataset.convert_rows(rows, tree, device="cuda:0")
    want = do.convert(rows, do.load_urdf_tree(URDF, do.G1_JOINT_NAMES))
    for k in ("dof_positions", "dof_velocities", "body_positions", "body_rotations", "body_linear_velocities", "body_angular_velocities"):
        assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
    assert int(got["fps"]) == 60 and list(got["dof_names"]) == list(want["dof_names"]) and list(got["body_names"]) == list(want["body_names"])
    assert np.array_equal(got["dof_positions"], want["dof_positions"])
    assert np.allclose(got["dof_velocities"], want["dof_velocities"], rtol=1e-12, atol=1e-12)
    assert close32(got["body_positions"], want["body_positions"])
    assert close32(got["body_rotations"], want["body_rotations"])
    assert np.abs(got["body_linear_velocities"] - want["body_linear_velocities"]).max() <= 1e-5
    # stage-wise: the oracle's own poses in -> the velocity stages alone
    dt = 1.0 / 60
    dv, lv, av = dataset.velocities(torch.from_numpy(want["dof_positions"]).cuda(), torch.from_numpy(want["body_positions"]).cuda(),
                                    torch.from_numpy(want["body_rotations"]).cuda(), dt)
    assert np.array_equal(lv.cpu().numpy(), want["body_linear_velocities"])
    assert np.allclose(dv.cpu().numpy(), want["dof_velocities"], rtol=1e-12, atol=1e-12)
    _, _, exact = do.velocity_stages(want["dof_positions"], want["body_positions"], want["body_rotations"], dt, exact_angular=True)
    assert np.abs(av.cpu().numpy() - exact).max() <= 1e-4
    assert np.abs(av.cpu().numpy() - want["body_angular_velocities"]).max() <= 5e-2
    return got, want


def test_shipped_walk_slice_vs_oracle_and_vs_the_reference_output():
    rows = np.loadtxt(CSV, delimiter=",", dtype=np.float64, ndmin=2).astype(np.float32)
    got, _ = check_against_oracle(rows)
    # and against the reference's OWN output of this conversion (motions/custom_motion.npz, made with the real Pinocchio)
    ref = np.load(os.path.join(HERE, "golden", "clips_full", "custom_motion.npz"))
    assert close32(got["body_positions"], ref["body_positions"]) and close32(got["body_rotations"], ref["body_rotations"])
    assert np.abs(got["dof_positions"] - ref["dof_positions"]).max() <= 1e-7
    assert np.abs(got["body_linear_velocities"] - ref["body_linear_velocities"]).max() <= 1e-5
    assert np.abs(got["body_angular_velocities"] - ref["body_angular_velocities"]).max() <= 5e-2


@pytest.mark.parametrize("n_in", [2, 3, 5, 700])
def test_synthetic_motion_and_short_clips(n_in):
    """Smooth random motion (large root rotations: every branch of the matrix -> quaternion conversion, both branches of the
    small-angle series) and clips shorter than the gaussian's 9 taps (reflect boundary wraps several times)."""
    rng = np.random.default_rng(n_in)
    t = np.linspace(0, n_in / 30, n_in)
    rows = np.zeros((n_in, 36), dtype=np.float64)
    rows[:, 0:3] = np.stack([0.5 * t, 0.2 * np.sin(t), 0.8 + 0.05 * np.cos(3 * t)], axis=1)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = 2.5 * np.sin(1.3 * t) + (1e-5 * t if n_in == 700 else 0.0)  # sweeps through pi/2..: negative traces too
    rows[:, 3:6] = axis[None, :] * np.sin(ang / 2)[:, None]
    rows[:, 6] = np.cos(ang / 2)
    rows[n_in // 2 :, 3:7] *= -1.0  # the double cover: Slerp must take the short way
    rows[:, 7:] = 0.6 * np.sin(t[:, None] * rng.uniform(0.5, 3.0, 29)[None, :] + rng.uniform(0, 6, 29)[None, :])
    check_against_oracle(rows.astype(np.float32))


def test_argument_checks():
    from humanoid_amp_b200 import AmpB200Error, dataset

    tree = dataset.UrdfTree(URDF, dataset.JOINT_NAMES)
    with pytest.raises(ValueError, match="at least two"):
        dataset.convert_rows(np.zeros((1, 36), dtype=np.float32), tree, device="cuda:0")
    with pytest.raises(ValueError, match="joint columns"):
        dataset.convert_rows(np.zeros((4, 30), dtype=np.float32), tree, device="cuda:0")
    with pytest.raises(AmpB200Error):
        dataset.convert_rows(np.zeros((4, 36), dtype=np.float32), tree, device="cpu")


def test_command_line_writes_a_clip_the_motion_loader_reads(tmp_path):
    """``python -m humanoid_amp_b200.dataset`` keeps the reference tool's command line (data_convert.py:133-158) and writes the
    same ``.npz`` layout: the file loads into the product ``MotionLoader`` and into the oracle loader with identical tensors."""
    import humanoid_amp_b200 as amp
    from humanoid_amp_b200 import dataset
    from oracle import OracleMotionLoader

    out = str(tmp_path / "clip.npz")
    dataset.main(["--csv", CSV, "--urdf", URDF, "--meshes", str(tmp_path), "--output", out, "--start", "10", "--end", "90"])
    d = np.load(out)
    assert set(d.files) == {"fps", "dof_names", "body_names", "dof_positions", "dof_velocities", "body_positions", "body_rotations",
                            "body_linear_velocities", "body_angular_velocities"}  # fmt: skip
    assert d["dof_positions"].shape == (159, 29) and d["body_rotations"].shape == (159, 25, 4) and int(d["fps"]) == 60
    assert d["dof_positions"].dtype == np.float64 and d["body_positions"].dtype == np.float32
    loader = amp.MotionLoader(out, "cuda:0")
    ora = OracleMotionLoader([out])
    assert loader.num_frames == ora.num_frames == 159 and loader.dof_names == list(ora.dof_names)
    assert torch.equal(loader.body_positions.cpu(), ora.body_positions) and torch.equal(loader.dof_positions.cpu(), ora.dof_positions)
    # the G1 env path runs on the converted clip (pelvis reference body, the four key bodies are among the 25 recorded links)
    env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file=out, num_envs=8, num_amp_observations=2, robot=amp.G1), "cuda:0", motion_loader=loader)
    rows = env.collect_reference_motions(64)
    assert rows.shape == (64, 2 * 83) and bool(torch.isfinite(rows).all())


def test_random_urdf_tree_with_oblique_axes_and_rotated_origins(tmp_path):
    """The shipped G1 URDF only has axis-aligned joints and two rotated origins.  A random tree (oblique unit axes, rpy origins,
    fixed and continuous joints, branches) exercises the general Rodrigues / rpy paths of the kernel against the oracle."""
    from humanoid_amp_b200 import dataset
    from oracle import dataset_oracle as do

    rng = np.random.default_rng(11)
    n_rev, links, lines = 9, ["base"], ['<?xml version="1.0"?>', '<robot name="rnd">', '  <link name="base"/>']
    joint_names, bodies = [], ["base"]
    for j in range(14):
        parent = links[int(rng.integers(0, len(links)))]
        child = f"l{j}"
        kind = "fixed" if j % 3 == 2 else ("continuous" if j % 5 == 0 else "revolute")
        xyz = " ".join(f"{v:.4f}" for v in rng.uniform(-0.3, 0.3, 3))
        rpy = " ".join(f"{v:.4f}" for v in rng.uniform(-1.5, 1.5, 3))
        axis = rng.normal(size=3)
        lines += [f'  <link name="{child}"/>', f'  <joint name="j{j}" type="{kind}">', f'    <origin xyz="{xyz}" rpy="{rpy}"/>',
                  f'    <parent link="{parent}"/>', f'    <child link="{child}"/>']  # fmt: skip
        if kind != "fixed":
            lines.append(f'    <axis xyz="{axis[0]:.5f} {axis[1]:.5f} {axis[2]:.5f}"/>')  # NOT normalised: both sides normalise
            joint_names.append(f"j{j}")
        lines.append("  </joint>")
        links.append(child)
        bodies.append(child)
    lines.append("</robot>")
    urdf = tmp_path / "rnd.urdf"
    urdf.write_text("\n".join(lines))
    D = len(joint_names)
    assert D >= n_rev - 2
    n_in = 60
    t = np.linspace(0, 2, n_in)
    rows = np.zeros((n_in, 7 + D))
    rows[:, 0:3] = rng.normal(size=(1, 3)) + 0.3 * np.sin(t)[:, None]
    q = rng.normal(size=(n_in, 4)) * 0.05 + np.array([0.3, -0.5, 0.2, 0.8])
    rows[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    rows[:, 7:] = 1.5 * np.sin(t[:, None] * rng.uniform(0.5, 3, D)[None, :] + rng.uniform(0, 6, D)[None, :])
    rows = rows.astype(np.float32)
    got = dataset.convert_rows(rows, dataset.UrdfTree(str(urdf), joint_names), body_names=bodies, joint_names=joint_names, device="cuda:0")
    want = do.convert(rows, do.load_urdf_tree(str(urdf), joint_names), joint_names=joint_names, body_names=bodies)
    assert np.array_equal(got["dof_positions"], want["dof_positions"])
    assert close32(got["body_positions"], want["body_positions"]) and close32(got["body_rotations"], want["body_rotations"])
    assert np.abs(got["body_linear_velocities"] - want["body_linear_velocities"]).max() <= 2e-5
