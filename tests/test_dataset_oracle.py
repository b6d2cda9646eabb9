"""Offline dataset pipeline (reference motions/data_convert.py:161-379): the oracle against the reference's
own outputs, and the product's host-side tables against scipy.  CPU only."""

from __future__ import annotations

import os

import numpy as np
import pytest

from oracle import dataset_oracle as do

HERE = os.path.dirname(os.path.abspath(__file__))
DS = os.path.join(HERE, "golden", "dataset")
URDF = os.path.join(DS, "g1_29dof_kinematic.urdf")
CSV = os.path.join(DS, "walk1_rows_110_265.csv")
SHIPPED = os.path.join(HERE, "golden", "clips_full", "custom_motion.npz")  # the reference's own output (real Pinocchio)
REF_TEXT = os.path.join(DS, "data_convert_output.npz")  # the unmodified reference text run on the restated FK
ARRAYS = ("dof_positions", "dof_velocities", "body_positions", "body_rotations", "body_linear_velocities", "body_angular_velocities")


@pytest.fixture(scope="module")
def converted():
    rows = np.loadtxt(CSV, delimiter=",", dtype=np.float64, ndmin=2).astype(np.float32)
    tree = do.load_urdf_tree(URDF, do.G1_JOINT_NAMES)
    return rows, tree, do.convert(rows, tree)


def test_oracle_equals_the_unmodified_reference_text_bit_for_bit(converted):
    """Everything but the forward kinematics (interpolation, differences, gaussian smoothing, angular velocities, the
    .npz layout) is pinned on the reference's own text: tests/golden/make_dataset_golden.py executed motions/data_convert.py
    unmodified, with this oracle's FK standing in for Pinocchio."""
    _, _, out = converted
    ref = np.load(REF_TEXT)
    assert int(out["fps"]) == int(ref["fps"]) == 60
    assert list(out["dof_names"]) == list(ref["dof_names"]) and list(out["body_names"]) == list(ref["body_names"])
    for k in ARRAYS:
        assert out[k].dtype == ref[k].dtype and out[k].shape == ref[k].shape, k
        assert np.array_equal(out[k], ref[k]), k


def test_restated_forward_kinematics_against_the_shipped_reference_output(converted):
    """motions/custom_motion.npz is an output of the reference tool with the real Pinocchio (rows 110:265 of
    datasets/walk1_subject1.csv): the restated FK + Eigen quaternion conversion reproduce its poses to float32 rounding."""
    _, _, out = converted
    ref = np.load(SHIPPED)
    assert list(out["body_names"]) == list(ref["body_names"]) and list(out["dof_names"]) == list(ref["dof_names"])
    assert out["body_positions"].shape == ref["body_positions"].shape == (309, 25, 3)
    assert np.abs(out["body_positions"] - ref["body_positions"]).max() <= 2.4e-7  # 2 float32 ulp at |x| < 2
    assert np.abs(out["body_rotations"] - ref["body_rotations"]).max() <= 1.2e-7  # 2 ulp at |q| <= 1, identical signs
    # the shipped file was interpolated by an older scipy (difference of the float32 samples taken in float64)
    assert np.abs(out["dof_positions"] - ref["dof_positions"]).max() <= 1e-7
    assert np.abs(out["dof_velocities"] - ref["dof_velocities"]).max() <= 1e-5
    assert np.abs(out["body_linear_velocities"] - ref["body_linear_velocities"]).max() <= 1e-5  # 1 ulp of position / (2 dt)


def test_velocity_stages_reproduce_the_shipped_output_from_its_own_poses():
    """Stage-wise pin: fed with the SHIPPED positions / joint angles, the velocity stages give the shipped velocities bit for
    bit.  The angular velocity is the exception and shows why its tolerance is what it is: 2 acos(w) / dt on float32
    quaternions of nearly equal rotations is ill-conditioned, so even the reference's own code run on the reference's own
    rotations differs from the file the reference shipped (another numpy / BLAS build) by ~1e-2 rad/s."""
    ref = np.load(SHIPPED)
    dv, lv, av = do.velocity_stages(ref["dof_positions"], ref["body_positions"], ref["body_rotations"], 1.0 / 60)
    assert np.array_equal(dv, ref["dof_velocities"])
    assert np.array_equal(lv, ref["body_linear_velocities"])
    d = np.abs(av - ref["body_angular_velocities"]).max()
    assert 1e-4 < d < 5e-2, d
    _, _, exact = do.velocity_stages(ref["dof_positions"], ref["body_positions"], ref["body_rotations"], 1.0 / 60, exact_angular=True)
    assert np.abs(exact - ref["body_angular_velocities"]).max() < 5e-2 and np.abs(exact - av).max() < 5e-2


def test_product_host_tables_are_scipys():
    """humanoid_amp_b200.dataset computes the gaussian weights and the interpolation tables on the host: same values as scipy."""
    from scipy.interpolate import interp1d
    from scipy.ndimage import gaussian_filter1d

    from humanoid_amp_b200 import dataset

    w = dataset.gaussian_weights()
    impulse = np.zeros(21)
    impulse[10] = 1.0
    assert np.array_equal(gaussian_filter1d(impulse, sigma=1)[10:15], w)
    for n_in in (2, 3, 155, 1000):
        t_orig, t_new, lo, ind, alpha = dataset.time_tables(n_in)
        assert len(t_new) == 2 * n_in - 1 and lo.min() >= 0 and lo.max() <= n_in - 2 and ind.min() >= 0 and ind.max() <= n_in - 2
        # the expression csrc/amp_dataset.cu evaluates per frame is scipy's (_call_linear), float32 samples included
        y = (np.arange(n_in, dtype=np.float64) ** 2 / 7.0).astype(np.float32)
        want = interp1d(t_orig, y, kind="linear")(t_new)
        x_lo, x_hi = t_orig[lo], t_orig[lo + 1]
        got = ((t_new - x_lo) / (x_hi - x_lo)) * y[lo + 1].astype(np.float64) + ((x_hi - t_new) / (x_hi - x_lo)) * y[lo].astype(np.float64)
        assert want.dtype == np.float64 and np.array_equal(got, want)
        assert np.all((alpha >= 0) & (alpha <= 1))


def test_urdf_tree_matches_the_oracle_tree_and_rejects_bad_input(tmp_path):
    from humanoid_amp_b200 import dataset

    tree = dataset.UrdfTree(URDF, dataset.JOINT_NAMES)
    ora = do.load_urdf_tree(URDF, do.G1_JOINT_NAMES)
    assert tree.root_link == ora.root_link == "pelvis" and tree.names == ora.names and tree.num_actuated == 29
    assert np.array_equal(tree.parent, ora.parent) and np.array_equal(tree.qidx, ora.qidx)
    assert np.array_equal(tree.origin_xyz, ora.origin_xyz) and np.array_equal(tree.axis, ora.axis)
    assert np.array_equal(tree.origin_rot.reshape(-1, 3, 3), ora.origin_rot)
    assert np.array_equal(tree.body_joints(dataset.BODY_NAMES), [ora.link_joint[b] for b in do.G1_BODY_NAMES])
    with pytest.raises(ValueError, match="links not found"):
        tree.body_joints(["no_such_link"])
    with pytest.raises(ValueError, match="no column"):
        dataset.UrdfTree(URDF, dataset.JOINT_NAMES[:-1])
    rows = dataset.read_csv_rows(CSV)
    assert rows.shape == (155, 36) and rows.dtype == np.float32
    try:
        import pandas as pd
    except ImportError:
        return
    assert np.array_equal(rows, pd.read_csv(CSV, header=None).to_numpy(dtype=np.float32))  # data_convert.py:164-178
