"""Offline dataset pipeline on the GPU: CSV motion capture -> the clip ``.npz`` the ``MotionLoader`` reads.

Mirrors the reference tool ``motions/data_convert.py:161-379`` (same command line, same ``.npz`` keys / dtypes / shapes):

    python -m humanoid_amp_b200.dataset --csv walk.csv --urdf g1_29dof_rev_1_0.urdf --meshes meshes/ --output motions/custom_motion.npz \
        [--start 0] [--end N] [--fps 60]

The reference interpolates with scipy, runs Pinocchio's forward kinematics frame by frame in Python and differentiates with
numpy; here the per-frame work (interpolation, FK over the URDF tree, quaternion conversion) and the per-element velocity
stages are CUDA kernels (``csrc/amp_dataset.cu``).  The host side only parses the CSV / URDF and lays out the per-frame time
tables.  There is no CPU path.
"""

from __future__ import annotations

import argparse
import ctypes as C
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib

# data_convert.py:228-258 (joint order of the CSV columns = Pinocchio's q order) and :301-327 (recorded links)
JOINT_NAMES = [
    "left_hip_pitch_joint", "left_hip_roll_joint", "left_hip_yaw_joint", "left_knee_joint", "left_ankle_pitch_joint", "left_ankle_roll_joint",
    "right_hip_pitch_joint", "right_hip_roll_joint", "right_hip_yaw_joint", "right_knee_joint", "right_ankle_pitch_joint", "right_ankle_roll_joint",
    "waist_yaw_joint", "waist_roll_joint", "waist_pitch_joint",
    "left_shoulder_pitch_joint", "left_shoulder_roll_joint", "left_shoulder_yaw_joint", "left_elbow_joint", "left_wrist_roll_joint",
    "left_wrist_pitch_joint", "left_wrist_yaw_joint",
    "right_shoulder_pitch_joint", "right_shoulder_roll_joint", "right_shoulder_yaw_joint", "right_elbow_joint", "right_wrist_roll_joint",
    "right_wrist_pitch_joint", "right_wrist_yaw_joint",
]  # fmt: skip
BODY_NAMES = [
    "pelvis", "head_link", "torso_link", "left_shoulder_pitch_link", "left_shoulder_roll_link", "left_shoulder_yaw_link", "left_elbow_link",
    "right_shoulder_pitch_link", "right_shoulder_roll_link", "right_shoulder_yaw_link", "right_elbow_link", "left_hip_yaw_link",
    "left_hip_roll_link", "left_hip_pitch_link", "left_knee_link", "right_hip_yaw_link", "right_hip_roll_link", "right_hip_pitch_link",
    "right_knee_link", "right_rubber_hand", "left_rubber_hand", "right_ankle_roll_link", "left_ankle_roll_link", "waist_yaw_link",
    "waist_roll_link",
]  # fmt: skip
MAX_JOINTS = 64


class UrdfTree:
    """Kinematic part of a URDF: joints in topological order with parent links resolved to joint indices."""

    def __init__(self, urdf_path: str, joint_names: Sequence[str]):
        robot = ET.parse(urdf_path).getroot()
        joints = []
        for j in robot.findall("joint"):
            origin, axis = j.find("origin"), j.find("axis")
            joints.append({
                "name": j.get("name"), "type": j.get("type"), "parent": j.find("parent").get("link"), "child": j.find("child").get("link"),
                "xyz": [float(v) for v in (origin.get("xyz", "0 0 0") if origin is not None else "0 0 0").split()],
                "rpy": [float(v) for v in (origin.get("rpy", "0 0 0") if origin is not None else "0 0 0").split()],
                "axis": [float(v) for v in (axis.get("xyz") if axis is not None else "1 0 0").split()],
            })  # fmt: skip
        children = {j["child"] for j in joints}
        roots = [link.get("name") for link in robot.findall("link") if link.get("name") not in children]
        if len(roots) != 1:
            raise ValueError(f"the URDF must have exactly one root link, found {roots}")
        root = roots[0]
        # a `floating` joint from a world link to the base is the free-flyer itself (JointModelFreeFlyer in the reference)
        floating = [j for j in joints if j["type"] == "floating" and j["parent"] == root]
        if floating:
            root = floating[0]["child"]
            joints = [j for j in joints if j is not floating[0]]
        self.root_link = root
        order: List[dict] = []
        self.link_joint: Dict[str, int] = {root: -1}
        frontier = [root]
        while frontier:
            link = frontier.pop(0)
            for j in joints:
                if j["parent"] == link:
                    self.link_joint[j["child"]] = len(order)
                    order.append(j)
                    frontier.append(j["child"])
        if len(order) > MAX_JOINTS:
            raise ValueError(f"at most {MAX_JOINTS} URDF joints are supported, found {len(order)}")
        qmap = {n: i for i, n in enumerate(joint_names)}
        J = len(order)
        self.names = [j["name"] for j in order]
        self.parent = np.array([self.link_joint[j["parent"]] for j in order], dtype=np.int32)
        self.qidx = np.full(J, -1, dtype=np.int32)
        self.origin_xyz = np.array([j["xyz"] for j in order], dtype=np.float64).reshape(J, 3)
        self.origin_rot = np.zeros((J, 9), dtype=np.float64)
        self.axis = np.array([j["axis"] for j in order], dtype=np.float64).reshape(J, 3)
        for i, j in enumerate(order):
            r, p, y = j["rpy"]
            cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
            self.origin_rot[i] = [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr, -sp, cp * sr, cp * cr]
            if j["type"] in ("revolute", "continuous"):
                if j["name"] not in qmap:
                    raise ValueError(f"URDF joint {j['name']} has no column in the joint-name list")
                self.qidx[i] = qmap[j["name"]]
                self.axis[i] /= np.linalg.norm(self.axis[i])
            elif j["type"] != "fixed":
                raise ValueError(f"URDF joint type {j['type']!r} is not supported ({j['name']})")
        self.num_actuated = int((self.qidx >= 0).sum())

    def body_joints(self, body_names: Sequence[str]) -> np.ndarray:
        missing = [b for b in body_names if b not in self.link_joint]
        if missing:
            raise ValueError(f"links not found in the URDF: {missing}")
        return np.array([self.link_joint[b] for b in body_names], dtype=np.int32)


def time_tables(n_in: int, fps_orig: int = 30):
    """The per-frame time arrays and knot indices of data_convert.py:186-215, as numpy / scipy derive them:
    ``t_orig``, ``t_new`` (``np.linspace``), interp1d's lower knot, Slerp's interval index and ``alpha``."""
    dt_orig = 1.0 / fps_orig
    t_orig = np.linspace(0, (n_in - 1) * dt_orig, n_in)
    n_out = 2 * n_in - 1
    t_new = np.linspace(0, (n_in - 1) * dt_orig, n_out)
    lerp_lo = (np.clip(np.searchsorted(t_orig, t_new), 1, n_in - 1) - 1).astype(np.int32)  # scipy interp1d._call_linear
    ind = np.searchsorted(t_orig, t_new) - 1  # scipy Slerp.__call__
    ind[t_new == t_orig[0]] = 0
    alpha = (t_new - t_orig[ind]) / np.diff(t_orig)[ind]
    return t_orig, t_new, lerp_lo, ind.astype(np.int32), alpha


def gaussian_weights(sigma: float = 1.0, truncate: float = 4.0) -> np.ndarray:
    """scipy.ndimage ``_gaussian_kernel1d`` (order 0): weights for offsets 0..radius (the kernel is symmetric)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x**2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:], dtype=np.float64)


def convert_rows(rows: np.ndarray, tree: UrdfTree, body_names: Sequence[str] = BODY_NAMES, joint_names: Sequence[str] = JOINT_NAMES,
                 fps: int = 60, device="cuda", return_root_pose: bool = False) -> Dict[str, np.ndarray]:
    """``main`` of data_convert.py on already sliced CSV rows ``(N, 7 + D)`` float32; returns the ``.npz`` dictionary."""
    dev = _lib.require_cuda(device)
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n_in, n_cols = rows.shape
    D = n_cols - 7
    if n_in < 2:
        raise ValueError("need at least two CSV frames")
    if D != len(joint_names):
        raise ValueError(f"the CSV has {D} joint columns, the joint-name list {len(joint_names)}")
    if 7 + D != 7 + tree.num_actuated:  # data_convert.py:321-324 prints a warning; the layouts cannot be reconciled, so refuse
        raise ValueError(f"CSV columns={7 + D}, but the URDF model has nq={7 + tree.num_actuated}")
    t_orig, t_new, lerp_lo, ind, alpha = time_tables(n_in)
    n_out, B = len(t_new), len(body_names)
    dt = 1.0 / fps
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    held = [up(rows), up(t_orig), up(t_new), up(lerp_lo), up(ind), up(alpha), up(tree.parent), up(tree.qidx), up(tree.origin_xyz),
            up(tree.origin_rot), up(tree.axis), up(tree.body_joints(body_names))]  # fmt: skip
    desc = _lib.DatasetDesc(n_in, n_cols, n_out, D, B, len(tree.names), *[t.data_ptr() for t in held])
    dof_pos = torch.empty((n_out, D), dtype=torch.float64, device=dev)
    body_pos = torch.empty((n_out, B, 3), dtype=torch.float32, device=dev)
    body_rot = torch.empty((n_out, B, 4), dtype=torch.float32, device=dev)
    root_pose = torch.empty((n_out, 7), dtype=torch.float64, device=dev) if return_root_pose else None
    lib, stream = _lib.enter(dev)
    scratch = torch.empty(int(lib.amp_dataset_scratch_bytes(n_in, n_out, B)), dtype=torch.uint8, device=dev)
    _lib.check(lib.amp_dataset_interp_fk(C.byref(desc), _lib.ptr(dof_pos), _lib.ptr(body_pos), _lib.ptr(body_rot), _lib.ptr(root_pose),
                                         _lib.ptr(scratch), scratch.numel(), stream))
    out = velocities(dof_pos, body_pos, body_rot, dt, scratch=scratch)
    res = {
        "fps": np.int64(fps), "dof_names": np.array(list(joint_names), dtype=np.str_), "body_names": np.array(list(body_names), dtype=np.str_),
        "dof_positions": dof_pos.cpu().numpy(), "dof_velocities": out[0].cpu().numpy(), "body_positions": body_pos.cpu().numpy(),
        "body_rotations": body_rot.cpu().numpy(), "body_linear_velocities": out[1].cpu().numpy(), "body_angular_velocities": out[2].cpu().numpy(),
    }  # fmt: skip
    if return_root_pose:
        res["root_pose"] = root_pose.cpu().numpy()
    return res


def velocities(dof_positions: torch.Tensor, body_positions: torch.Tensor, body_rotations: torch.Tensor, dt: float,
               scratch: Optional[torch.Tensor] = None):
    """data_convert.py:290-297, :349-371 on device tensors: ``(dof_velocities f64, body_linear_velocities f32,
    body_angular_velocities f32)``."""
    dev = _lib.require_cuda(dof_positions.device)
    dof_positions = dof_positions.to(torch.float64).contiguous()
    body_positions = body_positions.to(torch.float32).contiguous()
    body_rotations = body_rotations.to(torch.float32).contiguous()
    n_out, D = dof_positions.shape
    B = body_positions.shape[1]
    need = n_out * B * 3 * 4
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    dof_vel = torch.empty_like(dof_positions)
    lin_vel = torch.empty_like(body_positions)
    ang_vel = torch.empty_like(body_positions)
    w = gaussian_weights()
    lib, stream = _lib.enter(dev)
    _lib.check(lib.amp_dataset_velocities(n_out, D, B, float(dt), _lib.ptr(w), _lib.ptr(dof_positions), _lib.ptr(body_positions),
                                          _lib.ptr(body_rotations), _lib.ptr(dof_vel), _lib.ptr(lin_vel), _lib.ptr(ang_vel),
                                          _lib.ptr(scratch), scratch.numel(), stream))
    return dof_vel, lin_vel, ang_vel


def read_csv_rows(path: str, start: int = 0, end: Optional[int] = None) -> np.ndarray:
    """``pd.read_csv(path, header=None).iloc[start:end].to_numpy(dtype=np.float32)`` (data_convert.py:164-178) without pandas:
    the values are parsed as float64 and narrowed once, as pandas does."""
    data = np.loadtxt(path, delimiter=",", dtype=np.float64, ndmin=2)
    return data[start:end].astype(np.float32)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="Convert motion data to NPZ for Isaac Lab (GPU build of motions/data_convert.py).")
    ap.add_argument("--csv", type=str, required=True, help="Path to input CSV motion file")
    ap.add_argument("--urdf", type=str, required=True, help="Path to robot URDF file")
    ap.add_argument("--meshes", type=str, required=False, default=None, help="Path to mesh directory (kinematics need no meshes; accepted for compatibility)")
    ap.add_argument("--output", type=str, default="motions/custom_motion.npz", help="Output NPZ filename")
    ap.add_argument("--start", type=int, default=0, help="Start frame index")
    ap.add_argument("--end", type=int, default=None, help="End frame index (default: end of file)")
    ap.add_argument("--fps", type=int, default=60, help="Target FPS (default: 60)")
    ap.add_argument("--device", type=str, default="cuda")
    return ap.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    rows = read_csv_rows(args.csv, args.start, args.end)
    print(f"Loading CSV: {args.csv}, planning to extract frames [{args.start}:{args.end if args.end is not None else 'end'}]")
    print(f"Actual loaded frames: {rows.shape[0]}")
    tree = UrdfTree(args.urdf, JOINT_NAMES)
    data = convert_rows(rows, tree, fps=args.fps, device=args.device)
    np.savez(args.output, **data)
    print(f"Conversion completed, data saved to {args.output}")
    for k in ("dof_names", "body_names", "dof_positions", "dof_velocities", "body_positions", "body_rotations", "body_linear_velocities", "body_angular_velocities"):
        print(f"{k}:", data[k].shape)


if __name__ == "__main__":
    main()
