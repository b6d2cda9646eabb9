"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the reference's offline dataset
pipeline, ``/root/reference/motions/data_convert.py:161-379``: CSV at 30 fps -> 60 fps interpolation ->
forward kinematics -> central-difference + gaussian velocities -> quaternion-log angular velocities -> the clip ``.npz``.

The reference calls scipy (``interp1d``, ``Rotation`` / ``Slerp``, ``gaussian_filter1d``) -- installed here and on the GPU
box, so those steps are the SAME library calls in the SAME order.  The one dependency that is absent is Pinocchio 2.x/3.x
(``import pinocchio as pin``, ``data_convert.py:62``; not pinned by the reference, no requirements file lists it): its
forward kinematics over the URDF tree and Eigen's rotation-matrix -> quaternion conversion are restated below from their
published algorithms (``forward_kinematics``, ``matrix_to_quat_wxyz``).

Pinning: the reference ships one OUTPUT of this pipeline, ``motions/custom_motion.npz`` (rows 110:265 of
``datasets/walk1_subject1.csv`` with ``g1_model/urdf/g1_29dof_rev_1_0.urdf``): ``tests/test_oracle_pins.py`` checks this
restatement against it (joint positions bit for bit, everything else to float32 rounding), and in the build container the
UNMODIFIED text of ``data_convert.py`` is executed with this module's FK standing in for Pinocchio (``oracle/ref_harness.py``).
"""

from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

# data_convert.py:258-288 (joint order of the CSV = Pinocchio's q order) and :301-327 (recorded links)
G1_JOINT_NAMES = [
    "left_hip_pitch_joint", "left_hip_roll_joint", "left_hip_yaw_joint", "left_knee_joint", "left_ankle_pitch_joint", "left_ankle_roll_joint",
    "right_hip_pitch_joint", "right_hip_roll_joint", "right_hip_yaw_joint", "right_knee_joint", "right_ankle_pitch_joint", "right_ankle_roll_joint",
    "waist_yaw_joint", "waist_roll_joint", "waist_pitch_joint",
    "left_shoulder_pitch_joint", "left_shoulder_roll_joint", "left_shoulder_yaw_joint", "left_elbow_joint", "left_wrist_roll_joint",
    "left_wrist_pitch_joint", "left_wrist_yaw_joint",
    "right_shoulder_pitch_joint", "right_shoulder_roll_joint", "right_shoulder_yaw_joint", "right_elbow_joint", "right_wrist_roll_joint",
    "right_wrist_pitch_joint", "right_wrist_yaw_joint",
]  # fmt: skip
G1_BODY_NAMES = [
    "pelvis", "head_link", "torso_link", "left_shoulder_pitch_link", "left_shoulder_roll_link", "left_shoulder_yaw_link", "left_elbow_link",
    "right_shoulder_pitch_link", "right_shoulder_roll_link", "right_shoulder_yaw_link", "right_elbow_link", "left_hip_yaw_link",
    "left_hip_roll_link", "left_hip_pitch_link", "left_knee_link", "right_hip_yaw_link", "right_hip_roll_link", "right_hip_pitch_link",
    "right_knee_link", "right_rubber_hand", "left_rubber_hand", "right_ankle_roll_link", "left_ankle_roll_link", "waist_yaw_link",
    "waist_roll_link",
]  # fmt: skip


@dataclass
class KinematicTree:
    """URDF joints in topological order.  ``parent[i]`` indexes the joint whose child link is joint i's parent link (-1 =
    the root link); ``qidx[i]`` is the column of the joint-angle array driving joint i (-1 = fixed); ``link_joint[name]``
    is the joint whose child link is ``name`` (-1 for the root link)."""

    root_link: str
    names: List[str]
    parent: np.ndarray  # (J,) int32
    qidx: np.ndarray  # (J,) int32
    origin_xyz: np.ndarray  # (J, 3) float64
    origin_rot: np.ndarray  # (J, 3, 3) float64, from rpy
    axis: np.ndarray  # (J, 3) float64 (unit)
    link_joint: Dict[str, int]


def rpy_to_matrix(rpy: Sequence[float]) -> np.ndarray:
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array(
        [[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr], [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr], [-sp, cp * sr, cp * cr]],
        dtype=np.float64,
    )


def load_urdf_tree(urdf_path: str, joint_names: Sequence[str]) -> KinematicTree:
    """Kinematic part of a URDF (links, joints: parent / child / origin / axis / type).  A ``floating`` joint from a
    ``world`` link to the base (``g1_29dof_rev_1_0.urdf:8-11``) is the free-flyer itself: its child is the root link."""
    robot = ET.parse(urdf_path).getroot()
    joints = []
    for j in robot.findall("joint"):
        origin = j.find("origin")
        axis = j.find("axis")
        joints.append(
            dict(
                name=j.get("name"), type=j.get("type"), parent=j.find("parent").get("link"), child=j.find("child").get("link"),
                xyz=[float(v) for v in (origin.get("xyz", "0 0 0") if origin is not None else "0 0 0").split()],
                rpy=[float(v) for v in (origin.get("rpy", "0 0 0") if origin is not None else "0 0 0").split()],
                axis=[float(v) for v in (axis.get("xyz") if axis is not None else "1 0 0").split()],
            )
        )  # fmt: skip
    children = {j["child"] for j in joints}
    roots = [l.get("name") for l in robot.findall("link") if l.get("name") not in children]
    if len(roots) != 1:
        raise ValueError(f"URDF must have exactly one root link, found {roots}")
    root = roots[0]
    floating = [j for j in joints if j["type"] == "floating" and j["parent"] == root]
    if floating:  # world -> base through a floating joint: the base link is the free-flyer body
        root = floating[0]["child"]
        joints = [j for j in joints if j is not floating[0]]
    order, link_joint, frontier = [], {root: -1}, [root]
    while frontier:
        link = frontier.pop(0)
        for j in joints:
            if j["parent"] == link:
                link_joint[j["child"]] = len(order)
                order.append(j)
                frontier.append(j["child"])
    qmap = {n: i for i, n in enumerate(joint_names)}
    J = len(order)
    tree = KinematicTree(
        root_link=root, names=[j["name"] for j in order], parent=np.array([link_joint[j["parent"]] for j in order], dtype=np.int32),
        qidx=np.full(J, -1, dtype=np.int32), origin_xyz=np.array([j["xyz"] for j in order], dtype=np.float64).reshape(J, 3),
        origin_rot=np.stack([rpy_to_matrix(j["rpy"]) for j in order]) if J else np.zeros((0, 3, 3)),
        axis=np.array([j["axis"] for j in order], dtype=np.float64).reshape(J, 3), link_joint=link_joint,
    )  # fmt: skip
    for i, j in enumerate(order):
        if j["type"] in ("revolute", "continuous"):
            if j["name"] not in qmap:
                raise ValueError(f"URDF joint {j['name']} has no column in the joint-name list")
            tree.qidx[i] = qmap[j["name"]]
            n = np.linalg.norm(tree.axis[i])
            tree.axis[i] /= n
        elif j["type"] != "fixed":
            raise ValueError(f"URDF joint type {j['type']} is not supported ({j['name']})")
    return tree


def quat_xyzw_to_matrix(q: np.ndarray) -> np.ndarray:
    """Eigen ``Quaternion::toRotationMatrix`` on the NORMALISED quaternion (Pinocchio's free-flyer normalises q[3:7])."""
    x, y, z, w = (q / np.linalg.norm(q)).tolist()
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz, txx, txy, txz, tyy, tyz, tzz = tx * w, ty * w, tz * w, tx * x, ty * x, tz * x, ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy], [txy + twz, 1 - (txx + tzz), tyz - twx], [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def axis_angle_matrix(axis: np.ndarray, angle: float) -> np.ndarray:
    """Rodrigues (Eigen ``AngleAxis::toRotationMatrix``)."""
    c, s = np.cos(angle), np.sin(angle)
    x, y, z = axis.tolist()
    C = 1 - c
    return np.array([[c + x * x * C, x * y * C - z * s, x * z * C + y * s], [y * x * C + z * s, c + y * y * C, y * z * C - x * s], [z * x * C - y * s, z * y * C + x * s, c + z * z * C]])


def matrix_to_quat_wxyz(m: np.ndarray) -> np.ndarray:
    """Eigen ``Quaternion(Matrix3)`` (``pin.Quaternion(link_tf.rotation)``, data_convert.py:340-343), returned (w, x, y, z)."""
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)  # x, y, z, w
    if t > 0.0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0], q[1], q[2] = (m[2, 1] - m[1, 2]) * t, (m[0, 2] - m[2, 0]) * t, (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (m[k, j] - m[j, k]) * t
        q[j] = (m[j, i] + m[i, j]) * t
        q[k] = (m[k, i] + m[i, k]) * t
    return np.array([q[3], q[0], q[1], q[2]])


def forward_kinematics(tree: KinematicTree, root_pos: np.ndarray, root_quat_xyzw: np.ndarray, joint_pos: np.ndarray, bodies: Sequence[str]):
    """World placement of every link in ``bodies`` for ONE configuration (float64): ``(positions (B,3), rotations (B,3,3))``.
    ``oMi[child] = oMi[parent] * (origin * R(axis, q))`` (Pinocchio ``forwardKinematics`` + ``updateFramePlacements``)."""
    J = len(tree.names)
    R0, p0 = quat_xyzw_to_matrix(np.asarray(root_quat_xyzw, dtype=np.float64)), np.asarray(root_pos, dtype=np.float64)
    R, p = np.zeros((J, 3, 3)), np.zeros((J, 3))
    for i in range(J):
        Rp, pp = (R0, p0) if tree.parent[i] < 0 else (R[tree.parent[i]], p[tree.parent[i]])
        Rl = tree.origin_rot[i]
        if tree.qidx[i] >= 0:
            Rl = Rl @ axis_angle_matrix(tree.axis[i], float(joint_pos[tree.qidx[i]]))
        R[i] = Rp @ Rl
        p[i] = pp + Rp @ tree.origin_xyz[i]
    pos, rot = np.zeros((len(bodies), 3)), np.zeros((len(bodies), 3, 3))
    for b, name in enumerate(bodies):
        j = tree.link_joint[name]
        pos[b], rot[b] = (p0, R0) if j < 0 else (p[j], R[j])
    return pos, rot


# ---- data_convert.py:68-108 (quaternion helpers, literal) ------------------------------------------------------------------
def quaternion_inverse(q):
    w, x, y, z = q
    norm_sq = w * w + x * x + y * y + z * z
    if norm_sq < 1e-8:
        norm_sq = 1e-8
    return np.array([w, -x, -y, -z], dtype=q.dtype) / norm_sq


def quaternion_multiply(q1, q2):
    w1, x1, y1, z1 = q1
    w2, x2, y2, z2 = q2
    return np.array(
        [w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2],
        dtype=q1.dtype,
    )


def compute_angular_velocity(q_prev, q_next, dt, eps=1e-8):
    q_rel = quaternion_multiply(quaternion_inverse(q_prev), q_next)
    norm_q_rel = np.linalg.norm(q_rel)
    if norm_q_rel < eps:
        return np.zeros(3, dtype=np.float32)
    q_rel /= norm_q_rel
    if q_rel[0] < 0.0:
        q_rel = -q_rel
    w = np.clip(q_rel[0], -1.0, 1.0)
    angle = 2.0 * np.arccos(w)
    sin_half = np.sqrt(1.0 - w * w)
    if sin_half < eps:
        return np.zeros(3, dtype=np.float32)
    return (angle / dt) * (q_rel[1:] / sin_half)


def velocity_stages(dof_positions, body_positions, body_rotations, dt, exact_angular=False):
    """data_convert.py:290-297, :349-371 on given positions / rotations: ``(dof_velocities, body_linear_velocities,
    body_angular_velocities)``.  ``exact_angular``: the SAME angular-velocity expression evaluated in float64 on the float32
    rotations (the rotations are widened first, so every helper above runs in float64) -- the value the reference's float32
    evaluation scatters around: one ulp of the relative quaternion's w moves 2 acos(w) / dt by up to ~0.04 rad/s at 60 fps."""
    from scipy.ndimage import gaussian_filter1d

    N, B = body_positions.shape[:2]
    dof_velocities = np.zeros_like(dof_positions)
    dof_velocities[1:-1] = (dof_positions[2:] - dof_positions[:-2]) / (2 * dt)
    dof_velocities[0] = (dof_positions[1] - dof_positions[0]) / dt
    dof_velocities[-1] = (dof_positions[-1] - dof_positions[-2]) / dt
    dof_velocities = gaussian_filter1d(dof_velocities, sigma=1, axis=0)
    lin = np.zeros_like(body_positions)
    lin[1:-1] = (body_positions[2:] - body_positions[:-2]) / (2 * dt)
    lin[0] = (body_positions[1] - body_positions[0]) / dt
    lin[-1] = (body_positions[-1] - body_positions[-2]) / dt
    lin = gaussian_filter1d(lin, sigma=1, axis=0)
    ang = np.zeros((N, B, 3), dtype=np.float32)
    rot = body_rotations.astype(np.float64) if exact_angular else body_rotations
    for j in range(B):
        quats = rot[:, j, :]
        angular_vels = np.zeros((N, 3), dtype=np.float32)
        if N > 1:
            angular_vels[0] = compute_angular_velocity(quats[0], quats[1], dt)
            angular_vels[-1] = compute_angular_velocity(quats[-2], quats[-1], dt)
        for k in range(1, N - 1):
            angular_vels[k] = 0.5 * (compute_angular_velocity(quats[k - 1], quats[k], dt) + compute_angular_velocity(quats[k], quats[k + 1], dt))
        ang[:, j, :] = gaussian_filter1d(angular_vels, sigma=1, axis=0)
    return dof_velocities, lin, ang


def convert(csv_rows: np.ndarray, tree: KinematicTree, joint_names: Sequence[str] = G1_JOINT_NAMES, body_names: Sequence[str] = G1_BODY_NAMES,
            fps: int = 60) -> Dict[str, np.ndarray]:
    """``main`` of data_convert.py (:161-379) on the already sliced CSV rows (``df.iloc[start:end].to_numpy(float32)``)."""
    from scipy.interpolate import interp1d
    from scipy.ndimage import gaussian_filter1d
    from scipy.spatial.transform import Rotation as R
    from scipy.spatial.transform import Slerp

    data_orig = np.asarray(csv_rows, dtype=np.float32)
    N_orig = data_orig.shape[0]
    root_data_orig, joint_data_orig = data_orig[:, :7], data_orig[:, 7:]
    dt_orig = 1.0 / 30
    t_orig = np.linspace(0, (N_orig - 1) * dt_orig, N_orig)
    dt = 1.0 / fps
    N = 2 * N_orig - 1
    t_new = np.linspace(0, (N_orig - 1) * dt_orig, N)
    root_pos = interp1d(t_orig, root_data_orig[:, 0:3], axis=0, kind="linear")(t_new)
    root_quat = Slerp(t_orig, R.from_quat(root_data_orig[:, 3:7]))(t_new).as_quat()
    root_data = np.hstack((root_pos, root_quat))
    joint_data = interp1d(t_orig, joint_data_orig, axis=0, kind="linear")(t_new)
    dof_positions = joint_data.copy()
    dof_velocities = np.zeros_like(dof_positions)
    dof_velocities[1:-1] = (dof_positions[2:] - dof_positions[:-2]) / (2 * dt)
    dof_velocities[0] = (dof_positions[1] - dof_positions[0]) / dt
    dof_velocities[-1] = (dof_positions[-1] - dof_positions[-2]) / dt
    dof_velocities_smoothed = gaussian_filter1d(dof_velocities, sigma=1, axis=0)
    B = len(body_names)
    body_positions = np.zeros((N, B, 3), dtype=np.float32)
    body_rotations = np.zeros((N, B, 4), dtype=np.float32)
    for i in range(N):
        pos, rot = forward_kinematics(tree, root_data[i, 0:3], root_data[i, 3:7], joint_data[i], body_names)
        body_positions[i] = pos
        for j in range(B):
            body_rotations[i, j] = matrix_to_quat_wxyz(rot[j])
    body_linear_velocities = np.zeros_like(body_positions)
    body_linear_velocities[1:-1] = (body_positions[2:] - body_positions[:-2]) / (2 * dt)
    body_linear_velocities[0] = (body_positions[1] - body_positions[0]) / dt
    body_linear_velocities[-1] = (body_positions[-1] - body_positions[-2]) / dt
    body_linear_velocities = gaussian_filter1d(body_linear_velocities, sigma=1, axis=0)
    body_angular_velocities = np.zeros((N, B, 3), dtype=np.float32)
    for j in range(B):
        quats = body_rotations[:, j, :]
        angular_vels = np.zeros((N, 3), dtype=np.float32)
        if N > 1:
            angular_vels[0] = compute_angular_velocity(quats[0], quats[1], dt)
            angular_vels[-1] = compute_angular_velocity(quats[-2], quats[-1], dt)
        for k in range(1, N - 1):
            angular_vels[k] = 0.5 * (compute_angular_velocity(quats[k - 1], quats[k], dt) + compute_angular_velocity(quats[k], quats[k + 1], dt))
        body_angular_velocities[:, j, :] = gaussian_filter1d(angular_vels, sigma=1, axis=0)
    return {
        "fps": np.int64(fps), "dof_names": np.array(list(joint_names), dtype=np.str_), "body_names": np.array(list(body_names), dtype=np.str_),
        "dof_positions": dof_positions, "dof_velocities": dof_velocities_smoothed, "body_positions": body_positions,
        "body_rotations": body_rotations, "body_linear_velocities": body_linear_velocities, "body_angular_velocities": body_angular_velocities,
    }  # fmt: skip


# ---- a stand-in for the absent Pinocchio, so the UNMODIFIED data_convert.py text can run on this module's FK ---------------
class _SE3:
    def __init__(self, R, p):
        self.rotation, self.translation = R, p


class _Quat:
    def __init__(self, m):
        self.w, self.x, self.y, self.z = matrix_to_quat_wxyz(np.asarray(m)).tolist()


def make_pinocchio_stub(joint_names: Sequence[str] = G1_JOINT_NAMES):
    """A module object exposing exactly the Pinocchio calls data_convert.py makes (:125-128, :331-343)."""
    import types

    pin = types.ModuleType("pinocchio")

    class Model:
        def __init__(self, tree):
            self.tree = tree
            self.nq = 7 + int((tree.qidx >= 0).sum())
            self.frames = [tree.root_link] + [None] * 0
            self._links = list(tree.link_joint)

        def getFrameId(self, name):
            return self._links.index(name)

    class Data:
        def __init__(self):
            self.oMf = []

    class RobotWrapper:
        @staticmethod
        def BuildFromURDF(urdf_path, mesh_dir, root_joint=None):
            rw = RobotWrapper()
            rw.model, rw.data = Model(load_urdf_tree(urdf_path, joint_names)), Data()
            return rw

    def forwardKinematics(model, data, q):
        pos, rot = forward_kinematics(model.tree, q[0:3], q[3:7], q[7:], model._links)
        data._fk = [_SE3(rot[i], pos[i]) for i in range(len(model._links))]

    def updateFramePlacements(model, data):
        data.oMf = data._fk

    def neutral(model):
        q = np.zeros(model.nq)
        q[6] = 1.0
        return q

    pin.RobotWrapper, pin.JointModelFreeFlyer = RobotWrapper, (lambda: None)
    pin.forwardKinematics, pin.updateFramePlacements, pin.neutral, pin.Quaternion = forwardKinematics, updateFramePlacements, neutral, _Quat
    return pin
