"""Robot-side constants the AMP path needs from the simulator (PhysX / Isaac Lab are out of scope).

In the reference these come from ``self.robot.data.joint_names / body_names`` at env construction
(``g1_amp_env.py:40-60``, ``humanoid_amp_env.py:42-56``).  The only in-tree record of the G1 orders is the trailing
comment of ``motions/test/get_joint_name.py:231-232``; they are restated here as data.
"""

from __future__ import annotations

from dataclasses import dataclass, field


@dataclass(frozen=True)
class RobotSpec:
    name: str
    joint_names: tuple  # simulator (robot) joint order -> defines the dof permutation applied to clip columns
    body_names: tuple  # simulator body order (Bsim entries) -> indexes body_pos_w / body_quat_w
    reference_body: str  # cfg.reference_body (g1_amp_env_cfg.py:53, humanoid_amp_env_cfg.py:42)
    key_body_names: tuple  # order matters: RH, LH, RF, LF (g1_amp_env.py:40-45)
    # _reset_strategy_random takes the root transform from a HARD-CODED clip body and lifts it off the ground:
    # G1 "pelvis" + 0.05 (g1_amp_env.py:398, 403-405), 28-DoF humanoid "torso" + 0.15 (humanoid_amp_env.py:194, 199-201)
    reset_root_body: str = "pelvis"
    reset_root_lift: float = 0.05

    @property
    def num_joints(self) -> int:
        return len(self.joint_names)

    @property
    def amp_observation_space(self) -> int:
        """A = 2*D + 13 + 3*Kb, the width ``compute_obs`` produces (g1_amp_env.py:545-561)."""
        return 2 * len(self.joint_names) + 13 + 3 * len(self.key_body_names)


G1 = RobotSpec(
    name="g1",
    joint_names=(
        "left_hip_pitch_joint", "right_hip_pitch_joint", "waist_yaw_joint", "left_hip_roll_joint",
        "right_hip_roll_joint", "waist_roll_joint", "left_hip_yaw_joint", "right_hip_yaw_joint", "waist_pitch_joint",
        "left_knee_joint", "right_knee_joint", "left_shoulder_pitch_joint", "right_shoulder_pitch_joint",
        "left_ankle_pitch_joint", "right_ankle_pitch_joint", "left_shoulder_roll_joint", "right_shoulder_roll_joint",
        "left_ankle_roll_joint", "right_ankle_roll_joint", "left_shoulder_yaw_joint", "right_shoulder_yaw_joint",
        "left_elbow_joint", "right_elbow_joint", "left_wrist_roll_joint", "right_wrist_roll_joint",
        "left_wrist_pitch_joint", "right_wrist_pitch_joint", "left_wrist_yaw_joint", "right_wrist_yaw_joint",
    ),
    body_names=(
        "pelvis", "imu_in_pelvis", "left_hip_pitch_link", "pelvis_contour_link", "right_hip_pitch_link",
        "waist_yaw_link", "left_hip_roll_link", "right_hip_roll_link", "waist_roll_link", "left_hip_yaw_link",
        "right_hip_yaw_link", "torso_link", "left_knee_link", "right_knee_link", "d435_link", "head_link",
        "imu_in_torso", "left_shoulder_pitch_link", "logo_link", "mid360_link", "right_shoulder_pitch_link",
        "left_ankle_pitch_link", "right_ankle_pitch_link", "left_shoulder_roll_link", "right_shoulder_roll_link",
        "left_ankle_roll_link", "right_ankle_roll_link", "left_shoulder_yaw_link", "right_shoulder_yaw_link",
        "left_elbow_link", "right_elbow_link", "left_wrist_roll_link", "right_wrist_roll_link",
        "left_wrist_pitch_link", "right_wrist_pitch_link", "left_wrist_yaw_link", "right_wrist_yaw_link",
        "left_rubber_hand", "right_rubber_hand",
    ),
    reference_body="pelvis",
    key_body_names=("right_rubber_hand", "left_rubber_hand", "right_ankle_roll_link", "left_ankle_roll_link"),
)  # fmt: skip

# Isaac Lab's 28-DoF humanoid: the shipped clips already use the simulator's joint / body order.
HUMANOID28 = RobotSpec(
    name="humanoid28",
    joint_names=(
        "abdomen_x", "abdomen_y", "abdomen_z", "neck_x", "neck_y", "neck_z", "right_shoulder_x", "right_shoulder_y",
        "right_shoulder_z", "right_elbow", "left_shoulder_x", "left_shoulder_y", "left_shoulder_z", "left_elbow",
        "right_hip_x", "right_hip_y", "right_hip_z", "right_knee", "right_ankle_x", "right_ankle_y", "right_ankle_z",
        "left_hip_x", "left_hip_y", "left_hip_z", "left_knee", "left_ankle_x", "left_ankle_y", "left_ankle_z",
    ),
    body_names=(
        "pelvis", "torso", "head", "right_upper_arm", "right_lower_arm", "right_hand", "left_upper_arm",
        "left_lower_arm", "left_hand", "right_thigh", "right_shin", "right_foot", "left_thigh", "left_shin", "left_foot",
    ),
    reference_body="torso",
    key_body_names=("right_hand", "left_hand", "right_foot", "left_foot"),
    reset_root_body="torso",
    reset_root_lift=0.15,
)  # fmt: skip


def robot_for_clip(dof_names) -> RobotSpec:
    """Pick the robot whose joint set matches a clip's ``dof_names`` (order-insensitive)."""
    names = set(dof_names)
    for spec in (G1, HUMANOID28):
        if names == set(spec.joint_names):
            return spec
    raise ValueError(f"no known robot has the joint set {sorted(names)}")
