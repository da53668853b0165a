"""Generate the bundled, hand-authored robot descriptions (URDF XML).

The reference downloads PR2 / Fetch / JAXON URDFs at import time (skmp/__init__.py:7-15,
skmp/robot/jaxon.py:7); none of them exists in an offline container.  These tables are
*structure-faithful stand-ins*: joint and link names, tree topology, joint types and axes are
the ones the reference's configs and sphere tables rely on (skmp/robot/pr2.py:63-124,
fetch.py:27-35, jaxon.py:86-96, *_coll_spheres.yaml link names); offsets and limits follow the
public robot descriptions as closely as they could be restated from memory.  When the real URDF
is on disk (~/.skrobot/...), skmp_b200.robot.* loads that instead.

Run:  python scikit-motionplan_b200/tools/make_urdfs.py
"""
from pathlib import Path

OUT = Path(__file__).resolve().parent.parent / "skmp_b200" / "robot" / "urdf"

CONT = (None, None)


def urdf(name, root, joints, masses=None):
    """joints: (joint_name, type, parent, child, xyz, rpy, axis, (lo, hi))"""
    masses = masses or {}
    links = [root] + [j[3] for j in joints]
    out = ['<?xml version="1.0"?>', f'<robot name="{name}">']
    for l in links:
        m = masses.get(l, 1.0)
        out.append(f'  <link name="{l}"><inertial><origin xyz="0 0 0" rpy="0 0 0"/><mass value="{m}"/></inertial></link>')
    for jn, jt, par, ch, xyz, rpy, axis, lim in joints:
        out.append(f'  <joint name="{jn}" type="{jt}">')
        out.append(f'    <origin xyz="{xyz[0]} {xyz[1]} {xyz[2]}" rpy="{rpy[0]} {rpy[1]} {rpy[2]}"/>')
        out.append(f'    <parent link="{par}"/><child link="{ch}"/>')
        if jt != "fixed":
            out.append(f'    <axis xyz="{axis[0]} {axis[1]} {axis[2]}"/>')
        if jt in ("revolute", "prismatic"):
            out.append(f'    <limit lower="{lim[0]}" upper="{lim[1]}" effort="100" velocity="1"/>')
        out.append("  </joint>")
    out.append("</robot>")
    return "\n".join(out) + "\n"


Z0 = (0, 0, 0)
X, Y, Z = (1, 0, 0), (0, 1, 0), (0, 0, 1)


def pr2():
    J = []
    J.append(("base_footprint_joint", "fixed", "base_footprint", "base_link", (0, 0, 0.051), Z0, X, CONT))
    J.append(("torso_lift_joint", "prismatic", "base_link", "torso_lift_link", (-0.05, 0, 0.739675), Z0, Z, (0.0, 0.33)))
    J.append(("head_pan_joint", "revolute", "torso_lift_link", "head_pan_link", (-0.01707, 0, 0.38145), Z0, Z, (-3.007, 3.007)))
    J.append(("head_tilt_joint", "revolute", "head_pan_link", "head_tilt_link", (0.068, 0, 0), Z0, Y, (-0.471238, 1.39626)))
    J.append(("head_plate_frame_joint", "fixed", "head_tilt_link", "head_plate_frame", (0.0232, 0, 0.0645), Z0, X, CONT))
    J.append(("laser_tilt_mount_joint", "revolute", "torso_lift_link", "laser_tilt_mount_link", (0.09893, 0, 0.227), Z0, Y, (-0.7854, 1.48353)))
    for s, sy in (("r", -1.0), ("l", 1.0)):
        pan = (-2.2853981634, 0.714601836603) if s == "r" else (-0.714601836603, 2.2853981634)
        roll = (-3.9, 0.8) if s == "r" else (-0.8, 3.9)
        J.append((f"{s}_shoulder_pan_joint", "revolute", "torso_lift_link", f"{s}_shoulder_pan_link", (0, 0.188 * sy, 0), Z0, Z, pan))
        J.append((f"{s}_shoulder_lift_joint", "revolute", f"{s}_shoulder_pan_link", f"{s}_shoulder_lift_link", (0.1, 0, 0), Z0, Y, (-0.5236, 1.3963)))
        J.append((f"{s}_upper_arm_roll_joint", "revolute", f"{s}_shoulder_lift_link", f"{s}_upper_arm_roll_link", Z0, Z0, X, roll))
        J.append((f"{s}_upper_arm_joint", "fixed", f"{s}_upper_arm_roll_link", f"{s}_upper_arm_link", Z0, Z0, X, CONT))
        J.append((f"{s}_elbow_flex_joint", "revolute", f"{s}_upper_arm_link", f"{s}_elbow_flex_link", (0.4, 0, 0), Z0, Y, (-2.3213, 0.0)))
        J.append((f"{s}_forearm_roll_joint", "continuous", f"{s}_elbow_flex_link", f"{s}_forearm_roll_link", Z0, Z0, X, CONT))
        J.append((f"{s}_forearm_joint", "fixed", f"{s}_forearm_roll_link", f"{s}_forearm_link", Z0, Z0, X, CONT))
        J.append((f"{s}_wrist_flex_joint", "revolute", f"{s}_forearm_link", f"{s}_wrist_flex_link", (0.321, 0, 0), Z0, Y, (-2.18, 0.0)))
        J.append((f"{s}_wrist_roll_joint", "continuous", f"{s}_wrist_flex_link", f"{s}_wrist_roll_link", Z0, Z0, X, CONT))
        J.append((f"{s}_gripper_palm_joint", "fixed", f"{s}_wrist_roll_link", f"{s}_gripper_palm_link", Z0, Z0, X, CONT))
        J.append((f"{s}_gripper_tool_joint", "fixed", f"{s}_gripper_palm_link", f"{s}_gripper_tool_frame", (0.18, 0, 0), Z0, X, CONT))
        J.append((f"{s}_gripper_l_finger_joint", "revolute", f"{s}_gripper_palm_link", f"{s}_gripper_l_finger_link", (0.07691, 0.01, 0), Z0, Z, (0.0, 0.548)))
        J.append((f"{s}_gripper_r_finger_joint", "revolute", f"{s}_gripper_palm_link", f"{s}_gripper_r_finger_link", (0.07691, -0.01, 0), Z0, (0, 0, -1), (0.0, 0.548)))
        J.append((f"{s}_gripper_l_finger_tip_joint", "revolute", f"{s}_gripper_l_finger_link", f"{s}_gripper_l_finger_tip_link", (0.09137, 0.00495, 0), Z0, (0, 0, -1), (0.0, 0.548)))
        J.append((f"{s}_gripper_r_finger_tip_joint", "revolute", f"{s}_gripper_r_finger_link", f"{s}_gripper_r_finger_tip_link", (0.09137, -0.00495, 0), Z0, Z, (0.0, 0.548)))
    return urdf("pr2", "base_footprint", J)


def fetch():
    J = []
    J.append(("torso_lift_joint", "prismatic", "base_link", "torso_lift_link", (-0.086875, 0, 0.37743), Z0, Z, (0.0, 0.38615)))
    J.append(("torso_fixed_joint", "fixed", "base_link", "torso_fixed_link", (-0.086875, 0, 0.377425), Z0, X, CONT))
    J.append(("bellows_joint", "fixed", "torso_lift_link", "bellows_link2", Z0, Z0, X, CONT))
    J.append(("estop_joint", "fixed", "base_link", "estop_link", (-0.12465, 0.23892, 0.31127), (1.5708, 0, 0), X, CONT))
    J.append(("laser_joint", "fixed", "base_link", "laser_link", (0.235, 0, 0.2878), (3.14159265359, 0, 0), X, CONT))
    J.append(("r_wheel_joint", "continuous", "base_link", "r_wheel_link", (0.0012914, -0.18738, 0.055325), Z0, Y, CONT))
    J.append(("l_wheel_joint", "continuous", "base_link", "l_wheel_link", (0.0012914, 0.18738, 0.055325), Z0, Y, CONT))
    J.append(("head_pan_joint", "revolute", "torso_lift_link", "head_pan_link", (0.053125, 0, 0.603001417713939), Z0, Z, (-1.57, 1.57)))
    J.append(("head_tilt_joint", "revolute", "head_pan_link", "head_tilt_link", (0.14253, 0, 0.057999), Z0, Y, (-0.76, 1.45)))
    J.append(("shoulder_pan_joint", "revolute", "torso_lift_link", "shoulder_pan_link", (0.119525, 0, 0.34858), Z0, Z, (-1.6056, 1.6056)))
    J.append(("shoulder_lift_joint", "revolute", "shoulder_pan_link", "shoulder_lift_link", (0.117, 0, 0.06), Z0, Y, (-1.221, 1.518)))
    J.append(("upperarm_roll_joint", "continuous", "shoulder_lift_link", "upperarm_roll_link", (0.219, 0, 0), Z0, X, CONT))
    J.append(("elbow_flex_joint", "revolute", "upperarm_roll_link", "elbow_flex_link", (0.133, 0, 0), Z0, Y, (-2.251, 2.251)))
    J.append(("forearm_roll_joint", "continuous", "elbow_flex_link", "forearm_roll_link", (0.197, 0, 0), Z0, X, CONT))
    J.append(("wrist_flex_joint", "revolute", "forearm_roll_link", "wrist_flex_link", (0.1245, 0, 0), Z0, Y, (-2.16, 2.16)))
    J.append(("wrist_roll_joint", "continuous", "wrist_flex_link", "wrist_roll_link", (0.1385, 0, 0), Z0, X, CONT))
    J.append(("gripper_axis", "fixed", "wrist_roll_link", "gripper_link", (0.16645, 0, 0), Z0, X, CONT))
    J.append(("r_gripper_finger_joint", "prismatic", "gripper_link", "r_gripper_finger_link", (0, 0.015425, 0), Z0, Y, (0.0, 0.05)))
    J.append(("l_gripper_finger_joint", "prismatic", "gripper_link", "l_gripper_finger_link", (0, -0.015425, 0), Z0, (0, -1, 0), (0.0, 0.05)))
    return urdf("fetch", "base_link", J)


def jaxon():
    J = []
    masses = {"BODY": 15.0}
    for s, sy in (("R", -1.0), ("L", 1.0)):
        p = f"{s}LEG"
        J.append((f"{p}_JOINT0", "revolute", "BODY", f"{p}_LINK0", (0, 0.1 * sy, 0), Z0, Z, (-0.75, 0.75) if s == "L" else (-0.75, 0.75)))
        J.append((f"{p}_JOINT1", "revolute", f"{p}_LINK0", f"{p}_LINK1", Z0, Z0, X, (-0.72, 0.52) if s == "R" else (-0.52, 0.72)))
        J.append((f"{p}_JOINT2", "revolute", f"{p}_LINK1", f"{p}_LINK2", Z0, Z0, Y, (-2.09, 0.63)))
        J.append((f"{p}_JOINT3", "revolute", f"{p}_LINK2", f"{p}_LINK3", (0, 0, -0.38), Z0, Y, (0.0, 2.78)))
        J.append((f"{p}_JOINT4", "revolute", f"{p}_LINK3", f"{p}_LINK4", (0, 0, -0.38), Z0, Y, (-1.4, 1.4)))
        J.append((f"{p}_JOINT5", "revolute", f"{p}_LINK4", f"{p}_LINK5", (0, 0, -0.04), Z0, X, (-1.04, 1.04)))
        for i, m in enumerate((2.0, 1.0, 6.0, 4.0, 1.0, 2.0)):
            masses[f"{p}_LINK{i}"] = m
    J.append(("CHEST_JOINT0", "revolute", "BODY", "CHEST_LINK0", (0.006, 0, 0.182), Z0, X, (-0.15, 0.15)))
    J.append(("CHEST_JOINT1", "revolute", "CHEST_LINK0", "CHEST_LINK1", Z0, Z0, Y, (-0.2, 0.6)))
    J.append(("CHEST_JOINT2", "revolute", "CHEST_LINK1", "CHEST_LINK2", (0, 0, 0.12), Z0, Z, (-0.75, 0.75)))
    masses.update({"CHEST_LINK0": 1.0, "CHEST_LINK1": 1.5, "CHEST_LINK2": 20.0})
    J.append(("HEAD_JOINT0", "revolute", "CHEST_LINK2", "HEAD_LINK0", (0.03, 0, 0.43), Z0, Z, (-1.0, 1.0)))
    J.append(("HEAD_JOINT1", "revolute", "HEAD_LINK0", "HEAD_LINK1", (0, 0, 0.05), Z0, Y, (-0.35, 0.65)))
    masses.update({"HEAD_LINK0": 0.5, "HEAD_LINK1": 1.5})
    for s, sy in (("R", -1.0), ("L", 1.0)):
        p = f"{s}ARM"
        J.append((f"{p}_JOINT0", "revolute", "CHEST_LINK2", f"{p}_LINK0", (0.05, 0.1 * sy, 0.33), Z0, Z, (-0.3, 1.0) if s == "R" else (-1.0, 0.3)))
        J.append((f"{p}_JOINT1", "revolute", f"{p}_LINK0", f"{p}_LINK1", (0, 0.14 * sy, 0.03), Z0, Y, (-3.0, 1.0)))
        J.append((f"{p}_JOINT2", "revolute", f"{p}_LINK1", f"{p}_LINK2", Z0, Z0, X, (-1.9, 0.3) if s == "R" else (-0.3, 1.9)))
        J.append((f"{p}_JOINT3", "revolute", f"{p}_LINK2", f"{p}_LINK3", (0, 0, -0.02), Z0, Z, (-1.6, 1.6)))
        J.append((f"{p}_JOINT4", "revolute", f"{p}_LINK3", f"{p}_LINK4", (0, 0, -0.255), Z0, Y, (-2.4, 0.0)))
        J.append((f"{p}_JOINT5", "revolute", f"{p}_LINK4", f"{p}_LINK5", (0, 0, -0.02), Z0, Z, (-1.6, 1.6)))
        J.append((f"{p}_JOINT6", "revolute", f"{p}_LINK5", f"{p}_LINK6", (0, 0, -0.245), Z0, X, (-1.4, 1.4)))
        J.append((f"{p}_JOINT7", "revolute", f"{p}_LINK6", f"{p}_LINK7", Z0, Z0, Y, (-1.4, 1.4)))
        J.append((f"{p}_F_JOINT0", "revolute", f"{p}_LINK7", f"{p}_FINGER0", (0, 0.03 * sy, -0.17), (0, 1.5708, 0), Z, (-0.5, 0.5)))
        J.append((f"{p}_F_JOINT1", "revolute", f"{p}_LINK7", f"{p}_FINGER1", (0, -0.03 * sy, -0.17), (0, 1.5708, 0), Z, (-0.5, 0.5)))
        for i, m in enumerate((1.0, 2.0, 1.5, 2.0, 1.0, 1.5, 0.5, 1.5)):
            masses[f"{p}_LINK{i}"] = m
        masses[f"{p}_FINGER0"] = 0.2
        masses[f"{p}_FINGER1"] = 0.2
    return urdf("JAXON_RED", "BODY", J, masses)


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "pr2.urdf").write_text(pr2())
    (OUT / "fetch.urdf").write_text(fetch())
    (OUT / "jaxon.urdf").write_text(jaxon())
    print("wrote", sorted(p.name for p in OUT.iterdir()))
