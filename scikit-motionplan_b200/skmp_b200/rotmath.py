"""Rotation helpers with scikit-robot's conventions (skrobot.coordinates.math), which the
reference imports at skmp/kinematics.py:7, skmp/constraint.py:22-23 and skmp/robot/utils.py:10.

skrobot's ``rpy_matrix(az, ay, ax)`` is ``Rz(az) @ Ry(ay) @ Rx(ax)`` and ``rpy_angle(R)`` returns
two (yaw, pitch, roll) solutions; the reference always takes ``[0]`` and flips it to (roll, pitch, yaw).
"""
import numpy as np


def rpy_matrix(az: float, ay: float, ax: float) -> np.ndarray:
    cz, sz, cy, sy, cx, sx = np.cos(az), np.sin(az), np.cos(ay), np.sin(ay), np.cos(ax), np.sin(ax)
    rz = np.array([[cz, -sz, 0.0], [sz, cz, 0.0], [0.0, 0.0, 1.0]])
    ry = np.array([[cy, 0.0, sy], [0.0, 1.0, 0.0], [-sy, 0.0, cy]])
    rx = np.array([[1.0, 0.0, 0.0], [0.0, cx, -sx], [0.0, sx, cx]])
    return rz @ ry @ rx


def rpy_angle(matrix: np.ndarray):
    """(yaw, pitch, roll) of ``matrix`` and the alternative solution, like skrobot."""
    m = np.asarray(matrix, dtype=np.float64)
    a = np.arctan2(m[1, 0], m[0, 0])
    sa, ca = np.sin(a), np.cos(a)
    b = np.arctan2(-m[2, 0], ca * m[0, 0] + sa * m[1, 0])
    c = np.arctan2(sa * m[0, 2] - ca * m[1, 2], -sa * m[0, 1] + ca * m[1, 1])
    a2 = a + np.pi
    sa, ca = np.sin(a2), np.cos(a2)
    b2 = np.arctan2(-m[2, 0], ca * m[0, 0] + sa * m[1, 0])
    c2 = np.arctan2(sa * m[0, 2] - ca * m[1, 2], -sa * m[0, 1] + ca * m[1, 1])
    return np.array([a, b, c]), np.array([a2, b2, c2])


def rotation_matrix(theta: float, axis) -> np.ndarray:
    """Rodrigues rotation of ``theta`` about ``axis``."""
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    k = np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])
    return np.eye(3) + np.sin(theta) * k + (1.0 - np.cos(theta)) * (k @ k)


def matrix2quaternion(m: np.ndarray) -> np.ndarray:
    """Rotation matrix -> (w, x, y, z) with w >= 0 branch selection (skrobot convention)."""
    m = np.asarray(m, dtype=np.float64)
    q0_2 = (1 + m[0, 0] + m[1, 1] + m[2, 2]) / 4.0
    q1_2 = (1 + m[0, 0] - m[1, 1] - m[2, 2]) / 4.0
    q2_2 = (1 - m[0, 0] + m[1, 1] - m[2, 2]) / 4.0
    q3_2 = (1 - m[0, 0] - m[1, 1] + m[2, 2]) / 4.0
    mq_2 = max(q0_2, q1_2, q2_2, q3_2)
    if np.isclose(mq_2, q0_2):
        q0 = np.sqrt(q0_2)
        q1 = (m[2, 1] - m[1, 2]) / (4.0 * q0)
        q2 = (m[0, 2] - m[2, 0]) / (4.0 * q0)
        q3 = (m[1, 0] - m[0, 1]) / (4.0 * q0)
    elif np.isclose(mq_2, q1_2):
        q1 = np.sqrt(q1_2)
        q0 = (m[2, 1] - m[1, 2]) / (4.0 * q1)
        q2 = (m[1, 0] + m[0, 1]) / (4.0 * q1)
        q3 = (m[0, 2] + m[2, 0]) / (4.0 * q1)
    elif np.isclose(mq_2, q2_2):
        q2 = np.sqrt(q2_2)
        q0 = (m[0, 2] - m[2, 0]) / (4.0 * q2)
        q1 = (m[1, 0] + m[0, 1]) / (4.0 * q2)
        q3 = (m[1, 2] + m[2, 1]) / (4.0 * q2)
    else:
        q3 = np.sqrt(q3_2)
        q0 = (m[1, 0] - m[0, 1]) / (4.0 * q3)
        q1 = (m[0, 2] + m[2, 0]) / (4.0 * q3)
        q2 = (m[1, 2] + m[2, 1]) / (4.0 * q3)
    q = np.array([q0, q1, q2, q3])
    if q[0] < 0:
        q = -q
    return q


def wxyz2xyzw(q: np.ndarray) -> np.ndarray:
    q = np.asarray(q)
    return np.array([q[1], q[2], q[3], q[0]])


def urdf_rpy_matrix(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    return rpy_matrix(yaw, pitch, roll)
