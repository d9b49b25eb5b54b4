"""Small rigid-body helpers of the facade (host side, NumPy).

Same names and conventions as the reference's helper module (python/simpleicp/mathutils.py):
R = Rx(alpha1) @ Ry(alpha2) @ Rz(alpha3), H = [[R, t], [0, 1]].  The device code evaluates the
same closed form (csrc/common.cuh: euler_to_R).
"""
from typing import Sequence, Tuple

import numpy as np


def euler_angles_to_rotation_matrix(alpha1: float, alpha2: float, alpha3: float) -> np.ndarray:
    """Closed form of Rx(alpha1) Ry(alpha2) Rz(alpha3) (reference: mathutils.py:39-68)."""
    s1, c1 = np.sin(alpha1), np.cos(alpha1)
    s2, c2 = np.sin(alpha2), np.cos(alpha2)
    s3, c3 = np.sin(alpha3), np.cos(alpha3)
    R = np.empty((3, 3))
    R[0] = (c2 * c3, -c2 * s3, s2)
    R[1] = (c1 * s3 + s1 * s2 * c3, c1 * c3 - s1 * s2 * s3, -s1 * c2)
    R[2] = (s1 * s3 - c1 * s2 * c3, s1 * c3 + c1 * s2 * s3, c1 * c2)
    return R


def euler_angles_to_linearized_rotation_matrix(alpha1: float, alpha2: float, alpha3: float) -> np.ndarray:
    """First-order rotation I + [alpha]x (reference: mathutils.py:29-36)."""
    dR = np.eye(3)
    dR[0, 1], dR[0, 2] = -alpha3, alpha2
    dR[1, 0], dR[1, 2] = alpha3, -alpha1
    dR[2, 0], dR[2, 1] = -alpha2, alpha1
    return dR


def rotation_matrix_to_euler_angles(R: np.ndarray) -> Tuple[float, float, float]:
    """Inverse of euler_angles_to_rotation_matrix (reference: mathutils.py:71-78)."""
    return (float(np.arctan2(-R[1, 2], R[2, 2])), float(np.arcsin(R[0, 2])),
            float(np.arctan2(-R[0, 1], R[0, 0])))


def create_homogeneous_transformation_matrix(R: np.ndarray, t: Sequence[float]) -> np.ndarray:
    """4 x 4 matrix [[R, t], [0, 1]] (reference: mathutils.py:81-93)."""
    H = np.zeros((4, 4))
    H[:3, :3] = R
    H[:3, 3] = np.asarray(t, dtype=float)[:3]
    H[3, 3] = 1.0
    return H


def euler_coord_to_homogeneous_coord(Xe: np.ndarray) -> np.ndarray:
    """Append a column of ones (reference: mathutils.py:10-16)."""
    Xe = np.asarray(Xe)
    return np.hstack((Xe, np.ones((Xe.shape[0], 1))))


def homogeneous_coord_to_euler_coord(Xh: np.ndarray) -> np.ndarray:
    """Divide by the homogeneous coordinate (reference: mathutils.py:19-26)."""
    Xh = np.asarray(Xh)
    return Xh[:, :3] / Xh[:, 3:4]
