"""Quaternion <-> rotation matrix, as the reference gets them from pyquaternion.

The lifting scripts turn dataset poses into matrices with
`pyquaternion.Quaternion(q).rotation_matrix` (src/nuscenes/2d_to_3d.py:451,456,571,577;
src/waymo/2d_to_3d.py:575,689) and boxes back into quaternions with
`Quaternion(matrix=align_mat)` (src/nuscenes/2d_to_3d.py:796,806).  pyquaternion
(pinned 0.9.9, environment.yml:131) is an un-vendored dependency absent from the
reference tree: this file restates its published algorithm in fp64 numpy -
normalise when |1 - q.q| >= 1e-14, rotation matrix = rows/cols 1..3 of
Q(q) . Qbar(q)^T, matrix -> quaternion by the four-branch trace method on the
transposed matrix.  PARITY UNPINNED against pyquaternion itself (not installable
here); pinned against scipy's Rotation within 1e-15 in tests/test_host_logic.py.
"""
from __future__ import annotations

from math import sqrt

import numpy as np


class Quaternion:
    """Minimal stand-in with pyquaternion's element order (w, x, y, z)."""

    def __init__(self, *args, matrix=None):
        if matrix is not None:
            self.q = _from_matrix(np.asarray(matrix, dtype=np.float64))
        elif len(args) == 1:
            a = args[0]
            self.q = (a.q.copy() if isinstance(a, Quaternion) else np.asarray(a, dtype=np.float64).reshape(4).copy())
        elif len(args) == 4:
            self.q = np.asarray(args, dtype=np.float64)
        elif len(args) == 0:
            self.q = np.array([1.0, 0.0, 0.0, 0.0])
        else:
            raise ValueError("Quaternion(w, x, y, z) | Quaternion([w, x, y, z]) | Quaternion(matrix=M)")

    def __iter__(self):
        return iter(self.q)

    def __getitem__(self, i):
        return self.q[i]

    def __repr__(self):
        return "Quaternion({!r}, {!r}, {!r}, {!r})".format(*self.q)

    @property
    def elements(self):
        return self.q

    def _sum_of_squares(self):
        return np.dot(self.q, self.q)

    def _normalise(self):
        if not abs(1.0 - self._sum_of_squares()) < 1e-14:
            n = sqrt(self._sum_of_squares())
            if n > 0:
                self.q = self.q / n

    @property
    def rotation_matrix(self) -> np.ndarray:
        self._normalise()
        w, x, y, z = self.q
        q_mat = np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])
        q_bar = np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]])
        return np.dot(q_mat, q_bar.conj().transpose())[1:][:, 1:]


def _from_matrix(matrix: np.ndarray, rtol=1e-05, atol=1e-08) -> np.ndarray:
    if matrix.shape != (3, 3) and matrix.shape != (4, 4):
        raise ValueError("Invalid matrix shape: Input must be a 3x3 or 4x4 numpy array or matrix")
    R = matrix[:3, :3]
    if not np.allclose(np.dot(R, R.conj().transpose()), np.eye(3), rtol=rtol, atol=atol):
        raise ValueError("Matrix must be orthogonal, i.e. its transpose should be its inverse")
    if not np.isclose(np.linalg.det(R), 1.0, rtol=rtol, atol=atol):
        raise ValueError("Matrix must be special orthogonal i.e. its determinant must be +1.0")
    m = R.conj().transpose()
    if m[2, 2] < 0:
        if m[0, 0] > m[1, 1]:
            t = 1 + m[0, 0] - m[1, 1] - m[2, 2]
            q = [m[1, 2] - m[2, 1], t, m[0, 1] + m[1, 0], m[2, 0] + m[0, 2]]
        else:
            t = 1 - m[0, 0] + m[1, 1] - m[2, 2]
            q = [m[2, 0] - m[0, 2], m[0, 1] + m[1, 0], t, m[1, 2] + m[2, 1]]
    else:
        if m[0, 0] < -m[1, 1]:
            t = 1 - m[0, 0] - m[1, 1] + m[2, 2]
            q = [m[0, 1] - m[1, 0], m[2, 0] + m[0, 2], m[1, 2] + m[2, 1], t]
        else:
            t = 1 + m[0, 0] + m[1, 1] + m[2, 2]
            q = [t, m[1, 2] - m[2, 1], m[2, 0] - m[0, 2], m[0, 1] - m[1, 0]]
    q = np.array(q).astype("float64")
    q *= 0.5 / sqrt(t)
    return q


def quats_from_matrices(mats: np.ndarray) -> np.ndarray:
    """`_from_matrix` over a stack of proper rotation matrices (K,3,3) -> (K,4) (w,x,y,z): the same four
    branches and the same arithmetic per element (the validity checks are the caller's)."""
    m = np.swapaxes(np.asarray(mats, dtype=np.float64)[:, :3, :3], 1, 2)
    m00, m11, m22 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    neg = m22 < 0
    b0 = neg & (m00 > m11)
    b1 = neg & ~b0
    b2 = ~neg & (m00 < -m11)
    t = np.where(b0, 1 + m00 - m11 - m22, np.where(b1, 1 - m00 + m11 - m22, np.where(b2, 1 - m00 - m11 + m22, 1 + m00 + m11 + m22)))
    d12, d20, d01 = m[:, 1, 2] - m[:, 2, 1], m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] - m[:, 1, 0]
    s12, s20, s01 = m[:, 1, 2] + m[:, 2, 1], m[:, 2, 0] + m[:, 0, 2], m[:, 0, 1] + m[:, 1, 0]
    pick = lambda a, b, c, d: np.where(b0, a, np.where(b1, b, np.where(b2, c, d)))
    q = np.stack([pick(d12, d20, d01, t), pick(t, s01, s20, d12), pick(s01, t, s12, d20), pick(s20, s12, t, d01)], axis=1)
    return q * (0.5 / np.sqrt(t))[:, None]
