"""ORACLE (test infrastructure): stand-in for `pyquaternion.Quaternion` (pinned 0.9.9,
environment.yml:131), the un-vendored dependency the reference uses to turn dataset poses into
matrices (src/nuscenes/2d_to_3d.py:451,456,571,577; src/waymo/2d_to_3d.py:575,689) and lane-aligned
matrices into box quaternions (nuscenes:796,806; waymo Quaternion(matrix=...)).

Restated from pyquaternion's published source: elements (w,x,y,z); `rotation_matrix` first
normalises (only when |1 - |q|^2| >= 1e-14), then returns rows/columns 1..3 of Q(q) Qbar(q)^H;
`Quaternion(matrix=M)` checks orthogonality / determinant with numpy's default tolerances and uses
the four-branch trace method on M^T.  PARITY UNPINNED against the package itself (absent here).
Independent of cm3d_b200/quat.py; tests compare the two and scipy's Rotation.
"""
from __future__ import annotations

from math import sqrt

import numpy as np


class Quaternion:
    def __init__(self, *args, **kwargs):
        if "matrix" in kwargs:
            self.q = self._from_matrix(np.asarray(kwargs["matrix"], dtype=np.float64))
            return
        if len(args) == 0:
            self.q = np.array([1.0, 0.0, 0.0, 0.0])
        elif len(args) == 1:
            a = args[0]
            if isinstance(a, Quaternion):
                self.q = a.q.copy()
            else:
                v = np.array([float(e) for e in a], dtype=np.float64)
                if v.shape != (4,):
                    raise ValueError("a quaternion needs 4 elements")
                self.q = v
        elif len(args) == 4:
            self.q = np.array([float(e) for e in args], dtype=np.float64)
        else:
            raise ValueError("unsupported Quaternion constructor arguments")

    # ---- sequence protocol: list(q) == [w, x, y, z]
    def __iter__(self):
        return iter(self.q)

    def __len__(self):
        return 4

    def __getitem__(self, k):
        return self.q[k]

    @property
    def elements(self):
        return self.q

    @property
    def rotation_matrix(self):
        ss = float(np.dot(self.q, self.q))
        if not abs(1.0 - ss) < 1e-14:
            n = sqrt(ss)
            if n > 0:
                self.q = self.q / n
        w, x, y, z = self.q
        Q = np.array([[w, -x, -y, -z],
                      [x, w, -z, y],
                      [y, z, w, -x],
                      [z, -y, x, w]])
        Qbar = np.array([[w, -x, -y, -z],
                         [x, w, z, -y],
                         [y, -z, w, x],
                         [z, y, -x, w]])
        P = np.dot(Q, Qbar.conj().transpose())
        return P[1:][:, 1:]

    @staticmethod
    def _from_matrix(M, rtol=1e-05, atol=1e-08):
        if M.shape not in ((3, 3), (4, 4)):
            raise ValueError("Invalid matrix shape: Input must be a 3x3 or 4x4 numpy array or matrix")
        Rm = M[:3, :3]
        if not np.allclose(np.dot(Rm, Rm.conj().transpose()), np.eye(3), rtol=rtol, atol=atol):
            raise ValueError("Matrix must be orthogonal, i.e. its transpose should be its inverse")
        if not np.isclose(np.linalg.det(Rm), 1.0, rtol=rtol, atol=atol):
            raise ValueError("Matrix must be special orthogonal i.e. its determinant must be +1.0")
        m = Rm.conj().transpose()
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
