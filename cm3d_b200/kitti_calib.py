"""KITTI calibration, host side (mirrors src/kitti/kitti_utils.py:114-191,368-375).

Only the constants are built here - with the same torch ops the reference uses,
so V2C / C2V / R0 / f_u.. carry the reference's fp32 bits.  The point math
(`project_velo_to_ref`, `project_ref_to_velo`, `project_velo_to_rect`) is not
re-implemented on the CPU: it runs as 'A'/'R' ops in the CUDA chain
(csrc/lift.cu), see `cam_ops()` / `sweep_ops()`.
"""
from __future__ import annotations

import numpy as np
import torch

from .frames import op_A, op_R


def inverse_rigid_trans(Tr: torch.Tensor) -> torch.Tensor:
    """[R|t] -> [R'|-R't], 3x4 (kitti_utils.py:368-375)."""
    inv_Tr = torch.zeros_like(Tr)
    Rt = Tr[0:3, 0:3].transpose(0, 1)
    inv_Tr[0:3, 0:3] = Rt
    inv_Tr[0:3, 3] = torch.matmul(-Rt, Tr[0:3, 3])
    return inv_Tr


def read_calib_text(text: str) -> dict:
    """Calibration-file lines `key: v0 v1 ...` -> fp32 tensors (kitti_utils.py:172-191)."""
    data = {}
    for line in text.splitlines():
        line = line.rstrip()
        if not line:
            continue
        key, value = line.split(":", 1)
        try:
            data[key] = torch.from_numpy(np.array([float(x) for x in value.split()])).to(dtype=torch.float32)
        except ValueError:
            pass
    return data


class Calibration:
    """Same attribute names as the reference class (P, V2C, C2V, R0, c_u, c_v, f_u, f_v, b_x, b_y)."""

    def __init__(self, calib_filepath=None, text=None):
        if text is None:
            with open(calib_filepath, "r") as f:
                text = f.read()
        calibs = read_calib_text(text)
        self.P = calibs["P2"].view([3, 4])
        self.V2C = calibs["Tr_velo_to_cam"].view([3, 4])
        self.C2V = inverse_rigid_trans(self.V2C)
        self.R0 = calibs["R0_rect"].view([3, 3])
        self.c_u = self.P[0, 2]
        self.c_v = self.P[1, 2]
        self.f_u = self.P[0, 0]
        self.f_v = self.P[1, 1]
        self.b_x = self.P[0, 3] / (-self.f_u)
        self.b_y = self.P[1, 3] / (-self.f_v)

    def sweep_ops(self):
        """velo -> reference camera, once per frame (kitti/2d_to_3d.py:1066-1077)."""
        return [op_A(self.V2C.numpy())]

    def cam_ops(self):
        """ref -> velo -> ref -> rect, redone per mask in the reference (kitti/2d_to_3d.py:1238-1240)."""
        return [op_A(self.C2V.numpy()), op_A(self.V2C.numpy()), op_R(self.R0.numpy())]

    def scaled_intrinsic(self, ratio: float) -> np.ndarray:
        """[[f_u,0,c_u],[0,f_v,c_v],[0,0,1]] * ratio, [2,2]=1 (kitti/2d_to_3d.py:1259-1266)."""
        K = torch.Tensor([[self.f_u, 0, self.c_u], [0, self.f_v, self.c_v], [0, 0, 1]]).to(dtype=torch.float32)
        K = K * ratio
        K[2, 2] = 1
        return K.numpy()
