"""ORACLE (test infrastructure): numpy restatement of the KITTI box / yaw step
(src/kitti/2d_to_3d.py:855-876 `get_depth_bbox`, :1524 yaw).

PARITY UNPINNED: the reference calls open3d 0.15.2 `PointCloud.get_oriented_bounding_box()`
(environment.yml:111), an un-vendored dependency that is absent here; its published algorithm is a
PCA of the convex-hull vertices (Qhull).  This restates the convention csrc/obb.cu documents - PCA of
the member points themselves - so it pins the CUDA kernel to a written-down definition, not to open3d.
Everything after the eigen-decomposition (axis shuffle by axis-aligned size, `as_euler('zyx')[0]`)
follows the reference line by line and uses scipy's Rotation like the reference does.
"""
import numpy as np
from scipy.spatial.transform import Rotation


def principal_axes(pts3d: np.ndarray):
    p = np.asarray(pts3d, np.float64)
    mean = p.mean(0)
    d = p - mean
    cov = d.T @ d / len(p)
    w, v = np.linalg.eigh(cov)
    order = np.argsort(-w, kind="stable")
    R = np.zeros((3, 3))
    for col in range(2):
        e = v[:, order[col]] / np.linalg.norm(v[:, order[col]])
        if e[np.argmax(np.abs(e))] < 0:
            e = -e
        R[:, col] = e
    R[:, 2] = np.cross(R[:, 0], R[:, 1])
    q = d @ R
    lo, hi = q.min(0), q.max(0)
    center = R @ (0.5 * (lo + hi)) + mean
    return center, hi - lo, R


def get_depth_bbox(pts3d: np.ndarray):
    """kitti:855-876 with `obb` replaced by principal_axes()."""
    center, extent, R = principal_axes(pts3d)
    x_size = pts3d[:, 0].max() - pts3d[:, 0].min()
    y_size = pts3d[:, 1].max() - pts3d[:, 1].min()
    z_size = pts3d[:, 2].max() - pts3d[:, 2].min()
    axis = [ax[1] for ax in sorted([(x_size, "x"), (y_size, "y"), (z_size, "z")], key=lambda x: x[0])]
    wlh = extent.tolist()
    wlh = [wlh[axis.index("x")], wlh[axis.index("y")], wlh[axis.index("z")]]
    R = np.stack([R[:, axis.index("z")], R[:, axis.index("y")], R[:, axis.index("x")]], axis=1)
    return center.tolist(), wlh, R


def scipy_from_matrix_quat(m: np.ndarray) -> np.ndarray:
    """scipy==1.11.4 (the reference's pin, environment.yml) `Rotation.from_matrix`: quaternion
    (x,y,z,w) from the matrix entries by the largest of (m00, m11, m22, trace), then normalised.
    No determinant check in that version: the reference feeds it the axis-shuffled R', which is
    left-handed whenever the shuffle is an odd permutation, and gets whatever this arithmetic gives.
    (Newer scipy raises on such input, so this published algorithm is restated, not imported.)"""
    m = np.asarray(m, np.float64)
    decision = np.array([m[0, 0], m[1, 1], m[2, 2], m[0, 0] + m[1, 1] + m[2, 2]])
    choice = int(np.argmax(decision))
    q = np.empty(4)
    if choice != 3:
        i = choice
        j = (i + 1) % 3
        k = (j + 1) % 3
        q[i] = 1 - decision[3] + 2 * m[i, i]
        q[j] = m[j, i] + m[i, j]
        q[k] = m[k, i] + m[i, k]
        q[3] = m[k, j] - m[j, k]
    else:
        q[0] = m[2, 1] - m[1, 2]
        q[1] = m[0, 2] - m[2, 0]
        q[2] = m[1, 0] - m[0, 1]
        q[3] = 1 + decision[3]
    return q / np.linalg.norm(q)


def yaw_of(R: np.ndarray) -> float:
    """kitti:1524 - `R.from_matrix(bbox[2]).as_euler('zyx')[0]`: first extrinsic z angle of the
    rotation the quaternion above stands for, atan2(-M01, M00) with M its rotation matrix."""
    x, y, z, w = scipy_from_matrix_quat(R)
    return float(np.arctan2(2.0 * (z * w - x * y), 1.0 - 2.0 * (y * y + z * z)))
