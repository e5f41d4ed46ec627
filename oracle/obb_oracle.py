"""ORACLE (test infrastructure): the KITTI box / yaw step
(src/kitti/2d_to_3d.py:855-876 `get_depth_bbox`, :1481-1484 fallback, :1524 yaw).

The reference calls open3d 0.15.2 `PointCloud.get_oriented_bounding_box()` (environment.yml:111),
an un-vendored dependency that is absent here.  `open3d_obb` restates its published algorithm
(open3d/geometry/BoundingVolume.cpp, OrientedBoundingBox::CreateFromPoints, v0.15):

    hull      = Qhull convex hull of the points (option Qt)      -> here scipy.spatial.ConvexHull,
                which wraps the same Qhull library with Qt always on
    mean, cov = ComputeMeanAndCovariance of the HULL VERTICES    (cov = E[xx^T] - E[x]E[x]^T, /N)
    R         = eigenvectors of cov (Eigen SelfAdjointEigenSolver), columns reordered so that the
                eigenvalues descend, col0 and col1 normalised, col2 = col0 x col1
    extent    = axis-aligned extent of R^T (hull - mean);  center = R * aabb_centre + mean

PARITY UNPINNED against the open3d binary: the SIGN of Eigen's eigenvectors is implementation
defined; here (and in csrc/obb.cu) an eigenvector is oriented so that its largest-magnitude
component is positive.  A hull Qhull rejects (flat / fewer than 4 points) makes the reference's
bare `except` substitute the identity (kitti:1483-1484); `get_depth_bbox_or_fallback` does that.
Everything after the box (axis shuffle by axis-aligned size, `as_euler('zyx')[0]`) follows the
reference line by line.  `allpoint_axes` is round 1's estimator (PCA of all member points), kept only to
report how far it was from the hull-vertex one.
"""
import numpy as np


def hull_vertex_indices(pts3d: np.ndarray) -> np.ndarray:
    """Indices (ascending) of the convex-hull vertices, by Qhull; raises on a flat / tiny input."""
    from scipy.spatial import ConvexHull
    p = np.asarray(pts3d, np.float64)
    return np.sort(ConvexHull(p).vertices)


def _oriented(e):
    e = e / np.linalg.norm(e)
    return -e if e[np.argmax(np.abs(e))] < 0 else e


def axes_from_vertices(v: np.ndarray):
    """mean / covariance / eigenvectors / extent of a vertex set, as CreateFromPoints does them."""
    v = np.asarray(v, np.float64)
    mean = v.mean(0)
    cov = (v.T @ v) / len(v) - np.outer(mean, mean)
    w, vec = np.linalg.eigh(cov)
    order = np.argsort(-w, kind="stable")
    R = np.zeros((3, 3))
    R[:, 0] = _oriented(vec[:, order[0]])
    R[:, 1] = _oriented(vec[:, order[1]])
    R[:, 2] = np.cross(R[:, 0], R[:, 1])
    q = (v - mean) @ R
    lo, hi = q.min(0), q.max(0)
    center = R @ (0.5 * (lo + hi)) + mean
    return center, hi - lo, R


def open3d_obb(pts3d: np.ndarray):
    """(center, extent, R) of open3d 0.15's oriented bounding box of the points."""
    p = np.asarray(pts3d, np.float64)
    return axes_from_vertices(p[hull_vertex_indices(p)])


def allpoint_axes(pts3d: np.ndarray):
    """Round-1 estimator: PCA of ALL member points (not what open3d does)."""
    return axes_from_vertices(np.asarray(pts3d, np.float64))


def get_depth_bbox(pts3d: np.ndarray, obb=open3d_obb):
    """kitti:855-876 with `point_cloud.get_oriented_bounding_box()` = obb(pts3d)."""
    center, extent, R = obb(pts3d)
    x_size = pts3d[:, 0].max() - pts3d[:, 0].min()
    y_size = pts3d[:, 1].max() - pts3d[:, 1].min()
    z_size = pts3d[:, 2].max() - pts3d[:, 2].min()
    axis = [ax[1] for ax in sorted([(x_size, "x"), (y_size, "y"), (z_size, "z")], key=lambda x: x[0])]
    wlh = np.asarray(extent).tolist()
    wlh = [wlh[axis.index("x")], wlh[axis.index("y")], wlh[axis.index("z")]]
    R = np.stack([R[:, axis.index("z")], R[:, axis.index("y")], R[:, axis.index("x")]], axis=1)
    return np.asarray(center).tolist(), wlh, R


def get_depth_bbox_or_fallback(pts3d: np.ndarray):
    """kitti:1481-1484: any exception inside get_depth_bbox (Qhull on a flat cloud) -> identity box."""
    try:
        return get_depth_bbox(pts3d)
    except Exception:
        return [pts3d[0], np.array([1, 1, 1]), np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]])]


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
