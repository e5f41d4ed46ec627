"""ORACLE (test infrastructure): numpy restatement of the default-off extensions in
cm3d_b200/csrc/extras.cu and of the ground threshold in k_aggregate.  The reference never executes
these steps (its `clusters_hdbscan`, src/kitti/2d_to_3d.py:159-174, is dead code and its ground
filter, :1186-1190, is commented out), so there is nothing of the reference to pin them to:
PARITY UNPINNED; the GPU kernels are graded against this file only.
"""
import numpy as np


def ground_filter(aggr, floor_thresh):
    """`aggr_pc_points[:, aggr_pc_points[2] > floor_thresh]` (kitti:1186-1190, commented out there)."""
    aggr = np.asarray(aggr)
    return aggr[:, aggr[2] > np.float32(floor_thresh)]


def neighbor_keep(xyz_3m, radius, min_neighbors, chunk=512):
    """keep[j] = #{i : (dx*dx + dy*dy) + dz*dz <= fl(r*r)} >= min_neighbors, all in float32
    (bit-exact with k_neighbor_count: every operation is one IEEE binary32 rounding)."""
    p = np.ascontiguousarray(xyz_3m, np.float32)
    m = p.shape[1]
    r2 = np.float32(radius) * np.float32(radius)
    cnt = np.zeros(m, np.int64)
    for a in range(0, m, chunk):
        dx = p[0][:, None] - p[0][None, a:a + chunk]
        dy = p[1][:, None] - p[1][None, a:a + chunk]
        dz = p[2][:, None] - p[2][None, a:a + chunk]
        d2 = (dx * dx + dy * dy) + dz * dz
        cnt[a:a + chunk] = (d2 <= r2).sum(0)
    return cnt >= min_neighbors


def box_search(xyz_3m, up_axis=2, n_angles=90):
    """Min-footprint heading search: (centre xyz, extents (along, across, up), theta, area), float32."""
    p = np.ascontiguousarray(xyz_3m, np.float32)
    ia, ib = (up_axis + 1) % 3, (up_axis + 2) % 3
    a, b = p[ia], p[ib]
    best = None
    for k in range(n_angles):
        th = np.float32(k) * (np.float32(1.57079632679489662) / np.float32(n_angles))
        cs, sn = np.cos(th), np.sin(th)
        u = a * cs + b * sn
        v = b * cs - a * sn
        ext = (u.min(), u.max(), v.min(), v.max())
        area = (ext[1] - ext[0]) * (ext[3] - ext[2])
        if best is None or area < best[0]:
            best = (area, th, ext, cs, sn)
    area, th, ext, cs, sn = best
    uc, vc = np.float32(0.5) * (ext[0] + ext[1]), np.float32(0.5) * (ext[2] + ext[3])
    c = np.zeros(3, np.float32)
    c[ia] = uc * cs - vc * sn
    c[ib] = uc * sn + vc * cs
    c[up_axis] = np.float32(0.5) * (p[up_axis].min() + p[up_axis].max())
    return c, np.array([ext[1] - ext[0], ext[3] - ext[2], p[up_axis].max() - p[up_axis].min()], np.float32), float(th), float(area)
