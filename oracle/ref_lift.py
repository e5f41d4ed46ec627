"""ORACLE (test infrastructure, not product): torch-CPU restatement of the
reference's per-frame lifting body.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this.  It keeps the reference's cost structure (the
whole cloud is cloned, transformed and projected again for EVERY mask; the
medoid materialises an MxM `torch.cdist`) and its torch calls, so it inherits
torch-CPU numerics.  Each block cites the reference lines it follows:

  nuScenes  /root/reference/src/nuscenes/2d_to_3d.py:433-465 (sweeps),
            :510-628 (mask loop), :116-119,641-663 (medoid)
  KITTI     /root/reference/src/kitti/2d_to_3d.py:1066-1083,1129-1524
  Waymo     /root/reference/src/waymo/2d_to_3d.py:472-486,510-653
  helpers   /root/reference/src/nuscenes/utils/pcd.py:159-172,262-284,
            /root/reference/src/kitti/kitti_utils.py:212-249

Parity pin: oracle/make_golden.py runs these functions with the reference's own
`utils/pcd.py` and `kitti_utils.Calibration` injected (imported from
/root/reference) and commits the inputs+outputs under tests/golden/; the
restated helpers below and the C oracle (lift_oracle.c) are checked against
those fixtures bit for bit.  Not pinned by any reference test (it has none):
pycocotools RLE decoding, open3d's OBB (KITTI yaw) - see DESIGN.md.
"""
from __future__ import annotations

import time
from types import SimpleNamespace

import numpy as np
import torch

try:  # cv2 is what the reference erodes with (nuscenes:526-527); present in this image
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

from oracle.coco_rle import counts_to_runs as rle_counts_to_runs

# what row 3 of aggr_pc_points holds (the FrameSpec data model's `fourth` field): nothing (KITTI keeps
# (N,3) rows), column 3 of the raw scan (nuScenes intensity), or ones (waymo:477-479)
FOURTH_NONE, FOURTH_COL3, FOURTH_ONES = 0, 1, 2
FrameSpec = object      # duck-typed: sweeps, sweep_ops, cams[].ops/.K, cam_nums, masks, fourth, close_thresh, min_dist


# --------------------------------------------------------------------------- restated L1 helpers
class _PointCloud:
    """pcd.py:148-172 - in-place translate / rotate on a (4,N) fp32 tensor."""

    def __init__(self, points):
        self.points = points

    def translate(self, x):
        for i in range(3):
            self.points[i, :] = self.points[i, :] + x[i]

    def rotate(self, rot_matrix):
        self.points[:3, :] = torch.matmul(rot_matrix, self.points[:3, :])


def _view_points(points, view, normalize, device="cpu"):
    """pcd.py:262-284 - pad K into a 4x4, multiply homogeneous points, divide by row 2."""
    assert view.shape[0] <= 4 and view.shape[1] <= 4 and points.shape[0] == 3
    viewpad = torch.eye(4).to(device=device, dtype=torch.float32)
    viewpad[:view.shape[0], :view.shape[1]] = view
    n = points.shape[1]
    points = torch.concatenate((points, torch.ones((1, n)).to(device=device, dtype=torch.float32)))
    points = torch.matmul(viewpad, points)
    points = points[:3, :]
    depths = torch.clone(points[2, :])
    if normalize:
        points = points / points[2:3, :].repeat(3, 1).reshape(3, n)
    return points, depths


PCD = SimpleNamespace(LidarPointCloud=_PointCloud, view_points=_view_points)


class _KittiCalib:
    """kitti_utils.py:212-249 - the three projections the hot path calls."""

    def __init__(self, V2C, C2V, R0):
        self.V2C, self.C2V, self.R0 = V2C, C2V, R0

    def cart2hom(self, pts):
        return torch.hstack((pts, torch.ones((pts.shape[0], 1), device=pts.device)))

    def project_velo_to_ref(self, pts):
        return torch.matmul(self.cart2hom(pts), self.V2C.transpose(0, 1))

    def project_ref_to_velo(self, pts):
        return torch.matmul(self.cart2hom(pts), self.C2V.transpose(0, 1))

    def project_ref_to_rect(self, pts):
        return torch.matmul(self.R0, pts.T).T

    def project_velo_to_rect(self, pts):
        return self.project_ref_to_rect(self.project_velo_to_ref(pts))


def get_medoid(points):
    """nuscenes:116-119 / kitti:177-180 / waymo:120-122."""
    dist_matrix = torch.cdist(points.T, points.T, p=2)
    return torch.argmin(dist_matrix.sum(axis=0))


# --------------------------------------------------------------------------- masks
def decode_masks(frame: FrameSpec):
    """-> list of (H,W) uint8, what `depth_images[i]` is after nuscenes:425-428."""
    if isinstance(frame.masks, np.ndarray):
        return [frame.masks[i] for i in range(frame.n_instances)]
    out = []
    for rle in frame.masks:
        W, H = rle.size
        runs = rle_counts_to_runs(rle.counts)
        vals = np.zeros(len(runs), np.uint8)
        vals[1::2] = 1
        out.append(np.repeat(vals, runs).reshape(H, W))
    return out


def erode_to_wh_bool(maskarr):
    """nuscenes:526-527,543-544 - 3x3 erosion, bool, transposed to (W,H)."""
    kernel = np.ones((3, 3), np.uint8)
    maskarr = cv2.erode(np.ascontiguousarray(maskarr), kernel)
    maskarr = maskarr[:, :].astype(bool)
    return torch.transpose(torch.from_numpy(maskarr).to(dtype=bool), 1, 0)


def _apply_ops(pc, ops):
    for kind, m in ops:
        if kind == "R":
            pc.rotate(torch.from_numpy(np.array(m)).to(dtype=torch.float32))
        elif kind == "T":
            pc.translate(torch.from_numpy(np.array(m)).to(dtype=torch.float32))
        else:
            raise ValueError("column-major clouds only take R/T ops")


def _membership(points, depths, image_mask, min_dist, n_total):
    """nuscenes:592-617 - bounds/depth test, floor, mask lookup with the `!=0` quirk.

    Returns (track_points, in_image_idx, floored[:2])."""
    masked_pixels = (image_mask == 1)
    track_points = np.array(range(n_total))
    points_within_image = torch.logical_and(torch.logical_and(torch.logical_and(torch.logical_and(
        depths > min_dist,
        points[0] > 0),
        points[0] < image_mask.shape[0] - 1),
        points[1] > 0),
        points[1] < image_mask.shape[1] - 1)
    floored_points = torch.floor(points[:, points_within_image]).to(dtype=int)
    track_points = track_points[points_within_image.cpu()]
    in_image = track_points
    points_within_mask = torch.logical_and(
        floored_points,
        masked_pixels[floored_points[0], floored_points[1]])
    indices_within_mask = torch.where(torch.logical_and(torch.logical_and(
        points_within_mask[0, :], points_within_mask[1, :]), points_within_mask[2, :]))[0]
    track_points = track_points[indices_within_mask.cpu()]
    return np.atleast_1d(track_points), in_image, floored_points[:2].numpy()


def _new_result(frame, aggr, n):
    I = frame.n_instances
    return {
        "aggr": aggr, "n_points": n,
        "idx": [np.zeros(0, np.int64) for _ in range(I)],
        "medoid_local": np.full(I, -1, np.int64),
        "medoid_point_idx": np.full(I, -1, np.int64),
        "centroids": np.full((I, 3), np.nan, np.float32),
        "pix": {},      # cam -> (in_image_idx, fx, fy)
        "timer": {"points in mask": 0.0, "medoid": 0.0, "aggregate": 0.0},
    }


# --------------------------------------------------------------------------- nuScenes / Waymo
def _lift_columns(frame: FrameSpec, pcd, record_pix, do_medoid):
    """Column-major (4,N) clouds: nuScenes and Waymo bodies."""
    LidarPointCloud, view_points = pcd.LidarPointCloud, pcd.view_points
    t0 = time.perf_counter()
    aggr_set = []
    for raw, ops in zip(frame.sweeps, frame.sweep_ops):
        if frame.fourth == FOURTH_COL3:
            # LidarPointCloud.from_file: flat scan -> (-1,5)[:, :4].T   (pcd.py:246-257)
            scan = torch.from_numpy(raw.reshape(-1))
            lidar_points = scan.reshape((-1, raw.shape[1]))[:, :4].T
        elif frame.fourth == FOURTH_ONES:
            # waymo:477-479 - xyz + a ones row, float64 hstack then cast
            ones = np.ones(raw.shape[0]).reshape(raw.shape[0], 1)
            lidar_points = torch.from_numpy(np.hstack([raw[:, :3], ones]).transpose()).to(dtype=torch.float32)
        else:
            raise ValueError("row-major frames go through lift_kitti")
        if frame.close_thresh is not None:
            # nuscenes:441-446 (np.sqrt(min_dist) is a float64 scalar; torch compares in fp32)
            thr = np.sqrt(frame.min_dist)
            mask = torch.ones(lidar_points.shape[1])
            mask = torch.logical_and(mask, torch.abs(lidar_points[0, :]) < thr)
            mask = torch.logical_and(mask, torch.abs(lidar_points[1, :]) < thr)
            lidar_points = lidar_points[:, ~mask]
        elif not lidar_points.is_contiguous():
            lidar_points = lidar_points.clone()
        pc = LidarPointCloud(lidar_points)
        _apply_ops(pc, ops)                                  # nuscenes:450-457
        aggr_set.append(pc.points)
    aggr = torch.hstack(tuple(aggr_set))                     # nuscenes:465
    n = aggr.shape[1]
    res = _new_result(frame, aggr.numpy(), n)
    res["timer"]["aggregate"] = time.perf_counter() - t0

    depth_images = decode_masks(frame)
    for i in range(frame.n_instances):
        c = int(frame.cam_nums[i])
        cam = frame.cams[c]
        image_mask = erode_to_wh_bool(depth_images[i])       # (W,H)
        t1 = time.perf_counter()
        cam_pc = LidarPointCloud(torch.clone(aggr))          # nuscenes:553
        _apply_ops(cam_pc, cam.ops)                          # nuscenes:569-577 / waymo:573-575
        depths = cam_pc.points[2, :]
        K = torch.from_numpy(cam.K).to(dtype=torch.float32)
        points, _ = view_points(cam_pc.points[:3, :], K, normalize=True, device="cpu")
        track, in_img, floored = _membership(points, depths, image_mask, frame.min_dist, n)
        if record_pix and c not in res["pix"]:
            res["pix"][c] = (in_img.astype(np.int64), floored[0].copy(), floored[1].copy())
        global_masked_points = aggr[:, track]                # nuscenes:620
        res["timer"]["points in mask"] += time.perf_counter() - t1
        res["idx"][i] = track.astype(np.int64)
        if global_masked_points.numel() == 0:
            continue
        if do_medoid:
            t2 = time.perf_counter()
            m = int(get_medoid(global_masked_points[:3, :].to(dtype=torch.float32)))
            res["medoid_local"][i] = m
            res["medoid_point_idx"][i] = track[m]
            res["centroids"][i] = global_masked_points[:3, m].numpy()
            res["timer"]["medoid"] += time.perf_counter() - t2
    return res


# --------------------------------------------------------------------------- KITTI
def _lift_kitti(frame: FrameSpec, calib, pcd, record_pix, do_medoid):
    view_points = pcd.view_points
    t0 = time.perf_counter()
    raw = frame.sweeps[0]
    velo = torch.from_numpy(raw).to(dtype=torch.float32)
    pc = calib.project_velo_to_ref(velo[:, :3])              # kitti:1066-1077
    aggr = torch.hstack((pc,))
    n = aggr.shape[0]
    res = _new_result(frame, aggr.numpy(), n)
    res["timer"]["aggregate"] = time.perf_counter() - t0
    cam = frame.cams[0]
    depth_images = decode_masks(frame)
    for i in range(frame.n_instances):
        image_mask = erode_to_wh_bool(depth_images[i])
        t1 = time.perf_counter()
        depths_ = calib.project_ref_to_velo(aggr)            # kitti:1238-1240
        cam_pc_pts = calib.project_velo_to_rect(depths_)
        depths = cam_pc_pts[:, 2]
        K = torch.from_numpy(cam.K).to(dtype=torch.float32)
        cam_pc_pts = cam_pc_pts.T[:3, :]
        points, _ = view_points(cam_pc_pts, K, normalize=True, device="cpu")
        track, in_img, floored = _membership(points, depths, image_mask, frame.min_dist, n)
        if record_pix and 0 not in res["pix"]:
            res["pix"][0] = (in_img.astype(np.int64), floored[0].copy(), floored[1].copy())
        global_masked_points = aggr[track, :]                # kitti:1369
        res["timer"]["points in mask"] += time.perf_counter() - t1
        res["idx"][i] = track.astype(np.int64)
        if global_masked_points.numel() == 0:
            continue
        gmp = global_masked_points.T                         # (3,M)
        if gmp.shape[1] <= 3:                                # kitti:1479-1480
            continue
        if do_medoid:
            t2 = time.perf_counter()
            m = int(get_medoid(gmp[:3, :].to(dtype=torch.float32)))
            res["medoid_local"][i] = m
            res["medoid_point_idx"][i] = track[m]
            res["centroids"][i] = gmp[:3, m].numpy()
            res["timer"]["medoid"] += time.perf_counter() - t2
    return res


def kitti_calib_from_frame(frame: FrameSpec):
    ops = frame.cams[0].ops
    V2C = torch.from_numpy(frame.sweep_ops[0][0][1])
    C2V = torch.from_numpy(ops[0][1])
    R0 = torch.from_numpy(ops[2][1])
    return _KittiCalib(V2C, C2V, R0)


def lift_frame(frame: FrameSpec, pcd=PCD, calib=None, record_pix=True, do_medoid=True):
    """One frame through the reference body.  `pcd` / `calib` let make_golden.py
    inject the reference's own utils/pcd.py module and kitti_utils.Calibration."""
    if frame.dataset == "kitti":
        return _lift_kitti(frame, calib or kitti_calib_from_frame(frame), pcd, record_pix, do_medoid)
    return _lift_columns(frame, pcd, record_pix, do_medoid)
