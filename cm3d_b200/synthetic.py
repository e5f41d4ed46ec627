"""Seeded synthetic frames of the shapes BASELINE.json names.

numpy only.  The generator builds what the reference's dataset accessors would
hand to the per-frame body: raw sweeps in the sensor frame, fp64 poses cast to
fp32 the way the reference casts them, scaled intrinsics, and Detic-style
instance masks (the projected object hull, dilated a few pixels, overlaps
allowed).  Points inside a sweep are in firing (azimuth) order like a real
spinning-LiDAR file.

Configs: "c1" nuScenes 1 sweep x 6 cams x 20 inst; "c2" nuScenes 10 sweeps x
6 cams x 50 inst; "c3" KITTI 120k pts x 1 cam x 15 inst; "c4" Waymo 180k pts x
5 cams x 80 inst.  `scale` shrinks point counts (and `mask_div` the mask
resolution) for fast parity cases.
"""
from __future__ import annotations

import numpy as np

from .frames import (CamSpec, FrameSpec, FOURTH_COL3, FOURTH_NONE, FOURTH_ONES,
                     op_A, op_R, op_T)

SHAPE_PRIORS = {  # cfg/shape_priors_chatgpt.json: [w, l, h]
    "car": [1.8, 4.5, 1.4], "truck": [2.6, 8.0, 3.6], "bus": [2.5, 12.0, 4.0],
    "trailer": [2.6, 12.0, 3.6], "construction_vehicle": [2.0, 4.5, 2.5],
    "pedestrian": [0.4, 0.7, 1.7], "motorcycle": [0.8, 2.1, 1.7],
    "bicycle": [0.6, 1.8, 1.4], "traffic_cone": [0.3, 0.3, 0.7],
    "barrier": [0.5, 1.2, 0.9],
}
CLASSES = list(SHAPE_PRIORS)

# camera axes (x right, y down, z forward) expressed in a body frame (x fwd, y left, z up)
_CAM_BASE = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])


def _rotz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _roty(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def _rotx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def _pose_R(yaw, pitch, roll):
    return _rotz(yaw) @ _roty(pitch) @ _rotx(roll)


def _beam_table(n, lo_deg, hi_deg):
    return np.deg2rad(np.linspace(lo_deg, hi_deg, n))


def _background(rng, n, beams, height, rmin=1.0, rmax=80.0, near_frac=0.0):
    """Ground-plane hits of a spinning LiDAR at `height`, sensor frame (z up);
    `near_frac` of them are ego-vehicle returns 0.4..2.2 m away (close-point filter food)."""
    az = rng.uniform(-np.pi, np.pi, n)
    ring = rng.integers(0, len(beams), n)
    el = beams[ring] + rng.normal(0.0, 1e-3, n)
    with np.errstate(divide="ignore"):
        r = np.where(el < -1e-3, height / np.sin(-el), rmax)
    r = np.clip(r * (1.0 + rng.normal(0.0, 0.01, n)), rmin, rmax)
    if near_frac > 0.0:
        near = rng.uniform(0.0, 1.0, n) < near_frac
        r = np.where(near, rng.uniform(0.4, 2.2, n), r)
    ce = np.cos(el)
    pts = np.stack([r * ce * np.cos(az), r * ce * np.sin(az), r * np.sin(el)], 1)
    return pts, ring


def _box_surface(rng, n, center, wlh, yaw):
    """n points on the 4 sides + top of a box (w along local y, l along local x)."""
    w, l, h = wlh
    areas = np.array([l * h, l * h, w * h, w * h, l * w])
    face = rng.choice(5, size=n, p=areas / areas.sum())
    u = rng.uniform(-0.5, 0.5, n)
    v = rng.uniform(-0.5, 0.5, n)
    x = np.select([face == 0, face == 1, face == 2, face == 3], [u * l, u * l, 0.5 * l, -0.5 * l], u * l)
    y = np.select([face == 0, face == 1, face == 2, face == 3], [0.5 * w, -0.5 * w, u * w, u * w], v * w)
    z = np.where(face == 4, 0.5 * h, v * h)
    loc = np.stack([x, y, z], 1) + rng.normal(0.0, 0.02, (n, 3))
    return loc @ _rotz(yaw).T + center


def _box_corners(center, wlh, yaw):
    w, l, h = wlh
    s = np.array([[sx, sy, sz] for sx in (-0.5, 0.5) for sy in (-0.5, 0.5) for sz in (-0.5, 0.5)])
    return (s * np.array([l, w, h])) @ _rotz(yaw).T + center


def _convex_hull(pts):
    """Andrew monotone chain, counter-clockwise, pts (n,2)."""
    p = sorted(map(tuple, pts))
    if len(p) <= 2:
        return np.array(p)

    def half(seq):
        out = []
        for q in seq:
            while len(out) >= 2 and ((out[-1][0] - out[-2][0]) * (q[1] - out[-2][1])
                                     - (out[-1][1] - out[-2][1]) * (q[0] - out[-2][0])) <= 0:
                out.pop()
            out.append(q)
        return out

    lo, up = half(p), half(reversed(p))
    return np.array(lo[:-1] + up[:-1])


def _raster_hull(hull, W, H, margin):
    """uint8 (H,W) mask of pixels whose centre is within `margin` px of the hull."""
    m = np.zeros((H, W), np.uint8)
    if len(hull) < 3:
        return m
    x0 = int(max(0, np.floor(hull[:, 0].min() - margin)))
    x1 = int(min(W, np.ceil(hull[:, 0].max() + margin) + 1))
    y0 = int(max(0, np.floor(hull[:, 1].min() - margin)))
    y1 = int(min(H, np.ceil(hull[:, 1].max() + margin) + 1))
    if x1 <= x0 or y1 <= y0:
        return m
    xs, ys = np.meshgrid(np.arange(x0, x1) + 0.5, np.arange(y0, y1) + 0.5)
    inside = np.ones(xs.shape, bool)
    nv = len(hull)
    for k in range(nv):
        a, b = hull[k], hull[(k + 1) % nv]
        e = b - a
        ln = np.hypot(e[0], e[1])
        if ln < 1e-9:
            continue
        # signed distance, positive on the interior side of a CCW polygon
        d = (e[0] * (ys - a[1]) - e[1] * (xs - a[0])) / ln
        inside &= d >= -margin
    m[y0:y1, x0:x1] = inside
    return m


def _project_mask(corners_cam, K_scaled, W, H, margin):
    z = corners_cam[:, 2]
    if np.any(z < 0.5):
        return np.zeros((H, W), np.uint8)
    uv = (corners_cam @ K_scaled.T)
    uv = uv[:, :2] / uv[:, 2:3]
    return _raster_hull(_convex_hull(uv), W, H, margin)


def _alloc_counts(rng, ranges, total, lo=5, hi=6000):
    wgt = 1.0 / np.maximum(ranges, 3.0) ** 2
    wgt = wgt * rng.uniform(0.5, 1.5, len(ranges))
    cnt = np.clip(np.round(total * wgt / wgt.sum()), lo, hi).astype(int)
    return cnt


def _sample_range(rng, wlh, f_scaled, rmin=4.0, rmax=60.0, max_px=300.0):
    """Log-uniform range whose lower end keeps the projected box under ~max_px
    pixels, so mask areas stay within the 200..60,000 px band of SURVEY 8(d)."""
    lo = min(max(rmin, float(max(wlh)) * f_scaled / max_px), 0.8 * rmax)
    return float(np.exp(rng.uniform(np.log(lo), np.log(rmax))))


def _sort_firing(pts_sensor, extra):
    """Firing order: ascending azimuth in the sensor frame."""
    order = np.argsort(np.arctan2(pts_sensor[:, 1], pts_sensor[:, 0]), kind="stable")
    return pts_sensor[order], [e[order] for e in extra]


# ----------------------------------------------------------------------------- nuScenes
_NUSC_CAM_YAWS = np.deg2rad([0.0, -55.0, -110.0, 180.0, 110.0, 55.0])  # CAM_LIST order, nuscenes:62-69


def make_nuscenes_frame(seed, n_sweeps=3, pts_per_sweep=34720, n_inst=20, mask_div=1,
                        obj_frac=0.30, dense_masks=True) -> FrameSpec:
    rng = np.random.default_rng(seed)
    W, H = 1024 // mask_div, 576 // mask_div
    ratio = np.float32(0.64 / mask_div)
    beams = _beam_table(32, -30.0, 10.0)

    # ego trajectory: sample pose + sweeps 50 ms apart going forward in time (nuscenes:437,460-463)
    t0 = np.array([rng.uniform(300, 2000), rng.uniform(300, 2000), rng.normal(0.0, 0.2)])
    yaw0 = rng.uniform(-np.pi, np.pi)
    speed = rng.uniform(0.0, 15.0)
    yaw_rate = rng.normal(0.0, 0.05)

    def ego_at(dt):
        yaw = yaw0 + yaw_rate * dt
        R = _pose_R(yaw, rng.normal(0, 0.01), rng.normal(0, 0.01))
        t = t0 + speed * dt * np.array([np.cos(yaw0), np.sin(yaw0), 0.0])
        return R, t

    R_ls = _rotz(np.deg2rad(-90.0))
    t_ls = np.array([0.94, 0.0, 1.84])
    sweep_pose = [ego_at(0.05 * s) for s in range(n_sweeps)]

    # cameras (own ego pose each: camera timestamps differ from the LiDAR's)
    cams_f64 = []
    for ci, cy in enumerate(_NUSC_CAM_YAWS):
        R_cs = _rotz(cy + rng.normal(0, 0.005)) @ _roty(rng.normal(0, 0.005)) @ _CAM_BASE
        t_cs = np.array([1.5 * np.cos(cy), 0.5 * np.sin(cy), 1.5]) + rng.normal(0, 0.02, 3)
        R_e, t_e = ego_at(rng.uniform(-0.025, 0.025))
        f = 809.2 if ci == 3 else 1266.4
        K = np.array([[f, 0.0, 816.3], [0.0, f, 491.5], [0.0, 0.0, 1.0]])
        cams_f64.append((R_cs, t_cs, R_e, t_e, K))

    # objects, static in the global frame, each placed inside one camera's FOV
    R0, tg0 = sweep_pose[0]
    labels, scores, cam_nums, objs = [], [], [], []
    for i in range(n_inst):
        c = int(rng.integers(0, 6))
        half_fov = np.deg2rad(40.0 if c == 3 else 26.0)
        az = _NUSC_CAM_YAWS[c] + rng.uniform(-half_fov, half_fov)
        name = CLASSES[i % len(CLASSES)]
        wlh = np.array(SHAPE_PRIORS[name])
        rg = _sample_range(rng, wlh, (809.2 if c == 3 else 1266.4) * float(ratio))
        ctr_ego = np.array([rg * np.cos(az), rg * np.sin(az), 0.5 * wlh[2] + rng.normal(0, 0.05)])
        objs.append((R0 @ ctr_ego + tg0, wlh, rng.uniform(-np.pi, np.pi), rg))
        labels.append(name)
        scores.append(float(rng.uniform(0.1, 1.0)))
        cam_nums.append(c)

    n_obj_total = int(round(obj_frac * pts_per_sweep * n_sweeps))
    per_obj = _alloc_counts(rng, np.array([o[3] for o in objs]), n_obj_total) if n_inst else np.zeros(0, int)

    sweeps, sweep_ops = [], []
    for s in range(n_sweeps):
        R_e, t_e = sweep_pose[s]
        obj_pts = []
        for (ctr, wlh, yaw, _), cnt in zip(objs, per_obj):
            k = cnt // n_sweeps + (1 if s < cnt % n_sweeps else 0)
            if k:
                g = _box_surface(rng, k, ctr, wlh, yaw)
                e = (g - t_e) @ R_e                     # R_e^T (g - t_e)
                obj_pts.append((e - t_ls) @ R_ls)
        obj_pts = np.concatenate(obj_pts) if obj_pts else np.zeros((0, 3))
        n_bg = max(pts_per_sweep - len(obj_pts), 0)
        bg, ring = _background(rng, n_bg, beams, 1.84, near_frac=0.015)
        pts = np.concatenate([bg, obj_pts])
        ring = np.concatenate([ring, rng.integers(0, 32, len(obj_pts))])
        inten = rng.uniform(0.0, 255.0, len(pts))
        pts, (ring, inten) = _sort_firing(pts, [ring, inten])
        raw = np.concatenate([pts, inten[:, None], ring[:, None]], 1).astype(np.float32)
        sweeps.append(raw)
        sweep_ops.append([op_R(R_ls), op_T(t_ls), op_R(R_e), op_T(t_e)])     # nuscenes:450-457

    cams, masks = [], np.zeros((n_inst, H, W), np.uint8)
    for (R_cs, t_cs, R_e, t_e, K) in cams_f64:
        K32 = K.astype(np.float32) * ratio                                   # nuscenes:585-587
        K32[2, 2] = 1.0
        cams.append(CamSpec([op_T(-t_e), op_R(R_e.T), op_T(-t_cs), op_R(R_cs.T)], K32))  # :569-577
    for i, ((ctr, wlh, yaw, _), c) in enumerate(zip(objs, cam_nums)):
        R_cs, t_cs, R_e, t_e, K = cams_f64[c]
        cc = ((_box_corners(ctr, wlh, yaw) - t_e) @ R_e - t_cs) @ R_cs
        Ks = K * float(ratio)
        Ks[2, 2] = 1.0
        masks[i] = _project_mask(cc, Ks, W, H, rng.uniform(2.0, 6.0) / mask_div + 1.0)

    return FrameSpec("nuscenes", sweeps, sweep_ops, cams, np.array(cam_nums, np.int32),
                     masks if dense_masks else dense_to_rle(masks), labels, scores,
                     fourth=FOURTH_COL3, close_thresh=float(np.float32(np.sqrt(2.3))),
                     min_dist=2.3, token=f"synth-nusc-{seed}")


# ----------------------------------------------------------------------------- KITTI
_KITTI_CALIB = {
    "P2": "7.215377e+02 0.000000e+00 6.095593e+02 4.485728e+01 0.000000e+00 7.215377e+02 "
          "1.728540e+02 2.163791e-01 0.000000e+00 0.000000e+00 1.000000e+00 2.745884e-03",
    "R0_rect": "9.999239e-01 9.837760e-03 -7.445048e-03 -9.869795e-03 9.999421e-01 "
               "-4.278459e-03 7.402527e-03 4.351614e-03 9.999631e-01",
    "Tr_velo_to_cam": "7.533745e-03 -9.999714e-01 -6.166020e-04 -4.069766e-03 1.480249e-02 "
                      "7.280733e-04 -9.998902e-01 -7.631618e-02 9.998621e-01 7.523790e-03 "
                      "1.480755e-02 -2.717806e-01",
}


def kitti_calib_text() -> str:
    return "".join(f"{k}: {v}\n" for k, v in _KITTI_CALIB.items())


def kitti_calib():
    """The canonical calibration as the host-side `Calibration` object."""
    from .kitti_calib import Calibration
    return Calibration(text=kitti_calib_text())


def make_kitti_frame(seed, n_pts=120000, n_inst=15, mask_div=1, obj_frac=0.30,
                     dense_masks=True) -> FrameSpec:
    rng = np.random.default_rng(seed)
    W, H = 1024 // mask_div, 309 // mask_div
    ratio = np.float32(0.8366 / mask_div)
    beams = _beam_table(64, -24.8, 2.0)
    calib = kitti_calib()
    V2C64, R064 = calib.V2C.numpy().astype(np.float64), calib.R0.numpy().astype(np.float64)

    labels, scores, objs = [], [], []
    for i in range(n_inst):
        az = rng.uniform(-np.deg2rad(33.0), np.deg2rad(33.0))
        name = CLASSES[i % len(CLASSES)]
        wlh = np.array(SHAPE_PRIORS[name])
        rg = _sample_range(rng, wlh, 721.54 * float(ratio), rmin=5.0)
        ctr = np.array([rg * np.cos(az), rg * np.sin(az), -1.73 + 0.5 * wlh[2] + rng.normal(0, 0.05)])
        objs.append((ctr, wlh, rng.uniform(-np.pi, np.pi), rg))
        labels.append(name)
        scores.append(float(rng.uniform(0.1, 1.0)))
    per_obj = _alloc_counts(rng, np.array([o[3] for o in objs]), int(round(obj_frac * n_pts))) if n_inst else []
    obj_pts = [_box_surface(rng, int(k), c, wlh, yaw) for (c, wlh, yaw, _), k in zip(objs, per_obj)]
    obj_pts = np.concatenate(obj_pts) if obj_pts else np.zeros((0, 3))
    bg, _ = _background(rng, max(n_pts - len(obj_pts), 0), beams, 1.73)
    pts = np.concatenate([bg, obj_pts])
    inten = rng.uniform(0.0, 1.0, len(pts))
    pts, (inten,) = _sort_firing(pts, [inten])
    raw = np.concatenate([pts, inten[:, None]], 1).astype(np.float32)

    K32 = calib.scaled_intrinsic(float(ratio))                         # kitti:1259-1266
    cam = CamSpec(calib.cam_ops(), K32)                                # kitti:1238-1240
    masks = np.zeros((n_inst, H, W), np.uint8)
    Ks = K32.astype(np.float64)
    for i, (ctr, wlh, yaw, _) in enumerate(objs):
        cv = _box_corners(ctr, wlh, yaw)
        ref = cv @ V2C64[:, :3].T + V2C64[:, 3]
        masks[i] = _project_mask(ref @ R064.T, Ks, W, H, rng.uniform(2.0, 6.0) / mask_div + 1.0)
    return FrameSpec("kitti", [raw], [calib.sweep_ops()], [cam], np.zeros(n_inst, np.int32),
                     masks if dense_masks else dense_to_rle(masks), labels, scores,
                     fourth=FOURTH_NONE, close_thresh=None, min_dist=2.3,
                     token=f"synth-kitti-{seed}")


# ----------------------------------------------------------------------------- Waymo
_WAYMO_CAM_YAWS = np.deg2rad([0.0, 45.0, -45.0, 90.0, -90.0])   # FRONT, FRONT_LEFT, FRONT_RIGHT, SIDE_LEFT, SIDE_RIGHT


def make_waymo_frame(seed, n_pts=180000, n_inst=80, mask_div=1, obj_frac=0.30,
                     dense_masks=False) -> FrameSpec:
    """Waymo masks differ in size per camera (1024x683 front, 1024x473 side), so
    they are always carried as RLE (the on-disk format, waymo:451-457,520)."""
    rng = np.random.default_rng(seed)
    beams = _beam_table(64, -17.6, 2.4)
    ratio64 = (1024 / 1920) / mask_div                                 # waymo:523, fp64
    sizes = [(1024 // mask_div, (683 if c < 3 else 473) // mask_div) for c in range(5)]
    t_lidar = np.array([1.43, 0.0, 2.18])

    cams_f64, cams = [], []
    for c, cy in enumerate(_WAYMO_CAM_YAWS):
        R = _rotz(cy + rng.normal(0, 0.005)) @ _roty(rng.normal(0, 0.005)) @ _CAM_BASE
        t = np.array([1.5 * np.cos(cy), 0.6 * np.sin(cy), 2.1]) + rng.normal(0, 0.02, 3)
        intr = np.array([2055.0 + rng.normal(0, 5), 2055.0 + rng.normal(0, 5), 940.0 + rng.normal(0, 5),
                         (640.0 if c < 3 else 440.0) + rng.normal(0, 5)], np.float32).astype(np.float64)
        K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1]]) * ratio64
        K[2, 2] = 1.0                                                  # waymo:586-593 (fp64, then cast)
        cams_f64.append((R, t, K))
        cams.append(CamSpec([op_T(-(t.astype(np.float32))), op_R(R.T)], K.astype(np.float32)))  # :573-575

    labels, scores, cam_nums, objs = [], [], [], []
    for i in range(n_inst):
        c = int(rng.integers(0, 5))
        az = _WAYMO_CAM_YAWS[c] + rng.uniform(-np.deg2rad(22.0), np.deg2rad(22.0))
        name = CLASSES[i % len(CLASSES)]
        wlh = np.array(SHAPE_PRIORS[name])
        rg = _sample_range(rng, wlh, 2055.0 * ratio64)
        ctr = np.array([rg * np.cos(az), rg * np.sin(az), 0.5 * wlh[2] + rng.normal(0, 0.05)])
        objs.append((ctr, wlh, rng.uniform(-np.pi, np.pi), rg))
        labels.append(name)
        scores.append(float(rng.uniform(0.1, 1.0)))
        cam_nums.append(c)
    per_obj = _alloc_counts(rng, np.array([o[3] for o in objs]), int(round(obj_frac * n_pts))) if n_inst else []
    obj_pts = [_box_surface(rng, int(k), c, wlh, yaw) for (c, wlh, yaw, _), k in zip(objs, per_obj)]
    obj_pts = np.concatenate(obj_pts) if obj_pts else np.zeros((0, 3))
    bg, _ = _background(rng, max(n_pts - len(obj_pts), 0), beams, 2.18)
    pts_sensor = np.concatenate([bg, obj_pts - t_lidar])
    pts_sensor, _ = _sort_firing(pts_sensor, [])
    raw = (pts_sensor + t_lidar).astype(np.float32)                    # vehicle frame, (N,3)

    rles = []
    for i, ((ctr, wlh, yaw, _), c) in enumerate(zip(objs, cam_nums)):
        R, t, K = cams_f64[c]
        W, H = sizes[c]
        m = _project_mask((_box_corners(ctr, wlh, yaw) - t) @ R, K, W, H,
                          rng.uniform(2.0, 6.0) / mask_div + 1.0)
        rles.append(m)
    masks = [dense_to_rle(m[None])[0] for m in rles]
    return FrameSpec("waymo", [raw], [[]], cams, np.array(cam_nums, np.int32), masks, labels, scores,
                     fourth=FOURTH_ONES, close_thresh=None, min_dist=2.3,
                     token=f"synth-waymo-{seed}")


# ----------------------------------------------------------------------------- RLE helpers
def compress_rles(rles):
    """RLEMasks with uint32 run arrays -> RLEMasks with pycocotools' compressed `counts` bytes,
    the on-disk format of {f}_masks.pkl (gen_2d_masks_detic.py:471,506)."""
    from .frames import RLEMask
    from .rle import rle_counts_to_runs, runs_to_rle_string
    return [RLEMask(r.size, runs_to_rle_string(rle_counts_to_runs(r.counts))) for r in rles]


def dense_to_rle(masks_hw: np.ndarray):
    """(I,H,W) uint8 -> list[RLEMask] with uncompressed uint32 counts.

    COCO RLE of the (W,H) array in column-major order == row-major scan of the
    (H,W) image (gen_2d_masks_detic.py:468-471)."""
    from .frames import RLEMask
    out = []
    for m in masks_hw:
        H, W = m.shape
        flat = (m.reshape(-1) != 0).astype(np.int8)
        edges = np.flatnonzero(np.diff(flat)) + 1
        bounds = np.concatenate([[0], edges, [flat.size]])
        runs = np.diff(bounds).astype(np.uint32)
        if flat.size and flat[0] == 1:
            runs = np.concatenate([np.zeros(1, np.uint32), runs])
        out.append(RLEMask((W, H), runs))
    return out


def rle_to_dense(rle) -> np.ndarray:
    """RLEMask -> (H,W) uint8 (host helper for tests; the product decodes on the GPU)."""
    from .rle import rle_counts_to_runs
    W, H = rle.size
    runs = rle_counts_to_runs(rle.counts)
    vals = np.zeros(len(runs), np.uint8)
    vals[1::2] = 1
    flat = np.repeat(vals, runs)
    if flat.size != W * H:
        raise ValueError("RLE does not cover the mask")
    return flat.reshape(H, W)


# ----------------------------------------------------------------------------- configs
def make_frame(config: str, index: int = 0, scale: float = 1.0, mask_div: int = 1,
               dense_masks=None) -> FrameSpec:
    """Frame `index` of BASELINE config c1..c5; seed = 1000*config + index (c5: 5_000_000+index)."""
    k = config.lower()
    sc = lambda n: max(int(round(n * scale)), 64)
    if k == "c1":
        return make_nuscenes_frame(1000 + index, 1, sc(34720), 20, mask_div,
                                   dense_masks=True if dense_masks is None else dense_masks)
    if k in ("c2", "c5"):
        seed = 2000 + index if k == "c2" else 5_000_000 + index
        return make_nuscenes_frame(seed, 10, sc(34720), 50, mask_div,
                                   dense_masks=True if dense_masks is None else dense_masks)
    if k == "c3":
        return make_kitti_frame(3000 + index, sc(120000), 15, mask_div,
                                dense_masks=True if dense_masks is None else dense_masks)
    if k == "c4":
        return make_waymo_frame(4000 + index, sc(180000), 80, mask_div)
    raise ValueError(f"unknown config {config!r}")
