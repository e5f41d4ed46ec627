"""ORACLE tooling: execute the reference's own helper FUNCTIONS (the `FunctionDef`s of
/root/reference/src/{nuscenes,kitti,waymo}/2d_to_3d.py, extracted with `ast` and exec'd unmodified with
numpy / scipy / torch in their namespace) on seeded inputs and commit inputs + outputs as
tests/golden/pass2_functions.json.  The tests grade cm3d_b200/boxes.py, cm3d_nearest_lane, the
medoid kernels and oracle/ref_boxes.py against that file instead of against a restatement.

    python -m oracle.refrun.functions          # build container only (/root/reference)

Functions: get_medoid (nuscenes:116-119, kitti:177-180, waymo:120-122), get_detection_name,
get_shape_prior, push_centroid, lane_yaws_distances_and_coords (nuscenes:277-302), circle_nms
(:309-332), get_yaws_from_lane_coords (waymo:374-388), get_depth_bbox + save_pred (kitti:855-885;
open3d stubbed by oracle/obb_oracle.open3d_obb).
"""
from __future__ import annotations

import ast
import json
import os
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF_SRC = "/root/reference/src"
OUT = os.path.join(ROOT, "tests", "golden", "pass2_functions.json")


def load_functions(ds, names, extra=None):
    """FunctionDefs `names` (+ module-level constant Assigns they read) of src/<ds>/2d_to_3d.py, exec'd."""
    import scipy
    import scipy.spatial.distance
    import torch
    from scipy.spatial.transform import Rotation
    path = os.path.join(REF_SRC, ds, "2d_to_3d.py")
    tree = ast.parse(open(path).read(), filename=path)
    keep = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name in names) or
            (isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Name) and n.targets[0].id in ("KITTI_CLASS_MAPS", "ATTRIBUTE_NAMES"))]
    found = {n.name for n in keep if isinstance(n, ast.FunctionDef)}
    assert found == set(names), (ds, set(names) - found)
    ns = {"np": np, "torch": torch, "scipy": scipy, "R": Rotation, "time": time, "os": os,
          "timer": {"closest lane": 0}, "__name__": f"ref_{ds}"}
    ns.update(extra or {})
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


def _j(o):
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    if isinstance(o, (list, tuple)):
        return [_j(v) for v in o]
    if isinstance(o, dict):
        return {k: _j(v) for k, v in o.items()}
    return o


def main():
    import torch
    torch.set_num_threads(1)
    from oracle import obb_oracle, pyquat
    rng = np.random.default_rng(20261018)
    names = ["get_medoid", "get_detection_name", "get_shape_prior", "push_centroid"]
    nusc = load_functions("nuscenes", names + ["lane_yaws_distances_and_coords", "circle_nms"])

    class _PC:
        points = None

        def get_oriented_bounding_box(self):
            c, e, Rm = obb_oracle.open3d_obb(np.asarray(self.points, np.float64))
            return types.SimpleNamespace(center=np.asarray(c), extent=np.asarray(e), R=Rm)
    o3d = types.SimpleNamespace(geometry=types.SimpleNamespace(PointCloud=_PC),
                                utility=types.SimpleNamespace(Vector3dVector=lambda a: np.asarray(a, np.float64)))
    kitti = load_functions("kitti", names + ["circle_nms", "get_depth_bbox", "save_pred"], {"o3d": o3d})
    waymo = load_functions("waymo", names + ["circle_nms", "get_yaws_from_lane_coords"])
    fix = {"provenance": "FunctionDefs of /root/reference/src/<ds>/2d_to_3d.py exec'd unmodified (oracle/refrun/functions.py); "
                         f"torch {torch.__version__} cpu cap={torch.backends.cpu.get_cpu_capability()} threads=1, numpy {np.__version__}"}
    pri = json.load(open(os.path.join(REF_SRC, "nuscenes", "cfg", "shape_priors_chatgpt.json")))
    pri_old = json.load(open(os.path.join(REF_SRC, "nuscenes", "cfg", "shape_priors.json")))

    # ---- names and priors
    labels = list(pri) + ["trafficcone", "constructionvehicle", "human"]
    fix["get_detection_name"] = {ds: {l: ns["get_detection_name"](l) for l in labels}
                                 for ds, ns in (("nuscenes", nusc), ("kitti", kitti), ("waymo", waymo))}
    fix["get_shape_prior"] = {
        "chatgpt": {ds: {l: ns["get_shape_prior"](pri, l) for l in pri} for ds, ns in (("nuscenes", nusc), ("kitti", kitti), ("waymo", waymo))},
        "waymo_types": {l: waymo["get_shape_prior"](pri, l) for l in ("vehicle", "pedestrian", "cyclist")},
        "not_chatgpt": {l: nusc["get_shape_prior"](pri_old, l, chatgpt=False)
                        for l in ("car", "bicycle", "bus", "truck", "pedestrian", "traffic_cone", "construction_vehicle",
                                  "motorcycle", "trailer", "child", "stroller", "barrier")},
    }

    # ---- push_centroid: all quadrants, lane yaws around the circle, ego_frame on/off (waymo)
    cases = []
    for k in range(48):
        yaw = np.float32(rng.uniform(-np.pi, np.pi))                 # the reference's lane yaw is a numpy float32 (nuscenes:756,789)
        align = np.eye(3)
        align[0:2, 0:2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        q = pyquat.Quaternion(matrix=align)
        av = rng.uniform(300, 2000, 3)
        cen = av + np.array([rng.uniform(2, 60) * rng.choice([-1, 1]), rng.uniform(2, 60) * rng.choice([-1, 1]), rng.uniform(-1, 2)])
        if k == 0:
            cen = av + np.array([10.0, 0.0, 0.5])                    # on the x axis
        if k == 1:
            cen = av + np.array([0.0, -7.0, 0.5])                    # on the y axis: arctan(inf)
        ext = pri[list(pri)[k % len(pri)]]
        with np.errstate(all="ignore"):
            a = nusc["push_centroid"](np.float32(cen)[None], ext, q, {"translation": list(av)})
            ego = np.float32(cen - av)
            b = waymo["push_centroid"](ego, ext, q, ego_frame=True)
            c = kitti["push_centroid"](np.float32(cen)[None], ext, q, {"translation": list(av)})
        assert np.array_equal(a, c, equal_nan=True)
        cases.append({"lane_yaw": float(yaw), "quat_wxyz": list(q), "av_translation": list(av), "centroid_f32": np.float32(cen).tolist(),
                      "centroid_ego_f32": ego.tolist(),
                      "extents": ext, "pushed_global": a, "pushed_ego_frame": b})
    fix["push_centroid"] = cases

    # ---- closest lane: centroids x lane points (incl. exact ties: duplicated lane points)
    lane = np.concatenate([np.stack([np.cumsum(np.full(400, 0.5)) + rng.uniform(300, 900), np.linspace(0, 60, 400) + rng.uniform(500, 900),
                                     np.full(400, rng.uniform(-3, 3))], 1) for _ in range(5)])
    lane = np.concatenate([lane, lane[100:110]])                      # duplicates -> argmin must take the first
    cents = np.float32(np.concatenate([lane[rng.integers(0, len(lane), 40), :3] + rng.normal(0, 6, (40, 3)), lane[105:107, :3]]))
    yaws, dists, coords = nusc["lane_yaws_distances_and_coords"](torch.from_numpy(cents), [tuple(p) for p in lane.tolist()])
    fix["lane_yaws_distances_and_coords"] = {"centroids_f32": cents, "lane_pts": lane, "yaws": yaws, "distances": dists, "coords": coords}

    # ---- circle NMS: clustered boxes, equal scores, per-class thresholds
    thr = {"barrier": 1, "traffic_cone": 0.175, "bicycle": 0.85, "motorcycle": 0.85, "pedestrian": 0.175, "car": 4, "bus": 10,
           "construction_vehicle": 12, "trailer": 10, "truck": 12}
    nms = []
    for k in range(6):
        n = [0, 1, 12, 40, 40, 80][k]
        centres = rng.uniform(0, 30, (max(n // 4, 1), 2))
        xy = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 0.8, (n, 2))
        sc = np.round(rng.uniform(0.1, 1.0, n), 2 if k == 4 else 6)   # k == 4: many equal scores
        lab = [list(thr)[j] for j in rng.integers(0, len(thr), n)]
        dets = np.concatenate([xy, sc[:, None]], 1) if n else np.zeros((0, 3))
        keep = nusc["circle_nms"](dets, lab, thr) if n else []
        assert keep == (waymo["circle_nms"](dets, lab, thr) if n else []) == (kitti["circle_nms"](dets, lab, thr) if n else [])
        nms.append({"dets": dets, "labels": lab, "keep": [int(i) for i in keep]})
    fix["circle_nms"] = {"threshs_by_label": thr, "cases": nms}

    # ---- Waymo lane yaws by finite differences
    poly = [types.SimpleNamespace(x=float(x), y=float(y), z=0.0) for x, y in np.cumsum(rng.normal(0.4, 0.3, (30, 2)), 0) + 500]
    fix["get_yaws_from_lane_coords"] = {"polyline_xy": [[p.x, p.y] for p in poly], "out": waymo["get_yaws_from_lane_coords"](poly),
                                        "single": waymo["get_yaws_from_lane_coords"](poly[:1])}

    # ---- get_medoid on point sets in the three frames' coordinate ranges (global ~1e3 m, sensor ~1e1 m)
    med = []
    for k, (m, centre) in enumerate([(1, (1200, 900, 1)), (2, (1200, 900, 1)), (7, (0, 0, 0)), (25, (600, 1500, 0)), (26, (600, 1500, 0)),
                                     (33, (5, -20, 1)), (96, (1800, 400, 2)), (257, (-8, 30, 0)), (700, (1024.5, 512.25, 0.5)),
                                     (1500, (310, 1990, 1)), (2077, (15, 3, -1))]):
        pts = np.float32(np.asarray(centre) + rng.normal(0, 1.5, (m, 3)) * (1, 1, 0.4))
        if k == 4:
            pts[7] = pts[3]                                           # duplicate point: tie between two columns
        t = torch.from_numpy(np.ascontiguousarray(pts.T))
        out = [int(ns["get_medoid"](t)) for ns in (nusc, kitti, waymo)]
        assert out[0] == out[1] == out[2]
        med.append({"points_f32": pts, "medoid": out[0]})
    fix["get_medoid"] = med

    # ---- KITTI: get_depth_bbox (open3d stubbed) + yaw + save_pred lines
    from scipy.spatial.transform import Rotation
    from oracle.refrun import stubs
    stubs.install_scipy_1_11_from_matrix()
    boxes = []
    with tempfile.TemporaryDirectory() as td:
        for k in range(12):
            m = int(rng.integers(8, 400))
            ext = np.array([[4.5, 1.5, 1.8], [1.7, 0.6, 0.6], [12.0, 3.5, 2.5], [0.8, 1.9, 0.7]][k % 4])
            Rm = Rotation.from_euler("zyx", rng.uniform(-np.pi, np.pi, 3) * (1, 0.05, 0.05)).as_matrix()
            pts = np.float32((rng.uniform(-0.5, 0.5, (m, 3)) * ext) @ Rm.T + rng.uniform(-30, 30, 3))
            center, wlh, Rb = kitti["get_depth_bbox"](pts)
            yaw = Rotation.from_matrix(Rb).as_euler("zyx")[0]
            p = os.path.join(td, f"{k}.txt")
            kitti["save_pred"](p, "Car", [0, 0, 0, 0], [1.4, 1.8, 4.5], center, yaw, 0.5 if k % 2 else None)
            boxes.append({"points_f32": pts, "center": center, "wlh": wlh, "R": Rb, "yaw": float(yaw), "line": open(p).read()})
    fix["get_depth_bbox"] = boxes

    with open(OUT, "w") as f:
        json.dump(_j(fix), f)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
