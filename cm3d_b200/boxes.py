"""Pass 2 of the lifting scripts: centroid -> 3D box (class name, shape prior, lane yaw,
push-back, circle NMS).  Host side; only the centroid x lane-point nearest neighbour runs on
the GPU (csrc/boxes.cu through the C ABI).

Mirrors, function for function:
  get_detection_name / get_shape_prior   src/nuscenes/2d_to_3d.py:122-161, kitti:183-230, waymo:125-172
  push_centroid                          src/nuscenes/2d_to_3d.py:164-198, waymo:175-211
  lane_yaws_distances_and_coords         src/nuscenes/2d_to_3d.py:277-302
  circle_nms                             src/nuscenes/2d_to_3d.py:309-332 (CenterPoint's circle NMS)
  box assembly / NMS over the results    src/nuscenes/2d_to_3d.py:733-924

scipy's Rotation is called exactly where the reference calls it (including the quirk that a
pyquaternion (w,x,y,z) list is handed to `from_quat`, which reads it as (x,y,z,w)).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import numpy as np

from .quat import Quaternion, quats_from_matrices

ATTRIBUTE_NAMES = {     # src/nuscenes/2d_to_3d.py:70-81
    "barrier": "", "traffic_cone": "", "bicycle": "cycle.without_rider", "motorcycle": "cycle.without_rider",
    "pedestrian": "pedestrian.standing", "car": "vehicle.stopped", "bus": "vehicle.stopped",
    "construction_vehicle": "vehicle.stopped", "trailer": "vehicle.stopped", "truck": "vehicle.stopped",
}
THRESHS_BY_LABEL = {    # src/nuscenes/2d_to_3d.py:850-861 ("borrowed from centerpoint"), squared metres
    "barrier": 1, "traffic_cone": 0.175, "bicycle": 0.85, "motorcycle": 0.85, "pedestrian": 0.175,
    "car": 4, "bus": 10, "construction_vehicle": 12, "trailer": 10, "truck": 12,
}
KITTI_CLASS_MAPS = {    # src/kitti/2d_to_3d.py:105-116
    "car": "Car", "pedestrian": "Pedestrian", "truck": "Truck", "bus": "Tram", "traffic_cone": "Misc",
    "construction_vehicle": "Misc", "bicycle": "Cyclist", "motorcycle": "Cyclist", "trailer": "Misc", "barrier": "Misc",
}
NUSC_TO_WAYMO = {       # src/waymo/cfg/prompt_cfg.py:286-296
    "car": "vehicle", "truck": "vehicle", "bus": "vehicle", "bicycle": "cyclist", "pedestrian": "pedestrian",
    "trailer": "vehicle", "barrier": "", "construction_vehicle": "vehicle", "traffic_cone": "", "motorcycle": "vehicle",
}
VEHICLE_NAMES = ["car", "truck", "bus", "construction_vehicle", "trailer", "barrier"]   # nuscenes:763
_NON_CHATGPT = {
    "car": "vehicle.car", "bicycle": "vehicle.bicycle", "bus": "vehicle.bus.rigid", "truck": "vehicle.truck",
    "pedestrian": "human.pedestrian.adult", "traffic_cone": "movable_object.trafficcone",
    "construction_vehicle": "vehicle.construction", "motorcycle": "vehicle.motorcycle", "trailer": "vehicle.trailer",
    "child": "human.pedestrian.child", "stroller": "human.pedestrian.adult",
}


def get_detection_name(name: str, class_map: Optional[Dict[str, str]] = None) -> str:
    """Detic label -> detection class (nuscenes:122-132); KITTI maps on through KITTI_CLASS_MAPS (kitti:195)."""
    detection_name = {"trafficcone": "traffic_cone", "constructionvehicle": "construction_vehicle",
                      "human": "pedestrian"}.get(name, name)
    return class_map[detection_name] if class_map is not None else detection_name


def get_shape_prior(shape_priors: dict, name: str, chatgpt: bool = True, waymo: bool = False):
    """[w, l, h] of a class (nuscenes:134-161); Waymo's variant also accepts its own type names (waymo:162-172)."""
    if not chatgpt:
        key = _NON_CHATGPT.get(name)
        return None if key is None else shape_priors[key]
    if waymo:
        name = {"vehicle": "car", "cyclist": "bicycle"}.get(name, name)
    return shape_priors[name]


def push_centroid(centroid, extents, rot_quaternion, poserecord=None, ego_frame=False):
    """Move a surface medoid back along the viewing ray by the box's half-depth (nuscenes:164-198,
    waymo:175-211).  `rot_quaternion` iterates (w,x,y,z); scipy reads the list as (x,y,z,w)."""
    from scipy.spatial.transform import Rotation as R
    centroid = np.squeeze(centroid)
    ego_centroid = centroid if ego_frame else centroid - poserecord["translation"]
    l, w = extents[0], extents[1]
    angle = R.from_quat(list(rot_quaternion)).as_euler("xyz", degrees=False)
    theta = -angle[0]
    if np.isnan(theta):
        theta = 0.5 * np.pi
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = np.arctan(np.abs(ego_centroid[1]) / np.abs(ego_centroid[0]))
        if ego_centroid[0] < 0:
            alpha = -np.pi + alpha if ego_centroid[1] < 0 else np.pi - alpha
        elif ego_centroid[1] < 0:
            alpha = -alpha
        offset = np.min([np.abs(w / (2 * np.sin(theta - alpha))), np.abs(l / (2 * np.cos(theta - alpha)))])
    x_dash = centroid[0] + offset * np.cos(alpha)
    y_dash = centroid[1] + offset * np.sin(alpha)
    return np.array([x_dash, y_dash, centroid[2]])


def lane_yaws_distances_and_coords(all_centroids, all_lane_pts, device="cuda:0"):
    """Closest lane point per centroid: (yaws, distances, coords), nuscenes:277-302.

    Like the reference, both inputs go through float32 (`torch.Tensor(...)`) and the distances
    are binary64 (`scipy.spatial.distance.cdist`); argmin/min run on the GPU (cm3d_nearest_lane)
    without materialising the matrix."""
    import torch
    from . import _native as N
    lanes32 = np.ascontiguousarray(np.asarray(all_lane_pts, dtype=np.float32).reshape(-1, 3))
    cent32 = np.ascontiguousarray(np.asarray(all_centroids, dtype=np.float32).reshape(-1, 3))
    n, m = cent32.shape[0], lanes32.shape[0]
    if m == 0:
        raise ValueError("attempt to get argmin of an empty sequence")      # numpy's error in the reference
    dev = torch.device(device)
    with torch.cuda.device(dev):                     # the C ABI launches on the current CUDA device
        cxy = torch.from_numpy(cent32[:, :2].astype(np.float64)).to(dev)
        lxy = torch.from_numpy(lanes32[:, :2].astype(np.float64)).to(dev)
        idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        dist = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
        N.call("cm3d_nearest_lane", ctypes.c_void_p(cxy.data_ptr()), n, ctypes.c_void_p(lxy.data_ptr()), m,
               ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(dist.data_ptr()),
               ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    idx = idx[:n].cpu().numpy().astype(np.int64)
    return lanes32[idx, 2], dist[:n].cpu().numpy(), lanes32[idx, :2]


def circle_nms(dets: np.ndarray, det_labels: Sequence, threshs_by_label: dict) -> List[int]:
    """Greedy centre-distance NMS, class aware (nuscenes:309-332, CenterPoint's circle NMS).  dets rows: x, y, score.

    The reference walks all ordered pairs in Python; a box can only suppress boxes of its own class, so the
    classes are independent greedy passes: per class, in descending score order (the reference's
    `argsort()[::-1]`, ties included), a surviving box removes every later one within the class radius with one
    array comparison.  Returns the kept indices in descending score order, like the reference."""
    dets = np.asarray(dets)
    n = dets.shape[0]
    if n == 0:
        return []
    order = dets[:, 2].argsort()[::-1]
    rank = np.empty(n, dtype=np.int64)
    rank[order] = np.arange(n)
    labels = np.asarray(det_labels, dtype=object)
    alive = np.ones(n, dtype=bool)
    for label in dict.fromkeys(det_labels):
        idx = order[labels[order] == label]              # this class, best score first
        x, y, thr = dets[idx, 0], dets[idx, 1], threshs_by_label[label]
        live = np.ones(idx.size, dtype=bool)
        if idx.size <= 1024:                             # all pairs of the class at once, then the greedy pass over rows
            far = ~(((x[:, None] - x[None, :]) ** 2 + (y[:, None] - y[None, :]) ** 2) <= thr)
            for a in range(idx.size - 1):
                if live[a]:
                    live[a + 1:] &= far[a, a + 1:]
        else:
            for a in range(idx.size - 1):
                if live[a]:
                    d = (x[a] - x[a + 1:]) ** 2 + (y[a] - y[a + 1:]) ** 2
                    live[a + 1:] &= ~(d <= thr)
        alive[idx[~live]] = False
    keep = np.flatnonzero(alive)
    return [int(k) for k in keep[np.argsort(rank[keep])]]


def lane_align_matrix(lane_yaw) -> np.ndarray:
    """Rotation about z by the lane yaw, built like nuscenes:788-789 (the yaw keeps its float32 type)."""
    align_mat = np.eye(3)
    align_mat[0:2, 0:2] = [[np.cos(lane_yaw), -np.sin(lane_yaw)], [np.sin(lane_yaw), np.cos(lane_yaw)]]
    return align_mat


def nuscenes_box(sample_token: str, label: str, score, centroid: np.ndarray, lane_yaw, shape_priors: dict,
                 lidar_poserecord: dict, attribute_names: dict = ATTRIBUTE_NAMES) -> dict:
    """One entry of predictions["results"][sample_token] (nuscenes:745-817)."""
    detection_name = get_detection_name(label)
    centroid = np.squeeze(np.asarray(centroid))
    extents = get_shape_prior(shape_priors, detection_name)
    if detection_name in VEHICLE_NAMES:
        align_mat = lane_align_matrix(lane_yaw)
        pushed_centroid = push_centroid(centroid, extents, Quaternion(matrix=align_mat), lidar_poserecord)
    else:
        align_mat = np.eye(3)
        pushed_centroid = centroid
    rot_quaternion = Quaternion(matrix=align_mat)
    return {
        "sample_token": sample_token,
        "translation": [float(i) for i in pushed_centroid],
        "size": list(extents),
        "rotation": [float(v) for v in rot_quaternion],
        "velocity": [0, 0],
        "detection_name": detection_name,
        "detection_score": score,
        "attribute_name": attribute_names[detection_name],
    }


def lane_align_matrices(lane_yaws, where=None) -> np.ndarray:
    """`lane_align_matrix` for K yaws at once (rows where `where` is False stay the identity)."""
    yaw = np.asarray(lane_yaws)
    k = yaw.shape[0]
    where = np.ones(k, dtype=bool) if where is None else where
    cs, sn = np.cos(yaw), np.sin(yaw)                       # float32 yaws give float32 values, like the scalars of :788-789
    mats = np.tile(np.eye(3), (k, 1, 1))
    mats[where, 0, 0] = cs[where]
    mats[where, 0, 1] = -sn[where]
    mats[where, 1, 0] = sn[where]
    mats[where, 1, 1] = cs[where]
    return mats


def push_centroids(centroids: np.ndarray, ego_centroids: np.ndarray, extents_lw: np.ndarray, quats_wxyz: np.ndarray) -> np.ndarray:
    """`push_centroid` for K boxes at once: the same scipy / numpy element operations on arrays (bit-identical
    per box).  `extents_lw[:, 0]` is what the reference calls l (= extents[0]), `[:, 1]` its w."""
    from scipy.spatial.transform import Rotation as R
    c = np.asarray(centroids, dtype=np.float64)
    ego = np.asarray(ego_centroids, dtype=np.float64)
    l, w = extents_lw[:, 0], extents_lw[:, 1]
    theta = -R.from_quat(quats_wxyz).as_euler("xyz", degrees=False)[:, 0]      # (w,x,y,z) read as (x,y,z,w), like the reference
    theta = np.where(np.isnan(theta), 0.5 * np.pi, theta)
    out = c.copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = np.arctan(np.abs(ego[:, 1]) / np.abs(ego[:, 0]))
        alpha = np.where(ego[:, 0] < 0, np.where(ego[:, 1] < 0, -np.pi + alpha, np.pi - alpha),
                         np.where(ego[:, 1] < 0, -alpha, alpha))
        offset = np.minimum(np.abs(w / (2 * np.sin(theta - alpha))), np.abs(l / (2 * np.cos(theta - alpha))))
        out[:, 0] = c[:, 0] + offset * np.cos(alpha)
        out[:, 1] = c[:, 1] + offset * np.sin(alpha)
    return out


def nuscenes_boxes(sample_tokens: Sequence[str], labels: Sequence[str], scores: Sequence, centroids: np.ndarray, lane_yaws,
                   shape_priors: dict, pose_translations: np.ndarray, attribute_names: dict = ATTRIBUTE_NAMES) -> List[dict]:
    """`nuscenes_box` over all K boxes of a scene at once (the reference's per-box loop nuscenes:745-817 costs
    ~0.2 ms of interpreter time per box): the same numpy / scipy element operations on arrays, so the
    results are the per-box ones bit for bit (tests/test_host_logic.py).  `centroids` (K,3) float32,
    `lane_yaws` (K,) float32, `pose_translations` (K,3) the lidar ego_pose translation of each box's sample."""
    k = len(labels)
    if k == 0:
        return []
    names = [get_detection_name(l) for l in labels]
    extents = [get_shape_prior(shape_priors, n) for n in names]
    veh = np.fromiter((n in VEHICLE_NAMES for n in names), dtype=bool, count=k)
    c32 = np.asarray(centroids, dtype=np.float32).reshape(k, 3)
    mats = lane_align_matrices(lane_yaws, veh)
    quats = quats_from_matrices(mats)
    trans = c32.astype(np.float64)
    if veh.any():
        c = trans[veh]
        ego = c - np.asarray(pose_translations, dtype=np.float64).reshape(k, 3)[veh]
        ext = np.asarray([extents[i][:2] for i in np.flatnonzero(veh)], dtype=np.float64)
        trans[veh] = push_centroids(c, ego, ext, quats[veh])
    trans, quats = trans.tolist(), quats.tolist()
    return [{"sample_token": sample_tokens[i], "translation": trans[i], "size": list(extents[i]), "rotation": quats[i],
             "velocity": [0, 0], "detection_name": names[i], "detection_score": scores[i],
             "attribute_name": attribute_names[names[i]]} for i in range(k)]


def nms_predictions(predictions: dict, threshs_by_label: dict = THRESHS_BY_LABEL) -> dict:
    """Per-sample circle NMS over the assembled boxes (nuscenes:833-924).  Samples without any
    prediction keep an empty list, as in the reference."""
    final = {"meta": dict(predictions["meta"]), "results": {}}
    for sample, boxes in predictions["results"].items():
        final["results"][sample] = []
        if len(boxes) == 0:
            continue
        dets = np.array([np.array([b["translation"][0], b["translation"][1], b["detection_score"]]) for b in boxes])
        det_labels = [b["detection_name"] for b in boxes]
        keep = set(int(k) for k in circle_nms(dets, det_labels, threshs_by_label))
        for c, b in enumerate(boxes):
            if c in keep:
                final["results"][sample].append({
                    "sample_token": sample, "translation": b["translation"], "size": b["size"],
                    "rotation": b["rotation"], "velocity": [0, 0], "detection_name": b["detection_name"],
                    "detection_score": b["detection_score"], "attribute_name": b["attribute_name"]})
    return final
