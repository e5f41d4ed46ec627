"""ORACLE (test infrastructure, not product code): CPU restatement of pass 2 of the reference's
lifting scripts - closest lane, box assembly, push-back, circle NMS - line by line, with the
same library calls (scipy cdist / Rotation, numpy).

Follows src/nuscenes/2d_to_3d.py:164-198 (push_centroid), :277-302
(lane_yaws_distances_and_coords), :309-332 (circle_nms), :733-825 (pass 2), :844-924 (NMS);
src/kitti/2d_to_3d.py:1524-1536 (label line); src/waymo/2d_to_3d.py:785-870 (pass 2).
pyquaternion is absent here, so `Quaternion(matrix=M)` is restated with scipy's Rotation
(w,x,y,z order, sign fixed to w >= 0 like the trace method gives for these yaw-only matrices):
PARITY UNPINNED against pyquaternion itself.  Only tests/ and bench.py may import this module.
"""
import numpy as np
import scipy.spatial.distance
from scipy.spatial.transform import Rotation as R

ATTRIBUTE_NAMES = {
    "barrier": "", "traffic_cone": "", "bicycle": "cycle.without_rider", "motorcycle": "cycle.without_rider",
    "pedestrian": "pedestrian.standing", "car": "vehicle.stopped", "bus": "vehicle.stopped",
    "construction_vehicle": "vehicle.stopped", "trailer": "vehicle.stopped", "truck": "vehicle.stopped",
}
THRESHS = {"barrier": 1, "traffic_cone": 0.175, "bicycle": 0.85, "motorcycle": 0.85, "pedestrian": 0.175,
           "car": 4, "bus": 10, "construction_vehicle": 12, "trailer": 10, "truck": 12}


def quat_wxyz_from_matrix(m):
    x, y, z, w = R.from_matrix(np.asarray(m, np.float64)).as_quat()
    q = np.array([w, x, y, z])
    return -q if w < 0 else q


def get_detection_name(name):
    if name not in ["trafficcone", "constructionvehicle", "human"]:
        return name
    return {"trafficcone": "traffic_cone", "constructionvehicle": "construction_vehicle", "human": "pedestrian"}[name]


def push_centroid(centroid, extents, rot_quaternion_wxyz, av_translation=None, ego_frame=False):
    centroid = np.squeeze(centroid)
    ego_centroid = centroid if ego_frame else centroid - av_translation
    l = extents[0]
    w = extents[1]
    angle = R.from_quat(list(rot_quaternion_wxyz)).as_euler('xyz', degrees=False)   # (w,x,y,z) read as (x,y,z,w)
    theta = -angle[0]
    if np.isnan(theta):
        theta = 0.5 * np.pi
    alpha = np.arctan(np.abs(ego_centroid[1]) / np.abs(ego_centroid[0]))
    if ego_centroid[0] < 0:
        if ego_centroid[1] < 0:
            alpha = -np.pi + alpha
        else:
            alpha = np.pi - alpha
    else:
        if ego_centroid[1] < 0:
            alpha = -alpha
    offset = np.min([np.abs(w / (2 * np.sin(theta - alpha))), np.abs(l / (2 * np.cos(theta - alpha)))])
    return np.array([centroid[0] + offset * np.cos(alpha), centroid[1] + offset * np.sin(alpha), centroid[2]])


def lane_yaws_distances_and_coords(all_centroids, all_lane_pts):
    import torch
    all_lane_pts = torch.Tensor(np.asarray(all_lane_pts)).to(device='cpu')
    all_centroids = torch.Tensor(np.asarray(all_centroids)).to(device='cpu')
    DistMat = scipy.spatial.distance.cdist(all_centroids[:, :2], all_lane_pts[:, :2])
    min_lane_indices = np.argmin(DistMat, axis=1)
    distances = np.min(DistMat, axis=1)
    all_lane_pts = all_lane_pts.numpy()
    min_lanes = all_lane_pts[min_lane_indices]
    return min_lanes[:, 2], distances, min_lanes[:, :2], min_lane_indices


def circle_nms(dets, det_labels, threshs_by_label):
    x1 = dets[:, 0]
    y1 = dets[:, 1]
    scores = dets[:, 2]
    order = scores.argsort()[::-1].astype(np.int32)
    ndets = dets.shape[0]
    suppressed = np.zeros((ndets), dtype=np.int32)
    keep = []
    for _i in range(ndets):
        i = order[_i]
        if suppressed[i] == 1:
            continue
        keep.append(i)
        for _j in range(_i + 1, ndets):
            j = order[_j]
            if suppressed[j] == 1:
                continue
            dist = (x1[i] - x1[j]) ** 2 + (y1[i] - y1[j]) ** 2
            if dist <= threshs_by_label[det_labels[j]] and det_labels[j] == det_labels[i]:
                suppressed[j] = 1
    return keep


def nuscenes_scene(samples, datas, lidar_translations, centroid_ids, centroids, lane_pts, shape_priors):
    """Pass 2 of one scene: sample token -> box dicts (before NMS)."""
    results = {tok: [] for tok in samples}
    if len(centroid_ids) == 0:
        return results
    yaw_list, _, _, _ = lane_yaws_distances_and_coords(centroids, lane_pts)
    id_offset = -1
    for tok, data, av_t in zip(samples, datas, lidar_translations):
        for label, score, c in zip(data["labels"], data["detection_scores"], data["cam_nums"]):
            id_offset += 1
            if id_offset not in centroid_ids:
                continue
            k = centroid_ids.index(id_offset)
            detection_name = get_detection_name(label)
            centroid = np.squeeze(np.array(centroids[k]))
            lane_yaw = yaw_list[k]
            extents = shape_priors[detection_name]
            if detection_name in ["car", "truck", "bus", "construction_vehicle", "trailer", "barrier"]:
                align_mat = np.eye(3)
                align_mat[0:2, 0:2] = [[np.cos(lane_yaw), -np.sin(lane_yaw)], [np.sin(lane_yaw), np.cos(lane_yaw)]]
                pushed = push_centroid(centroid, extents, quat_wxyz_from_matrix(align_mat), np.asarray(av_t))
            else:
                align_mat = np.eye(3)
                pushed = centroid
            results[tok].append({
                "sample_token": tok, "translation": [float(i) for i in pushed], "size": list(extents),
                "rotation": [float(v) for v in quat_wxyz_from_matrix(align_mat)], "velocity": [0, 0],
                "detection_name": detection_name, "detection_score": score,
                "attribute_name": ATTRIBUTE_NAMES[detection_name]})
    return results


def nuscenes_nms(results):
    final = {}
    for sample, boxes in results.items():
        final[sample] = []
        if not boxes:
            continue
        dets = np.array([np.array([b["translation"][0], b["translation"][1], b["detection_score"]]) for b in boxes])
        labels = [b["detection_name"] for b in boxes]
        keep = list(circle_nms(dets, labels, THRESHS))
        final[sample] = [b for c, b in enumerate(boxes) if c in keep]
    return final


# ------------------------------------------------------------------ Waymo pass 2 (waymo:684-699, 803-858)
NUSC_TO_WAYMO = {"car": "vehicle", "truck": "vehicle", "bus": "vehicle", "bicycle": "cyclist", "pedestrian": "pedestrian",
                 "trailer": "vehicle", "barrier": "", "construction_vehicle": "vehicle", "traffic_cone": "",
                 "motorcycle": "vehicle"}
WAYMO_TYPES = {"vehicle": 1, "pedestrian": 2, "cyclist": 4}
WAYMO_THRESHS = {0: 1, 3: 0.175, 4: 0.85, 2: 0.175, 1: 4}


def waymo_centroid_to_global(centroid_xyz, pose_transform):
    """fp64 restatement of the reference's fp32 rotate + translate (tolerance 1e-3 m in the tests)."""
    tm = np.array(pose_transform, np.float32).reshape(4, 4).astype(np.float64)
    rot = R.from_matrix(tm[:3, :3]).as_matrix()          # quaternion round trip = projection onto SO(3)
    return rot @ np.asarray(centroid_xyz, np.float64).reshape(3) + tm[:3, 3]


def waymo_object(context_name, timestamp, pose_transform, label, score, centroid_global, global_lane_yaw, shape_priors):
    detection_name = get_detection_name(label)
    transform_matrix = np.array(pose_transform, np.float32).reshape(4, 4)
    transform_matrix = np.linalg.inv(transform_matrix)
    centroid_pc = np.hstack([np.squeeze(np.array(centroid_global)), [1]])
    centroid = np.dot(transform_matrix, centroid_pc)[:3]
    extents = shape_priors[{"vehicle": "car", "cyclist": "bicycle"}.get(detection_name, detection_name)]
    if detection_name in ["car", "truck", "bus", "construction_vehicle", "trailer", "barrier"]:
        global_align_mat = np.eye(3)
        global_align_mat[0:2, 0:2] = [[np.cos(global_lane_yaw), -np.sin(global_lane_yaw)],
                                      [np.sin(global_lane_yaw), np.cos(global_lane_yaw)]]
        align_mat = np.dot(transform_matrix[:3, :3], global_align_mat)
        pushed = push_centroid(centroid, extents, quat_wxyz_from_matrix(global_align_mat), ego_frame=True)
        heading = R.from_matrix(align_mat).as_euler('xyz', degrees=False)[2]
    else:
        pushed = centroid
        heading = R.from_matrix(np.eye(3)).as_euler('xyz', degrees=False)[2]
    name = NUSC_TO_WAYMO[detection_name]
    if name not in WAYMO_TYPES:
        raise ValueError
    return {"context_name": context_name, "frame_timestamp_micros": timestamp, "center_x": float(pushed[0]),
            "center_y": float(pushed[1]), "center_z": float(pushed[2]), "length": float(extents[1]),
            "width": float(extents[0]), "height": float(extents[2]), "heading": float(heading),
            "score": float(np.float32(score)), "type": WAYMO_TYPES[name]}


def waymo_nms(objects):
    by_ts = {}
    for o in objects:
        by_ts.setdefault(o["frame_timestamp_micros"], []).append(o)
    final = []
    for ts, objs in by_ts.items():
        dets = np.array([np.array([o["center_x"], o["center_y"], o["score"]]) for o in objs])
        keep = list(circle_nms(dets, [o["type"] for o in objs], WAYMO_THRESHS))
        final.extend(o for c, o in enumerate(objs) if c in keep)
    return final
