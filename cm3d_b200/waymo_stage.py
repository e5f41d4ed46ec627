"""Waymo lifting stage: the `__main__` of the reference's src/waymo/2d_to_3d.py as a function, the
per-frame / per-mask body replaced by the CUDA path.

Follows src/waymo/2d_to_3d.py:395-1305.  Pass 1 per frame (:445-702): masks + data (missing files
skip the frame, :450-455), lanes with finite-difference yaws from frame 0's map features
(:459-468 -> :374-388), TOP-LiDAR first-return points + a ones row in the vehicle frame (:472-481);
per mask: camera `c + 1` (:513-518), extrinsic . inv(axes) -> scipy quaternion -> pyquaternion
matrix (:557-575), intrinsics x 1024/1920 in fp64 (:586-593), membership, medoid (:649-653),
centroid to the global frame through `frame.pose` (:684-699).  Closest lane per centroid (:751).
Pass 2 (:785-870): back to the vehicle frame with inv(pose), shape prior, lane-aligned heading,
push-back, `metrics_pb2.Object`; per-timestamp circle NMS (:1108-1275); serialised Objects file
(:1300-1305).

`scenes` is an iterable of (scene_name, frames); frames are `dataset_pb2.Frame`s or duck-typed
equivalents.  `points_fn(frame)` returns the (N,3) vehicle-frame points; the default parses the
range images with waymo_open_dataset exactly like the reference.
"""
from __future__ import annotations

import json
import os
import time
from types import SimpleNamespace
from typing import Callable, Iterable, List, Optional

import numpy as np

from . import boxes as B
from . import waymo_proto as WP
from .frames import CamSpec, FrameSpec, FOURTH_ONES, op_R, op_T
from .nuscenes_stage import load_frame_masks, new_timer
from .quat import Quaternion

DEFAULTS = dict(
    INPUT_PATH="../../data/waymo/training/", OUTPUT_DIR="../../outputs/waymo/",
    INPUT_DIR="../../mask_outputs/waymo-detic/", ATTRIBUTE_NAMES=B.ATTRIBUTE_NAMES, DEVICE="cuda:0",
    CAM_LIST=["FRONT", "FRONT_LEFT", "FRONT_RIGHT", "SIDE_LEFT", "SIDE_RIGHT"],
    min_dist=2.3, floor_thresh=-0.6, ratio=1024 / 1920, scene_slice=(680, 710),
    shape_priors_path="cfg/shape_priors_chatgpt.json", output_path="../../outputs/waymo/pred_0307_detic_train_680_710.bin",
    batch_frames=32,
)
WAYMO_THRESHS = {WP.TYPE_UNKNOWN: 1, WP.TYPE_SIGN: 0.175, WP.TYPE_CYCLIST: 0.85, WP.TYPE_PEDESTRIAN: 0.175,
                 WP.TYPE_VEHICLE: 4}          # waymo:1119-1125


def make_cfg(**overrides) -> SimpleNamespace:
    d = dict(DEFAULTS)
    d.update(overrides)
    return SimpleNamespace(**d)


def get_yaws_from_lane_coords(lane_list) -> np.ndarray:
    """waymo:374-388: yaw of each polyline vertex from the step to it; vertex 0 copies vertex 1."""
    prev_x, prev_y = 0, 0
    out = []
    for xyz in lane_list:
        x, y = xyz.x, xyz.y
        out.append([x, y, np.arctan2(y - prev_y, x - prev_x)])
        prev_x, prev_y = x, y
    if len(out) > 1:
        out[0][2] = out[1][2]
    return np.array(out)


def lanes_of_frame(frame) -> np.ndarray:
    lane_pt_list = []
    for feature in frame.map_features:                                   # waymo:461-468
        if feature.HasField("lane"):
            lane_pt_list.append(get_yaws_from_lane_coords(list(feature.lane.polyline)))
    return np.vstack(lane_pt_list)


def default_points_fn(frame) -> np.ndarray:
    from waymo_open_dataset.utils import frame_utils                     # waymo:472-476
    range_images, camera_projections, _, range_image_top_pose = frame_utils.parse_range_image_and_camera_projection(frame)
    point_clouds, _ = frame_utils.convert_range_image_to_point_cloud(frame, range_images, camera_projections,
                                                                     range_image_top_pose, 0, False)
    return point_clouds[0]


def _wxyz_of_matrix32(rot32):
    """`quat = R.from_matrix(rotation_matrix.cpu()); rotation = (q[3], q[0], q[1], q[2])` (waymo:568-571)."""
    from scipy.spatial.transform import Rotation as R
    q = R.from_matrix(np.asarray(rot32)).as_quat()
    return (q[3], q[0], q[1], q[2])


_CAM_SPECS = {}


def cam_spec(cam_calib, ratio: float) -> CamSpec:
    """Chain + scaled intrinsics of one camera.  A segment's calibration is the same in every frame (the reference
    recomputes it per mask, waymo:557-593): kept by value."""
    key = (tuple(cam_calib.extrinsic.transform), tuple(cam_calib.intrinsic), float(ratio))
    spec = _CAM_SPECS.get(key)
    if spec is None:
        if len(_CAM_SPECS) > 4096:
            _CAM_SPECS.clear()
        spec = _CAM_SPECS[key] = _cam_spec(cam_calib, ratio)
    return spec


def _cam_spec(cam_calib, ratio: float) -> CamSpec:
    import torch
    axes = torch.Tensor([[0, -1, 0, 0], [0, 0, -1, 0], [1, 0, 0, 0], [0, 0, 0, 1]]).to(dtype=torch.float32)
    axes = torch.linalg.inv(axes)                                        # waymo:557-562
    tm = torch.from_numpy(np.array(cam_calib.extrinsic.transform).reshape(4, 4)).to(dtype=torch.float32)
    tm = torch.matmul(tm, axes)
    rotation = _wxyz_of_matrix32(tm[:3, :3].numpy())
    t = (-tm[:3, 3]).numpy()                                             # cam_pc.translate(-T[:3,3])
    Rt = np.asarray(Quaternion(rotation).rotation_matrix.T, np.float64).astype(np.float32)
    matrix = np.array(cam_calib.intrinsic, np.float32).tolist()          # waymo:586-593: fp64 maths, then cast
    K = np.array([[matrix[0], 0, matrix[2]], [0, matrix[1], matrix[3]], [0, 0, 1]]) * ratio
    K[2, 2] = 1
    return CamSpec([op_T(t), op_R(Rt)], K.astype(np.float32))


def frame_spec(frame, masks, data, cfg, points_fn: Callable) -> FrameSpec:
    pts = np.ascontiguousarray(np.asarray(points_fn(frame))[:, :3], np.float32)
    by_name = {int(c.name): c for c in frame.context.camera_calibrations}
    n = len(data["labels"])
    cam_nums = [int(c) for c in data["cam_nums"][:n]]
    for c in cam_nums:
        if c + 1 not in by_name:
            print("Invalid cam_num")                                     # waymo:516-518
            raise SystemExit
    used = sorted(set(cam_nums))
    cams = [cam_spec(by_name[c + 1], cfg.ratio) for c in used]
    remap = {c: k for k, c in enumerate(used)}
    return FrameSpec("waymo", [pts], [[]], cams if cams else [CamSpec([], np.eye(3, dtype=np.float32))],
                     np.asarray([remap[c] for c in cam_nums], np.int32), masks[:n], list(data["labels"]),
                     list(data["detection_scores"]), fourth=FOURTH_ONES, close_thresh=None, min_dist=cfg.min_dist,
                     token=f"{frame.context.name}:{frame.timestamp_micros}")


def centroid_to_global(centroid_xyz1: np.ndarray, frame) -> np.ndarray:
    """Vehicle -> global with the reference's fp32 ops (waymo:684-699): rotate by the pyquaternion
    matrix of the pose rotation, then add the pose translation."""
    import torch
    tm = torch.from_numpy(np.array(frame.pose.transform, np.float32).reshape(4, 4)).to(dtype=torch.float32)
    rotation = _wxyz_of_matrix32(tm[:3, :3].numpy())
    Rm = torch.from_numpy(Quaternion(rotation).rotation_matrix).to(dtype=torch.float32)
    p = torch.from_numpy(np.asarray(centroid_xyz1, np.float32).reshape(-1, 1)[:3].copy())
    p = torch.matmul(Rm, p)                                              # pcd.py rotate
    for i in range(3):                                                   # pcd.py translate
        p[i, :] = p[i, :] + tm[i, 3]
    return p[:3, 0].numpy()


def object_of(frame, label: str, score, centroid_global: np.ndarray, global_lane_yaw, shape_priors: dict) -> dict:
    """Pass 2 for one instance (waymo:803-858) -> a dict with the metrics_pb2.Object fields."""
    from scipy.spatial.transform import Rotation as R
    detection_name = B.get_detection_name(label)
    transform_matrix = np.array(frame.pose.transform, np.float32).reshape(4, 4)
    transform_matrix = np.linalg.inv(transform_matrix)
    centroid_pc = np.hstack([np.squeeze(np.array(centroid_global)), [1]])
    centroid = np.dot(transform_matrix, centroid_pc)[:3]
    extents = B.get_shape_prior(shape_priors, detection_name, waymo=True)
    if detection_name in B.VEHICLE_NAMES:
        global_align_mat = B.lane_align_matrix(global_lane_yaw)
        align_mat = np.dot(transform_matrix[:3, :3], global_align_mat)
        pushed = B.push_centroid(centroid, extents, Quaternion(matrix=global_align_mat), ego_frame=True)
        heading = R.from_matrix(align_mat).as_euler("xyz", degrees=False)[2]
    else:
        pushed = centroid
        heading = R.from_matrix(np.eye(3)).as_euler("xyz", degrees=False)[2]
    wname = B.NUSC_TO_WAYMO[detection_name]
    if wname not in WP.TYPE_BY_NAME:
        raise ValueError(detection_name)                                 # waymo:1060-1061 (barrier / traffic_cone)
    return {"context_name": frame.context.name, "frame_timestamp_micros": int(frame.timestamp_micros),
            "center_x": float(pushed[0]), "center_y": float(pushed[1]), "center_z": float(pushed[2]),
            "length": float(extents[1]), "width": float(extents[0]), "height": float(extents[2]),
            "heading": float(heading), "score": float(np.float32(float(score))), "type": WP.TYPE_BY_NAME[wname],
            "id": "unique object tracking ID"}


def centroids_to_global(centroids_xyz: np.ndarray, frame) -> np.ndarray:
    """`centroid_to_global` for all K centroids of a frame at once, bit for bit: torch 2.11's CPU `matmul` of a
    3x3 by a 3x1 float32 matrix evaluates row i as (R[i,1] p1 + R[i,2] p2) + R[i,0] p0 with every product and sum
    rounded to binary32 (no FMA; probed, and pinned by tests/test_host_logic.py against the per-instance torch
    calls), the translation is one more rounded add (waymo:684-699)."""
    tm = np.array(frame.pose.transform, np.float32).reshape(4, 4)
    rotation = _wxyz_of_matrix32(tm[:3, :3])
    Rm = np.asarray(Quaternion(rotation).rotation_matrix, np.float64).astype(np.float32)
    p = np.asarray(centroids_xyz, np.float32).reshape(-1, 3)
    out = np.empty_like(p)
    for i in range(3):
        out[:, i] = ((Rm[i, 1] * p[:, 1] + Rm[i, 2] * p[:, 2]) + Rm[i, 0] * p[:, 0]) + tm[i, 3]
    return out


def frame_objects(frame, labels, scores, centroids_global: np.ndarray, global_lane_yaws, shape_priors: dict) -> List[dict]:
    """`object_of` for all K instances of one frame (see `scene_objects`)."""
    return scene_objects([frame], np.zeros(len(labels), np.int64), labels, scores, centroids_global, global_lane_yaws, shape_priors)


def scene_objects(frames, frame_of, labels, scores, centroids_global: np.ndarray, global_lane_yaws, shape_priors: dict) -> List[dict]:
    """`object_of` (waymo:803-858) for all K instances of a scene in one numpy / scipy pass instead of ~0.5 ms of
    interpreter time per object: instance k belongs to `frames[frame_of[k]]`, whose inv(pose) takes it back to the
    vehicle frame.  Equal to the per-object function within 1e-10 m / rad (tests/test_host_logic.py); the class of an
    object decides, as there, whether it is pushed and lane aligned."""
    from scipy.spatial.transform import Rotation as R
    k = len(labels)
    if k == 0:
        return []
    names = [B.get_detection_name(l) for l in labels]
    wnames = [B.NUSC_TO_WAYMO[n] for n in names]
    for n, wn in zip(names, wnames):
        if wn not in WP.TYPE_BY_NAME:
            raise ValueError(n)                                          # waymo:1060-1061 (barrier / traffic_cone)
    extents = [B.get_shape_prior(shape_priors, n, waymo=True) for n in names]
    veh = np.fromiter((n in B.VEHICLE_NAMES for n in names), dtype=bool, count=k)
    fo = np.asarray(frame_of, dtype=np.int64)
    T32 = np.stack([np.linalg.inv(np.array(f.pose.transform, np.float32).reshape(4, 4)) for f in frames])    # float32, like :806-807
    T = T32.astype(np.float64)[fo]                                       # (K,4,4)
    cg = np.asarray(centroids_global).reshape(k, 3)
    pc = np.hstack([cg.astype(np.float64), np.ones((k, 1))])
    cents = np.einsum("kij,kj->ki", T, pc)[:, :3]
    gmats = B.lane_align_matrices(np.asarray(global_lane_yaws), veh)
    heading = np.full(k, R.from_matrix(np.eye(3)).as_euler("xyz", degrees=False)[2])
    if veh.any():
        ext = np.asarray([extents[i][:2] for i in np.flatnonzero(veh)], dtype=np.float64)
        cents[veh] = B.push_centroids(cents[veh], cents[veh], ext, B.quats_from_matrices(gmats[veh]))
        align = np.einsum("kij,kjl->kil", T[veh][:, :3, :3], gmats[veh])
        heading[veh] = R.from_matrix(align).as_euler("xyz", degrees=False)[:, 2]
    ctx = [(f.context.name, int(f.timestamp_micros)) for f in frames]
    cents, heading, fo = cents.tolist(), heading.tolist(), fo.tolist()
    return [{"context_name": ctx[fo[i]][0], "frame_timestamp_micros": ctx[fo[i]][1],
             "center_x": cents[i][0], "center_y": cents[i][1], "center_z": cents[i][2],
             "length": float(extents[i][1]), "width": float(extents[i][0]), "height": float(extents[i][2]),
             "heading": heading[i], "score": float(np.float32(float(scores[i]))), "type": WP.TYPE_BY_NAME[wnames[i]],
             "id": "unique object tracking ID"} for i in range(k)]


def nms_objects(objects: List[dict]) -> List[dict]:
    """Per-timestamp circle NMS (waymo:1108-1275)."""
    by_ts = {}
    for o in objects:
        by_ts.setdefault(o["frame_timestamp_micros"], []).append(o)
    final = []
    for ts, objs in by_ts.items():
        dets = np.array([[o["center_x"], o["center_y"], o["score"]] for o in objs], dtype=np.float64).reshape(-1, 3)
        keep = set(int(k) for k in B.circle_nms(dets, [o["type"] for o in objs], WAYMO_THRESHS))
        final.extend(o for c, o in enumerate(objs) if c in keep)
    return final


def run(cfg, scenes: Iterable, points_fn: Optional[Callable] = None, lifter=None, write: bool = True) -> List[dict]:
    from .lifter import Lifter
    from .shard import gather_labels, init_distributed, stage_device
    total_start = time.time()
    timer = new_timer()
    rank, world, local_rank = init_distributed()
    cfg.DEVICE = stage_device(cfg.DEVICE, world, local_rank)
    points_fn = points_fn or default_points_fn
    lifter = lifter or Lifter(cfg.DEVICE)
    with open(cfg.shape_priors_path) as f:
        shape_priors = json.load(f)
    local = {}
    n_scenes = 0
    for scene_num, (scene_name, scene_frames) in enumerate(scenes):
        n_scenes += 1
        if scene_num % world != rank:
            continue
        kept, lanes = [], [None]                  # (frame, data) of the frames that have mask files

        def numbered():
            for frame_num, frame in enumerate(scene_frames):             # the TFRecord is parsed here, in order
                if frame_num == 0:
                    # the reference takes the lanes inside its `try` (waymo:459-468), so a scene whose frame 0 has
                    # no mask files dies later with a NameError / stale lanes; here frame 0's map is read either way
                    lanes[0] = lanes_of_frame(frame)
                yield frame_num, frame

        def build(item, scene_name=scene_name):   # masks, range image -> points, calibration: on reader threads
            frame_num, frame = item
            t0 = time.time()
            try:
                masks, data = load_frame_masks(cfg.INPUT_DIR, scene_name, frame_num)
            except FileNotFoundError:
                return None                                              # waymo:453-455
            for label in data["labels"]:                                 # waymo:1060-1061 raises in pass 2, after all the
                if B.NUSC_TO_WAYMO.get(B.get_detection_name(label), "") not in WP.TYPE_BY_NAME:     # lifting: fail early
                    raise ValueError(f"{scene_name} frame {frame_num}: label {label!r} has no Waymo type")
            return frame, data, frame_spec(frame, masks, data, cfg, points_fn), time.time() - t0

        def frames():
            from .lifter import prefetch_map
            for item in prefetch_map(build, numbered(), getattr(cfg, "reader_threads", 8)):
                if item is None:
                    continue
                frame, data, spec, dt = item
                kept.append((frame, data))
                timer["io"] += dt
                yield spec

        cents, owners, per_frame = [], [], []     # global centroids, (kept index, instance index), (kept index, first, count)
        k = 0
        for res_batch in lifter.lift_frame_stream(frames(), batch_frames=cfg.batch_frames, timer=timer):
            for r in res_batch:
                frame, _ = kept[k]
                has = np.flatnonzero(np.asarray(r.medoid_local) >= 0)
                if has.size:
                    per_frame.append((k, len(cents), int(has.size)))
                    cents += list(centroids_to_global(np.asarray(r.centroids)[has, :3], frame))
                    owners += [(k, int(i)) for i in has]
                k += 1
        objs = []
        if cents:
            t0 = time.time()
            yaw_list, _, _ = B.lane_yaws_distances_and_coords(np.asarray(cents, np.float32), lanes[0], cfg.DEVICE)
            timer["closest lane"] += time.time() - t0
            frames_, frame_of, labels, scores = [], [], [], []
            for k, first, count in per_frame:
                frame, data = kept[k]
                frames_.append(frame)
                for _, i in owners[first:first + count]:
                    frame_of.append(len(frames_) - 1)
                    labels.append(data["labels"][i])
                    scores.append(data["detection_scores"][i])
            objs = scene_objects(frames_, frame_of, labels, scores, np.asarray(cents), np.asarray(yaw_list), shape_priors)
        local[scene_num] = objs
    merged = gather_labels(local, n_scenes) if world > 1 else [local.get(i) for i in range(n_scenes)]
    if rank != 0:
        return []
    objects = [o for part in merged if part for o in part]
    print("\nRunning NMS on the predictions.\n")
    t0 = time.time()
    final = nms_objects(objects)
    timer["nms"] += time.time() - t0
    print(len(final), len(objects))
    if write:
        os.makedirs(os.path.dirname(os.path.abspath(cfg.output_path)), exist_ok=True)
        with open(cfg.output_path, "wb") as f:
            f.write(WP.serialize_objects(final))
    timer["total"] += time.time() - total_start
    for operation in timer:
        print(operation, ":\t\t", timer[operation])
    return final
