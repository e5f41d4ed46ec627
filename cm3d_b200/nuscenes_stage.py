"""nuScenes lifting stage: the `__main__` of the reference's src/nuscenes/2d_to_3d.py
(README: 2d_to_3d_new.py) as a function, with the per-frame / per-mask body replaced by the
CUDA path (`cm3d_b200.Lifter`).

`run(cfg, nusc, nusc_map_factory, scene_names)` follows src/nuscenes/2d_to_3d.py:343-938:
  pass 1 (:413-695)  per frame: read `{f}_masks.pkl` + `{f}_data.json`, aggregate sweeps in the
                     global frame, per mask find the LiDAR points inside it and their medoid
                     -> HERE: build one FrameSpec per frame, lift whole batches on the GPU;
  lane yaw (:704-706) closest discretised lane point per centroid -> cm3d_nearest_lane;
  pass 2 (:733-825)  class name, shape prior, lane-aligned rotation, push-back, box dict;
  NMS (:844-924)     per-sample circle NMS;  JSON (:929-930).
`cfg` carries the script's module constants (VER_NAME, INPUT_PATH, OUTPUT_DIR, INPUT_DIR,
CAM_LIST, ATTRIBUTE_NAMES, DEVICE) and the literals of its `__main__` (min_dist, ratio,
n_sweeps, pointsensor_channel, output_name, threshs_by_label, batch_frames).

`nusc` is duck-typed on the nuscenes-devkit calls the reference makes: `get(table, token)`,
`field2token(table, field, value)`, `dataroot`.  `nusc_map_factory(nusc, scene)` returns an
object with `lane`, `lane_connector` and `discretize_lanes(tokens, resolution)`.
"""
from __future__ import annotations

import json
import os
import pickle
import time
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import boxes as B
from .frames import CamSpec, FrameSpec, FOURTH_COL3, RLEMask, op_R, op_T
from .quat import Quaternion

DEFAULTS = dict(
    VER_NAME="v1.0-trainval", INPUT_PATH="../../data/nuScenes/", OUTPUT_DIR="../../outputs/nuscenes/",
    INPUT_DIR="../../mask_outputs/nuscenes-detic/",
    CAM_LIST=["CAM_FRONT", "CAM_FRONT_RIGHT", "CAM_BACK_RIGHT", "CAM_BACK", "CAM_BACK_LEFT", "CAM_FRONT_LEFT"],
    ATTRIBUTE_NAMES=B.ATTRIBUTE_NAMES, DEVICE="cuda:0",
    min_dist=2.3, floor_thresh=0.6, ratio=0.64, n_sweeps=3, pointsensor_channel="LIDAR_TOP",
    shape_priors_path="cfg/shape_priors_chatgpt.json", output_name="pseudolabels_minival.json",
    threshs_by_label=B.THRESHS_BY_LABEL, batch_frames=32, reader_threads=8,
)


def make_cfg(**overrides) -> SimpleNamespace:
    d = dict(DEFAULTS)
    d.update(overrides)
    return SimpleNamespace(**d)


def new_timer() -> Dict[str, float]:
    """The reference's stopwatch keys (nuscenes:368-378)."""
    return {k: 0 for k in ("io", "points in mask", "mvp", "medoid", "drivable", "closest lane", "lane pose", "nms", "total")}


# ------------------------------------------------------------------------------------- helpers
def count_frames(nusc, sample) -> int:
    """nuscenes:88-101."""
    n = 1
    while sample["next"] != "":
        n += 1
        sample = nusc.get("sample", sample["next"])
    return n


def load_frame_masks(input_dir: str, scene_name: Optional[str], frame_num: int):
    """`{f}_masks.pkl` (list of COCO RLE dicts, size=[W,H]) + `{f}_data.json` (nuscenes:422-423).
    The compressed `counts` strings are NOT decoded here: they go to the GPU as they are."""
    d = input_dir if scene_name is None else os.path.join(input_dir, scene_name)
    with open(os.path.join(d, f"{frame_num}_masks.pkl"), "rb") as f:
        masks_compressed = pickle.load(f)
    with open(os.path.join(d, f"{frame_num}_data.json")) as f:
        data = json.load(f)
    masks = [RLEMask((int(m["size"][0]), int(m["size"][1])), m["counts"]) for m in masks_compressed]
    return masks, data


def load_lidar_bin(path: str) -> np.ndarray:
    """LidarPointCloud.from_file (utils/pcd.py:246-257): (N,5) float32 rows x,y,z,intensity,ring."""
    assert path.endswith(".bin"), "Unsupported filetype {}".format(path)
    scan = np.fromfile(path, dtype=np.float32)
    return scan.reshape((-1, 5))


def _f32(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32)


_ROT = {}


def _rot32(rotation, transpose=False) -> np.ndarray:
    """float32 of `Quaternion(rotation).rotation_matrix` (or its transpose).  A scene's calibrated_sensor records
    repeat for every sweep and camera of every frame, so the matrices are kept by quaternion value."""
    key = (tuple(rotation), transpose)
    m = _ROT.get(key)
    if m is None:
        if len(_ROT) > 8192:
            _ROT.clear()
        m = Quaternion(rotation).rotation_matrix
        m = _ROT[key] = _f32(m.T if transpose else m)
        m.setflags(write=False)
    return m


def frame_spec(nusc, sample, masks: List[RLEMask], data: dict, cfg) -> FrameSpec:
    """Everything nuscenes:430-503 and the per-mask constants of :569-587 read from the devkit."""
    pointsensor_next = nusc.get("sample_data", sample["data"][cfg.pointsensor_channel])
    sweeps, sweep_ops = [], []
    for _ in range(cfg.n_sweeps):                                                 # :437
        sweeps.append(load_lidar_bin(os.path.join(nusc.dataroot, pointsensor_next["filename"])))
        cs_record = nusc.get("calibrated_sensor", pointsensor_next["calibrated_sensor_token"])
        poserecord = nusc.get("ego_pose", pointsensor_next["ego_pose_token"])
        sweep_ops.append([op_R(_rot32(cs_record["rotation"])),                                   # :451-457
                          op_T(_f32(np.array(cs_record["translation"]))),
                          op_R(_rot32(poserecord["rotation"])),
                          op_T(_f32(np.array(poserecord["translation"])))])
        try:
            pointsensor_next = nusc.get("sample_data", pointsensor_next["next"])  # :460-463
        except KeyError:
            break
    cams = []
    ratio32 = np.float32(cfg.ratio)
    for camera in cfg.CAM_LIST:                                                   # :490-503
        cam_data = nusc.get("sample_data", sample["data"][camera])
        poserecord = nusc.get("ego_pose", cam_data["ego_pose_token"])
        cs_record = nusc.get("calibrated_sensor", cam_data["calibrated_sensor_token"])
        K = _f32(np.array(cs_record["camera_intrinsic"])) * ratio32               # :585-587
        K[2, 2] = 1
        cams.append(CamSpec([op_T(_f32(-np.array(poserecord["translation"]))),    # :569-577
                             op_R(_rot32(poserecord["rotation"], True)),
                             op_T(_f32(-np.array(cs_record["translation"]))),
                             op_R(_rot32(cs_record["rotation"], True))], K))
    n = len(data["labels"])
    return FrameSpec("nuscenes", sweeps, sweep_ops, cams, np.asarray(data["cam_nums"][:n], np.int32), masks[:n],
                     list(data["labels"]), list(data["detection_scores"]), fourth=FOURTH_COL3,
                     close_thresh=float(np.float32(np.sqrt(cfg.min_dist))), min_dist=cfg.min_dist,
                     token=sample["token"])


def get_all_lane_points_in_scene(nusc_map):
    """nuscenes:228-240."""
    lane_records = nusc_map.lane + nusc_map.lane_connector
    lane_tokens = [lane["token"] for lane in lane_records]
    lane_pt_dict = nusc_map.discretize_lanes(lane_tokens, 0.5)
    all_lane_pts = []
    for lane_pts in lane_pt_dict.values():
        for lane_pt in lane_pts:
            all_lane_pts.append(lane_pt)
    return lane_pt_dict, all_lane_pts


def default_map_factory(input_path: str) -> Callable:
    maps = {}                                                                      # one NuScenesMap per location, not per scene

    def factory(nusc, scene):                                                      # nuscenes:216-224
        from nuscenes.map_expansion.map_api import NuScenesMap
        log = nusc.get("log", scene["log_token"])
        if log["location"] not in maps:
            maps[log["location"]] = NuScenesMap(dataroot=input_path, map_name=log["location"])
        return maps[log["location"]]
    return factory


# ------------------------------------------------------------------------------------- the stage
def lift_scenes(nusc, scene_items: Sequence[tuple], cfg, lifter, timer):
    """Pass 1 over the scenes `[(scene_num, scene_name), ...]` through ONE frame stream: the reader threads are
    already on the next scene's files while the GPU finishes this one.  Yields `(scene_num, scene)` in order as
    each scene's last frame comes back; scene = {"samples": [token...], "data": [data json...],
    "centroid_ids": [...], "centroids": (K,3) float32, "lidar_pose": [poserecord per frame]}."""
    from collections import deque
    from .lifter import prefetch_map
    plans, work = [], []
    for p, (scene_num, scene_name) in enumerate(scene_items):
        scene = nusc.get("scene", nusc.field2token("scene", "name", scene_name)[0])
        sample = nusc.get("sample", scene["first_sample_token"])
        samples = [sample]                              # the scene's sample chain (nuscenes:413-415, :693-694)
        for _ in range(count_frames(nusc, sample) - 1):
            samples.append(nusc.get("sample", samples[-1]["next"]))
        plans.append((scene_num, scene_name, samples))
        work += [(p, frame_num) for frame_num in range(len(samples))]
    outs = [{"samples": [], "data": [], "lidar_pose": [], "centroid_ids": [], "centroids": [], "_done": 0, "_ids": 0}
            for _ in plans]
    owners = deque()

    def build(item):
        t0 = time.time()
        p, frame_num = item
        s = plans[p][2][frame_num]
        masks, data = load_frame_masks(cfg.INPUT_DIR, plans[p][1], frame_num)
        spec = frame_spec(nusc, s, masks, data, cfg)
        ps = nusc.get("sample_data", s["data"][cfg.pointsensor_channel])
        return p, spec, s["token"], data, nusc.get("ego_pose", ps["ego_pose_token"]), time.time() - t0

    def frames():
        for p, spec, token, data, pose, dt in prefetch_map(build, work, getattr(cfg, "reader_threads", 8)):
            out = outs[p]
            out["samples"].append(token)
            out["data"].append(data)
            out["lidar_pose"].append(pose)
            owners.append(p)
            timer["io"] += dt
            yield spec

    for res_batch in lifter.lift_frame_stream(frames(), batch_frames=cfg.batch_frames, timer=timer):
        for r in res_batch:
            p = owners.popleft()
            out = outs[p]
            has = np.flatnonzero(np.asarray(r.medoid_local) >= 0)       # empty mask -> `continue` (:626-628)
            out["centroid_ids"] += (out["_ids"] + has).tolist()
            out["centroids"] += [r.centroids[i] for i in has]
            out["_ids"] += len(r.medoid_local)
            out["_done"] += 1
            if out["_done"] == len(plans[p][2]):
                out["centroids"] = np.asarray(out["centroids"], np.float32).reshape(-1, 3)
                del out["_done"], out["_ids"]
                yield plans[p][0], out


def lift_scene(nusc, scene_name: str, cfg, lifter, timer) -> dict:
    """Pass 1 of one scene (see `lift_scenes`)."""
    return next(lift_scenes(nusc, [(0, scene_name)], cfg, lifter, timer))[1]


def scene_boxes(scene: dict, lane_pt_list, cfg, shape_priors: dict, timer) -> Dict[str, list]:
    """Lane yaw + pass 2 of one scene (nuscenes:704-825): sample token -> list of box dicts."""
    results = {tok: [] for tok in scene["samples"]}
    if len(scene["centroid_ids"]) == 0:
        return results
    t0 = time.time()
    yaw_list, min_distance_list, _ = B.lane_yaws_distances_and_coords(scene["centroids"], lane_pt_list, cfg.DEVICE)
    timer["closest lane"] += time.time() - t0
    where = {cid: k for k, cid in enumerate(scene["centroid_ids"])}
    toks, labels, scores, ks, poses = [], [], [], [], []
    id_offset = -1
    for tok, data, pose in zip(scene["samples"], scene["data"], scene["lidar_pose"]):
        for label, score, c in zip(data["labels"], data["detection_scores"], data["cam_nums"]):
            id_offset += 1
            k = where.get(id_offset)
            if k is None:
                continue
            toks.append(tok)
            labels.append(label)
            scores.append(score)
            ks.append(k)
            poses.append(pose["translation"])
    t0 = time.time()
    ks = np.asarray(ks, dtype=np.int64)
    for box in B.nuscenes_boxes(toks, labels, scores, scene["centroids"][ks], np.asarray(yaw_list)[ks], shape_priors,
                                np.asarray(poses, dtype=np.float64).reshape(-1, 3), cfg.ATTRIBUTE_NAMES):
        results[box["sample_token"]].append(box)
    timer["boxes"] = timer.get("boxes", 0.0) + time.time() - t0
    return results


def run(cfg, nusc, nusc_map_factory: Callable, scene_names: Sequence[str], lifter=None, write: bool = True) -> dict:
    """The whole script.  Under torchrun (WORLD_SIZE > 1) scenes are sharded over the ranks
    (scene i -> rank i mod world), rank 0 merges the per-sample results and writes the file."""
    from .lifter import Lifter
    from .shard import gather_labels, init_distributed, stage_device
    total_start = time.time()
    timer = new_timer()
    rank, world, local_rank = init_distributed()
    cfg.DEVICE = stage_device(cfg.DEVICE, world, local_rank)
    lifter = lifter or Lifter(cfg.DEVICE)
    with open(cfg.shape_priors_path) as f:
        shape_priors = json.load(f)
    predictions = {"meta": {"use_camera": True, "use_lidar": False, "use_radar": False, "use_map": True,
                            "use_external": False}, "results": {}}
    local, lanes = {}, {}
    items = [(scene_num, scene_name) for scene_num, scene_name in enumerate(scene_names) if scene_num % world == rank]
    for scene_num, scene in lift_scenes(nusc, items, cfg, lifter, timer):
        scene_rec = nusc.get("scene", nusc.field2token("scene", "name", scene_names[scene_num])[0])
        nusc_map = nusc_map_factory(nusc, scene_rec)
        if id(nusc_map) not in lanes:                   # discretised once per map object (the factory keeps one per location)
            lanes = {id(nusc_map): (nusc_map, get_all_lane_points_in_scene(nusc_map)[1])}
        local[scene_num] = scene_boxes(scene, lanes[id(nusc_map)][1], cfg, shape_priors, timer)
    merged = gather_labels(local, len(scene_names)) if world > 1 else [local.get(i) for i in range(len(scene_names))]
    if rank != 0:
        return {}
    for part in merged:
        if part:
            predictions["results"].update(part)
    print("\nRunning NMS on the predictions.\n")
    t0 = time.time()
    final_predictions = B.nms_predictions(predictions, cfg.threshs_by_label)
    timer["nms"] += time.time() - t0
    if write:
        os.makedirs(cfg.OUTPUT_DIR, exist_ok=True)
        with open(os.path.join(cfg.OUTPUT_DIR, cfg.output_name), "w") as f:
            json.dump(final_predictions, f)
        print(f"wrote {len(final_predictions['results'])} samples.")
    timer["total"] += time.time() - total_start
    for operation in timer:
        print(operation, ":\t\t", timer[operation])
    return final_predictions
