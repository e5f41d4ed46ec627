"""On-disk synthetic datasets in the layouts the reference's scripts read, built from the
seeded FrameSpecs of `synthetic.py`.  Used by the drop-in script tests and demos: the scripts
then run end to end (dataset accessors -> masks on disk -> lifting -> label files) without
nuscenes-devkit / KITTI downloads / TFRecords.

  write_nuscenes(...)  -> FakeNuScenes  (get / field2token / dataroot), FakeNuScenesMap factory,
                          sweeps as (N,5) float32 .bin, `{scene}/{f}_masks.pkl` + `{f}_data.json`
  write_kitti(...)     -> KITTI object folders: training/velodyne/%06d.bin, training/calib/%06d.txt,
                          `{f}_masks.pkl` + `{f}_data.json` (no scene directory, no cam_nums)
  waymo_frames(...)    -> duck-typed Waymo frames (context, pose, camera calibrations, points)
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Dict, List, Sequence

import numpy as np

from . import synthetic as S
from .frames import FrameSpec
from .rle import rle_counts_to_runs, runs_to_rle_string


def _masks_pkl(frame: FrameSpec) -> list:
    """COCO RLE dicts as gen_2d_masks_detic.py:468-471,506 writes them: size=[W,H], counts=bytes."""
    rles = S.dense_to_rle(frame.masks) if isinstance(frame.masks, np.ndarray) else frame.masks
    return [{"size": [int(r.size[0]), int(r.size[1])], "counts": runs_to_rle_string(rle_counts_to_runs(r.counts))}
            for r in rles]


def _write_masks(dirpath: str, frame_num: int, frame: FrameSpec, with_cams: bool):
    os.makedirs(dirpath, exist_ok=True)
    with open(os.path.join(dirpath, f"{frame_num}_masks.pkl"), "wb") as f:
        pickle.dump(_masks_pkl(frame), f)
    data = {"labels": list(frame.labels), "detection_scores": [float(s) for s in frame.scores]}
    if with_cams:
        data["cam_nums"] = [int(c) for c in frame.cam_nums]
    with open(os.path.join(dirpath, f"{frame_num}_data.json"), "w") as f:
        json.dump(data, f)


def _quat(R32) -> list:
    """(w,x,y,z) of an fp32 rotation matrix (orthogonal only to ~1e-7, so scipy's projection is used)."""
    from scipy.spatial.transform import Rotation
    x, y, z, w = Rotation.from_matrix(np.asarray(R32, np.float64)).as_quat()
    return [float(w), float(x), float(y), float(z)]


# ------------------------------------------------------------------------------------- nuScenes
class FakeNuScenes:
    """The slice of nuscenes-devkit's NuScenes the lifting script touches."""

    def __init__(self, dataroot: str):
        self.dataroot = dataroot
        self.tables: Dict[str, Dict[str, dict]] = {k: {} for k in
                                                   ("scene", "sample", "sample_data", "calibrated_sensor", "ego_pose", "log")}

    def get(self, table: str, token: str) -> dict:
        return self.tables[table][token]            # KeyError on '' like the devkit (nuscenes:460-463)

    def field2token(self, table: str, field: str, value) -> List[str]:
        return [t for t, r in self.tables[table].items() if r[field] == value]

    def add(self, table: str, rec: dict) -> str:
        self.tables[table][rec["token"]] = rec
        return rec["token"]


class FakeNuScenesMap:
    """lane / lane_connector records + discretize_lanes, as used at nuscenes:228-240."""

    def __init__(self, polylines: Dict[str, np.ndarray]):
        toks = list(polylines)
        self.lane = [{"token": t} for t in toks[::2]]
        self.lane_connector = [{"token": t} for t in toks[1::2]]
        self._poly = polylines
        self.drivable_area = []

    def discretize_lanes(self, tokens: Sequence[str], resolution_meters: float) -> Dict[str, list]:
        return {t: [tuple(float(v) for v in p) for p in self._poly[t]] for t in tokens}


def _lanes_around(rng, centre_xy, n_lanes=6, length=120.0, step=0.5) -> Dict[str, np.ndarray]:
    out = {}
    for k in range(n_lanes):
        yaw = rng.uniform(-np.pi, np.pi)
        off = rng.uniform(-30, 30, 2)
        s = np.arange(-length / 2, length / 2, step)
        curv = rng.normal(0, 0.004)
        th = yaw + curv * s
        x = centre_xy[0] + off[0] + np.cumsum(np.cos(th)) * step
        y = centre_xy[1] + off[1] + np.cumsum(np.sin(th)) * step
        out[f"lane-{k}"] = np.stack([x, y, th], 1)
    return out


NUSC_CAM_LIST = ("CAM_FRONT", "CAM_FRONT_RIGHT", "CAM_BACK_RIGHT", "CAM_BACK", "CAM_BACK_LEFT", "CAM_FRONT_LEFT")


def write_nuscenes(root: str, input_dir: str, scenes: Dict[str, List[FrameSpec]], ratio: float = 0.64,
                   cam_list: Sequence[str] = NUSC_CAM_LIST):
    """scenes: name -> FrameSpecs (from synthetic.make_nuscenes_frame).  Returns (nusc, map_factory)."""
    nusc = FakeNuScenes(root)
    maps = {}
    os.makedirs(os.path.join(root, "sweeps", "LIDAR_TOP"), exist_ok=True)
    for si, (scene_name, frames) in enumerate(scenes.items()):
        rng = np.random.default_rng(9000 + si)
        sample_tokens = [f"{scene_name}-sample-{f}" for f in range(len(frames))]
        nusc.add("log", {"token": f"{scene_name}-log", "location": f"synthetic-town-{si}"})
        nusc.add("scene", {"token": f"{scene_name}-tok", "name": scene_name, "first_sample_token": sample_tokens[0],
                           "log_token": f"{scene_name}-log"})
        centre = None
        for f, frame in enumerate(frames):
            data = {}
            # LiDAR sweeps: a chain of sample_data records linked by `next`
            for s, (raw, ops) in enumerate(zip(frame.sweeps, frame.sweep_ops)):
                tok = f"{scene_name}-{f}-lidar-{s}"
                rel = os.path.join("sweeps", "LIDAR_TOP", f"{scene_name}_{f}_{s}.bin")
                np.ascontiguousarray(raw, np.float32).tofile(os.path.join(root, rel))
                (_, R_ls), (_, t_ls), (_, R_e), (_, t_e) = ops
                cs = nusc.add("calibrated_sensor", {"token": tok + "-cs", "rotation": _quat(R_ls),
                                                    "translation": [float(v) for v in t_ls]})
                ep = nusc.add("ego_pose", {"token": tok + "-pose", "rotation": _quat(R_e),
                                           "translation": [float(v) for v in t_e]})
                nxt = f"{scene_name}-{f}-lidar-{s + 1}" if s + 1 < len(frame.sweeps) else ""
                nusc.add("sample_data", {"token": tok, "filename": rel, "calibrated_sensor_token": cs,
                                         "ego_pose_token": ep, "next": nxt})
                if s == 0:
                    data["LIDAR_TOP"] = tok
                    if centre is None:
                        centre = np.asarray(t_e[:2], np.float64)
            for c, cam in enumerate(frame.cams):
                tok = f"{scene_name}-{f}-cam-{c}"
                (_, nt_e), (_, R_eT), (_, nt_cs), (_, R_csT) = cam.ops
                K = np.asarray(cam.K, np.float64) / ratio
                K[2, 2] = 1.0
                cs = nusc.add("calibrated_sensor", {"token": tok + "-cs", "rotation": _quat(R_csT.T),
                                                    "translation": [float(-v) for v in nt_cs],
                                                    "camera_intrinsic": K.tolist()})
                ep = nusc.add("ego_pose", {"token": tok + "-pose", "rotation": _quat(R_eT.T),
                                           "translation": [float(-v) for v in nt_e]})
                nusc.add("sample_data", {"token": tok, "filename": "", "calibrated_sensor_token": cs,
                                         "ego_pose_token": ep, "next": ""})
                data[cam_list[c]] = tok
            nusc.add("sample", {"token": sample_tokens[f], "data": data,
                                "next": sample_tokens[f + 1] if f + 1 < len(frames) else ""})
            _write_masks(os.path.join(input_dir, scene_name), f, frame, with_cams=True)
        maps[f"synthetic-town-{si}"] = FakeNuScenesMap(_lanes_around(rng, centre))

    def map_factory(nusc_, scene):
        return maps[nusc_.get("log", scene["log_token"])["location"]]
    return nusc, map_factory


# ------------------------------------------------------------------------------------- KITTI
def write_kitti(root: str, input_dir: str, frames: List[FrameSpec]):
    """KITTI object layout (kitti_object.py:27-79) + per-frame mask files (kitti/2d_to_3d.py:1001-1002)."""
    for sub in ("velodyne", "calib", "image_2", "label_2"):
        os.makedirs(os.path.join(root, "training", sub), exist_ok=True)
    for f, frame in enumerate(frames):
        np.ascontiguousarray(frame.sweeps[0][:, :4], np.float32).tofile(
            os.path.join(root, "training", "velodyne", "%06d.bin" % f))
        with open(os.path.join(root, "training", "calib", "%06d.txt" % f), "w") as fh:
            fh.write(S.kitti_calib_text())
        _write_masks(input_dir, f, frame, with_cams=False)


# ------------------------------------------------------------------------------------- Waymo
class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def waymo_frames(scene_name: str, input_dir: str, frames: List[FrameSpec], ratio: float = 1024 / 1920):
    """Duck-typed `dataset_pb2.Frame`s: .context.name/.camera_calibrations[].name/.extrinsic.transform/
    .intrinsic, .pose.transform, .timestamp_micros, .map_features, plus `.points_vehicle` (N,3) standing
    in for `convert_range_image_to_point_cloud(...)[0]` (waymo/2d_to_3d.py:472-476)."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(scene_name.encode()))      # stable across processes (str hash is salted)
    axes = np.array([[0, -1, 0, 0], [0, 0, -1, 0], [1, 0, 0, 0], [0, 0, 0, 1]], np.float64)
    out = []
    yaw0 = rng.uniform(-np.pi, np.pi)
    for f, frame in enumerate(frames):
        calibs = []
        for c, cam in enumerate(frame.cams):
            (_, nt), (_, R_T) = cam.ops                     # T(-t), R(R^T)   (waymo:573-575)
            # the script computes transform @ inv(axes): store transform = [R|t] @ axes so that it gets [R|t] back
            M = np.eye(4)
            M[:3, :3] = np.asarray(R_T, np.float64).T
            M[:3, 3] = -np.asarray(nt, np.float64)
            ext = M @ axes
            K = np.asarray(cam.K, np.float64) / ratio
            intr = [float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), 0.0, 0.0, 0.0, 0.0, 0.0]
            calibs.append(_NS(name=c + 1, extrinsic=_NS(transform=[float(v) for v in ext.reshape(-1)]), intrinsic=intr))
        pose = np.eye(4)
        yaw = yaw0 + 0.01 * f
        pose[:3, :3] = S._rotz(yaw)
        pose[:3, 3] = [500.0 + 3.0 * f * np.cos(yaw0), 800.0 + 3.0 * f * np.sin(yaw0), 10.0]
        feats = []
        if f == 0:
            for k, poly in enumerate(_lanes_around(rng, pose[:2, 3]).values()):
                pts = [_NS(x=float(p[0]), y=float(p[1]), z=0.0) for p in poly]
                feats.append(_NS(lane=_NS(polyline=pts), HasField=lambda name: name == "lane"))
        out.append(_NS(context=_NS(name=scene_name, camera_calibrations=calibs),
                       pose=_NS(transform=[float(v) for v in pose.reshape(-1)]),
                       timestamp_micros=1_000_000 + 100_000 * f, map_features=feats,
                       points_vehicle=np.ascontiguousarray(frame.sweeps[0][:, :3], np.float32)))
        _write_masks(os.path.join(input_dir, scene_name), f, frame, with_cams=True)
    return out
