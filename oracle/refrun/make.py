"""ORACLE tooling: build the synthetic on-disk datasets, EXECUTE the reference's three lifting
scripts on them (oracle/refrun/run.py, one subprocess each) and commit what they wrote as
tests/golden/ref_script_<ds>.json (+ the Waymo .bin).  Runs in the build container only
(/root/reference is needed); the GPU tests re-create the same datasets from the same seeds, run
the drop-in scripts of this repo and compare with these files.

    python -m oracle.refrun.make [nuscenes] [kitti] [waymo]

The dataset WRITERS are the product's test-support generators (cm3d_b200/synthetic.py,
synthetic_datasets.py): they only produce inputs.  `inputs_sha256` in every fixture is the digest
of those inputs; the tests recompute it, so a generator change cannot silently desynchronise them.
"""
from __future__ import annotations

import hashlib
import json
import os
import pickle
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")


# ------------------------------------------------------------------------------------- the datasets (shared with tests)
def nuscenes_scenes():
    from cm3d_b200 import synthetic as S
    return {f"scene-{k:04d}": [S.make_nuscenes_frame(7100 + 10 * k + f, n_sweeps=3, pts_per_sweep=8000, n_inst=16,
                                                     mask_div=2, dense_masks=False) for f in range(3)]
            for k in range(2)}


def kitti_frames():
    from cm3d_b200 import synthetic as S
    return [S.make_kitti_frame(8300 + f, n_pts=30000, n_inst=12, mask_div=1, dense_masks=False) for f in range(3)]


def waymo_scenes():
    from cm3d_b200 import synthetic as S
    out = {}
    for k in range(2):
        frames = [S.make_waymo_frame(8400 + 10 * k + f, n_pts=30000, n_inst=16, mask_div=2) for f in range(3)]
        for fr in frames:       # barrier / traffic_cone have no Waymo type: the reference raises ValueError on them
            fr.labels = [{"barrier": "car", "traffic_cone": "pedestrian"}.get(l, l) for l in fr.labels]
        out[f"segment-{k}"] = frames
    return out


def frames_digest(frames) -> str:
    from cm3d_b200.frames import frame_to_arrays
    h = hashlib.sha256()
    for f in frames:
        d = frame_to_arrays(f)
        for k in sorted(d):
            a = np.ascontiguousarray(d[k])
            h.update(k.encode())
            h.update(str(a.dtype).encode())
            h.update(a.tobytes())
    return h.hexdigest()


def write_nuscenes_tree(work, scenes):
    """WORK/data/nuScenes (+ fake_devkit.pkl), WORK/mask_outputs/nuscenes-detic.  Returns (nusc, map_factory)."""
    from cm3d_b200 import synthetic_datasets as SD
    root = os.path.join(work, "data", "nuScenes")
    input_dir = os.path.join(work, "mask_outputs", "nuscenes-detic")
    nusc, map_factory = SD.write_nuscenes(root, input_dir, scenes, ratio=0.64)
    maps = {}
    for name in scenes:
        scene = nusc.get("scene", nusc.field2token("scene", "name", name)[0])
        maps[nusc.get("log", scene["log_token"])["location"]] = {t: np.asarray(p).tolist() for t, p in map_factory(nusc, scene)._poly.items()}
    with open(os.path.join(root, "fake_devkit.pkl"), "wb") as f:
        pickle.dump({"tables": nusc.tables, "maps": maps, "scene_names": list(scenes)}, f)
    return nusc, map_factory, root, input_dir


def write_kitti_tree(work, frames):
    from cm3d_b200 import synthetic_datasets as SD
    root = os.path.join(work, "data", "kitti")
    input_dir = os.path.join(work, "mask_outputs", "kitti")
    SD.write_kitti(root, input_dir, frames)
    return root, input_dir


def _plain(o):
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    if isinstance(o, np.ndarray) or not hasattr(o, "__dict__"):
        return o
    return {k: _plain(v) for k, v in o.__dict__.items() if not callable(v)}


def write_waymo_tree(work, scenes, drop=(("segment-1", 1),)):
    """WORK/data/waymo-v1.4.2/waymo_format/training/<scene> (pickled frame trees standing in for the
    TFRecords, after 680 pad entries: the reference iterates scene_list[680:710]), masks under
    WORK/mask_outputs/waymo-detic/<scene>/.  Returns [(scene_name, duck-typed frames)]."""
    from cm3d_b200 import synthetic_datasets as SD
    root = os.path.join(work, "data", "waymo-v1.4.2", "waymo_format", "training")
    input_dir = os.path.join(work, "mask_outputs", "waymo-detic")
    os.makedirs(root, exist_ok=True)
    for k in range(680):
        open(os.path.join(root, f"000-pad-{k:04d}"), "w").close()
    out = []
    for name, frames in scenes.items():
        duck = SD.waymo_frames(name, input_dir, frames, ratio=1024 / 1920)
        with open(os.path.join(root, name), "wb") as f:
            pickle.dump([_plain(fr) for fr in duck], f)
        out.append((name, duck))
    for scene, fnum in drop:                 # a frame without mask files is skipped (waymo:450-455, :789-792)
        # BOTH files go: pass 1 of the reference skips a frame when either is missing, its pass 2 only when the
        # json is - with just the pkl gone its two instance counters drift apart (a reference bug, not a target)
        os.remove(os.path.join(input_dir, scene, f"{fnum}_masks.pkl"))
        os.remove(os.path.join(input_dir, scene, f"{fnum}_data.json"))
    return out, root, input_dir


# ------------------------------------------------------------------------------------- run + collect
def _run(ds, work):
    env = dict(os.environ, PYTHONPATH=ROOT, PYTHONDONTWRITEBYTECODE="1")
    log = subprocess.run([sys.executable, "-m", "oracle.refrun.run", ds, work], cwd=ROOT, env=env, capture_output=True, text=True)
    if log.returncode != 0:
        sys.stderr.write(log.stdout[-3000:] + "\n" + log.stderr[-6000:])
        raise SystemExit(f"reference {ds} script failed")
    with open(os.path.join(work, "refrun_info.json")) as f:
        return json.load(f)


def _provenance(info):
    import scipy
    import torch
    return {"script": info["script"], "source_edits": info["edits"], "ended": info["ended"],
            "stubs": "oracle/refrun/stubs.py (absent pip dependencies only)",
            "made_with": f"torch {torch.__version__} cpu cap={torch.backends.cpu.get_cpu_capability()} threads=1, "
                         f"numpy {np.__version__}, scipy {scipy.__version__}"}


def make_nuscenes():
    scenes = nuscenes_scenes()
    work = tempfile.mkdtemp(prefix="refrun_nusc_")
    try:
        write_nuscenes_tree(work, scenes)
        info = _run("nuscenes", work)
        with open(os.path.join(work, "outputs", "nuscenes", "pseudolabels_minival.json")) as f:
            out = json.load(f)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    fix = {"provenance": _provenance(info), "inputs_sha256": frames_digest([f for fs in scenes.values() for f in fs]),
           "last_scene_centroid_ids": info["namespace"].get("centroid_ids"),
           "last_scene_centroids": info["namespace"].get("all_centroids_list"), "pseudolabels": out}
    with open(os.path.join(GOLDEN, "ref_script_nuscenes.json"), "w") as f:
        json.dump(fix, f)
    n = sum(len(v) for v in out["results"].values())
    print(f"nuscenes: {len(out['results'])} samples, {n} boxes after NMS, {len(fix['last_scene_centroid_ids'])} centroids in the last scene")


def make_kitti():
    frames = kitti_frames()
    work = tempfile.mkdtemp(prefix="refrun_kitti_")
    try:
        root, _ = write_kitti_tree(work, frames)
        info = _run("kitti", work)
        files = {}
        for d in ("pred", "pseudo"):
            for k in range(len(frames)):
                with open(os.path.join(root, "training", d, f"{k:06}.txt")) as f:
                    files[f"{d}/{k:06}.txt"] = f.read().splitlines()
    finally:
        shutil.rmtree(work, ignore_errors=True)
    fix = {"provenance": _provenance(info), "inputs_sha256": frames_digest(frames), "files": files}
    with open(os.path.join(GOLDEN, "ref_script_kitti.json"), "w") as f:
        json.dump(fix, f)
    print("kitti:", {k: len(v) for k, v in files.items()}, "|", info["ended"])


def make_waymo():
    scenes = waymo_scenes()
    work = tempfile.mkdtemp(prefix="refrun_waymo_")
    try:
        write_waymo_tree(work, scenes)
        info = _run("waymo", work)
        with open(os.path.join(work, "outputs", "waymo", "pred_0307_detic_train_680_710.bin"), "rb") as f:
            blob = f.read()
    finally:
        shutil.rmtree(work, ignore_errors=True)
    with open(os.path.join(GOLDEN, "ref_script_waymo.bin"), "wb") as f:
        f.write(blob)
    fix = {"provenance": _provenance(info), "inputs_sha256": frames_digest([f for fs in scenes.values() for f in fs]),
           "bin": "ref_script_waymo.bin", "bin_sha256": hashlib.sha256(blob).hexdigest(),
           "last_scene_centroid_ids": info["namespace"].get("centroid_ids"),
           "last_scene_centroids": info["namespace"].get("all_centroids_list")}
    with open(os.path.join(GOLDEN, "ref_script_waymo.json"), "w") as f:
        json.dump(fix, f)
    print(f"waymo: {len(blob)} bytes of metrics.Objects")


def main(which=None):
    which = which or ["nuscenes", "kitti", "waymo"]
    os.makedirs(GOLDEN, exist_ok=True)
    for ds in which:
        {"nuscenes": make_nuscenes, "kitti": make_kitti, "waymo": make_waymo}[ds]()


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main(sys.argv[1:] or None)
