#!/usr/bin/env python
"""Runs the three drop-in scripts on deterministic synthetic on-disk datasets and writes their outputs
under --out.  Single process, or under torchrun (scenes / frames sharded over the ranks):

    python tools/run_synthetic_scripts.py --out /tmp/a
    torchrun --nproc-per-node 2 tools/run_synthetic_scripts.py --out /tmp/b
    diff -r /tmp/a /tmp/b        # sharding must not change a byte of the labels
"""
import argparse
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--data", default=None, help="where the synthetic datasets go (default: <out>/_data_rank<r>)")
    args = ap.parse_args()
    from cm3d_b200 import synthetic as S
    from cm3d_b200 import synthetic_datasets as SD
    rank = int(os.environ.get("RANK", "0"))
    data = args.data or os.path.join(args.out, f"_data_rank{rank}")     # every rank writes its own identical copy
    os.makedirs(args.out, exist_ok=True)

    # nuScenes: 3 scenes x 3 frames
    scenes = {f"scene-{k:04d}": [S.make_nuscenes_frame(7000 + 10 * k + f, n_sweeps=3, pts_per_sweep=6000, n_inst=12,
                                                       mask_div=2, dense_masks=False) for f in range(3)] for k in range(3)}
    nusc, map_factory = SD.write_nuscenes(os.path.join(data, "nusc"), os.path.join(data, "nusc_masks"), scenes, ratio=0.32)
    m = load("src/nuscenes/2d_to_3d.py", "nusc_script")
    m.INPUT_PATH, m.INPUT_DIR, m.OUTPUT_DIR = os.path.join(data, "nusc"), os.path.join(data, "nusc_masks"), os.path.join(args.out, "nuscenes")
    m.ratio, m.BATCH_FRAMES = 0.32, 2
    m.main(nusc, map_factory, list(scenes))

    # KITTI: 5 frames
    frames = [S.make_kitti_frame(8100 + f, n_pts=30000, n_inst=10, mask_div=1, dense_masks=False) for f in range(5)]
    SD.write_kitti(os.path.join(data, "kitti"), os.path.join(data, "kitti_masks"), frames)
    k = load("src/kitti/2d_to_3d.py", "kitti_script")
    k.INPUT_PATH, k.INPUT_DIR = os.path.join(data, "kitti"), os.path.join(data, "kitti_masks")
    k.PRED_DIR, k.PSEUDO_DIR = os.path.join(args.out, "kitti", "pred"), os.path.join(args.out, "kitti", "pseudo")
    k.NUM_SAMPLES, k.BATCH_FRAMES = 5, 2
    k.main()

    # Waymo: 3 segments x 2 frames
    wscenes = []
    for s in range(3):
        fr = [S.make_waymo_frame(8200 + 10 * s + f, n_pts=30000, n_inst=16, mask_div=2) for f in range(2)]
        for x in fr:
            x.labels = [{"barrier": "car", "traffic_cone": "pedestrian"}.get(l, l) for l in x.labels]
        wscenes.append((f"segment-{s}", SD.waymo_frames(f"segment-{s}", os.path.join(data, "waymo_masks"), fr, ratio=(1024 / 1920) / 2)))
    w = load("src/waymo/2d_to_3d.py", "waymo_script")
    w.INPUT_DIR, w.OUTPUT_FILE = os.path.join(data, "waymo_masks"), os.path.join(args.out, "waymo", "pred.bin")
    w.ratio, w.BATCH_FRAMES = (1024 / 1920) / 2, 2
    w.main(wscenes, lambda f: f.points_vehicle)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
