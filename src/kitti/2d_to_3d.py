"""KITTI 2D-mask -> 3D pseudo-label lifting (the reference's `2d_to_3d.py` / README `2d_to_3d_new.py`),
B200-native.  Run like the reference: `cd src/kitti && python 2d_to_3d.py` after editing the
variables below (names and defaults of src/kitti/2d_to_3d.py:77-126 of the reference).
Inputs: KITTI object folders under INPUT_PATH (training/velodyne, training/calib) and
`{INPUT_DIR}/{f}_masks.pkl` + `{f}_data.json`; outputs: KITTI label files `{f:06}.txt` in PRED_DIR
(with score) and PSEUDO_DIR (without).  CUDA only (cm3d_b200, sm_100a): no CPU fallback.
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")))

import torch  # noqa: E402

VER_NAME = "v1.0-trainval"
INPUT_PATH = "/data2/mehark/kitti/"

KITTI_CLASS_MAPS = {
    "car": "Car",
    "pedestrian": "Pedestrian",
    "truck": "Truck",
    "bus": "Tram",
    "traffic_cone": "Misc",
    "construction_vehicle": "Misc",
    "bicycle": "Cyclist",
    "motorcycle": "Cyclist",
    "trailer": "Misc",
    "barrier": "Misc",
}

OUTPUT_DIR = "../../outputs/kitti/"
PRED_DIR = "/data2/mehark/kitti/training/pred/"
PSEUDO_DIR = "/data2/mehark/kitti/training/pseudo/"
INPUT_DIR = "/data2/mehark/zs3d_outputs/kitti_detic_wo_2d_nms/"
DEVICE = "cuda:0" if torch.cuda.is_available() else "cpu"     # the reference forces "cpu"; this build is CUDA only

# literals of the reference's __main__ (src/kitti/2d_to_3d.py:903,909,993-994)
min_dist = 2.3
floor_thresh = 0.6            # unused by the reference as shipped
ratio = 0.8366
SPLIT = "training"
NUM_SAMPLES = None            # None = the reference's hard-coded 7481 (training) / 7518 (testing)
BATCH_FRAMES = 64
READER_THREADS = 8            # threads that read scans / masks ahead of the GPU (not in the reference)


def main(kitti=None, frame_range=None, lifter=None):
    from cm3d_b200 import kitti_stage as stage
    if DEVICE == "cpu":
        raise RuntimeError("cm3d_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    cfg = stage.make_cfg(INPUT_PATH=INPUT_PATH, OUTPUT_DIR=OUTPUT_DIR, PRED_DIR=PRED_DIR, PSEUDO_DIR=PSEUDO_DIR,
                         INPUT_DIR=INPUT_DIR, KITTI_CLASS_MAPS=KITTI_CLASS_MAPS, DEVICE=DEVICE, min_dist=min_dist,
                         floor_thresh=floor_thresh, ratio=ratio, split=SPLIT, num_samples=NUM_SAMPLES,
                         batch_frames=BATCH_FRAMES, reader_threads=READER_THREADS,
                         shape_priors_path=os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg",
                                                        "shape_priors_chatgpt.json"))
    return stage.run(cfg, kitti, frame_range, lifter=lifter)


if __name__ == "__main__":
    main()
