"""Waymo 2D-mask -> 3D pseudo-label lifting (the reference's `2d_to_3d.py` / README `2d_to_3d_new.py`),
B200-native.  Run like the reference: `cd src/waymo && python 2d_to_3d.py` after editing the
variables below (names of src/waymo/2d_to_3d.py:47-67,352-372 of the reference).  Inputs: Waymo
Open Dataset TFRecords under INPUT_PATH (parsed with tensorflow + waymo_open_dataset, like the
reference) and `{INPUT_DIR}/{scene}/{f}_masks.pkl` + `{f}_data.json`; output: a serialised
`metrics_pb2.Objects` file.  CUDA only (cm3d_b200, sm_100a): no CPU fallback.
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")))

import torch  # noqa: E402

INPUT_PATH = "../../data/waymo/training/"
ATTRIBUTE_NAMES = {
    "barrier": "",
    "traffic_cone": "",
    "bicycle": "cycle.without_rider",
    "motorcycle": "cycle.without_rider",
    "pedestrian": "pedestrian.standing",
    "car": "vehicle.stopped",
    "bus": "vehicle.stopped",
    "construction_vehicle": "vehicle.stopped",
    "trailer": "vehicle.stopped",
    "truck": "vehicle.stopped",
}
OUTPUT_DIR = "../../outputs/waymo/"
INPUT_DIR = "../../mask_outputs/waymo-detic/"
DEVICE = "cuda:0" if torch.cuda.is_available() else "cpu"       # the reference hard-codes cuda:1
RGB_Name = ["FRONT_IMAGE", "FRONT_LEFT_IMAGE", "FRONT_RIGHT_IMAGE", "SIDE_LEFT_IMAGE", "SIDE_RIGHT_IMAGE"]
CAM_LIST = ["FRONT", "FRONT_LEFT", "FRONT_RIGHT", "SIDE_LEFT", "SIDE_RIGHT"]
Lidar_Name = ["TOP", "FRONT", "SIDE_LEFT", "SIDE_RIGHT", "REAR"]

# literals of the reference's __main__ (src/waymo/2d_to_3d.py:400,407,431,523,1300)
min_dist = 2.3
floor_thresh = -0.6           # unused by the reference as shipped
ratio = 1024 / 1920
SCENE_SLICE = (680, 710)      # scene_list[680:710]
OUTPUT_FILE = "../../outputs/waymo/pred_0307_detic_train_680_710.bin"
BATCH_FRAMES = 32
READER_THREADS = 8            # threads that read scans / masks ahead of the GPU (not in the reference)


def _tfrecord_scenes(scene_names):
    import tensorflow as tf
    from waymo_open_dataset import dataset_pb2

    def frames_of(scene_name):
        for frame_data in tf.data.TFRecordDataset(INPUT_PATH + scene_name, compression_type=""):
            frame = dataset_pb2.Frame()
            frame.ParseFromString(bytearray(frame_data.numpy()))
            yield frame
    for scene_name in scene_names:
        yield scene_name, frames_of(scene_name)


def main(scenes=None, points_fn=None, lifter=None):
    from cm3d_b200 import waymo_stage as stage
    if DEVICE == "cpu":
        raise RuntimeError("cm3d_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    cfg = stage.make_cfg(INPUT_PATH=INPUT_PATH, OUTPUT_DIR=OUTPUT_DIR, INPUT_DIR=INPUT_DIR, ATTRIBUTE_NAMES=ATTRIBUTE_NAMES,
                         DEVICE=DEVICE, CAM_LIST=CAM_LIST, min_dist=min_dist, floor_thresh=floor_thresh, ratio=ratio,
                         scene_slice=SCENE_SLICE, output_path=OUTPUT_FILE, batch_frames=BATCH_FRAMES, reader_threads=READER_THREADS,
                         shape_priors_path=os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg",
                                                        "shape_priors_chatgpt.json"))
    if scenes is None:
        scene_list = sorted(os.listdir(INPUT_PATH))
        print(len(scene_list))
        scenes = _tfrecord_scenes(scene_list[SCENE_SLICE[0]:SCENE_SLICE[1]])
    return stage.run(cfg, scenes, points_fn, lifter=lifter)


if __name__ == "__main__":
    main()
