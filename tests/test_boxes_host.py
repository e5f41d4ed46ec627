"""Host logic of pass 2 (no GPU): quaternion restatement, push-back, circle NMS, box assembly -
cm3d_b200.boxes / cm3d_b200.quat against the oracle restatement (oracle/ref_boxes.py) and scipy."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation as R

from cm3d_b200 import boxes as B
from cm3d_b200.quat import Quaternion
from oracle import ref_boxes as RB


def test_quaternion_rotation_matrix_matches_scipy():
    rng = np.random.default_rng(0)
    for _ in range(200):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        q *= 1 + rng.normal(0, 1e-9)                     # devkit records are only nearly unit
        m = Quaternion(list(q)).rotation_matrix
        ref = R.from_quat([q[1], q[2], q[3], q[0]]).as_matrix()
        assert np.allclose(m, ref, atol=1e-8)            # unnormalised input: pyquaternion renormalises
        qn = q / np.linalg.norm(q)
        assert np.allclose(Quaternion(list(qn)).rotation_matrix, ref, atol=1e-14)


def test_quaternion_from_matrix_round_trip_all_branches():
    rng = np.random.default_rng(1)
    for _ in range(300):
        m = R.random(random_state=int(rng.integers(1 << 30))).as_matrix()
        q = Quaternion(matrix=m)
        assert np.allclose(q.rotation_matrix, m, atol=1e-12)
        ref = RB.quat_wxyz_from_matrix(m)
        got = np.array(list(q))
        assert np.allclose(got, ref, atol=1e-12) or np.allclose(got, -ref, atol=1e-12)
    with pytest.raises(ValueError):
        Quaternion(matrix=np.diag([1.0, 1.0, -1.0]))    # det -1
    with pytest.raises(ValueError):
        Quaternion(matrix=np.ones((3, 3)))


def test_detection_names_and_priors():
    assert B.get_detection_name("trafficcone") == "traffic_cone"
    assert B.get_detection_name("constructionvehicle") == "construction_vehicle"
    assert B.get_detection_name("human") == "pedestrian"
    assert B.get_detection_name("car") == "car"
    assert B.get_detection_name("bus", B.KITTI_CLASS_MAPS) == "Tram"
    pri = {"car": [1.8, 4.5, 1.4], "bicycle": [0.6, 1.8, 1.4], "pedestrian": [0.4, 0.7, 1.7]}
    assert B.get_shape_prior(pri, "car") == [1.8, 4.5, 1.4]
    assert B.get_shape_prior(pri, "vehicle", waymo=True) == [1.8, 4.5, 1.4]
    assert B.get_shape_prior(pri, "cyclist", waymo=True) == [0.6, 1.8, 1.4]
    with pytest.raises(KeyError):
        B.get_shape_prior(pri, "vehicle")


def test_push_centroid_matches_oracle():
    rng = np.random.default_rng(2)
    for _ in range(300):
        yaw = np.float32(rng.uniform(-np.pi, np.pi))
        centroid = rng.uniform(-50, 50, 3).astype(np.float32) + np.array([1000, 900, 0], np.float32)
        av = list(rng.uniform(-1, 1, 3) + np.array([1000.0, 900.0, 0.0]))
        ext = [1.8, 4.5, 1.4]
        m = B.lane_align_matrix(yaw)
        got = B.push_centroid(centroid[:, None], ext, Quaternion(matrix=m), {"translation": av})
        ref = RB.push_centroid(centroid[:, None], ext, RB.quat_wxyz_from_matrix(m), np.asarray(av))
        assert np.allclose(got, ref, rtol=0, atol=1e-9)
        got_e = B.push_centroid(centroid - np.asarray(av, np.float32), ext, Quaternion(matrix=m), ego_frame=True)
        ref_e = RB.push_centroid(centroid - np.asarray(av, np.float32), ext, RB.quat_wxyz_from_matrix(m), ego_frame=True)
        assert np.allclose(got_e, ref_e, rtol=0, atol=1e-9)


def test_circle_nms_matches_oracle_and_properties():
    rng = np.random.default_rng(3)
    names = list(B.THRESHS_BY_LABEL)
    for trial in range(50):
        n = int(rng.integers(1, 60))
        dets = np.stack([rng.uniform(0, 20, n), rng.uniform(0, 20, n), rng.uniform(0, 1, n)], 1)
        labels = [names[int(k)] for k in rng.integers(0, len(names), n)]
        keep = B.circle_nms(dets, labels, B.THRESHS_BY_LABEL)
        assert [int(k) for k in keep] == [int(k) for k in RB.circle_nms(dets, labels, RB.THRESHS)]
        # idempotent: survivors do not suppress each other
        again = B.circle_nms(dets[keep], [labels[k] for k in keep], B.THRESHS_BY_LABEL)
        assert len(again) == len(keep)
        # every suppressed box is within its class radius of a kept, higher-scored box of its class
        for j in set(range(n)) - set(int(k) for k in keep):
            assert any(labels[k] == labels[j] and dets[k, 2] >= dets[j, 2] and
                       (dets[k, 0] - dets[j, 0]) ** 2 + (dets[k, 1] - dets[j, 1]) ** 2 <= B.THRESHS_BY_LABEL[labels[j]]
                       for k in keep)


def test_nuscenes_box_and_nms_format():
    pri = dict(RB.THRESHS)
    pri = {k: [1.0, 2.0, 1.5] for k in pri}
    pose = {"translation": [100.0, 200.0, 0.0]}
    b = B.nuscenes_box("tok", "human", 0.7, np.array([[110.0], [205.0], [1.0]], np.float32), np.float32(0.3), pri, pose)
    assert b["detection_name"] == "pedestrian" and b["rotation"] == [1.0, 0.0, 0.0, 0.0]
    assert b["translation"] == [110.0, 205.0, 1.0] and b["attribute_name"] == "pedestrian.standing"
    c = B.nuscenes_box("tok", "car", 0.9, np.array([110.0, 205.0, 1.0], np.float32), np.float32(0.3), pri, pose)
    assert c["detection_name"] == "car" and c["velocity"] == [0, 0] and len(c["rotation"]) == 4
    assert abs(2 * np.arctan2(c["rotation"][3], c["rotation"][0]) - 0.3) < 1e-6
    preds = {"meta": {"use_camera": True}, "results": {"tok": [c, dict(c, detection_score=0.5), b], "empty": []}}
    out = B.nms_predictions(preds)
    assert [x["detection_score"] for x in out["results"]["tok"]] == [0.9, 0.7]
    assert out["results"]["empty"] == []
    assert set(out["results"]["tok"][0]) == {"sample_token", "translation", "size", "rotation", "velocity",
                                              "detection_name", "detection_score", "attribute_name"}


def test_waymo_objects_wire_format_matches_protobuf_library():
    """The hand-written serialiser against google.protobuf on the same (restated) schema."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    from cm3d_b200 import waymo_proto as W
    fd = descriptor_pb2.FileDescriptorProto(name="m.proto", package="t", syntax="proto2")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, tname, label in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if tname:
                f.type_name = tname
    msg("Box", [(n, k, 1, None, 1) for n, k in W._BOX_FIELDS])
    msg("Label", [("box", 1, 11, ".t.Box", 1), ("type", 3, 5, None, 1), ("id", 4, 9, None, 1)])
    msg("Object", [("object", 1, 11, ".t.Label", 1), ("score", 2, 2, None, 1), ("context_name", 3, 9, None, 1),
                   ("frame_timestamp_micros", 4, 3, None, 1)])
    msg("Objects", [("objects", 1, 11, ".t.Object", 3)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    cls = message_factory.GetMessageClass(pool.FindMessageTypeByName("t.Objects"))
    rng = np.random.default_rng(5)
    objs = [dict(context_name=f"seg-{i}", frame_timestamp_micros=int(rng.integers(1, 1 << 52)), score=float(rng.uniform()),
                 type=int(rng.choice([1, 2, 4])), center_x=float(rng.normal(0, 50)), center_y=float(rng.normal(0, 50)),
                 center_z=float(rng.normal()), length=4.5, width=1.8, height=1.4, heading=float(rng.uniform(-3, 3)))
            for i in range(40)]
    data = W.serialize_objects(objs)
    m = cls()
    m.ParseFromString(data)
    assert m.SerializeToString() == data and len(m.objects) == 40
    back = W.parse_objects(data)
    for a, b, c in zip(objs, back, m.objects):
        assert a["context_name"] == b["context_name"] == c.context_name
        assert a["frame_timestamp_micros"] == b["frame_timestamp_micros"] == c.frame_timestamp_micros
        assert a["center_x"] == b["center_x"] == c.object.box.center_x and a["length"] == c.object.box.length
        assert a["width"] == c.object.box.width and a["heading"] == b["heading"] == c.object.box.heading
        assert b["score"] == c.score == np.float32(a["score"]) and a["type"] == b["type"] == c.object.type
        assert c.object.id == "unique object tracking ID"


def _members(frame):
    from oracle import c_oracle as CO
    o = CO.lift_frame_c(frame, record_pix=False, do_medoid=False)
    return o["n_points"], [len(x) for x in o["idx"]]


def test_stage_frame_specs_from_on_disk_datasets(tmp_path):
    """Host logic of the three scripts without a GPU: the FrameSpecs they build from the on-disk
    layouts (devkit-style tables + quaternions, KITTI calib files, Waymo calibration protos, masks as
    pycocotools strings in `{f}_masks.pkl`) see the same points in the same masks as the frames the
    datasets were written from."""
    from cm3d_b200 import kitti_stage, nuscenes_stage, waymo_stage
    from cm3d_b200 import synthetic as S
    from cm3d_b200 import synthetic_datasets as SD
    # nuScenes: quaternion round trip + `next` chain + KeyError at the end of the sweep chain
    fr = [S.make_nuscenes_frame(7100 + f, n_sweeps=2, pts_per_sweep=4000, n_inst=8, mask_div=2, dense_masks=False) for f in range(2)]
    nusc, mf = SD.write_nuscenes(str(tmp_path / "n"), str(tmp_path / "nm"), {"scene-0001": fr}, ratio=0.32)
    cfg = nuscenes_stage.make_cfg(INPUT_DIR=str(tmp_path / "nm"), ratio=0.32, n_sweeps=3)      # asks for 3, chain has 2
    scene = nusc.get("scene", nusc.field2token("scene", "name", "scene-0001")[0])
    sample = nusc.get("sample", scene["first_sample_token"])
    assert nuscenes_stage.count_frames(nusc, sample) == 2
    for f in range(2):
        masks, data = nuscenes_stage.load_frame_masks(str(tmp_path / "nm"), "scene-0001", f)
        assert isinstance(masks[0].counts, bytes) and data["cam_nums"] == [int(c) for c in fr[f].cam_nums]
        spec = nuscenes_stage.frame_spec(nusc, sample, masks, data, cfg)
        assert len(spec.sweeps) == 2 and spec.sweeps[0].shape[1] == 5 and spec.token == sample["token"]
        assert _members(spec) == _members(fr[f])
        if sample["next"]:
            sample = nusc.get("sample", sample["next"])
    _, lanes = nuscenes_stage.get_all_lane_points_in_scene(mf(nusc, scene))
    assert len(lanes) > 100 and len(lanes[0]) == 3
    # KITTI: calib text -> Calibration -> chains
    kf = [S.make_kitti_frame(8300, n_pts=8000, n_inst=6, mask_div=1, dense_masks=False)]
    SD.write_kitti(str(tmp_path / "k"), str(tmp_path / "km"), kf)
    kcfg = kitti_stage.make_cfg(INPUT_PATH=str(tmp_path / "k"), INPUT_DIR=str(tmp_path / "km"), num_samples=1)
    kitti = kitti_stage.kitti_object(str(tmp_path / "k"), "training", 1)
    assert len(kitti) == 1 and len(kitti_stage.kitti_object(str(tmp_path / "k"))) == 7481
    masks, data = kitti_stage.load_frame_masks(str(tmp_path / "km"), None, 0)
    assert "cam_nums" not in data
    assert _members(kitti_stage.frame_spec(kitti, 0, masks, data, kcfg)) == _members(kf[0])
    # Waymo: extrinsic . inv(axes) -> scipy quaternion -> pyquaternion matrix; fp64 intrinsics
    wf = [S.make_waymo_frame(8400, n_pts=8000, n_inst=10, mask_div=2)]
    frames = SD.waymo_frames("segment-x", str(tmp_path / "wm"), wf, ratio=(1024 / 1920) / 2)
    wcfg = waymo_stage.make_cfg(INPUT_DIR=str(tmp_path / "wm"), ratio=(1024 / 1920) / 2)
    masks, data = waymo_stage.load_frame_masks(str(tmp_path / "wm"), "segment-x", 0)
    spec = waymo_stage.frame_spec(frames[0], masks, data, wcfg, lambda f: f.points_vehicle)
    n0, m0 = _members(wf[0])
    n1, m1 = _members(spec)
    assert n0 == n1 and sum(abs(a - b) for a, b in zip(m0, m1)) <= 0.01 * sum(m0)     # matrices differ in the last bits
    lanes = waymo_stage.lanes_of_frame(frames[0])
    assert lanes.shape[1] == 3 and lanes[0, 2] == lanes[1, 2]                          # vertex 0 copies vertex 1's yaw
    g = waymo_stage.centroid_to_global(np.array([10.0, 2.0, 1.0, 1.0], np.float32), frames[0])
    assert np.allclose(g, RB.waymo_centroid_to_global([10.0, 2.0, 1.0], frames[0].pose.transform), atol=1e-3)


def test_kitti_label_line_format(tmp_path):
    from cm3d_b200 import kitti_stage
    p = str(tmp_path / "000007.txt")
    kitti_stage.save_pred(p, "Car", [0, 0, 0, 0], [1.4, 1.8, 4.5], [1.0, 2.5, 30.0], -1.5, 0.75)
    kitti_stage.save_pred(p, "Cyclist", [0, 0, 0, 0], [1.4, 0.6, 1.8], [1.0, 2.5, 30.0], 0.25, None)
    assert open(p).read() == ("Car -1 -1 -10 0 0 0 0 1.4 1.8 4.5 1.0 2.5 30.0 -1.5 0.75\n"
                              "Cyclist -1 -1 -10 0 0 0 0 1.4 0.6 1.8 1.0 2.5 30.0 0.25\n")
