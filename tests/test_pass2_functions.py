"""The product's pass-2 host code, the GPU nearest-lane / medoid / hull-box kernels and the oracle
restatement, graded against tests/golden/pass2_functions.json: outputs of the reference's OWN functions
(`FunctionDef`s of /root/reference/src/<ds>/2d_to_3d.py exec'd unmodified by oracle/refrun/functions.py)."""
import ctypes
import json
import os

import numpy as np
import pytest

from cm3d_b200 import boxes as B
from cm3d_b200.quat import Quaternion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fix():
    with open(os.path.join(ROOT, "tests", "golden", "pass2_functions.json")) as f:
        return json.load(f)


def test_detection_names_and_priors_match_reference_functions(fix):
    from oracle import ref_boxes as RB
    pri = json.load(open(os.path.join(ROOT, "src/nuscenes/cfg/shape_priors_chatgpt.json")))
    for label, want in fix["get_detection_name"]["nuscenes"].items():
        assert B.get_detection_name(label) == want == RB.get_detection_name(label)
    assert fix["get_detection_name"]["waymo"] == fix["get_detection_name"]["nuscenes"]
    for label, want in fix["get_detection_name"]["kitti"].items():
        assert B.get_detection_name(label, B.KITTI_CLASS_MAPS) == want
    for ds in ("nuscenes", "kitti", "waymo"):
        for label, want in fix["get_shape_prior"]["chatgpt"][ds].items():
            assert B.get_shape_prior(pri, label, waymo=(ds == "waymo")) == want
    for label, want in fix["get_shape_prior"]["waymo_types"].items():
        assert B.get_shape_prior(pri, label, waymo=True) == want
    old = json.load(open("/root/reference/src/nuscenes/cfg/shape_priors.json")) if os.path.exists("/root/reference") else None
    if old is not None:
        for label, want in fix["get_shape_prior"]["not_chatgpt"].items():
            assert B.get_shape_prior(old, label, chatgpt=False) == want


def test_push_centroid_matches_reference_function(fix):
    from oracle import ref_boxes as RB
    for c in fix["push_centroid"]:
        m = B.lane_align_matrix(np.float32(c["lane_yaw"]))
        q = Quaternion(matrix=m)
        assert np.allclose(list(q), c["quat_wxyz"], rtol=0, atol=1e-15)
        cen = np.asarray(c["centroid_f32"], np.float32)
        av = c["av_translation"]
        got = B.push_centroid(cen[None], c["extents"], q, {"translation": av})
        assert np.array_equal(got, np.asarray(c["pushed_global"]), equal_nan=True)
        got_e = B.push_centroid(np.asarray(c["centroid_ego_f32"], np.float32), c["extents"], q, ego_frame=True)
        assert np.array_equal(got_e, np.asarray(c["pushed_ego_frame"]), equal_nan=True)
        with np.errstate(all="ignore"):
            ref = RB.push_centroid(cen[None], c["extents"], c["quat_wxyz"], np.asarray(av))
        assert np.array_equal(ref, np.asarray(c["pushed_global"]), equal_nan=True)


def test_circle_nms_matches_reference_function(fix):
    from oracle import ref_boxes as RB
    thr = fix["circle_nms"]["threshs_by_label"]
    assert thr == {k: v for k, v in B.THRESHS_BY_LABEL.items()}
    for c in fix["circle_nms"]["cases"]:
        dets = np.asarray(c["dets"], np.float64).reshape(-1, 3)
        if len(dets) == 0:
            continue
        assert [int(k) for k in B.circle_nms(dets, c["labels"], thr)] == c["keep"]
        assert [int(k) for k in RB.circle_nms(dets, c["labels"], thr)] == c["keep"]


def test_waymo_lane_yaws_match_reference_function(fix):
    from types import SimpleNamespace
    from cm3d_b200 import waymo_stage as W
    g = fix["get_yaws_from_lane_coords"]
    poly = [SimpleNamespace(x=x, y=y, z=0.0) for x, y in g["polyline_xy"]]
    assert np.array_equal(W.get_yaws_from_lane_coords(poly), np.asarray(g["out"]))
    assert np.array_equal(W.get_yaws_from_lane_coords(poly[:1]), np.asarray(g["single"]))


def test_oracle_closest_lane_and_medoid_match_reference_functions(fix):
    """The restatements the other parity tests lean on: ref_boxes' closest lane, the C oracle's medoid."""
    from oracle import c_oracle as CO
    from oracle import ref_boxes as RB
    g = fix["lane_yaws_distances_and_coords"]
    yaws, dist, coords, idx = RB.lane_yaws_distances_and_coords(np.asarray(g["centroids_f32"], np.float32), np.asarray(g["lane_pts"]))
    assert np.array_equal(yaws, np.asarray(g["yaws"], np.float32))
    assert np.array_equal(dist, np.asarray(g["distances"])) and np.array_equal(coords, np.asarray(g["coords"], np.float32))
    for c in fix["get_medoid"]:
        pts = np.asarray(c["points_f32"], np.float32).reshape(-1, 3)
        assert CO.medoid(np.ascontiguousarray(pts.T)) == c["medoid"], len(pts)


def test_oracle_open3d_box_matches_reference_get_depth_bbox(fix):
    from oracle import obb_oracle as O
    for c in fix["get_depth_bbox"]:
        pts = np.asarray(c["points_f32"], np.float32)
        center, wlh, Rb = O.get_depth_bbox(pts)
        assert np.allclose(center, c["center"], atol=1e-12) and np.allclose(wlh, c["wlh"], atol=1e-12)
        assert np.allclose(Rb, c["R"], atol=1e-12)
        assert abs(O.yaw_of(Rb) - c["yaw"]) < 1e-12
        assert c["line"].split()[14] == str(c["yaw"]) or abs(float(c["line"].split()[14]) - c["yaw"]) < 1e-15


# ------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_nearest_lane_kernel_matches_reference_function(fix):
    g = fix["lane_yaws_distances_and_coords"]
    yaws, dist, coords = B.lane_yaws_distances_and_coords(np.asarray(g["centroids_f32"], np.float32), np.asarray(g["lane_pts"]), "cuda:0")
    assert np.array_equal(yaws, np.asarray(g["yaws"], np.float32))
    assert np.array_equal(dist, np.asarray(g["distances"]))            # binary64, bit for bit
    assert np.array_equal(coords, np.asarray(g["coords"], np.float32))


def _segments_on_device(point_sets):
    """Instances given directly as gathered segments (SoA, stride seg_cap) for the per-instance kernels."""
    import torch
    m = [len(p) for p in point_sets]
    seg_off = np.concatenate([[0], np.cumsum(m)]).astype(np.int32)
    seg_cap = (int(seg_off[-1]) + 3) & ~3
    xyzw = np.zeros((4, seg_cap), np.float32)
    for p, o in zip(point_sets, seg_off[:-1]):
        xyzw[:3, o:o + len(p)] = np.asarray(p, np.float32).T
    return torch.from_numpy(xyzw.reshape(-1)).cuda(), torch.from_numpy(seg_off).cuda(), seg_cap, seg_off


@pytest.mark.gpu
def test_hull_box_kernel_matches_reference_get_depth_bbox(fix):
    """cm3d_hull_obb against the reference's own get_depth_bbox (open3d stubbed by the hull-vertex oracle):
    yaw 1e-3 rad, centre / extents 1e-3 m, and the hull vertex count equals Qhull's."""
    import torch
    from cm3d_b200 import _native as N
    from scipy.spatial import ConvexHull
    sets = [np.asarray(c["points_f32"], np.float32) for c in fix["get_depth_bbox"]]
    xyzw, seg_off, seg_cap, _ = _segments_on_device(sets)
    I = len(sets)
    obb = torch.empty(16 * I, dtype=torch.float32, device="cuda")
    info = torch.empty(I, dtype=torch.int32, device="cuda")
    err = torch.zeros(4, dtype=torch.int32, device="cuda")
    words = int(N.load().cm3d_hull_obb_ws_words(seg_cap))
    ws = torch.empty(words, dtype=torch.int32, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    N.call("cm3d_hull_obb", p(xyzw), seg_cap, p(seg_off), I, 4, 0, None, p(ws), words, p(obb), p(info), p(err),
           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    obb, info = obb.cpu().numpy().reshape(I, 16), info.cpu().numpy()
    for k, (c, pts) in enumerate(zip(fix["get_depth_bbox"], sets)):
        assert info[k] == len(ConvexHull(pts.astype(np.float64)).vertices), k
        d = abs(((float(obb[k, 0]) - c["yaw"] + np.pi) % (2 * np.pi)) - np.pi)
        assert d < 1e-3, (k, obb[k, 0], c["yaw"])
        assert np.allclose(obb[k, 1:4], c["center"], atol=1e-3) and np.allclose(obb[k, 4:7], c["wlh"], atol=1e-3)
        assert np.allclose(obb[k, 7:16].reshape(3, 3), c["R"], atol=1e-5)


@pytest.mark.gpu
def test_medoid_kernels_match_reference_get_medoid(fix):
    """cm3d_medoid (all-exact, and screen + verify forced down to 32-point instances and at its default
    threshold) on the point sets the reference's own get_medoid was run on (torch 2.11 CPU cdist)."""
    from test_gpu_parity import _medoid_abi
    sets = [np.ascontiguousarray(np.asarray(c["points_f32"], np.float32).reshape(-1, 3).T) for c in fix["get_medoid"]]
    want = [c["medoid"] for c in fix["get_medoid"]]
    for screen_min in (0, 32, 512):
        got, _, _ = _medoid_abi(sets, screen_min)
        assert got.tolist() == want, screen_min
