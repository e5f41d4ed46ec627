"""CPU: host-side logic - RLE codec, frame (de)serialisation, batch packing, sharding (gloo, 2 ranks)."""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rle_string_codec_round_trip():
    from cm3d_b200.rle import rle_string_to_runs, runs_to_rle_string
    rng = np.random.default_rng(0)
    for _ in range(50):
        runs = rng.integers(0, 5000, rng.integers(1, 60)).astype(np.uint32)
        s = runs_to_rle_string(runs)
        assert all(48 <= c < 112 for c in s)
        assert np.array_equal(rle_string_to_runs(s), runs)
    # known vector: pycocotools encodes a 4x4 mask with a 2x2 block of ones at rows 1-2, cols 1-2
    # (column-major) as counts [5,2,2,2,5] -> "52203"
    assert runs_to_rle_string([5, 2, 2, 2, 5]) == b"52203"
    assert rle_string_to_runs(b"52203").tolist() == [5, 2, 2, 2, 5]


def test_dense_rle_round_trip():
    from cm3d_b200.synthetic import dense_to_rle, rle_to_dense
    rng = np.random.default_rng(1)
    m = (rng.uniform(size=(3, 37, 53)) > 0.6).astype(np.uint8)
    m[1] = 1
    m[2] = 0
    for dense, rle in zip(m, dense_to_rle(m)):
        assert rle.size == (53, 37)
        assert np.array_equal(rle_to_dense(rle), dense)


def test_frame_arrays_round_trip():
    from cm3d_b200.frames import frame_from_arrays, frame_to_arrays
    frame, _ = load_golden("waymo_small")
    d = frame_to_arrays(frame)
    f2 = frame_from_arrays(d)
    d2 = frame_to_arrays(f2)
    assert sorted(d) == sorted(d2)
    for k in d:
        assert np.array_equal(d[k], d2[k], equal_nan=d[k].dtype.kind == "f"), k


def test_synthetic_frames_are_deterministic_and_shaped():
    from cm3d_b200 import synthetic as S
    a = S.make_frame("c1", 5, scale=0.1)
    b = S.make_frame("c1", 5, scale=0.1)
    assert all(np.array_equal(x, y) for x, y in zip(a.sweeps, b.sweeps))
    assert np.array_equal(a.masks, b.masks)
    assert a.n_instances == 20 and len(a.cams) == 6 and a.sweeps[0].shape[1] == 5
    k = S.make_frame("c3", 0, scale=0.05)
    assert k.dataset == "kitti" and len(k.cams) == 1 and k.masks.shape[1:] == (309, 1024)
    w = S.make_frame("c4", 0, scale=0.05)
    assert w.dataset == "waymo" and {m.size for m in w.masks} <= {(1024, 683), (1024, 473)}


def test_pack_frames_tables():
    from cm3d_b200 import batch as B
    frames = [load_golden(n)[0] for n in ("nusc_small", "kitti_small", "waymo_small")]
    pb = B.pack_frames(frames)
    fd = pb.table("frame_desc", B.FR_WORDS)
    sd = pb.table("sweep_desc", B.SW_WORDS)
    ts = pb.table("tile_sweep")
    assert pb.n_frames == 3 and pb.n_sweeps == 5 and pb.masks_kind == "rle"
    # tiles: contiguous per frame, ceil(npts/1024) per sweep, sweep bases consistent
    assert fd[0, 0] == 0 and np.array_equal(fd[1:, 0], fd[:-1, 1]) and fd[-1, 1] == pb.n_tiles
    t = 0
    for s in range(pb.n_sweeps):
        nt = -(-int(sd[s, 2]) // B.TILE)
        assert sd[s, 5] == t and (ts[t:t + nt] == s).all()
        o = int(np.uint32(sd[s, 0])) | (int(sd[s, 1]) << 32)
        assert o % 4 == 0                                   # 16-byte aligned sweep starts (bulk copy)
        t += nt
    # raw points: the columns the kernels read, verbatim (KITTI: x, y, z - its reflectance column is never read)
    o = int(np.uint32(sd[3, 0]))
    kxyz = frames[1].sweeps[0][:, :3]
    assert sd[3, 3] == 3 and np.array_equal(pb.raw[o:o + kxyz.size], kxyz.reshape(-1))
    o = int(np.uint32(sd[0, 0]))
    assert sd[0, 3] == 4 and np.array_equal(pb.raw[o:o + 4 * frames[0].sweeps[0].shape[0]], frames[0].sweeps[0][:, :4].reshape(-1))
    # instances: frame-major, vcam lists partition the frame's instances
    assert pb.n_inst == sum(f.n_instances for f in frames)
    vd = pb.table("vcam_desc", B.VC_WORDS)
    lst = pb.table("cam_inst_list")
    for f in range(3):
        v0, nv, i0, ni = fd[f, 2], fd[f, 3], fd[f, 4], fd[f, 5]
        got = []
        for v in range(v0, v0 + nv):
            got += lst[i0 + vd[v, 15]: i0 + vd[v, 15] + vd[v, 16]].tolist()
        assert sorted(got) == list(range(ni))
    assert fd[1, 10] == 4 and fd[0, 10] == 1               # KITTI skips M <= 3 (kitti:1479-1480)
    assert fd[0, 7] == 1 and fd[1, 7] == 0                 # close-point filter is nuScenes only
    assert pb.cnt_total == sum((fd[f, 1] - fd[f, 0]) * fd[f, 5] for f in range(3))


def test_pack_limits():
    from cm3d_b200 import batch as B
    from cm3d_b200 import synthetic as S
    f = S.make_nuscenes_frame(1, n_sweeps=1, pts_per_sweep=500, n_inst=2, mask_div=8)
    f.cam_nums = np.zeros(300, np.int32)
    f.masks = np.zeros((300,) + f.masks.shape[1:], np.uint8)
    with pytest.raises(ValueError, match="CM3D_ELIMIT"):
        B.pack_frames([f])


def test_shard_indices_and_merge():
    from cm3d_b200 import shard as SH
    for n in (0, 1, 7, 64, 28130):
        for world in (1, 2, 4, 8):
            for mode in ("interleaved", "blocked"):
                parts = [SH.shard_indices(n, r, world, mode) for r in range(world)]
                assert sorted(sum(parts, [])) == list(range(n))
                assert max(map(len, parts)) - min(map(len, parts)) <= 1
    merged = SH.merge_shards([{0: "a", 2: "c"}, {1: "b"}], 3)
    assert merged == ["a", "b", "c"]
    with pytest.raises(ValueError):
        SH.merge_shards([{0: "a"}, {0: "b"}], 1)
    with pytest.raises(ValueError):
        SH.merge_shards([{0: "a"}], 2)


_WORKER = r'''
import os, sys, pickle
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from cm3d_b200 import shard as SH, synthetic as S
from oracle import c_oracle as CO

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
N = 6
load = lambda i: S.make_nuscenes_frame(100 + i, n_sweeps=1, pts_per_sweep=1500, n_inst=4, mask_div=4)
def lift(frames):      # the oracle stands in for the GPU lifter on this CPU-only box
    return [CO.lift_frame_c(f, record_pix=False)["medoid_point_idx"].tolist() for f in frames]
local = SH.lift_sharded(N, load, lift, batch=2, rank=dist.get_rank(), world=2)
merged = SH.gather_labels(local, N)
if dist.get_rank() == 0:
    single = SH.lift_sharded(N, load, lift, batch=4, rank=0, world=1)
    assert merged == [single[i] for i in range(N)], (merged, single)
    open(sys.argv[2], "w").write("ok %d" % len(merged))
else:
    assert merged is None
dist.destroy_process_group()
'''


def test_sharded_lift_two_ranks_gloo(tmp_path):
    """world_size 2 over gloo: each rank lifts its share, rank 0 gathers and the merged labels
    equal the single-process result."""
    import socket
    import subprocess
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    out = tmp_path / "out.txt"
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(out)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    assert out.read_text() == "ok 6"


def test_cull_planes_never_drop_an_in_image_point():
    """The host-built frustum planes (cm3d_b200.batch.cull_planes, CM3D_VC_PLANES) are conservative:
    every point the oracle finds inside a camera image keeps a plane value above -S*2^-17 with half
    the margin to spare (fp64 evaluation of the fp32 coefficients), on all three datasets - and the
    planes do cull most (point, camera) pairs."""
    from cm3d_b200 import synthetic as S
    from cm3d_b200.batch import cull_planes
    from oracle import c_oracle as CO
    culled = pairs = 0
    for cfg, scale in (("c1", 0.5), ("c2", 0.1), ("c3", 0.25), ("c4", 0.2)):
        f = S.make_frame(cfg, 2, scale=scale, mask_div=2)
        aggr = CO.aggregate(f)                                   # (rows, N) reference cloud
        xyz = np.ascontiguousarray(aggr[:3], np.float32)
        first = f.cams[0].ops
        tref = np.asarray(first[0][1], np.float32) if first and first[0][0] == "T" else np.zeros(3, np.float32)
        q = (xyz + tref[:, None]).astype(np.float32).astype(np.float64)
        Sq = np.abs(q).sum(0)
        for c, cam in enumerate(f.cams):
            sizes = {f.mask_size(i) for i in range(f.n_instances) if f.cam_nums[i] == c}
            for (W, H) in sizes:
                planes, flags = cull_planes(cam, W, H, f.min_dist_f32(), tref)
                assert flags == 1 and not np.array_equal(planes[0], [0, 0, 0, 1])
                val = planes[:, :3].astype(np.float64) @ q + planes[:, 3:4].astype(np.float64)   # (5, N)
                pix = CO.project(xyz, cam, W, H, f.min_dist)
                inside = pix >= 0
                assert inside.any()
                assert (val[:, inside] > -0.5 * Sq[inside] * 2.0 ** -17).all(), (cfg, c)
                culled += int((val < -Sq * 2.0 ** -17).any(0).sum())
                pairs += xyz.shape[1]
    assert culled > 0.6 * pairs


def test_cull_planes_random_cameras_and_boundary_points():
    """Randomised: arbitrary chains (T-R-T-R, T-R, A-A-R, R-T, A), poses up to 3 km from the origin,
    points pushed onto the frustum faces.  No point the C oracle puts inside the image may be culled
    (fp64 evaluation of the fp32 planes must stay above -S*2^-17 with half the margin to spare)."""
    from scipy.spatial.transform import Rotation
    from cm3d_b200.batch import cull_planes
    from cm3d_b200.frames import CamSpec, op_A, op_R, op_T
    from oracle import c_oracle as CO
    rng = np.random.default_rng(77)
    n_inside = n_culled = n_pairs = 0
    for trial in range(60):
        kind = trial % 5
        Rw = Rotation.random(random_state=int(rng.integers(1 << 30))).as_matrix()
        Rc = Rotation.random(random_state=int(rng.integers(1 << 30))).as_matrix()
        tw = rng.uniform(-3000, 3000, 3) * (1 if kind in (0, 3) else 0.001)
        tc = rng.uniform(-2, 2, 3)
        if kind == 0:
            ops = [op_T(-tw), op_R(Rw.T), op_T(-tc), op_R(Rc.T)]
        elif kind == 1:
            ops = [op_T(-tc), op_R(Rc.T)]
        elif kind == 2:
            A1 = np.concatenate([Rw, tc[:, None]], 1)
            A2 = np.concatenate([Rw.T, (-Rw.T @ tc)[:, None]], 1)
            ops = [op_A(A2), op_A(A1), op_R(Rc)]
        elif kind == 3:
            ops = [op_R(Rw), op_T(tw * 0.01)]
        else:
            ops = [op_A(np.concatenate([Rc, tc[:, None]], 1))]
        W, H = int(rng.integers(200, 2000)), int(rng.integers(100, 1300))
        f = rng.uniform(0.3, 2.5) * W
        K = np.array([[f, 0, rng.uniform(0.3, 0.7) * W], [0, f * rng.uniform(0.9, 1.1), rng.uniform(0.3, 0.7) * H], [0, 0, 1]])
        cam = CamSpec(ops, K)
        # points: random in the camera frame incl. exactly-on-boundary rays, mapped back to the cloud frame in fp64
        n = 4000
        z = np.exp(rng.uniform(np.log(0.5), np.log(150.0), n))
        u = rng.uniform(-0.2 * W, 1.2 * W, n)
        v = rng.uniform(-0.2 * H, 1.2 * H, n)
        edge = rng.integers(0, 6, n)
        u = np.where(edge == 0, rng.uniform(0, 1e-3, n), np.where(edge == 1, W - 1 - rng.uniform(0, 1e-3, n), u))
        v = np.where(edge == 2, rng.uniform(0, 1e-3, n), np.where(edge == 3, H - 1 - rng.uniform(0, 1e-3, n), v))
        z = np.where(edge == 4, 2.3 + rng.uniform(0, 1e-4, n), z)
        pc = np.stack([(u - K[0, 2]) * z / K[0, 0], (v - K[1, 2]) * z / K[1, 1], z])
        M, c = np.eye(3), np.zeros(3)                       # p_cam = M p + c
        for kd, m in ops:
            m = np.asarray(m, np.float64)
            if kd == "T":
                c = c + m
            elif kd == "R":
                M, c = m @ M, m @ c
            else:
                M, c = m[:, :3] @ M, m[:, :3] @ c + m[:, 3]
        p = np.linalg.solve(M, pc - c[:, None]).astype(np.float32)
        tref = np.asarray(ops[0][1], np.float32) if ops[0][0] == "T" else np.zeros(3, np.float32)
        planes, flags = cull_planes(cam, W, H, np.float32(2.3), tref)
        assert flags == 1
        if kind == 0:
            # the camera that defines tref has NO residual first translation: its margin must not scale with the
            # kilometres of |tw| (tau counts |t_c - tref|, not |t_c + tref|); depth plane: d = (c2_z - min_dist) / 1.75
            c2 = c - M @ tref.astype(np.float64)
            margin = float(planes[0, 3]) - (c2[2] - float(np.float32(2.3))) / 1.75
            assert 0 < margin < 1e-3, (trial, margin)
        q = (p + tref[:, None]).astype(np.float32).astype(np.float64)
        Sq = np.abs(q).sum(0)
        val = planes[:, :3].astype(np.float64) @ q + planes[:, 3:4].astype(np.float64)
        inside = CO.project(p, cam, W, H, 2.3) >= 0
        n_inside += int(inside.sum())
        assert (val[:, inside] > -0.5 * Sq[inside] * 2.0 ** -17).all(), (trial, kind)
        n_culled += int((val < -Sq * 2.0 ** -17).any(0).sum())
        n_pairs += n
    assert n_inside > 20000 and n_culled > 0.2 * n_pairs


def test_cull_planes_many_equals_per_camera():
    """The per-frame (stacked) cull-plane computation gives the per-camera function's planes and flags,
    including the switched-off planes of a camera whose chain is not rigid."""
    from cm3d_b200 import synthetic as S
    from cm3d_b200.batch import cull_planes, cull_planes_many
    from cm3d_b200.frames import op_R
    for cfg in ("c1", "c4"):
        f = S.make_frame(cfg, 3, scale=0.05, mask_div=4)
        cams = list(f.cams.values()) if isinstance(f.cams, dict) else list(f.cams)
        first = cams[0].ops
        tref = np.asarray(first[0][1], np.float32) if first[0][0] == "T" else np.zeros(3, np.float32)
        # make one camera's rotation non-orthogonal: its planes must come back switched off
        bad = cams[1]
        k = [i for i, (kind, _) in enumerate(bad.ops) if kind == "R"][0]
        bad.ops[k] = op_R(np.asarray(bad.ops[k][1]) * 1.01)
        sizes = [(1024, 576 + 7 * i) for i in range(len(cams))]
        pm, fm = cull_planes_many(cams, sizes, f.min_dist_f32(), tref)
        for v, c in enumerate(cams):
            p1, f1 = cull_planes(c, sizes[v][0], sizes[v][1], f.min_dist_f32(), tref)
            assert f1 == fm[v]
            assert np.allclose(p1, pm[v], rtol=1e-6, atol=0)
        assert np.array_equal(pm[1], np.tile(np.array([0, 0, 0, 1], np.float32), (5, 1)))


def test_native_packer_equals_python_packer():
    """csrc/pack.cu (cm3d_pack_plan / cm3d_pack_fill through pack_frames_native) writes the buffers the Python
    packer writes: every descriptor table, the raw sweeps, the counts strings and the scalar geometry are
    identical; the fp64 cull planes agree to fp32 rounding (numpy's 3x3 products may associate differently)."""
    import dataclasses
    from cm3d_b200 import synthetic as S
    from cm3d_b200 import batch as B
    from cm3d_b200.synthetic import compress_rles, dense_to_rle
    frames = [S.make_frame("c1", 0, scale=0.1, mask_div=2), S.make_frame("c2", 1, scale=0.05, mask_div=2),
              S.make_frame("c3", 2, scale=0.1, mask_div=2), S.make_frame("c4", 3, scale=0.05, mask_div=2)]
    for f in frames:                      # the on-disk mask format: compressed counts strings
        if isinstance(f.masks, np.ndarray):
            f.masks = compress_rles(dense_to_rle(f.masks))
        elif not all(isinstance(m.counts, (bytes, str)) for m in f.masks):
            f.masks = compress_rles(f.masks)
    empty = dataclasses.replace(frames[0], sweeps=[np.zeros((0, 5), np.float32)] + frames[0].sweeps[1:])
    for batch, keep in ((frames, True), (frames[:1], True), ([frames[2]], True), ([empty, frames[1]], True),
                        (frames, False), ([empty, frames[1]], False)):
        a, b = B.pack_frames(batch, keep_fourth=keep), B.pack_frames_native(batch, keep_fourth=keep)
        assert b.masks_kind == "rle_str" == a.masks_kind
        # only the columns the kernels read are packed: 4 floats per point when column 3 becomes row 3, else 3
        want = sum((-(-s.size // s.shape[1] * (4 if (keep and f.fourth == 1) else 3) // 4)) * 4 for f in batch for s in f.sweeps if s.shape[0])
        assert a.raw.size == want + 4
        for name in ("n_frames", "n_sweeps", "n_tiles", "n_vcams", "n_inst", "n_chains", "max_inst_per_frame", "cnt_total",
                     "bits_words", "max_words", "n_raw_points", "max_runs", "grid_words", "max_cells", "any_kitti",
                     "frame_datasets", "frame_vcam_cams"):
            assert getattr(a, name) == getattr(b, name), name
        assert np.array_equal(a.frame_inst, b.frame_inst)
        assert np.array_equal(a.raw.view(np.uint32), b.raw.view(np.uint32))
        assert np.array_equal(a.mask, b.mask) and np.array_equal(a.mask_off, b.mask_off)
        for name, words in (("tile_sweep", 1), ("sweep_desc", B.SW_WORDS), ("frame_desc", B.FR_WORDS), ("cam_inst_list", 1),
                            ("inst_desc", B.IN_WORDS), ("chains", B.CHAIN_WORDS)):
            assert a.off[name + "_n"] == b.off[name + "_n"], name
            assert np.array_equal(a.table(name, words), b.table(name, words)), name
        va, vb = a.table("vcam_desc", B.VC_WORDS), b.table("vcam_desc", B.VC_WORDS)
        assert np.array_equal(va[:, :20], vb[:, :20]) and np.array_equal(va[:, 40:], vb[:, 40:])
        pa, pb_ = va[:, 20:40].copy().view(np.float32), vb[:, 20:40].copy().view(np.float32)
        assert np.allclose(pa, pb_, rtol=2e-6, atol=0)
        assert (pa != pb_).mean() < 0.2          # mostly bit-identical


def test_vectorised_box_assembly_equals_the_per_box_path():
    """`boxes.nuscenes_boxes` (one numpy pass per scene) against `boxes.nuscenes_box` (the reference's per-box
    loop, nuscenes:745-817) on every class, quadrant, the zero-yaw / pi-yaw branches of the trace method and a
    centroid at the ego origin (0/0 -> NaN offsets): identical JSON."""
    import json
    from cm3d_b200 import boxes as B
    from cm3d_b200.quat import Quaternion, quats_from_matrices
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sp = json.load(open(os.path.join(root, "src", "nuscenes", "cfg", "shape_priors_chatgpt.json")))
    rng = np.random.default_rng(11)
    k = 600
    labs = list(B.ATTRIBUTE_NAMES) + ["trafficcone", "human", "constructionvehicle"]
    labels = [labs[i] for i in rng.integers(0, len(labs), k)]
    labels[5] = "car"
    c = rng.normal(0, 30, (k, 3)).astype(np.float32) + np.float32([600, 1200, 0])
    pose = np.tile([600.0, 1200.0, 0.0], (k, 1)) + rng.normal(0, 1, (k, 3))
    c[5], pose[5] = np.float32([600, 1200, 0]), [600.0, 1200.0, 0.0]
    yaw = rng.uniform(-np.pi, np.pi, k).astype(np.float32)
    yaw[7], yaw[9], yaw[11] = 0, np.float32(np.pi), np.float32(-np.pi / 2)
    scores = rng.random(k).tolist()
    toks = [f"tok{i % 7}" for i in range(k)]
    got = B.nuscenes_boxes(toks, labels, scores, c, yaw, sp, pose)
    want = [B.nuscenes_box(toks[i], labels[i], scores[i], c[i], yaw[i], sp, {"translation": pose[i].tolist()}) for i in range(k)]
    assert [json.dumps(g) for g in got] == [json.dumps(w) for w in want]
    assert B.nuscenes_boxes([], [], [], np.zeros((0, 3), np.float32), np.zeros(0, np.float32), sp, np.zeros((0, 3))) == []
    from scipy.spatial.transform import Rotation
    mats = Rotation.random(200, random_state=3).as_matrix()          # all four branches of the trace method
    q = quats_from_matrices(mats)
    assert np.array_equal(q, np.stack([Quaternion(matrix=m).q for m in mats]))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores, no GPU): one JSON line with the arm's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "pseudo_label_frames_per_s" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"].startswith("C1:")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_vectorised_waymo_pass2_equals_the_per_object_path():
    """`waymo_stage.centroids_to_global` / `frame_objects` (one numpy pass per frame) against the per-instance
    functions that call torch / scipy exactly where the reference does (waymo:684-699, 803-858): the global
    centroids bit for bit, the objects within 1e-10 (a batched fp64 matrix product may order its terms
    differently)."""
    import json
    from types import SimpleNamespace as NS
    from scipy.spatial.transform import Rotation as R
    from cm3d_b200 import boxes as B
    from cm3d_b200 import waymo_stage as W
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sp = json.load(open(os.path.join(root, "src", "waymo", "cfg", "shape_priors_chatgpt.json")))
    rng = np.random.default_rng(5)
    for trial in range(3):
        pose = np.eye(4)
        pose[:3, :3] = R.from_euler("xyz", [0.01 * trial, -0.02, rng.uniform(-3, 3)]).as_matrix()
        pose[:3, 3] = [512.3 + 1000 * trial, -811.7, 12.5]
        frame = NS(pose=NS(transform=[float(v) for v in pose.reshape(-1)]), context=NS(name="seg"), timestamp_micros=123456 + trial)
        k = 300
        labs = [l for l in B.NUSC_TO_WAYMO if B.NUSC_TO_WAYMO[l]] + ["human"]
        labels = [labs[i] for i in rng.integers(0, len(labs), k)]
        local = np.column_stack([rng.normal(0, 30, k), rng.normal(0, 30, k), rng.normal(0, 1, k)]).astype(np.float32)
        local[3], local[4], local[5] = [0, 0, 0], [5, 0, 1], [0, -7, 1]
        cg = W.centroids_to_global(local, frame)
        one = np.array([W.centroid_to_global(c, frame) for c in local])
        assert np.array_equal(cg.view(np.uint32), one.view(np.uint32))
        yaw = rng.uniform(-np.pi, np.pi, k).astype(np.float32)
        scores = rng.random(k)
        got = W.frame_objects(frame, labels, scores, cg, yaw, sp)
        want = [W.object_of(frame, labels[i], scores[i], cg[i], yaw[i], sp) for i in range(k)]
        for a, b in zip(got, want):
            assert set(a) == set(b)
            for key in a:
                if isinstance(a[key], float):
                    assert (np.isnan(a[key]) and np.isnan(b[key])) or abs(a[key] - b[key]) <= 1e-10, key
                else:
                    assert a[key] == b[key], key
    assert W.frame_objects(frame, [], [], np.zeros((0, 3)), np.zeros(0, np.float32), sp) == []
    with pytest.raises(ValueError):
        W.frame_objects(frame, ["barrier"], [0.5], cg[:1], yaw[:1], sp)


def test_prefetch_map_keeps_order_bounds_lookahead_and_raises_in_place():
    """`lifter.prefetch_map` (the stages' reader threads): results in input order whatever the completion order, the
    input iterator is never drained further than the look-ahead, and an exception surfaces at its item's position."""
    import threading
    import time
    torch = pytest.importorskip("torch")
    from cm3d_b200.lifter import prefetch_map
    pulled = []

    def items():
        for k in range(40):
            pulled.append(k)
            yield k

    def work(k):
        time.sleep(0.002 * ((k * 7) % 5))
        if k == 25:
            raise KeyError("frame 25")
        return k * k

    got = []
    with pytest.raises(KeyError):
        for v in prefetch_map(work, items(), workers=4):
            got.append(v)
            assert len(pulled) <= len(got) + 2 * 4 + 1
    assert got == [k * k for k in range(25)]
    assert list(prefetch_map(lambda k: k + 1, range(5), workers=1)) == [1, 2, 3, 4, 5]
    assert list(prefetch_map(lambda k: threading.get_ident(), range(3), workers=1)) == [threading.get_ident()] * 3


def test_native_packer_column_selection_every_stride_and_tail():
    """The C packer's column selection (whole-vector loads + shuffles, four rows at a time) against plain slicing
    for 5-, 4- and 3-column sweeps of 0..13 rows, with and without the 4th column, including values whose bit
    patterns a float copy must not touch (NaN payloads, -0.0, denormals)."""
    import dataclasses
    from cm3d_b200 import batch as B
    from cm3d_b200 import synthetic as S
    from cm3d_b200.synthetic import compress_rles, dense_to_rle
    rng = np.random.default_rng(3)
    base = {5: S.make_frame("c2", 1, scale=0.05, mask_div=2), 4: S.make_frame("c3", 2, scale=0.1, mask_div=2),
            3: S.make_frame("c4", 3, scale=0.05, mask_div=2)}
    for cols, f in base.items():
        if isinstance(f.masks, np.ndarray):
            f.masks = compress_rles(dense_to_rle(f.masks))
        elif not all(isinstance(m.counts, (bytes, str)) for m in f.masks):
            f.masks = compress_rles(f.masks)
        assert f.sweeps[0].shape[1] == cols
        for n in range(14):
            bits = rng.integers(0, 2 ** 32, (n, cols), dtype=np.uint64).astype(np.uint32)
            if n > 2:
                bits[1, 0], bits[2, 1], bits[0, 2] = 0x7FC12345, 0x80000000, 0x00000001
            sweep = bits.view(np.float32)
            g = dataclasses.replace(f, sweeps=[sweep] + [np.zeros((0, cols), np.float32)] * (len(f.sweeps) - 1))
            for keep in (True, False):
                pb = B.pack_frames_native([g], keep_fourth=keep)
                stride = 4 if (keep and g.fourth == 1 and cols >= 4) else 3
                got = pb.raw.view(np.uint32)[:n * stride].reshape(n, stride)
                assert np.array_equal(got, bits[:, :stride]), (cols, n, keep)
                assert not pb.raw.view(np.uint32)[n * stride:(n * stride + 3) // 4 * 4].any()      # zero padding to 16 bytes
