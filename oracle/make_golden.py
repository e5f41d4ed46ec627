"""ORACLE tooling: generate tests/golden/*.npz by running the reference's OWN L1
code (imported from /root/reference, never copied) inside the restated per-frame
body of oracle/ref_lift.py.

    python -m oracle.make_golden            # needs /root/reference; run in the build container

What comes from the reference unmodified:
  * src/nuscenes/utils/pcd.py  LidarPointCloud (from_file/rotate/translate), view_points
  * src/waymo/utils/pcd.py     same, Waymo copy (torch.cat variant)
  * src/kitti/kitti_utils.py   Calibration (read_calib_file, inverse_rigid_trans,
                               project_velo_to_ref/ref_to_velo/velo_to_rect), load_velo_scan
pcd.py imports matplotlib/pyquaternion/nuscenes at module top only for type
hints and dead code; three stub modules satisfy those imports.  torch runs with
1 thread; its version and CPU capability are recorded in each fixture.

Each fixture holds the frame INPUTS (so nothing is regenerated at test time) and
the reference-derived OUTPUTS: aggregated cloud, per-camera in-image point index
+ floored pixel, per-instance point index lists, medoid index, centroid.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]
    mod("matplotlib")
    mod("matplotlib.axes", Axes=type("Axes", (), {}))
    mod("pyquaternion", Quaternion=type("Quaternion", (), {}))
    mod("nuscenes")
    mod("nuscenes.utils")
    mod("nuscenes.utils.geometry_utils", view_points=None, transform_matrix=None)


def load_reference_module(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def pack_result(frame, res):
    from cm3d_b200.frames import frame_to_arrays
    d = {"in_" + k: v for k, v in frame_to_arrays(frame).items()}
    d["aggr"] = res["aggr"]
    d["n_points"] = np.array(res["n_points"])
    cnt = np.array([len(x) for x in res["idx"]], np.int64)
    d["seg_offsets"] = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    d["seg_point_idx"] = (np.concatenate(res["idx"]) if len(res["idx"]) else np.zeros(0)).astype(np.int32)
    d["medoid_local"] = res["medoid_local"].astype(np.int32)
    d["medoid_point_idx"] = res["medoid_point_idx"].astype(np.int32)
    d["centroids"] = res["centroids"]
    cams = sorted(res["pix"])
    d["pix_cams"] = np.array(cams, np.int32)
    for c in cams:
        idx, fx, fy = res["pix"][c]
        d[f"pix_{c}_idx"] = idx.astype(np.int32)
        d[f"pix_{c}_fx"] = fx.astype(np.int16)
        d[f"pix_{c}_fy"] = fy.astype(np.int16)
    d["made_with"] = np.array(f"torch {torch.__version__} cap={torch.backends.cpu.get_cpu_capability()} threads=1")
    return d


def edge_frame():
    """nuScenes-shaped frame with hand-made masks: empty, full image (exercises the
    floor==0 quirk and border pixels), a 1-px line (erodes to nothing), a 3x3 block
    (erodes to one pixel), and a mask touching the image border."""
    from cm3d_b200 import synthetic as S
    f = S.make_nuscenes_frame(777, n_sweeps=2, pts_per_sweep=6000, n_inst=8, mask_div=2)
    m = f.masks
    H, W = m.shape[1:]
    m[0] = 0
    m[1] = 1
    m[2] = 0
    m[2, H // 2, :] = 1
    m[3] = 0
    m[3, 100:103, 200:203] = 1
    m[4] = 0
    m[4, :40, :] = 1
    m[5] = 0
    m[5, :, :30] = 1
    f.cam_nums[:6] = [0, 0, 1, 2, 3, 4]
    return f


def main():
    torch.set_num_threads(1)
    _stub_modules()
    from cm3d_b200 import synthetic as S
    from oracle import ref_lift as RL
    pcd_nusc = load_reference_module("nuscenes/utils/pcd.py", "ref_pcd_nusc")
    pcd_waymo = load_reference_module("waymo/utils/pcd.py", "ref_pcd_waymo")
    kutils = load_reference_module("kitti/kitti_utils.py", "ref_kitti_utils")
    os.makedirs(OUT, exist_ok=True)

    cases = {
        "nusc_c1": (S.make_frame("c1", 0), pcd_nusc),
        "nusc_small": (S.make_nuscenes_frame(11, n_sweeps=3, pts_per_sweep=3000, n_inst=12, mask_div=2), pcd_nusc),
        "nusc_edge": (edge_frame(), pcd_nusc),
        "kitti_small": (S.make_kitti_frame(33, n_pts=12000, n_inst=15, mask_div=1), None),
        "waymo_small": (S.make_waymo_frame(44, n_pts=18000, n_inst=80, mask_div=2), pcd_waymo),
    }
    for name, (frame, pcd) in cases.items():
        if frame.dataset == "kitti":
            with tempfile.TemporaryDirectory() as td:
                cpath = os.path.join(td, "000000.txt")
                with open(cpath, "w") as fh:
                    fh.write(S.kitti_calib_text())
                calib = kutils.Calibration(cpath, device="cpu")
                vpath = os.path.join(td, "000000.bin")
                frame.sweeps[0].tofile(vpath)
                assert np.array_equal(kutils.load_velo_scan(vpath), frame.sweeps[0])
            ours = RL.kitti_calib_from_frame(frame)
            for a in ("V2C", "C2V", "R0"):
                assert torch.equal(getattr(calib, a), getattr(ours, a)), a
            res = RL.lift_frame(frame, pcd=pcd_nusc, calib=calib)
        else:
            if frame.dataset == "nuscenes":
                with tempfile.TemporaryDirectory() as td:
                    p = os.path.join(td, "s.pcd.bin")
                    frame.sweeps[0].tofile(p)
                    pts = pcd.LidarPointCloud.from_file(p, "cpu").points
                    assert torch.equal(pts, torch.from_numpy(frame.sweeps[0][:, :4].T.copy()))
            res = RL.lift_frame(frame, pcd=pcd)
        # the restated helpers must agree with the reference's on the same frame
        res2 = RL.lift_frame(frame)
        assert np.array_equal(res["aggr"].view(np.uint32), res2["aggr"].view(np.uint32)), name
        assert all(np.array_equal(a, b) for a, b in zip(res["idx"], res2["idx"])), name
        assert np.array_equal(res["medoid_local"], res2["medoid_local"]), name
        d = pack_result(frame, res)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        cnt = np.diff(d["seg_offsets"])
        print(f"{name}: N={int(d['n_points'])} I={len(cnt)} M min/max={cnt.min()}/{cnt.max()} "
              f"empty={int((cnt == 0).sum())} -> {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
