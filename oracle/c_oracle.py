"""ORACLE (test infrastructure): ctypes front end of oracle/lift_oracle.c.

`lift_frame_c(frame)` returns the same dict layout as oracle.ref_lift.lift_frame,
computed with fixed IEEE binary32 arithmetic (see the header of lift_oracle.c).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from oracle.coco_rle import counts_to_runs

FOURTH_NONE = 0
FrameSpec = object      # duck-typed (see oracle/ref_lift.py)
_OP_KIND = {"T": 1, "R": 2, "A": 3}     # op codes of lift_oracle.c: 16 words per op = kind, 12 floats, 3 pad


def encode_chain(ops):
    """A transform chain [("R"|"T"|"A", matrix), ...] as the 4 x 16 words lift_oracle.c reads."""
    if len(ops) > 4:
        raise ValueError("chain longer than 4 ops")
    out = np.zeros(64, dtype=np.uint32)
    for k, (kind, m) in enumerate(ops):
        out[16 * k] = _OP_KIND[kind]
        flat = np.ascontiguousarray(m, dtype=np.float32).reshape(-1)
        out[16 * k + 1: 16 * k + 1 + flat.size] = flat.view(np.uint32)
    return out


def rle_counts_to_runs(counts):
    return np.asarray(counts_to_runs(counts), dtype=np.uint32)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblift_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "lift_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_aggregate_sweep.restype = ctypes.c_long
        L.oracle_membership.restype = ctypes.c_long
        L.oracle_medoid.restype = ctypes.c_long
        L.oracle_medoid_mt.restype = ctypes.c_long
        L.oracle_rle_decode.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


_F, _U32, _I32, _U8 = ctypes.c_float, ctypes.c_uint32, ctypes.c_int32, ctypes.c_uint8


def aggregate(frame: FrameSpec):
    """-> (rows, N) fp32 aggr_pc_points (nuscenes:433-465 / kitti:1066-1077 / waymo:472-486)."""
    L = lib()
    n_raw = frame.n_raw_points
    rows = frame.point_rows
    out = np.empty((4, max(n_raw, 1)), np.float32)
    k = 0
    for raw, ops in zip(frame.sweeps, frame.sweep_ops):
        chain = encode_chain(ops)
        use_close = frame.close_thresh is not None
        thr = np.float32(frame.close_thresh if use_close else 0.0)
        kept = L.oracle_aggregate_sweep(
            _p(raw, _F), ctypes.c_long(raw.shape[0]), ctypes.c_int(raw.shape[1]), _p(chain, _U32),
            ctypes.c_int(frame.fourth), ctypes.c_int(int(use_close)), ctypes.c_float(thr),
            _p(out[0, k:], _F), _p(out[1, k:], _F), _p(out[2, k:], _F), _p(out[3, k:], _F))
        k += kept
    return np.ascontiguousarray(out[:rows, :k])


def project(aggr_xyz, cam, W, H, min_dist):
    L = lib()
    x, y, z = (np.ascontiguousarray(aggr_xyz[i]) for i in range(3))
    n = x.shape[0]
    pix = np.empty(n, np.int32)
    chain = encode_chain(cam.ops)
    vp = cam.viewpad34()
    L.oracle_project(_p(x, _F), _p(y, _F), _p(z, _F), ctypes.c_long(n), _p(chain, _U32), _p(vp, _F),
                     ctypes.c_float(np.float32(min_dist)), ctypes.c_int(W), ctypes.c_int(H), _p(pix, _I32))
    return pix


def decode_rle(rle) -> np.ndarray:
    L = lib()
    W, H = rle.size
    runs = rle_counts_to_runs(rle.counts)
    out = np.empty((H, W), np.uint8)
    rc = L.oracle_rle_decode(_p(runs, _U32), ctypes.c_long(len(runs)), ctypes.c_int(W), ctypes.c_int(H), _p(out, _U8))
    if rc != 0:
        raise ValueError("RLE does not cover the mask")
    return out


def erode(mask_hw: np.ndarray) -> np.ndarray:
    L = lib()
    m = np.ascontiguousarray(mask_hw, np.uint8)
    out = np.empty_like(m)
    L.oracle_erode3x3(_p(m, _U8), ctypes.c_int(m.shape[1]), ctypes.c_int(m.shape[0]), _p(out, _U8))
    return out


def membership(pix, eroded_hw):
    L = lib()
    idx = np.empty(pix.shape[0], np.int32)
    k = L.oracle_membership(_p(pix, _I32), ctypes.c_long(pix.shape[0]), _p(eroded_hw, _U8),
                            ctypes.c_int(eroded_hw.shape[1]), ctypes.c_int(eroded_hw.shape[0]), _p(idx, _I32))
    return idx[:k].copy()


def medoid(xyz_3m, want_sums=False):
    L = lib()
    x, y, z = (np.ascontiguousarray(xyz_3m[i], np.float32) for i in range(3))
    m = x.shape[0]
    sums = np.empty(m, np.float32) if want_sums else None
    j = L.oracle_medoid_mt(_p(x, _F), _p(y, _F), _p(z, _F), ctypes.c_long(m),
                           _p(sums, _F) if want_sums else None, ctypes.c_int(os.cpu_count() or 1))
    return (int(j), sums) if want_sums else int(j)


def sqdist(xyz_3m):
    """m x m matrix of the matmul formula's squared distances before the clamp (row i, column j)."""
    L = lib()
    x, y, z = (np.ascontiguousarray(xyz_3m[i], np.float32) for i in range(3))
    m = x.shape[0]
    out = np.empty((m, m), np.float32)
    L.oracle_sqdist(_p(x, _F), _p(y, _F), _p(z, _F), ctypes.c_long(m), _p(out, _F))
    return out


def lift_frame_c(frame: FrameSpec, record_pix=True, do_medoid=True):
    aggr = aggregate(frame)
    n = aggr.shape[1]
    I = frame.n_instances
    res = {
        "aggr": aggr if frame.fourth != FOURTH_NONE else np.ascontiguousarray(aggr.T),
        "n_points": n,
        "idx": [np.zeros(0, np.int64) for _ in range(I)],
        "medoid_local": np.full(I, -1, np.int64),
        "medoid_point_idx": np.full(I, -1, np.int64),
        "centroids": np.full((I, 3), np.nan, np.float32),
        "pix": {},
    }
    pix_cache = {}
    for i in range(I):
        c = int(frame.cam_nums[i])
        W, H = frame.mask_size(i)
        key = (c, W, H)
        if key not in pix_cache:
            pix_cache[key] = project(aggr[:3], frame.cams[c], W, H, frame.min_dist)
            if record_pix and c not in res["pix"]:
                p = pix_cache[key]
                sel = np.flatnonzero(p >= 0)
                res["pix"][c] = (sel.astype(np.int64), (p[sel] & 0xFFFF).astype(np.int64), (p[sel] >> 16).astype(np.int64))
        dense = frame.masks[i] if isinstance(frame.masks, np.ndarray) else decode_rle(frame.masks[i])
        idx = membership(pix_cache[key], erode(dense))
        res["idx"][i] = idx.astype(np.int64)
        min_pts = 4 if frame.dataset == "kitti" else 1            # kitti:1479-1480 skips M<=3
        if do_medoid and idx.size >= min_pts:
            m = medoid(aggr[:3, idx])
            res["medoid_local"][i] = m
            res["medoid_point_idx"][i] = idx[m]
            res["centroids"][i] = aggr[:3, idx[m]]
    return res
