"""Frame data model for the 2D-mask -> 3D lifting path.

A `FrameSpec` is everything the reference's per-frame body consumes once the
dataset accessors have run (reference: src/nuscenes/2d_to_3d.py:417-503,
src/kitti/2d_to_3d.py:990-1083, src/waymo/2d_to_3d.py:445-507): raw LiDAR
sweeps, the rigid-transform chain each sweep goes through, the per-camera
chain + intrinsics, and the instance masks with their camera numbers.

All matrices are fp32 *exactly as the reference casts them* (fp64 pose ->
`.to(dtype=torch.float32)`), so neither the oracle nor the CUDA kernels ever
see fp64.  A transform chain is a list of ops applied left to right:

    ("R", M[3,3])   points[:3] = M @ points[:3]          (pcd.py:166-172)
    ("T", t[3])     points[i] += t[i]                    (pcd.py:159-165)
    ("A", M[3,4])   p' = [p,1] @ M.T                     (kitti_utils.py:224-230)

Each output coordinate of R/A is the k-ordered FMA chain torch's CPU matmul
produces (r = a0*b0; r = fma(a1,b1,r); ...), see DESIGN.md "Numerics".
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

OP_END, OP_T, OP_R, OP_A = 0, 1, 2, 3
_KIND = {"T": OP_T, "R": OP_R, "A": OP_A}
MAX_CHAIN = 4          # longest chain in the reference: T,R,T,R (nuscenes:569-577)
OP_WORDS = 16          # int kind + 12 floats + 3 pad
CHAIN_WORDS = MAX_CHAIN * OP_WORDS

FOURTH_NONE, FOURTH_COL3, FOURTH_ONES = 0, 1, 2


def _f32(a, shape):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


def op_R(m) -> Tuple[str, np.ndarray]:
    return ("R", _f32(m, (3, 3)))


def op_T(t) -> Tuple[str, np.ndarray]:
    return ("T", _f32(np.asarray(t).reshape(-1), (3,)))


def op_A(m) -> Tuple[str, np.ndarray]:
    return ("A", _f32(m, (3, 4)))


def encode_chain(ops: Sequence[Tuple[str, np.ndarray]]) -> np.ndarray:
    """Pack a chain into CHAIN_WORDS 32-bit words (kind as int bits, floats raw)."""
    if len(ops) > MAX_CHAIN:
        raise ValueError(f"chain longer than {MAX_CHAIN} ops")
    out = np.zeros(CHAIN_WORDS, dtype=np.uint32)
    for k, (kind, m) in enumerate(ops):
        base = k * OP_WORDS
        out[base] = _KIND[kind]
        flat = np.asarray(m, dtype=np.float32).reshape(-1)
        out[base + 1: base + 1 + flat.size] = flat.view(np.uint32)
    return out


@dataclass
class CamSpec:
    """One camera of a frame: world->camera chain and the scaled intrinsic.

    K is the 3x3 *after* the reference's `K*ratio; K[2,2]=1`
    (nuscenes:585-587, kitti:1259-1266, waymo:586-593)."""
    ops: List[Tuple[str, np.ndarray]]
    K: np.ndarray

    def __post_init__(self):
        self.K = _f32(self.K, (3, 3))

    def viewpad34(self) -> np.ndarray:
        """Rows 0..2 of view_points' 4x4 `viewpad` (pcd.py:269-270)."""
        v = np.zeros((3, 4), dtype=np.float32)
        v[:, :3] = self.K
        return v


@dataclass
class RLEMask:
    """COCO run-length mask as written by gen_2d_masks_detic.py:468-471:
    `size=[W,H]`, column-major runs over a (W,H) array == row-major over the
    (H,W) image.  `counts` is either the pycocotools compressed bytes/str or an
    already-decoded uint32 run-length array (alternating 0-run, 1-run, ...)."""
    size: Tuple[int, int]
    counts: object


@dataclass
class FrameSpec:
    dataset: str                               # "nuscenes" | "kitti" | "waymo"
    sweeps: List[np.ndarray]                   # each (N_s, stride) fp32, C-contiguous
    sweep_ops: List[List[Tuple[str, np.ndarray]]]
    cams: List[CamSpec]
    cam_nums: np.ndarray                       # (I,) camera index of each instance
    masks: object                              # (I,H,W) uint8 dense, or list[RLEMask]
    labels: List[str] = field(default_factory=list)
    scores: List[float] = field(default_factory=list)
    fourth: int = FOURTH_COL3                  # what row 3 of aggr_pc_points holds
    close_thresh: Optional[float] = None       # fp32(sqrt(min_dist)) or None
    min_dist: float = 2.3                      # depth threshold (python float in the ref)
    token: str = ""
    meta: dict = field(default_factory=dict)
    floor_thresh: Optional[float] = None       # default-off extension: keep aggregated points with z > floor_thresh

    def __post_init__(self):
        self.sweeps = [np.ascontiguousarray(s, dtype=np.float32) for s in self.sweeps]
        for s in self.sweeps:
            if s.ndim != 2 or s.shape[1] not in (3, 4, 5):
                raise ValueError("sweep must be (N, 3|4|5) fp32")
        if len(self.sweeps) != len(self.sweep_ops):
            raise ValueError("one op chain per sweep")
        self.cam_nums = np.asarray(self.cam_nums, dtype=np.int32).reshape(-1)
        if self.n_instances and (self.cam_nums.min() < 0 or self.cam_nums.max() >= len(self.cams)):
            raise ValueError("cam_nums out of range")
        if self.fourth == FOURTH_COL3 and any(s.shape[1] < 4 for s in self.sweeps):
            raise ValueError("fourth=col3 needs >=4 columns")

    @property
    def n_instances(self) -> int:
        return int(self.cam_nums.shape[0])

    @property
    def n_raw_points(self) -> int:
        return int(sum(s.shape[0] for s in self.sweeps))

    @property
    def point_rows(self) -> int:
        return 3 if self.fourth == FOURTH_NONE else 4

    def mask_size(self, i: int) -> Tuple[int, int]:
        """(W, H) of instance i's mask."""
        if isinstance(self.masks, np.ndarray):
            return int(self.masks.shape[2]), int(self.masks.shape[1])
        w, h = self.masks[i].size
        return int(w), int(h)

    def min_dist_f32(self) -> np.float32:
        # `depths > min_dist` compares an fp32 tensor with a python float: torch
        # casts the scalar to fp32 (nuscenes:598).
        return np.float32(self.min_dist)


@dataclass
class LiftResult:
    """Per-frame output of the lifting path (what nuscenes:510-665 leaves behind)."""
    n_points: int                      # N after close-point removal
    seg_offsets: np.ndarray            # (I+1,) int32, instance i owns [off[i], off[i+1])
    seg_point_idx: Optional[np.ndarray]  # (sum M,) int32 ascending per instance (track_points)
    medoid_local: np.ndarray           # (I,) int32 index inside the instance, -1 if empty
    medoid_point_idx: np.ndarray       # (I,) int32 index into aggr_pc_points, -1 if empty
    centroids: np.ndarray              # (I,3) fp32 xyz of the medoid point (nan if empty)
    aggr_points: Optional[np.ndarray] = None   # (rows,N) fp32, only when requested
    pix: Optional[np.ndarray] = None           # (C,N) int32 packed fx|fy<<16 or -1, debug only
    yaw: Optional[np.ndarray] = None           # (I,) fp32 KITTI OBB yaw, nan if n/a
    obb: Optional[np.ndarray] = None           # (I,16) yaw, centre xyz, wlh, R' row-major (KITTI)
    box: Optional[np.ndarray] = None           # (I,8) orientation search: centre, extents (along, across, up), heading, area
    raw_counts: Optional[np.ndarray] = None    # (I,) member counts before the neighbour-count filter (when it ran)

    @property
    def counts(self) -> np.ndarray:
        return np.diff(self.seg_offsets)

    def instance_points(self, i: int) -> np.ndarray:
        return self.seg_point_idx[self.seg_offsets[i]: self.seg_offsets[i + 1]]


# --------------------------------------------------------------------------- (de)serialisation
def frame_to_arrays(frame: FrameSpec) -> dict:
    """Flatten a FrameSpec into numpy arrays (np.savez-able); masks become COCO runs."""
    from .rle import rle_counts_to_runs
    d = {
        "dataset": np.array(frame.dataset),
        "token": np.array(frame.token),
        "n_sweeps": np.array(len(frame.sweeps)),
        "fourth": np.array(frame.fourth),
        "close_thresh": np.array(np.nan if frame.close_thresh is None else frame.close_thresh, np.float64),
        "min_dist": np.array(frame.min_dist, np.float64),
        "cam_nums": frame.cam_nums,
        "labels": np.array(frame.labels),
        "scores": np.array(frame.scores, np.float64),
        "sweep_chains": np.stack([encode_chain(o) for o in frame.sweep_ops]) if frame.sweeps else np.zeros((0, CHAIN_WORDS), np.uint32),
        "cam_chains": np.stack([encode_chain(c.ops) for c in frame.cams]),
        "cam_K": np.stack([c.K for c in frame.cams]),
    }
    for s, raw in enumerate(frame.sweeps):
        d[f"sweep_{s}"] = raw
    if isinstance(frame.masks, np.ndarray):
        from .synthetic import dense_to_rle
        rles = dense_to_rle(frame.masks)
        d["masks_dense"] = np.array(1)
    else:
        rles = frame.masks
        d["masks_dense"] = np.array(0)
    runs = [rle_counts_to_runs(r.counts) for r in rles]
    d["mask_wh"] = np.array([r.size for r in rles], np.int32).reshape(-1, 2)
    d["mask_run_off"] = np.concatenate([[0], np.cumsum([len(r) for r in runs])]).astype(np.int64)
    d["mask_runs"] = np.concatenate(runs).astype(np.uint32) if runs else np.zeros(0, np.uint32)
    return d


def decode_chain(words: np.ndarray):
    ops = []
    for k in range(MAX_CHAIN):
        op = words[k * OP_WORDS:(k + 1) * OP_WORDS]
        kind = int(op[0])
        f = op[1:13].view(np.float32)
        if kind == OP_END:
            break
        if kind == OP_T:
            ops.append(("T", f[:3].copy()))
        elif kind == OP_R:
            ops.append(("R", f[:9].reshape(3, 3).copy()))
        else:
            ops.append(("A", f[:12].reshape(3, 4).copy()))
    return ops


def frame_from_arrays(d) -> FrameSpec:
    n_sweeps = int(d["n_sweeps"])
    sweeps = [np.asarray(d[f"sweep_{s}"]) for s in range(n_sweeps)]
    sweep_ops = [decode_chain(np.asarray(w, np.uint32)) for w in d["sweep_chains"]]
    cams = [CamSpec(decode_chain(np.asarray(w, np.uint32)), np.asarray(k))
            for w, k in zip(d["cam_chains"], d["cam_K"])]
    off = d["mask_run_off"]
    rles = [RLEMask((int(w), int(h)), np.asarray(d["mask_runs"][off[i]:off[i + 1]], np.uint32))
            for i, (w, h) in enumerate(d["mask_wh"])]
    if int(d["masks_dense"]):
        from .synthetic import rle_to_dense
        masks = np.stack([rle_to_dense(r) for r in rles]) if rles else np.zeros((0, 1, 1), np.uint8)
    else:
        masks = rles
    ct = float(d["close_thresh"])
    return FrameSpec(str(d["dataset"]), sweeps, sweep_ops, cams, np.asarray(d["cam_nums"]), masks,
                     [str(x) for x in d["labels"]], [float(x) for x in d["scores"]],
                     fourth=int(d["fourth"]), close_thresh=None if np.isnan(ct) else ct,
                     min_dist=float(d["min_dist"]), token=str(d["token"]))
