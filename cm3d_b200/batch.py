"""Host-side packing of FrameSpecs into the flat buffers the C-ABI consumes.

A `PackedBatch` is three host buffers (pinned when CUDA is present) that go to
the GPU with one copy each:

    raw    float32   every sweep's raw points back to back (16-byte aligned starts)
    meta   int32     all descriptor tables of include/cm3d_b200.h, back to back
    mask   uint8     dense (H,W) uint8 masks, or COCO run lengths (uint32) + offsets

plus the scalar geometry (tile / instance / word counts) that sizes the device
workspace.  Nothing here touches point coordinates: the arithmetic of the
reference (src/nuscenes/2d_to_3d.py:433-665) happens in the CUDA kernels.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np

from .frames import CHAIN_WORDS, FOURTH_COL3, FOURTH_NONE, FrameSpec, RLEMask, encode_chain
from .rle import rle_counts_to_runs

TILE = 1024
MAX_INST = 254
MAX_VCAMS = 16
MEDOID_COLS = 256
SCREEN_MIN_PTS = 512          # CM3D_SCREEN_MIN_PTS: medoid instances this large are screened, then verified
CELL = 32
SW_WORDS, FR_WORDS, VC_WORDS, IN_WORDS, ERR_WORDS = 8, 20, 44, 8, 4


def _split64(v: int):
    """int64 -> (lo, hi) as signed int32 words (device side: join64)."""
    lo, hi = v & 0xFFFFFFFF, (v >> 32) & 0xFFFFFFFF
    return (lo - (1 << 32) if lo >= (1 << 31) else lo), (hi - (1 << 32) if hi >= (1 << 31) else hi)


def _f32_bits(x) -> int:
    return int(np.float32(x).view(np.int32))


_I3 = np.eye(3)
_OFF_PLANES = np.tile(np.array([0.0, 0.0, 0.0, 1.0], np.float32), (5, 1))


def _compose(ops):
    """fp64 composition of a transform chain: p_out = M p + c; also sum of |translations|_1 after
    the first op and whether every linear part is orthogonal to 1e-3."""
    M, c, tau, ok = _I3, np.zeros(3), 0.0, True
    for k, (kind, m) in enumerate(ops):
        m = np.asarray(m, np.float64)
        if kind == "T":
            c = c + m
            if k:
                tau += abs(float(m[0])) + abs(float(m[1])) + abs(float(m[2]))
        else:
            L = m if kind == "R" else m[:, :3]
            ok = ok and bool(np.abs(L @ L.T - _I3).max() < 1e-3)
            M, c = L @ M, L @ c
            if kind == "A":
                c = c + m[:, 3]
                tau += abs(float(m[0, 3])) + abs(float(m[1, 3])) + abs(float(m[2, 3]))
    return M, c, tau, ok


def cull_planes(cam, W: int, H: int, min_dist32, tref):
    """Conservative frustum planes of one vcam in q = p + tref coordinates (include/cm3d_b200.h,
    CM3D_VC_PLANES).  Error budget (u = 2^-24, S = |q|_1, tau = |translations|_1): the reference's
    fp32 chain is within 96u(S+tau) of the exact composition per coordinate, its pixel numerators
    within 4u of theirs, and evaluating a plane scaled to |n|_1 <= 1 in fp32 costs <= 6uS + 4u|d|;
    all of it is below S*2^-17 + mg0, and mg0 is folded into d.  Returns (planes[5,4] f32, flags).
    (The five planes are evaluated as one 5x3 product: this runs once per camera and frame on the host.)"""
    K = np.asarray(cam.K, np.float64)
    k00, k01, k02 = float(K[0, 0]), float(K[0, 1]), float(K[0, 2])
    k10, k11, k12 = float(K[1, 0]), float(K[1, 1]), float(K[1, 2])
    bottom_ok = K[2, 0] == 0 and K[2, 1] == 0 and K[2, 2] == 1
    simple = bool(k01 == 0 and k10 == 0 and k00 != 0 and k11 != 0 and bottom_ok)
    flags = 1 if simple else 0
    M, c, tau, ok = _compose(cam.ops)
    if not ok or not bottom_ok or not np.all(np.isfinite(M)):
        return _OFF_PLANES.copy(), flags
    tref = np.asarray(tref, np.float64)
    first_T = len(cam.ops) and cam.ops[0][0] == "T"
    # q = p + tref, so a camera whose first op is T(t_c) sees q + (t_c - tref): the residual translation
    t0 = np.asarray(cam.ops[0][1], np.float64) - tref if first_T else tref
    tau += abs(float(t0[0])) + abs(float(t0[1])) + abs(float(t0[2]))
    c2 = c - M @ tref                                   # p = q - tref
    abc = np.array([[0.0, 0.0, 1.0],                                  # z > min_dist
                    [k00, k01, k02], [-k00, -k01, W - k02],           # 0 < r0, r0 < W z
                    [k10, k11, k12], [-k10, -k11, H - k12]])          # 0 < r1, r1 < H z
    d0 = np.array([-float(min_dist32), 0.0, 0.0, 0.0, 0.0])
    kappa = 1.75 * np.abs(abc).sum(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        n = (abc @ M) / kappa[:, None]
        d = (abc @ c2 + d0) / kappa
    if not bool((np.abs(n).sum(1) <= 1.0).all()) or not bool(np.isfinite(d).all()):
        return _OFF_PLANES.copy(), flags
    out = np.empty((5, 4), np.float64)
    out[:, :3], out[:, 3] = n, d
    mg0 = 2.0 ** -17 * tau + 2.0 ** -20 * float(np.abs(d).max()) + 1e-30
    out[:, 3] += mg0
    out32 = out.astype(np.float32)
    out32[:, 3] = np.nextafter(out32[:, 3], np.float32(np.inf))   # rounding of d never tightens a plane
    return out32, flags


def cull_planes_many(cams, sizes, min_dist32, tref):
    """cull_planes for all vcams of a frame at once (same arithmetic on stacked arrays; they share the
    kinds of their chain, the caller checks that).  cams: CamSpecs, sizes: (W, H) per vcam.
    Returns (planes[V,5,4] f32, flags[V])."""
    V = len(cams)
    K = np.stack([np.asarray(c.K, np.float64) for c in cams])                   # (V,3,3)
    bottom_ok = (K[:, 2, 0] == 0) & (K[:, 2, 1] == 0) & (K[:, 2, 2] == 1)
    simple = (K[:, 0, 1] == 0) & (K[:, 1, 0] == 0) & (K[:, 0, 0] != 0) & (K[:, 1, 1] != 0) & bottom_ok
    M = np.broadcast_to(_I3, (V, 3, 3))
    c = np.zeros((V, 3))
    tau = np.zeros(V)
    ok = np.ones(V, bool)
    kinds = [k for k, _ in cams[0].ops]
    for k, kind in enumerate(kinds):
        m = np.stack([np.asarray(cam.ops[k][1], np.float64) for cam in cams])
        if kind == "T":
            c = c + m
            if k:
                tau = tau + np.abs(m).sum(1)
        else:
            L = m if kind == "R" else m[:, :, :3]
            ok &= np.abs(L @ L.transpose(0, 2, 1) - _I3).reshape(V, 9).max(1) < 1e-3
            M, c = L @ M, (L @ c[:, :, None])[:, :, 0]
            if kind == "A":
                c = c + m[:, :, 3]
                tau = tau + np.abs(m[:, :, 3]).sum(1)
    tref = np.asarray(tref, np.float64)
    if kinds and kinds[0] == "T":
        t0 = np.stack([np.asarray(cam.ops[0][1], np.float64) for cam in cams]) - tref      # residual, see cull_planes
        tau = tau + np.abs(t0).sum(1)
    else:
        tau = tau + np.abs(tref).sum()
    c2 = c - M @ tref                                                           # p = q - tref
    W = np.array([s[0] for s in sizes], np.float64)
    H = np.array([s[1] for s in sizes], np.float64)
    abc = np.empty((V, 5, 3))
    abc[:, 0] = (0.0, 0.0, 1.0)
    abc[:, 1] = K[:, 0]
    abc[:, 2, :2], abc[:, 2, 2] = -K[:, 0, :2], W - K[:, 0, 2]
    abc[:, 3] = K[:, 1]
    abc[:, 4, :2], abc[:, 4, 2] = -K[:, 1, :2], H - K[:, 1, 2]
    d0 = np.zeros((V, 5))
    d0[:, 0] = -float(min_dist32)
    kappa = 1.75 * np.abs(abc).sum(2)                                            # (V,5)
    with np.errstate(divide="ignore", invalid="ignore"):
        n = (abc @ M) / kappa[:, :, None]
        d = ((abc @ c2[:, :, None])[:, :, 0] + d0) / kappa
    good = ok & bottom_ok & np.isfinite(M).reshape(V, 9).all(1) & (np.abs(n).sum(2) <= 1.0).all(1) & np.isfinite(d).all(1)
    with np.errstate(invalid="ignore"):
        mg0 = 2.0 ** -17 * tau + 2.0 ** -20 * np.abs(d).max(1) + 1e-30
    out = np.empty((V, 5, 4))
    out[:, :, :3] = n
    out[:, :, 3] = d + mg0[:, None]
    with np.errstate(invalid="ignore", over="ignore"):
        out32 = out.astype(np.float32)
    out32[:, :, 3] = np.nextafter(out32[:, :, 3], np.float32(np.inf))           # rounding of d never tightens a plane
    out32[~good] = _OFF_PLANES
    return out32, simple.astype(np.int32)


def _chain_sig(ops) -> int:
    kinds = {"T": 1, "R": 2, "A": 3}
    return sum(kinds[k] << (2 * i) for i, (k, _) in enumerate(ops))


class PinnedPool:
    """Reusable pinned host buffers for the packers (cudaHostAlloc is slow and serialises the packer threads).
    take() hands out a pinned uint8 tensor of at least `nbytes`; give() returns it once the batch's
    host->device copies are done (PackedBatch.release)."""

    def __init__(self, max_bytes: int = 24 << 30):
        import threading
        self._lock = threading.Lock()
        self._free = []
        self._free_bytes = 0
        self.max_bytes = int(max_bytes)
        self.allocations = 0                  # cudaHostAlloc calls so far (a warm stream makes none)

    def take(self, nbytes: int):
        import torch
        nbytes = max(int(nbytes), 16)
        with self._lock:
            best = -1
            for k, t in enumerate(self._free):           # smallest buffer that fits, but not a giant for a small request
                if nbytes <= t.numel() <= max(2 * nbytes, 1 << 16) and (best < 0 or t.numel() < self._free[best].numel()):
                    best = k
            if best >= 0:
                t = self._free.pop(best)
                self._free_bytes -= t.numel()
                return t
            self.allocations += 1
        size = max(nbytes * 5 // 4, 1 << 16)             # head room: batches vary; small tables share one size class
        return torch.empty((size + 4095) & ~4095, dtype=torch.uint8, pin_memory=True)

    def give(self, tensors):
        with self._lock:
            for t in tensors:
                if self._free_bytes + t.numel() <= self.max_bytes:
                    self._free.append(t)
                    self._free_bytes += t.numel()


def _alloc(n, dtype, pin, pool=None, taken=None):
    """Host buffer of n elements; pinned torch memory when asked and possible (from `pool` when given)."""
    n = max(int(n), 1)
    if pin and pool is not None:
        import torch
        base = pool.take(n * np.dtype(dtype).itemsize)
        taken.append(base)
        t = base[:n * np.dtype(dtype).itemsize].view(getattr(torch, np.dtype(dtype).name))
        return t.numpy(), t
    if pin:
        import torch
        t = torch.empty(n, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
        return t.numpy(), t
    a = np.empty(n, dtype)
    return a, None


@dataclass
class PackedBatch:
    n_frames: int
    n_sweeps: int
    n_tiles: int
    n_vcams: int
    n_inst: int
    n_chains: int
    max_inst_per_frame: int
    cnt_total: int                  # sum over frames of tiles(f) * instances(f)
    bits_words: int                 # total words of one set of bit planes
    max_words: int                  # largest single plane
    n_raw_points: int
    masks_kind: str                 # "dense" | "rle" (uint32 runs) | "rle_str" (pycocotools counts bytes)
    max_runs: int
    raw: np.ndarray                 # float32
    meta: np.ndarray                # int32
    mask: np.ndarray                # uint8 (dense bytes / counts strings) or uint32 (runs)
    mask_off: np.ndarray            # int64: src_off[n_inst] or run_off[n_inst+1]
    off: Dict[str, int] = field(default_factory=dict)       # table name -> int32 offset in meta
    frame_inst: np.ndarray = None   # (F+1,) instance ranges
    frame_vcam_cams: List[List[tuple]] = field(default_factory=list)  # per frame [(cam, W, H)]
    tensors: dict = field(default_factory=dict)             # pinned torch views of raw/meta/mask/mask_off
    grid_words: int = 0             # words of the per-vcam instance lookup grids
    max_cells: int = 0
    any_kitti: bool = False         # a KITTI frame is in the batch -> Lifter also computes the OBB yaw
    frame_datasets: List[str] = field(default_factory=list)

    def table(self, name: str, words: int = 1) -> np.ndarray:
        o, n = self.off[name], self.off[name + "_n"]
        v = self.meta[o:o + n]
        return v.reshape(-1, words) if words > 1 else v

    def release(self):
        """Hand the pinned buffers back to the pool they came from (after the host->device copies are done;
        the batch's host arrays must not be used afterwards).  No-op for batches that own their memory."""
        pool, taken = getattr(self, "_pool", None), getattr(self, "_taken", None)
        if pool is not None and taken:
            self._pool = self._taken = None
            pool.give(taken)

    @property
    def h2d_bytes(self) -> int:
        return int(self.raw.nbytes + self.meta.nbytes + self.mask.nbytes + self.mask_off.nbytes)


def _eff_fourth(f: FrameSpec, keep_fourth: bool) -> int:
    """What row 3 of the aggregated cloud holds ON THE DEVICE: the frame's own `fourth`, or nothing when the
    caller does not read point rows back.  No output of the reference depends on it (the medoid takes
    `[:3]`, the centroid `points[:3]`: nuscenes:645,656), and it is a quarter of the sweep bytes."""
    return f.fourth if keep_fourth else FOURTH_NONE


def pack_frames(frames: Sequence[FrameSpec], pin: bool = False, keep_fourth: bool = True) -> PackedBatch:
    F = len(frames)
    if F == 0:
        raise ValueError("empty batch")
    def _kind(f):
        if isinstance(f.masks, np.ndarray):
            return "dense"
        return "rle_str" if all(isinstance(m.counts, (bytes, str)) for m in f.masks) else "rle"
    kinds = {_kind(f) for f in frames if f.n_instances}
    # dense uint8 | pycocotools `counts` strings (decoded on the GPU) | uint32 run lengths
    masks_kind = kinds.pop() if len(kinds) == 1 else "rle"

    # ---- geometry pass
    n_sweeps = sum(len(f.sweeps) for f in frames)
    n_inst = sum(f.n_instances for f in frames)
    sweep_tiles, raw_off = [], []
    ro = 0
    def packed(s, fourth):      # only the columns the kernels read travel: x, y, z (+ column 3 when it becomes row 3)
        ncol = 4 if fourth == FOURTH_COL3 else 3
        return s[:, :ncol] if s.shape[1] != ncol else s
    for f in frames:
        for s in f.sweeps:
            s = packed(s, _eff_fourth(f, keep_fourth))
            if s.shape[0] == 0:
                sweep_tiles.append(0)
                raw_off.append(ro)
                continue
            sweep_tiles.append(-(-s.shape[0] // TILE))
            raw_off.append(ro)
            ro += (s.size + 3) & ~3
    n_tiles = int(sum(sweep_tiles))
    raw, raw_t = _alloc(ro + 4, np.float32, pin)
    raw[ro:] = 0

    sweep_desc = np.zeros((max(n_sweeps, 1), SW_WORDS), np.int32)
    frame_desc = np.zeros((F, FR_WORDS), np.int32)
    tile_sweep = np.zeros(max(n_tiles, 1), np.int32)
    inst_desc = np.zeros((max(n_inst, 1), IN_WORDS), np.int32)
    cam_inst_list = np.zeros(max(n_inst, 1), np.int32)
    chains: List[np.ndarray] = []
    vcam_rows: List[np.ndarray] = []
    frame_vcam_cams = []
    frame_inst = np.zeros(F + 1, np.int64)

    si = ti = ii = 0
    cnt_total = bits_words = max_words = grid_words = max_cells = 0
    max_inst_pf = 0
    mask_chunks, mask_off = [], [0]
    max_runs = 0
    for fi, f in enumerate(frames):
        I = f.n_instances
        if I > MAX_INST:
            raise ValueError(f"frame {fi}: {I} instances > {MAX_INST} (CM3D_ELIMIT)")
        max_inst_pf = max(max_inst_pf, I)
        t_begin = ti
        for s_local, (s, ops) in enumerate(zip(f.sweeps, f.sweep_ops)):
            s = packed(s, _eff_fourth(f, keep_fourth))
            nt = sweep_tiles[si]
            o = raw_off[si]
            raw[o:o + s.size] = s.reshape(-1)
            pad = ((s.size + 3) & ~3) - s.size
            if pad:
                raw[o + s.size:o + s.size + pad] = 0
            sweep_desc[si] = [*_split64(o), s.shape[0], s.shape[1], fi, ti, len(chains), _eff_fourth(f, keep_fourth)]
            chains.append(encode_chain(ops))
            tile_sweep[ti:ti + nt] = si
            ti += nt
            si += 1
        ntf = ti - t_begin
        # vcams: unique (camera, W, H) among the frame's instances
        sizes = [f.mask_size(i) for i in range(I)]
        keys = sorted({(int(f.cam_nums[i]),) + sizes[i] for i in range(I)})
        if len(keys) > MAX_VCAMS:
            raise ValueError(f"frame {fi}: {len(keys)} (camera, mask size) pairs > {MAX_VCAMS} (CM3D_ELIMIT)")
        v_begin = len(vcam_rows)
        lb = 0
        key_index = {k: n for n, k in enumerate(keys)}
        members = [[] for _ in keys]
        for i in range(I):
            members[key_index[(int(f.cam_nums[i]),) + sizes[i]]].append(i)
        first = f.cams[keys[0][0]].ops if keys else []
        tref = np.asarray(first[0][1], np.float32) if (first and first[0][0] == "T") else np.zeros(3, np.float32)
        sigs = {_chain_sig(f.cams[k[0]].ops) for k in keys}
        vcams = [f.cams[k[0]] for k in keys]
        if len(vcams) > 1 and len({tuple(kind for kind, _ in c.ops) for c in vcams}) == 1:
            planes_all, flags_all = cull_planes_many(vcams, [(k[1], k[2]) for k in keys], f.min_dist_f32(), tref)
        else:
            pf = [cull_planes(c, k[1], k[2], f.min_dist_f32(), tref) for c, k in zip(vcams, keys)]
            planes_all, flags_all = [a for a, _ in pf], [b for _, b in pf]
        for vi, (k, mem) in enumerate(zip(keys, members)):
            cam = f.cams[k[0]]
            row = np.zeros(VC_WORDS, np.int32)
            planes, flags = planes_all[vi], int(flags_all[vi])
            row[20:40] = planes.reshape(-1).view(np.int32)
            row[40] = flags
            row[0] = len(chains)
            chains.append(encode_chain(cam.ops))
            row[1:13] = cam.viewpad34().reshape(-1).view(np.int32)
            row[13], row[14], row[15], row[16] = k[1], k[2], lb, len(mem)
            gnx, gny = -(-k[1] // CELL), -(-k[2] // CELL)
            row[17], row[18], row[19] = fi, grid_words, gnx
            grid_words += gnx * gny * ((len(mem) + 31) // 32)
            max_cells = max(max_cells, gnx * gny)
            cam_inst_list[ii + lb: ii + lb + len(mem)] = mem
            lb += len(mem)
            vcam_rows.append(row)
        frame_vcam_cams.append(keys)
        use_close = f.close_thresh is not None
        min_pts = 4 if f.dataset == "kitti" else 1          # kitti/2d_to_3d.py:1479-1480 skips M <= 3
        frame_desc[fi] = [t_begin, ti, v_begin, len(keys), ii, I,
                          _f32_bits(f.close_thresh if use_close else 0.0), int(use_close),
                          _f32_bits(f.min_dist_f32()), cnt_total, min_pts, ii,
                          _f32_bits(tref[0]), _f32_bits(tref[1]), _f32_bits(tref[2]),
                          sigs.pop() if len(sigs) == 1 else -1,
                          _f32_bits(0.0 if f.floor_thresh is None else f.floor_thresh), int(f.floor_thresh is not None),
                          0, 0]
        cnt_total += ntf * I
        for i in range(I):
            W, H = sizes[i]
            pitch = (W + 31) // 32
            words = pitch * H
            inst_desc[ii + i] = [*_split64(bits_words), W, H, pitch,
                                 v_begin + key_index[(int(f.cam_nums[i]),) + sizes[i]], fi, i]
            bits_words += words
            max_words = max(max_words, words)
            if masks_kind == "dense":
                m = f.masks[i]
                if m.shape != (H, W):
                    raise ValueError("dense mask shape mismatch")
                mask_chunks.append(np.ascontiguousarray(m, np.uint8).reshape(-1))
                mask_off.append(mask_off[-1] + ((H * W + 15) & ~15))
            elif masks_kind == "rle_str":
                c = f.masks[i].counts
                c = np.frombuffer(c.encode("ascii") if isinstance(c, str) else c, np.uint8)
                mask_chunks.append(c)
                mask_off.append(mask_off[-1] + len(c))
                max_runs = max(max_runs, len(c))
            else:
                rle = f.masks[i]
                if isinstance(f.masks, np.ndarray):
                    from .synthetic import dense_to_rle
                    rle = dense_to_rle(f.masks[i][None])[0]
                runs = rle_counts_to_runs(rle.counts)
                mask_chunks.append(runs)
                mask_off.append(mask_off[-1] + len(runs))
                max_runs = max(max_runs, len(runs))
        ii += I
        frame_inst[fi + 1] = ii

    if masks_kind == "dense":
        mask, mask_t = _alloc(mask_off[-1], np.uint8, pin)
        for c, o in zip(mask_chunks, mask_off[:-1]):
            mask[o:o + c.size] = c
        mask_off_arr = np.asarray(mask_off[:-1] if n_inst else [0], np.int64)
    else:
        mask, mask_t = _alloc(mask_off[-1], np.uint8 if masks_kind == "rle_str" else np.uint32, pin)
        if mask_chunks:
            np.concatenate(mask_chunks, out=mask[:mask_off[-1]])
        mask_off_arr = np.asarray(mask_off, np.int64)

    chains_arr = np.stack(chains).astype(np.uint32) if chains else np.zeros((1, CHAIN_WORDS), np.uint32)
    vcam_desc = np.stack(vcam_rows) if vcam_rows else np.zeros((1, VC_WORDS), np.int32)
    tables = [("tile_sweep", tile_sweep), ("sweep_desc", sweep_desc), ("frame_desc", frame_desc),
              ("vcam_desc", vcam_desc), ("cam_inst_list", cam_inst_list), ("inst_desc", inst_desc),
              ("chains", chains_arr.view(np.int32))]
    off, pos = {}, 0
    for name, arr in tables:
        off[name] = pos
        off[name + "_n"] = arr.size
        pos += (arr.size + 3) & ~3            # keep every table 16-byte aligned
    meta, meta_t = _alloc(pos, np.int32, pin)
    for name, arr in tables:
        meta[off[name]: off[name] + arr.size] = arr.reshape(-1)

    mo, mo_t = _alloc(mask_off_arr.size, np.int64, pin)
    mo[:mask_off_arr.size] = mask_off_arr

    pb = PackedBatch(F, n_sweeps, n_tiles, len(vcam_rows), n_inst, len(chains), max_inst_pf, cnt_total,
                     bits_words, max_words, sum(f.n_raw_points for f in frames), masks_kind, max_runs,
                     raw, meta, mask, mo[:mask_off_arr.size], off, frame_inst, frame_vcam_cams,
                     {"raw": raw_t, "meta": meta_t, "mask": mask_t, "mask_off": mo_t})
    pb.any_kitti = any(f.dataset == "kitti" for f in frames)
    pb.frame_datasets = [f.dataset for f in frames]
    pb.grid_words, pb.max_cells = grid_words, max_cells
    return pb


# ------------------------------------------------------------------------------------------ native packer
_KIND_CODE = {"T": 1, "R": 2, "A": 3}
_PACK_PTR_FIELDS = ("fr_n_sweeps", "fr_n_cams", "fr_n_inst", "fr_fourth", "fr_min_pts", "fr_use_close", "fr_use_floor",
                    "fr_close", "fr_min_dist", "fr_floor", "sw_ptr", "sw_npts", "sw_stride", "op_begin", "op_kind",
                    "op_ptr", "cam_K", "in_cam", "in_W", "in_H", "in_counts_off", "counts")


def _pack_input_type():
    """ctypes mirror of `cm3d_pack_input` (include/cm3d_b200.h)."""
    import ctypes

    class PackInput(ctypes.Structure):
        _fields_ = [("n_frames", ctypes.c_int32), ("n_sweeps", ctypes.c_int32), ("n_cams", ctypes.c_int32),
                    ("n_inst", ctypes.c_int32)] + [(k, ctypes.c_void_p) for k in _PACK_PTR_FIELDS]
    return PackInput


_PackInput = None


def _ptr_of(a: np.ndarray) -> int:
    return a.__array_interface__["data"][0]


def pack_frames_native(frames: Sequence[FrameSpec], pin: bool = False, pool: "PinnedPool" = None,
                       keep_fourth: bool = True) -> PackedBatch:
    """pack_frames through the C packer of the library (csrc/pack.cu): the FrameSpecs are flattened into
    pointer / size arrays here, everything else - the copy of the raw sweeps into (pinned) memory, the
    descriptor tables, the fp64 cull planes - happens in one ctypes call that holds no GIL, so several
    batches are packed at once by worker threads.  Same buffers as pack_frames (tests/test_host_logic.py).
    Batches it does not cover (dense masks, decoded run lengths, no instances at all) go to pack_frames."""
    import ctypes
    from . import _native as N
    F = len(frames)
    if F == 0:
        raise ValueError("empty batch")
    for f in frames:
        if isinstance(f.masks, np.ndarray) or any(not isinstance(m.counts, (bytes, str)) for m in f.masks):
            return pack_frames(frames, pin, keep_fourth)
    if not any(f.n_instances for f in frames):
        return pack_frames(frames, pin, keep_fourth)
    if any(len(o) > 4 for f in frames for o in f.sweep_ops) or any(len(c.ops) > 4 for f in frames for c in f.cams):
        return pack_frames(frames, pin, keep_fourth)              # raises the chain-length error with its message
    lib = N.load()

    # Every small matrix of the batch (chain ops, intrinsics) is gathered into ONE float32 blob with a single
    # concatenate; the C side gets pointers into it (base + offset).  A per-matrix pointer lookup costs 2 us of
    # GIL time, 2.6k of them per 32-frame batch would be a third of the stream's per-frame budget.
    sw_ptr, sw_npts, sw_stride, op_begin, op_kind, mats, k_slot = [], [], [], [0], [], [], []
    for f in frames:
        for s, ops in zip(f.sweeps, f.sweep_ops):
            sw_ptr.append(_ptr_of(s)); sw_npts.append(s.shape[0]); sw_stride.append(s.shape[1])
            for kind, m in ops:
                op_kind.append(_KIND_CODE[kind]); mats.append(m.reshape(-1))
            op_begin.append(len(op_kind))
    for f in frames:
        for cam in f.cams:
            for kind, m in cam.ops:
                op_kind.append(_KIND_CODE[kind]); mats.append(m.reshape(-1))
            op_begin.append(len(op_kind))
            k_slot.append(len(mats)); mats.append(cam.K.reshape(-1))
    sizes_ = np.fromiter(map(len, mats), np.int64, len(mats))
    starts = np.zeros(len(mats) + 1, np.int64)
    np.cumsum(sizes_, out=starts[1:])
    mat_blob = np.concatenate(mats).astype(np.float32, copy=False) if mats else np.zeros(1, np.float32)
    mat_ptr = (np.uint64(_ptr_of(mat_blob)) + (starts[:-1] * 4).astype(np.uint64))
    is_k = np.zeros(len(mats), bool)
    is_k[k_slot] = True
    op_ptr, cam_K = mat_ptr[~is_k], mat_ptr[is_k]
    counts = []
    for f in frames:
        for m in f.masks:
            c = m.counts
            counts.append(c.encode("ascii") if isinstance(c, str) else c)
    blob = b"".join(counts)
    in_counts_off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(np.fromiter(map(len, counts), np.int64, len(counts)), out=in_counts_off[1:])
    n_inst = len(counts)
    in_cam = np.concatenate([f.cam_nums for f in frames]).astype(np.int32) if n_inst else np.zeros(1, np.int32)
    sizes = np.array([m.size for f in frames for m in f.masks], np.int32).reshape(-1, 2) if n_inst else np.zeros((1, 2), np.int32)
    in_W, in_H = np.ascontiguousarray(sizes[:, 0]), np.ascontiguousarray(sizes[:, 1])

    i32 = lambda v: np.asarray(v, np.int32)
    f32 = lambda v: np.asarray(v, np.float32)
    arrs = dict(
        fr_n_sweeps=i32([len(f.sweeps) for f in frames]), fr_n_cams=i32([len(f.cams) for f in frames]),
        fr_n_inst=i32([f.n_instances for f in frames]), fr_fourth=i32([_eff_fourth(f, keep_fourth) for f in frames]),
        fr_min_pts=i32([4 if f.dataset == "kitti" else 1 for f in frames]),
        fr_use_close=i32([f.close_thresh is not None for f in frames]),
        fr_use_floor=i32([f.floor_thresh is not None for f in frames]),
        fr_close=f32([0.0 if f.close_thresh is None else f.close_thresh for f in frames]),
        fr_min_dist=f32([f.min_dist_f32() for f in frames]),
        fr_floor=f32([0.0 if f.floor_thresh is None else f.floor_thresh for f in frames]),
        sw_ptr=np.asarray(sw_ptr, np.uint64), sw_npts=i32(sw_npts), sw_stride=i32(sw_stride),
        op_begin=i32(op_begin), op_kind=i32(op_kind), op_ptr=np.ascontiguousarray(op_ptr, np.uint64),
        cam_K=np.ascontiguousarray(cam_K, np.uint64), in_cam=in_cam, in_W=in_W, in_H=in_H, in_counts_off=in_counts_off)
    global _PackInput
    if _PackInput is None:
        _PackInput = _pack_input_type()
        lib.cm3d_pack_input_size.restype = ctypes.c_int
        if lib.cm3d_pack_input_size() != ctypes.sizeof(_PackInput):
            raise N.Cm3dError("cm3d_pack_input: the library's struct and batch.py's mirror have drifted apart")
    inp = _PackInput(F, len(sw_ptr), int(cam_K.size), n_inst)
    for k, a in arrs.items():
        setattr(inp, k, _ptr_of(a))
    inp.counts = ctypes.cast(ctypes.c_char_p(blob), ctypes.c_void_p).value
    plan = np.zeros(16, np.int64)
    rc = lib.cm3d_pack_plan(ctypes.byref(inp), ctypes.c_void_p(_ptr_of(plan)))
    if rc != 0:
        return pack_frames(frames, pin, keep_fourth)          # raises the per-frame limit error with its message
    n_tiles, n_vcams, n_chains = int(plan[0]), int(plan[1]), int(plan[2])
    taken = []
    raw, raw_t = _alloc(int(plan[3]), np.float32, pin, pool, taken)
    meta, meta_t = _alloc(int(plan[4]), np.int32, pin, pool, taken)
    mask, mask_t = _alloc(int(plan[5]), np.uint8, pin, pool, taken)
    mo, mo_t = _alloc(n_inst + 1, np.int64, pin, pool, taken)
    vkeys = np.zeros((max(n_vcams, 1), 4), np.int32)
    out = np.zeros(8, np.int64)
    rc = lib.cm3d_pack_fill(ctypes.byref(inp), ctypes.c_void_p(_ptr_of(plan)), ctypes.c_void_p(_ptr_of(raw)),
                            ctypes.c_void_p(_ptr_of(meta)), ctypes.c_void_p(_ptr_of(mask)), ctypes.c_void_p(_ptr_of(mo)),
                            ctypes.c_void_p(_ptr_of(vkeys)), ctypes.c_void_p(_ptr_of(out)))
    N.check(rc, "cm3d_pack_fill")
    names = ["tile_sweep", "sweep_desc", "frame_desc", "vcam_desc", "cam_inst_list", "inst_desc", "chains"]
    lens = [max(n_tiles, 1), max(len(sw_ptr), 1) * SW_WORDS, F * FR_WORDS, max(n_vcams, 1) * VC_WORDS, max(n_inst, 1),
            max(n_inst, 1) * IN_WORDS, max(n_chains, 1) * CHAIN_WORDS]
    off = {}
    for k, (name, n) in enumerate(zip(names, lens)):
        off[name] = int(plan[6 + k])
        off[name + "_n"] = int(n)
    frame_inst = np.zeros(F + 1, np.int64)
    np.cumsum(arrs["fr_n_inst"], out=frame_inst[1:])
    frame_vcam_cams = [[] for _ in range(F)]
    for fr, cam, W, H in vkeys[:n_vcams].tolist():
        frame_vcam_cams[fr].append((cam, W, H))
    pb = PackedBatch(F, len(sw_ptr), n_tiles, n_vcams, n_inst, n_chains, int(plan[13]), int(out[0]), int(out[1]),
                     int(out[2]), int(plan[14]), "rle_str", int(out[5]), raw, meta, mask, mo[:n_inst + 1], off, frame_inst,
                     frame_vcam_cams, {"raw": raw_t, "meta": meta_t, "mask": mask_t, "mask_off": mo_t})
    pb.any_kitti = any(f.dataset == "kitti" for f in frames)
    pb.frame_datasets = [f.dataset for f in frames]
    pb.grid_words, pb.max_cells = int(out[3]), int(out[4])
    if pool is not None and taken:
        pb._pool, pb._taken = pool, taken
    return pb
