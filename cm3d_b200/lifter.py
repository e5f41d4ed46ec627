"""`Lifter`: the host side of the lifting path - uploads a PackedBatch, launches the
CUDA kernels through the C-ABI (include/cm3d_b200.h) and reads the labels back.

It stands where the reference has the body of its frame loop
(src/nuscenes/2d_to_3d.py:433-665; kitti:1066-1542; waymo:472-702): per frame,
aggregate the sweeps, and per mask find the LiDAR points inside it and their
medoid.  PyTorch is used for device memory, streams and pinned host buffers only.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .batch import (ERR_WORDS, FR_WORDS, MAX_INST, MEDOID_COLS, SCREEN_MIN_PTS, TILE, PackedBatch, pack_frames,
                    pack_frames_native, PinnedPool)
from .frames import FrameSpec, LiftResult


def _ptr(t):
    """Device pointer of a tensor (or an address already taken, or None) as a ctypes argument."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t if isinstance(t, int) else t.data_ptr())


@dataclass
class DeviceBatch:
    pb: PackedBatch
    raw: torch.Tensor
    meta: torch.Tensor
    mask: torch.Tensor
    mask_off: torch.Tensor

    def tab(self, name: str) -> torch.Tensor:
        o = self.pb.off[name]
        return self.meta[o:o + self.pb.off[name + "_n"]]

    def tab_ptr(self, name: str) -> int:
        """Address of a descriptor table (no tensor view: the launch sequence asks for ~30 of these per batch)."""
        return self.meta.data_ptr() + 4 * self.pb.off[name]


@dataclass
class DeviceOutputs:
    db: DeviceBatch
    out: torch.Tensor                 # packed int32 result block (see Lifter._out_layout)
    layout: dict
    seg_cap: int
    seg_point_idx: torch.Tensor
    seg_xyzw: torch.Tensor
    xyzw: torch.Tensor
    tile_cnt: torch.Tensor
    tile_prefix: torch.Tensor
    pix: Optional[torch.Tensor] = None
    col_sums: Optional[torch.Tensor] = None
    hits: Optional[torch.Tensor] = None
    bits: Optional[torch.Tensor] = None
    bbox: Optional[torch.Tensor] = None
    obb: Optional[torch.Tensor] = None     # (I,16): yaw, centre, wlh, R' (KITTI frames / want_obb)
    box: Optional[torch.Tensor] = None     # (I,8): orientation search (box_search=n_angles)
    seg_off_raw: Optional[torch.Tensor] = None   # (I+1,) offsets before the neighbour-count filter
    done: Optional[object] = None                # overlap mode: event recorded after the last kernel (on the medoid stream)


def split_oversize(frames: Sequence[FrameSpec], limit: int):
    """Frames with more than `limit` instances (the kernels keep instance ids in one byte) are split
    into sub-frames over the same sweeps and cameras with disjoint instance ranges; the reference has
    no such limit.  Returns (flat list of frames, parts) with parts[k] = number of sub-frames of input k."""
    import copy
    flat, parts = [], []
    for f in frames:
        n = f.n_instances
        if n <= limit:
            flat.append(f)
            parts.append(1)
            continue
        k = -(-n // limit)
        for a in range(0, n, limit):
            g = copy.copy(f)
            sl = slice(a, min(a + limit, n))
            g.cam_nums = f.cam_nums[sl]
            g.masks = f.masks[sl]
            g.labels = list(f.labels[sl])
            g.scores = list(f.scores[sl])
            flat.append(g)
        parts.append(k)
    return flat, parts


def merge_split(results: List[LiftResult], parts: Sequence[int]) -> List[LiftResult]:
    out, p = [], 0
    for k in parts:
        grp = results[p:p + k]
        p += k
        if k == 1:
            out.append(grp[0])
            continue
        cat = lambda name: (None if getattr(grp[0], name) is None else np.concatenate([getattr(g, name) for g in grp]))
        offs, base = [np.zeros(1, np.int32)], 0
        for g in grp:
            offs.append(g.seg_offsets[1:] + base)
            base += int(g.seg_offsets[-1])
        r = LiftResult(n_points=grp[0].n_points, seg_offsets=np.concatenate(offs).astype(np.int32),
                       seg_point_idx=cat("seg_point_idx"), medoid_local=cat("medoid_local"),
                       medoid_point_idx=cat("medoid_point_idx"), centroids=cat("centroids"),
                       aggr_points=grp[0].aggr_points, pix=None, yaw=cat("yaw"), obb=cat("obb"), box=cat("box"),
                       raw_counts=cat("raw_counts"))
        out.append(r)
    return out


def prefetch_map(fn, items, workers: int = 8, depth: int = 0):
    """Ordered `map(fn, items)` on reader threads with a bounded look-ahead: the drop-in scripts build their
    FrameSpecs with it (np.fromfile of the sweeps, pickle / json of the masks and the devkit record
    lookups run while earlier frames are being packed, copied and lifted).  The reference reads its
    files serially inside the frame loop (src/nuscenes/2d_to_3d.py:422-438)."""
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor
    workers = max(1, int(workers))
    depth = depth or 2 * workers
    if workers == 1:
        for x in items:
            yield fn(x)
        return
    with ThreadPoolExecutor(max_workers=workers) as pool:
        pending = deque()
        for x in items:
            pending.append(pool.submit(fn, x))
            if len(pending) >= depth:
                yield pending.popleft().result()
        while pending:
            yield pending.popleft().result()


def _round_up(n: int) -> int:
    """n rounded up to 1, 1.125, ... 1.875 times a power of two (above 64 Ki): the workspace sizes of a stream
    vary from batch to batch (mask sizes, member counts), and a caching allocator only reuses a block
    for a request that is not larger - a handful of size classes keeps cudaMalloc out of the steady state."""
    n = int(n)
    if n < (1 << 16):
        return n
    q = 1 << (n.bit_length() - 4)
    return (n + q - 1) // q * q


def _on_device(fn):
    """Run a Lifter method with the lifter's GPU as the current CUDA device (generators included)."""
    import functools
    import inspect
    if inspect.isgeneratorfunction(fn):
        @functools.wraps(fn)
        def gen(self, *a, **k):
            it = fn(self, *a, **k)
            while True:
                with torch.cuda.device(self.device):
                    try:
                        item = next(it)
                    except StopIteration:
                        return
                yield item
        return gen

    @functools.wraps(fn)
    def call(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return call


class Lifter:
    """One per process / GPU.  All methods enqueue on the current torch CUDA stream."""

    def __init__(self, device="cuda:0", seg_factor: float = 2.0):
        if not torch.cuda.is_available():
            raise N.Cm3dError("cm3d_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        # one process per GPU: the C ABI launches on the CURRENT CUDA device, so make this one current
        # (run()/upload()/the stream paths also guard it for callers that switch devices in between)
        torch.cuda.set_device(self.device)
        self.seg_factor = float(seg_factor)
        N.load()
        self._med_stream = None     # second-phase stream of run(overlap=True)
        self._graph_runner = None   # lift_frame_graph(): CUDA graphs by batch geometry
        self._streams = None        # (copy, compute) streams of the pipelined path, created once: torch's
        #                             caching allocator pools memory per stream, so fresh streams per call
        #                             would cudaMalloc the whole workspace again (~100 ms)
        self._pin_pool = None       # pinned host buffers of the streaming path, reused across batches
        self._seg_ratio = None      # largest member-points / raw-points ratio the streaming path has seen
        self.denoise = None         # default-off extensions, see run()
        self.box_search = None
        self.launches = 0           # kernels launched by this object (bench.py reports it)
        self.cap_retries = 0        # streamed batches rerun because the segment buffers were too small
        self.stream_stats = {}      # seconds the streaming path spent waiting for packers / the GPU / enqueueing (diagnostics)
        self.screen_min_pts = SCREEN_MIN_PTS   # medoid: instances this large are screened, then verified; 0 = all exact
        self.screen_flags = 0                  # bit 0: no symmetric screen (every screened instance does all M^2 pairs);
        #                                        bit 1: no grouped symmetric screen for instances that straddle two binades;
        #                                        bit 2: no exact column pruning for sensor-frame clouds
        self.last_screen_stats = None          # device int32[1]: columns the last run() verified exactly
        self.last_screen_modes = None          # device int32[3 I]: mode (0 exact, 1 all pairs, 2 symmetric, 3 grouped symmetric,
        #                                        4 all pairs over pruned columns) | points in group 0 (mode 3) or columns kept
        #                                        (mode 4) | points in the sliver group (mode 3)
        self.timing = None          # dict label -> [(start_event, end_event)] when bench.py profiles
        self.obb_mode = 0           # KITTI box: 0 = hull vertices (open3d's algorithm), 1 = all member points (diagnostics)
        self.last_hull_info = None  # device int32[I]: hull vertex count per instance (-1: flat cloud -> the reference's fallback box)
        self.fused = True           # run(): one C call over one workspace (cm3d_lift_batch) when nothing asks for the call-by-call path
        self._launch_count = ctypes.c_int32(0)

    def _buf(self, n, dtype=torch.int32):
        """Uninitialised device buffer of n elements, from an allocation of a rounded-up size class."""
        n = max(int(n), 1)
        return torch.empty(_round_up(n), dtype=dtype, device=self.device)[:n]

    def _call(self, label: str, name: str, *args):
        """One C-ABI call; with `self.timing` set, bracketed by CUDA events on the launch stream."""
        if self.timing is None:
            N.call(name, *args)
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        N.call(name, *args)
        b.record()
        self.timing.setdefault(label, []).append((a, b))

    # ------------------------------------------------------------------ host -> device
    def pack(self, frames: Sequence[FrameSpec], keep_fourth: bool = True) -> PackedBatch:
        """FrameSpecs -> pinned host buffers.  keep_fourth=False leaves the 4th point column (nuScenes
        intensity, KITTI reflectance) on the host: nothing the reference outputs reads it, and it is a
        quarter of the bytes that cross PCIe; only LiftResult.aggr_points row 3 needs it."""
        # the C packer (csrc/pack.cu) for batches whose masks are counts strings, the Python packer otherwise
        return pack_frames_native(frames, pin=True, keep_fourth=keep_fourth)

    def _pack_pooled(self, frames: Sequence[FrameSpec]) -> PackedBatch:
        """pack() into pinned buffers of this Lifter's pool; the streaming path releases them after use.
        The streaming path returns labels only, so the 4th point column stays on the host."""
        if self._pin_pool is None:
            self._pin_pool = PinnedPool()
        return pack_frames_native(frames, pin=True, pool=self._pin_pool, keep_fourth=False)

    @_on_device
    def upload(self, pb: PackedBatch) -> DeviceBatch:
        def up(name, arr):
            t = pb.tensors.get(name)
            src = t if t is not None else torch.from_numpy(arr)
            if arr.dtype == np.uint32:          # torch has no uint32 arithmetic; move the bytes as int32
                src = src.view(torch.int32) if t is not None else torch.from_numpy(arr.view(np.int32))
            return src.to(self.device, non_blocking=True)
        return DeviceBatch(pb, up("raw", pb.raw), up("meta", pb.meta), up("mask", pb.mask),
                           up("mask_off", pb.mask_off))

    # ------------------------------------------------------------------ launch sequence
    @staticmethod
    def _out_layout(F: int, I: int) -> dict:
        lay, pos = {}, 0
        for name, n in (("frame_n", F), ("seg_off", I + 1), ("item_off", I + 1), ("medoid_local", I),
                        ("medoid_point_idx", I), ("centroid", 4 * I), ("errflags", ERR_WORDS)):
            lay[name] = (pos, n)
            pos += (n + 3) & ~3
        lay["_words"] = pos
        return lay

    @_on_device
    def run(self, db: DeviceBatch, seg_cap: Optional[int] = None, want_pix: bool = False,
            want_col_sums: bool = False, do_medoid: bool = True, want_obb: Optional[bool] = None,
            denoise=None, box_search: Optional[int] = None, overlap: bool = False) -> DeviceOutputs:
        """One launch sequence over a packed batch.  `overlap=True` (the streaming paths and bench.py): the
        medoid and box kernels go to a second stream behind an event, so that the HBM-bound front end
        (masks, aggregation, projection, gather) of the NEXT batch runs next to this batch's XU-bound medoid
        when the caller launches the front ends on a higher-priority stream; `DeviceOutputs.done` is then the
        event to wait for before touching any output.  Default-off extensions (not executed by the
        reference, parity unpinned): `denoise=(radius, min_neighbors)` drops member points with fewer
        than min_neighbors members of their instance within radius before the medoid / boxes;
        `box_search=n_angles` adds the orientation / extent search (LiftResult.box)."""
        denoise = denoise if denoise is not None else self.denoise
        box_search = box_search if box_search is not None else self.box_search
        if (self.fused and self.timing is None and do_medoid and not want_pix and not want_col_sums and denoise is None
                and not box_search):
            return self._run_fused(db, seg_cap, want_obb, overlap)
        pb, dev = db.pb, self.device
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        F, I, T = pb.n_frames, pb.n_inst, pb.n_tiles
        n_slots = max(T, 1) * TILE
        if seg_cap is None:
            # members per raw point: the configured bound until the streaming path has seen batches, then 1.5 x the
            # largest ratio seen (+5 %); a batch that does not fit is rerun with its exact size (check_flags)
            factor = self.seg_factor if self._seg_ratio is None else min(self.seg_factor, 1.5 * self._seg_ratio + 0.05)
            seg_cap = int(factor * pb.n_raw_points) + 1024
        seg_cap = _round_up((int(seg_cap) + 3) & ~3) if overlap else (int(seg_cap) + 3) & ~3     # streams: few distinct sizes
        i32 = dict(dtype=torch.int32, device=dev)
        lay = self._out_layout(F, I)
        out = torch.zeros(lay["_words"], **i32)

        def o(name):
            a, n = lay[name]
            return out[a:a + max(n, 1)]

        out_base = out.data_ptr()

        def op(name):                       # address of a field of the label block
            return out_base + 4 * lay[name][0]

        # ---- masks -> eroded bit planes (+ bbox)
        bits_raw = self._buf(max(pb.bits_words, 1))
        bits = self._buf(max(pb.bits_words, 1))
        bbox = self._buf(max(I, 1) * 4)
        inst_desc = db.tab_ptr("inst_desc")
        row_range = None
        if I:
            if pb.masks_kind == "dense":
                self._call("masks_pack", "cm3d_masks_pack_dense", _ptr(db.mask), _ptr(db.mask_off), _ptr(inst_desc), I,
                       pb.max_words, _ptr(bits_raw), st)
                self.launches += 1
            else:
                bits_raw.zero_()
                runs = db.mask
                if pb.masks_kind == "rle_str":      # pycocotools counts strings -> run lengths on the device
                    runs = self._buf(max(db.mask.numel(), 1))
                    self._call("masks_decode", "cm3d_masks_decode_counts", _ptr(db.mask), _ptr(db.mask_off), I, _ptr(runs), st)
                    self.launches += 1
                run_start = self._buf(max(runs.numel(), 1))
                row_range = self._buf(2 * I)
                self._call("masks_rle", "cm3d_masks_fill_rle", _ptr(runs), _ptr(db.mask_off), _ptr(run_start), _ptr(inst_desc),
                       I, pb.max_runs, _ptr(bits_raw), _ptr(row_range), _ptr(op("errflags")), st)
                self.launches += 2              # k_rle_prefix, k_rle_fill (torch's zero fill of the planes is not ours)
            self._call("masks_erode", "cm3d_masks_erode3x3", _ptr(bits_raw), _ptr(inst_desc), _ptr(row_range), I,
                       pb.max_words, _ptr(bits), _ptr(bbox), st)
            self.launches += 2

        vcam_grid = self._buf(max(pb.grid_words, 1))
        if I:
            self._call("vcam_grid", "cm3d_build_vcam_grid", _ptr(db.tab_ptr("vcam_desc")), pb.n_vcams, pb.max_cells,
                       _ptr(db.tab_ptr("frame_desc")), _ptr(db.tab_ptr("cam_inst_list")), _ptr(bbox), _ptr(vcam_grid), st)
            self.launches += 1

        # ---- sweeps -> aggregated cloud
        xyzw = self._buf(4 * n_slots, torch.float32)
        tile_cnt = self._buf(max(T, 1))
        tile_prefix = self._buf(max(T, 1))
        self._call("aggregate", "cm3d_aggregate_sweeps", _ptr(db.raw), _ptr(db.tab_ptr("tile_sweep")), T, _ptr(db.tab_ptr("sweep_desc")),
               _ptr(db.tab_ptr("frame_desc")), _ptr(db.tab_ptr("chains")), _ptr(xyzw), _ptr(tile_cnt), st)
        self.launches += 1 if T else 0

        # ---- projection + membership (count pass)
        hits = self._buf(n_slots)
        tile_inst_cnt = self._buf(max(pb.cnt_total, 1), torch.int16)
        tile_inst_base = self._buf(max(pb.cnt_total, 1))
        pix = self._buf(16 * n_slots) if want_pix else None
        self._call("project_count", "cm3d_project_membership", _ptr(xyzw), _ptr(tile_cnt), _ptr(db.tab_ptr("tile_sweep")), T,
               _ptr(db.tab_ptr("sweep_desc")), _ptr(db.tab_ptr("frame_desc")), _ptr(db.tab_ptr("vcam_desc")),
               _ptr(db.tab_ptr("cam_inst_list")), _ptr(inst_desc), _ptr(bbox), _ptr(db.tab_ptr("chains")), _ptr(bits),
               _ptr(vcam_grid), _ptr(hits), _ptr(tile_inst_cnt), _ptr(pix), st)
        self.launches += 1 if T else 0

        # ---- scans, ordered compaction + gather
        medoid_best = self._buf(max(I, 1), torch.int64)
        item_inst = self._buf(max(I, 1))
        self._call("scan", "cm3d_scan_segments", _ptr(tile_cnt), _ptr(tile_inst_cnt), _ptr(db.tab_ptr("frame_desc")), F,
               pb.max_inst_per_frame, I, _ptr(inst_desc), seg_cap, _ptr(tile_prefix), _ptr(op("frame_n")),
               _ptr(tile_inst_base), _ptr(op("seg_off")), _ptr(op("item_off")), _ptr(item_inst),
               _ptr(medoid_best), _ptr(op("errflags")), st)
        self.launches += 2
        seg_point_idx = self._buf(seg_cap)
        seg_xyzw = self._buf(4 * seg_cap, torch.float32)
        self._call("compact", "cm3d_compact_segments", _ptr(xyzw), _ptr(tile_cnt), _ptr(tile_prefix), _ptr(db.tab_ptr("tile_sweep")),
               T, _ptr(db.tab_ptr("sweep_desc")), _ptr(db.tab_ptr("frame_desc")), _ptr(db.tab_ptr("vcam_desc")),
               _ptr(db.tab_ptr("cam_inst_list")), _ptr(inst_desc), _ptr(bbox), _ptr(db.tab_ptr("chains")), _ptr(bits),
               _ptr(vcam_grid), _ptr(hits), _ptr(tile_inst_base), _ptr(op("seg_off")), _ptr(seg_point_idx), _ptr(seg_xyzw),
               seg_cap, pb.max_inst_per_frame, _ptr(op("errflags")), st)
        self.launches += 1 if T else 0

        # ---- default-off: neighbour-count outlier filter -> filtered segments
        seg_off_raw = None
        if denoise is not None and I:
            radius, min_nb = float(denoise[0]), int(denoise[1])
            keep = self._buf(seg_cap, torch.uint8)
            item_first = self._buf(I + 1)
            seg_off2 = torch.zeros(I + 2, **i32)
            kept = ctypes.c_void_p(seg_off2.data_ptr() + 4)          # counts staged at seg_off2[1..]
            self._call("denoise", "cm3d_neighbor_filter", _ptr(seg_xyzw), seg_cap, _ptr(op("seg_off")), I,
                       seg_cap // MEDOID_COLS + I, radius, min_nb, _ptr(item_first), _ptr(keep), kept, st)
            self._call("denoise", "cm3d_schedule_segments", _ptr(db.tab_ptr("frame_desc")), _ptr(inst_desc), I, seg_cap,
                       _ptr(seg_off2), _ptr(op("item_off")), _ptr(item_inst), _ptr(medoid_best), _ptr(op("errflags")), st)
            seg_point_idx2 = self._buf(seg_cap)
            seg_xyzw2 = self._buf(4 * seg_cap, torch.float32)
            self._call("denoise", "cm3d_filter_segments", _ptr(seg_xyzw), _ptr(seg_point_idx), seg_cap, _ptr(op("seg_off")),
                       _ptr(keep), _ptr(seg_off2), I, _ptr(seg_xyzw2), _ptr(seg_point_idx2), _ptr(op("errflags")), st)
            self.launches += 4
            seg_off_raw = o("seg_off")[:I + 1].clone()
            o("seg_off")[:I + 1].copy_(seg_off2[:I + 1])
            seg_point_idx, seg_xyzw = seg_point_idx2, seg_xyzw2

        # ---- second phase (medoid, box search): same stream, or the lifter's medoid stream behind an event
        front = torch.cuda.current_stream(dev)
        phase2 = None
        st2 = st
        if overlap:
            if self._med_stream is None:
                self._med_stream = torch.cuda.Stream(dev)
            phase2 = self._med_stream
            ev = torch.cuda.Event()
            ev.record(front)                         # the segments are gathered: the medoid may start
            phase2.wait_event(ev)
            for t in (out, seg_point_idx, seg_xyzw, medoid_best, item_inst):
                t.record_stream(phase2)              # allocated on the front stream, read / written on the other one
            st2 = ctypes.c_void_p(phase2.cuda_stream)

        # ---- KITTI: open3d's box of the hull vertices + yaw (kitti/2d_to_3d.py:855-876,1524; M <= 3 skipped, :1479).
        # Stays on the front stream: in overlap mode it runs NEXT TO this batch's medoid (both only read the segments).
        obb = hull_info = ev_hull = None
        if want_obb is None:
            want_obb = pb.any_kitti
        if want_obb and I:
            obb = self._buf(I * 16, torch.float32)
            hull_info = self._buf(I)
            ws_words = int(N.load().cm3d_hull_obb_ws_words(seg_cap)) if self.obb_mode == 0 else 1
            hull_ws = self._buf(ws_words)
            self._call("hull_obb", "cm3d_hull_obb", _ptr(seg_xyzw), seg_cap, _ptr(op("seg_off")), I, 4, int(self.obb_mode),
                       _ptr(item_inst), _ptr(hull_ws), ws_words, _ptr(obb), _ptr(hull_info), _ptr(op("errflags")), st)
            self.launches += 1
            if phase2 is not None:
                ev_hull = torch.cuda.Event()
                ev_hull.record(front)
        self.last_hull_info = hull_info

        with torch.cuda.stream(phase2 if phase2 is not None else front):
            do = self._run_phase2(db, out, lay, seg_cap, seg_point_idx, seg_xyzw, medoid_best, item_inst, st2,
                                  want_col_sums, do_medoid, box_search)
            if phase2 is not None:
                if ev_hull is not None:
                    phase2.wait_event(ev_hull)       # `done` covers the boxes too
                do.done = torch.cuda.Event()
                do.done.record(phase2)
        do.xyzw, do.tile_cnt, do.tile_prefix, do.pix, do.hits, do.bits, do.bbox, do.seg_off_raw, do.obb = \
            xyzw, tile_cnt, tile_prefix, pix, hits, bits, bbox, seg_off_raw, obb
        return do

    _MASK_KIND = {"dense": 0, "rle": 1, "rle_str": 2}

    def _run_fused(self, db: DeviceBatch, seg_cap, want_obb, overlap: bool) -> DeviceOutputs:
        """run() as ONE C call (cm3d_lift_batch, csrc/batch.cu) over ONE workspace allocation: the same entry points in
        the same order with the same arguments, minus ~15 foreign-function calls and ~30 allocations per batch."""
        pb, dev = db.pb, self.device
        F, I, T = pb.n_frames, pb.n_inst, pb.n_tiles
        n_slots = max(T, 1) * TILE
        if seg_cap is None:
            factor = self.seg_factor if self._seg_ratio is None else min(self.seg_factor, 1.5 * self._seg_ratio + 0.05)
            seg_cap = int(factor * pb.n_raw_points) + 1024
        seg_cap = _round_up((int(seg_cap) + 3) & ~3) if overlap else (int(seg_cap) + 3) & ~3
        lay = self._out_layout(F, I)
        kind = self._MASK_KIND[pb.masks_kind]
        if want_obb is None:
            want_obb = pb.any_kitti
        want_obb = bool(want_obb and I)
        screen = self.screen_min_pts > 0
        sym = screen and not (self.screen_flags & 3)
        max_items = seg_cap // MEDOID_COLS + 2 * I
        hull_words = int(N.load().cm3d_hull_obb_ws_words(seg_cap)) if (want_obb and self.obb_mode == 0) else 1
        mask_n = int(db.mask.numel())
        I1 = max(I, 1)
        sizes = (("out", 4 * lay["_words"]), ("runs", 4 * mask_n if kind == 2 else 0), ("run_start", 4 * mask_n if kind else 0),
                 ("row_range", 8 * I1), ("bits_raw", 4 * max(pb.bits_words, 1)), ("bits", 4 * max(pb.bits_words, 1)),
                 ("bbox", 16 * I1), ("vcam_grid", 4 * max(pb.grid_words, 1)), ("xyzw", 16 * n_slots), ("tile_cnt", 4 * max(T, 1)),
                 ("tile_prefix", 4 * max(T, 1)), ("hits", 4 * n_slots), ("tile_inst_cnt", 2 * max(pb.cnt_total, 1)),
                 ("tile_inst_base", 4 * max(pb.cnt_total, 1)), ("medoid_best", 8 * I1), ("item_inst", 4 * I1),
                 ("seg_point_idx", 4 * seg_cap), ("seg_xyzw", 16 * seg_cap), ("screen_sums", 4 * seg_cap if screen else 0),
                 ("screen_min", 20 * I1 if screen else 0), ("sym_ws", 20 * seg_cap if sym else 0), ("screen_stats", 4 if screen else 0),
                 ("item_info", 16 * max(max_items, 1)), ("obb", 64 * I1 if want_obb else 0), ("hull_info", 4 * I1 if want_obb else 0),
                 ("hull_ws", 4 * hull_words if want_obb else 0))
        off, pos = {}, 0
        for name, nbytes in sizes:
            off[name] = (pos, nbytes)
            pos += (nbytes + 255) & ~255
        ws = torch.empty(_round_up(max(pos, 256)), dtype=torch.uint8, device=dev)
        base = ws.data_ptr()

        def view(name, dtype):
            a, n = off[name]
            return ws[a:a + n].view(dtype) if n else None

        front = torch.cuda.current_stream(dev)
        phase2 = None
        if overlap:
            if self._med_stream is None:
                self._med_stream = torch.cuda.Stream(dev)
            phase2 = self._med_stream
            ws.record_stream(phase2)
        a = N.BatchArgs()
        a.n_frames, a.n_inst, a.n_tiles, a.n_vcams = F, I, T, pb.n_vcams
        a.max_cells, a.max_words, a.max_runs, a.max_inst_per_frame = pb.max_cells, pb.max_words, pb.max_runs, pb.max_inst_per_frame
        a.masks_kind, a.want_obb, a.obb_mode, a.obb_min_pts = kind, int(want_obb), int(self.obb_mode), 4
        a.screen_min_pts, a.screen_flags, a.max_items = (int(self.screen_min_pts) if screen else 0), int(self.screen_flags), max_items
        a.bits_words, a.seg_cap, a.hull_ws_words, a.out_words, a.mask_bytes = max(pb.bits_words, 1), seg_cap, hull_words, lay["_words"], mask_n
        a.raw, a.mask, a.mask_off = db.raw.data_ptr(), db.mask.data_ptr(), db.mask_off.data_ptr()
        meta = db.meta.data_ptr()
        for name in ("tile_sweep", "sweep_desc", "frame_desc", "vcam_desc", "cam_inst_list", "inst_desc", "chains"):
            setattr(a, name, meta + 4 * pb.off[name])
        out_base = base + off["out"][0]
        for name in ("frame_n", "seg_off", "item_off", "medoid_local", "medoid_point_idx", "centroid", "errflags"):
            setattr(a, name, out_base + 4 * lay[name][0])
        for name, (p0, nbytes) in off.items():
            setattr(a, name, base + p0 if nbytes else None)
        a.stream = front.cuda_stream
        a.stream_medoid = phase2.cuda_stream if phase2 is not None else None
        a.launches = ctypes.addressof(self._launch_count)
        N.call("cm3d_lift_batch", ctypes.byref(a))
        self.launches += self._launch_count.value
        do = DeviceOutputs(db, view("out", torch.int32), lay, seg_cap, view("seg_point_idx", torch.int32), view("seg_xyzw", torch.float32),
                           view("xyzw", torch.float32), view("tile_cnt", torch.int32), view("tile_prefix", torch.int32))
        do.hits, do.bits, do.bbox = view("hits", torch.int32), view("bits", torch.int32), view("bbox", torch.int32)
        do.obb = view("obb", torch.float32)
        if phase2 is not None:
            do.done = torch.cuda.Event()
            do.done.record(phase2)
        self.last_hull_info = view("hull_info", torch.int32)
        self.last_screen_stats = view("screen_stats", torch.int32)
        smin = view("screen_min", torch.int32)
        self.last_screen_modes = smin[I:4 * I] if (I and smin is not None) else None
        return do

    def _run_phase2(self, db, out, lay, seg_cap, seg_point_idx, seg_xyzw, medoid_best, item_inst, st,
                    want_col_sums, do_medoid, box_search) -> DeviceOutputs:
        pb, dev = db.pb, self.device
        I = pb.n_inst
        i32 = dict(dtype=torch.int32, device=dev)

        def o(name):
            a, n = lay[name]
            return out[a:a + max(n, 1)]

        out_base = out.data_ptr()

        def op(name):                       # address of a field of the label block
            return out_base + 4 * lay[name][0]

        # ---- medoid (screen + verify for instances of >= screen_min_pts points, see csrc/medoid.cu)
        col_sums = self._buf(seg_cap, torch.float32) if want_col_sums else None
        screen_stats = screen_min = None
        if do_medoid and I:
            max_items = seg_cap // MEDOID_COLS + 2 * I
            screen = self.screen_min_pts > 0 and not want_col_sums
            screen_sums = self._buf(seg_cap, torch.float32) if screen else None
            screen_min = self._buf(5 * I) if screen else None
            sym_ws = self._buf(5 * seg_cap, torch.float32) if screen and not (self.screen_flags & 3) else None
            screen_stats = torch.zeros(1, **i32) if screen else None
            item_pos = self._buf(4 * max_items)     # item_info: {instance, q, o, m} per item
            self._call("medoid", "cm3d_medoid", _ptr(seg_xyzw), seg_cap, _ptr(op("seg_off")), _ptr(seg_point_idx),
                   _ptr(op("item_off")), _ptr(item_inst), I, max_items, _ptr(medoid_best), _ptr(col_sums),
                   _ptr(screen_sums), _ptr(screen_min), int(self.screen_min_pts) if screen else 0, int(self.screen_flags), _ptr(sym_ws),
                   _ptr(screen_stats), _ptr(item_pos),
                   _ptr(op("medoid_local")), _ptr(op("medoid_point_idx")), _ptr(op("centroid")), _ptr(op("errflags")), st)
            # expand_items, k_medoid, finalize; + classify, screen, verify; + screen_sym, screen_min; + permute
            self.launches += ((6 if self.screen_flags & 1 else (8 if self.screen_flags & 2 or sym_ws is None else 9)) +
                              (1 if sym_ws is not None and not self.screen_flags & 4 else 0)) if screen else 3   # + prune
        self.last_screen_stats = screen_stats
        self.last_screen_modes = screen_min[I:4 * I] if (do_medoid and I and screen_min is not None) else None
        box = None
        if box_search and I:
            kitti_flags = {f == "kitti" for f in pb.frame_datasets}
            if len(kitti_flags) > 1:
                raise ValueError("box_search: a batch must not mix KITTI (y up) with nuScenes/Waymo (z up) frames")
            up_axis = 1 if pb.any_kitti else 2
            box = self._buf(I * 8, torch.float32)
            self._call("box_search", "cm3d_box_search", _ptr(seg_xyzw), seg_cap, _ptr(op("seg_off")), I, up_axis,
                       int(box_search), 1, _ptr(box), _ptr(op("errflags")), st)
            self.launches += 1
        return DeviceOutputs(db, out, lay, seg_cap, seg_point_idx, seg_xyzw, None, None, None,
                             None, col_sums, None, None, None, None, box, None)

    # ------------------------------------------------------------------ device -> host
    @_on_device
    def fetch_labels(self, do: DeviceOutputs, pinned: Optional[torch.Tensor] = None) -> dict:
        """D2H of the small per-instance result block (sizes, medoids, centroids, flags)."""
        if pinned is not None:
            pinned[:do.out.numel()].copy_(do.out, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            h = pinned[:do.out.numel()].numpy()
        else:
            h = do.out.cpu().numpy()
        res = {}
        for k, v in do.layout.items():
            if k.startswith("_"):
                continue
            res[k] = h[v[0]:v[0] + v[1]]
        res["centroid"] = res["centroid"].view(np.float32).reshape(-1, 4)
        return res

    @staticmethod
    def _split_labels(h: np.ndarray, layout: dict) -> dict:
        res = {k: h[v[0]:v[0] + v[1]] for k, v in layout.items() if not k.startswith("_")}
        res["centroid"] = res["centroid"].view(np.float32).reshape(-1, 4)
        return res

    @_on_device
    def lift_frame_stream(self, frames, batch_frames: int = 32, timer: Optional[dict] = None, depth: int = 3,
                          pack_workers: int = 4):
        """Drop-in scripts' entry: an iterator of FrameSpecs in, lists of LiftResult (one list per
        batch of `batch_frames` frames, frame order kept) out.  Batches are packed into pinned
        buffers and pipelined through `lift_packed_stream`; per-instance point lists stay on the
        device (the scripts only need sizes, medoids, centroids and the KITTI yaw).  `timer`, when
        given, gets the wall time spent here under the reference's "points in mask" key."""
        import time

        part_queue = []                         # per batch: sub-frame counts of its (possibly split) frames

        def groups():
            # the first batches are small (8, 8, 16 frames): the first pack + copy is the only part of a stream
            # that nothing overlaps, so it should be short; from then on batches of batch_frames
            ramp = [min(batch_frames, n) for n in (8, 8, 16)] if batch_frames > 16 else []
            cur, parts = [], []
            for f in frames:
                sub, k = split_oversize([f], MAX_INST)
                cur.extend(sub)
                parts.extend(k)
                if len(parts) >= (ramp[0] if ramp else batch_frames):
                    if ramp:
                        ramp.pop(0)
                    part_queue.append(parts)
                    yield cur
                    cur, parts = [], []
            if cur:
                part_queue.append(parts)
                yield cur

        def batches():
            # pack the next batches (numpy copies into pinned memory release the GIL, the descriptor
            # tables do not) on worker threads while the current one is on the GPU; order is kept
            from collections import deque
            from concurrent.futures import ThreadPoolExecutor
            workers = max(1, int(pack_workers))
            with ThreadPoolExecutor(max_workers=workers) as pool:
                pending = deque()
                def take():
                    t = time.perf_counter()
                    pb = pending.popleft().result()
                    self.stream_stats["wait_pack"] = self.stream_stats.get("wait_pack", 0.0) + time.perf_counter() - t
                    return pb
                for g in groups():
                    pending.append(pool.submit(self._pack_pooled, g))
                    if len(pending) >= workers:
                        yield take()
                while pending:
                    yield take()

        # The packer threads hold the GIL while they flatten a batch (a few ms of Python per batch); with the
        # default 5 ms switch interval the launching thread would queue behind them every time one of its ~60
        # C calls per batch hands the GIL back, and the GPU would idle.  A short interval keeps the launches flowing.
        import sys
        old_interval = sys.getswitchinterval()
        sys.setswitchinterval(min(old_interval, 2e-4))
        try:
            t0 = time.time()
            for pb, do, lab in self.lift_packed_stream(batches(), depth=depth, with_handles=True):
                res = merge_split(self.results(do, lab, with_points=False), part_queue.pop(0))
                pb.release()                        # its pinned buffers go back to the pool: the copies are long done
                if timer is not None:
                    timer["points in mask"] += time.time() - t0
                yield res
                t0 = time.time()
        finally:
            sys.setswitchinterval(old_interval)

    @_on_device
    def lift_packed_stream(self, batches, seg_cap: Optional[int] = None, depth: int = 3, with_handles: bool = False):
        """Pipelined bulk path (config C5: tens of thousands of frames): yields the label dict of
        every PackedBatch in order.  Batch k+1 is copied host->device on a copy stream while batch
        k runs on the compute stream; the small result block comes back asynchronously into
        pinned memory.  A batch whose segment buffers overflow is rerun with the exact size."""
        dev = self.device
        if self._streams is None:
            # front-end stream at high priority: its blocks take the SM slots the previous batch's medoid frees
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1))
        copy_s, comp_s = self._streams
        cur = torch.cuda.current_stream(dev)
        copy_s.wait_stream(cur)             # whatever the caller queued so far comes first
        comp_s.wait_stream(cur)
        inflight = []                       # (pb, db, do, pinned, done_event)
        pool = []                           # pinned result buffers, reused (cudaHostAlloc is slow)

        import time as _t
        st = self.stream_stats

        def finish(item):
            pb, db, do, pinned, done = item
            t = _t.perf_counter()
            done.synchronize()
            st["wait_gpu"] = st.get("wait_gpu", 0.0) + _t.perf_counter() - t
            lab = self._split_labels(pinned[:do.out.numel()].numpy().copy(), do.layout)
            pool.append(pinned)
            need = self.check_flags(lab)
            if pb.n_raw_points:
                ratio = max(need, int(lab["seg_off"][-1])) / pb.n_raw_points
                self._seg_ratio = ratio if self._seg_ratio is None else max(0.95 * self._seg_ratio, ratio)
            if need:                        # rare: rerun this batch synchronously with exact capacity
                self.cap_retries += 1
                with torch.cuda.stream(comp_s):
                    do = self.run(db, seg_cap=need)
                    lab = self.fetch_labels(do)
                if self.check_flags(lab):
                    raise N.Cm3dError("segment capacity retry failed")
            return (pb, do, lab) if with_handles else lab

        for pb in batches:
            t_enq = _t.perf_counter()
            with torch.cuda.stream(copy_s):
                db = self.upload(pb)
                up = torch.cuda.Event()
                up.record(copy_s)
            with torch.cuda.stream(comp_s):
                comp_s.wait_event(up)
                do = self.run(db, seg_cap=seg_cap, overlap=True)
            with torch.cuda.stream(self._med_stream):          # behind the medoid: labels back, then the batch is done
                k = next((k for k, b in enumerate(pool) if b.numel() >= do.out.numel()), -1)
                if k < 0:
                    pinned = torch.empty(max(do.out.numel(), 1 << 14), dtype=torch.int32, pin_memory=True)
                else:
                    pinned = pool.pop(k)       # (list.remove would compare tensors element-wise)
                pinned[:do.out.numel()].copy_(do.out, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self._med_stream)
            st["enqueue"] = st.get("enqueue", 0.0) + _t.perf_counter() - t_enq
            st["batches"] = st.get("batches", 0) + 1
            inflight.append((pb, db, do, pinned, done))
            if len(inflight) >= depth:
                yield finish(inflight.pop(0))
        while inflight:
            yield finish(inflight.pop(0))

    def lift_frame_graph(self):
        """Low-latency entry for one frame (or one small batch) per call: the launch sequence is captured
        once per batch geometry as a CUDA graph over a preallocated workspace and replayed
        (cm3d_b200/graph.py).  Returns a runner: `runner.lift(frame_or_frames) -> [LiftResult]`."""
        from .graph import GraphRunner
        if self._graph_runner is None:
            self._graph_runner = GraphRunner(self)
        return self._graph_runner

    def check_flags(self, labels: dict):
        e = labels["errflags"]
        if e[1]:
            raise ValueError(f"instance {int(e[1]) - 1}: COCO run lengths do not cover the mask")
        return int(e[0])            # members needed when the segment buffers were too small, else 0

    @_on_device
    def results(self, do: DeviceOutputs, labels: dict, with_points: bool = True, with_pix: bool = False) -> List[LiftResult]:
        pb = do.db.pb
        seg_off = labels["seg_off"].astype(np.int64)
        total = int(seg_off[-1])
        idx_all = do.seg_point_idx[:total].cpu().numpy() if with_points else None
        aggr_all = tile_cnt = None
        if with_points:
            aggr_all = do.xyzw.view(4, -1).cpu().numpy()
            tile_cnt = do.tile_cnt.cpu().numpy()
        pix_all = do.pix.view(16, -1).cpu().numpy() if (with_pix and do.pix is not None) else None
        fdesc = pb.table("frame_desc", FR_WORDS)
        obb_all = do.obb.view(-1, 16).cpu().numpy() if do.obb is not None else None
        box_all = do.box.view(-1, 8).cpu().numpy() if do.box is not None else None
        raw_off = do.seg_off_raw.cpu().numpy().astype(np.int64) if do.seg_off_raw is not None else None
        out = []
        for f in range(pb.n_frames):
            i0, i1 = int(pb.frame_inst[f]), int(pb.frame_inst[f + 1])
            offs = seg_off[i0:i1 + 1] - seg_off[i0]
            r = LiftResult(
                n_points=int(labels["frame_n"][f]),
                seg_offsets=offs.astype(np.int32),
                seg_point_idx=idx_all[seg_off[i0]:seg_off[i1]].copy() if with_points else None,
                medoid_local=labels["medoid_local"][i0:i1].copy(),
                medoid_point_idx=labels["medoid_point_idx"][i0:i1].copy(),
                centroids=labels["centroid"][i0:i1, :3].copy())
            if obb_all is not None:
                r.yaw = obb_all[i0:i1, 0].copy()
                r.obb = obb_all[i0:i1].copy()
            if box_all is not None:
                r.box = box_all[i0:i1].copy()
            if raw_off is not None:
                r.raw_counts = np.diff(raw_off[i0:i1 + 1]).astype(np.int32)
            if with_points:
                tb, te = int(fdesc[f, 0]), int(fdesc[f, 1])
                keep = (np.arange(TILE)[None, :] < tile_cnt[tb:te, None]).reshape(-1)
                r.aggr_points = aggr_all[:, tb * TILE:te * TILE][:, keep]
                if pix_all is not None:
                    r.pix = pix_all[:len(pb.frame_vcam_cams[f]), tb * TILE:te * TILE][:, keep]
            out.append(r)
        return out

    @_on_device
    def lift_frames(self, frames: Sequence[FrameSpec], with_points: bool = True, with_pix: bool = False,
                    want_col_sums: bool = False, denoise=None, box_search: Optional[int] = None,
                    keep_fourth: Optional[bool] = None) -> List[LiftResult]:
        """Synchronous convenience: pack, upload, run, read back; retries once with exact
        segment capacity if the default guess was too small.  The 4th point column travels only when
        point rows are read back (keep_fourth defaults to with_points)."""
        frames, parts = split_oversize(frames, MAX_INST)
        pb = self.pack(frames, keep_fourth=with_points if keep_fourth is None else keep_fourth)
        db = self.upload(pb)
        kw = dict(want_pix=with_pix, want_col_sums=want_col_sums, denoise=denoise, box_search=box_search)
        do = self.run(db, **kw)
        labels = self.fetch_labels(do)
        need = self.check_flags(labels)
        if need:
            do = self.run(db, seg_cap=need, **kw)
            labels = self.fetch_labels(do)
            if self.check_flags(labels):
                raise N.Cm3dError("segment capacity retry failed")
        res = self.results(do, labels, with_points, with_pix)
        self.last = do
        return merge_split(res, parts) if any(k > 1 for k in parts) else res
