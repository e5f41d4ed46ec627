#!/usr/bin/env python
"""Benchmark of the 2D-mask -> 3D lifting path (BASELINE.json metric: pseudo-label frames/s on
nuScenes-shaped 10-sweep x 6-camera frames, plus achieved HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A step is one pass of the whole hot path (mask decode+erosion, sweep aggregation, projection,
membership, ordered gather, medoid) over one batch of B synthetic C2 frames per GPU.
  value  frames/s with the packed batch already resident in HBM (CUDA events, max over ranks)
  e2e    frames/s through the public API (Lifter.lift_packed_stream) from pinned HOST buffers:
         every step copies its inputs host->device and its labels device->host inside the timed
         region (the copy of step k+1 overlaps the kernels of step k on a second stream)
Under torchrun (N>1) every rank lifts its own frames (sharded by sample index, no collective
on the data path); NCCL is only used for the barrier and the max-over-ranks of the timings.
`--impl reference` times the reference's own CPU algorithm (oracle/ref_lift.py: the restated
per-frame body with the same torch calls) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pseudo_label_frames_per_s"
UNIT = "frames/s"
MASK_INPUT = "COCO counts strings (pycocotools format), decoded on the GPU"
# BASELINE.json configs.  `gen` is the synthetic generator's config name (c5 = independent C2-shaped frames),
# `batch` the frames per GPU and launch sequence, `stream` the DISTINCT frames per GPU of the streaming legs.
CONFIGS = {
    "c2": dict(gen="c5", batch=64, stream=1024, stream_total=8192,
               workload="C2: nuScenes-shaped 10 sweeps x 34,720 pts (~347k pts) x 6 cams 1024x576 masks x 50 instances per frame"),
    "c1": dict(gen="c1", batch=64, stream=512, stream_total=32768,
               workload="C1: nuScenes-shaped single sample, 1 sweep (~34.7k pts) x 6 cams 1024x576 masks x 20 instances per frame"),
    "c3": dict(gen="c3", batch=32, stream=256, stream_total=8192,
               workload="C3: KITTI-shaped 64-beam (~120k pts) x 1 cam, 1024x309 masks x 15 instances per frame"),
    "c4": dict(gen="c4", batch=16, stream=128, stream_total=8192,
               workload="C4: Waymo-shaped top LiDAR (~180k pts) x 5 cams, 1024x683 / 1024x473 masks x 80 instances per frame"),
}


def _gen_frame(arg):
    gen, index = arg
    from cm3d_b200 import synthetic as S
    f = S.make_frame(gen, index, dense_masks=False)
    f.masks = S.compress_rles(f.masks)                       # masks as in {f}_masks.pkl: COCO counts strings
    return f


def make_frames(gen, first, count, workers):
    idx = [(gen, i) for i in range(first, first + count)]
    if workers <= 1 or count <= 2:
        return [_gen_frame(i) for i in idx]
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(workers) as pool:
        return pool.map(_gen_frame, idx, chunksize=max(1, count // (8 * workers)))


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML from a thread (no process
    spawn, sub-millisecond queries), `nvidia-smi -lms` as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu, self.nvml, self.stop_flag = [], None, gpu_index, None, False

    def _nvml_loop(self):
        import pynvml as nv
        h = self.nvml
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), float(mx), int(mask)))
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self.nvml = nv.nvmlDeviceGetHandleByIndex(idx)
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            sm = sorted(r[0] for r in self.rows)
            reasons = sorted({name for r in self.rows for bit, name in self.REASONS if r[2] & bit})
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.rows[-1][1] if self.rows else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(frames, threads):
    """The reference's per-frame body (oracle/ref_lift.py, torch CPU) over `frames`; (seconds, results)."""
    import torch
    from oracle import ref_lift as RL
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    out = [RL.lift_frame(f, record_pix=False) for f in frames]
    return time.perf_counter() - t0, out


def _ref_worker(arg):
    """One frame through the CPU port on ONE thread (frame-parallel CPU baseline); returns seconds."""
    import torch
    torch.set_num_threads(1)
    from oracle import ref_lift as RL
    f = _gen_frame(arg)
    RL.lift_frame(_gen_frame_small(), record_pix=False)          # page in torch, untimed
    t0 = time.perf_counter()
    RL.lift_frame(f, record_pix=False)
    return time.perf_counter() - t0


def _gen_frame_small():
    from cm3d_b200 import synthetic as S
    return S.make_frame("c1", 0, scale=0.1)


def cpu_frame_parallel(gen, n_procs):
    """Frames/s of the CPU port run one process per core, one frame each, all at the same time."""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(n_procs) as pool:
        secs = pool.map(_ref_worker, [(gen, i) for i in range(100000, 100000 + n_procs)], chunksize=1)
    return sum(1.0 / s for s in secs), secs


def run_reference(args, rank):
    if rank != 0:
        return
    import torch
    cfg = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    nf = max(1, args.ref_frames)
    frames = make_frames(cfg["gen"], 0, nf, 1)
    for _ in range(args.warmup):
        cpu_reference_run(frames[:1], threads)
    times = [cpu_reference_run(frames, threads)[0] for _ in range(args.steps)]
    total = sum(times)
    v = nf * args.steps / total
    sample = f"{nf} {args.config.upper()} frame(s) per step x {args.steps} steps, oracle/ref_lift.py (torch {torch.__version__} CPU)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "mask_input": MASK_INPUT},         # our arm's workload, a bounded sample of it per step
        "run": {"frames_per_step": nf, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def _frame_labels(lab, pb, k):
    """(member counts, medoid point index, centroids) of frame k of a packed batch from its label block."""
    import numpy as np
    i0, i1 = int(pb.frame_inst[k]), int(pb.frame_inst[k + 1])
    so = lab["seg_off"].astype(np.int64)
    return np.diff(so[i0:i1 + 1]), lab["medoid_point_idx"][i0:i1].copy(), lab["centroid"][i0:i1, :3].copy()


def _labels_equal_oracle(gpu, o):
    """GPU labels of one frame against the CPU oracle's (oracle/ref_lift.py): member counts, medoid point
    index, centroid bit patterns (the medoid is a copy of an input point, so tolerance 0)."""
    import numpy as np
    cnt, mpi, cen = gpu
    want_cnt = np.array([len(x) for x in o["idx"]], np.int64)
    has = o["medoid_local"] >= 0
    return bool(np.array_equal(cnt, want_cnt) and np.array_equal(mpi.astype(np.int64), o["medoid_point_idx"].astype(np.int64)) and
                np.array_equal(cen[has].view(np.uint32), o["centroids"][has].astype(np.float32).view(np.uint32)) and
                bool(np.isnan(cen[~has]).all()))


def _disk_leg(frames, lifter, n_sweeps, reader_threads=8):
    """The drop-in nuScenes script (src/nuscenes/2d_to_3d.py -> nuscenes_stage.run) over an on-disk synthetic
    dataset: .bin sweeps, {f}_masks.pkl, {f}_data.json in, pseudolabels JSON out; seconds per frame include
    every file read, the packer, H2D, kernels, pass 2, NMS and the JSON write."""
    import importlib.util
    import shutil
    import tempfile
    from cm3d_b200 import synthetic_datasets as SD
    per_scene = 32
    n_scenes = max(1, min(4, len(frames) // per_scene))
    scenes = {f"scene-{k:04d}": frames[k * per_scene:(k + 1) * per_scene] for k in range(n_scenes)}
    work = tempfile.mkdtemp(prefix="cm3d_disk_")
    try:
        t_w = time.perf_counter()
        nusc, map_factory = SD.write_nuscenes(os.path.join(work, "nusc"), os.path.join(work, "masks"), scenes, ratio=0.64)
        t_w = time.perf_counter() - t_w
        spec = importlib.util.spec_from_file_location("nusc_script_bench", os.path.join(ROOT, "src", "nuscenes", "2d_to_3d.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.INPUT_PATH, mod.INPUT_DIR, mod.OUTPUT_DIR = os.path.join(work, "nusc"), os.path.join(work, "masks"), os.path.join(work, "out")
        mod.n_sweeps, mod.BATCH_FRAMES, mod.READER_THREADS = n_sweeps, 32, reader_threads
        import contextlib
        import io
        names = list(scenes)
        with contextlib.redirect_stdout(io.StringIO()):
            mod.main(nusc, map_factory, names[:1], lifter=lifter)               # warm-up: one scene
            t0 = time.perf_counter()
            final = mod.main(nusc, map_factory, names, lifter=lifter)
            dt = time.perf_counter() - t0
        n_frames = n_scenes * per_scene
        n_boxes = sum(len(v) for v in final["results"].values())
        nbytes = sum(os.path.getsize(os.path.join(dp, fn)) for dp, _, fns in os.walk(work) for fn in fns)
        return {"value": n_frames / dt, "unit": UNIT, "frames": n_frames, "scenes": n_scenes, "boxes_after_nms": n_boxes,
                "dataset_bytes": nbytes, "dataset_write_seconds": round(t_w, 2), "reader_threads": reader_threads,
                "note": "src/nuscenes/2d_to_3d.py end to end on an on-disk synthetic nuScenes tree (page-cache warm): devkit record "
                        "lookups, np.fromfile of the sweeps, pickle/json of the masks on reader threads, C packer, H2D, kernels, "
                        "lane lookup, pass 2, circle NMS and the JSON write inside the timed region"}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def _disk_leg_kitti(frames, lifter, reader_threads=8):
    """The drop-in KITTI script (src/kitti/2d_to_3d.py -> kitti_stage.run) over an on-disk synthetic KITTI object tree:
    velodyne .bin, calib .txt, {f}_masks.pkl, {f}_data.json in; pred/ and pseudo/ label files out."""
    import contextlib
    import importlib.util
    import io
    import shutil
    import tempfile
    from cm3d_b200 import synthetic_datasets as SD
    n = min(len(frames), 128)
    work = tempfile.mkdtemp(prefix="cm3d_disk_")
    try:
        root, input_dir = os.path.join(work, "kitti"), os.path.join(work, "masks")
        SD.write_kitti(root, input_dir, frames[:n])
        spec = importlib.util.spec_from_file_location("kitti_script_bench", os.path.join(ROOT, "src", "kitti", "2d_to_3d.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.INPUT_PATH, mod.INPUT_DIR = root, input_dir
        mod.PRED_DIR, mod.PSEUDO_DIR = os.path.join(work, "pred"), os.path.join(work, "pseudo")
        mod.NUM_SAMPLES, mod.BATCH_FRAMES, mod.READER_THREADS = n, 32, reader_threads
        with contextlib.redirect_stdout(io.StringIO()):
            mod.main(frame_range=range(min(n, 32)), lifter=lifter)                # warm-up
            t0 = time.perf_counter()
            written = mod.main(lifter=lifter)
            dt = time.perf_counter() - t0
        nbytes = sum(os.path.getsize(os.path.join(dp, fn)) for dp, _, fns in os.walk(work) for fn in fns)
        return {"value": n / dt, "unit": UNIT, "frames": n, "objects_written": int(written), "dataset_bytes": nbytes,
                "reader_threads": reader_threads,
                "note": "src/kitti/2d_to_3d.py end to end on an on-disk synthetic KITTI tree (page-cache warm): velodyne / calib / mask "
                        "files read on reader threads, C packer, H2D, kernels (hull boxes included), two label files per frame written"}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def _script_leg_waymo(frames, lifter, reader_threads=8):
    """The drop-in Waymo script (src/waymo/2d_to_3d.py -> waymo_stage.run) over synthetic segments: mask files on disk,
    the LiDAR frames as parsed `dataset_pb2.Frame`-shaped objects (reading TFRecords needs tensorflow, absent here);
    metrics.Objects file out."""
    import contextlib
    import importlib.util
    import io
    import shutil
    import tempfile
    from cm3d_b200 import synthetic_datasets as SD
    import copy
    from cm3d_b200 import boxes as BX
    per_scene = 32
    n_scenes = max(1, min(4, len(frames) // per_scene))
    relabelled = []
    for f in frames[:n_scenes * per_scene]:      # the reference raises on classes without a Waymo type (waymo:1060-1061): none here
        g = copy.copy(f)
        g.labels = [l if BX.NUSC_TO_WAYMO.get(BX.get_detection_name(l), "") else "car" for l in f.labels]
        relabelled.append(g)
    frames = relabelled
    work = tempfile.mkdtemp(prefix="cm3d_disk_")
    try:
        input_dir = os.path.join(work, "masks")
        scenes = [(f"segment-{k:03d}", SD.waymo_frames(f"segment-{k:03d}", input_dir, frames[k * per_scene:(k + 1) * per_scene]))
                  for k in range(n_scenes)]
        spec = importlib.util.spec_from_file_location("waymo_script_bench", os.path.join(ROOT, "src", "waymo", "2d_to_3d.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.INPUT_DIR, mod.OUTPUT_FILE = input_dir, os.path.join(work, "out", "pred.bin")
        mod.BATCH_FRAMES, mod.READER_THREADS = 16, reader_threads
        points_fn = lambda fr: fr.points_vehicle
        with contextlib.redirect_stdout(io.StringIO()):
            mod.main(scenes[:1], points_fn, lifter=lifter)                      # warm-up: one segment
            t0 = time.perf_counter()
            final = mod.main(scenes, points_fn, lifter=lifter)
            dt = time.perf_counter() - t0
        n = n_scenes * per_scene
        return {"value": n / dt, "unit": UNIT, "frames": n, "segments": n_scenes, "objects_after_nms": len(final),
                "reader_threads": reader_threads,
                "note": "src/waymo/2d_to_3d.py end to end: mask files from disk on reader threads, calibration -> FrameSpec, C packer, "
                        "H2D, kernels, vehicle -> global, lane lookup, pass 2, per-timestamp NMS and the Objects file inside the timed region"}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def run_ours(args, rank, world, local_rank):
    cfg = CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    n_stream = max(B, (args.stream_frames or cfg["stream"]) // B * B)
    # frames first (spawned workers), CUDA afterwards
    workers = max(1, min(args.workers or (os.cpu_count() or 1) // max(world, 1), 32))
    t_gen = time.perf_counter()
    frames = make_frames(cfg["gen"], rank * n_stream, n_stream, workers)
    t_gen = time.perf_counter() - t_gen

    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cm3d_b200.lifter import Lifter

    lifter = Lifter(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident batches: up to 8 DISTINCT batches of B frames live in HBM; a step lifts one of them
    n_res = max(1, min(8, n_stream // B))
    pbs = [lifter.pack(frames[k * B:(k + 1) * B], keep_fourth=False) for k in range(n_res)]
    dbs = [lifter.upload(pb) for pb in pbs]
    torch.cuda.synchronize()
    first_labels, caps, n_points, seg_totals, pairs = [], [], [], [], []
    for db in dbs:                      # segment capacity of every resident batch (exact, found once, untimed)
        do = lifter.run(db)
        lab = lifter.fetch_labels(do)
        need = lifter.check_flags(lab)
        if need:
            do = lifter.run(db, seg_cap=need)
            lab = lifter.fetch_labels(do)
            assert lifter.check_flags(lab) == 0
        first_labels.append(lab)
        caps.append(int(lab["seg_off"][-1]))
        n_points.append(int(lab["frame_n"].sum()))
        m = np.diff(lab["seg_off"].astype(np.int64))
        seg_totals.append(int(m.sum()))
        pairs.append(float((m.astype(np.float64) ** 2).sum()))
        del do
    seg_cap = max(caps) + 4096
    for w in range(max(args.warmup, 3)):
        lifter.run(dbs[w % n_res], seg_cap=seg_cap)
    barrier()

    # ---- device-resident timed region: CUDA events on the launch stream, nothing else inside
    sampler = ClockSampler(local_rank)
    sampler.start()
    lifter.timing = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    overlap = not args.no_overlap
    cur = torch.cuda.current_stream(dev)
    # the front end (masks, aggregation, projection, gather: HBM-bound) is launched on a high-priority stream and
    # the medoid (XU-bound) on the lifter's second stream, so step k+1's front end runs next to step k's medoid
    front = torch.cuda.Stream(dev, priority=-1) if overlap else cur
    import collections
    # the last few steps' buffers (--ring) stay referenced (as in the streaming path's pipeline): a workspace that is
    # freed while the other stream still uses it cannot be reused, and the allocator would cudaMalloc a new one
    ring = collections.deque(maxlen=max(1, args.ring) if overlap else 1)
    # warm-up: every resident batch at least once in the timed configuration (their workspace sizes differ, and the
    # caching allocator must have seen them all), and never fewer steps than asked for
    for w in range(max(args.warmup, n_res + ring.maxlen + 1)):
        with torch.cuda.stream(front):
            ring.append(lifter.run(dbs[w % n_res], seg_cap=seg_cap, overlap=overlap))
    barrier()
    lifter.launches = 0                     # counts the timed region only
    e0.record(cur)
    front.wait_stream(cur)
    t_host = time.perf_counter()
    with torch.cuda.stream(front):
        for k in range(args.steps):
            if overlap and len(ring) == ring.maxlen:
                ring[0].done.synchronize()      # the host stays at most `ring` steps ahead (the GPU always has work queued):
                #                                 buffers are then freed AFTER their last use and recycled without cudaMalloc
            ring.append(lifter.run(dbs[k % n_res], seg_cap=seg_cap, overlap=overlap))
    t_host = (time.perf_counter() - t_host) / args.steps * 1e3
    do = ring[-1]
    if overlap:
        cur.wait_event(do.done)             # the medoid stream is in order: the last step's event covers every step
    cur.wait_stream(front)
    e1.record(cur)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = lifter.launches
    final = lifter.fetch_labels(do)
    assert lifter.check_flags(final) == 0
    last = (args.steps - 1) % n_res
    assert np.array_equal(final["medoid_point_idx"], first_labels[last]["medoid_point_idx"])       # deterministic labels
    del do
    ring.clear()

    # ---- per-kernel CUDA events: a separate pass over every resident batch (not inside the headline region)
    lifter.timing = {}
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for db in dbs:
        do = lifter.run(db, seg_cap=seg_cap)
        del do
    p1.record()
    torch.cuda.synchronize()
    pass_ms = p0.elapsed_time(p1) / n_res
    timing = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in lifter.timing.items()}
    lifter.timing = None
    do = lifter.run(dbs[0], seg_cap=seg_cap)
    screen_modes = lifter.last_screen_modes.cpu().numpy() if lifter.last_screen_modes is not None else None
    screen_verified = int(lifter.last_screen_stats.item()) if lifter.last_screen_stats is not None else None
    del do

    # ---- host->device copy rate of one packed batch (explains e2e when PCIe, not the kernels, bounds it)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lifter.upload(pbs[0])
    torch.cuda.synchronize()
    h0.record()
    for _ in range(3):
        lifter.upload(pbs[0])
    h1.record()
    torch.cuda.synchronize()
    h2d_gbps = 3 * pbs[0].h2d_bytes / (h0.elapsed_time(h1) * 1e-3) / 1e9

    # ---- end to end: pinned host buffers -> H2D -> kernels -> D2H of the labels, every step, through the public
    # streaming API.  A step's B frames travel as sub-batches (same frames, same order): the copy of sub-batch
    # k+1 overlaps the kernels of sub-batch k; segment capacity is the lifter's own estimate (a batch that does
    # not fit is rerun inside the timed region)
    sub = max(1, min(args.e2e_sub or B, B))
    if sub < B:
        subs = [[lifter.pack(frames[k * B + i:k * B + i + sub], keep_fourth=False) for i in range(0, B, sub)] for k in range(n_res)]
    else:
        subs = [[pb] for pb in pbs]
    # pipeline ramp-up: the very first copy of a stream is the only one nothing overlaps, so the first step's batch
    # travels in small pieces (same frames, same order); every later step is one whole batch
    ramp = max(0, min(args.e2e_ramp, B))
    first = [lifter.pack(frames[i:i + ramp], keep_fourth=False) for i in range(0, B, ramp)] if 0 < ramp < B else subs[0]
    seq = lambda steps: [p for k in range(steps) for p in (first if k == 0 else subs[k % n_res])]
    lifter.cap_retries = 0
    for lab in lifter.lift_packed_stream(seq(max(3, min(n_res, 8)))):
        pass
    barrier()
    retries0 = lifter.cap_retries
    t0 = time.perf_counter()
    labs = []
    for lab in lifter.lift_packed_stream(seq(args.steps)):
        labs.append(lab["medoid_point_idx"])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    per = len(subs[0])
    assert np.array_equal(np.concatenate(labs[-per:]), first_labels[last]["medoid_point_idx"])
    d2h_bytes = sum(int(lifter._out_layout(p.n_frames, p.n_inst)["_words"]) * 4 for p in subs[0])
    h2d_bytes = sum(p.h2d_bytes for p in subs[0])
    e2e_retries = lifter.cap_retries - retries0

    # ---- C5-style stream: every one of this rank's DISTINCT FrameSpecs (what the drop-in scripts hand over: numpy
    # sweeps + counts strings + calibration), cycled until the timed region is seconds long - host packing, pinned
    # pool reuse, capacity estimates and retries included
    fs = None
    if not args.no_framespec_leg:
        # measured on the 16-vCPU box: 6 packer threads feed the GPU (3.8 k frames/s); 8 or 12 slow the launching thread down
        pw = max(2, min(args.pack_workers or 6, (os.cpu_count() or 2) // max(world, 1) - 1))
        fs_kw = dict(batch_frames=args.stream_batch, pack_workers=pw)
        import itertools
        warm = max(16 * args.stream_batch, 256)                 # fills the pinned-buffer pool and the allocator's size classes
        for _ in lifter.lift_frame_stream(itertools.islice(itertools.cycle(frames), warm), **fs_kw):
            pass
        barrier()
        cycles = max(1, args.stream_cycles or -(-cfg["stream_total"] // n_stream))
        lifter.stream_stats.clear()
        r0 = lifter.cap_retries
        t0 = time.perf_counter()
        n_fs = n_boxes = 0
        for res in lifter.lift_frame_stream(itertools.chain.from_iterable(frames for _ in range(cycles)), **fs_kw):     # ONE stream
            n_fs += len(res)
            n_boxes += sum(int((r.medoid_local >= 0).sum()) for r in res)
        torch.cuda.synchronize()
        fs = {"seconds": time.perf_counter() - t0, "frames": n_fs, "centroids": n_boxes, "retries": lifter.cap_retries - r0,
              "pack_workers": pw, "cycles": cycles, "stats": {k: round(v, 3) for k, v in lifter.stream_stats.items()},
              "pinned_allocations": lifter._pin_pool.allocations if lifter._pin_pool is not None else None}

    # ---- parity witnesses
    parity = {}
    if world > 1:       # one sampled frame of another rank, recomputed alone on rank 0
        src = world - 1
        mine = np.full(256, -2, np.int32)
        mp_ = _frame_labels(first_labels[0], pbs[0], 0)[1]
        mine[:min(256, mp_.size)] = mp_[:256]
        t_all = [torch.empty(256, dtype=torch.int32, device=dev) for _ in range(world)]
        dist.all_gather(t_all, torch.from_numpy(mine).to(dev))
        if rank == 0:
            f_src = _gen_frame((cfg["gen"], src * n_stream))
            r = lifter.lift_frames([f_src], with_points=False)[0]
            want = np.full(256, -2, np.int32)
            want[:min(256, r.medoid_point_idx.size)] = r.medoid_point_idx[:256]
            ok = bool(np.array_equal(t_all[src].cpu().numpy(), want))
            parity["cross_rank"] = {"rank": src, "frame": src * n_stream, "equal": ok,
                                    "what": "medoid point indices of that rank's first frame == a single-frame recomputation on rank 0"}
            assert ok, "cross-rank parity witness failed"

    t = torch.tensor([dev_ms, e2e_s * 1e3, fs["seconds"] * 1e3 if fs else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, fs_ms = float(t[0]), float(t[1]), float(t[2])
    cnt = torch.tensor([fs["frames"] if fs else 0, fs["retries"] if fs else 0, e2e_retries], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    fs_frames_all, fs_retries_all, e2e_retries_all = (int(v) for v in cnt)

    if rank == 0:
        total_frames = B * world * args.steps
        value = total_frames / (dev_ms * 1e-3)
        e2e = total_frames / (e2e_ms * 1e-3)
        peak, peak_src = measured_peak()
        mean = lambda v: float(np.mean(v))
        raw_bytes = mean([int(pb.raw.nbytes) for pb in pbs])
        npts, segt = mean(n_points), mean(seg_totals)
        bits_words = mean([pb.bits_words for pb in pbs])
        mask_bytes = mean([int(pb.mask.nbytes) for pb in pbs])
        algo = {   # algorithmic bytes per launch (DESIGN.md "Kernels and rooflines"), mean over the resident batches
            "aggregate": raw_bytes + 12 * npts,
            "project_count": 16 * npts,
            "compact": 4 * npts + 28 * segt,          # hit words in; per member: xyz in, index + xyz out
            "masks_decode": mask_bytes * 5,
            "masks_rle": mask_bytes * 4 + 4 * bits_words,
            "masks_erode": 8 * bits_words,
        }
        step_ms = dev_ms / args.steps
        kern = {}
        for k, ms in timing.items():
            kern[k] = {"ms": ms, "share": ms / pass_ms}
            if k in algo:
                kern[k]["algo_bytes"] = int(algo[k])
                kern[k]["gbps"] = algo[k] / (ms * 1e-3) / 1e9
        hbm_k = max((k for k in kern if k in ("aggregate", "project_count", "compact")), key=lambda k: kern[k]["ms"])
        traffic, traffic_src, erode_dram = None, None, None
        try:        # DRAM bytes per launch from the newest committed ncu capture of this same command
            for fn in sorted((fn for fn in os.listdir(os.path.join(ROOT, "profiles")) if fn.endswith("_traffic.json")), reverse=True):
                with open(os.path.join(ROOT, "profiles", fn)) as f:
                    tj = json.load(f)
                if tj.get("frames_per_launch") == B and tj.get("config", "c2") == args.config:
                    traffic, traffic_src = tj["dram_bytes_per_launch"].get("k_" + hbm_k), tj["source"]
                    erode_dram = tj["dram_bytes_per_launch"].get("k_erode3x3")
                    break
        except Exception:
            pass
        if erode_dram and "masks_erode" in kern:     # rows without set pixels are written unread: report the DRAM rate too
            kern["masks_erode"]["dram_bytes_ncu"] = erode_dram
            kern["masks_erode"]["dram_gbps"] = erode_dram / (kern["masks_erode"]["ms"] * 1e-3) / 1e9
        roof = {"kernel": hbm_k, "bound": "hbm", "achieved": kern[hbm_k]["gbps"], "peak": peak, "unit": "GB/s",
                "frac": kern[hbm_k]["gbps"] / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algo_bytes": int(algo[hbm_k]), "peak_source": peak_src,
                "timing": "CUDA events around the C-ABI call on its launch stream, mean over a separate pass across the resident batches"}
        if "medoid" in timing:
            prs = mean(pairs)
            kern["medoid"]["pair_distances"] = prs
            kern["medoid"]["gpairs_per_s"] = prs / (timing["medoid"] * 1e-3) / 1e9
            # XU-pipe ceiling: one MUFU square root per EVALUATED pair and the XU pipe retires
            # 16 lanes per SM and clock.  Instances whose squared-distance matrix is exactly symmetric
            # (mode 2) are screened over the pairs i <= j of 256-column strips: about half the roots.
            n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
            mhz = clocks.get("sm_mhz") or 1965.0
            ceil_roots = n_sm * 16 * mhz * 1e6 / 1e9
            m0 = np.diff(first_labels[0]["seg_off"].astype(np.int64))
            roots = float((m0.astype(np.float64) ** 2).sum())
            if screen_modes is not None:
                I_ = m0.size
                modes, g0, g1 = screen_modes[:I_], screen_modes[I_:2 * I_].astype(np.float64), screen_modes[2 * I_:3 * I_].astype(np.float64)
                mm = m0.astype(np.float64)

                def tri(x):      # roots of a symmetric strip sweep over x points: pairs i <= j by 256-column strips
                    fb = np.floor(x / 256.0)
                    return 256.0 * 256.0 * fb * (fb + 1) / 2.0 + (x - 256.0 * fb) * x
                g2 = np.maximum(mm - g0 - g1, 0.0)
                grouped = tri(g0) + tri(g2) + (mm * mm - g0 * g0 - g2 * g2)      # both groups symmetric, the rest in both orders
                roots = float(np.where(modes == 2, tri(mm), np.where(modes == 3, grouped, np.where(modes == 4, g0 * mm, mm * mm))).sum())
                kern["medoid"]["instances_by_mode"] = {"exact": int((modes == 0).sum()), "screen_all_pairs": int((modes == 1).sum()),
                                                       "screen_symmetric": int((modes == 2).sum()),
                                                       "screen_grouped_symmetric": int((modes == 3).sum()),
                                                       "screen_pruned_columns": int((modes == 4).sum())}
                if (modes == 4).any():
                    kern["medoid"]["pruned_columns_kept"] = float(g0[modes == 4].sum() / mm[modes == 4].sum())
            roots *= prs / max(pairs[0], 1.0)        # batch 0's root count scaled to the mean batch
            kern["medoid"]["bound"] = ("XU pipe: one MUFU.SQRT per evaluated pair distance in the screen pass (reads only "
                                       "sum M points, L2-resident); symmetric instances evaluate the pairs i <= j only")
            kern["medoid"]["square_roots"] = roots
            kern["medoid"]["groots_per_s"] = roots / (timing["medoid"] * 1e-3) / 1e9
            kern["medoid"]["peak_groots_per_s"] = ceil_roots
            kern["medoid"]["frac"] = kern["medoid"]["groots_per_s"] / ceil_roots
            if screen_verified is not None:
                kern["medoid"]["screen_min_pts"] = lifter.screen_min_pts
                kern["medoid"]["verified_columns_per_step"] = screen_verified
                kern["medoid"]["instances_per_step"] = int(m0.size)
        inter_mb = 4 * 5 * mean([pb.n_tiles for pb in pbs]) * 1024 / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "mask_input": MASK_INPUT,
                       "l2": f"a step reads one of {n_res} distinct resident batches: {pbs[0].h2d_bytes / 1e6:.0f} MB inputs + {inter_mb:.0f} MB "
                             f"intermediates per step > 126 MB L2, no explicit flush",
                       "parallelism": f"frames sharded by sample index, {world} process(es), no collective"},
            "run": {"warmup_steps_done": max(args.warmup, n_res + max(1, args.ring) + 1), "frames_per_gpu_per_step": B, "distinct_resident_batches_per_gpu": n_res, "distinct_frames_per_gpu": n_stream,
                    "points_per_step": int(npts), "member_points_per_step": int(segt),
                    "point_columns_shipped": "x, y, z (the 4th column never reaches a label: nuscenes:645,656)",
                    "streams": ("front end of step k+1 (high-priority stream) overlaps the medoid of step k (second stream)"
                                if overlap else "one stream")},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": d2h_bytes * world, "ms_per_step": e2e_ms / args.steps,
                    "h2d_gbps_rank0": round(h2d_gbps, 1), "sub_batch_frames": sub, "first_step_piece_frames": ramp or B,
                    "capacity_retries": e2e_retries_all,
                    "api": "Lifter.lift_packed_stream (pinned host buffers in, label block out), distinct batches, the lifter's own capacity estimate"},
            "gpu_launches": launches,
            "host_ms_per_step_rank0": round(t_host, 3),      # launch-side time of a step in the device-timed loop
            "e2e_from_framespecs": None if fs is None else {
                "value": fs_frames_all / (fs_ms * 1e-3), "unit": UNIT, "frames": fs_frames_all, "seconds": fs_ms * 1e-3,
                "distinct_frames_per_gpu": n_stream, "cycles": fs["cycles"], "capacity_retries": fs_retries_all,
                "pack_workers_per_gpu": fs["pack_workers"], "vs_value": fs_frames_all / (fs_ms * 1e-3) / value,
                "rank0_seconds_waiting": fs["stats"], "rank0_pinned_allocations": fs["pinned_allocations"],
                "note": "Lifter.lift_frame_stream over every distinct FrameSpec of the rank (numpy sweeps + counts strings + calibration), "
                        "cycled: C packer (csrc/pack.cu, GIL released) on worker threads into pooled pinned buffers, "
                        "H2D / kernels / D2H pipelined; wall clock, max over ranks"},
            "roofline": roof,
            "path_hbm": {"algorithmic_bytes_per_step": int(sum(algo.values())) * world,
                         "achieved_gbps": sum(algo.values()) * world / (step_ms * 1e-3) / 1e9,
                         "frac_of_peak_per_gpu": sum(algo.values()) / (step_ms * 1e-3) / 1e9 / peak,
                         "note": "sum of the kernels' algorithmic bytes over the whole step; the step is bound by the "
                                 "medoid's square roots (XU pipe), not by HBM (kernels.medoid)"},
            "kernels": kern,
            "kernel_pass_ms_per_step": pass_ms,
            "gen_seconds": round(t_gen, 1),
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            nf = max(1, args.ref_frames)
            from cm3d_b200 import synthetic as S
            cpu_reference_run([S.make_frame("c1", 0, scale=0.25)], threads)      # spin up the torch thread pool
            secs, outs = cpu_reference_run(frames[:nf], threads)
            line["cpu_baseline"] = {"value": nf / secs, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{nf} of this run's {args.config.upper()} frames through oracle/ref_lift.py "
                                              f"(torch {torch.__version__} CPU, {threads} threads), after a small warm-up frame"}
            eq = [_labels_equal_oracle(_frame_labels(first_labels[0], pbs[0], k), outs[k]) for k in range(nf)]
            parity["oracle"] = {"frames": nf, "instances": int(sum(len(o["idx"]) for o in outs)), "equal": bool(all(eq)),
                                "what": "GPU labels of the cpu_baseline frames == oracle/ref_lift.py: member counts, medoid point index, centroid bits"}
            assert all(eq), "bench frames: GPU labels differ from the CPU oracle"
            if not args.no_frame_parallel:
                try:      # one single-threaded process per core, one frame each, concurrently
                    fp, per = cpu_frame_parallel(cfg["gen"], threads)
                    line["cpu_baseline"]["frame_parallel"] = {
                        "value": fp, "unit": UNIT, "processes": threads,
                        "sample": f"{threads} {args.config.upper()} frames, one per process, 1 torch thread each, run concurrently; "
                                  f"sum of 1/seconds (mean {sum(per) / len(per):.1f} s per frame)"}
                except Exception as e:
                    line["cpu_baseline"]["frame_parallel"] = {"error": repr(e)[:200]}
        line["parity_checked"] = parity
        if world == 1 and args.config in ("c2", "c3", "c4") and not args.no_disk_leg:
            try:
                leg = {"c2": lambda: _disk_leg(frames, lifter, n_sweeps=10, reader_threads=args.reader_threads),
                       "c3": lambda: _disk_leg_kitti(frames, lifter, reader_threads=args.reader_threads),
                       "c4": lambda: _script_leg_waymo(frames, lifter, reader_threads=args.reader_threads)}[args.config]
                line["e2e_from_disk"] = leg()
            except Exception as e:
                line["e2e_from_disk"] = {"error": repr(e)[:300]}
        if world == 1 and not args.no_latency_leg:
            try:
                line["latency_batch1"] = _latency_leg(frames, lifter)
            except Exception as e:
                line["latency_batch1"] = {"error": repr(e)[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _latency_leg(frames, lifter):
    """One frame per call (BASELINE config 1's "single sample"): host FrameSpec in, labels out, synchronous."""
    import numpy as np
    import torch
    fr = frames[:16]
    for f in fr[:4]:
        lifter.lift_frames([f], with_points=False, keep_fourth=False)
    torch.cuda.synchronize()
    ts = []
    for f in fr:
        t0 = time.perf_counter()
        lifter.lift_frames([f], with_points=False, keep_fourth=False)
        ts.append(time.perf_counter() - t0)
    out = {"ms_median": 1e3 * float(np.median(ts)), "ms_min": 1e3 * float(np.min(ts)), "frames": len(ts),
           "api": "Lifter.lift_frames([frame]) - pack, H2D, launch sequence, D2H, synchronous"}
    g = lifter.lift_frame_graph()
    for f in fr[:4]:
        g.lift(f)
    tg = []
    for f in fr:
        t0 = time.perf_counter()
        res = g.lift(f)
        tg.append(time.perf_counter() - t0)
    ref = lifter.lift_frames([fr[-1]], with_points=False, keep_fourth=False)[0]
    assert np.array_equal(res[0].medoid_point_idx, ref.medoid_point_idx) and np.array_equal(res[0].seg_offsets, ref.seg_offsets)
    out["graph_ms_median"] = 1e3 * float(np.median(tg))
    out["graph_ms_min"] = 1e3 * float(np.min(tg))
    out["graphs_captured"] = g.captures
    out["graph_api"] = ("Lifter.lift_frame_graph().lift(frame): C packer, 4 H2D copies into static inputs, the launch sequence "
                        "replayed as one CUDA graph over a preallocated workspace, label block back; one graph per batch geometry")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json config (default: the one the metric is quoted on)")
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (0 = the config's default: 64 for C2)")
    ap.add_argument("--stream-frames", type=int, default=0, help="distinct frames per GPU of the streaming legs (0 = config default: 1024 for C2)")
    ap.add_argument("--stream-cycles", type=int, default=0, help="passes over the distinct frames in the FrameSpec leg (0 = config default)")
    ap.add_argument("--e2e-sub", type=int, default=0, help="frames per pipelined sub-batch of the end-to-end leg (0 = the whole batch; "
                    "32 measured 2.8 %% slower on one GPU: two launch sequences per step have two medoid tails)")
    ap.add_argument("--e2e-ramp", type=int, default=16, help="frames per piece of the FIRST step of the end-to-end leg (0 = no ramp-up)")
    ap.add_argument("--pack-workers", type=int, default=0)
    ap.add_argument("--stream-batch", type=int, default=40, help="frames per launch sequence of the FrameSpec stream leg")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-frames", type=int, default=2, help="frames of the CPU sample")
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--no-overlap", action="store_true", help="launch front end and medoid of a step on one stream (A/B of the two-stream overlap)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frame-parallel", action="store_true", help="skip the one-process-per-core CPU baseline")
    ap.add_argument("--no-framespec-leg", action="store_true", help="skip the FrameSpec-level (host packing included) leg")
    ap.add_argument("--ring", type=int, default=3, help="steps the host may run ahead of the GPU in the device-timed loop")
    ap.add_argument("--reader-threads", type=int, default=8, help="reader threads of the on-disk script leg")
    ap.add_argument("--no-disk-leg", action="store_true", help="skip the on-disk drop-in script leg")
    ap.add_argument("--no-latency-leg", action="store_true", help="skip the one-frame-per-call latency leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        args.ref_frames = 1 if args.ref_frames == 2 else args.ref_frames
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
