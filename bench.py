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
WORKLOAD = "C2: nuScenes-shaped 10 sweeps x 34,720 pts (~347k pts) x 6 cams 1024x576 masks x 50 instances per frame"


def _gen_frame(index):
    from cm3d_b200 import synthetic as S
    f = S.make_frame("c5", index, dense_masks=False)         # C5 = independent C2-shaped frames
    f.masks = S.compress_rles(f.masks)                       # masks as in {f}_masks.pkl: COCO counts strings
    return f


def make_frames(first, count, workers):
    idx = list(range(first, first + count))
    if workers <= 1 or count <= 2:
        return [_gen_frame(i) for i in idx]
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(workers) as pool:
        return pool.map(_gen_frame, idx)


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
    """The reference's per-frame body (oracle/ref_lift.py, torch CPU) over `frames`; seconds."""
    import torch
    from oracle import ref_lift as RL
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    for f in frames:
        RL.lift_frame(f, record_pix=False)
    return time.perf_counter() - t0


def _ref_worker(index):
    """One frame through the CPU port on ONE thread (frame-parallel CPU baseline); returns seconds."""
    import torch
    torch.set_num_threads(1)
    from oracle import ref_lift as RL
    f = _gen_frame(index)
    RL.lift_frame(_gen_frame_small(), record_pix=False)          # page in torch, untimed
    t0 = time.perf_counter()
    RL.lift_frame(f, record_pix=False)
    return time.perf_counter() - t0


def _gen_frame_small():
    from cm3d_b200 import synthetic as S
    return S.make_frame("c1", 0, scale=0.1)


def cpu_frame_parallel(n_procs):
    """Frames/s of the CPU port run one process per core, one frame each, all at the same time."""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(n_procs) as pool:
        secs = pool.map(_ref_worker, list(range(1000, 1000 + n_procs)), chunksize=1)
    return sum(1.0 / s for s in secs), secs


def run_reference(args, rank):
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    nf = max(1, args.ref_frames)
    frames = make_frames(0, nf, 1)
    for _ in range(args.warmup):
        cpu_reference_run(frames[:1], threads)
    times = [cpu_reference_run(frames, threads) for _ in range(args.steps)]
    total = sum(times)
    v = nf * args.steps / total
    sample = f"{nf} C2 frame(s) per step x {args.steps} steps, oracle/ref_lift.py (torch {torch.__version__} CPU)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": nf, "mask_input": "COCO counts strings (pycocotools format), decoded on the GPU"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local_rank):
    # frames first (spawned workers), CUDA afterwards
    workers = max(1, min(args.workers or (os.cpu_count() or 1) // max(world, 1), 16))
    t_gen = time.perf_counter()
    frames = make_frames(rank * args.batch, args.batch, workers)
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
    pb = lifter.pack(frames)
    db = lifter.upload(pb)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- find the segment capacity once (exact), then warm up
    do = lifter.run(db)
    labels = lifter.fetch_labels(do)
    need = lifter.check_flags(labels)
    seg_total = int(labels["seg_off"][-1])
    seg_cap = max(need, seg_total) + 4096
    n_points = int(labels["frame_n"].sum())
    items = int(labels["item_off"][-1])
    so = labels["seg_off"].astype(np.int64)
    m = np.diff(so)
    pairs = float((m.astype(np.float64) ** 2).sum())
    del do
    for _ in range(max(args.warmup, 3)):
        lifter.run(db, seg_cap=seg_cap)
    barrier()

    # ---- device-resident timed region (CUDA events on the launch stream, per-kernel events inside)
    sampler = ClockSampler(local_rank)
    sampler.start()
    lifter.timing = {}
    lifter.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    do = None
    for _ in range(args.steps):
        del do                          # free the previous step's buffers first: no second workspace, no cudaMalloc
        do = lifter.run(db, seg_cap=seg_cap)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = lifter.launches
    timing = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in lifter.timing.items()}
    lifter.timing = None
    final = lifter.fetch_labels(do)
    assert lifter.check_flags(final) == 0
    screen_modes = lifter.last_screen_modes.cpu().numpy() if lifter.last_screen_modes is not None else None
    screen_verified = int(lifter.last_screen_stats.item()) if lifter.last_screen_stats is not None else None

    # ---- host->device copy rate of one packed batch (explains e2e when PCIe, not the kernels, bounds it)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lifter.upload(pb)
    torch.cuda.synchronize()
    h0.record()
    for _ in range(3):
        lifter.upload(pb)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbps = 3 * pb.h2d_bytes / (h0.elapsed_time(h1) * 1e-3) / 1e9

    # ---- end to end: pinned host buffers -> H2D -> kernels -> D2H of the labels, every step.  A step's
    # frames go through the public streaming API in sub-batches (same frames, same order): the copy of
    # sub-batch k+1 overlaps the kernels of sub-batch k, and only the very first copy is exposed
    sub = max(1, min(args.e2e_sub or args.batch, args.batch))
    subs = [lifter.pack(frames[i:i + sub]) for i in range(0, len(frames), sub)] if sub < args.batch else [pb]
    sub_cap = seg_cap
    if len(subs) > 1:                   # segment capacity of the largest sub-batch (exact, found once, untimed)
        sub_cap = 4096 + max(int(lifter.fetch_labels(lifter.run(lifter.upload(p)))["seg_off"][-1]) for p in subs)
    for lab in lifter.lift_packed_stream(subs * 3, seg_cap=sub_cap):
        pass
    barrier()
    t0 = time.perf_counter()
    labs = []
    for lab in lifter.lift_packed_stream(subs * args.steps, seg_cap=sub_cap):
        labs.append(lab["medoid_point_idx"])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    assert np.array_equal(np.concatenate(labs[-len(subs):]), final["medoid_point_idx"])
    d2h_bytes = sum(int(lifter._out_layout(p.n_frames, p.n_inst)["_words"]) * 4 for p in subs)
    h2d_bytes = sum(p.h2d_bytes for p in subs)

    # ---- the same frames as FrameSpecs (what the drop-in scripts hand over): host packing included
    fs_fps = None
    if world == 1 and not args.no_framespec_leg:
        fs_kw = dict(batch_frames=16, pack_workers=4)       # small batches: the packers start the pipeline sooner
        for _ in lifter.lift_frame_stream(iter(frames * 2), **fs_kw):     # also fills the pinned-buffer pool
            pass
        t0 = time.perf_counter()
        n_fs = 0
        for res in lifter.lift_frame_stream(iter(frames * 4), **fs_kw):
            n_fs += len(res)
        torch.cuda.synchronize()
        fs_fps = n_fs / (time.perf_counter() - t0)

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        total_frames = args.batch * world * args.steps
        value = total_frames / (dev_ms * 1e-3)
        e2e = total_frames / (e2e_ms * 1e-3)
        peak, peak_src = measured_peak()
        raw_bytes = int(pb.raw.nbytes)
        algo = {   # algorithmic bytes per launch (DESIGN.md "Kernels and rooflines")
            "aggregate": raw_bytes + 16 * n_points,
            "project_count": 16 * n_points,
            "compact": 4 * n_points + 36 * seg_total,
            "masks_decode": int(pb.mask.nbytes) * 5,
            "masks_rle": int(pb.mask.nbytes) * 4 + 4 * pb.bits_words,
            "masks_erode": 8 * pb.bits_words,
        }
        kern = {}
        for k, ms in timing.items():
            kern[k] = {"ms": ms, "share": ms * args.steps / dev_ms}
            if k in algo:
                kern[k]["algo_bytes"] = algo[k]
                kern[k]["gbps"] = algo[k] / (ms * 1e-3) / 1e9
        hbm_k = max((k for k in kern if k in ("aggregate", "project_count", "compact")), key=lambda k: kern[k]["ms"])
        traffic, traffic_src = None, None
        try:        # DRAM bytes per launch of that kernel from the newest committed ncu capture of this same command
            cands = sorted(fn for fn in os.listdir(os.path.join(ROOT, "profiles")) if fn.endswith("_traffic.json"))
            with open(os.path.join(ROOT, "profiles", cands[-1])) as f:
                tj = json.load(f)
            if tj.get("frames_per_launch") == args.batch:
                traffic, traffic_src = tj["dram_bytes_per_launch"].get("k_" + hbm_k), tj["source"]
        except Exception:
            pass
        roof = {"kernel": hbm_k, "bound": "hbm", "achieved": kern[hbm_k]["gbps"], "peak": peak, "unit": "GB/s",
                "frac": kern[hbm_k]["gbps"] / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algo_bytes": algo[hbm_k], "peak_source": peak_src}
        if "medoid" in timing:
            kern["medoid"]["pair_distances"] = pairs
            kern["medoid"]["gpairs_per_s"] = pairs / (timing["medoid"] * 1e-3) / 1e9
            # XU-pipe ceiling: one MUFU square root per EVALUATED pair and the XU pipe retires
            # 16 lanes per SM and clock.  Instances whose squared-distance matrix is exactly symmetric
            # (mode 2) are screened over the pairs i <= j of 256-column strips: about half the roots.
            n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
            mhz = clocks.get("sm_mhz") or 1965.0
            ceil_roots = n_sm * 16 * mhz * 1e6 / 1e9
            roots = pairs
            if screen_modes is not None:
                I_ = m.size
                modes, g0, g1 = screen_modes[:I_], screen_modes[I_:2 * I_].astype(np.float64), screen_modes[2 * I_:3 * I_].astype(np.float64)
                mm = m.astype(np.float64)

                def tri(x):      # roots of a symmetric strip sweep over x points: pairs i <= j by 256-column strips
                    fb = np.floor(x / 256.0)
                    return 256.0 * 256.0 * fb * (fb + 1) / 2.0 + (x - 256.0 * fb) * x
                g2 = np.maximum(mm - g0 - g1, 0.0)
                grouped = tri(g0) + tri(g2) + (mm * mm - g0 * g0 - g2 * g2)      # both groups symmetric, the rest in both orders
                roots = float(np.where(modes == 2, tri(mm), np.where(modes == 3, grouped, mm * mm)).sum())
                kern["medoid"]["instances_by_mode"] = {"exact": int((modes == 0).sum()), "screen_all_pairs": int((modes == 1).sum()),
                                                       "screen_symmetric": int((modes == 2).sum()),
                                                       "screen_grouped_symmetric": int((modes == 3).sum())}
            kern["medoid"]["bound"] = ("XU pipe: one MUFU.SQRT per evaluated pair distance in the screen pass (reads only "
                                       "sum M points, L2-resident); symmetric instances evaluate the pairs i <= j only")
            kern["medoid"]["square_roots"] = roots
            kern["medoid"]["groots_per_s"] = roots / (timing["medoid"] * 1e-3) / 1e9
            kern["medoid"]["peak_groots_per_s"] = ceil_roots
            kern["medoid"]["frac"] = kern["medoid"]["groots_per_s"] / ceil_roots
            if screen_verified is not None:
                kern["medoid"]["screen_min_pts"] = lifter.screen_min_pts
                kern["medoid"]["verified_columns_per_step"] = screen_verified
                kern["medoid"]["instances_per_step"] = int(m.size)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": args.batch, "mask_input": "COCO counts strings (pycocotools format), decoded on the GPU",
                       "points_per_step": n_points, "member_points_per_step": seg_total,
                       "l2": f"inputs per step {pb.h2d_bytes / 1e6:.0f} MB + {4 * 5 * pb.n_tiles * 1024 / 1e6:.0f} MB "
                             f"intermediates > 126 MB L2, no explicit flush",
                       "parallelism": f"frames sharded by sample index, {world} process(es), no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": d2h_bytes * world, "ms_per_step": e2e_ms / args.steps,
                    "h2d_gbps_rank0": round(h2d_gbps, 1), "sub_batch_frames": sub,
                    "api": "Lifter.lift_packed_stream (pinned host buffers in, label block out)"},
            "gpu_launches": launches,
            "e2e_from_framespecs": None if fs_fps is None else {
                "value": fs_fps, "unit": UNIT,
                "note": "Lifter.lift_frame_stream over 4 x the step's FrameSpecs (numpy sweeps + RLE masks + calibration): "
                        "C packer (csrc/pack.cu, ~0.9 ms per frame and thread, GIL released) on 4 worker threads into "
                        "pooled pinned buffers, 16-frame batches"},
            "roofline": roof,
            "path_hbm": {"algorithmic_bytes_per_step": int(sum(algo.values())) * world,
                         "achieved_gbps": sum(algo.values()) * world / (dev_ms / args.steps * 1e-3) / 1e9,
                         "frac_of_peak_per_gpu": sum(algo.values()) / (dev_ms / args.steps * 1e-3) / 1e9 / peak,
                         "note": "sum of the kernels' algorithmic bytes over the whole step; the step is bound by the "
                                 "medoid's square roots (XU pipe), not by HBM (kernels.medoid)"},
            "kernels": kern,
            "gen_seconds": round(t_gen, 1),
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            nf = max(1, args.ref_frames)
            from cm3d_b200 import synthetic as S
            cpu_reference_run([S.make_frame("c1", 0, scale=0.25)], threads)      # spin up the torch thread pool
            secs = cpu_reference_run(frames[:nf], threads)
            line["cpu_baseline"] = {"value": nf / secs, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{nf} of this step's C2 frames through oracle/ref_lift.py "
                                              f"(torch {torch.__version__} CPU, {threads} threads), after a small warm-up frame"}
            if not args.no_frame_parallel:
                try:      # one single-threaded process per core, one C2 frame each, concurrently
                    fp, per = cpu_frame_parallel(threads)
                    line["cpu_baseline"]["frame_parallel"] = {
                        "value": fp, "unit": UNIT, "processes": threads,
                        "sample": f"{threads} C2 frames, one per process, 1 torch thread each, run concurrently; "
                                  f"sum of 1/seconds (mean {sum(per) / len(per):.1f} s per frame)"}
                except Exception as e:
                    line["cpu_baseline"]["frame_parallel"] = {"error": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--e2e-sub", type=int, default=0,
                    help="frames per pipelined sub-batch of the end-to-end leg (0 = the whole batch; 16 measured 1.3 %% "
                         "slower on one GPU: the host's per-batch enqueue work stops hiding behind 5.6 ms of kernels)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-frames", type=int, default=2, help="frames of the CPU sample")
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frame-parallel", action="store_true", help="skip the one-process-per-core CPU baseline")
    ap.add_argument("--no-framespec-leg", action="store_true", help="skip the FrameSpec-level (host packing included) leg")
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
