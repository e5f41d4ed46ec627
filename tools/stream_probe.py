#!/usr/bin/env python
"""Where does the FrameSpec stream lose time?  Packer threads alone, packed stream alone, full stream."""
import os, sys, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    frames = bench.make_frames("c5", 0, n, 16)
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from cm3d_b200.lifter import Lifter
    lifter = Lifter("cuda:0")
    for bf in (32,):
        groups = [frames[i:i + bf] for i in range(0, n, bf)]
        for workers in (8,):
            with ThreadPoolExecutor(workers) as pool:
                list(pool.map(lifter._pack_pooled, groups[:workers]))
                for pb in list(pool.map(lifter._pack_pooled, groups)):
                    pb.release()
                t0 = time.perf_counter()
                pbs = list(pool.map(lifter._pack_pooled, groups))
                dt = time.perf_counter() - t0
                for pb in pbs:
                    pb.release()
            print(f"pack only  batch {bf} workers {workers}: {n / dt:8.0f} frames/s", flush=True)
    for bf in (32, 64):
        groups = [frames[i:i + bf] for i in range(0, n, bf)]
        pbs = [lifter.pack(g, keep_fourth=False) for g in groups]
        for _ in lifter.lift_packed_stream(pbs):
            pass
        t0 = time.perf_counter()
        for _ in lifter.lift_packed_stream(pbs * 2):
            pass
        torch.cuda.synchronize()
        print(f"packed stream batch {bf}: {2 * n / (time.perf_counter() - t0):8.0f} frames/s", flush=True)
        del pbs
    for bf in (32, 64):
        for workers in (4, 8, 12):
            for _ in lifter.lift_frame_stream(iter(frames), batch_frames=bf, pack_workers=workers):
                pass
            lifter.stream_stats.clear()
            t0 = time.perf_counter()
            for _ in lifter.lift_frame_stream(iter(frames * 2), batch_frames=bf, pack_workers=workers):
                pass
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"frame stream batch {bf} workers {workers}: {2 * n / dt:8.0f} frames/s  total {dt:.3f}s", {k: round(v, 3) for k, v in lifter.stream_stats.items()}, flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in lifter.lift_frame_stream(iter(frames * 2), batch_frames=64, pack_workers=8):
        pass
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
    print(s.getvalue())


if __name__ == "__main__":
    main()
