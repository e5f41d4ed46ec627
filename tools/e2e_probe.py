#!/usr/bin/env python
"""Where does the end-to-end step time go?  Host timestamps around the phases of
Lifter.lift_packed_stream for a bench-sized batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cm3d_b200.lifter import Lifter

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    frames = bench.make_frames(0, B, 16)
    lifter = Lifter("cuda:0")
    pb = lifter.pack(frames)
    db = lifter.upload(pb); do = lifter.run(db); lab = lifter.fetch_labels(do)
    seg_cap = int(lab["seg_off"][-1]) + 4096
    for _ in range(3): lifter.run(db, seg_cap=seg_cap)
    torch.cuda.synchronize()
    # phase timing, serialised
    for rep in range(3):
        t0 = time.perf_counter(); d = lifter.upload(pb); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        o = lifter.run(d, seg_cap=seg_cap); t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
        l = lifter.fetch_labels(o); t5 = time.perf_counter()
        print(f"serial rep{rep}: upload enqueue {1e3*(t1-t0):.2f} ms, copy done +{1e3*(t2-t1):.2f}, run enqueue {1e3*(t3-t2):.2f}, kernels done +{1e3*(t4-t3):.2f}, fetch {1e3*(t5-t4):.2f}", flush=True)
    for depth in (2, 3):
        for rep in range(2):
            ts = [time.perf_counter()]
            for lab in lifter.lift_packed_stream([pb] * 8, seg_cap=seg_cap, depth=depth):
                ts.append(time.perf_counter())
            torch.cuda.synchronize()
            print(f"stream depth={depth} rep{rep}: per-yield ms", [round(1e3 * (b - a), 1) for a, b in zip(ts, ts[1:])], flush=True)
    print(torch.cuda.memory_summary(abbreviated=True)[:1500])

if __name__ == "__main__":
    main()
