#!/usr/bin/env python
"""Where the FrameSpec-level path spends its host time: pack (C packer vs Python packer, pinned buffers),
upload enqueue, run enqueue, results construction.  16 C2 frames per batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    from cm3d_b200.batch import pack_frames, pack_frames_native
    from cm3d_b200.lifter import Lifter
    from concurrent.futures import ThreadPoolExecutor
    frames = bench.make_frames(0, 16, 16)
    lifter = Lifter("cuda:0")
    for name, fn in (("python packer", pack_frames), ("C packer", pack_frames_native)):
        fn(frames, pin=True)
        t = time.perf_counter()
        for _ in range(8):
            pb = fn(frames, pin=True)
        print(f"{name}: {(time.perf_counter() - t) / 8 / 16 * 1e3:.2f} ms per frame (pinned)", flush=True)
    def pooled(_):
        p = lifter._pack_pooled(frames)
        p.release()
    for workers in (1, 2, 4, 8):
        with ThreadPoolExecutor(workers) as pool:
            list(pool.map(pooled, range(2 * workers)))
            t = time.perf_counter()
            list(pool.map(pooled, range(32)))
            print(f"C packer, pooled pinned buffers, {workers} threads: {(time.perf_counter() - t) / 32 / 16 * 1e3:.2f} ms per frame effective", flush=True)
    db = lifter.upload(pb)
    do = lifter.run(db)
    lab = lifter.fetch_labels(do)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(8):
        res = lifter.results(do, lab, with_points=False)
    print(f"results(): {(time.perf_counter() - t) / 8 / 16 * 1e3:.2f} ms per frame", flush=True)
    t = time.perf_counter()
    for _ in range(8):
        do = lifter.run(db)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"run() enqueue: {(t1 - t) / 8 / 16 * 1e3:.3f} ms per frame; kernels {(time.perf_counter() - t) / 8 / 16 * 1e3:.3f} ms per frame", flush=True)


if __name__ == "__main__":
    main()
