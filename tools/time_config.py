#!/usr/bin/env python
"""Device-resident frames/s of the lifting path on a batch of KITTI- (c3) or Waymo-shaped (c4) frames,
with the per-kernel CUDA-event times (the bench line is quoted on C2 only).

    python tools/time_config.py c3 32        # config, frames in the batch
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    from cm3d_b200 import synthetic as S
    from cm3d_b200.lifter import Lifter
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    frames = [S.make_frame(cfg, 500 + i) for i in range(n)]
    lifter = Lifter("cuda:0")
    db = lifter.upload(lifter.pack(frames))
    lab = lifter.fetch_labels(lifter.run(db))
    cap = int(lab["seg_off"][-1]) + 4096
    for _ in range(3):
        lifter.run(db, seg_cap=cap)
    torch.cuda.synchronize()
    lifter.timing = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        lifter.run(db, seg_cap=cap)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = {k: round(float(np.mean([a.elapsed_time(b) for a, b in v])), 3) for k, v in lifter.timing.items()}
    m = np.diff(lab["seg_off"].astype(np.int64))
    modes = lifter.last_screen_modes.cpu().numpy()[:m.size]
    print(f"{cfg}: {n} frames, {ms:.2f} ms per batch = {n / ms * 1e3:.0f} frames/s; points {int(lab['frame_n'].sum())}, "
          f"members {int(m.sum())}, sum M^2 {float((m.astype(float) ** 2).sum()):.3g}; modes exact/all/sym/grouped "
          f"{[int((modes == k).sum()) for k in range(4)]}; kernels ms {t}", flush=True)


if __name__ == "__main__":
    main()
