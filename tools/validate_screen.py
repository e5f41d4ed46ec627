#!/usr/bin/env python
"""Screened medoid (screen + verify) against the all-exact kernel on bench-sized batches.

    python tools/validate_screen.py [n_frames=64] [first_index=0] [config=c2|c3|c4]

Prints the number of instances compared, how many went through the screen, the number of columns the
verify pass summed exactly, and any disagreement (there must be none)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import bench
    from cm3d_b200.lifter import Lifter
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    cfg = sys.argv[3] if len(sys.argv) > 3 else "c2"
    frames = bench.make_frames({"c2": "c5"}.get(cfg, cfg), first, n, 16)
    lifter = Lifter("cuda:0")
    pb = lifter.pack(frames)
    db = lifter.upload(pb)
    out = {}
    for thr in (0, 512, 32):
        lifter.screen_min_pts = thr
        do = lifter.run(db)
        lab = lifter.fetch_labels(do)
        assert lifter.check_flags(lab) == 0
        out[thr] = (lab["medoid_local"].copy(), lab["medoid_point_idx"].copy(),
                    int(lifter.last_screen_stats.item()) if lifter.last_screen_stats is not None else 0,
                    np.diff(lab["seg_off"].astype(np.int64)))
    m = out[0][3]
    for thr in (512, 32):
        bad = np.flatnonzero(out[thr][0] != out[0][0])
        modes = lifter.last_screen_modes.cpu().numpy()[:m.size]
        print(f"modes (0 exact, 1 all pairs, 2 symmetric, 3 grouped, 4 pruned): {np.bincount(modes, minlength=5).tolist()}", flush=True)
        print(f"{cfg} frames {first}..{first + n - 1} screen_min_pts={thr}: {m.size} instances, {int((m >= max(thr, 32)).sum())} screened, "
              f"{out[thr][2]} columns verified, {bad.size} disagreements {bad[:10].tolist()}", flush=True)
        assert bad.size == 0 and np.array_equal(out[thr][1], out[0][1])


if __name__ == "__main__":
    main()
