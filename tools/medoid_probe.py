#!/usr/bin/env python
"""Times cm3d_medoid (through the C ABI) on controlled segment-size distributions, to separate the
main-loop rate from tail-column / partial-block / load-balance losses.

    python tools/medoid_probe.py            # prints cycles per warp-row-col and G pairs/s per case
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cm3d_b200 import _native as N  # noqa: E402


def run_case(name, sizes, reps=3, centre=(1200.0, 950.0, 1.0), order_desc=False, tiny_in=0):
    rng = np.random.default_rng(0)
    sizes = np.asarray(sizes, np.int64)
    seg_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    cap = (int(seg_off[-1]) + 3) & ~3
    xyzw = np.zeros((4, cap), np.float32)
    xyzw[:3, :seg_off[-1]] = (rng.normal(0, 3, (3, int(seg_off[-1]))) + np.array(centre)[:, None]).astype(np.float32)
    for k in range(tiny_in):                 # a coordinate below 2^-20 in the first instances: range-check fallback
        xyzw[2, seg_off[k] + 3] = 1e-9
    items = np.array([N.load().cm3d_medoid_items(int(m), 1) for m in sizes])
    if order_desc:
        order = np.argsort(-items, kind="stable").astype(np.int32)
    else:
        order = np.arange(len(sizes), dtype=np.int32)
    items = items[order]
    item_off = np.concatenate([[0], np.cumsum(items)]).astype(np.int32)
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(a).to(dev)
    d_xyzw, d_off, d_item = t(xyzw.reshape(-1)), t(seg_off), t(item_off)
    d_idx = torch.arange(cap, dtype=torch.int32, device=dev)
    d_inst = t(order)
    n = len(sizes)
    best = torch.full((n,), -1, dtype=torch.int64, device=dev)
    ml = torch.zeros(n, dtype=torch.int32, device=dev)
    mp = torch.zeros(n, dtype=torch.int32, device=dev)
    cen = torch.zeros(4 * n, dtype=torch.float32, device=dev)
    err = torch.zeros(4, dtype=torch.int32, device=dev)
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ssum = torch.zeros(cap, dtype=torch.float32, device=dev)
    smin = torch.zeros(5 * n, dtype=torch.int32, device=dev)
    ws = torch.zeros(5 * cap, dtype=torch.float32, device=dev)
    flags = int(os.environ.get("CM3D_SCREEN_FLAGS", "0"))
    ipos = torch.zeros(4 * (int(item_off[-1]) + 1), dtype=torch.int32, device=dev)
    screen = int(os.environ.get("CM3D_SCREEN_MIN_PTS", "512"))

    def launch():
        best.fill_(-1)
        N.call("cm3d_medoid", p(d_xyzw), cap, p(d_off), p(d_idx), p(d_item), p(d_inst), n, int(item_off[-1]), p(best),
               None, p(ssum), p(smin), screen, flags, p(ws), None, p(ipos), p(ml), p(mp), p(cen), p(err), st)
    launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pairs = float((sizes.astype(np.float64) ** 2).sum())
    rate = pairs / (ms * 1e-3)
    cyc = 148 * 4 * 32 * 1.965e9 / rate
    print(f"{name:<44s} {ms:8.3f} ms  {rate / 1e9:8.1f} Gpairs/s  {cyc:6.2f} cycles/warp-row-col", flush=True)
    return ml.cpu().numpy()


def split_cases():
    """Which part of the screened medoid pays for what (run under `ncu --metrics gpu__time_duration.sum`)."""
    big = [2976] * 3000
    run_case("A 3000 x M=2976", big, reps=1, order_desc=True)
    run_case("B A + 300 x M=300", big + [300] * 300, reps=1, order_desc=True)
    run_case("C A + 300 x M=20", big + [20] * 300, reps=1, order_desc=True)
    run_case("D A, 8 instances fail the range check", big, reps=1, order_desc=True, tiny_in=8)
    run_case("E 300 x M=8000, 4 fail the range check", [8000] * 300, reps=1, order_desc=True, tiny_in=4)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "split":
        return split_cases()
    rng = np.random.default_rng(1)
    run_case("592 x M=4096 (no tails, full blocks)", [4096] * 592)
    run_case("1184 x M=2048", [2048] * 1184)
    run_case("592 x M=4111 (tail 15 cols, partial block)", [4111] * 592)
    run_case("592 x M=4352 (=17*256, no tail)", [4352] * 592)
    run_case("148 x M=8192", [8192] * 148)
    m = np.clip(rng.lognormal(np.log(2500), 0.9, 3200), 5, 9000).astype(np.int64)
    run_case("3200 x lognormal(2500, 0.9) (bench-like)", m)
    run_case("same, scheduled largest-first", m, order_desc=True)
    run_case("kitti-like coords 3200 x lognormal", m, centre=(10.0, 1.0, 20.0))


if __name__ == "__main__":
    main()
