"""COCO run-length masks: the `counts` string codec (host side).

The reference reads `{f}_masks.pkl` with `pycocotools.mask.decode`
(src/nuscenes/2d_to_3d.py:422-425, src/kitti/2d_to_3d.py:1001-1004,
src/waymo/2d_to_3d.py:520).  pycocotools (pinned 2.0.7, environment.yml:127) is
an un-vendored dependency; this restates its published string format
(maskApi.c rleToString/rleFrString): each run length is written LEB128-like
with 5 payload bits per char (chars 48..111, bit 0x20 = continuation, sign
extended from bit 0x10), and from the 4th run on as a delta against the run two
places earlier.  Only the tiny string -> run-length step runs on the host; the
runs are expanded, eroded and bit-packed on the GPU (csrc/masks.cu).
"""
from __future__ import annotations

import numpy as np


def rle_string_to_runs(s) -> np.ndarray:
    """Compressed COCO `counts` (bytes or str) -> uint32 run lengths (0-run first)."""
    if isinstance(s, str):
        s = s.encode("ascii")
    b = np.frombuffer(s, dtype=np.uint8).astype(np.int64) - 48
    runs = []
    p, n = 0, b.shape[0]
    while p < n:
        x, k, more = 0, 0, True
        while more:
            c = int(b[p])
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(runs) > 2:
            x += runs[-2]
        runs.append(x)
    return np.asarray(runs, dtype=np.uint32)


def runs_to_rle_string(runs) -> bytes:
    """uint32 run lengths -> compressed COCO `counts` bytes (inverse of the above)."""
    runs = [int(r) for r in np.asarray(runs).reshape(-1)]
    out = bytearray()
    for i, r in enumerate(runs):
        x = r - runs[i - 2] if i > 2 else r
        more = True
        while more:
            c = x & 0x1F
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(c + 48)
    return bytes(out)


def rle_counts_to_runs(counts) -> np.ndarray:
    """Accept either the compressed string or an already-decoded run array."""
    if isinstance(counts, (bytes, str)):
        return rle_string_to_runs(counts)
    return np.ascontiguousarray(np.asarray(counts, dtype=np.uint32).reshape(-1))
