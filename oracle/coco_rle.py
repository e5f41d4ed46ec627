"""ORACLE (test infrastructure, not product): the COCO run-length mask codec, restated from the
published C source of pycocotools 2.0.7 (common/maskApi.c: rleEncode, rleDecode, rleToString,
rleFrString) - the un-vendored dependency the reference calls at src/nuscenes/2d_to_3d.py:425,
src/kitti/2d_to_3d.py:1004, src/waymo/2d_to_3d.py:520 (`pycocotools.mask.decode`) and its mask
generator at src/nuscenes/gen_2d_masks_detic.py:471 (`encode`).

This file shares no code with the product's cm3d_b200/rle.py or csrc/masks.cu; the tests compare
the two.  PARITY UNPINNED against the pycocotools binary itself (not installable here): the pins
are (1) hand-worked known-answer strings in tests/test_rle_vectors.py, derived character by
character from the format below, and (2) the run lists an independent third-party implementation
produces for the same masks (transformers' SAM `_mask_to_rle`, "the format expected by pycoco
tools"), committed as tests/golden/rle_vectors.npz by oracle/make_golden.py.

Format (maskApi.c).  A mask of h rows and w columns is flattened COLUMN-major (Fortran order) and
written as run lengths cnts[0..m), starting with a run of zeros (possibly of length 0).  The
string form writes each count as a little-endian base-32 varint, 5 payload bits per character:
    x = cnts[i] - (i > 2 ? cnts[i-2] : 0)
    do { c = x & 0x1f; x >>= 5;  more = (c & 0x10) ? x != -1 : x != 0;
         if (more) c |= 0x20;  emit char(c + 48); } while (more);
so characters lie in 48..111, bit 0x20 is "continue", and the last character's bit 0x10 is the sign.
"""
from __future__ import annotations

import numpy as np


def fr_string(s) -> list:
    """rleFrString: compressed `counts` (bytes or str) -> python list of run lengths."""
    if isinstance(s, str):
        s = s.encode("ascii")
    cnts = []
    p = 0
    n = len(s)
    while p < n:
        x = 0
        k = 0
        more = 1
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = c & 0x20
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(cnts) > 2:
            x += cnts[len(cnts) - 2]
        cnts.append(x)
    return cnts


def to_string(cnts) -> bytes:
    """rleToString: run lengths -> compressed `counts` bytes."""
    out = bytearray()
    cnts = [int(v) for v in cnts]
    for i in range(len(cnts)):
        x = cnts[i]
        if i > 2:
            x -= cnts[i - 2]
        more = 1
        while more:
            c = x & 0x1F
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(c + 48)
    return bytes(out)


def encode_runs(mask_hw: np.ndarray) -> list:
    """rleEncode of one (h, w) mask: run lengths over the column-major flattening, zeros first."""
    flat = np.asarray(mask_hw).astype(bool).reshape(-1, order="F")
    cnts = []
    prev = False
    c = 0
    for v in flat.tolist():
        if v != prev:
            cnts.append(c)
            c = 0
            prev = v
        c += 1
    cnts.append(c)
    return cnts


def decode_runs(cnts, h: int, w: int) -> np.ndarray:
    """rleDecode: run lengths -> (h, w) uint8 mask (column-major fill, value toggles per run)."""
    flat = np.zeros(h * w, np.uint8)
    pos = 0
    v = 0
    for c in cnts:
        c = int(c)
        if c < 0 or pos + c > h * w:
            raise ValueError("run lengths do not fit the mask")
        if v:
            flat[pos:pos + c] = 1
        pos += c
        v ^= 1
    if pos != h * w:
        raise ValueError("run lengths do not cover the mask")
    return flat.reshape((h, w), order="F")


def counts_to_runs(counts) -> list:
    """`counts` as found in a COCO RLE dict: compressed bytes/str, or an uncompressed list."""
    if isinstance(counts, (bytes, str)):
        return fr_string(counts)
    return [int(v) for v in np.asarray(counts).reshape(-1)]


def decode(rle_objs):
    """pycocotools.mask.decode: one RLE dict -> (h, w) uint8; a list of n dicts -> (h, w, n)."""
    if isinstance(rle_objs, dict):
        h, w = rle_objs["size"]
        return decode_runs(counts_to_runs(rle_objs["counts"]), int(h), int(w))
    return np.stack([decode(r) for r in rle_objs], axis=2) if len(rle_objs) else np.zeros((0, 0, 0), np.uint8)


def encode(mask: np.ndarray):
    """pycocotools.mask.encode: (h, w) or (h, w, n) uint8 -> RLE dict(s) with compressed counts."""
    m = np.asarray(mask)
    if m.ndim == 2:
        return {"size": [int(m.shape[0]), int(m.shape[1])], "counts": to_string(encode_runs(m))}
    return [encode(m[:, :, k]) for k in range(m.shape[2])]
