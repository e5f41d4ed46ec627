"""COCO run-length masks: the string codec and the run semantics, three ways.

  * hand-worked known answers of the published format (pycocotools maskApi.c rleToString / rleFrString),
    derived character by character in the comments below - independent of every implementation here;
  * tests/golden/rle_vectors.npz: run lists an independent third-party implementation (transformers'
    SAM `_mask_to_rle`) produced for 35 masks (oracle/refrun/rle_vectors.py);
  * product (cm3d_b200/rle.py host codec, csrc/masks.cu device decoder) vs oracle (oracle/coco_rle.py).
"""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# count -> characters, by hand from `c = x & 0x1f; x >>= 5; more = (c & 0x10) ? x != -1 : x != 0; if more c |= 0x20; chr(c+48)`
#   0   -> c=0,  x=0, bit4 clear, more = (0 != 0) = 0          -> chr(48)      "0"
#   5   -> c=5                                                 -> chr(53)      "5"
#   15  -> c=15, bit4 clear, x=0 -> stop                        -> chr(63)      "?"
#   16  -> c=16, bit4 SET, x=0, more = (0 != -1) = 1 -> c|=32=48 -> chr(96) "`"; then c=0, x=0, stop -> "0"      "`0"
#   31  -> c=31, bit4 set, x=0 -> more -> 63 -> chr(111) "o"; then c=0 -> "0"                                     "o0"
#   32  -> c=0, x=1, bit4 clear, more = (1 != 0) -> c=32 -> chr(80) "P"; then c=1, x=0 -> chr(49) "1"            "P1"
#   1000-> 1000 = 31*32 + 8: c=8, x=31, more -> 40 -> chr(88) "X"; c=31, x=0, bit4 set, more (0 != -1) -> 63 -> "o"; c=0 -> "0"   "Xo0"
#   -1  -> c=31, x=-1, bit4 set, more = (-1 != -1) = 0          -> chr(79)      "O"
#   -3  -> c=29 (0b11101), x=-1, bit4 set -> stop               -> chr(77)      "M"
#   -17 -> -17 & 31 = 15, x = -1, bit4 CLEAR, more = (-1 != 0) -> 47 -> chr(95) "_"; then c=31, x=-1, stop -> "O"  "_O"
SINGLE = {0: b"0", 5: b"5", 15: b"?", 16: b"`0", 31: b"o0", 32: b"P1", 1000: b"Xo0"}
# runs -> string: from the 4th run on (index > 2) the DIFFERENCE to the run two places back is written
#   [5, 2, 2, 0, 3]     -> 5, 2, 2, 0-2=-2 ("N": -2&31=30 -> chr(78)), 3-2=1      "522N1"
#   [0, 6]              -> "06"   (mask starts with a set pixel: leading zero-run of length 0)
#   [7, 1, 7, 18, 6]    -> 7, 1, 7, 18-1=17 (c=17, bit4 set, x=0 -> more -> 49 -> chr(97) "a", then "0"), 6-7=-1 "O"   "717a0O"
#   [100, 3, 97]        -> 100 = 3*32+4: c=4, x=3 -> more -> 36 -> chr(84) "T"; c=3 -> "3";  3 -> "3";  97 = 3*32+1: c=1|32=33 -> chr(81) "Q"; "3"   "T33Q3"
MULTI = {(5, 2, 2, 0, 3): b"522N1", (0, 6): b"06", (7, 1, 7, 18, 6): b"717a0O", (100, 3, 97): b"T33Q3"}
NEGATIVE_DELTAS = {(40, 1, 39, 0, 22): None}     # round-trip only


def test_known_answer_strings_oracle_and_product():
    from cm3d_b200 import rle as PR
    from oracle import coco_rle as OR
    for v, s in SINGLE.items():
        assert OR.to_string([v]) == s and OR.fr_string(s) == [v]
        assert PR.runs_to_rle_string([v]) == s and PR.rle_string_to_runs(s).tolist() == [v]
    for runs, s in MULTI.items():
        assert OR.to_string(runs) == s and OR.fr_string(s) == list(runs)
        assert PR.runs_to_rle_string(runs) == s and PR.rle_string_to_runs(s).tolist() == list(runs)
    for runs in NEGATIVE_DELTAS:
        assert OR.fr_string(OR.to_string(runs)) == list(runs)
        assert PR.rle_string_to_runs(OR.to_string(runs)).tolist() == list(runs)


def _vectors():
    d = np.load(os.path.join(ROOT, "tests", "golden", "rle_vectors.npz"))
    for k in range(int(d["n"])):
        h, w = (int(v) for v in d[f"hw_{k}"])
        mask = np.unpackbits(d[f"bits_{k}"])[:h * w].reshape(h, w).astype(np.uint8)
        yield k, h, w, mask, d[f"runs_{k}"].astype(np.int64), d[f"counts_{k}"].tobytes()


def test_third_party_run_lists_oracle_and_product():
    """Column-major, zeros-first run semantics: transformers' run lists == the oracle's encoder, and both
    decoders give the mask back; the product's host codec reads the same strings."""
    from cm3d_b200 import rle as PR
    from oracle import coco_rle as OR
    n = 0
    for k, h, w, mask, runs, counts in _vectors():
        if h * w <= 70000:
            assert OR.encode_runs(mask) == runs.tolist(), k
        assert np.array_equal(OR.decode_runs(runs, h, w), mask), k
        assert OR.fr_string(counts) == runs.tolist(), k
        assert PR.rle_string_to_runs(counts).tolist() == runs.tolist(), k
        assert np.array_equal(OR.decode({"size": [h, w], "counts": counts}), mask), k
        n += 1
    assert n == 35


@pytest.mark.gpu
def test_device_decoder_on_third_party_vectors():
    """csrc/masks.cu: counts strings -> run lengths -> bit planes, against the fixture masks.  The
    reference reads masks of `size=[W,H]` whose column-major runs are the row-major (H,W) image
    (gen_2d_masks_detic.py:468-471), so a fixture mask of shape (h,w) is the image (H=w... transposed)."""
    import torch
    from cm3d_b200 import _native as N
    vecs = list(_vectors())
    I = len(vecs)
    blob = b"".join(v[5] for v in vecs)
    off = np.concatenate([[0], np.cumsum([len(v[5]) for v in vecs])]).astype(np.int64)
    descs, word_off = [], 0
    for k, h, w, mask, runs, counts in vecs:
        W, H = h, w                           # RLE size = [h, w] = [W, H] of the image the lifter sees
        pitch = (W + 31) // 32
        descs.append([word_off & 0xFFFFFFFF, word_off >> 32, W, H, pitch, 0, 0, k])
        word_off += pitch * H
    desc = np.array(descs, np.int64).astype(np.int32)
    dev = "cuda:0"
    d_blob = torch.from_numpy(np.frombuffer(blob, np.uint8).copy()).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    d_desc = torch.from_numpy(desc.reshape(-1)).to(dev)
    d_runs = torch.zeros(max(len(blob), 1), dtype=torch.int32, device=dev)
    run_start = torch.empty(max(len(blob), 1), dtype=torch.int32, device=dev)
    row_range = torch.empty(2 * I, dtype=torch.int32, device=dev)
    bits = torch.zeros(word_off + 4, dtype=torch.int32, device=dev)
    err = torch.zeros(4, dtype=torch.int32, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    max_runs = max(len(v[5]) for v in vecs)
    N.call("cm3d_masks_decode_counts", p(d_blob), p(d_off), I, p(d_runs), st)
    N.call("cm3d_masks_fill_rle", p(d_runs), p(d_off), p(run_start), p(d_desc), I, int(max_runs), p(bits), p(row_range), p(err), st)
    torch.cuda.synchronize()
    assert err.cpu().numpy()[1] == 0
    got_runs = d_runs.cpu().numpy().view(np.uint32)
    words = bits.cpu().numpy().view(np.uint32)
    for (k, h, w, mask, runs, counts), dsc in zip(vecs, descs):
        o0 = int(off[k])
        assert np.array_equal(got_runs[o0:o0 + len(runs)], runs.astype(np.uint32)), k
        W, H, pitch = dsc[2], dsc[3], dsc[4]
        plane = words[dsc[0]:dsc[0] + pitch * H].reshape(H, pitch)
        img = ((plane[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(H, pitch * 32)[:, :W].astype(np.uint8)
        assert np.array_equal(img, mask.T), k          # (H,W) image = the (W,H)=(h,w) array transposed
