"""ORACLE tooling: independent run-length vectors for the COCO mask format.

pycocotools is absent, so the run lists come from a THIRD implementation that shares no code with
this repo: transformers' SAM post-processing `_mask_to_rle` ("Encodes masks the run-length encoding
(RLE), in the format expected by pycoco tools": column-major runs, zeros first), and its inverse
`_rle_to_mask`.  Committed as tests/golden/rle_vectors.npz: masks (bit-packed), their run lists by
transformers, and the compressed `counts` strings oracle/coco_rle.to_string makes of those runs.
The string codec itself is pinned by hand-worked known-answer tests (tests/test_rle_vectors.py).

    python -m oracle.refrun.rle_vectors
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
OUT = os.path.join(ROOT, "tests", "golden", "rle_vectors.npz")


def masks(rng):
    out = []
    for (h, w) in [(1, 1), (3, 5), (7, 4), (33, 65), (64, 31), (100, 257), (576, 1024)]:
        out.append(np.zeros((h, w), bool))
        out.append(np.ones((h, w), bool))
        m = rng.uniform(size=(h, w)) < 0.5
        out.append(m)
        m = np.zeros((h, w), bool)
        m[h // 4: h // 4 + max(h // 2, 1), w // 3: w // 3 + max(w // 2, 1)] = True     # a blob: long runs (multi-char varints)
        m &= rng.uniform(size=(h, w)) < 0.97
        out.append(m)
        m = np.zeros((h, w), bool)
        m[0, 0] = True                                                                  # first pixel set: leading 0-run
        m[-1, -1] = True
        out.append(m)
    return out


def main():
    import torch
    from transformers.models.sam.image_processing_sam import _mask_to_rle, _rle_to_mask
    from oracle import coco_rle
    rng = np.random.default_rng(425)
    d = {}
    ms = masks(rng)
    for k, m in enumerate(ms):
        rle = _mask_to_rle(torch.from_numpy(m)[None])[0]
        h, w = rle["size"]
        runs = np.asarray(rle["counts"], np.int64)
        assert (h, w) == m.shape and runs.sum() == h * w
        assert np.array_equal(_rle_to_mask(rle).numpy(), m)
        d[f"hw_{k}"] = np.array([h, w], np.int32)
        d[f"bits_{k}"] = np.packbits(m.reshape(-1))
        d[f"runs_{k}"] = runs
        d[f"counts_{k}"] = np.frombuffer(coco_rle.to_string(runs), np.uint8)
    d["n"] = np.array(len(ms))
    import transformers
    d["made_with"] = np.array(f"transformers {transformers.__version__} models/sam/image_processing_sam.py _mask_to_rle")
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB,", len(ms), "masks")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
