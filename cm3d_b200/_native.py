"""ctypes binding of the C-ABI library (include/cm3d_b200.h).

The library is built in-tree by `cm3d_b200.build.build()` (nvcc, sm_100a) and
loaded from `cm3d_b200/_lib/libcm3d_b200.so`.  There is no CPU fallback: if the
library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CM3D_B200_LIB") or os.path.join(_HERE, "_lib", "libcm3d_b200.so")   # env: kernel experiments
ABI_VERSION = 13

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64

# name -> argtypes (restype is int for all but the two listed below); mirrors include/cm3d_b200.h
PROTOTYPES = {
    "cm3d_masks_pack_dense": [_P, _P, _P, _I, _I, _P, _P],
    "cm3d_masks_decode_counts": [_P, _P, _I, _P, _P],
    "cm3d_masks_fill_rle": [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P],
    "cm3d_masks_erode3x3": [_P, _P, _P, _I, _I, _P, _P, _P],
    "cm3d_aggregate_sweeps": [_P, _P, _I, _P, _P, _P, _P, _P, _P],
    "cm3d_build_vcam_grid": [_P, _I, _I, _P, _P, _P, _P, _P],
    "cm3d_project_membership": [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "cm3d_scan_segments": [_P, _P, _P, _I, _I, _I, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "cm3d_compact_segments": [_P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P],
    "cm3d_medoid": [_P, _L, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "cm3d_medoid_items": [_I, _I],
    "cm3d_hull_obb": [_P, _L, _P, _I, _I, _I, _P, _P, _L, _P, _P, _P, _P],
    "cm3d_neighbor_filter": [_P, _L, _P, _I, _I, ctypes.c_float, _I, _P, _P, _P, _P],
    "cm3d_schedule_segments": [_P, _P, _I, _L, _P, _P, _P, _P, _P, _P],
    "cm3d_filter_segments": [_P, _P, _L, _P, _P, _P, _I, _P, _P, _P, _P],
    "cm3d_box_search": [_P, _L, _P, _I, _I, _I, _I, _P, _P, _P],
    "cm3d_nearest_lane": [_P, _I, _P, _I, _P, _P, _P],
    "cm3d_selftest_sqrt": [_P, _P],
    "cm3d_selftest_sqrt_approx": [_P, _P],
    "cm3d_lift_batch": [_P],
    "cm3d_pack_plan": [_P, _P],
    "cm3d_pack_fill": [_P, _P, _P, _P, _P, _P, _P, _P],
}
EXPORTS = ["cm3d_abi_version", "cm3d_error_string", "cm3d_hull_obb_ws_words", "cm3d_batch_args_size", "cm3d_pack_input_size"] + list(PROTOTYPES)



class BatchArgs(ctypes.Structure):
    """`cm3d_batch_args` of include/cm3d_b200.h (field for field)."""
    _fields_ = ([(n, ctypes.c_int32) for n in ("n_frames", "n_inst", "n_tiles", "n_vcams", "max_cells", "max_words", "max_runs",
                                               "max_inst_per_frame", "masks_kind", "want_obb", "obb_mode", "obb_min_pts",
                                               "screen_min_pts", "screen_flags", "max_items", "reserved")] +
                [(n, ctypes.c_int64) for n in ("bits_words", "seg_cap", "hull_ws_words", "out_words", "mask_bytes")] +
                [(n, ctypes.c_void_p) for n in (
                    "raw", "tile_sweep", "sweep_desc", "frame_desc", "vcam_desc", "cam_inst_list", "inst_desc", "chains", "mask",
                    "mask_off", "out", "frame_n", "seg_off", "item_off", "medoid_local", "medoid_point_idx", "centroid", "errflags",
                    "runs", "run_start", "row_range", "bits_raw", "bits", "bbox", "vcam_grid", "xyzw", "tile_cnt", "tile_prefix",
                    "hits", "tile_inst_cnt", "tile_inst_base", "medoid_best", "item_inst", "seg_point_idx", "seg_xyzw",
                    "screen_sums", "screen_min", "sym_ws", "screen_stats", "item_info", "obb", "hull_info", "hull_ws", "stream",
                    "stream_medoid", "launches")])


_lib = None


class Cm3dError(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises (never falls back) when it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Cm3dError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  cm3d_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.cm3d_abi_version.restype = ctypes.c_int
    lib.cm3d_error_string.restype = ctypes.c_char_p
    lib.cm3d_error_string.argtypes = [ctypes.c_int]
    lib.cm3d_hull_obb_ws_words.restype = ctypes.c_int64
    lib.cm3d_hull_obb_ws_words.argtypes = [ctypes.c_int64]
    v = lib.cm3d_abi_version()
    if v != ABI_VERSION:
        raise Cm3dError(f"libcm3d_b200.so has ABI {v}, python side expects {ABI_VERSION}: rebuild")
    lib.cm3d_batch_args_size.restype = ctypes.c_int
    if lib.cm3d_batch_args_size() != ctypes.sizeof(BatchArgs):
        raise Cm3dError(f"cm3d_batch_args is {lib.cm3d_batch_args_size()} bytes in the library, {ctypes.sizeof(BatchArgs)} in "
                        "cm3d_b200/_native.py: BatchArgs - the two declarations have drifted apart")
    # Kernel launches return in microseconds: they are bound through PyDLL, which keeps the GIL, so the launching
    # thread does not queue behind the packer threads' Python sections a dozen times per batch.  The host packer
    # (cm3d_pack_*: milliseconds of memcpy) stays on CDLL, which releases it.
    pylib = ctypes.PyDLL(LIB_PATH)
    for name, args in PROTOTYPES.items():
        for handle in (lib, pylib):
            fn = getattr(handle, name)
            fn.restype = ctypes.c_int
            fn.argtypes = args
    for name in PROTOTYPES:
        if not name.startswith("cm3d_pack_"):
            setattr(lib, name, getattr(pylib, name))
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().cm3d_error_string(rc).decode()
        raise Cm3dError(f"{what} failed: {msg} (code {rc})")


def call(name: str, *args):
    rc = getattr(load(), name)(*args)
    check(rc, name)
