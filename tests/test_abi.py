"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/cm3d_b200.h
declares (no compute calls - there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from cm3d_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cm3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cm3d_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path(lib_path):
    syms = declared_symbols()
    for s in ["cm3d_masks_pack_dense", "cm3d_masks_fill_rle", "cm3d_masks_erode3x3", "cm3d_aggregate_sweeps",
              "cm3d_project_membership", "cm3d_scan_segments", "cm3d_compact_segments", "cm3d_medoid"]:
        assert s in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/cm3d_b200.h but not exported"


def test_python_prototypes_cover_the_header(lib_path):
    from cm3d_b200 import _native as N
    assert sorted(N.EXPORTS) == declared_symbols()
    lib = N.load()
    assert lib.cm3d_abi_version() == N.ABI_VERSION
    assert lib.cm3d_error_string(0) == b"ok"
    assert b"invalid" in lib.cm3d_error_string(-1)


def test_header_constants_match_python():
    from cm3d_b200 import batch as B, frames as F
    text = open(os.path.join(ROOT, "include", "cm3d_b200.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(CM3D_[A-Z_]+)\s+(\d+)\b", text)}
    assert defs["CM3D_TILE"] == B.TILE and defs["CM3D_MAX_INST"] == B.MAX_INST
    assert defs["CM3D_MAX_VCAMS"] == B.MAX_VCAMS and defs["CM3D_MEDOID_COLS"] == B.MEDOID_COLS
    assert defs["CM3D_MAX_CHAIN"] == F.MAX_CHAIN and defs["CM3D_OP_WORDS"] == F.OP_WORDS


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from cm3d_b200 import _native as N
    from cm3d_b200.lifter import Lifter
    with pytest.raises(N.Cm3dError):
        Lifter("cuda:0")


def test_sass_is_sm100a_with_bulk_copy_and_packed_fma(lib_path):
    """The built library holds sm_100a code; the aggregate kernel stages tiles with a bulk async
    copy (UBLKCP) and the medoid kernel uses packed FFMA2."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out
    assert "FFMA2" in out


def test_batch_args_mirror_and_null_call(lib_path):
    """The ctypes mirror of `cm3d_batch_args` has the library's size (checked at load, and here), and the one-call
    entry point rejects a null argument block before it touches CUDA."""
    import ctypes
    from cm3d_b200 import _native as N
    lib = N.load()
    assert lib.cm3d_batch_args_size() == ctypes.sizeof(N.BatchArgs)
    assert lib.cm3d_lift_batch(None) == -1          # CM3D_EINVAL
