import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    import numpy as np
    from cm3d_b200.frames import frame_from_arrays
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    frame = frame_from_arrays({k[3:]: d[k] for k in d.files if k.startswith("in_")})
    return frame, d


GOLDEN = ["nusc_small", "nusc_edge", "nusc_c1", "kitti_small", "waymo_small"]
