"""cm3d_b200: B200-native 2D-mask -> 3D pseudo-label lifting (the CM3D `2d_to_3d` stage)."""
from .frames import CamSpec, FrameSpec, LiftResult, RLEMask, op_A, op_R, op_T  # noqa: F401

__version__ = "0.1.0"
