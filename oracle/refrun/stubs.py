"""ORACLE tooling: stand-ins for the third-party packages the reference's lifting scripts import
but this container does not have, so that `/root/reference/src/<ds>/2d_to_3d.py` can be EXECUTED
(oracle/refrun/run.py).  Nothing here restates reference code; each stub restates, or routes to a
restatement of, the PUBLISHED behaviour of a pip dependency pinned in the reference's
environment.yml:

  pycocotools.mask.decode        -> oracle/coco_rle.py           (maskApi.c, 2.0.7)
  pyquaternion.Quaternion        -> oracle/pyquat.py             (0.9.9)
  open3d ...get_oriented_bounding_box -> oracle/obb_oracle.py    (0.15 CreateFromPoints; eigenvector SIGN unpinned)
  nuscenes.nuscenes.NuScenes, map_api.NuScenesMap, utils.splits  -> a pickled record store written next
                                    to the synthetic dataset (get / field2token / dataroot /
                                    lane, lane_connector, discretize_lanes, drivable_area)
  tensorflow.compat.v1 TFRecordDataset, waymo_open_dataset dataset_pb2.Frame, frame_utils,
  label_pb2, metrics_pb2         -> pickled duck-typed frames; the label/metrics messages are REAL
                                    google.protobuf classes built from the restated public schema
  shapely.geometry.Point, matplotlib, hdbscan, groundingdino, segment_anything, trimesh -> inert
  scipy Rotation.from_matrix     -> the reference pins scipy 1.11.4, whose from_matrix has no
                                    determinant check; the installed scipy raises on the left-handed
                                    matrices kitti:1524 feeds it, so for det < 0 ONLY the 1.11.4
                                    arithmetic (oracle/obb_oracle.scipy_from_matrix_quat) is used
"""
from __future__ import annotations

import importlib.machinery
import os
import pickle
import sys
import types

import numpy as np


class _Inert(types.ModuleType):
    """A module whose every attribute exists and does nothing (imports of unused names succeed)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, v)
        return v


def _mod(name, inert=False, **attrs):
    m = (_Inert if inert else types.ModuleType)(name)
    m.__dict__.update(attrs)
    m.__path__ = []          # every stub may have submodules
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)     # importlib.util.find_spec() wants one
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


def install_common():
    from oracle import coco_rle, pyquat
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.axes", "hdbscan", "groundingdino", "groundingdino.datasets",
                 "groundingdino.datasets.transforms", "groundingdino.models", "groundingdino.util",
                 "groundingdino.util.slconfig", "groundingdino.util.utils", "segment_anything", "trimesh",
                 "shapely", "nuscenes", "nuscenes.utils", "nuscenes.utils.data_classes",
                 "nuscenes.utils.geometry_utils", "nuscenes.map_expansion", "nuscenes.map_expansion.arcline_path_utils",
                 "nuscenes.map_expansion.bitmap"):
        _mod(name, inert=True)

    class Point:
        def __init__(self, x, y):
            self.x, self.y = x, y

        def within(self, polygon):
            return bool(polygon.contains_xy(self.x, self.y))
    _mod("shapely.geometry", inert=True, Point=Point)
    _mod("pycocotools", mask=None)
    _mod("pycocotools.mask", decode=coco_rle.decode, encode=coco_rle.encode)
    _mod("pyquaternion", Quaternion=pyquat.Quaternion)


# ------------------------------------------------------------------------------------- nuScenes devkit
class FakeNuScenes:
    def __init__(self, version, dataroot, verbose=True):
        self.dataroot = dataroot
        with open(os.path.join(dataroot, "fake_devkit.pkl"), "rb") as f:
            blob = pickle.load(f)
        self.tables, self._maps = blob["tables"], blob["maps"]

    def get(self, table, token):
        return self.tables[table][token]

    def field2token(self, table, field, value):
        return [t for t, r in self.tables[table].items() if r[field] == value]


class FakeNuScenesMap:
    def __init__(self, dataroot, map_name):
        with open(os.path.join(dataroot, "fake_devkit.pkl"), "rb") as f:
            poly = pickle.load(f)["maps"][map_name]
        toks = list(poly)
        self.lane = [{"token": t} for t in toks[::2]]
        self.lane_connector = [{"token": t} for t in toks[1::2]]
        self._poly = poly
        self.drivable_area = []

    def discretize_lanes(self, tokens, resolution_meters):
        return {t: [tuple(float(v) for v in p) for p in self._poly[t]] for t in tokens}


def install_nuscenes(dataroot):
    with open(os.path.join(dataroot, "fake_devkit.pkl"), "rb") as f:
        scenes = pickle.load(f)["scene_names"]
    _mod("nuscenes.nuscenes", inert=True, NuScenes=FakeNuScenes)
    _mod("nuscenes.map_expansion.map_api", inert=True, NuScenesMap=FakeNuScenesMap)
    _mod("nuscenes.utils.splits", mini_val=list(scenes), mini_train=[], train_detect=[], train=[], val=[])


# ------------------------------------------------------------------------------------- open3d (KITTI)
def install_open3d():
    from oracle import obb_oracle

    class _Obb:
        def __init__(self, center, extent, R):
            self.center, self.extent, self.R = np.asarray(center), np.asarray(extent), np.asarray(R)

    class PointCloud:
        points = None

        def get_oriented_bounding_box(self):
            return _Obb(*obb_oracle.open3d_obb(np.asarray(self.points, np.float64)))
    o3d = _mod("open3d")
    o3d.geometry = _mod("open3d.geometry", PointCloud=PointCloud)
    o3d.utility = _mod("open3d.utility", Vector3dVector=lambda a: np.asarray(a, np.float64))


def install_scipy_1_11_from_matrix():
    import scipy.spatial.transform as st
    from oracle import obb_oracle
    orig = st.Rotation.from_matrix

    def from_matrix(matrix, *a, **k):
        m = np.asarray(matrix, np.float64)
        if m.shape == (3, 3) and np.linalg.det(m) < 0:
            return st.Rotation.from_quat(obb_oracle.scipy_from_matrix_quat(m))
        return orig(matrix, *a, **k)
    st.Rotation.from_matrix = staticmethod(from_matrix)


# ------------------------------------------------------------------------------------- Waymo
class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _ns(d):
    """dict tree -> attribute tree (lists kept); map_features get their HasField."""
    if isinstance(d, dict):
        o = _NS(**{k: _ns(v) for k, v in d.items()})
        if "lane" in d and "polyline" in (d["lane"] or {}):
            o.HasField = lambda name: name == "lane"
        return o
    if isinstance(d, list):
        return [_ns(v) for v in d]
    return d


def _waymo_protos():
    """label_pb2 / metrics_pb2 as real protobuf classes from the restated public schema
    (waymo_open_dataset/label.proto, protos/metrics.proto; the field numbers are the published ones)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="waymo_restated.proto", package="waymo.open_dataset", syntax="proto2")
    label = fd.message_type.add(name="Label")
    box = label.nested_type.add(name="Box")
    for k, n in enumerate(("center_x", "center_y", "center_z", "width", "length", "height", "heading"), 1):
        box.field.add(name=n, number=k, type=1, label=1)                       # double, optional
    en = label.enum_type.add(name="Type")
    for n, v in (("TYPE_UNKNOWN", 0), ("TYPE_VEHICLE", 1), ("TYPE_PEDESTRIAN", 2), ("TYPE_SIGN", 3), ("TYPE_CYCLIST", 4)):
        en.value.add(name=n, number=v)
    label.field.add(name="box", number=1, type=11, label=1, type_name=".waymo.open_dataset.Label.Box")
    label.field.add(name="type", number=3, type=14, label=1, type_name=".waymo.open_dataset.Label.Type")
    label.field.add(name="id", number=4, type=9, label=1)
    obj = fd.message_type.add(name="Object")
    obj.field.add(name="object", number=1, type=11, label=1, type_name=".waymo.open_dataset.Label")
    obj.field.add(name="score", number=2, type=2, label=1)                     # float
    obj.field.add(name="context_name", number=3, type=9, label=1)
    obj.field.add(name="frame_timestamp_micros", number=4, type=3, label=1)    # int64
    objs = fd.message_type.add(name="Objects")
    objs.field.add(name="objects", number=1, type=11, label=3, type_name=".waymo.open_dataset.Object")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("waymo.open_dataset." + n))
    return get("Label"), get("Object"), get("Objects")


def install_waymo(input_path):
    """INPUT_PATH holds one pickle per scene (list of frame dict trees)."""
    Label, Object, Objects = _waymo_protos()

    class _Record:
        def __init__(self, key):
            self._key = key

        def numpy(self):
            return self._key

    registry = {}

    class TFRecordDataset:
        def __init__(self, path, compression_type=""):
            with open(path, "rb") as f:
                frames = pickle.load(f)
            self._keys = []
            for k, fr in enumerate(frames):
                key = f"{os.path.basename(path)}#{k}".encode()
                registry[bytes(key)] = fr
                self._keys.append(key)

        def __iter__(self):
            return iter(_Record(k) for k in self._keys)

    class Frame:
        def ParseFromString(self, data):
            self.__dict__.update(_ns(registry[bytes(data)]).__dict__)

    def parse_range_image_and_camera_projection(frame):
        return None, None, None, None

    def convert_range_image_to_point_cloud(frame, range_images, camera_projections, range_image_top_pose, ri_index=0,
                                           keep_polar_features=False):
        return [np.asarray(frame.points_vehicle, np.float32)], None

    tf = _mod("tensorflow", inert=True)
    _mod("tensorflow.compat", inert=True)
    v1 = _mod("tensorflow.compat.v1", inert=True)
    v1.data = _NS(TFRecordDataset=TFRecordDataset)
    tf.data = v1.data
    _mod("waymo_open_dataset", inert=True)
    _mod("waymo_open_dataset.utils", inert=True)
    _mod("waymo_open_dataset.utils.range_image_utils", inert=True)
    _mod("waymo_open_dataset.utils.transform_utils", inert=True)
    _mod("waymo_open_dataset.utils.frame_utils", parse_range_image_and_camera_projection=parse_range_image_and_camera_projection,
         convert_range_image_to_point_cloud=convert_range_image_to_point_cloud)
    _mod("waymo_open_dataset.dataset_pb2", Frame=Frame)
    _mod("waymo_open_dataset.label_pb2", Label=Label)
    _mod("waymo_open_dataset.protos", inert=True)
    _mod("waymo_open_dataset.protos.metrics_pb2", Object=Object, Objects=Objects)
