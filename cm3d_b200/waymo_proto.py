"""Serialiser for the Waymo Open Dataset `metrics.Objects` message the reference writes
(src/waymo/2d_to_3d.py:1034-1065,1244-1275,1300-1305) - `objects.SerializeToString()`.

When `waymo_open_dataset` is importable its generated classes are used.  Otherwise (this image:
not installed, no network) the protobuf wire format is written directly from the published
schema (waymo_open_dataset/protos/metrics.proto, label.proto):

    Objects { repeated Object objects = 1; }
    Object  { Label object = 1; float score = 2; string context_name = 3; int64 frame_timestamp_micros = 4; }
    Label   { Box box = 1; Type type = 3; string id = 4; }
    Box     { double center_x = 1, center_y = 2, center_z = 3, width = 4, length = 5, height = 6, heading = 7; }
    Type    { UNKNOWN = 0, VEHICLE = 1, PEDESTRIAN = 2, SIGN = 3, CYCLIST = 4 }

PARITY UNPINNED: the .proto files are not in the reference tree; field numbers are restated from
the published schema.  `parse_objects` reads the same subset back (tests round-trip through it).
"""
from __future__ import annotations

import struct
from typing import List

TYPE_UNKNOWN, TYPE_VEHICLE, TYPE_PEDESTRIAN, TYPE_SIGN, TYPE_CYCLIST = 0, 1, 2, 3, 4
TYPE_BY_NAME = {"vehicle": TYPE_VEHICLE, "cyclist": TYPE_CYCLIST, "pedestrian": TYPE_PEDESTRIAN}
_BOX_FIELDS = (("center_x", 1), ("center_y", 2), ("center_z", 3), ("width", 4), ("length", 5), ("height", 6), ("heading", 7))


def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def _ld(field: int, payload: bytes) -> bytes:
    return _key(field, 2) + _varint(len(payload)) + payload


_BOX = struct.Struct("<" + "Bd" * 7)                         # seven (key, double) pairs: keys (n << 3) | 1
_BOX_KEYS = tuple((n << 3) | 1 for _, n in _BOX_FIELDS)
_TAILS = {}                                                   # per (context_name, timestamp) / per id: the constant bytes


_F32 = struct.Struct("<f")
_HEADS = {}                                                   # per (type, id): the bytes before and after the box


def encode_object(o: dict) -> bytes:
    """One `Object` (field order and presence as `SerializeToString` of the proto2 message with every field set)."""
    k = _BOX_KEYS
    box = _BOX.pack(k[0], float(o["center_x"]), k[1], float(o["center_y"]), k[2], float(o["center_z"]), k[3], float(o["width"]),
                    k[4], float(o["length"]), k[5], float(o["height"]), k[6], float(o["heading"]))
    hk = (int(o["type"]), o.get("id", "unique object tracking ID"))
    head = _HEADS.get(hk)
    if head is None:                        # Object.object = Label { box = 1 (63 bytes), type = 3, id = 4 }
        after = b"\x18" + _varint(hk[0]) + _ld(4, hk[1].encode())
        head = _HEADS[hk] = (b"\x0a" + _varint(2 + _BOX.size + len(after)) + b"\x0a" + _varint(_BOX.size), after)
    ck = (o["context_name"], int(o["frame_timestamp_micros"]))
    tail = _TAILS.get(ck)
    if tail is None:
        if len(_TAILS) > 65536:
            _TAILS.clear()
        tail = _TAILS[ck] = _ld(3, ck[0].encode()) + _key(4, 0) + _varint(ck[1])
    return head[0] + box + head[1] + b"\x15" + _F32.pack(float(o["score"])) + tail


def serialize_objects(objects: List[dict]) -> bytes:
    """objects: dicts with context_name, frame_timestamp_micros, score, type, center_x/y/z, length, width,
    height, heading."""
    try:
        from waymo_open_dataset import label_pb2
        from waymo_open_dataset.protos import metrics_pb2
    except Exception:
        return b"".join(_ld(1, encode_object(o)) for o in objects)
    msg = metrics_pb2.Objects()
    for d in objects:
        o = metrics_pb2.Object()
        o.context_name = d["context_name"]
        o.frame_timestamp_micros = int(d["frame_timestamp_micros"])
        b = label_pb2.Label.Box()
        for name, _ in _BOX_FIELDS:
            setattr(b, name, float(d[name]))
        o.object.box.CopyFrom(b)
        o.score = float(d["score"])
        o.object.id = d.get("id", "unique object tracking ID")
        o.object.type = int(d["type"])
        msg.objects.append(o)
    return msg.SerializeToString()


# ------------------------------------------------------------------ reader (tests, merging shards)
def _read_varint(buf: bytes, p: int):
    v, s = 0, 0
    while True:
        b = buf[p]
        p += 1
        v |= (b & 0x7F) << s
        s += 7
        if not b & 0x80:
            return v, p


def _fields(buf: bytes):
    p = 0
    while p < len(buf):
        k, p = _read_varint(buf, p)
        f, w = k >> 3, k & 7
        if w == 0:
            v, p = _read_varint(buf, p)
        elif w == 1:
            v, p = buf[p:p + 8], p + 8
        elif w == 5:
            v, p = buf[p:p + 4], p + 4
        elif w == 2:
            n, p = _read_varint(buf, p)
            v, p = buf[p:p + n], p + n
        else:
            raise ValueError("unsupported wire type")
        yield f, w, v


def parse_objects(data: bytes) -> List[dict]:
    out = []
    names = {n: name for name, n in _BOX_FIELDS}
    for f, w, v in _fields(data):
        if f != 1:
            continue
        o = {}
        for f2, w2, v2 in _fields(v):
            if f2 == 1:
                for f3, w3, v3 in _fields(v2):
                    if f3 == 1:
                        for f4, w4, v4 in _fields(v3):
                            o[names[f4]] = struct.unpack("<d", v4)[0]
                    elif f3 == 3:
                        o["type"] = v3
                    elif f3 == 4:
                        o["id"] = v3.decode()
            elif f2 == 2:
                o["score"] = struct.unpack("<f", v2)[0]
            elif f2 == 3:
                o["context_name"] = v2.decode()
            elif f2 == 4:
                o["frame_timestamp_micros"] = v2 - (1 << 64) if v2 >= (1 << 63) else v2
        out.append(o)
    return out
