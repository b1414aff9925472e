"""`.caffemodel` wire-format writer / reader (no protoc in the image) and seeded synthetic weights.

The reference loads weights with `Net::CopyTrainedLayersFrom` (caffe/src/caffe/net.cpp:788-803,
750-785: layers matched by *name*, blob count and shape must agree).  On disk that is a binary
`NetParameter` (caffe/src/caffe/proto/caffe.proto:64-96): `name=1`, repeated `layer=100`
(`LayerParameter`: `name=1`, `type=2`, repeated `blobs=7`), each `BlobProto` (:10-22) carrying
`shape=7 { dim=1 packed int64 }` and `data=5 packed float` (legacy `num/channels/height/width=1..4`).

The shipped `*.caffemodel` files are Git-LFS stubs, so tests and the bench use weights drawn from
the prototxt's own fillers (`msra` / `xavier`) with a fixed seed -- see `synth_weights`.
The product-side reader is C++ (`sivo_b200/csrc/caffemodel.cc`); `read_caffemodel` here is the
Python twin used by the oracle.
"""
from __future__ import annotations

import struct
from typing import Dict, List

import numpy as np

from .prototxt import NetSpec, param_shapes


def _varint(n: int) -> bytes:
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _blob(arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr, dtype="<f4")
    dims = b"".join(_varint(int(d)) for d in arr.shape)
    shape = _ld(1, dims)  # BlobShape.dim packed
    return _ld(7, shape) + _ld(5, arr.tobytes())


def write_caffemodel(path: str, net_name: str, layers: Dict[str, List[np.ndarray]],
                     types: Dict[str, str]) -> None:
    body = _ld(1, net_name.encode())
    for name, blobs in layers.items():
        lp = _ld(1, name.encode()) + _ld(2, types[name].encode())
        for b in blobs:
            lp += _ld(7, _blob(b))
        body += _ld(100, lp)
    with open(path, "wb") as f:
        f.write(body)


def _read_varint(buf: memoryview, pos: int):
    shift = 0
    val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not (b & 0x80):
            return val, pos
        shift += 7


def _fields(buf: memoryview):
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
            yield field, wt, v
        elif wt == 1:
            yield field, wt, bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            yield field, wt, buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            yield field, wt, bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"caffemodel: unsupported wire type {wt}")


def _parse_blob(buf: memoryview) -> np.ndarray:
    shape = None
    legacy = {}
    data = None
    scalars: List[float] = []
    for field, wt, v in _fields(buf):
        if field == 7 and wt == 2:
            dims = []
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 2:
                    p = 0
                    while p < len(v2):
                        d, p = _read_varint(v2, p)
                        dims.append(d)
                elif f2 == 1 and w2 == 0:
                    dims.append(v2)
            shape = tuple(dims)
        elif field == 5 and wt == 2:
            data = np.frombuffer(bytes(v), dtype="<f4")
        elif field == 5 and wt == 5:
            scalars.append(struct.unpack("<f", v)[0])
        elif field in (1, 2, 3, 4) and wt == 0:
            legacy[field] = v
    if data is None:
        data = np.asarray(scalars, dtype=np.float32)
    if shape is None:
        shape = tuple(legacy.get(i, 1) for i in (1, 2, 3, 4))
    return data.reshape(shape).copy()


def read_caffemodel(path: str) -> Dict[str, List[np.ndarray]]:
    with open(path, "rb") as f:
        raw = f.read()
    if raw.startswith(b"version https://git-lfs"):
        raise ValueError(f"{path} is a Git-LFS pointer stub, not a caffemodel")
    out: Dict[str, List[np.ndarray]] = {}
    for field, wt, v in _fields(memoryview(raw)):
        if field == 100 and wt == 2:
            name = None
            blobs = []
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 2:
                    name = bytes(v2).decode()
                elif f2 == 7 and w2 == 2:
                    blobs.append(_parse_blob(v2))
            out[name] = blobs
        elif field == 2 and wt == 2:
            raise ValueError("caffemodel: V1LayerParameter ('layers' = 2) is not supported")
    return out


def shipped_scales(kind: str) -> Dict[str, float]:
    """The committed per-layer calibration factors of the two shipped topologies (configs/synth_scales.json, written by
    tools/calibrate_synth.py): with them the seeded nets keep O(1) activations and logits of std ~2.5 at any input size."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "configs", "synth_scales.json")
    return {k: float(v) for k, v in json.load(open(path))[kind].items()}


def synth_weights(net: NetSpec, seed: int = 0, scales: Dict[str, float] = None) -> Dict[str, List[np.ndarray]]:
    """Seeded weights from the prototxt's fillers: `msra` = N(0, sqrt(2/fan_in)),
    `xavier` = U(+-sqrt(3/fan_in)) (caffe/include/caffe/filler.hpp); biases N(0, 0.1) and BN
    scale 1+N(0,0.1), shift N(0,0.1) so that neither path is trivially the identity.
    `scales` (conv layer name -> factor) multiplies that layer's weight and bias blobs: the calibration that keeps the
    untrained net's activations O(1) and its softmax unsaturated (tools/calibrate_synth.py)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, List[np.ndarray]] = {}
    ltypes = {ly.name: ly for ly in net.layers}
    for name, shapes in param_shapes(net).items():
        ly = ltypes[name]
        if ly.type == "Convolution":
            co, ci, kh, kw = shapes[0]
            fan_in = ci * kh * kw
            if ly.weight_filler == "xavier":
                lim = np.sqrt(3.0 / fan_in)
                w = rng.uniform(-lim, lim, size=shapes[0])
            else:
                w = rng.normal(0.0, np.sqrt(2.0 / fan_in), size=shapes[0])
            blobs = [w.astype(np.float32)]
            if len(shapes) > 1:
                blobs.append(rng.normal(0.0, 0.1, size=shapes[1]).astype(np.float32))
            if scales and name in scales:
                blobs = [b * np.float32(scales[name]) for b in blobs]
            out[name] = blobs
        else:  # BN: blobs[0] = scale, blobs[1] = shift (caffe/src/caffe/layers/bn_layer.cpp:62-74)
            out[name] = [(1.0 + rng.normal(0.0, 0.1, size=shapes[0])).astype(np.float32),
                         rng.normal(0.0, 0.1, size=shapes[1]).astype(np.float32)]
    return out


def write_synth_model(net: NetSpec, path: str, seed: int = 0, scales: Dict[str, float] = None) -> Dict[str, List[np.ndarray]]:
    w = synth_weights(net, seed, scales)
    write_caffemodel(path, net.name, w, {ly.name: ly.type for ly in net.layers})
    return w
