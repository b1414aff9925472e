"""Packed per-frame record exchanged between ranks when frames are sharded one per GPU: one
fixed-size byte buffer per frame so a single all-gather per step moves everything the host SLAM consumes:
header | classes u8 [H*W] | confidence f32 [H*W] | entropy f32 [H*W] | 2 x (keypoints[cap] 28 B, descriptors[cap] 32 B)
-- 3.5 MB per frame at 1024x352 with nfeatures = 2000.  The two maps travel in single precision; the operator's own
double maps stay on the producing rank's device."""
from __future__ import annotations

import numpy as np

from .orb import KP_DTYPE

HEADER = 32  # frame_id i64, n_left i64, n_right i64, reserved i64


def _maps_end(hw: int) -> int:
    return HEADER + (hw + 15) // 16 * 16 + hw * 8


def record_bytes(hw: int, kp_cap: int) -> int:
    n = _maps_end(hw) + 2 * kp_cap * 60
    return (n + 255) // 256 * 256


def offsets(hw: int, kp_cap: int):
    c = HEADER + (hw + 15) // 16 * 16  # the f32 maps start on a 16-byte boundary
    o = {"classes": HEADER, "confidence": c, "entropy": c + hw * 4}
    base = _maps_end(hw)
    o["kp_left"], o["desc_left"] = base, base + kp_cap * 28
    o["kp_right"], o["desc_right"] = base + kp_cap * 60, base + kp_cap * 60 + kp_cap * 28
    return o


def pack_host_part(buf: np.ndarray, hw: int, kp_cap: int, frame_id: int, kl, dl, kr, dr) -> None:
    """Header + keypoints/descriptors (they live on the host after the quad tree); the three maps are copied
    device-to-device into the same record by the caller."""
    o = offsets(hw, kp_cap)
    nl, nr = len(kl), len(kr)
    if nl > kp_cap or nr > kp_cap:
        raise ValueError("more keypoints than the record holds")
    buf[:HEADER].view(np.int64)[:] = (frame_id, nl, nr, 0)
    buf[o["kp_left"]:o["kp_left"] + nl * 28] = np.ascontiguousarray(kl, KP_DTYPE).view(np.uint8).reshape(-1)
    buf[o["desc_left"]:o["desc_left"] + nl * 32] = np.ascontiguousarray(dl, np.uint8).reshape(-1)
    buf[o["kp_right"]:o["kp_right"] + nr * 28] = np.ascontiguousarray(kr, KP_DTYPE).view(np.uint8).reshape(-1)
    buf[o["desc_right"]:o["desc_right"] + nr * 32] = np.ascontiguousarray(dr, np.uint8).reshape(-1)


def pack_host(buf: np.ndarray, hw: int, kp_cap: int, frame_id: int, classes, conf, ent, kl, dl, kr, dr) -> None:
    """The whole record from host results (the e2e path, where segmentImage has already copied the maps to the host)."""
    o = offsets(hw, kp_cap)
    buf[o["classes"]:o["classes"] + hw] = classes.reshape(-1)
    buf[o["confidence"]:o["confidence"] + hw * 4].view(np.float32)[:] = conf.reshape(-1)  # f64 -> f32, round to nearest
    buf[o["entropy"]:o["entropy"] + hw * 4].view(np.float32)[:] = ent.reshape(-1)
    pack_host_part(buf, hw, kp_cap, frame_id, kl, dl, kr, dr)


def unpack(buf: np.ndarray, h: int, w: int, kp_cap: int):
    hw = h * w
    o = offsets(hw, kp_cap)
    frame_id, nl, nr, _ = (int(v) for v in buf[:HEADER].view(np.int64))
    out = {"frame_id": frame_id,
           "classes": buf[o["classes"]:o["classes"] + hw].reshape(h, w),
           "confidence": buf[o["confidence"]:o["confidence"] + hw * 4].view(np.float32).reshape(h, w),
           "entropy": buf[o["entropy"]:o["entropy"] + hw * 4].view(np.float32).reshape(h, w),
           "kp_left": buf[o["kp_left"]:o["kp_left"] + nl * 28].view(KP_DTYPE),
           "desc_left": buf[o["desc_left"]:o["desc_left"] + nl * 32].reshape(nl, 32),
           "kp_right": buf[o["kp_right"]:o["kp_right"] + nr * 28].view(KP_DTYPE),
           "desc_right": buf[o["desc_right"]:o["desc_right"] + nr * 32].reshape(nr, 32)}
    return out
