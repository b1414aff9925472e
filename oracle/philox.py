"""TEST INFRASTRUCTURE -- CPU oracle, never imported by the product path.

Philox4x32-10 counter-based generator (Salmon et al., SC'11; Random123 `philox4x32_R(10, ...)`)
and the dropout-mask rule this build specifies on top of it.

Why a specified generator: the reference samples MC-dropout masks at test time from an
*unseeded* stream -- `caffe_rng_bernoulli` on boost::mt19937 (caffe/src/caffe/layers/dropout_layer.cpp:37-42,
caffe/src/caffe/util/math_functions.cpp:309-320) or cuRAND XORWOW (dropout_layer.cu:9-35), seeded from
/dev/urandom (caffe/src/caffe/common.cpp:23-40) because SIVO never calls `Caffe::set_random_seed`.
Mask *bits* are therefore unpinnable; what is pinned is the distribution (Bernoulli(1-ratio) per
element, independent per MC sample) and the survivor rule `y = x * mask * 1/(1-ratio)`
(caffe/src/caffe/test/test_neuron_layer.cpp:63-88).  Oracle and CUDA path share this rule:

    keep(seed, frame, layer, n, h, w, c) = bit (c & 31) of word ((c >> 5) & 3) of
        philox4x32_10(ctr = (h*W + w, n | layer << 16 | (c >> 7) << 24, frame_lo, frame_hi),
                      key = (seed_lo, seed_hi))

`layer` is the ordinal of the Dropout layer in the prototxt, `n` the MC sample, (h, w, c) the
element of that layer's blob.  One Philox call yields the 128 keep-bits of 128 consecutive channels
of one pixel, which is what one epilogue thread of the CUDA conv kernel owns.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (broadcast).  Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(np.asarray(x & MASK32).astype(np.uint32) for x in (c0, c1, c2, c3))


def dropout_keep(seed: int, frame: int, layer: int, T: int, C: int, H: int, W: int) -> np.ndarray:
    """bool [T, C, H, W] keep-mask for Dropout layer ordinal `layer` (ratio 0.5)."""
    assert T < (1 << 16) and layer < 256 and C <= 256 * 128
    pix = np.arange(H * W, dtype=np.uint32)[None, :, None]
    n = np.arange(T, dtype=np.uint32)[:, None, None]
    ngroups = (C + 127) // 128
    g = np.arange(ngroups, dtype=np.uint32)[None, None, :]
    c1 = n | np.uint32(layer << 16) | (g << np.uint32(24))
    words = philox4x32_10(pix, c1, np.uint32(frame & 0xFFFFFFFF), np.uint32((frame >> 32) & 0xFFFFFFFF),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    # words[i]: [T, H*W, ngroups]  -> bits [T, H*W, ngroups, 4, 32]
    w = np.stack(words, axis=-1)
    bits = (w[..., None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
    bits = bits.reshape(T, H * W, ngroups * 128)[:, :, :C]
    return bits.astype(bool).transpose(0, 2, 1).reshape(T, C, H, W)
