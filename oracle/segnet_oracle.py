"""TEST INFRASTRUCTURE -- CPU oracle for `SIVO::BayesianSegNet::segmentImage`; never imported by
the product path (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference).

Restates, layer by layer in fp32 NCHW (float64 for the MC reduction), what the reference executes:

  preprocess      src/bayesian_segnet/bayesian_segnet.cpp:142-178  (centre crop, u8 BGR -> f32 planar, T copies)
  Convolution     caffe/src/caffe/layers/conv_layer.cpp:25-40 + base_conv_layer.cpp:257-280 (im2col GEMM + bias)
  ReLU            caffe/src/caffe/layers/relu_layer.cpp:9-19
  BN (INFERENCE)  caffe/src/caffe/layers/bn_layer.cpp:199-223      y = x*scale[c] + shift[c]  (mul, then add)
  LRN             caffe/src/caffe/layers/lrn_layer.cpp:108-152     y = x * (k + alpha/n * sum_win x^2)^-beta
  Pooling+mask    caffe/src/caffe/layers/pooling_layer.cpp:140-187 strict '>' scan => first max wins; mask = h*W+w
  Upsample        caffe/src/caffe/layers/upsample_layer.cpp:74-103 zero fill, top[mask[i]] = bottom[i]
  Dropout         caffe/src/caffe/layers/dropout_layer.cpp:31-46   y = x * mask * 1/(1-ratio)  (mask: oracle/philox.py)
  Softmax         caffe/src/caffe/layers/softmax_layer.cpp:27-60   max-subtract, exp, sum, divide (fp32)
  MC reduction    src/bayesian_segnet/bayesian_segnet.cpp:278-318,38-44  (double: mean over T, first-max argmax,
                                                                    max, -sum p log2 p with 0 log 0 := 0)

PARITY STATUS.  Pinned by the reference's own known-answer tests (tests/test_oracle_segnet.py):
pool+mask 3x5 KAT (caffe/src/caffe/test/test_pooling_layer.cpp:57-118), upsample KAT and pool->unpool
round trip (test_upsample_layer.cpp:58-105,193-246), softmax sums (test_softmax_layer.cpp:43-75), LRN vs
the reference loop (test_lrn_layer.cpp:55-...), conv vs the naive loop (test_convolution_layer.cpp:22-139),
dropout survivor rule (test_neuron_layer.cpp:63-88).  UNPINNED (the reference holds no vector and its Caffe
cannot be built here -- SURVEY 8c): dropout mask bits (reference is unseeded), GEMM summation order (cuDNN7 /
BLAS internal), end-to-end `segmentImage` outputs (reference test checks sizes only,
tests/test_bayesian_segnet.cpp:152-168; weights are Git-LFS stubs).

`precision="fp16"` models the tensor-core path exactly up to fp32 summation order: every convolution's
input activations and weights are rounded to IEEE half first (products are then exact in fp32), and every
convolution output is rounded to half after bias/BN/ReLU -- except the one feeding Softmax, which stays fp32.
Pool / unpool / dropout(x2) are exact on half values.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from oracle.philox import dropout_keep


def center_crop(image: np.ndarray, width: int, height: int) -> np.ndarray:
    """`BayesianSegNet::resizeImage` (bayesian_segnet.cpp:142-162).  Returns None (the reference
    returns an empty Mat) if the image is smaller than the network input."""
    rows, cols = image.shape[:2]
    if (cols, rows) == (width, height):
        return image
    if rows >= height and cols >= width:
        x_tl = cols // 2 - width // 2
        y_tl = rows // 2 - height // 2
        return np.ascontiguousarray(image[y_tl:y_tl + height, x_tl:x_tl + width])
    return None


def _h(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float16).to(torch.float32)


def lrn_across(x: torch.Tensor, size: int, alpha: float, beta: float, k: float) -> torch.Tensor:
    n, c, h, w = x.shape
    pre = (size - 1) // 2
    sq = F.pad(x * x, (0, 0, 0, 0, pre, size - 1 - pre))
    acc = torch.zeros_like(x)
    for i in range(size):
        acc = acc + sq[:, i:i + c]
    scale = np.float32(k) + acc * np.float32(alpha / size)
    return x * torch.pow(scale, np.float32(-beta))


def pool_with_mask(x: torch.Tensor):
    """2x2/2 max pool; mask = plane-local index h*W+w of the first maximum (as Caffe stores it)."""
    n, c, h, w = x.shape
    best = torch.full((n, c, h // 2, w // 2), -torch.inf, dtype=x.dtype)
    idx = torch.full((n, c, h // 2, w // 2), -1, dtype=torch.int64)
    hh = torch.arange(0, h, 2).view(-1, 1)
    ww = torch.arange(0, w, 2).view(1, -1)
    for dh in (0, 1):
        for dw in (0, 1):
            v = x[:, :, dh::2, dw::2]
            take = v > best
            best = torch.where(take, v, best)
            idx = torch.where(take, ((hh + dh) * w + ww + dw).expand_as(idx), idx)
    return best, idx


def unpool(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    n, c, h, w = x.shape
    out = torch.zeros((n, c, h * 2 * w * 2), dtype=x.dtype)
    out.scatter_(2, idx.reshape(n, c, -1).expand(n, c, -1), x.reshape(n, c, -1))
    return out.view(n, c, h * 2, w * 2)


def softmax_channels(x: torch.Tensor) -> torch.Tensor:
    m = x.max(dim=1, keepdim=True).values
    e = torch.exp(x - m)
    return e / e.sum(dim=1, keepdim=True)


def mc_reduce(prob: np.ndarray):
    """prob float32 [T, C, H, W] -> classes u8 [H,W], confidence f64 [H,W], entropy f64 [H,W]."""
    p = prob.astype(np.float64)
    mean = p.mean(axis=0)
    classes = np.argmax(mean, axis=0).astype(np.uint8)  # first max wins, as Eigen's argmax
    conf = mean.max(axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.where(mean == 0.0, 0.0, -mean * np.log2(mean))
    return classes, conf, e.sum(axis=0)


def forward(net, weights: Dict[str, List[np.ndarray]], image_bgr: np.ndarray, seed: int = 1234,
            frame: int = 0, precision: str = "fp32", T: Optional[int] = None, dedup: bool = True,
            threads: Optional[int] = None, return_blobs: bool = False, masks: Optional[Dict[str, np.ndarray]] = None,
            stop_after: Optional[str] = None):
    """Runs the net on one (already cropped or larger) BGR u8 image; returns prob [T,C,H,W] float32.

    `masks` (blob name -> Caffe-style plane index array) replaces the argmax decisions of the named pooling
    layers: the pooled value is gathered at the given index.  Tests use it to hand the oracle the device's
    tie-breaks -- two window entries that differ by <= 1 half ulp may legitimately swap order between two fp32
    summation orders, and one swapped position moves a whole activation under the next 7x7 filter -- so that
    everything downstream can be compared tightly; the swapped positions themselves are checked to be such ties.

    `stop_after` (layer name) ends the pass after that layer and returns the blob dict so far (used by the synthetic-weight
    calibration, tools/calibrate_synth.py)."""
    if threads:
        torch.set_num_threads(threads)
    T = T or net.T
    if T is None:
        raise ValueError("T (first input dim) is not set")
    _, C, H, W = net.input_dims
    img = center_crop(image_bgr, W, H)
    if img is None:
        raise ValueError("image smaller than the network input")
    half = precision == "fp16"
    x0 = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)).astype(np.float32))[None]
    blobs: Dict[str, torch.Tensor] = {net.input_name: x0 if dedup else x0.repeat(T, 1, 1, 1)}
    consumers: Dict[str, List[str]] = {}
    for ly in net.layers:
        for b in ly.bottoms:
            consumers.setdefault(b, []).append(ly.type)
    drop_idx = 0
    with torch.no_grad():
        for li, ly in enumerate(net.layers):
            x = blobs[ly.bottoms[0]]
            t = ly.type
            if t == "Convolution":
                w = torch.from_numpy(weights[ly.name][0])
                b = torch.from_numpy(weights[ly.name][1]) if len(weights[ly.name]) > 1 else None
                if half:
                    x, w = _h(x), _h(w)
                y = F.conv2d(x, w, None, padding=ly.pad)
                if b is not None:
                    y = y + b.view(1, -1, 1, 1)
                # in-place followers (BN / ReLU) are applied by their own layers below; the half
                # rounding of the stored activation happens after the last in-place follower.
                blobs[ly.tops[0]] = y
                blobs["__pending_round__" + ly.tops[0]] = torch.tensor(1)
            elif t == "BN":
                s = torch.from_numpy(weights[ly.name][0]).view(1, -1, 1, 1)
                sh = torch.from_numpy(weights[ly.name][1]).view(1, -1, 1, 1)
                blobs[ly.tops[0]] = x * s + sh
            elif t == "ReLU":
                blobs[ly.tops[0]] = torch.clamp_min(x, 0) + ly.negative_slope * torch.clamp_max(x, 0)
            elif t == "LRN":
                blobs[ly.tops[0]] = lrn_across(x, ly.local_size, ly.alpha, ly.beta, ly.k)
            elif t == "Pooling":
                v, m = pool_with_mask(x)
                if masks is not None and ly.tops[1] in masks:
                    m = torch.from_numpy(np.asarray(masks[ly.tops[1]]).astype(np.int64))
                    if m.shape[0] != x.shape[0]:
                        m = m[:x.shape[0]]
                    v = torch.gather(x.reshape(x.shape[0], x.shape[1], -1), 2, m.reshape(m.shape[0], m.shape[1], -1)).view(m.shape)
                blobs[ly.tops[0]], blobs[ly.tops[1]] = v, m
            elif t == "Upsample":
                m = blobs[ly.bottoms[1]]
                if m.shape[0] != x.shape[0]:
                    m = m.expand(x.shape[0], -1, -1, -1)
                blobs[ly.tops[0]] = unpool(x, m)
            elif t == "Dropout":
                if not ly.sample_weights_test:
                    blobs[ly.tops[0]] = x  # plain test-time dropout is the identity
                else:
                    if x.shape[0] == 1 and T > 1:
                        x = x.repeat(T, 1, 1, 1)
                        # masks produced upstream stay sample-invariant (batch 1) and are broadcast on use
                    keep = torch.from_numpy(dropout_keep(seed, frame, drop_idx, T, *x.shape[1:]))
                    scale = np.float32(1.0 / (1.0 - ly.dropout_ratio))
                    blobs[ly.tops[0]] = x * keep.to(torch.float32) * scale
                drop_idx += 1
            elif t == "Softmax":
                blobs[ly.tops[0]] = softmax_channels(x)
            else:
                raise ValueError(t)
            # fp16 model: round the conv output once its in-place chain (BN, ReLU) is finished
            if half:
                top = ly.tops[0]
                key = "__pending_round__" + top
                if key in blobs:
                    nxt = net.layers[li + 1] if li + 1 < len(net.layers) else None
                    chain_continues = nxt is not None and nxt.type in ("BN", "ReLU") and nxt.bottoms[0] == top \
                        and nxt.tops[0] == top
                    if not chain_continues:
                        del blobs[key]
                        if "Softmax" not in consumers.get(top, []):
                            blobs[top] = _h(blobs[top])
            if stop_after is not None and ly.name == stop_after:
                return {k: v for k, v in blobs.items() if not k.startswith("__")}
    prob = blobs[net.layers[-1].tops[0]]
    if prob.shape[0] == 1 and T > 1:
        prob = prob.repeat(T, 1, 1, 1)
    out = prob.numpy()
    if return_blobs:
        return out, {k: v for k, v in blobs.items() if not k.startswith("__")}
    return out


def segment_image(net, weights, image_bgr, **kw):
    """Oracle of the operator: (classes u8, confidence f64, entropy f64), each [H, W]."""
    return mc_reduce(forward(net, weights, image_bgr, **kw))
