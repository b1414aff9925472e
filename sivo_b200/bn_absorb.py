"""Inference-time BN absorption: the on-disk transformation either side of the path.

Restates `caffe/scripts/BN-absorber.py` (dependencies/caffe-segnet-cudnn7/scripts/BN-absorber.py) without pycaffe:

* weights (`bn_absorber_weights`, :38-90): for every `BN` layer whose *preceding* layer in the prototxt is a `Convolution`,
  with gamma = BN blob 0 and beta = BN blob 1 (the inference scale / shift `compute_bn_statistics.py` leaves there),
  `W'[j] = W[j] * gamma[j]`, `b'[j] = b[j] * gamma[j] + beta[j]`, evaluated in float64 on copies (:67-68, 81-84) and stored
  back into float32 blobs; the BN blobs are then zeroed (:87-88) and stay in the saved model.
* prototxt (`bn_absorber_prototxt`, :93-107): the `BN` layers are removed.  (The script removes while iterating the repeated
  field, which skips the element after each removed one; in SegNet a BN is always followed by a ReLU, so every BN is found.
  Here all BN layers that were absorbed are removed; a BN that does not follow a Convolution is kept, as its weights were.)

The library runs either form: with BN layers it applies them as an epilogue affine after the bias (bn_layer.cpp:199-223 order),
without them the folded convolution.  The two differ by fp32 rounding of W * gamma only.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

import numpy as np

from .prototxt import load_net


def absorb_weights(prototxt_text: str, weights: Dict[str, List[np.ndarray]]) -> Tuple[Dict[str, List[np.ndarray]], List[str]]:
    """Returns (new weights, names of the BN layers absorbed)."""
    net = load_net(prototxt_text)
    out = {k: [np.array(b, copy=True) for b in v] for k, v in weights.items()}
    absorbed = []
    for i, layer in enumerate(net.layers):
        if layer.type != "BN" or i == 0 or net.layers[i - 1].type != "Convolution":
            continue
        conv, bn = net.layers[i - 1].name, layer.name
        if conv not in out or bn not in out or len(out[conv]) < 2 or len(out[bn]) < 2:
            raise ValueError(f"BN absorber: layers '{conv}' / '{bn}' need weight+bias and scale+shift blobs")
        w = np.array(out[conv][0], dtype=np.float64)
        b = np.array(out[conv][1], dtype=np.float64).reshape(-1)
        gamma = np.asarray(out[bn][0], dtype=np.float32).reshape(-1)
        beta = np.asarray(out[bn][1], dtype=np.float32).reshape(-1)
        if not (w.shape[0] == b.size == gamma.size == beta.size):
            raise ValueError(f"BN absorber: channel counts of '{conv}' and '{bn}' differ")
        g64 = gamma.astype(np.float64)
        out[conv][0] = (w * g64.reshape(-1, 1, 1, 1)).astype(np.float32)
        out[conv][1] = (b * g64 + beta.astype(np.float64)).astype(np.float32).reshape(out[conv][1].shape)
        out[bn][0] = np.zeros_like(out[bn][0])
        out[bn][1] = np.zeros_like(out[bn][1])
        absorbed.append(bn)
    return out, absorbed


def _layer_blocks(text: str):
    """Yields (start, end, body) of every top-level `layer { ... }` / `layers { ... }` block."""
    for m in re.finditer(r"\blayers?\s*\{", text):
        depth, j = 1, m.end()
        while j < len(text) and depth:
            depth += {"{": 1, "}": -1}.get(text[j], 0)
            j += 1
        yield m.start(), j, text[m.end():j - 1]


def absorb_prototxt(prototxt_text: str, absorbed: List[str]) -> str:
    """Drops the layer blocks of the absorbed BN layers (by name) from the prototxt text."""
    drop = set(absorbed)
    out, pos = [], 0
    for a, b, body in _layer_blocks(prototxt_text):
        name = re.search(r'\bname\s*:\s*"([^"]*)"', body)
        if name and name.group(1) in drop and re.search(r'\btype\s*:\s*"?BN"?', body):
            out.append(prototxt_text[pos:a])
            pos = b
    out.append(prototxt_text[pos:])
    return "".join(out)


def absorb(prototxt_text: str, weights: Dict[str, List[np.ndarray]]):
    """(merged prototxt text, merged weights) -- the pair BN-absorber.py writes as bn_conv_merged_model.prototxt /
    bn_conv_merged_weights.caffemodel."""
    new_w, absorbed = absorb_weights(prototxt_text, weights)
    return absorb_prototxt(prototxt_text, absorbed), new_w


def main(argv=None):
    import argparse
    import os
    from .caffemodel import read_caffemodel, write_caffemodel
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--model", required=True)
    ap.add_argument("--weights", required=True)
    ap.add_argument("--out_dir", required=True)
    a = ap.parse_args(argv)
    text = open(a.model).read()
    new_text, new_w = absorb(text, read_caffemodel(a.weights))
    os.makedirs(a.out_dir, exist_ok=True)
    open(os.path.join(a.out_dir, "bn_conv_merged_model.prototxt"), "w").write(new_text)
    net = load_net(text)
    types = {l.name: l.type for l in net.layers}
    write_caffemodel(os.path.join(a.out_dir, "bn_conv_merged_weights.caffemodel"), net.name, new_w,
                     {k: types.get(k, "Convolution") for k in new_w})


if __name__ == "__main__":
    main()
