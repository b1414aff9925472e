"""TEST / BENCH TOOLING -- calibrates the seeded synthetic weights so that the untrained nets behave like trained ones
numerically: every convolution's output (after its in-place BN / ReLU chain) has rms ~1 and the logits have a standard
deviation of ~2.5, so the softmax is not saturated and the entropy map is spread over [0, log2 15] instead of being 0
almost everywhere (round-1 weights: logits rms 169 -> one-hot softmax -> `0 == 0` entropy checks).

The reference's weights are Git-LFS stubs (SURVEY "weights"), so there is nothing to calibrate *to*; this is LSUV-style
data-dependent scaling (Mishkin & Matas 2016) of the filler-drawn weights: layer by layer, in net order, run the oracle up to
the layer on a calibration image and multiply the layer's weights and bias by target / measured rms.

The per-layer factors for the two shipped topologies (seed 0) are committed as `configs/synth_scales.json`, so the product
side (`sivo_b200.caffemodel.synth_weights(..., scales=...)`) only multiplies -- it never imports the oracle.

  python tools/calibrate_synth.py        -> configs/synth_scales.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

LOGIT_STD = 2.5


def lsuv_scales(net, weights, image_bgr, T=2, seed=1234, iters=2, verbose=False):
    """Returns {conv layer name: factor}; `weights` is modified in place (conv weight and bias blobs multiplied)."""
    from oracle import segnet_oracle as S
    scales = {}
    layers = net.layers
    for li, ly in enumerate(layers):
        if ly.type != "Convolution":
            continue
        end = li  # last in-place follower (BN / ReLU) of this convolution
        while end + 1 < len(layers) and layers[end + 1].type in ("BN", "ReLU") and layers[end + 1].bottoms[0] == ly.tops[0] \
                and layers[end + 1].tops[0] == ly.tops[0]:
            end += 1
        feeds_softmax = any(l.type == "Softmax" and l.bottoms[0] == ly.tops[0] for l in layers[end + 1:])
        total = 1.0
        for _ in range(iters):
            blobs = S.forward(net, weights, image_bgr, seed=seed, frame=0, precision="fp32", T=T, stop_after=layers[end].name)
            y = blobs[ly.tops[0]].numpy().astype(np.float64)
            if feeds_softmax:  # spread of the logits across classes is what shapes the softmax
                cur, target = float((y - y.mean(axis=1, keepdims=True)).std()), LOGIT_STD
            else:
                cur, target = float(np.sqrt((y * y).mean())), 1.0
            f = np.float32(target / max(cur, 1e-30))
            for b in weights[ly.name]:
                b *= f
            total *= float(f)
        scales[ly.name] = total
        if verbose:
            print(f"{ly.name:28s} x {total:.6g}")
    return scales


def calibration_image(H, W):
    from sivo_b200.synth import stereo_frame
    left, _ = stereo_frame(0)
    return np.ascontiguousarray(left[100:100 + H, 109:109 + W])


def main():
    import gen_prototxt
    from sivo_b200.caffemodel import synth_weights
    from sivo_b200.prototxt import load_net
    out = {"_doc": "per-layer multipliers for sivo_b200.caffemodel.synth_weights(seed=0); written by tools/calibrate_synth.py "
                   "(LSUV on synthetic frame 0, crop 160x512 at (109,100), T=2, dropout seed 1234)"}
    H, W = 160, 512
    for kind in ("basic", "standard"):
        net = load_net(getattr(gen_prototxt, kind)(T=2, H=H, W=W))
        w = synth_weights(net, 0)
        out[kind] = lsuv_scales(net, w, calibration_image(H, W), verbose=True)
    path = os.path.join(ROOT, "configs", "synth_scales.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
