"""-m gpu: `BayesianSegNet::segmentImage` through the C-ABI against the oracle -- blob by blob on small
nets (identical weights and dropout masks), against the committed golden fixtures, and at the full
1024x352 geometry through size-independent properties."""
import os

import numpy as np
import pytest

from conftest import make_model, GOLDEN
from oracle import segnet_oracle as S
from sivo_b200 import BayesianSegNet, BayesianSegNetParams
from sivo_b200.synth import stereo_frame

pytestmark = pytest.mark.gpu

ENGINES = ["simt", "auto"]


def _crop(kitti_bgr, h, w):
    return np.ascontiguousarray(kitti_bgr[100:100 + h, 300:300 + w])


@pytest.mark.parametrize("kind,kw", [("basic", dict(T=3, H=64, W=128)),
                                     ("standard", dict(T=2, H=64, W=128, widths=(64, 64, 64, 64, 64)))])
@pytest.mark.parametrize("prec", ["fp32", "fp16"])
@pytest.mark.parametrize("engine", ENGINES)
def test_blobs_match_oracle(model_dir, kitti_bgr, kind, kw, prec, engine):
    if prec == "fp32" and engine != "simt":
        pytest.skip("fp32 operands run on the SIMT engine only")
    net, w, proto, model = make_model(model_dir, kind, seed=0, **kw)
    img = _crop(kitti_bgr, kw["H"], kw["W"])
    seg = BayesianSegNet(BayesianSegNetParams(proto, model), seed=1234, precision=prec, engine=engine, keep_blobs=True)
    seg.set_frame(0)
    cls, conf, ent = seg.segmentImage(img)
    prob, blobs = S.forward(net, w, img, seed=1234, frame=0, precision=prec, return_blobs=True)
    # every blob up to the logits: identical inputs per layer only up to accumulated rounding, so the
    # tolerance is relative to the blob's scale; pooling masks must agree except at near-ties
    tol = 2e-4 if prec == "fp32" else 4e-3
    for ly in net.layers:
        if ly.type == "Softmax":
            continue
        for top in ly.tops:
            ref = blobs[top].numpy()
            got = seg.blob(top)
            if ref.shape[0] == 1 and got.shape[0] > 1:
                ref = np.repeat(ref, got.shape[0], axis=0)
            assert got.shape == ref.shape, top
            if top.endswith("_mask"):
                assert (got != ref).mean() < 2e-3, top
            else:
                scale = max(1.0, float(np.abs(ref).max()))
                bad = np.abs(got - ref) > tol * scale
                assert bad.mean() < 2e-3, (top, float(np.abs(got - ref).max()), scale)
    rc, rf, re = S.mc_reduce(prob)
    assert (cls != rc).mean() < 5e-3
    ok = cls == rc
    assert np.abs(conf - rf)[ok].max() < (1e-4 if prec == "fp32" else 5e-3)
    assert np.median(np.abs(ent - re)) < 1e-4


@pytest.mark.parametrize("kind", ["basic", "standard"])
@pytest.mark.parametrize("engine", ENGINES)
def test_matches_golden(model_dir, kitti_bgr, kind, engine):
    g = np.load(os.path.join(GOLDEN, "segnet_small.npz"))
    kw = dict(T=3, H=64, W=128) if kind == "basic" else dict(T=2, H=64, W=128, widths=(64, 64, 64, 64, 64))
    _, _, proto, model = make_model(model_dir, kind, seed=0, **kw)
    img = _crop(kitti_bgr, 64, 128)
    seg = BayesianSegNet(BayesianSegNetParams(proto, model), seed=1234, precision="fp16", engine=engine)
    seg.set_frame(0)
    cls, conf, ent = seg.segmentImage(img)
    assert (cls != g[f"{kind}_fp16_classes"]).mean() < 5e-3
    assert np.median(np.abs(ent - g[f"{kind}_fp16_entropy"])) < 1e-4
    assert np.median(np.abs(conf - g[f"{kind}_fp16_confidence"])) < 1e-4
    # and the fp16-operand model stays close to the fp32 reference semantics
    assert (cls != g[f"{kind}_fp32_classes"]).mean() < 0.05


def test_output_sizes_and_crop_like_the_reference(model_dir, kitti_bgr):
    # tests/test_bayesian_segnet.cpp:152-168 (sizes == H*W) + resizeImage's centre crop of the 1242x375 frame
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from sivo_b200.caffemodel import write_synth_model
    from sivo_b200.prototxt import load_net
    proto = os.path.join(root, "configs", "bayesian_segnet_basic.prototxt")
    net = load_net(open(proto).read())
    model = os.path.join(str(model_dir), "basic_full.caffemodel")
    write_synth_model(net, model, 0)
    seg = BayesianSegNet(BayesianSegNetParams(proto, model), seed=1234, T=2)
    assert seg.getInputGeometry() == (1024, 352)
    seg.set_frame(7)
    cls, conf, ent = seg.segmentImage(kitti_bgr)
    assert cls.size == conf.size == ent.size == 352 * 1024
    seg.set_frame(7)
    cls2, conf2, ent2 = seg.segmentImage(np.ascontiguousarray(kitti_bgr[11:11 + 352, 109:109 + 1024]))
    assert np.array_equal(cls, cls2) and np.array_equal(conf, conf2) and np.array_equal(ent, ent2)  # deterministic + crop origin (109, 11)
    # properties that hold at any size: probabilities, entropy bounds, argmax consistency
    assert conf.min() >= 1.0 / 15 - 1e-9 and conf.max() <= 1.0 + 1e-9
    assert ent.min() >= 0 and ent.max() <= np.log2(15) + 1e-9
    assert cls.max() < 15
    # a different frame index draws different masks
    cls3, _, ent3 = seg.segmentImage(kitti_bgr)
    assert not np.array_equal(ent, ent3)
    with pytest.raises(Exception):
        seg.segmentImage(np.zeros((100, 100, 3), np.uint8))


def test_full_size_basic_against_oracle(model_dir):
    """Config C1 geometry (Basic, T=2, 1024x352) on a synthetic frame: the whole operator vs the oracle."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from sivo_b200.caffemodel import write_synth_model
    from sivo_b200.prototxt import load_net
    proto = os.path.join(root, "configs", "bayesian_segnet_basic.prototxt")
    net = load_net(open(proto).read(), T=2)
    model = os.path.join(str(model_dir), "basic_full.caffemodel")
    w = write_synth_model(net, model, 0)
    left, _ = stereo_frame(0)
    seg = BayesianSegNet(BayesianSegNetParams(proto, model), seed=1234, T=2, precision="fp16")
    seg.set_frame(3)
    cls, conf, ent = seg.segmentImage(left)
    rc, rf, re = S.segment_image(net, w, left, seed=1234, frame=3, precision="fp16", T=2)
    mism = (cls != rc).mean()
    assert mism < 5e-3, mism
    assert np.median(np.abs(ent - re)) < 1e-4
    assert np.quantile(np.abs(ent - re), 0.99) < 5e-2
