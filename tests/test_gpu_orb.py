"""-m gpu: `ORBextractor::operator()` and the stereo Hamming stage through the C-ABI, bit-exact against
the oracle and the committed golden fixtures."""
import os
import threading

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import orb_oracle as O
from sivo_b200 import ORBextractor, stereo_hamming, KP_DTYPE
from sivo_b200.synth import bgr_to_gray, stereo_frame

pytestmark = pytest.mark.gpu


def kp_matrix(kps):
    return np.stack([kps["x"], kps["y"], kps["size"], kps["angle"], kps["response"], kps["octave"].astype(np.float32),
                     kps["class_id"].astype(np.float32)], axis=1)


@pytest.mark.parametrize("name,nf", [("orb_kitti_2000", 2000), ("orb_kitti_1000", 1000)])
def test_kitti_fixture_bit_exact_vs_golden(kitti_gray_crop, name, nf):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    ext = ORBextractor(nf, 1.2, 8, 20, 7)
    kps, desc = ext(kitti_gray_crop, None)
    for lvl in range(8):
        assert len(ext.candidates(lvl)[0]) == g["cand_counts"][lvl]
    assert len(kps) == len(g["keypoints"])
    assert np.array_equal(kp_matrix(kps), g["keypoints"])
    assert np.array_equal(desc, g["descriptors"])


def test_pyramid_and_candidates_vs_oracle(kitti_gray_crop):
    ext = ORBextractor(2000, 1.2, 8, 20, 7)
    ext(kitti_gray_crop, None)
    r = O.extract(kitti_gray_crop)
    for lvl in range(8):
        assert np.array_equal(ext._bordered[lvl], r.pyramid[lvl]), lvl
        xs, ys, rs = ext.candidates(lvl)
        assert np.array_equal(xs, r.candidates[lvl][0]) and np.array_equal(ys, r.candidates[lvl][1])
        assert np.array_equal(rs, r.candidates[lvl][2])
    assert ext.mvImagePyramid[0].shape == (352, 1024) and np.array_equal(ext.mvImagePyramid[0], kitti_gray_crop)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_synthetic_frames_bit_exact_vs_oracle(seed):
    left, right = stereo_frame(seed)
    for img in (left, right):
        gray = np.ascontiguousarray(bgr_to_gray(img)[11:11 + 352, 109:109 + 1024])
        ext = ORBextractor(2000, 1.2, 8, 20, 7)
        kps, desc = ext(gray, None)
        r = O.extract(gray)
        assert np.array_equal(kp_matrix(kps), r.keypoints.astype(np.float32))
        assert np.array_equal(desc, r.descriptors)


def test_edge_cases():
    ext = ORBextractor(500, 1.2, 8, 20, 7)
    k, d = ext(np.zeros((0, 0), np.uint8), None)           # empty image returns silently (:1023-1024)
    assert len(k) == 0 and d.shape == (0, 32)
    flat = np.full((352, 1024), 128, np.uint8)              # no corners -> zero keypoints, descriptors released
    k, d = ext(flat, None)
    assert len(k) == 0 and d.shape == (0, 32)
    rng = np.random.default_rng(0)                          # odd, non-multiple-of-anything size
    img = rng.integers(0, 256, size=(301, 517), dtype=np.uint8)
    k, d = ext(img, None)
    r = O.extract(img, O.ExtractorParams(500))
    assert np.array_equal(kp_matrix(k), r.keypoints.astype(np.float32)) and np.array_equal(d, r.descriptors)
    with pytest.raises(Exception):
        ext(np.zeros((40, 40), np.uint8), None)             # too small for 8 levels


def test_two_extractors_concurrently(kitti_gray_crop):
    # Frame.cc:126-129 runs left and right on two std::threads
    g = np.load(os.path.join(GOLDEN, "orb_kitti_2000.npz"))
    exts = [ORBextractor(2000, 1.2, 8, 20, 7) for _ in range(2)]
    out = [None, None]

    def work(i):
        for _ in range(5):
            out[i] = exts[i](kitti_gray_crop, None)
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for kps, desc in out:
        assert np.array_equal(kp_matrix(kps), g["keypoints"]) and np.array_equal(desc, g["descriptors"])


def test_stereo_hamming_vs_brute_force():
    left, right = stereo_frame(0)
    gl = np.ascontiguousarray(bgr_to_gray(left)[11:11 + 352, 109:109 + 1024])
    gr = np.ascontiguousarray(bgr_to_gray(right)[11:11 + 352, 109:109 + 1024])
    el, er = ORBextractor(2000, 1.2, 8, 20, 7), ORBextractor(2000, 1.2, 8, 20, 7)
    kl, dl = el(gl, None)
    kr, dr = er(gr, None)
    sf = el.GetScaleFactors()
    max_d = 100.0
    idx, dist = stereo_hamming(kl, dl, kr, dr, sf, 352, 0.0, max_d)
    # restatement of Frame.cc:452-533 (row table, octave gate, disparity window, first strictly smaller distance)
    rows = [[] for _ in range(352)]
    for i in range(len(kr)):
        r = np.float32(2.0) * sf[kr["octave"][i]]
        lo, hi = int(np.floor(kr["y"][i] - r)), int(np.ceil(kr["y"][i] + r))
        for y in range(lo, hi + 1):
            rows[y].append(i)
    xor = np.unpackbits(dl[:, None, :] ^ dr[None, :, :], axis=2).sum(axis=2) if len(kl) * len(kr) < 6e6 else None
    matched = 0
    for i in range(len(kl)):
        best, bi = 100, -1
        u, v, lv = kl["x"][i], kl["y"][i], kl["octave"][i]
        if u - np.float32(0) >= 0:
            for j in rows[int(v)]:
                if kr["octave"][j] < lv - 1 or kr["octave"][j] > lv + 1:
                    continue
                if np.float32(u - np.float32(max_d)) <= kr["x"][j] <= u:
                    d = int(xor[i, j]) if xor is not None else O.descriptor_distance(dl[i], dr[j])
                    if d < best:
                        best, bi = d, j
        assert dist[i] == best and idx[i] == bi, i
        matched += bi >= 0
    assert matched > 200


def test_stereo_match_full_vs_oracle():
    """The whole ComputeStereoMatches (Hamming + SAD slide + parabola + median cut) on the device pyramids."""
    from sivo_b200 import stereo_match
    left, right = stereo_frame(1)
    gl = np.ascontiguousarray(bgr_to_gray(left)[11:11 + 352, 109:109 + 1024])
    gr = np.ascontiguousarray(bgr_to_gray(right)[11:11 + 352, 109:109 + 1024])
    el, er = ORBextractor(2000, 1.2, 8, 20, 7), ORBextractor(2000, 1.2, 8, 20, 7)
    kl, dl = el(gl, None)
    kr, dr = er(gr, None)
    mbf = 386.1448  # Camera.bf of config/kitti/KITTI00-02.yaml; mb = bf / fx
    mb = mbf / 718.856
    u, z = stereo_match(el, er, kl, dl, kr, dr, mb, mbf)
    ru, rz = O.compute_stereo_matches(kl, dl, kr, dr, el._bordered, er._bordered, el.GetScaleFactors(), el.GetInverseScaleFactors(), mb, mbf)
    assert np.array_equal(u, ru) and np.array_equal(z, rz)
    assert (u >= 0).sum() > 150


def test_hamming_best2_equals_the_sequential_loop():
    """Next row 8(f)-4: best / second-best over candidate lists, ties and empty lists included (bit-exact: integers)."""
    from sivo_b200.orb import hamming_best2
    rng = np.random.default_rng(11)
    nt, nq = 700, 400
    train = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
    # queries near some train descriptors (a few bits flipped) so that small distances and exact ties occur
    src = rng.integers(0, nt, nq)
    query = train[src].copy()
    for i in range(nq):
        flips = rng.integers(0, 256, size=rng.integers(0, 40))
        for b in flips:
            query[i, b >> 3] ^= np.uint8(1 << (b & 7))
    train[5] = train[9]  # duplicated descriptors: equal distances inside one list
    train[17] = train[9]
    level = rng.integers(0, 8, nt).astype(np.int32)
    lens = rng.integers(0, 90, nq)
    lens[:5] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cand = np.concatenate([rng.permutation(nt)[:n] for n in lens] + [np.empty(0, np.int64)]).astype(np.int32)
    for i in range(5, 40):  # make sure the true source and the duplicates are among the candidates of some queries
        if lens[i] >= 3:
            cand[off[i]:off[i] + 3] = (17, 5, 9)
            query[i] = train[9]
    want = O.hamming_best2(query, train, off, cand, level)
    got = hamming_best2(query, train, off, cand, level)
    assert np.array_equal(got, want)
    assert np.array_equal(hamming_best2(query, train, off, cand, None)[:, [0, 1, 3]], want[:, [0, 1, 3]])
    assert (got[:5] == np.array([-1, 256, -1, 256, -1])).all()
