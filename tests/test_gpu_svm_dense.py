"""GPU parity of the tensor-core SVM (csrc/svm_dense.cu: tcgen05.mma kind::i8 + float64 epilogue) against the oracle
(SvmClassifier.cpp:55-60, RbfKernel.hpp:32-40,78-108, HistEq64Filter.cpp:32-125) and against the per-window kernel
of svm.cu. Distances are float64 sums of up to 1024 kernel values: the integer part (dot products, squared norms) is
exact, the kernel value differs from glibc's exp by a few 1e-16 relative -> tolerance 1e-9 (north_star allows 1e-4)."""
import os

import numpy as np
import pytest

from featuredetection_b200 import synthetic as syn
from featuredetection_b200.detector import SlidingWindowCascade, ProbabilisticSvmClassifier

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _oracle():
    from oracle import fdoracle
    return fdoracle


class _per_window_kernel:
    """FDB_SVM_DENSE=0 routes the same calls through svm_kernel (one CTA per vector)"""

    def __enter__(self):
        os.environ["FDB_SVM_DENSE"] = "0"

    def __exit__(self, *a):
        os.environ.pop("FDB_SVM_DENSE", None)


@pytest.mark.parametrize("n,num_sv,dim", [(128, 64, 400), (300, 128, 400), (1000, 1024, 400), (257, 200, 380), (513, 300, 64),
                                          (4096, 1024, 400)])
def test_dense_vectors_against_oracle(ctx, n, num_sv, dim):
    fo = _oracle()
    rng = np.random.default_rng(n + num_sv)
    sv = rng.integers(0, 256, (num_sv, dim), dtype=np.uint8)
    sv[0] = 0; sv[-1] = 255
    model = syn.SvmModel(sv, rng.normal(0, 1, num_sv).astype(np.float32), gamma=7.689e-7,
                         bias=0.25, threshold=0.1)
    x = rng.integers(0, 256, (n, dim), dtype=np.uint8)
    x[0] = 0; x[1] = 255; x[2] = sv[3]
    # near vectors make kernel values close to 1 (small ssd) as well as the usual ~exp(-10)
    x[3:40] = np.clip(sv[rng.integers(0, num_sv, 37)].astype(np.int32) + rng.integers(-3, 4, (37, dim)), 0, 255).astype(np.uint8)
    cls = ProbabilisticSvmClassifier(ctx, model)
    assert cls.has_dense, "the tensor-core form must exist for this model"
    dist, prob, pos = cls.get_probability(x)
    m = min(n, 400)  # the scalar oracle needs ~0.2 ms per (vector, 1024 support vectors)
    idx = np.concatenate([np.arange(min(m, 48)), rng.choice(n, m - min(m, 48), replace=False)]) if n > m else np.arange(n)
    rd, rp, rpos = fo.Svm(model).eval(x[idx])
    assert np.max(np.abs(dist[idx] - rd)) <= TOL, np.max(np.abs(dist[idx] - rd))
    assert np.max(np.abs(prob[idx] - rp)) <= TOL
    assert np.array_equal(pos[idx], rpos)
    with _per_window_kernel():
        assert not cls.has_dense
        d2, p2, q2 = cls.get_probability(x)
    assert np.max(np.abs(dist - d2)) <= TOL, np.max(np.abs(dist - d2))
    assert np.array_equal(pos, q2)


def test_dense_unavailable_models_keep_the_per_window_kernel(ctx):
    """large gamma (exp table finer than shared memory holds) and float32 support vectors have no tensor-core form"""
    fo = _oracle()
    rng = np.random.default_rng(3)
    sv = rng.integers(0, 256, (64, 400), dtype=np.uint8)
    model = syn.SvmModel(sv, rng.normal(0, 1, 64).astype(np.float32), gamma=0.01, bias=0.0, threshold=0.0)
    cls = ProbabilisticSvmClassifier(ctx, model)
    assert not cls.has_dense
    x = np.clip(sv[rng.integers(0, 64, 200)].astype(np.int32) + rng.integers(-2, 3, (200, 400)), 0, 255).astype(np.uint8)
    dist, _, _ = cls.get_probability(x)
    rd, _, _ = fo.Svm(model).eval(x)
    assert np.max(np.abs(dist - rd)) <= TOL


def test_single_detector_dense_windows(ctx, face_models):
    """`single` psvm detector (ffpDetectApp.cpp:427-500): every window of every layer through ONE kernel launch per
    chunk; all distances + positives against the oracle and against the per-window kernel; ragged last pass; batch of 5
    frames in chunks"""
    fo = _oracle()
    det_kw, _, _ = face_models
    kw = dict(det_kw, min_scale_factor=0.09, max_scale_factor=0.2)
    frames = np.ascontiguousarray(syn.synthetic_frames(40, 5)[:, :240, :320])
    svm = syn.make_svm(20, 20, seed=5, num_sv=300)
    casc = SlidingWindowCascade(ctx, kw, None, svm)
    casc.prepare(320, 240, 5)
    assert casc.single_dense
    dets, dist = casc.detect_single(frames)
    so = fo.Svm(svm)
    for k in (0, 4):
        ref = fo.detect_frame(kw, None, so, frames[k], frame_index=k)
        assert dist.shape[1] == ref["windows"]
        assert np.max(np.abs(dist[k] - ref["svm_dense"])) <= TOL, np.abs(dist[k] - ref["svm_dense"]).max()
        mine = dets[dets["frame"] == k]
        assert list(mine["window"]) == list(ref["detections"]["window"])
        assert np.array_equal(mine["center_x"], ref["detections"]["center_x"])
        assert np.allclose(mine["probability"], ref["detections"]["probability"], rtol=0, atol=TOL)
    assert len(dets) > 0 and casc.last_counts()[0] == 5 * casc.windows_per_frame
    with _per_window_kernel():
        d2, dist2 = casc.detect_single(frames)
    assert np.max(np.abs(dist - dist2)) <= TOL
    assert list(d2["window"]) == list(dets["window"]) and list(d2["frame"]) == list(dets["frame"])
    # single frame, partial batch, no distances wanted
    d3, none = casc.detect_single(frames[1], want_distances=False)
    assert none is None and list(d3["window"]) == list(dets[dets["frame"] == 1]["window"])


def test_single_detector_dense_full_size(ctx, face_models):
    """640x480 FaceFrontal geometry (16 185 windows per frame, 13 layers incl. window steps at layer borders), 1024
    support vectors: sampled windows against the oracle, everything against the per-window kernel"""
    fo = _oracle()
    det_kw, _, svm = face_models
    frames = syn.synthetic_frames(3, 3)
    casc = SlidingWindowCascade(ctx, det_kw, None, svm)
    casc.prepare(640, 480, 3)
    assert casc.single_dense
    dets, dist = casc.detect_single(frames)
    with _per_window_kernel():
        d2, dist2 = casc.detect_single(frames)
    assert np.max(np.abs(dist - dist2)) <= TOL, np.max(np.abs(dist - dist2))
    assert list(d2["window"]) == list(dets["window"])
    # oracle on a sample of windows of frame 2 (the scalar SVM costs ~0.2 ms per window)
    rng = np.random.default_rng(0)
    layers = casc.layers()
    _, pyr = fo.pyramid(frames[2], det_kw["incremental_scale_factor"], det_kw["min_scale_factor"], det_kw["max_scale_factor"])
    so = fo.Svm(svm)
    patches, wins = [], []
    for li, L in enumerate(layers):
        img = pyr[li][2]
        for _ in range(12):
            ix, iy = int(rng.integers(0, L["windows_x"])), int(rng.integers(0, L["windows_y"]))
            patches.append(fo.hq64(img[iy:iy + 20, ix:ix + 20]).ravel())
            wins.append(L["first_window"] + iy * L["windows_x"] + ix)
    rd, _, _ = so.eval(np.stack(patches))
    assert np.max(np.abs(dist[2][wins] - rd)) <= TOL


def test_single_detector_dense_other_patch_size(ctx):
    """16x24 patches (the ear detectors): the generic producer path of svm_dense_kernel against the per-window kernel and,
    on a sample, the oracle"""
    fo = _oracle()
    det_kw, _, _ = syn.landmark_models("LeftEarCenter")
    kw = dict(det_kw, min_scale_factor=0.3, max_scale_factor=0.5)
    svm = syn.make_svm(16, 24, seed=8, num_sv=200)
    frames = np.ascontiguousarray(syn.synthetic_frames(60, 2)[:, :240, :320])
    casc = SlidingWindowCascade(ctx, kw, None, svm)
    casc.prepare(320, 240, 2)
    assert casc.single_dense
    dets, dist = casc.detect_single(frames)
    with _per_window_kernel():
        d2, dist2 = casc.detect_single(frames)
    assert np.max(np.abs(dist - dist2)) <= TOL, np.max(np.abs(dist - dist2))
    assert list(d2["window"]) == list(dets["window"])
    ref = fo.detect_frame(kw, None, fo.Svm(svm), frames[1], frame_index=1)
    assert np.max(np.abs(dist[1] - ref["svm_dense"])) <= TOL
