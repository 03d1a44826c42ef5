"""detection::NonMaximumSuppression (NonMaximumSuppression.cpp:27-112, SURVEY 8(f) rank 2's suppression step): the host
function of the library against the oracle's C restatement and the reference's own class compiled into oracle/_ref."""
import ctypes as C

import numpy as np
import pytest

from featuredetection_b200 import capi


def _boxes(rng, n, distinct_scores=True):
    centres = rng.integers(40, 600, (max(n // 6, 1), 2))
    c = centres[rng.integers(0, len(centres), n)] + rng.integers(-25, 26, (n, 2))
    wh = rng.integers(30, 90, (n, 1)) + rng.integers(-5, 6, (n, 2))
    rects = np.concatenate([c - wh // 2, wh], axis=1).astype(np.int32)
    scores = rng.uniform(0.1, 5.0, n).astype(np.float32)
    if distinct_scores:
        scores = (np.argsort(np.argsort(scores)) * 0.01 + 0.1).astype(np.float32)
    else:
        scores = np.round(scores * 2) / 2            # many ties
    return scores.astype(np.float32), rects


def _run(fn, scores, rects, thr, kind, via_n_out):
    s, r = scores.copy(), np.ascontiguousarray(rects.copy())
    if via_n_out:
        n = C.c_int64()
        assert fn(s.ctypes.data, r.ctypes.data, len(s), thr, kind, C.byref(n)) == 0
        k = n.value
    else:
        k = fn(s.ctypes.data, r.ctypes.data, len(s), thr, kind)
    return s[:k], r[:k]


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("thr", [0.3, 0.5, 0.0, 1.0])
def test_iou_nms_matches_oracle_and_reference(built, kind, thr):
    from oracle import fdoracle as fo
    lib = capi.load_library()
    L = fo.lib()
    L.fdo_non_maximum_suppression.restype = C.c_int64
    L.fdo_non_maximum_suppression.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int]
    rng = np.random.default_rng(int(thr * 10) + kind)
    for n in (0, 1, 2, 57, 400):
        scores, rects = _boxes(rng, n)
        ps, pr = _run(lib.fdb_non_maximum_suppression, scores, rects, thr, kind, True)
        os_, or_ = _run(L.fdo_non_maximum_suppression, scores, rects, thr, kind, False)
        assert np.array_equal(ps, os_) and np.array_equal(pr, or_), (n, kind, thr)
        if fo.ref_available():
            R = fo.ref()
            R.ref_non_maximum_suppression.restype = C.c_int64
            R.ref_non_maximum_suppression.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int]
            rs, rr = _run(R.ref_non_maximum_suppression, scores, rects, thr, kind, False)
            assert np.array_equal(ps, rs) and np.array_equal(pr, rr), (n, kind, thr)
        if n and thr < 1.0:
            assert len(ps) <= n and np.all(np.diff(ps) <= 0)      # clusters come out best first
        if thr == 1.0:
            assert np.array_equal(ps, scores)                     # NonMaximumSuppression.cpp:28-29: returned unchanged


def test_iou_nms_ties_and_errors(built):
    from oracle import fdoracle as fo
    lib = capi.load_library()
    L = fo.lib()
    L.fdo_non_maximum_suppression.restype = C.c_int64
    L.fdo_non_maximum_suppression.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int]
    rng = np.random.default_rng(9)
    scores, rects = _boxes(rng, 300, distinct_scores=False)   # ties: product and oracle both keep the input order (std::sort leaves it open)
    for kind in (0, 1, 2):
        a = _run(lib.fdb_non_maximum_suppression, scores, rects, 0.4, kind, True)
        b = _run(L.fdo_non_maximum_suppression, scores, rects, 0.4, kind, False)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    n = C.c_int64()
    assert lib.fdb_non_maximum_suppression(scores.ctypes.data, rects.ctypes.data, 300, 0.4, 7, C.byref(n)) != 0
    assert lib.fdb_non_maximum_suppression(None, None, 3, 0.4, 0, C.byref(n)) != 0
