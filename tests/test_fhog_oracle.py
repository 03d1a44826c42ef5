"""Groundwork for SURVEY 8(f) rank 2: the FHOG layer filter of AggregatedFeaturesDetector's feature pyramid. The C
restatement (oracle/fd_fhog.c) against the reference's own FhogFilter / FhogAggregationFilter compiled into oracle/_ref -
bit for bit (float32, the reference's order of operations). No product kernel exists yet for this row."""
import ctypes as C

import numpy as np
import pytest

from featuredetection_b200 import synthetic as syn

ARGS = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]


def _fhog(fn, img, cell, bins, ib, ic, alpha):
    img = np.ascontiguousarray(img, np.uint8)
    rows, cols = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    D = 3 * bins + 4
    out = np.full((rows // cell, cols // cell, D), np.nan, np.float32)
    n = fn(img.ctypes.data, cols, rows, ch, cell, bins, int(ib), int(ic), alpha, out.ctypes.data)
    assert n == out.size
    return out


@pytest.mark.parametrize("cell,bins,ib,ic,alpha", [(4, 9, True, True, 0.2), (8, 9, False, True, 0.2), (4, 6, True, False, 0.2),
                                                   (5, 9, False, False, 0.5), (6, 8, True, True, 1.0)])
def test_fhog_restatement_equals_the_compiled_reference(built, cell, bins, ib, ic, alpha):
    from oracle import fdoracle as fo
    if not fo.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference)")
    L, R = fo.lib(), fo.ref()
    L.fdo_fhog.restype = C.c_int64; L.fdo_fhog.argtypes = ARGS
    R.ref_fhog.restype = C.c_int64; R.ref_fhog.argtypes = ARGS
    gray = syn.synthetic_frame(3)[:131, :203]                       # sizes that are not multiples of the cell
    rng = np.random.default_rng(1)
    bgr = np.stack([gray, np.roll(gray, 3, 1), rng.integers(0, 256, gray.shape, dtype=np.uint8)], axis=2)
    flat = np.full((40, 48), 77, np.uint8)                           # zero gradients: the eps of the normalisers decides
    for img in (gray, bgr, flat, gray[:cell, :cell * 2]):
        a = _fhog(L.fdo_fhog, img, cell, bins, ib, ic, alpha)
        b = _fhog(R.ref_fhog, img, cell, bins, ib, ic, alpha)
        assert a.shape == b.shape and np.array_equal(a, b), float(np.nanmax(np.abs(a - b)))
        assert np.isfinite(a).all() and a.min() >= 0 and a[..., :3 * bins].max() <= 2 * alpha + 1e-6


def test_fhog_properties(built):
    from oracle import fdoracle as fo
    L = fo.lib()
    L.fdo_fhog.restype = C.c_int64; L.fdo_fhog.argtypes = ARGS
    img = syn.synthetic_frame(1)[:64, :96]
    d = _fhog(L.fdo_fhog, img, 4, 9, True, True, 0.2)
    assert d.shape == (16, 24, 31)
    # contrast-insensitive bins are bounded by the clamp: 0.5 * 4 * alpha
    assert d[..., 18:27].max() <= 0.4 + 1e-6
    # inverting the image flips every gradient: signed bins rotate by half a turn, unsigned bins and energies stay
    e = _fhog(L.fdo_fhog, 255 - img, 4, 9, True, True, 0.2)
    assert np.allclose(e[..., 18:], d[..., 18:], atol=1e-5)
    assert np.allclose(e[..., :9], d[..., 9:18], atol=2e-4) and np.allclose(e[..., 9:18], d[..., :9], atol=2e-4)
    assert L.fdo_fhog(img.ctypes.data, 96, 64, 2, 4, 9, 1, 1, 0.2, d.ctypes.data) == -1


def test_aggregated_features_detector_restatement_finds_a_planted_template(built):
    """the whole rank-2 chain of the oracle (pyramid -> FHOG -> linear-SVM score map -> threshold -> bounds -> IoU NMS): a
    template cut out of a frame's own FHOG map must be found where it was cut, at the scale it was cut from"""
    from oracle import fdoracle as fo
    frame = syn.synthetic_frame(4)[:240, :320].copy()
    rng = np.random.default_rng(2)
    frame[60:140, 100:180] = rng.integers(0, 256, (80, 80), dtype=np.uint8)      # a textured 80x80 object
    cell, kh, kw = 4, 10, 10                                                       # 40x40 px window at scale 1
    inc = 0.5 ** (1.0 / 5)
    _, layers = fo.pyramid(frame, inc, 0.5, 1.0)
    idx5 = [l for l in layers if l[0] == 5][0]                                     # scale 0.5: the object is 40x40 px there
    feat = fo.fhog(idx5[2], cell)
    y0, x0 = 30 // cell, 50 // cell
    w = feat[y0:y0 + kh, x0:x0 + kw].copy()
    w -= w.mean()
    self_score = float((feat[y0:y0 + kh, x0:x0 + kw] * w).sum())
    rects, scores, maps = fo.aggregated_features_detect(frame, w, bias=0.0, threshold=0.8 * self_score, cell=cell, octave_layer_count=5,
                                                        nms_threshold=0.3, want_scores=True)
    assert len(rects) >= 1 and scores[0] >= 0.99 * self_score * 0.8
    x, y, bw, bh = rects[0]
    assert abs(x - 2 * x0 * cell) <= 8 and abs(y - 2 * y0 * cell) <= 8 and abs(bw - 80) <= 2 and abs(bh - 80) <= 2
    assert np.all(np.diff(scores) <= 0)
    # min_window_width removes the layers whose windows would be smaller
    r2, s2 = fo.aggregated_features_detect(frame, w, bias=0.0, threshold=0.8 * self_score, cell=cell, octave_layer_count=5, min_window_width=100)
    assert all(r[2] >= 80 for r in r2)


def test_host_chain_of_the_library_matches_the_restatement(built):
    """the host-only product pieces of rank 2 - fdb_aggdet_windows (positives -> boxes) and fdb_non_maximum_suppression -
    fed with the oracle's score maps reproduce aggregated_features_detect's detections exactly"""
    import ctypes as C
    from featuredetection_b200 import capi
    from oracle import fdoracle as fo
    lib = capi.load_library()
    frame = np.ascontiguousarray(syn.synthetic_frame(8)[:200, :280])
    rng = np.random.default_rng(6)
    w = rng.normal(0, 0.15, (6, 5, 31)).astype(np.float32)
    kw = dict(cell=4, octave_layer_count=4, nms_threshold=0.35, nms_type=2, width_scale=0.9, height_scale=1.1)
    _, _, maps = fo.aggregated_features_detect(frame, w, bias=0.1, threshold=1e9, want_scores=True, **kw)
    thr = float(np.quantile(np.concatenate([m.ravel() for m in maps if m.size]), 0.97))
    rects, scores, maps = fo.aggregated_features_detect(frame, w, bias=0.1, threshold=thr, want_scores=True, **kw)
    assert len(rects) > 3
    import math
    inc = 0.5 ** (1.0 / 4)
    patch_w, patch_h = 5 * 4, 6 * 4
    aspect, image_aspect = patch_h / patch_w, 200 / 280
    max_width = int(200 / aspect) if aspect > image_aspect else 280
    min_scale = math.pow(inc, int(math.log(patch_w / max_width) / math.log(inc)))
    _, layers = fo.pyramid(frame, inc, min_scale, 1.0)
    all_s, all_r = [], []
    for (_, _, img), m in zip(layers, maps):
        if m.size == 0:
            continue
        cap = m.size
        s = np.zeros(cap, np.float32); r = np.zeros((cap, 4), np.int32); n = C.c_int64()
        capi.check(lib, lib.fdb_aggdet_windows(np.ascontiguousarray(m).ctypes.data, m.shape[0], m.shape[1], thr, 6, 5, 4,
                                               img.shape[1] / 280.0, img.shape[0] / 200.0, 0.9, 1.1, s.ctypes.data, r.ctypes.data, cap, C.byref(n)))
        all_s.append(s[:n.value]); all_r.append(r[:n.value])
    s = np.concatenate(all_s); r = np.ascontiguousarray(np.concatenate(all_r))
    n = C.c_int64()
    capi.check(lib, lib.fdb_non_maximum_suppression(s.ctypes.data, r.ctypes.data, len(s), 0.35, 2, C.byref(n)))
    assert np.array_equal(r[:n.value], rects) and np.array_equal(s[:n.value], scores)
