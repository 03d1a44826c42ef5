"""detection::AggregatedFeaturesDetector on the GPU (csrc/aggdet.cu, fdb_aggdet_*) against the oracle's restatement
(oracle/fdoracle.py:aggregated_features_detect - FHOG and the IoU suppression pinned to the reference's own classes compiled
into oracle/_ref, the glue restated from AggregatedFeaturesDetector.cpp / AggregatedFeaturesExtractor.cpp). FHOG feature maps:
bit-exact. Score maps: the kernel adds in the oracle's order (channel, kernel row, kernel column) -> identical floats here;
against the real cv::filter2D the bound is 1e-4 (its summation order is OpenCV's)."""
import numpy as np
import pytest

from featuredetection_b200 import capi, synthetic as syn
from featuredetection_b200.detector import AggregatedFeaturesDetector


def _fo():
    from oracle import fdoracle as fo
    fo.build()
    return fo


def _weights(kh, kw, D, seed):
    rng = np.random.default_rng(seed)
    return rng.normal(0, 0.1, (kh, kw, D)).astype(np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("cell,kh,kw,olc,ib,ic,size", [(4, 5, 6, 3, False, True, (200, 160)), (8, 3, 3, 2, True, True, (320, 240)),
                                                       (4, 6, 6, 5, False, False, (253, 187)), (6, 4, 5, 4, True, False, (300, 220))])
def test_feature_and_score_maps_equal_the_oracle(ctx, cell, kh, kw, olc, ib, ic, size):
    fo = _fo()
    W, H = size
    frame = np.ascontiguousarray(syn.synthetic_frame(6)[:H, :W])
    w = _weights(kh, kw, 31, 4)
    det = AggregatedFeaturesDetector(ctx, w, bias=0.25, threshold=1e9, cell=cell, octave_layer_count=olc, interpolate_bins=ib,
                                     interpolate_cells=ic)
    det.prepare(W, H, 1)
    _, _, maps = fo.aggregated_features_detect(frame, w, bias=0.25, threshold=1e9, cell=cell, octave_layer_count=olc,
                                               interpolate_bins=ib, interpolate_cells=ic, want_scores=True)
    got = det.score_maps(frame)
    assert len(got) == len(maps) and len(got) >= 3
    import math
    inc = math.pow(0.5, 1.0 / olc)
    # the layers the oracle used (its pyramid restatement is pinned to cv2): recompute the scale limits the same way
    pw, ph = kw * cell, kh * cell
    max_width = int(H / (ph / pw)) if ph / pw > H / W else W
    min_scale = math.pow(inc, int(math.log(pw / max_width) / math.log(inc)))
    _, layers = fo.pyramid(frame, inc, min_scale, 1.0)
    checked = 0
    for (info, feat, scores), want, (_, _, img) in zip(got, maps, layers):
        assert (info["width"], info["height"]) == (img.shape[1], img.shape[0])
        ref_feat = fo.fhog(img, cell, 9, ib, ic, 0.2)
        assert feat.shape == ref_feat.shape and np.array_equal(feat, ref_feat)          # FHOG: bit for bit
        assert scores.size == want.size
        if want.size:
            assert np.array_equal(scores.reshape(want.shape), want)                      # same summation order: same floats
            checked += 1
    assert checked >= 3


@pytest.mark.gpu
@pytest.mark.parametrize("nms_type", [0, 2])
def test_detections_equal_the_oracle(ctx, nms_type):
    """whole chain on a batch: windows above the threshold -> boxes in image pixels -> NonMaximumSuppression, 3 frames"""
    fo = _fo()
    W, H = 320, 240
    w = _weights(5, 5, 31, 7)
    frames = np.ascontiguousarray(syn.synthetic_frames(20, 3)[:, :H, :W])
    # a threshold that lets a few hundred windows through
    _, _, maps = fo.aggregated_features_detect(frames[0], w, bias=0.1, threshold=1e9, cell=4, octave_layer_count=4, want_scores=True)
    allscores = np.concatenate([m.ravel() for m in maps if m.size])
    thr = float(np.quantile(allscores, 0.98))
    det = AggregatedFeaturesDetector(ctx, w, bias=0.1, threshold=thr, cell=4, octave_layer_count=4, nms_threshold=0.3, nms_type=nms_type,
                                     width_scale=0.9, height_scale=1.1)
    det.prepare(W, H, 3)
    rects, scores, fr = det.detect(frames)
    assert len(rects) > 0
    for k in range(3):
        r_ref, s_ref = fo.aggregated_features_detect(frames[k], w, bias=0.1, threshold=thr, cell=4, octave_layer_count=4, nms_threshold=0.3,
                                                     nms_type=nms_type, width_scale=0.9, height_scale=1.1)
        mine = fr == k
        assert np.array_equal(rects[mine], r_ref) and np.array_equal(scores[mine], s_ref)


@pytest.mark.gpu
def test_min_window_width_and_empty_cases(ctx):
    fo = _fo()
    w = _weights(5, 5, 31, 9)
    frame = np.ascontiguousarray(syn.synthetic_frame(3)[:200, :260])
    det = AggregatedFeaturesDetector(ctx, w, bias=0.0, threshold=0.05, cell=4, octave_layer_count=3, min_window_width=40)
    det.prepare(260, 200, 1)
    rects, scores, _ = det.detect(frame)
    r_ref, s_ref = fo.aggregated_features_detect(frame, w, bias=0.0, threshold=0.05, cell=4, octave_layer_count=3, min_window_width=40)
    assert np.array_equal(rects, r_ref) and np.array_equal(scores, s_ref)
    assert (rects[:, 2] >= 40).all() if len(rects) else True
    # an image smaller than the window: no layer, no detection
    tiny = AggregatedFeaturesDetector(ctx, w, bias=0.0, threshold=-1e9, cell=4, octave_layer_count=3)
    tiny.prepare(16, 16, 1)
    rects, scores, _ = tiny.detect(np.zeros((16, 16), np.uint8))
    assert len(rects) == 0
    with pytest.raises(capi.FdbError):
        AggregatedFeaturesDetector(ctx, w, bias=0.0, threshold=0.0, cell=0)
