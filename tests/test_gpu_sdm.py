"""GPU parity of the supervised-descent regressor (sdm.cu through the C ABI; BASELINE configs[4]) against the CPU oracle
(oracle/fd_sdm.c, pinned against the reference's own hog.c and cv2) and the committed golden fits (tests/golden/sdm.npz).

Bar: descriptors bit-exact (every float32 / float64 operation of hog.c in its order); fitted shapes within 1e-4 px
(north_star) - the regressor product accumulates float32 x float32 products in float64 like cv::gemm, in a tiled order,
so a shape coordinate can differ from the oracle by one float32 ulp; faces whose window leaves the image are flagged
in `status` where the reference would throw."""
import numpy as np
import pytest

from featuredetection_b200 import synthetic as syn
from featuredetection_b200.detector import SdmLandmarkModel

pytestmark = pytest.mark.gpu
TOL = 1e-4
BOXES = np.array([[220, 140, 200, 200], [100, 60, 260, 260], [300, 200, 150, 150], [5, 5, 200, 200], [400, 250, 230, 220]], np.int32)


def _oracle():
    from oracle import fdoracle as fo
    return fo


@pytest.fixture(scope="module")
def models(built):
    return {L: syn.make_sdm(L, 5, 500) for L in (68, 15)}


def test_descriptors_match_oracle(ctx, models):
    fo = _oracle()
    sdm = SdmLandmarkModel(ctx, models[68])
    frame = syn.synthetic_frame(2)
    rng = np.random.default_rng(5)
    for wsh in (6, 9, 12, 15, 18, 21, 30, 33, 45, 60):
        pts = np.stack([rng.uniform(wsh + 1, 638 - wsh, 100), rng.uniform(wsh + 1, 478 - wsh, 100)], axis=1).astype(np.float32)
        assert np.array_equal(sdm.descriptors(frame, pts, wsh), fo.sdm_descriptors(frame, pts, wsh)), wsh
    # border windows: black canvas and the reference's row-offset quirk
    for pt in ([4.0, 200.0], [3.0, 100.5], [10.0, 240.0]):
        assert np.array_equal(sdm.descriptors(frame, [pt], 15), fo.sdm_descriptors(frame, [pt], 15)), pt
    from featuredetection_b200.capi import FdbError
    with pytest.raises(FdbError):
        sdm.descriptors(frame, [[636.0, 2.0]], 15)
    # flat and degenerate inputs
    flat = np.full((480, 640), 93, np.uint8)
    assert np.array_equal(sdm.descriptors(flat, [[320, 240]], 15), fo.sdm_descriptors(flat, [[320, 240]], 15))


@pytest.mark.parametrize("L", [68, 15])
def test_fit_matches_oracle(ctx, models, L):
    fo = _oracle()
    sdm = SdmLandmarkModel(ctx, models[L])
    ora = fo.Sdm(models[L])
    frames = syn.synthetic_frames(0, 4)
    boxes = np.tile(BOXES, (4, 1))
    face_frame = np.repeat(np.arange(4, dtype=np.int32), len(BOXES))
    start = sdm.align_rigid(boxes)
    for i, b in enumerate(boxes):
        assert np.array_equal(start[i], ora.align_rigid(b))
    shapes, status, feats = sdm.optimize(frames, start, face_frame, want_features=True)
    n_bad = 0
    for i in range(len(boxes)):
        try:
            want, wfeat = ora.optimize(frames[face_frame[i]], start[i], want_features=True)
        except RuntimeError:
            assert status[i] != 0, i
            n_bad += 1
            continue
        assert status[i] == 0, i
        assert np.array_equal(feats[0, i], wfeat[0]), i          # first step: same start shape -> identical descriptors
        assert np.abs(shapes[i] - want).max() <= TOL, (i, np.abs(shapes[i] - want).max())
    print("faces %d, out of image %d, max |shape - oracle| = %.3g px" % (
        len(boxes), n_bad, max(np.abs(shapes[i] - ora.optimize(frames[face_frame[i]], start[i])).max() for i in range(len(boxes)) if status[i] == 0)))


def test_fit_matches_golden(ctx, models):
    """the committed oracle fits of the synthetic 68-landmark model (tests/golden/make_sdm_golden.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdm.npz"))
    sdm = SdmLandmarkModel(ctx, models[68])
    frames = syn.synthetic_frames(0, 4)
    face_frame = np.repeat(np.arange(4, dtype=np.int32), len(g["fit_boxes"]))
    assert np.array_equal(sdm.align_rigid(np.tile(g["fit_boxes"], (4, 1))), g["fit_synth_start"])
    shapes, status = sdm.optimize(frames, g["fit_synth_start"], face_frame)
    want = g["fit_synth_shapes"]
    failed = np.isnan(want[:, 0])
    assert np.array_equal(status != 0, failed)
    assert np.abs(shapes[~failed] - want[~failed]).max() <= TOL


def test_face_per_frame_default_and_empty(ctx, models):
    fo = _oracle()
    sdm = SdmLandmarkModel(ctx, models[15])
    ora = fo.Sdm(models[15])
    frames = syn.synthetic_frames(3, 3)
    start = sdm.align_rigid(np.tile(BOXES[0], (3, 1)))
    shapes, status = sdm.optimize(frames, start)          # face i lies in frame i
    for i in range(3):
        assert status[i] == 0 and np.abs(shapes[i] - ora.optimize(frames[i], start[i])).max() <= TOL
    s, st = sdm.optimize(frames, np.zeros((0, 30), np.float32))
    assert s.shape == (0, 30) and st.shape == (0,)


def test_batch_is_order_independent(ctx, models):
    """size-independent property at a larger batch: a face's fit does not depend on its position in the batch"""
    sdm = SdmLandmarkModel(ctx, models[68])
    frames = syn.synthetic_frames(0, 8)
    rng = np.random.default_rng(9)
    n = 512
    boxes = np.stack([rng.integers(60, 300, n), rng.integers(40, 180, n), rng.integers(120, 260, n), rng.integers(120, 260, n)], axis=1).astype(np.int32)
    boxes[:, 2] = np.minimum(boxes[:, 2], 600 - boxes[:, 0]); boxes[:, 3] = np.minimum(boxes[:, 3], 450 - boxes[:, 1])
    ff = rng.integers(0, 8, n).astype(np.int32)
    start = sdm.align_rigid(boxes)
    a, sa = sdm.optimize(frames, start, ff)
    perm = rng.permutation(n)
    b, sb = sdm.optimize(frames, start[perm], ff[perm])
    assert np.array_equal(sa[perm], sb) and np.array_equal(a[perm], b)


def test_pipelined_upload_matches_plain_call(ctx, models):
    """>= 64 MiB of frames: the host call uploads the frames in chunks, each followed by the fit of its faces on its own
    stream (sdm.cu, SDM_CHUNKS); the result must be what the plain call gives for the same faces"""
    sdm = SdmLandmarkModel(ctx, models[68])
    base = syn.synthetic_frames(0, 8)
    frames = np.concatenate([base] * 28)            # 224 frames, 68.8 MB
    rng = np.random.default_rng(11)
    n = 700
    boxes = np.stack([rng.integers(60, 300, n), rng.integers(40, 180, n), rng.integers(120, 260, n), rng.integers(120, 260, n)], axis=1).astype(np.int32)
    boxes[:, 2] = np.minimum(boxes[:, 2], 600 - boxes[:, 0]); boxes[:, 3] = np.minimum(boxes[:, 3], 450 - boxes[:, 1])
    boxes[::97] = [5, 5, 200, 200]                  # some faces leave the image: status must travel through the permutation
    ff = rng.integers(0, 224, n).astype(np.int32)
    start = sdm.align_rigid(boxes)
    a, sa = sdm.optimize(frames, start, ff)         # pipelined
    b, sb = sdm.optimize(base, start, ff % 8)       # plain
    assert np.array_equal(sa, sb) and np.array_equal(a, b)
    assert (sa != 0).any() and (sa == 0).any()
