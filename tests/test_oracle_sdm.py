"""The oracle's supervised-descent restatement (oracle/fd_sdm.c) against tests/golden/sdm.npz - the reference's own hog.c
(compiled into oracle/_ref) and cv2 4.13 for cv::resize / cv::gemm on CV_32F - and, where this environment has them,
against oracle/_ref and cv2 live; plus the product's host-side model reader (fdb_sdm_file_load, no GPU involved) against
the reference's in-repo model file.

Tolerance: HOG and resize are bit-exact; cv::gemm accumulates CV_32F products in double in an order that depends on its
blocking, so the product row is compared within 1e-6 relative (observed: identical floats)."""
import ctypes as C
import os

import numpy as np
import pytest

from featuredetection_b200 import capi, synthetic as syn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REAL_MODEL = "/root/reference/detect-landmarks/share/models/SDM_Model_HOG_Zhenhua_22072014.txt"


@pytest.fixture(scope="module")
def fo(built):
    from oracle import fdoracle
    return fdoracle


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "sdm.npz"))


def test_vlhog_matches_reference_hog_golden(fo, g):
    for p, ref in zip(g["hog_patches"], g["hog_ref"]):
        assert np.array_equal(fo.vlhog_uoctti(p, 10, 9), ref)


def test_vlhog_matches_reference_hog_live(fo):
    if not fo.ref_available():
        pytest.skip("oracle/_ref not built here")
    rng = np.random.default_rng(31)
    for i in range(40):
        side = int(rng.integers(12, 64))
        p = (rng.random((side, side)) * 255).astype(np.float32) if i % 2 else rng.integers(0, 256, (side, side)).astype(np.float32)
        cell = int(rng.choice([4, 8, 10]))
        assert np.array_equal(fo.vlhog_uoctti(p, cell, 9), fo.vlhog_uoctti(p, cell, 9, use_ref=True)), (i, side, cell)


def test_resize_f32_matches_cv2_golden(fo, g):
    for s in g["resize_sizes"]:
        assert np.array_equal(fo.resize_linear_f32(g["resize_src_%d" % s], 30, 30), g["resize_dst_%d" % s]), int(s)


def test_resize_f32_matches_cv2_live(fo):
    """cv2's baseline (SSE, no FMA3) code path = the arithmetic of the OpenCV 2.4.3 the reference pins; the dispatched
    AVX2 + FMA3 path of 4.13 differs in the last bit (see tests/golden/make_sdm_golden.py)"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(32)
    was = cv2.useOptimized()
    cv2.setUseOptimized(False)
    try:
        for s in range(6, 140, 2):
            src = rng.integers(0, 256, (s, s)).astype(np.float32)
            assert np.array_equal(fo.resize_linear_f32(src, 30, 30), cv2.resize(src, (30, 30), interpolation=cv2.INTER_LINEAR)), s
    finally:
        cv2.setUseOptimized(was)


def _oracle_product(f, R):
    K = R.shape[0] - 1
    acc = np.zeros(R.shape[1], np.float64)
    for k in range(K):
        acc += np.float64(f[0, k]) * R[k].astype(np.float64)
    return (acc + R[K].astype(np.float64)).astype(np.float32)


def test_regressor_product_matches_cv2_gemm_golden(g):
    mine = _oracle_product(g["gemm_f"], g["gemm_R"])
    assert np.allclose(mine, g["gemm_out"][0], rtol=1e-6, atol=0)


def test_descriptor_layout_and_quirks(fo):
    """window crop at the image border: black canvas + the reference's row-offset quirk (DescriptorExtractor.hpp:161-173)"""
    frame = syn.synthetic_frame(1)
    inside = fo.sdm_descriptors(frame, [[320.4, 240.6]], 15)
    patch = frame[241 - 15:241 + 15, 320 - 15:320 + 15].astype(np.float32)
    hog = fo.vlhog_uoctti(patch, 10, 9)
    assert np.array_equal(inside[0].reshape(31, 3, 3), hog.transpose(0, 2, 1))  # per dimension: column-major cells
    left = fo.sdm_descriptors(frame, [[4.0, 200.0]], 15)  # needs a left border only: rows are not shifted
    canvas = np.zeros((480, 640 + 11), np.float32)
    canvas[:, 11:] = frame
    assert np.array_equal(left[0].reshape(31, 3, 3), fo.vlhog_uoctti(canvas[185:215, 0:30], 10, 9).transpose(0, 2, 1))
    with pytest.raises(RuntimeError):
        fo.sdm_descriptors(frame, [[636.0, 2.0]], 15)  # right + top border: ry = y - h + borderRight leaves no valid roi ... or shifts
    

def test_whole_fit_golden(fo, g):
    frames = syn.synthetic_frames(0, 4)
    real = None
    if os.path.exists(REAL_MODEL):
        real = fo.Sdm(path=REAL_MODEL)
    synth = fo.Sdm(syn.make_sdm(68, 5, 500))
    for name, m in (("real", real), ("synth", synth)):
        if m is None:
            continue
        i = 0
        for k in range(4):
            for b in g["fit_boxes"]:
                s0 = m.align_rigid(b)
                assert np.array_equal(s0, g["fit_%s_start" % name][i])
                want = g["fit_%s_shapes" % name][i]
                if np.isnan(want[0]):
                    with pytest.raises(RuntimeError):
                        m.optimize(frames[k], s0)
                else:
                    assert np.array_equal(m.optimize(frames[k], s0), want), (name, i)
                i += 1


def test_product_model_reader_matches_reference_file(built, fo, g):
    """fdb_sdm_file_load (csrc/model_io.cpp, host only) on the reference's in-repo model"""
    if not os.path.exists(REAL_MODEL):
        pytest.skip("/root/reference not present")
    lib = capi.load_library()
    f = C.c_void_p()
    capi.check(lib, lib.fdb_sdm_file_load(REAL_MODEL.encode(), C.byref(f)))
    try:
        d = lib.fdb_sdm_file_desc(f).contents
        assert (d.num_landmarks, d.num_cascade_steps) == (15, 5)
        mean = np.ctypeslib.as_array(d.mean_landmarks, shape=(30,))
        assert np.array_equal(mean, g["real_mean"])
        om = fo.Sdm(path=REAL_MODEL).to_model()
        for s in range(5):
            R = np.ctypeslib.as_array(C.cast(d.regressors[s], C.POINTER(C.c_float)), shape=(15 * 279 + 1, 30))
            assert np.array_equal(R, om.regressors[s])
            assert R.astype(np.float64).sum() == g["real_reg_sums"][s]
            assert np.array_equal(R[-2:, :4], g["real_reg_corner"][s])
    finally:
        lib.fdb_sdm_file_free(f)


def test_product_model_reader_errors(built, tmp_path):
    lib = capi.load_library()
    f = C.c_void_p()
    assert lib.fdb_sdm_file_load(str(tmp_path / "missing.txt").encode(), C.byref(f)) == 2  # FDB_ERR_RUNTIME, like the reference's throw
    assert b"could not be opened" in lib.fdb_last_error()
    p = tmp_path / "sift.txt"
    L = 13
    p.write_text("desc\nnumLandmarks %d\n" % L + "".join("lm%d\n" % i for i in range(L)) + "".join("0.1\n" for _ in range(2 * L))
                 + "numCascadeSteps 1\ncascadeStep 1 rows 10 cols 26\ndescriptorType OpenCVSift\ndescriptorPostprocessing none\ndescriptorParameters \n")
    assert lib.fdb_sdm_file_load(str(p).encode(), C.byref(f)) == 5  # FDB_ERR_UNSUPPORTED


def test_product_model_reader_roundtrip(built, tmp_path):
    """a small model written in the reference's save() layout (SdmLandmarkModel.cpp:98-128) reads back bit-exactly"""
    lib = capi.load_library()
    m = syn.make_sdm(13, 2, 900)
    p = tmp_path / "m.txt"
    with open(p, "w") as fh:
        fh.write("# comment\r\nnumLandmarks 13\r\n")
        for i in range(13):
            fh.write("id%d\r\n" % i)
        for v in m.mean:
            fh.write("%r\r\n" % float(v))
        fh.write("numCascadeSteps 2\r\n")
        for s, R in enumerate(m.regressors):
            fh.write("cascadeStep %d rows %d cols %d\r\ndescriptorType vlhog-uoctti\r\ndescriptorPostprocessing none\r\ndescriptorParameters \r\n"
                     % (s, R.shape[0], R.shape[1]))
            for row in R:
                fh.write(" ".join(repr(float(v)) for v in row) + " \r\n")
    f = C.c_void_p()
    capi.check(lib, lib.fdb_sdm_file_load(str(p).encode(), C.byref(f)))
    try:
        d = lib.fdb_sdm_file_desc(f).contents
        assert np.array_equal(np.ctypeslib.as_array(d.mean_landmarks, shape=(26,)), m.mean)
        for s in range(2):
            R = np.ctypeslib.as_array(C.cast(d.regressors[s], C.POINTER(C.c_float)), shape=m.regressors[s].shape)
            assert np.array_equal(R, m.regressors[s])
    finally:
        lib.fdb_sdm_file_free(f)
