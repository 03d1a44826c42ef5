"""The MATLAB classifier-file readers (csrc/matfile.cpp + fdb_wvm_file_load / fdb_svm_mat_load in csrc/model_io.cpp; host
only, no GPU) on Level-5 MAT-files written by scipy.io.savemat in the variable layout WvmClassifier::loadFromMatlab
(WvmClassifier.cpp:348-770) and SvmClassifier::loadFromMatlab (SvmClassifier.cpp:240-335) read.  The reference has no .mat
file in its tree (every cfg points at absent files) and reads them through MATLAB's libmat, so the expected values are
the loader's unit conversions restated in numpy from the numbers that were written."""
import ctypes as C

import numpy as np
import pytest

from featuredetection_b200 import capi, synthetic as syn

sio = pytest.importorskip("scipy.io")


write_wvm_mat = syn.write_wvm_mat


def _arr(ptr, n, dtype):
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True) if n else np.zeros(0, dtype)


@pytest.mark.parametrize("compress", [False, True])
def test_wvm_mat_reader(built, tmp_path, compress):
    lib = capi.load_library()
    m = syn.make_wvm(20, 20, 4, 3, 0.04, seed=7)
    m = m.with_thresholds(np.linspace(-1, 1, m.n).astype(np.float32))
    cpath, tpath = str(tmp_path / "wvm.mat"), str(tmp_path / "thr.mat")
    written = write_wvm_mat(m, cpath, tpath, compress, rvm_param=compress)
    f = C.c_void_p()
    capi.check(lib, lib.fdb_wvm_file_load(cpath.encode(), tpath.encode(), C.byref(f)))
    try:
        d = lib.fdb_wvm_file_desc(f).contents
        n = m.n
        assert (d.filter_size_x, d.filter_size_y, d.num_lin_filters, d.num_filters_per_level, d.num_levels) == (20, 20, n, 4, 3)
        assert d.num_used_filters == 280 and d.limit_reliability_filter == 0.0
        p = written["param_nonlin1_rvm" if compress else "param_nonlin1"][0]
        assert d.basis_param == np.float32(p[2] / 65025.0)                                   # WvmClassifier.cpp:555
        assert np.array_equal(_arr(d.lin_thresholds, n, np.float32), np.full(n, np.float32(p[0])))   # :572-575
        assert np.array_equal(_arr(d.hk_weights, n * (n + 1) // 2, np.float32), m.hk_weights)
        assert np.array_equal(_arr(d.app_rsv_convol, n, np.float64), written["app_rsv_convol"][0] * 65025.0)  # :696
        assert np.array_equal(_arr(d.hierarchical_thresholds, n, np.float32), m.thresholds)
        assert np.array_equal(_arr(d.area_cntval, n, np.int32), m.cntval)
        nv = int(m.cntval.sum())
        want_val = np.concatenate([written["area"][0, k]["val_u"][0] for k in range(n)]) * np.float64(np.float32(255.0))  # :644
        assert np.array_equal(_arr(d.area_val, nv, np.float64), want_val)
        want_cnt = m.cntrec.copy()
        want_cnt[np.concatenate([[0], np.cumsum(m.cntval)[:-1]])] = 0
        assert np.array_equal(_arr(d.area_cntrec, nv, np.int32), want_cnt)
        rec = np.ctypeslib.as_array(C.cast(d.area_rec, C.POINTER(C.c_int32)), shape=(len(m.rec), 4))
        assert np.array_equal(rec, m.rec)
        assert (d.logistic_a, d.logistic_b) == (m.logistic_a, m.logistic_b)                  # posterior_wrvm = {B, A}
    finally:
        lib.fdb_wvm_file_free(f)


def test_wvm_mat_reader_errors(built, tmp_path):
    lib = capi.load_library()
    f = C.c_void_p()
    m = syn.make_wvm(20, 20, 2, 2, 0.04, seed=8)
    cpath, tpath = str(tmp_path / "wvm.mat"), str(tmp_path / "thr.mat")
    written = write_wvm_mat(m, cpath, tpath, False)
    assert lib.fdb_wvm_file_load(str(tmp_path / "none.mat").encode(), tpath.encode(), C.byref(f)) == 1   # invalid_argument (:371)
    assert lib.fdb_wvm_file_load(cpath.encode(), str(tmp_path / "none.mat").encode(), C.byref(f)) == 2  # runtime_error (:716)
    for drop, msg in (("num_hk", b"num_hk"), ("area", b"'area' not found"), ("app_rsv_convol", b"app_rsv_convol"), ("num_lev_wvm", b"num_lev_wvm"),
                      ("support_hk3", b"support_hk3"), ("weight_hk2", b"weight_hk2")):
        bad = {k: v for k, v in written.items() if k != drop}
        sio.savemat(str(tmp_path / "bad.mat"), bad, format="5")
        assert lib.fdb_wvm_file_load(str(tmp_path / "bad.mat").encode(), tpath.encode(), C.byref(f)) == 2, drop
        assert msg in lib.fdb_last_error(), (drop, lib.fdb_last_error())
    sio.savemat(str(tmp_path / "thr2.mat"), {"hierar_thresh": np.zeros((1, 3)), "posterior_wrvm": np.zeros((1, 2))}, format="5")
    assert lib.fdb_wvm_file_load(cpath.encode(), str(tmp_path / "thr2.mat").encode(), C.byref(f)) == 2
    assert b"hierarchicalThresholdsFromFile.size() != numLinFilters" in lib.fdb_last_error()
    (tmp_path / "junk.mat").write_bytes(b"not a mat file" * 20)
    assert lib.fdb_wvm_file_load(str(tmp_path / "junk.mat").encode(), tpath.encode(), C.byref(f)) == 1
    (tmp_path / "hdf.mat").write_bytes(b"\x89HDF\r\n\x1a\n" + bytes(600))
    assert lib.fdb_wvm_file_load(str(tmp_path / "hdf.mat").encode(), tpath.encode(), C.byref(f)) == 1
    assert b"v7.3" in lib.fdb_last_error()


@pytest.mark.parametrize("compress", [False, True])
def test_svm_mat_reader(built, tmp_path, compress):
    lib = capi.load_library()
    rng = np.random.default_rng(3)
    h, w, nsv = 16, 32, 40
    sv = rng.random((h, w, nsv))                       # MATLAB: [h][w][numSV] column-major, grey values / 255
    coef = rng.standard_normal(nsv)
    path, lpath = str(tmp_path / "svm.mat"), str(tmp_path / "log.mat")
    sio.savemat(path, {"param_nonlin1": np.array([[0.75, 2.0, 0.05, 0.0, 1.0]]), "support_nonlin1": sv, "weight_nonlin1": coef.reshape(1, -1)},
                format="5", do_compression=compress)
    sio.savemat(lpath, {"posterior_svm": np.array([[-1.25, 0.5]])}, format="5", do_compression=compress)
    f = C.c_void_p()
    capi.check(lib, lib.fdb_svm_mat_load(path.encode(), lpath.encode(), C.byref(f)))
    try:
        d = lib.fdb_svm_file_desc(f).contents
        assert (d.kernel, d.num_sv, d.dim, d.sv_type) == (capi.FDB_KERNEL_RBF, nsv, w * h, capi.FDB_SV_U8)
        assert d.gamma == float(np.float32(0.05 / 65025.0)) and d.bias == np.float32(0.75) and d.threshold == 0.0
        got = np.ctypeslib.as_array(C.cast(d.support_vectors, C.POINTER(C.c_uint8)), shape=(nsv, h, w))
        want = np.floor(255.0 * sv).astype(np.uint8).transpose(2, 0, 1)      # (uchar)(255.0 * v): truncation (SvmClassifier.cpp:308)
        assert np.array_equal(got, want)
        assert np.array_equal(_arr(d.coefficients, nsv, np.float32), coef.astype(np.float32))
        assert (d.logistic_a, d.logistic_b) == (0.5, -1.25)
    finally:
        lib.fdb_svm_file_free(f)
    # no posterior_svm: the reference warns and continues with A = B = 0 (ProbabilisticSvmClassifier.cpp:139-143)
    sio.savemat(lpath, {"something_else": np.zeros((1, 2))}, format="5")
    capi.check(lib, lib.fdb_svm_mat_load(path.encode(), lpath.encode(), C.byref(f)))
    d = lib.fdb_svm_file_desc(f).contents
    assert (d.logistic_a, d.logistic_b) == (0.0, 0.0)
    lib.fdb_svm_file_free(f)
    # polynomial kernel (SvmClassifier.cpp:270-272): PolynomialKernel(1 / divisor, basisParam / divisor, polyPower) in float arithmetic
    sio.savemat(path, {"param_nonlin1": np.array([[0.75, 1.0, 0.05, 2.0, 4.0]]), "support_nonlin1": sv, "weight_nonlin1": coef.reshape(1, -1)}, format="5")
    capi.check(lib, lib.fdb_svm_mat_load(path.encode(), None, C.byref(f)))
    d = lib.fdb_svm_file_desc(f).contents
    basis = np.float32(0.05 / 65025.0)
    assert (d.kernel, d.poly_degree) == (capi.FDB_KERNEL_POLYNOMIAL, 2)
    assert d.poly_alpha == float(np.float32(1) / np.float32(4)) and d.poly_constant == float(basis / np.float32(4))
    lib.fdb_svm_file_free(f)
    # any other kernel type: the reference throws
    sio.savemat(path, {"param_nonlin1": np.array([[0.75, 3.0, 0.05, 2.0, 4.0]]), "support_nonlin1": sv, "weight_nonlin1": coef.reshape(1, -1)}, format="5")
    assert lib.fdb_svm_mat_load(path.encode(), None, C.byref(f)) != 0


def test_storage_types_and_small_elements(built, tmp_path):
    """MATLAB stores double-class arrays in the smallest integer type that holds them (and tiny ones in the 8-byte "small
    data element" form); scipy writes the numpy dtype as is, which exercises the same storage types"""
    lib = capi.load_library()
    m = syn.make_wvm(20, 20, 2, 1, 0.04, seed=9)
    cpath, tpath = str(tmp_path / "wvm.mat"), str(tmp_path / "thr.mat")
    written = write_wvm_mat(m, cpath, tpath, False)
    written["num_hk"] = np.array([[2]], np.uint8)
    written["num_hk_wvm"] = np.array([[2]], np.int16)
    written["num_lev_wvm"] = np.array([[1]], np.int32)
    written["app_rsv_convol"] = written["app_rsv_convol"].astype(np.float32)
    sio.savemat(cpath, written, format="5")
    f = C.c_void_p()
    capi.check(lib, lib.fdb_wvm_file_load(cpath.encode(), tpath.encode(), C.byref(f)))
    d = lib.fdb_wvm_file_desc(f).contents
    assert (d.num_lin_filters, d.num_filters_per_level, d.num_levels) == (2, 2, 1)
    assert np.array_equal(_arr(d.app_rsv_convol, 2, np.float64), written["app_rsv_convol"][0].astype(np.float64) * 65025.0)
    lib.fdb_wvm_file_free(f)


@pytest.mark.parametrize("kernel_type", [2, 1])
def test_rvm_mat_reader(built, tmp_path, kernel_type):
    """RvmClassifier::loadFromMatlab (RvmClassifier.cpp:141-319) + posterior_wrvm: float32 vectors in row-major order (no
    grey-value scaling), i + 1 coefficients for level i, thresholds, kernel parameters as the SVM loader computes them"""
    lib = capi.load_library()
    rng = np.random.default_rng(4)
    n, h, w = 5, 6, 4
    svs = [rng.uniform(0, 1, (h, w)) for _ in range(n)]
    weights = [rng.normal(0, 1, i + 1) for i in range(n)]
    thr = rng.normal(0, 1, n)
    cpath, tpath = str(tmp_path / "rvm.mat"), str(tmp_path / "thr.mat")
    d = {"num_hk": np.array([[float(n)]]), "param_nonlin1_rvm": np.array([[0.5, float(kernel_type), 0.04, 3.0, 2.0]])}
    for i in range(n):
        d["support_hk%d" % (i + 1)] = svs[i]
        d["weight_hk%d" % (i + 1)] = weights[i].reshape(1, -1) if i % 2 else weights[i].reshape(-1, 1)
    sio.savemat(cpath, d, format="5", do_compression=True)
    sio.savemat(tpath, {"hierar_thresh": thr.reshape(1, -1), "posterior_wrvm": np.array([[-2.5, 0.25]])}, format="5")
    f = C.c_void_p()
    capi.check(lib, lib.fdb_rvm_file_load(cpath.encode(), tpath.encode(), C.byref(f)))
    try:
        r = lib.fdb_rvm_file_desc(f).contents
        assert (r.num_filters, r.num_filters_to_use, r.dim, r.sv_type) == (n, n, w * h, capi.FDB_SV_F32)
        basis = np.float32(0.04 / 65025.0)
        if kernel_type == 2:
            assert r.kernel == capi.FDB_KERNEL_RBF and r.gamma == float(basis)
        else:
            assert (r.kernel, r.poly_degree) == (capi.FDB_KERNEL_POLYNOMIAL, 3)
            assert r.poly_alpha == float(np.float32(1) / np.float32(2)) and r.poly_constant == float(basis / np.float32(2))
        assert r.bias == np.float32(0.5) and (r.logistic_a, r.logistic_b) == (0.25, -2.5)
        got = np.ctypeslib.as_array(C.cast(r.support_vectors, C.POINTER(C.c_float)), shape=(n, h, w))
        assert np.array_equal(got, np.stack(svs).astype(np.float32))
        coef = _arr(r.coefficients, n * (n + 1) // 2, np.float32)
        assert np.array_equal(coef, np.concatenate(weights).astype(np.float32))
        assert np.array_equal(_arr(r.hierarchical_thresholds, n, np.float32), thr.astype(np.float32))
    finally:
        lib.fdb_rvm_file_free(f)
    # error paths of the reference: missing thresholds / wrong count / missing kernel parameters
    sio.savemat(tpath, {"hierar_thresh": thr[:3].reshape(1, -1), "posterior_wrvm": np.array([[-2.5, 0.25]])}, format="5")
    assert lib.fdb_rvm_file_load(cpath.encode(), tpath.encode(), C.byref(f)) != 0
    assert b"hierarchicalThresholds.size() != coefficients.size()" in lib.fdb_last_error()
    d.pop("param_nonlin1_rvm")
    sio.savemat(cpath, d, format="5")
    assert lib.fdb_rvm_file_load(cpath.encode(), tpath.encode(), C.byref(f)) != 0
    assert b"Could not find kernel parameters" in lib.fdb_last_error()


def test_malformed_files_fail_cleanly(built, tmp_path):
    """variables with the expected DIMENSIONS but no numeric data (cell arrays), sizes from the file that would size a huge
    allocation: every loader returns a status, nothing reads out of bounds and no C++ exception crosses the C ABI"""
    lib = capi.load_library()
    f = C.c_void_p()
    m = syn.make_wvm(20, 20, 2, 2, 0.04, seed=8)
    cpath, tpath = str(tmp_path / "wvm.mat"), str(tmp_path / "thr.mat")
    written = write_wvm_mat(m, cpath, tpath, False)
    cell = np.empty((1, 1), dtype=object)
    cell[0, 0] = np.array([[1.0]])
    bad = dict(written)
    bad["weight_hk1"] = cell                       # 1x1 cell: right dims, real data empty
    sio.savemat(cpath, bad, format="5")
    assert lib.fdb_wvm_file_load(cpath.encode(), tpath.encode(), C.byref(f)) == capi.FDB_ERR_RUNTIME
    assert b"weight_hk1" in lib.fdb_last_error()
    # RVM: the same pattern
    rng = np.random.default_rng(1)
    d = {"num_hk": np.array([[2.0]]), "param_nonlin1_rvm": np.array([[0.5, 2.0, 0.04, 3.0, 2.0]]),
         "support_hk1": rng.uniform(0, 1, (4, 4)), "support_hk2": rng.uniform(0, 1, (4, 4)),
         "weight_hk1": cell, "weight_hk2": np.array([[1.0, 2.0]])}
    rpath = str(tmp_path / "rvm.mat")
    sio.savemat(rpath, d, format="5")
    assert lib.fdb_rvm_file_load(rpath.encode(), tpath.encode(), C.byref(f)) == capi.FDB_ERR_RUNTIME
    # SVM: a 3-D cell array as support_nonlin1
    cell3 = np.empty((2, 2, 2), dtype=object)
    for idx in np.ndindex(2, 2, 2):
        cell3[idx] = np.array([[0.5]])
    spath = str(tmp_path / "svm.mat")
    sio.savemat(spath, {"param_nonlin1": np.array([[0.75, 2.0, 0.05, 0.0, 1.0]]), "support_nonlin1": cell3,
                        "weight_nonlin1": np.array([[1.0, 2.0]])}, format="5")
    assert lib.fdb_svm_mat_load(spath.encode(), None, C.byref(f)) == capi.FDB_ERR_RUNTIME
    assert b"support_nonlin1" in lib.fdb_last_error()
    # text container: sizes that would ask for 8e18 elements, and a truncated vector block
    tpath2 = str(tmp_path / "svm.txt")
    with open(tpath2, "w") as fh:
        fh.write("Kernel RBF 0.5\nBias 0.1\nCoefficients 2 1.0 2.0\nSupportVectors 2 2000000000 2000000000 1 0\n")
    assert lib.fdb_svm_file_load(tpath2.encode(), C.byref(f)) == capi.FDB_ERR_RUNTIME
    with open(tpath2, "w") as fh:
        fh.write("Kernel RBF 0.5\nBias 0.1\nCoefficients 2 1.0 2.0\nSupportVectors 2 1 4 1 5\n1 2 3 4 5 6\n")
    assert lib.fdb_svm_file_load(tpath2.encode(), C.byref(f)) == capi.FDB_ERR_RUNTIME
