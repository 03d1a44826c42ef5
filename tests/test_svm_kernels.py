"""SvmClassifier with the reference's other kernels: PolynomialKernel.hpp:38-40,62-70,
HistogramIntersectionKernel.hpp:31-39,59-83, LinearKernel.hpp:27-29 next to RbfKernel. CPU: the restatement against the
reference's own kernel classes compiled into oracle/_ref (bit-exact; cv::Mat::dot of the float case is OpenCV's and is
restated from the 2.4.3 source in both, i.e. unpinned). GPU: svm_kernel against the oracle."""
import numpy as np
import pytest

from featuredetection_b200 import synthetic as syn

KERNELS = [("rbf", {}), ("polynomial", dict(alpha=1.0 / 65025.0, constant=0.5, degree=3)), ("polynomial", dict(alpha=2e-5, constant=1.0, degree=2)),
           ("hik", {}), ("linear", {})]


def _model(kind, kw, dtype, seed, num_sv=96, dim=400):
    rng = np.random.default_rng(seed)
    if dtype == np.uint8:
        sv = rng.integers(0, 256, (num_sv, dim), dtype=np.uint8)
        x = rng.integers(0, 256, (60, dim), dtype=np.uint8)
        x[0] = 0; x[1] = 255; x[2] = sv[5]
        gamma = 7.689e-7
        scale = {"rbf": 1.0, "polynomial": 1.0, "hik": 1e-4, "linear": 1e-7}[kind]
    else:
        sv = rng.normal(0, 1, (num_sv, 147)).astype(np.float32)     # odd length: the 4-way unrolled dot has a tail
        x = rng.normal(0, 1, (60, 147)).astype(np.float32)
        x[2] = sv[5]
        gamma = 0.2
        kw = dict(kw, alpha=0.5) if kind == "polynomial" else kw
        scale = 1.0 if kind != "hik" else 0.1
    coef = (rng.normal(0, 1, num_sv) * scale).astype(np.float32)
    return syn.SvmModel(sv, coef, gamma=gamma, bias=0.125, threshold=0.0, kernel=kind, **kw), x


@pytest.mark.parametrize("kind,kw", KERNELS)
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_oracle_kernels_equal_the_compiled_reference(built, kind, kw, dtype):
    from oracle import fdoracle as fo
    if not fo.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference)")
    model, x = _model(kind, kw, dtype, 11)
    d, p, q = fo.Svm(model).eval(x)
    rd, rp, rq = fo.Svm(model, use_ref=True).eval(x)
    assert np.array_equal(d, rd), np.max(np.abs(d - rd))          # bit for bit
    assert np.array_equal(p, rp) and np.array_equal(q, rq)
    assert np.isfinite(d).all() and d.std() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind,kw", KERNELS)
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_gpu_kernels_match_oracle(ctx, kind, kw, dtype):
    from oracle import fdoracle as fo
    from featuredetection_b200.detector import ProbabilisticSvmClassifier
    model, x = _model(kind, kw, dtype, 12)
    d, p, q = ProbabilisticSvmClassifier(ctx, model).get_probability(x)
    rd, rp, rq = fo.Svm(model).eval(x)
    if kind == "rbf":   # CUDA exp vs glibc exp: <= 1 ulp per kernel value
        assert np.max(np.abs(d - rd)) <= 1e-9
    else:               # integer / float64 arithmetic in the reference's order: identical bits
        assert np.array_equal(d, rd), np.max(np.abs(d - rd))
    assert np.array_equal(q, rq) and np.max(np.abs(p - rp)) <= 1e-9


@pytest.mark.gpu
def test_text_model_with_polynomial_kernel(ctx, tmp_path):
    """SvmClassifier::load text container (SvmClassifier.cpp:101-160) with a Polynomial kernel line"""
    import ctypes as C
    from featuredetection_b200 import capi
    path = tmp_path / "svm.txt"
    sv = np.arange(24, dtype=np.float32).reshape(3, 8) / 10
    with open(path, "w") as f:
        f.write("Kernel Polynomial 3 0.5 0.25\nBias 0.1\nCoefficients 3\n0.5\n-0.25\n1\nSupportVectors 3 1 8 1 5\n")
        for row in sv:
            f.write(" ".join(repr(float(v)) for v in row) + "\n")
    h = C.c_void_p()
    capi.check(ctx.lib, ctx.lib.fdb_svm_file_load(str(path).encode(), C.byref(h)))
    d = ctx.lib.fdb_svm_file_desc(h).contents
    assert (d.kernel, d.poly_degree, d.poly_constant, d.poly_alpha) == (capi.FDB_KERNEL_POLYNOMIAL, 3, 0.5, 0.25)
    ctx.lib.fdb_svm_file_free(h)
