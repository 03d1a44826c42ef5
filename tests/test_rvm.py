"""RvmClassifier / ProbabilisticRvmClassifier (RvmClassifier.cpp:66-112, ProbabilisticRvmClassifier.cpp:52-64):
the cascade as the reference's live code path evaluates it. CPU: the C restatement against the reference's own classes
compiled into oracle/_ref (bit-exact). GPU: fdb_rvm_get_probability and the `single` prvm detector against the oracle."""
import numpy as np
import pytest

from featuredetection_b200 import synthetic as syn


def _patches(rng, model, n):
    sv = model.sv
    x = np.clip(sv[rng.integers(0, len(sv), n)].astype(np.int64) + rng.integers(-60, 61, (n, sv.shape[1])), 0, 255).astype(np.uint8)
    x[0] = 0; x[1] = 255; x[2] = sv[0]
    return x


@pytest.mark.parametrize("use", [0, 1, 7, 24, 99])
def test_oracle_rvm_equals_the_compiled_reference(built, use):
    from oracle import fdoracle as fo
    if not fo.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference)")
    model = syn.make_rvm(20, 20, seed=3)
    model.use = use
    x = _patches(np.random.default_rng(use), model, 200)
    lv, d, p, q = fo.Rvm(model).eval(x)
    rl, rd, rp, rq = fo.Rvm(model, use_ref=True).eval(x)
    assert np.array_equal(lv, rl) and np.array_equal(d, rd) and np.array_equal(p, rp) and np.array_equal(q, rq)
    assert lv.max() == model.filters_to_use - 1 and len(np.unique(lv)) > min(3, model.filters_to_use - 1)   # a spread of exits
    if model.filters_to_use > 1:
        assert 0 < q.sum() < len(q)


def test_oracle_rvm_other_kernels_and_float_vectors(built):
    from oracle import fdoracle as fo
    if not fo.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference)")
    rng = np.random.default_rng(5)
    n = 12
    sv = rng.normal(0, 1, (n, 147)).astype(np.float32)
    coef = rng.normal(0, 1, n * (n + 1) // 2).astype(np.float32)
    for kind, kw in (("polynomial", dict(alpha=0.1, constant=1.0, degree=3)), ("linear", {}), ("hik", {}), ("rbf", {})):
        model = syn.RvmModel(sv, coef, np.full(n, -50.0, np.float32), gamma=0.05, bias=0.25, kernel=kind, **kw)
        x = rng.normal(0, 1, (40, 147)).astype(np.float32)
        a = fo.Rvm(model).eval(x)
        b = fo.Rvm(model, use_ref=True).eval(x)
        for u, v in zip(a, b):
            assert np.array_equal(u, v), kind


@pytest.mark.gpu
@pytest.mark.parametrize("use", [0, 1, 7])
def test_gpu_rvm_get_probability(ctx, use):
    from oracle import fdoracle as fo
    from featuredetection_b200.detector import ProbabilisticRvmClassifier
    model = syn.make_rvm(20, 20, seed=3)
    model.use = use
    x = _patches(np.random.default_rng(10 + use), model, 300)
    lv, d, p, q = ProbabilisticRvmClassifier(ctx, model).get_probability(x)
    rl, rd, rp, rq = fo.Rvm(model).eval(x)
    assert np.array_equal(lv, rl), int((lv != rl).sum())
    assert np.max(np.abs(d - rd)) <= 1e-9 and np.max(np.abs(p - rp)) <= 1e-9 and np.array_equal(q, rq)


@pytest.mark.gpu
def test_gpu_rvm_float_kernels(ctx):
    from oracle import fdoracle as fo
    from featuredetection_b200.detector import ProbabilisticRvmClassifier
    rng = np.random.default_rng(6)
    n = 12
    sv = rng.normal(0, 1, (n, 147)).astype(np.float32)
    coef = rng.normal(0, 1, n * (n + 1) // 2).astype(np.float32)
    for kind, kw in (("polynomial", dict(alpha=0.1, constant=1.0, degree=3)), ("linear", {}), ("hik", {})):
        model = syn.RvmModel(sv, coef, np.full(n, -50.0, np.float32), bias=0.25, kernel=kind, **kw)
        x = rng.normal(0, 1, (40, 147)).astype(np.float32)
        a = ProbabilisticRvmClassifier(ctx, model).get_probability(x)
        b = fo.Rvm(model).eval(x)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3]), kind


@pytest.mark.gpu
def test_single_prvm_detector(ctx, face_models):
    """ffpDetectApp `single` detector with classifier prvm: every window of the pyramid through the RVM cascade"""
    from oracle import fdoracle as fo
    from featuredetection_b200.detector import SlidingWindowCascade
    det_kw, _, _ = face_models
    kw = dict(det_kw, min_scale_factor=0.09, max_scale_factor=0.16)
    frames = np.ascontiguousarray(syn.synthetic_frames(50, 2)[:, :240, :320])
    model = syn.make_rvm(20, 20, seed=3, num_filters=12, survival=0.6)
    casc = SlidingWindowCascade(ctx, kw, None, rvm_model=model)
    casc.prepare(320, 240, 2)
    dets, dist = casc.detect_single(frames)
    ro = fo.Rvm(model)
    import ctypes as C
    from featuredetection_b200 import capi
    desc = syn.detector_desc(**kw)
    for k in range(2):
        # the oracle's own window grid (fdo_enumerate: DirectPyramidFeatureExtractor.cpp:83-121, bounds from the theoretical scale)
        L = fo.lib()
        p = L.fdo_pyramid_build(frames[k].ctypes.data, 320, 240, 320, desc.incremental_scale_factor, desc.min_scale_factor, desc.max_scale_factor)
        infos = (capi.LayerInfo * p.contents.n_layers)()
        total = L.fdo_enumerate(p, 20, 20, 1, 1, 0, 0, 0, 0, infos, p.contents.n_layers)
        L.fdo_pyramid_free(p)
        _, layers = fo.pyramid(frames[k], kw["incremental_scale_factor"], kw["min_scale_factor"], kw["max_scale_factor"])
        patches = np.stack([fo.hq64(img[y:y + 20, x:x + 20]).ravel() for (_, _, img), info in zip(layers, infos)
                            for y in range(info.windows_y) for x in range(info.windows_x)])
        assert len(patches) == total
        rl, rd, rp, rq = ro.eval(patches)
        assert dist.shape[1] == len(patches)
        assert np.max(np.abs(dist[k] - rd)) <= 1e-9
        mine = dets[dets["frame"] == k]
        assert list(mine["window"]) == list(np.nonzero(rq)[0])
        assert np.array_equal(mine["wvm_level"], rl[rq.astype(bool)])
        assert np.max(np.abs(mine["probability"] - rp[rq.astype(bool)])) <= 1e-9
    assert len(dets) > 0
