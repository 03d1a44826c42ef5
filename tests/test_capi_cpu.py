"""CPU-side checks of the C ABI library: it loads, exports every declared symbol, refuses to run
without a GPU (no CPU fallback), and its host-only entry points agree with the oracle."""
import ctypes as C
import re
import os

import numpy as np
import pytest

from featuredetection_b200 import capi, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(built):
    header = open(os.path.join(ROOT, "include", "fdb200.h")).read()
    declared = set(re.findall(r"FDB_API\s+[\w\s\*]+?\b(fdb_\w+)\s*\(", header))
    bound = {name for name, _, _ in capi.SYMBOLS}
    assert declared == bound, declared ^ bound
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.fdb_abi_version() == 2


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="GPU present")
def test_no_cpu_fallback(built):
    lib = capi.load_library()
    h = C.c_void_p()
    status = lib.fdb_ctx_create(-1, C.byref(h))
    assert status == 4 and not h.value  # FDB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.fdb_last_error()


def _plan(lib, kw, W, H, roi=(0, 0, 0, 0)):
    desc = syn.detector_desc(**kw)
    infos = (capi.LayerInfo * 64)()
    nl, nw = C.c_int32(), C.c_int64()
    capi.check(lib, lib.fdb_plan_layers(C.byref(desc), W, H, roi[0], roi[1], roi[2], roi[3], infos, 64, C.byref(nl), C.byref(nw)))
    return [infos[i] for i in range(nl.value)], nw.value


# windows per 640x480 / 1920x1080 frame for the 15 ffpDetectApp cfgs
WINDOWS = {
    "FaceFrontal": (16185, 190616), "FaceLeftProfile": (59721, 538963), "RightEyeCenter": (158393, 1181132),
    "LeftEyeCenter": (366100, 2708968), "CenterLipUpperOuter": (363372, 2693536), "NoseTip": (355172, 2674136),
    "LeftEarCenter": (371572, 2712936),
}


@pytest.mark.parametrize("name", sorted(WINDOWS))
def test_window_counts_match_survey(built, name):
    lib = capi.load_library()
    idx, (nm, pw, ph, inc, mn, mx, per, lev, r) = syn.landmark_config(name)
    kw = dict(incremental_scale_factor=float(np.float32(inc)), min_scale_factor=float(np.float32(mn)),
              max_scale_factor=float(np.float32(mx)), patch_width=pw, patch_height=ph, step_x=1, step_y=1)
    assert _plan(lib, kw, 640, 480)[1] == WINDOWS[name][0]
    assert _plan(lib, kw, 1920, 1080)[1] == WINDOWS[name][1]


def test_fifteen_models_total(built):
    lib = capi.load_library()
    tot = [0, 0]
    for (nm, pw, ph, inc, mn, mx, per, lev, r) in syn.LANDMARK_CONFIGS:
        kw = dict(incremental_scale_factor=float(np.float32(inc)), min_scale_factor=float(np.float32(mn)),
                  max_scale_factor=float(np.float32(mx)), patch_width=pw, patch_height=ph, step_x=1, step_y=1)
        tot[0] += _plan(lib, kw, 640, 480)[1]
        tot[1] += _plan(lib, kw, 1920, 1080)[1]
    assert tot == [4302040, 32113402]


@pytest.mark.parametrize("size", [(640, 480), (1920, 1080), (333, 250), (97, 61), (21, 21)])
@pytest.mark.parametrize("roi", [(0, 0, 0, 0), (50, 30, 200, 150), (-5, -5, 5000, 5000)])
def test_plan_matches_oracle_enumeration(built, size, roi):
    """product plan.cpp vs oracle fd_oracle.c: same layers, same window grid."""
    from oracle import fdoracle as fo
    lib = capi.load_library()
    W, H = size
    for (nm, pw, ph, inc, mn, mx, per, lev, r) in syn.LANDMARK_CONFIGS[:5]:
        kw = dict(incremental_scale_factor=float(np.float32(inc)), min_scale_factor=float(np.float32(mn)),
                  max_scale_factor=float(np.float32(mx)), patch_width=pw, patch_height=ph, step_x=1, step_y=2)
        mine, nw = _plan(lib, kw, W, H, roi)
        frame = np.zeros((H, W), np.uint8)
        L = fo.lib()
        p = L.fdo_pyramid_build(frame.ctypes.data, W, H, W, kw["incremental_scale_factor"], kw["min_scale_factor"], kw["max_scale_factor"])
        infos = (capi.LayerInfo * 64)()
        tot = L.fdo_enumerate(p, pw, ph, 1, 2, roi[0], roi[1], roi[2], roi[3], infos, 64)
        n = p.contents.n_layers
        L.fdo_pyramid_free(p)
        assert nw == tot and len(mine) == n
        for a, i in zip(mine, range(n)):
            b = infos[i]
            for f, _ in capi.LayerInfo._fields_:
                assert getattr(a, f) == getattr(b, f), (nm, f)


def test_invalid_arguments(built):
    lib = capi.load_library()
    base = dict(incremental_scale_factor=0.9, min_scale_factor=0.1, max_scale_factor=0.5, patch_width=20, patch_height=20)
    for bad in (dict(incremental_scale_factor=1.0), dict(incremental_scale_factor=0.0), dict(min_scale_factor=0.0),
                dict(max_scale_factor=1.5), dict(step_x=-1)):
        desc = syn.detector_desc(**dict(base, **bad))
        nl, nw = C.c_int32(), C.c_int64()
        assert lib.fdb_plan_layers(C.byref(desc), 640, 480, 0, 0, 0, 0, None, 0, C.byref(nl), C.byref(nw)) == 1
        assert lib.fdb_last_error()


def _random_dets(rng, n, W=640, H=480, ties=False):
    from featuredetection_b200.detector import DETECTION_DTYPE
    d = np.zeros(n, DETECTION_DTYPE)
    d["center_x"] = rng.integers(0, W, n); d["center_y"] = rng.integers(0, H, n)
    d["width"] = rng.choice([135, 147, 160, 174, 190], n); d["height"] = d["width"]
    d["window"] = np.arange(n)
    p = rng.uniform(0.2, 0.9, n)
    if ties:
        p = np.round(p, 1)
    d["probability"] = p
    return d


@pytest.mark.parametrize("seed", range(6))
def test_overlap_elimination_vs_oracle_and_reference(built, seed):
    from oracle import fdoracle as fo
    lib = capi.load_library()
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    d = _random_dets(rng, n, 200, 150)
    dist, ratio = [(5.0, 0.0), (0.5, 0.7), (12.0, 0.9)][seed % 3]
    mine = d.copy(); cnt = C.c_int64()
    capi.check(lib, lib.fdb_overlap_eliminate(mine.ctypes.data, n, dist, ratio, C.byref(cnt)))
    mine = mine[:cnt.value]
    if fo.ref_available():
        keep = np.zeros(n, np.int32)
        cx, cy, w, p = (np.ascontiguousarray(d[f]) for f in ("center_x", "center_y", "width", "probability"))
        m = fo.ref().ref_overlap_eliminate(dist, ratio, n, cx.ctypes.data, cy.ctypes.data, w.ctypes.data, p.ctypes.data, keep.ctypes.data)
        assert list(keep[:m]) == list(mine["window"])  # distinct probabilities: no tie ambiguity


@pytest.mark.parametrize("seed", range(3))
def test_overlap_elimination_grid_path_vs_reference(built, seed):
    """hundreds to thousands of candidates per frame (the eye / lip / nose detectors): the product's grid-bucketed elimination
    (absolute distance, n > 64) returns the survivors of the reference's own OverlapElimination class"""
    from oracle import fdoracle as fo
    lib = capi.load_library()
    rng = np.random.default_rng(100 + seed)
    n = [700, 3000, 5000][seed]
    d = _random_dets(rng, n, 640, 480)
    d["width"] = rng.choice([34, 38, 44, 48], n); d["height"] = d["width"]
    d["probability"] = rng.permutation(n) / float(n)  # distinct: no tie ambiguity
    dist, ratio = [(5.0, 0.0), (5.0, 0.85), (9.5, 0.0)][seed]
    mine = d.copy(); cnt = C.c_int64()
    capi.check(lib, lib.fdb_overlap_eliminate(mine.ctypes.data, n, dist, ratio, C.byref(cnt)))
    mine = mine[:cnt.value]
    assert 0 < cnt.value < n
    if fo.ref_available():
        keep = np.zeros(n, np.int32)
        cx, cy, w, p = (np.ascontiguousarray(d[f]) for f in ("center_x", "center_y", "width", "probability"))
        m = fo.ref().ref_overlap_eliminate(dist, ratio, n, cx.ctypes.data, cy.ctypes.data, w.ctypes.data, p.ctypes.data, keep.ctypes.data)
        assert list(keep[:m]) == list(mine["window"])


def test_host_nms_vs_oracle(built):
    """fdb_five_stage_nms (sparse, product) vs the oracle's dense-map restatement, incl. the
    all-0.5 probabilities the five-stage detector really produces and the no-maximum fallback."""
    from oracle import fdoracle as fo
    lib = capi.load_library()
    det_kw, wvm, svm = syn.landmark_models("FaceFrontal")
    wo, so = fo.Wvm(wvm), fo.Svm(svm)
    for k in range(6):
        frame = syn.synthetic_frame(60 + k)
        pre = fo.detect_frame(det_kw, wo, so, frame, stage=capi.FDB_STAGE_SVM)["detections"]
        ref = fo.detect_frame(det_kw, wo, so, frame, stage=capi.FDB_STAGE_NMS)["detections"]
        # stage-3 output is sorted; NMS input order in the reference is the OE order = same for ties-stable sort
        mine = pre.copy(); cnt = C.c_int64()
        capi.check(lib, lib.fdb_five_stage_nms(mine.ctypes.data, len(mine), 640, 480, C.byref(cnt)))
        assert list(mine[:cnt.value]["window"]) == list(ref["window"])
