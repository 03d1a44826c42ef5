"""fdb_detector_set (csrc/detector_set.cu): all 15 ffpDetectApp landmark detectors on every frame over shared pyramids
(ffpDetectApp.cpp:391-500 builds them, :548-596 runs them). The set must return exactly what its members return one by one -
and that is checked against the CPU oracle for every one of the 15 cfg models, at 640x480 (BASELINE configs[3] shape) and on a
1920x1080 frame (configs[2] geometry). Scores: 1e-4 (north_star); levels, windows, detections: exact."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from featuredetection_b200 import capi, synthetic as syn
from featuredetection_b200.detector import SlidingWindowCascade, DetectorSet, SCORE_DTYPE

TOL = 1e-4
NAMES = [c[0] for c in syn.LANDMARK_CONFIGS]


def _oracle():
    from oracle import fdoracle as fo
    fo.build()
    return fo


@pytest.fixture(scope="module")
def fifteen(ctx):
    models = [syn.landmark_models(nm) for nm in NAMES]
    # room for the stage-1 positives of a 1920x1080 frame (the default list holds 4096 per frame)
    cascs = [SlidingWindowCascade(ctx, dict(kw, max_positives_per_frame=65536), wvm, svm) for kw, wvm, svm in models]
    return models, cascs, DetectorSet(ctx, cascs)


def _dense_buffers(cascs, n):
    import torch
    bufs = [torch.full((n, c.windows_per_frame, 2), -1, dtype=torch.int32, device="cuda") for c in cascs]
    return bufs, [b.data_ptr() for b in bufs]


def _dense_host(buf):
    return buf.cpu().numpy().view(SCORE_DTYPE)[..., 0]


def _check_against_oracle(fo, models, dets, dense, frames, first_frame_index=0):
    def one(args):
        d, k = args
        kw, wvm, svm = models[d]
        return d, k, fo.detect_frame(kw, fo.Wvm(wvm), fo.Svm(svm), frames[k], stage=capi.FDB_STAGE_NMS, frame_index=k)
    jobs = [(d, k) for d in range(len(models)) for k in range(len(frames))]
    with ThreadPoolExecutor(16) as ex:  # the oracle's C functions release the GIL
        for d, k, ref in ex.map(one, jobs):
            got = _dense_host(dense[d])[k]
            assert ref["windows"] == got.shape[0], NAMES[d]
            assert np.array_equal(got["level"], ref["dense"]["level"]), NAMES[d]
            assert np.max(np.abs(got["fout"] - ref["dense"]["fout"])) <= TOL, NAMES[d]
            mine = dets[(dets["reserved"] == d) & (dets["frame"] == k)]
            assert list(mine["window"]) == list(ref["detections"]["window"]), NAMES[d]
            assert np.allclose(mine["svm_distance"], ref["detections"]["svm_distance"], rtol=0, atol=TOL), NAMES[d]
            for f in ("layer", "x", "y", "center_x", "center_y", "width", "height"):
                assert np.array_equal(mine[f], ref["detections"][f]), (NAMES[d], f)


@pytest.mark.gpu
def test_fifteen_models_640x480_against_the_oracle(ctx, fifteen):
    """every cfg model, whole five-stage cascade, 2 frames: dense stage-1 records and detections of the set = the oracle's"""
    import torch
    fo = _oracle()
    models, cascs, dset = fifteen
    dset.prepare(640, 480, 2)
    info = dset.info()
    assert info["fast_members"] == 15
    assert dset.windows_per_frame == 4302040  # sum over the 15 cfgs
    frames = syn.synthetic_frames(60, 2)
    dev = torch.from_numpy(frames).cuda()
    dense, ptrs = _dense_buffers(cascs, 2)
    dets = dset.detect_device(dev.data_ptr(), 2, stage=capi.FDB_STAGE_NMS, dense_ptrs=ptrs)
    torch.cuda.synchronize()
    _check_against_oracle(fo, models, dets, dense, frames)
    # host-frame entry point: the same detections
    dets_h = dset.detect(frames, stage=capi.FDB_STAGE_NMS)
    assert np.array_equal(dets_h, dets)


@pytest.mark.gpu
def test_set_equals_its_members_over_a_chunked_batch(ctx, fifteen):
    """7 frames through the slot pipeline (chunks of 2): the set's detections = each member's own fdb_detect_batch"""
    models, cascs, dset = fifteen
    dset.prepare(640, 480, 7)
    frames = syn.synthetic_frames(70, 7)
    dets = dset.detect(frames, stage=capi.FDB_STAGE_NMS)
    for d, c in enumerate(cascs):
        own = c.detect(frames, stage=capi.FDB_STAGE_NMS)
        mine = dets[dets["reserved"] == d].copy()
        mine["reserved"] = 0
        assert np.array_equal(mine, own), NAMES[d]
        assert list(c.last_counts())[0] == c.windows_per_frame * 7


@pytest.mark.gpu
def test_shared_pyramid_is_the_union_of_the_members(ctx, fifteen):
    """4 distinct pyramid parameter sets in the 15 cfgs: the set builds each distinct image once"""
    models, cascs, dset = fifteen
    dset.prepare(640, 480, 1)
    info = dset.info()
    separate = sum(c.pyramid_bytes for c in cascs)
    assert info["pyramid_bytes"] < separate / 3          # 12 members share the 4-layer pyramid
    distinct = {}
    for (kw, _, _), c in zip(models, cascs):
        distinct[(kw["incremental_scale_factor"], kw["min_scale_factor"], kw["max_scale_factor"])] = c.pyramid_bytes
    assert info["pyramid_bytes"] <= sum(distinct.values())  # never more than the 4 distinct pyramids


@pytest.mark.gpu
def test_fifteen_models_on_a_full_hd_frame(ctx, fifteen):
    """BASELINE configs[2] geometry: 15 models on one 1920x1080 frame (32.1 M windows), stage-1 records and detections
    against the oracle"""
    import torch
    fo = _oracle()
    models, cascs, dset = fifteen
    dset.prepare(1920, 1080, 1)
    assert dset.windows_per_frame == 32113402
    frame = syn.synthetic_frame(80, 1920, 1080)[None]
    dev = torch.from_numpy(frame).cuda()
    dense, ptrs = _dense_buffers(cascs, 1)
    dets = dset.detect_device(dev.data_ptr(), 1, stage=capi.FDB_STAGE_NMS, dense_ptrs=ptrs, det_cap=1 << 18)
    torch.cuda.synchronize()
    _check_against_oracle(fo, models, dets, dense, frame)


@pytest.mark.gpu
def test_member_outside_the_group_kernels_runs_alone(ctx):
    """a member with a window step of 2 cannot use the strip kernels: the set runs it through its own pipeline"""
    kw, wvm, svm = syn.landmark_models("FaceFrontal")
    a = SlidingWindowCascade(ctx, kw, wvm, svm)
    b = SlidingWindowCascade(ctx, dict(kw, step_x=2, step_y=2), wvm, svm)
    dset = DetectorSet(ctx, [a, b])
    dset.prepare(640, 480, 2)
    assert dset.info()["fast_members"] == 1
    frames = syn.synthetic_frames(5, 2)
    dets = dset.detect(frames)
    for d, c in enumerate((a, b)):
        own = c.detect(frames)
        mine = dets[dets["reserved"] == d].copy()
        mine["reserved"] = 0
        assert np.array_equal(mine, own)


@pytest.mark.gpu
def test_set_argument_errors(ctx):
    kw, wvm, svm = syn.landmark_models("FaceFrontal")
    a = SlidingWindowCascade(ctx, kw, wvm, svm)
    with pytest.raises(capi.FdbError):
        DetectorSet(ctx, [a, a])
    single = SlidingWindowCascade(ctx, kw, None, svm)
    with pytest.raises(capi.FdbError):
        DetectorSet(ctx, [a, single])
    dset = DetectorSet(ctx, [a])
    dset.width, dset.height = 640, 480  # not prepared: the library refuses
    with pytest.raises(capi.FdbError):
        dset.detect(np.zeros((1, 480, 640), np.uint8))


@pytest.mark.gpu
def test_both_window_kernels_return_the_same_records(ctx, fifteen, monkeypatch):
    """wvm_group_tc.cu (tcgen05, packs of up to four models) and wvm_group.cu (mma.sync, packs of two) are two schedules of the same
    arithmetic: dense stage-1 records (float bits and levels) and detections must be identical, also for a frame count that
    leaves the last group of four frames incomplete. FDB_WINDOW_KERNEL is read at prepare (packing) and at every launch."""
    import torch
    models, cascs, dset = fifteen
    frames = syn.synthetic_frames(90, 5)
    dev = torch.from_numpy(frames).cuda()
    got = {}
    for choice in ("mma", "tc", None):
        if choice is None:
            monkeypatch.delenv("FDB_WINDOW_KERNEL", raising=False)
        else:
            monkeypatch.setenv("FDB_WINDOW_KERNEL", choice)
        dset.prepare(640, 480, 5)
        dense, ptrs = _dense_buffers(cascs, 5)
        dets = dset.detect_device(dev.data_ptr(), 5, stage=capi.FDB_STAGE_NMS, dense_ptrs=ptrs)
        torch.cuda.synchronize()
        got[choice] = (dets, [b.cpu().numpy() for b in dense])
    for choice in ("tc", None):
        assert np.array_equal(got[choice][0], got["mma"][0]), choice
        for d in range(len(cascs)):
            assert np.array_equal(got[choice][1][d], got["mma"][1][d]), (choice, NAMES[d])


@pytest.mark.gpu
def test_packs_of_three_without_early_exit(ctx, monkeypatch):
    """three models of one geometry whose cascades never reject (every window reaches the deep kernel's queue, which overflows and
    sends the members down the generic path): the set still equals its members, on either window kernel"""
    names = ("LeftLipCorner", "RightLipCorner", "LeftNoseCorner")
    models = [syn.landmark_models(nm, profile="no-exit") for nm in names]
    frames = syn.synthetic_frames(95, 2)[:, :120, :160].copy()
    ref = None
    for choice in ("mma", "tc"):
        monkeypatch.setenv("FDB_WINDOW_KERNEL", choice)
        cascs = [SlidingWindowCascade(ctx, dict(kw, min_scale_factor=0.5, max_scale_factor=1.0, max_positives_per_frame=65536), wvm, svm)
                 for kw, wvm, svm in models]
        dset = DetectorSet(ctx, cascs)
        dset.prepare(160, 120, 2)
        dets = dset.detect(frames, stage=capi.FDB_STAGE_WVM, det_cap=1 << 20)
        for d, c in enumerate(cascs):
            own = c.detect(frames, stage=capi.FDB_STAGE_WVM, det_cap=1 << 20)
            mine = dets[dets["reserved"] == d].copy()
            mine["reserved"] = 0
            assert mine.tobytes() == own.tobytes(), (choice, names[d])  # bytes: stage-1 records carry NaN for the SVM fields
        if ref is None:
            ref = dets
        else:
            assert dets.tobytes() == ref.tobytes()
