"""Thin Python front-end over the C ABI (include/fdb200.h) for tests, bench.py and smoke().

The product's host side is the C++ inside libfdb200.so (api.cu / plan.cpp / hostpost.cpp) plus the
C++ adapters in adapters/ that implement the reference's own interfaces; this module only moves
numpy arrays across the ABI. It mirrors the reference's object graph
(ffpDetectApp.cpp:391-425): ProbabilisticWvmClassifier, ProbabilisticSvmClassifier and a
FiveStageSlidingWindowDetector built from them.
"""
import ctypes as C

import numpy as np

from . import capi
from .synthetic import detector_desc

DETECTION_DTYPE = np.dtype(
    [(name, {C.c_int32: "i4", C.c_int64: "i8", C.c_float: "f4", C.c_double: "f8"}[ct])
     for name, ct in capi.Detection._fields_], align=True)
assert DETECTION_DTYPE.itemsize == C.sizeof(capi.Detection)
SCORE_DTYPE = np.dtype([("fout", "f4"), ("level", "i4")])
LAYER_FIELDS = [f for f, _ in capi.LayerInfo._fields_]


def feature_shape(feature, patch_w, patch_h):
    """(dim, numpy dtype) of a feature space for a patch size (fdb_feature_shape, host only)"""
    lib = capi.load_library()
    dim, is_float = C.c_int32(), C.c_int32()
    capi.check(lib, lib.fdb_feature_shape(C.byref(feature), patch_w, patch_h, C.byref(dim), C.byref(is_float)))
    return dim.value, (np.float32 if is_float.value else np.uint8)


class Context:
    """fdb_ctx: one CUDA device + stream. Raises FdbError when no sm_100 GPU is usable."""

    def __init__(self, device=-1):
        self.lib = capi.load_library()
        h = C.c_void_p()
        capi.check(self.lib, self.lib.fdb_ctx_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.fdb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        capi.check(self.lib, self.lib.fdb_ctx_synchronize(self.h))

    def stream(self):
        return self.lib.fdb_ctx_stream(self.h)

    def launch_count(self):
        return int(self.lib.fdb_ctx_launch_count(self.h))

    def timer_start(self):
        capi.check(self.lib, self.lib.fdb_ctx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        capi.check(self.lib, self.lib.fdb_ctx_timer_stop(self.h, C.byref(ms)))
        return ms.value


def gray_from_bgr(ctx, frames_bgr, pitch=None):
    """GrayscaleFilter::applyTo on [n, H, W, 3] interleaved BGR frames (fdb_gray_from_bgr) -> [n, H, W] u8"""
    frames = np.ascontiguousarray(frames_bgr, np.uint8)
    if frames.ndim == 3:
        frames = frames[None]
    n, H, W, ch = frames.shape
    assert ch == 3
    out = np.empty((n, H, W), np.uint8)
    capi.check(ctx.lib, ctx.lib.fdb_gray_from_bgr(ctx.h, frames.ctypes.data, pitch or 3 * W, W, H, n, out.ctypes.data))
    return out


class ProbabilisticWvmClassifier:
    """classification::ProbabilisticWvmClassifier (ProbabilisticWvmClassifier.cpp:42-54) on the GPU."""

    def __init__(self, ctx, model):
        self.ctx, self.model = ctx, model
        self._desc = model.desc()
        h = C.c_void_p()
        capi.check(ctx.lib, ctx.lib.fdb_wvm_create(ctx.h, C.byref(self._desc), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.fdb_wvm_destroy(self.h)
        except Exception:
            pass

    def set_limit_reliability_filter(self, value):
        capi.check(self.ctx.lib, self.ctx.lib.fdb_wvm_set_limit_reliability_filter(self.h, value))

    def get_probability(self, patches):
        """patches [n, w*h] u8 -> (level, fout, probability, positive)"""
        p = np.ascontiguousarray(patches, np.uint8).reshape(-1, self.model.w * self.model.h)
        n = p.shape[0]
        level = np.empty(n, np.int32); fout = np.empty(n, np.float32)
        prob = np.empty(n, np.float64); pos = np.empty(n, np.uint8)
        capi.check(self.ctx.lib, self.ctx.lib.fdb_wvm_get_probability(
            self.h, p.ctypes.data, n, level.ctypes.data, fout.ctypes.data, prob.ctypes.data, pos.ctypes.data))
        return level, fout, prob, pos


class ProbabilisticSvmClassifier:
    """classification::ProbabilisticSvmClassifier (ProbabilisticSvmClassifier.cpp:42-58) on the GPU."""

    def __init__(self, ctx, model):
        self.ctx, self.model = ctx, model
        self._desc = model.desc()
        h = C.c_void_p()
        capi.check(ctx.lib, ctx.lib.fdb_svm_create(ctx.h, C.byref(self._desc), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.fdb_svm_destroy(self.h)
        except Exception:
            pass

    def set_threshold(self, t):
        capi.check(self.ctx.lib, self.ctx.lib.fdb_svm_set_threshold(self.h, t))

    @property
    def has_dense(self):
        """True when batches run on the tensor cores (csrc/svm_dense.cu)"""
        return bool(self.ctx.lib.fdb_svm_has_dense(self.h))

    def get_probability(self, vectors):
        v = np.ascontiguousarray(vectors, self.model.sv.dtype).reshape(-1, self.model.sv.shape[1])
        n = v.shape[0]
        dist = np.empty(n, np.float64); prob = np.empty(n, np.float64); pos = np.empty(n, np.uint8)
        capi.check(self.ctx.lib, self.ctx.lib.fdb_svm_get_probability(
            self.h, v.ctypes.data, n, dist.ctypes.data, prob.ctypes.data, pos.ctypes.data))
        return dist, prob, pos


class ProbabilisticRvmClassifier:
    """classification::ProbabilisticRvmClassifier over an RvmClassifier (fdb_rvm)."""

    def __init__(self, ctx, model):
        self.ctx, self.model = ctx, model
        self._desc = model.desc()
        self.h = None
        h = C.c_void_p()
        capi.check(ctx.lib, ctx.lib.fdb_rvm_create(ctx.h, C.byref(self._desc), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.fdb_rvm_destroy(self.h)
        except Exception:
            pass

    def set_num_filters_to_use(self, n):
        capi.check(self.ctx.lib, self.ctx.lib.fdb_rvm_set_num_filters_to_use(self.h, n))

    def get_probability(self, vectors):
        """-> (level i32[n], distance f64[n], probability f64[n], positive u8[n])"""
        v = np.ascontiguousarray(vectors, self.model.sv.dtype).reshape(-1, self.model.sv.shape[1])
        n = v.shape[0]
        level = np.empty(n, np.int32); dist = np.empty(n, np.float64); prob = np.empty(n, np.float64); pos = np.empty(n, np.uint8)
        capi.check(self.ctx.lib, self.ctx.lib.fdb_rvm_get_probability(
            self.h, v.ctypes.data, n, level.ctypes.data, dist.ctypes.data, prob.ctypes.data, pos.ctypes.data))
        return level, dist, prob, pos


class SlidingWindowCascade:
    """detection::FiveStageSlidingWindowDetector / SlidingWindowDetector over frame batches."""

    def __init__(self, ctx, det_kwargs, wvm_model, svm_model=None, feature=None, rvm_model=None):
        """wvm_model None + svm_model: the `single` psvm detector (every window through the SVM); rvm_model alone: the
        `single` prvm detector. feature: capi.FeatureDesc - feature space of the SVM (default: the HistEq64 patch)."""
        self.ctx = ctx
        self.wvm = ProbabilisticWvmClassifier(ctx, wvm_model) if wvm_model is not None else None
        self.svm = ProbabilisticSvmClassifier(ctx, svm_model) if svm_model is not None else None
        self.rvm = ProbabilisticRvmClassifier(ctx, rvm_model) if rvm_model is not None else None
        self._desc = detector_desc(**det_kwargs)
        h = C.c_void_p()
        if self.rvm is not None:
            capi.check(ctx.lib, ctx.lib.fdb_detector_create_rvm(ctx.h, C.byref(self._desc), self.rvm.h, C.byref(h)))
        else:
            capi.check(ctx.lib, ctx.lib.fdb_detector_create(ctx.h, C.byref(self._desc), self.wvm.h if self.wvm else None,
                                                            self.svm.h if self.svm else None, C.byref(h)))
        self.h = h
        self.width = self.height = None
        self.patch = (self._desc.patch_width, self._desc.patch_height)
        self.feature = feature
        self.feature_dim, self.feature_dtype = self.patch[0] * self.patch[1], np.uint8
        if feature is not None:
            capi.check(ctx.lib, ctx.lib.fdb_detector_set_feature(self.h, C.byref(feature)))
            self.feature_dim, self.feature_dtype = feature_shape(feature, *self.patch)

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.fdb_detector_destroy(self.h)
        except Exception:
            pass

    def prepare(self, width, height, max_batch):
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_prepare(self.h, width, height, max_batch))
        self.width, self.height, self.max_batch = width, height, max_batch

    @property
    def windows_per_frame(self):
        return int(self.ctx.lib.fdb_detector_windows_per_frame(self.h))

    @property
    def pyramid_bytes(self):
        return int(self.ctx.lib.fdb_detector_pyramid_bytes(self.h))

    def layers(self):
        n = C.c_int32()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_layers(self.h, None, 0, C.byref(n)))
        buf = (capi.LayerInfo * max(n.value, 1))()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_layers(self.h, buf, n.value, C.byref(n)))
        return [{f: getattr(buf[i], f) for f in LAYER_FIELDS} for i in range(n.value)]

    def detect(self, frames, stage=capi.FDB_STAGE_NMS, want_dense=False, det_cap=None):
        """frames [n, H, W] u8 (host) -> detections (structured array) [, dense [n, windows]]"""
        frames = np.ascontiguousarray(frames, np.uint8)
        if frames.ndim == 2:
            frames = frames[None]
        n, H, W = frames.shape
        assert (W, H) == (self.width, self.height), "prepare() was called for another frame size"
        det_cap = det_cap or max(1024, n * 4096)
        dets = np.zeros(det_cap, DETECTION_DTYPE)
        dense = np.zeros((n, self.windows_per_frame), SCORE_DTYPE) if want_dense else None
        cnt = C.c_int64()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detect_batch(
            self.h, frames.ctypes.data, W, n, stage, dense.ctypes.data if want_dense else None,
            dets.ctypes.data, det_cap, C.byref(cnt)))
        out = dets[:cnt.value].copy()
        return (out, dense) if want_dense else out

    def detect_bgr(self, frames_bgr, stage=capi.FDB_STAGE_NMS, want_dense=False, det_cap=None):
        """frames [n, H, W, 3] u8 interleaved BGR (host): GrayscaleFilter's cvtColor branch runs on the device first"""
        frames = np.ascontiguousarray(frames_bgr, np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        n, H, W, ch = frames.shape
        assert ch == 3 and (W, H) == (self.width, self.height)
        det_cap = det_cap or max(1024, n * 4096)
        dets = np.zeros(det_cap, DETECTION_DTYPE)
        dense = np.zeros((n, self.windows_per_frame), SCORE_DTYPE) if want_dense else None
        cnt = C.c_int64()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detect_batch_bgr(
            self.h, frames.ctypes.data, 3 * W, n, stage, dense.ctypes.data if want_dense else None,
            dets.ctypes.data, det_cap, C.byref(cnt)))
        out = dets[:cnt.value].copy()
        return (out, dense) if want_dense else out

    def detect_device(self, frames_ptr, n, stage=capi.FDB_STAGE_NMS, dense_ptr=None, det_cap=None):
        """frames already in device memory ([n, H, W] u8 at frames_ptr); dense_ptr: optional device
        buffer of n * windows_per_frame records"""
        det_cap = det_cap or max(1024, n * 4096)
        dets = np.zeros(det_cap, DETECTION_DTYPE)
        cnt = C.c_int64()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detect_batch_device(
            self.h, frames_ptr, n, stage, dense_ptr, dets.ctypes.data, det_cap, C.byref(cnt)))
        return dets[:cnt.value].copy()

    def enqueue_device(self, frames_ptr, n, dense_ptr=None):
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detect_enqueue_device(self.h, frames_ptr, n, dense_ptr))

    def profile_device(self, frames_ptr, n):
        """(resize ms, pyrDown ms, window-kernel ms, deep-kernel ms, stage-1 ms, launches per kernel) from CUDA
        events between the kernels, summed over the internal chunks"""
        ms = (C.c_double * 6)()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detect_profile_device(self.h, frames_ptr, n, ms))
        return list(ms)

    def detect_roi(self, frame, roi, stage=capi.FDB_STAGE_SVM, det_cap=1 << 16):
        frame = np.ascontiguousarray(frame, np.uint8)
        dets = np.zeros(det_cap, DETECTION_DTYPE)
        cnt = C.c_int64()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detect_roi(
            self.h, frame.ctypes.data, frame.shape[1], roi[0], roi[1], roi[2], roi[3], stage,
            dets.ctypes.data, det_cap, C.byref(cnt)))
        return dets[:cnt.value].copy()

    def extract_patches(self, frame):
        """PyramidFeatureExtractor::extract(1, 1): hq64 patch of every window [windows, w*h]"""
        frame = np.ascontiguousarray(frame, np.uint8)
        nw = self.windows_per_frame
        out = np.zeros((nw, self.patch[0] * self.patch[1]), np.uint8)
        cnt = C.c_int64()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_extract_patches(
            self.h, frame.ctypes.data, frame.shape[1], out.ctypes.data, nw, C.byref(cnt)))
        return out

    def pyramid_layer(self, frame, layer_index):
        frame = np.ascontiguousarray(frame, np.uint8)
        info = [L for L in self.layers() if L["index"] == layer_index]
        if not info:
            raise KeyError(layer_index)
        out = np.zeros((info[0]["height"], info[0]["width"]), np.uint8)
        capi.check(self.ctx.lib, self.ctx.lib.fdb_pyramid_layer(
            self.h, frame.ctypes.data, frame.shape[1], layer_index, out.ctypes.data, out.size))
        return out

    def extract_features(self, frame, layer_x_y):
        """PyramidFeatureExtractor::extract(layer, x, y) in the SVM's feature space: [n, dim]"""
        frame = np.ascontiguousarray(frame, np.uint8)
        lxy = np.ascontiguousarray(layer_x_y, np.int32).reshape(-1, 3)
        out = np.zeros((lxy.shape[0], self.feature_dim), self.feature_dtype)
        capi.check(self.ctx.lib, self.ctx.lib.fdb_extract_features(
            self.h, frame.ctypes.data, frame.shape[1], lxy.ctypes.data, lxy.shape[0], out.ctypes.data))
        return out

    def detect_single(self, frames, want_distances=True, det_cap=None):
        """`single` psvm detector: (detections, distances [n, windows] or None)"""
        frames = np.ascontiguousarray(frames, np.uint8)
        if frames.ndim == 2:
            frames = frames[None]
        n, H, W = frames.shape
        det_cap = det_cap or max(1024, n * self.windows_per_frame)
        dets = np.zeros(det_cap, DETECTION_DTYPE)
        dist = np.zeros((n, self.windows_per_frame), np.float64) if want_distances else None
        cnt = C.c_int64()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detect_single(
            self.h, frames.ctypes.data, W, n, dist.ctypes.data if want_distances else None,
            dets.ctypes.data, det_cap, C.byref(cnt)))
        return dets[:cnt.value].copy(), dist

    def evaluate_samples(self, frame, samples_xywh, max_svm_patches=8):
        """condensation::WvmSvmModel::evaluate(image, samples): samples [n, 4] = centre x, centre y, width, height ->
        (target [n] bool, weight [n] float64)"""
        frame = np.ascontiguousarray(frame, np.uint8)
        smp = np.ascontiguousarray(samples_xywh, np.int32).reshape(-1, 4)
        target = np.zeros(len(smp), np.uint8); weight = np.zeros(len(smp), np.float64)
        capi.check(self.ctx.lib, self.ctx.lib.fdb_evaluate_samples(
            self.h, frame.ctypes.data, frame.shape[1], smp.ctypes.data, len(smp), max_svm_patches, target.ctypes.data, weight.ctypes.data))
        return target.astype(bool), weight

    @property
    def single_dense(self):
        """True when detect_single runs as one tensor-core launch per chunk of frames (csrc/svm_dense.cu)"""
        return bool(self.ctx.lib.fdb_detector_single_dense(self.h))

    def single_dense_profile(self):
        """(svm_dense_kernel milliseconds, launches) of the last detect_single / detect_single_device call"""
        ms, n = C.c_double(), C.c_int32()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_single_dense_profile(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def detect_single_device(self, frames_ptr, n, distance_ptr=None, det_cap=None):
        """detect_single for frames resident in device memory; distances (if wanted) stay on the device"""
        det_cap = det_cap or max(1024, 64 * n)
        dets = np.zeros(det_cap, DETECTION_DTYPE)
        cnt = C.c_int64()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detect_single_device(
            self.h, frames_ptr, n, distance_ptr, dets.ctypes.data, det_cap, C.byref(cnt)))
        return dets[:cnt.value].copy()

    def last_counts(self):
        c = (C.c_int64 * 5)()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_last_counts(self.h, c))
        return list(c)


class DetectorSet:
    """fdb_detector_set: every detector of an application on every frame (ffpDetectApp.cpp:548-596) over shared pyramids.
    Results equal those of the members run one by one; `reserved` of a detection = member index."""

    def __init__(self, ctx, cascades):
        self.ctx, self.members = ctx, list(cascades)
        arr = (C.c_void_p * len(self.members))(*[c.h for c in self.members])
        h = C.c_void_p()
        capi.check(ctx.lib, ctx.lib.fdb_detector_set_create(ctx.h, arr, len(self.members), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.fdb_detector_set_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def prepare(self, width, height, max_batch):
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_set_prepare(self.h, width, height, max_batch))
        self.width, self.height, self.max_batch = width, height, max_batch
        for c in self.members:
            c.width, c.height, c.max_batch = width, height, max_batch

    @property
    def windows_per_frame(self):
        return int(self.ctx.lib.fdb_detector_set_windows_per_frame(self.h))

    def info(self):
        ni, nb, nl, nf = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_set_info(self.h, C.byref(ni), C.byref(nb), C.byref(nl), C.byref(nf)))
        return {"pyramid_images": ni.value, "pyramid_bytes": nb.value, "window_launches": nl.value, "fast_members": nf.value}

    def detect(self, frames, stage=capi.FDB_STAGE_NMS, det_cap=None):
        frames = np.ascontiguousarray(frames, np.uint8)
        if frames.ndim == 2:
            frames = frames[None]
        n, H, W = frames.shape
        assert (W, H) == (self.width, self.height), "prepare() was called for another frame size"
        det_cap = det_cap or max(1024, n * 4096 * len(self.members))
        dets = np.zeros(det_cap, DETECTION_DTYPE)
        cnt = C.c_int64()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_set_detect_batch(
            self.h, frames.ctypes.data, W, n, stage, dets.ctypes.data, det_cap, C.byref(cnt)))
        return dets[:cnt.value].copy()

    def detect_device(self, frames_ptr, n, stage=capi.FDB_STAGE_NMS, dense_ptrs=None, det_cap=None):
        """frames in device memory; dense_ptrs: optional list (one device pointer or None per member)"""
        det_cap = det_cap or max(1024, n * 4096 * len(self.members))
        dets = np.zeros(det_cap, DETECTION_DTYPE)
        cnt = C.c_int64()
        arr = None
        if dense_ptrs is not None:
            arr = (C.c_void_p * len(self.members))(*[p if p else None for p in dense_ptrs])
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_set_detect_batch_device(
            self.h, frames_ptr, n, stage, arr, dets.ctypes.data, det_cap, C.byref(cnt)))
        return dets[:cnt.value].copy()

    def profile_device(self, frames_ptr, n):
        ms = (C.c_double * 6)()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_set_profile_device(self.h, frames_ptr, n, ms))
        return list(ms)

    def last_host_ms(self):
        """host wall clock of the last detect call: enqueue, phase A (wait + overlap elimination + SVM launch), phase B, whole call"""
        ms = (C.c_double * 8)()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_detector_set_last_host_ms(self.h, ms))
        return dict(zip(("enqueue", "phase_a", "phase_b", "call", "wait_stage1", "phase_a_fetch", "phase_a_cpu", "phase_a_launch"), ms))


class AggregatedFeaturesDetector:
    """detection::AggregatedFeaturesDetector (AggregatedFeaturesDetector.cpp:37-118) with a GrayscaleFilter image filter and a
    FhogFilter layer filter: weights [window_rows, window_cols, 3 * unsigned_bins + 4] float32 = the linear SVM's support vector."""

    def __init__(self, ctx, weights, bias, threshold, cell=4, octave_layer_count=5, min_window_width=0, nms_threshold=0.3,
                 nms_type=capi.FDB_NMS_MAX_SCORE, width_scale=1.0, height_scale=1.0, unsigned_bins=9, interpolate_bins=False,
                 interpolate_cells=True, alpha=0.2):
        self.ctx = ctx
        w = np.ascontiguousarray(weights, np.float32)
        assert w.ndim == 3 and w.shape[2] == 3 * unsigned_bins + 4
        d = capi.AggdetDesc()
        d.cell_size, d.window_rows, d.window_cols, d.octave_layer_count = cell, w.shape[0], w.shape[1], octave_layer_count
        d.min_window_width, d.width_scale, d.height_scale = min_window_width, width_scale, height_scale
        d.unsigned_bins, d.interpolate_bins, d.interpolate_cells, d.alpha = unsigned_bins, int(interpolate_bins), int(interpolate_cells), alpha
        d.weights = w.ctypes.data_as(C.POINTER(C.c_float))
        d.bias, d.threshold, d.nms_overlap_threshold, d.nms_type = bias, threshold, nms_threshold, nms_type
        h = C.c_void_p()
        capi.check(ctx.lib, ctx.lib.fdb_aggdet_create(ctx.h, C.byref(d), C.byref(h)))
        self.h, self.dim = h, w.shape[2]

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.fdb_aggdet_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def prepare(self, width, height, max_batch):
        capi.check(self.ctx.lib, self.ctx.lib.fdb_aggdet_prepare(self.h, width, height, max_batch))
        self.width, self.height = width, height

    def layers(self):
        n = C.c_int32()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_aggdet_layers(self.h, C.byref(n), None, 0))
        info = np.zeros((max(n.value, 1), 6), np.int32)
        capi.check(self.ctx.lib, self.ctx.lib.fdb_aggdet_layers(self.h, C.byref(n), info.ctypes.data, n.value))
        return [dict(zip(("index", "width", "height", "cells_x", "cells_y", "positions"), map(int, r))) for r in info[:n.value]]

    @property
    def positions_per_frame(self):
        return int(self.ctx.lib.fdb_aggdet_positions_per_frame(self.h))

    def detect(self, frames, cap=None):
        """frames [n, H, W] u8 -> (rects [k, 4] x y w h, scores [k], frame [k])"""
        frames = np.ascontiguousarray(frames, np.uint8)
        if frames.ndim == 2:
            frames = frames[None]
        n, H, W = frames.shape
        assert (W, H) == (self.width, self.height)
        cap = cap or max(4096, 1024 * n)
        scores, rects, fr = np.zeros(cap, np.float32), np.zeros((cap, 4), np.int32), np.zeros(cap, np.int32)
        cnt = C.c_int64()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_aggdet_detect_batch(self.h, frames.ctypes.data, W, n, scores.ctypes.data, rects.ctypes.data,
                                                                      fr.ctypes.data, cap, C.byref(cnt)))
        k = cnt.value
        return rects[:k].copy(), scores[:k].copy(), fr[:k].copy()

    def detect_device(self, frames_ptr, n, cap=None):
        cap = cap or max(4096, 1024 * n)
        scores, rects, fr = np.zeros(cap, np.float32), np.zeros((cap, 4), np.int32), np.zeros(cap, np.int32)
        cnt = C.c_int64()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_aggdet_detect_batch_device(self.h, frames_ptr, n, scores.ctypes.data, rects.ctypes.data,
                                                                             fr.ctypes.data, cap, C.byref(cnt)))
        k = cnt.value
        return rects[:k].copy(), scores[:k].copy(), fr[:k].copy()

    def score_maps(self, frame):
        """per layer: (feature map [cells_y, cells_x, D], score map [valid rows, valid cols])"""
        frame = np.ascontiguousarray(frame, np.uint8)
        L = self.layers()
        npos, ncell = sum(l["positions"] for l in L), sum(l["cells_x"] * l["cells_y"] for l in L)
        sc, ft = np.zeros(max(npos, 1), np.float32), np.zeros(max(ncell * self.dim, 1), np.float32)
        capi.check(self.ctx.lib, self.ctx.lib.fdb_aggdet_score_maps(self.h, frame.ctypes.data, frame.shape[1], sc.ctypes.data, sc.size,
                                                                    ft.ctypes.data, ft.size))
        out, so, fo_ = [], 0, 0
        for l in L:
            nc = l["cells_x"] * l["cells_y"]
            feat = ft[fo_:fo_ + nc * self.dim].reshape(l["cells_y"], l["cells_x"], self.dim)
            fo_ += nc * self.dim
            out.append((l, feat, sc[so:so + l["positions"]]))
            so += l["positions"]
        return out

    def profile_device(self, frames_ptr, n):
        ms = (C.c_double * 6)()
        capi.check(self.ctx.lib, self.ctx.lib.fdb_aggdet_profile_device(self.h, frames_ptr, n, ms))
        return dict(zip(("pyramid", "histograms", "descriptors", "score_maps", "total", "chunks"), ms))


def detect_face_features(face, features, frame, cap=4096):
    """ffpDetectApp.cpp:553-596: face detector on the frame, then each feature detector inside the first face's bounds.
    face / features: prepared SlidingWindowCascade objects. -> (face detections, [feature detections per detector])"""
    frame = np.ascontiguousarray(frame, np.uint8)
    lib = face.ctx.lib
    fdets = np.zeros(cap, DETECTION_DTYPE)
    out = np.zeros((max(len(features), 1), cap), DETECTION_DTYPE)
    nf = C.c_int64()
    counts = (C.c_int64 * max(len(features), 1))()
    handles = (C.c_void_p * max(len(features), 1))(*[f.h for f in features])
    capi.check(lib, lib.fdb_detect_face_features(face.h, handles, len(features), frame.ctypes.data, frame.shape[1], fdets.ctypes.data, cap,
                                                 C.byref(nf), out.ctypes.data, cap, counts))
    return fdets[:nf.value].copy(), [out[i, :counts[i]].copy() for i in range(len(features))]


class SdmLandmarkModel:
    """fdb_sdm: superviseddescent::SdmLandmarkModel + SdmLandmarkModelFitting (SdmLandmarkModel.hpp:44-256) for batches of
    faces. `model` is a featuredetection_b200.synthetic.SdmModel; `path` a reference text model (SdmLandmarkModel::load)."""

    def __init__(self, ctx, model=None, path=None):
        self.ctx, self.lib = ctx, ctx.lib
        self.h = None
        h = C.c_void_p()
        if path is not None:
            f = C.c_void_p()
            capi.check(self.lib, self.lib.fdb_sdm_file_load(str(path).encode(), C.byref(f)))
            try:
                capi.check(self.lib, self.lib.fdb_sdm_create(ctx.h, self.lib.fdb_sdm_file_desc(f), C.byref(h)))
            finally:
                self.lib.fdb_sdm_file_free(f)
        else:
            regs = (C.c_void_p * len(model.regressors))(*[r.ctypes.data for r in model.regressors])
            desc = capi.SdmDesc(model.num_landmarks, len(model.regressors), model.mean.ctypes.data_as(C.POINTER(C.c_float)), regs)
            capi.check(self.lib, self.lib.fdb_sdm_create(ctx.h, C.byref(desc), C.byref(h)))
        self.h = h
        self.num_landmarks = int(self.lib.fdb_sdm_num_landmarks(h))
        self.num_cascade_steps = int(self.lib.fdb_sdm_num_cascade_steps(h))

    def __del__(self):
        try:
            if self.h:
                self.lib.fdb_sdm_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def align_rigid(self, boxes_xywh):
        """alignRigid(mean, faceBox) for [n][4] integer boxes -> [n][2L] float32 start shapes"""
        boxes = np.ascontiguousarray(boxes_xywh, np.int32).reshape(-1, 4)
        out = np.empty((boxes.shape[0], 2 * self.num_landmarks), np.float32)
        capi.check(self.lib, self.lib.fdb_sdm_align_rigid(self.h, boxes.ctypes.data, boxes.shape[0], out.ctypes.data))
        return out

    def optimize(self, frames, shapes, face_frame=None, want_features=False):
        """optimize(modelShape, image) for every face: frames [n][H][W] u8, shapes [faces][2L] -> (shapes, status[, features])"""
        frames = np.ascontiguousarray(frames, np.uint8)
        if frames.ndim == 2:
            frames = frames[None]
        shapes = np.ascontiguousarray(shapes, np.float32).reshape(-1, 2 * self.num_landmarks).copy()
        n_faces = shapes.shape[0]
        ff = None if face_frame is None else np.ascontiguousarray(face_frame, np.int32)
        status = np.zeros(n_faces, np.int32)
        feats = np.zeros((self.num_cascade_steps, n_faces, 279 * self.num_landmarks), np.float32) if want_features else None
        capi.check(self.lib, self.lib.fdb_sdm_optimize_batch(
            self.h, frames.ctypes.data, frames.shape[2], frames.shape[2], frames.shape[1], frames.shape[0],
            None if ff is None else ff.ctypes.data, n_faces, shapes.ctypes.data, status.ctypes.data,
            None if feats is None else feats.ctypes.data))
        return (shapes, status, feats) if want_features else (shapes, status)

    def optimize_device(self, frames_ptr, width, height, n_frames, face_frame_ptr, n_faces, shapes_ptr, status_ptr=None):
        capi.check(self.lib, self.lib.fdb_sdm_optimize_batch_device(self.h, frames_ptr, width, height, n_frames, face_frame_ptr, n_faces,
                                                                    shapes_ptr, status_ptr))

    def profile_device(self, frames_ptr, width, height, n_frames, face_frame_ptr, n_faces, shapes_ptr, status_ptr=None):
        ms = (C.c_double * 4)()
        capi.check(self.lib, self.lib.fdb_sdm_profile_device(self.h, frames_ptr, width, height, n_frames, face_frame_ptr, n_faces,
                                                             shapes_ptr, status_ptr, ms))
        return dict(hog=ms[0], gemm=ms[1], update=ms[2], total=ms[3])

    def descriptors(self, frame, points_xy, window_size_half):
        """VlHogDescriptorExtractor::getDescriptors -> [n][279] float32"""
        frame = np.ascontiguousarray(frame, np.uint8)
        pts = np.ascontiguousarray(points_xy, np.float32).reshape(-1, 2)
        out = np.zeros((pts.shape[0], 279), np.float32)
        capi.check(self.lib, self.lib.fdb_sdm_descriptors(self.h, frame.ctypes.data, frame.shape[1], frame.shape[1], frame.shape[0],
                                                          pts.ctypes.data, pts.shape[0], window_size_half, out.ctypes.data))
        return out


def load_wvm_mat(classifier_path, thresholds_path):
    """WvmClassifier::loadFromMatlab + the posterior_wrvm logistic through the library's MAT-file reader (fdb_wvm_file_load, host
    only) -> synthetic.WvmModel holding copies of the descriptor arrays (evaluator units)."""
    from .synthetic import WvmModel
    lib = capi.load_library()
    f = C.c_void_p()
    capi.check(lib, lib.fdb_wvm_file_load(str(classifier_path).encode(), str(thresholds_path).encode(), C.byref(f)))
    try:
        d = lib.fdb_wvm_file_desc(f).contents
        n = d.num_lin_filters
        arr = lambda p, k, t: np.ctypeslib.as_array(p, shape=(k,)).astype(t, copy=True) if k else np.zeros(0, t)
        cntval = arr(d.area_cntval, n, np.int32)
        nv = int(cntval.sum())
        cntrec = arr(d.area_cntrec, nv, np.int32)
        first = np.concatenate([[0], np.cumsum(cntval)[:-1]]).astype(np.int64)
        nrec = int(cntrec.sum() - cntrec[first].sum())
        rec = np.ctypeslib.as_array(C.cast(d.area_rec, C.POINTER(C.c_int32)), shape=(nrec, 4)).copy() if nrec else np.zeros((0, 4), np.int32)
        return WvmModel(d.filter_size_x, d.filter_size_y, d.num_filters_per_level, d.num_levels, d.basis_param,
                        arr(d.lin_thresholds, n, np.float32), arr(d.hk_weights, n * (n + 1) // 2, np.float32),
                        arr(d.app_rsv_convol, n, np.float64), arr(d.hierarchical_thresholds, n, np.float32), cntval,
                        arr(d.area_val, nv, np.float64), cntrec, rec, limit_reliability_filter=d.limit_reliability_filter,
                        num_used=d.num_used_filters, logistic_a=d.logistic_a, logistic_b=d.logistic_b)
    finally:
        lib.fdb_wvm_file_free(f)


def load_svm_mat(classifier_path, logistic_path=None):
    """SvmClassifier::loadFromMatlab + posterior_svm (fdb_svm_mat_load, host only) -> synthetic.SvmModel"""
    from .synthetic import SvmModel
    lib = capi.load_library()
    f = C.c_void_p()
    capi.check(lib, lib.fdb_svm_mat_load(str(classifier_path).encode(), None if logistic_path is None else str(logistic_path).encode(), C.byref(f)))
    try:
        d = lib.fdb_svm_file_desc(f).contents
        sv = np.ctypeslib.as_array(C.cast(d.support_vectors, C.POINTER(C.c_uint8)), shape=(d.num_sv, d.dim)).copy()
        coef = np.ctypeslib.as_array(d.coefficients, shape=(d.num_sv,)).copy()
        kernel = {capi.FDB_KERNEL_RBF: "rbf", capi.FDB_KERNEL_POLYNOMIAL: "polynomial", capi.FDB_KERNEL_HIK: "hik",
                  capi.FDB_KERNEL_LINEAR: "linear"}[d.kernel]
        return SvmModel(sv, coef, d.gamma, bias=d.bias, threshold=d.threshold, logistic_a=d.logistic_a, logistic_b=d.logistic_b,
                        kernel=kernel, alpha=d.poly_alpha, constant=d.poly_constant, degree=d.poly_degree)
    finally:
        lib.fdb_svm_file_free(f)
