"""ctypes front-end of the CPU oracle (libfdoracle.so) and of the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py. The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from featuredetection_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "libfdoracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libfdref.so")


def build(force=False):
    """make -C oracle (compiles fd_oracle.c and, when /root/reference exists, oracle/_ref)."""
    if force or not os.path.exists(ORACLE_LIB) or (os.path.isdir("/root/reference") and not os.path.exists(REF_LIB)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


class _Pyramid(C.Structure):
    _fields_ = [
        ("octave_layer_count", C.c_int), ("incremental_scale_factor", C.c_double),
        ("min_scale_factor", C.c_double), ("max_scale_factor", C.c_double),
        ("image_width", C.c_int), ("image_height", C.c_int),
        ("n_layers", C.c_int), ("layers", C.c_void_p),
        ("n_resize", C.c_int), ("n_pyrdown", C.c_int),
        ("px_resize", C.c_int64), ("px_pyrdown", C.c_int64),
    ]


class _Layer(C.Structure):
    _fields_ = [("index", C.c_int), ("scale", C.c_double), ("width", C.c_int), ("height", C.c_int),
                ("data", C.POINTER(C.c_uint8))]


_lib = None


def _bind_sdm(L):
    """fd_sdm.c entry points (present in libfdoracle.so and, with the reference's own hog.c inside, in oracle/_ref)"""
    L.fdo_resize_linear_f32.restype = None
    L.fdo_resize_linear_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.fdo_vlhog_uoctti.restype = None
    L.fdo_vlhog_uoctti.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fdo_sdm_descriptors.restype = C.c_int
    L.fdo_sdm_descriptors.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.fdo_sdm_create.restype = C.c_void_p
    L.fdo_sdm_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.fdo_sdm_load.restype = C.c_void_p; L.fdo_sdm_load.argtypes = [C.c_char_p]
    L.fdo_sdm_free.restype = None; L.fdo_sdm_free.argtypes = [C.c_void_p]
    L.fdo_sdm_num_landmarks.restype = C.c_int; L.fdo_sdm_num_landmarks.argtypes = [C.c_void_p]
    L.fdo_sdm_num_steps.restype = C.c_int; L.fdo_sdm_num_steps.argtypes = [C.c_void_p]
    L.fdo_sdm_mean.restype = C.POINTER(C.c_float); L.fdo_sdm_mean.argtypes = [C.c_void_p]
    L.fdo_sdm_regressor.restype = C.POINTER(C.c_float); L.fdo_sdm_regressor.argtypes = [C.c_void_p, C.c_int]
    L.fdo_sdm_align_rigid.restype = None
    L.fdo_sdm_align_rigid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.fdo_sdm_window.restype = None
    L.fdo_sdm_window.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.fdo_sdm_optimize.restype = C.c_int
    L.fdo_sdm_optimize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_LIB)
        L.fdo_cvround.restype = C.c_int; L.fdo_cvround.argtypes = [C.c_double]
        L.fdo_resize_linear_u8.restype = None
        L.fdo_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.fdo_pyrdown_u8.restype = None
        L.fdo_pyrdown_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.fdo_pyramid_build.restype = C.POINTER(_Pyramid)
        L.fdo_pyramid_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.fdo_pyramid_free.restype = None; L.fdo_pyramid_free.argtypes = [C.POINTER(_Pyramid)]
        L.fdo_hq64.restype = None; L.fdo_hq64.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.fdo_wvm_create.restype = C.c_void_p; L.fdo_wvm_create.argtypes = [C.POINTER(capi.WvmDesc)]
        L.fdo_wvm_free.restype = None; L.fdo_wvm_free.argtypes = [C.c_void_p]
        L.fdo_wvm_eval.restype = None
        L.fdo_wvm_eval.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.fdo_wvm_eval_all_levels.restype = None
        L.fdo_wvm_eval_all_levels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fdo_wvm_classify.restype = C.c_int; L.fdo_wvm_classify.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.fdo_wvm_probability.restype = C.c_double; L.fdo_wvm_probability.argtypes = [C.c_void_p, C.c_float]
        L.fdo_svm_create.restype = C.c_void_p; L.fdo_svm_create.argtypes = [C.POINTER(capi.SvmDesc)]
        L.fdo_svm_free.restype = None; L.fdo_svm_free.argtypes = [C.c_void_p]
        L.fdo_svm_distance.restype = C.c_double; L.fdo_svm_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.fdo_svm_classify.restype = C.c_int; L.fdo_svm_classify.argtypes = [C.c_void_p, C.c_double]
        L.fdo_svm_probability.restype = C.c_double; L.fdo_svm_probability.argtypes = [C.c_void_p, C.c_double]
        L.fdo_enumerate.restype = C.c_int64
        L.fdo_enumerate.argtypes = [C.POINTER(_Pyramid), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.POINTER(capi.LayerInfo), C.c_int]
        L.fdo_detect_frame.restype = C.c_int64
        L.fdo_detect_frame.argtypes = [C.POINTER(capi.DetectorDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.fdo_gradient_u8.restype = None
        L.fdo_gradient_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.fdo_gradient_bin_luts.restype = None
        L.fdo_gradient_bin_luts.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.fdo_lbp_u8.restype = None
        L.fdo_lbp_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.fdo_equalize_hist_u8.restype = None
        L.fdo_equalize_hist_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.fdo_whitening_filter.restype = None
        L.fdo_whitening_filter.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.fdo_whitening_u8.restype = None
        L.fdo_whitening_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fdo_features_create.restype = C.c_void_p
        L.fdo_features_create.argtypes = [C.POINTER(capi.FeatureDesc), C.c_int, C.c_int]
        L.fdo_features_free.restype = None; L.fdo_features_free.argtypes = [C.c_void_p]
        for nm in ("fdo_features_dim", "fdo_features_is_float", "fdo_features_layer_channels"):
            getattr(L, nm).restype = C.c_int; getattr(L, nm).argtypes = [C.c_void_p]
        L.fdo_features_filter_layer.restype = None
        L.fdo_features_filter_layer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.fdo_features_patch.restype = None
        L.fdo_features_patch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.fdo_extract_features.restype = C.c_int
        L.fdo_extract_features.argtypes = [C.POINTER(capi.DetectorDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_int64, C.c_void_p]
        L.fdo_detect_frame_ex.restype = C.c_int64
        L.fdo_detect_frame_ex.argtypes = [C.POINTER(capi.DetectorDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        _bind_sdm(L)
        _lib = L
    return _lib


_ref = None


def ref_available():
    return os.path.exists(REF_LIB)


def ref():
    """The compiled, unmodified reference sources (oracle/_ref/libfdref.so)."""
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(REF_LIB)
        R.ref_hq64.restype = None; R.ref_hq64.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        R.ref_wvm_create.restype = C.c_void_p; R.ref_wvm_create.argtypes = [C.POINTER(capi.WvmDesc)]
        R.ref_wvm_free.restype = None; R.ref_wvm_free.argtypes = [C.c_void_p]
        R.ref_wvm_eval.restype = None
        R.ref_wvm_eval.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                   C.POINTER(C.c_double), C.POINTER(C.c_int)]
        R.ref_svm_create.restype = C.c_void_p; R.ref_svm_create.argtypes = [C.POINTER(capi.SvmDesc)]
        R.ref_svm_free.restype = None; R.ref_svm_free.argtypes = [C.c_void_p]
        R.ref_svm_eval.restype = None
        R.ref_svm_eval.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        R.ref_overlap_eliminate.restype = C.c_int
        R.ref_overlap_eliminate.argtypes = [C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]
        R.ref_svm_store.restype = C.c_int; R.ref_svm_store.argtypes = [C.c_void_p, C.c_char_p]
        R.ref_detect_frame.restype = C.c_int64
        R.ref_detect_frame.argtypes = [C.POINTER(capi.DetectorDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.POINTER(C.c_double)]
        R.ref_detect_frame_ex.restype = C.c_int64
        R.ref_detect_frame_ex.argtypes = [C.POINTER(capi.DetectorDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.POINTER(C.c_double)]
        R.ref_detect_layers_ex.restype = C.c_int64
        R.ref_detect_layers_ex.argtypes = [C.POINTER(capi.DetectorDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.POINTER(C.c_double)]
        _bind_sdm(R)
        R.ref_vlhog_uoctti.restype = C.c_int
        R.ref_vlhog_uoctti.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        R.ref_gradient_bin_luts.restype = None
        R.ref_gradient_bin_luts.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        R.ref_lbp.restype = None; R.ref_lbp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        R.ref_patch_histogram.restype = C.c_int
        R.ref_patch_histogram.argtypes = [C.POINTER(capi.FeatureDesc), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_int]
        _ref = R
    return _ref


def ref_gradient_bin_luts(bins, signed):
    one = np.empty((65536, 2), np.uint8); two = np.empty((65536, 4), np.uint8)
    ref().ref_gradient_bin_luts(bins, int(signed), one.ctypes.data, two.ctypes.data)
    return one, two


def ref_lbp(gray, lbp_type):
    gray = np.ascontiguousarray(gray, np.uint8)
    out = np.empty(gray.shape, np.uint8)
    ref().ref_lbp(gray.ctypes.data, gray.shape[1], gray.shape[0], lbp_type, out.ctypes.data)
    return out


def ref_patch_histogram(feature_desc, bins, filtered_layer, x, y, pw, ph):
    """The reference's own SpatialHistogramFilter / HogFilter / ExtendedHogFilter on the window (x, y, pw, ph) of a
    binned layer [H, W, channels] u8."""
    fl = np.ascontiguousarray(filtered_layer, np.uint8)
    if fl.ndim == 2:
        fl = fl[:, :, None]
    ch = fl.shape[2]
    roi = fl[y:y + ph, x:x + pw]
    out = np.empty(1 << 16, np.float32)
    n = ref().ref_patch_histogram(C.byref(feature_desc), bins, roi.ctypes.data, fl.strides[0], ph, pw, ch, out.ctypes.data, out.size)
    if n < 0:
        raise RuntimeError("ref_patch_histogram: buffer too small")
    return out[:n].copy()


# ------------------------------------------------------------------------------------------------
# numpy-level helpers
# ------------------------------------------------------------------------------------------------
def resize_linear(img, dw, dh):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty((dh, dw), np.uint8)
    lib().fdo_resize_linear_u8(img.ctypes.data, img.shape[1], img.shape[0], img.shape[1], out.ctypes.data, dw, dh)
    return out


def pyrdown(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(((img.shape[0] + 1) // 2, (img.shape[1] + 1) // 2), np.uint8)
    lib().fdo_pyrdown_u8(img.ctypes.data, img.shape[1], img.shape[0], out.ctypes.data)
    return out


def hq64(patch, use_ref=False):
    """HistEq64 of a 2-D u8 array (may be a strided view with contiguous rows)."""
    assert patch.dtype == np.uint8 and patch.strides[1] == 1
    h, w = patch.shape
    out = np.empty((h, w), np.uint8)
    fn = ref().ref_hq64 if use_ref else lib().fdo_hq64
    fn(patch.ctypes.data, patch.strides[0], w, h, out.ctypes.data)
    return out


def pyramid(frame, inc, mn, mx):
    """Returns (octave_layer_count, [(index, scale, image), ...])."""
    frame = np.ascontiguousarray(frame, np.uint8)
    p = lib().fdo_pyramid_build(frame.ctypes.data, frame.shape[1], frame.shape[0], frame.shape[1], inc, mn, mx)
    if not p:
        raise ValueError("invalid pyramid parameters")
    try:
        layers = C.cast(p.contents.layers, C.POINTER(_Layer))
        out = []
        for i in range(p.contents.n_layers):
            L = layers[i]
            img = np.ctypeslib.as_array(L.data, shape=(L.height, L.width)).copy()
            out.append((L.index, L.scale, img))
        return p.contents.octave_layer_count, out
    finally:
        lib().fdo_pyramid_free(p)


def cv2_pyramid(frame, inc, mn, mx):
    """ImagePyramid::createLayers (ImagePyramid.cpp:170-198) with OpenCV's own cv::resize / cv::pyrDown (python cv2), in the
    reference's loop order: for every octave layer i a resize of the frame, then pyrDown while the scale stays >= the minimum;
    layers inside [min, max] are kept and sorted by index. Returns (octave_layer_count, [(index, scale, image), ...]) - the
    same values as pyramid() (tests/test_oracle_golden.py pins the restatement to cv2 bit for bit). Used by the CPU
    reference arm of bench.py so that the reference is timed with the SIMD pyramid it really links."""
    import cv2
    import math
    frame = np.ascontiguousarray(frame, np.uint8)
    H, W = frame.shape
    olc = int(round(math.log(0.5) / math.log(inc)))
    incr = math.pow(0.5, 1.0 / olc)
    L = lib()
    out = []
    for i in range(olc):
        sf = math.pow(incr, i)
        w, h = L.fdo_cvround(W * sf), L.fdo_cvround(H * sf)
        scaled = cv2.resize(frame, (w, h), interpolation=cv2.INTER_LINEAR)
        if mn <= sf <= mx:
            out.append((i, sf, scaled))
        prev, sf, j = scaled, sf * 0.5, 1
        while sf >= mn and prev.shape[1] > 1:
            prev = cv2.pyrDown(prev)
            if sf <= mx:
                out.append((i + j * olc, sf, prev))
            sf *= 0.5
            j += 1
    out.sort(key=lambda t: t[0])
    return olc, out


class Wvm:
    def __init__(self, model, use_ref=False):
        self.model = model
        self.use_ref = use_ref
        self._desc = model.desc()
        self.h = (ref().ref_wvm_create if use_ref else lib().fdo_wvm_create)(C.byref(self._desc))

    def __del__(self):
        try:
            (ref().ref_wvm_free if self.use_ref else lib().fdo_wvm_free)(self.h)
        except Exception:
            pass

    def eval(self, patches):
        """patches: [n, h*w] u8 -> (level int32[n], fout float32[n], prob float64[n], positive uint8[n])"""
        patches = np.ascontiguousarray(patches, np.uint8).reshape(-1, self.model.w * self.model.h)
        n = patches.shape[0]
        level = np.empty(n, np.int32); fout = np.empty(n, np.float32)
        prob = np.empty(n, np.float64); pos = np.empty(n, np.uint8)
        lv, fo, pr, po = C.c_int(), C.c_float(), C.c_double(), C.c_int()
        if self.use_ref:
            R = ref()
            for i in range(n):
                R.ref_wvm_eval(self.h, patches[i].ctypes.data, C.byref(lv), C.byref(fo), C.byref(pr), C.byref(po))
                level[i], fout[i], prob[i], pos[i] = lv.value, fo.value, pr.value, po.value
        else:
            L = lib()
            for i in range(n):
                L.fdo_wvm_eval(self.h, patches[i].ctypes.data, C.byref(lv), C.byref(fo))
                level[i], fout[i] = lv.value, fo.value
                prob[i] = L.fdo_wvm_probability(self.h, fo)
                pos[i] = L.fdo_wvm_classify(self.h, lv, fo)
        return level, fout, prob, pos

    def eval_all_levels(self, patches):
        assert not self.use_ref
        patches = np.ascontiguousarray(patches, np.uint8).reshape(-1, self.model.w * self.model.h)
        out = np.empty((patches.shape[0], self.model.n), np.float32)
        L = lib()
        for i in range(patches.shape[0]):
            L.fdo_wvm_eval_all_levels(self.h, patches[i].ctypes.data, out[i].ctypes.data)
        return out


class Svm:
    def __init__(self, model, use_ref=False):
        self.model = model
        self.use_ref = use_ref
        self._desc = model.desc()
        self.h = (ref().ref_svm_create if use_ref else lib().fdo_svm_create)(C.byref(self._desc))

    def __del__(self):
        try:
            (ref().ref_svm_free if self.use_ref else lib().fdo_svm_free)(self.h)
        except Exception:
            pass

    def eval(self, vectors):
        """vectors [n, dim] of the SV dtype -> (distance f64[n], probability f64[n], positive u8[n])"""
        vectors = np.ascontiguousarray(vectors, self.model.sv.dtype).reshape(-1, self.model.sv.shape[1])
        n = vectors.shape[0]
        dist = np.empty(n, np.float64); prob = np.empty(n, np.float64); pos = np.empty(n, np.uint8)
        d, p, q = C.c_double(), C.c_double(), C.c_int()
        for i in range(n):
            if self.use_ref:
                ref().ref_svm_eval(self.h, vectors[i].ctypes.data, C.byref(d), C.byref(p), C.byref(q))
                dist[i], prob[i], pos[i] = d.value, p.value, q.value
            else:
                L = lib()
                dist[i] = L.fdo_svm_distance(self.h, vectors[i].ctypes.data)
                prob[i] = L.fdo_svm_probability(self.h, dist[i])
                pos[i] = L.fdo_svm_classify(self.h, dist[i])
        return dist, prob, pos


class Rvm:
    """RvmClassifier + ProbabilisticRvmClassifier: the C restatement (fdo_rvm_*) or - use_ref - the reference's own classes"""

    def __init__(self, model, use_ref=False):
        self.model, self.use_ref = model, use_ref
        self._desc = model.desc()
        if use_ref:
            R = ref()
            R.ref_rvm_create.restype = C.c_void_p; R.ref_rvm_create.argtypes = [C.POINTER(capi.RvmDesc)]
            R.ref_rvm_free.restype = None; R.ref_rvm_free.argtypes = [C.c_void_p]
            R.ref_rvm_eval.restype = C.c_int
            R.ref_rvm_eval.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
            self.h = R.ref_rvm_create(C.byref(self._desc))
        else:
            L = lib()
            L.fdo_rvm_create.restype = C.c_void_p; L.fdo_rvm_create.argtypes = [C.POINTER(capi.RvmDesc)]
            L.fdo_rvm_free.restype = None; L.fdo_rvm_free.argtypes = [C.c_void_p]
            L.fdo_rvm_eval.restype = C.c_int; L.fdo_rvm_eval.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
            L.fdo_rvm_classify.restype = C.c_int; L.fdo_rvm_classify.argtypes = [C.c_void_p, C.c_int, C.c_double]
            L.fdo_rvm_probability.restype = C.c_double; L.fdo_rvm_probability.argtypes = [C.c_void_p, C.c_double]
            self.h = L.fdo_rvm_create(C.byref(self._desc))

    def __del__(self):
        try:
            (ref().ref_rvm_free if self.use_ref else lib().fdo_rvm_free)(self.h)
        except Exception:
            pass

    def eval(self, vectors):
        """vectors [n, dim] -> (level i32[n], distance f64[n], probability f64[n], positive u8[n])"""
        vectors = np.ascontiguousarray(vectors, self.model.sv.dtype).reshape(-1, self.model.sv.shape[1])
        n = vectors.shape[0]
        level = np.empty(n, np.int32); dist = np.empty(n, np.float64); prob = np.empty(n, np.float64); pos = np.empty(n, np.uint8)
        d, p, q = C.c_double(), C.c_double(), C.c_int()
        for i in range(n):
            if self.use_ref:
                level[i] = ref().ref_rvm_eval(self.h, vectors[i].ctypes.data, C.byref(d), C.byref(p), C.byref(q))
                dist[i], prob[i], pos[i] = d.value, p.value, q.value
            else:
                L = lib()
                level[i] = L.fdo_rvm_eval(self.h, vectors[i].ctypes.data, C.byref(d))
                dist[i] = d.value
                prob[i] = L.fdo_rvm_probability(self.h, d.value)
                pos[i] = L.fdo_rvm_classify(self.h, int(level[i]), d.value)
        return level, dist, prob, pos


class Features:
    """A prepared feature space (fdo_features): layer filters + patch filter chain."""

    def __init__(self, feature_desc, patch_w, patch_h):
        self._desc = feature_desc
        self.pw, self.ph = patch_w, patch_h
        self.h = lib().fdo_features_create(C.byref(feature_desc), patch_w, patch_h)
        if not self.h:
            raise ValueError("unsupported feature descriptor")
        self.dim = lib().fdo_features_dim(self.h)
        self.is_float = bool(lib().fdo_features_is_float(self.h))
        self.layer_channels = lib().fdo_features_layer_channels(self.h)
        self.dtype = np.float32 if self.is_float else np.uint8

    def __del__(self):
        try:
            lib().fdo_features_free(self.h)
        except Exception:
            pass

    def filter_layer(self, gray):
        gray = np.ascontiguousarray(gray, np.uint8)
        h, w = gray.shape
        ch = max(self.layer_channels, 1)
        out = np.empty((h, w, ch), np.uint8)
        lib().fdo_features_filter_layer(self.h, gray.ctypes.data, w, h, out.ctypes.data)
        return out

    def patch(self, filtered_layer, x, y):
        fl = np.ascontiguousarray(filtered_layer, np.uint8)
        out = np.empty(self.dim, self.dtype)
        lib().fdo_features_patch(self.h, fl.ctypes.data, fl.shape[1], x, y, out.ctypes.data)
        return out

    def extract(self, det_kwargs, frame, layer_x_y):
        """feature vectors of the windows [(layer index, x, y), ...] of one frame -> [n, dim]"""
        from featuredetection_b200.synthetic import detector_desc
        desc = detector_desc(**det_kwargs)
        frame = np.ascontiguousarray(frame, np.uint8)
        lxy = np.ascontiguousarray(layer_x_y, np.int32).reshape(-1, 3)
        out = np.empty((lxy.shape[0], self.dim), self.dtype)
        rc = lib().fdo_extract_features(C.byref(desc), self.h, frame.ctypes.data, frame.shape[1], frame.shape[0],
                                        frame.shape[1], lxy.ctypes.data, lxy.shape[0], out.ctypes.data)
        if rc != 0:
            raise RuntimeError("fdo_extract_features failed (%d)" % rc)
        return out


def gradient(gray, ksize=1):
    gray = np.ascontiguousarray(gray, np.uint8)
    out = np.empty(gray.shape + (2,), np.uint8)
    lib().fdo_gradient_u8(gray.ctypes.data, gray.shape[1], gray.shape[0], gray.shape[1], ksize, out.ctypes.data)
    return out


def gradient_bin_luts(bins, signed):
    one = np.empty((65536, 2), np.uint8); two = np.empty((65536, 4), np.uint8)
    lib().fdo_gradient_bin_luts(bins, int(signed), one.ctypes.data, two.ctypes.data)
    return one, two


def lbp(gray, lbp_type):
    gray = np.ascontiguousarray(gray, np.uint8)
    out = np.empty(gray.shape, np.uint8)
    lib().fdo_lbp_u8(gray.ctypes.data, gray.shape[1], gray.shape[0], gray.shape[1], lbp_type, out.ctypes.data)
    return out


def equalize_hist(patch):
    assert patch.dtype == np.uint8 and patch.strides[1] == 1
    h, w = patch.shape
    out = np.empty((h, w), np.uint8)
    lib().fdo_equalize_hist_u8(patch.ctypes.data, patch.strides[0], w, h, out.ctypes.data)
    return out


def whitening_filter(w, h, alpha=1.0, cutoff=0.390625):
    out = np.empty((h, w), np.float32)
    lib().fdo_whitening_filter(w, h, alpha, cutoff, out.ctypes.data)
    return out


def whitening(patch, alpha=1.0, cutoff=0.390625):
    """-> (u8 whitened patch, float32 image before convertTo)"""
    assert patch.dtype == np.uint8 and patch.strides[1] == 1
    h, w = patch.shape
    filt = whitening_filter(w, h, alpha, cutoff)
    out = np.empty((h, w), np.uint8); real = np.empty((h, w), np.float32)
    lib().fdo_whitening_u8(patch.ctypes.data, patch.strides[0], w, h, filt.ctypes.data, out.ctypes.data, real.ctypes.data)
    return out, real


def bgr_to_gray(bgr):
    """GrayscaleFilter::applyTo on one [H, W, 3] BGR frame (OpenCV 2.4.3 cvtColor arithmetic)"""
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    out = np.empty((h, w), np.uint8)
    L = lib()
    L.fdo_bgr_to_gray.restype = None
    L.fdo_bgr_to_gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.fdo_bgr_to_gray(bgr.ctypes.data, w, h, 3 * w, out.ctypes.data)
    return out


def resize_linear_f32(img, dw, dh):
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty((dh, dw), np.float32)
    lib().fdo_resize_linear_f32(img.ctypes.data, img.shape[1], img.shape[0], out.ctypes.data, dw, dh)
    return out


def vlhog_uoctti(img, cell_size=10, num_orientations=9, use_ref=False):
    """VLFeat HOG (UoCTTI variant) of a float32 image -> [3 no + 4, hogHeight, hogWidth]"""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    hw, hh = (w + cell_size // 2) // cell_size, (h + cell_size // 2) // cell_size
    out = np.zeros((3 * num_orientations + 4, hh, hw), np.float32)
    a, b = C.c_int(), C.c_int()
    fn = ref().ref_vlhog_uoctti if use_ref else lib().fdo_vlhog_uoctti
    fn(img.ctypes.data, w, h, cell_size, num_orientations, out.ctypes.data, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (hw, hh)
    return out


class Sdm:
    """SdmLandmarkModel + SdmLandmarkModelFitting (fd_sdm.c)."""

    def __init__(self, model=None, path=None, use_ref=False):
        """model: featuredetection_b200.synthetic.SdmModel; or path of a reference text model.
        use_ref: descriptors come from the reference's own hog.c (oracle/_ref) instead of the restatement."""
        L_ = self._L = ref() if use_ref else lib()
        if path is not None:
            self.h = L_.fdo_sdm_load(path.encode())
            if not self.h:
                raise ValueError("cannot load SDM model %s" % path)
        else:
            regs = (C.c_void_p * len(model.regressors))(*[r.ctypes.data for r in model.regressors])
            self.h = L_.fdo_sdm_create(model.num_landmarks, len(model.regressors), model.mean.ctypes.data, regs)
        self.L = L_.fdo_sdm_num_landmarks(self.h)
        self.steps = L_.fdo_sdm_num_steps(self.h)

    def __del__(self):
        try:
            self._L.fdo_sdm_free(self.h)
        except Exception:
            pass

    def to_model(self):
        from featuredetection_b200.synthetic import SdmModel
        L_ = self._L
        mean = np.ctypeslib.as_array(L_.fdo_sdm_mean(self.h), shape=(2 * self.L,)).copy()
        rows = self.L * 279 + 1
        regs = [np.ctypeslib.as_array(L_.fdo_sdm_regressor(self.h, s), shape=(rows, 2 * self.L)).copy() for s in range(self.steps)]
        return SdmModel(mean, regs)

    def align_rigid(self, box):
        shape = np.empty(2 * self.L, np.float32)
        self._L.fdo_sdm_align_rigid(self.h, int(box[0]), int(box[1]), int(box[2]), int(box[3]), shape.ctypes.data)
        return shape

    def optimize(self, image, shape, want_features=False):
        image = np.ascontiguousarray(image, np.uint8)
        shape = np.ascontiguousarray(shape, np.float32).copy()
        feats = np.zeros((self.steps, self.L * 279), np.float32) if want_features else None
        rc = self._L.fdo_sdm_optimize(self.h, image.ctypes.data, image.shape[1], image.shape[0], image.shape[1], shape.ctypes.data,
                                    feats.ctypes.data if want_features else None)
        if rc != 0:
            raise RuntimeError("fdo_sdm_optimize: region of interest outside the image (%d)" % rc)
        return (shape, feats) if want_features else shape

    def fit(self, image, box):
        return self.optimize(image, self.align_rigid(box))


def sdm_descriptors(image, pts_xy, window_half):
    image = np.ascontiguousarray(image, np.uint8)
    pts = np.ascontiguousarray(pts_xy, np.float32).reshape(-1, 2)
    out = np.zeros((pts.shape[0], 279), np.float32)
    rc = lib().fdo_sdm_descriptors(image.ctypes.data, image.shape[1], image.shape[0], image.shape[1], pts.ctypes.data, pts.shape[0],
                                   window_half, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("fdo_sdm_descriptors: region of interest outside the image")
    return out


def detections_to_array(buf, n):
    """ctypes Detection array -> numpy structured array copy"""
    dt = np.dtype([(name, np.ctypeslib.as_ctypes_type(np.dtype(_np_of(ct)))) for name, ct in capi.Detection._fields_], align=True)
    return np.frombuffer(buf, dtype=dt, count=n).copy()


def _np_of(ct):
    return {C.c_int32: "i4", C.c_int64: "i8", C.c_float: "f4", C.c_double: "f8"}[ct]


DETECTION_DTYPE = np.dtype([(name, _np_of(ct)) for name, ct in capi.Detection._fields_], align=True)
assert DETECTION_DTYPE.itemsize == C.sizeof(capi.Detection)
SCORE_DTYPE = np.dtype([("fout", "f4"), ("level", "i4")])


def detect_frame(det_kwargs, wvm, svm, frame, stage=capi.FDB_STAGE_NMS, roi=(0, 0, 0, 0), frame_index=0,
                 want_dense=True, want_patches=False, det_cap=1 << 16, timing=False, svm_features=None):
    """Whole reference path on one frame (fdo_detect_frame). Returns a dict."""
    from featuredetection_b200.synthetic import detector_desc
    desc = detector_desc(**det_kwargs)
    frame = np.ascontiguousarray(frame, np.uint8)
    H, W = frame.shape
    L = lib()
    # window count
    p = L.fdo_pyramid_build(frame.ctypes.data, W, H, W, desc.incremental_scale_factor, desc.min_scale_factor, desc.max_scale_factor)
    if not p:
        raise ValueError("invalid pyramid parameters")
    infos = (capi.LayerInfo * max(p.contents.n_layers, 1))()
    total = L.fdo_enumerate(p, desc.patch_width, desc.patch_height, max(desc.step_x, 1), max(desc.step_y, 1),
                            roi[0], roi[1], roi[2], roi[3], infos, p.contents.n_layers)
    n_layers = p.contents.n_layers
    L.fdo_pyramid_free(p)
    dense = np.zeros(total, SCORE_DTYPE) if want_dense else None
    patches = np.zeros((total, desc.patch_width * desc.patch_height), np.uint8) if want_patches else None
    dets = np.zeros(det_cap, DETECTION_DTYPE)
    counts = (C.c_int64 * 5)()
    tim = (C.c_double * 5)()
    svm_dense = np.zeros(total, np.float64) if (wvm is None and want_dense) else None
    n = L.fdo_detect_frame_ex(C.byref(desc), wvm.h if wvm is not None else None, svm.h if svm is not None else None,
                              svm_features.h if svm_features is not None else None, frame.ctypes.data, W, H, W,
                              frame_index, roi[0], roi[1], roi[2], roi[3], stage,
                              dense.ctypes.data if (dense is not None and wvm is not None) else None,
                              patches.ctypes.data if (patches is not None and wvm is not None) else None,
                              svm_dense.ctypes.data if svm_dense is not None else None,
                              dets.ctypes.data, det_cap, counts, tim if timing else None)
    if n < 0:
        raise RuntimeError("fdo_detect_frame failed (%d)" % n)
    layers = [{f: getattr(infos[i], f) for f, _ in capi.LayerInfo._fields_} for i in range(n_layers)]
    return dict(windows=int(total), dense=dense, patches=patches, detections=dets[:n].copy(),
                counts=list(counts), timing=list(tim), layers=layers, svm_dense=svm_dense)


def ref_detect_frame(det_kwargs, wvm, svm, frame, stage=capi.FDB_STAGE_NMS, want_dense=True, det_cap=1 << 16, svm_features=None,
                     pyramid_impl="restated"):
    """One frame through the reference's own classes (oracle/_ref; see ref_driver.cpp:ref_detect_frame).
    wvm / svm are Wvm / Svm objects created with use_ref=True. pyramid_impl "cv2": the pyramid comes from cv2_pyramid()
    (timing[0] = its wall time). Returns a dict."""
    from featuredetection_b200.synthetic import detector_desc
    desc = detector_desc(**det_kwargs)
    frame = np.ascontiguousarray(frame, np.uint8)
    H, W = frame.shape
    nwin = C.c_int64()
    tim = (C.c_double * 5)()
    wins = np.zeros(det_cap, np.int64)
    # first call sizes the dense buffer
    dense = None
    if want_dense:
        from featuredetection_b200 import capi as _c
        lib_ = _c.load_library() if False else None  # (no product code involved)
        p = lib().fdo_pyramid_build(frame.ctypes.data, W, H, W, desc.incremental_scale_factor, desc.min_scale_factor, desc.max_scale_factor)
        infos = (capi.LayerInfo * max(p.contents.n_layers, 1))()
        total = lib().fdo_enumerate(p, desc.patch_width, desc.patch_height, max(desc.step_x, 1), max(desc.step_y, 1), 0, 0, 0, 0, infos, p.contents.n_layers)
        lib().fdo_pyramid_free(p)
        dense = np.zeros(total, SCORE_DTYPE)
    if pyramid_impl == "cv2":
        import time as _time
        t0 = _time.perf_counter()
        _, layers = cv2_pyramid(frame, desc.incremental_scale_factor, desc.min_scale_factor, desc.max_scale_factor)
        arr = (_Layer * max(len(layers), 1))()
        keep = []
        for i, (idx, sc, img) in enumerate(layers):
            img = np.ascontiguousarray(img)
            keep.append(img)
            arr[i].index, arr[i].scale, arr[i].width, arr[i].height = idx, sc, img.shape[1], img.shape[0]
            arr[i].data = img.ctypes.data_as(C.POINTER(C.c_uint8))
        t_pyr = _time.perf_counter() - t0
        n = ref().ref_detect_layers_ex(C.byref(desc), wvm.h, svm.h if svm is not None else None,
                                       svm_features.h if svm_features is not None else None, arr, len(layers), W, H, stage,
                                       dense.ctypes.data if dense is not None else None, C.byref(nwin), wins.ctypes.data, det_cap, tim)
        tim[0] = t_pyr
    else:
        n = ref().ref_detect_frame_ex(C.byref(desc), wvm.h, svm.h if svm is not None else None,
                                      svm_features.h if svm_features is not None else None, frame.ctypes.data, W, H, stage,
                                      dense.ctypes.data if dense is not None else None, C.byref(nwin), wins.ctypes.data, det_cap, tim)
    if n < 0:
        raise RuntimeError("ref_detect_frame failed (%d)" % n)
    return dict(windows=int(nwin.value), dense=dense, det_windows=wins[:n].copy(), timing=list(tim))


def evaluate_samples(det_kwargs, wvm, svm, frame, samples_xywh, max_svm_patches=8):
    """condensation::WvmSvmModel::evaluate(image, samples) restated (libCondensation/src/condensation/WvmSvmModel.cpp:74-119)
    over a caching extractor: DirectPyramidFeatureExtractor::extract(x, y, w, h) (DirectPyramidFeatureExtractor.cpp:67-73,
    133-147), ImagePyramid::getLayer(double) (ImagePyramid.cpp:307-310), ImagePyramidLayer::getScaled
    (ImagePyramidLayer.hpp:65-67). wvm / svm: Wvm / Svm oracle objects (svm may be None). Small inputs only (python loop).
    Ties of equal WVM probability in the top-k cut keep first-seen order (std::sort leaves them open)."""
    import math
    pw, ph = det_kwargs["patch_width"], det_kwargs["patch_height"]
    olc, layers = pyramid(frame, float(det_kwargs["incremental_scale_factor"]), det_kwargs["min_scale_factor"], det_kwargs["max_scale_factor"])
    inc = math.pow(0.5, 1.0 / olc)  # ImagePyramid.cpp:90-91: the pyramid keeps the factor recomputed from the octave layer count
    by_index = {idx: (scale, img) for idx, scale, img in layers}
    n = len(samples_xywh)
    target = np.zeros(n, bool); weight = np.zeros(n, np.float64)
    cache, order, sample_key = {}, [], [None] * n
    L = lib()
    for i, (x, y, w, h) in enumerate(np.asarray(samples_xywh, np.int64).tolist()):
        if w <= 0:
            continue
        power = math.log(float(pw) / float(w)) / math.log(inc)
        index = int(math.floor(abs(power) + 0.5)) * (1 if power >= 0 else -1)  # std::round: half away from zero
        if index not in by_index:
            continue
        scale, img = by_index[index]
        half_w, half_h = int(w / 2), int(h / 2)  # C++ integer division truncates
        px, py = L.fdo_cvround((x - half_w) * scale), L.fdo_cvround((y - half_h) * scale)
        if px < 0 or py < 0 or px + pw > img.shape[1] or py + ph > img.shape[0]:
            continue
        key = (index, px, py)
        if key not in cache:
            patch = hq64(img[py:py + ph, px:px + pw]).ravel()
            level, fout, prob, pos = wvm.eval(patch[None])
            cache[key] = (bool(pos[0]), float(prob[0]), patch)
            if pos[0]:
                order.append(key)
        sample_key[i] = key
        weight[i] = 0.5 * cache[key][1]
    if order and svm is not None:
        if max_svm_patches > 0 and len(order) > max_svm_patches:
            order = sorted(order, key=lambda k: -cache[k][1])[:max_svm_patches]  # sorted() is stable
        res = {}
        for key in order:
            d, p, q = svm.eval(cache[key][2][None])
            res[key] = (bool(q[0]), float(p[0]))
        for i in range(n):
            if sample_key[i] in res:
                target[i] = res[sample_key[i]][0]
                weight[i] = 2 * weight[i] * res[sample_key[i]][1]
    return target, weight


def fhog(image, cell=4, unsigned_bins=9, interpolate_bins=False, interpolate_cells=True, alpha=0.2, use_ref=False):
    """FhogFilter::applyTo (filtering/FhogFilter.cpp:59-67): u8 image [H, W] or [H, W, 3] -> float32 [H // cell, W // cell,
    3 * unsigned_bins + 4]. use_ref: the reference's own classes (oracle/_ref)."""
    img = np.ascontiguousarray(image, np.uint8)
    rows, cols = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    args = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
    if use_ref:
        fn = ref().ref_fhog
    else:
        fn = lib().fdo_fhog
    fn.restype = C.c_int64; fn.argtypes = args
    out = np.zeros((rows // cell, cols // cell, 3 * unsigned_bins + 4), np.float32)
    n = fn(img.ctypes.data, cols, rows, ch, cell, unsigned_bins, int(interpolate_bins), int(interpolate_cells), alpha, out.ctypes.data)
    if n != out.size:
        raise ValueError("fhog: invalid arguments")
    return out


def aggregated_features_detect(frame, weights, bias, threshold, cell=4, octave_layer_count=5, min_window_width=0,
                               nms_threshold=0.3, nms_type=0, width_scale=1.0, height_scale=1.0, unsigned_bins=9,
                               interpolate_bins=False, interpolate_cells=True, alpha=0.2, want_scores=False):
    """detection::AggregatedFeaturesDetector::detectWithScores on a gray frame, restated for the next
    round's kernels - no product code exists for it yet:
      AggregatedFeaturesExtractor ctor / update / getMinScaleFactor / getMaxScaleFactor / computeBoundsInImagePixels
                                                             extraction/AggregatedFeaturesExtractor.cpp:22-131
      ImagePyramid(size_t, double, double), createLayers(Mat)  ImagePyramid.cpp:60-76,170-198 (fdo_pyramid_build)
      FhogFilter as the layer filter                           filtering/FhogFilter.cpp (fdo_fhog)
      ConvolutionFilter::applyTo as the score layer filter     ConvolutionFilter.cpp:31-49: -bias + sum_c filter2D(channel_c, w_c),
                                                               correlation, anchor (0, 0), zero border
      getPositiveWindows / rescaleWindow                       AggregatedFeaturesDetector.cpp:92-118
      NonMaximumSuppression                                    NonMaximumSuppression.cpp:27-112 (fdo_non_maximum_suppression)
    weights [kh, kw, D] float32 = the support vector of the linear SVM. The score map is float32; cv::filter2D's own summation
    order (and its DFT path for kernels of >= 50 elements) is OpenCV's: UNPINNED, compare at 1e-4.
    Returns (rects [n, 4] int32 x y w h, scores [n] float32) after suppression (+ the per-layer score maps)."""
    import math
    frame = np.ascontiguousarray(frame, np.uint8)
    H, W = frame.shape
    weights = np.ascontiguousarray(weights, np.float32)
    kh, kw, D = weights.shape
    inc = math.pow(0.5, 1.0 / octave_layer_count)                      # ImagePyramid.cpp:75
    patch_w, patch_h = kw * cell, kh * cell                            # patchSizeInPixels
    max_scale = 1.0
    if min_window_width > patch_w:                                     # AggregatedFeaturesExtractor.cpp:30-31,46-51
        ms = float(patch_w) / min_window_width
        max_scale = math.pow(inc, int(math.ceil(math.log(ms) / math.log(inc))))
    aspect, image_aspect = float(patch_h) / float(patch_w), float(H) / float(W)   # getMaxWidth, :66-73
    max_width = int(H / aspect) if aspect > image_aspect else W
    mn = float(patch_w) / max_width                                    # getMinScaleFactor, :60-64
    min_scale = math.pow(inc, int(math.log(mn) / math.log(inc)))
    _, layers = pyramid(frame, inc, min_scale, max_scale)
    scores_list, rects_list, maps = [], [], []
    for index, scale, img in layers:
        feat = fhog(img, cell, unsigned_bins, interpolate_bins, interpolate_cells, alpha)
        rows, cols = feat.shape[:2]
        vh, vw = rows - kh + 1, cols - kw + 1                          # AggregatedFeaturesDetector.cpp:95-96
        if vh <= 0 or vw <= 0:
            maps.append(np.zeros((max(vh, 0), max(vw, 0)), np.float32))
            continue
        score = np.full((vh, vw), np.float32(-bias), np.float32)       # filtered = delta
        for c in range(D):                                             # filtered += filter2D(channel_c, w_c)
            tmp = np.zeros((vh, vw), np.float32)
            for i in range(kh):
                for j in range(kw):
                    tmp += feat[i:i + vh, j:j + vw, c] * weights[i, j, c]
            score += tmp
        maps.append(score)
        scale_x, scale_y = float(img.shape[1]) / float(W), float(img.shape[0]) / float(H)   # ImagePyramid.cpp:178-179,187-188
        ys, xs = np.nonzero(score > np.float32(threshold))
        for y, x in zip(ys.tolist(), xs.tolist()):
            # computeBoundsInImagePixels (:121-128): std::round = half away from zero
            def rnd(v):
                return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)
            bx, by = rnd((x * cell) / scale_x), rnd((y * cell) / scale_y)
            bw, bh = rnd((kw * cell) / scale_x), rnd((kh * cell) / scale_y)
            cx, cy = bx + bw // 2, by + bh // 2                        # Patch::computeCenter (non-negative sizes)
            rw, rh = int(np.float32(width_scale) * np.float32(bw)), int(np.float32(height_scale) * np.float32(bh))   # Size(float, float) -> int
            rects_list.append((cx - rw // 2, cy - rh // 2, rw, rh))     # Patch::computeBounds
            scores_list.append(score[y, x])
    scores = np.array(scores_list, np.float32)
    rects = np.array(rects_list, np.int32).reshape(-1, 4)
    if len(scores):
        L = lib()
        L.fdo_non_maximum_suppression.restype = C.c_int64
        L.fdo_non_maximum_suppression.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int]
        rects = np.ascontiguousarray(rects)
        n = L.fdo_non_maximum_suppression(scores.ctypes.data, rects.ctypes.data, len(scores), nms_threshold, nms_type)
        scores, rects = scores[:n].copy(), rects[:n].copy()
    return (rects, scores, maps) if want_scores else (rects, scores)
