"""Seeded synthetic frames and WVM/SVM models.

The reference ships no model weights for this path (every ffpDetectApp .cfg points at absent
MATLAB files, e.g. ffpDetectApp/FaceFrontal.cfg:7-8,15-16), so benchmarks and parity tests run on
synthetic models whose *shapes* come from the .cfg files and whose numbers are generated here,
deterministically, with integer / IEEE-exact numpy arithmetic only (no cv2, no libm-dependent
steps in the frame generator) so that the GPU box regenerates bit-identical inputs.
"""
import ctypes as C
import json
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
DATA_DIR = os.path.join(_HERE, "data")

# cfg order of SURVEY.md section 8 (ffpDetectApp/*.cfg:17-23; filter counts from classifierFile names)
# name: (patch_w, patch_h, inc, min, max, filters_per_level, levels, rbf_r)
LANDMARK_CONFIGS = [
    ("FaceFrontal", 20, 20, 0.92, 0.05, 0.16, 14, 20, 0.04),
    ("FaceLeftProfile", 20, 20, 0.9, 0.09, 0.25, 14, 7, 0.09),
    ("FaceRightProfile", 20, 20, 0.9, 0.09, 0.25, 14, 7, 0.09),
    ("RightEyeCenter", 32, 16, 0.85, 0.5, 0.7, 20, 8, 0.0325),
    ("LeftEyeCenter", 32, 16, 0.9, 0.5, 0.7, 20, 8, 0.0325),
    ("CenterLipUpperOuter", 24, 24, 0.9, 0.5, 0.7, 30, 8, 0.0195),
    ("LeftEyeOuterCorner", 24, 24, 0.9, 0.5, 0.7, 20, 11, 0.013),
    ("RightEyeOuterCorner", 24, 24, 0.9, 0.5, 0.7, 20, 11, 0.013),
    ("LeftLipCorner", 24, 24, 0.9, 0.5, 0.7, 30, 8, 0.0325),
    ("RightLipCorner", 24, 24, 0.9, 0.5, 0.7, 30, 8, 0.0325),
    ("LeftNoseCorner", 24, 24, 0.9, 0.5, 0.7, 30, 8, 0.0325),
    ("RightNoseCorner", 24, 24, 0.9, 0.5, 0.7, 30, 8, 0.0325),
    ("NoseTip", 32, 24, 0.9, 0.5, 0.7, 30, 7, 0.026),
    ("LeftEarCenter", 16, 24, 0.9, 0.5, 0.7, 20, 10, 0.0325),
    ("RightEarCenter", 16, 24, 0.9, 0.5, 0.7, 20, 10, 0.0325),
]


def landmark_config(name):
    for i, c in enumerate(LANDMARK_CONFIGS):
        if c[0] == name:
            return i, c
    raise KeyError(name)


# ------------------------------------------------------------------------------------------------
# frames
# ------------------------------------------------------------------------------------------------
def _box_sum(a, r):
    """Exact integer box sum of radius r with edge replication (separable, via cumsum)."""
    for axis in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (r + 1, r)
        p = np.pad(a, pad, mode="edge")
        c = np.cumsum(p, axis=axis, dtype=np.int64)
        n = a.shape[axis]
        hi = np.take(c, np.arange(2 * r + 1, 2 * r + 1 + n), axis=axis)
        lo = np.take(c, np.arange(0, n), axis=axis)
        a = hi - lo
    return a


def synthetic_frame(k, width=640, height=480):
    """8-bit 1-channel frame k: smooth random field (3x box blur, radius 8) stretched to 0..255
    plus +-12 noise. Integer arithmetic only => identical on every machine."""
    rng = np.random.default_rng(1234 + int(k))
    base = rng.integers(0, 256, (height, width), dtype=np.int64)
    img = base
    for _ in range(3):
        img = _box_sum(img, 8)
    lo, hi = int(img.min()), int(img.max())
    img = (img - lo) * 255 // max(hi - lo, 1)
    noise = rng.integers(-12, 13, (height, width), dtype=np.int64)
    return np.clip(img + noise, 0, 255).astype(np.uint8)


def synthetic_frames(first, count, width=640, height=480):
    return np.stack([synthetic_frame(first + i, width, height) for i in range(count)])


# ------------------------------------------------------------------------------------------------
# hq64 in numpy (used only to draw support vectors; float32 ops are IEEE-exact in numpy)
# ------------------------------------------------------------------------------------------------
def _hq64_np(patch):
    h, w = patch.shape
    s = np.float32(255.0) / np.float32(w * h)
    counts = np.bincount((patch.ravel() >> 2), minlength=64).astype(np.float32)
    pdf = counts * s
    cdf = np.zeros(64, np.float32)
    acc = np.float32(0)
    for i in range(64):
        acc = np.float32(acc + pdf[i]) if i else pdf[0]
        cdf[i] = acc
    eq = np.floor(cdf.astype(np.float64) + 0.5).astype(np.uint8)
    return eq[patch >> 2]


# ------------------------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------------------------
class WvmModel:
    """Host-side state of a WvmClassifier + logistic in evaluator units
    (WvmClassifier.hpp:77-130, ProbabilisticWvmClassifier.hpp:35)."""

    def __init__(self, w, h, per_level, levels, basis_param, lin_thresholds, hk_weights,
                 app_rsv_convol, thresholds, cntval, val, cntrec, rec,
                 limit_reliability_filter=0.0, num_used=0, logistic_a=0.00556, logistic_b=-2.95):
        self.w, self.h = int(w), int(h)
        self.per_level, self.levels = int(per_level), int(levels)
        self.n = self.per_level * self.levels
        self.basis_param = np.float32(basis_param)
        self.lin_thresholds = np.ascontiguousarray(lin_thresholds, np.float32)
        self.hk_weights = np.ascontiguousarray(hk_weights, np.float32)
        self.app_rsv_convol = np.ascontiguousarray(app_rsv_convol, np.float64)
        self.thresholds = np.ascontiguousarray(thresholds, np.float32)
        self.cntval = np.ascontiguousarray(cntval, np.int32)
        self.val = np.ascontiguousarray(val, np.float64)
        self.cntrec = np.ascontiguousarray(cntrec, np.int32)
        self.rec = np.ascontiguousarray(rec, np.int32).reshape(-1, 4)  # x1,y1,x2,y2
        self.limit_reliability_filter = float(limit_reliability_filter)
        self.num_used = int(num_used)
        self.logistic_a, self.logistic_b = float(logistic_a), float(logistic_b)

    def desc(self):
        d = capi.WvmDesc()
        d.filter_size_x, d.filter_size_y = self.w, self.h
        d.num_lin_filters, d.num_filters_per_level, d.num_levels = self.n, self.per_level, self.levels
        d.num_used_filters = self.num_used
        d.basis_param = float(self.basis_param)
        d.limit_reliability_filter = self.limit_reliability_filter
        d.lin_thresholds = self.lin_thresholds.ctypes.data_as(C.POINTER(C.c_float))
        d.hk_weights = self.hk_weights.ctypes.data_as(C.POINTER(C.c_float))
        d.app_rsv_convol = self.app_rsv_convol.ctypes.data_as(C.POINTER(C.c_double))
        d.hierarchical_thresholds = self.thresholds.ctypes.data_as(C.POINTER(C.c_float))
        d.area_cntval = self.cntval.ctypes.data_as(C.POINTER(C.c_int32))
        d.area_val = self.val.ctypes.data_as(C.POINTER(C.c_double))
        d.area_cntrec = self.cntrec.ctypes.data_as(C.POINTER(C.c_int32))
        d.area_rec = self.rec.ctypes.data_as(C.POINTER(capi.Rect4))
        d.logistic_a, d.logistic_b = self.logistic_a, self.logistic_b
        d._keepalive = self
        return d

    def with_thresholds(self, thresholds):
        m = WvmModel.__new__(WvmModel)
        m.__dict__.update(self.__dict__)
        m.thresholds = np.ascontiguousarray(thresholds, np.float32)
        assert m.thresholds.shape == (self.n,)
        return m


class SvmModel:
    """Host-side state of an SvmClassifier (RBF) + logistic (SvmClassifier.hpp:165-167)."""

    def __init__(self, support_vectors, coefficients, gamma, bias=0.0, threshold=0.0,
                 logistic_a=0.00556, logistic_b=-2.95, kernel="rbf", alpha=1.0, constant=0.0, degree=2):
        sv = np.ascontiguousarray(support_vectors)
        assert sv.ndim == 2 and sv.dtype in (np.uint8, np.float32)
        self.sv = sv
        self.coef = np.ascontiguousarray(coefficients, np.float32)
        self.gamma = float(gamma)
        self.kernel = {"rbf": capi.FDB_KERNEL_RBF, "polynomial": capi.FDB_KERNEL_POLYNOMIAL, "hik": capi.FDB_KERNEL_HIK,
                       "linear": capi.FDB_KERNEL_LINEAR}[kernel]
        self.alpha, self.constant, self.degree = float(alpha), float(constant), int(degree)
        self.bias, self.threshold = float(np.float32(bias)), float(np.float32(threshold))
        self.logistic_a, self.logistic_b = float(logistic_a), float(logistic_b)

    def desc(self):
        d = capi.SvmDesc()
        d.kernel = self.kernel
        d.gamma = self.gamma
        d.poly_alpha, d.poly_constant, d.poly_degree = self.alpha, self.constant, self.degree
        d.num_sv, d.dim = self.sv.shape
        d.sv_type = capi.FDB_SV_U8 if self.sv.dtype == np.uint8 else capi.FDB_SV_F32
        d.support_vectors = self.sv.ctypes.data
        d.coefficients = self.coef.ctypes.data_as(C.POINTER(C.c_float))
        d.bias, d.threshold = self.bias, self.threshold
        d.logistic_a, d.logistic_b = self.logistic_a, self.logistic_b
        d._keepalive = self
        return d


def make_wvm(w, h, per_level, levels, rbf_r, seed, cntval=5, rects_per_value=4):
    """Synthetic WVM: per filter `cntval` grey values and `rects_per_value` inclusive rectangles
    for v >= 1 (2..w/3 px extents, so overlaps are rare and the
    rectangle image stays in the 0..255 range). Filters at wavelet level l >= 1 are zero-mean residuals and app_rsv_convol is the
    exact squared norm of the CUMULATIVE rectangle image (WvmClassifier.cpp:313-314 accumulates
    x.p across wavelet levels), so norm = |x - P_cum|^2 >= 0. Thresholds are -inf ("no-exit")
    until calibrated (see load_thresholds)."""
    rng = np.random.default_rng(int(seed))
    n = per_level * levels
    cv = np.full(n, cntval, np.int32)
    val = np.zeros((n, cntval), np.float64)
    cntrec = np.zeros((n, cntval), np.int32)
    cntrec[:, 1:] = rects_per_value
    rec = np.zeros((n, cntval - 1, rects_per_value, 4), np.int32)
    cum = np.zeros((per_level, h, w), np.float64)
    convol = np.zeros(n, np.float64)
    for f in range(n):
        lev, k = divmod(f, per_level)
        if lev == 0:
            val[f] = rng.uniform(64.0, 192.0, cntval)
        else:
            val[f] = rng.normal(0.0, 24.0 * 0.7 ** lev, cntval)
        img = np.full((h, w), val[f, 0])
        for v in range(1, cntval):
            for r in range(rects_per_value):
                rw = int(rng.integers(2, max(3, w // 3) + 1)); rh = int(rng.integers(2, max(3, h // 3) + 1))
                x1 = int(rng.integers(0, w - rw + 1)); x2 = x1 + rw - 1
                y1 = int(rng.integers(0, h - rh + 1)); y2 = y1 + rh - 1
                rec[f, v - 1, r] = (x1, y1, x2, y2)
                img[y1:y2 + 1, x1:x2 + 1] += val[f, v] - val[f, 0]
        cum[k] += img
        convol[f] = float(np.sum(cum[k] * cum[k]))
    weights = np.concatenate([rng.normal(0.0, (l + 1) ** -0.5, l + 1) for l in range(n)]).astype(np.float32)
    return WvmModel(w, h, per_level, levels, np.float32(rbf_r) / np.float32(65025.0),
                    np.zeros(n, np.float32), weights, convol,
                    np.full(n, -np.inf, np.float32), cv, val.ravel(), cntrec.ravel(),
                    rec.reshape(-1, 4))


class RvmModel:
    """Host-side state of an RvmClassifier (RvmClassifier.hpp:120-124) + logistic (ProbabilisticRvmClassifier.hpp)."""

    def __init__(self, support_vectors, coefficients, thresholds, gamma=7.689e-7, bias=0.0, num_filters_to_use=0,
                 logistic_a=0.00556, logistic_b=-2.95, kernel="rbf", alpha=1.0, constant=0.0, degree=2):
        sv = np.ascontiguousarray(support_vectors)
        assert sv.ndim == 2 and sv.dtype in (np.uint8, np.float32)
        n = sv.shape[0]
        self.sv = sv
        self.coef = np.ascontiguousarray(coefficients, np.float32)      # packed lower triangle, level l at l (l + 1) / 2
        assert self.coef.shape == (n * (n + 1) // 2,)
        self.thresholds = np.ascontiguousarray(thresholds, np.float32)
        assert self.thresholds.shape == (n,)
        self.gamma, self.bias, self.use = float(gamma), float(np.float32(bias)), int(num_filters_to_use)
        self.kernel = {"rbf": capi.FDB_KERNEL_RBF, "polynomial": capi.FDB_KERNEL_POLYNOMIAL, "hik": capi.FDB_KERNEL_HIK,
                       "linear": capi.FDB_KERNEL_LINEAR}[kernel]
        self.alpha, self.constant, self.degree = float(alpha), float(constant), int(degree)
        self.logistic_a, self.logistic_b = float(logistic_a), float(logistic_b)

    @property
    def filters_to_use(self):
        n = self.sv.shape[0]
        return n if self.use <= 0 or self.use > n else self.use

    def desc(self):
        d = capi.RvmDesc()
        d.kernel, d.gamma = self.kernel, self.gamma
        d.poly_alpha, d.poly_constant, d.poly_degree = self.alpha, self.constant, self.degree
        d.num_filters, d.dim = self.sv.shape
        d.num_filters_to_use = self.use
        d.sv_type = capi.FDB_SV_U8 if self.sv.dtype == np.uint8 else capi.FDB_SV_F32
        d.support_vectors = self.sv.ctypes.data
        d.coefficients = self.coef.ctypes.data_as(C.POINTER(C.c_float))
        d.hierarchical_thresholds = self.thresholds.ctypes.data_as(C.POINTER(C.c_float))
        d.bias = self.bias
        d.logistic_a, d.logistic_b = self.logistic_a, self.logistic_b
        d._keepalive = self
        return d


def make_rvm(w, h, seed, num_filters=24, gamma=7.689e-7, survival=0.7):
    """Synthetic u8 RBF RVM cascade: reduced set vectors from make_svm's crops, random coefficient rows, thresholds set so
    that about `survival` of random equalised patches pass each level (calibrated here with numpy on the level sums the
    reference's live code path uses: the diagonal coefficients)."""
    rng = np.random.default_rng(int(seed) + 7)
    svm = make_svm(w, h, seed, num_sv=num_filters, gamma=gamma)
    coef = rng.normal(0, 1, num_filters * (num_filters + 1) // 2).astype(np.float32)
    diag = np.array([coef[l * (l + 1) // 2 + l] for l in range(num_filters)], np.float64)
    probe = svm.sv[rng.integers(0, num_filters, 256)].astype(np.int64)
    probe = np.clip(probe + rng.integers(-40, 41, probe.shape), 0, 255)
    ssd = ((probe[:, None, :] - svm.sv[None].astype(np.int64)) ** 2).sum(axis=2)
    dist = np.cumsum(diag[None] * np.exp(-gamma * ssd), axis=1)
    thr = np.empty(num_filters, np.float32)
    alive = np.ones(len(probe), bool)
    for l in range(num_filters):
        v = dist[alive, l] if alive.sum() > 8 else dist[:, l]
        thr[l] = np.float32(np.quantile(v, 1.0 - survival))
        alive &= dist[:, l] >= thr[l]
    return RvmModel(svm.sv, coef, thr, gamma=gamma)


def make_svm(w, h, seed, num_sv=1024, gamma=7.689e-7):
    """Synthetic u8 RBF-SVM: support vectors are hq64-equalised crops of block-averaged frames
    1000..1003 (gamma from adaptiveTrackingApp/default.cfg:148)."""
    rng = np.random.default_rng(int(seed))
    svs = np.zeros((num_sv, w * h), np.uint8)
    small = []
    for k in range(4):
        f = synthetic_frame(1000 + k).astype(np.int64)
        H, W = f.shape
        f = f[: H // 6 * 6, : W // 6 * 6].reshape(H // 6, 6, W // 6, 6).sum(axis=(1, 3)) // 36
        small.append(f.astype(np.uint8))
    for i in range(num_sv):
        img = small[i % 4]
        y = int(rng.integers(0, img.shape[0] - h)); x = int(rng.integers(0, img.shape[1] - w))
        svs[i] = _hq64_np(img[y:y + h, x:x + w]).ravel()
    coef = rng.normal(0.0, 1.0, num_sv).astype(np.float32)
    return SvmModel(svs, coef, gamma)


def thresholds_path(name, profile="realistic"):
    return os.path.join(DATA_DIR, "thresholds_%s_%s.json" % (name, profile))


def load_thresholds(name, profile="realistic"):
    """Hierarchical thresholds calibrated offline by oracle/tools/calibrate_thresholds.py to the
    survival profile S(l) = max(0.5^(l+1), 2e-3) (stored as float32 bit patterns)."""
    with open(thresholds_path(name, profile)) as fh:
        d = json.load(fh)
    return np.array(d["thresholds_u32"], dtype=np.uint32).view(np.float32)


def landmark_models(name, profile="realistic"):
    """(detector kwargs, WvmModel, SvmModel) for one ffpDetectApp landmark cfg.
    profile: "realistic" (calibrated thresholds) or "no-exit" (all thresholds -inf)."""
    idx, (nm, pw, ph, inc, mn, mx, per_level, levels, r) = landmark_config(name)
    wvm = make_wvm(pw, ph, per_level, levels, r, seed=100 + idx)
    svm = make_svm(pw, ph, seed=300 + idx)
    if profile == "realistic":
        wvm = wvm.with_thresholds(load_thresholds(name, profile))
    elif profile != "no-exit":
        raise ValueError(profile)
    det = dict(incremental_scale_factor=float(np.float32(inc)), min_scale_factor=float(np.float32(mn)),
               max_scale_factor=float(np.float32(mx)), patch_width=pw, patch_height=ph,
               step_x=1, step_y=1, oe_dist=5.0, oe_ratio=0.0)
    return det, wvm, svm


def detector_desc(**kw):
    d = capi.DetectorDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    return d


# ------------------------------------------------------------------------------------------------
# feature spaces (parameters of adaptiveTrackingApp/default.cfg:105-129,148-155)
# ------------------------------------------------------------------------------------------------
FEATURE_KINDS = {"hq64": capi.FDB_FEATURE_HQ64, "gray": capi.FDB_FEATURE_GRAY, "histeq": capi.FDB_FEATURE_HISTEQ,
                 "whi": capi.FDB_FEATURE_WHI, "hog": capi.FDB_FEATURE_HOG, "ehog": capi.FDB_FEATURE_EHOG,
                 "lbp": capi.FDB_FEATURE_LBP}
NORMALIZATIONS = {"none": capi.FDB_NORM_NONE, "l2norm": capi.FDB_NORM_L2NORM, "l2hys": capi.FDB_NORM_L2HYS,
                  "l1norm": capi.FDB_NORM_L1NORM, "l1sqrt": capi.FDB_NORM_L1SQRT}
LBP_TYPES = {"lbp8": capi.FDB_LBP8, "lbp8uniform": capi.FDB_LBP8_UNIFORM, "lbp4": capi.FDB_LBP4,
             "lbp4rotated": capi.FDB_LBP4_ROTATED}
# rbf gamma per feature space (adaptiveTrackingApp/default.cfg:148-155)
FEATURE_GAMMA = {"hq64": 7.689e-7, "gray": 7.689e-7, "histeq": 7.689e-7, "whi": 0.002, "hog": 0.2, "ehog": 0.2, "lbp": 0.6}


def feature_desc(kind="hq64", gradient_kernel=1, blur_kernel=0, bins=9, signed_gradients=False, interpolate_bins=True,
                 cell_size=5, block_size=1, interpolate_cells=False, concatenate=False, signed_and_unsigned=False,
                 normalization="l2norm", lbp_type="lbp8uniform", ehog_alpha=0.2, whi_alpha=1.0, whi_cutoff=0.390625):
    """fdb_feature_desc with the canonical parameters of the reference's cfg files as defaults."""
    d = capi.FeatureDesc()
    d.kind = FEATURE_KINDS[kind] if isinstance(kind, str) else kind
    d.gradient_kernel, d.blur_kernel, d.bins = gradient_kernel, blur_kernel, bins
    d.signed_gradients, d.interpolate_bins = int(signed_gradients), int(interpolate_bins)
    d.cell_size, d.block_size = cell_size, block_size
    d.interpolate_cells, d.concatenate, d.signed_and_unsigned = int(interpolate_cells), int(concatenate), int(signed_and_unsigned)
    d.normalization = NORMALIZATIONS[normalization] if isinstance(normalization, str) else normalization
    d.lbp_type = LBP_TYPES[lbp_type] if isinstance(lbp_type, str) else lbp_type
    d.ehog_alpha, d.whi_alpha, d.whi_cutoff = ehog_alpha, whi_alpha, whi_cutoff
    return d


def feature_sample_windows(layers, patch_w, patch_h, per_layer, seed):
    """seeded window corners {layer index, x, y} inside every pyramid layer (layers: dicts with index/width/height)"""
    rng = np.random.default_rng(seed)
    out = []
    for L in layers:
        for _ in range(per_layer):
            out.append((L["index"], int(rng.integers(0, L["width"] - patch_w + 1)), int(rng.integers(0, L["height"] - patch_h + 1))))
    return np.array(out, np.int32)


def make_feature_svm(vectors, seed, num_sv=1024, gamma=0.2, center=False):
    """RBF SVM whose support vectors are drawn (seeded) from the given feature vectors [n, dim] (u8 or f32):
    SURVEY.md 8(d) - coefficients ~ N(0,1) float32, bias 0, threshold 0, logistic defaults.
    center: subtract the mean coefficient so that distances scatter around the threshold 0."""
    rng = np.random.default_rng(seed)
    vectors = np.ascontiguousarray(vectors)
    idx = rng.integers(0, vectors.shape[0], num_sv)
    sv = vectors[idx].copy()
    if sv.dtype != np.uint8:
        # decorrelate: blend pairs so support vectors are not exact copies of scanned windows
        idx2 = rng.integers(0, vectors.shape[0], num_sv)
        sv = (np.float32(0.5) * (sv.astype(np.float32) + vectors[idx2].astype(np.float32))).astype(np.float32)
    coef = rng.standard_normal(num_sv).astype(np.float32)
    if center:
        coef = (coef - np.float32(coef.astype(np.float64).mean())).astype(np.float32)
    return SvmModel(sv, coef, gamma=gamma)


# ------------------------------------------------------------------------------------------------
# supervised-descent regressor (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------
SDM_DESC = 279  # 3 x 3 cells x (3 * 9 + 4) UoCTTI dimensions (DescriptorExtractor.hpp:140-144)


class SdmModel:
    """SdmLandmarkModel state (SdmLandmarkModel.hpp:123-131): mean shape (all x, then all y, in [-0.5, 0.5]) and one
    (L * 279 + 1) x 2L regressor per cascade step (last row = bias)."""

    def __init__(self, mean, regressors):
        self.mean = np.ascontiguousarray(mean, np.float32)
        self.num_landmarks = self.mean.size // 2
        self.regressors = [np.ascontiguousarray(r, np.float32) for r in regressors]
        for r in self.regressors:
            assert r.shape == (self.num_landmarks * SDM_DESC + 1, 2 * self.num_landmarks)


def make_sdm(num_landmarks=68, num_steps=5, seed=500, sigma=1e-3):
    """Synthetic model: mean shape = a grid in [-0.4, 0.4]^2 with landmarks 8, 9 (inner eye corners) and 11, 12 (mouth
    corners) placed where SdmLandmarkModelFitting::optimize expects them (SdmLandmarkModel.hpp:212-216);
    regressors ~ N(0, sigma) seeded 500 + step."""
    L = num_landmarks
    assert L >= 13
    side = int(np.ceil(np.sqrt(L)))
    gx, gy = np.meshgrid(np.linspace(-0.4, 0.4, side), np.linspace(-0.4, 0.4, side))
    mean = np.concatenate([gx.ravel()[:L], gy.ravel()[:L]]).astype(np.float32)
    for idx, (x, y) in {8: (-0.08, -0.15), 9: (0.08, -0.15), 11: (-0.15, 0.25), 12: (0.15, 0.25)}.items():
        mean[idx], mean[L + idx] = x, y
    regs = []
    for s in range(num_steps):
        rng = np.random.default_rng(seed + s)
        regs.append((rng.standard_normal((L * SDM_DESC + 1, 2 * L)) * sigma).astype(np.float32))
    return SdmModel(mean, regs)


# ------------------------------------------------------------------------------------------------
# MATLAB model files (test support): the variable layout WvmClassifier::loadFromMatlab / SvmClassifier::loadFromMatlab read
# ------------------------------------------------------------------------------------------------
def _sio():
    import scipy.io
    return scipy.io


def _cell(v):
    return np.array([[float(v)]])


def write_wvm_mat(m, classifier_path, thresholds_path, compress, rvm_param=True):
    """inverse of the loader's conversions: grey values / 255, app_rsv_convol / 65025, basisParam * 65025"""
    n = m.n
    out = {"num_hk": _cell(n), "num_hk_wvm": _cell(m.per_level), "num_lev_wvm": _cell(m.levels)}
    for i in range(n):
        out["support_hk%d" % (i + 1)] = np.zeros((m.h, m.w))
        w = m.hk_weights[i * (i + 1) // 2: i * (i + 1) // 2 + i + 1].astype(np.float64)
        out["weight_hk%d" % (i + 1)] = w.reshape(1, -1) if i % 2 == 0 else w.reshape(-1, 1)   # both orientations are legal (:528)
    out["param_nonlin1_rvm" if rvm_param else "param_nonlin1"] = np.array([[float(m.lin_thresholds[0]), 2.0, float(m.basis_param) * 65025.0, 0.0, 1.0]])
    area = np.empty((1, n), dtype=[("val_u", "O"), ("cntrec_u", "O"), ("crec", "O")])
    vo = ro = 0
    for f in range(n):
        cv = int(m.cntval[f])
        cnt = m.cntrec[vo:vo + cv].astype(np.float64).copy()
        cnt[0] = 0
        maxrec = max(1, int(cnt.max()))
        crec = np.empty((cv, maxrec), dtype=[("x1", "O"), ("y1", "O"), ("x2", "O"), ("y2", "O")])
        for v in range(cv):
            for r in range(maxrec):
                crec[v, r] = (_cell(0), _cell(0), _cell(0), _cell(0))
        for v in range(1, cv):
            for r in range(int(cnt[v])):
                x1, y1, x2, y2 = m.rec[ro]
                crec[v, r] = (_cell(x1), _cell(y1), _cell(x2), _cell(y2))
                ro += 1
        area[0, f] = ((m.val[vo:vo + cv] / 255.0).reshape(1, -1), cnt.reshape(1, -1), crec)
        vo += cv
    out["area"] = area
    out["app_rsv_convol"] = (m.app_rsv_convol / 65025.0).reshape(1, -1)
    _sio().savemat(classifier_path, out, format="5", do_compression=compress, oned_as="row")
    _sio().savemat(thresholds_path, {"hierar_thresh": m.thresholds.astype(np.float64).reshape(1, -1),
                                  "posterior_wrvm": np.array([[m.logistic_b, m.logistic_a]]),
                                  "posterior_svm": np.array([[-1.25, 0.5]])}, format="5", do_compression=compress)
    return out


def write_svm_mat(m, w, h, classifier_path, logistic_path=None, compress=True):
    """u8 SvmModel -> support_nonlin1 [h][w][numSV] in grey / 255 units ((v + 0.5) / 255 survives the loader's truncation),
    weight_nonlin1, param_nonlin1 = [bias, 2 (rbf), gamma * 65025, 0, 1]; logistic file: posterior_svm = [B, A]"""
    sv = ((m.sv.astype(np.float64) + 0.5) / 255.0).reshape(-1, h, w).transpose(1, 2, 0)
    _sio().savemat(classifier_path, {"param_nonlin1": np.array([[m.bias, 2.0, m.gamma * 65025.0, 0.0, 1.0]]), "support_nonlin1": sv,
                                     "weight_nonlin1": m.coef.astype(np.float64).reshape(1, -1)}, format="5", do_compression=compress)
    if logistic_path is not None:
        _sio().savemat(logistic_path, {"posterior_svm": np.array([[m.logistic_b, m.logistic_a]])}, format="5", do_compression=compress)
