"""ctypes mirror of include/fdb200.h and the loader of the CUDA library.

There is no CPU fallback: :func:`load_library` raises if ``libfdb200.so`` (the hand-written
sm_100a kernels + C ABI, built in-tree by ``__graft_entry__.build()``) is missing.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDB_LIB") or os.path.join(_HERE, "csrc", "libfdb200.so")  # FDB_LIB: tuning builds only

FDB_OK = 0
FDB_ERR_INVALID_ARGUMENT, FDB_ERR_RUNTIME, FDB_ERR_CUDA, FDB_ERR_NO_DEVICE, FDB_ERR_UNSUPPORTED, FDB_ERR_OVERFLOW = 1, 2, 3, 4, 5, 6
FDB_STAGE_WVM, FDB_STAGE_OE, FDB_STAGE_SVM, FDB_STAGE_NMS = 1, 2, 3, 4
FDB_SV_U8, FDB_SV_F32 = 0, 1
FDB_KERNEL_RBF, FDB_KERNEL_POLYNOMIAL, FDB_KERNEL_HIK, FDB_KERNEL_LINEAR = range(4)
(FDB_FEATURE_HQ64, FDB_FEATURE_GRAY, FDB_FEATURE_HISTEQ, FDB_FEATURE_WHI, FDB_FEATURE_HOG, FDB_FEATURE_EHOG,
 FDB_FEATURE_LBP) = range(7)
FDB_NORM_NONE, FDB_NORM_L2NORM, FDB_NORM_L2HYS, FDB_NORM_L1NORM, FDB_NORM_L1SQRT = range(5)
FDB_LBP8, FDB_LBP8_UNIFORM, FDB_LBP4, FDB_LBP4_ROTATED = range(4)


class Rect4(C.Structure):
    _fields_ = [("x1", C.c_int32), ("y1", C.c_int32), ("x2", C.c_int32), ("y2", C.c_int32)]


class WvmDesc(C.Structure):
    _fields_ = [
        ("filter_size_x", C.c_int32), ("filter_size_y", C.c_int32),
        ("num_lin_filters", C.c_int32), ("num_filters_per_level", C.c_int32),
        ("num_levels", C.c_int32), ("num_used_filters", C.c_int32),
        ("basis_param", C.c_float), ("limit_reliability_filter", C.c_float),
        ("lin_thresholds", C.POINTER(C.c_float)),
        ("hk_weights", C.POINTER(C.c_float)),
        ("app_rsv_convol", C.POINTER(C.c_double)),
        ("hierarchical_thresholds", C.POINTER(C.c_float)),
        ("area_cntval", C.POINTER(C.c_int32)),
        ("area_val", C.POINTER(C.c_double)),
        ("area_cntrec", C.POINTER(C.c_int32)),
        ("area_rec", C.POINTER(Rect4)),
        ("logistic_a", C.c_double), ("logistic_b", C.c_double),
    ]


class SvmDesc(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32), ("gamma", C.c_double),
        ("num_sv", C.c_int32), ("dim", C.c_int32), ("sv_type", C.c_int32),
        ("support_vectors", C.c_void_p),
        ("coefficients", C.POINTER(C.c_float)),
        ("bias", C.c_float), ("threshold", C.c_float),
        ("logistic_a", C.c_double), ("logistic_b", C.c_double),
        ("poly_alpha", C.c_double), ("poly_constant", C.c_double), ("poly_degree", C.c_int32),
    ]


class RvmDesc(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32), ("gamma", C.c_double),
        ("poly_alpha", C.c_double), ("poly_constant", C.c_double), ("poly_degree", C.c_int32),
        ("num_filters", C.c_int32), ("num_filters_to_use", C.c_int32),
        ("dim", C.c_int32), ("sv_type", C.c_int32),
        ("support_vectors", C.c_void_p),
        ("coefficients", C.POINTER(C.c_float)),
        ("hierarchical_thresholds", C.POINTER(C.c_float)),
        ("bias", C.c_float),
        ("logistic_a", C.c_double), ("logistic_b", C.c_double),
    ]


class DetectorDesc(C.Structure):
    _fields_ = [
        ("incremental_scale_factor", C.c_double),
        ("min_scale_factor", C.c_double), ("max_scale_factor", C.c_double),
        ("patch_width", C.c_int32), ("patch_height", C.c_int32),
        ("step_x", C.c_int32), ("step_y", C.c_int32),
        ("oe_dist", C.c_float), ("oe_ratio", C.c_float),
        ("max_positives_per_frame", C.c_int32),
    ]


class FeatureDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("gradient_kernel", C.c_int32), ("blur_kernel", C.c_int32), ("bins", C.c_int32),
        ("signed_gradients", C.c_int32), ("interpolate_bins", C.c_int32), ("cell_size", C.c_int32),
        ("block_size", C.c_int32), ("interpolate_cells", C.c_int32), ("concatenate", C.c_int32),
        ("signed_and_unsigned", C.c_int32), ("normalization", C.c_int32), ("lbp_type", C.c_int32),
        ("ehog_alpha", C.c_float), ("whi_alpha", C.c_float), ("whi_cutoff", C.c_float),
    ]


class SdmDesc(C.Structure):
    _fields_ = [("num_landmarks", C.c_int32), ("num_cascade_steps", C.c_int32),
                ("mean_landmarks", C.POINTER(C.c_float)), ("regressors", C.POINTER(C.c_void_p))]


class AggdetDesc(C.Structure):
    _fields_ = [("cell_size", C.c_int32), ("window_cols", C.c_int32), ("window_rows", C.c_int32), ("octave_layer_count", C.c_int32),
                ("min_window_width", C.c_int32), ("width_scale", C.c_float), ("height_scale", C.c_float), ("unsigned_bins", C.c_int32),
                ("interpolate_bins", C.c_int32), ("interpolate_cells", C.c_int32), ("alpha", C.c_float), ("weights", C.POINTER(C.c_float)),
                ("bias", C.c_float), ("threshold", C.c_float), ("nms_overlap_threshold", C.c_double), ("nms_type", C.c_int32)]


FDB_NMS_MAX_SCORE, FDB_NMS_AVERAGE, FDB_NMS_WEIGHTED_AVERAGE = 0, 1, 2


class WindowScore(C.Structure):
    _fields_ = [("fout", C.c_float), ("level", C.c_int32)]


class LayerInfo(C.Structure):
    _fields_ = [
        ("index", C.c_int32), ("scale", C.c_double),
        ("width", C.c_int32), ("height", C.c_int32),
        ("orig_patch_width", C.c_int32), ("orig_patch_height", C.c_int32),
        ("windows_x", C.c_int32), ("windows_y", C.c_int32),
        ("first_window", C.c_int64),
    ]


class Detection(C.Structure):
    _fields_ = [
        ("frame", C.c_int32), ("layer", C.c_int32), ("x", C.c_int32), ("y", C.c_int32),
        ("center_x", C.c_int32), ("center_y", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32),
        ("window", C.c_int64),
        ("wvm_level", C.c_int32), ("wvm_fout", C.c_float),
        ("wvm_probability", C.c_double),
        ("svm_distance", C.c_double), ("svm_probability", C.c_double),
        ("probability", C.c_double),
        ("positive", C.c_int32), ("reserved", C.c_int32),
    ]


# every symbol include/fdb200.h declares: (name, restype, argtypes)
_P = C.POINTER
SYMBOLS = [
    ("fdb_abi_version", C.c_int, []),
    ("fdb_last_error", C.c_char_p, []),
    ("fdb_status_string", C.c_char_p, [C.c_int]),
    ("fdb_ctx_create", C.c_int, [C.c_int, _P(C.c_void_p)]),
    ("fdb_ctx_destroy", None, [C.c_void_p]),
    ("fdb_ctx_stream", C.c_void_p, [C.c_void_p]),
    ("fdb_ctx_synchronize", C.c_int, [C.c_void_p]),
    ("fdb_ctx_launch_count", C.c_int64, [C.c_void_p]),
    ("fdb_ctx_timer_start", C.c_int, [C.c_void_p]),
    ("fdb_ctx_timer_stop", C.c_int, [C.c_void_p, _P(C.c_double)]),
    ("fdb_host_alloc", C.c_int, [C.c_size_t, _P(C.c_void_p)]),
    ("fdb_host_free", None, [C.c_void_p]),
    ("fdb_wvm_create", C.c_int, [C.c_void_p, _P(WvmDesc), _P(C.c_void_p)]),
    ("fdb_wvm_destroy", None, [C.c_void_p]),
    ("fdb_wvm_set_limit_reliability_filter", C.c_int, [C.c_void_p, C.c_float]),
    ("fdb_wvm_get_probability", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("fdb_svm_create", C.c_int, [C.c_void_p, _P(SvmDesc), _P(C.c_void_p)]),
    ("fdb_svm_destroy", None, [C.c_void_p]),
    ("fdb_svm_set_threshold", C.c_int, [C.c_void_p, C.c_float]),
    ("fdb_svm_get_probability", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("fdb_svm_has_dense", C.c_int, [C.c_void_p]),
    ("fdb_rvm_create", C.c_int, [C.c_void_p, _P(RvmDesc), _P(C.c_void_p)]),
    ("fdb_rvm_destroy", None, [C.c_void_p]),
    ("fdb_rvm_set_num_filters_to_use", C.c_int, [C.c_void_p, C.c_int32]),
    ("fdb_rvm_get_probability", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("fdb_detector_create_rvm", C.c_int, [C.c_void_p, _P(DetectorDesc), C.c_void_p, _P(C.c_void_p)]),
    ("fdb_detector_create", C.c_int, [C.c_void_p, _P(DetectorDesc), C.c_void_p, C.c_void_p, _P(C.c_void_p)]),
    ("fdb_detector_destroy", None, [C.c_void_p]),
    ("fdb_detector_prepare", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    ("fdb_detector_layers", C.c_int, [C.c_void_p, _P(LayerInfo), C.c_int32, _P(C.c_int32)]),
    ("fdb_detector_windows_per_frame", C.c_int64, [C.c_void_p]),
    ("fdb_detector_pyramid_bytes", C.c_int64, [C.c_void_p]),
    ("fdb_detect_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, _P(C.c_int64)]),
    ("fdb_detect_batch_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, _P(C.c_int64)]),
    ("fdb_detect_enqueue_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    ("fdb_detect_profile_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _P(C.c_double)]),
    ("fdb_detect_roi", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, _P(C.c_int64)]),
    ("fdb_extract_patches", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _P(C.c_int64)]),
    ("fdb_pyramid_layer", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64]),
    ("fdb_plan_layers", C.c_int, [_P(DetectorDesc), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P(LayerInfo), C.c_int32, _P(C.c_int32), _P(C.c_int64)]),
    ("fdb_overlap_eliminate", C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_float, _P(C.c_int64)]),
    ("fdb_non_maximum_suppression", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int32, _P(C.c_int64)]),
    ("fdb_five_stage_nms", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, _P(C.c_int64)]),
    ("fdb_svm_file_load", C.c_int, [C.c_char_p, _P(C.c_void_p)]),
    ("fdb_svm_file_desc", _P(SvmDesc), [C.c_void_p]),
    ("fdb_svm_file_free", None, [C.c_void_p]),
    ("fdb_wvm_file_load", C.c_int, [C.c_char_p, C.c_char_p, _P(C.c_void_p)]),
    ("fdb_wvm_file_desc", _P(WvmDesc), [C.c_void_p]),
    ("fdb_wvm_file_free", None, [C.c_void_p]),
    ("fdb_rvm_file_load", C.c_int, [C.c_char_p, C.c_char_p, _P(C.c_void_p)]),
    ("fdb_rvm_file_desc", _P(RvmDesc), [C.c_void_p]),
    ("fdb_rvm_file_free", None, [C.c_void_p]),
    ("fdb_svm_mat_load", C.c_int, [C.c_char_p, C.c_char_p, _P(C.c_void_p)]),
    ("fdb_gray_from_bgr", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    ("fdb_detect_batch_bgr", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, _P(C.c_int64)]),
    ("fdb_detector_last_counts", C.c_int, [C.c_void_p, _P(C.c_int64)]),
    ("fdb_feature_shape", C.c_int, [_P(FeatureDesc), C.c_int32, C.c_int32, _P(C.c_int32), _P(C.c_int32)]),
    ("fdb_detector_set_feature", C.c_int, [C.c_void_p, _P(FeatureDesc)]),
    ("fdb_extract_features", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    ("fdb_detect_single", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, _P(C.c_int64)]),
    ("fdb_detect_face_features", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _P(C.c_int64),
                                          C.c_void_p, C.c_int64, C.c_void_p]),
    ("fdb_fhog", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                          C.c_float, C.c_void_p]),
    ("fdb_fhog_score_map", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_float, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    ("fdb_aggdet_windows", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_float,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_int64, _P(C.c_int64)]),
    ("fdb_evaluate_samples", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    ("fdb_detect_single_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, _P(C.c_int64)]),
    ("fdb_detector_single_dense", C.c_int, [C.c_void_p]),
    ("fdb_detector_single_dense_profile", C.c_int, [C.c_void_p, _P(C.c_double), _P(C.c_int32)]),
    ("fdb_detect_single_roi", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                        _P(C.c_int64)]),
    ("fdb_extract_windows", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    ("fdb_detector_set_create", C.c_int, [C.c_void_p, _P(C.c_void_p), C.c_int32, _P(C.c_void_p)]),
    ("fdb_detector_set_destroy", None, [C.c_void_p]),
    ("fdb_detector_set_prepare", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    ("fdb_detector_set_windows_per_frame", C.c_int64, [C.c_void_p]),
    ("fdb_detector_set_info", C.c_int, [C.c_void_p, _P(C.c_int32), _P(C.c_int64), _P(C.c_int32), _P(C.c_int32)]),
    ("fdb_detector_set_last_host_ms", C.c_int, [C.c_void_p, _P(C.c_double)]),
    ("fdb_detector_set_detect_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                                _P(C.c_int64)]),
    ("fdb_detector_set_detect_batch_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, _P(C.c_void_p), C.c_void_p,
                                                       C.c_int64, _P(C.c_int64)]),
    ("fdb_detector_set_profile_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _P(C.c_double)]),
    ("fdb_aggdet_create", C.c_int, [C.c_void_p, _P(AggdetDesc), _P(C.c_void_p)]),
    ("fdb_aggdet_destroy", None, [C.c_void_p]),
    ("fdb_aggdet_prepare", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    ("fdb_aggdet_layers", C.c_int, [C.c_void_p, _P(C.c_int32), C.c_void_p, C.c_int32]),
    ("fdb_aggdet_positions_per_frame", C.c_int64, [C.c_void_p]),
    ("fdb_aggdet_detect_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          _P(C.c_int64)]),
    ("fdb_aggdet_detect_batch_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                 _P(C.c_int64)]),
    ("fdb_aggdet_score_maps", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    ("fdb_aggdet_profile_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _P(C.c_double)]),
    ("fdb_sdm_create", C.c_int, [C.c_void_p, _P(SdmDesc), _P(C.c_void_p)]),
    ("fdb_sdm_destroy", None, [C.c_void_p]),
    ("fdb_sdm_num_landmarks", C.c_int32, [C.c_void_p]),
    ("fdb_sdm_num_cascade_steps", C.c_int32, [C.c_void_p]),
    ("fdb_sdm_file_load", C.c_int, [C.c_char_p, _P(C.c_void_p)]),
    ("fdb_sdm_file_desc", _P(SdmDesc), [C.c_void_p]),
    ("fdb_sdm_file_free", None, [C.c_void_p]),
    ("fdb_sdm_align_rigid", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("fdb_sdm_optimize_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    ("fdb_sdm_optimize_batch_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                                C.c_void_p, C.c_void_p]),
    ("fdb_sdm_profile_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p, _P(C.c_double)]),
    ("fdb_sdm_descriptors", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_void_p]),
]

_lib = None


def load_library(path=None):
    """Load libfdb200.so and bind every declared symbol. Raises if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            "featuredetection_b200: CUDA library %s is missing - run `python -c "
            "'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if path is None:
        _lib = lib
    return lib


class FdbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("fdb200 status %d: %s" % (status, message))
        self.status = status


def check(lib, status):
    if status != FDB_OK:
        msg = lib.fdb_last_error()
        raise FdbError(status, msg.decode("utf-8", "replace") if msg else "")
