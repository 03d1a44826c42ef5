/*
 * fdb200.h - C ABI of the B200-native sliding-window landmark detector hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch/OpenCV types.
 * The reference (elador/FeatureDetection) has no FFI of its own; its seams are the C++
 * abstract classes
 *     imageprocessing::FeatureExtractor / PyramidFeatureExtractor
 *         (libImageProcessing/include/imageprocessing/FeatureExtractor.hpp:32-53,
 *          PyramidFeatureExtractor.hpp:52-118)
 *     classification::ProbabilisticClassifier
 *         (libClassification/include/classification/ProbabilisticClassifier.hpp:33,
 *          BinaryClassifier.hpp:32,42)
 *     detection::Detector
 *         (libDetection/include/detection/Detector.hpp:59-79)
 * Every entry point below names the reference call it replaces.  The C++ adapters that
 * implement those three interfaces on top of this ABI live in adapters/ and
 * INTEGRATION.md shows how a maintainer links them.
 *
 * Conventions
 *  - every function returns an fdb_status (0 = ok), never throws, never takes ownership
 *    of caller memory; fdb_last_error() returns a thread-local message for the last failure.
 *  - "host" pointers are ordinary (pageable or pinned) CPU memory; "device" pointers are
 *    CUDA device memory of the context's device.
 *  - a context is bound to one CUDA device; objects created from it may be used by one
 *    host thread at a time (same contract as the reference objects, which are not re-entrant).
 *  - there is NO CPU fallback: if the CUDA device is missing the create call fails.
 */
#ifndef FDB200_H_
#define FDB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define FDB_API __declspec(dllexport)
#else
#define FDB_API __attribute__((visibility("default")))
#endif

#define FDB_ABI_VERSION 2

typedef enum fdb_status {
	FDB_OK = 0,
	FDB_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument in the reference */
	FDB_ERR_RUNTIME = 2,          /* std::runtime_error in the reference */
	FDB_ERR_CUDA = 3,             /* CUDA runtime/driver failure */
	FDB_ERR_NO_DEVICE = 4,        /* no usable sm_100 device: there is no CPU fallback */
	FDB_ERR_UNSUPPORTED = 5,      /* model outside the exactness envelope */
	FDB_ERR_OVERFLOW = 6          /* a fixed-capacity result buffer was too small */
} fdb_status;

typedef struct fdb_ctx fdb_ctx;
typedef struct fdb_wvm fdb_wvm;
typedef struct fdb_svm fdb_svm;
typedef struct fdb_rvm fdb_rvm;
typedef struct fdb_detector fdb_detector;
typedef struct fdb_detector_set fdb_detector_set;
typedef struct fdb_aggdet fdb_aggdet;

/* ------------------------------------------------------------------------------------------
 * Model descriptors (host memory, copied during create)
 * ---------------------------------------------------------------------------------------- */

/* One rectangle of a rectangle-approximated reduced set vector; corners are INCLUSIVE.
 * Mirrors WvmClassifier::TRec {x1,x2,y1,y2} (WvmClassifier.hpp:103-105). */
typedef struct fdb_rect4 {
	int32_t x1, y1, x2, y2;
} fdb_rect4;

/* Wavelet reduced vector machine + logistic.
 * Replaces the state that WvmClassifier::loadFromMatlab (WvmClassifier.cpp:348-770) and
 * ProbabilisticWvmClassifier::load (ProbabilisticWvmClassifier.cpp:75-85) build:
 * numbers are given already in evaluator units (the loader's x255, /65025 conversions applied). */
typedef struct fdb_wvm_desc {
	int32_t filter_size_x;         /* WvmClassifier::filter_size_x */
	int32_t filter_size_y;
	int32_t num_lin_filters;       /* numLinFilters (e.g. 280) */
	int32_t num_filters_per_level; /* numFiltersPerLevel (e.g. 14) */
	int32_t num_levels;            /* numLevels (e.g. 20) */
	int32_t num_used_filters;      /* 0 or > num_lin_filters: use all (setNumUsedFilters, WvmClassifier.cpp:151) */
	float basis_param;             /* basisParam */
	float limit_reliability_filter;/* cfg "threshold"; added to every hierarchical threshold (WvmClassifier.cpp:165) */
	const float* lin_thresholds;   /* [num_lin_filters] */
	const float* hk_weights;       /* packed lower triangle: filter l owns l+1 weights at offset l(l+1)/2 */
	const double* app_rsv_convol;  /* [num_lin_filters] */
	const float* hierarchical_thresholds; /* [num_lin_filters] hierarchicalThresholdsFromFile */
	/* Area of filter f: grey values val[0..cntval) and, for v >= 1, cntrec rectangles.
	 * Flattened: value slot s = val_offset[f] + v. */
	const int32_t* area_cntval;    /* [num_lin_filters] */
	const double* area_val;        /* [sum cntval] */
	const int32_t* area_cntrec;    /* [sum cntval]; entry for v == 0 is ignored */
	const fdb_rect4* area_rec;     /* [sum cntrec over v >= 1], in (f, v, r) order */
	double logistic_a;             /* ProbabilisticWvmClassifier::logisticA (default 0.00556) */
	double logistic_b;             /* ProbabilisticWvmClassifier::logisticB (default -2.95) */
} fdb_wvm_desc;

typedef enum fdb_kernel_kind {
	FDB_KERNEL_RBF = 0,        /* classification::RbfKernel (RbfKernel.hpp:32-40): exp(-gamma * sum (x - y)^2) */
	FDB_KERNEL_POLYNOMIAL = 1, /* classification::PolynomialKernel (PolynomialKernel.hpp:38-40,62-70): powi(alpha * x.y + constant, degree) */
	FDB_KERNEL_HIK = 2,        /* classification::HistogramIntersectionKernel (HistogramIntersectionKernel.hpp:31-39,59-83): sum min(x, y) */
	FDB_KERNEL_LINEAR = 3      /* classification::LinearKernel (LinearKernel.hpp:27-29): x.y (cv::Mat::dot) */
} fdb_kernel_kind;

typedef enum fdb_sv_type {
	FDB_SV_U8 = 0, /* CV_8U support vectors: exact integer SSD (RbfKernel.hpp:78-88) */
	FDB_SV_F32 = 1 /* CV_32F support vectors: float32 sequential SSD (RbfKernel.hpp:97-108) */
} fdb_sv_type;

/* Support vector machine + logistic.
 * Replaces SvmClassifier::setSvmParameters (SvmClassifier.cpp:62-66),
 * VectorMachineClassifier::setThreshold and ProbabilisticSvmClassifier's logistic
 * (ProbabilisticSvmClassifier.cpp:50-58). */
typedef struct fdb_svm_desc {
	int32_t kernel;          /* fdb_kernel_kind */
	double gamma;            /* RbfKernel::gamma */
	int32_t num_sv;
	int32_t dim;             /* elements per support vector (patch w*h for u8) */
	int32_t sv_type;         /* fdb_sv_type */
	const void* support_vectors; /* [num_sv][dim] row-major, u8 or f32 */
	const float* coefficients;   /* [num_sv] */
	float bias;              /* VectorMachineClassifier::bias */
	float threshold;         /* VectorMachineClassifier::threshold */
	double logistic_a;
	double logistic_b;
	double poly_alpha;       /* PolynomialKernel::alpha (ABI version 2) */
	double poly_constant;    /* PolynomialKernel::constant */
	int32_t poly_degree;     /* PolynomialKernel::degree */
} fdb_svm_desc;

/* Reduced vector machine cascade + logistic.
 * Replaces RvmClassifier (RvmClassifier.hpp:33-125: supportVectors, coefficients per level, hierarchicalThresholds,
 * numFiltersToUse; the MATLAB loader RvmClassifier.cpp:131-319 is the only way the reference fills them) and
 * ProbabilisticRvmClassifier's logistic (ProbabilisticRvmClassifier.cpp:56-64). */
typedef struct fdb_rvm_desc {
	int32_t kernel;              /* fdb_kernel_kind */
	double gamma;                /* RbfKernel::gamma */
	double poly_alpha, poly_constant; int32_t poly_degree;
	int32_t num_filters;         /* reduced set vectors = cascade levels */
	int32_t num_filters_to_use;  /* RvmClassifier::setNumFiltersToUse (RvmClassifier.cpp:119-126): 0 or > num_filters = all */
	int32_t dim;
	int32_t sv_type;             /* fdb_sv_type */
	const void* support_vectors; /* [num_filters][dim] */
	const float* coefficients;   /* RvmClassifier::coefficients packed: level l at l (l + 1) / 2, entries [l][0..l] */
	const float* hierarchical_thresholds; /* [num_filters] */
	float bias;
	double logistic_a, logistic_b;
} fdb_rvm_desc;

/* Image pyramid + window extraction parameters.
 * Replaces ImagePyramid(double inc, double min, double max) (ImagePyramid.cpp:79-92),
 * DirectPyramidFeatureExtractor(pyramid, w, h) (DirectPyramidFeatureExtractor.cpp:36-37) with
 * a GrayscaleFilter image filter and a HistEq64Filter patch filter (ffpDetectApp.cpp:407-415),
 * SlidingWindowDetector(classifier, extractor, stepX, stepY) (SlidingWindowDetector.cpp:27) and
 * OverlapElimination(dist, ratio) (OverlapElimination.cpp:30). */
typedef struct fdb_detector_desc {
	double incremental_scale_factor; /* cfg pyramid.incrementalScaleFactor (float widened, as the app does) */
	double min_scale_factor;
	double max_scale_factor;
	int32_t patch_width;
	int32_t patch_height;
	int32_t step_x;   /* SlidingWindowDetector stepSizeX (default 1) */
	int32_t step_y;
	float oe_dist;    /* overlapElimination.dist (default 5.0) */
	float oe_ratio;   /* overlapElimination.ratio (default 0.0) */
	int32_t max_positives_per_frame; /* capacity of the per-frame stage-1 candidate list; 0 = default 4096 */
} fdb_detector_desc;

/* Feature space of a classifier = the patch-filter chain (and pyramid layer filters) the reference
 * attaches to its feature extractor:
 *   hq64 / gray / histeq / whi   ffpDetectApp.cpp:446-461 ("feature" node of a `single` detector) and
 *                                patchConverter.cpp:218-227
 *   hog / ehog / lbp             adaptiveTrackingApp/AdaptiveTracking.cpp:183-204,222-227 (+ createHogFilter
 *                                :241-253, createLbpFilter :255-268, createHistogramFilter :270-292), parameters as
 *                                in adaptiveTrackingApp/default.cfg:105-129 */
typedef enum fdb_feature_kind {
	FDB_FEATURE_HQ64 = 0,   /* HistEq64Filter (HistEq64Filter.cpp:32-125); u8, w*h */
	FDB_FEATURE_GRAY = 1,   /* no patch filter; u8, w*h */
	FDB_FEATURE_HISTEQ = 2, /* HistogramEqualizationFilter = cv::equalizeHist (HistogramEqualizationFilter.cpp:17-20); u8, w*h */
	FDB_FEATURE_WHI = 3,    /* WhiteningFilter (WhiteningFilter.cpp:20-81) + HistogramEqualizationFilter +
	                         * ConversionFilter(CV_32F, 1/127.5, -1) + UnitNormFilter(NORM_L2); f32, w*h */
	FDB_FEATURE_HOG = 4,    /* layer filters GradientFilter (GradientFilter.cpp:38-59) + GradientBinningFilter
	                         * (GradientBinningFilter.cpp:18-92); patch filter SpatialHistogramFilter
	                         * (SpatialHistogramFilter.cpp:56-94) when block_size == 1 and !signed_and_unsigned,
	                         * else HogFilter (HogFilter.cpp:58-122); f32 */
	FDB_FEATURE_EHOG = 5,   /* same layer filters; patch filter ExtendedHogFilter (ExtendedHogFilter.cpp:54-209); f32 */
	FDB_FEATURE_LBP = 6     /* layer filter LbpFilter (LbpFilter.cpp:56-85, LbpFilter.hpp:88-178); patch filter
	                         * SpatialHistogramFilter; f32 */
} fdb_feature_kind;

typedef enum fdb_normalization { /* HistogramFilter::Normalization (HistogramFilter.hpp:26, HistogramFilter.cpp:222-252) */
	FDB_NORM_NONE = 0, FDB_NORM_L2NORM = 1, FDB_NORM_L2HYS = 2, FDB_NORM_L1NORM = 3, FDB_NORM_L1SQRT = 4
} fdb_normalization;

typedef enum fdb_lbp_type { /* LbpFilter::Type */
	FDB_LBP8 = 0, FDB_LBP8_UNIFORM = 1, FDB_LBP4 = 2, FDB_LBP4_ROTATED = 3
} fdb_lbp_type;

typedef struct fdb_feature_desc {
	int32_t kind;                /* fdb_feature_kind */
	int32_t gradient_kernel;     /* hog/ehog: GradientFilter kernelSize, 1 or 3 (cfg gradientKernel) */
	int32_t blur_kernel;         /* hog/ehog: GradientFilter blurKernelSize; only 0 is supported */
	int32_t bins;                /* hog/ehog: GradientBinningFilter bins */
	int32_t signed_gradients;    /* hog/ehog: cfg "signed" */
	int32_t interpolate_bins;    /* hog/ehog: cfg "interpolate" (two bins per pixel) */
	int32_t cell_size;           /* histogram.cellSize */
	int32_t block_size;          /* histogram.blockSize */
	int32_t interpolate_cells;   /* histogram.interpolate (bilinear between cells) */
	int32_t concatenate;         /* histogram.concatenate (SpatialHistogramFilter blocks) */
	int32_t signed_and_unsigned; /* histogram.signedAndUnsigned */
	int32_t normalization;       /* fdb_normalization (SpatialHistogramFilter) */
	int32_t lbp_type;            /* fdb_lbp_type */
	float ehog_alpha;            /* histogram.alpha (ExtendedHogFilter clamp) */
	float whi_alpha;             /* WhiteningFilter alpha (default 1) */
	float whi_cutoff;            /* WhiteningFilter cutoffFrequency (default 0.390625) */
} fdb_feature_desc;

/* ------------------------------------------------------------------------------------------
 * Result records
 * ---------------------------------------------------------------------------------------- */

/* Dense stage-1 record, one per window in canonical order (layer index asc, y, x);
 * = the pair<int,double> of WvmClassifier::computeHyperplaneDistance (WvmClassifier.cpp:100-149),
 * fout kept in its native float. */
typedef struct fdb_window_score {
	float fout;
	int32_t level; /* index of the last evaluated filter */
} fdb_window_score;

/* Geometry of one pyramid layer as ImagePyramidLayer / PyramidFeatureExtractor expose it. */
typedef struct fdb_layer_info {
	int32_t index;        /* ImagePyramidLayer::getIndex */
	double scale;         /* theoretical scale factor (getScaleFactor) */
	int32_t width, height;/* layer image size (getLayerSizes) */
	int32_t orig_patch_width, orig_patch_height; /* getPatchSizes: cvRound(patch / scale) */
	int32_t windows_x, windows_y; /* window grid of the full-frame scan */
	int64_t first_window; /* canonical index of this layer's first window */
} fdb_layer_info;

/* One classified patch = detection::ClassifiedPatch + imageprocessing::Patch geometry
 * (ClassifiedPatch.hpp:18-97, Patch.hpp:241-243). */
typedef struct fdb_detection {
	int32_t frame;
	int32_t layer;          /* pyramid layer index */
	int32_t x, y;           /* top-left corner inside the layer */
	int32_t center_x, center_y;   /* Patch::getX/getY (original image coordinates) */
	int32_t width, height;  /* Patch::getWidth/getHeight (original image size) */
	int64_t window;         /* canonical window index inside the frame */
	int32_t wvm_level;
	float wvm_fout;
	double wvm_probability; /* ProbabilisticWvmClassifier::getProbability */
	double svm_distance;    /* SvmClassifier::computeHyperplaneDistance (NaN if stage not run) */
	double svm_probability; /* ProbabilisticSvmClassifier::getProbability */
	double probability;     /* ClassifiedPatch::getProbability as the reference returns it */
	int32_t positive;
	int32_t reserved;       /* results of a detector set: index of the detector inside the set; 0 otherwise */
} fdb_detection;

typedef enum fdb_stage {
	FDB_STAGE_WVM = 1,      /* SlidingWindowDetector::detect: windows with wvm positive */
	FDB_STAGE_OE = 2,       /* + OverlapElimination::eliminate */
	FDB_STAGE_SVM = 3,      /* + SVM classify, positives only, sorted (FiveStage detect(Mat,Rect)) */
	FDB_STAGE_NMS = 4       /* + grid NMS of FiveStageSlidingWindowDetector::detect(Mat) */
} fdb_stage;

/* ------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------- */
FDB_API int fdb_abi_version(void);
FDB_API const char* fdb_last_error(void);
FDB_API const char* fdb_status_string(int status);

/* device < 0: use the current CUDA device. Fails with FDB_ERR_NO_DEVICE without a GPU. */
FDB_API int fdb_ctx_create(int device, fdb_ctx** out);
FDB_API void fdb_ctx_destroy(fdb_ctx* ctx);
/* The CUDA stream (cudaStream_t) all work of this context is enqueued on. */
FDB_API void* fdb_ctx_stream(fdb_ctx* ctx);
FDB_API int fdb_ctx_synchronize(fdb_ctx* ctx);
/* Number of kernel launches issued by this context so far (for bench accounting). */
FDB_API int64_t fdb_ctx_launch_count(fdb_ctx* ctx);

/* CUDA-event stopwatch on the context stream (device time, for bench.py):
 * start records an event; stop records a second one, waits for it and returns the elapsed ms. */
FDB_API int fdb_ctx_timer_start(fdb_ctx* ctx);
FDB_API int fdb_ctx_timer_stop(fdb_ctx* ctx, double* elapsed_ms);

/* Pinned host memory for frame/result staging (optional; any host memory is accepted). */
FDB_API int fdb_host_alloc(size_t bytes, void** out);
FDB_API void fdb_host_free(void* p);

/* ------------------------------------------------------------------------------------------
 * Classifiers  (classification::ProbabilisticClassifier surface)
 * ---------------------------------------------------------------------------------------- */
FDB_API int fdb_wvm_create(fdb_ctx* ctx, const fdb_wvm_desc* desc, fdb_wvm** out);
FDB_API void fdb_wvm_destroy(fdb_wvm* wvm);
/* WvmClassifier::setLimitReliabilityFilter (WvmClassifier.cpp:165-181) */
FDB_API int fdb_wvm_set_limit_reliability_filter(fdb_wvm* wvm, float value);

/* ProbabilisticWvmClassifier::getProbability (ProbabilisticWvmClassifier.cpp:42-54) over a batch
 * of n already-extracted feature vectors (u8, filter_size_x*filter_size_y each, continuous).
 * Any output pointer may be NULL. */
FDB_API int fdb_wvm_get_probability(fdb_wvm* wvm, const uint8_t* patches_host, int64_t n,
		int32_t* level_out, float* fout_out, double* probability_out, uint8_t* positive_out);

FDB_API int fdb_svm_create(fdb_ctx* ctx, const fdb_svm_desc* desc, fdb_svm** out);
FDB_API void fdb_svm_destroy(fdb_svm* svm);
FDB_API int fdb_svm_set_threshold(fdb_svm* svm, float threshold);

/* ProbabilisticSvmClassifier::getProbability / SvmClassifier::computeHyperplaneDistance
 * (ProbabilisticSvmClassifier.cpp:50-58, SvmClassifier.cpp:55-60) over n feature vectors
 * of the SVM's sv_type (u8 or f32), dim elements each. */
FDB_API int fdb_svm_get_probability(fdb_svm* svm, const void* vectors_host, int64_t n,
		double* distance_out, double* probability_out, uint8_t* positive_out);

/* 1 when the SVM has a tensor-core form (csrc/svm_dense.cu: u8 support vectors, RBF kernel, gamma and dimension within
 * the shared-memory budget): batches of >= 128 vectors in fdb_svm_get_probability and the `single` detector then
 * evaluate windows x support vectors as one exact u8 matrix product (tcgen05.mma kind::i8) with a float64 epilogue.
 * Same loop as above (SvmClassifier.cpp:55-60, RbfKernel.hpp:32-40); distances agree with it to ~1e-13.
 * CONSEQUENCE: the same feature vector may get distances ~1e-12 apart depending on the batch size (the tensor-core
 * kernel sums the support vectors in a different order than the per-window kernel, which follows the reference's order),
 * so a `distance >= threshold` result can differ for a distance within ~1e-12 of the threshold. Everything else in the
 * library is bit-exact against the reference. Setting the environment variable FDB_SVM_DENSE=0 before the library is
 * loaded forces the reference-order kernel everywhere. */
FDB_API int fdb_svm_has_dense(const fdb_svm* svm);

/* RvmClassifier / ProbabilisticRvmClassifier. The cascade is evaluated the way the reference's live code path does
 * (computeHyperplaneDistance -> computeHyperplaneDistanceCached, RvmClassifier.cpp:75-112): level 0 = -bias + c[0][0] k_0,
 * level l = distance of level l - 1 + c[l][l] k_l (only the diagonal coefficients take part), stopping at the first level
 * whose distance is below its threshold. */
FDB_API int fdb_rvm_create(fdb_ctx* ctx, const fdb_rvm_desc* desc, fdb_rvm** out);
FDB_API void fdb_rvm_destroy(fdb_rvm* rvm);
FDB_API int fdb_rvm_set_num_filters_to_use(fdb_rvm* rvm, int32_t num_filters);
/* ProbabilisticRvmClassifier::getProbability (ProbabilisticRvmClassifier.cpp:52-64) over n feature vectors: level reached,
 * hyperplane distance, probability 1 / (1 + exp(A + B distance)), positive = RvmClassifier::classify (RvmClassifier.cpp:66-73) */
FDB_API int fdb_rvm_get_probability(fdb_rvm* rvm, const void* vectors_host, int64_t n, int32_t* level_out,
		double* distance_out, double* probability_out, uint8_t* positive_out);

/* ------------------------------------------------------------------------------------------
 * Detector  (PyramidFeatureExtractor + Detector surface)
 * ---------------------------------------------------------------------------------------- */
/* svm may be NULL (plain SlidingWindowDetector, ffpDetectApp "single" with pwvm); wvm may be NULL when svm is
 * given ("single" with psvm: see fdb_detect_single). */
FDB_API int fdb_detector_create(fdb_ctx* ctx, const fdb_detector_desc* desc,
		fdb_wvm* wvm, fdb_svm* svm, fdb_detector** out);
/* `single` detector of ffpDetectApp.cpp:427-500 with classifier prvm: every window through the RVM cascade on its HistEq64
 * patch. Use fdb_detect_single / fdb_detect_batch; detections carry wvm_level = level reached, svm_distance = distance. */
FDB_API int fdb_detector_create_rvm(fdb_ctx* ctx, const fdb_detector_desc* desc, fdb_rvm* rvm, fdb_detector** out);
FDB_API void fdb_detector_destroy(fdb_detector* det);

/* Fix the frame geometry (ImagePyramid::update / createLayers sizing, ImagePyramid.cpp:170-198)
 * and allocate device buffers for up to max_batch frames. Must be called before detect. */
FDB_API int fdb_detector_prepare(fdb_detector* det, int32_t width, int32_t height, int32_t max_batch);

/* PyramidFeatureExtractor getters (getLayerScales/getLayerSizes/getPatchSizes,
 * PyramidFeatureExtractor.hpp:72-118). Returns the number of layers; fills up to cap entries. */
FDB_API int fdb_detector_layers(fdb_detector* det, fdb_layer_info* out, int32_t cap, int32_t* n_layers);
FDB_API int64_t fdb_detector_windows_per_frame(fdb_detector* det);
/* Bytes of all pyramid images of one frame (the materialised pyramid) */
FDB_API int64_t fdb_detector_pyramid_bytes(fdb_detector* det);

/* Detector::detect over a batch of frames in HOST memory (8-bit, 1 channel, row pitch in bytes).
 *   stage           how far to run the five-stage cascade (fdb_stage)
 *   dense_out       NULL or host [n_frames * windows_per_frame] stage-1 records
 *   detections_out  host array of capacity det_cap; *n_detections receives the count
 * Copies frames H2D, runs the whole path, copies results D2H and returns after completion.
 * Replaces FiveStageSlidingWindowDetector::detect(Mat) (FiveStageSlidingWindowDetector.cpp:187-322)
 * / SlidingWindowDetector::detect(Mat) (SlidingWindowDetector.cpp:40-50). */
FDB_API int fdb_detect_batch(fdb_detector* det, const uint8_t* frames_host, int64_t pitch,
		int32_t n_frames, int32_t stage, fdb_window_score* dense_out,
		fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections);

/* Same, frames already resident in device memory (pitch == width, frames contiguous);
 * no input copy inside. dense_out_device may be NULL or a DEVICE pointer. */
FDB_API int fdb_detect_batch_device(fdb_detector* det, const uint8_t* frames_device,
		int32_t n_frames, int32_t stage, fdb_window_score* dense_out_device,
		fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections);

/* Stage-1 only on device-resident frames, results stay on the device; returns without
 * synchronising (used for kernel timing). dense_out_device may be NULL. */
FDB_API int fdb_detect_enqueue_device(fdb_detector* det, const uint8_t* frames_device,
		int32_t n_frames, fdb_window_score* dense_out_device);

/* Same enqueue with CUDA events between the kernels of the step (summed over the internal chunks);
 * after completion ms_out[0..4] = resize kernel, pyrDown kernels, fused hq64+WVM window kernel,
 * deep-cascade kernel, whole stage 1 (ms); ms_out[5] = number of launches of each kernel. */
FDB_API int fdb_detect_profile_device(fdb_detector* det, const uint8_t* frames_device,
		int32_t n_frames, double ms_out[6]);

/* ROI variant: Detector::detect(Mat, Rect) (FiveStageSlidingWindowDetector.cpp:331-380,
 * SlidingWindowDetector.cpp:53-79): one frame, windows restricted by roi {x,y,w,h}. */
FDB_API int fdb_detect_roi(fdb_detector* det, const uint8_t* frame_host, int64_t pitch,
		int32_t roi_x, int32_t roi_y, int32_t roi_w, int32_t roi_h, int32_t stage,
		fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections);

/* PyramidFeatureExtractor::extract(stepX, stepY) (DirectPyramidFeatureExtractor.cpp:75-123):
 * writes the HistEq64-filtered patch of every window of one frame, canonical order,
 * patch_w*patch_h bytes each, to patches_out (host). */
FDB_API int fdb_extract_patches(fdb_detector* det, const uint8_t* frame_host, int64_t pitch,
		uint8_t* patches_out, int64_t cap_windows, int64_t* n_windows);

/* ImagePyramid::getLayers image data of one frame: copies layer `layer_index`
 * (width*height bytes) to out (host). */
FDB_API int fdb_pyramid_layer(fdb_detector* det, const uint8_t* frame_host, int64_t pitch,
		int32_t layer_index, uint8_t* out, int64_t cap);

/* ------------------------------------------------------------------------------------------
 * Host-only entry points (no GPU needed)
 * ---------------------------------------------------------------------------------------- */
/* Pyramid geometry and window grid for a width x height frame without touching the device:
 * ImagePyramid::createLayers sizing (ImagePyramid.cpp:170-198) + the window loops of
 * DirectPyramidFeatureExtractor::extract(stepX, stepY, roi) (DirectPyramidFeatureExtractor.cpp:75-123).
 * roi all-zero = whole image. Fills up to cap layer records. */
FDB_API int fdb_plan_layers(const fdb_detector_desc* desc, int32_t width, int32_t height,
		int32_t roi_x, int32_t roi_y, int32_t roi_w, int32_t roi_h,
		fdb_layer_info* out, int32_t cap, int32_t* n_layers, int64_t* n_windows);

/* OverlapElimination::eliminate (OverlapElimination.cpp:44-105) on n classified patches, in place;
 * *n_out receives the surviving count (survivors first, in the reference's output order). */
FDB_API int fdb_overlap_eliminate(fdb_detection* dets, int64_t n, float dist, float ratio, int64_t* n_out);

/* Grid NMS + final ordering of FiveStageSlidingWindowDetector::detect(Mat)
 * (FiveStageSlidingWindowDetector.cpp:143-184,276-311) on the SVM-positive patches of ONE frame. */
FDB_API int fdb_five_stage_nms(fdb_detection* dets, int64_t n, int32_t width, int32_t height, int64_t* n_out);

/* detection::NonMaximumSuppression::eliminateRedundantDetections (libDetection/src/detection/NonMaximumSuppression.cpp:27-112;
 * used by AggregatedFeaturesDetector.cpp:104-112): intersection-over-union suppression of n scored boxes {x, y, w, h}, in
 * place; the first *n_out entries are the surviving boxes, best cluster first. maximum_type: fdb_nms_maximum_type. */
typedef enum fdb_nms_maximum_type { FDB_NMS_MAX_SCORE = 0, FDB_NMS_AVERAGE = 1, FDB_NMS_WEIGHTED_AVERAGE = 2 } fdb_nms_maximum_type;
FDB_API int fdb_non_maximum_suppression(float* scores, int32_t* rects_xywh, int64_t n, double overlap_threshold,
		int32_t maximum_type, int64_t* n_out);

/* The reference's SVM text container (SvmClassifier::store/load(std::ifstream&), SvmClassifier.cpp:68-158,
 * followed by ProbabilisticSvmClassifier's "Logistic a b" line, ProbabilisticSvmClassifier.cpp:65-78).
 * The returned descriptor points into the file object and stays valid until fdb_svm_file_free. */
typedef struct fdb_svm_file fdb_svm_file;
FDB_API int fdb_svm_file_load(const char* path, fdb_svm_file** out);
FDB_API const fdb_svm_desc* fdb_svm_file_desc(const fdb_svm_file* file);
FDB_API void fdb_svm_file_free(fdb_svm_file* file);

/* The MATLAB classifier files every ffpDetectApp .cfg names (classifierFile + thresholdsFile): Level-5 MAT-files read with
 * the library's own reader (the reference uses MATLAB's libmat).
 *   fdb_wvm_file_load  WvmClassifier::loadFromMatlab (WvmClassifier.cpp:348-770) + ProbabilisticWvmClassifier::
 *                      loadSigmoidParamsFromMatlab (ProbabilisticWvmClassifier.cpp:95-137): variables num_hk, support_hk%d,
 *                      weight_hk%d, param_nonlin1[_rvm], num_hk_wvm, num_lev_wvm, area{val_u, cntrec_u, crec{x1,y1,x2,y2}},
 *                      app_rsv_convol; thresholds file: hierar_thresh, posterior_wrvm. Units converted as the loader does.
 *   fdb_svm_mat_load   SvmClassifier::loadFromMatlab (SvmClassifier.cpp:240-335) + ProbabilisticSvmClassifier::
 *                      loadSigmoidParamsFromMatlab (ProbabilisticSvmClassifier.cpp:114-162): param_nonlin1, support_nonlin1,
 *                      weight_nonlin1; logistic file (may be NULL): posterior_svm. Returns the same object type as
 *                      fdb_svm_file_load (use fdb_svm_file_desc / fdb_svm_file_free).
 * The cfg's "threshold" values are applied afterwards with fdb_wvm_set_limit_reliability_filter / fdb_svm_set_threshold,
 * as ProbabilisticWvmClassifier::load / ProbabilisticSvmClassifier::load do. */
typedef struct fdb_wvm_file fdb_wvm_file;
FDB_API int fdb_wvm_file_load(const char* classifier_path, const char* thresholds_path, fdb_wvm_file** out);
FDB_API const fdb_wvm_desc* fdb_wvm_file_desc(const fdb_wvm_file* file);
FDB_API void fdb_wvm_file_free(fdb_wvm_file* file);
/* RvmClassifier::loadFromMatlab (RvmClassifier.cpp:141-319) + the posterior_wrvm logistic (ProbabilisticRvmClassifier.cpp:92-125):
 * float32 reduced set vectors, coefficient rows, hierar_thresh; the descriptor borrows the file object's arrays */
typedef struct fdb_rvm_file fdb_rvm_file;
FDB_API int fdb_rvm_file_load(const char* classifier_path, const char* thresholds_path, fdb_rvm_file** out);
FDB_API const fdb_rvm_desc* fdb_rvm_file_desc(const fdb_rvm_file* file);
FDB_API void fdb_rvm_file_free(fdb_rvm_file* file);
FDB_API int fdb_svm_mat_load(const char* classifier_path, const char* logistic_path, fdb_svm_file** out);

/* ------------------------------------------------------------------------------------------
 * Feature spaces
 * ---------------------------------------------------------------------------------------- */
/* Length and element type (0: u8, 1: float32) of the feature vector a patch of patch_width x patch_height
 * becomes in this feature space (host only). */
FDB_API int fdb_feature_shape(const fdb_feature_desc* desc, int32_t patch_width, int32_t patch_height,
		int32_t* dim, int32_t* is_float);

/* Gives the detector's SVM (the secondClassifier of a fiveStageCascade, or the classifier of a `single`
 * detector) its own feature space: FilteringPyramidFeatureExtractor::addPatchFilter / DirectPyramidFeatureExtractor::
 * addLayerFilter (ffpDetectApp.cpp:445-461, AdaptiveTracking.cpp:183-227). The WVM stage always sees HistEq64
 * patches. Call before fdb_detector_prepare; the SVM's support vectors must have the feature's type and length. */
FDB_API int fdb_detector_set_feature(fdb_detector* det, const fdb_feature_desc* desc);

/* PyramidFeatureExtractor::extract(layer, x, y) (DirectPyramidFeatureExtractor.cpp:125-143 +
 * FilteringPyramidFeatureExtractor.hpp:61-66) for n windows {layer index, x, y} of one frame: writes n feature
 * vectors (fdb_feature_shape elements each) to out (host). Windows must lie inside their layer. */
FDB_API int fdb_extract_features(fdb_detector* det, const uint8_t* frame_host, int64_t pitch,
		const int32_t* layer_x_y, int64_t n, void* out);

/* `single` detector of ffpDetectApp.cpp:427-500 with classifier psvm: a detector created with wvm == NULL
 * classifies EVERY window with the SVM in its feature space (SlidingWindowDetector.cpp:87-98) and returns the
 * positives in canonical order, probability = ProbabilisticSvmClassifier::getProbability.
 * distance_out: NULL or host [n_frames * windows_per_frame] hyperplane distances.
 * (fdb_detect_batch on such a detector does the same without distance_out.) */
FDB_API int fdb_detect_single(fdb_detector* det, const uint8_t* frames_host, int64_t pitch, int32_t n_frames,
		double* distance_out, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections);

/* The same call for frames resident in device memory ([n_frames][height][width], tight pitch); distance_device: NULL
 * or device memory for n_frames * windows_per_frame doubles. Only for detectors whose SVM runs on the tensor cores
 * (fdb_detector_single_dense() == 1), FDB_ERR_UNSUPPORTED otherwise. */
FDB_API int fdb_detect_single_device(fdb_detector* det, const uint8_t* frames_device, int32_t n_frames,
		double* distance_device, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections);
/* SlidingWindowDetector::detect(image, roi) (SlidingWindowDetector.cpp:53-79) of a `single` detector: the windows
 * PyramidFeatureExtractor::extract(stepX, stepY, roi) visits inside the region of interest, all classified; positives in
 * extract order. This is what ffpDetectApp.cpp:591 calls on every feature detector with the face box. */
FDB_API int fdb_detect_single_roi(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, int32_t roi_x, int32_t roi_y,
		int32_t roi_w, int32_t roi_h, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections);
/* PyramidFeatureExtractor::extract(layer, x, y) / extract(x, y, w, h) (DirectPyramidFeatureExtractor.cpp:67-73,125-143) for n
 * windows of one frame: layer_x_y = n triples {pyramid layer index, window corner x, y inside the layer image}; out = n
 * vectors of the detector's patch filter (HistEq64 patches of patch_width * patch_height bytes, or the feature space set with
 * fdb_detector_set_feature); valid_out[i] = 0 where the reference returns an empty pointer (no such layer, window not inside
 * the layer image; the vector is left untouched). */
FDB_API int fdb_extract_windows(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, const int32_t* layer_x_y, int64_t n,
		void* out, uint8_t* valid_out);
FDB_API int fdb_detector_single_dense(fdb_detector* det);
/* time spent in svm_dense_kernel during the last fdb_detect_single[_device] call (CUDA events on the library stream)
 * and the number of launches (one per chunk of frames): the roofline numerator of bench.py --workload single-psvm */
FDB_API int fdb_detector_single_dense_profile(fdb_detector* det, double* kernel_ms, int32_t* launches);

/* imageprocessing::filtering::FhogFilter::applyTo (libImageProcessing/src/imageprocessing/filtering/FhogFilter.cpp:59-67 with
 * FhogAggregationFilter.cpp:43-150): the FHOG feature map of one 8-bit image (1 or 3 interleaved channels), the layer filter
 * of detection::AggregatedFeaturesDetector's feature pyramid. out_host: (height / cell_size) x
 * (width / cell_size) x (3 * unsigned_bins + 4) float32, bit-identical to the reference's filter (GPU tests in
 * tests/test_fhog_host_emulation.py). One image per call (simple kernels); the batched path is fdb_aggdet_*. */
FDB_API int fdb_fhog(fdb_ctx* ctx, const uint8_t* image_host, int64_t pitch, int32_t width, int32_t height, int32_t channels,
		int32_t cell_size, int32_t unsigned_bins, int32_t interpolate_bins, int32_t interpolate_cells, float alpha, float* out_host);

/* The score map AggregatedFeaturesDetector computes for one pyramid layer (AggregatedFeaturesDetector.cpp:60-65,92-98;
 * ConvolutionFilter.cpp:31-49): FHOG of the layer image, then -bias + the correlation with the linear SVM's support vector
 * weights_host [kernel_rows][kernel_cols][3 * unsigned_bins + 4]. scores_host: (cells_y - kernel_rows + 1) x (cells_x -
 * kernel_cols + 1) float32 (nothing is written when no window fits). EXPERIMENTAL like fdb_fhog: host-verified arithmetic,
 * kernels not yet run on a B200; parity with cv::filter2D's own summation order is 1e-4 by design. */
FDB_API int fdb_fhog_score_map(fdb_ctx* ctx, const uint8_t* image_host, int64_t pitch, int32_t width, int32_t height, int32_t channels,
		int32_t cell_size, int32_t unsigned_bins, int32_t interpolate_bins, int32_t interpolate_cells, float alpha,
		const float* weights_host, int32_t kernel_rows, int32_t kernel_cols, float bias, float* scores_host);

/* AggregatedFeaturesDetector::getPositiveWindows for ONE layer's score map (AggregatedFeaturesDetector.cpp:92-118;
 * bounds: AggregatedFeaturesExtractor.cpp:121-128; rescaleWindow :114-118): positions with score > threshold become boxes in
 * image pixels. scale_x / scale_y = layer size / image size. *n_out may exceed cap (only cap entries are written). Host only;
 * feed the boxes of all layers to fdb_non_maximum_suppression. */
FDB_API int fdb_aggdet_windows(const float* score_map, int32_t valid_rows, int32_t valid_cols, float threshold, int32_t kernel_rows,
		int32_t kernel_cols, int32_t cell_size, double scale_x, double scale_y, float width_scale, float height_scale,
		float* scores_out, int32_t* rects_xywh_out, int64_t cap, int64_t* n_out);

/* ------------------------------------------------------------------------------------------
 * detection::AggregatedFeaturesDetector (the reference's newer detector family)
 * ------------------------------------------------------------------------------------------
 * AggregatedFeaturesDetector(imageFilter = GrayscaleFilter, layerFilter = FhogFilter, cellSize, windowSize, octaveLayerCount,
 * svm (LinearKernel), nms, widthScale, heightScale, minWindowWidth) - AggregatedFeaturesDetector.cpp:37-66 - on batches of
 * 8-bit 1-channel frames: image pyramid with one layer per octave step (AggregatedFeaturesExtractor.cpp:22-73: scale limits
 * from the window and image size), FHOG of every layer, score map = -bias + correlation of the feature map with the SVM's
 * support vector (ConvolutionFilter.cpp:31-49), windows with score > threshold -> boxes in image pixels
 * (computeBoundsInImagePixels + rescaleWindow) -> NonMaximumSuppression per frame. Everything up to the candidate list runs on
 * the GPU (csrc/aggdet.cu). Not covered: colour input to the layer filter, the approximated in-between layers of the FPDW
 * variant (ImagePyramid.cpp:200-289) and the LUV / ACF channel filters. */
typedef struct fdb_aggdet_desc {
	int32_t cell_size;                /* cellSizeInPixels */
	int32_t window_cols, window_rows; /* windowSize in cells = size of the SVM's support vector */
	int32_t octave_layer_count;
	int32_t min_window_width;         /* minWindowWidth in pixels; 0: none */
	float width_scale, height_scale;  /* rescaleWindow */
	int32_t unsigned_bins;            /* FhogFilter(cellSize, unsignedBinCount, interpolateBins, interpolateCells, alpha) */
	int32_t interpolate_bins, interpolate_cells;
	float alpha;
	const float* weights;             /* [window_rows][window_cols][3 * unsigned_bins + 4] float32: svm->getSupportVectors()[0] */
	float bias, threshold;            /* svm->getBias(), svm->getThreshold() */
	double nms_overlap_threshold;     /* NonMaximumSuppression(overlapThreshold, maximumType) */
	int32_t nms_type;                 /* fdb_nms_maximum_type */
} fdb_aggdet_desc;
FDB_API int fdb_aggdet_create(fdb_ctx* ctx, const fdb_aggdet_desc* desc, fdb_aggdet** out);
FDB_API void fdb_aggdet_destroy(fdb_aggdet* det);
FDB_API int fdb_aggdet_prepare(fdb_aggdet* det, int32_t width, int32_t height, int32_t max_batch);
/* pyramid layers of the prepared size: info_out rows {layer index, width, height, cells x, cells y, score positions} */
FDB_API int fdb_aggdet_layers(fdb_aggdet* det, int32_t* n_layers, int32_t* info_out, int32_t cap);
FDB_API int64_t fdb_aggdet_positions_per_frame(fdb_aggdet* det); /* windows scored per frame (all layers) */
/* AggregatedFeaturesDetector::detectWithScores on n_frames host frames: scores_out [cap], rects_xywh_out [cap][4] (image
 * pixels), frame_out [cap] (may be NULL); per frame in the order NonMaximumSuppression returns them */
FDB_API int fdb_aggdet_detect_batch(fdb_aggdet* det, const uint8_t* frames_host, int64_t pitch, int32_t n_frames, float* scores_out,
		int32_t* rects_xywh_out, int32_t* frame_out, int64_t cap, int64_t* n_out);
FDB_API int fdb_aggdet_detect_batch_device(fdb_aggdet* det, const uint8_t* frames_device, int32_t n_frames, float* scores_out,
		int32_t* rects_xywh_out, int32_t* frame_out, int64_t cap, int64_t* n_out);
/* parity / debugging: the valid score positions of every layer of one frame, concatenated in layer order (scores_out, may be
 * NULL) and the FHOG feature maps [cells][3 * unsigned_bins + 4] of every layer, concatenated (features_out, may be NULL) */
FDB_API int fdb_aggdet_score_maps(fdb_aggdet* det, const uint8_t* frame_host, int64_t pitch, float* scores_out, int64_t cap,
		float* features_out, int64_t feat_cap);
/* bench.py: kernels serialised between CUDA events: ms_out = {pyramid, histograms, descriptors, score maps, total, chunks} */
FDB_API int fdb_aggdet_profile_device(fdb_aggdet* det, const uint8_t* frames_device, int32_t n_frames, double ms_out[6]);

/* The per-frame flow of ffpDetectApp (ffpDetectApp.cpp:553-596): the face detector on the whole frame, then every feature
 * detector restricted to the bounds of the FIRST (most probable) face patch - Patch::getBounds() = {x - w / 2, y - h / 2, w, h}
 * - through Detector::detect(img, roi). face_out receives the face detections; feature_out has feature_cap_each slots per
 * feature detector, n_feature[i] their counts. Without a face nothing else runs (all n_feature = 0). All detectors must be
 * prepared for the frame's size. */
FDB_API int fdb_detect_face_features(fdb_detector* face, fdb_detector* const* features, int32_t n_features,
		const uint8_t* frame_host, int64_t pitch, fdb_detection* face_out, int64_t face_cap, int64_t* n_face,
		fdb_detection* feature_out, int64_t feature_cap_each, int64_t* n_feature);

/* condensation::WvmSvmModel::evaluate(image, samples) (libCondensation/src/condensation/WvmSvmModel.cpp:74-119): the
 * tracker's sparse use of the two classifiers on one frame. samples_xywh[i] = {centre x, centre y, width, height} (Sample.hpp);
 * each sample's patch is DirectPyramidFeatureExtractor::extract(x, y, w, h) (DirectPyramidFeatureExtractor.cpp:67-73,133-147;
 * layer choice ImagePyramid.cpp:307-310); equal patches are classified once (CachingPyramidFeatureExtractor + the model's
 * cache). target_out[i] / weight_out[i] = Sample::isTarget / getWeight afterwards: no patch -> (0, 0); WVM-negative or not
 * among the max_svm_patches (reference: 8) most probable WVM positives -> (0, 0.5 P_wvm); else (SVM positive, P_wvm P_svm).
 * max_svm_patches <= 0 disables the cut (= MeasurementModel::evaluate(Sample&) for every sample). Needs a detector with a WVM
 * (the SVM may be NULL: no second stage) working on HistEq64 patches. */
FDB_API int fdb_evaluate_samples(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, const int32_t* samples_xywh,
		int64_t n, int32_t max_svm_patches, uint8_t* target_out, double* weight_out);

/* ------------------------------------------------------------------------------------------
 * Detector set: every detector of an application on every frame
 * ------------------------------------------------------------------------------------------
 * ffpDetectApp creates one detector per landmark cfg, each with its own ImagePyramid (ffpDetectApp.cpp:391-500: loop over the
 * "detectors" nodes; pyramids at :407 and :435), and runs all of them on every frame (ffpDetectApp.cpp:548-596). A set produces
 * exactly the detections its members would produce one by one (fdb_detect_batch on each), but builds every distinct pyramid
 * image once per frame and equalises windows of the same layer and size once for all members that scan them.
 * Members: cascade detectors (WVM first stage) of the same context; the set does not own them (destroy the set first).
 * Results are ordered by member, then frame, then as fdb_detect_batch orders them; fdb_detection.reserved = member index;
 * per-member counters through fdb_detector_last_counts.
 * Environment (diagnostics, never needed): FDB_WINDOW_KERNEL=mma|tc forces one of the two window kernels (default: the tcgen05
 * kernel for packs of three and four detectors of one window geometry, the mma.sync kernel otherwise - same results);
 * FDB_HOST_THREADS=n sizes the post-processing pool (default: the affinity mask, at most 16); FDB_SET_TRACE=1 prints the host
 * phases and the stage-1 intervals of every chunk on the device clock to stderr. */
FDB_API int fdb_detector_set_create(fdb_ctx* ctx, fdb_detector* const* detectors, int32_t n_detectors, fdb_detector_set** out);
FDB_API void fdb_detector_set_destroy(fdb_detector_set* set);
/* prepares every member (fdb_detector_prepare) and the shared pyramid / work tables for frames of width x height */
FDB_API int fdb_detector_set_prepare(fdb_detector_set* set, int32_t width, int32_t height, int32_t max_batch);
FDB_API int64_t fdb_detector_set_windows_per_frame(fdb_detector_set* set); /* all members */
/* shared pyramid: images built per frame and their bytes, window-kernel launches per chunk, members on the shared kernels */
FDB_API int fdb_detector_set_info(fdb_detector_set* set, int32_t* n_images, int64_t* pyramid_bytes, int32_t* n_window_launches,
		int32_t* n_fast_members);
/* host wall clock of the last detect call in milliseconds: {enqueueing the chunks, phase A = candidate lists + overlap
 * elimination + SVM launch, phase B = classification + NMS, whole call, waiting for stage 1, and phase A split into fetching
 * long candidate lists, the members' CPU work (thread pool), launching} - where a batch spends its time outside the kernels */
FDB_API int fdb_detector_set_last_host_ms(fdb_detector_set* set, double ms_out[8]);
/* Detector::detect of every member on n_frames host frames (8-bit, 1 channel, row pitch `pitch`) */
FDB_API int fdb_detector_set_detect_batch(fdb_detector_set* set, const uint8_t* frames_host, int64_t pitch, int32_t n_frames,
		int32_t stage, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections);
/* the same with frames resident in device memory (contiguous); dense_out_device: NULL or one pointer per member
 * (NULL or device memory [n_frames][windows of that member]) receiving the stage-1 record of every window */
FDB_API int fdb_detector_set_detect_batch_device(fdb_detector_set* set, const uint8_t* frames_device, int32_t n_frames, int32_t stage,
		fdb_window_score* const* dense_out_device, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections);
/* bench.py: stage-1 kernels of the set serialised between CUDA events, summed over the chunks: ms_out = {resize, pyrDown,
 * window kernels, deep kernel, total, window-kernel launches} */
FDB_API int fdb_detector_set_profile_device(fdb_detector_set* set, const uint8_t* frames_device, int32_t n_frames, double ms_out[6]);

/* Per-stage counters of the last detect call (TOT/TACC-style counters, ffpDetectApp.cpp:650-657):
 * [0] windows, [1] wvm positives, [2] after OE, [3] svm positives, [4] after NMS. */
FDB_API int fdb_detector_last_counts(fdb_detector* det, int64_t counts[5]);

/* GrayscaleFilter::applyTo (GrayscaleFilter.cpp:18-24) for colour input: cv::cvtColor(CV_BGR2GRAY) of n_frames interleaved
 * 8-bit BGR frames (row pitch >= 3 * width bytes, frame k at k * pitch * height), OpenCV 2.4.3 arithmetic
 * (1868 B + 9617 G + 4899 R + 8192) >> 14. fdb_gray_from_bgr returns the gray frames (host, width * height bytes each);
 * fdb_detect_batch_bgr is fdb_detect_batch on such frames (the conversion runs on the device in front of the pyramid).
 * 1-channel frames take the filter's copy branch: pass them to fdb_detect_batch directly. */
FDB_API int fdb_gray_from_bgr(fdb_ctx* ctx, const uint8_t* bgr_host, int64_t pitch, int32_t width, int32_t height,
		int32_t n_frames, uint8_t* gray_host);
FDB_API int fdb_detect_batch_bgr(fdb_detector* det, const uint8_t* bgr_host, int64_t pitch, int32_t n_frames, int32_t stage,
		fdb_window_score* dense_out, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections);

/* ------------------------------------------------------------------------------------------
 * Supervised-descent landmark regressor (libSupervisedDescent; BASELINE configs[4])
 * ---------------------------------------------------------------------------------------- */
typedef struct fdb_sdm fdb_sdm;

/* SdmLandmarkModel state (SdmLandmarkModel.hpp:123-131). Every cascade step uses the "vlhog-uoctti" descriptor with
 * adaptive parameters (SdmLandmarkModel.cpp:181-215 with empty descriptorParameters; DescriptorExtractor.hpp:140-144:
 * 3 x 3 cells, 9 orientations, patch resized to 30 x 30 -> 279 values per landmark). */
typedef struct fdb_sdm_desc {
	int32_t num_landmarks;          /* L >= 13: optimize() reads landmarks 8, 9, 11, 12 (SdmLandmarkModel.hpp:212-216) */
	int32_t num_cascade_steps;      /* <= 8 */
	const float* mean_landmarks;    /* 2 L: all x, then all y */
	const float* const* regressors; /* [steps] -> (279 L + 1) x 2 L row-major float32, last row = bias */
} fdb_sdm_desc;

/* SdmLandmarkModel::SdmLandmarkModel(meanLandmarks, ..., regressorData, ...) (SdmLandmarkModel.cpp:34-41): uploads the model. */
FDB_API int fdb_sdm_create(fdb_ctx* ctx, const fdb_sdm_desc* desc, fdb_sdm** out);
FDB_API void fdb_sdm_destroy(fdb_sdm* sdm);
/* getNumLandmarks / getNumCascadeSteps (SdmLandmarkModel.cpp:43-51) */
FDB_API int32_t fdb_sdm_num_landmarks(const fdb_sdm* sdm);
FDB_API int32_t fdb_sdm_num_cascade_steps(const fdb_sdm* sdm);

/* SdmLandmarkModel::load (SdmLandmarkModel.cpp:130-232), text format with a "descriptorType vlhog-uoctti" line per step.
 * The descriptor points into the file object and stays valid until fdb_sdm_file_free. */
typedef struct fdb_sdm_file fdb_sdm_file;
FDB_API int fdb_sdm_file_load(const char* path, fdb_sdm_file** out);
FDB_API const fdb_sdm_desc* fdb_sdm_file_desc(const fdb_sdm_file* file);
FDB_API void fdb_sdm_file_free(fdb_sdm_file* file);

/* SdmLandmarkModelFitting::alignRigid(modelShape = mean, faceBox) (SdmLandmarkModel.hpp:156-192) for n_faces boxes
 * {x, y, width, height}: shapes_out[n_faces][2 L] (host; a handful of float operations per landmark, done on the host). */
FDB_API int fdb_sdm_align_rigid(const fdb_sdm* sdm, const int32_t* boxes_xywh, int64_t n_faces, float* shapes_out);

/* SdmLandmarkModelFitting::optimize(modelShape, image) (SdmLandmarkModel.hpp:199-256) for a batch of faces.
 *  frames_host   n_frames 8-bit 1-channel images of width x height, row pitch `pitch` bytes, frame k at k * pitch * height
 *  face_frame    [n_faces] index of the frame each face lies in (NULL: face i lies in frame i)
 *  shapes        [n_faces][2 L] in: start shapes (alignRigid), out: fitted shapes
 *  status_out    NULL or [n_faces]: 0 = ok, s + 1 = a HOG window of cascade step s left the (extended) image, where the
 *                reference's Mat::operator()(Rect) throws (DescriptorExtractor.hpp:173); that face's shape is what it was
 *                before step s
 *  features_out  NULL or [steps][n_faces][279 L]: the descriptor rows (for stage-wise checks) */
FDB_API int fdb_sdm_optimize_batch(fdb_sdm* sdm, const uint8_t* frames_host, int64_t pitch, int32_t width, int32_t height,
		int32_t n_frames, const int32_t* face_frame, int64_t n_faces, float* shapes, int32_t* status_out, float* features_out);
/* Same with everything resident on the device (frames contiguous, pitch == width); asynchronous on the context's stream. */
FDB_API int fdb_sdm_optimize_batch_device(fdb_sdm* sdm, const uint8_t* frames_device, int32_t width, int32_t height,
		int32_t n_frames, const int32_t* face_frame_device, int64_t n_faces, float* shapes_device, int32_t* status_device);
/* Same as the device call, synchronous, with CUDA-event times per kernel family summed over the cascade steps:
 * ms_out = {HOG descriptors, regressor product, shape update, total}. */
FDB_API int fdb_sdm_profile_device(fdb_sdm* sdm, const uint8_t* frames_device, int32_t width, int32_t height,
		int32_t n_frames, const int32_t* face_frame_device, int64_t n_faces, float* shapes_device, int32_t* status_device,
		double ms_out[4]);

/* VlHogDescriptorExtractor::getDescriptors(image, locations, windowSizeHalf) (DescriptorExtractor.hpp:106-219), adaptive
 * parameters: n_points x 279 float32 to out (host). Returns FDB_ERR_RUNTIME when a window leaves the extended image. */
FDB_API int fdb_sdm_descriptors(fdb_sdm* sdm, const uint8_t* frame_host, int64_t pitch, int32_t width, int32_t height,
		const float* points_xy, int32_t n_points, int32_t window_size_half, float* out);

#ifdef __cplusplus
}
#endif
#endif /* FDB200_H_ */
