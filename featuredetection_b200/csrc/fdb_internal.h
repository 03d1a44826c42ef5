/*
 * fdb_internal.h - internal structures shared by the C ABI (api.cu), the host-side plan
 * (plan.cpp), the host post-processing (hostpost.cpp) and the kernel launchers (*.cu).
 */
#ifndef FDB_INTERNAL_H_
#define FDB_INTERNAL_H_

#include <cstdint>
#include <string>
#include <vector>

#include "fdb200.h"

namespace fdb {

/* ---- error plumbing -------------------------------------------------------------------- */
void set_error(const std::string& msg);
int fail(int status, const std::string& msg);
/* status of the exception in flight (call inside a catch block): no C++ exception may cross the C ABI */
int api_exception() noexcept;
#define FDB_API_CATCH catch (...) { return fdb::api_exception(); }

/* ---- pyramid plan (host) --------------------------------------------------------------- */
enum ImageKind { IMG_FRAME = 0, IMG_RESIZE = 1, IMG_PYRDOWN = 2 };

struct PyrImage {
	int kind;        /* ImageKind */
	int src;         /* index of the source image in the plan (-1: the input frame) */
	int width, height;
	int pitch;       /* row pitch in bytes: a multiple of 16 inside the arena (TMA), the frame width for IMG_FRAME */
	int64_t offset;  /* byte offset inside the per-frame arena (IMG_FRAME: unused) */
	int octave, down;/* i and j of ImagePyramid::createLayers */
	double scale;    /* theoretical scale factor */
	bool kept;       /* becomes an ImagePyramidLayer */
	int layer_index; /* i + j * octaveLayerCount */
	int level;       /* dependency depth: resize = 0, pyrDown = down */
};

struct PlanLayer {
	int image;          /* index into images */
	int index;          /* ImagePyramidLayer index */
	double scale;
	int width, height;
	int orig_patch_w, orig_patch_h;
	int begin_x, begin_y;     /* first window corner (ROI scaled) */
	int windows_x, windows_y;
	int64_t first_window;
};

struct Plan {
	int width = 0, height = 0;
	int octave_layer_count = 0;
	double incremental_scale_factor = 0;
	double min_scale = 0, max_scale = 0;
	std::vector<PyrImage> images; /* dependency order */
	std::vector<PlanLayer> layers; /* sorted by index */
	int64_t arena_bytes = 0;      /* per frame */
	int64_t windows = 0;          /* per frame, whole-image scan */
	int max_down = 0;
};

int cv_round(double v);

/* ImagePyramid(double,double,double) + createLayers sizing (ImagePyramid.cpp:79-92,170-198);
 * returns FDB_OK or FDB_ERR_INVALID_ARGUMENT (message set) */
int build_plan(const fdb_detector_desc& d, int width, int height, Plan* out);

/* DirectPyramidFeatureExtractor::extract window grid for a ROI (all-zero = whole image);
 * fills begin/windows/first_window of every layer, returns the window count */
int64_t enumerate_windows(Plan* plan, int patch_w, int patch_h, int step_x, int step_y,
		int roi_x, int roi_y, int roi_w, int roi_h);

/* ---- device-side tables ----------------------------------------------------------------- */
struct ResizeJob {       /* one cv::resize target */
	int dst_w, dst_h;
	int dst_pitch;
	int64_t dst_offset;  /* arena offset */
	int xtab, ytab;      /* offsets into the coefficient tables */
	int area2x;          /* 1: exact 2x decimation (INTER_AREA fast path) */
	int words_ok;        /* 1: aligned-word source path usable (W % 4 == 0, every quad spans <= 11 source bytes) */
};

struct DownJob {         /* one cv::pyrDown */
	int src_w, src_h, dst_w, dst_h;
	int src_pitch, dst_pitch;
	int64_t src_offset, dst_offset; /* arena offsets; src_offset < 0: source is the input frame */
};

struct DevLayer {        /* per layer, read by the window kernels */
	int64_t offset;      /* arena offset, or -1 when the layer is the frame itself */
	int width, height;
	int pitch;           /* row pitch in bytes */
	int begin_x, begin_y;
	int windows_x, windows_y;
	int first_window;
	int tma_ok;          /* a TMA tensor map exists for this layer (arena image, 16-byte pitch) */
};

#define FDB_MAX_LAYERS 64

/* candidate record written by the stage-1 kernel for WVM-positive windows */
struct Candidate {
	int32_t window;
	int32_t level;
	float fout;
	int32_t frame; /* frame index inside the launch */
};

/* ---- host post-processing (hostpost.cpp) ---------------------------------------------------- */
void stable_sort_desc(std::vector<fdb_detection>& v);
void overlap_eliminate(std::vector<fdb_detection>& v, float dist, float ratio);
/* grid NMS + final ordering of FiveStageSlidingWindowDetector::detect(Mat) on one frame's
 * SVM-positive patches (probability already set to what the reference stores) */
void five_stage_nms(std::vector<fdb_detection>& v, int width, int height);
double wvm_probability(double logistic_a, double logistic_b, float fout);
double svm_probability(double logistic_a, double logistic_b, double distance);
double rvm_probability(double logistic_a, double logistic_b, double distance);

} // namespace fdb

#endif
