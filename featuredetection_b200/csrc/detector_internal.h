/*
 * detector_internal.h - the detector object and the pipeline pieces shared by detector.cu (one detector) and
 * detector_set.cu (all detectors of an application over shared pyramids).
 */
#ifndef FDB_DETECTOR_INTERNAL_H_
#define FDB_DETECTOR_INTERNAL_H_

#include <cuda.h>
#include <cuda_runtime.h>

#include <vector>

#include "fdb_internal.h"
#include "wvm_device.h"
#include "api_types.h"
#include "features_device.h"
#include "wvm_group.h"

#define PIPE_SLOTS 4
#define FDB_NCOUNTERS 8 /* per-slot device counters: [0] candidates, [1] deep queue, [2] deep cursor, [3] group-kernel cursor */
#define OPT_CAND 4096 /* smallest candidate list capacity; det->opt_cand (<= 65536) candidates are fetched together with the counters (one D2H in stream order), more need a second copy */
#define FEAT_BATCH 8192 /* feature vectors materialised at a time (feature-space SVM stage) */

namespace fdb {

struct Slot {
	cudaStream_t st = nullptr;
	cudaStream_t st_copy = nullptr; /* late D2H of long candidate lists: must not queue behind SVM kernels on `st` */
	cudaEvent_t ev_stage1 = nullptr, ev_svm = nullptr;
	uint8_t* d_frames = nullptr;
	uint8_t* d_arena = nullptr;
	const uint8_t* arena = nullptr; /* pyramid arena of the chunk in flight: d_arena, or the shared arena of a detector set */
	int64_t arena_stride = 0;
	CUtensorMap* d_tmaps = nullptr; /* one TMA descriptor per pyramid layer of this slot's arena (strip kernel) */
	fdb_window_score* d_dense = nullptr;
	int* d_counters = nullptr;     /* FDB_NCOUNTERS counters (see above), followed by the candidate list */
	Candidate* d_cand = nullptr;   /* = (Candidate*)(d_counters + FDB_NCOUNTERS) */
	DeepQueue deep{};
	SvmItem* d_items = nullptr;
	double* d_dist = nullptr;
	uint8_t* d_farena = nullptr;   /* filtered pyramid layers of the chunk (feature spaces with layer filters) */
	void* d_feat = nullptr;        /* FEAT_BATCH feature vectors */
	int* h_counters = nullptr;     /* pinned mirror: the counters + OPT_CAND candidates */
	Candidate* h_cand_big = nullptr; /* pinned, cand_cap entries (second copy when > OPT_CAND) */
	SvmItem* h_items = nullptr;
	double* h_dist = nullptr;
	/* state of the chunk in flight */
	int n = 0, base = 0;
	const Candidate* cand_src = nullptr; /* host copy of the chunk's candidate list (valid after phase_a_fetch[_wait]) */
	bool cand_pending = false;           /* a late copy of the list is in flight on st_copy */
	const uint8_t* frames_dev = nullptr;
	std::vector<std::vector<fdb_detection>> per_frame;
	size_t svm_items = 0;
	bool busy = false;
};

/* cv::resize / cv::pyrDown jobs of a set of pyramid images on the device */
struct PyramidJobs {
	ResizeJob* d_resize = nullptr; int n_resize = 0; int max_quads = 0;
	std::vector<DownJob*> d_down; std::vector<int> n_down; std::vector<int> max_down_px;
	int4* d_xy_tab = nullptr; /* bilinear tables: {source offset, a0, a1, word-path info} */
};
} // namespace fdb

using namespace fdb; /* internal header: only detector.cu and detector_set.cu include it */

struct fdb_detector {
	fdb_ctx* ctx = nullptr;
	fdb_detector_desc desc{};
	fdb_wvm* wvm = nullptr;
	fdb_svm* svm = nullptr;
	Plan plan;
	bool prepared = false;
	uint8_t* d_bgr = nullptr;      /* fdb_detect_batch_bgr staging: interleaved frames and their gray conversion */
	uint8_t* d_gray = nullptr;
	int64_t bgr_cap_px = 0;
	int max_batch = 0, chunk = 0, n_slots = 0;
	int cand_cap = 0, items_cap = 0;
	int opt_cand = OPT_CAND;          /* candidates copied with the counters */
	std::vector<void*> owned, owned_host;
	Slot slots[PIPE_SLOTS];
	cudaEvent_t ev_begin = nullptr;
	uint8_t* d_patches = nullptr; int64_t d_patches_bytes = 0;
	DevLayer* d_layers = nullptr;     /* whole-image scan */
	DevLayer* d_layers_roi = nullptr; /* scratch table for ROI scans */
	fdb::PyramidJobs jobs;            /* cv::resize / cv::pyrDown job tables of this detector's pyramid */
	GroupItem* d_gitems = nullptr; int n_gitems = 0; /* strips of the whole-image scan (group kernel work items, one model) */
	GroupImage* d_gimages = nullptr;                 /* image table of the group kernels: entry li = image of layer li */
	bool use_tma = false;             /* strip tiles staged by TMA (tensor maps encoded) */
	bool use_strips = false;          /* fast path usable (and not yet overflowed) */
	bool has_feature = false;         /* the SVM works in its own feature space (fdb_detector_set_feature) */
	fdb_feature_desc fdesc{};
	DevFeature feat{};
	int64_t farena_bytes = 0;         /* per frame */
	SvmItem* d_all_items = nullptr;   /* every window of a frame as an SVM item (`single` detector without a WVM) */
	double* d_all_dist = nullptr; double* h_all_dist = nullptr;
	int* d_all_level = nullptr; int* h_all_level = nullptr; /* RVM (`single` prvm): level reached per window */
	/* `single` detector on the tensor cores (svm_dense.cu): distances of a chunk, positives list */
	double* d_sd_dist = nullptr; int* d_sd_count = nullptr; DensePositive* d_sd_pos = nullptr;
	int* h_sd_count = nullptr; DensePositive* h_sd_pos = nullptr; int sd_pos_cap = 0;
	/* fdb_evaluate_samples scratch (grow-only; the tracker calls it every frame) */
	std::vector<void*> es_owned; int es_cap = 0;
	SvmItem* d_es_items = nullptr; uint8_t* d_es_patches = nullptr; fdb_window_score* d_es_scores = nullptr; double* d_es_dist = nullptr;
	double sd_kernel_ms = 0; int sd_kernel_launches = 0; /* svm_dense_kernel time of the last call (CUDA events) */
	int64_t counts[5] = {0, 0, 0, 0, 0};
};

namespace fdb {

int build_pyramid_jobs(const std::vector<PyrImage>& images, int max_down, int width, int height, PyramidJobs* out, std::vector<void*>& owned);
/* resize + pyrDown launches for n frames; ev_mid (optional) is recorded between the two */
int enqueue_pyramid(fdb_ctx* c, cudaStream_t st, const PyramidJobs& jobs, const uint8_t* d_frames, int W, int H, int n, uint8_t* d_arena,
		int64_t arena_stride, cudaEvent_t ev_mid = nullptr);
/* TMA descriptors of the group kernel's tiles: image k of `images` inside `arena` ({width, height, frames} u8, strides {pitch,
 * arena_stride}); entries of the frame itself stay zero. false: the driver entry point is missing */
bool encode_tile_maps(const std::vector<PyrImage>& images, const std::vector<int>& which, uint8_t* arena, int64_t arena_stride, int frames,
		std::vector<CUtensorMap>* out);
void append_strip_items(const PlanLayer& L, int image, int patch_h, int n_models, const int* models, const int* first_windows, int max_pack,
		std::vector<GroupItem>* out);
void fill_detection(fdb_detection* d, const Plan& plan, const fdb_detector_desc& desc, int frame, int64_t window);
int copy_out(const std::vector<fdb_detection>& dets, fdb_detection* out, int64_t cap, int64_t* n_out);
#define STATUS_REDO (-1000) /* internal: the fast path overflowed, run the call again on the generic kernels */
/* phase A of the host post-processing: wait for stage 1 of the slot's chunk, per-frame candidate lists, overlap
 * elimination, SVM launch on the survivors (async); phase B: SVM distances -> classify, grid NMS, append in frame order */
int phase_a(fdb_detector* det, Slot& sl, cudaStream_t st, const Plan& plan, const DevLayer* d_layers, int stage,
		int fast_path = -1 /* 1: stage 1 ran on the group kernels (deep-queue overflow -> STATUS_REDO); -1: decide from d_layers */);
int phase_a_fetch(fdb_detector* det, Slot& sl, cudaStream_t st, const DevLayer* d_layers, int fast_path);
int phase_a_fetch_wait(Slot& sl, cudaStream_t st);
int phase_a_host(fdb_detector* det, Slot& sl, const Plan& plan, int stage); /* pure CPU, thread-safe per (det, sl); FDB_OK or FDB_ERR_OVERFLOW */
int phase_a_launch(fdb_detector* det, Slot& sl, cudaStream_t st, const Plan& plan, const DevLayer* d_layers, int stage);
void phase_b_host(fdb_detector* det, Slot& sl, const Plan& plan, int stage, bool is_roi, std::vector<fdb_detection>& out); /* pure CPU */
int detect_impl(fdb_detector* det, const uint8_t* frames, bool frames_on_device, int64_t pitch, int32_t n_frames,
		int32_t stage, fdb_window_score* dense_out, bool dense_on_device, fdb_detection* dets_out, int64_t det_cap,
		int64_t* n_dets);
int phase_b(fdb_detector* det, Slot& sl, const Plan& plan, int stage, bool is_roi, std::vector<fdb_detection>& out);

} // namespace fdb

#endif
