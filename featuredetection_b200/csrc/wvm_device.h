/*
 * wvm_device.h - device-side view of the classifiers and launcher prototypes.
 */
#ifndef FDB_WVM_DEVICE_H_
#define FDB_WVM_DEVICE_H_

#include <cuda_runtime.h>
#include <cstdint>

#include "fdb_internal.h"

namespace fdb {

#define WVM_THREADS 128
#define FDB_MAX_FILTERS 512   /* hk_kernel_eval capacity per window (e.g. 280 used) */
#define FDB_MAX_PER_LEVEL 64  /* u_kernel_eval capacity (numFiltersPerLevel, e.g. 14..30) */
#define FDB_MAX_VALUES 8      /* grey values v >= 1 per filter (cntval - 1) */
#ifndef WVM_KA
#define WVM_KA 8              /* filters evaluated by the window kernel before a survivor is queued for wvm_deep_kernel */
#endif

/* a window that survived the first WVM_KA filters (state of WvmClassifier::computeHyperplaneDistance so far) */
struct DeepRec {
	int frame, window;
	int image, x, y;   /* group path: pyramid image and window corner (the deep kernel re-equalises the window from the image) */
	float total_f, sum_xx;
	float hk[WVM_KA];
	float u[WVM_KA];
};

struct DeepQueue {
	int* count;        /* device counter (may run past cap) */
	int* next;         /* work-distribution cursor of wvm_deep_warp_kernel */
	int cap;
	DeepRec* rec;      /* [cap]; nullptr disables the queue */
	uint32_t* patch;   /* [nwords][cap] equalised patch words (generic kernels only) */
};

/* WvmClassifier state in evaluator form; all pointers are device memory */
struct DevWvm {
	int fsx, fsy, nwords;
	int num_lin, per_level, num_used;
	int step_x, step_y;             /* window step of the launching detector */
	float basis_param;
	const float* lin_thresholds;    /* [num_lin] */
	const float* hk_weights;        /* packed triangle */
	const double* app_rsv_convol;   /* [num_lin] */
	const float* thresholds;        /* hierarchicalThresholds incl. limitReliabilityFilter */
	const int* cntval;              /* [num_lin] */
	const int* val_off;             /* [num_lin] */
	const double* val;              /* grey values */
	const uint32_t* masks;          /* rectangle coverage counts, 4 pixels per word: filter f at mask_off[f], [nwords][cntval-1] */
	const int* mask_off;            /* [num_lin] */
	const uint4* bfrag;             /* masks of the first WVM_KA filters as mma.m16n8k32 B fragments: [k-step][lane][2], one k-step per
	                                 * patch row padded to 32 bytes (two rows for 16-wide windows) - wvm_group.cu; nullptr if unavailable */
	const uint8_t* btc;             /* the same coverage counts as UMMA core matrices for wvm_group_tc.cu: [k-step][1 KB] = [4 groups of 8 columns]
	                                 * [2 chunks of 16 operand bytes][8 columns][16 bytes], column = 4 * filter + grey value */
	const float* hk_weights_t;      /* hkWeights of the deep kernel's rounds of 32 filters, transposed: round r at hk_t_off[r] (float4
	                                 * units), [p / 4][lane][4] = w[WVM_KA + 32 r + lane][p .. p + 3] (0 beyond the triangle) */
	const int* hk_t_off;
	const uint2* rects;             /* rectangles of all filters: {x1 | y1 << 8 | x2 << 16 | y2 << 24, grey value index v - 1} */
	const int* rect_off;            /* [num_lin + 1] first rectangle of filter l */
};

/* strips of the window grid (wvm_group.cu): `cols` adjacent window columns x `nsub` runs of <= WVM_RUN window rows */
#ifndef WVM_RUN
#define WVM_RUN 12   /* longest run of window rows one lane walks down */
#endif
#define WVM_MAXSUB 4
#define STRIP_TILE_ROWS 42 /* rows of a warp's bin tile: nsub * run + patch_h - 1 <= 42 */

/* SvmClassifier (RBF) state; support vectors transposed to [word][sv] for coalesced reads */
struct DevSvm {
	int num_sv, dim, nwords, sv_type;
	int kernel;                     /* fdb_kernel_kind */
	double gamma;
	double poly_alpha, poly_constant; int poly_degree;
	float bias, threshold;
	const uint32_t* sv_words;       /* u8: [nwords][num_sv] packed 4 px per word */
	const float* sv_f32;            /* f32: [dim][num_sv] */
	const float* coef;              /* [num_sv]; RVM: the diagonal coefficients c[l][l] */
	int rvm_filters;                /* > 0: RvmClassifier cascade over the first rvm_filters vectors (numFiltersToUse) */
	const float* rvm_thresholds;    /* [num_sv] hierarchicalThresholds */
};

/* tensor-core form of an u8 RBF SVM (svm_dense.cu): support vectors as UMMA core matrices, |sv|^2, float64
 * coefficients, the exp table; all pointers device memory */
struct DevSvmDense {
	int num_sv_pad;            /* support vectors padded to a multiple of 256 (zero vectors with coefficient 0) */
	int dim, chunks;           /* elements per vector; 16-byte k chunks per row (2 * ceil(dim / 32), zero padded) */
	int shift, tab_n;          /* ssd = hi << shift | lo; exp_tab[hi] = exp(-gamma * (hi << shift)), tab_n entries */
	int clamp;                 /* 1: the table ends at the underflow point (hi is clamped to the last, zero, entry) */
	float threshold;
	double neg_bias;
	double poly[4];            /* (-gamma)^k / k!, k = 1..4: exp(-gamma * lo) */
	const uint8_t* b_blocks;   /* [n block of 128][k block]{[16 row groups][chunks in block][8 rows][16 bytes]} */
	const int* ssq;            /* [num_sv_pad] */
	const double* coef;        /* [num_sv_pad] */
	const double* exp_tab;     /* [tab_n] */
};
struct DensePositive {        /* a row (window of the batch) at or above the threshold */
	int64_t row;
	double distance;
};
}
#include <vector>
namespace fdb {
struct SvmDenseHost {
	DevSvmDense dev{};
	std::vector<uint8_t> b_blocks;
	std::vector<int> ssq;
	std::vector<double> coef, tab;
};
#define FDB_SVM_DENSE_MIN_VECTORS 128 /* fdb_svm_get_probability: batches from this size use the tensor-core kernel */
int svm_dense_configure();
bool svm_dense_enabled(); /* false when FDB_SVM_DENSE=0 (debugging: forces the per-window kernel) */
bool svm_dense_build(const uint8_t* sv, const float* coef, int num_sv, int dim, double gamma, float bias, float threshold,
		SvmDenseHost* out);
void launch_svm_dense_windows(cudaStream_t st, const DevSvmDense& s, int patch_w, int patch_h, int step_x, int step_y,
		const uint8_t* frames, int W, int H, int n_frames, const uint8_t* arena, int64_t arena_stride, const DevLayer* layers,
		int n_layers, int64_t windows_per_frame, double* distance_out, int* pos_count, DensePositive* pos, int pos_cap);
void launch_svm_dense_vectors(cudaStream_t st, const DevSvmDense& s, const uint8_t* vectors, int64_t n, double* distance_out);

int wvm_configure();
size_t wvm_smem_bytes(const DevWvm& m);
void launch_wvm_windows(cudaStream_t st, const DevWvm& m, const uint8_t* frames, int W, int H, int n_frames,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, int n_layers, int windows_per_frame,
		fdb_window_score* dense, uint8_t* patches_out, Candidate* cand, int* cand_count, int cand_cap, const DeepQueue& q);
void launch_wvm_patches(cudaStream_t st, const DevWvm& m, const uint8_t* patches, int n, fdb_window_score* dense);

/* fast path (wvm_group.cu): whole-image scans with step 1 and one of the ffpDetectApp window sizes */
#define STRIP_TILE_PITCH 64
void launch_resize(cudaStream_t st, const uint8_t* frames, int W, int H, int n_frames, uint8_t* arena,
		int64_t arena_stride, const ResizeJob* jobs_dev, int n_jobs, int max_tiles, const int4* xy_tab);
int resize_tiles(int dst_w, int dst_h);
void launch_pyrdown(cudaStream_t st, const uint8_t* frames, int W, int H, int n_frames, uint8_t* arena,
		int64_t arena_stride, const DownJob* jobs_dev, int n_jobs, int max_tiles);
int pyrdown_tiles(int dst_w, int dst_h);
void launch_bgr2gray(cudaStream_t st, const uint8_t* bgr, uint8_t* gray, int64_t n_px);

/* one SVM work item: a window of a frame (geometry resolved on the host) */
struct SvmItem {
	int frame;
	int layer;   /* index into the DevLayer table */
	int x, y;    /* window corner inside the layer */
};
int svm_configure();
void launch_svm_windows(cudaStream_t st, const DevSvm& s, int patch_w, int patch_h, const uint8_t* frames, int W, int H,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, const SvmItem* items, int n_items,
		double* distance_out, int* level_out = nullptr /* RVM: level reached */);
void launch_svm_vectors(cudaStream_t st, const DevSvm& s, const void* vectors, int n, double* distance_out, int* level_out = nullptr);
void launch_hq64_items(cudaStream_t st, int patch_w, int patch_h, const uint8_t* frames, int W, int H, const uint8_t* arena,
		int64_t arena_stride, const DevLayer* layers, const SvmItem* items, int n_items, uint8_t* out);

} // namespace fdb
#endif
