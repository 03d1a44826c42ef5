/*
 * aggdet.cu - detection::AggregatedFeaturesDetector on the GPU: gray image pyramid -> FHOG feature
 * map of every layer -> linear-SVM score map of every layer -> windows above the threshold -> boxes in image pixels ->
 * IoU non-maximum suppression.
 *
 *   AggregatedFeaturesDetector::detectWithScores / getPositiveWindows / rescaleWindow   libDetection/src/detection/AggregatedFeaturesDetector.cpp:37-118
 *   AggregatedFeaturesExtractor (scale limits, computeBoundsInImagePixels)              libImageProcessing/src/imageprocessing/extraction/AggregatedFeaturesExtractor.cpp:22-131
 *   ImagePyramid(size_t octaveLayerCount, double min, double max), createLayers          ImagePyramid.cpp:60-76,170-198
 *   FhogFilter (layer filter) + FhogAggregationFilter                                    filtering/FhogFilter.cpp:20-122, FhogFilter.hpp:112-208, FhogAggregationFilter.cpp:43-150
 *   ConvolutionFilter (the SVM's support vector as a correlation kernel, anchor (0,0))   ConvolutionFilter.cpp:31-49
 *   NonMaximumSuppression                                                                libDetection/src/detection/NonMaximumSuppression.cpp:27-112
 *
 * Kernels (all layers of all frames of a chunk per launch; the pyramid comes from pyramid.cu):
 *   aggdet_hist_kernel   a CTA owns a tile of cells of one layer. Phase 1: every pixel of the tile (plus one cell of halo: with
 *                        cell interpolation a pixel feeds the 4 nearest cells) gets its gradient-LUT entry {bin, bin + 1, weights}
 *                        ONCE, into shared memory. Phase 2: one thread per cell walks the pixels that feed its cell in raster
 *                        order and accumulates its signed histogram - float32 sums are order dependent and the reference adds a
 *                        bin's contributions in raster order (FhogFilter.hpp:118-123,170-207), which a per-cell walk reproduces
 *                        exactly for every bin at once. Energy of the cell (FhogAggregationFilter.cpp:60-68) on the way out.
 *   aggdet_desc_kernel   thread = cell: 4 normalisers from the 3 x 3 energies, 3 B + 4 features (fhog_descriptor of fhog_core.h),
 *                        rows padded to a multiple of 4 floats.
 *   aggdet_score_kernel  a CTA owns 8 x 32 score positions of one layer, a thread 4 adjacent ones: 8 channels at a time the
 *                        feature tile (+ kernel halo) and the weights are staged in shared memory; per kernel tap one weight
 *                        load and one new feature serve the thread's 4 positions; score = -bias + sum over channels of the
 *                        correlation in the order of the oracle's restatement (channel, kernel row, kernel column); positions
 *                        above the threshold are appended to the candidate list (atomic cursor), optionally the dense map.
 * Bound: HBM by construction (1 byte per pixel in, (3 B + 4) * 4 / cell^2 bytes per pixel out); the per-cell replay of phase 2
 * costs ~5 warp instructions per pixel, which keeps the kernel issue-bound below that roofline (measured: profiles/).
 * cv::filter2D's own summation order (and its DFT path for kernels of >= 50 elements) belongs to OpenCV: score parity is 1e-4.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "detector_internal.h"
#include "fhog_core.h"

namespace fdb {
void fhog_build_lut(int unsigned_bins, int interpolate_bins, std::vector<FhogLutEntry>* out); /* fhog.cu */
}

using namespace fdb;

namespace {

#define AGG_MAX_LAYERS 96
#define AGG_SCORE_TY 8
#define AGG_SCORE_TX 32

struct AggLayer {            /* one pyramid layer and where its intermediates live (per frame) */
	int64_t img_offset;      /* arena offset of the layer image; < 0: the frame itself */
	int img_pitch, rows, cols;       /* pixels */
	int crow, ccol;                  /* cells */
	int vh, vw;                      /* valid score positions */
	int64_t hist_off, energy_off, feat_off, score_off; /* element offsets into the per-frame arrays */
	int tiles_x, tiles_y, first_tile;       /* histogram tiles (cells) */
	int stiles_x, stiles_y, first_stile;    /* score tiles */
	int index; double scale_x, scale_y;     /* ImagePyramidLayer index, layer size / image size */
};

struct AggParams {
	int cell, ubins, interp_bins, interp_cells;
	int D, Dp;               /* features per cell, padded row length (multiple of 4) */
	int hstride;             /* floats between the histograms of neighbouring cells in shared memory (odd: conflict free) */
	int tc;                  /* cells per histogram tile edge */
	float alpha;
	int kh, kw;
	float bias, threshold;
};

struct AggCandidate { int frame, layer, x, y; float score; };

typedef FhogPix PixEntry; /* fhog_core.h */

__global__ void __launch_bounds__(128) aggdet_hist_kernel(const AggParams P, const AggLayer* __restrict__ layers, const int* __restrict__ tile_layer,
		const FhogLutEntry* __restrict__ lut, const uint8_t* __restrict__ frames, int W, int H, const uint8_t* __restrict__ arena,
		int64_t arena_stride, float* __restrict__ hist, int64_t hist_stride, float* __restrict__ energy, int64_t energy_stride) {
	extern __shared__ __align__(16) uint8_t smem[];
	const int li = tile_layer[blockIdx.x];
	const AggLayer L = layers[li];
	const int frame = blockIdx.y;
	const int tile = blockIdx.x - L.first_tile;
	const int ty = tile / L.tiles_x, tx = tile - ty * L.tiles_x;
	const int cell = P.cell, tc = P.tc;
	const int cr0 = ty * tc, cc0 = tx * tc;                        /* first cell of the tile */
	const int halo = P.interp_cells ? cell : 0;
	const int pr0 = cr0 * cell - halo, pc0 = cc0 * cell - halo;    /* first pixel of the staged region */
	const int region = tc * cell + 2 * halo;                        /* pixels per edge */
	PixEntry* const s_px = reinterpret_cast<PixEntry*>(smem);
	float* const s_hist = reinterpret_cast<float*>(smem + (size_t)region * region * sizeof(PixEntry));
	const uint8_t* __restrict__ img = L.img_offset < 0 ? frames + (int64_t)frame * W * H : arena + (int64_t)frame * arena_stride + L.img_offset;
	const int rows_used = L.crow * cell, cols_used = L.ccol * cell;
	/* phase 1: the gradient-LUT entry of every pixel of the region (FhogFilter.hpp:127-168); gradients look at the neighbours
	 * in the IMAGE (clamped at its border), not in the tile */
	for (int i = threadIdx.x; i < region * region; i += blockDim.x) {
		const int rr = i / region, cc = i - rr * region;
		const int r = pr0 + rr, c = pc0 + cc;
		PixEntry e; e.i1 = 0; e.i2 = 0; e.valid = 0; e.w1 = 0.f; e.w2 = 0.f;
		if (r >= 0 && c >= 0 && r < rows_used && c < cols_used) {
			const FhogLutEntry* q = fhog_pixel_entry(lut, img, L.img_pitch, L.rows, L.cols, 1, r, c);
			e.i1 = (uint8_t)q->index1; e.i2 = (uint8_t)q->index2; e.w1 = q->weight1; e.w2 = q->weight2; e.valid = 1;
		}
		s_px[i] = e;
	}
	const int sb = 2 * P.ubins;
	for (int i = threadIdx.x; i < tc * tc * P.hstride; i += blockDim.x) s_hist[i] = 0.f;
	__syncthreads();
	/* phase 2: thread = cell; pixels that feed the cell, in raster order (addToSignedHistograms, FhogFilter.hpp:170-207) */
	if (threadIdx.x < tc * tc) {
		const int lr = threadIdx.x / tc, lc = threadIdx.x - lr * tc;
		const int cr = cr0 + lr, cc = cc0 + lc;
		if (cr < L.crow && cc < L.ccol) {
			float* const h = s_hist + threadIdx.x * P.hstride;
			fhog_cell_histogram(s_px, pr0, pc0, region, cell, L.crow, L.ccol, P.interp_bins, P.interp_cells, cr, cc, h);
			const int64_t cellidx = (int64_t)cr * L.ccol + cc;
			float* const out = hist + (int64_t)frame * hist_stride + L.hist_off + cellidx * sb;
			for (int b = 0; b < sb; ++b) out[b] = h[b];
			energy[(int64_t)frame * energy_stride + L.energy_off + cellidx] = fhog_energy(h, P.ubins);
		}
	}
}

__global__ void __launch_bounds__(128) aggdet_desc_kernel(const AggParams P, const AggLayer* __restrict__ layers, int n_layers,
		const float* __restrict__ hist, int64_t hist_stride, const float* __restrict__ energy, int64_t energy_stride,
		float* __restrict__ feat, int64_t feat_stride, int64_t cells_per_frame) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cells_per_frame) return;
	const int frame = blockIdx.y;
	int li = 0;
	while (li + 1 < n_layers && i >= layers[li + 1].energy_off) ++li; /* energy_off = first cell of the layer */
	const AggLayer L = layers[li];
	const int local = (int)(i - L.energy_off);
	const int r = local / L.ccol, c = local - r * L.ccol;
	float out[3 * 64 + 4];
	fhog_descriptor(hist + (int64_t)frame * hist_stride + L.hist_off + (int64_t)local * 2 * P.ubins,
			energy + (int64_t)frame * energy_stride + L.energy_off, L.crow, L.ccol, r, c, P.ubins, P.alpha, out);
	float* dst = feat + (int64_t)frame * feat_stride + L.feat_off + (int64_t)local * P.Dp;
	for (int k = 0; k < P.D; ++k) dst[k] = out[k];
	for (int k = P.D; k < P.Dp; ++k) dst[k] = 0.f;
}

#define AGG_SCORE_XB 4   /* adjacent score positions per thread: one weight load and a sliding feature window serve 4 positions */
#define AGG_SCORE_CB 8   /* channels staged per pass */
#define AGG_SCORE_THREADS (AGG_SCORE_TY * AGG_SCORE_TX / AGG_SCORE_XB)

__global__ void __launch_bounds__(AGG_SCORE_THREADS) aggdet_score_kernel(const AggParams P, const AggLayer* __restrict__ layers,
		const int* __restrict__ stile_layer, const float* __restrict__ feat, int64_t feat_stride, const float* __restrict__ weights,
		float* __restrict__ scores /* nullable */, int64_t score_stride, AggCandidate* __restrict__ cand, int* __restrict__ cand_count, int cand_cap) {
	extern __shared__ __align__(16) float sf[];
	const int li = stile_layer[blockIdx.x];
	const AggLayer L = layers[li];
	const int frame = blockIdx.y;
	const int tile = blockIdx.x - L.first_stile;
	const int ty0 = (tile / L.stiles_x) * AGG_SCORE_TY, tx0 = (tile % L.stiles_x) * AGG_SCORE_TX;
	const int th = AGG_SCORE_TY + P.kh - 1, tw = AGG_SCORE_TX + P.kw - 1;
	constexpr int CS = AGG_SCORE_CB + 1;              /* odd row length in shared memory: neighbouring positions hit different banks */
	float* const s_w = sf;                            /* [kh][kw][CB] weights of the pass */
	float* const s_f = sf + P.kh * P.kw * AGG_SCORE_CB; /* [th][tw][CS] features of the pass */
	const float* __restrict__ f = feat + (int64_t)frame * feat_stride + L.feat_off;
	const int ly = threadIdx.x / (AGG_SCORE_TX / AGG_SCORE_XB), lx = (threadIdx.x % (AGG_SCORE_TX / AGG_SCORE_XB)) * AGG_SCORE_XB;
	/* ConvolutionFilter::applyTo as AggregatedFeaturesDetector configures it (ConvolutionFilter.cpp:31-49): delta = -bias, then
	 * channel by channel the correlation with that channel's kernel, anchor (0, 0) */
	float score[AGG_SCORE_XB];
#pragma unroll
	for (int k = 0; k < AGG_SCORE_XB; ++k) score[k] = -P.bias;
	for (int c0 = 0; c0 < P.D; c0 += AGG_SCORE_CB) {
		const int nc = min(AGG_SCORE_CB, P.D - c0);
		__syncthreads(); /* the previous pass is consumed */
		for (int i = threadIdx.x; i < P.kh * P.kw * AGG_SCORE_CB; i += blockDim.x) {
			const int c = i % AGG_SCORE_CB, tap = i / AGG_SCORE_CB;
			s_w[i] = c < nc ? weights[tap * P.D + c0 + c] : 0.f;
		}
		for (int i = threadIdx.x; i < th * tw * (AGG_SCORE_CB / 4); i += blockDim.x) {
			const int q = i % (AGG_SCORE_CB / 4), cellidx = i / (AGG_SCORE_CB / 4);
			const int yy = cellidx / tw, xx = cellidx - yy * tw;
			const int y = ty0 + yy, x = tx0 + xx;
			float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
			if (y < L.crow && x < L.ccol && c0 + 4 * q < P.Dp) v = *reinterpret_cast<const float4*>(f + ((int64_t)y * L.ccol + x) * P.Dp + c0 + 4 * q);
			float* d = s_f + (yy * tw + xx) * CS + 4 * q;
			d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
		}
		__syncthreads();
		for (int c = 0; c < nc; ++c) {
			float tmp[AGG_SCORE_XB];
#pragma unroll
			for (int k = 0; k < AGG_SCORE_XB; ++k) tmp[k] = 0.f;
			for (int i = 0; i < P.kh; ++i) {
				const float* row = s_f + ((ly + i) * tw + lx) * CS + c;
				const float* wrow = s_w + i * P.kw * AGG_SCORE_CB + c;
				float win[AGG_SCORE_XB]; /* features under tap j of the 4 positions: a window sliding with j */
#pragma unroll
				for (int k = 0; k < AGG_SCORE_XB - 1; ++k) win[k + 1] = row[k * CS];
				for (int j = 0; j < P.kw; ++j) {
#pragma unroll
					for (int k = 0; k < AGG_SCORE_XB - 1; ++k) win[k] = win[k + 1];
					win[AGG_SCORE_XB - 1] = row[(j + AGG_SCORE_XB - 1) * CS];
					const float w = wrow[j * AGG_SCORE_CB];
#pragma unroll
					for (int k = 0; k < AGG_SCORE_XB; ++k) tmp[k] = FHOG_ADD(tmp[k], FHOG_MUL(win[k], w));
				}
			}
#pragma unroll
			for (int k = 0; k < AGG_SCORE_XB; ++k) score[k] = FHOG_ADD(score[k], tmp[k]);
		}
	}
	const int y = ty0 + ly;
	if (y >= L.vh) return;
#pragma unroll
	for (int k = 0; k < AGG_SCORE_XB; ++k) {
		const int x = tx0 + lx + k;
		if (x >= L.vw) continue;
		if (scores) scores[(int64_t)frame * score_stride + L.score_off + (int64_t)y * L.vw + x] = score[k];
		if (score[k] > P.threshold) { /* AggregatedFeaturesDetector.cpp:100 */
			const int slot = atomicAdd(cand_count, 1);
			if (slot < cand_cap) { AggCandidate q; q.frame = frame; q.layer = li; q.x = x; q.y = y; q.score = score[k]; cand[slot] = q; }
		}
	}
}

} // namespace

struct fdb_aggdet {
	fdb_ctx* ctx = nullptr;
	fdb_aggdet_desc desc{};
	std::vector<float> weights;
	AggParams P{};
	bool prepared = false;
	int W = 0, H = 0, max_batch = 0, chunk = 0;
	Plan plan;
	PyramidJobs jobs;
	std::vector<AggLayer> layers;
	int n_tiles = 0, n_stiles = 0;
	int64_t hist_stride = 0, energy_stride = 0, feat_stride = 0, score_stride = 0;
	size_t hist_smem = 0, score_smem = 0;
	int cand_cap = 0;
	std::vector<void*> owned, owned_host;
	FhogLutEntry* d_lut = nullptr;
	float* d_weights = nullptr;
	AggLayer* d_layers = nullptr; int* d_tile_layer = nullptr; int* d_stile_layer = nullptr;
	uint8_t* d_frames = nullptr; uint8_t* d_arena = nullptr;
	float* d_hist = nullptr; float* d_energy = nullptr; float* d_feat = nullptr; float* d_scores = nullptr;
	AggCandidate* d_cand = nullptr; int* d_count = nullptr;
	AggCandidate* h_cand = nullptr; int* h_count = nullptr;
	int64_t last_candidates = 0, last_positions = 0;
};

namespace {

void agg_release(fdb_aggdet* d) {
	free_all(d->owned, &d->owned_host);
	d->jobs = PyramidJobs();
	d->layers.clear();
	d->prepared = false;
}

/* AggregatedFeaturesExtractor's scale limits for an image (AggregatedFeaturesExtractor.cpp:46-51,60-73) */
void agg_scale_limits(const fdb_aggdet_desc& a, int W, int H, double* inc, double* mn, double* mx) {
	*inc = std::pow(0.5, 1.0 / a.octave_layer_count);                 /* ImagePyramid.cpp:75 */
	const int patch_w = a.window_cols * a.cell_size, patch_h = a.window_rows * a.cell_size;
	*mx = 1.0;
	if (a.min_window_width > patch_w) {
		const double ms = (double)patch_w / a.min_window_width;
		*mx = std::pow(*inc, (int)std::ceil(std::log(ms) / std::log(*inc)));
	}
	const double aspect = (double)patch_h / (double)patch_w, image_aspect = (double)H / (double)W;
	const int max_width = aspect > image_aspect ? (int)(H / aspect) : W;
	const double m = (double)patch_w / max_width;
	*mn = std::pow(*inc, (int)(std::log(m) / std::log(*inc)));
}

/* pyramid + FHOG + score map of n frames resident at d_frames; candidates are left on the device. ev: 5 marks or null */
int agg_enqueue(fdb_aggdet* d, const uint8_t* d_frames, int n, bool want_scores, cudaEvent_t* ev) {
	fdb_ctx* c = d->ctx;
	cudaStream_t st = c->stream;
	CUDA_TRY(cudaMemsetAsync(d->d_count, 0, sizeof(int), st));
	if (ev) CUDA_TRY(cudaEventRecord(ev[0], st));
	{ const int r = enqueue_pyramid(c, st, d->jobs, d_frames, d->W, d->H, n, d->d_arena, d->plan.arena_bytes); if (r) return r; }
	if (ev) CUDA_TRY(cudaEventRecord(ev[1], st));
	const AggParams& P = d->P;
	if (d->n_tiles) {
		aggdet_hist_kernel<<<dim3((unsigned)d->n_tiles, (unsigned)n), 128, d->hist_smem, st>>>(P, d->d_layers, d->d_tile_layer, d->d_lut, d_frames,
				d->W, d->H, d->d_arena, d->plan.arena_bytes, d->d_hist, d->hist_stride, d->d_energy, d->energy_stride);
		c->launches++;
		if (ev) CUDA_TRY(cudaEventRecord(ev[2], st));
		aggdet_desc_kernel<<<dim3((unsigned)((d->energy_stride + 127) / 128), (unsigned)n), 128, 0, st>>>(P, d->d_layers, (int)d->layers.size(),
				d->d_hist, d->hist_stride, d->d_energy, d->energy_stride, d->d_feat, d->feat_stride, d->energy_stride);
		c->launches++;
	} else if (ev) CUDA_TRY(cudaEventRecord(ev[2], st));
	if (ev) CUDA_TRY(cudaEventRecord(ev[3], st));
	if (d->n_stiles) {
		aggdet_score_kernel<<<dim3((unsigned)d->n_stiles, (unsigned)n), AGG_SCORE_THREADS, d->score_smem, st>>>(P, d->d_layers,
				d->d_stile_layer, d->d_feat, d->feat_stride, d->d_weights, want_scores ? d->d_scores : nullptr, d->score_stride, d->d_cand,
				d->d_count, d->cand_cap);
		c->launches++;
	}
	if (ev) CUDA_TRY(cudaEventRecord(ev[4], st));
	CUDA_TRY(cudaGetLastError());
	return FDB_OK;
}

int agg_detect(fdb_aggdet* d, const uint8_t* frames, bool on_device, int64_t pitch, int32_t n_frames, float* scores_out, int32_t* rects_out,
		int32_t* frame_out, int64_t cap, int64_t* n_out) {
	if (!d || !d->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "aggregated-features detector not prepared (call fdb_aggdet_prepare)");
	int s = check_ctx(d->ctx); if (s) return s;
	if (n_frames < 0 || (n_frames > 0 && !frames) || !n_out || cap < 0 || (cap > 0 && (!scores_out || !rects_out)))
		return fail(FDB_ERR_INVALID_ARGUMENT, "bad arguments");
	if (!on_device && pitch < d->W) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	cudaStream_t st = d->ctx->stream;
	const fdb_aggdet_desc& a = d->desc;
	int64_t total = 0;
	d->last_candidates = 0;
	d->last_positions = d->score_stride * n_frames;
	std::vector<float> sc; std::vector<int32_t> rc;
	for (int base = 0; base < n_frames; base += d->chunk) {
		const int n = std::min(d->chunk, n_frames - base);
		const uint8_t* df = frames + (int64_t)base * d->W * d->H;
		if (!on_device) {
			CUDA_TRY(cudaMemcpy2DAsync(d->d_frames, (size_t)d->W, frames + (int64_t)base * pitch * d->H, (size_t)pitch, (size_t)d->W,
					(size_t)d->H * n, cudaMemcpyHostToDevice, st));
			df = d->d_frames;
		}
		s = agg_enqueue(d, df, n, false, nullptr); if (s) return s;
		CUDA_TRY(cudaMemcpyAsync(d->h_count, d->d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		const int nc = d->h_count[0];
		if (nc > d->cand_cap) return fail(FDB_ERR_OVERFLOW, "more windows above the threshold than the candidate list holds (raise the threshold)");
		if (nc > 0) {
			CUDA_TRY(cudaMemcpyAsync(d->h_cand, d->d_cand, sizeof(AggCandidate) * (size_t)nc, cudaMemcpyDeviceToHost, st));
			CUDA_TRY(cudaStreamSynchronize(st));
		}
		d->last_candidates += nc;
		/* the reference visits layers in pyramid order, then y, then x (AggregatedFeaturesDetector.cpp:94-98) */
		std::sort(d->h_cand, d->h_cand + nc, [](const AggCandidate& p, const AggCandidate& q) {
			if (p.frame != q.frame) return p.frame < q.frame;
			if (p.layer != q.layer) return p.layer < q.layer;
			return p.y != q.y ? p.y < q.y : p.x < q.x; });
		int i = 0;
		for (int f = 0; f < n; ++f) {
			sc.clear(); rc.clear();
			for (; i < nc && d->h_cand[i].frame == f; ++i) {
				const AggCandidate& k = d->h_cand[i];
				const AggLayer& L = d->layers[(size_t)k.layer];
				/* computeBoundsInImagePixels (AggregatedFeaturesExtractor.cpp:121-128): std::round = half away from zero */
				const int bx = (int)std::round((double)(k.x * a.cell_size) / L.scale_x), by = (int)std::round((double)(k.y * a.cell_size) / L.scale_y);
				const int bw = (int)std::round((double)(a.window_cols * a.cell_size) / L.scale_x), bh = (int)std::round((double)(a.window_rows * a.cell_size) / L.scale_y);
				const int cx = bx + bw / 2, cy = by + bh / 2;                              /* Patch::computeCenter */
				const int rw = (int)(a.width_scale * bw), rh = (int)(a.height_scale * bh); /* Size(float, float) -> Size_<int> truncates */
				sc.push_back(k.score);
				rc.push_back(cx - rw / 2); rc.push_back(cy - rh / 2); rc.push_back(rw); rc.push_back(rh); /* Patch::computeBounds */
			}
			int64_t kept = (int64_t)sc.size();
			if (kept) { s = fdb_non_maximum_suppression(sc.data(), rc.data(), kept, a.nms_overlap_threshold, a.nms_type, &kept); if (s) return s; }
			for (int64_t k = 0; k < kept; ++k, ++total) {
				if (total >= cap) continue;
				scores_out[total] = sc[(size_t)k];
				std::memcpy(rects_out + 4 * total, rc.data() + 4 * k, 4 * sizeof(int32_t));
				if (frame_out) frame_out[total] = base + f;
			}
		}
	}
	*n_out = total;
	if (total > cap) return fail(FDB_ERR_OVERFLOW, "detections_out capacity too small");
	return FDB_OK;
}

} // namespace

extern "C" {

int fdb_aggdet_create(fdb_ctx* ctx, const fdb_aggdet_desc* desc, fdb_aggdet** out) try {
	int s = check_ctx(ctx); if (s) return s;
	if (!desc || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	const fdb_aggdet_desc& a = *desc;
	if (a.cell_size < 1 || a.cell_size > 32) return fail(FDB_ERR_INVALID_ARGUMENT, "AggregatedFeaturesDetector: cell size must be in 1..32");
	if (a.window_cols < 1 || a.window_rows < 1 || a.window_cols > 32 || a.window_rows > 32) return fail(FDB_ERR_INVALID_ARGUMENT, "AggregatedFeaturesDetector: window size (cells) must be in 1..32");
	if (a.octave_layer_count < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "ImagePyramid: the number of layers per octave must be greater than zero");
	if (a.unsigned_bins < 1 || a.unsigned_bins > 64) return fail(FDB_ERR_INVALID_ARGUMENT, "FhogFilter: unsignedBinCount must be bigger than zero (at most 64 here)");
	if (!(a.alpha > 0)) return fail(FDB_ERR_INVALID_ARGUMENT, "FhogAggregationFilter: alpha must be bigger than zero");
	if (!a.weights) return fail(FDB_ERR_INVALID_ARGUMENT, "AggregatedFeaturesDetector: the SVM must use a LinearKernel (one support vector = the weights)");
	if (a.nms_type < FDB_NMS_MAX_SCORE || a.nms_type > FDB_NMS_WEIGHTED_AVERAGE) return fail(FDB_ERR_INVALID_ARGUMENT, "NonMaximumSuppression: unsupported maximum type");
	fdb_aggdet* d = new fdb_aggdet;
	d->ctx = ctx; d->desc = a;
	const int D = 3 * a.unsigned_bins + 4;
	d->weights.assign(a.weights, a.weights + (size_t)a.window_rows * a.window_cols * D);
	d->desc.weights = d->weights.data();
	*out = d;
	return FDB_OK;
} FDB_API_CATCH

void fdb_aggdet_destroy(fdb_aggdet* d) {
	if (!d) return;
	cudaSetDevice(d->ctx->device);
	cudaStreamSynchronize(d->ctx->stream);
	agg_release(d);
	delete d;
}

int fdb_aggdet_prepare(fdb_aggdet* d, int32_t width, int32_t height, int32_t max_batch) try {
	if (!d) return fail(FDB_ERR_INVALID_ARGUMENT, "null detector");
	int s = check_ctx(d->ctx); if (s) return s;
	if (max_batch < 1 || width < 1 || height < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "bad frame geometry");
	CUDA_TRY(cudaStreamSynchronize(d->ctx->stream));
	agg_release(d);
	const fdb_aggdet_desc& a = d->desc;
	double inc, mn, mx;
	agg_scale_limits(a, width, height, &inc, &mn, &mx);
	fdb_detector_desc pd{};
	pd.incremental_scale_factor = inc; pd.min_scale_factor = mn; pd.max_scale_factor = mx;
	pd.patch_width = a.window_cols * a.cell_size; pd.patch_height = a.window_rows * a.cell_size; pd.step_x = pd.step_y = 1;
	if (!(mn > 0) || mn > mx) { /* the window does not fit the image at any scale: no layers, no detections */
		d->W = width; d->H = height; d->max_batch = max_batch; d->chunk = 1;
		s = dev_alloc(&d->d_count, 4, d->owned); if (s) return s;
		s = host_alloc(&d->h_count, 4, d->owned_host); if (s) return s;
		d->plan = Plan(); d->plan.width = width; d->plan.height = height; d->plan.arena_bytes = 128;
		d->prepared = true;
		return FDB_OK;
	}
	s = build_plan(pd, width, height, &d->plan); if (s) return s;
	if ((int)d->plan.layers.size() > AGG_MAX_LAYERS) return fail(FDB_ERR_UNSUPPORTED, "more than 96 pyramid layers");
	d->W = width; d->H = height; d->max_batch = max_batch;
	d->chunk = std::min(max_batch, 16);
	AggParams& P = d->P;
	P.cell = a.cell_size; P.ubins = a.unsigned_bins; P.interp_bins = a.interpolate_bins != 0; P.interp_cells = a.interpolate_cells != 0;
	P.D = 3 * a.unsigned_bins + 4; P.Dp = (P.D + 3) / 4 * 4;
	P.hstride = 2 * a.unsigned_bins + 1;
	P.alpha = a.alpha; P.kh = a.window_rows; P.kw = a.window_cols; P.bias = a.bias; P.threshold = a.threshold;
	/* cells per histogram tile edge: the staged pixel region ((tc + 2) cell)^2 entries of 12 bytes must fit 48 KB */
	P.tc = std::max(1, std::min(8, (int)std::floor(std::sqrt(48.0 * 1024 / sizeof(PixEntry)) / a.cell_size) - 2));
	const int region = P.tc * a.cell_size + 2 * a.cell_size;
	d->hist_smem = (size_t)region * region * sizeof(PixEntry) + (size_t)P.tc * P.tc * P.hstride * sizeof(float);
	d->score_smem = ((size_t)P.kh * P.kw * AGG_SCORE_CB + (size_t)(AGG_SCORE_TY + P.kh - 1) * (AGG_SCORE_TX + P.kw - 1) * (AGG_SCORE_CB + 1)) * sizeof(float);
	if (d->hist_smem > 200 * 1024 || d->score_smem > 200 * 1024) return fail(FDB_ERR_UNSUPPORTED, "cell or window size too large for the shared-memory tiles");
	CUDA_TRY(cudaFuncSetAttribute(aggdet_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->hist_smem));
	CUDA_TRY(cudaFuncSetAttribute(aggdet_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->score_smem));
	std::vector<int> tile_layer, stile_layer;
	int64_t cells = 0, positions = 0;
	for (size_t li = 0; li < d->plan.layers.size(); ++li) {
		const PlanLayer& pl = d->plan.layers[li];
		const PyrImage& im = d->plan.images[(size_t)pl.image];
		AggLayer L{};
		L.img_offset = im.kind == IMG_FRAME ? -1 : im.offset; L.img_pitch = im.pitch; L.rows = im.height; L.cols = im.width;
		L.crow = im.height / a.cell_size; L.ccol = im.width / a.cell_size;
		L.vh = std::max(0, L.crow - a.window_rows + 1); L.vw = std::max(0, L.ccol - a.window_cols + 1);
		if (L.vh == 0 || L.vw == 0) L.vh = L.vw = 0;
		L.energy_off = cells; L.hist_off = cells * 2 * a.unsigned_bins; L.feat_off = cells * P.Dp; L.score_off = positions;
		L.tiles_x = (L.ccol + P.tc - 1) / P.tc; L.tiles_y = (L.crow + P.tc - 1) / P.tc; L.first_tile = (int)tile_layer.size();
		for (int t = 0; t < L.tiles_x * L.tiles_y; ++t) tile_layer.push_back((int)li);
		L.stiles_x = (L.vw + AGG_SCORE_TX - 1) / AGG_SCORE_TX; L.stiles_y = (L.vh + AGG_SCORE_TY - 1) / AGG_SCORE_TY; L.first_stile = (int)stile_layer.size();
		for (int t = 0; t < L.stiles_x * L.stiles_y; ++t) stile_layer.push_back((int)li);
		L.index = pl.index;
		L.scale_x = (double)im.width / (double)width; L.scale_y = (double)im.height / (double)height; /* ImagePyramid.cpp:178-179,187-188 */
		cells += (int64_t)L.crow * L.ccol; positions += (int64_t)L.vh * L.vw;
		d->layers.push_back(L);
	}
	d->n_tiles = (int)tile_layer.size(); d->n_stiles = (int)stile_layer.size();
	d->energy_stride = std::max<int64_t>(cells, 1); d->hist_stride = d->energy_stride * 2 * a.unsigned_bins; d->feat_stride = d->energy_stride * P.Dp;
	d->score_stride = std::max<int64_t>(positions, 1);
	d->cand_cap = (int)std::min<int64_t>(std::max<int64_t>(positions, 1) * d->chunk, (int64_t)1 << 22);
	s = build_pyramid_jobs(d->plan.images, d->plan.max_down, width, height, &d->jobs, d->owned); if (s) return s;
	std::vector<FhogLutEntry> lut;
	fhog_build_lut(a.unsigned_bins, a.interpolate_bins != 0, &lut);
	s = upload(lut.data(), lut.size(), &d->d_lut, d->owned); if (s) return s;
	s = upload(d->weights.data(), d->weights.size(), &d->d_weights, d->owned); if (s) return s;
	s = upload(d->layers.data(), d->layers.size(), &d->d_layers, d->owned); if (s) return s;
	s = upload(tile_layer.data(), tile_layer.size(), &d->d_tile_layer, d->owned); if (s) return s;
	s = upload(stile_layer.data(), stile_layer.size(), &d->d_stile_layer, d->owned); if (s) return s;
	s = dev_alloc(&d->d_frames, (size_t)d->chunk * width * height, d->owned); if (s) return s;
	s = dev_alloc(&d->d_arena, (size_t)d->chunk * (size_t)d->plan.arena_bytes, d->owned); if (s) return s;
	s = dev_alloc(&d->d_hist, (size_t)d->chunk * (size_t)d->hist_stride, d->owned); if (s) return s;
	s = dev_alloc(&d->d_energy, (size_t)d->chunk * (size_t)d->energy_stride, d->owned); if (s) return s;
	s = dev_alloc(&d->d_feat, (size_t)d->chunk * (size_t)d->feat_stride, d->owned); if (s) return s;
	s = dev_alloc(&d->d_scores, (size_t)d->chunk * (size_t)d->score_stride, d->owned); if (s) return s;
	s = dev_alloc(&d->d_cand, (size_t)d->cand_cap, d->owned); if (s) return s;
	s = dev_alloc(&d->d_count, 4, d->owned); if (s) return s;
	s = host_alloc(&d->h_cand, (size_t)d->cand_cap, d->owned_host); if (s) return s;
	s = host_alloc(&d->h_count, 4, d->owned_host); if (s) return s;
	d->prepared = true;
	return FDB_OK;
} FDB_API_CATCH

int fdb_aggdet_layers(fdb_aggdet* d, int32_t* n_layers, int32_t* info_out /* [cap][6]: index, width, height, cells x, cells y, positions */, int32_t cap) try {
	if (!d || !d->prepared || !n_layers) return fail(FDB_ERR_INVALID_ARGUMENT, "aggregated-features detector not prepared");
	*n_layers = (int32_t)d->layers.size();
	for (size_t i = 0; i < d->layers.size() && (int32_t)i < cap && info_out; ++i) {
		const AggLayer& L = d->layers[i];
		int32_t* o = info_out + 6 * i;
		o[0] = L.index; o[1] = L.cols; o[2] = L.rows; o[3] = L.ccol; o[4] = L.crow; o[5] = L.vh * L.vw;
	}
	return FDB_OK;
} FDB_API_CATCH

int64_t fdb_aggdet_positions_per_frame(fdb_aggdet* d) {
	if (!d || !d->prepared) return -1;
	int64_t n = 0;
	for (const AggLayer& L : d->layers) n += (int64_t)L.vh * L.vw;
	return n;
}

int fdb_aggdet_detect_batch(fdb_aggdet* d, const uint8_t* frames_host, int64_t pitch, int32_t n_frames, float* scores_out,
		int32_t* rects_xywh_out, int32_t* frame_out, int64_t cap, int64_t* n_out) try {
	return agg_detect(d, frames_host, false, pitch, n_frames, scores_out, rects_xywh_out, frame_out, cap, n_out);
} FDB_API_CATCH

int fdb_aggdet_detect_batch_device(fdb_aggdet* d, const uint8_t* frames_device, int32_t n_frames, float* scores_out,
		int32_t* rects_xywh_out, int32_t* frame_out, int64_t cap, int64_t* n_out) try {
	return agg_detect(d, frames_device, true, 0, n_frames, scores_out, rects_xywh_out, frame_out, cap, n_out);
} FDB_API_CATCH

int fdb_aggdet_score_maps(fdb_aggdet* d, const uint8_t* frame_host, int64_t pitch, float* scores_out, int64_t cap, float* features_out, int64_t feat_cap) try {
	if (!d || !d->prepared || !frame_host) return fail(FDB_ERR_INVALID_ARGUMENT, "bad arguments");
	int s = check_ctx(d->ctx); if (s) return s;
	const int64_t npos = fdb_aggdet_positions_per_frame(d);
	int64_t ncell = 0;
	for (const AggLayer& L : d->layers) ncell += (int64_t)L.crow * L.ccol;
	if ((scores_out && cap < npos) || (features_out && feat_cap < ncell * d->P.D)) return fail(FDB_ERR_OVERFLOW, "output buffer too small");
	if (d->layers.empty()) return FDB_OK;
	cudaStream_t st = d->ctx->stream;
	CUDA_TRY(cudaMemcpy2DAsync(d->d_frames, (size_t)d->W, frame_host, (size_t)pitch, (size_t)d->W, (size_t)d->H, cudaMemcpyHostToDevice, st));
	s = agg_enqueue(d, d->d_frames, 1, true, nullptr); if (s) return s;
	if (scores_out && npos) CUDA_TRY(cudaMemcpyAsync(scores_out, d->d_scores, sizeof(float) * (size_t)npos, cudaMemcpyDeviceToHost, st));
	std::vector<float> padded;
	if (features_out && ncell) {
		padded.resize((size_t)ncell * d->P.Dp);
		CUDA_TRY(cudaMemcpyAsync(padded.data(), d->d_feat, sizeof(float) * padded.size(), cudaMemcpyDeviceToHost, st));
	}
	CUDA_TRY(cudaStreamSynchronize(st));
	if (features_out) for (int64_t i = 0; i < ncell; ++i) std::memcpy(features_out + i * d->P.D, padded.data() + i * d->P.Dp, sizeof(float) * d->P.D);
	return FDB_OK;
} FDB_API_CATCH

int fdb_aggdet_profile_device(fdb_aggdet* d, const uint8_t* frames_device, int32_t n_frames, double ms_out[6]) try {
	if (!d || !d->prepared || !frames_device || !ms_out || n_frames < 0) return fail(FDB_ERR_INVALID_ARGUMENT, "bad arguments");
	int s = check_ctx(d->ctx); if (s) return s;
	cudaEvent_t ev[5];
	for (int k = 0; k < 5; ++k) CUDA_TRY(cudaEventCreate(&ev[k]));
	for (int k = 0; k < 6; ++k) ms_out[k] = 0;
	for (int base = 0; base < n_frames && s == FDB_OK; base += d->chunk) {
		const int n = std::min(d->chunk, n_frames - base);
		s = agg_enqueue(d, frames_device + (int64_t)base * d->W * d->H, n, false, ev);
		if (s) break;
		if (cudaEventSynchronize(ev[4]) != cudaSuccess) { s = fail(FDB_ERR_CUDA, "profile: event synchronize failed"); break; }
		float p = 0, h = 0, de = 0, sc = 0, t = 0;
		cudaEventElapsedTime(&p, ev[0], ev[1]); cudaEventElapsedTime(&h, ev[1], ev[2]); cudaEventElapsedTime(&de, ev[2], ev[3]);
		cudaEventElapsedTime(&sc, ev[3], ev[4]); cudaEventElapsedTime(&t, ev[0], ev[4]);
		ms_out[0] += p; ms_out[1] += h; ms_out[2] += de; ms_out[3] += sc; ms_out[4] += t; ms_out[5] += 1;
	}
	for (int k = 0; k < 5; ++k) cudaEventDestroy(ev[k]);
	return s;
} FDB_API_CATCH

} // extern "C"
