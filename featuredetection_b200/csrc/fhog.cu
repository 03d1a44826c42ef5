/*
 * fhog.cu - FHOG layer filter on the GPU (the feature map that
 * detection::AggregatedFeaturesDetector convolves its linear SVM with). Kernels around the host/device functions of
 * fhog_core.h; the gradient look-up table is built on the host with the reference's formulas (FhogFilter.cpp:36-57).
 *
 * STATUS: NOT YET RUN ON A B200 (written after the round's GPU budget was spent). The arithmetic is verified on the host
 * (tests/test_fhog_host_emulation.py runs fhog_core.h under g++ against the pinned oracle); the launch code below is not.
 * Nothing else in the library calls into this file.
 *
 *   fhog_hist_kernel   thread = (cell, signed bin): replays the cell's contributing pixels in raster order (float32 sums are
 *                      order dependent); 2 * bins threads per cell, the LUT (5.2 MB) is L2 resident
 *   fhog_desc_kernel   thread = cell: energies of the 3 x 3 neighbourhood -> 4 normalisers -> 3 * bins + 4 features
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "api_types.h"
#include "fhog_core.h"

namespace fdb {

__global__ void __launch_bounds__(256) fhog_hist_kernel(const FhogLutEntry* __restrict__ lut, const uint8_t* __restrict__ image,
		int pitch, int rows, int cols, int channels, int cell, int crow, int ccol, int unsigned_bins, int interpolate_bins,
		int interpolate_cells, float* __restrict__ hist /* [crow * ccol][2 * unsigned_bins] */) {
	const int signed_bins = 2 * unsigned_bins;
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (int64_t)crow * ccol * signed_bins) return;
	const int bin = (int)(idx % signed_bins);
	const int cellidx = (int)(idx / signed_bins);
	const int cr = cellidx / ccol, cc = cellidx - cr * ccol;
	hist[idx] = fhog_signed_bin(lut, image, pitch, rows, cols, channels, cell, crow, ccol, interpolate_bins, interpolate_cells, cr, cc, bin);
}

__global__ void __launch_bounds__(256) fhog_energy_kernel(const float* __restrict__ hist, int n_cells, int unsigned_bins,
		float* __restrict__ energies) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_cells) energies[i] = fhog_energy(hist + (int64_t)i * 2 * unsigned_bins, unsigned_bins);
}

__global__ void __launch_bounds__(128) fhog_desc_kernel(const float* __restrict__ hist, const float* __restrict__ energies, int crow,
		int ccol, int unsigned_bins, float alpha, float* __restrict__ out /* [crow * ccol][3 * unsigned_bins + 4] */) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= crow * ccol) return;
	const int r = i / ccol, c = i - r * ccol;
	fhog_descriptor(hist + (int64_t)i * 2 * unsigned_bins, energies, crow, ccol, r, c, unsigned_bins, alpha,
			out + (int64_t)i * (3 * unsigned_bins + 4));
}

/* FhogFilter::createGradientLut (FhogFilter.cpp:36-57) with GradientOrientationFilter::computeOrientation (full
 * orientations, GradientOrientationFilter.cpp:137-144) and GradientMagnitudeFilter::computeMagnitude (:80-86) */
void fhog_build_lut(int unsigned_bins, int interpolate_bins, std::vector<FhogLutEntry>* out) {
	const int signed_bins = 2 * unsigned_bins;
	const float two_pi = static_cast<float>(2 * M_PI);
	const float value2bin = signed_bins / two_pi;
	out->assign(512 * 512, FhogLutEntry{0, 0, 0.f, 0.f, 0.f});
	for (int cx = 1; cx < 512; ++cx) {
		const float gx = (cx - 256) / (255.0f * 2.0f);
		for (int cy = 1; cy < 512; ++cy) {
			const float gy = (cy - 256) / (255.0f * 2.0f);
			FhogLutEntry e{0, 0, 0.f, 0.f, 0.f};
			e.magnitude = std::sqrt(gx * gx + gy * gy);
			float orientation = std::atan2(gy, gx);
			if (orientation < 0) orientation += two_pi;
			if (interpolate_bins) {
				const float bin = orientation * value2bin;
				e.index1 = static_cast<int>(bin);
				e.index2 = e.index1 + 1;
				if (e.index2 == signed_bins) e.index2 = 0;
				e.weight2 = e.magnitude * (bin - e.index1);
				e.weight1 = e.magnitude - e.weight2;
			} else {
				int bin = static_cast<int>(orientation * value2bin + 0.5f);
				if (bin == signed_bins) bin = 0;
				e.index1 = bin;
				e.weight1 = e.magnitude;
			}
			(*out)[(size_t)cy * 512 + cx] = e;
		}
	}
}

/* thread = score-map position; weights are read through the read-only path (the same address across a warp's lanes) */
__global__ void __launch_bounds__(128) aggdet_score_kernel(const float* __restrict__ feat, int rows, int cols, int D,
		const float* __restrict__ weights, int kh, int kw, float bias, float* __restrict__ scores /* [rows - kh + 1][cols - kw + 1] */) {
	const int vw = cols - kw + 1, vh = rows - kh + 1;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= vw * vh) return;
	const int y = i / vw, x = i - y * vw;
	scores[i] = aggdet_score(feat, cols, D, weights, kh, kw, bias, y, x);
}

} // namespace fdb

using namespace fdb;

extern "C" int fdb_fhog(fdb_ctx* ctx, const uint8_t* image_host, int64_t pitch, int32_t width, int32_t height, int32_t channels,
		int32_t cell_size, int32_t unsigned_bins, int32_t interpolate_bins, int32_t interpolate_cells, float alpha, float* out_host) try {
	int s = check_ctx(ctx); if (s) return s;
	if (!image_host || !out_host) return fail(FDB_ERR_INVALID_ARGUMENT, "null buffer");
	if (channels != 1 && channels != 3) return fail(FDB_ERR_INVALID_ARGUMENT, "FhogFilter: the image type must be CV_8UC1 or CV_8UC3");
	if (unsigned_bins < 1 || unsigned_bins > 64) return fail(FDB_ERR_INVALID_ARGUMENT, "FhogFilter: unsignedBinCount must be bigger than zero");
	if (!(alpha > 0)) return fail(FDB_ERR_INVALID_ARGUMENT, "FhogAggregationFilter: alpha must be bigger than zero");
	if (cell_size < 1 || width < 1 || height < 1 || pitch < (int64_t)width * channels) return fail(FDB_ERR_INVALID_ARGUMENT, "bad image geometry");
	const int crow = height / cell_size, ccol = width / cell_size;
	if (crow == 0 || ccol == 0) return FDB_OK; /* an image smaller than a cell has no descriptor */
	const int signed_bins = 2 * unsigned_bins, D = signed_bins + unsigned_bins + 4;
	std::vector<FhogLutEntry> lut;
	fhog_build_lut(unsigned_bins, interpolate_bins != 0, &lut);
	std::vector<void*> tmp;
	FhogLutEntry* d_lut; uint8_t* d_img; float* d_hist; float* d_energy; float* d_out;
	const size_t row_bytes = (size_t)width * channels;
	s = upload(lut.data(), lut.size(), &d_lut, tmp);
	if (!s) s = dev_alloc(&d_img, row_bytes * height, tmp);
	if (!s) s = dev_alloc(&d_hist, (size_t)crow * ccol * signed_bins, tmp);
	if (!s) s = dev_alloc(&d_energy, (size_t)crow * ccol, tmp);
	if (!s) s = dev_alloc(&d_out, (size_t)crow * ccol * D, tmp);
	if (s) { free_all(tmp); return s; }
	cudaStream_t st = ctx->stream;
	cudaError_t e = cudaMemcpy2DAsync(d_img, row_bytes, image_host, (size_t)pitch, row_bytes, (size_t)height, cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess) {
		const int64_t n_bins = (int64_t)crow * ccol * signed_bins;
		fhog_hist_kernel<<<(unsigned)((n_bins + 255) / 256), 256, 0, st>>>(d_lut, d_img, (int)row_bytes, height, width, channels, cell_size,
				crow, ccol, unsigned_bins, interpolate_bins != 0, interpolate_cells != 0, d_hist);
		fhog_energy_kernel<<<(unsigned)((crow * ccol + 255) / 256), 256, 0, st>>>(d_hist, crow * ccol, unsigned_bins, d_energy);
		fhog_desc_kernel<<<(unsigned)((crow * ccol + 127) / 128), 128, 0, st>>>(d_hist, d_energy, crow, ccol, unsigned_bins, alpha, d_out);
		ctx->launches += 3;
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d_out, sizeof(float) * (size_t)crow * ccol * D, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	free_all(tmp);
	if (e != cudaSuccess) return fail(FDB_ERR_CUDA, std::string("fdb_fhog: ") + cudaGetErrorString(e));
	return FDB_OK;
} FDB_API_CATCH

/* FHOG + linear-SVM score map of ONE pyramid layer (see fdb200.h) */
extern "C" int fdb_fhog_score_map(fdb_ctx* ctx, const uint8_t* image_host, int64_t pitch, int32_t width, int32_t height, int32_t channels,
		int32_t cell_size, int32_t unsigned_bins, int32_t interpolate_bins, int32_t interpolate_cells, float alpha,
		const float* weights_host, int32_t kernel_rows, int32_t kernel_cols, float bias, float* scores_host) try {
	int s = check_ctx(ctx); if (s) return s;
	if (!weights_host || !scores_host || kernel_rows < 1 || kernel_cols < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "bad kernel");
	if (!image_host) return fail(FDB_ERR_INVALID_ARGUMENT, "null buffer");
	if (channels != 1 && channels != 3) return fail(FDB_ERR_INVALID_ARGUMENT, "FhogFilter: the image type must be CV_8UC1 or CV_8UC3");
	if (unsigned_bins < 1 || unsigned_bins > 64) return fail(FDB_ERR_INVALID_ARGUMENT, "FhogFilter: unsignedBinCount must be bigger than zero");
	if (!(alpha > 0)) return fail(FDB_ERR_INVALID_ARGUMENT, "FhogAggregationFilter: alpha must be bigger than zero");
	if (cell_size < 1 || width < 1 || height < 1 || pitch < (int64_t)width * channels) return fail(FDB_ERR_INVALID_ARGUMENT, "bad image geometry");
	const int crow = height / cell_size, ccol = width / cell_size, D = 3 * unsigned_bins + 4;
	const int vh = crow - kernel_rows + 1, vw = ccol - kernel_cols + 1;
	if (vh <= 0 || vw <= 0) return FDB_OK; /* no window fits this layer */
	std::vector<float> feat((size_t)crow * ccol * D);
	s = fdb_fhog(ctx, image_host, pitch, width, height, channels, cell_size, unsigned_bins, interpolate_bins, interpolate_cells, alpha, feat.data());
	if (s) return s;
	std::vector<void*> tmp;
	float* d_feat; float* d_w; float* d_scores;
	s = upload(feat.data(), feat.size(), &d_feat, tmp);
	if (!s) s = upload(weights_host, (size_t)kernel_rows * kernel_cols * D, &d_w, tmp);
	if (!s) s = dev_alloc(&d_scores, (size_t)vh * vw, tmp);
	if (s) { free_all(tmp); return s; }
	cudaStream_t st = ctx->stream;
	aggdet_score_kernel<<<(unsigned)((vh * vw + 127) / 128), 128, 0, st>>>(d_feat, crow, ccol, D, d_w, kernel_rows, kernel_cols, bias, d_scores);
	ctx->launches++;
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess) e = cudaMemcpyAsync(scores_host, d_scores, sizeof(float) * (size_t)vh * vw, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	free_all(tmp);
	if (e != cudaSuccess) return fail(FDB_ERR_CUDA, std::string("fdb_fhog_score_map: ") + cudaGetErrorString(e));
	return FDB_OK;
} FDB_API_CATCH
