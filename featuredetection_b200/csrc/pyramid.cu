/*
 * pyramid.cu - image pyramid construction on the GPU (sm_100a).
 *
 * Replaces ImagePyramid::createLayers (libImageProcessing/src/imageprocessing/ImagePyramid.cpp:170-198):
 * for every octave offset i one bilinear cv::resize of the frame (:177) and then a chain of
 * cv::pyrDown (:186).  Both are 8-bit fixed-point operations whose intermediate roundings must
 * be reproduced exactly, so the kernels use the same integer formulas:
 *   resize INTER_LINEAR 8UC1: coefficients rint(f * 2048) as int16, horizontal pass in int32,
 *       vertical pass ((b0*(H0>>4))>>16 + (b1*(H1>>4))>>16 + 2) >> 2; exact 2x2 decimation takes
 *       OpenCV's area path (s00+s01+s10+s11+2)>>2; same size is a copy (the plan aliases the frame).
 *   pyrDown 8U: 5x5 [1 4 6 4 1]^2, BORDER_REFLECT_101, (sum + 128) >> 8.
 *
 * Data layout: frames are [n][H][W] u8; every other pyramid image lives at a fixed offset of a
 * per-frame arena (u8, row pitch == width, 16-byte aligned starts).  One launch covers all
 * frames of the batch and all images of one dependency level; each thread produces 4 horizontally
 * adjacent output pixels where the width allows and stores them as one 32-bit word.
 */
#include <cuda_runtime.h>
#include <cstdint>

#include "fdb_internal.h"

namespace fdb {

__device__ __forceinline__ int reflect101(int p, int len) {
	if (len == 1) return 0;
	while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
	return p;
}

/* ---------------------------------------------------------------------------------------------
 * resize: grid = (pixel-quad blocks, job, frame)
 * ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) resize_kernel(const uint8_t* __restrict__ frames, int W, int H,
		uint8_t* __restrict__ arena, int64_t arena_stride,
		const ResizeJob* __restrict__ jobs, const int* __restrict__ ofs_tab, const short2* __restrict__ coef_tab) {
	const ResizeJob job = jobs[blockIdx.y];
	const uint8_t* __restrict__ src = frames + (int64_t)blockIdx.z * W * H;
	uint8_t* __restrict__ dst = arena + (int64_t)blockIdx.z * arena_stride + job.dst_offset;
	const int quads_per_row = (job.dst_w + 3) >> 2;
	const int total = quads_per_row * job.dst_h;
	for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
		const int dy = q / quads_per_row;
		const int dx0 = (q - dy * quads_per_row) << 2;
		uint32_t packed = 0;
		int nvalid = min(4, job.dst_w - dx0);
		if (job.area2x) {
			const uint8_t* s0 = src + (int64_t)(2 * dy) * W;
			const uint8_t* s1 = s0 + W;
			for (int k = 0; k < nvalid; ++k) {
				const int x = 2 * (dx0 + k);
				const int v = (s0[x] + s0[x + 1] + s1[x] + s1[x + 1] + 2) >> 2;
				packed |= (uint32_t)v << (8 * k);
			}
		} else {
			const int sy0 = ofs_tab[job.ytab + dy];
			const short2 b = coef_tab[job.ytab + dy];
			const int y0 = min(max(sy0, 0), H - 1), y1 = min(max(sy0 + 1, 0), H - 1);
			const uint8_t* s0 = src + (int64_t)y0 * W;
			const uint8_t* s1 = src + (int64_t)y1 * W;
			for (int k = 0; k < nvalid; ++k) {
				const int sx = ofs_tab[job.xtab + dx0 + k];
				const short2 a = coef_tab[job.xtab + dx0 + k];
				const int sx1 = min(sx + 1, W - 1);
				const int h0 = s0[sx] * a.x + s0[sx1] * a.y;
				const int h1 = s1[sx] * a.x + s1[sx1] * a.y;
				const int v = (((b.x * (h0 >> 4)) >> 16) + ((b.y * (h1 >> 4)) >> 16) + 2) >> 2;
				packed |= (uint32_t)(v & 255) << (8 * k);
			}
		}
		uint8_t* o = dst + (int64_t)dy * job.dst_w + dx0;
		if (nvalid == 4 && ((job.dst_w & 3) == 0)) {
			*reinterpret_cast<uint32_t*>(o) = packed;
		} else {
			for (int k = 0; k < nvalid; ++k) o[k] = (uint8_t)(packed >> (8 * k));
		}
	}
}

/* ---------------------------------------------------------------------------------------------
 * pyrDown: grid = (pixel blocks, job, frame); one thread per output pixel, direct 25 taps with
 * the horizontal 5-tap sums shared between the 5 rows through registers.
 * ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) pyrdown_kernel(const uint8_t* __restrict__ frames, int W, int H,
		uint8_t* __restrict__ arena, int64_t arena_stride, const DownJob* __restrict__ jobs) {
	const DownJob job = jobs[blockIdx.y];
	const uint8_t* __restrict__ src = job.src_offset < 0
			? frames + (int64_t)blockIdx.z * W * H
			: arena + (int64_t)blockIdx.z * arena_stride + job.src_offset;
	uint8_t* __restrict__ dst = arena + (int64_t)blockIdx.z * arena_stride + job.dst_offset;
	const int total = job.dst_w * job.dst_h;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
		const int y = p / job.dst_w, x = p - y * job.dst_w;
		int xs[5];
		const bool interior_x = 2 * x - 2 >= 0 && 2 * x + 2 < job.src_w;
#pragma unroll
		for (int u = 0; u < 5; ++u) xs[u] = interior_x ? 2 * x - 2 + u : reflect101(2 * x - 2 + u, job.src_w);
		int acc = 0;
#pragma unroll
		for (int t = 0; t < 5; ++t) {
			const int sy = reflect101(2 * y - 2 + t, job.src_h);
			const uint8_t* r = src + (int64_t)sy * job.src_w;
			const int h = r[xs[0]] + 4 * r[xs[1]] + 6 * r[xs[2]] + 4 * r[xs[3]] + r[xs[4]];
			const int wt = (t == 0 || t == 4) ? 1 : ((t == 2) ? 6 : 4);
			acc += wt * h;
		}
		dst[p] = (uint8_t)((acc + 128) >> 8);
	}
}

/* ---------------------------------------------------------------------------------------------
 * launchers
 * ------------------------------------------------------------------------------------------- */
void launch_resize(cudaStream_t st, const uint8_t* frames, int W, int H, int n_frames, uint8_t* arena,
		int64_t arena_stride, const ResizeJob* jobs_dev, int n_jobs, int max_quads,
		const int* ofs_tab, const short2* coef_tab) {
	if (n_jobs == 0 || n_frames == 0) return;
	dim3 grid((unsigned)((max_quads + 255) / 256), (unsigned)n_jobs, (unsigned)n_frames);
	resize_kernel<<<grid, 256, 0, st>>>(frames, W, H, arena, arena_stride, jobs_dev, ofs_tab, coef_tab);
}

void launch_pyrdown(cudaStream_t st, const uint8_t* frames, int W, int H, int n_frames, uint8_t* arena,
		int64_t arena_stride, const DownJob* jobs_dev, int n_jobs, int max_pixels) {
	if (n_jobs == 0 || n_frames == 0) return;
	dim3 grid((unsigned)((max_pixels + 255) / 256), (unsigned)n_jobs, (unsigned)n_frames);
	pyrdown_kernel<<<grid, 256, 0, st>>>(frames, W, H, arena, arena_stride, jobs_dev);
}

} // namespace fdb
