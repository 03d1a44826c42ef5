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
 * resize: grid = (tiles, job, frame); a CTA of 256 threads = 32 pixel-quads x 8 rows, i.e. a
 * 128 x 8 output tile; each thread produces 4 horizontally adjacent pixels and stores one word.
 * Tables: xy_tab[k] = {source offset, a0 | a1 << 16} (one 8-byte load per output column / row).
 * ------------------------------------------------------------------------------------------- */
#define RS_QX 32
#define RS_TY 8

__global__ void __launch_bounds__(RS_QX * RS_TY) resize_kernel(const uint8_t* __restrict__ frames, int W, int H,
		uint8_t* __restrict__ arena, int64_t arena_stride,
		const ResizeJob* __restrict__ jobs, const int2* __restrict__ xy_tab) {
	const ResizeJob job = jobs[blockIdx.y];
	const int quads_per_row = (job.dst_w + 3) >> 2;
	const int tiles_x = (quads_per_row + RS_QX - 1) / RS_QX, tiles_y = (job.dst_h + RS_TY - 1) / RS_TY;
	if ((int)blockIdx.x >= tiles_x * tiles_y) return;
	const int tile_y = (int)blockIdx.x / tiles_x, tile_x = (int)blockIdx.x - tile_y * tiles_x;
	const int dy = tile_y * RS_TY + ((int)threadIdx.x >> 5);
	const int dx0 = (tile_x * RS_QX + ((int)threadIdx.x & 31)) << 2;
	if (dy >= job.dst_h || dx0 >= job.dst_w) return;
	const uint8_t* __restrict__ src = frames + (int64_t)blockIdx.z * W * H;
	uint8_t* __restrict__ o = arena + (int64_t)blockIdx.z * arena_stride + job.dst_offset + dy * job.dst_w + dx0;
	const int nvalid = min(4, job.dst_w - dx0);
	uint32_t packed = 0;
	if (job.area2x) {
		const uint8_t* s0 = src + (2 * dy) * W;
		const uint8_t* s1 = s0 + W;
#pragma unroll
		for (int k = 0; k < 4; ++k)
			if (k < nvalid) {
				const int x = 2 * (dx0 + k);
				packed |= (uint32_t)((s0[x] + s0[x + 1] + s1[x] + s1[x + 1] + 2) >> 2) << (8 * k);
			}
	} else {
		const int2 ty = __ldg(xy_tab + job.ytab + dy);
		const int b0 = (short)(ty.y & 0xffff), b1 = ty.y >> 16;
		const int y0 = min(max(ty.x, 0), H - 1), y1 = min(max(ty.x + 1, 0), H - 1);
		const uint8_t* s0 = src + y0 * W;
		const uint8_t* s1 = src + y1 * W;
#pragma unroll
		for (int k = 0; k < 4; ++k)
			if (k < nvalid) {
				const int2 tx = __ldg(xy_tab + job.xtab + dx0 + k);
				const int a0 = (short)(tx.y & 0xffff), a1 = tx.y >> 16;
				const int sx = tx.x, sx1 = min(sx + 1, W - 1);
				const int h0 = s0[sx] * a0 + s0[sx1] * a1;
				const int h1 = s1[sx] * a0 + s1[sx1] * a1;
				const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
				packed |= (uint32_t)(v & 255) << (8 * k);
			}
	}
	if (nvalid == 4 && ((job.dst_w & 3) == 0)) {
		*reinterpret_cast<uint32_t*>(o) = packed;
	} else {
		for (int k = 0; k < nvalid; ++k) o[k] = (uint8_t)(packed >> (8 * k));
	}
}

/* ---------------------------------------------------------------------------------------------
 * pyrDown: grid = (tiles, job, frame); a CTA of 256 threads produces a 32x8 output tile.
 * The 67x19 input footprint is staged once in shared memory (border pixels resolved with
 * BORDER_REFLECT_101, only on tiles that touch the image border), the separable filter runs as a
 * horizontal pass into shared memory and a vertical pass from it: 5 + 5 shared-memory taps per
 * output instead of 25 global loads.  Integer sums are identical to the 25-tap form (OpenCV's
 * pyrDown has no intermediate rounding).
 * ------------------------------------------------------------------------------------------- */
#define PD_TW 32
#define PD_TH 8
#define PD_IW (2 * PD_TW + 3)
#define PD_IH (2 * PD_TH + 3)

__global__ void __launch_bounds__(PD_TW * PD_TH) pyrdown_kernel(const uint8_t* __restrict__ frames, int W, int H,
		uint8_t* __restrict__ arena, int64_t arena_stride, const DownJob* __restrict__ jobs) {
	__shared__ uint8_t s_in[PD_IH][PD_IW + 1];
	__shared__ uint16_t s_h[PD_IH][PD_TW];
	const DownJob job = jobs[blockIdx.y];
	const int tiles_x = (job.dst_w + PD_TW - 1) / PD_TW, tiles_y = (job.dst_h + PD_TH - 1) / PD_TH;
	if ((int)blockIdx.x >= tiles_x * tiles_y) return;
	const int tile_y = (int)blockIdx.x / tiles_x;
	const int ty0 = tile_y * PD_TH, tx0 = ((int)blockIdx.x - tile_y * tiles_x) * PD_TW;
	const uint8_t* __restrict__ src = job.src_offset < 0
			? frames + (int64_t)blockIdx.z * W * H
			: arena + (int64_t)blockIdx.z * arena_stride + job.src_offset;
	uint8_t* __restrict__ dst = arena + (int64_t)blockIdx.z * arena_stride + job.dst_offset;
	const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
	const int ix0 = 2 * tx0 - 2, iy0 = 2 * ty0 - 2;
	const bool interior = ix0 >= 0 && iy0 >= 0 && ix0 + PD_IW <= job.src_w && iy0 + PD_IH <= job.src_h;
	if (interior) {
		for (int r = ly; r < PD_IH; r += PD_TH) {
			const uint8_t* row = src + (iy0 + r) * job.src_w + ix0;
			s_in[r][lx] = row[lx];
			s_in[r][lx + 32] = row[lx + 32];
			if (lx < PD_IW - 64) s_in[r][lx + 64] = row[lx + 64];
		}
	} else {
		for (int r = ly; r < PD_IH; r += PD_TH) {
			const uint8_t* row = src + reflect101(iy0 + r, job.src_h) * job.src_w;
			for (int c = lx; c < PD_IW; c += 32) s_in[r][c] = row[reflect101(ix0 + c, job.src_w)];
		}
	}
	__syncthreads();
	for (int r = ly; r < PD_IH; r += PD_TH) {
		const uint8_t* p = &s_in[r][2 * lx];
		s_h[r][lx] = (uint16_t)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
	}
	__syncthreads();
	if (tx0 + lx < job.dst_w && ty0 + ly < job.dst_h) {
		const int acc = s_h[2 * ly][lx] + 4 * s_h[2 * ly + 1][lx] + 6 * s_h[2 * ly + 2][lx] + 4 * s_h[2 * ly + 3][lx] + s_h[2 * ly + 4][lx];
		dst[(ty0 + ly) * job.dst_w + tx0 + lx] = (uint8_t)((acc + 128) >> 8);
	}
}

/* ---------------------------------------------------------------------------------------------
 * launchers
 * ------------------------------------------------------------------------------------------- */
void launch_resize(cudaStream_t st, const uint8_t* frames, int W, int H, int n_frames, uint8_t* arena,
		int64_t arena_stride, const ResizeJob* jobs_dev, int n_jobs, int max_tiles, const int2* xy_tab) {
	if (n_jobs == 0 || n_frames == 0) return;
	dim3 grid((unsigned)max_tiles, (unsigned)n_jobs, (unsigned)n_frames);
	resize_kernel<<<grid, RS_QX * RS_TY, 0, st>>>(frames, W, H, arena, arena_stride, jobs_dev, xy_tab);
}

int resize_tiles(int dst_w, int dst_h) {
	const int quads = (dst_w + 3) / 4;
	return ((quads + RS_QX - 1) / RS_QX) * ((dst_h + RS_TY - 1) / RS_TY);
}

void launch_pyrdown(cudaStream_t st, const uint8_t* frames, int W, int H, int n_frames, uint8_t* arena,
		int64_t arena_stride, const DownJob* jobs_dev, int n_jobs, int max_tiles) {
	if (n_jobs == 0 || n_frames == 0) return;
	dim3 grid((unsigned)max_tiles, (unsigned)n_jobs, (unsigned)n_frames);
	pyrdown_kernel<<<grid, PD_TW * PD_TH, 0, st>>>(frames, W, H, arena, arena_stride, jobs_dev);
}

int pyrdown_tiles(int dst_w, int dst_h) {
	return ((dst_w + PD_TW - 1) / PD_TW) * ((dst_h + PD_TH - 1) / PD_TH);
}

} // namespace fdb
