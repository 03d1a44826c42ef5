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
 * per-frame arena (u8, row pitch = width rounded up to 16 bytes, 128-byte aligned starts).  One launch covers all
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
 * Tables: xy_tab[k] = {source offset, a0, a1, word-path info} (one 16-byte load per output column / row);
 * word-path info of column dx = word select | bit shift << 8 | base word of its quad << 16.
 * ------------------------------------------------------------------------------------------- */
#define RS_QX 32
#define RS_TY 8

#define RS_ROWS 4   /* output rows per thread: the column tables are loaded once and reused */

__global__ void __launch_bounds__(RS_QX * RS_TY) resize_kernel(const uint8_t* __restrict__ frames, int W, int H,
		uint8_t* __restrict__ arena, int64_t arena_stride,
		const ResizeJob* __restrict__ jobs, const int4* __restrict__ xy_tab) {
	const ResizeJob job = jobs[blockIdx.y];
	const int quads_per_row = (job.dst_w + 3) >> 2;
	const int tiles_x = (quads_per_row + RS_QX - 1) / RS_QX, tiles_y = (job.dst_h + RS_TY * RS_ROWS - 1) / (RS_TY * RS_ROWS);
	if ((int)blockIdx.x >= tiles_x * tiles_y) return;
	const int tile_y = (int)blockIdx.x / tiles_x, tile_x = (int)blockIdx.x - tile_y * tiles_x;
	const int dx0 = (tile_x * RS_QX + ((int)threadIdx.x & 31)) << 2;
	if (dx0 >= job.dst_w) return;
	const uint8_t* __restrict__ src = frames + (int64_t)blockIdx.z * W * H;
	uint8_t* __restrict__ dst = arena + (int64_t)blockIdx.z * arena_stride + job.dst_offset;
	const int nvalid = min(4, job.dst_w - dx0);
	const bool word_store = nvalid == 4; /* rows are 16-byte aligned (dst_pitch), dx0 is a multiple of 4 */
	const int4* __restrict__ xt = xy_tab + job.xtab + dx0;
	const bool fast = nvalid == 4 && job.words_ok && !job.area2x;
	int4 t0 = make_int4(0, 0, 0, 0), t1 = t0, t2 = t0, t3 = t0;
	if (fast) { t0 = __ldg(xt); t1 = __ldg(xt + 1); t2 = __ldg(xt + 2); t3 = __ldg(xt + 3); }
	const int wlast = (W >> 2) - 1;
	const int b0 = (int)((uint32_t)t0.w >> 16), b1 = min(b0 + 1, wlast), b2 = min(b0 + 2, wlast);
	/* word path: the 8 source bytes of a row that 4 adjacent outputs need lie within 12 bytes of a 4-aligned base (scale < 2):
	 * three aligned 32-bit loads per row; the byte pairs (sx, sx + 1) of two outputs are gathered by two PRMT - one over the first
	 * two words, one over its result and the third word - whose selectors depend on the column only (computed once per thread).
	 * t.w = word select | bit shift << 8 | base word << 16, so the byte offset of sx in the 12 bytes is 4 * select + shift / 8 */
	uint32_t sel1[2] = {0u, 0u}, sel2[2] = {0u, 0u};
	if (fast) {
		const int off[4] = {(t0.w & 255) * 4 + ((t0.w >> 8) & 31) / 8, (t1.w & 255) * 4 + ((t1.w >> 8) & 31) / 8,
				(t2.w & 255) * 4 + ((t2.w >> 8) & 31) / 8, (t3.w & 255) * 4 + ((t3.w >> 8) & 31) / 8};
#pragma unroll
		for (int pr = 0; pr < 2; ++pr)
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const int j = off[2 * pr + (i >> 1)] + (i & 1); /* source byte of output byte i of the pair: 0 .. 11 */
				sel1[pr] |= (uint32_t)(j < 8 ? j : 0) << (4 * i);
				sel2[pr] |= (uint32_t)(j < 8 ? i : j - 4) << (4 * i);
			}
	}
	const uint32_t cxy0 = (uint32_t)t0.y | ((uint32_t)t0.z << 16), cxy1 = (uint32_t)t1.y | ((uint32_t)t1.z << 16); /* a0, a1 <= 2048 as two u16 */
	const uint32_t cxy2 = (uint32_t)t2.y | ((uint32_t)t2.z << 16), cxy3 = (uint32_t)t3.y | ((uint32_t)t3.z << 16);
#pragma unroll
	for (int i = 0; i < RS_ROWS; ++i) {
		const int dy = tile_y * (RS_TY * RS_ROWS) + i * RS_TY + ((int)threadIdx.x >> 5);
		if (dy >= job.dst_h) break;
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
			const int4 ty = __ldg(xy_tab + job.ytab + dy);
			const int y0 = min(max(ty.x, 0), H - 1), y1 = min(max(ty.x + 1, 0), H - 1);
			const uint8_t* __restrict__ s0 = src + y0 * W;
			const uint8_t* __restrict__ s1 = src + y1 * W;
			if (fast) {
				const uint32_t* __restrict__ r0 = reinterpret_cast<const uint32_t*>(s0);
				const uint32_t* __restrict__ r1 = reinterpret_cast<const uint32_t*>(s1);
				const int by0 = ty.y << 16, by1 = ty.z << 16; /* coefficients <= 2048: no overflow; all factors are non-negative */
				const uint32_t p0 = __ldg(r0 + b0), p1 = __ldg(r0 + b1), p2 = __ldg(r0 + b2);
				const uint32_t q0 = __ldg(r1 + b0), q1 = __ldg(r1 + b1), q2 = __ldg(r1 + b2);
				/* {S[sx_a], S[sx_a + 1], S[sx_b], S[sx_b + 1]} of the upper and the lower source row, outputs (0, 1) and (2, 3) */
				const uint32_t up01 = __byte_perm(__byte_perm(p0, p1, sel1[0]), p2, sel2[0]), lo01 = __byte_perm(__byte_perm(q0, q1, sel1[0]), q2, sel2[0]);
				const uint32_t up23 = __byte_perm(__byte_perm(p0, p1, sel1[1]), p2, sel2[1]), lo23 = __byte_perm(__byte_perm(q0, q1, sel1[1]), q2, sel2[1]);
				/* S[sx] * a0 + S[sx + 1] * a1 (dp2a on the low / high byte pair), then (b * (h >> 4)) >> 16 as the high word of (b << 16) * (h >> 4) */
#define FDB_RS_OUT(H0, H1) ((__mulhi(by0, (int)(H0) >> 4) + __mulhi(by1, (int)(H1) >> 4) + 2) >> 2)
				const int v0 = FDB_RS_OUT(__dp2a_lo(cxy0, up01, 0u), __dp2a_lo(cxy0, lo01, 0u));
				const int v1 = FDB_RS_OUT(__dp2a_hi(cxy1, up01, 0u), __dp2a_hi(cxy1, lo01, 0u));
				const int v2 = FDB_RS_OUT(__dp2a_lo(cxy2, up23, 0u), __dp2a_lo(cxy2, lo23, 0u));
				const int v3 = FDB_RS_OUT(__dp2a_hi(cxy3, up23, 0u), __dp2a_hi(cxy3, lo23, 0u));
#undef FDB_RS_OUT
				packed = (uint32_t)(v0 & 255) | ((uint32_t)(v1 & 255) << 8) | ((uint32_t)(v2 & 255) << 16) | ((uint32_t)(v3 & 255) << 24);
		} else {
				for (int k = 0; k < nvalid; ++k) {
					const int4 tx = __ldg(xt + k);
					const int sx = tx.x, sx1 = min(sx + 1, W - 1);
					const int h0 = s0[sx] * tx.y + s0[sx1] * tx.z;
					const int h1 = s1[sx] * tx.y + s1[sx1] * tx.z;
					const int v = (((ty.y * (h0 >> 4)) >> 16) + ((ty.z * (h1 >> 4)) >> 16) + 2) >> 2;
					packed |= (uint32_t)(v & 255) << (8 * k);
				}
			}
		}
		uint8_t* o = dst + dy * job.dst_pitch + dx0;
		if (word_store) {
			*reinterpret_cast<uint32_t*>(o) = packed;
		} else {
			for (int k = 0; k < nvalid; ++k) o[k] = (uint8_t)(packed >> (8 * k));
		}
	}
}

/* ---------------------------------------------------------------------------------------------
 * pyrDown: grid = (tiles, job, frame); a CTA of 256 threads = 32 x 8 threads, each thread
 * produces a 4 (x) x 4 (y) block of output pixels => a 128 x 32 output tile per CTA.
 * Interior threads read their 11 x 11 input footprint as aligned 32-bit words straight from global
 * memory (L1 serves the overlap with the neighbours), realign with funnel shifts and evaluate the
 * horizontal [1 4 6 4 1] taps with dp4a on packed bytes; the vertical taps run on registers.
 * Threads whose footprint touches the image border (BORDER_REFLECT_101) take a byte-wise path.
 * Integer sums are identical to the 25-tap form (OpenCV's pyrDown has no intermediate rounding).
 * ------------------------------------------------------------------------------------------- */
#define PD_BX 32
#define PD_BY 8
#define PD_TW (PD_BX * 4)
#ifndef PD_RY
#define PD_RY 4              /* output rows per thread: 2 * PD_RY + 3 input rows feed PD_RY output rows */
#endif
#define PD_TH (PD_BY * PD_RY)

__device__ __forceinline__ void pd_hrow_fast(const uint8_t* __restrict__ rowp, int col, int* h) {
	/* bytes col .. col+10 of the row as three words b[0..3], b[4..7], b[8..11] */
	const uintptr_t addr = reinterpret_cast<uintptr_t>(rowp + col);
	const uint32_t* __restrict__ wp = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
	const unsigned sh = (unsigned)(addr & 3) * 8;
	const uint32_t a0 = __ldg(wp), a1 = __ldg(wp + 1), a2 = __ldg(wp + 2), a3 = __ldg(wp + 3);
	const uint32_t w0 = __funnelshift_r(a0, a1, sh), w1 = __funnelshift_r(a1, a2, sh), w2 = __funnelshift_r(a2, a3, sh);
	h[0] = (int)__dp4a(w0, 0x04060401u, w1 & 0xffu);
	h[1] = (int)__dp4a(w1, 0x00010406u, __dp4a(w0, 0x04010000u, 0u));
	h[2] = (int)__dp4a(w1, 0x04060401u, w2 & 0xffu);
	h[3] = (int)__dp4a(w2, 0x00010406u, __dp4a(w1, 0x04010000u, 0u));
}

/* border threads: the 11 reflected column indices are resolved once per thread (cx), the row once per call */
__device__ __forceinline__ void pd_hrow_border(const uint8_t* __restrict__ rowp, const int* cx, int* h) {
	int b[11];
#pragma unroll
	for (int i = 0; i < 11; ++i) b[i] = rowp[cx[i]];
#pragma unroll
	for (int k = 0; k < 4; ++k) h[k] = b[2 * k] + 4 * b[2 * k + 1] + 6 * b[2 * k + 2] + 4 * b[2 * k + 3] + b[2 * k + 4];
}

__global__ void __launch_bounds__(PD_BX * PD_BY) pyrdown_kernel(const uint8_t* __restrict__ frames, int W, int H,
		uint8_t* __restrict__ arena, int64_t arena_stride, const DownJob* __restrict__ jobs) {
	const DownJob job = jobs[blockIdx.y];
	const int tiles_x = (job.dst_w + PD_TW - 1) / PD_TW, tiles_y = (job.dst_h + PD_TH - 1) / PD_TH;
	if ((int)blockIdx.x >= tiles_x * tiles_y) return;
	const int tile_y = (int)blockIdx.x / tiles_x;
	const int x0 = (((int)blockIdx.x - tile_y * tiles_x) * PD_BX + ((int)threadIdx.x & 31)) * 4;
	const int y0 = (tile_y * PD_BY + ((int)threadIdx.x >> 5)) * PD_RY;
	if (x0 >= job.dst_w || y0 >= job.dst_h) return;
	const uint8_t* __restrict__ src = job.src_offset < 0
			? frames + (int64_t)blockIdx.z * W * H
			: arena + (int64_t)blockIdx.z * arena_stride + job.src_offset;
	uint8_t* __restrict__ dst = arena + (int64_t)blockIdx.z * arena_stride + job.dst_offset;
	const int col = 2 * x0 - 2, row = 2 * y0 - 2;
	/* fast path: footprint strictly inside the image and not on its last row (aligned word reads may
	 * run a few bytes past the footprint) */
	constexpr int NR = 2 * PD_RY + 3; /* input rows of a thread */
	const bool interior = col >= 0 && row >= 0 && col + 16 <= job.src_w && row + NR < job.src_h;
	int h[NR][4];
	if (interior) {
#pragma unroll
		for (int r = 0; r < NR; ++r) pd_hrow_fast(src + (row + r) * job.src_pitch, col, h[r]);
	} else {
		int cx[11];
#pragma unroll
		for (int i = 0; i < 11; ++i) cx[i] = reflect101(col + i, job.src_w);
#pragma unroll
		for (int r = 0; r < NR; ++r) pd_hrow_border(src + reflect101(row + r, job.src_h) * job.src_pitch, cx, h[r]);
	}
	const int nx = min(4, job.dst_w - x0);
#pragma unroll
	for (int oy = 0; oy < PD_RY; ++oy) {
		if (y0 + oy >= job.dst_h) break;
		uint32_t packed = 0;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const int acc = h[2 * oy][k] + 4 * h[2 * oy + 1][k] + 6 * h[2 * oy + 2][k] + 4 * h[2 * oy + 3][k] + h[2 * oy + 4][k];
			packed |= (uint32_t)((acc + 128) >> 8) << (8 * k);
		}
		uint8_t* o = dst + (y0 + oy) * job.dst_pitch + x0;
		if (nx == 4) { /* 16-byte aligned rows, x0 is a multiple of 4 */
			*reinterpret_cast<uint32_t*>(o) = packed;
		} else {
			for (int k = 0; k < nx; ++k) o[k] = (uint8_t)(packed >> (8 * k));
		}
	}
}

/* ---------------------------------------------------------------------------------------------
 * GrayscaleFilter::applyTo (GrayscaleFilter.cpp:18-24): cv::cvtColor(CV_BGR2GRAY) on 8-bit frames. OpenCV 2.4.3 (the pinned
 * version; imgproc/src/color.cpp RGB2Gray<uchar>, yuv_shift = 14, B2Y = 1868, G2Y = 9617, R2Y = 4899):
 *     gray = (1868 B + 9617 G + 4899 R + 8192) >> 14
 * Pure streaming: 3 bytes read + 1 byte written per pixel, so this is the one kernel of the path that sits on the HBM roofline.
 * A thread converts 16 pixels: three 16-byte loads (48 interleaved bytes), one 16-byte store; rows are handled as a flat
 * array when pitch == 3 W (always inside the library), with a scalar tail.
 * ------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) { return (1868u * b + 9617u * g + 4899u * r + 8192u) >> 14; }

__global__ void __launch_bounds__(256) bgr2gray_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, int64_t n_px) {
	const int64_t groups = n_px >> 4;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups; gidx += stride) {
		const uint4* src = reinterpret_cast<const uint4*>(bgr) + 3 * gidx;
		const uint4 a = __ldcs(src), b = __ldcs(src + 1), c = __ldcs(src + 2);
		const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
		uint32_t out[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) { /* 4 pixels = 12 bytes = words 3q .. 3q + 2 */
			const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
			const uint32_t p0 = gray_of(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
			const uint32_t p1 = gray_of(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
			const uint32_t p2 = gray_of((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
			const uint32_t p3 = gray_of((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
			out[q] = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
		}
		__stcs(reinterpret_cast<uint4*>(gray) + gidx, make_uint4(out[0], out[1], out[2], out[3]));
	}
	if (blockIdx.x == 0) /* tail: fewer than 16 pixels */
		for (int64_t i = (groups << 4) + threadIdx.x; i < n_px; i += blockDim.x) gray[i] = (uint8_t)gray_of(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
}

void launch_bgr2gray(cudaStream_t st, const uint8_t* bgr, uint8_t* gray, int64_t n_px) {
	if (n_px <= 0) return;
	const int64_t groups = n_px >> 4;
	const int64_t want = (groups + 255) / 256;
	const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want)); /* grid-stride over 148 SMs x 16 resident CTAs */
	bgr2gray_kernel<<<grid, 256, 0, st>>>(bgr, gray, n_px);
}

/* ---------------------------------------------------------------------------------------------
 * launchers
 * ------------------------------------------------------------------------------------------- */
void launch_resize(cudaStream_t st, const uint8_t* frames, int W, int H, int n_frames, uint8_t* arena,
		int64_t arena_stride, const ResizeJob* jobs_dev, int n_jobs, int max_tiles, const int4* xy_tab) {
	if (n_jobs == 0 || n_frames == 0) return;
	dim3 grid((unsigned)max_tiles, (unsigned)n_jobs, (unsigned)n_frames);
	resize_kernel<<<grid, RS_QX * RS_TY, 0, st>>>(frames, W, H, arena, arena_stride, jobs_dev, xy_tab);
}

int resize_tiles(int dst_w, int dst_h) {
	const int quads = (dst_w + 3) / 4;
	return ((quads + RS_QX - 1) / RS_QX) * ((dst_h + RS_TY * RS_ROWS - 1) / (RS_TY * RS_ROWS));
}

void launch_pyrdown(cudaStream_t st, const uint8_t* frames, int W, int H, int n_frames, uint8_t* arena,
		int64_t arena_stride, const DownJob* jobs_dev, int n_jobs, int max_tiles) {
	if (n_jobs == 0 || n_frames == 0) return;
	dim3 grid((unsigned)max_tiles, (unsigned)n_jobs, (unsigned)n_frames);
	pyrdown_kernel<<<grid, PD_BX * PD_BY, 0, st>>>(frames, W, H, arena, arena_stride, jobs_dev);
}

int pyrdown_tiles(int dst_w, int dst_h) {
	return ((dst_w + PD_TW - 1) / PD_TW) * ((dst_h + PD_TH - 1) / PD_TH);
}

} // namespace fdb
