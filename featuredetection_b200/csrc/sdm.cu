/*
 * sdm.cu - supervised-descent landmark regressor on the GPU (sm_100a), BASELINE configs[4]:
 *   SdmLandmarkModelFitting::alignRigid / optimize   libSupervisedDescent/include/superviseddescent/SdmLandmarkModel.hpp:156-256
 *   VlHogDescriptorExtractor::getDescriptors         libSupervisedDescent/include/superviseddescent/DescriptorExtractor.hpp:106-219
 *   vl_hog_put_image / vl_hog_extract (UoCTTI)       libSupervisedDescent/src/superviseddescent/hog.c:595-727,857-1063
 *
 * One cascade step over a batch of faces is three launches:
 *   sdm_hog_kernel     one WARP per (landmark, face): crop (black canvas outside the image, with the reference's
 *                      row-offset quirk), float32 bilinear resize to 30x30, VLFeat HOG (3x3 cells x 31) written
 *                      straight into the face's feature row [face][landmark * 279 + ...]
 *   sdm_gemm_dmma_kernel  delta[faces x 2L] = features[faces x 279 L] * R[0:-1] + R[-1] on the FP64 tensor cores: float32 inputs,
 *                      FLOAT64 accumulation, which is what cv::gemm does for CV_32F - the product of two floats is exact in
 *                      double, so the result agrees with the sequential reference sum to ~1e-16 relative
 *   sdm_update_kernel  shape += delta^T * eye-mouth distance
 * Why FP64 and not bf16/tf32 tensor cores: the fit is a feedback loop through cvRound(landmark) - a 1e-5 px difference in a
 * shape moves a HOG window by a whole pixel with probability ~1e-5 per coordinate, and 2L x steps coordinates per face turn
 * that into visibly different fits for ~1 % of the faces.  The float64-accumulated product stays below that, and on the DMMA
 * pipe [4096 x 18972] x [18972 x 136] costs ~1 ms per step against ~3.8 ms of HOG.
 *
 * Exactness: every float32/float64 operation of the reference is issued in its order with _rn intrinsics (no FMA
 * contraction); histogram cells are accumulated by one thread per (cell, orientation) in pixel raster order.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fdb_internal.h"
#include "api_types.h"
#include "sdm_device.h"

namespace fdb {

#define SDM_THREADS 128
#define SDM_P 30          /* patch side after the resize (3 cells of 10 px) */
#define SDM_CELLS 3
#define SDM_NO 9          /* undirected orientations */
#define SDM_DIM 31        /* 3 * 9 + 4 */
#define SDM_DESC 279

/* eye-mouth distance and HOG window half size of a cascade step (SdmLandmarkModel.hpp:212-229) */
__device__ __forceinline__ void sdm_window(const float* __restrict__ shape, int L, double step_factor, float* d_out, int* wsh_out) {
	const float a1x = __fdiv_rn(__fadd_rn(shape[8], shape[9]), 2.0f), a1y = __fdiv_rn(__fadd_rn(shape[8 + L], shape[9 + L]), 2.0f);
	const float a2x = __fdiv_rn(__fadd_rn(shape[11], shape[12]), 2.0f), a2y = __fdiv_rn(__fadd_rn(shape[11 + L], shape[12 + L]), 2.0f);
	const float dx = __fsub_rn(a1x, a2x), dy = __fsub_rn(a1y, a2y);
	const float d = (float)sqrt(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)));
	float wsh = __fdiv_rn(__fdiv_rn(d, 2.0f), 2.0f);
	wsh = (float)round(__dmul_rn((double)wsh, step_factor));
	const int w = (int)wsh;
	*wsh_out = w + SDM_CELLS - (w % SDM_CELLS);
	*d_out = d;
}

/* pixel of the image extended by a black canvas (copyMakeBorder, DescriptorExtractor.hpp:166-168) */
__device__ __forceinline__ float sdm_px(const uint8_t* __restrict__ img, int W, int H, int x, int y) {
	return (x >= 0 && y >= 0 && x < W && y < H) ? (float)img[(int64_t)y * W + x] : 0.f;
}

/* One WARP per descriptor (landmark of a face); HW_WARPS descriptors per CTA, no CTA-wide barrier after the table load.
 *   pass 1  lane = patch column: rows stream through three registers (resized rows r-1, r, r+1), so the resized patch is
 *           never stored; gradient, dominant orientation (hog.c:617-682); every contributing pixel gets its rank inside
 *           the raster-ordered list of its orientation (__match_any_sync + running counts)
 *   pass 2  scatter pixel ids into the 18 orientation lists (stable: raster order is kept)
 *   pass 3  bilinear cell accumulation (hog.c:697-722): lane = (orientation, cell row); it walks the contiguous list
 *           segment of the rows that touch its cell row and feeds the three cells of that row; columns outside a cell have
 *           weight 0, and acc + 0 == acc exactly, so every cell sees its own pixels in raster order like the reference
 *   pass 4  block normalisation + UoCTTI features (hog.c:879-1060), staged in shared memory and written as one 279-float row
 * mode 0: points come from the face's current shape, window from sdm_window; mode 1: explicit points + window */
#define HW_WARPS 4
#define HW_RS 32 /* row stride of the per-pixel arrays */

struct HogWarpSmem {
	float grad[SDM_P * HW_RS];     /* gradient magnitude per pixel; reused as the output staging row */
	uint16_t bp[SDM_P * HW_RS];    /* (orientation << 10) | rank in the orientation's list; 0xffff: pixel contributes nothing */
	uint16_t list[SDM_P * SDM_P];  /* pixel ids (y * 32 + x), the 18 orientation lists back to back */
	float hog[SDM_CELLS * SDM_CELLS * 2 * SDM_NO]; /* [o][cy][cx] */
	float norm[12];
	double fac[SDM_CELLS * SDM_CELLS][4];
	int cnt[2 * SDM_NO];
	int seg[3][2 * SDM_NO];        /* list fill at the start of rows 5, 15, 25 */
	int base[2 * SDM_NO + 2];
};

__global__ void __launch_bounds__(HW_WARPS * 32) sdm_hog_kernel(const DevSdm m, const uint8_t* __restrict__ frames, int W, int H,
		const int* __restrict__ face_frame, const float* __restrict__ shapes, int step, const float* __restrict__ pts_xy,
		int window_half, int n_desc, float* __restrict__ features, int* __restrict__ status) {
	__shared__ HogWarpSmem s_all[HW_WARPS];
	__shared__ float4 s_wx[SDM_P];            /* column weight for cell columns 0, 1, 2 */
	__shared__ float s_wy[SDM_CELLS][SDM_P];
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, L = m.L;
	if (threadIdx.x < SDM_P) {
		const int x = threadIdx.x, b = m.bin_of[x];
		const float w1 = m.w1_of[x], w2 = m.w2_of[x];
		s_wx[x] = make_float4(b == 0 ? w1 : (b + 1 == 0 ? w2 : 0.f), b == 1 ? w1 : (b + 1 == 1 ? w2 : 0.f), b == 2 ? w1 : (b + 1 == 2 ? w2 : 0.f), 0.f);
		for (int c = 0; c < SDM_CELLS; ++c) s_wy[c][x] = b == c ? w1 : (b + 1 == c ? w2 : 0.f);
	}
	__syncthreads();
	const int desc = blockIdx.x * HW_WARPS + warp;
	if (desc >= n_desc) return;
	const int face = desc / L, lm = desc - face * L;
	if (status && status[face] != 0) return; /* the reference threw at an earlier cascade step of this face */
	HogWarpSmem& s = s_all[warp];
	const uint8_t* __restrict__ img = frames + (int64_t)(face_frame ? face_frame[face] : 0) * W * H;
	float* __restrict__ out = features + (int64_t)desc * SDM_DESC;

	/* ---- window geometry (DescriptorExtractor.hpp:157-173), computed by every lane ---- */
	int x0, y0, side;
	{
		float px, py; int wsh;
		if (pts_xy) { px = pts_xy[2 * (int64_t)desc]; py = pts_xy[2 * (int64_t)desc + 1]; wsh = window_half; }
		else {
			const float* shape = shapes + (int64_t)face * 2 * L;
			float d;
			sdm_window(shape, L, m.step_factor[step], &d, &wsh);
			px = shape[lm]; py = shape[lm + L];
		}
		const int x = __float2int_rn(px), y = __float2int_rn(py); /* cvRound */
		int rx = x - wsh, ry = y - wsh, bl = 0, bt = 0, br = 0, bb = 0;
		if (x - wsh < 0 || y - wsh < 0 || x + wsh >= W || y + wsh >= H) {
			bl = (x - wsh) < 0 ? abs(x - wsh) : 0;
			bt = (y - wsh) < 0 ? abs(y - wsh) : 0;
			br = (x + wsh) >= W ? abs(W - (x + wsh)) : 0;
			bb = (y + wsh) >= H ? abs(H - (y + wsh)) : 0;
			rx = (x - wsh) + bl;
			ry = (y - wsh) + br; /* sic (:169) */
		}
		side = 2 * wsh;
		const bool ok = side >= 4 && rx >= 0 && ry >= 0 && rx + side <= W + bl + br && ry + side <= H + bt + bb;
		x0 = rx - bl; y0 = ry - bt;
		if (!ok) { /* the reference's Mat::operator()(roi) would throw: flag the face with step + 1, emit zeros */
			if (lane == 0 && status) atomicCAS(status + face, 0, step + 1);
			for (int i = lane; i < SDM_DESC; i += 32) out[i] = 0.f;
			return;
		}
	}
	const bool inside = x0 >= 0 && y0 >= 0 && x0 + side <= W && y0 + side <= H;
	const int mode = side == SDM_P ? 0 : (side == 2 * SDM_P ? 1 : 2);
	const int x = lane;
	const bool xin = x < SDM_P;

	/* ---- cv::resize(CV_32F, INTER_LINEAR): column coefficients per lane, row coefficients per (warp-uniform) row ---- */
	const double scale = __ddiv_rn((double)side, (double)SDM_P);
	int sx = 0, sx1 = 0;
	float a0 = 1.f, a1 = 0.f;
	if (mode == 2 && xin) {
		float f = (float)__dsub_rn(__dmul_rn((double)x + 0.5, scale), 0.5);
		int si = (int)floorf(f);
		f = __fsub_rn(f, (float)si);
		if (si < 0) { f = 0.f; si = 0; }                 /* columns: coefficient and index clamped */
		if (si >= side - 1) { f = 0.f; si = side - 1; }
		sx = si; sx1 = min(si + 1, side - 1);
		a0 = __fsub_rn(1.f, f); a1 = f;
	}
	/* source columns / rows of the (up to) four pixels behind one resized pixel */
	const int cA = x0 + (mode == 0 ? x : (mode == 1 ? 2 * x : sx)), cB = x0 + (mode == 0 ? x : (mode == 1 ? 2 * x + 1 : sx1));
	struct Row { int y0, y1; float b0, b1; };
	auto row_of = [&](int r) -> Row { /* rows: the coefficient is kept, only the row index is clamped */
		Row q;
		if (mode == 0) { q.y0 = q.y1 = y0 + r; q.b0 = 1.f; q.b1 = 0.f; return q; }
		if (mode == 1) { q.y0 = y0 + 2 * r; q.y1 = q.y0 + 1; q.b0 = 1.f; q.b1 = 0.f; return q; }
		float f = (float)__dsub_rn(__dmul_rn((double)r + 0.5, scale), 0.5);
		const int si = (int)floorf(f);
		f = __fsub_rn(f, (float)si);
		q.y0 = y0 + min(max(si, 0), side - 1); q.y1 = y0 + min(max(si + 1, 0), side - 1);
		q.b0 = __fsub_rn(1.f, f); q.b1 = f;
		return q;
	};
	auto px = [&](int xx, int yy) -> uint32_t { /* black canvas outside the image (copyMakeBorder) */
		return (xx >= 0 && yy >= 0 && xx < W && yy < H) ? (uint32_t)__ldg(img + (int64_t)yy * W + xx) : 0u;
	};
	auto load_raw = [&](const Row& q) -> uint32_t { /* the four source pixels, packed; issued one row ahead of their use */
		if (!xin) return 0u;
		if (inside) {
			const uint8_t* r0 = img + (int64_t)q.y0 * W;
			if (mode == 0) return (uint32_t)__ldg(r0 + cA);
			const uint8_t* r1 = img + (int64_t)q.y1 * W;
			return (uint32_t)__ldg(r0 + cA) | ((uint32_t)__ldg(r0 + cB) << 8) | ((uint32_t)__ldg(r1 + cA) << 16) | ((uint32_t)__ldg(r1 + cB) << 24);
		}
		if (mode == 0) return px(cA, q.y0);
		return px(cA, q.y0) | (px(cB, q.y0) << 8) | (px(cA, q.y1) << 16) | (px(cB, q.y1) << 24);
	};
	auto finish = [&](uint32_t raw, const Row& q) -> float { /* resized pixel from its packed sources */
		const float p00 = (float)(raw & 255u), p01 = (float)((raw >> 8) & 255u), p10 = (float)((raw >> 16) & 255u), p11 = (float)(raw >> 24);
		if (mode == 0) return p00;
		if (mode == 1) return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p00, p01), p10), p11), 0.25f); /* INTER_AREA fast path of an exact 2x decimation */
		const float r0 = __fadd_rn(__fmul_rn(p00, a0), __fmul_rn(p01, a1));
		const float r1 = __fadd_rn(__fmul_rn(p10, a0), __fmul_rn(p11, a1));
		return __fadd_rn(__fmul_rn(r0, q.b0), __fmul_rn(r1, q.b1));
	};

	/* ---- pass 1: gradient + dominant directed orientation per pixel, ranks inside the orientation lists ---- */
	if (lane < 2 * SDM_NO) s.cnt[lane] = 0;
	__syncwarp();
	const bool xvalid = x >= 1 && x < SDM_P - 1;
	float rowA, rowB;
	{ const Row q0 = row_of(0), q1 = row_of(1); const uint32_t w0 = load_raw(q0), w1 = load_raw(q1); rowA = finish(w0, q0); rowB = finish(w1, q1); }
	Row qn = row_of(2);
	uint32_t pending = load_raw(qn);
	for (int r = 1; r < SDM_P - 1; ++r) {
		const Row qc = qn;
		const uint32_t rawc = pending;
		if (r + 2 < SDM_P) { qn = row_of(r + 2); pending = load_raw(qn); } /* prefetch: consumed in the next iteration */
		const float rowC = finish(rawc, qc);
		const float left = __shfl_up_sync(FULL, rowB, 1), right = __shfl_down_sync(FULL, rowB, 1);
		int key = 31;
		float grad = 0.f;
		if (xvalid) {
			float gx = __fsub_rn(right, left), gy = __fsub_rn(rowC, rowA);
			float g2 = __fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy));
			if (g2 > 0.f) {
				grad = __fsqrt_rn(g2);
				if (grad > 1e-10f) { /* float / double rounded to float == float / float (53 >= 2 * 24 + 2: no double rounding) */
					gx = __fdiv_rn(gx, grad); gy = __fdiv_rn(gy, grad);
				} else {
					const double den = (double)grad > 1e-10 ? (double)grad : 1e-10;
					gx = (float)__ddiv_rn((double)gx, den); gy = (float)__ddiv_rn((double)gy, den);
				}
				float best = 0.f;
				int bk = -1;
#pragma unroll
				for (int k = 0; k < SDM_NO; ++k) { /* first strict maximum of |score| (hog.c:656-672) */
					const float score = __fadd_rn(__fmul_rn(gx, m.ox[k]), __fmul_rn(gy, m.oy[k]));
					if (fabsf(score) > fabsf(best)) { best = score; bk = k; }
				}
				if (bk >= 0) key = best < 0.f ? bk + SDM_NO : bk;
			}
		}
		if (r == 5 || r == 15 || r == 25) { if (lane < 2 * SDM_NO) s.seg[r / 10][lane] = s.cnt[lane]; }
		const unsigned grp = __match_any_sync(FULL, key);
		const int rank = __popc(grp & ((1u << lane) - 1u));
		const int fill = key < 2 * SDM_NO ? s.cnt[key] : 0;
		__syncwarp();
		if (key < 2 * SDM_NO) {
			s.bp[r * HW_RS + x] = (uint16_t)((key << 10) | (fill + rank));
			s.grad[r * HW_RS + x] = grad;
			if (rank == 0) s.cnt[key] = fill + __popc(grp);
		} else if (xin) s.bp[r * HW_RS + x] = 0xffffu;
		__syncwarp();
		rowA = rowB; rowB = rowC;
	}
	/* ---- pass 2: exclusive prefix over the 18 list lengths, scatter the pixel ids ---- */
	{
		const int c = lane < 2 * SDM_NO ? s.cnt[lane] : 0;
		int incl = c;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += v; }
		if (lane < 2 * SDM_NO) s.base[lane] = incl - c;
		__syncwarp();
		if (xvalid)
			for (int r = 1; r < SDM_P - 1; ++r) {
				const unsigned v = s.bp[r * HW_RS + x];
				if (v != 0xffffu) s.list[s.base[v >> 10] + (v & 1023u)] = (uint16_t)(r * HW_RS + x);
			}
		__syncwarp();
	}
	/* ---- pass 3: cell accumulation; pair p = (orientation o, cell row cy) ---- */
	for (int p = lane; p < 2 * SDM_NO * SDM_CELLS; p += 32) {
		const int o = p / SDM_CELLS, cy = p - o * SDM_CELLS;
		const int b = s.base[o];
		const int beg = b + (cy == 0 ? 0 : s.seg[cy - 1][o]);            /* rows 1.., 5.., 15.. */
		const int end = b + (cy == 2 ? s.cnt[o] : s.seg[cy + 1][o]);     /* ..14, ..24, ..28 */
		float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
		for (int i = beg; i < end; ++i) {
			const int id = s.list[i];
			const float g = s.grad[id];
			const float4 wx = s_wx[id & 31];
			const float wy = s_wy[cy][id >> 5];
			acc0 = __fadd_rn(acc0, __fmul_rn(__fmul_rn(g, wx.x), wy));
			acc1 = __fadd_rn(acc1, __fmul_rn(__fmul_rn(g, wx.y), wy));
			acc2 = __fadd_rn(acc2, __fmul_rn(__fmul_rn(g, wx.z), wy));
		}
		float* h = s.hog + o * 9 + cy * 3;
		h[0] = acc0; h[1] = acc1; h[2] = acc2;
	}
	__syncwarp();
	/* ---- pass 4: vl_hog_extract (hog.c:879-1060) ---- */
	if (lane < SDM_CELLS * SDM_CELLS) {
		float nrm = 0.f;
		for (int k = 0; k < SDM_NO; ++k) {
			const float h = __fadd_rn(s.hog[k * 9 + lane], s.hog[(k + SDM_NO) * 9 + lane]);
			nrm = __fadd_rn(nrm, __fmul_rn(h, h));
		}
		s.norm[lane] = nrm;
	}
	__syncwarp();
	if (lane < SDM_CELLS * SDM_CELLS * 4) { /* (cell, block) -> 1 / sqrt(sum of the block's four cell norms + 1e-4) */
		const int c = lane >> 2, q = lane & 3, y = c / SDM_CELLS, xx = c - y * SDM_CELLS;
		const int xa = q & 1 ? xx : max(xx - 1, 0), xb = q & 1 ? min(xx + 1, SDM_CELLS - 1) : xx;   /* block columns, left to right */
		const int ya = q & 2 ? y : max(y - 1, 0), yb = q & 2 ? min(y + 1, SDM_CELLS - 1) : y;
		const double n_aa = s.norm[ya * SDM_CELLS + xa], n_ab = s.norm[ya * SDM_CELLS + xb];
		const double n_ba = s.norm[yb * SDM_CELLS + xa], n_bb = s.norm[yb * SDM_CELLS + xb];
		s.fac[c][q] = __ddiv_rn(1.0, sqrt(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(n_aa, n_ab), n_ba), n_bb), 1e-4)));
	}
	if (lane + 32 < SDM_CELLS * SDM_CELLS * 4) {
		const int e = lane + 32, c = e >> 2, q = e & 3, y = c / SDM_CELLS, xx = c - y * SDM_CELLS;
		const int xa = q & 1 ? xx : max(xx - 1, 0), xb = q & 1 ? min(xx + 1, SDM_CELLS - 1) : xx;
		const int ya = q & 2 ? y : max(y - 1, 0), yb = q & 2 ? min(y + 1, SDM_CELLS - 1) : y;
		const double n_aa = s.norm[ya * SDM_CELLS + xa], n_ab = s.norm[ya * SDM_CELLS + xb];
		const double n_ba = s.norm[yb * SDM_CELLS + xa], n_bb = s.norm[yb * SDM_CELLS + xb];
		s.fac[c][q] = __ddiv_rn(1.0, sqrt(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(n_aa, n_ab), n_ba), n_bb), 1e-4)));
	}
	__syncwarp();
	float* stage = s.grad; /* descriptor layout (DescriptorExtractor.hpp:196-204): dimension j, then cell column, then cell row */
	for (int e = lane; e < SDM_CELLS * SDM_CELLS * SDM_NO; e += 32) {
		const int c = e / SDM_NO, k = e - c * SDM_NO, cy = c / SDM_CELLS, cx = c - cy * SDM_CELLS;
		const double ha = s.hog[k * 9 + c], hb = s.hog[(k + SDM_NO) * 9 + c];
		double sa = 0, sb = 0, sc = 0;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const double f = s.fac[c][q];
			const double haq = __dmul_rn(f, ha), hbq = __dmul_rn(f, hb);
			const double hcm = fmin(0.2, __dadd_rn(haq, hbq));
			sa = q == 0 ? fmin(0.2, haq) : __dadd_rn(sa, fmin(0.2, haq));
			sb = q == 0 ? fmin(0.2, hbq) : __dadd_rn(sb, fmin(0.2, hbq));
			sc = q == 0 ? hcm : __dadd_rn(sc, hcm);
		}
		const int pos = cx * SDM_CELLS + cy;
		stage[k * 9 + pos] = (float)__dmul_rn(0.5, sa);
		stage[(k + SDM_NO) * 9 + pos] = (float)__dmul_rn(0.5, sb);
		stage[(k + 2 * SDM_NO) * 9 + pos] = (float)__dmul_rn(0.5, sc);
	}
	for (int e = lane; e < SDM_CELLS * SDM_CELLS * 4; e += 32) { /* texture features: t_q = sum over k of the clamped hc (hog.c:1012-1015,1046-1049) */
		const int c = e >> 2, q = e & 3, cy = c / SDM_CELLS, cx = c - cy * SDM_CELLS;
		const double f = s.fac[c][q];
		double t = 0;
		for (int k = 0; k < SDM_NO; ++k) {
			const double ha = s.hog[k * 9 + c], hb = s.hog[(k + SDM_NO) * 9 + c];
			t = __dadd_rn(t, fmin(0.2, __dadd_rn(__dmul_rn(f, ha), __dmul_rn(f, hb))));
		}
		stage[(3 * SDM_NO + q) * 9 + cx * SDM_CELLS + cy] = (float)__dmul_rn((double)m.tex, t);
	}
	__syncwarp();
	for (int i = lane; i < SDM_DESC; i += 32) out[i] = stage[i];
}

/* ------------------------------------------------------------------------------------------------
 * delta = F * R[0:K] + R[K] on the FP64 tensor cores (DMMA).
 *
 * cv::gemm on CV_32F accumulates float x float products in double (SdmLandmarkModel.hpp:241 is a MatExpr -> cv::gemm);
 * a float x float product is exact in double, so a double-accumulating GEMM in any summation order agrees with the
 * sequential reference sum to ~1e-16 relative - below one float32 ulp of the result except with probability ~1e-9.
 * mma.sync ... f64 keeps that: operands are converted float -> double once while they are staged into shared memory.
 *
 * CTA tile: 32 faces x (8 NT8) columns, 4 warps: warp w owns rows 16 (w & 1) .. +15 and the k16 block (w >> 1) of every
 * 32-wide k stage; the two k halves are added in a fixed order at the end (deterministic).  Global loads of the next stage
 * are issued into registers before the current stage is multiplied (register double buffering).
 * Shared layout (doubles): A [32][36], B [32][8 NT8 + 12]: both strides = 4 mod 16, so the 16 lanes of a half warp
 * (g = 0..3, t = 0..3 -> g * stride + t) hit 16 distinct 8-byte banks.
 * ---------------------------------------------------------------------------------------------- */
#define GM_BM 32
#define GM_BK 32
#define GM_AS 36 /* A row stride in doubles */

__device__ __forceinline__ void dmma_m16n8k16(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
	asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
			: "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
			: "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

/* VEC: rows of F and R are 16-byte aligned (K % 4 == 0 and N % 4 == 0); K16: use the m16n8k16 shape (else m8n8k4) */
template <int NT8, bool VEC, bool K16>
__global__ void __launch_bounds__(128) sdm_gemm_dmma_kernel(const float* __restrict__ F, const float* __restrict__ R, int n_faces, int K, int N,
		float* __restrict__ delta) {
	constexpr int BN = 8 * NT8, BS = BN + 12;
	constexpr int A_V4 = GM_BM * GM_BK / 4 / 128;                /* float4 loads of A per thread and stage (2) */
	constexpr int B_V4 = (GM_BK * BN / 4 + 127) / 128;           /* float4 loads of B per thread and stage */
	extern __shared__ __align__(16) double gm_smem[];
	double* sA = gm_smem;                 /* [32][GM_AS] */
	double* sB = gm_smem + GM_BM * GM_AS; /* [32][BS] */
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
	const int mh = warp & 1, kh = warp >> 1;
	const int face0 = blockIdx.x * GM_BM, col0 = blockIdx.y * BN;
	double acc[NT8][4];
#pragma unroll
	for (int j = 0; j < NT8; ++j) { acc[j][0] = 0; acc[j][1] = 0; acc[j][2] = 0; acc[j][3] = 0; }
	float4 ra[A_V4], rb[B_V4];

	auto load_stage = [&](int k0) {
#pragma unroll
		for (int i = 0; i < A_V4; ++i) { /* A: 32 rows x 8 float4 */
			const int idx = tid + i * 128, row = idx >> 3, c4 = (idx & 7) * 4;
			const int f = face0 + row, k = k0 + c4;
			float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
			if (f < n_faces) {
				const float* src = F + (int64_t)f * K + k;
				if (VEC && k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(src));
				else { if (k < K) v.x = __ldg(src); if (k + 1 < K) v.y = __ldg(src + 1); if (k + 2 < K) v.z = __ldg(src + 2); if (k + 3 < K) v.w = __ldg(src + 3); }
			}
			ra[i] = v;
		}
#pragma unroll
		for (int i = 0; i < B_V4; ++i) { /* B: 32 k rows x BN / 4 float4 */
			const int idx = tid + i * 128, row = idx / (BN / 4), c4 = (idx - row * (BN / 4)) * 4;
			const int k = k0 + row, c = col0 + c4;
			float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
			if (row < GM_BK && k < K) {
				const float* src = R + (int64_t)k * N + c;
				if (VEC && c + 3 < N) v = __ldg(reinterpret_cast<const float4*>(src));
				else { if (c < N) v.x = __ldg(src); if (c + 1 < N) v.y = __ldg(src + 1); if (c + 2 < N) v.z = __ldg(src + 2); if (c + 3 < N) v.w = __ldg(src + 3); }
			}
			rb[i] = v;
		}
	};
	auto store_stage = [&]() {
#pragma unroll
		for (int i = 0; i < A_V4; ++i) {
			const int idx = tid + i * 128, row = idx >> 3, c4 = (idx & 7) * 4;
			double2* d = reinterpret_cast<double2*>(sA + row * GM_AS + c4);
			d[0] = make_double2((double)ra[i].x, (double)ra[i].y);
			d[1] = make_double2((double)ra[i].z, (double)ra[i].w);
		}
#pragma unroll
		for (int i = 0; i < B_V4; ++i) {
			const int idx = tid + i * 128, row = idx / (BN / 4), c4 = (idx - row * (BN / 4)) * 4;
			if (row < GM_BK) {
				double2* d = reinterpret_cast<double2*>(sB + row * BS + c4);
				d[0] = make_double2((double)rb[i].x, (double)rb[i].y);
				d[1] = make_double2((double)rb[i].z, (double)rb[i].w);
			}
		}
	};

	load_stage(0);
	for (int k0 = 0; k0 < K; k0 += GM_BK) {
		store_stage();
		__syncthreads();
		if (k0 + GM_BK < K) load_stage(k0 + GM_BK);
		const double* A0 = sA + (mh * 16 + g) * GM_AS + kh * 16 + t;
		const double* B0 = sB + (kh * 16 + t) * BS + g;
		if (K16) {
			double a[8];
#pragma unroll
			for (int i = 0; i < 8; ++i) a[i] = A0[(i & 1) * 8 * GM_AS + (i >> 1) * 4];
#pragma unroll
			for (int j = 0; j < NT8; ++j) {
				double b[4];
#pragma unroll
				for (int i = 0; i < 4; ++i) b[i] = B0[i * 4 * BS + 8 * j];
				dmma_m16n8k16(acc[j], a, b);
			}
		} else {
#pragma unroll
			for (int q = 0; q < 4; ++q) { /* four k4 steps of this warp's k16 block */
				const double a_lo = A0[q * 4], a_hi = A0[8 * GM_AS + q * 4];
#pragma unroll
				for (int j = 0; j < NT8; ++j) {
					const double b = B0[q * 4 * BS + 8 * j];
					dmma_m8n8k4(acc[j][0], acc[j][1], a_lo, b);
					dmma_m8n8k4(acc[j][2], acc[j][3], a_hi, b);
				}
			}
		}
		__syncthreads();
	}
	/* add the two k halves in a fixed order (kh = 0 first), then the bias row, round to float32 once */
	double* red = gm_smem; /* [2 m halves][NT8][4][32 lanes] */
	if (kh == 1) {
#pragma unroll
		for (int j = 0; j < NT8; ++j)
#pragma unroll
			for (int i = 0; i < 4; ++i) red[((mh * NT8 + j) * 4 + i) * 32 + lane] = acc[j][i];
	}
	__syncthreads();
	if (kh == 0) {
#pragma unroll
		for (int j = 0; j < NT8; ++j)
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const int f = face0 + mh * 16 + g + (i >> 1) * 8, c = col0 + 8 * j + 2 * t + (i & 1);
				if (f < n_faces && c < N) {
					const double sum = __dadd_rn(acc[j][i], red[((mh * NT8 + j) * 4 + i) * 32 + lane]);
					delta[(int64_t)f * N + c] = (float)__dadd_rn(sum, (double)R[(int64_t)K * N + c]);
				}
			}
	}
}

template <int NT8>
static size_t gemm_smem_bytes() {
	const size_t stage = sizeof(double) * (GM_BM * GM_AS + GM_BK * (8 * NT8 + 12));
	const size_t red = sizeof(double) * 2 * NT8 * 4 * 32;
	return stage > red ? stage : red;
}

/* scalar FP64 FMA version (reference for the tensor-core kernel; FDB_SDM_GEMM=scalar): BM faces per CTA, thread (tx, ty)
 * owns faces ty*4..+3 and columns tx + 16 j */
#define GEMM_BM 32
#define GEMM_BK 16
template <int NC>
__global__ void __launch_bounds__(128) sdm_gemm_kernel(const float* __restrict__ F, const float* __restrict__ R, int n_faces, int K, int N,
		float* __restrict__ delta) {
	__shared__ float s_f[GEMM_BK][GEMM_BM + 1];
	__shared__ float s_r[GEMM_BK][16 * NC];
	const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
	const int face0 = blockIdx.x * GEMM_BM;
	double acc[4][NC];
#pragma unroll
	for (int i = 0; i < 4; ++i)
#pragma unroll
		for (int j = 0; j < NC; ++j) acc[i][j] = 0.0;
	for (int k0 = 0; k0 < K; k0 += GEMM_BK) {
		for (int i = tid; i < GEMM_BM * GEMM_BK; i += 128) {
			const int f = i / GEMM_BK, k = i - f * GEMM_BK;
			s_f[k][f] = (face0 + f < n_faces && k0 + k < K) ? F[(int64_t)(face0 + f) * K + k0 + k] : 0.f;
		}
		for (int i = tid; i < GEMM_BK * 16 * NC; i += 128) {
			const int k = i / (16 * NC), c = i - k * (16 * NC);
			s_r[k][c] = (c < N && k0 + k < K) ? R[(int64_t)(k0 + k) * N + c] : 0.f;
		}
		__syncthreads();
#pragma unroll 4
		for (int k = 0; k < GEMM_BK; ++k) { /* k ascending: the reference's accumulation order; float x float is exact in double */
			double a[4], b[NC];
#pragma unroll
			for (int i = 0; i < 4; ++i) a[i] = (double)s_f[k][ty * 4 + i];
#pragma unroll
			for (int j = 0; j < NC; ++j) b[j] = (double)s_r[k][tx + 16 * j];
#pragma unroll
			for (int i = 0; i < 4; ++i)
#pragma unroll
				for (int j = 0; j < NC; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
		}
		__syncthreads();
	}
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int f = face0 + ty * 4 + i;
		if (f >= n_faces) continue;
#pragma unroll
		for (int j = 0; j < NC; ++j) {
			const int c = tx + 16 * j;
			if (c < N) delta[(int64_t)f * N + c] = (float)__dadd_rn(acc[i][j], (double)R[(int64_t)K * N + c]);
		}
	}
}

__global__ void sdm_update_kernel(const DevSdm m, int step, const float* __restrict__ delta, float* __restrict__ shapes,
		const int* __restrict__ status, int n_faces) {
	const int face = blockIdx.x * blockDim.x + threadIdx.x;
	if (face >= n_faces || status[face] != 0) return; /* the reference threw in getDescriptors: the shape keeps its value */
	float* shape = shapes + (int64_t)face * 2 * m.L;
	float d; int wsh;
	sdm_window(shape, m.L, m.step_factor[step], &d, &wsh);
	for (int j = 0; j < 2 * m.L; ++j) shape[j] = __fadd_rn(shape[j], __fmul_rn(delta[(int64_t)face * 2 * m.L + j], d)); /* :243 */
}

/* ------------------------------------------------------------------------------------------------
 * host side
 * ---------------------------------------------------------------------------------------------- */
void sdm_fill_tables(DevSdm* m, int L, int steps) {
	m->L = L; m->steps = steps; m->K = L * SDM_DESC; m->N = 2 * L;
	for (int o = 0; o < SDM_NO; ++o) { /* hog.c:193-202 */
		const double angle = o * 3.141592653589793 / SDM_NO;
		m->ox[o] = (float)std::cos(angle); m->oy[o] = (float)std::sin(angle);
	}
	for (int x = 0; x < SDM_P; ++x) { /* hog.c:697-708 with cellSize 10 */
		const float hx = (x + 0.5) / 10 - 0.5;
		const int bin = (int)std::floor(hx);
		const float w2 = hx - bin;
		const float w1 = 1.0 - w2;
		m->bin_of[x] = bin; m->w1_of[x] = w1; m->w2_of[x] = w2;
	}
	m->tex = 1.0f / std::sqrt(18.0f);
	for (int s = 0; s < steps && s < SDM_MAX_STEPS; ++s)
		m->step_factor[s] = 1 / (1 + std::exp((double)((s + 1) - steps))); /* SdmLandmarkModel.hpp:226 */
}

void launch_sdm_hog(cudaStream_t st, const DevSdm& m, const uint8_t* frames, int W, int H, const int* face_frame, const float* shapes,
		int step, const float* pts_xy, int window_half, int n_faces, float* features, int* status) {
	if (n_faces == 0) return;
	const int n_desc = n_faces * m.L;
	sdm_hog_kernel<<<(unsigned)((n_desc + HW_WARPS - 1) / HW_WARPS), HW_WARPS * 32, 0, st>>>(m, frames, W, H, face_frame, shapes, step, pts_xy,
			window_half, n_desc, features, status);
}

static int gemm_mode() { /* FDB_SDM_GEMM = scalar | m8 | m16 (default): tuning / cross-check hook */
	static int mode = -1;
	if (mode < 0) {
		const char* e = getenv("FDB_SDM_GEMM");
		mode = !e ? 2 : (!strcmp(e, "scalar") ? 0 : (!strcmp(e, "m8") ? 1 : 2));
	}
	return mode;
}

template <int NT8, bool VEC, bool K16>
static void launch_dmma(cudaStream_t st, const DevSdm& m, int step, const float* features, int n_faces, float* delta) {
	static bool configured = false;
	const size_t smem = gemm_smem_bytes<NT8>();
	if (!configured) { cudaFuncSetAttribute(sdm_gemm_dmma_kernel<NT8, VEC, K16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = true; }
	dim3 grid((unsigned)((n_faces + GM_BM - 1) / GM_BM), (unsigned)((m.N + 8 * NT8 - 1) / (8 * NT8)));
	sdm_gemm_dmma_kernel<NT8, VEC, K16><<<grid, 128, smem, st>>>(features, m.R[step], n_faces, m.K, m.N, delta);
}

template <int NT8>
static void launch_dmma_nt(cudaStream_t st, const DevSdm& m, int step, const float* features, int n_faces, float* delta, bool k16) {
	const bool vec = (m.K % 4 == 0) && (m.N % 4 == 0);
	if (vec) { if (k16) launch_dmma<NT8, true, true>(st, m, step, features, n_faces, delta); else launch_dmma<NT8, true, false>(st, m, step, features, n_faces, delta); }
	else { if (k16) launch_dmma<NT8, false, true>(st, m, step, features, n_faces, delta); else launch_dmma<NT8, false, false>(st, m, step, features, n_faces, delta); }
}

void launch_sdm_gemm(cudaStream_t st, const DevSdm& m, int step, const float* features, int n_faces, float* delta) {
	if (n_faces == 0) return;
	const int mode = gemm_mode();
	if (mode == 0) {
		const unsigned grid = (unsigned)((n_faces + GEMM_BM - 1) / GEMM_BM);
		const int nc = (m.N + 15) / 16;
		if (nc <= 2) sdm_gemm_kernel<2><<<grid, 128, 0, st>>>(features, m.R[step], n_faces, m.K, m.N, delta);
		else if (nc <= 5) sdm_gemm_kernel<5><<<grid, 128, 0, st>>>(features, m.R[step], n_faces, m.K, m.N, delta);
		else if (nc <= 9) sdm_gemm_kernel<9><<<grid, 128, 0, st>>>(features, m.R[step], n_faces, m.K, m.N, delta);
		else launch_dmma_nt<9>(st, m, step, features, n_faces, delta, true);
		return;
	}
	/* column tile: 72 (two tiles cover the 136 columns of a 68-landmark model), 32 for small models */
	if (m.N <= 32) launch_dmma_nt<4>(st, m, step, features, n_faces, delta, mode == 2);
	else launch_dmma_nt<9>(st, m, step, features, n_faces, delta, mode == 2);
}

void launch_sdm_update(cudaStream_t st, const DevSdm& m, int step, const float* delta, float* shapes, const int* status, int n_faces) {
	if (n_faces == 0) return;
	sdm_update_kernel<<<(unsigned)((n_faces + 127) / 128), 128, 0, st>>>(m, step, delta, shapes, status, n_faces);
}

} // namespace fdb

/* ------------------------------------------------------------------------------------------------
 * C ABI (include/fdb200.h, "Supervised-descent landmark regressor")
 * ---------------------------------------------------------------------------------------------- */
#define SDM_CHUNKS 4
#define SDM_CHUNK_MIN_BYTES (64ll << 20) /* pipeline the upload only when the frames are big enough to matter */

struct fdb_sdm {
	fdb_ctx* ctx = nullptr;
	fdb::DevSdm dev{};
	std::vector<float> mean;
	std::vector<void*> owned;     /* model */
	std::vector<void*> work;      /* workspace, grows with the batch */
	int64_t cap_faces = 0, cap_frame_bytes = 0;
	float* d_features = nullptr;  /* [faces][K] */
	float* d_delta = nullptr;     /* [faces][N] */
	float* d_shapes = nullptr;    /* [faces][N] (host-call staging) */
	int* d_status = nullptr;      /* [faces] */
	int* d_face_frame = nullptr;  /* [faces] */
	uint8_t* d_frames = nullptr;  /* host-call staging */
	cudaEvent_t ev[2 * SDM_MAX_STEPS * 3 + 2] = {};
	cudaStream_t chunk_stream[SDM_CHUNKS] = {};  /* host-call pipeline: upload of frame chunk c + 1 overlaps the fit of chunk c */
	cudaEvent_t chunk_ev[SDM_CHUNKS + 1] = {};
};

using namespace fdb;

namespace {

int sdm_reserve(fdb_sdm* m, int64_t n_faces, int64_t frame_bytes) {
	if (n_faces > m->cap_faces) {
		for (void* p : {(void*)m->d_features, (void*)m->d_delta, (void*)m->d_shapes, (void*)m->d_status, (void*)m->d_face_frame}) if (p) cudaFree(p);
		m->d_features = nullptr; m->d_delta = nullptr; m->d_shapes = nullptr; m->d_status = nullptr; m->d_face_frame = nullptr;
		m->cap_faces = 0;
		CUDA_TRY(cudaMalloc((void**)&m->d_features, sizeof(float) * (size_t)n_faces * m->dev.K));
		CUDA_TRY(cudaMalloc((void**)&m->d_delta, sizeof(float) * (size_t)n_faces * m->dev.N));
		CUDA_TRY(cudaMalloc((void**)&m->d_shapes, sizeof(float) * (size_t)n_faces * m->dev.N));
		CUDA_TRY(cudaMalloc((void**)&m->d_status, sizeof(int) * (size_t)n_faces));
		CUDA_TRY(cudaMalloc((void**)&m->d_face_frame, sizeof(int) * (size_t)n_faces));
		m->cap_faces = n_faces;
	}
	if (frame_bytes > m->cap_frame_bytes) {
		if (m->d_frames) cudaFree(m->d_frames);
		m->d_frames = nullptr; m->cap_frame_bytes = 0;
		CUDA_TRY(cudaMalloc((void**)&m->d_frames, (size_t)frame_bytes));
		m->cap_frame_bytes = frame_bytes;
	}
	return FDB_OK;
}

/* the cascade (SdmLandmarkModel.hpp:231-249) on device-resident data; features_host: NULL or [steps][faces][K];
 * ev: NULL or 6 * steps + 2 events recorded around each kernel family */
int sdm_run(fdb_sdm* m, const uint8_t* d_frames, int W, int H, const int* d_face_frame, int64_t n_faces, float* d_shapes, int* d_status,
		float* features_host, cudaEvent_t* ev, cudaStream_t st = nullptr, int64_t first_face = 0) {
	if (!st) st = m->ctx->stream;
	int* status = d_status ? d_status : m->d_status + first_face;
	float* const features = m->d_features + first_face * m->dev.K; /* workspace rows of this face range */
	float* const delta = m->d_delta + first_face * m->dev.N;
	CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int) * (size_t)n_faces, st));
	if (ev) CUDA_TRY(cudaEventRecord(ev[0], st));
	for (int s = 0; s < m->dev.steps; ++s) {
		launch_sdm_hog(st, m->dev, d_frames, W, H, d_face_frame, d_shapes, s, nullptr, 0, (int)n_faces, features, status);
		if (ev) CUDA_TRY(cudaEventRecord(ev[1 + 3 * s], st));
		launch_sdm_gemm(st, m->dev, s, features, (int)n_faces, delta);
		if (ev) CUDA_TRY(cudaEventRecord(ev[2 + 3 * s], st));
		launch_sdm_update(st, m->dev, s, delta, d_shapes, status, (int)n_faces);
		if (ev) CUDA_TRY(cudaEventRecord(ev[3 + 3 * s], st));
		m->ctx->launches += 3;
		if (features_host)
			CUDA_TRY(cudaMemcpyAsync(features_host + (size_t)s * n_faces * m->dev.K, features, sizeof(float) * (size_t)n_faces * m->dev.K,
					cudaMemcpyDeviceToHost, st));
	}
	CUDA_TRY(cudaGetLastError());
	return FDB_OK;
}

int sdm_check_batch(const fdb_sdm* m, int W, int H, int n_frames, int64_t n_faces) {
	if (!m) return fail(FDB_ERR_INVALID_ARGUMENT, "null sdm");
	if (W < 1 || H < 1 || n_frames < 1 || n_faces < 0) return fail(FDB_ERR_INVALID_ARGUMENT, "bad image size or batch");
	if (n_faces > 65535) return fail(FDB_ERR_INVALID_ARGUMENT, "at most 65535 faces per call");
	return FDB_OK;
}

} // namespace

extern "C" {

int fdb_sdm_create(fdb_ctx* ctx, const fdb_sdm_desc* d, fdb_sdm** out) try {
	if (!out) return fail(FDB_ERR_INVALID_ARGUMENT, "out is null");
	*out = nullptr;
	int s = check_ctx(ctx); if (s) return s;
	if (!d || !d->mean_landmarks || !d->regressors) return fail(FDB_ERR_INVALID_ARGUMENT, "null descriptor field");
	if (d->num_landmarks < 13) return fail(FDB_ERR_INVALID_ARGUMENT, "SdmLandmarkModelFitting::optimize reads landmarks 8, 9, 11 and 12: at least 13 landmarks");
	if (d->num_landmarks > 1024) return fail(FDB_ERR_UNSUPPORTED, "more than 1024 landmarks");
	if (d->num_cascade_steps < 1 || d->num_cascade_steps > SDM_MAX_STEPS) return fail(FDB_ERR_UNSUPPORTED, "1..8 cascade steps");
	fdb_sdm* m = new fdb_sdm;
	m->ctx = ctx;
	sdm_fill_tables(&m->dev, d->num_landmarks, d->num_cascade_steps);
	m->mean.assign(d->mean_landmarks, d->mean_landmarks + 2 * (size_t)d->num_landmarks);
	for (int k = 0; k < d->num_cascade_steps; ++k) {
		float* p = nullptr;
		if (!d->regressors[k]) { fdb_sdm_destroy(m); return fail(FDB_ERR_INVALID_ARGUMENT, "null regressor"); }
		s = upload(d->regressors[k], (size_t)(m->dev.K + 1) * m->dev.N, &p, m->owned);
		if (s) { fdb_sdm_destroy(m); return s; }
		m->dev.R[k] = p;
	}
	for (cudaEvent_t& e : m->ev) if (cudaEventCreate(&e) != cudaSuccess) { fdb_sdm_destroy(m); return fail(FDB_ERR_CUDA, "cudaEventCreate"); }
	for (cudaEvent_t& e : m->chunk_ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { fdb_sdm_destroy(m); return fail(FDB_ERR_CUDA, "cudaEventCreate"); }
	for (cudaStream_t& t : m->chunk_stream) if (cudaStreamCreateWithFlags(&t, cudaStreamNonBlocking) != cudaSuccess) { fdb_sdm_destroy(m); return fail(FDB_ERR_CUDA, "cudaStreamCreate"); }
	*out = m;
	return FDB_OK;
} FDB_API_CATCH

void fdb_sdm_destroy(fdb_sdm* m) {
	if (!m) return;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	free_all(m->owned);
	for (void* p : {(void*)m->d_features, (void*)m->d_delta, (void*)m->d_shapes, (void*)m->d_status, (void*)m->d_face_frame, (void*)m->d_frames}) if (p) cudaFree(p);
	for (cudaEvent_t e : m->ev) if (e) cudaEventDestroy(e);
	for (cudaEvent_t e : m->chunk_ev) if (e) cudaEventDestroy(e);
	for (cudaStream_t t : m->chunk_stream) if (t) cudaStreamDestroy(t);
	delete m;
}

int32_t fdb_sdm_num_landmarks(const fdb_sdm* m) { return m ? m->dev.L : 0; }
int32_t fdb_sdm_num_cascade_steps(const fdb_sdm* m) { return m ? m->dev.steps : 0; }

int fdb_sdm_align_rigid(const fdb_sdm* m, const int32_t* boxes, int64_t n_faces, float* shapes_out) try {
	if (!m || !boxes || !shapes_out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	const int L = m->dev.L;
	/* SdmLandmarkModel.hpp:156-192 with modelShape = mean. The source line `(xCoords + 0.5f) * faceBox.width + faceBox.x` is a
	 * cv::MatExpr that OpenCV folds into convertTo(CV_32F, alpha = w, beta = 0.5 w + x): one float32 multiply, one float32 add
	 * (core/src/matop.cpp MatOp_AddEx, core/src/convert.cpp cvtScale_<float, float, float>). */
	for (int64_t f = 0; f < n_faces; ++f) {
		const float ax = (float)boxes[4 * f + 2], bx = (float)(0.5 * boxes[4 * f + 2] + boxes[4 * f]);
		const float ay = (float)boxes[4 * f + 3], by = (float)(0.5 * boxes[4 * f + 3] + boxes[4 * f + 1]);
		float* shape = shapes_out + f * 2 * L;
		for (int i = 0; i < L; ++i) {
			volatile float tx = m->mean[i] * ax, ty = m->mean[L + i] * ay; /* volatile: one rounding per operation, no contraction */
			shape[i] = tx + bx;
			shape[L + i] = ty + by;
		}
	}
	return FDB_OK;
} FDB_API_CATCH

int fdb_sdm_optimize_batch_device(fdb_sdm* m, const uint8_t* frames, int32_t W, int32_t H, int32_t n_frames, const int32_t* face_frame,
		int64_t n_faces, float* shapes, int32_t* status) try {
	int s = sdm_check_batch(m, W, H, n_frames, n_faces); if (s) return s;
	if (!frames || !shapes) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	if (!face_frame && n_faces > n_frames) return fail(FDB_ERR_INVALID_ARGUMENT, "face_frame is null but there are more faces than frames");
	s = check_ctx(m->ctx); if (s) return s;
	if (n_faces == 0) return FDB_OK;
	s = sdm_reserve(m, n_faces, 0); if (s) return s;
	if (!face_frame) { /* face i lies in frame i */
		std::vector<int> ident((size_t)n_faces);
		for (int64_t i = 0; i < n_faces; ++i) ident[i] = (int)i;
		CUDA_TRY(cudaMemcpyAsync(m->d_face_frame, ident.data(), sizeof(int) * (size_t)n_faces, cudaMemcpyHostToDevice, m->ctx->stream));
		CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
		face_frame = m->d_face_frame;
	}
	return sdm_run(m, frames, W, H, face_frame, n_faces, shapes, status, nullptr, nullptr);
} FDB_API_CATCH

int fdb_sdm_profile_device(fdb_sdm* m, const uint8_t* frames, int32_t W, int32_t H, int32_t n_frames, const int32_t* face_frame,
		int64_t n_faces, float* shapes, int32_t* status, double ms_out[4]) try {
	int s = sdm_check_batch(m, W, H, n_frames, n_faces); if (s) return s;
	if (!frames || !shapes || !face_frame || !ms_out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	s = check_ctx(m->ctx); if (s) return s;
	s = sdm_reserve(m, n_faces, 0); if (s) return s;
	s = sdm_run(m, frames, W, H, face_frame, n_faces, shapes, status, nullptr, m->ev); if (s) return s;
	CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
	ms_out[0] = ms_out[1] = ms_out[2] = ms_out[3] = 0;
	for (int k = 0; k < m->dev.steps; ++k)
		for (int j = 0; j < 3; ++j) {
			float ms = 0;
			CUDA_TRY(cudaEventElapsedTime(&ms, m->ev[3 * k + j], m->ev[3 * k + j + 1]));
			ms_out[j] += ms; ms_out[3] += ms;
		}
	return FDB_OK;
} FDB_API_CATCH

int fdb_sdm_optimize_batch(fdb_sdm* m, const uint8_t* frames, int64_t pitch, int32_t W, int32_t H, int32_t n_frames, const int32_t* face_frame,
		int64_t n_faces, float* shapes, int32_t* status_out, float* features_out) try {
	int s = sdm_check_batch(m, W, H, n_frames, n_faces); if (s) return s;
	if (!frames || !shapes || pitch < W) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument or pitch < width");
	if (!face_frame && n_faces > n_frames) return fail(FDB_ERR_INVALID_ARGUMENT, "face_frame is null but there are more faces than frames");
	for (int64_t i = 0; face_frame && i < n_faces; ++i)
		if (face_frame[i] < 0 || face_frame[i] >= n_frames) return fail(FDB_ERR_INVALID_ARGUMENT, "face_frame entry out of range");
	s = check_ctx(m->ctx); if (s) return s;
	if (n_faces == 0) return FDB_OK;
	s = sdm_reserve(m, n_faces, (int64_t)n_frames * W * H); if (s) return s;
	cudaStream_t st = m->ctx->stream;
	const int N = m->dev.N;
	const size_t frame_bytes = (size_t)W * H;
	auto upload_frames = [&](int f0, int f1, cudaStream_t t) -> cudaError_t {
		if (pitch == W) return cudaMemcpyAsync(m->d_frames + f0 * frame_bytes, frames + f0 * frame_bytes, (size_t)(f1 - f0) * frame_bytes, cudaMemcpyHostToDevice, t);
		return cudaMemcpy2DAsync(m->d_frames + f0 * frame_bytes, (size_t)W, frames + (size_t)f0 * pitch * H, (size_t)pitch, (size_t)W, (size_t)(f1 - f0) * H,
				cudaMemcpyHostToDevice, t);
	};
	std::vector<int> ident;
	if (!face_frame) { ident.resize((size_t)n_faces); for (int64_t i = 0; i < n_faces; ++i) ident[i] = (int)i; face_frame = ident.data(); }
	const bool pipelined = !features_out && n_frames >= SDM_CHUNKS && (int64_t)n_frames * (int64_t)frame_bytes >= SDM_CHUNK_MIN_BYTES && n_faces >= 64 * SDM_CHUNKS;
	if (!pipelined) {
		CUDA_TRY(upload_frames(0, n_frames, st));
		CUDA_TRY(cudaMemcpyAsync(m->d_face_frame, face_frame, sizeof(int) * (size_t)n_faces, cudaMemcpyHostToDevice, st));
		CUDA_TRY(cudaMemcpyAsync(m->d_shapes, shapes, sizeof(float) * (size_t)n_faces * N, cudaMemcpyHostToDevice, st));
		s = sdm_run(m, m->d_frames, W, H, m->d_face_frame, n_faces, m->d_shapes, nullptr, features_out, nullptr);
		if (s) { cudaStreamSynchronize(st); return s; }
		CUDA_TRY(cudaMemcpyAsync(shapes, m->d_shapes, sizeof(float) * (size_t)n_faces * N, cudaMemcpyDeviceToHost, st));
		if (status_out) CUDA_TRY(cudaMemcpyAsync(status_out, m->d_status, sizeof(int) * (size_t)n_faces, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		return FDB_OK;
	}
	/* Pipelined host call: the frames go up in SDM_CHUNKS pieces, each on its own stream followed by the whole cascade of the
	 * faces that lie in that piece, so the copy engine works on piece c + 1 while the SMs fit the faces of piece c (faces are
	 * independent: SURVEY.md 8(e)).  Faces are grouped by piece with a stable counting sort and scattered back at the end. */
	auto chunk_of = [&](int frame) { return (int)((int64_t)frame * SDM_CHUNKS / n_frames); };
	int64_t begin[SDM_CHUNKS + 1] = {0};
	for (int64_t i = 0; i < n_faces; ++i) ++begin[chunk_of(face_frame[i]) + 1];
	for (int c = 0; c < SDM_CHUNKS; ++c) begin[c + 1] += begin[c];
	std::vector<int64_t> order((size_t)n_faces);
	{
		int64_t fill[SDM_CHUNKS];
		for (int c = 0; c < SDM_CHUNKS; ++c) fill[c] = begin[c];
		for (int64_t i = 0; i < n_faces; ++i) order[(size_t)fill[chunk_of(face_frame[i])]++] = i;
	}
	std::vector<float> hshapes((size_t)n_faces * N);
	std::vector<int> hframe((size_t)n_faces), hstatus((size_t)n_faces);
	for (int64_t k = 0; k < n_faces; ++k) {
		std::memcpy(&hshapes[(size_t)k * N], shapes + order[(size_t)k] * N, sizeof(float) * N);
		hframe[(size_t)k] = face_frame[order[(size_t)k]];
	}
	CUDA_TRY(cudaMemcpyAsync(m->d_face_frame, hframe.data(), sizeof(int) * (size_t)n_faces, cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(m->d_shapes, hshapes.data(), sizeof(float) * (size_t)n_faces * N, cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaEventRecord(m->chunk_ev[SDM_CHUNKS], st));
	for (int c = 0; c < SDM_CHUNKS; ++c) {
		cudaStream_t t = m->chunk_stream[c];
		const int f0 = (int)(((int64_t)c * n_frames + SDM_CHUNKS - 1) / SDM_CHUNKS), f1 = (int)(((int64_t)(c + 1) * n_frames + SDM_CHUNKS - 1) / SDM_CHUNKS);
		CUDA_TRY(cudaStreamWaitEvent(t, m->chunk_ev[SDM_CHUNKS], 0));
		if (f1 > f0) CUDA_TRY(upload_frames(f0, f1, t));
		const int64_t nf = begin[c + 1] - begin[c];
		if (nf > 0) {
			s = sdm_run(m, m->d_frames, W, H, m->d_face_frame + begin[c], nf, m->d_shapes + begin[c] * N, nullptr, nullptr, nullptr, t, begin[c]);
			if (s) { cudaDeviceSynchronize(); return s; }
		}
		CUDA_TRY(cudaEventRecord(m->chunk_ev[c], t));
		CUDA_TRY(cudaStreamWaitEvent(st, m->chunk_ev[c], 0));
	}
	CUDA_TRY(cudaMemcpyAsync(hshapes.data(), m->d_shapes, sizeof(float) * (size_t)n_faces * N, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaMemcpyAsync(hstatus.data(), m->d_status, sizeof(int) * (size_t)n_faces, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	for (int64_t k = 0; k < n_faces; ++k) {
		std::memcpy(shapes + order[(size_t)k] * N, &hshapes[(size_t)k * N], sizeof(float) * N);
		if (status_out) status_out[order[(size_t)k]] = hstatus[(size_t)k];
	}
	return FDB_OK;
} FDB_API_CATCH

int fdb_sdm_descriptors(fdb_sdm* m, const uint8_t* frame, int64_t pitch, int32_t W, int32_t H, const float* pts, int32_t n_points,
		int32_t window_half, float* out) try {
	if (!m || !frame || !pts || !out || W < 1 || H < 1 || pitch < W || n_points < 0) return fail(FDB_ERR_INVALID_ARGUMENT, "bad argument");
	int s = check_ctx(m->ctx); if (s) return s;
	if (n_points == 0) return FDB_OK;
	/* one pseudo-face per ceil(n_points / L) group so that the kernel's (landmark, face) grid covers the points */
	const int L = m->dev.L;
	const int64_t groups = (n_points + L - 1) / L;
	s = sdm_reserve(m, groups, (int64_t)W * H); if (s) return s;
	cudaStream_t st = m->ctx->stream;
	CUDA_TRY(cudaMemcpy2DAsync(m->d_frames, (size_t)W, frame, (size_t)pitch, (size_t)W, (size_t)H, cudaMemcpyHostToDevice, st));
	std::vector<float> padded((size_t)groups * L * 2, 0.f);
	for (int i = 0; i < n_points; ++i) { padded[2 * i] = pts[2 * i]; padded[2 * i + 1] = pts[2 * i + 1]; }
	for (int64_t i = n_points; i < groups * L; ++i) { padded[2 * i] = pts[0]; padded[2 * i + 1] = pts[1]; } /* padding repeats a valid point */
	float* d_pts = m->d_delta; /* [groups][2 L] floats: same size as the delta buffer */
	CUDA_TRY(cudaMemcpyAsync(d_pts, padded.data(), sizeof(float) * padded.size(), cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaMemsetAsync(m->d_status, 0, sizeof(int) * (size_t)groups, st));
	launch_sdm_hog(st, m->dev, m->d_frames, W, H, nullptr, nullptr, 0, d_pts, window_half, (int)groups, m->d_features, m->d_status);
	m->ctx->launches += 1;
	std::vector<int> status((size_t)groups);
	CUDA_TRY(cudaMemcpyAsync(out, m->d_features, sizeof(float) * (size_t)n_points * SDM_DESC, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaMemcpyAsync(status.data(), m->d_status, sizeof(int) * (size_t)groups, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	CUDA_TRY(cudaGetLastError());
	for (int v : status) if (v) return fail(FDB_ERR_RUNTIME, "VlHogDescriptorExtractor::getDescriptors: region of interest outside the image");
	return FDB_OK;
} FDB_API_CATCH

} // extern "C"
