/*
 * sdm.cu - supervised-descent landmark regressor on the GPU (sm_100a), BASELINE configs[4]:
 *   SdmLandmarkModelFitting::alignRigid / optimize   libSupervisedDescent/include/superviseddescent/SdmLandmarkModel.hpp:156-256
 *   VlHogDescriptorExtractor::getDescriptors         libSupervisedDescent/include/superviseddescent/DescriptorExtractor.hpp:106-219
 *   vl_hog_put_image / vl_hog_extract (UoCTTI)       libSupervisedDescent/src/superviseddescent/hog.c:595-727,857-1063
 *
 * One cascade step over a batch of faces is three launches:
 *   sdm_hog_kernel     one CTA per (landmark, face): crop (black canvas outside the image, with the reference's
 *                      row-offset quirk), float32 bilinear resize to 30x30, VLFeat HOG (3x3 cells x 31) written
 *                      straight into the face's feature row [face][landmark * 279 + ...]
 *   sdm_gemm_kernel    delta[faces x 2L] = features[faces x 279 L] * R[0:-1] + R[-1]: float32 inputs, FLOAT64
 *                      accumulation in k order, exactly what cv::gemm does for CV_32F - the products of two floats
 *                      are exact in double, so the tiled kernel is bit-identical to the sequential reference sum
 *   sdm_update_kernel  shape += delta^T * eye-mouth distance
 * Why not the tensor cores: the fit is a feedback loop through cvRound(landmark) - a 1e-5 px difference in a shape
 * moves a HOG window by a whole pixel with probability ~1e-5 per coordinate, and 2L x steps coordinates per face
 * turn that into visibly different fits for ~1 % of the faces.  bf16/tf32 products cannot stay below that; the
 * float64-accumulated product can, and at [4096 x 18972] x [18972 x 136] it costs ~1 ms per step against ~3 ms of HOG.
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

/* mode 0: points come from the face's current shape, window from sdm_window; mode 1: explicit points + window */
__global__ void __launch_bounds__(SDM_THREADS) sdm_hog_kernel(const DevSdm m, const uint8_t* __restrict__ frames, int W, int H,
		const int* __restrict__ face_frame, const float* __restrict__ shapes, int step, const float* __restrict__ pts_xy,
		int window_half, float* __restrict__ features, int* __restrict__ status) {
	__shared__ float s_img[SDM_P * SDM_P];
	__shared__ float s_grad[SDM_P * SDM_P];
	__shared__ uint8_t s_bin[SDM_P * SDM_P];
	__shared__ uint32_t s_mask[SDM_P][2 * SDM_NO];         /* per row: pixels (bit x) whose orientation bin is o */
	__shared__ float s_hog[SDM_CELLS * SDM_CELLS * 2 * SDM_NO]; /* [o][cy][cx] */
	__shared__ float s_norm[SDM_CELLS * SDM_CELLS];
	__shared__ double s_fac[SDM_CELLS * SDM_CELLS][4];
	__shared__ double s_hc[SDM_CELLS * SDM_CELLS][SDM_NO][4];
	__shared__ int s_geo[8];
	const int tid = threadIdx.x, lm = blockIdx.x, face = blockIdx.y, L = m.L;
	const uint8_t* __restrict__ img = frames + (int64_t)(face_frame ? face_frame[face] : 0) * W * H;
	float* __restrict__ out = features + ((int64_t)face * L + lm) * SDM_DESC;

	if (tid == 0) {
		float px, py; int wsh;
		if (pts_xy) { px = pts_xy[2 * ((int64_t)face * L + lm)]; py = pts_xy[2 * ((int64_t)face * L + lm) + 1]; wsh = window_half; }
		else {
			const float* shape = shapes + (int64_t)face * 2 * L;
			float d;
			sdm_window(shape, L, m.step_factor[step], &d, &wsh);
			px = shape[lm]; py = shape[lm + L];
		}
		const int x = __float2int_rn(px), y = __float2int_rn(py); /* cvRound */
		int rx = x - wsh, ry = y - wsh, bl = 0, bt = 0, br = 0, bb = 0;
		if (x - wsh < 0 || y - wsh < 0 || x + wsh >= W || y + wsh >= H) { /* DescriptorExtractor.hpp:161-169 */
			bl = (x - wsh) < 0 ? abs(x - wsh) : 0;
			bt = (y - wsh) < 0 ? abs(y - wsh) : 0;
			br = (x + wsh) >= W ? abs(W - (x + wsh)) : 0;
			bb = (y + wsh) >= H ? abs(H - (y + wsh)) : 0;
			rx = (x - wsh) + bl;
			ry = (y - wsh) + br; /* sic (:169) */
		}
		const int side = 2 * wsh;
		const bool ok = side >= 4 && rx >= 0 && ry >= 0 && rx + side <= W + bl + br && ry + side <= H + bt + bb;
		s_geo[0] = rx - bl; s_geo[1] = ry - bt; s_geo[2] = side; s_geo[3] = ok ? 1 : 0;
	}
	__syncthreads();
	const int x0 = s_geo[0], y0 = s_geo[1], side = s_geo[2];
	if (!s_geo[3]) { /* the reference's Mat::operator()(roi) would throw: flag the face with step + 1, emit zeros */
		if (tid == 0 && status) atomicCAS(status + face, 0, step + 1);
		for (int i = tid; i < SDM_DESC; i += SDM_THREADS) out[i] = 0.f;
		return;
	}

	/* ---- crop + convertTo(CV_32F) + cv::resize(30x30, INTER_LINEAR) ---- */
	if (side == SDM_P) {
		for (int i = tid; i < SDM_P * SDM_P; i += SDM_THREADS) { const int r = i / SDM_P, c = i - r * SDM_P; s_img[i] = sdm_px(img, W, H, x0 + c, y0 + r); }
	} else if (side == 2 * SDM_P) { /* exact 2x decimation: INTER_AREA fast path, ((a + b) + c) + d) * 0.25 */
		for (int i = tid; i < SDM_P * SDM_P; i += SDM_THREADS) {
			const int r = i / SDM_P, c = i - r * SDM_P;
			const float a = sdm_px(img, W, H, x0 + 2 * c, y0 + 2 * r), b = sdm_px(img, W, H, x0 + 2 * c + 1, y0 + 2 * r);
			const float cc = sdm_px(img, W, H, x0 + 2 * c, y0 + 2 * r + 1), d = sdm_px(img, W, H, x0 + 2 * c + 1, y0 + 2 * r + 1);
			s_img[i] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), cc), d), 0.25f);
		}
	} else {
		const double scale = __ddiv_rn((double)side, (double)SDM_P);
		for (int i = tid; i < SDM_P * SDM_P; i += SDM_THREADS) {
			const int dy = i / SDM_P, dx = i - dy * SDM_P;
			float fx = (float)__dsub_rn(__dmul_rn((double)dx + 0.5, scale), 0.5);
			int sx = (int)floorf(fx);
			fx = __fsub_rn(fx, (float)sx);
			if (sx < 0) { fx = 0.f; sx = 0; }
			if (sx >= side - 1) { fx = 0.f; sx = side - 1; }
			float fy = (float)__dsub_rn(__dmul_rn((double)dy + 0.5, scale), 0.5);
			const int sy = (int)floorf(fy);
			fy = __fsub_rn(fy, (float)sy);
			const float a0 = __fsub_rn(1.f, fx), a1 = fx, b0 = __fsub_rn(1.f, fy), b1 = fy;
			const int sx1 = min(sx + 1, side - 1);
			const int sy0 = min(max(sy, 0), side - 1), sy1 = min(max(sy + 1, 0), side - 1);
			const float r0 = __fadd_rn(__fmul_rn(sdm_px(img, W, H, x0 + sx, y0 + sy0), a0), __fmul_rn(sdm_px(img, W, H, x0 + sx1, y0 + sy0), a1));
			const float r1 = __fadd_rn(__fmul_rn(sdm_px(img, W, H, x0 + sx, y0 + sy1), a0), __fmul_rn(sdm_px(img, W, H, x0 + sx1, y0 + sy1), a1));
			s_img[i] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
		}
	}
	for (int i = tid; i < SDM_CELLS * SDM_CELLS * 2 * SDM_NO; i += SDM_THREADS) s_hog[i] = 0.f;
	__syncthreads();

	/* ---- vl_hog_put_image: gradient, dominant directed orientation (hog.c:617-682) ---- */
	for (int i = tid; i < SDM_P * SDM_P; i += SDM_THREADS) {
		const int y = i / SDM_P, x = i - y * SDM_P;
		int bin = 255;
		float grad = 0.f;
		if (x >= 1 && x < SDM_P - 1 && y >= 1 && y < SDM_P - 1) {
			float gx = __fsub_rn(s_img[i + 1], s_img[i - 1]), gy = __fsub_rn(s_img[i + SDM_P], s_img[i - SDM_P]);
			float g2 = __fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy));
			if (!(g2 > 0.f)) { gx = 0.f; gy = 0.f; g2 = 0.f; }
			grad = __fsqrt_rn(g2);
			const double den = (double)grad > 1e-10 ? (double)grad : 1e-10;
			gx = (float)__ddiv_rn((double)gx, den);
			gy = (float)__ddiv_rn((double)gy, den);
			float w0 = 0.f, w1 = 0.f;
			int b0 = -1;
#pragma unroll
			for (int k = 0; k < SDM_NO; ++k) {
				float score = __fadd_rn(__fmul_rn(gx, m.ox[k]), __fmul_rn(gy, m.oy[k]));
				int b = k;
				if (score < 0.f) { score = -score; b += SDM_NO; }
				if (score > w0) { w1 = w0; b0 = b; w0 = score; }
				else if (score > w1) { w1 = score; }
			}
			if (b0 >= 0) bin = b0;
		}
		s_bin[i] = (uint8_t)bin;
		s_grad[i] = grad;
	}
	__syncthreads();
	{ /* per row and orientation: which columns hold that bin */
		const int lane = tid & 31, warp = tid >> 5;
		for (int y = warp; y < SDM_P; y += SDM_THREADS / 32) {
			const int b = lane < SDM_P ? s_bin[y * SDM_P + lane] : 255;
#pragma unroll
			for (int o = 0; o < 2 * SDM_NO; ++o) {
				const uint32_t msk = __ballot_sync(0xffffffffu, b == o);
				if (lane == 0) s_mask[y][o] = msk;
			}
		}
	}
	__syncthreads();
	/* bilinear cell accumulation (hog.c:697-722): entry (o, cy, cx) visits its pixels in raster order */
	for (int e = tid; e < SDM_CELLS * SDM_CELLS * 2 * SDM_NO; e += SDM_THREADS) {
		const int o = e / (SDM_CELLS * SDM_CELLS), c = e - o * (SDM_CELLS * SDM_CELLS), cy = c / SDM_CELLS, cx = c - cy * SDM_CELLS;
		float acc = 0.f;
		for (int y = 1; y < SDM_P - 1; ++y) {
			const int by = m.bin_of[y];
			float wy;
			if (by == cy) wy = m.w1_of[y];
			else if (by + 1 == cy) wy = m.w2_of[y];
			else continue;
			uint32_t msk = s_mask[y][o];
			while (msk) {
				const int x = __ffs(msk) - 1;
				msk &= msk - 1;
				const int bx = m.bin_of[x];
				float wx;
				if (bx == cx) wx = m.w1_of[x];
				else if (bx + 1 == cx) wx = m.w2_of[x];
				else continue;
				acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(s_grad[y * SDM_P + x], wx), wy));
			}
		}
		s_hog[e] = acc;
	}
	__syncthreads();
	/* ---- vl_hog_extract (hog.c:879-1060) ---- */
	if (tid < SDM_CELLS * SDM_CELLS) {
		float nrm = 0.f;
		for (int k = 0; k < SDM_NO; ++k) {
			const float h = __fadd_rn(s_hog[k * 9 + tid], s_hog[(k + SDM_NO) * 9 + tid]);
			nrm = __fadd_rn(nrm, __fmul_rn(h, h));
		}
		s_norm[tid] = nrm;
	}
	__syncthreads();
	if (tid < SDM_CELLS * SDM_CELLS) {
		const int y = tid / SDM_CELLS, x = tid - y * SDM_CELLS;
		const int xm = max(x - 1, 0), xp = min(x + 1, SDM_CELLS - 1), ym = max(y - 1, 0), yp = min(y + 1, SDM_CELLS - 1);
#define NRM(xx, yy) ((double)s_norm[(yy) * SDM_CELLS + (xx)])
		const double n1 = NRM(xm, ym), n2 = NRM(x, ym), n3 = NRM(xp, ym), n4 = NRM(xm, y), n5 = NRM(x, y), n6 = NRM(xp, y);
		const double n7 = NRM(xm, yp), n8 = NRM(x, yp), n9 = NRM(xp, yp);
#undef NRM
		s_fac[tid][0] = __ddiv_rn(1.0, sqrt(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(n1, n2), n4), n5), 1e-4)));
		s_fac[tid][1] = __ddiv_rn(1.0, sqrt(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(n2, n3), n5), n6), 1e-4)));
		s_fac[tid][2] = __ddiv_rn(1.0, sqrt(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(n4, n5), n7), n8), 1e-4)));
		s_fac[tid][3] = __ddiv_rn(1.0, sqrt(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(n5, n6), n8), n9), 1e-4)));
	}
	__syncthreads();
	/* descriptor layout (DescriptorExtractor.hpp:196-204): dimension j, then cell column x, then cell row y */
	if (tid < SDM_CELLS * SDM_CELLS * SDM_NO) {
		const int c = tid / SDM_NO, k = tid - c * SDM_NO, cy = c / SDM_CELLS, cx = c - cy * SDM_CELLS;
		const double ha = s_hog[k * 9 + c], hb = s_hog[(k + SDM_NO) * 9 + c];
		double sa = 0, sb = 0, sc = 0;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const double f = s_fac[c][q];
			const double haq = __dmul_rn(f, ha), hbq = __dmul_rn(f, hb);
			const double hcq = __dadd_rn(haq, hbq);
			const double hcm = fmin(0.2, hcq);
			s_hc[c][k][q] = hcm;
			sa = q == 0 ? fmin(0.2, haq) : __dadd_rn(sa, fmin(0.2, haq));
			sb = q == 0 ? fmin(0.2, hbq) : __dadd_rn(sb, fmin(0.2, hbq));
			sc = q == 0 ? hcm : __dadd_rn(sc, hcm);
		}
		const int pos = cx * SDM_CELLS + cy;
		out[k * 9 + pos] = (float)__dmul_rn(0.5, sa);
		out[(k + SDM_NO) * 9 + pos] = (float)__dmul_rn(0.5, sb);
		out[(k + 2 * SDM_NO) * 9 + pos] = (float)__dmul_rn(0.5, sc);
	}
	__syncthreads();
	if (tid < SDM_CELLS * SDM_CELLS * 4) { /* texture features: t_q = sum over k of the clamped hc (hog.c:1012-1015,1046-1049) */
		const int c = tid >> 2, q = tid & 3, cy = c / SDM_CELLS, cx = c - cy * SDM_CELLS;
		double t = 0;
		for (int k = 0; k < SDM_NO; ++k) t = __dadd_rn(t, s_hc[c][k][q]);
		out[(3 * SDM_NO + q) * 9 + cx * SDM_CELLS + cy] = (float)__dmul_rn((double)m.tex, t);
	}
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
	dim3 grid((unsigned)m.L, (unsigned)n_faces);
	sdm_hog_kernel<<<grid, SDM_THREADS, 0, st>>>(m, frames, W, H, face_frame, shapes, step, pts_xy, window_half, features, status);
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
		float* features_host, cudaEvent_t* ev) {
	cudaStream_t st = m->ctx->stream;
	int* status = d_status ? d_status : m->d_status;
	CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int) * (size_t)n_faces, st));
	if (ev) CUDA_TRY(cudaEventRecord(ev[0], st));
	for (int s = 0; s < m->dev.steps; ++s) {
		launch_sdm_hog(st, m->dev, d_frames, W, H, d_face_frame, d_shapes, s, nullptr, 0, (int)n_faces, m->d_features, status);
		if (ev) CUDA_TRY(cudaEventRecord(ev[1 + 3 * s], st));
		launch_sdm_gemm(st, m->dev, s, m->d_features, (int)n_faces, m->d_delta);
		if (ev) CUDA_TRY(cudaEventRecord(ev[2 + 3 * s], st));
		launch_sdm_update(st, m->dev, s, m->d_delta, d_shapes, status, (int)n_faces);
		if (ev) CUDA_TRY(cudaEventRecord(ev[3 + 3 * s], st));
		m->ctx->launches += 3;
		if (features_host)
			CUDA_TRY(cudaMemcpyAsync(features_host + (size_t)s * n_faces * m->dev.K, m->d_features, sizeof(float) * (size_t)n_faces * m->dev.K,
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

int fdb_sdm_create(fdb_ctx* ctx, const fdb_sdm_desc* d, fdb_sdm** out) {
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
	*out = m;
	return FDB_OK;
}

void fdb_sdm_destroy(fdb_sdm* m) {
	if (!m) return;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	free_all(m->owned);
	for (void* p : {(void*)m->d_features, (void*)m->d_delta, (void*)m->d_shapes, (void*)m->d_status, (void*)m->d_face_frame, (void*)m->d_frames}) if (p) cudaFree(p);
	for (cudaEvent_t e : m->ev) if (e) cudaEventDestroy(e);
	delete m;
}

int32_t fdb_sdm_num_landmarks(const fdb_sdm* m) { return m ? m->dev.L : 0; }
int32_t fdb_sdm_num_cascade_steps(const fdb_sdm* m) { return m ? m->dev.steps : 0; }

int fdb_sdm_align_rigid(const fdb_sdm* m, const int32_t* boxes, int64_t n_faces, float* shapes_out) {
	if (!m || !boxes || !shapes_out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	const int L = m->dev.L;
	for (int64_t f = 0; f < n_faces; ++f) { /* SdmLandmarkModel.hpp:156-192 with modelShape = mean: (mean + 0.5) * size + corner, float32 */
		const float x = (float)boxes[4 * f], y = (float)boxes[4 * f + 1], w = (float)boxes[4 * f + 2], h = (float)boxes[4 * f + 3];
		float* shape = shapes_out + f * 2 * L;
		for (int i = 0; i < L; ++i) {
			volatile float tx = m->mean[i] + 0.5f, ty = m->mean[L + i] + 0.5f; /* volatile: one rounding per operation */
			volatile float sx = tx * w, sy = ty * h;
			shape[i] = sx + x;
			shape[L + i] = sy + y;
		}
	}
	return FDB_OK;
}

int fdb_sdm_optimize_batch_device(fdb_sdm* m, const uint8_t* frames, int32_t W, int32_t H, int32_t n_frames, const int32_t* face_frame,
		int64_t n_faces, float* shapes, int32_t* status) {
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
}

int fdb_sdm_profile_device(fdb_sdm* m, const uint8_t* frames, int32_t W, int32_t H, int32_t n_frames, const int32_t* face_frame,
		int64_t n_faces, float* shapes, int32_t* status, double ms_out[4]) {
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
}

int fdb_sdm_optimize_batch(fdb_sdm* m, const uint8_t* frames, int64_t pitch, int32_t W, int32_t H, int32_t n_frames, const int32_t* face_frame,
		int64_t n_faces, float* shapes, int32_t* status_out, float* features_out) {
	int s = sdm_check_batch(m, W, H, n_frames, n_faces); if (s) return s;
	if (!frames || !shapes || pitch < W) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument or pitch < width");
	if (!face_frame && n_faces > n_frames) return fail(FDB_ERR_INVALID_ARGUMENT, "face_frame is null but there are more faces than frames");
	for (int64_t i = 0; face_frame && i < n_faces; ++i)
		if (face_frame[i] < 0 || face_frame[i] >= n_frames) return fail(FDB_ERR_INVALID_ARGUMENT, "face_frame entry out of range");
	s = check_ctx(m->ctx); if (s) return s;
	if (n_faces == 0) return FDB_OK;
	s = sdm_reserve(m, n_faces, (int64_t)n_frames * W * H); if (s) return s;
	cudaStream_t st = m->ctx->stream;
	if (pitch == W) CUDA_TRY(cudaMemcpyAsync(m->d_frames, frames, (size_t)n_frames * W * H, cudaMemcpyHostToDevice, st));
	else CUDA_TRY(cudaMemcpy2DAsync(m->d_frames, (size_t)W, frames, (size_t)pitch, (size_t)W, (size_t)n_frames * H, cudaMemcpyHostToDevice, st));
	std::vector<int> ident;
	if (!face_frame) { ident.resize((size_t)n_faces); for (int64_t i = 0; i < n_faces; ++i) ident[i] = (int)i; face_frame = ident.data(); }
	CUDA_TRY(cudaMemcpyAsync(m->d_face_frame, face_frame, sizeof(int) * (size_t)n_faces, cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(m->d_shapes, shapes, sizeof(float) * (size_t)n_faces * m->dev.N, cudaMemcpyHostToDevice, st));
	s = sdm_run(m, m->d_frames, W, H, m->d_face_frame, n_faces, m->d_shapes, nullptr, features_out, nullptr);
	if (s) { cudaStreamSynchronize(st); return s; }
	CUDA_TRY(cudaMemcpyAsync(shapes, m->d_shapes, sizeof(float) * (size_t)n_faces * m->dev.N, cudaMemcpyDeviceToHost, st));
	if (status_out) CUDA_TRY(cudaMemcpyAsync(status_out, m->d_status, sizeof(int) * (size_t)n_faces, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	return FDB_OK;
}

int fdb_sdm_descriptors(fdb_sdm* m, const uint8_t* frame, int64_t pitch, int32_t W, int32_t H, const float* pts, int32_t n_points,
		int32_t window_half, float* out) {
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
}

} // extern "C"
