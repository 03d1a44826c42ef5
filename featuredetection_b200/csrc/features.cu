/*
 * features.cu - feature spaces of the second-stage / `single` classifier on the GPU (sm_100a):
 * pyramid layer filters and per-window patch-filter chains of the reference.
 *
 *   layer filters (ImagePyramid.cpp:182,191 apply them to every layer):
 *     GradientFilter           GradientFilter.cpp:38-59    cv::Sobel(k = 1 | 3, scale, delta 127) -> {gx, gy} u8
 *     GradientBinningFilter    GradientBinningFilter.cpp:18-92   65536-entry table -> {bin, w} or {bin0, w0, bin1, w1}
 *     LbpFilter                LbpFilter.cpp:20-85, LbpFilter.hpp:88-178   3x3 codes, BORDER_REPLICATE, uniform map
 *   patch filters (FilteringPyramidFeatureExtractor.hpp:46-66, DirectPyramidFeatureExtractor.cpp:117):
 *     HistogramEqualizationFilter  HistogramEqualizationFilter.cpp:17-20   cv::equalizeHist
 *     WhiteningFilter chain        WhiteningFilter.cpp:20-81, ConversionFilter.cpp:16-19, UnitNormFilter.cpp:20-38
 *     SpatialHistogramFilter       SpatialHistogramFilter.cpp:56-94 (+ HistogramFilter.cpp:23-252)
 *     HogFilter                    HogFilter.cpp:58-122
 *     ExtendedHogFilter            ExtendedHogFilter.cpp:54-209
 *
 * Exactness: every float32 accumulation keeps the reference's order - one thread owns one accumulator
 * (a (cell, bin) histogram entry, a cell energy, a block normaliser) and visits its contributions in the
 * reference's pixel / corner / bin order with _rn intrinsics (no FMA contraction).  Whole-vector norms are
 * double-precision sequential sums by one thread, as cv::norm accumulates.  The whitening transforms are
 * evaluated in double (direct sums, exact-argument twiddle tables): cv::dft is a float32 FFT whose rounding
 * no other implementation reproduces bit for bit; the u8 quantisation after it absorbs the difference except
 * when a value lies within ~1e-5 of a rounding boundary (measured: tests/test_gpu_features.py).
 *
 * One CTA per window; candidates of the cascade are few (tens per frame), so this kernel is latency-bound
 * by design; the all-windows `single` detector reuses it with one CTA per window of the frame.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "fdb_internal.h"
#include "wvm_device.h"
#include "api_types.h"
#include "features_device.h"

namespace fdb {

#define FEAT_THREADS 128
#define HIST_EPS 1e-4f /* HistogramFilter::eps, UnitNormFilter::eps */

/* ------------------------------------------------------------------------------------------------
 * layer filters: one thread per layer pixel, all kept layers of all frames in one launch
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ int refl101(int p, int len) {
	if (len == 1) return 0;
	while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
	return p;
}

__device__ __forceinline__ uint32_t sat_u8_rint(float v) {
	const int i = __float2int_rn(v); /* cvRound: half to even */
	return (uint32_t)min(max(i, 0), 255);
}

__global__ void __launch_bounds__(256) feature_layer_kernel(const DevFeature f, const uint8_t* __restrict__ frames, int W, int H,
		const uint8_t* __restrict__ arena, int64_t arena_stride, const DevLayer* __restrict__ layers,
		uint8_t* __restrict__ farena, int64_t farena_stride) {
	const int frame = blockIdx.y;
	const int gid = blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= f.px_prefix[f.n_layers]) return;
	int li = 0;
	while (gid >= f.px_prefix[li + 1]) ++li;
	const DevLayer L = layers[li];
	const int p = gid - f.px_prefix[li];
	const int y = p / L.width, x = p - y * L.width;
	const uint8_t* __restrict__ img = (L.offset < 0 ? frames + (int64_t)frame * W * H : arena + (int64_t)frame * arena_stride + L.offset);
	uint8_t* __restrict__ out = farena + (int64_t)frame * farena_stride + f.layer_offset[li];
	if (f.kind == FDB_FEATURE_LBP) {
		const int ym = max(y - 1, 0), yp = min(y + 1, L.height - 1), xm = max(x - 1, 0), xp = min(x + 1, L.width - 1);
		const uint8_t* r0 = img + (int64_t)ym * L.pitch;
		const uint8_t* r1 = img + (int64_t)y * L.pitch;
		const uint8_t* r2 = img + (int64_t)yp * L.pitch;
		const int c = r1[x];
		int code;
		if (f.lbp_type == FDB_LBP8 || f.lbp_type == FDB_LBP8_UNIFORM)
			code = ((r0[xm] > c) << 7) | ((r0[x] > c) << 6) | ((r0[xp] > c) << 5) | ((r1[xp] > c) << 4)
					| ((r2[xp] > c) << 3) | ((r2[x] > c) << 2) | ((r2[xm] > c) << 1) | (r1[xm] > c);
		else if (f.lbp_type == FDB_LBP4)
			code = ((r0[x] > c) << 3) | ((r1[xp] > c) << 2) | ((r2[x] > c) << 1) | (r1[xm] > c);
		else
			code = ((r0[xm] > c) << 3) | ((r0[xp] > c) << 2) | ((r2[xp] > c) << 1) | (r2[xm] > c);
		out[p] = f.lbp_map[code];
		return;
	}
	/* Sobel with BORDER_REFLECT_101; all intermediates are exact in float32 */
	const int ym = refl101(y - 1, L.height), yp = refl101(y + 1, L.height);
	const int xm = refl101(x - 1, L.width), xp = refl101(x + 1, L.width);
	const uint8_t* r0 = img + (int64_t)ym * L.pitch;
	const uint8_t* r1 = img + (int64_t)y * L.pitch;
	const uint8_t* r2 = img + (int64_t)yp * L.pitch;
	float gx, gy;
	if (f.gradient_kernel == 1) {
		gx = __fadd_rn(__fmul_rn((float)((int)r1[xp] - (int)r1[xm]), 0.5f), 127.f);
		gy = __fadd_rn(__fmul_rn((float)((int)r2[x] - (int)r0[x]), 0.5f), 127.f);
	} else {
		const int dx = ((int)r0[xp] - (int)r0[xm]) + 2 * ((int)r1[xp] - (int)r1[xm]) + ((int)r2[xp] - (int)r2[xm]);
		const int dy = ((int)r2[xm] - (int)r0[xm]) + 2 * ((int)r2[x] - (int)r0[x]) + ((int)r2[xp] - (int)r0[xp]);
		gx = __fadd_rn(__fmul_rn((float)dx, 0.125f), 127.f);
		gy = __fadd_rn(__fmul_rn((float)dy, 0.125f), 127.f);
	}
	const uint32_t idx = sat_u8_rint(gx) | (sat_u8_rint(gy) << 8);
	if (f.layer_channels == 4) reinterpret_cast<uint32_t*>(out)[p] = reinterpret_cast<const uint32_t*>(f.lut)[idx];
	else reinterpret_cast<uint16_t*>(out)[p] = reinterpret_cast<const uint16_t*>(f.lut)[idx];
}

/* ------------------------------------------------------------------------------------------------
 * patch filters: one CTA per window
 * ---------------------------------------------------------------------------------------------- */
/* cv::equalizeHist on the w*h bytes at src (u8, continuous) -> dst; s_hist: 256 ints of scratch */
__device__ void equalize_hist_block(const uint8_t* src, uint8_t* dst, int n, int* s_hist, uint8_t* s_lut) {
	const int tid = threadIdx.x;
	for (int i = tid; i < 256; i += FEAT_THREADS) s_hist[i] = 0;
	__syncthreads();
	for (int i = tid; i < n; i += FEAT_THREADS) atomicAdd(&s_hist[src[i]], 1);
	__syncthreads();
	if (tid == 0) {
		int i = 0;
		while (!s_hist[i]) ++i;
		if (s_hist[i] == n) {
			for (int k = 0; k < 256; ++k) s_lut[k] = (uint8_t)i;
		} else {
			const float scale = __fdiv_rn(255.f, (float)(n - s_hist[i]));
			int sum = 0;
			for (int k = 0; k <= i; ++k) s_lut[k] = 0;
			for (++i; i < 256; ++i) {
				sum += s_hist[i];
				s_lut[i] = (uint8_t)sat_u8_rint(__fmul_rn((float)sum, scale));
			}
		}
	}
	__syncthreads();
	for (int i = tid; i < n; i += FEAT_THREADS) dst[i] = s_lut[src[i]];
	__syncthreads();
}

/* v[0..n) /= s the way cv::Mat / double does: multiply by (float)(1 / s) */
__device__ void scale_block(float* v, int n, double s) {
	const float sc = (float)(1.0 / s);
	for (int i = threadIdx.x; i < n; i += FEAT_THREADS) v[i] = __fmul_rn(v[i], sc);
	__syncthreads();
}

/* cv::norm: double accumulation in element order; every thread returns the value */
__device__ double norm_block(const float* v, int n, bool l1, double* s_tmp) {
	__syncthreads();
	if (threadIdx.x == 0) {
		double s = 0;
		if (l1) for (int i = 0; i < n; ++i) s += fabs((double)v[i]);
		else for (int i = 0; i < n; ++i) s = __dadd_rn(s, __dmul_rn((double)v[i], (double)v[i]));
		*s_tmp = l1 ? s : sqrt(s);
	}
	__syncthreads();
	return *s_tmp;
}

/* HistogramFilter::normalize (HistogramFilter.cpp:222-252) on v[0..n) by the whole CTA */
__device__ void normalize_block(float* v, int n, int normalization, double* s_tmp) {
	if (normalization == FDB_NORM_NONE) return;
	const bool l1 = normalization == FDB_NORM_L1NORM || normalization == FDB_NORM_L1SQRT;
	float nf = (float)norm_block(v, n, l1, s_tmp);
	scale_block(v, n, (double)__fadd_rn(nf, HIST_EPS));
	if (normalization == FDB_NORM_L2HYS) {
		for (int i = threadIdx.x; i < n; i += FEAT_THREADS) v[i] = fminf(v[i], 0.2f);
		nf = (float)norm_block(v, n, false, s_tmp);
		scale_block(v, n, (double)__fadd_rn(nf, HIST_EPS));
	} else if (normalization == FDB_NORM_L1SQRT) {
		for (int i = threadIdx.x; i < n; i += FEAT_THREADS) v[i] = __fsqrt_rn(v[i]);
		__syncthreads();
	}
}

/* same for one thread (per-block normalisation of SpatialHistogramFilter::createBlockHistograms) */
__device__ void normalize_thread(float* v, int n, int normalization) {
	if (normalization == FDB_NORM_NONE) return;
	const bool l1 = normalization == FDB_NORM_L1NORM || normalization == FDB_NORM_L1SQRT;
	for (int pass = 0; pass < 2; ++pass) {
		double s = 0;
		if (l1) for (int i = 0; i < n; ++i) s += fabs((double)v[i]);
		else for (int i = 0; i < n; ++i) s = __dadd_rn(s, __dmul_rn((double)v[i], (double)v[i]));
		const float nf = (float)(l1 ? s : sqrt(s));
		const float sc = (float)(1.0 / (double)__fadd_rn(nf, HIST_EPS));
		for (int i = 0; i < n; ++i) v[i] = __fmul_rn(v[i], sc);
		if (normalization != FDB_NORM_L2HYS || pass == 1) break;
		for (int i = 0; i < n; ++i) v[i] = fminf(v[i], 0.2f);
	}
	if (normalization == FDB_NORM_L1SQRT) for (int i = 0; i < n; ++i) v[i] = __fsqrt_rn(v[i]);
}

__global__ void __launch_bounds__(FEAT_THREADS) feature_patch_kernel(const DevFeature f, const uint8_t* __restrict__ frames, int W, int H,
		const uint8_t* __restrict__ arena, int64_t arena_stride, const DevLayer* __restrict__ layers,
		const uint8_t* __restrict__ farena, int64_t farena_stride, const SvmItem* __restrict__ items, void* __restrict__ out) {
	extern __shared__ __align__(16) unsigned char fsm[];
	__shared__ int s_hist[256];
	__shared__ uint8_t s_lut[256];
	__shared__ double s_tmp;
	const int tid = threadIdx.x;
	const SvmItem it = items[blockIdx.x];
	const DevLayer L = layers[it.layer];
	const int pw = f.pw, ph = f.ph, npx = pw * ph;

	if (f.layer_channels == 0) {
		/* ---- chains on the gray window: gray / histeq / whi ---- */
		const uint8_t* __restrict__ img = (L.offset < 0 ? frames + (int64_t)it.frame * W * H
				: arena + (int64_t)it.frame * arena_stride + L.offset) + (int64_t)it.y * L.pitch + it.x;
		uint8_t* s_px = fsm;                 /* [npx] window pixels */
		uint8_t* s_px2 = fsm + ((npx + 15) & ~15);
		for (int i = tid; i < npx; i += FEAT_THREADS) { const int r = i / pw; s_px[i] = img[(int64_t)r * L.pitch + (i - r * pw)]; }
		__syncthreads();
		if (f.kind == FDB_FEATURE_GRAY) {
			uint8_t* o = reinterpret_cast<uint8_t*>(out) + (int64_t)blockIdx.x * f.dim;
			for (int i = tid; i < npx; i += FEAT_THREADS) o[i] = s_px[i];
			return;
		}
		if (f.kind == FDB_FEATURE_HISTEQ) {
			equalize_hist_block(s_px, s_px2, npx, s_hist, s_lut);
			uint8_t* o = reinterpret_cast<uint8_t*>(out) + (int64_t)blockIdx.x * f.dim;
			for (int i = tid; i < npx; i += FEAT_THREADS) o[i] = s_px2[i];
			return;
		}
		/* whitening (WhiteningFilter.cpp:20-56): forward DFT / (w h), x filter, Hermitian inverse over columns 0..w/2 */
		const int hw = pw / 2 + 1, nh = ph * hw;
		double* a_re = reinterpret_cast<double*>(fsm + 2 * ((npx + 15) & ~15));
		double* a_im = a_re + nh; double* b_re = a_im + nh; double* b_im = b_re + nh;
		const double* __restrict__ cw = f.twiddle; const double* __restrict__ sw = cw + pw;
		const double* __restrict__ ch = sw + pw; const double* __restrict__ sh = ch + ph;
		for (int i = tid; i < nh; i += FEAT_THREADS) { /* forward rows */
			const int y = i / hw, u = i - y * hw;
			double re = 0, im = 0;
			for (int x = 0, k = 0; x < pw; ++x) {
				const double v = (double)s_px[y * pw + x];
				re += v * cw[k]; im -= v * sw[k];
				k += u; if (k >= pw) k -= pw;
			}
			a_re[i] = re; a_im[i] = im;
		}
		__syncthreads();
		const double inv_n = 1.0 / ((double)pw * ph);
		for (int i = tid; i < nh; i += FEAT_THREADS) { /* forward columns, DFT_SCALE, float spectrum x float filter */
			const int v = i / hw, u = i - v * hw;
			double re = 0, im = 0;
			for (int y = 0, k = 0; y < ph; ++y) {
				const double ar = a_re[y * hw + u], ai = a_im[y * hw + u];
				re += ar * ch[k] + ai * sh[k];
				im += ai * ch[k] - ar * sh[k];
				k += v; if (k >= ph) k -= ph;
			}
			const float fl = f.whi_filter[v * pw + u];
			b_re[i] = (double)__fmul_rn((float)(re * inv_n), fl);
			b_im[i] = (double)__fmul_rn((float)(im * inv_n), fl);
		}
		__syncthreads();
		for (int i = tid; i < nh; i += FEAT_THREADS) { /* inverse columns */
			const int y = i / hw, u = i - y * hw;
			double re = 0, im = 0;
			for (int v = 0, k = 0; v < ph; ++v) {
				const double br = b_re[v * hw + u], bi = b_im[v * hw + u];
				re += br * ch[k] - bi * sh[k];
				im += bi * ch[k] + br * sh[k];
				k += y; if (k >= ph) k -= ph;
			}
			a_re[i] = re; a_im[i] = im;
		}
		__syncthreads();
		for (int i = tid; i < npx; i += FEAT_THREADS) { /* inverse rows, complex to real; convertTo(CV_8U, 1, 127) */
			const int y = i / pw, x = i - y * pw;
			double s = a_re[y * hw];
			for (int u = 1, k = x; u < hw; ++u) {
				if (2 * u == pw) s += a_re[y * hw + u] * cw[k];
				else s += 2 * (a_re[y * hw + u] * cw[k] - a_im[y * hw + u] * sw[k]);
				k += x; if (k >= pw) k -= pw;
			}
			s_px[i] = (uint8_t)sat_u8_rint(__fadd_rn((float)s, 127.f));
		}
		__syncthreads();
		equalize_hist_block(s_px, s_px2, npx, s_hist, s_lut);
		float* vec = reinterpret_cast<float*>(a_re); /* transforms are done: reuse */
		const float alpha = (float)(1.0 / 127.5), beta = -1.f; /* ConversionFilter(CV_32F, 1/127.5, -1) */
		for (int i = tid; i < npx; i += FEAT_THREADS) vec[i] = __fadd_rn(__fmul_rn((float)s_px2[i], alpha), beta);
		const double nrm = norm_block(vec, npx, false, &s_tmp);
		scale_block(vec, npx, nrm + (double)HIST_EPS); /* UnitNormFilter.cpp:35-38: double norm + float eps */
		float* o = reinterpret_cast<float*>(out) + (int64_t)blockIdx.x * f.dim;
		for (int i = tid; i < npx; i += FEAT_THREADS) o[i] = vec[i];
		return;
	}

	/* ---- histogram features on the binned layer: hog / ehog / lbp ---- */
	const int chn = f.layer_channels, bins = f.bins, R = f.cell_rows, Cc = f.cell_cols, ncell = R * Cc;
	const uint8_t* __restrict__ fl = farena + (int64_t)it.frame * farena_stride + f.layer_offset[it.layer]
			+ ((int64_t)it.y * L.width + it.x) * chn;
	const int fpitch = L.width * chn;
	float* cells = reinterpret_cast<float*>(fsm);                 /* [ncell][bins] */
	float* vec = cells + ncell * bins;                            /* [dim] */
	float* energies = vec + f.dim;                                /* [ncell] */
	const float factor = __fdiv_rn(1.f, 255.f);
	/* HistogramFilter::createCellHistograms: entry (cell, bin) <- its contributions in reference order */
	for (int e = tid; e < ncell * bins; e += FEAT_THREADS) {
		const int cell = e / bins, bin = e - cell * bins;
		const int cr = cell / Cc, cc = cell - cr * Cc;
		float acc = 0.f;
		if (f.interpolate_cells) {
			for (int r = 0; r < ph; ++r) {
				const FeatCache rc = f.row_cache[r];
				const bool r0 = rc.index1 == cr && rc.index1 >= 0, r1 = rc.index2 == cr && rc.index2 < R;
				if (!r0 && !r1) continue;
				const uint8_t* row = fl + (int64_t)r * fpitch;
				for (int c = 0; c < pw; ++c) {
					const FeatCache ccache = f.col_cache[c];
					const bool c0 = ccache.index1 == cc && ccache.index1 >= 0, c1 = ccache.index2 == cc && ccache.index2 < Cc;
					if (!c0 && !c1) continue;
					const uint8_t* px = row + c * chn;
					const int nb = chn == 4 ? 2 : 1;
#pragma unroll
					for (int corner = 0; corner < 4; ++corner) {
						const bool hit = ((corner < 2) ? r0 : r1) && ((corner & 1) ? c1 : c0);
						if (!hit) continue;
						const float rw = corner < 2 ? rc.weight1 : rc.weight2, cwt = (corner & 1) ? ccache.weight2 : ccache.weight1;
						if (chn == 1) { if (px[0] == bin) acc = __fadd_rn(acc, __fmul_rn(rw, cwt)); }
						else
							for (int b = 0; b < nb; ++b)
								if (px[2 * b] == bin)
									acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(__fmul_rn(factor, (float)px[2 * b + 1]), rw), cwt));
					}
				}
			}
		} else {
			const int sr = (cr * ph) / R, er = ((cr + 1) * ph) / R, sc = (cc * pw) / Cc, ec = ((cc + 1) * pw) / Cc;
			for (int r = sr; r < er; ++r) {
				const uint8_t* row = fl + (int64_t)r * fpitch;
				for (int c = sc; c < ec; ++c) {
					const uint8_t* px = row + c * chn;
					if (chn == 1) { if (px[0] == bin) acc = __fadd_rn(acc, 1.f); }
					else {
						if (px[0] == bin) acc = __fadd_rn(acc, __fmul_rn(factor, (float)px[1]));
						if (chn == 4 && px[2] == bin) acc = __fadd_rn(acc, __fmul_rn(factor, (float)px[3]));
					}
				}
			}
		}
		cells[e] = acc;
	}
	__syncthreads();
	float* o = reinterpret_cast<float*>(out) + (int64_t)blockIdx.x * f.dim;
	const int half = bins / 2, su = f.signed_and_unsigned, bs = f.block_size;
	if (f.kind == FDB_FEATURE_EHOG || f.use_hog_filter) {
		/* cell energies (HogFilter.cpp:102-122, ExtendedHogFilter.cpp:72-80,152-156) */
		for (int cell = tid; cell < ncell; cell += FEAT_THREADS) {
			const float* c = cells + cell * bins;
			float energy = 0.f;
			if (su) for (int b = 0; b < half; ++b) { const float u = __fadd_rn(c[b], c[half + b]); energy = __fadd_rn(energy, __fmul_rn(u, u)); }
			else for (int b = 0; b < bins; ++b) energy = __fadd_rn(energy, __fmul_rn(c[b], c[b]));
			energies[cell] = energy;
		}
		__syncthreads();
	}
	if (f.kind == FDB_FEATURE_EHOG) {
		const int per = bins + (su ? half : 0) + 4;
		const float alpha = f.ehog_alpha;
		for (int cell = tid; cell < ncell; cell += FEAT_THREADS) { /* ExtendedHogFilter::createDescriptors, one cell per thread */
			const int r1 = cell / Cc, c1 = cell - r1 * Cc;
			const int r0 = max(r1 - 1, 0), r2 = min(r1 + 1, R - 1), c0 = max(c1 - 1, 0), c2 = min(c1 + 1, Cc - 1);
#define EN(r, c) energies[(r) * Cc + (c)]
			const float n1 = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(EN(r0, c0), EN(r0, c1)), EN(r1, c0)), EN(r1, c1)), HIST_EPS)));
			const float n2 = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(EN(r0, c1), EN(r0, c2)), EN(r1, c1)), EN(r1, c2)), HIST_EPS)));
			const float n3 = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(EN(r1, c0), EN(r1, c1)), EN(r2, c0)), EN(r2, c1)), HIST_EPS)));
			const float n4 = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(EN(r1, c1), EN(r1, c2)), EN(r2, c1)), EN(r2, c2)), HIST_EPS)));
#undef EN
			const float* c = cells + cell * bins;
			float* v = o + cell * per;
			float t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f;
			for (int b = 0; b < bins; ++b) {
				const float h1 = fminf(alpha, __fmul_rn(c[b], n1)), h2 = fminf(alpha, __fmul_rn(c[b], n2));
				const float h3 = fminf(alpha, __fmul_rn(c[b], n3)), h4 = fminf(alpha, __fmul_rn(c[b], n4));
				v[b] = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fadd_rn(h1, h2), h3), h4)); /* 0.5 * float sum: exact in either precision */
				t1 = __fadd_rn(t1, h1); t2 = __fadd_rn(t2, h2); t3 = __fadd_rn(t3, h3); t4 = __fadd_rn(t4, h4);
			}
			v += bins;
			if (su) {
				for (int b = 0; b < half; ++b) {
					const float s = __fadd_rn(c[b], c[b + half]);
					const float h1 = fminf(alpha, __fmul_rn(s, n1)), h2 = fminf(alpha, __fmul_rn(s, n2));
					const float h3 = fminf(alpha, __fmul_rn(s, n3)), h4 = fminf(alpha, __fmul_rn(s, n4));
					v[b] = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fadd_rn(h1, h2), h3), h4));
				}
				v += half;
			}
			/* 0.2357 * t: double product rounded to float (ExtendedHogFilter.cpp:141-144) */
			v[0] = (float)__dmul_rn(0.2357, (double)t1); v[1] = (float)__dmul_rn(0.2357, (double)t2);
			v[2] = (float)__dmul_rn(0.2357, (double)t3); v[3] = (float)__dmul_rn(0.2357, (double)t4);
		}
		return;
	}
	const int br = R - bs + 1, bc = Cc - bs + 1;
	if (f.use_hog_filter) { /* HogFilter::createBlockHistograms (HogFilter.cpp:69-100): one block per thread */
		const int per_cell = bins + (su ? half : 0);
		for (int blk = tid; blk < br * bc; blk += FEAT_THREADS) {
			const int brow = blk / bc, bcol = blk - brow * bc;
			float energy = 0.f;
			for (int r = brow; r < brow + bs; ++r)
				for (int c = bcol; c < bcol + bs; ++c) energy = __fadd_rn(energy, energies[r * Cc + c]);
			const float normalizer = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(energy, HIST_EPS)));
			float* v = o + (int64_t)blk * bs * bs * per_cell;
			for (int r = brow; r < brow + bs; ++r)
				for (int c = bcol; c < bcol + bs; ++c) {
					const float* ch = cells + (r * Cc + c) * bins;
					for (int b = 0; b < bins; ++b) v[b] = __fmul_rn(normalizer, ch[b]);
					v += bins;
					if (su) {
						for (int b = 0; b < half; ++b) v[b] = __fmul_rn(normalizer, __fadd_rn(ch[b], ch[half + b]));
						v += half;
					}
				}
		}
		return;
	}
	if (bs == 1) { /* SpatialHistogramFilter.cpp:59-61: normalise the whole concatenated histogram */
		normalize_block(cells, f.dim, f.normalization, &s_tmp);
		__syncthreads();
		for (int i = tid; i < f.dim; i += FEAT_THREADS) o[i] = cells[i];
		return;
	}
	/* SpatialHistogramFilter::createBlockHistograms (SpatialHistogramFilter.cpp:69-94): one block per thread */
	const int size = f.concatenate ? bs * bs * bins : bins;
	for (int blk = tid; blk < br * bc; blk += FEAT_THREADS) {
		const int brow = blk / bc, bcol = blk - brow * bc;
		float* v = vec + (int64_t)blk * size;
		for (int i = 0; i < size; ++i) v[i] = 0.f;
		float* ins = v;
		for (int r = brow; r < brow + bs; ++r)
			for (int c = bcol; c < bcol + bs; ++c) {
				const float* ch = cells + (r * Cc + c) * bins;
				for (int b = 0; b < bins; ++b) ins[b] = __fadd_rn(ins[b], ch[b]);
				if (f.concatenate) ins += bins;
			}
		normalize_thread(v, size, f.normalization);
		for (int i = 0; i < size; ++i) o[(int64_t)blk * size + i] = v[i];
	}
}

/* ------------------------------------------------------------------------------------------------
 * host side
 * ---------------------------------------------------------------------------------------------- */
static int cv_round_host(double v) { return (int)std::nearbyint(v); }
static uint8_t sat_u8_host(double v) { const int i = cv_round_host(v); return (uint8_t)(i < 0 ? 0 : (i > 255 ? 255 : i)); }

int feature_shape(const fdb_feature_desc& d, int pw, int ph, FeatureShape* s) {
	FeatureShape o{};
	switch (d.kind) {
	case FDB_FEATURE_HQ64: case FDB_FEATURE_GRAY: case FDB_FEATURE_HISTEQ:
		o.dim = pw * ph; o.is_float = 0; o.layer_channels = 0; break;
	case FDB_FEATURE_WHI:
		o.dim = pw * ph; o.is_float = 1; o.layer_channels = 0; break;
	case FDB_FEATURE_HOG: case FDB_FEATURE_EHOG: case FDB_FEATURE_LBP: {
		if (d.cell_size <= 0) return fail(FDB_ERR_INVALID_ARGUMENT, "SpatialHistogramFilter: cellSize must be greater than zero");
		if (d.kind != FDB_FEATURE_EHOG && d.block_size <= 0) return fail(FDB_ERR_INVALID_ARGUMENT, "SpatialHistogramFilter: blockSize must be greater than zero");
		if (d.kind == FDB_FEATURE_LBP) {
			if (d.lbp_type < FDB_LBP8 || d.lbp_type > FDB_LBP4_ROTATED) return fail(FDB_ERR_INVALID_ARGUMENT, "invalid LBP type");
			o.bins = d.lbp_type == FDB_LBP8 ? 256 : (d.lbp_type == FDB_LBP8_UNIFORM ? 59 : 16);
			o.layer_channels = 1;
		} else {
			if (d.bins <= 0 || d.bins > 255) return fail(FDB_ERR_INVALID_ARGUMENT, "HogFilter: binCount must be greater than zero");
			if (d.gradient_kernel != 1 && d.gradient_kernel != 3)
				return fail(FDB_ERR_UNSUPPORTED, "GradientFilter: only kernel sizes 1 and 3 are implemented");
			if (d.blur_kernel != 0) return fail(FDB_ERR_UNSUPPORTED, "GradientFilter: blurKernelSize must be 0");
			if (d.signed_and_unsigned && d.bins % 2 != 0)
				return fail(FDB_ERR_INVALID_ARGUMENT, "HogFilter: the bin size must be even for signed and unsigned gradients to be combined");
			if (d.kind == FDB_FEATURE_EHOG && !(d.ehog_alpha > 0)) return fail(FDB_ERR_INVALID_ARGUMENT, "ExtendedHogFilter: alpha must be greater than zero");
			o.bins = d.bins;
			o.layer_channels = d.interpolate_bins ? 4 : 2;
		}
		o.cell_rows = cv_round((double)ph / (double)d.cell_size);
		o.cell_cols = cv_round((double)pw / (double)d.cell_size);
		if (o.cell_rows < 1 || o.cell_cols < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "feature cells larger than the patch");
		const int half = o.bins / 2;
		if (d.kind == FDB_FEATURE_EHOG) {
			o.dim = o.cell_rows * o.cell_cols * (o.bins + (d.signed_and_unsigned ? half : 0) + 4);
		} else {
			o.use_hog_filter = d.kind == FDB_FEATURE_HOG && !(d.block_size == 1 && !d.signed_and_unsigned);
			const int br = o.cell_rows - d.block_size + 1, bc = o.cell_cols - d.block_size + 1;
			if (br < 1 || bc < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "feature blocks larger than the cell grid");
			if (o.use_hog_filter) o.dim = br * bc * d.block_size * d.block_size * (o.bins + (d.signed_and_unsigned ? half : 0));
			else if (d.block_size == 1) o.dim = o.cell_rows * o.cell_cols * o.bins;
			else o.dim = br * bc * (d.concatenate ? d.block_size * d.block_size * o.bins : o.bins);
		}
		o.is_float = 1;
		break; }
	default: return fail(FDB_ERR_INVALID_ARGUMENT, "unknown feature kind");
	}
	if (s) *s = o;
	return FDB_OK;
}

static void build_cache(std::vector<FeatCache>& cache, int size, int count) { /* HistogramFilter::createCache */
	cache.resize((size_t)size);
	for (int i = 0; i < size; ++i) {
		const double real = (double)count * ((double)i + 0.5) / (double)size - 0.5;
		FeatCache e;
		e.index1 = (int)std::floor(real); e.index2 = e.index1 + 1;
		e.weight2 = (float)(real - e.index1); e.weight1 = 1.f - e.weight2;
		if (e.index1 < 0) { e.index1 = e.index2; e.weight1 = 0; }
		else if (e.index2 >= count) { e.index2 = e.index1; e.weight2 = 0; }
		cache[(size_t)i] = e;
	}
}

size_t feature_smem_bytes(const DevFeature& f) {
	if (f.layer_channels == 0) {
		const size_t px = (size_t)((f.pw * f.ph + 15) & ~15) * 2;
		if (f.kind != FDB_FEATURE_WHI) return px;
		return px + sizeof(double) * 4 * (size_t)f.ph * (f.pw / 2 + 1) + 64;
	}
	return sizeof(float) * ((size_t)f.cell_rows * f.cell_cols * f.bins + (size_t)f.dim + (size_t)f.cell_rows * f.cell_cols) + 64;
}

/* builds the device tables of a feature space for a prepared plan (layer sizes) */
int feature_build(const fdb_feature_desc& d, int pw, int ph, const Plan& plan, DevFeature* out, int64_t* farena_bytes,
		std::vector<void*>& owned) {
	FeatureShape sh;
	int s = feature_shape(d, pw, ph, &sh); if (s) return s;
	DevFeature f{};
	f.kind = d.kind; f.pw = pw; f.ph = ph; f.dim = sh.dim; f.is_float = sh.is_float; f.layer_channels = sh.layer_channels;
	f.bins = sh.bins; f.cell_rows = sh.cell_rows; f.cell_cols = sh.cell_cols; f.use_hog_filter = sh.use_hog_filter;
	f.block_size = d.block_size; f.concatenate = d.concatenate; f.signed_and_unsigned = d.signed_and_unsigned;
	f.normalization = d.normalization; f.interpolate_cells = d.interpolate_cells; f.gradient_kernel = d.gradient_kernel;
	f.lbp_type = d.lbp_type; f.ehog_alpha = d.ehog_alpha;
	if (plan.layers.size() > FDB_MAX_LAYERS) return fail(FDB_ERR_UNSUPPORTED, "too many pyramid layers");
	f.n_layers = (int)plan.layers.size();
	int64_t off = 0; int px = 0;
	for (size_t i = 0; i < plan.layers.size(); ++i) {
		f.layer_offset[i] = off; f.px_prefix[i] = px;
		off += (((int64_t)plan.layers[i].width * plan.layers[i].height * std::max(sh.layer_channels, 1)) + 15) & ~(int64_t)15;
		px += plan.layers[i].width * plan.layers[i].height;
	}
	f.px_prefix[plan.layers.size()] = px;
	*farena_bytes = sh.layer_channels ? off : 0;
	if (feature_smem_bytes(f) > 200 * 1024) return fail(FDB_ERR_UNSUPPORTED, "feature vector too large for the patch kernel");
	if (d.kind == FDB_FEATURE_LBP) {
		/* LbpFilter::LbpFilter (LbpFilter.cpp:20-44): uniform patterns get bins 1.., the rest share bin 0 */
		for (int i = 0; i < 256; ++i) f.lbp_map[i] = (uint8_t)i;
		if (d.lbp_type == FDB_LBP8_UNIFORM) {
			int next = 1;
			for (int i = 0; i < 256; ++i) {
				int transitions = 0, prev = (i >> 7) & 1;
				for (int pos = 0; pos < 8; ++pos) { const int cur = (i >> pos) & 1; if (cur != prev) { ++transitions; prev = cur; } }
				f.lbp_map[i] = transitions <= 2 ? (uint8_t)next++ : (uint8_t)0;
			}
		}
	} else if (sh.layer_channels) {
		/* GradientBinningFilter::GradientBinningFilter (GradientBinningFilter.cpp:18-59), index = gx | gy << 8 */
		const int chn = sh.layer_channels;
		std::vector<uint8_t> lut((size_t)65536 * chn);
		const double pi = 3.1415926535897932384626433832795;
		for (int gx = 0; gx < 256; ++gx)
			for (int gy = 0; gy < 256; ++gy) {
				const double dx = ((double)gx - 127) / 255, dy = ((double)gy - 127) / 255;
				double direction = std::atan2(dy, dx);
				const double magnitude = std::sqrt(dx * dx + dy * dy);
				double bin;
				if (d.signed_gradients) { direction += pi; bin = direction * (unsigned)d.bins / (2 * pi); }
				else { if (direction < 0) direction += pi; bin = direction * (unsigned)d.bins / pi; }
				uint8_t* e = &lut[(size_t)(gx | (gy << 8)) * chn];
				if (chn == 2) {
					e[0] = (uint8_t)((uint8_t)std::round(bin) % d.bins);
					e[1] = sat_u8_host(255 * magnitude);
				} else {
					e[0] = (uint8_t)((uint8_t)std::floor(bin) % d.bins);
					e[2] = (uint8_t)((uint8_t)std::ceil(bin) % d.bins);
					e[3] = sat_u8_host(255 * magnitude * (bin - std::floor(bin)));
					e[1] = sat_u8_host(255 * magnitude - e[3]);
				}
			}
		uint8_t* dl = nullptr;
		s = upload(lut.data(), lut.size(), &dl, owned); if (s) return s;
		f.lut = dl;
	}
	if (sh.layer_channels && d.interpolate_cells) {
		std::vector<FeatCache> rc, cc;
		build_cache(rc, ph, sh.cell_rows); build_cache(cc, pw, sh.cell_cols);
		FeatCache* p = nullptr;
		s = upload(rc.data(), rc.size(), &p, owned); if (s) return s; f.row_cache = p;
		s = upload(cc.data(), cc.size(), &p, owned); if (s) return s; f.col_cache = p;
	}
	if (d.kind == FDB_FEATURE_WHI) {
		/* WhiteningFilter::getFilter (WhiteningFilter.cpp:62-81): float32 grid, |f|^alpha x exp(-(rho/cutoff)^4) */
		std::vector<float> filt((size_t)pw * ph);
		const float nyq = 0.5f;
		for (int row = 0; row < ph; ++row)
			for (int col = 0; col < pw; ++col) {
				const int srow = (row + ph / 2) % ph, scol = (col + pw / 2) % pw;
				const float fx = -nyq + scol * (2 * nyq) / (pw - 1);
				const float fy = -nyq + srow * (2 * nyq) / (ph - 1);
				const float rho = std::sqrt(fx * fx + fy * fy);
				float v = std::pow(rho, d.whi_alpha);
				if (d.whi_cutoff > 0) v = (float)((double)v * std::exp(-std::pow((double)(rho / d.whi_cutoff), 4)));
				filt[(size_t)row * pw + col] = v;
			}
		float* df = nullptr;
		s = upload(filt.data(), filt.size(), &df, owned); if (s) return s; f.whi_filter = df;
		std::vector<double> tw((size_t)2 * (pw + ph));
		const double pi = 3.14159265358979323846;
		for (int k = 0; k < pw; ++k) { tw[(size_t)k] = std::cos(2 * pi * k / pw); tw[(size_t)pw + k] = std::sin(2 * pi * k / pw); }
		for (int k = 0; k < ph; ++k) { tw[(size_t)2 * pw + k] = std::cos(2 * pi * k / ph); tw[(size_t)2 * pw + ph + k] = std::sin(2 * pi * k / ph); }
		double* dt = nullptr;
		s = upload(tw.data(), tw.size(), &dt, owned); if (s) return s; f.twiddle = dt;
	}
	*out = f;
	return FDB_OK;
}

int feature_configure() {
	return (int)cudaFuncSetAttribute(feature_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

void launch_feature_layers(cudaStream_t st, const DevFeature& f, const uint8_t* frames, int W, int H, int n_frames,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, uint8_t* farena, int64_t farena_stride) {
	if (!f.layer_channels || n_frames == 0 || f.px_prefix[f.n_layers] == 0) return;
	dim3 grid((unsigned)((f.px_prefix[f.n_layers] + 255) / 256), (unsigned)n_frames);
	feature_layer_kernel<<<grid, 256, 0, st>>>(f, frames, W, H, arena, arena_stride, layers, farena, farena_stride);
}

void launch_feature_patches(cudaStream_t st, const DevFeature& f, const uint8_t* frames, int W, int H,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, const uint8_t* farena, int64_t farena_stride,
		const SvmItem* items, int n_items, void* out) {
	if (n_items == 0) return;
	feature_patch_kernel<<<(unsigned)n_items, FEAT_THREADS, feature_smem_bytes(f), st>>>(f, frames, W, H, arena, arena_stride, layers,
			farena, farena_stride, items, out);
}

} // namespace fdb
