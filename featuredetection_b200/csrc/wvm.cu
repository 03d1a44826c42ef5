/*
 * wvm.cu - stage 1 of the cascade on the GPU: per-window HistEq64 + WVM evaluation (sm_100a).
 *
 * One thread owns one sliding window and runs, fused in one kernel,
 *   DirectPyramidFeatureExtractor::extract's window addressing  (DirectPyramidFeatureExtractor.cpp:110-118)
 *   HistEq64Filter::applyTo                                      (HistEq64Filter.cpp:32-125)
 *   IImg::calIImgPatch (only the two totals the evaluator reads)  (IImg.cpp:26-65)
 *   WvmClassifier::computeHyperplaneDistance / linEvalWvmHisteq64 (WvmClassifier.cpp:100-149,191-346)
 *   WvmClassifier::classify(pair)                                 (WvmClassifier.cpp:91-98)
 *
 * Exactness notes (see DESIGN.md "numeric ledger"):
 *  - the 64-bin cumulative histogram is accumulated sequentially in float32 exactly like
 *    HistEq64Filter.cpp:77-81; floor(cdf + 0.5) is taken in exact arithmetic.
 *  - the reference evaluates x.p through rectangle sums on a float integral image; every one of
 *    those partial sums is an integer below 2^24 (checked at model creation), i.e. exact, so the
 *    kernel computes the same integers as dot products of the equalised patch with per-grey-level
 *    rectangle-coverage masks (dp4a) - no integral image is materialised.
 *  - sum(x^2) is accumulated row by row in float32 in the order of IImg.cpp:33-47 (it exceeds 2^24).
 *  - the double-precision chain (sum_xp, norm, exp) and the float32 weighted kernel sum keep the
 *    reference's operation order; FMA contraction is disabled with explicit _rn intrinsics.
 *
 * Shared memory per thread (T threads per CTA), column layout [word][thread] => conflict-free:
 *   32 words  histogram (two 16-bit counts per word), later overwritten by the 64-entry eq LUT
 *   nwords    equalised patch, 4 pixels per word
 */
#include <cuda_runtime.h>
#include <cstdint>

#include "fdb_internal.h"
#include "wvm_device.h"

namespace fdb {

template <bool FROM_PATCHES>
__global__ void __launch_bounds__(WVM_THREADS) wvm_window_kernel(const DevWvm m,
		const uint8_t* __restrict__ frames, int W, int H,
		const uint8_t* __restrict__ arena, int64_t arena_stride,
		const DevLayer* __restrict__ layers, int n_layers, int windows_per_frame,
		const uint8_t* __restrict__ patches_in,     /* FROM_PATCHES: [n][npix] feature vectors */
		fdb_window_score* __restrict__ dense,       /* nullable: [frame][window] */
		uint8_t* __restrict__ patches_out,          /* nullable: [frame][window][npix] */
		Candidate* __restrict__ cand, int* __restrict__ cand_count, int cand_cap) {
	extern __shared__ uint32_t smem[];
	__shared__ DevLayer s_layers[FDB_MAX_LAYERS];
	const int tid = threadIdx.x;
	const int T = WVM_THREADS;
	uint32_t* s_hist = smem;                 /* [32][T] */
	uint32_t* s_patch = smem + 32 * T;       /* [nwords][T] */

	if (!FROM_PATCHES) {
		for (int i = tid; i < n_layers; i += T) s_layers[i] = layers[i];
		__syncthreads();
	}
	const int frame = blockIdx.y;
	const int win = blockIdx.x * T + tid;
	if (win >= windows_per_frame) return;

	const int pw = m.fsx, ph = m.fsy, npix = pw * ph;
	float sum_xx = 0.f;   /* iimg_xx->data[dr] */
	int total = 0;        /* iimg_x->data[dr] (exact integer) */

	if (FROM_PATCHES) {
		const uint8_t* p = patches_in + ((int64_t)frame * windows_per_frame + win) * npix;
		uint32_t word = 0; int rowsq = 0, c = 0, r = 0;
		for (int i = 0; i < npix; ++i) {
			const uint32_t e = p[i];
			word |= e << (8 * (i & 3));
			if ((i & 3) == 3 || i == npix - 1) { s_patch[(i >> 2) * T + tid] = word; word = 0; }
			rowsq += e * e; total += e;
			if (++c == pw) {
				sum_xx = r == 0 ? (float)rowsq : __fadd_rn(sum_xx, (float)rowsq);
				rowsq = 0; c = 0; ++r;
			}
		}
	} else {
		/* window -> (layer, x, y): canonical order = layer index asc, y, x */
		int li = 0;
		while (li + 1 < n_layers && win >= s_layers[li + 1].first_window) ++li;
		const DevLayer L = s_layers[li];
		const int local = win - L.first_window;
		const int iy = local / L.windows_x, ix = local - iy * L.windows_x;
		const int x = L.begin_x + ix * m.step_x, y = L.begin_y + iy * m.step_y;
		const uint8_t* __restrict__ img = (L.offset < 0 ? frames + (int64_t)frame * W * H
				: arena + (int64_t)frame * arena_stride + L.offset) + (int64_t)y * L.width + x;

		/* --- HistEq64: 64-bin histogram (HistEq64Filter.cpp:59-67) --- */
#pragma unroll
		for (int k = 0; k < 32; ++k) s_hist[k * T + tid] = 0;
		for (int r = 0; r < ph; ++r) {
			const uint8_t* row = img + (int64_t)r * L.width;
			for (int c = 0; c < pw; ++c) {
				const int b = row[c] >> 2;
				s_hist[(b >> 1) * T + tid] += 1u << ((b & 1) * 16);
			}
		}
		/* --- stretch + sequential float cumsum + rounding (:34,:70-87,:97) --- */
		const float stretch = __fdiv_rn(255.0f, (float)npix);
		float cdf = 0.f;
		uint32_t eqw = 0;
#pragma unroll 4
		for (int k = 0; k < 32; ++k) {
			const uint32_t hw = s_hist[k * T + tid];
#pragma unroll
			for (int half = 0; half < 2; ++half) {
				const float cnt = (float)((hw >> (16 * half)) & 0xffffu);
				cdf = __fadd_rn(cdf, __fmul_rn(cnt, stretch));
				/* (uchar)floor((double)cdf + 0.5), exactly: floor(cdf) + (frac >= 0.5) */
				const float fl = floorf(cdf);
				const int e = (int)fl + (__fsub_rn(cdf, fl) >= 0.5f ? 1 : 0);
				eqw |= (uint32_t)(e & 255) << (8 * ((2 * k + half) & 3));
			}
			if (k & 1) { s_hist[(k >> 1) * T + tid] = eqw; eqw = 0; } /* LUT word k/2 overwrites consumed counts */
		}
		/* --- apply the LUT, pack 4 pixels per word, accumulate the two integral-image totals --- */
		uint32_t word = 0; int i = 0;
		for (int r = 0; r < ph; ++r) {
			const uint8_t* row = img + (int64_t)r * L.width;
			int rowsq = 0;
			for (int c = 0; c < pw; ++c, ++i) {
				const int b = row[c] >> 2;
				const uint32_t e = (s_hist[(b >> 2) * T + tid] >> ((b & 3) * 8)) & 255u;
				word |= e << (8 * (i & 3));
				if ((i & 3) == 3 || i == npix - 1) { s_patch[(i >> 2) * T + tid] = word; word = 0; }
				rowsq += e * e; total += e;
			}
			sum_xx = r == 0 ? (float)rowsq : __fadd_rn(sum_xx, (float)rowsq);
		}
	}

	if (patches_out) {
		uint8_t* o = patches_out + ((int64_t)frame * windows_per_frame + win) * npix;
		for (int i = 0; i < npix; ++i) o[i] = (uint8_t)(s_patch[(i >> 2) * T + tid] >> (8 * (i & 3)));
	}
	if (m.num_lin == 0) return; /* extraction only */

	/* --- WvmClassifier::computeHyperplaneDistance (WvmClassifier.cpp:129-138) --- */
	float hk[FDB_MAX_FILTERS];       /* hk_kernel_eval */
	float u[FDB_MAX_PER_LEVEL];      /* u_kernel_eval */
	for (int n = 0; n < m.per_level; ++n) u[n] = 0.f;
	const int nwords = m.nwords;
	const float total_f = (float)total;
	int level = -1;
	float fout = 0.f;
	do {
		++level;
		const int n = level % m.per_level;
		/* linEvalWvmHisteq64 (WvmClassifier.cpp:191-346) */
		const int nv = m.cntval[level] - 1;
		const uint32_t* __restrict__ mk = m.masks + m.mask_off[level];
		uint32_t acc[FDB_MAX_VALUES];
#pragma unroll
		for (int v = 0; v < FDB_MAX_VALUES; ++v) acc[v] = 0;
		for (int j = 0; j < nwords; ++j) {
			const uint32_t xw = s_patch[j * T + tid];
#pragma unroll
			for (int v = 0; v < FDB_MAX_VALUES; ++v)
				if (v < nv) acc[v] = __dp4a(xw, mk[j * nv + v], acc[v]);
		}
		const double* __restrict__ val = m.val + m.val_off[level];
		float sumv0 = total_f;
		double sum_xp = 0.0;
#pragma unroll
		for (int v = 0; v < FDB_MAX_VALUES; ++v)
			if (v < nv) {
				const float sumv = (float)acc[v];                     /* exact: < 2^24 */
				sumv0 = __fsub_rn(sumv0, sumv);                       /* :308 */
				sum_xp = __dadd_rn(sum_xp, __dmul_rn((double)sumv, val[v + 1])); /* :309 */
			}
		sum_xp = __dadd_rn(sum_xp, __dmul_rn((double)sumv0, val[0]));  /* :312 */
		sum_xp = __dadd_rn(sum_xp, (double)u[n]);                      /* :313 */
		u[n] = (float)sum_xp;                                          /* :314 */
		double norm = __dsub_rn((double)sum_xx, __dmul_rn(2.0, sum_xp)); /* :316 */
		norm = __dadd_rn(norm, m.app_rsv_convol[level]);               /* :322 */
		hk[level] = (float)exp(__dmul_rn((double)(-m.basis_param), norm)); /* :333 */
		const float* __restrict__ wgt = m.hk_weights + (size_t)level * (level + 1) / 2;
		float res = -m.lin_thresholds[level];                          /* :201 */
		for (int p = 0; p <= level; ++p)                               /* :340-341 */
			res = __fadd_rn(res, __fmul_rn(wgt[p], hk[p]));
		fout = res;
	} while (fout >= m.thresholds[level] && level + 1 < m.num_used);

	const int64_t gw = (int64_t)frame * windows_per_frame + win;
	if (dense) { fdb_window_score s; s.fout = fout; s.level = level; dense[gw] = s; }
	/* WvmClassifier::classify(pair) (WvmClassifier.cpp:91-98) */
	if (cand && level + 1 == m.num_lin && fout >= m.thresholds[level]) {
		const int slot = atomicAdd(cand_count, 1); /* one list per launch; the host restores (frame, window) order */
		if (slot < cand_cap) {
			Candidate c; c.window = win; c.level = level; c.fout = fout; c.frame = frame;
			cand[slot] = c;
		}
	}
}

size_t wvm_smem_bytes(const DevWvm& m) {
	return (size_t)(32 + m.nwords) * WVM_THREADS * sizeof(uint32_t);
}

int wvm_configure() {
	cudaError_t e = cudaFuncSetAttribute(wvm_window_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(wvm_window_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	return (int)e;
}

void launch_wvm_windows(cudaStream_t st, const DevWvm& m, const uint8_t* frames, int W, int H, int n_frames,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, int n_layers, int windows_per_frame,
		fdb_window_score* dense, uint8_t* patches_out, Candidate* cand, int* cand_count, int cand_cap) {
	if (windows_per_frame == 0 || n_frames == 0) return;
	dim3 grid((unsigned)((windows_per_frame + WVM_THREADS - 1) / WVM_THREADS), (unsigned)n_frames);
	wvm_window_kernel<false><<<grid, WVM_THREADS, wvm_smem_bytes(m), st>>>(m, frames, W, H, arena, arena_stride,
			layers, n_layers, windows_per_frame, nullptr, dense, patches_out, cand, cand_count, cand_cap);
}

void launch_wvm_patches(cudaStream_t st, const DevWvm& m, const uint8_t* patches, int n,
		fdb_window_score* dense) {
	if (n == 0) return;
	dim3 grid((unsigned)((n + WVM_THREADS - 1) / WVM_THREADS), 1);
	wvm_window_kernel<true><<<grid, WVM_THREADS, wvm_smem_bytes(m), st>>>(m, nullptr, 0, 0, nullptr, 0,
			nullptr, 0, n, patches, dense, nullptr, nullptr, nullptr, 0);
}

} // namespace fdb
