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
#include <algorithm>
#include <cstdint>

#include "fdb_internal.h"
#include "wvm_device.h"

namespace fdb {

/* equalised patch words of one window: shared memory column (stride = threads per CTA) or a
 * column of the deep queue (stride = queue capacity) */
template <bool GLOBAL>
struct XWords {
	const uint32_t* base;
	int stride;
	__device__ __forceinline__ uint32_t operator[](int j) const { return GLOBAL ? __ldg(base + (size_t)j * stride) : base[j * stride]; }
};

/* linEvalWvmHisteq64 (WvmClassifier.cpp:191-346) for one filter. *un is u_kernel_eval[level % per_level],
 * hk[0..level] the kernel values so far (hk[level] is written); hk_len only bounds unrolling. */
template <class X>
__device__ __forceinline__ float wvm_level(const DevWvm& m, int level, const X& xs, float total_f, float sum_xx,
		float* un, float* hk, int static_level) {
	const int nv = m.cntval[level] - 1;
	const uint32_t* __restrict__ mk = m.masks + m.mask_off[level];
	const double* __restrict__ val = m.val + m.val_off[level];
	float sumv0 = total_f;
	double sum_xp = 0.0;
	/* any number of grey values per filter (WvmClassifier.hpp:107-124 sets no limit): FDB_MAX_VALUES at a time, each group one
	 * pass over the patch words, folded into sumv0 / sum_xp in the reference's order v = 1, 2, ... (:277-309) */
	for (int v0 = 0; v0 < nv; v0 += FDB_MAX_VALUES) {
		const int nvc = min(FDB_MAX_VALUES, nv - v0);
		uint32_t acc[FDB_MAX_VALUES];
#pragma unroll
		for (int v = 0; v < FDB_MAX_VALUES; ++v) acc[v] = 0;
		if (nv == 4) { /* common case: one 16-byte load brings the four masks of a word */
			const uint4* __restrict__ mk4 = reinterpret_cast<const uint4*>(mk);
#pragma unroll 4
			for (int j = 0; j < m.nwords; ++j) {
				const uint32_t xw = xs[j];
				const uint4 k = __ldg(mk4 + j);
				acc[0] = __dp4a(xw, k.x, acc[0]); acc[1] = __dp4a(xw, k.y, acc[1]);
				acc[2] = __dp4a(xw, k.z, acc[2]); acc[3] = __dp4a(xw, k.w, acc[3]);
			}
		} else {
			for (int j = 0; j < m.nwords; ++j) {
				const uint32_t xw = xs[j];
#pragma unroll
				for (int v = 0; v < FDB_MAX_VALUES; ++v)
					if (v < nvc) acc[v] = __dp4a(xw, __ldg(mk + j * nv + v0 + v), acc[v]);
			}
		}
#pragma unroll
		for (int v = 0; v < FDB_MAX_VALUES; ++v)
			if (v < nvc) {
				const float sumv = (float)acc[v];                                 /* exact: < 2^24 */
				sumv0 = __fsub_rn(sumv0, sumv);                                   /* :308 */
				sum_xp = __dadd_rn(sum_xp, __dmul_rn((double)sumv, __ldg(val + v0 + v + 1))); /* :309 */
			}
	}
	sum_xp = __dadd_rn(sum_xp, __dmul_rn((double)sumv0, __ldg(val)));         /* :312 */
	sum_xp = __dadd_rn(sum_xp, (double)*un);                                  /* :313 */
	*un = (float)sum_xp;                                                      /* :314 */
	double norm = __dsub_rn((double)sum_xx, __dmul_rn(2.0, sum_xp));          /* :316 */
	norm = __dadd_rn(norm, __ldg(m.app_rsv_convol + level));                  /* :322 */
	const float k = (float)exp(__dmul_rn((double)(-m.basis_param), norm));    /* :333 */
	const float* __restrict__ wgt = m.hk_weights + (size_t)level * (level + 1) / 2;
	float res = -__ldg(m.lin_thresholds + level);                             /* :201 */
	if (static_level >= 0) {
#pragma unroll
		for (int p = 0; p < WVM_KA; ++p) {                                    /* :340-341, registers */
			if (p == static_level) hk[p] = k;
			if (p <= static_level) res = __fadd_rn(res, __fmul_rn(__ldg(wgt + p), hk[p]));
		}
	} else {
		hk[level] = k;
		for (int p = 0; p <= level; ++p)                                      /* :340-341 */
			res = __fadd_rn(res, __fmul_rn(__ldg(wgt + p), hk[p]));
	}
	return res;
}

/* continues the cascade after the first WVM_KA filters (WvmClassifier.cpp:134-138) */
template <class X>
__device__ __noinline__ void wvm_deep(const DevWvm& m, const X& xs, float total_f, float sum_xx,
		const float* hk_init, const float* u_init, int* level_io, float* fout_io) {
	float hk[FDB_MAX_FILTERS];
	float u[FDB_MAX_PER_LEVEL];
	for (int n = 0; n < m.per_level; ++n) u[n] = 0.f;
#pragma unroll
	for (int i = 0; i < WVM_KA; ++i) {
		hk[i] = hk_init[i];
		if (i < m.per_level) u[i] = u_init[i];
	}
	int level = WVM_KA - 1;
	float fout;
	do {
		++level;
		fout = wvm_level(m, level, xs, total_f, sum_xx, &u[level % m.per_level], hk, -1);
	} while (fout >= __ldg(m.thresholds + level) && level + 1 < m.num_used);
	*level_io = level;
	*fout_io = fout;
}

template <bool FROM_PATCHES>
__global__ void __launch_bounds__(WVM_THREADS) wvm_window_kernel(const DevWvm m,
		const uint8_t* __restrict__ frames, int W, int H,
		const uint8_t* __restrict__ arena, int64_t arena_stride,
		const DevLayer* __restrict__ layers, int n_layers, int windows_per_frame,
		const uint8_t* __restrict__ patches_in,     /* FROM_PATCHES: [n][npix] feature vectors */
		fdb_window_score* __restrict__ dense,       /* nullable: [frame][window] */
		uint8_t* __restrict__ patches_out,          /* nullable: [frame][window][npix] */
		Candidate* __restrict__ cand, int* __restrict__ cand_count, int cand_cap, const DeepQueue q) {
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
				: arena + (int64_t)frame * arena_stride + L.offset) + (int64_t)y * L.pitch + x;

		/* --- HistEq64: 64-bin histogram (HistEq64Filter.cpp:59-67) --- */
#pragma unroll
		for (int k = 0; k < 32; ++k) s_hist[k * T + tid] = 0;
		for (int r = 0; r < ph; ++r) {
			const uint8_t* row = img + (int64_t)r * L.pitch;
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
			const uint8_t* row = img + (int64_t)r * L.pitch;
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

	/* --- WvmClassifier::computeHyperplaneDistance (WvmClassifier.cpp:129-138), first WVM_KA filters.
	 * The loop is fully unrolled so that hk_kernel_eval / u_kernel_eval of these levels live in
	 * registers. Most windows leave the cascade here; the few that survive all WVM_KA filters are
	 * handed to wvm_deep_kernel through a queue so that they do not stall the other 31 lanes. --- */
	const float total_f = (float)total;
	float hk[WVM_KA], u[WVM_KA];
#pragma unroll
	for (int i = 0; i < WVM_KA; ++i) { hk[i] = 0.f; u[i] = 0.f; }
	int level = -1;
	float fout = 0.f;
	bool alive = true;
	const XWords<false> xs = {s_patch + tid, T};
#pragma unroll
	for (int L = 0; L < WVM_KA; ++L) {
		if (alive) {
			level = L;
			const int n = L % m.per_level;
			float un = 0.f;
#pragma unroll
			for (int i = 0; i < WVM_KA; ++i) if (i == n) un = u[i];
			fout = wvm_level(m, L, xs, total_f, sum_xx, &un, hk, L);
#pragma unroll
			for (int i = 0; i < WVM_KA; ++i) if (i == n) u[i] = un;
			alive = fout >= m.thresholds[L] && L + 1 < m.num_used;
		}
	}
	const int64_t gw = (int64_t)frame * windows_per_frame + win;
	if (alive) {
		const int slot = q.rec ? atomicAdd(q.count, 1) : q.cap;
		if (slot < q.cap) {
			DeepRec r;
			r.frame = frame; r.window = win; r.total_f = total_f; r.sum_xx = sum_xx;
#pragma unroll
			for (int i = 0; i < WVM_KA; ++i) { r.hk[i] = hk[i]; r.u[i] = u[i]; }
			q.rec[slot] = r;
			for (int j = 0; j < m.nwords; ++j) q.patch[(size_t)j * q.cap + slot] = s_patch[j * T + tid];
			return; /* wvm_deep_kernel finishes this window */
		}
		/* queue full (e.g. a model without early exits): finish inline */
		wvm_deep(m, xs, total_f, sum_xx, hk, u, &level, &fout);
	}
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

/* Second half of the cascade for the windows that survived the first WVM_KA filters: one thread
 * per queued window; all lanes of a warp now run long, similar filter chains (no divergence). */
__global__ void __launch_bounds__(WVM_THREADS) wvm_deep_kernel(const DevWvm m, const DeepQueue q, int windows_per_frame,
		fdb_window_score* __restrict__ dense, Candidate* __restrict__ cand, int* __restrict__ cand_count, int cand_cap) {
	const int n = min(*q.count, q.cap);
	for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
		const DeepRec r = q.rec[slot];
		const XWords<true> xs = {q.patch + slot, q.cap};
		int level = WVM_KA - 1;
		float fout = 0.f;
		wvm_deep(m, xs, r.total_f, r.sum_xx, r.hk, r.u, &level, &fout);
		if (dense) { fdb_window_score s; s.fout = fout; s.level = level; dense[(int64_t)r.frame * windows_per_frame + r.window] = s; }
		if (cand && level + 1 == m.num_lin && fout >= m.thresholds[level]) {
			const int c_slot = atomicAdd(cand_count, 1);
			if (c_slot < cand_cap) {
				Candidate c; c.window = r.window; c.level = level; c.fout = fout; c.frame = r.frame;
				cand[c_slot] = c;
			}
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
		fdb_window_score* dense, uint8_t* patches_out, Candidate* cand, int* cand_count, int cand_cap, const DeepQueue& q) {
	if (windows_per_frame == 0 || n_frames == 0) return;
	dim3 grid((unsigned)((windows_per_frame + WVM_THREADS - 1) / WVM_THREADS), (unsigned)n_frames);
	wvm_window_kernel<false><<<grid, WVM_THREADS, wvm_smem_bytes(m), st>>>(m, frames, W, H, arena, arena_stride,
			layers, n_layers, windows_per_frame, nullptr, dense, patches_out, cand, cand_count, cand_cap, q);
	if (q.rec && m.num_lin > 0) {
		const int blocks = std::min((q.cap + WVM_THREADS - 1) / WVM_THREADS, 148 * 8);
		wvm_deep_kernel<<<blocks, WVM_THREADS, 0, st>>>(m, q, windows_per_frame, dense, cand, cand_count, cand_cap);
	}
}

void launch_wvm_patches(cudaStream_t st, const DevWvm& m, const uint8_t* patches, int n,
		fdb_window_score* dense) {
	if (n == 0) return;
	dim3 grid((unsigned)((n + WVM_THREADS - 1) / WVM_THREADS), 1);
	DeepQueue q{}; /* no queue: deep windows finish inline */
	wvm_window_kernel<true><<<grid, WVM_THREADS, wvm_smem_bytes(m), st>>>(m, nullptr, 0, 0, nullptr, 0,
			nullptr, 0, n, patches, dense, nullptr, nullptr, nullptr, 0, q);
}

} // namespace fdb
