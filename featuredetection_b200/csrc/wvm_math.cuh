/*
 * wvm_math.cuh - the per-filter arithmetic of the WVM shared by all stage-1 kernels.
 *
 * Everything after the integer rectangle sums of linEvalWvmHisteq64
 * (libClassification/src/classification/WvmClassifier.cpp:308-341) in the reference's
 * operation order: float sumv0, double sum_xp / norm, (float)exp(double), float weighted sum.
 * Explicit _rn intrinsics keep nvcc from contracting multiplies and adds into FMAs.
 */
#ifndef FDB_WVM_MATH_CUH_
#define FDB_WVM_MATH_CUH_

#include <cuda_runtime.h>
#include <cstdint>

#include "wvm_device.h"

namespace fdb {

/* acc[v] = exact integer sum over the rectangles of grey value v+1 (v < nv).
 * Returns sum_xp after :313 (the value u_kernel_eval[level % per_level] takes at :314, as a double). */
template <int MAXV>
__device__ __forceinline__ double wvm_sum_xp(const DevWvm& m, int level, const uint32_t* acc, int nv, float total_f, float un) {
	const double* __restrict__ val = m.val + __ldg(m.val_off + level);
	float sumv0 = total_f;
	double sum_xp = 0.0;
#pragma unroll
	for (int v = 0; v < MAXV; ++v)
		if (v < nv) {
			const float sumv = (float)acc[v];                                        /* exact: < 2^24 */
			sumv0 = __fsub_rn(sumv0, sumv);                                          /* :308 */
			sum_xp = __dadd_rn(sum_xp, __dmul_rn((double)sumv, __ldg(val + v + 1))); /* :309 */
		}
	sum_xp = __dadd_rn(sum_xp, __dmul_rn((double)sumv0, __ldg(val)));                /* :312 */
	return __dadd_rn(sum_xp, (double)un);                                            /* :313 */
}

/* *un is u_kernel_eval[level % per_level] (read, then overwritten, :313-314).
 * Returns hk_kernel_eval[level] (:333). */
template <int MAXV>
__device__ __forceinline__ float wvm_kernel_value(const DevWvm& m, int level, const uint32_t* acc, int nv,
		float total_f, float sum_xx, float* un) {
	const double sum_xp = wvm_sum_xp<MAXV>(m, level, acc, nv, total_f, *un);
	*un = (float)sum_xp;                                                             /* :314 */
	double norm = __dsub_rn((double)sum_xx, __dmul_rn(2.0, sum_xp));                 /* :316 */
	norm = __dadd_rn(norm, __ldg(m.app_rsv_convol + level));                         /* :322 */
	return (float)exp(__dmul_rn((double)(-m.basis_param), norm));                    /* :333 */
}

/* same with only four grey values and the sums passed by value (register friendly) */
__device__ __forceinline__ float wvm_kernel_value4(const DevWvm& m, int level, uint32_t a0, uint32_t a1, uint32_t a2,
		uint32_t a3, int nv, float total_f, float sum_xx, float* un) {
	const uint32_t acc[4] = {a0, a1, a2, a3};
	return wvm_kernel_value<4>(m, level, acc, nv, total_f, sum_xx, un);
}

/* WvmClassifier::classify(pair) (WvmClassifier.cpp:91-98) + candidate append */
__device__ __forceinline__ void wvm_emit(const DevWvm& m, int frame, int win, int windows_per_frame, int level, float fout,
		fdb_window_score* __restrict__ dense, Candidate* __restrict__ cand, int* __restrict__ cand_count, int cand_cap) {
	if (dense) {
		fdb_window_score s; s.fout = fout; s.level = level;
		dense[(int64_t)frame * windows_per_frame + win] = s;
	}
	if (cand && level + 1 == m.num_lin && fout >= __ldg(m.thresholds + level)) {
		const int slot = atomicAdd(cand_count, 1); /* one list per launch; the host restores (frame, window) order */
		if (slot < cand_cap) {
			Candidate c; c.window = win; c.level = level; c.fout = fout; c.frame = frame;
			cand[slot] = c;
		}
	}
}

} // namespace fdb
#endif
