/*
 * wvm_strip_mma.cu - stage 1 fast path, second generation (sm_100a): the rectangle sums of the first
 * WVM_KA filters of ALL 32 windows of a warp row as ONE integer matrix product on the tensor cores.
 *
 * Same arithmetic as wvm_strip.cu / wvm.cu (see there and DESIGN.md for the reference citations:
 * HistEq64Filter.cpp:32-125, WvmClassifier.cpp:129-138,191-346, IImg.cpp:33-47).  What changes is who
 * computes what:
 *
 *   lane = window (as before)        sliding 64-bin histogram (u32 counts, column layout, bank = lane),
 *                                    sequential float cumsum -> equalisation LUT, sum(x), sum(x^2)
 *   warp = 32 x K x 32 matrix product D[window][filter, value] = A[window][pixel] * B[pixel][filter, value]
 *                                    A = equalised pixels (u8), built straight into mma fragments from the
 *                                    bin tile and the LUTs: lane (g, t) of a quad serves windows g, g+8,
 *                                    g+16, g+24, so every LUT is stored 4 times (once per t) in a
 *                                    [bin][reader lane] word layout whose 4 bytes belong to those 4
 *                                    windows: bank = reader lane, conflict free, and the byte arrives
 *                                    without any extraction arithmetic.  B = rectangle coverage counts
 *                                    (u8), pre-swizzled into fragment order on the host (DevWvm::bfrag).
 *                                    mma.sync.m16n8k32.u8.u8.s32: exact (sums < 2^24).
 *   lane = window again              D goes through shared memory (aliasing the LUT), then the scalar
 *                                    cascade (double-precision kernel value, float weighted sum, threshold)
 *                                    runs per lane until its window is rejected; survivors of all WVM_KA
 *                                    filters rebuild their patch words and enter the deep queue.
 *
 * The 400-register-byte patch per lane, the 100 uniform mask loads per filter and the divergent dp4a
 * loops of the first generation are gone; what diverges now is only the ~150-instruction scalar tail.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>

#include "fdb_internal.h"
#include "wvm_device.h"
#include "wvm_math.cuh"

namespace fdb {

#ifndef STRIP2_WARPS
#define STRIP2_WARPS 4
#endif
#ifndef STRIP2_MIN_CTAS
#define STRIP2_MIN_CTAS 3
#endif
#define STRIP2_PITCH 64      /* bytes per tile row */
#define STRIP2_STAGE 40      /* ints per window row of the D staging area: 16-byte aligned rows; the 16 int2 fragment stores per
                              * window are conflict-free (half-warp rows 8 banks apart), the rarer uint4 reads 2-way */
#define STRIP2_HKU (32 * STRIP2_STAGE) /* float offset of the cascade state behind the staging area: hk[8][32], u[8][32] */

static_assert(WVM_KA == 8, "the fragment table holds 8 filters x 4 grey values = 32 columns");

template <int PW, int PH>
struct MmaCfg {
	static constexpr int NW = PW * PH / 4;
	static constexpr int KS = (NW + 7) / 8;  /* k-steps of 32 pixels */
	static constexpr int TILE_ROWS = STRIP_TILE_ROWS;
	static constexpr int TILE_BYTES = TILE_ROWS * STRIP2_PITCH;
	static constexpr int HIST_BYTES = 64 * 32 * 4;
	static constexpr int LUT_BYTES = 64 * 32 * 4;
	static constexpr size_t SMEM = (size_t)STRIP2_WARPS * (TILE_BYTES + HIST_BYTES + LUT_BYTES + 8 /* mbarrier */);
};

__device__ __forceinline__ void mma_u8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
	asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
			: "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

/* one step of the sequential cumulative histogram (HistEq64Filter.cpp:70-87,97) */
__device__ __forceinline__ uint32_t hq_step(float& cdf, uint32_t cnt, float stretch) {
	cdf = __fadd_rn(cdf, __fmul_rn((float)cnt, stretch));
	const float fl = floorf(cdf); /* (uchar)floor((double)cdf + 0.5) == floor(cdf) + (frac >= 0.5) */
	return ((uint32_t)(int)fl + (__fsub_rn(cdf, fl) >= 0.5f ? 1u : 0u)) & 255u;
}

template <int PW, int PH>
__global__ void __launch_bounds__(STRIP2_WARPS * 32, STRIP2_MIN_CTAS) wvm_strip_mma_kernel(const DevWvm m,
		const uint8_t* __restrict__ frames, int W, int H,
		const uint8_t* __restrict__ arena, int64_t arena_stride,
		const DevLayer* __restrict__ layers, const Strip* __restrict__ strips, int n_strips, int windows_per_frame,
		fdb_window_score* __restrict__ dense,
		Candidate* __restrict__ cand, int* __restrict__ cand_count, int cand_cap, const DeepQueue q,
		const CUtensorMap* __restrict__ tmaps) {
	static_assert(PW % 4 == 0 && PW <= 32, "patch width must be a multiple of 4, at most 32");
	static_assert((STRIP2_HKU + 2 * WVM_KA * 32) * 4 <= MmaCfg<PW, PH>::LUT_BYTES, "cascade state must fit behind the staging area");
	using Cfg = MmaCfg<PW, PH>;
	constexpr int NW = Cfg::NW, WPR = PW / 4, KS = Cfg::KS, TR = Cfg::TILE_ROWS;
	extern __shared__ __align__(128) uint8_t smem8[];
	uint64_t* const s_mbar = reinterpret_cast<uint64_t*>(smem8 + STRIP2_WARPS * (Cfg::TILE_BYTES + Cfg::HIST_BYTES + Cfg::LUT_BYTES));
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint8_t* const s_tile = smem8 + warp * Cfg::TILE_BYTES;
	uint32_t* const s_hist = reinterpret_cast<uint32_t*>(smem8 + STRIP2_WARPS * Cfg::TILE_BYTES) + warp * 2048 + lane; /* bin b: s_hist[b * 32] */
	uint8_t* const s_lut = smem8 + STRIP2_WARPS * (Cfg::TILE_BYTES + Cfg::HIST_BYTES) + warp * Cfg::LUT_BYTES;          /* [64][32] words */
	int* const s_stage = reinterpret_cast<int*>(s_lut);                                                                   /* [32][STRIP2_STAGE] */

	const int strip_id = blockIdx.x * STRIP2_WARPS + warp;
	if (strip_id >= n_strips) return; /* whole warp leaves; only warp-level synchronisation below */
	const Strip st = strips[strip_id];
	const DevLayer L = layers[st.layer];
	const int frame = blockIdx.y;
	const int tx0 = L.begin_x + st.ix0, ty0 = L.begin_y + st.iy0;

	/* --- stage the strip's pixels as histogram bins (v >> 2, HistEq64Filter.cpp:14-25); zero outside the image --- */
	if (tmaps != nullptr && L.tma_ok) {
		const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_mbar[warp]);
		const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_tile);
		if (lane == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(Cfg::TILE_BYTES) : "memory");
			asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
					:: "r"(dst), "l"(reinterpret_cast<uint64_t>(tmaps + st.layer)), "r"(tx0), "r"(ty0), "r"(frame), "r"(bar) : "memory");
		}
		__syncwarp();
		uint32_t done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
					: "=r"(done) : "r"(bar) : "memory");
		}
		uint32_t* const tw32 = reinterpret_cast<uint32_t*>(s_tile);
		for (int i = lane; i < TR * (STRIP2_PITCH / 4); i += 32) tw32[i] = (tw32[i] >> 2) & 0x3f3f3f3fu;
	} else {
		const uint8_t* __restrict__ img = (L.offset < 0 ? frames + (int64_t)frame * W * H
				: arena + (int64_t)frame * arena_stride + L.offset);
		for (int r = 0; r < TR; ++r) {
			const bool row_ok = ty0 + r < L.height;
			const uint8_t* row = img + (int64_t)(ty0 + r) * L.pitch + tx0;
			for (int c = lane; c < STRIP2_PITCH; c += 32)
				s_tile[r * STRIP2_PITCH + c] = (row_ok && tx0 + c < L.width) ? (uint8_t)(row[c] >> 2) : (uint8_t)0;
		}
	}
	__syncwarp();

	/* --- roles --- */
	const int col = lane % st.cols, sub = lane / st.cols;
	const int iy_first = st.iy0 + sub * st.run;
	const bool valid = sub < st.nsub && iy_first < L.windows_y;
	const int nrows = valid ? min(st.run, L.windows_y - iy_first) : 0;
	const int maxrows = min(st.run, L.windows_y - st.iy0);
	const int org = valid ? (sub * st.run) * STRIP2_PITCH + col : 0; /* tile offset of this lane's first window */
	const int g = lane >> 2, t = lane & 3;    /* mma fragment coordinates: the quad g serves windows g + 8 i */
	const int gw = lane & 7, iw = lane >> 3;  /* as a window: this lane is window gw + 8 iw */
	int org_i[4];
#pragma unroll
	for (int i = 0; i < 4; ++i) org_i[i] = __shfl_sync(0xffffffffu, org, g + 8 * i);
	float* const s_hk = reinterpret_cast<float*>(s_lut) + STRIP2_HKU + lane; /* hk_kernel_eval[i]: s_hk[i * 32] */
	float* const s_u = s_hk + WVM_KA * 32;                                   /* u_kernel_eval[i]: s_u[i * 32] */
	/* LUT sharing: after a 4x4 byte transpose over the lanes gw, gw+8, gw+16, gw+24 this lane holds the word of
	 * bin 4 q + iw; it stores it for the 4 reader lanes 4 gw + t' with one 16-byte store */
	const uint32_t sel1 = (iw & 1) ? 0x3715u : 0x6240u, sel2 = (iw & 2) ? 0x3276u : 0x5410u;
	const uint32_t woff = (uint32_t)(iw * 128 + gw * 16); /* the 4 reader copies 4 gw + t' are 16 contiguous bytes */
	const uint8_t* const lut_r = s_lut + lane * 4;          /* reader view: + bin * 128 + i */
	const uint8_t* const lut_own = s_lut + (4 * gw) * 4 + iw; /* this window's own entries (copy t' = 0) */
	const float stretch = __fdiv_rn(255.0f, (float)(PW * PH)); /* HistEq64Filter.cpp:34 */

	for (int w = 0; w < maxrows; ++w) {
		__syncwarp();
		const bool active = w < nrows;
		const uint8_t* const tw = s_tile + org + w * STRIP2_PITCH; /* top-left bin of this lane's window */
		uint32_t total = 0, sxx = 0;
		if (active) {
			if (w == 0) { /* histogram of the first window of the run */
#pragma unroll
				for (int k = 0; k < 64; ++k) s_hist[k * 32] = 0;
				for (int r = 0; r < PH; ++r) {
#pragma unroll
					for (int c = 0; c < PW; ++c) atomicAdd(s_hist + tw[r * STRIP2_PITCH + c] * 32, 1u);
				}
			} else { /* slide down: row w-1 leaves, row w+PH-1 enters */
				const uint8_t* const r_out = tw - STRIP2_PITCH;
				const uint8_t* const r_in = tw + (PH - 1) * STRIP2_PITCH;
#pragma unroll
				for (int c = 0; c < PW; ++c) {
					atomicAdd(s_hist + r_out[c] * 32, 0xffffffffu); /* fire-and-forget: no load-add-store chain */
					atomicAdd(s_hist + r_in[c] * 32, 1u);
				}
			}
		}
		/* --- equalisation LUT (sequential float cumsum) of 4 bins at a time, shared with the 4 reader lanes at once.
		 * Rolled on purpose: the fully unrolled form made the kernel 100 KB of code and instruction-fetch bound --- */
		{
			float cdf = 0.f;
#pragma unroll 2
			for (int k = 0; k < 16; ++k) {
				uint32_t wq = 0;
				if (active) {
#pragma unroll
					for (int b = 0; b < 4; ++b) {
						const uint32_t cnt = s_hist[(4 * k + b) * 32];
						const uint32_t e = hq_step(cdf, cnt, stretch);
						wq |= e << (8 * b);
						total += cnt * e;
						sxx += cnt * e * e;
					}
				}
				const uint32_t x1 = __shfl_xor_sync(0xffffffffu, wq, 8);
				const uint32_t v1 = __byte_perm(wq, x1, sel1);
				const uint32_t x2 = __shfl_xor_sync(0xffffffffu, v1, 16);
				const uint32_t v2 = __byte_perm(v1, x2, sel2);
				*reinterpret_cast<uint4*>(s_lut + (4 * k) * 128 + woff) = make_uint4(v2, v2, v2, v2);
			}
		}
		__syncwarp();
		/* iimg_xx->data[dr]: float32 accumulation in row order (IImg.cpp:33-47); exact unless >= 2^24 */
		float sum_xx = (float)sxx;
		if (active && sxx >= (1u << 24)) {
			sum_xx = 0.f;
			for (int r = 0; r < PH; ++r) {
				uint32_t rowsq = 0;
#pragma unroll
				for (int c = 0; c < PW; ++c) {
					const uint32_t e = lut_own[tw[r * STRIP2_PITCH + c] * 128];
					rowsq += e * e;
				}
				sum_xx = r == 0 ? (float)rowsq : __fadd_rn(sum_xx, (float)rowsq);
			}
		}
		const float total_f = (float)total;

		/* --- D[window][8 filters x 4 values] on the tensor cores --- */
		int acc[2][4][4];
#pragma unroll
		for (int a = 0; a < 2; ++a)
#pragma unroll
			for (int b = 0; b < 4; ++b)
#pragma unroll
				for (int c = 0; c < 4; ++c) acc[a][b][c] = 0;
		const uint8_t* const trow = s_tile + w * STRIP2_PITCH;
#pragma unroll 2
		for (int s = 0; s < KS; ++s) {
			const uint4 bA = __ldg(m.bfrag + (s * 32 + lane) * 2), bB = __ldg(m.bfrag + (s * 32 + lane) * 2 + 1);
			/* tile offsets of the two patch words this lane feeds in this k-step: j0 = 8 s + t, j1 = j0 + 4 (0 past the patch:
			 * the B fragment is zero there) */
			const int j0 = 8 * s + t, j1 = j0 + 4;
			const int off2[2] = {j0 < NW ? (j0 / WPR) * STRIP2_PITCH + (j0 % WPR) * 4 : 0, j1 < NW ? (j1 / WPR) * STRIP2_PITCH + (j1 % WPR) * 4 : 0};
			uint32_t a[4][2];
#pragma unroll
			for (int i = 0; i < 4; ++i)
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					const uint8_t* p = trow + org_i[i] + off2[h];
					const uint32_t e0 = lut_r[p[0] * 128 + i], e1 = lut_r[p[1] * 128 + i];
					const uint32_t e2 = lut_r[p[2] * 128 + i], e3 = lut_r[p[3] * 128 + i];
					a[i][h] = e0 | (e1 << 8) | (e2 << 16) | (e3 << 24);
				}
			mma_u8(acc[0][0], a[0][0], a[1][0], a[0][1], a[1][1], bA.x, bA.y);
			mma_u8(acc[0][1], a[0][0], a[1][0], a[0][1], a[1][1], bA.z, bA.w);
			mma_u8(acc[0][2], a[0][0], a[1][0], a[0][1], a[1][1], bB.x, bB.y);
			mma_u8(acc[0][3], a[0][0], a[1][0], a[0][1], a[1][1], bB.z, bB.w);
			mma_u8(acc[1][0], a[2][0], a[3][0], a[2][1], a[3][1], bA.x, bA.y);
			mma_u8(acc[1][1], a[2][0], a[3][0], a[2][1], a[3][1], bA.z, bA.w);
			mma_u8(acc[1][2], a[2][0], a[3][0], a[2][1], a[3][1], bB.x, bB.y);
			mma_u8(acc[1][3], a[2][0], a[3][0], a[2][1], a[3][1], bB.z, bB.w);
		}
		__syncwarp(); /* every lane is done with the LUTs: the staging area may overwrite them */
#pragma unroll
		for (int mt = 0; mt < 2; ++mt)
#pragma unroll
			for (int nt = 0; nt < 4; ++nt)
#pragma unroll
				for (int h = 0; h < 2; ++h)
					*reinterpret_cast<int2*>(s_stage + (g + 8 * h + 16 * mt) * STRIP2_STAGE + 8 * nt + 2 * t) =
							make_int2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
		__syncwarp();

		/* --- scalar cascade over the first WVM_KA filters (WvmClassifier.cpp:129-138, 191-346); hk_kernel_eval and
		 * u_kernel_eval live in shared memory (one column per lane) so that the level loop stays rolled --- */
#pragma unroll
		for (int i = 0; i < WVM_KA; ++i) s_u[i * 32] = 0.f;                          /* :129-131 */
		int level = -1;
		float fout = 0.f;
		bool alive = active;
#pragma unroll 1
		for (int lv = 0; lv < WVM_KA; ++lv) {
			if (!__any_sync(0xffffffffu, alive)) break;
			if (alive) {
				level = lv;
				const int nv = __ldg(m.cntval + lv) - 1;
				const uint4 d4 = *reinterpret_cast<const uint4*>(s_stage + lane * STRIP2_STAGE + 4 * lv);
				const int n = lv % m.per_level;
				float un = s_u[n * 32];
				const float kv = wvm_kernel_value4(m, lv, d4.x, d4.y, d4.z, d4.w, nv, total_f, sum_xx, &un);
				s_u[n * 32] = un;
				s_hk[lv * 32] = kv;
				const float* __restrict__ wgt = m.hk_weights + lv * (lv + 1) / 2;
				float res = -__ldg(m.lin_thresholds + lv);                      /* :201 */
#pragma unroll 1
				for (int p = 0; p <= lv; ++p) res = __fadd_rn(res, __fmul_rn(__ldg(wgt + p), s_hk[p * 32])); /* :340-341 */
				fout = res;
				alive = fout >= __ldg(m.thresholds + lv) && lv + 1 < m.num_used;
			}
		}
		__syncwarp(); /* staging area consumed */
		if (active) {
			const int win = L.first_window + (iy_first + w) * L.windows_x + st.ix0 + col;
			if (alive) { /* survived every filter of this kernel: rebuild the LUT (own column) and hand the patch over */
				const int slot = atomicAdd(q.count, 1);
				if (slot < q.cap) {
					DeepRec r;
					r.frame = frame; r.window = win; r.total_f = total_f; r.sum_xx = sum_xx;
#pragma unroll
					for (int i = 0; i < WVM_KA; ++i) { r.hk[i] = s_hk[i * 32]; r.u[i] = s_u[i * 32]; }
					q.rec[slot] = r;
				}
				/* the LUT area (staging + cascade state) is free again: rebuild this window's own LUT column */
				uint32_t* const mylut = reinterpret_cast<uint32_t*>(s_lut) + lane; /* bin b: mylut[b * 32] */
				float cdf = 0.f;
#pragma unroll 4
				for (int k = 0; k < 64; ++k) mylut[k * 32] = hq_step(cdf, s_hist[k * 32], stretch);
				if (slot < q.cap) {
					for (int rr = 0; rr < PH; ++rr) {
#pragma unroll
						for (int k = 0; k < WPR; ++k) {
							const uint8_t* p = tw + rr * STRIP2_PITCH + 4 * k;
							const uint32_t e0 = mylut[p[0] * 32], e1 = mylut[p[1] * 32], e2 = mylut[p[2] * 32], e3 = mylut[p[3] * 32];
							q.patch[(size_t)(rr * WPR + k) * q.cap + slot] = e0 | (e1 << 8) | (e2 << 16) | (e3 << 24);
						}
					}
				}
				/* slot >= cap: counted in *q.count; the host re-runs the launch on the generic path */
			} else {
				wvm_emit(m, frame, win, windows_per_frame, level, fout, dense, cand, cand_count, cand_cap);
			}
		}
	}
}

template <int PW, int PH>
static cudaError_t mma_configure() {
	return cudaFuncSetAttribute(wvm_strip_mma_kernel<PW, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MmaCfg<PW, PH>::SMEM);
}

int strip_mma_configure_all() {
	cudaError_t e = mma_configure<20, 20>();
	if (e == cudaSuccess) e = mma_configure<24, 24>();
	if (e == cudaSuccess) e = mma_configure<32, 16>();
	if (e == cudaSuccess) e = mma_configure<32, 24>();
	if (e == cudaSuccess) e = mma_configure<16, 24>();
	return (int)e;
}

int strip_mma_ksteps(int pw, int ph) { return (pw * ph / 4 + 7) / 8; }

template <int PW, int PH>
static void mma_launch(cudaStream_t st, const DevWvm& m, const uint8_t* frames, int W, int H, int n_frames,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, const Strip* strips, int n_strips,
		int windows_per_frame, fdb_window_score* dense, Candidate* cand, int* cand_count, int cand_cap, const DeepQueue& q,
		const CUtensorMap* tmaps) {
	dim3 grid((unsigned)((n_strips + STRIP2_WARPS - 1) / STRIP2_WARPS), (unsigned)n_frames);
	wvm_strip_mma_kernel<PW, PH><<<grid, STRIP2_WARPS * 32, MmaCfg<PW, PH>::SMEM, st>>>(m, frames, W, H, arena, arena_stride, layers,
			strips, n_strips, windows_per_frame, dense, cand, cand_count, cand_cap, q, tmaps);
}

/* true when the launch was issued (supported patch size and the model carries a fragment table) */
bool launch_strip_mma(cudaStream_t st, const DevWvm& m, const uint8_t* frames, int W, int H, int n_frames,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, const Strip* strips, int n_strips,
		int windows_per_frame, fdb_window_score* dense, Candidate* cand, int* cand_count, int cand_cap, const DeepQueue& q,
		const void* tmaps_v) {
	if (m.bfrag == nullptr) return false;
	const CUtensorMap* tmaps = reinterpret_cast<const CUtensorMap*>(tmaps_v);
#define FDB_MMA_CASE(PW, PH) if (m.fsx == PW && m.fsy == PH) { mma_launch<PW, PH>(st, m, frames, W, H, n_frames, arena, arena_stride, \
		layers, strips, n_strips, windows_per_frame, dense, cand, cand_count, cand_cap, q, tmaps); return true; }
	FDB_MMA_CASE(20, 20) FDB_MMA_CASE(24, 24) FDB_MMA_CASE(32, 16) FDB_MMA_CASE(32, 24) FDB_MMA_CASE(16, 24)
#undef FDB_MMA_CASE
	return false;
}

} // namespace fdb
