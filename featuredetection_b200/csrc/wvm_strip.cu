/*
 * wvm_strip.cu - the fast path of stage 1 (sm_100a): column-strip window kernel + warp-per-window
 * deep kernel.  Same arithmetic as wvm.cu (which stays as the generic fallback for unusual patch
 * sizes, window steps != 1, ROI scans and direct feature-vector classification); see wvm.cu and
 * DESIGN.md for the reference citations and the exactness argument.
 *
 * wvm_strip_kernel<PW, PH>
 *   A warp owns a strip of the window grid of one pyramid layer of one frame: up to 32 adjacent
 *   window columns (narrow layers pack several row runs side by side) by WVM_RUN window rows.  The
 *   warp stages the strip's pixels once as 6-bit histogram bins in shared memory; every lane then
 *   walks DOWN its column: the 64-bin histogram of the window below differs by one pixel row
 *   leaving and one entering (2*PW updates instead of PW*PH).  Per window: sequential float
 *   cumsum -> 64-entry LUT (+ sum(x), sum(x^2) from the histogram), LUT application into PW*PH/4
 *   REGISTERS (4 pixels per register), then up to WVM_KA filters as dp4a dot products against the
 *   rectangle-coverage masks.  Survivors of all WVM_KA filters go to the deep queue.
 *   Shared memory per lane: 64 u16 histogram counts + 64 u16 LUT entries, column layout [bin][lane]
 *   (two lanes per 32-bit word) so that every access is bank-conflict free.
 *
 * wvm_deep_warp_kernel
 *   One WARP per queued window.  The 32 lanes split the patch words; 16 filters are evaluated per
 *   round: per-lane partial dp4a sums, a recursive-halving reduce-scatter over the lanes (62
 *   shuffles per round), the double-precision kernel values on 16 owner lanes (in wavelet-level
 *   order where filters share u_kernel_eval), the float weighted sums as 16 independent sequential
 *   chains, then a ballot finds the first filter that rejects.  Filters past the rejecting one are
 *   computed speculatively and discarded, so results equal the sequential cascade exactly.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "fdb_internal.h"
#include "wvm_device.h"
#include "wvm_math.cuh"

namespace fdb {

#define STRIP_T 128          /* threads per CTA (4 warps, one strip each) */
#ifndef STRIP_MIN_CTAS
#define STRIP_MIN_CTAS 3
#endif
#define STRIP_TILE_PITCH 64  /* bytes per tile row: 32 columns + PW - 1 <= 63 */

template <int PW, int PH>
struct StripCfg {
	static constexpr int NW = PW * PH / 4;
	static constexpr int TILE_ROWS = STRIP_TILE_ROWS; /* even: each warp's tile is a multiple of 128 bytes (TMA destination) */
	static constexpr size_t SMEM = (size_t)(64 + 64) * STRIP_T * 2 + (size_t)4 * TILE_ROWS * STRIP_TILE_PITCH;
};

template <int PW, int PH>
__global__ void __launch_bounds__(STRIP_T, (PW * PH <= 416 ? STRIP_MIN_CTAS : (PW * PH <= 600 ? 2 : 1))) wvm_strip_kernel(const DevWvm m,
		const uint8_t* __restrict__ frames, int W, int H,
		const uint8_t* __restrict__ arena, int64_t arena_stride,
		const DevLayer* __restrict__ layers, const Strip* __restrict__ strips, int n_strips, int windows_per_frame,
		fdb_window_score* __restrict__ dense,
		Candidate* __restrict__ cand, int* __restrict__ cand_count, int cand_cap, const DeepQueue q,
		const CUtensorMap* __restrict__ tmaps) {
	static_assert(PW % 4 == 0 && PW <= 32, "patch width must be a multiple of 4, at most 32");
	constexpr int NW = StripCfg<PW, PH>::NW;
	constexpr int WPR = PW / 4; /* words per patch row */
	constexpr int T = STRIP_T;
	extern __shared__ __align__(1024) uint32_t smem[];
	__shared__ __align__(8) uint64_t s_mbar[4];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint16_t* const s_hist = reinterpret_cast<uint16_t*>(smem) + tid;          /* [64][T] u16 counts -> bin b at s_hist[b*T] */
	uint16_t* const s_lut = reinterpret_cast<uint16_t*>(smem) + 64 * T + tid;  /* [64][T] u16 equalised values */
	uint8_t* const s_tile = reinterpret_cast<uint8_t*>(smem) + 128 * T * 2 + warp * StripCfg<PW, PH>::TILE_ROWS * STRIP_TILE_PITCH;

	const int strip_id = blockIdx.x * 4 + warp;
	if (strip_id >= n_strips) return; /* whole warp leaves; only __syncwarp is used below */
	const Strip st = strips[strip_id];
	const DevLayer L = layers[st.layer];
	const int frame = blockIdx.y;
	const uint8_t* __restrict__ img = (L.offset < 0 ? frames + (int64_t)frame * W * H
			: arena + (int64_t)frame * arena_stride + L.offset);
	/* --- stage the strip's pixels as histogram bins (v >> 2, HistEq64Filter.cpp:14-25) --- */
	const int tx0 = L.begin_x + st.ix0, ty0 = L.begin_y + st.iy0;
	const int tcols = min(st.cols + PW - 1, L.width - tx0);
	const int trows = min(st.nsub * st.run + PH - 1, L.height - ty0);
	if (tmaps != nullptr && L.tma_ok) {
		/* TMA: one bulk tensor copy (64 bytes x TILE_ROWS rows of the layer, zero-filled outside the image)
		 * lands the tile in shared memory and signals the warp's mbarrier */
		const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_mbar[warp]);
		const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_tile);
		if (lane == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(StripCfg<PW, PH>::TILE_ROWS * STRIP_TILE_PITCH) : "memory");
			asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
					:: "r"(dst), "l"(reinterpret_cast<uint64_t>(tmaps + st.layer)), "r"(tx0), "r"(ty0), "r"(frame), "r"(bar) : "memory");
		}
		__syncwarp();
		uint32_t done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
					: "=r"(done) : "r"(bar) : "memory");
		}
		/* pixels -> bins, in place, 4 per word */
		uint32_t* const tw32 = reinterpret_cast<uint32_t*>(s_tile);
		for (int i = lane; i < trows * (STRIP_TILE_PITCH / 4); i += 32) tw32[i] = (tw32[i] >> 2) & 0x3f3f3f3fu;
	} else {
		for (int r = 0; r < trows; ++r) {
			const uint8_t* row = img + (int64_t)(ty0 + r) * L.pitch + tx0;
			for (int c = lane; c < tcols; c += 32) s_tile[r * STRIP_TILE_PITCH + c] = row[c] >> 2;
		}
	}
	__syncwarp();
	const int col = lane % st.cols, sub = lane / st.cols;
	const int iy_first = st.iy0 + sub * st.run;
	if (sub >= st.nsub || iy_first >= L.windows_y) return;
	const int nrows = min(st.run, L.windows_y - iy_first);
	const uint8_t* const tcol = s_tile + (sub * st.run) * STRIP_TILE_PITCH + col;
	const float stretch = __fdiv_rn(255.0f, (float)(PW * PH)); /* HistEq64Filter.cpp:34 */

	/* histogram of the first window of the run */
#pragma unroll
	for (int k = 0; k < 64; ++k) s_hist[k * T] = 0;
	for (int r = 0; r < PH; ++r) {
#pragma unroll
		for (int c = 0; c < PW; ++c) s_hist[tcol[r * STRIP_TILE_PITCH + c] * T] += 1;
	}

	for (int w = 0; w < nrows; ++w) {
		const uint8_t* const tw = tcol + w * STRIP_TILE_PITCH; /* top-left bin of this window */
		if (w > 0) { /* slide down: row w-1 leaves, row w+PH-1 enters */
			const uint8_t* const r_out = tw - STRIP_TILE_PITCH;
			const uint8_t* const r_in = tw + (PH - 1) * STRIP_TILE_PITCH;
#pragma unroll
			for (int c = 0; c < PW; ++c) {
				s_hist[r_out[c] * T] -= 1;
				s_hist[r_in[c] * T] += 1;
			}
		}
		/* --- sequential float cumsum -> LUT (HistEq64Filter.cpp:70-87,97); sums from the histogram --- */
		float cdf = 0.f;
		uint32_t total = 0, sxx = 0;
#pragma unroll 16
		for (int k = 0; k < 64; ++k) {
			const uint32_t cnt = s_hist[k * T];
			cdf = __fadd_rn(cdf, __fmul_rn((float)cnt, stretch));
			const float fl = floorf(cdf); /* (uchar)floor((double)cdf + 0.5) == floor(cdf) + (frac >= 0.5) */
			const uint32_t e = ((uint32_t)(int)fl + (__fsub_rn(cdf, fl) >= 0.5f ? 1u : 0u)) & 255u;
			s_lut[k * T] = (uint16_t)e;
			total += cnt * e;
			sxx += cnt * e * e;
		}
		/* --- equalised patch into registers, 4 pixels per word --- */
		uint32_t x[NW];
#pragma unroll
		for (int r = 0; r < PH; ++r) {
#pragma unroll
			for (int k = 0; k < WPR; ++k) {
				const uint8_t* p = tw + r * STRIP_TILE_PITCH + 4 * k;
				const uint32_t e0 = s_lut[p[0] * T], e1 = s_lut[p[1] * T], e2 = s_lut[p[2] * T], e3 = s_lut[p[3] * T];
				x[r * WPR + k] = e0 | (e1 << 8) | (e2 << 16) | (e3 << 24);
			}
		}
		/* iimg_xx->data[dr]: float32 accumulation in row order (IImg.cpp:33-47); exact unless >= 2^24 */
		float sum_xx;
		if (sxx < (1u << 24)) {
			sum_xx = (float)sxx;
		} else {
			sum_xx = 0.f;
#pragma unroll
			for (int r = 0; r < PH; ++r) {
				uint32_t rowsq = 0;
#pragma unroll
				for (int k = 0; k < WPR; ++k) rowsq = __dp4a(x[r * WPR + k], x[r * WPR + k], rowsq);
				sum_xx = r == 0 ? (float)rowsq : __fadd_rn(sum_xx, (float)rowsq);
			}
		}
		const float total_f = (float)total;

		/* --- first WVM_KA filters (WvmClassifier.cpp:129-138, 191-346) --- */
		float hk[WVM_KA], u[WVM_KA];
#pragma unroll
		for (int i = 0; i < WVM_KA; ++i) { hk[i] = 0.f; u[i] = 0.f; }
		int level = -1;
		float fout = 0.f;
		bool alive = true;
#pragma unroll 1
		for (int lv = 0; lv < WVM_KA && alive; ++lv) {
			level = lv;
			const int nv = __ldg(m.cntval + lv) - 1;
			const uint4* __restrict__ mk4 = reinterpret_cast<const uint4*>(m.masks4) + (size_t)lv * NW;
			uint32_t acc[FDB_MAX_VALUES];
#pragma unroll
			for (int v = 0; v < FDB_MAX_VALUES; ++v) acc[v] = 0;
#pragma unroll
			for (int j = 0; j < NW; ++j) { /* one 16-byte load (uniform address) brings the four masks of a word */
				const uint4 k4 = __ldg(mk4 + j);
				acc[0] = __dp4a(x[j], k4.x, acc[0]); acc[1] = __dp4a(x[j], k4.y, acc[1]);
				acc[2] = __dp4a(x[j], k4.z, acc[2]); acc[3] = __dp4a(x[j], k4.w, acc[3]);
			}
			const int n = lv % m.per_level;
			float un = 0.f;
#pragma unroll
			for (int i = 0; i < WVM_KA; ++i) if (i == n) un = u[i];
			const float kv = wvm_kernel_value(m, lv, acc, nv, total_f, sum_xx, &un);
#pragma unroll
			for (int i = 0; i < WVM_KA; ++i) { if (i == n) u[i] = un; if (i == lv) hk[i] = kv; }
			const float* __restrict__ wgt = m.hk_weights + lv * (lv + 1) / 2;
			float res = -__ldg(m.lin_thresholds + lv);                      /* :201 */
#pragma unroll
			for (int p = 0; p < WVM_KA; ++p)                                /* :340-341 */
				if (p <= lv) res = __fadd_rn(res, __fmul_rn(__ldg(wgt + p), hk[p]));
			fout = res;
			alive = fout >= __ldg(m.thresholds + lv) && lv + 1 < m.num_used;
		}
		const int win = L.first_window + (iy_first + w) * L.windows_x + st.ix0 + col;
		if (alive) { /* survived every filter of this kernel: hand over (the queue always has room, see api.cu) */
			const int slot = atomicAdd(q.count, 1);
			if (slot < q.cap) {
				DeepRec r;
				r.frame = frame; r.window = win; r.total_f = total_f; r.sum_xx = sum_xx;
#pragma unroll
				for (int i = 0; i < WVM_KA; ++i) { r.hk[i] = hk[i]; r.u[i] = u[i]; }
				q.rec[slot] = r;
#pragma unroll
				for (int j = 0; j < NW; ++j) q.patch[(size_t)j * q.cap + slot] = x[j];
			}
			/* slot >= cap: counted in *q.count; the host re-runs the launch on the generic path */
		} else {
			wvm_emit(m, frame, win, windows_per_frame, level, fout, dense, cand, cand_count, cand_cap);
		}
	}
}

/* ---------------------------------------------------------------------------------------------
 * deep kernel: one warp per queued window, 32 filters per round, rectangle sums from an integral
 * image of the equalised patch (the reference's own evaluation scheme, WvmClassifier.cpp:277-306;
 * every entry is an exact integer < 2^24, so int32 arithmetic reproduces the float sums bit for bit)
 * ------------------------------------------------------------------------------------------- */
#define DEEP_WARPS 4
#define DEEP_IIMG 1024   /* (w + 1) * (h + 1) <= 1024 ints per warp (32x24 -> 825) */

__global__ void __launch_bounds__(DEEP_WARPS * 32) wvm_deep_warp_kernel(const DevWvm m, const DeepQueue q, int windows_per_frame,
		fdb_window_score* __restrict__ dense, Candidate* __restrict__ cand, int* __restrict__ cand_count, int cand_cap) {
	__shared__ __align__(16) float s_hk[DEEP_WARPS][FDB_MAX_FILTERS];
	__shared__ float s_u[DEEP_WARPS][FDB_MAX_PER_LEVEL];
	__shared__ int s_ii[DEEP_WARPS][DEEP_IIMG];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	float* const hk = s_hk[warp];
	float* const us = s_u[warp];
	int* const ii = s_ii[warp];
	const int n = min(*q.count, q.cap);
	const int pw = m.fsx, ph = m.fsy, pitch = pw + 1;
	const int wpr = (pw + 3) >> 2; /* patch words per row (pw is a multiple of 4 on this path) */
	for (;;) {
		/* dynamic work distribution: windows differ by 50x in cost (first-round exits vs. full depth) */
		int slot = 0;
		if (lane == 0) slot = atomicAdd(q.next, 1);
		slot = __shfl_sync(0xffffffffu, slot, 0);
		if (slot >= n) break;
		const DeepRec rec = q.rec[slot];
		/* --- integral image with a zero first row and column: ii[(y+1)*pitch + x+1] = sum of x[0..y][0..x] --- */
		for (int i = lane; i < pitch; i += 32) ii[i] = 0;
		int colsum = 0; /* lane = column */
		for (int r = 0; r < ph; ++r) {
			uint32_t v = 0;
			if (lane < pw) v = (__ldg(q.patch + (size_t)(r * wpr + (lane >> 2)) * q.cap + slot) >> ((lane & 3) * 8)) & 255u;
			/* inclusive prefix over the row */
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
				if (lane >= d) v += t;
			}
			/* column-wise accumulation of the row prefixes: ii[r+1][c+1] = ii[r][c+1] + prefix(r, c) */
			colsum += (int)v;
			if (lane < pw) ii[(r + 1) * pitch + lane + 1] = colsum;
			if (lane == 0) ii[(r + 1) * pitch] = 0;
		}
		for (int i = lane; i < m.per_level; i += 32) us[i] = 0.f;
		__syncwarp();
		if (lane < WVM_KA) { hk[lane] = rec.hk[lane]; if (lane < m.per_level) us[lane] = rec.u[lane]; }
		__syncwarp();
		int final_level = -1;
		float final_fout = 0.f;
		for (int base = WVM_KA; base < m.num_used && final_level < 0; base += 32) {
			const int cnt = min(32, m.num_used - base);
			const int level = base + lane;
			const bool owner = lane < cnt;
			/* rectangle sums of the filter this lane owns */
			uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
			int nv = 0;
			if (owner) {
				nv = __ldg(m.cntval + level) - 1;
				const int r0 = __ldg(m.rect_off + level), r1 = __ldg(m.rect_off + level + 1);
				for (int r = r0; r < r1; ++r) {
					const uint2 rc = __ldg(m.rects + r); /* {x1 | y1 << 8 | x2 << 16 | y2 << 24, grey value index} */
					const int x1 = rc.x & 255, y1 = (rc.x >> 8) & 255, x2 = (rc.x >> 16) & 255, y2 = rc.x >> 24;
					const int sum = ii[(y2 + 1) * pitch + x2 + 1] - ii[y1 * pitch + x2 + 1] - ii[(y2 + 1) * pitch + x1] + ii[y1 * pitch + x1];
					s0 += rc.y == 0 ? (uint32_t)sum : 0u; s1 += rc.y == 1 ? (uint32_t)sum : 0u;
					s2 += rc.y == 2 ? (uint32_t)sum : 0u; s3 += rc.y == 3 ? (uint32_t)sum : 0u;
				}
			}
			/* kernel values; filters sharing u_kernel_eval go in wavelet-level order */
			const int rounds = (cnt + m.per_level - 1) / m.per_level;
			for (int r = 0; r < rounds; ++r) {
				if (owner && lane / m.per_level == r) {
					float un = us[level % m.per_level];
					const float kv = wvm_kernel_value4(m, level, s0, s1, s2, s3, nv, rec.total_f, rec.sum_xx, &un);
					us[level % m.per_level] = un;
					hk[level] = kv;
				}
				__syncwarp();
			}
			/* float weighted sums (WvmClassifier.cpp:340-341): one sequential chain per owner lane; weights come
			 * four at a time from the 16-byte aligned row copy, kernel values four at a time from shared memory */
			float res = 0.f;
			bool pass = true;
			if (owner) {
				const float* __restrict__ wrow = m.hk_weights4 + __ldg(m.hk_row4 + level);
				const float4* __restrict__ w4 = reinterpret_cast<const float4*>(wrow);
				const float4* h4 = reinterpret_cast<const float4*>(hk);
				res = -__ldg(m.lin_thresholds + level);
				const int groups = (level + 1) >> 2;
				int g = 0;
				for (; g + 2 <= groups; g += 2) {
					const float4 wa = __ldg(w4 + g), wb = __ldg(w4 + g + 1);
					const float4 ha = h4[g], hb = h4[g + 1];
					res = __fadd_rn(res, __fmul_rn(wa.x, ha.x)); res = __fadd_rn(res, __fmul_rn(wa.y, ha.y));
					res = __fadd_rn(res, __fmul_rn(wa.z, ha.z)); res = __fadd_rn(res, __fmul_rn(wa.w, ha.w));
					res = __fadd_rn(res, __fmul_rn(wb.x, hb.x)); res = __fadd_rn(res, __fmul_rn(wb.y, hb.y));
					res = __fadd_rn(res, __fmul_rn(wb.z, hb.z)); res = __fadd_rn(res, __fmul_rn(wb.w, hb.w));
				}
				for (int p = 4 * g; p <= level; ++p) res = __fadd_rn(res, __fmul_rn(__ldg(wrow + p), hk[p]));
				pass = res >= __ldg(m.thresholds + level) && level + 1 < m.num_used;
			}
			/* the cascade stops at the first rejecting filter; later ones were speculative */
			const unsigned fails = __ballot_sync(0xffffffffu, owner && !pass);
			if (fails) {
				const int src = __ffs(fails) - 1;
				final_level = base + src;
				final_fout = __shfl_sync(0xffffffffu, res, src);
			}
			__syncwarp();
		}
		if (lane == 0)
			wvm_emit(m, rec.frame, rec.window, windows_per_frame, final_level, final_fout, dense, cand, cand_count, cand_cap);
		__syncwarp();
	}
}

/* ---------------------------------------------------------------------------------------------
 * launchers
 * ------------------------------------------------------------------------------------------- */
template <int PW, int PH>
static cudaError_t strip_configure() {
	return cudaFuncSetAttribute(wvm_strip_kernel<PW, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StripCfg<PW, PH>::SMEM);
}

int strip_configure_all() {
	cudaError_t e = strip_configure<20, 20>();
	if (e == cudaSuccess) e = strip_configure<24, 24>();
	if (e == cudaSuccess) e = strip_configure<32, 16>();
	if (e == cudaSuccess) e = strip_configure<32, 24>();
	if (e == cudaSuccess) e = strip_configure<16, 24>();
	return (int)e;
}

bool strip_supported(int pw, int ph) {
	return (pw == 20 && ph == 20) || (pw == 24 && ph == 24) || (pw == 32 && ph == 16) || (pw == 32 && ph == 24) || (pw == 16 && ph == 24);
}

template <int PW, int PH>
static void strip_launch(cudaStream_t st, const DevWvm& m, const uint8_t* frames, int W, int H, int n_frames,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, const Strip* strips, int n_strips,
		int windows_per_frame, fdb_window_score* dense, Candidate* cand, int* cand_count, int cand_cap, const DeepQueue& q,
		const CUtensorMap* tmaps) {
	dim3 grid((unsigned)((n_strips + 3) / 4), (unsigned)n_frames);
	wvm_strip_kernel<PW, PH><<<grid, STRIP_T, StripCfg<PW, PH>::SMEM, st>>>(m, frames, W, H, arena, arena_stride, layers, strips,
			n_strips, windows_per_frame, dense, cand, cand_count, cand_cap, q, tmaps);
}

void launch_wvm_strips(cudaStream_t st, const DevWvm& m, const uint8_t* frames, int W, int H, int n_frames,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, const Strip* strips, int n_strips,
		int windows_per_frame, fdb_window_score* dense, Candidate* cand, int* cand_count, int cand_cap, const DeepQueue& q,
		cudaEvent_t ev_mid, const void* tmaps_v) {
	if (n_strips == 0 || n_frames == 0) return;
	const CUtensorMap* tmaps = reinterpret_cast<const CUtensorMap*>(tmaps_v);
	static const bool use_mma = []{ const char* e = std::getenv("FDB_STRIP_V1"); return !(e && e[0] == '1'); }();
	if (!(use_mma && launch_strip_mma(st, m, frames, W, H, n_frames, arena, arena_stride, layers, strips, n_strips,
			windows_per_frame, dense, cand, cand_count, cand_cap, q, tmaps_v))) {
#define FDB_STRIP_CASE(PW, PH) if (m.fsx == PW && m.fsy == PH) { strip_launch<PW, PH>(st, m, frames, W, H, n_frames, arena, arena_stride, \
		layers, strips, n_strips, windows_per_frame, dense, cand, cand_count, cand_cap, q, tmaps); }
	FDB_STRIP_CASE(20, 20) else FDB_STRIP_CASE(24, 24) else FDB_STRIP_CASE(32, 16) else FDB_STRIP_CASE(32, 24) else FDB_STRIP_CASE(16, 24)
#undef FDB_STRIP_CASE
	}
	if (ev_mid) cudaEventRecord(ev_mid, st); /* profiling mark between the two kernels */
	const int blocks = std::min((q.cap + DEEP_WARPS - 1) / DEEP_WARPS, 148 * 8);
	wvm_deep_warp_kernel<<<blocks, DEEP_WARPS * 32, 0, st>>>(m, q, windows_per_frame, dense, cand, cand_count, cand_cap);
}

} // namespace fdb

namespace fdb {
int strip_tile_rows(int patch_h) { (void)patch_h; return STRIP_TILE_ROWS; }
int strip_tile_pitch() { return STRIP_TILE_PITCH; }
}
