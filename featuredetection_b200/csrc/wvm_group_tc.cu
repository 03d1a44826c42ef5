/*
 * wvm_group_tc.cu - the window kernel of stage 1 on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in
 * tensor memory), sm_100a only. Same work items, same arithmetic and same results as wvm_group_kernel (wvm_group.cu); what
 * changes is who holds what:
 *
 *   - a CTA = 4 producer warps + 1 issue warp; it takes one work item (a strip of <= 32 window columns of one pyramid image
 *     and a pack of <= GRP_MAX_PACK models) for FOUR consecutive frames, one frame per producer warp, so the four warps walk
 *     the same window rows in step and their 4 x 32 windows are the 128 rows of one UMMA tile;
 *   - producers (lane = window): sliding 64-bin histogram, sequential float32 cumulative histogram -> equalisation table,
 *     then per k-step (one patch row padded to 32 bytes, two rows for 16-wide windows) the equalised pixels of the lane's
 *     window go to the A slab in shared memory as UMMA core matrices (8 rows x 16 bytes); sum(x^2) accumulates in the
 *     reference's float32 order (IImg.cpp:33-47);
 *   - the issue warp (one lane) streams the models' rectangle-coverage slabs (B, 1 KB per model and k-step, laid out on the
 *     host as core matrices) from L2 with cp.async.bulk into a ring, and per k-step issues ONE
 *     tcgen05.mma M128 N(32 x models) K32: D[window][model, filter, grey value] += A[window][pixel] . B[pixel][...].
 *     The accumulators of every model of the pack live in tensor memory - no accumulator registers, no B fragments through
 *     the load/store pipe, no ldmatrix: the mma.sync kernel was bound by that pipe (85 % of its wavefront peak, a quarter of
 *     it B fragments), and its register accumulators limited a pack to two models;
 *   - after the last k-step the producers read their own row of D straight from tensor memory (tcgen05.ld 32x32b: lane =
 *     window, 4 columns = the grey-value sums of one filter) and run the scalar cascade tail per model as before
 *     (WvmClassifier.cpp:129-138,191-346); survivors of all WVM_KA filters are queued for the deep kernel.
 *
 * Measured on the B200 (profiles/NOTES_r2.md): packs of 4 + 3 detectors of one 24 x 24 grid take 6.8 ms per 16-frame chunk
 * against 9.2 ms as 2 + 2 + 2 + 1 on the mma.sync kernel; for packs of 1-2 the mma.sync kernel is faster, so launch_wvm_group
 * sends only packs of 3-4 here. 4 CTAs per SM (55.8 KB shared memory, 96 registers, 128 tensor-memory columns each) beat 3 by
 * 16 %; deeper slab / B rings changed nothing - a producer's waiting time is the tile's slowest warp, whose cascade tail is
 * longer. Two rewrites of the tail that looked better on paper measured slower and were dropped: compacting the surviving
 * (window, model) pairs of a row into full passes (45 % fewer filter evaluations, but gathering sums and state across lanes
 * lengthens the dependency chain of each pass), and interleaving two models' cascades in straight-line code.
 *
 * Exactness: u8 x u8 -> s32 with sums < 2^24, identical to the integral-image rectangle sums of the reference.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>

#include "fdb_internal.h"
#include "wvm_device.h"
#include "wvm_group.h"
#include "wvm_group_dev.cuh"
#include "wvm_math.cuh"

namespace fdb {

namespace {

#ifndef GTC_NA_SMALL_
#define GTC_NA_SMALL_ 3   /* A slabs in flight, packs of 1 and 2 */
#endif
#ifndef GTC_NA_BIG_
#define GTC_NA_BIG_ 2     /* packs of 3 and 4 */
#endif
#ifndef GTC_CTAS_SMALL_
#define GTC_CTAS_SMALL_ 4 /* CTAs per SM, packs of 1 and 2 */
#endif
#ifndef GTC_CTAS_BIG_
#define GTC_CTAS_BIG_ 4   /* packs of 3 and 4 */
#endif
#ifndef GTC_NB_SMALL_
#define GTC_NB_SMALL_ 4   /* B ring stages */
#endif
#ifndef GTC_NB_BIG_
#define GTC_NB_BIG_ 3
#endif
#ifndef GTC_B_AHEAD_
#define GTC_B_AHEAD_ 2
#endif
#ifndef GTC_SLEEP_ISSUE_
#define GTC_SLEEP_ISSUE_ 32   /* ns between polls of the issue lane (it shares a scheduler with three producer warps) */
#endif
#ifndef GTC_SLEEP_PROD_
#define GTC_SLEEP_PROD_ 0
#endif
__host__ __device__ constexpr int gtc_na(int msub) { return msub <= 2 ? GTC_NA_SMALL_ : GTC_NA_BIG_; } /* A slabs in flight */
constexpr int GTC_NA_MAX = GTC_NA_SMALL_ > GTC_NA_BIG_ ? GTC_NA_SMALL_ : GTC_NA_BIG_;
__host__ __device__ constexpr int gtc_nb(int msub) { return msub <= 2 ? GTC_NB_SMALL_ : GTC_NB_BIG_; } /* B ring stages */
constexpr int GTC_NB_MAX = GTC_NB_SMALL_ > GTC_NB_BIG_ ? GTC_NB_SMALL_ : GTC_NB_BIG_;
constexpr int GTC_B_AHEAD = GTC_B_AHEAD_;      /* k-steps the B stream runs ahead of the MMA issue (an L2 -> shared bulk copy takes longer than a k-step) */
static_assert(GTC_B_AHEAD < GTC_NB_SMALL_ && GTC_B_AHEAD < GTC_NB_BIG_, "the B stream cannot run further ahead than the ring is deep");
constexpr int GTC_SLAB = 128 * 32;             /* A slab: 128 windows x 32 bytes */
constexpr int GTC_THREADS = 160;
static_assert(GRP_LUT_BYTES == GRP_HKU_BYTES, "the cascade tail's hk / u columns overlay the equalisation table (dead after the last k-step of a row)");
constexpr int GTC_WARP_BYTES = GRP_TILE_BYTES + GRP_HIST_BYTES + GRP_LUT_BYTES;
constexpr int GTC_SPIN_LIMIT = 1 << 26;
static_assert(GTC_WARP_BYTES % 128 == 0, "TMA destinations are 128-byte aligned");

enum { GB_TILE = 0, GB_A_FULL = 4, GB_A_EMPTY = GB_A_FULL + GTC_NA_MAX, GB_B_FULL = GB_A_EMPTY + GTC_NA_MAX, GB_B_EMPTY = GB_B_FULL + GTC_NB_MAX,
	GB_ACC_FULL = GB_B_EMPTY + GTC_NB_MAX, GB_ACC_EMPTY, GB_UNIT, GB_COUNT };

__host__ __device__ constexpr int gtc_b_stage(int msub) { return msub * 1024; }
__host__ __device__ constexpr int gtc_smem(int msub) { return 4 * GTC_WARP_BYTES + gtc_na(msub) * GTC_SLAB + gtc_nb(msub) * gtc_b_stage(msub) + GB_COUNT * 8 + 32; }
__host__ __device__ constexpr int gtc_tmem_cols(int msub) { return msub <= 1 ? 32 : (msub == 2 ? 64 : (msub <= 4 ? 128 : 256)); }
__host__ __device__ constexpr int gtc_ctas_per_sm(int msub) { return msub <= 2 ? GTC_CTAS_SMALL_ : GTC_CTAS_BIG_; }

__device__ __forceinline__ uint32_t gtc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gtc_mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void gtc_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void gtc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
/* bounded wait: a protocol error traps instead of hanging the device. SLEEP_NS > 0 backs off between polls: try_wait returns at
 * once on this part, and a polling warp takes issue slots from the warps that have work */
template <int SLEEP_NS = 0>
__device__ __forceinline__ void gtc_mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t done;
	int spins = 0;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
				: "=r"(done) : "r"(bar), "r"(parity) : "memory");
		if (!done) {
			if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
			if (++spins > GTC_SPIN_LIMIT) {
				printf("wvm_group_tc_kernel: barrier %u timed out (block %d thread %d)\n", bar, blockIdx.x, threadIdx.x);
				__trap();
			}
		}
	} while (!done);
}
__device__ __forceinline__ void gtc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gtc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gtc_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
/* D[tmem] (+)= A[smem] . B[smem]^T, u8 x u8 -> s32 */
__device__ __forceinline__ void gtc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
			:: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
/* shared-memory matrix descriptor: K-major, no swizzle, version 1 (sm_100); core matrix = 8 rows x 16 bytes = 128 contiguous
 * bytes, `lbo` bytes between core matrices adjacent in K, `sbo` bytes between groups of 8 rows */
__device__ __forceinline__ uint64_t gtc_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
	return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32)
			| ((uint64_t)1 << 46);
}
/* this lane's row, 4 consecutive accumulator columns */
__device__ __forceinline__ uint4 gtc_ld4(uint32_t taddr) {
	uint4 v;
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
			: "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(taddr) : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
	return v;
}

template <int PW, int PH, int MSUB>
__global__ void __launch_bounds__(GTC_THREADS, gtc_ctas_per_sm(MSUB)) wvm_group_tc_kernel(const __grid_constant__ GroupArgs a) {
	static_assert(PW % 4 == 0 && PW >= 16 && PW <= 32, "window width: a multiple of 4 in 16..32");
	static_assert(MSUB >= 1 && MSUB <= GRP_MAX_PACK, "pack size");
	constexpr int WPR = PW / 4;                 /* words per patch row */
	constexpr int RPK = PW <= 16 ? 2 : 1;       /* patch rows per k-step (32 operand bytes) */
	static_assert(PH % RPK == 0, "window height must split into k-steps");
	static_assert(grp_stretch_is_safe(PW * PH), "grp_hq_step's shortcut does not hold for this window size");
	constexpr int KS = PH / RPK;
	constexpr int BSTAGE = gtc_b_stage(MSUB);
	constexpr int TCOLS = gtc_tmem_cols(MSUB);
	constexpr int GTC_NA = gtc_na(MSUB), GTC_NB = gtc_nb(MSUB);
	extern __shared__ __align__(128) uint8_t smem8[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint8_t* const s_slabs = smem8 + 4 * GTC_WARP_BYTES;
	uint8_t* const s_bring = s_slabs + GTC_NA * GTC_SLAB;
	uint64_t* const s_bars = reinterpret_cast<uint64_t*>(s_bring + GTC_NB * BSTAGE);
	uint32_t* const s_tmem = reinterpret_cast<uint32_t*>(s_bars + GB_COUNT);
	volatile int* const s_unit = reinterpret_cast<volatile int*>(s_tmem + 1); /* [2] */
	const uint32_t bar0 = gtc_smem_u32(s_bars);
	auto bar = [bar0](int i) { return bar0 + 8u * (uint32_t)i; };

	if (tid == 0) {
		for (int i = 0; i < 4; ++i) gtc_mbar_init(bar(GB_TILE + i), 1);
		for (int i = 0; i < GTC_NA; ++i) { gtc_mbar_init(bar(GB_A_FULL + i), 4); gtc_mbar_init(bar(GB_A_EMPTY + i), 1); }
		for (int i = 0; i < GTC_NB; ++i) { gtc_mbar_init(bar(GB_B_FULL + i), 1); gtc_mbar_init(bar(GB_B_EMPTY + i), 1); }
		gtc_mbar_init(bar(GB_ACC_FULL), 1); gtc_mbar_init(bar(GB_ACC_EMPTY), 4); gtc_mbar_init(bar(GB_UNIT), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 4) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(gtc_smem_u32(s_tmem)), "n"(TCOLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	gtc_fence_before();
	__syncthreads();
	gtc_fence_after();
	const uint32_t tmem = *s_tmem;

	const int quads = (a.n_frames + 3) >> 2;
	const int total_units = a.n_items * quads;
	uint32_t kcount = 0, rcount = 0, ucount = 0; /* k-steps, window rows and units so far: every role counts the same way */

	/* producer state */
	uint8_t* const s_base = smem8 + (warp & 3) * GTC_WARP_BYTES;
	uint8_t* const s_tile = s_base;
	uint32_t* const s_histw = reinterpret_cast<uint32_t*>(s_base + GRP_TILE_BYTES);       /* word of bin b: [b * 16 + lane / 2] */
	const uint16_t* const s_hist = reinterpret_cast<const uint16_t*>(s_histw) + lane;      /* count of bin b: [b * 32] */
	uint32_t* const s_lutw = reinterpret_cast<uint32_t*>(s_base + GRP_TILE_BYTES + GRP_HIST_BYTES) + lane; /* word of bins 4q..4q+3: [q * 32] */
	const uint8_t* const s_lutb = s_base + GRP_TILE_BYTES + GRP_HIST_BYTES + lane * 4;    /* the same column, as bytes */
	float* const s_hk = reinterpret_cast<float*>(s_base + GRP_TILE_BYTES + GRP_HIST_BYTES) + lane; /* [i * 32], over the table */
	float* const s_u = s_hk + WVM_KA * 32;
	const uint32_t tile_bar = bar(GB_TILE + (warp & 3));
	const uint32_t tile_s = gtc_smem_u32(s_tile);
	const CUtensorMap* const tmaps = reinterpret_cast<const CUtensorMap*>(a.tmaps);
	uint32_t tile_phase = 0;
	const uint32_t hinc = 1u << (16 * (lane & 1));
	uint32_t* const hword = s_histw + (lane >> 1);
	const float stretch = __fdiv_rn(255.0f, (float)(PW * PH));     /* HistEq64Filter.cpp:34 */
	const int arow_off = ((warp * 32 + lane) >> 3) * 256 + (lane & 7) * 16; /* this window's row of a slab: chunk c at + c * 128 */
	const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);            /* this warp's quarter of tensor memory */

	/* work units are handed out by the issue lane: it draws unit k + 1 from the counter when it starts unit k and publishes it when
	 * it has issued the last MMA of unit k - by then every producer has read unit k's slot (it took part in all of its k-steps), and
	 * the barrier is never more than one phase ahead of a waiting producer. No CTA-wide barrier between units. */
	int my_unit = 0;
	if (tid == 128) { my_unit = atomicAdd(a.cursor, 1); s_unit[0] = my_unit; gtc_mbar_arrive(bar(GB_UNIT)); }
	if (warp != 4 || lane == 0)
	for (;; ++ucount) {
		int unit = my_unit, next_unit = 0;
		if (warp != 4) {
			gtc_mbar_wait<GTC_SLEEP_PROD_>(bar(GB_UNIT), ucount & 1);
			unit = s_unit[ucount & 1];
		}
		if (unit >= total_units) break;
		if (warp == 4) next_unit = atomicAdd(a.cursor, 1);
		const int item_id = unit / quads, quad = unit - item_id * quads;
		const GroupItem it = a.items[item_id];
		const int maxrows = min(it.run, it.windows_y - it.iy0);

		if (warp == 4) {
			/* ===== B stream + MMA issue (one lane) ===== */
			if (lane == 0) {
				const uint32_t idesc = (2u << 4) | ((uint32_t)((32 * it.nm) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); /* D s32, A = B = u8, K-major */
				const uint8_t* btc[MSUB];
#pragma unroll
				for (int k = 0; k < MSUB; ++k) btc[k] = a.models[k < it.nm ? it.model[k] : it.model[0]].m.btc;
				const int total = maxrows * KS;
				const uint32_t a_addr = gtc_smem_u32(s_slabs), b_addr = gtc_smem_u32(s_bring);
				auto issue_b = [&](int j) {
					const uint32_t kc = kcount + (uint32_t)j;
					const uint32_t stage = kc % GTC_NB;
					gtc_mbar_wait<GTC_SLEEP_ISSUE_>(bar(GB_B_EMPTY + stage), ((kc / GTC_NB) & 1) ^ 1);
					gtc_mbar_expect_tx(bar(GB_B_FULL + stage), (uint32_t)it.nm * 1024u);
					const int s = j % KS;
#pragma unroll
					for (int k = 0; k < MSUB; ++k)
						if (k < it.nm)
							asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
									:: "r"(b_addr + stage * BSTAGE + k * 1024), "l"(btc[k] + (size_t)s * 1024), "r"(1024u), "r"(bar(GB_B_FULL + stage)) : "memory");
				};
				for (int j = 0; j < min(GTC_B_AHEAD, total); ++j) issue_b(j);
				for (int j = 0; j < total; ++j) {
					if (j + GTC_B_AHEAD < total) issue_b(j + GTC_B_AHEAD);
					const int s = j % KS;
					if (s == 0) { /* the producers have read the previous row's accumulators */
						gtc_mbar_wait<GTC_SLEEP_ISSUE_>(bar(GB_ACC_EMPTY), (rcount & 1) ^ 1);
						gtc_fence_after();
					}
					const uint32_t kc = kcount + (uint32_t)j;
					const uint32_t stage = kc % GTC_NB, buf = kc % GTC_NA;
					gtc_mbar_wait<GTC_SLEEP_ISSUE_>(bar(GB_B_FULL + stage), (kc / GTC_NB) & 1);
					gtc_mbar_wait<GTC_SLEEP_ISSUE_>(bar(GB_A_FULL + buf), (kc / GTC_NA) & 1);
					gtc_fence_after();
					gtc_mma_i8(tmem, gtc_desc(a_addr + buf * GTC_SLAB, 128, 256), gtc_desc(b_addr + stage * BSTAGE, 128, 256), idesc, s != 0 ? 1u : 0u);
					gtc_commit(bar(GB_A_EMPTY + buf));
					gtc_commit(bar(GB_B_EMPTY + stage));
					if (s == KS - 1) { gtc_commit(bar(GB_ACC_FULL)); ++rcount; }
				}
				kcount += (uint32_t)total;
				s_unit[(ucount + 1) & 1] = next_unit;
				gtc_mbar_arrive(bar(GB_UNIT));
				my_unit = next_unit;
			}
			continue;
		}

		/* ===== producers: warp = frame of the quad, lane = window ===== */
		const int frame = quad * 4 + warp;
		const bool have = frame < a.n_frames;
		const GroupImage im = a.images[it.image];
		const int tx0 = it.begin_x + it.ix0, ty0 = it.begin_y + it.iy0;
		if (have) {
			/* --- stage the strip's pixels as histogram bins (v >> 2, HistEq64Filter.cpp:14-25); zero outside the image --- */
			if (tmaps != nullptr && im.tma_ok) {
				if (lane == 0) {
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the previous unit's generic accesses to the tile */
					asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(tile_bar), "r"(GRP_TILE_BYTES) : "memory");
					asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
							:: "r"(tile_s), "l"(reinterpret_cast<uint64_t>(tmaps + it.image)), "r"(tx0), "r"(ty0), "r"(frame), "r"(tile_bar) : "memory");
				}
				__syncwarp();
				gtc_mbar_wait<0>(tile_bar, tile_phase);
				tile_phase ^= 1u;
				uint32_t* const tw32 = reinterpret_cast<uint32_t*>(s_tile);
				for (int i = lane; i < GRP_TILE_BYTES / 4; i += 32) tw32[i] = (tw32[i] >> 2) & 0x3f3f3f3fu;
			} else {
				const uint8_t* __restrict__ img = (im.offset < 0 ? a.frames + (int64_t)frame * a.W * a.H
						: a.arena + (int64_t)frame * a.arena_stride + im.offset);
				for (int r = 0; r < STRIP_TILE_ROWS; ++r) {
					const bool row_ok = ty0 + r < im.height;
					const uint8_t* row = img + (int64_t)(ty0 + r) * im.pitch + tx0;
					for (int c = lane; c < GRP_PITCH; c += 32)
						s_tile[r * GRP_PITCH + c] = (row_ok && tx0 + c < im.width) ? (uint8_t)(row[c] >> 2) : (uint8_t)0;
				}
			}
		}
		__syncwarp();

		/* --- roles: lane = window column `col` of row run `sub` --- */
		const int col = lane % it.cols, sub = lane / it.cols;
		const int iy_first = it.iy0 + sub * it.run;
		const bool valid = have && sub < it.nsub && iy_first < it.windows_y;
		const int nrows = valid ? min(it.run, it.windows_y - iy_first) : 0;
		const int org = valid ? (sub * it.run) * GRP_PITCH + col : 0; /* tile offset of this lane's first window */
		const int sh = (org & 3) * 8;                                   /* misalignment of the lane's column */

		for (int w = 0; w < maxrows; ++w, ++rcount) {
			const bool active = w < nrows;
			const uint8_t* const tw = s_tile + org + w * GRP_PITCH; /* top-left bin of this lane's window */
			uint32_t total = 0;
			if (active) {
				if (w == 0) { /* histogram of the first window of the run */
					uint16_t* const mine = const_cast<uint16_t*>(s_hist);
#pragma unroll
					for (int k = 0; k < 64; ++k) mine[k * 32] = 0;
					__syncwarp(__activemask());
					for (int r = 0; r < PH; ++r) {
#pragma unroll
						for (int c = 0; c < PW; ++c) atomicAdd(hword + tw[r * GRP_PITCH + c] * 16, hinc);
					}
				} else { /* slide down: row w-1 leaves, row w+PH-1 enters (enter first: a count never drops below zero) */
					const uint8_t* const r_out = tw - GRP_PITCH;
					const uint8_t* const r_in = tw + (PH - 1) * GRP_PITCH;
#pragma unroll
					for (int c = 0; c < PW; ++c) atomicAdd(hword + r_in[c] * 16, hinc);
#pragma unroll
					for (int c = 0; c < PW; ++c) atomicAdd(hword + r_out[c] * 16, 0u - hinc);
				}
			}
			__syncwarp(); /* both lanes of a histogram word are done with it */
			/* --- equalisation table: sequential float32 cumulative histogram, one byte column per lane --- */
			if (active) total = grp_build_table(s_hist, s_lutw, stretch);
			const float total_f = (float)total;

			/* --- per k-step: this lane's window row(s) -> its row of the A slab --- */
			float sum_xx = 0.f; /* iimg_xx->data[last]: float32 accumulation of the integer row sums in row order (IImg.cpp:33-47) */
			const uint32_t* const trow_px = reinterpret_cast<const uint32_t*>(s_tile + ((org + w * GRP_PITCH) & ~3));
#pragma unroll 1
			for (int s = 0; s < KS; ++s, ++kcount) {
				const uint32_t buf = kcount % GTC_NA;
				gtc_mbar_wait<GTC_SLEEP_PROD_>(bar(GB_A_EMPTY + buf), ((kcount / GTC_NA) & 1) ^ 1); /* the MMA that read this slab is done */
				uint8_t* const arow = s_slabs + buf * GTC_SLAB + arow_off;
#pragma unroll
				for (int pr = 0; pr < RPK; ++pr) {
					const uint32_t* const src = trow_px + (s * RPK + pr) * (GRP_PITCH / 4);
					uint32_t wd[8];
					const uint32_t rowsq = grp_equalise_row<WPR>(src, sh, s_lutb, wd);
					sum_xx = __fadd_rn(sum_xx, (float)rowsq); /* 0.f + x == x: the first row starts the sum */
					if (RPK == 2) {
						*reinterpret_cast<uint4*>(arow + pr * 128) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
					} else {
						*reinterpret_cast<uint4*>(arow) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
						*reinterpret_cast<uint4*>(arow + 128) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
					}
				}
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy stores -> visible to the tensor core */
				__syncwarp();
				if (lane == 0) gtc_mbar_arrive(bar(GB_A_FULL + buf));
			}

			/* --- this lane's row of D from tensor memory; scalar cascade over the first WVM_KA filters per model --- */
			gtc_mbar_wait<GTC_SLEEP_PROD_>(bar(GB_ACC_FULL), rcount & 1);
			gtc_fence_after();
			const int wx = it.ix0 + col, wy = iy_first + w; /* window coordinates in the layer's grid */
#pragma unroll
			for (int mi = 0; mi < MSUB; ++mi) {
				if (mi >= it.nm) break;
				const GroupModel& gm = a.models[it.model[mi]];
				const DevWvm& m = gm.m;
#pragma unroll
				for (int i = 0; i < WVM_KA; ++i) s_u[i * 32] = 0.f;                          /* WvmClassifier.cpp:129-131 */
				int level = -1;
				float fout = 0.f;
				bool alive = active;
#pragma unroll 1
				for (int lv = 0; lv < WVM_KA; ++lv) {
					if (!__any_sync(0xffffffffu, alive)) break;
					const uint4 d4 = gtc_ld4(trow + (uint32_t)(32 * mi + 4 * lv));
					if (alive) {
						level = lv;
						const int nv = __ldg(m.cntval + lv) - 1;
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
				if (active) {
					const int win = it.first_window[mi] + wy * it.windows_x + wx;
					if (alive) { /* survived every filter of this kernel: the deep kernel finishes the window */
						const int slot = atomicAdd(gm.q.count, 1);
						if (slot < gm.q.cap) {
							DeepRec r;
							r.frame = frame; r.window = win; r.total_f = total_f; r.sum_xx = sum_xx;
							r.image = it.image; r.x = it.begin_x + wx; r.y = it.begin_y + wy;
#pragma unroll
							for (int i = 0; i < WVM_KA; ++i) { r.hk[i] = s_hk[i * 32]; r.u[i] = s_u[i * 32]; }
							gm.q.rec[slot] = r;
						}
						/* slot >= cap: counted in *q.count; the host re-runs the launch on the generic path */
					} else {
						wvm_emit(m, frame, win, gm.windows_per_frame, level, fout, gm.dense, gm.cand, gm.cand_count, gm.cand_cap);
					}
				}
			}
			gtc_fence_before();
			__syncwarp();
			if (lane == 0) gtc_mbar_arrive(bar(GB_ACC_EMPTY)); /* the next row's first MMA may overwrite the accumulators */
		}
	}

	gtc_fence_before();
	__syncthreads();
	if (warp == 4) {
		gtc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TCOLS) : "memory");
	}
}

template <int PW, int PH, int MSUB>
cudaError_t gtc_configure() {
	return cudaFuncSetAttribute(wvm_group_tc_kernel<PW, PH, MSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, gtc_smem(MSUB));
}

int gtc_sm_count() {
	static int sms = 0;
	if (!sms) {
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (sms <= 0) sms = 148;
	}
	return sms;
}

template <int PW, int PH, int MSUB>
void gtc_launch(cudaStream_t st, const GroupArgs& args) {
	const int64_t units = (int64_t)args.n_items * ((args.n_frames + 3) / 4);
	const int blocks = (int)std::min<int64_t>((int64_t)gtc_sm_count() * gtc_ctas_per_sm(MSUB), units); /* one persistent CTA per resident slot */
	wvm_group_tc_kernel<PW, PH, MSUB><<<blocks, GTC_THREADS, gtc_smem(MSUB), st>>>(args);
}

} // namespace

#define GTC_SIZES(X) X(20, 20) X(24, 24) X(32, 16) X(32, 24) X(16, 24)

int group_tc_configure_all() {
	cudaError_t e = cudaSuccess;
#define GTC_CFG(PW, PH) if (e == cudaSuccess) e = gtc_configure<PW, PH, 1>(); if (e == cudaSuccess) e = gtc_configure<PW, PH, 2>(); \
	if (e == cudaSuccess) e = gtc_configure<PW, PH, 4>();
	GTC_SIZES(GTC_CFG)
#undef GTC_CFG
	return (int)e;
}

void launch_wvm_group_tc(cudaStream_t st, int pw, int ph, int pack, const GroupArgs& args) {
	if (args.n_items == 0 || args.n_frames == 0) return;
#define GTC_CASE(PW, PH) if (pw == PW && ph == PH) { if (pack <= 1) gtc_launch<PW, PH, 1>(st, args); else if (pack == 2) gtc_launch<PW, PH, 2>(st, args); \
	else gtc_launch<PW, PH, 4>(st, args); return; }
	GTC_SIZES(GTC_CASE)
#undef GTC_CASE
}

} // namespace fdb
