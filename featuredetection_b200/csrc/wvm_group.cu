/*
 * wvm_group.cu - stage 1 of the cascade for one detector or for a whole set of detectors (sm_100a):
 * HistEq64 + the first WVM_KA filters of every window (wvm_group_kernel), then the rest of the cascade for the few
 * survivors of all detectors in one launch (wvm_deep_group_kernel).
 *
 * Arithmetic and citations are those of wvm.cu / wvm_math.cuh (HistEq64Filter.cpp:32-125, IImg.cpp:26-65,
 * WvmClassifier.cpp:100-149,191-346); this file only decides who computes what.
 *
 * wvm_group_kernel<PW, PH, MSUB>   persistent warps; a warp takes (item, frame) units from an atomic counter. An item
 *   is a strip of the window grid of one pyramid image (<= 32 adjacent window columns, narrow layers pack several row
 *   runs side by side) and a pack of <= MSUB models that scan it. The warp stages the strip's pixels once by TMA as
 *   6-bit histogram bins; lane = window throughout:
 *     - sliding 64-bin histogram down the lane's column (one pixel row leaves, one enters: 2 PW fire-and-forget shared
 *       atomics), u16 counts packed two lanes per word, column layout (conflict free);
 *     - sequential float32 cumulative histogram -> the window's equalisation table, one column of words per lane (a bank
 *       per lane: look-ups never conflict; wvm_group_dev.cuh);
 *     - per patch row: equalised pixels of the lane's window (aligned word loads + funnel shift, 4 table look-ups per
 *       word) -> 32 bytes of the window's row of the A operand in shared memory; sum(x^2) accumulates per patch row in
 *       the reference's float32 order (IImg.cpp:33-47) from dp4a row sums;
 *     - ldmatrix -> mma.sync.m16n8k32.u8.u8.s32: D[window][8 filters x 4 grey values] = A[window][pixel] .
 *       B[pixel][filter, value] (rectangle coverage counts, fragment order prepared on the host) for every model of the
 *       pack from the SAME A fragments - the equalisation is shared by the models of a pack. Exact: sums < 2^24.
 *       K runs over patch rows padded to 32 bytes (one k-step per patch row; two rows per k-step for 16-wide windows),
 *       so no index arithmetic separates the window from the operand.  This kernel serves packs of one and two models
 *       (32 accumulator registers per model); packs of three and four run on wvm_group_tc.cu (tcgen05.mma, accumulators in
 *       tensor memory). For small packs mma.sync measures faster on the B200: its warps are independent (16 per SM), while the
 *       four warps of a UMMA tile walk in step and wait for the slowest one - launch_wvm_group picks by pack size.
 *     - D goes through shared memory (over the table and the A buffers, which are dead by then) and the scalar cascade
 *       tail runs per lane and model; survivors of all WVM_KA filters are queued for the deep kernel.
 *
 * wvm_deep_group_kernel   one warp per queued window of any detector of the table: re-equalises the window from the
 *   pyramid image (cheaper than carrying 400..768 bytes per survivor through HBM), integral image in shared memory,
 *   32 filters per round (one filter per lane, rectangles the reference's way), kernel values in wavelet order, the
 *   float weighted sums as 32 sequential chains whose weights arrive coalesced (weights transposed per round on the
 *   host), ballot for the first rejecting filter.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "fdb_internal.h"
#include "wvm_device.h"
#include "wvm_group.h"
#include "wvm_group_dev.cuh"
#include "wvm_math.cuh"

namespace fdb {

__device__ __forceinline__ void grp_mma_u8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
	asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
			: "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void grp_ldmatrix4(uint32_t (&r)[4], uint32_t saddr) {
	asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
			: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}

template <int PW, int PH, int MSUB>
__global__ void __launch_bounds__(GRP_WARPS * 32, MSUB == 1 ? 4 : (MSUB == 2 ? 3 : 2)) wvm_group_kernel(const __grid_constant__ GroupArgs a) {
	static_assert(PW % 4 == 0 && PW >= 16 && PW <= 32, "window width: a multiple of 4 in 16..32");
	constexpr int WPR = PW / 4;                 /* words per patch row */
	constexpr int RPK = PW <= 16 ? 2 : 1;       /* patch rows per k-step (32 operand bytes) */
	static_assert(PH % RPK == 0, "window height must split into k-steps");
	static_assert(grp_stretch_is_safe(PW * PH), "grp_hq_step's shortcut does not hold for this window size");
	constexpr int KS = PH / RPK;
#ifndef GRP_KUNROLL
#define GRP_KUNROLL 1
#endif
	constexpr int KUNROLL = GRP_KUNROLL; /* k-steps per loop iteration (tuning builds) */
	extern __shared__ __align__(128) uint8_t smem8[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint8_t* const s_base = smem8 + warp * GRP_WARP_BYTES;
	uint8_t* const s_tile = s_base;
	uint32_t* const s_histw = reinterpret_cast<uint32_t*>(s_base + GRP_TILE_BYTES);       /* word of bin b: [b * 16 + lane / 2] */
	const uint16_t* const s_hist = reinterpret_cast<const uint16_t*>(s_histw) + lane;      /* count of bin b: [b * 32] */
	uint8_t* const s_R = s_base + GRP_TILE_BYTES + GRP_HIST_BYTES;
	uint32_t* const s_lutw = reinterpret_cast<uint32_t*>(s_R) + lane;                     /* word of bins 4q..4q+3: [q * 32] */
	const uint8_t* const s_lutb = s_R + lane * 4;                                         /* the same column, as bytes */
	uint8_t* const s_abuf = s_R + GRP_LUT_BYTES;
	int* const s_stage = reinterpret_cast<int*>(s_R);                                     /* [32][GRP_STAGE] */
	float* const s_hk = reinterpret_cast<float*>(s_R + GRP_R_BYTES) + lane;               /* hk_kernel_eval[i]: [i * 32] */
	float* const s_u = s_hk + WVM_KA * 32;                                                /* u_kernel_eval[i]: [i * 32] */
	uint64_t* const s_mbar = reinterpret_cast<uint64_t*>(smem8 + GRP_WARPS * GRP_WARP_BYTES) + warp;
	const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_mbar);
	const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(s_tile);
	const uint32_t abuf_s = (uint32_t)__cvta_generic_to_shared(s_abuf);
	const CUtensorMap* const tmaps = reinterpret_cast<const CUtensorMap*>(a.tmaps);
	if (lane == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	uint32_t phase = 0;
	const int g = lane >> 2, t = lane & 3;
	const uint32_t hinc = 1u << (16 * (lane & 1));                 /* this lane's half of a histogram word */
	uint32_t* const hword = s_histw + (lane >> 1);
	const float stretch = __fdiv_rn(255.0f, (float)(PW * PH));     /* HistEq64Filter.cpp:34 */
	/* ldmatrix row addresses: matrices {rows 0-7, bytes 0-15}, {rows 8-15, 0-15}, {rows 0-7, 16-31}, {rows 8-15, 16-31} */
	const uint32_t ldm_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * GRP_AROW + (lane >> 4) * 16);
	const int total_units = a.n_items * a.n_frames;

	for (;;) {
		int unit = 0;
		if (lane == 0) unit = atomicAdd(a.cursor, 1);
		unit = __shfl_sync(0xffffffffu, unit, 0);
		if (unit >= total_units) break;
		const int item_id = unit / a.n_frames, frame = unit - item_id * a.n_frames;
		const GroupItem it = a.items[item_id];
		const GroupImage im = a.images[it.image];
		const int tx0 = it.begin_x + it.ix0, ty0 = it.begin_y + it.iy0;

		/* --- stage the strip's pixels as histogram bins (v >> 2, HistEq64Filter.cpp:14-25); zero outside the image --- */
		if (tmaps != nullptr && im.tma_ok) {
			if (lane == 0) {
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the previous unit's generic accesses to the tile */
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(GRP_TILE_BYTES) : "memory");
				asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
						:: "r"(tile_s), "l"(reinterpret_cast<uint64_t>(tmaps + it.image)), "r"(tx0), "r"(ty0), "r"(frame), "r"(bar) : "memory");
			}
			__syncwarp();
			uint32_t done = 0;
			while (!done) {
				asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
						: "=r"(done) : "r"(bar), "r"(phase) : "memory");
			}
			phase ^= 1u;
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
		__syncwarp();

		/* --- roles: lane = window column `col` of row run `sub` --- */
		const int col = lane % it.cols, sub = lane / it.cols;
		const int iy_first = it.iy0 + sub * it.run;
		const bool valid = sub < it.nsub && iy_first < it.windows_y;
		const int nrows = valid ? min(it.run, it.windows_y - iy_first) : 0;
		const int maxrows = min(it.run, it.windows_y - it.iy0);
		const int org = valid ? (sub * it.run) * GRP_PITCH + col : 0; /* tile offset of this lane's first window */
		const int sh = (org & 3) * 8;                                   /* misalignment of the lane's column */

		for (int w = 0; w < maxrows; ++w) {
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
			__syncwarp(); /* both lanes of a histogram word are done with it; the previous row's staging area is consumed */
			/* --- equalisation table: sequential float32 cumulative histogram, one byte column per lane --- */
			if (active) total = grp_build_table(s_hist, s_lutw, stretch);
			const float total_f = (float)total;

			/* --- per k-step: this lane's window row(s) -> A buffer; all models of the pack multiply the same fragments --- */
			int acc[MSUB][2][4][4];
#pragma unroll
			for (int mi = 0; mi < MSUB; ++mi)
#pragma unroll
				for (int x = 0; x < 2; ++x)
#pragma unroll
					for (int y = 0; y < 4; ++y)
#pragma unroll
						for (int z = 0; z < 4; ++z) acc[mi][x][y][z] = 0;
			float sum_xx = 0.f; /* iimg_xx->data[last]: float32 accumulation of the integer row sums in row order (IImg.cpp:33-47) */
			const uint4* bf[MSUB];
#pragma unroll
			for (int mi = 0; mi < MSUB; ++mi) bf[mi] = a.models[it.model[mi < it.nm ? mi : 0]].m.bfrag + lane * 2;
			const uint32_t* const trow = reinterpret_cast<const uint32_t*>(s_tile + ((org + w * GRP_PITCH) & ~3));
#pragma unroll KUNROLL
			for (int s = 0; s < KS; ++s) {
				/* the models' B fragments of this k-step: requested first, they arrive while the A rows are built */
				uint4 bA[MSUB], bB[MSUB];
#pragma unroll
				for (int mi = 0; mi < MSUB; ++mi) { bA[mi] = __ldg(bf[mi] + s * 64); bB[mi] = __ldg(bf[mi] + s * 64 + 1); }
				uint8_t* const arow = s_abuf + (s & 1) * GRP_ABUF_BYTES + lane * GRP_AROW;
#pragma unroll
				for (int pr = 0; pr < RPK; ++pr) {
					const uint32_t* const src = trow + (s * RPK + pr) * (GRP_PITCH / 4);
					uint32_t wd[8];
					const uint32_t rowsq = grp_equalise_row<WPR>(src, sh, s_lutb, wd);
					sum_xx = (s == 0 && pr == 0) ? (float)rowsq : __fadd_rn(sum_xx, (float)rowsq);
					if (RPK == 2) {
						*reinterpret_cast<uint4*>(arow + pr * 16) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
					} else {
						*reinterpret_cast<uint4*>(arow) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
						*reinterpret_cast<uint4*>(arow + 16) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
					}
				}
				__syncwarp();
				uint32_t af[2][4];
				grp_ldmatrix4(af[0], abuf_s + (s & 1) * GRP_ABUF_BYTES + ldm_off);
				grp_ldmatrix4(af[1], abuf_s + (s & 1) * GRP_ABUF_BYTES + 16 * GRP_AROW + ldm_off);
#pragma unroll
				for (int mi = 0; mi < MSUB; ++mi) {
					grp_mma_u8(acc[mi][0][0], af[0], bA[mi].x, bA[mi].y);
					grp_mma_u8(acc[mi][0][1], af[0], bA[mi].z, bA[mi].w);
					grp_mma_u8(acc[mi][0][2], af[0], bB[mi].x, bB[mi].y);
					grp_mma_u8(acc[mi][0][3], af[0], bB[mi].z, bB[mi].w);
					grp_mma_u8(acc[mi][1][0], af[1], bA[mi].x, bA[mi].y);
					grp_mma_u8(acc[mi][1][1], af[1], bA[mi].z, bA[mi].w);
					grp_mma_u8(acc[mi][1][2], af[1], bB[mi].x, bB[mi].y);
					grp_mma_u8(acc[mi][1][3], af[1], bB[mi].z, bB[mi].w);
				}
			}

			const int wx = it.ix0 + col, wy = iy_first + w; /* window coordinates in the layer's grid */
#pragma unroll
			for (int mi = 0; mi < MSUB; ++mi) {
				if (mi >= it.nm) break;
				const GroupModel& gm = a.models[it.model[mi]];
				const DevWvm& m = gm.m;
				__syncwarp(); /* table, A buffers and the previous model's staging area are consumed */
#pragma unroll
				for (int mt = 0; mt < 2; ++mt)
#pragma unroll
					for (int nt = 0; nt < 4; ++nt)
#pragma unroll
						for (int h = 0; h < 2; ++h)
							*reinterpret_cast<int2*>(s_stage + (g + 8 * h + 16 * mt) * GRP_STAGE + 8 * nt + 2 * t) =
									make_int2(acc[mi][mt][nt][2 * h], acc[mi][mt][nt][2 * h + 1]);
				__syncwarp();
				/* --- scalar cascade over the first WVM_KA filters (WvmClassifier.cpp:129-138, 191-346) --- */
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
						const uint4 d4 = *reinterpret_cast<const uint4*>(s_stage + lane * GRP_STAGE + 4 * lv);
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
		}
		__syncwarp();
	}
}

/* ---------------------------------------------------------------------------------------------
 * deep kernel: the rest of the cascade for the survivors of the window kernel, all models of the table in one launch.
 * One warp per queued window, 32 filters per round (lane = filter). A variant that finished 8 windows per warp together to
 * load every round's weights once (the float weighted sums need 193 KB of weights per window for 280 filters) measured
 * SLOWER on the B200 (90 vs 61 ms per 256-frame step): the weights of the one model a CTA works on stay L1 resident, so the
 * kernel is bound by its ~9 000 warp instructions per full-depth window, not by the weight stream.
 * ------------------------------------------------------------------------------------------- */
#define GDEEP_WARPS 4
#define GDEEP_IIMG 1024   /* (w + 1) * (h + 1) <= 1024 ints per warp (32 x 24 -> 825) */
#define GDEEP_PX 768      /* w * h <= 768 */

/* HistEq64 of the window straight from the pyramid image (HistEq64Filter.cpp:32-125), then its integral image with a zero
 * first row and column: ii[(y+1)*pitch + x+1] = sum of x[0..y][0..x] (exact integers < 2^24) */
__device__ __forceinline__ void gdeep_integral_image(const DeepArgs& a, const DeepRec& rec, int pw, int ph, float stretch, int lane,
		uint32_t* hist, uint8_t* lut, uint8_t* px, int* ii) {
	const GroupImage im = a.images[rec.image];
	const uint8_t* __restrict__ img = (im.offset < 0 ? a.frames + (int64_t)rec.frame * a.W * a.H
			: a.arena + (int64_t)rec.frame * a.arena_stride + im.offset) + (int64_t)rec.y * im.pitch + rec.x;
	const int pitch = pw + 1;
	hist[lane] = 0; hist[lane + 32] = 0;
	__syncwarp();
	for (int r = 0; r < ph; ++r)
		if (lane < pw) {
			const uint8_t b = __ldg(img + (int64_t)r * im.pitch + lane) >> 2;
			px[r * pw + lane] = b;
			atomicAdd(hist + b, 1u);
		}
	__syncwarp();
	if (lane == 0) {
		float cdf = 0.f;
		for (int k = 0; k < 64; ++k) lut[k] = (uint8_t)grp_hq_step_any(cdf, hist[k], stretch);
	}
	__syncwarp();
	for (int i = lane; i < pitch; i += 32) ii[i] = 0;
	int colsum = 0; /* lane = column */
	for (int r = 0; r < ph; ++r) {
		uint32_t v = lane < pw ? (uint32_t)lut[px[r * pw + lane]] : 0u;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { /* inclusive prefix over the row */
			const uint32_t tt = __shfl_up_sync(0xffffffffu, v, d);
			if (lane >= d) v += tt;
		}
		colsum += (int)v; /* ii[r+1][c+1] = ii[r][c+1] + prefix(r, c) */
		if (lane < pw) ii[(r + 1) * pitch + lane + 1] = colsum;
		if (lane == 0) ii[(r + 1) * pitch] = 0;
	}
	__syncwarp();
}

/* kernel values of the round's filters base .. base + cnt - 1 (lane = filter): rectangle sums the reference's way
 * (WvmClassifier.cpp:277-306), then hk_kernel_eval in wavelet order where filters share u_kernel_eval (:313-314) */
__device__ __forceinline__ void gdeep_kernel_values(const DevWvm& m, const DeepRec& rec, int base, int cnt, int lane, const int* ii, int pitch,
		float* us, float* hk) {
	const int level = base + lane;
	const bool owner = lane < cnt;
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
}

__global__ void __launch_bounds__(GDEEP_WARPS * 32) wvm_deep_group_kernel(const __grid_constant__ DeepArgs a) {
	__shared__ __align__(16) float s_hk[GDEEP_WARPS][FDB_MAX_FILTERS + 4];
	__shared__ float s_u[GDEEP_WARPS][FDB_MAX_PER_LEVEL];
	__shared__ int s_ii[GDEEP_WARPS][GDEEP_IIMG];
	__shared__ uint8_t s_px[GDEEP_WARPS][GDEEP_PX];
	__shared__ uint32_t s_hist[GDEEP_WARPS][64];
	__shared__ uint8_t s_lut[GDEEP_WARPS][64];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	float* const hk = s_hk[warp];
	float* const us = s_u[warp];
	int* const ii = s_ii[warp];
	/* every CTA works through the queues of all models in turn: the queues differ in length by orders of magnitude (a face detector
	 * scans a few small layers, a landmark detector three large ones), so a grid split by model would leave most CTAs idle */
	for (int mi = 0; mi < a.n_models; ++mi) {
	const GroupModel& gm = a.models[mi];
	const DevWvm& m = gm.m;
	const DeepQueue& q = gm.q;
	const int n = min(*q.count, q.cap);
	const int pw = m.fsx, ph = m.fsy, pitch = pw + 1;
	const float stretch = __fdiv_rn(255.0f, (float)(pw * ph));
	for (;;) {
		/* dynamic work distribution: windows differ by 50x in cost (first-round exits vs. full depth) */
		int slot = 0;
		if (lane == 0) slot = atomicAdd(q.next, 1);
		slot = __shfl_sync(0xffffffffu, slot, 0);
		if (slot >= n) break;
		const DeepRec rec = q.rec[slot];
		gdeep_integral_image(a, rec, pw, ph, stretch, lane, s_hist[warp], s_lut[warp], s_px[warp], ii);
		for (int i = lane; i < m.per_level; i += 32) us[i] = 0.f;
		__syncwarp();
		if (lane < WVM_KA) { hk[lane] = rec.hk[lane]; if (lane < m.per_level) us[lane] = rec.u[lane]; }
		__syncwarp();
		int final_level = -1;
		float final_fout = 0.f;
		int round = 0;
		for (int base = WVM_KA; base < m.num_used && final_level < 0; base += 32, ++round) {
			const int cnt = min(32, m.num_used - base);
			const int level = base + lane;
			const bool owner = lane < cnt;
			gdeep_kernel_values(m, rec, base, cnt, lane, ii, pitch, us, hk);
			/* float weighted sums (WvmClassifier.cpp:340-341): one sequential chain per lane; the weights of the round's 32
			 * filters are stored [p / 4][lane][4] so that a warp load is one coalesced 512-byte request; the next group of
			 * weights is in flight while the current one is added */
			float res = 0.f;
			bool pass = true;
			{
				const float4* __restrict__ w4 = reinterpret_cast<const float4*>(m.hk_weights_t) + __ldg(m.hk_t_off + round) + lane;
				const float4* h4 = reinterpret_cast<const float4*>(hk);
				res = owner ? -__ldg(m.lin_thresholds + level) : 0.f;
				const int full = base >> 2;                 /* groups every lane adds completely (p < base <= level) */
				const int groups = (base + cnt + 3) >> 2;
				float4 wn = __ldg(w4);
				int gq = 0;
				for (; gq < full; ++gq) {
					const float4 wc = wn;
					wn = __ldg(w4 + (size_t)(gq + 1) * 32); /* the table is padded by one group */
					const float4 hc = h4[gq];
					res = __fadd_rn(res, __fmul_rn(wc.x, hc.x)); res = __fadd_rn(res, __fmul_rn(wc.y, hc.y));
					res = __fadd_rn(res, __fmul_rn(wc.z, hc.z)); res = __fadd_rn(res, __fmul_rn(wc.w, hc.w));
				}
				for (; gq < groups; ++gq) { /* the triangle's edge: a lane stops at p == level */
					const float4 wc = wn;
					wn = __ldg(w4 + (size_t)(gq + 1) * 32);
					const float4 hc = h4[gq];
					const int p = 4 * gq;
					if (owner && p <= level) res = __fadd_rn(res, __fmul_rn(wc.x, hc.x));
					if (owner && p + 1 <= level) res = __fadd_rn(res, __fmul_rn(wc.y, hc.y));
					if (owner && p + 2 <= level) res = __fadd_rn(res, __fmul_rn(wc.z, hc.z));
					if (owner && p + 3 <= level) res = __fadd_rn(res, __fmul_rn(wc.w, hc.w));
				}
				if (owner) pass = res >= __ldg(m.thresholds + level) && level + 1 < m.num_used;
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
			wvm_emit(m, rec.frame, rec.window, gm.windows_per_frame, final_level, final_fout, gm.dense, gm.cand, gm.cand_count, gm.cand_cap);
		__syncwarp();
	}
	}
}

/* ---------------------------------------------------------------------------------------------
 * launchers
 * ------------------------------------------------------------------------------------------- */
#ifndef GRP_MMA_PACK
#define GRP_MMA_PACK 2 /* register accumulators (32 per model) limit the mma.sync kernel to packs of two */
#endif

template <int PW, int PH, int MSUB>
static cudaError_t grp_configure() {
	return cudaFuncSetAttribute(wvm_group_kernel<PW, PH, MSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRP_SMEM);
}

#define GRP_SIZES(X) X(20, 20) X(24, 24) X(32, 16) X(32, 24) X(16, 24)

int group_configure_all() {
	cudaError_t e = cudaSuccess;
#if GRP_MMA_PACK > 2
#define GRP_CFG(PW, PH) if (e == cudaSuccess) e = grp_configure<PW, PH, 1>(); if (e == cudaSuccess) e = grp_configure<PW, PH, 2>(); \
	if (e == cudaSuccess) e = grp_configure<PW, PH, GRP_MMA_PACK>();
#else
#define GRP_CFG(PW, PH) if (e == cudaSuccess) e = grp_configure<PW, PH, 1>(); if (e == cudaSuccess) e = grp_configure<PW, PH, 2>();
#endif
	GRP_SIZES(GRP_CFG)
#undef GRP_CFG
	if (e == cudaSuccess) return group_tc_configure_all();
	return (int)e;
}

bool group_supported(int pw, int ph) {
#define GRP_SUP(PW, PH) if (pw == PW && ph == PH) return true;
	GRP_SIZES(GRP_SUP)
#undef GRP_SUP
	return false;
}

static int grp_sm_count() {
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
static void grp_launch(cudaStream_t st, const GroupArgs& args) {
	const int64_t units = (int64_t)args.n_items * args.n_frames;
	const int resident = grp_sm_count() * (MSUB == 1 ? 4 : (MSUB == 2 ? 3 : 2)); /* one persistent CTA per resident slot */
	const int blocks = (int)std::min<int64_t>(resident, (units + GRP_WARPS - 1) / GRP_WARPS);
	wvm_group_kernel<PW, PH, MSUB><<<blocks, GRP_WARPS * 32, GRP_SMEM, st>>>(args);
}

void launch_wvm_group_mma(cudaStream_t st, int pw, int ph, int pack, const GroupArgs& args) {
	if (args.n_items == 0 || args.n_frames == 0) return;
#if GRP_MMA_PACK > 2
#define GRP_CASE(PW, PH) if (pw == PW && ph == PH) { if (pack <= 1) grp_launch<PW, PH, 1>(st, args); else if (pack == 2) grp_launch<PW, PH, 2>(st, args); \
	else grp_launch<PW, PH, GRP_MMA_PACK>(st, args); return; }
#else
#define GRP_CASE(PW, PH) if (pw == PW && ph == PH) { if (pack <= 1) grp_launch<PW, PH, 1>(st, args); else grp_launch<PW, PH, 2>(st, args); return; }
#endif
	GRP_SIZES(GRP_CASE)
#undef GRP_CASE
}

/* which window kernel runs a pack: measured on the B200 (profiles/), the mma.sync kernel wins for packs of one and two models
 * (16 independent warps per SM hide the table look-up latency best), the tcgen05 kernel for packs of three and four (one
 * equalisation and one A operand for up to four models; accumulators in tensor memory). FDB_WINDOW_KERNEL=mma | tc forces one. */
static int group_kernel_choice() { /* 0: by pack size, 1: mma.sync only, 2: tcgen05 only; read at every call (tests switch it) */
	const char* e = std::getenv("FDB_WINDOW_KERNEL");
	return !e ? 0 : (e[0] == 'm' ? 1 : (e[0] == 't' ? 2 : 0));
}

int group_max_pack(bool tc_ok) { return group_kernel_choice() == 1 || !tc_ok ? GRP_MMA_PACK : GRP_MAX_PACK; }

void launch_wvm_group(cudaStream_t st, int pw, int ph, int pack, const GroupArgs& args, bool tc_ok) {
	const int choice = group_kernel_choice();
	if (tc_ok && (choice == 2 || (choice == 0 && pack > GRP_MMA_PACK))) launch_wvm_group_tc(st, pw, ph, pack, args);
	else launch_wvm_group_mma(st, pw, ph, pack, args);
}

void launch_wvm_deep_group(cudaStream_t st, const DeepArgs& args) {
	if (args.n_models == 0) return;
	wvm_deep_group_kernel<<<grp_sm_count() * 8, GDEEP_WARPS * 32, 0, st>>>(args);
}

} // namespace fdb
