/*
 * wvm_group_dev.cuh - constants and arithmetic shared by the window kernels of wvm_group.cu (mma.sync) and
 * wvm_group_tc.cu (tcgen05): the per-warp shared-memory pieces and one step of the HistEq64 cumulative histogram.
 */
#ifndef FDB_WVM_GROUP_DEV_CUH_
#define FDB_WVM_GROUP_DEV_CUH_

#include <cuda_runtime.h>
#include <cstdint>

#include "wvm_device.h"

namespace fdb {

#define GRP_WARPS 4
#define GRP_PITCH 64                    /* bytes per tile row: 32 columns + PW - 1 <= 63 */
#define GRP_TILE_BYTES (STRIP_TILE_ROWS * GRP_PITCH)
#define GRP_HIST_BYTES (64 * 32 * 2)    /* u16 [bin][lane] */
#define GRP_LUT_BYTES (16 * 32 * 4)     /* u32 [bin / 4][lane]: byte bin % 4 of the word is the entry of that bin */
#define GRP_AROW 48                     /* bytes between windows in an A buffer: 32 used; 3 x 16 keeps ldmatrix and the 16-byte stores conflict free */
#define GRP_ABUF_BYTES (32 * GRP_AROW)
#define GRP_STAGE 40                    /* ints per window row of the D staging area */
#define GRP_R_BYTES (GRP_LUT_BYTES + 2 * GRP_ABUF_BYTES) /* table + two A buffers = the staging area */
#define GRP_HKU_BYTES (2 * WVM_KA * 32 * 4)
#define GRP_WARP_BYTES (GRP_TILE_BYTES + GRP_HIST_BYTES + GRP_R_BYTES + GRP_HKU_BYTES)
#define GRP_SMEM (GRP_WARPS * GRP_WARP_BYTES + GRP_WARPS * 8)

static_assert(WVM_KA == 8, "the fragment table holds 8 filters x 4 grey values = 32 columns");
static_assert(GRP_R_BYTES == 32 * GRP_STAGE * 4, "the staging area overlays the table and the A buffers exactly");
static_assert(GRP_WARP_BYTES % 128 == 0, "TMA destinations are 128-byte aligned");

/* one step of the sequential cumulative histogram (HistEq64Filter.cpp:70-87,97): cdf += count * stretch in float32, then
 * (uchar)floor((double)cdf + 0.5). For 0 <= cdf < 256.5 that equals floor(cdf +f 0.5f) for EVERY float except the one just
 * below 0.5 (0x1.fffffep-2: the float sum rounds up to 1.0) - checked exhaustively over all 1.13e9 floats of the range.
 * A cumulative histogram below 0.5 is a single product count * stretch (stretch > 0.25 for windows of <= 1020 pixels), and
 * grp_stretch_is_safe() verifies at compile time that no such product is that float for the window sizes built here. */
__device__ __forceinline__ uint32_t grp_hq_step(float& cdf, uint32_t cnt, float stretch) {
	cdf = __fadd_rn(cdf, __fmul_rn((float)cnt, stretch));
	return (uint32_t)__float2int_rd(__fadd_rn(cdf, 0.5f));
}

__host__ __device__ constexpr bool grp_stretch_is_safe(int pixels) {
	const float stretch = 255.0f / (float)pixels;
	if (!(stretch > 0.25f)) return false;
	for (int cnt = 1; (float)cnt * stretch < 0.5f; ++cnt)
		if ((float)cnt * stretch > 0.4999999f) return false;
	return true;
}

/* the deep kernel equalises windows of any size: same value, exception handled explicitly */
__device__ __forceinline__ uint32_t grp_hq_step_any(float& cdf, uint32_t cnt, float stretch) {
	cdf = __fadd_rn(cdf, __fmul_rn((float)cnt, stretch));
	const float fl = floorf(cdf); /* (uchar)floor((double)cdf + 0.5) == floor(cdf) + (frac >= 0.5) */
	return ((uint32_t)(int)fl + (__fsub_rn(cdf, fl) >= 0.5f ? 1u : 0u)) & 255u;
}

/* ---- equalisation table of the lane's window -------------------------------------------------------------------------
 * 64 u8 entries per lane stored as 16 words in a column of words, word (bin / 4) of lane l at byte (bin / 4) * 128 + 4 l:
 * every lane owns one shared-memory bank, so the 4 x PW x PH look-ups of a window never conflict (a byte table [bin][lane]
 * measured 1.8 wavefronts per look-up: the four lanes of a word collide whenever their bins differ by a multiple of four). */

/* sequential float32 cumulative histogram (HistEq64Filter.cpp:70-87,97) of the lane's column of u16 counts -> the lane's table;
 * returns sum(count * entry) = the sum of the equalised window */
__device__ __forceinline__ uint32_t grp_build_table(const uint16_t* hist /* + lane, stride 32 */, uint32_t* lutw /* + lane, stride 32 */,
		float stretch) {
	float cdf = 0.f;
	uint32_t total = 0;
#pragma unroll 4
	for (int q = 0; q < 16; ++q) {
		uint32_t word = 0;
#pragma unroll
		for (int r = 0; r < 4; ++r) {
			const uint32_t cnt = hist[(4 * q + r) * 32];
			const uint32_t e = grp_hq_step(cdf, cnt, stretch);
			word |= e << (8 * r);
			total += cnt * e;
		}
		lutw[q * 32] = word;
	}
	return total;
}

/* equalised pixels of one patch row of the lane's window: `src` = the aligned words of the tile row that hold the lane's PW bins
 * (as v >> 2), `sh` = 8 x the lane's misalignment; wd[c] = 4 equalised pixels, zero beyond the row. Returns sum(x^2) of the row.
 * Per word: the four table words are addressed by bin / 4 (one PRMT + one scaled add each), the entries are cut out by the two
 * PRMTs that also pack them (selector nibbles = bin % 4, built for all four pixels with two shifts and two LOP3). */
template <int WPR>
__device__ __forceinline__ uint32_t grp_equalise_row(const uint32_t* src, int sh, const uint8_t* lutb /* the lane's column, as bytes */,
		uint32_t (&wd)[8]) {
	uint32_t x[WPR + 1];
#pragma unroll
	for (int c = 0; c <= WPR; ++c) x[c] = src[c];
	uint32_t rowsq = 0;
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		if (c < WPR) {
			const uint32_t b = __funnelshift_r(x[c], x[c + 1], sh); /* 4 bins of the lane's window */
			const uint32_t bq = b & 0xfcfcfcfcu;                    /* 4 * (bin / 4): times 32 = the byte offset of the table word */
			const uint32_t w0 = *reinterpret_cast<const uint32_t*>(lutb + __byte_perm(bq, 0, 0x4440) * 32);
			const uint32_t w1 = *reinterpret_cast<const uint32_t*>(lutb + __byte_perm(bq, 0, 0x4441) * 32);
			const uint32_t w2 = *reinterpret_cast<const uint32_t*>(lutb + __byte_perm(bq, 0, 0x4442) * 32);
			const uint32_t w3 = *reinterpret_cast<const uint32_t*>(lutb + __byte_perm(bq, 0, 0x4443) * 32);
			const uint32_t r = b & 0x03030303u;                     /* bin % 4 per pixel */
			const uint32_t t = r | (r >> 4) | 0x00400040u;          /* low half: r0 | (4 + r1) << 4, high half: r2 | (4 + r3) << 4 */
			const uint32_t lo = __byte_perm(w0, w1, t), hi = __byte_perm(w2, w3, t >> 16); /* bytes 0, 1 = the two entries */
			wd[c] = __byte_perm(lo, hi, 0x5410);
			rowsq = __dp4a(wd[c], wd[c], rowsq);
		} else wd[c] = 0u;
	}
	return rowsq;
}

} // namespace fdb
#endif
