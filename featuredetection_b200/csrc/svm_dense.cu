/*
 * svm_dense.cu - the RBF support vector machine over EVERY window of a batch as one integer matrix product on the
 * 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in tensor memory), sm_100a only.
 *
 * What it replaces: the `single` detector of ffpDetectApp with a psvm classifier (ffpDetectApp.cpp:427-500) -
 * SlidingWindowDetector::detect (SlidingWindowDetector.cpp:40-98) calls, for every window of every pyramid layer,
 *   HistEq64Filter::applyTo                         HistEq64Filter.cpp:32-125
 *   SvmClassifier::computeHyperplaneDistance        SvmClassifier.cpp:55-60      distance = -bias + sum_i coef_i k(x, sv_i)
 *   RbfKernel::compute                              RbfKernel.hpp:32-40,78-108   k = exp(-gamma * sum (x - sv)^2), u8 -> int
 * i.e. windows x support vectors x pixels multiply-adds (16 185 x 1024 x 400 per 640x480 FaceFrontal frame).
 *
 * The sum of squared differences of two u8 vectors is an exact integer: |x|^2 + |sv|^2 - 2 x.sv, and x.sv for all
 * (window, support vector) pairs is the matrix product [windows x pixels] . [pixels x support vectors] of u8 operands
 * with s32 accumulation - exact on the integer tensor cores (400 * 255 * 255 < 2^31).
 *
 * One persistent CTA per SM, 14 warps with fixed roles; a tile is 128 windows:
 *   warp 0      streams the support-vector blocks (pre-arranged on the host as UMMA core matrices) from L2 into a
 *               4-stage shared-memory ring with cp.async.bulk + mbarrier complete_tx; the whole model passes once per tile
 *   warp 1      owns tensor memory (512 columns) and issues tcgen05.mma (one lane): M128 N128 K32, 13 k-steps per block of
 *               128 support vectors; 4 accumulator blocks of 128 columns alternate between MMA and epilogue
 *   warps 2-5   producers of the A operand: thread = window; HistEq64 of the window straight from the pyramid layer
 *               (float32 cdf in the reference's order, like the other kernels), equalised pixels written to shared
 *               memory as 8x16-byte core matrices, |x|^2 from the histogram; two A buffers, so tile i + 1 is produced while
 *               tile i is multiplied and summed; no equalised patch ever touches HBM
 *   warps 6-13  epilogue: two threads per window row (tcgen05.ld 32 lanes x 16 columns), ssd = |x|^2 + |sv|^2 - 2 dot,
 *               k = exp(-gamma ssd) in float64, 8 independent float64 partial sums per thread (the float64 pipe has a
 *               long latency), added in a fixed order - deterministic, but not the reference's left-to-right order
 *
 * exp(-gamma * ssd) for an integer ssd: ssd = hi * 2^s + lo, exp(-gamma hi 2^s) from a table of glibc-computed doubles
 * in shared memory, exp(-gamma lo) by its degree-4 Taylor polynomial (gamma * 2^s <= 2^-7, truncation < 2.4e-13 relative).
 * The kernel value therefore differs from glibc's exp by at most 2.4e-13 relative - far inside the 1e-4 score
 * tolerance (tests compare distances at 1e-9) - while the integer part of the computation is exact.
 *
 * Shared-memory operand layout (no swizzle, K-major): a core matrix is 8 rows x 16 bytes stored as 128 contiguous
 * bytes; core matrices adjacent in K are 128 bytes apart (descriptor LBO), groups of 8 rows SBO bytes apart.
 */
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fdb_internal.h"
#include "wvm_device.h"

namespace fdb {

namespace {

constexpr int SD_ROWS = 128;      /* window rows per MMA (UMMA M) */
constexpr int SD_ABUF = 2;        /* window tiles in shared memory: one being multiplied, one being produced */
constexpr int SD_ACC = 4;         /* accumulator blocks in tensor memory (4 x 128 columns) */
constexpr int SD_EPI_SPLIT = 2;   /* epilogue threads per window row */
constexpr int SD_XX_RING = 8;     /* |x|^2 slots (tiles the producers may be ahead of the epilogue) */
constexpr int SD_N = 128;         /* support vectors per accumulator block (UMMA N) */
constexpr int SD_EPI_COLS = SD_N / SD_EPI_SPLIT; /* accumulator columns (support vectors) per epilogue thread and block */
constexpr int SD_KCH = 8;         /* 16-byte k chunks per streamed support-vector block */
constexpr int SD_STAGES = 4;
constexpr int SD_B_STAGE_BYTES = (SD_N / 8) * SD_KCH * 128;
constexpr int SD_PROD_SETS = 1;    /* sets of 4 producer warps (set s builds the tiles s, s + SD_PROD_SETS, ... of its CTA) */
constexpr int SD_PROD_WARPS = 4 * SD_PROD_SETS, SD_EPI_WARPS = 4 * SD_EPI_SPLIT;
constexpr int SD_THREADS = 32 * (2 + SD_PROD_WARPS + SD_EPI_WARPS);
constexpr int SD_TMEM_COLS = 512;
constexpr int SD_SPIN_LIMIT = 1 << 24;

/* barrier slots */
enum { BAR_B_FULL = 0, BAR_B_EMPTY = SD_STAGES, BAR_A_FULL = 2 * SD_STAGES, BAR_A_EMPTY = BAR_A_FULL + SD_ABUF,
	BAR_T_FULL = BAR_A_EMPTY + SD_ABUF, BAR_T_EMPTY = BAR_T_FULL + SD_ACC, BAR_COUNT = BAR_T_EMPTY + SD_ACC };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
/* bounded wait: a protocol error traps instead of hanging the device. SLEEP_NS > 0 backs off between polls so that
 * waiting warps leave the issue slots to the warps that have work (try_wait returns at once on this part). */
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t done;
	int spins = 0;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
				: "=r"(done) : "r"(bar), "r"(parity) : "memory");
		if (!done) {
			if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
			if (++spins > SD_SPIN_LIMIT) {
				printf("svm_dense_kernel: barrier %u timed out (block %d thread %d)\n", bar, blockIdx.x, threadIdx.x);
				__trap();
			}
		}
	} while (!done);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
/* D[tmem] (+)= A[smem] . B[smem]^T, u8 x u8 -> s32 */
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
			:: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
/* shared-memory matrix descriptor: K-major, no swizzle, version 1 (sm_100) */
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
	return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32)
			| ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
			"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
			: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
			  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
			: "r"(taddr) : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct SvmDenseArgs {
	DevSvmDense s;
	int patch_w, patch_h, step_x, step_y;
	const uint8_t* frames; int W, H;
	const uint8_t* arena; int64_t arena_stride;
	const DevLayer* layers; int n_layers;
	int64_t windows_per_frame;
	const uint8_t* vectors;       /* MODE 1: [total][dim] u8 */
	int64_t total;                /* rows = windows of the batch (or vectors) */
	double* distance_out;         /* [total] */
	int* pos_count;               /* positives appended (may run past pos_cap), or nullptr */
	DensePositive* pos;
	int pos_cap;
};

/* A row of one window: HistEq64 (HistEq64Filter.cpp:32-125) from the layer image into the core-matrix layout */
__device__ __forceinline__ int produce_window(const SvmDenseArgs& a, const DevLayer* sLayers, uint16_t* hist /* column of this thread */,
		uint8_t* arow /* A tile + row offset */, int64_t g) {
	const int pw = a.patch_w, ph = a.patch_h, npix = pw * ph;
	const int frame = (int)(g / a.windows_per_frame);
	const int w = (int)(g - (int64_t)frame * a.windows_per_frame);
	int li = 0;
	while (li + 1 < a.n_layers && w >= sLayers[li + 1].first_window) ++li;
	const DevLayer& L = sLayers[li];
	const int local = w - L.first_window;
	const int iy = local / L.windows_x, ix = local - iy * L.windows_x;
	const int pitch = L.pitch;
	const uint8_t* src = (L.offset < 0 ? a.frames + (int64_t)frame * a.W * a.H : a.arena + (int64_t)frame * a.arena_stride + L.offset)
			+ (int64_t)(L.begin_y + iy * a.step_y) * pitch + (L.begin_x + ix * a.step_x);
#pragma unroll 8
	for (int b = 0; b < 64; ++b) hist[b * SD_ROWS] = 0;
	for (int r = 0; r < ph; ++r) {
		const uint8_t* row = src + (int64_t)r * pitch;
		for (int c = 0; c < pw; ++c) hist[(row[c] >> 2) * SD_ROWS] += 1;
	}
	/* sequential float32 cdf, rounded: HistEq64Filter.cpp:70-87,97 */
	const float stretch = __fdiv_rn(255.0f, (float)npix);
	float cdf = 0.f;
	int xx = 0;
	for (int b = 0; b < 64; ++b) {
		const int cnt = hist[b * SD_ROWS];
		cdf = __fadd_rn(cdf, __fmul_rn((float)cnt, stretch));
		const float fl = floorf(cdf);
		const int eq = ((int)fl + (__fsub_rn(cdf, fl) >= 0.5f ? 1 : 0)) & 255; /* saturate_cast never triggers: cdf <= 255 */
		hist[b * SD_ROWS] = (uint16_t)eq;
		xx += cnt * eq * eq;
	}
	int r = 0, c = 0;
	const uint8_t* row = src;
	for (int ch = 0; ch < a.s.chunks; ++ch) {
		uint32_t wd[4] = {0u, 0u, 0u, 0u};
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			if (ch * 16 + k < npix) {
				const uint32_t e = hist[(row[c] >> 2) * SD_ROWS];
				wd[k >> 2] |= e << (8 * (k & 3));
				if (++c == pw) { c = 0; ++r; row += pitch; }
			}
		}
		*reinterpret_cast<uint4*>(arow + ch * 128) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
	}
	return xx;
}

/* 4-pixel words of one patch row from an arbitrarily aligned address: aligned word loads + funnel shifts */
template <int WPR>
__device__ __forceinline__ void load_row_words(const uint8_t* p, uint32_t (&x)[WPR]) {
	const uintptr_t ad = reinterpret_cast<uintptr_t>(p);
	const uint32_t* wp = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
	const uint32_t sh = (uint32_t)(ad & 3) * 8;
	uint32_t w[WPR + 1];
#pragma unroll
	for (int i = 0; i < WPR; ++i) w[i] = __ldg(wp + i);
	w[WPR] = sh ? __ldg(wp + WPR) : 0u; /* an aligned row ends inside word WPR - 1 */
#pragma unroll
	for (int i = 0; i < WPR; ++i) x[i] = __funnelshift_r(w[i], w[i + 1], sh);
}

/* the same for the patch sizes of ffpDetectApp (width and height multiples of 4): word loads, fire-and-forget
 * shared-memory atomics for the histogram (no read-modify-write chains), 4 rows = PW / 4 whole 16-byte chunks */
template <int PW, int PH>
__device__ __forceinline__ int produce_window_fast(const SvmDenseArgs& a, const DevLayer* sLayers, uint16_t* hist /* column of this thread */,
		uint32_t* hist_word /* the 32-bit word holding it (shared with the neighbouring thread) */, uint32_t hist_inc, uint8_t* arow, int64_t g) {
	static_assert(PW % 4 == 0 && PH % 4 == 0, "patch sides must be multiples of 4");
	constexpr int WPR = PW / 4, NPIX = PW * PH;
	const int frame = (int)(g / a.windows_per_frame);
	const int w = (int)(g - (int64_t)frame * a.windows_per_frame);
	int li = 0;
	while (li + 1 < a.n_layers && w >= sLayers[li + 1].first_window) ++li;
	const DevLayer& L = sLayers[li];
	const int local = w - L.first_window;
	const int iy = local / L.windows_x, ix = local - iy * L.windows_x;
	const int pitch = L.pitch;
	const uint8_t* src = (L.offset < 0 ? a.frames + (int64_t)frame * a.W * a.H : a.arena + (int64_t)frame * a.arena_stride + L.offset)
			+ (int64_t)(L.begin_y + iy * a.step_y) * pitch + (L.begin_x + ix * a.step_x);
#pragma unroll 16
	for (int b = 0; b < 64; ++b) hist[b * SD_ROWS] = 0;
#pragma unroll 2
	for (int r = 0; r < PH; ++r) {
		uint32_t x[WPR];
		load_row_words<WPR>(src + (int64_t)r * pitch, x);
#pragma unroll
		for (int i = 0; i < WPR; ++i)
#pragma unroll
			for (int k = 0; k < 4; ++k) atomicAdd(&hist_word[((x[i] >> (8 * k + 2)) & 63u) * (SD_ROWS / 2)], hist_inc); /* counts <= 400: no carry */
	}
	/* sequential float32 cdf, rounded: HistEq64Filter.cpp:70-87,97 */
	const float stretch = __fdiv_rn(255.0f, (float)NPIX);
	float cdf = 0.f;
	int xx = 0;
#pragma unroll 8
	for (int b = 0; b < 64; ++b) {
		const int cnt = (int)hist[b * SD_ROWS];
		cdf = __fadd_rn(cdf, __fmul_rn((float)cnt, stretch));
		const float fl = floorf(cdf);
		const int eq = ((int)fl + (__fsub_rn(cdf, fl) >= 0.5f ? 1 : 0)) & 255;
		hist[b * SD_ROWS] = (uint16_t)eq;
		xx += cnt * eq * eq;
	}
#pragma unroll 1
	for (int r4 = 0; r4 < PH; r4 += 4) {
		uint32_t o[4 * WPR];
#pragma unroll
		for (int rr = 0; rr < 4; ++rr) {
			uint32_t x[WPR];
			load_row_words<WPR>(src + (int64_t)(r4 + rr) * pitch, x);
#pragma unroll
			for (int i = 0; i < WPR; ++i) {
				uint32_t e = hist[((x[i] >> 2) & 63u) * SD_ROWS];
				e |= (uint32_t)hist[((x[i] >> 10) & 63u) * SD_ROWS] << 8;
				e |= (uint32_t)hist[((x[i] >> 18) & 63u) * SD_ROWS] << 16;
				e |= (uint32_t)hist[(x[i] >> 26) * SD_ROWS] << 24;
				o[rr * WPR + i] = e;
			}
		}
#pragma unroll
		for (int c = 0; c < WPR; ++c)
			*reinterpret_cast<uint4*>(arow + ((r4 >> 2) * WPR + c) * 128) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
	}
	for (int ch = NPIX / 16; ch < a.s.chunks; ++ch) *reinterpret_cast<uint4*>(arow + ch * 128) = make_uint4(0u, 0u, 0u, 0u);
	return xx;
}

__device__ __forceinline__ int produce_vector(const SvmDenseArgs& a, uint8_t* arow, int64_t g) {
	const int dim = a.s.dim;
	const uint8_t* v = a.vectors + g * dim;
	int xx = 0;
	for (int ch = 0; ch < a.s.chunks; ++ch) {
		uint32_t wd[4] = {0u, 0u, 0u, 0u};
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			const int i = ch * 16 + k;
			if (i < dim) wd[k >> 2] |= (uint32_t)v[i] << (8 * (k & 3));
		}
#pragma unroll
		for (int k = 0; k < 4; ++k) xx = __dp4a(wd[k], wd[k], (unsigned)xx);
		*reinterpret_cast<uint4*>(arow + ch * 128) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
	}
	return xx;
}

template <int MODE, bool CLAMP> /* MODE 0: windows of frames (HistEq64 built by the producers), 1: given u8 vectors;
                                   CLAMP: the exp table ends at the underflow point instead of the largest possible ssd */
__global__ void __launch_bounds__(SD_THREADS, 1) svm_dense_kernel(const __grid_constant__ SvmDenseArgs a) {
	extern __shared__ __align__(128) unsigned char sd_smem[];
	const int chunks = a.s.chunks;
	const int a_tile = (SD_ROWS / 8) * chunks * 128;
	uint8_t* sA = sd_smem;                                                             /* [SD_ABUF] window tiles */
	uint8_t* sB = sA + SD_ABUF * a_tile;                                               /* [SD_STAGES] support-vector blocks */
	uint16_t* sHist = reinterpret_cast<uint16_t*>(sB + SD_STAGES * SD_B_STAGE_BYTES);   /* [SD_PROD_SETS][64][SD_ROWS] */
	double* sTab = reinterpret_cast<double*>(sHist + SD_PROD_SETS * 64 * SD_ROWS);     /* [tab_n] */
	double* sPart = sTab + a.s.tab_n;                                                  /* [2][SD_EPI_SPLIT][SD_ROWS] */
	int* sXX = reinterpret_cast<int*>(sPart + 2 * SD_EPI_SPLIT * SD_ROWS);             /* [SD_XX_RING][SD_ROWS] */
	DevLayer* sLayers = reinterpret_cast<DevLayer*>(sXX + SD_XX_RING * SD_ROWS);
	uint64_t* bars = reinterpret_cast<uint64_t*>(sLayers + FDB_MAX_LAYERS);
	uint32_t* sTmem = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t bar0 = smem_u32(bars);
	auto bar = [bar0](int i) { return bar0 + 8u * (uint32_t)i; };

	if (tid == 0) {
		for (int i = 0; i < SD_STAGES; ++i) { mbar_init(bar(BAR_B_FULL + i), 1); mbar_init(bar(BAR_B_EMPTY + i), 1); }
		for (int i = 0; i < SD_ABUF; ++i) { mbar_init(bar(BAR_A_FULL + i), SD_ROWS); mbar_init(bar(BAR_A_EMPTY + i), 1); }
		for (int i = 0; i < SD_ACC; ++i) { mbar_init(bar(BAR_T_FULL + i), 1); mbar_init(bar(BAR_T_EMPTY + i), SD_EPI_WARPS); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(sTmem)), "n"(SD_TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	for (int i = tid; i < a.s.tab_n; i += SD_THREADS) sTab[i] = a.s.exp_tab[i];
	if (MODE == 0) for (int i = tid; i < a.n_layers; i += SD_THREADS) sLayers[i] = a.layers[i];
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = *sTmem;

	const int64_t ntiles = (a.total + SD_ROWS - 1) / SD_ROWS;
	const int NT = a.s.num_sv_pad / SD_N;
	const int KB = (chunks + SD_KCH - 1) / SD_KCH;
	const int nt_rot = (int)(blockIdx.x % (unsigned)NT);

	if (warp == 0) {
		/* ===== support-vector block stream: the whole model once per window tile, from L2 ===== */
		if (lane == 0) {
			uint32_t it = 0;
			const size_t nblock_bytes = (size_t)(SD_N / 8) * chunks * 128;
			for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
				for (int n0 = 0; n0 < NT; ++n0) {
					const int nt = (n0 + nt_rot) % NT; /* CTAs walk the model in different phases: spreads the L2 reads */
					const uint8_t* src = a.s.b_blocks + (size_t)nt * nblock_bytes;
					for (int kb = 0; kb < KB; ++kb, ++it) {
						const int cb = min(SD_KCH, chunks - kb * SD_KCH);
						const uint32_t bytes = (uint32_t)(SD_N / 8) * cb * 128;
						const int stage = it % SD_STAGES;
						mbar_wait<32>(bar(BAR_B_EMPTY + stage), ((it / SD_STAGES) & 1) ^ 1);
						mbar_expect_tx(bar(BAR_B_FULL + stage), bytes);
						asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
								:: "r"(smem_u32(sB + stage * SD_B_STAGE_BYTES)), "l"(src), "r"(bytes), "r"(bar(BAR_B_FULL + stage)) : "memory");
						src += bytes;
					}
				}
			}
		}
	} else if (warp == 1) {
		/* ===== MMA issue ===== */
		if (lane == 0) {
			/* instruction descriptor: D = s32, A = B = u8, both K-major, N = 128, M = 128 */
			const uint32_t idesc = (2u << 4) | ((uint32_t)(SD_N >> 3) << 17) | ((uint32_t)(SD_ROWS >> 4) << 24);
			const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
			const uint32_t a_sbo = (uint32_t)chunks * 128;
			uint32_t it = 0, acc_it = 0, tile_it = 0;
			for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
				const uint32_t abuf = tile_it % SD_ABUF;
				mbar_wait<32>(bar(BAR_A_FULL + abuf), (tile_it / SD_ABUF) & 1);
				tc_fence_after();
				for (int nt = 0; nt < NT; ++nt, ++acc_it) {
					const uint32_t acc = acc_it % SD_ACC;
					mbar_wait<32>(bar(BAR_T_EMPTY + acc), ((acc_it / SD_ACC) & 1) ^ 1);
					tc_fence_after();
					for (int kb = 0; kb < KB; ++kb, ++it) {
						const int cb = min(SD_KCH, chunks - kb * SD_KCH);
						const int stage = it % SD_STAGES;
						mbar_wait<0>(bar(BAR_B_FULL + stage), (it / SD_STAGES) & 1);
						tc_fence_after();
						const uint32_t b_sbo = (uint32_t)cb * 128;
						for (int ks = 0; ks < cb / 2; ++ks) {
							const uint32_t bstart = b_addr + stage * SD_B_STAGE_BYTES + ks * 256;
							const uint32_t astart = a_addr + abuf * a_tile + (kb * SD_KCH + 2 * ks) * 128;
							const uint64_t bdesc = tc_desc(bstart, 128, b_sbo), adesc = tc_desc(astart, 128, a_sbo);
							tc_mma_i8(tmem + acc * SD_N, adesc, bdesc, idesc, (kb | ks) != 0 ? 1u : 0u);
						}
						tc_commit(bar(BAR_B_EMPTY + stage)); /* the ring slot is free once these MMAs have read it */
					}
					tc_commit(bar(BAR_T_FULL + acc));
				}
				tc_commit(bar(BAR_A_EMPTY + abuf));
			}
		}
	} else if (warp < 2 + SD_PROD_WARPS) {
		/* ===== A operand producers: thread = window row; tile i + 1 is built while tile i is multiplied and summed ===== */
		const int set = (warp - 2) >> 2;
		const int t = (tid - 64) & (SD_ROWS - 1);
		uint16_t* hist = sHist + set * (64 * SD_ROWS) + t;
		uint32_t* hist_word = reinterpret_cast<uint32_t*>(sHist + set * (64 * SD_ROWS)) + (t >> 1);
		const uint32_t hist_inc = 1u << (16 * (t & 1));
		const bool fast20 = a.patch_w == 20 && a.patch_h == 20;
		uint32_t tile_it = (uint32_t)set;
		for (int64_t tile = blockIdx.x + (int64_t)set * gridDim.x; tile < ntiles; tile += (int64_t)SD_PROD_SETS * gridDim.x, tile_it += SD_PROD_SETS) {
			const uint32_t abuf = tile_it % SD_ABUF;
			mbar_wait<128>(bar(BAR_A_EMPTY + abuf), ((tile_it / SD_ABUF) & 1) ^ 1);
			const int64_t g = tile * SD_ROWS + t;
			uint8_t* arow = sA + abuf * a_tile + (t >> 3) * (chunks * 128) + (t & 7) * 16;
			int xx = 0;
			if (g < a.total) {
				if (MODE == 1) xx = produce_vector(a, arow, g);
				else if (fast20) xx = produce_window_fast<20, 20>(a, sLayers, hist, hist_word, hist_inc, arow, g);
				else xx = produce_window(a, sLayers, hist, arow, g);
			} else {
				for (int ch = 0; ch < chunks; ++ch) *reinterpret_cast<uint4*>(arow + ch * 128) = make_uint4(0u, 0u, 0u, 0u);
			}
			/* the ring is deeper than the producers can run ahead of the epilogue (SD_ABUF tiles + SD_ACC blocks) */
			sXX[(tile_it % SD_XX_RING) * SD_ROWS + t] = xx;
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy stores -> visible to the tensor core */
			mbar_arrive(bar(BAR_A_FULL + abuf));
		}
	} else {
		/* ===== epilogue: SD_EPI_SPLIT threads per window row, each owning every SD_EPI_SPLIT-th group of 32 support vectors ===== */
		const int e = warp - (2 + SD_PROD_WARPS);
		const int part = e >> 2, q = warp & 3;    /* a warp reads the tensor-memory lanes 32 * (warp id % 4) .. + 31 */
		const int row = q * 32 + lane;
		const int s_shift = a.s.shift;
		const uint32_t lo_mask = (1u << s_shift) - 1u;
		const int tab_last = a.s.tab_n - 1;
		const double c1 = a.s.poly[0], c2 = a.s.poly[1], c3 = a.s.poly[2], c4 = a.s.poly[3];
		const int* __restrict__ ssq = a.s.ssq;
		const double* __restrict__ coef = a.s.coef;
		uint32_t acc_it = 0, tile_it = 0;
		for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
			int xx = 0;
			double dacc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
			for (int nt = 0; nt < NT; ++nt, ++acc_it) {
				const uint32_t acc = acc_it % SD_ACC;
				mbar_wait<64>(bar(BAR_T_FULL + acc), (acc_it / SD_ACC) & 1);
				tc_fence_after();
				if (nt == 0) xx = sXX[(tile_it % SD_XX_RING) * SD_ROWS + row]; /* written before the tile's first MMA was issued */
				const int sv0 = ((nt + nt_rot) % NT) * SD_N + part * SD_EPI_COLS;
				const int4* __restrict__ s4p = reinterpret_cast<const int4*>(ssq + sv0);       /* warp-uniform vector loads */
				const double2* __restrict__ c2p = reinterpret_cast<const double2*>(coef + sv0);
#pragma unroll 1
				for (int h16 = 0; h16 < SD_EPI_COLS / 16; ++h16) {
					uint32_t v[16];
					tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + acc * SD_N + part * SD_EPI_COLS + h16 * 16, v);
					if (h16 == SD_EPI_COLS / 16 - 1) { /* accumulators are in registers: hand the buffer back to the MMA warp */
						tc_fence_before();
						__syncwarp();
						if (lane == 0) mbar_arrive(bar(BAR_T_EMPTY + acc));
					}
					/* 8 support vectors in lockstep: the float64 pipe has a long latency, so the eight Horner chains, the
					 * table loads and the eight partial sums are kept independent of each other */
#pragma unroll
					for (int j8 = 0; j8 < 2; ++j8) {
						const int4 sa = __ldg(s4p + h16 * 4 + 2 * j8), sb = __ldg(s4p + h16 * 4 + 2 * j8 + 1);
						const int ss[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
						double cc[8];
#pragma unroll
						for (int k = 0; k < 4; ++k) {
							const double2 c = __ldg(c2p + h16 * 8 + 4 * j8 + k);
							cc[2 * k] = c.x; cc[2 * k + 1] = c.y;
						}
						int ssd[8];
						double tv[8], l[8], p[8];
#pragma unroll
						for (int k = 0; k < 8; ++k) ssd[k] = xx + ss[k] - 2 * (int)v[8 * j8 + k];
#pragma unroll
						for (int k = 0; k < 8; ++k) {
							int hi = ssd[k] >> s_shift;
							if (CLAMP) hi = min(hi, tab_last);
							tv[k] = sTab[hi];
						}
#pragma unroll
						for (int k = 0; k < 8; ++k) /* exact int -> double without the conversion unit: 2^52 + lo as raw bits, minus 2^52 */
							l[k] = __dsub_rn(__hiloint2double(0x43300000, (int)(ssd[k] & lo_mask)), 4503599627370496.0);
#pragma unroll
						for (int k = 0; k < 8; ++k) p[k] = fma(l[k], c4, c3);
#pragma unroll
						for (int k = 0; k < 8; ++k) p[k] = fma(l[k], p[k], c2);
#pragma unroll
						for (int k = 0; k < 8; ++k) p[k] = fma(l[k], p[k], c1);
#pragma unroll
						for (int k = 0; k < 8; ++k) p[k] = fma(l[k], p[k], 1.0);
#pragma unroll
						for (int k = 0; k < 8; ++k) dacc[k] = fma(cc[k], __dmul_rn(tv[k], p[k]), dacc[k]); /* RbfKernel.hpp:39, SvmClassifier.cpp:58 */
					}
				}
			}
			const double dist = __dadd_rn(__dadd_rn(__dadd_rn(dacc[0], dacc[1]), __dadd_rn(dacc[2], dacc[3])),
					__dadd_rn(__dadd_rn(dacc[4], dacc[5]), __dadd_rn(dacc[6], dacc[7])));
			/* the SD_EPI_SPLIT partial sums of a row, added in a fixed order */
			double* pbuf = sPart + (tile_it & 1) * (SD_EPI_SPLIT * SD_ROWS);
			pbuf[part * SD_ROWS + row] = dist;
			asm volatile("bar.sync 1, %0;" :: "n"(SD_EPI_WARPS * 32) : "memory");
			const int64_t g = tile * SD_ROWS + row;
			if (part == 0 && g < a.total) {
				double d = a.s.neg_bias; /* SvmClassifier.cpp:56: double distance = -bias */
#pragma unroll
				for (int i = 0; i < SD_EPI_SPLIT; ++i) d = __dadd_rn(d, pbuf[i * SD_ROWS + row]);
				a.distance_out[g] = d;
				if (a.pos_count && d >= (double)a.s.threshold) { /* SvmClassifier::classify (SvmClassifier.cpp:44-46) */
					const int slot = atomicAdd(a.pos_count, 1);
					if (slot < a.pos_cap) {
						DensePositive dp;
						dp.row = g; dp.distance = d;
						a.pos[slot] = dp;
					}
				}
			}
		}
	}

	tc_fence_before();
	__syncthreads();
	if (warp == 1) {
		tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(SD_TMEM_COLS) : "memory");
	}
}

size_t sd_smem_bytes(const DevSvmDense& s) {
	const size_t a_tile = (size_t)(SD_ROWS / 8) * s.chunks * 128;
	return SD_ABUF * a_tile + (size_t)SD_STAGES * SD_B_STAGE_BYTES + SD_PROD_SETS * 64 * SD_ROWS * sizeof(uint16_t) + (size_t)s.tab_n * 8
			+ 2 * SD_EPI_SPLIT * SD_ROWS * sizeof(double) + SD_XX_RING * SD_ROWS * sizeof(int) + FDB_MAX_LAYERS * sizeof(DevLayer)
			+ BAR_COUNT * 8 + 16;
}

int g_sd_sms = 0;
size_t g_sd_smem_max = 0;

} // namespace

int svm_dense_configure() {
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	int v = 0;
	if (e == cudaSuccess) e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
	g_sd_sms = v;
	if (e == cudaSuccess) e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
	g_sd_smem_max = (size_t)v;
	if (e == cudaSuccess) e = cudaFuncSetAttribute(svm_dense_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(svm_dense_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(svm_dense_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(svm_dense_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
	return (int)e;
}

/* host-side preparation of the tensor-core form of an u8 RBF SVM; returns false when the model does not fit
 * (shared memory for the window tiles / the exp table) - callers then keep the per-window kernel of svm.cu */
bool svm_dense_build(const uint8_t* sv, const float* coef, int num_sv, int dim, double gamma, float bias, float threshold,
		SvmDenseHost* out) {
	if (!(gamma > 0.0) || num_sv < 1 || dim < 1 || (int64_t)dim * 65025 >= ((int64_t)1 << 31)) return false;
	SvmDenseHost& h = *out;
	DevSvmDense& d = h.dev;
	d.dim = dim;
	d.chunks = 2 * ((dim + 31) / 32);
	d.num_sv_pad = ((num_sv + 2 * SD_N - 1) / (2 * SD_N)) * (2 * SD_N); /* an even number of accumulator blocks */
	d.threshold = threshold;
	d.neg_bias = -(double)bias;
	/* exp table: ssd = hi * 2^shift + lo with gamma * 2^shift <= 2^-7 */
	int shift = 0;
	while (shift < 24 && gamma * (double)((int64_t)2 << shift) <= 0.0078125) ++shift;
	if (gamma > 0.0078125) shift = 0;
	const int64_t max_ssd = (int64_t)dim * 65025;
	const double under = 746.0 / gamma; /* exp(-746) == 0 in float64 */
	const int64_t top = under < (double)max_ssd ? (int64_t)under + 1 : max_ssd;
	const int64_t entries = (top >> shift) + 2;
	if (gamma > 0.0078125 || entries > 4096) return false; /* TODO(large gamma): needs a finer table than shared memory holds */
	d.shift = shift;
	d.tab_n = (int)entries;
	h.tab.resize((size_t)entries);
	for (int64_t i = 0; i < entries; ++i) h.tab[(size_t)i] = std::exp(-gamma * (double)(i << shift));
	d.clamp = top < max_ssd ? 1 : 0;
	if (d.clamp) h.tab[(size_t)entries - 1] = 0.0; /* everything past the underflow point */
	long double f = 1.0L, gk = 1.0L;
	for (int k = 1; k <= 4; ++k) { f *= k; gk *= -(long double)gamma; d.poly[k - 1] = (double)(gk / f); }
	if (sd_smem_bytes(d) > (g_sd_smem_max ? g_sd_smem_max : (size_t)232448)) return false;
	/* support vectors as core matrices: [n block][k block][group of 8 rows][k chunk][row][16 bytes] */
	const int NT = d.num_sv_pad / SD_N, KB = (d.chunks + SD_KCH - 1) / SD_KCH;
	h.b_blocks.assign((size_t)d.num_sv_pad * d.chunks * 16, 0);
	size_t o = 0;
	for (int nt = 0; nt < NT; ++nt)
		for (int kb = 0; kb < KB; ++kb) {
			const int cb = std::min(SD_KCH, d.chunks - kb * SD_KCH);
			for (int grp = 0; grp < SD_N / 8; ++grp)
				for (int c = 0; c < cb; ++c)
					for (int r = 0; r < 8; ++r, o += 16) {
						const int i = nt * SD_N + grp * 8 + r;
						if (i >= num_sv) continue;
						const int k0 = (kb * SD_KCH + c) * 16;
						for (int k = 0; k < 16 && k0 + k < dim; ++k) h.b_blocks[o + k] = sv[(size_t)i * dim + k0 + k];
					}
		}
	h.ssq.assign((size_t)d.num_sv_pad, 0);
	h.coef.assign((size_t)d.num_sv_pad, 0.0);
	for (int i = 0; i < num_sv; ++i) {
		int s = 0;
		for (int k = 0; k < dim; ++k) { const int v = sv[(size_t)i * dim + k]; s += v * v; }
		h.ssq[(size_t)i] = s;
		h.coef[(size_t)i] = (double)coef[i];
	}
	return true;
}

bool svm_dense_enabled() {
	const char* e = std::getenv("FDB_SVM_DENSE");
	return !(e && e[0] == '0');
}

static int dense_grid(int64_t total) {
	const int64_t ntiles = (total + SD_ROWS - 1) / SD_ROWS;
	return (int)std::min<int64_t>(ntiles, g_sd_sms > 0 ? g_sd_sms : 148);
}

void launch_svm_dense_windows(cudaStream_t st, const DevSvmDense& s, int patch_w, int patch_h, int step_x, int step_y,
		const uint8_t* frames, int W, int H, int n_frames, const uint8_t* arena, int64_t arena_stride, const DevLayer* layers,
		int n_layers, int64_t windows_per_frame, double* distance_out, int* pos_count, DensePositive* pos, int pos_cap) {
	const int64_t total = windows_per_frame * n_frames;
	if (total <= 0) return;
	SvmDenseArgs a{};
	a.s = s; a.patch_w = patch_w; a.patch_h = patch_h; a.step_x = step_x; a.step_y = step_y;
	a.frames = frames; a.W = W; a.H = H; a.arena = arena; a.arena_stride = arena_stride;
	a.layers = layers; a.n_layers = n_layers; a.windows_per_frame = windows_per_frame;
	a.total = total; a.distance_out = distance_out; a.pos_count = pos_count; a.pos = pos; a.pos_cap = pos_cap;
	if (s.clamp) svm_dense_kernel<0, true><<<dense_grid(total), SD_THREADS, sd_smem_bytes(s), st>>>(a);
	else svm_dense_kernel<0, false><<<dense_grid(total), SD_THREADS, sd_smem_bytes(s), st>>>(a);
}

void launch_svm_dense_vectors(cudaStream_t st, const DevSvmDense& s, const uint8_t* vectors, int64_t n, double* distance_out) {
	if (n <= 0) return;
	SvmDenseArgs a{};
	a.s = s; a.vectors = vectors; a.total = n; a.distance_out = distance_out; a.windows_per_frame = n;
	if (s.clamp) svm_dense_kernel<1, true><<<dense_grid(n), SD_THREADS, sd_smem_bytes(s), st>>>(a);
	else svm_dense_kernel<1, false><<<dense_grid(n), SD_THREADS, sd_smem_bytes(s), st>>>(a);
}

} // namespace fdb
