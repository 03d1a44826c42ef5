/*
 * api.cu - the C ABI of include/fdb200.h: contexts, model upload, the detector pipeline.
 *
 * Host side of the product (C++), mirroring the reference's object graph:
 *   fdb_wvm      <- ProbabilisticWvmClassifier -> WvmClassifier   (libClassification)
 *   fdb_svm      <- ProbabilisticSvmClassifier -> SvmClassifier -> RbfKernel
 *   fdb_detector <- FiveStageSlidingWindowDetector( SlidingWindowDetector( pwvm,
 *                     DirectPyramidFeatureExtractor( ImagePyramid + GrayscaleFilter, HistEq64Filter ) ),
 *                     OverlapElimination, psvm )                   (ffpDetectApp.cpp:391-425)
 * The per-window work runs in the CUDA kernels of pyramid.cu / wvm.cu / svm.cu; the few
 * surviving candidates per frame are post-processed in hostpost.cpp.  No CPU fallback exists.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "fdb_internal.h"
#include "wvm_device.h"
#include "api_types.h"
#include "features_device.h"
#include "wvm_group.h"

namespace fdb {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int status, const std::string& msg) { g_last_error = msg; return status; }
int api_exception() noexcept {
	try {
		try { throw; }
		catch (const std::bad_alloc&) { return fail(FDB_ERR_RUNTIME, "out of host memory"); }
		catch (const std::exception& e) { return fail(FDB_ERR_RUNTIME, std::string("internal error: ") + e.what()); }
		catch (...) { return fail(FDB_ERR_RUNTIME, "internal error: unknown exception"); }
	} catch (...) { return FDB_ERR_RUNTIME; } /* the message itself could not be allocated */
}

} // namespace fdb

using namespace fdb;

namespace {

} // namespace

extern "C" {

int fdb_abi_version(void) { return FDB_ABI_VERSION; }
const char* fdb_last_error(void) { return g_last_error.c_str(); }
const char* fdb_status_string(int s) {
	switch (s) {
	case FDB_OK: return "ok";
	case FDB_ERR_INVALID_ARGUMENT: return "invalid argument";
	case FDB_ERR_RUNTIME: return "runtime error";
	case FDB_ERR_CUDA: return "CUDA error";
	case FDB_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
	case FDB_ERR_UNSUPPORTED: return "unsupported model or geometry";
	case FDB_ERR_OVERFLOW: return "result buffer overflow";
	default: return "unknown status";
	}
}

int fdb_ctx_create(int device, fdb_ctx** out) try {
	if (!out) return fail(FDB_ERR_INVALID_ARGUMENT, "out is null");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(FDB_ERR_NO_DEVICE, std::string("no CUDA device available: ") + cudaGetErrorString(e) + " (fdb200 has no CPU fallback)");
	if (device < 0) CUDA_TRY(cudaGetDevice(&device));
	if (device >= count) return fail(FDB_ERR_INVALID_ARGUMENT, "device index out of range");
	CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10)
		return fail(FDB_ERR_NO_DEVICE, "fdb200 kernels are built for sm_100a only");
	fdb_ctx* c = new fdb_ctx;
	c->device = device;
	CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	for (cudaEvent_t& e : c->ev) CUDA_TRY(cudaEventCreate(&e));
	if (wvm_configure() != 0 || svm_configure() != 0 || svm_dense_configure() != 0 || group_configure_all() != 0 || feature_configure() != 0) {
		cudaStreamDestroy(c->stream); delete c;
		return fail(FDB_ERR_CUDA, "cudaFuncSetAttribute failed: libfdb200 kernels not loadable on this device");
	}
	*out = c;
	return FDB_OK;
} FDB_API_CATCH

void fdb_ctx_destroy(fdb_ctx* c) {
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	for (cudaEvent_t e : c->ev) if (e) cudaEventDestroy(e);
	cudaStreamDestroy(c->stream);
	delete c;
}

void* fdb_ctx_stream(fdb_ctx* c) { return c ? (void*)c->stream : nullptr; }

int fdb_ctx_synchronize(fdb_ctx* c) try {
	int s = check_ctx(c); if (s) return s;
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return FDB_OK;
} FDB_API_CATCH

int fdb_ctx_timer_start(fdb_ctx* c) try {
	int s = check_ctx(c); if (s) return s;
	CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
	return FDB_OK;
} FDB_API_CATCH

int fdb_ctx_timer_stop(fdb_ctx* c, double* elapsed_ms) try {
	int s = check_ctx(c); if (s) return s;
	CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
	CUDA_TRY(cudaEventSynchronize(c->ev[1]));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
	if (elapsed_ms) *elapsed_ms = ms;
	return FDB_OK;
} FDB_API_CATCH

int64_t fdb_ctx_launch_count(fdb_ctx* c) { return c ? c->launches : 0; }

int fdb_host_alloc(size_t bytes, void** out) try {
	if (!out) return fail(FDB_ERR_INVALID_ARGUMENT, "out is null");
	CUDA_TRY(cudaMallocHost(out, std::max<size_t>(bytes, 16)));
	return FDB_OK;
} FDB_API_CATCH
void fdb_host_free(void* p) { if (p) cudaFreeHost(p); }

/* ---------------------------------------------------------------------------------------------
 * WVM
 * ------------------------------------------------------------------------------------------- */
int fdb_wvm_create(fdb_ctx* ctx, const fdb_wvm_desc* d, fdb_wvm** out) try {
	int s = check_ctx(ctx); if (s) return s;
	if (!d || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	const int n = d->num_lin_filters, w = d->filter_size_x, h = d->filter_size_y;
	if (n < 1 || w < 1 || h < 1 || d->num_filters_per_level < 1)
		return fail(FDB_ERR_INVALID_ARGUMENT, "WVM: empty model");
	if (n > FDB_MAX_FILTERS) return fail(FDB_ERR_UNSUPPORTED, "WVM: more than 512 filters");
	if (d->num_filters_per_level > FDB_MAX_PER_LEVEL) return fail(FDB_ERR_UNSUPPORTED, "WVM: more than 64 filters per level");
	const int npix = w * h, nwords = (npix + 3) / 4;
	if ((size_t)(32 + nwords) * WVM_THREADS * 4 > 200 * 1024)
		return fail(FDB_ERR_UNSUPPORTED, "WVM: patch too large for the shared-memory layout");
	/* rectangle coverage masks; exactness envelope of the float integral-image arithmetic */
	std::vector<int> val_off(n), mask_off(n);
	std::vector<uint32_t> masks;
	int slot = 0; size_t rofs = 0;
	int max_nv = 0;
	for (int f = 0; f < n; ++f) {
		const int cntval = d->area_cntval[f];
		max_nv = std::max(max_nv, cntval - 1);
		if (cntval < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "WVM: filter without grey values");
		if (cntval > 256) return fail(FDB_ERR_UNSUPPORTED, "WVM: more than 255 rectangle grey values per filter");
		val_off[f] = slot;
		mask_off[f] = (int)masks.size();
		const int nv = cntval - 1;
		std::vector<int> cover((size_t)npix * std::max(nv, 1), 0);
		int64_t mass = 0;
		for (int v = 1; v < cntval; ++v)
			for (int r = 0; r < d->area_cntrec[slot + v]; ++r) {
				const fdb_rect4& q = d->area_rec[rofs++];
				if (q.x1 < 0 || q.y1 < 0 || q.x2 >= w || q.y2 >= h || q.x1 > q.x2 || q.y1 > q.y2)
					return fail(FDB_ERR_INVALID_ARGUMENT, "WVM: rectangle outside the filter window");
				for (int y = q.y1; y <= q.y2; ++y)
					for (int x = q.x1; x <= q.x2; ++x) cover[(size_t)(v - 1) * npix + y * w + x]++;
				mass += (int64_t)(q.x2 - q.x1 + 1) * (q.y2 - q.y1 + 1);
			}
		if (mass * 255 >= (1 << 24))
			return fail(FDB_ERR_UNSUPPORTED, "WVM: rectangle mass exceeds the exact float32 integer range of the reference arithmetic");
		for (int j = 0; j < nwords; ++j)
			for (int v = 0; v < nv; ++v) {
				uint32_t word = 0;
				for (int k = 0; k < 4; ++k) {
					const int px = 4 * j + k;
					const int cnt = px < npix ? cover[(size_t)v * npix + px] : 0;
					if (cnt > 255) return fail(FDB_ERR_UNSUPPORTED, "WVM: more than 255 overlapping rectangles on a pixel");
					word |= (uint32_t)cnt << (8 * k);
				}
				masks.push_back(word);
			}
		slot += cntval;
	}
	fdb_wvm* m = new fdb_wvm;
	m->ctx = ctx;
	m->logistic_a = d->logistic_a; m->logistic_b = d->logistic_b;
	m->thresholds_from_file.assign(d->hierarchical_thresholds, d->hierarchical_thresholds + n);
	DevWvm& dv = m->dev;
	dv.fsx = w; dv.fsy = h; dv.nwords = nwords;
	dv.num_lin = n; dv.per_level = d->num_filters_per_level;
	dv.num_used = (d->num_used_filters > n || d->num_used_filters == 0) ? n : d->num_used_filters; /* WvmClassifier.cpp:151-158 */
	dv.step_x = dv.step_y = 1;
	dv.basis_param = d->basis_param;
	float* fp; double* dp; int* ip; uint32_t* up;
#define UP(ptr, count, field, tmp) do { s = upload(ptr, (size_t)(count), &tmp, m->owned); if (s) { free_all(m->owned); delete m; return s; } dv.field = tmp; } while (0)
	UP(d->lin_thresholds, n, lin_thresholds, fp);
	UP(d->hk_weights, (size_t)n * (n + 1) / 2, hk_weights, fp);
	UP(d->app_rsv_convol, n, app_rsv_convol, dp);
	UP(d->area_cntval, n, cntval, ip);
	UP(val_off.data(), n, val_off, ip);
	UP(d->area_val, slot, val, dp);
	UP(masks.data(), masks.size(), masks, up);
	UP(mask_off.data(), n, mask_off, ip);
	{ /* rectangle table for the integral-image evaluation of the deep kernel */
		std::vector<uint2> rc;
		std::vector<int> roff(n + 1, 0);
		int sl = 0; size_t ro = 0;
		for (int f = 0; f < n; ++f) {
			roff[f] = (int)rc.size();
			for (int v = 1; v < d->area_cntval[f]; ++v)
				for (int r = 0; r < d->area_cntrec[sl + v]; ++r) {
					const fdb_rect4& q = d->area_rec[ro++];
					uint2 e;
					e.x = (uint32_t)q.x1 | ((uint32_t)q.y1 << 8) | ((uint32_t)q.x2 << 16) | ((uint32_t)q.y2 << 24);
					e.y = (uint32_t)(v - 1);
					rc.push_back(e);
				}
			sl += d->area_cntval[f];
		}
		roff[n] = (int)rc.size();
		uint2* rp; 
		s = upload(rc.data(), rc.size(), &rp, m->owned); if (s) { free_all(m->owned); delete m; return s; }
		dv.rects = rp;
		UP(roff.data(), n + 1, rect_off, ip);
	}
	dv.bfrag = nullptr; dv.btc = nullptr;
	dv.hk_weights_t = nullptr; dv.hk_t_off = nullptr;
	if (max_nv <= 4 && n > WVM_KA && group_supported(w, h)) {
		/* rectangle coverage counts of the first WVM_KA filters in mma.m16n8k32 B-fragment order (wvm_group.cu): k-step s is
		 * patch row s padded to 8 words (16-wide windows: rows 2 s and 2 s + 1); lane (g, t) holds, for n-tile nt, the words
		 * 4 h + t of that row (16-wide: word t of row 2 s + h), h = 0, 1, of filter 2 nt + (g >> 2), grey value g & 3 */
		const int wpr = w / 4, rpk = w <= 16 ? 2 : 1, ks = h / rpk;
		std::vector<uint32_t> bf((size_t)ks * 32 * 8, 0u);
		for (int s2 = 0; s2 < ks; ++s2)
			for (int lane = 0; lane < 32; ++lane)
				for (int nt = 0; nt < 4; ++nt)
					for (int hh = 0; hh < 2; ++hh) {
						const int g = lane >> 2, t = lane & 3;
						const int f = 2 * nt + (g >> 2), v = g & 3;
						const int row = rpk == 2 ? 2 * s2 + hh : s2, c4 = rpk == 2 ? t : 4 * hh + t;
						const int nv = d->area_cntval[f] - 1;
						if (c4 < wpr && v < nv)
							bf[((size_t)s2 * 32 + lane) * 8 + nt * 2 + hh] = masks[(size_t)mask_off[f] + (size_t)(row * wpr + c4) * nv + v];
					}
		uint32_t* bp;
		s = upload(bf.data(), bf.size(), &bp, m->owned); if (s) { free_all(m->owned); delete m; return s; }
		dv.bfrag = reinterpret_cast<const uint4*>(bp);
		/* the same counts as the B operand of tcgen05.mma (wvm_group_tc.cu): per k-step 32 columns (4 * filter + grey value) x 32
		 * operand bytes as UMMA core matrices, K-major without swizzle: byte (n, k) at (n / 8) * 256 + (k / 16) * 128 + (n % 8) * 16 + k % 16 */
		std::vector<uint32_t> bt((size_t)ks * 256, 0u);
		for (int s2 = 0; s2 < ks; ++s2)
			for (int col = 0; col < 32; ++col)
				for (int c = 0; c < 8; ++c) { /* word c of the k-step's 32 operand bytes */
					const int f = col >> 2, v = col & 3;
					const int row = rpk == 2 ? 2 * s2 + (c >> 2) : s2, c4 = rpk == 2 ? (c & 3) : c;
					const int nv = d->area_cntval[f] - 1;
					if (c4 < wpr && v < nv)
						bt[(size_t)s2 * 256 + (size_t)((col >> 3) * 256 + (c >> 2) * 128 + (col & 7) * 16 + (c & 3) * 4) / 4] =
								masks[(size_t)mask_off[f] + (size_t)(row * wpr + c4) * nv + v];
				}
		uint32_t* btp;
		s = upload(bt.data(), bt.size(), &btp, m->owned); if (s) { free_all(m->owned); delete m; return s; }
		dv.btc = reinterpret_cast<const uint8_t*>(btp);
		/* weights of the deep kernel's rounds, transposed for coalesced loads; one zero group of padding per round (prefetch) */
		std::vector<float> wt;
		std::vector<int> toff;
		for (int base = WVM_KA; base < n; base += 32) {
			toff.push_back((int)(wt.size() / 4));
			const int cnt = std::min(32, n - base), groups = (base + cnt + 3) / 4 + 1;
			const size_t at = wt.size();
			wt.resize(at + (size_t)groups * 32 * 4, 0.f);
			for (int lane = 0; lane < cnt; ++lane) {
				const int l = base + lane;
				for (int p = 0; p <= l; ++p) wt[at + ((size_t)(p / 4) * 32 + lane) * 4 + (p & 3)] = d->hk_weights[(size_t)l * (l + 1) / 2 + p];
			}
		}
		UP(wt.data(), wt.size(), hk_weights_t, fp);
		UP(toff.data(), toff.size(), hk_t_off, ip);
	}
#undef UP
	s = dev_alloc(&m->d_thresholds, (size_t)n, m->owned);
	if (s) { free_all(m->owned); delete m; return s; }
	dv.thresholds = m->d_thresholds;
	s = fdb_wvm_set_limit_reliability_filter(m, d->limit_reliability_filter);
	if (s) { free_all(m->owned); delete m; return s; }
	*out = m;
	return FDB_OK;
} FDB_API_CATCH

void fdb_wvm_destroy(fdb_wvm* m) {
	if (!m) return;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	free_all(m->owned);
	delete m;
}

int fdb_wvm_set_limit_reliability_filter(fdb_wvm* m, float value) try {
	if (!m) return fail(FDB_ERR_INVALID_ARGUMENT, "null wvm");
	int s = check_ctx(m->ctx); if (s) return s;
	/* WvmClassifier.cpp:165-181 */
	m->limit = value;
	m->thresholds = m->thresholds_from_file;
	if (value != 0.0f)
		for (float& t : m->thresholds) t = t + value;
	CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
	CUDA_TRY(cudaMemcpy(m->d_thresholds, m->thresholds.data(), sizeof(float) * m->thresholds.size(), cudaMemcpyHostToDevice));
	return FDB_OK;
} FDB_API_CATCH

int fdb_wvm_get_probability(fdb_wvm* m, const uint8_t* patches, int64_t n, int32_t* level_out, float* fout_out,
		double* prob_out, uint8_t* pos_out) try {
	if (!m) return fail(FDB_ERR_INVALID_ARGUMENT, "null wvm");
	int s = check_ctx(m->ctx); if (s) return s;
	if (n < 0 || (n > 0 && !patches)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad patch batch");
	if (n == 0) return FDB_OK;
	if (n > (1 << 30)) return fail(FDB_ERR_INVALID_ARGUMENT, "batch too large");
	const size_t npix = (size_t)m->dev.fsx * m->dev.fsy;
	std::vector<void*> tmp;
	uint8_t* d_p; fdb_window_score* d_s;
	s = dev_alloc(&d_p, npix * (size_t)n, tmp); if (s) { free_all(tmp); return s; }
	s = dev_alloc(&d_s, (size_t)n, tmp); if (s) { free_all(tmp); return s; }
	cudaStream_t st = m->ctx->stream;
	std::vector<fdb_window_score> host((size_t)n);
	cudaError_t e = cudaMemcpyAsync(d_p, patches, npix * (size_t)n, cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess) {
		launch_wvm_patches(st, m->dev, d_p, (int)n, d_s);
		m->ctx->launches++;
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(host.data(), d_s, sizeof(fdb_window_score) * (size_t)n, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	free_all(tmp);
	if (e != cudaSuccess) return fail(FDB_ERR_CUDA, std::string("wvm_get_probability: ") + cudaGetErrorString(e));
	for (int64_t i = 0; i < n; ++i) {
		const fdb_window_score& r = host[(size_t)i];
		if (level_out) level_out[i] = r.level;
		if (fout_out) fout_out[i] = r.fout;
		if (prob_out) prob_out[i] = wvm_probability(m->logistic_a, m->logistic_b, r.fout);
		if (pos_out) pos_out[i] = (r.level + 1 == m->dev.num_lin && r.fout >= m->thresholds[(size_t)r.level]) ? 1 : 0;
	}
	return FDB_OK;
} FDB_API_CATCH

/* ---------------------------------------------------------------------------------------------
 * SVM
 * ------------------------------------------------------------------------------------------- */
int fdb_svm_create(fdb_ctx* ctx, const fdb_svm_desc* d, fdb_svm** out) try {
	int s = check_ctx(ctx); if (s) return s;
	if (!d || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	if (d->kernel < FDB_KERNEL_RBF || d->kernel > FDB_KERNEL_LINEAR) return fail(FDB_ERR_INVALID_ARGUMENT, "SVM: unknown kernel kind");
	if (d->kernel == FDB_KERNEL_POLYNOMIAL && d->poly_degree < 0) return fail(FDB_ERR_INVALID_ARGUMENT, "SVM: negative polynomial degree");
	if (d->num_sv < 1 || d->dim < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "SVM: empty model");
	if (d->sv_type != FDB_SV_U8 && d->sv_type != FDB_SV_F32) return fail(FDB_ERR_INVALID_ARGUMENT, "SVM: bad sv_type");
	if ((size_t)d->dim * 4 > 96 * 1024) return fail(FDB_ERR_UNSUPPORTED, "SVM: feature vector too long");
	fdb_svm* m = new fdb_svm;
	m->ctx = ctx;
	m->logistic_a = d->logistic_a; m->logistic_b = d->logistic_b;
	DevSvm& dv = m->dev;
	dv.num_sv = d->num_sv; dv.dim = d->dim; dv.sv_type = d->sv_type;
	dv.nwords = (d->dim + 3) / 4;
	dv.kernel = d->kernel; dv.gamma = d->gamma; dv.bias = d->bias; dv.threshold = d->threshold;
	dv.poly_alpha = d->poly_alpha; dv.poly_constant = d->poly_constant; dv.poly_degree = d->poly_degree;
	dv.sv_words = nullptr; dv.sv_f32 = nullptr;
	float* fp;
	s = upload(d->coefficients, (size_t)d->num_sv, &fp, m->owned);
	dv.coef = fp;
	if (!s) {
		if (d->sv_type == FDB_SV_U8) {
			const uint8_t* sv = (const uint8_t*)d->support_vectors;
			std::vector<uint32_t> tr((size_t)dv.nwords * d->num_sv, 0);
			for (int i = 0; i < d->num_sv; ++i)
				for (int k = 0; k < d->dim; ++k)
					tr[(size_t)(k >> 2) * d->num_sv + i] |= (uint32_t)sv[(size_t)i * d->dim + k] << (8 * (k & 3));
			uint32_t* up;
			s = upload(tr.data(), tr.size(), &up, m->owned);
			dv.sv_words = up;
			/* tensor-core form for whole-batch evaluation (svm_dense.cu) */
			SvmDenseHost dh;
			if (!s && d->kernel == FDB_KERNEL_RBF && svm_dense_build(sv, d->coefficients, d->num_sv, d->dim, d->gamma, d->bias, d->threshold, &dh)) {
				uint8_t* bb; int* sq; double* cf; double* tb;
				s = upload(dh.b_blocks.data(), dh.b_blocks.size(), &bb, m->owned);
				if (!s) s = upload(dh.ssq.data(), dh.ssq.size(), &sq, m->owned);
				if (!s) s = upload(dh.coef.data(), dh.coef.size(), &cf, m->owned);
				if (!s) s = upload(dh.tab.data(), dh.tab.size(), &tb, m->owned);
				if (!s) {
					m->dense = dh.dev;
					m->dense.b_blocks = bb; m->dense.ssq = sq; m->dense.coef = cf; m->dense.exp_tab = tb;
					m->has_dense = true;
				}
			}
		} else {
			const float* sv = (const float*)d->support_vectors;
			std::vector<float> tr((size_t)d->dim * d->num_sv);
			for (int i = 0; i < d->num_sv; ++i)
				for (int k = 0; k < d->dim; ++k) tr[(size_t)k * d->num_sv + i] = sv[(size_t)i * d->dim + k];
			float* fp2;
			s = upload(tr.data(), tr.size(), &fp2, m->owned);
			dv.sv_f32 = fp2;
		}
	}
	if (s) { free_all(m->owned); delete m; return s; }
	*out = m;
	return FDB_OK;
} FDB_API_CATCH

void fdb_svm_destroy(fdb_svm* m) {
	if (!m) return;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	free_all(m->owned);
	delete m;
}

int fdb_svm_set_threshold(fdb_svm* m, float t) try {
	if (!m) return fail(FDB_ERR_INVALID_ARGUMENT, "null svm");
	m->dev.threshold = t;
	m->dense.threshold = t;
	return FDB_OK;
} FDB_API_CATCH

int fdb_svm_has_dense(const fdb_svm* m) { return m && m->has_dense && fdb::svm_dense_enabled() ? 1 : 0; }

int fdb_svm_get_probability(fdb_svm* m, const void* vectors, int64_t n, double* dist_out, double* prob_out, uint8_t* pos_out) try {
	if (!m) return fail(FDB_ERR_INVALID_ARGUMENT, "null svm");
	int s = check_ctx(m->ctx); if (s) return s;
	if (n < 0 || (n > 0 && !vectors)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad vector batch");
	if (n == 0) return FDB_OK;
	if (n > (1 << 30)) return fail(FDB_ERR_INVALID_ARGUMENT, "batch too large");
	const size_t es = m->dev.sv_type == FDB_SV_U8 ? 1 : 4;
	const size_t bytes = es * (size_t)m->dev.dim * (size_t)n;
	std::vector<void*> tmp;
	uint8_t* d_v; double* d_d;
	s = dev_alloc(&d_v, bytes, tmp); if (s) { free_all(tmp); return s; }
	s = dev_alloc(&d_d, (size_t)n, tmp); if (s) { free_all(tmp); return s; }
	cudaStream_t st = m->ctx->stream;
	std::vector<double> host((size_t)n);
	cudaError_t e = cudaMemcpyAsync(d_v, vectors, bytes, cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess) {
		if (m->has_dense && svm_dense_enabled() && n >= FDB_SVM_DENSE_MIN_VECTORS)
			launch_svm_dense_vectors(st, m->dense, d_v, n, d_d);
		else
			launch_svm_vectors(st, m->dev, d_v, (int)n, d_d);
		m->ctx->launches++;
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(host.data(), d_d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	free_all(tmp);
	if (e != cudaSuccess) return fail(FDB_ERR_CUDA, std::string("svm_get_probability: ") + cudaGetErrorString(e));
	for (int64_t i = 0; i < n; ++i) {
		const double dd = host[(size_t)i];
		if (dist_out) dist_out[i] = dd;
		if (prob_out) prob_out[i] = svm_probability(m->logistic_a, m->logistic_b, dd);
		if (pos_out) pos_out[i] = dd >= m->dev.threshold ? 1 : 0;
	}
	return FDB_OK;
} FDB_API_CATCH

/* ---------------------------------------------------------------------------------------------
 * RVM
 * ------------------------------------------------------------------------------------------- */
int fdb_rvm_create(fdb_ctx* ctx, const fdb_rvm_desc* d, fdb_rvm** out) try {
	int s = check_ctx(ctx); if (s) return s;
	if (!d || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	if (d->kernel < FDB_KERNEL_RBF || d->kernel > FDB_KERNEL_LINEAR) return fail(FDB_ERR_INVALID_ARGUMENT, "RVM: unknown kernel kind");
	if (d->num_filters < 1 || d->dim < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "RVM: empty model");
	if (d->sv_type != FDB_SV_U8 && d->sv_type != FDB_SV_F32) return fail(FDB_ERR_INVALID_ARGUMENT, "RVM: bad sv_type");
	if (!d->support_vectors || !d->coefficients || !d->hierarchical_thresholds) return fail(FDB_ERR_INVALID_ARGUMENT, "RVM: null array");
	if ((size_t)d->dim * 4 > 96 * 1024) return fail(FDB_ERR_UNSUPPORTED, "RVM: feature vector too long");
	/* the device form is the SVM's: support vectors transposed, one coefficient per vector (the diagonal c[l][l]) */
	std::vector<float> diag((size_t)d->num_filters);
	for (int l = 0; l < d->num_filters; ++l) diag[(size_t)l] = d->coefficients[(size_t)l * (l + 1) / 2 + l];
	fdb_svm_desc sd = fdb_svm_desc();
	sd.kernel = d->kernel; sd.gamma = d->gamma; sd.poly_alpha = d->poly_alpha; sd.poly_constant = d->poly_constant; sd.poly_degree = d->poly_degree;
	sd.num_sv = d->num_filters; sd.dim = d->dim; sd.sv_type = d->sv_type; sd.support_vectors = d->support_vectors;
	sd.coefficients = diag.data(); sd.bias = d->bias; sd.threshold = 0.f; sd.logistic_a = d->logistic_a; sd.logistic_b = d->logistic_b;
	fdb_svm* base = nullptr;
	s = fdb_svm_create(ctx, &sd, &base);
	if (s) return s;
	fdb_rvm* m = new fdb_rvm;
	static_cast<fdb_svm&>(*m) = *base;  /* takes over the device allocations */
	base->owned.clear();
	delete base;
	m->has_dense = false;
	m->rvm_thresholds.assign(d->hierarchical_thresholds, d->hierarchical_thresholds + d->num_filters);
	float* thr;
	s = upload(m->rvm_thresholds.data(), m->rvm_thresholds.size(), &thr, m->owned);
	if (s) { free_all(m->owned); delete m; return s; }
	m->dev.rvm_thresholds = thr;
	m->dev.rvm_filters = (d->num_filters_to_use <= 0 || d->num_filters_to_use > d->num_filters) ? d->num_filters : d->num_filters_to_use;
	*out = m;
	return FDB_OK;
} FDB_API_CATCH

void fdb_rvm_destroy(fdb_rvm* m) {
	if (!m) return;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	free_all(m->owned);
	delete m;
}

int fdb_rvm_set_num_filters_to_use(fdb_rvm* m, int32_t n) { /* RvmClassifier.cpp:119-126 */
	if (!m) return fail(FDB_ERR_INVALID_ARGUMENT, "null rvm");
	m->dev.rvm_filters = (n <= 0 || n > m->dev.num_sv) ? m->dev.num_sv : n;
	return FDB_OK;
}

int fdb_rvm_get_probability(fdb_rvm* m, const void* vectors, int64_t n, int32_t* level_out, double* dist_out, double* prob_out,
		uint8_t* pos_out) try {
	if (!m) return fail(FDB_ERR_INVALID_ARGUMENT, "null rvm");
	int s = check_ctx(m->ctx); if (s) return s;
	if (n < 0 || (n > 0 && !vectors)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad vector batch");
	if (n == 0) return FDB_OK;
	if (n > (1 << 30)) return fail(FDB_ERR_INVALID_ARGUMENT, "batch too large");
	const size_t es = m->dev.sv_type == FDB_SV_U8 ? 1 : 4;
	const size_t bytes = es * (size_t)m->dev.dim * (size_t)n;
	std::vector<void*> tmp;
	uint8_t* d_v; double* d_d; int* d_l;
	s = dev_alloc(&d_v, bytes, tmp); if (s) { free_all(tmp); return s; }
	s = dev_alloc(&d_d, (size_t)n, tmp); if (s) { free_all(tmp); return s; }
	s = dev_alloc(&d_l, (size_t)n, tmp); if (s) { free_all(tmp); return s; }
	cudaStream_t st = m->ctx->stream;
	std::vector<double> host((size_t)n);
	std::vector<int> lev((size_t)n);
	cudaError_t e = cudaMemcpyAsync(d_v, vectors, bytes, cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess) {
		launch_svm_vectors(st, m->dev, d_v, (int)n, d_d, d_l);
		m->ctx->launches++;
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(host.data(), d_d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaMemcpyAsync(lev.data(), d_l, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	free_all(tmp);
	if (e != cudaSuccess) return fail(FDB_ERR_CUDA, std::string("rvm_get_probability: ") + cudaGetErrorString(e));
	for (int64_t i = 0; i < n; ++i) {
		const double dd = host[(size_t)i];
		const int l = lev[(size_t)i];
		if (level_out) level_out[i] = l;
		if (dist_out) dist_out[i] = dd;
		if (prob_out) prob_out[i] = rvm_probability(m->logistic_a, m->logistic_b, dd);
		if (pos_out) pos_out[i] = (l + 1 == m->dev.rvm_filters && dd >= (double)m->rvm_thresholds[(size_t)l]) ? 1 : 0;
	}
	return FDB_OK;
} FDB_API_CATCH

int fdb_plan_layers(const fdb_detector_desc* desc, int32_t width, int32_t height, int32_t roi_x, int32_t roi_y,
		int32_t roi_w, int32_t roi_h, fdb_layer_info* out, int32_t cap, int32_t* n_layers, int64_t* n_windows) try {
	if (!desc) return fail(FDB_ERR_INVALID_ARGUMENT, "null descriptor");
	fdb_detector_desc d = *desc;
	if (d.step_x == 0) d.step_x = 1;
	if (d.step_y == 0) d.step_y = 1;
	if (d.step_x < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "DirectPyramidFeatureExtractor: stepX has to be greater than zero");
	if (d.step_y < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "DirectPyramidFeatureExtractor: stepY has to be greater than zero");
	Plan plan;
	int s = build_plan(d, width, height, &plan);
	if (s) return s;
	const int64_t windows = enumerate_windows(&plan, d.patch_width, d.patch_height, d.step_x, d.step_y, roi_x, roi_y, roi_w, roi_h);
	if (n_layers) *n_layers = (int32_t)plan.layers.size();
	if (n_windows) *n_windows = windows;
	for (size_t i = 0; i < plan.layers.size() && (int32_t)i < cap && out; ++i) {
		const PlanLayer& L = plan.layers[i];
		fdb_layer_info& o = out[i];
		o.index = L.index; o.scale = L.scale; o.width = L.width; o.height = L.height;
		o.orig_patch_width = L.orig_patch_w; o.orig_patch_height = L.orig_patch_h;
		o.windows_x = L.windows_x; o.windows_y = L.windows_y; o.first_window = L.first_window;
	}
	return FDB_OK;
} FDB_API_CATCH

int fdb_overlap_eliminate(fdb_detection* dets, int64_t n, float dist, float ratio, int64_t* n_out) try {
	if (n < 0 || (n > 0 && !dets)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad detection list");
	std::vector<fdb_detection> v(dets, dets + n);
	overlap_eliminate(v, dist, ratio);
	if (!v.empty()) std::memcpy(dets, v.data(), sizeof(fdb_detection) * v.size());
	if (n_out) *n_out = (int64_t)v.size();
	return FDB_OK;
} FDB_API_CATCH

int fdb_five_stage_nms(fdb_detection* dets, int64_t n, int32_t width, int32_t height, int64_t* n_out) try {
	if (n < 0 || (n > 0 && !dets) || width < 1 || height < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "bad detection list");
	std::vector<fdb_detection> v(dets, dets + n);
	five_stage_nms(v, width, height);
	if (!v.empty()) std::memcpy(dets, v.data(), sizeof(fdb_detection) * v.size());
	if (n_out) *n_out = (int64_t)v.size();
	return FDB_OK;
} FDB_API_CATCH

} // extern "C"
