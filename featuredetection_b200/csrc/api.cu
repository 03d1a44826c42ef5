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

namespace fdb {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int status, const std::string& msg) { g_last_error = msg; return status; }

#define CUDA_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
	return fdb::fail(FDB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } while (0)

template <class T>
static int upload(const T* host, size_t n, T** dev, std::vector<void*>& owned) {
	*dev = nullptr;
	void* p = nullptr;
	CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)));
	owned.push_back(p);
	if (n) CUDA_TRY(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
	*dev = (T*)p;
	return FDB_OK;
}

} // namespace fdb

using namespace fdb;

struct fdb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	int64_t launches = 0;
	cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; /* stopwatch + per-kernel profile marks */
};

struct fdb_wvm {
	fdb_ctx* ctx = nullptr;
	DevWvm dev{};
	std::vector<void*> owned;
	std::vector<float> thresholds_from_file;
	float* d_thresholds = nullptr;
	float limit = 0.f;
	double logistic_a = 0, logistic_b = 0;
	std::vector<float> thresholds; /* host copy incl. limit */
};

struct fdb_svm {
	fdb_ctx* ctx = nullptr;
	DevSvm dev{};
	std::vector<void*> owned;
	double logistic_a = 0, logistic_b = 0;
};

struct fdb_detector {
	fdb_ctx* ctx = nullptr;
	fdb_detector_desc desc{};
	fdb_wvm* wvm = nullptr;
	fdb_svm* svm = nullptr;
	Plan plan;
	bool prepared = false;
	int max_batch = 0;
	int cand_cap = 0; /* total candidate capacity of one chunk */
	std::vector<void*> owned;       /* device allocations */
	std::vector<void*> owned_host;  /* pinned host allocations */
	uint8_t* d_frames = nullptr;
	uint8_t* d_arena = nullptr;
	fdb_window_score* d_dense = nullptr;
	uint8_t* d_patches = nullptr; int64_t d_patches_bytes = 0;
	Candidate* d_cand = nullptr;
	int* d_cand_count = nullptr;      /* [0] candidate counter, [1] deep-queue counter */
	DeepQueue deep{};
	Strip* d_strips = nullptr; int n_strips = 0;
	bool use_strips = false;          /* fast path usable (and not yet overflowed) */
	DevLayer* d_layers = nullptr;     /* whole-image scan */
	DevLayer* d_layers_roi = nullptr; /* scratch table for ROI scans */
	ResizeJob* d_resize = nullptr; int n_resize = 0; int max_quads = 0;
	std::vector<DownJob*> d_down; std::vector<int> n_down; std::vector<int> max_down_px;
	int* d_ofs_tab = nullptr; short2* d_coef_tab = nullptr;
	SvmItem* d_items = nullptr; double* d_dist = nullptr; int items_cap = 0;
	Candidate* h_cand = nullptr; int* h_count = nullptr;
	SvmItem* h_items = nullptr; double* h_dist = nullptr;
	int64_t counts[5] = {0, 0, 0, 0, 0};
};

namespace {

int check_ctx(fdb_ctx* c) {
	if (!c) return fail(FDB_ERR_INVALID_ARGUMENT, "null context");
	CUDA_TRY(cudaSetDevice(c->device));
	return FDB_OK;
}

void free_all(std::vector<void*>& dev, std::vector<void*>* host = nullptr) {
	for (void* p : dev) cudaFree(p);
	dev.clear();
	if (host) { for (void* p : *host) cudaFreeHost(p); host->clear(); }
}

template <class T>
int dev_alloc(T** out, size_t n, std::vector<void*>& owned) {
	void* p = nullptr;
	CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)));
	owned.push_back(p);
	*out = (T*)p;
	return FDB_OK;
}

template <class T>
int host_alloc(T** out, size_t n, std::vector<void*>& owned) {
	void* p = nullptr;
	CUDA_TRY(cudaMallocHost(&p, std::max<size_t>(n * sizeof(T), 16)));
	owned.push_back(p);
	*out = (T*)p;
	return FDB_OK;
}

/* OpenCV bilinear coefficient tables for one axis (see pyramid.cu) */
void linear_tables(int src, int dst, bool clamp_fraction, std::vector<int>& ofs, std::vector<short2>& coef) {
	const double inv_scale = (double)dst / src;
	const double scale = 1. / inv_scale;
	for (int d = 0; d < dst; ++d) {
		float f = (float)((d + 0.5) * scale - 0.5);
		int s = (int)std::floor(f);
		f -= s;
		if (clamp_fraction) {
			if (s < 0) { f = 0; s = 0; }
			if (s >= src - 1) { f = 0; s = src - 1; }
		}
		ofs.push_back(s);
		short2 c;
		c.x = (short)std::nearbyint((1.f - f) * 2048.f);
		c.y = (short)std::nearbyint(f * 2048.f);
		coef.push_back(c);
	}
}

int upload_layers(fdb_detector* det, const Plan& plan, DevLayer* dst) {
	std::vector<DevLayer> L(plan.layers.size());
	for (size_t i = 0; i < plan.layers.size(); ++i) {
		const PlanLayer& p = plan.layers[i];
		L[i].offset = plan.images[p.image].offset;
		L[i].width = p.width; L[i].height = p.height;
		L[i].begin_x = p.begin_x; L[i].begin_y = p.begin_y;
		L[i].windows_x = p.windows_x; L[i].windows_y = p.windows_y;
		L[i].first_window = (int)p.first_window; L[i].pad = 0;
	}
	if (!L.empty())
		CUDA_TRY(cudaMemcpyAsync(dst, L.data(), sizeof(DevLayer) * L.size(), cudaMemcpyHostToDevice, det->ctx->stream));
	CUDA_TRY(cudaStreamSynchronize(det->ctx->stream)); /* L is a stack-owned staging buffer */
	return FDB_OK;
}

/* enqueue pyramid + stage-1 kernels for n frames resident at d_frames */
int enqueue_stage1(fdb_detector* det, const uint8_t* d_frames, int n, const Plan& plan, const DevLayer* d_layers,
		int64_t windows, fdb_window_score* d_dense, uint8_t* d_patches, bool want_candidates, bool marks = false) {
	fdb_ctx* c = det->ctx;
	cudaStream_t st = c->stream;
	const int W = plan.width, H = plan.height;
	CUDA_TRY(cudaMemsetAsync(det->d_cand_count, 0, 2 * sizeof(int), st));
	if (marks) CUDA_TRY(cudaEventRecord(c->ev[1], st));
	if (det->n_resize) {
		launch_resize(st, d_frames, W, H, n, det->d_arena, plan.arena_bytes, det->d_resize, det->n_resize,
				det->max_quads, det->d_ofs_tab, det->d_coef_tab);
		c->launches++;
	}
	if (marks) CUDA_TRY(cudaEventRecord(c->ev[2], st));
	for (size_t j = 0; j < det->d_down.size(); ++j) {
		if (!det->n_down[j]) continue;
		launch_pyrdown(st, d_frames, W, H, n, det->d_arena, plan.arena_bytes, det->d_down[j], det->n_down[j], det->max_down_px[j]);
		c->launches++;
	}
	if (marks) CUDA_TRY(cudaEventRecord(c->ev[3], st));
	if (windows > 0) {
		DevWvm m = det->wvm->dev;
		m.step_x = det->desc.step_x; m.step_y = det->desc.step_y;
		if (det->use_strips && d_layers == det->d_layers && !d_patches) {
			launch_wvm_strips(st, m, d_frames, W, H, n, det->d_arena, plan.arena_bytes, d_layers, det->d_strips, det->n_strips,
					(int)windows, d_dense, want_candidates ? det->d_cand : nullptr, det->d_cand_count, det->cand_cap, det->deep);
		} else {
			launch_wvm_windows(st, m, d_frames, W, H, n, det->d_arena, plan.arena_bytes, d_layers, (int)plan.layers.size(),
					(int)windows, d_dense, d_patches, want_candidates ? det->d_cand : nullptr, det->d_cand_count, det->cand_cap, det->deep);
		}
		c->launches += det->wvm->dev.num_lin > 0 ? 2 : 1;
	}
	if (marks) CUDA_TRY(cudaEventRecord(c->ev[4], st));
	CUDA_TRY(cudaGetLastError());
	return FDB_OK;
}

void fill_detection(fdb_detection* d, const Plan& plan, const fdb_detector_desc& desc, int frame, int64_t window) {
	/* DirectPyramidFeatureExtractor.cpp:115-118 */
	size_t li = 0;
	while (li + 1 < plan.layers.size() && window >= plan.layers[li + 1].first_window) ++li;
	const PlanLayer& L = plan.layers[li];
	const int64_t local = window - L.first_window;
	const int iy = (int)(local / L.windows_x), ix = (int)(local - (int64_t)iy * L.windows_x);
	std::memset(d, 0, sizeof(*d));
	d->frame = frame; d->layer = L.index;
	d->x = L.begin_x + ix * desc.step_x; d->y = L.begin_y + iy * desc.step_y;
	d->width = L.orig_patch_w; d->height = L.orig_patch_h;
	d->center_x = cv_round(d->x / L.scale) + L.orig_patch_w / 2;
	d->center_y = cv_round(d->y / L.scale) + L.orig_patch_h / 2;
	d->window = window;
	d->reserved = (int32_t)li;
}

/* candidates of one chunk -> per-frame post-processing -> detections appended to out */
int finish_chunk(fdb_detector* det, const uint8_t* d_frames, int n, int frame_base, const Plan& plan,
		const DevLayer* d_layers, int stage, bool is_roi, std::vector<fdb_detection>& out) {
	fdb_ctx* c = det->ctx;
	cudaStream_t st = c->stream;
	CUDA_TRY(cudaMemcpyAsync(det->h_count, det->d_cand_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	if (det->use_strips && d_layers == det->d_layers && det->h_count[1] > det->deep.cap) {
		/* more survivors than the deep queue holds (a model with hardly any early exits): the strip
		 * kernel cannot finish them inline, so this detector switches to the generic kernels for good */
		det->use_strips = false;
		return -1000;
	}
	const int ncand = *det->h_count;
	if (ncand > det->cand_cap)
		return fail(FDB_ERR_OVERFLOW, "stage-1 candidate list overflow: raise max_positives_per_frame");
	if (ncand) {
		CUDA_TRY(cudaMemcpyAsync(det->h_cand, det->d_cand, sizeof(Candidate) * (size_t)ncand, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	/* canonical order: (frame, window) - SlidingWindowDetector::detect() pushes in extract order */
	std::vector<Candidate> cand(det->h_cand, det->h_cand + ncand);
	std::sort(cand.begin(), cand.end(), [](const Candidate& a, const Candidate& b) {
		return a.frame != b.frame ? a.frame < b.frame : a.window < b.window; });
	det->counts[1] += ncand;
	std::vector<std::vector<fdb_detection>> per_frame((size_t)n);
	for (const Candidate& k : cand) {
		fdb_detection d;
		fill_detection(&d, plan, det->desc, frame_base + k.frame, k.window);
		d.wvm_level = k.level; d.wvm_fout = k.fout;
		d.wvm_probability = wvm_probability(det->wvm->logistic_a, det->wvm->logistic_b, k.fout);
		d.svm_distance = std::numeric_limits<double>::quiet_NaN();
		d.svm_probability = std::numeric_limits<double>::quiet_NaN();
		d.probability = d.wvm_probability;
		d.positive = 1;
		per_frame[(size_t)k.frame].push_back(d);
	}
	if (stage >= FDB_STAGE_OE)
		for (auto& v : per_frame) { overlap_eliminate(v, det->desc.oe_dist, det->desc.oe_ratio); det->counts[2] += (int64_t)v.size(); }
	if (stage >= FDB_STAGE_SVM && det->svm) {
		size_t total = 0;
		for (auto& v : per_frame) total += v.size();
		if ((int64_t)total > det->items_cap)
			return fail(FDB_ERR_OVERFLOW, "SVM work list overflow");
		size_t k = 0;
		for (int f = 0; f < n; ++f)
			for (const fdb_detection& d : per_frame[(size_t)f]) {
				SvmItem it; it.frame = f; it.layer = d.reserved; it.x = d.x; it.y = d.y;
				det->h_items[k++] = it;
			}
		if (total) {
			CUDA_TRY(cudaMemcpyAsync(det->d_items, det->h_items, sizeof(SvmItem) * total, cudaMemcpyHostToDevice, st));
			launch_svm_windows(st, det->svm->dev, det->desc.patch_width, det->desc.patch_height, d_frames, plan.width, plan.height,
					det->d_arena, plan.arena_bytes, d_layers, det->d_items, (int)total, det->d_dist);
			c->launches++;
			CUDA_TRY(cudaGetLastError());
			CUDA_TRY(cudaMemcpyAsync(det->h_dist, det->d_dist, sizeof(double) * total, cudaMemcpyDeviceToHost, st));
			CUDA_TRY(cudaStreamSynchronize(st));
		}
		k = 0;
		for (int f = 0; f < n; ++f) {
			std::vector<fdb_detection>& v = per_frame[(size_t)f];
			std::vector<fdb_detection> pos;
			for (fdb_detection& d : v) {
				d.svm_distance = det->h_dist[k++];
				d.svm_probability = svm_probability(det->svm->logistic_a, det->svm->logistic_b, d.svm_distance);
				/* FiveStageSlidingWindowDetector.cpp:260: ClassifiedPatch(patch, classify(...)) => probability 0.5 */
				d.positive = d.svm_distance >= det->svm->dev.threshold ? 1 : 0;
				d.probability = 0.5;
				if (d.positive) pos.push_back(d);
			}
			v.swap(pos);
			det->counts[3] += (int64_t)v.size();
			if (stage >= FDB_STAGE_NMS && !is_roi) five_stage_nms(v, plan.width, plan.height);
			else stable_sort_desc(v);
			det->counts[4] += (int64_t)v.size();
		}
	}
	for (auto& v : per_frame)
		for (fdb_detection& d : v) { d.reserved = 0; out.push_back(d); }
	return FDB_OK;
}

int ensure_dense(fdb_detector* det) {
	if (det->d_dense) return FDB_OK;
	return dev_alloc(&det->d_dense, (size_t)det->max_batch * (size_t)std::max<int64_t>(det->plan.windows, 1), det->owned);
}

int copy_out(const std::vector<fdb_detection>& dets, fdb_detection* out, int64_t cap, int64_t* n_out) {
	if (n_out) *n_out = (int64_t)dets.size();
	if ((int64_t)dets.size() > cap)
		return fail(FDB_ERR_OVERFLOW, "detections_out capacity too small");
	if (!dets.empty() && out) std::memcpy(out, dets.data(), sizeof(fdb_detection) * dets.size());
	return FDB_OK;
}

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

int fdb_ctx_create(int device, fdb_ctx** out) {
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
	if (wvm_configure() != 0 || svm_configure() != 0 || strip_configure_all() != 0) {
		cudaStreamDestroy(c->stream); delete c;
		return fail(FDB_ERR_CUDA, "cudaFuncSetAttribute failed: libfdb200 kernels not loadable on this device");
	}
	*out = c;
	return FDB_OK;
}

void fdb_ctx_destroy(fdb_ctx* c) {
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	for (cudaEvent_t e : c->ev) if (e) cudaEventDestroy(e);
	cudaStreamDestroy(c->stream);
	delete c;
}

void* fdb_ctx_stream(fdb_ctx* c) { return c ? (void*)c->stream : nullptr; }

int fdb_ctx_synchronize(fdb_ctx* c) {
	int s = check_ctx(c); if (s) return s;
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return FDB_OK;
}

int fdb_ctx_timer_start(fdb_ctx* c) {
	int s = check_ctx(c); if (s) return s;
	CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
	return FDB_OK;
}

int fdb_ctx_timer_stop(fdb_ctx* c, double* elapsed_ms) {
	int s = check_ctx(c); if (s) return s;
	CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
	CUDA_TRY(cudaEventSynchronize(c->ev[1]));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
	if (elapsed_ms) *elapsed_ms = ms;
	return FDB_OK;
}

int64_t fdb_ctx_launch_count(fdb_ctx* c) { return c ? c->launches : 0; }

int fdb_host_alloc(size_t bytes, void** out) {
	if (!out) return fail(FDB_ERR_INVALID_ARGUMENT, "out is null");
	CUDA_TRY(cudaMallocHost(out, std::max<size_t>(bytes, 16)));
	return FDB_OK;
}
void fdb_host_free(void* p) { if (p) cudaFreeHost(p); }

/* ---------------------------------------------------------------------------------------------
 * WVM
 * ------------------------------------------------------------------------------------------- */
int fdb_wvm_create(fdb_ctx* ctx, const fdb_wvm_desc* d, fdb_wvm** out) {
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
		if (cntval - 1 > FDB_MAX_VALUES) return fail(FDB_ERR_UNSUPPORTED, "WVM: more than 8 rectangle grey values per filter");
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
	dv.masks4 = nullptr;
	if (max_nv <= 4) { /* padded copy for the strip / deep-warp kernels: [filter][word][4] */
		std::vector<uint32_t> m4((size_t)n * nwords * 4, 0u);
		for (int f = 0; f < n; ++f) {
			const int nv = d->area_cntval[f] - 1;
			for (int j = 0; j < nwords; ++j)
				for (int v = 0; v < nv; ++v) m4[((size_t)f * nwords + j) * 4 + v] = masks[(size_t)mask_off[f] + (size_t)j * nv + v];
		}
		UP(m4.data(), m4.size(), masks4, up);
	}
#undef UP
	s = dev_alloc(&m->d_thresholds, (size_t)n, m->owned);
	if (s) { free_all(m->owned); delete m; return s; }
	dv.thresholds = m->d_thresholds;
	s = fdb_wvm_set_limit_reliability_filter(m, d->limit_reliability_filter);
	if (s) { free_all(m->owned); delete m; return s; }
	*out = m;
	return FDB_OK;
}

void fdb_wvm_destroy(fdb_wvm* m) {
	if (!m) return;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	free_all(m->owned);
	delete m;
}

int fdb_wvm_set_limit_reliability_filter(fdb_wvm* m, float value) {
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
}

int fdb_wvm_get_probability(fdb_wvm* m, const uint8_t* patches, int64_t n, int32_t* level_out, float* fout_out,
		double* prob_out, uint8_t* pos_out) {
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
}

/* ---------------------------------------------------------------------------------------------
 * SVM
 * ------------------------------------------------------------------------------------------- */
int fdb_svm_create(fdb_ctx* ctx, const fdb_svm_desc* d, fdb_svm** out) {
	int s = check_ctx(ctx); if (s) return s;
	if (!d || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	if (d->kernel != FDB_KERNEL_RBF) return fail(FDB_ERR_UNSUPPORTED, "SVM: only the RBF kernel is implemented");
	if (d->num_sv < 1 || d->dim < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "SVM: empty model");
	if (d->sv_type != FDB_SV_U8 && d->sv_type != FDB_SV_F32) return fail(FDB_ERR_INVALID_ARGUMENT, "SVM: bad sv_type");
	if ((size_t)d->dim * 4 > 96 * 1024) return fail(FDB_ERR_UNSUPPORTED, "SVM: feature vector too long");
	fdb_svm* m = new fdb_svm;
	m->ctx = ctx;
	m->logistic_a = d->logistic_a; m->logistic_b = d->logistic_b;
	DevSvm& dv = m->dev;
	dv.num_sv = d->num_sv; dv.dim = d->dim; dv.sv_type = d->sv_type;
	dv.nwords = (d->dim + 3) / 4;
	dv.gamma = d->gamma; dv.bias = d->bias; dv.threshold = d->threshold;
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
}

void fdb_svm_destroy(fdb_svm* m) {
	if (!m) return;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	free_all(m->owned);
	delete m;
}

int fdb_svm_set_threshold(fdb_svm* m, float t) {
	if (!m) return fail(FDB_ERR_INVALID_ARGUMENT, "null svm");
	m->dev.threshold = t;
	return FDB_OK;
}

int fdb_svm_get_probability(fdb_svm* m, const void* vectors, int64_t n, double* dist_out, double* prob_out, uint8_t* pos_out) {
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
}

/* ---------------------------------------------------------------------------------------------
 * Detector
 * ------------------------------------------------------------------------------------------- */
int fdb_detector_create(fdb_ctx* ctx, const fdb_detector_desc* desc, fdb_wvm* wvm, fdb_svm* svm, fdb_detector** out) {
	int s = check_ctx(ctx); if (s) return s;
	if (!desc || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	if (!wvm) return fail(FDB_ERR_INVALID_ARGUMENT, "detector needs a first-stage classifier");
	fdb_detector_desc d = *desc;
	if (d.step_x == 0) d.step_x = 1;
	if (d.step_y == 0) d.step_y = 1;
	/* DirectPyramidFeatureExtractor.cpp:77-80 */
	if (d.step_x < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "DirectPyramidFeatureExtractor: stepX has to be greater than zero");
	if (d.step_y < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "DirectPyramidFeatureExtractor: stepY has to be greater than zero");
	if (d.patch_width != wvm->dev.fsx || d.patch_height != wvm->dev.fsy)
		return fail(FDB_ERR_INVALID_ARGUMENT, "patch size differs from the WVM filter size");
	if (svm && (svm->dev.sv_type != FDB_SV_U8 || svm->dev.dim != d.patch_width * d.patch_height))
		return fail(FDB_ERR_INVALID_ARGUMENT, "second-stage SVM must take the u8 patch as its feature vector");
	if (d.max_positives_per_frame <= 0) d.max_positives_per_frame = 4096;
	Plan probe;
	s = build_plan(d, 64, 64, &probe); /* validates the pyramid parameters (ImagePyramid.cpp:84-89) */
	if (s) return s;
	fdb_detector* det = new fdb_detector;
	det->ctx = ctx; det->desc = d; det->wvm = wvm; det->svm = svm;
	*out = det;
	return FDB_OK;
}

void fdb_detector_destroy(fdb_detector* det) {
	if (!det) return;
	cudaSetDevice(det->ctx->device);
	cudaStreamSynchronize(det->ctx->stream);
	free_all(det->owned, &det->owned_host);
	delete det;
}

int fdb_detector_prepare(fdb_detector* det, int32_t width, int32_t height, int32_t max_batch) {
	if (!det) return fail(FDB_ERR_INVALID_ARGUMENT, "null detector");
	int s = check_ctx(det->ctx); if (s) return s;
	if (max_batch < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "max_batch must be positive");
	CUDA_TRY(cudaStreamSynchronize(det->ctx->stream));
	free_all(det->owned, &det->owned_host);
	det->prepared = false;
	det->d_dense = nullptr; det->d_patches = nullptr; det->d_patches_bytes = 0;
	det->d_down.clear(); det->n_down.clear(); det->max_down_px.clear();
	s = build_plan(det->desc, width, height, &det->plan);
	if (s) return s;
	const Plan& plan = det->plan;
	if (plan.windows >= (int64_t)1 << 31) return fail(FDB_ERR_UNSUPPORTED, "too many windows per frame");
	det->max_batch = max_batch;
	const int64_t cap64 = (int64_t)det->desc.max_positives_per_frame * max_batch;
	det->cand_cap = (int)std::min<int64_t>(cap64, (int64_t)1 << 26);
	det->items_cap = det->cand_cap;

	/* job tables */
	std::vector<ResizeJob> rj;
	std::vector<std::vector<DownJob>> dj((size_t)plan.max_down + 1);
	std::vector<int> ofs; std::vector<short2> coef;
	det->max_quads = 0;
	for (const PyrImage& im : plan.images) {
		if (im.kind == IMG_RESIZE) {
			ResizeJob j{};
			j.dst_w = im.width; j.dst_h = im.height; j.dst_offset = im.offset;
			j.area2x = (width == 2 * im.width && height == 2 * im.height) ? 1 : 0;
			j.xtab = (int)ofs.size();
			linear_tables(width, im.width, true, ofs, coef);
			j.ytab = (int)ofs.size();
			linear_tables(height, im.height, false, ofs, coef);
			rj.push_back(j);
			det->max_quads = std::max(det->max_quads, ((im.width + 3) / 4) * im.height);
		} else if (im.kind == IMG_PYRDOWN) {
			const PyrImage& src = plan.images[(size_t)im.src];
			DownJob j{};
			j.src_w = src.width; j.src_h = src.height; j.dst_w = im.width; j.dst_h = im.height;
			j.src_offset = src.kind == IMG_FRAME ? -1 : src.offset; j.dst_offset = im.offset;
			dj[(size_t)im.down].push_back(j);
		}
	}
	det->n_resize = (int)rj.size();
	s = upload(rj.data(), rj.size(), &det->d_resize, det->owned); if (s) return s;
	s = upload(ofs.data(), ofs.size(), &det->d_ofs_tab, det->owned); if (s) return s;
	s = upload(coef.data(), coef.size(), &det->d_coef_tab, det->owned); if (s) return s;
	for (size_t j = 1; j < dj.size(); ++j) {
		DownJob* p = nullptr;
		s = upload(dj[j].data(), dj[j].size(), &p, det->owned); if (s) return s;
		int mx = 0;
		for (const DownJob& q : dj[j]) mx = std::max(mx, q.dst_w * q.dst_h);
		det->d_down.push_back(p); det->n_down.push_back((int)dj[j].size()); det->max_down_px.push_back(mx);
	}
	s = dev_alloc(&det->d_frames, (size_t)max_batch * width * height, det->owned); if (s) return s;
	s = dev_alloc(&det->d_arena, (size_t)max_batch * (size_t)plan.arena_bytes, det->owned); if (s) return s;
	s = dev_alloc(&det->d_cand, (size_t)det->cand_cap, det->owned); if (s) return s;
	s = dev_alloc(&det->d_cand_count, 4, det->owned); if (s) return s;
	/* deep queue: room for 1/16 of the windows of a full batch (beyond that windows finish inline) */
	det->deep.count = det->d_cand_count + 1;
	det->deep.cap = (int)std::max<int64_t>(1024, std::min<int64_t>(plan.windows * max_batch / 16 + 1024, (int64_t)1 << 24));
	s = dev_alloc(&det->deep.rec, (size_t)det->deep.cap, det->owned); if (s) return s;
	s = dev_alloc(&det->deep.patch, (size_t)det->deep.cap * (size_t)det->wvm->dev.nwords, det->owned); if (s) return s;
	s = dev_alloc(&det->d_layers, FDB_MAX_LAYERS, det->owned); if (s) return s;
	s = dev_alloc(&det->d_layers_roi, FDB_MAX_LAYERS, det->owned); if (s) return s;
	s = dev_alloc(&det->d_items, (size_t)det->items_cap, det->owned); if (s) return s;
	s = dev_alloc(&det->d_dist, (size_t)det->items_cap, det->owned); if (s) return s;
	s = host_alloc(&det->h_cand, (size_t)det->cand_cap, det->owned_host); if (s) return s;
	s = host_alloc(&det->h_count, 4, det->owned_host); if (s) return s;
	s = host_alloc(&det->h_items, (size_t)det->items_cap, det->owned_host); if (s) return s;
	s = host_alloc(&det->h_dist, (size_t)det->items_cap, det->owned_host); if (s) return s;
	s = upload_layers(det, plan, det->d_layers); if (s) return s;
	/* strip table of the fast path: whole-image scan, step 1, supported patch size, <= 4 grey values, <= 256 words */
	det->use_strips = det->desc.step_x == 1 && det->desc.step_y == 1 && det->wvm->dev.masks4 != nullptr
			&& strip_supported(det->desc.patch_width, det->desc.patch_height) && det->wvm->dev.num_lin > WVM_KA
			&& det->wvm->dev.num_used > WVM_KA;
	std::vector<Strip> strips;
	if (det->use_strips) {
		for (size_t li = 0; li < plan.layers.size(); ++li) {
			const PlanLayer& L = plan.layers[li];
			for (int ix0 = 0; ix0 < L.windows_x; ix0 += 32) {
				const int cols = std::min(32, L.windows_x - ix0);
				const int nsub = std::min(WVM_MAXSUB, 32 / cols);
				for (int iy0 = 0; iy0 < L.windows_y; iy0 += nsub * WVM_RUN) {
					Strip st{};
					st.layer = (int)li; st.ix0 = ix0; st.iy0 = iy0; st.cols = cols;
					st.nsub = std::min(nsub, (L.windows_y - iy0 + WVM_RUN - 1) / WVM_RUN);
					strips.push_back(st);
				}
			}
		}
	}
	det->n_strips = (int)strips.size();
	s = upload(strips.data(), strips.size(), &det->d_strips, det->owned); if (s) return s;
	det->prepared = true;
	return FDB_OK;
}

int fdb_detector_layers(fdb_detector* det, fdb_layer_info* out, int32_t cap, int32_t* n_layers) {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared");
	const Plan& p = det->plan;
	if (n_layers) *n_layers = (int32_t)p.layers.size();
	for (size_t i = 0; i < p.layers.size() && (int32_t)i < cap && out; ++i) {
		const PlanLayer& L = p.layers[i];
		fdb_layer_info& o = out[i];
		o.index = L.index; o.scale = L.scale; o.width = L.width; o.height = L.height;
		o.orig_patch_width = L.orig_patch_w; o.orig_patch_height = L.orig_patch_h;
		o.windows_x = L.windows_x; o.windows_y = L.windows_y; o.first_window = L.first_window;
	}
	return FDB_OK;
}

int64_t fdb_detector_windows_per_frame(fdb_detector* det) { return det && det->prepared ? det->plan.windows : -1; }
int64_t fdb_detector_pyramid_bytes(fdb_detector* det) {
	if (!det || !det->prepared) return -1;
	int64_t b = 0;
	for (const PyrImage& im : det->plan.images) if (im.kind != IMG_FRAME) b += (int64_t)im.width * im.height;
	return b;
}

static int detect_impl(fdb_detector* det, const uint8_t* frames, bool frames_on_device, int64_t pitch, int32_t n_frames,
		int32_t stage, fdb_window_score* dense_out, bool dense_on_device, fdb_detection* dets_out, int64_t det_cap,
		int64_t* n_dets) {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (n_frames < 0 || (n_frames > 0 && !frames)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad frame batch");
	if (stage < FDB_STAGE_WVM || stage > FDB_STAGE_NMS) return fail(FDB_ERR_INVALID_ARGUMENT, "bad stage");
	const Plan& plan = det->plan;
	const int W = plan.width, H = plan.height;
	if (!frames_on_device && pitch < W) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	cudaStream_t st = det->ctx->stream;
	std::fill(det->counts, det->counts + 5, 0);
	det->counts[0] = plan.windows * n_frames;
	std::vector<fdb_detection> dets;
	if (dense_out && !dense_on_device) { s = ensure_dense(det); if (s) return s; }
	for (int base = 0; base < n_frames; base += det->max_batch) {
		const int n = std::min(det->max_batch, n_frames - base);
		const uint8_t* d_frames;
		if (frames_on_device) {
			d_frames = frames + (int64_t)base * W * H;
		} else {
			CUDA_TRY(cudaMemcpy2DAsync(det->d_frames, (size_t)W, frames + (int64_t)base * pitch * H, (size_t)pitch, (size_t)W,
					(size_t)H * n, cudaMemcpyHostToDevice, st));
			d_frames = det->d_frames;
		}
		fdb_window_score* d_dense = nullptr;
		if (dense_out) d_dense = dense_on_device ? dense_out + (int64_t)base * plan.windows : det->d_dense;
		s = enqueue_stage1(det, d_frames, n, plan, det->d_layers, plan.windows, d_dense, nullptr, true);
		if (s) return s;
		if (dense_out && !dense_on_device && plan.windows > 0)
			CUDA_TRY(cudaMemcpyAsync(dense_out + (int64_t)base * plan.windows, det->d_dense,
					sizeof(fdb_window_score) * (size_t)plan.windows * n, cudaMemcpyDeviceToHost, st));
		s = finish_chunk(det, d_frames, n, base, plan, det->d_layers, stage, false, dets);
		if (s == -1000) { /* deep-queue overflow: redo this chunk on the generic path */
			base -= det->max_batch;
			continue;
		}
		if (s) return s;
	}
	return copy_out(dets, dets_out, det_cap, n_dets);
}

int fdb_detect_batch(fdb_detector* det, const uint8_t* frames_host, int64_t pitch, int32_t n_frames, int32_t stage,
		fdb_window_score* dense_out, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) {
	return detect_impl(det, frames_host, false, pitch, n_frames, stage, dense_out, false, detections_out, det_cap, n_detections);
}

int fdb_detect_batch_device(fdb_detector* det, const uint8_t* frames_device, int32_t n_frames, int32_t stage,
		fdb_window_score* dense_out_device, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) {
	return detect_impl(det, frames_device, true, 0, n_frames, stage, dense_out_device, true, detections_out, det_cap, n_detections);
}

int fdb_detect_enqueue_device(fdb_detector* det, const uint8_t* frames_device, int32_t n_frames, fdb_window_score* dense_out_device) {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (n_frames < 0 || n_frames > det->max_batch) return fail(FDB_ERR_INVALID_ARGUMENT, "n_frames exceeds the prepared batch");
	return enqueue_stage1(det, frames_device, n_frames, det->plan, det->d_layers, det->plan.windows, dense_out_device, nullptr, true);
}

int fdb_detect_profile_device(fdb_detector* det, const uint8_t* frames_device, int32_t n_frames, double ms_out[4]) {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (n_frames < 0 || n_frames > det->max_batch || !ms_out) return fail(FDB_ERR_INVALID_ARGUMENT, "bad arguments");
	s = enqueue_stage1(det, frames_device, n_frames, det->plan, det->d_layers, det->plan.windows, nullptr, nullptr, true, true);
	if (s) return s;
	fdb_ctx* c = det->ctx;
	CUDA_TRY(cudaEventSynchronize(c->ev[4]));
	float a = 0, b = 0, w = 0, t = 0;
	CUDA_TRY(cudaEventElapsedTime(&a, c->ev[1], c->ev[2]));
	CUDA_TRY(cudaEventElapsedTime(&b, c->ev[2], c->ev[3]));
	CUDA_TRY(cudaEventElapsedTime(&w, c->ev[3], c->ev[4]));
	CUDA_TRY(cudaEventElapsedTime(&t, c->ev[1], c->ev[4]));
	ms_out[0] = a; ms_out[1] = b; ms_out[2] = w; ms_out[3] = t;
	return FDB_OK;
}

int fdb_detect_roi(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, int32_t roi_x, int32_t roi_y, int32_t roi_w,
		int32_t roi_h, int32_t stage, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (!frame_host) return fail(FDB_ERR_INVALID_ARGUMENT, "null frame");
	if (stage < FDB_STAGE_WVM || stage > FDB_STAGE_NMS) return fail(FDB_ERR_INVALID_ARGUMENT, "bad stage");
	Plan plan = det->plan;
	const int W = plan.width, H = plan.height;
	if (pitch < W) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	const bool is_roi = !(roi_x == 0 && roi_y == 0 && roi_w == 0 && roi_h == 0);
	const int64_t windows = enumerate_windows(&plan, det->desc.patch_width, det->desc.patch_height, det->desc.step_x,
			det->desc.step_y, roi_x, roi_y, roi_w, roi_h);
	plan.windows = windows;
	cudaStream_t st = det->ctx->stream;
	s = upload_layers(det, plan, det->d_layers_roi); if (s) return s;
	CUDA_TRY(cudaMemcpy2DAsync(det->d_frames, (size_t)W, frame_host, (size_t)pitch, (size_t)W, (size_t)H, cudaMemcpyHostToDevice, st));
	std::fill(det->counts, det->counts + 5, 0);
	det->counts[0] = windows;
	s = enqueue_stage1(det, det->d_frames, 1, plan, det->d_layers_roi, windows, nullptr, nullptr, true);
	if (s) return s;
	std::vector<fdb_detection> dets;
	s = finish_chunk(det, det->d_frames, 1, 0, plan, det->d_layers_roi, stage, is_roi, dets);
	if (s) return s;
	return copy_out(dets, detections_out, det_cap, n_detections);
}

int fdb_extract_patches(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, uint8_t* patches_out, int64_t cap_windows,
		int64_t* n_windows) {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	const Plan& plan = det->plan;
	if (n_windows) *n_windows = plan.windows;
	if (!frame_host || !patches_out) return fail(FDB_ERR_INVALID_ARGUMENT, "null buffer");
	if (cap_windows < plan.windows) return fail(FDB_ERR_OVERFLOW, "patches_out capacity too small");
	if (pitch < plan.width) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	const int64_t bytes = plan.windows * det->desc.patch_width * det->desc.patch_height;
	if (det->d_patches_bytes < bytes) {
		s = dev_alloc(&det->d_patches, (size_t)bytes, det->owned); if (s) return s;
		det->d_patches_bytes = bytes;
	}
	cudaStream_t st = det->ctx->stream;
	CUDA_TRY(cudaMemcpy2DAsync(det->d_frames, (size_t)plan.width, frame_host, (size_t)pitch, (size_t)plan.width, (size_t)plan.height,
			cudaMemcpyHostToDevice, st));
	s = enqueue_stage1(det, det->d_frames, 1, plan, det->d_layers, plan.windows, nullptr, det->d_patches, false);
	if (s) return s;
	if (bytes) CUDA_TRY(cudaMemcpyAsync(patches_out, det->d_patches, (size_t)bytes, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	return FDB_OK;
}

int fdb_pyramid_layer(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, int32_t layer_index, uint8_t* out, int64_t cap) {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	const Plan& plan = det->plan;
	const PlanLayer* L = nullptr;
	for (const PlanLayer& l : plan.layers) if (l.index == layer_index) L = &l;
	if (!L) return fail(FDB_ERR_INVALID_ARGUMENT, "no such pyramid layer");
	if (!frame_host || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null buffer");
	if (cap < (int64_t)L->width * L->height) return fail(FDB_ERR_OVERFLOW, "layer buffer too small");
	if (pitch < plan.width) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	cudaStream_t st = det->ctx->stream;
	CUDA_TRY(cudaMemcpy2DAsync(det->d_frames, (size_t)plan.width, frame_host, (size_t)pitch, (size_t)plan.width, (size_t)plan.height,
			cudaMemcpyHostToDevice, st));
	s = enqueue_stage1(det, det->d_frames, 1, plan, det->d_layers, 0, nullptr, nullptr, false);
	if (s) return s;
	const PyrImage& im = plan.images[(size_t)L->image];
	const uint8_t* src = im.kind == IMG_FRAME ? det->d_frames : det->d_arena + im.offset;
	CUDA_TRY(cudaMemcpyAsync(out, src, (size_t)L->width * L->height, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	return FDB_OK;
}

int fdb_plan_layers(const fdb_detector_desc* desc, int32_t width, int32_t height, int32_t roi_x, int32_t roi_y,
		int32_t roi_w, int32_t roi_h, fdb_layer_info* out, int32_t cap, int32_t* n_layers, int64_t* n_windows) {
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
}

int fdb_overlap_eliminate(fdb_detection* dets, int64_t n, float dist, float ratio, int64_t* n_out) {
	if (n < 0 || (n > 0 && !dets)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad detection list");
	std::vector<fdb_detection> v(dets, dets + n);
	overlap_eliminate(v, dist, ratio);
	if (!v.empty()) std::memcpy(dets, v.data(), sizeof(fdb_detection) * v.size());
	if (n_out) *n_out = (int64_t)v.size();
	return FDB_OK;
}

int fdb_five_stage_nms(fdb_detection* dets, int64_t n, int32_t width, int32_t height, int64_t* n_out) {
	if (n < 0 || (n > 0 && !dets) || width < 1 || height < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "bad detection list");
	std::vector<fdb_detection> v(dets, dets + n);
	five_stage_nms(v, width, height);
	if (!v.empty()) std::memcpy(dets, v.data(), sizeof(fdb_detection) * v.size());
	if (n_out) *n_out = (int64_t)v.size();
	return FDB_OK;
}

int fdb_detector_last_counts(fdb_detector* det, int64_t counts[5]) {
	if (!det || !counts) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	std::memcpy(counts, det->counts, sizeof(det->counts));
	return FDB_OK;
}

} // extern "C"
