/*
 * detector.cu - the detector half of the C ABI: pyramid plan upload, the software-pipelined
 * batch path and the single-frame entry points.
 *
 * Host side of the product (C++), mirroring
 *   fdb_detector <- FiveStageSlidingWindowDetector( SlidingWindowDetector( pwvm,
 *                     DirectPyramidFeatureExtractor( ImagePyramid + GrayscaleFilter, HistEq64Filter ) ),
 *                     OverlapElimination, psvm )                   (ffpDetectApp.cpp:391-425)
 *
 * A batch is cut into chunks of `chunk` frames that travel through PIPE_SLOTS independent slots
 * (own stream, frame staging, pyramid arena, candidate list, deep queue, pinned result buffers):
 *
 *   slot stream:  H2D frames -> resize -> pyrDown x4 -> window kernel -> deep kernel -> D2H counts+candidates
 *                 ... (host: overlap elimination) ... H2D SVM items -> SVM kernel -> D2H distances
 *   host:         chunk i is enqueued, then chunk i-1 gets its overlap elimination + SVM launch,
 *                 then chunk i-2 is classified / NMS'ed and appended - so copies, the latency-bound
 *                 deep/SVM kernels and the host post-processing of neighbouring chunks overlap.
 * Results are appended in chunk order, i.e. in frame order, exactly as the serial path would.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "detector_internal.h"

using namespace fdb;

namespace fdb {

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

int build_pyramid_jobs(const std::vector<PyrImage>& images, int max_down, int width, int height, PyramidJobs* out, std::vector<void*>& owned) {
	*out = PyramidJobs();
	std::vector<ResizeJob> rj;
	std::vector<std::vector<DownJob>> dj((size_t)max_down + 1);
	std::vector<int> ofs, winfo; std::vector<short2> coef;
	for (const PyrImage& im : images) {
		if (im.kind == IMG_RESIZE) {
			ResizeJob j{};
			j.dst_w = im.width; j.dst_h = im.height; j.dst_pitch = im.pitch; j.dst_offset = im.offset;
			j.area2x = (width == 2 * im.width && height == 2 * im.height) ? 1 : 0;
			j.xtab = (int)ofs.size();
			linear_tables(width, im.width, true, ofs, coef);
			/* word-path info per output column (see resize_kernel) */
			j.words_ok = (width % 4 == 0) ? 1 : 0;
			winfo.resize(ofs.size(), 0);
			for (int dx0 = 0; dx0 < im.width; dx0 += 4) {
				const int base = ofs[(size_t)j.xtab + dx0] & ~3;
				for (int k = 0; k < 4 && dx0 + k < im.width; ++k) {
					const int o = ofs[(size_t)j.xtab + dx0 + k] - base;
					if (o < 0 || o > 10) j.words_ok = 0;
					winfo[(size_t)j.xtab + dx0 + k] = (o >> 2) | (((o & 3) * 8) << 8) | ((base >> 2) << 16);
				}
			}
			j.ytab = (int)ofs.size();
			linear_tables(height, im.height, false, ofs, coef);
			winfo.resize(ofs.size(), 0);
			rj.push_back(j);
			out->max_quads = std::max(out->max_quads, resize_tiles(im.width, im.height));
		} else if (im.kind == IMG_PYRDOWN) {
			const PyrImage& src = images[(size_t)im.src];
			DownJob j{};
			j.src_w = src.width; j.src_h = src.height; j.dst_w = im.width; j.dst_h = im.height;
			j.src_pitch = src.pitch; j.dst_pitch = im.pitch;
			j.src_offset = src.kind == IMG_FRAME ? -1 : src.offset; j.dst_offset = im.offset;
			dj[(size_t)im.down].push_back(j);
		}
	}
	out->n_resize = (int)rj.size();
	int s = upload(rj.data(), rj.size(), &out->d_resize, owned); if (s) return s;
	std::vector<int4> xy(ofs.size());
	winfo.resize(ofs.size(), 0);
	for (size_t k = 0; k < ofs.size(); ++k) { xy[k].x = ofs[k]; xy[k].y = coef[k].x; xy[k].z = coef[k].y; xy[k].w = winfo[k]; }
	s = upload(xy.data(), xy.size(), &out->d_xy_tab, owned); if (s) return s;
	for (size_t j = 1; j < dj.size(); ++j) {
		DownJob* p = nullptr;
		s = upload(dj[j].data(), dj[j].size(), &p, owned); if (s) return s;
		int mx = 0;
		for (const DownJob& q : dj[j]) mx = std::max(mx, pyrdown_tiles(q.dst_w, q.dst_h));
		out->d_down.push_back(p); out->n_down.push_back((int)dj[j].size()); out->max_down_px.push_back(mx);
	}
	return FDB_OK;
}

int enqueue_pyramid(fdb_ctx* c, cudaStream_t st, const PyramidJobs& jobs, const uint8_t* d_frames, int W, int H, int n, uint8_t* d_arena,
		int64_t arena_stride, cudaEvent_t ev_mid) {
	if (jobs.n_resize) {
		launch_resize(st, d_frames, W, H, n, d_arena, arena_stride, jobs.d_resize, jobs.n_resize, jobs.max_quads, jobs.d_xy_tab);
		c->launches++;
	}
	if (ev_mid) CUDA_TRY(cudaEventRecord(ev_mid, st));
	for (size_t j = 0; j < jobs.d_down.size(); ++j) {
		if (!jobs.n_down[j]) continue;
		launch_pyrdown(st, d_frames, W, H, n, d_arena, arena_stride, jobs.d_down[j], jobs.n_down[j], jobs.max_down_px[j]);
		c->launches++;
	}
	return FDB_OK;
}

bool encode_tile_maps(const std::vector<PyrImage>& images, const std::vector<int>& which, uint8_t* arena, int64_t arena_stride, int frames,
		std::vector<CUtensorMap>* out) {
	/* encoded through the driver entry point (no link-time libcuda dependency) */
	typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
			const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
			CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
	if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess
			|| qres != cudaDriverEntryPointSuccess || !fn) return false;
	out->assign(which.size(), CUtensorMap());
	std::memset(out->data(), 0, sizeof(CUtensorMap) * out->size());
	for (size_t k = 0; k < which.size(); ++k) {
		const PyrImage& im = images[(size_t)which[k]];
		if (im.offset < 0) continue; /* the frame itself: plain loads */
		const cuuint64_t dims[3] = {(cuuint64_t)im.width, (cuuint64_t)im.height, (cuuint64_t)frames};
		const cuuint64_t strides[2] = {(cuuint64_t)im.pitch, (cuuint64_t)arena_stride};
		const cuuint32_t box[3] = {STRIP_TILE_PITCH, STRIP_TILE_ROWS, 1};
		const cuuint32_t estr[3] = {1, 1, 1};
		if (reinterpret_cast<EncodeFn>(fn)(&(*out)[k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, arena + im.offset, dims, strides, box, estr,
				CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
				CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
	}
	return true;
}

/* work items of the group kernel for one layer: strips of <= 32 window columns; narrow layers pack several row runs side
 * by side (all 32 lanes busy), the runs share the tile's rows. Balanced run length: the same number of window rows for
 * every lane of the layer. One item per strip and pack of <= max_pack models. */
void append_strip_items(const PlanLayer& L, int image, int patch_h, int n_models, const int* models, const int* first_windows,
		int max_pack, std::vector<GroupItem>* out) {
	if (L.windows_x <= 0 || L.windows_y <= 0) return;
	const int budget = STRIP_TILE_ROWS - (patch_h - 1); /* window rows per tile */
	/* balanced packs: 7 models -> 4 + 3, 5 -> 3 + 2 */
	const int cap = std::max(1, std::min(max_pack, GRP_MAX_PACK)), npacks = (n_models + cap - 1) / cap;
	for (int pk = 0, m0 = 0; pk < npacks; ++pk) {
		const int nm = n_models / npacks + (pk < n_models % npacks ? 1 : 0);
		for (int ix0 = 0; ix0 < L.windows_x; ix0 += 32) {
			const int cols = std::min(32, L.windows_x - ix0);
			const int nsub = std::min(WVM_MAXSUB, 32 / cols);
			const int run_cap = std::max(1, std::min(WVM_RUN, budget / nsub));
			const int nruns = (L.windows_y + run_cap - 1) / run_cap;
			const int run = (L.windows_y + nruns - 1) / nruns;
			for (int iy0 = 0; iy0 < L.windows_y; iy0 += nsub * run) {
				GroupItem it{};
				it.image = image; it.begin_x = L.begin_x; it.begin_y = L.begin_y; it.windows_x = L.windows_x; it.windows_y = L.windows_y;
				it.ix0 = ix0; it.iy0 = iy0; it.cols = cols; it.run = run;
				it.nsub = std::min(nsub, (L.windows_y - iy0 + run - 1) / run);
				it.nm = nm;
				for (int k = 0; k < nm; ++k) { it.model[k] = models[m0 + k]; it.first_window[k] = first_windows[m0 + k]; }
				out->push_back(it);
			}
		}
		m0 += nm;
	}
}

int upload_layers(fdb_detector* det, const Plan& plan, DevLayer* dst, cudaStream_t st) {
	std::vector<DevLayer> L(plan.layers.size());
	for (size_t i = 0; i < plan.layers.size(); ++i) {
		const PlanLayer& p = plan.layers[i];
		L[i].offset = plan.images[p.image].offset;
		L[i].width = p.width; L[i].height = p.height;
		L[i].pitch = plan.images[p.image].pitch;
		L[i].begin_x = p.begin_x; L[i].begin_y = p.begin_y;
		L[i].windows_x = p.windows_x; L[i].windows_y = p.windows_y;
		L[i].first_window = (int)p.first_window;
		L[i].tma_ok = det->use_tma && L[i].offset >= 0 ? 1 : 0;
	}
	if (!L.empty())
		CUDA_TRY(cudaMemcpyAsync(dst, L.data(), sizeof(DevLayer) * L.size(), cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaStreamSynchronize(st)); /* L is a stack-owned staging buffer */
	return FDB_OK;
}

/* enqueue pyramid + stage-1 kernels for n frames resident at d_frames, on stream st with slot buffers */
int enqueue_stage1(fdb_detector* det, Slot& sl, cudaStream_t st, const uint8_t* d_frames, int n, const Plan& plan,
		const DevLayer* d_layers, int64_t windows, fdb_window_score* d_dense, uint8_t* d_patches, bool want_candidates,
		bool marks = false) {
	fdb_ctx* c = det->ctx;
	const int W = plan.width, H = plan.height;
	sl.arena = sl.d_arena; sl.arena_stride = plan.arena_bytes;
	CUDA_TRY(cudaMemsetAsync(sl.d_counters, 0, FDB_NCOUNTERS * sizeof(int), st));
	if (marks) CUDA_TRY(cudaEventRecord(c->ev[1], st));
	{ const int r = enqueue_pyramid(c, st, det->jobs, d_frames, W, H, n, sl.d_arena, plan.arena_bytes, marks ? c->ev[2] : nullptr); if (r) return r; }
	if (marks) CUDA_TRY(cudaEventRecord(c->ev[3], st));
	if (windows > 0 && det->wvm) {
		DevWvm m = det->wvm->dev;
		m.step_x = det->desc.step_x; m.step_y = det->desc.step_y;
		if (det->use_strips && d_layers == det->d_layers && !d_patches) {
			GroupModel gm{};
			gm.m = m; gm.dense = d_dense; gm.windows_per_frame = (int)windows; gm.cand_cap = det->cand_cap;
			gm.cand = want_candidates ? sl.d_cand : nullptr; gm.cand_count = sl.d_counters; gm.q = sl.deep;
			GroupArgs ga{};
			ga.items = det->d_gitems; ga.n_items = det->n_gitems; ga.n_frames = n;
			ga.images = det->d_gimages; ga.tmaps = det->use_tma ? sl.d_tmaps : nullptr;
			ga.frames = d_frames; ga.W = W; ga.H = H; ga.arena = sl.d_arena; ga.arena_stride = plan.arena_bytes;
			ga.cursor = sl.d_counters + 3;
			ga.models[0] = gm;
			launch_wvm_group(st, det->desc.patch_width, det->desc.patch_height, 1, ga, m.btc != nullptr);
			if (marks) CUDA_TRY(cudaEventRecord(c->ev[5], st)); /* profiling mark between the two kernels */
			DeepArgs da{};
			da.images = det->d_gimages; da.frames = d_frames; da.W = W; da.H = H; da.arena = sl.d_arena; da.arena_stride = plan.arena_bytes;
			da.n_models = 1; da.models[0] = gm;
			launch_wvm_deep_group(st, da);
		} else {
			if (marks) CUDA_TRY(cudaEventRecord(c->ev[5], st)); /* generic path: no separate deep mark */
			launch_wvm_windows(st, m, d_frames, W, H, n, sl.d_arena, plan.arena_bytes, d_layers, (int)plan.layers.size(),
					(int)windows, d_dense, d_patches, want_candidates ? sl.d_cand : nullptr, sl.d_counters, det->cand_cap, sl.deep);
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

/* stage-1 results -> counters + candidates on their way to the host (async) */
int enqueue_fetch(fdb_detector* det, Slot& sl, cudaStream_t st) {
	CUDA_TRY(cudaMemcpyAsync(sl.h_counters, sl.d_counters, FDB_NCOUNTERS * sizeof(int) + (size_t)det->opt_cand * sizeof(Candidate), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaEventRecord(sl.ev_stage1, st));
	return FDB_OK;
}


/* the second classifier on n work items of the slot's chunk: hq64 patches are rebuilt inside the SVM kernel;
 * any other feature space runs its layer filters once per chunk, then feature kernel + SVM in batches */
void svm_stage(fdb_detector* det, Slot& sl, cudaStream_t st, const Plan& plan, const DevLayer* d_layers, const SvmItem* d_items,
		int n, double* d_dist, bool filter_layers, int* d_level = nullptr) {
	fdb_ctx* c = det->ctx;
	if (!det->has_feature || det->feat.kind == FDB_FEATURE_HQ64) {
		launch_svm_windows(st, det->svm->dev, det->desc.patch_width, det->desc.patch_height, sl.frames_dev, plan.width, plan.height,
				sl.arena, sl.arena_stride, d_layers, d_items, n, d_dist, d_level);
		c->launches++;
		return;
	}
	if (filter_layers && det->feat.layer_channels) {
		launch_feature_layers(st, det->feat, sl.frames_dev, plan.width, plan.height, sl.n, sl.arena, sl.arena_stride, d_layers,
				sl.d_farena, det->farena_bytes);
		c->launches++;
	}
	for (int off = 0; off < n; off += FEAT_BATCH) {
		const int m = std::min(FEAT_BATCH, n - off);
		launch_feature_patches(st, det->feat, sl.frames_dev, plan.width, plan.height, sl.arena, sl.arena_stride, d_layers,
				sl.d_farena, det->farena_bytes, d_items + off, m, sl.d_feat);
		launch_svm_vectors(st, det->svm->dev, sl.d_feat, m, d_dist + off, d_level ? d_level + off : nullptr);
		c->launches += 2;
	}
}

/* phase A of the host post-processing, in three parts so that a detector set can run the middle one for all its members
 * on several host threads:
 *   fetch   wait for stage 1 of the slot's chunk, make the whole candidate list available on the host
 *   host    (pure CPU, touches only det and sl) per-frame candidate lists in canonical order, overlap elimination, SVM work list
 *   launch  SVM on the survivors (async) */
int phase_a_fetch(fdb_detector* det, Slot& sl, cudaStream_t st, const DevLayer* d_layers, int fast_path) {
	CUDA_TRY(cudaEventSynchronize(sl.ev_stage1));
	if (fast_path < 0) fast_path = det->use_strips && d_layers == det->d_layers;
	if (fast_path && sl.h_counters[1] > sl.deep.cap) {
		/* more survivors than the deep queue holds (a model with hardly any early exits): the strip
		 * kernel cannot finish them inline, so this detector switches to the generic kernels for good */
		det->use_strips = false;
		return STATUS_REDO;
	}
	const int ncand = sl.h_counters[0];
	if (ncand > det->cand_cap)
		return fail(FDB_ERR_OVERFLOW, "stage-1 candidate list overflow: raise max_positives_per_frame");
	sl.cand_src = reinterpret_cast<const Candidate*>(sl.h_counters + FDB_NCOUNTERS);
	if (ncand > det->opt_cand) {
		/* the list is complete (stage 1 finished): fetch it on the copy stream - `st` may already hold other work of this chunk
		 * (a detector set queues its members' SVM kernels there) and the host must not wait for that */
		cudaStream_t cs = sl.st_copy ? sl.st_copy : st;
		CUDA_TRY(cudaMemcpyAsync(sl.h_cand_big, sl.d_cand, sizeof(Candidate) * (size_t)ncand, cudaMemcpyDeviceToHost, cs));
		sl.cand_src = sl.h_cand_big;
		sl.cand_pending = true;
	}
	return FDB_OK;
}

int phase_a_fetch_wait(Slot& sl, cudaStream_t st) {
	if (sl.cand_pending) { CUDA_TRY(cudaStreamSynchronize(sl.st_copy ? sl.st_copy : st)); sl.cand_pending = false; }
	return FDB_OK;
}

/* returns FDB_OK or FDB_ERR_OVERFLOW (no message: may run on a worker thread) */
int phase_a_host(fdb_detector* det, Slot& sl, const Plan& plan, int stage) {
	const int ncand = sl.h_counters[0];
	/* canonical order: (frame, window) - SlidingWindowDetector::detect() pushes in extract order */
	std::vector<Candidate> cand(sl.cand_src, sl.cand_src + ncand);
	std::sort(cand.begin(), cand.end(), [](const Candidate& a, const Candidate& b) {
		return a.frame != b.frame ? a.frame < b.frame : a.window < b.window; });
	det->counts[1] += ncand;
	sl.per_frame.assign((size_t)sl.n, std::vector<fdb_detection>());
	for (const Candidate& k : cand) {
		fdb_detection d;
		fill_detection(&d, plan, det->desc, sl.base + k.frame, k.window);
		d.wvm_level = k.level; d.wvm_fout = k.fout;
		d.wvm_probability = wvm_probability(det->wvm->logistic_a, det->wvm->logistic_b, k.fout);
		d.svm_distance = std::numeric_limits<double>::quiet_NaN();
		d.svm_probability = std::numeric_limits<double>::quiet_NaN();
		d.probability = d.wvm_probability;
		d.positive = 1;
		sl.per_frame[(size_t)k.frame].push_back(d);
	}
	if (stage >= FDB_STAGE_OE)
		for (auto& v : sl.per_frame) { overlap_eliminate(v, det->desc.oe_dist, det->desc.oe_ratio); det->counts[2] += (int64_t)v.size(); }
	sl.svm_items = 0;
	if (stage >= FDB_STAGE_SVM && det->svm) {
		size_t total = 0;
		for (auto& v : sl.per_frame) total += v.size();
		if ((int64_t)total > det->items_cap) return FDB_ERR_OVERFLOW;
		size_t k = 0;
		for (int f = 0; f < sl.n; ++f)
			for (const fdb_detection& d : sl.per_frame[(size_t)f]) {
				SvmItem it; it.frame = f; it.layer = d.reserved; it.x = d.x; it.y = d.y;
				sl.h_items[k++] = it;
			}
		sl.svm_items = total;
	}
	return FDB_OK;
}

int phase_a_launch(fdb_detector* det, Slot& sl, cudaStream_t st, const Plan& plan, const DevLayer* d_layers, int stage) {
	const size_t total = sl.svm_items;
	if (stage >= FDB_STAGE_SVM && det->svm && total) {
		CUDA_TRY(cudaMemcpyAsync(sl.d_items, sl.h_items, sizeof(SvmItem) * total, cudaMemcpyHostToDevice, st));
		svm_stage(det, sl, st, plan, d_layers, sl.d_items, (int)total, sl.d_dist, true);
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpyAsync(sl.h_dist, sl.d_dist, sizeof(double) * total, cudaMemcpyDeviceToHost, st));
	}
	CUDA_TRY(cudaEventRecord(sl.ev_svm, st));
	return FDB_OK;
}

int phase_a(fdb_detector* det, Slot& sl, cudaStream_t st, const Plan& plan, const DevLayer* d_layers, int stage, int fast_path) {
	int r = phase_a_fetch(det, sl, st, d_layers, fast_path); if (r) return r;
	r = phase_a_fetch_wait(sl, st); if (r) return r;
	r = phase_a_host(det, sl, plan, stage);
	if (r) return fail(r, "SVM work list overflow");
	return phase_a_launch(det, sl, st, plan, d_layers, stage);
}

/* phase B: SVM distances -> classify, grid NMS, append in frame order; phase_b_host is pure CPU (det, sl and out only) */
void phase_b_host(fdb_detector* det, Slot& sl, const Plan& plan, int stage, bool is_roi, std::vector<fdb_detection>& out) {
	if (stage >= FDB_STAGE_SVM && det->svm) {
		size_t k = 0;
		for (int f = 0; f < sl.n; ++f) {
			std::vector<fdb_detection>& v = sl.per_frame[(size_t)f];
			std::vector<fdb_detection> pos;
			for (fdb_detection& d : v) {
				d.svm_distance = sl.h_dist[k++];
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
	for (auto& v : sl.per_frame)
		for (fdb_detection& d : v) { d.reserved = 0; out.push_back(d); }
	sl.busy = false;
}

int phase_b(fdb_detector* det, Slot& sl, const Plan& plan, int stage, bool is_roi, std::vector<fdb_detection>& out) {
	CUDA_TRY(cudaEventSynchronize(sl.ev_svm));
	phase_b_host(det, sl, plan, stage, is_roi, out);
	return FDB_OK;
}

int ensure_dense(fdb_detector* det) {
	for (int i = 0; i < det->n_slots; ++i)
		if (!det->slots[i].d_dense) {
			int s = dev_alloc(&det->slots[i].d_dense, (size_t)det->chunk * (size_t)std::max<int64_t>(det->plan.windows, 1), det->owned);
			if (s) return s;
		}
	return FDB_OK;
}

int copy_out(const std::vector<fdb_detection>& dets, fdb_detection* out, int64_t cap, int64_t* n_out) {
	if (n_out) *n_out = (int64_t)dets.size();
	if ((int64_t)dets.size() > cap)
		return fail(FDB_ERR_OVERFLOW, "detections_out capacity too small");
	if (!dets.empty() && out) std::memcpy(out, dets.data(), sizeof(fdb_detection) * dets.size());
	return FDB_OK;
}

void release(fdb_detector* det) {
	for (Slot& sl : det->slots) {
		if (sl.st) { cudaStreamSynchronize(sl.st); cudaStreamDestroy(sl.st); }
		if (sl.st_copy) { cudaStreamSynchronize(sl.st_copy); cudaStreamDestroy(sl.st_copy); }
		if (sl.ev_stage1) cudaEventDestroy(sl.ev_stage1);
		if (sl.ev_svm) cudaEventDestroy(sl.ev_svm);
		sl = Slot();
	}
	if (det->ev_begin) { cudaEventDestroy(det->ev_begin); det->ev_begin = nullptr; }
	free_all(det->owned, &det->owned_host);
	free_all(det->es_owned); det->es_cap = 0;
	det->prepared = false;
	det->d_patches = nullptr; det->d_patches_bytes = 0;
	det->jobs = PyramidJobs();
}

bool single_dense_usable(const fdb_detector* det);
int detect_single_dense(fdb_detector* det, const uint8_t* frames, bool frames_on_device, int64_t pitch, int32_t n_frames,
		double* distance_out, bool distance_on_device, fdb_detection* dets_out, int64_t det_cap, int64_t* n_dets);

/* `single` detector of ffpDetectApp.cpp:427-500 with a psvm classifier: SlidingWindowDetector::detect
 * (SlidingWindowDetector.cpp:40-98) where the extractor is a FilteringPyramidFeatureExtractor (the feature space)
 * and the classifier a ProbabilisticSvmClassifier - every window is classified, positives are returned in
 * canonical order with the SVM probability. distance_out: NULL or host [n_frames * windows] distances. */
int detect_single(fdb_detector* det, const uint8_t* frames, bool frames_on_device, int64_t pitch, int32_t n_frames,
		double* distance_out, fdb_detection* dets_out, int64_t det_cap, int64_t* n_dets) {
	if (single_dense_usable(det))
		return detect_single_dense(det, frames, frames_on_device, pitch, n_frames, distance_out, false, dets_out, det_cap, n_dets);
	const Plan& plan = det->plan;
	const int W = plan.width, H = plan.height;
	cudaStream_t st = det->ctx->stream;
	Slot& sl = det->slots[0];
	std::fill(det->counts, det->counts + 5, 0);
	det->counts[0] = plan.windows * n_frames;
	std::vector<fdb_detection> dets;
	const int nwin = (int)plan.windows;
	for (int k = 0; k < n_frames; ++k) {
		if (frames_on_device) sl.frames_dev = frames + (int64_t)k * W * H;
		else {
			CUDA_TRY(cudaMemcpy2DAsync(sl.d_frames, (size_t)W, frames + (int64_t)k * pitch * H, (size_t)pitch, (size_t)W, (size_t)H,
					cudaMemcpyHostToDevice, st));
			sl.frames_dev = sl.d_frames;
		}
		sl.base = k; sl.n = 1;
		int s = enqueue_stage1(det, sl, st, sl.frames_dev, 1, plan, det->d_layers, 0, nullptr, nullptr, false);
		if (s) return s;
		if (nwin > 0) {
			const bool rvm_stage = det->svm->dev.rvm_filters > 0;
			svm_stage(det, sl, st, plan, det->d_layers, det->d_all_items, nwin, det->d_all_dist, true, rvm_stage ? det->d_all_level : nullptr);
			CUDA_TRY(cudaGetLastError());
			CUDA_TRY(cudaMemcpyAsync(det->h_all_dist, det->d_all_dist, sizeof(double) * (size_t)nwin, cudaMemcpyDeviceToHost, st));
			if (rvm_stage) CUDA_TRY(cudaMemcpyAsync(det->h_all_level, det->d_all_level, sizeof(int) * (size_t)nwin, cudaMemcpyDeviceToHost, st));
		}
		CUDA_TRY(cudaStreamSynchronize(st));
		if (distance_out && nwin) std::memcpy(distance_out + (int64_t)k * nwin, det->h_all_dist, sizeof(double) * (size_t)nwin);
		for (int w = 0; w < nwin; ++w) {
			const double dist = det->h_all_dist[w];
			const bool is_rvm = det->svm->dev.rvm_filters > 0;
			int level = -1;
			if (is_rvm) { /* RvmClassifier::classify(pair) (RvmClassifier.cpp:66-73) */
				level = det->h_all_level[w];
				if (!(level + 1 == det->svm->dev.rvm_filters && dist >= (double)det->svm->rvm_thresholds[(size_t)level])) continue;
			} else if (!(dist >= det->svm->dev.threshold)) continue; /* SvmClassifier::classify (SvmClassifier.cpp:44-46) */
			fdb_detection d;
			fill_detection(&d, plan, det->desc, k, w);
			d.reserved = 0;
			d.wvm_level = level;
			d.wvm_fout = std::numeric_limits<float>::quiet_NaN();
			d.wvm_probability = std::numeric_limits<double>::quiet_NaN();
			d.svm_distance = dist;
			d.svm_probability = is_rvm ? rvm_probability(det->svm->logistic_a, det->svm->logistic_b, dist)
					: svm_probability(det->svm->logistic_a, det->svm->logistic_b, dist);
			d.probability = d.svm_probability;
			d.positive = 1;
			dets.push_back(d);
		}
	}
	det->counts[1] = det->counts[2] = det->counts[3] = det->counts[4] = (int64_t)dets.size();
	return copy_out(dets, dets_out, det_cap, n_dets);
}

/* the same detector when the SVM has a tensor-core form and works on HistEq64 patches: whole chunks of frames go
 * through pyramid kernels + ONE svm_dense_kernel launch (every window x every support vector as an u8 matrix product);
 * only the positives (and, if asked for, the distances) come back.
 * distance_out: NULL, host [n_frames * windows] or - distance_on_device - device memory of that size. */
bool single_dense_usable(const fdb_detector* det) {
	return !det->wvm && det->svm && det->svm->has_dense && svm_dense_enabled() && det->d_sd_dist
			&& (!det->has_feature || det->feat.kind == FDB_FEATURE_HQ64)
			&& det->svm->dense.dim == det->desc.patch_width * det->desc.patch_height;
}

int detect_single_dense(fdb_detector* det, const uint8_t* frames, bool frames_on_device, int64_t pitch, int32_t n_frames,
		double* distance_out, bool distance_on_device, fdb_detection* dets_out, int64_t det_cap, int64_t* n_dets) {
	const Plan& plan = det->plan;
	const int W = plan.width, H = plan.height;
	fdb_ctx* c = det->ctx;
	cudaStream_t st = c->stream;
	Slot& sl = det->slots[0];
	std::fill(det->counts, det->counts + 5, 0);
	det->counts[0] = plan.windows * n_frames;
	std::vector<fdb_detection> dets;
	const int64_t nwin = plan.windows;
	det->sd_kernel_ms = 0; det->sd_kernel_launches = 0;
	/* host frames: the upload of chunk k + 1 (copy stream, the other slot's staging buffer) overlaps the kernels of chunk k */
	const bool overlap = !frames_on_device && det->n_slots >= 2;
	cudaStream_t cs = overlap ? det->slots[1].st : st;
	auto upload = [&](int base) -> cudaError_t {
		const int n = std::min(det->chunk, n_frames - base);
		Slot& dst = det->slots[overlap ? (base / det->chunk) & 1 : 0];
		cudaError_t e = cudaMemcpy2DAsync(dst.d_frames, (size_t)W, frames + (int64_t)base * pitch * H, (size_t)pitch, (size_t)W,
				(size_t)H * n, cudaMemcpyHostToDevice, cs);
		if (e == cudaSuccess && overlap) e = cudaEventRecord(dst.ev_stage1, cs);
		return e;
	};
	if (overlap && n_frames > 0) CUDA_TRY(upload(0));
	for (int base = 0; base < n_frames; base += det->chunk) {
		const int n = std::min(det->chunk, n_frames - base);
		if (frames_on_device) sl.frames_dev = frames + (int64_t)base * W * H;
		else {
			Slot& src = det->slots[overlap ? (base / det->chunk) & 1 : 0];
			if (overlap) CUDA_TRY(cudaStreamWaitEvent(st, src.ev_stage1, 0));
			else CUDA_TRY(upload(base));
			sl.frames_dev = src.d_frames;
		}
		sl.base = base; sl.n = n;
		CUDA_TRY(cudaMemsetAsync(det->d_sd_count, 0, sizeof(int), st));
		int s = enqueue_stage1(det, sl, st, sl.frames_dev, n, plan, det->d_layers, 0, nullptr, nullptr, false);
		if (s) return s;
		double* d_out = distance_out && distance_on_device ? distance_out + (int64_t)base * nwin : det->d_sd_dist;
		CUDA_TRY(cudaEventRecord(c->ev[1], st));
		launch_svm_dense_windows(st, det->svm->dense, det->desc.patch_width, det->desc.patch_height, det->desc.step_x,
				det->desc.step_y, sl.frames_dev, W, H, n, sl.d_arena, plan.arena_bytes, det->d_layers, (int)plan.layers.size(),
				nwin, d_out, det->d_sd_count, det->d_sd_pos, det->sd_pos_cap);
		c->launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaEventRecord(c->ev[2], st));
		CUDA_TRY(cudaMemcpyAsync(det->h_sd_count, det->d_sd_count, sizeof(int), cudaMemcpyDeviceToHost, st));
		if (overlap && base + det->chunk < n_frames) CUDA_TRY(upload(base + det->chunk)); /* its buffer was last read by chunk k - 1 (finished) */
		if (distance_out && !distance_on_device)
			CUDA_TRY(cudaMemcpyAsync(distance_out + (int64_t)base * nwin, d_out, sizeof(double) * (size_t)(nwin * n), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		{ float ms = 0.f; CUDA_TRY(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); det->sd_kernel_ms += ms; det->sd_kernel_launches++; }
		const int npos = det->h_sd_count[0];
		if (npos > det->sd_pos_cap) return fail(FDB_ERR_OVERFLOW, "positives list overflow: raise max_positives_per_frame");
		if (npos > 0) {
			CUDA_TRY(cudaMemcpyAsync(det->h_sd_pos, det->d_sd_pos, sizeof(DensePositive) * (size_t)npos, cudaMemcpyDeviceToHost, st));
			CUDA_TRY(cudaStreamSynchronize(st));
			std::sort(det->h_sd_pos, det->h_sd_pos + npos, [](const DensePositive& a, const DensePositive& b) { return a.row < b.row; });
		}
		for (int i = 0; i < npos; ++i) { /* canonical order: SlidingWindowDetector::detect pushes in extract order */
			const DensePositive& p = det->h_sd_pos[i];
			fdb_detection d;
			fill_detection(&d, plan, det->desc, base + (int)(p.row / nwin), p.row % nwin);
			d.reserved = 0;
			d.wvm_level = -1;
			d.wvm_fout = std::numeric_limits<float>::quiet_NaN();
			d.wvm_probability = std::numeric_limits<double>::quiet_NaN();
			d.svm_distance = p.distance;
			d.svm_probability = svm_probability(det->svm->logistic_a, det->svm->logistic_b, p.distance);
			d.probability = d.svm_probability;
			d.positive = 1;
			dets.push_back(d);
		}
	}
	det->counts[1] = det->counts[2] = det->counts[3] = det->counts[4] = (int64_t)dets.size();
	return copy_out(dets, dets_out, det_cap, n_dets);
}

/* condensation::WvmSvmModel::evaluate(image, samples) (WvmSvmModel.cpp:74-119) on one frame: the tracker's sparse use of the
 * same two classifiers. Sample i = {centre x, centre y, width, height} in image pixels:
 *   patch  = DirectPyramidFeatureExtractor::extract(x, y, w, h) (DirectPyramidFeatureExtractor.cpp:67-73,133-147): layer index
 *            round(log(patchWidth / w) / log(incrementalScaleFactor)) (ImagePyramid.cpp:307-310), corner
 *            cvRound((x - w / 2) * scale), cvRound((y - h / 2) * scale); no such layer / outside the layer -> weight 0
 *   equal (layer, corner) = one patch, classified once (CachingPyramidFeatureExtractor + the model's result cache)
 *   weight = 0.5 * P_wvm; the max_svm_patches (reference: 8) WVM-positive patches of highest probability also get the SVM:
 *   target = SVM positive, weight = 2 * weight * P_svm. max_svm_patches <= 0: no cut (= evaluate(Sample&) per sample).
 * std::sort in the reference leaves the order of equal probabilities open; ties keep first-seen order here. */
int evaluate_samples(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, const int32_t* xywh, int64_t n,
		int32_t max_svm_patches, uint8_t* target_out, double* weight_out) {
	const Plan& plan = det->plan;
	fdb_ctx* c = det->ctx;
	cudaStream_t st = c->stream;
	Slot& sl = det->slots[0];
	const int pw = det->desc.patch_width, ph = det->desc.patch_height, npix = pw * ph;
	std::vector<int> sample_patch((size_t)n, -1);
	std::vector<SvmItem> items;
	std::vector<std::pair<int64_t, int>> seen; /* (key, patch index), sorted insert */
	for (int64_t i = 0; i < n; ++i) {
		const int x = xywh[4 * i], y = xywh[4 * i + 1], w = xywh[4 * i + 2], h = xywh[4 * i + 3];
		target_out[i] = 0; weight_out[i] = 0.0;
		if (w <= 0) continue;
		const double power = std::log((double)pw / (double)w) / std::log(plan.incremental_scale_factor);
		const int index = (int)std::round(power);
		int li = -1;
		for (size_t k = 0; k < plan.layers.size(); ++k) if (plan.layers[k].index == index) li = (int)k;
		if (li < 0) continue;
		const PlanLayer& L = plan.layers[(size_t)li];
		const int px = cv_round((x - w / 2) * L.scale), py = cv_round((y - h / 2) * L.scale);
		if (px < 0 || py < 0 || px + pw > L.width || py + ph > L.height) continue; /* DirectPyramidFeatureExtractor.cpp:135-136 */
		const int64_t key = ((int64_t)li << 48) | ((int64_t)py << 24) | (int64_t)px;
		auto pos = std::lower_bound(seen.begin(), seen.end(), std::make_pair(key, -1));
		if (pos == seen.end() || pos->first != key) {
			SvmItem it; it.frame = 0; it.layer = li; it.x = px; it.y = py;
			pos = seen.insert(pos, std::make_pair(key, (int)items.size()));
			items.push_back(it);
		}
		sample_patch[(size_t)i] = pos->second;
	}
	const int m = (int)items.size();
	if (m == 0) return FDB_OK;
	int s = FDB_OK;
	if (m > det->es_cap) {
		free_all(det->es_owned);
		det->es_cap = 0;
		const size_t cap = (size_t)std::max(m, 1024) * 3 / 2;
		s = dev_alloc(&det->d_es_items, cap, det->es_owned);
		if (!s) s = dev_alloc(&det->d_es_patches, cap * npix, det->es_owned);
		if (!s) s = dev_alloc(&det->d_es_scores, cap, det->es_owned);
		if (!s) s = dev_alloc(&det->d_es_dist, cap, det->es_owned);
		if (s) { free_all(det->es_owned); return s; }
		det->es_cap = (int)cap;
	}
	SvmItem* d_items = det->d_es_items; uint8_t* d_patches = det->d_es_patches;
	fdb_window_score* d_scores = det->d_es_scores; double* d_dist = det->d_es_dist;
	std::vector<fdb_window_score> scores((size_t)m);
	auto run = [&]() -> int {
		CUDA_TRY(cudaMemcpy2DAsync(sl.d_frames, (size_t)plan.width, frame_host, (size_t)pitch, (size_t)plan.width, (size_t)plan.height,
				cudaMemcpyHostToDevice, st));
		sl.frames_dev = sl.d_frames; sl.base = 0; sl.n = 1;
		int r = enqueue_stage1(det, sl, st, sl.d_frames, 1, plan, det->d_layers, 0, nullptr, nullptr, false);
		if (r) return r;
		CUDA_TRY(cudaMemcpyAsync(d_items, items.data(), sizeof(SvmItem) * (size_t)m, cudaMemcpyHostToDevice, st));
		launch_hq64_items(st, pw, ph, sl.d_frames, plan.width, plan.height, sl.d_arena, plan.arena_bytes, det->d_layers, d_items, m, d_patches);
		DevWvm wm = det->wvm->dev;
		launch_wvm_patches(st, wm, d_patches, m, d_scores);
		c->launches += 2;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpyAsync(scores.data(), d_scores, sizeof(fdb_window_score) * (size_t)m, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		return FDB_OK;
	};
	s = run();
	if (s) return s;
	/* WVM results per patch (ProbabilisticWvmClassifier::getProbability) */
	std::vector<double> pwvm((size_t)m);
	std::vector<int> remaining; /* WVM-positive patches in first-seen order */
	const fdb_wvm* wv = det->wvm;
	for (int k = 0; k < m; ++k) {
		const fdb_window_score& r = scores[(size_t)k];
		pwvm[(size_t)k] = wvm_probability(wv->logistic_a, wv->logistic_b, r.fout);
		if (r.level + 1 == wv->dev.num_lin && r.fout >= wv->thresholds[(size_t)r.level]) remaining.push_back(k);
	}
	for (int64_t i = 0; i < n; ++i) if (sample_patch[(size_t)i] >= 0) weight_out[i] = 0.5 * pwvm[(size_t)sample_patch[(size_t)i]];
	if (remaining.empty() || !det->svm) return FDB_OK;
	if (max_svm_patches > 0 && (int)remaining.size() > max_svm_patches) {
		std::stable_sort(remaining.begin(), remaining.end(), [&](int a, int b) { return pwvm[(size_t)a] > pwvm[(size_t)b]; });
		remaining.resize((size_t)max_svm_patches);
	}
	/* SVM on the chosen patches: gather their equalised patches on the device side by re-running the item kernel on the subset */
	const int q = (int)remaining.size();
	std::vector<SvmItem> sub((size_t)q);
	for (int k = 0; k < q; ++k) sub[(size_t)k] = items[(size_t)remaining[(size_t)k]];
	std::vector<double> dist((size_t)q);
	auto run2 = [&]() -> int {
		CUDA_TRY(cudaMemcpyAsync(d_items, sub.data(), sizeof(SvmItem) * (size_t)q, cudaMemcpyHostToDevice, st));
		launch_svm_windows(st, det->svm->dev, pw, ph, sl.d_frames, plan.width, plan.height, sl.d_arena, plan.arena_bytes, det->d_layers,
				d_items, q, d_dist);
		c->launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpyAsync(dist.data(), d_dist, sizeof(double) * (size_t)q, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		return FDB_OK;
	};
	s = run2();
	if (s) return s;
	std::vector<int> patch_rank((size_t)m, -1);
	for (int k = 0; k < q; ++k) patch_rank[(size_t)remaining[(size_t)k]] = k;
	for (int64_t i = 0; i < n; ++i) {
		const int pidx = sample_patch[(size_t)i];
		if (pidx < 0 || patch_rank[(size_t)pidx] < 0) continue;
		const double d = dist[(size_t)patch_rank[(size_t)pidx]];
		target_out[i] = d >= det->svm->dev.threshold ? 1 : 0;
		weight_out[i] = 2 * weight_out[i] * svm_probability(det->svm->logistic_a, det->svm->logistic_b, d);
	}
	return FDB_OK;
}


int detect_pipeline(fdb_detector* det, const uint8_t* frames, bool frames_on_device, int64_t pitch, int32_t n_frames,
		int32_t stage, fdb_window_score* dense_out, bool dense_on_device, fdb_detection* dets_out, int64_t det_cap,
		int64_t* n_dets) {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (n_frames < 0 || (n_frames > 0 && !frames)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad frame batch");
	if (stage < FDB_STAGE_WVM || stage > FDB_STAGE_NMS) return fail(FDB_ERR_INVALID_ARGUMENT, "bad stage");
	const Plan& plan = det->plan;
	const int W = plan.width, H = plan.height;
	if (!frames_on_device && pitch < W) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	if (!det->wvm) {
		if (dense_out) return fail(FDB_ERR_INVALID_ARGUMENT, "a detector without a WVM has no stage-1 records (use fdb_detect_single)");
		return detect_single(det, frames, frames_on_device, pitch, n_frames, nullptr, dets_out, det_cap, n_dets);
	}
	std::fill(det->counts, det->counts + 5, 0);
	det->counts[0] = plan.windows * n_frames;
	std::vector<fdb_detection> dets;
	if (dense_out && !dense_on_device) { s = ensure_dense(det); if (s) return s; }
	/* the slot streams start after everything queued on the context stream (timer events, user copies) */
	CUDA_TRY(cudaEventRecord(det->ev_begin, det->ctx->stream));
	for (int i = 0; i < det->n_slots; ++i) CUDA_TRY(cudaStreamWaitEvent(det->slots[i].st, det->ev_begin, 0));

	const int n_chunks = (n_frames + det->chunk - 1) / det->chunk;
	/* software pipeline over chunks: enqueue(i), phase A(i-1), phase B(i-2) */
	int enq = 0, a_done = 0, retired = 0;
	auto do_a = [&]() -> int {
		Slot& sa = det->slots[a_done % det->n_slots];
		const int r = phase_a(det, sa, sa.st, plan, det->d_layers, stage);
		if (r == FDB_OK) ++a_done;
		return r;
	};
	auto do_b = [&]() -> int {
		const int r = phase_b(det, det->slots[retired % det->n_slots], plan, stage, false, dets);
		if (r == FDB_OK) ++retired;
		return r;
	};
	while (enq < n_chunks) {
		Slot& sl = det->slots[enq % det->n_slots];
		while (sl.busy) { /* the slot's previous chunk must be retired before its buffers are reused */
			if (a_done == retired) { s = do_a(); if (s) return s; }
			s = do_b(); if (s) return s;
		}
		sl.base = enq * det->chunk;
		sl.n = std::min(det->chunk, n_frames - sl.base);
		sl.busy = true;
		if (frames_on_device) {
			sl.frames_dev = frames + (int64_t)sl.base * W * H;
		} else {
			if (pitch == W) /* contiguous frames: one linear copy */
				CUDA_TRY(cudaMemcpyAsync(sl.d_frames, frames + (int64_t)sl.base * W * H, (size_t)W * H * sl.n, cudaMemcpyHostToDevice, sl.st));
			else
				CUDA_TRY(cudaMemcpy2DAsync(sl.d_frames, (size_t)W, frames + (int64_t)sl.base * pitch * H, (size_t)pitch, (size_t)W,
						(size_t)H * sl.n, cudaMemcpyHostToDevice, sl.st));
			sl.frames_dev = sl.d_frames;
		}
		fdb_window_score* d_dense = nullptr;
		if (dense_out) d_dense = dense_on_device ? dense_out + (int64_t)sl.base * plan.windows : sl.d_dense;
		s = enqueue_stage1(det, sl, sl.st, sl.frames_dev, sl.n, plan, det->d_layers, plan.windows, d_dense, nullptr, true);
		if (s) return s;
		if (dense_out && !dense_on_device && plan.windows > 0)
			CUDA_TRY(cudaMemcpyAsync(dense_out + (int64_t)sl.base * plan.windows, sl.d_dense,
					sizeof(fdb_window_score) * (size_t)plan.windows * sl.n, cudaMemcpyDeviceToHost, sl.st));
		s = enqueue_fetch(det, sl, sl.st);
		if (s) return s;
		++enq;
		/* while chunk enq-1 runs: post-process its predecessors */
		while (a_done < enq - 1) { s = do_a(); if (s) return s; }
		while (retired < a_done - 1) { s = do_b(); if (s) return s; }
	}
	while (a_done < n_chunks) {
		s = do_a(); if (s) return s;
		while (retired < a_done - 1) { s = do_b(); if (s) return s; }
	}
	while (retired < n_chunks) { s = do_b(); if (s) return s; }
	for (int i = 0; i < det->n_slots; ++i) CUDA_TRY(cudaStreamSynchronize(det->slots[i].st));
	return copy_out(dets, dets_out, det_cap, n_dets);
}

int detect_impl(fdb_detector* det, const uint8_t* frames, bool frames_on_device, int64_t pitch, int32_t n_frames,
		int32_t stage, fdb_window_score* dense_out, bool dense_on_device, fdb_detection* dets_out, int64_t det_cap,
		int64_t* n_dets) {
	int s = detect_pipeline(det, frames, frames_on_device, pitch, n_frames, stage, dense_out, dense_on_device, dets_out, det_cap, n_dets);
	if (s == STATUS_REDO) {
		/* deep-queue overflow on the fast path: drain the slots and run the whole call again on the generic kernels */
		for (int i = 0; i < det->n_slots; ++i) { cudaStreamSynchronize(det->slots[i].st); det->slots[i].busy = false; }
		s = detect_pipeline(det, frames, frames_on_device, pitch, n_frames, stage, dense_out, dense_on_device, dets_out, det_cap, n_dets);
	}
	if (s != FDB_OK && det && det->prepared)
		for (int i = 0; i < det->n_slots; ++i) { cudaStreamSynchronize(det->slots[i].st); det->slots[i].busy = false; }
	return s;
}

} // namespace fdb

extern "C" {

int fdb_detector_create(fdb_ctx* ctx, const fdb_detector_desc* desc, fdb_wvm* wvm, fdb_svm* svm, fdb_detector** out) try {
	int s = check_ctx(ctx); if (s) return s;
	if (!desc || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	if (!wvm && !svm) return fail(FDB_ERR_INVALID_ARGUMENT, "detector needs a classifier");
	fdb_detector_desc d = *desc;
	if (d.step_x == 0) d.step_x = 1;
	if (d.step_y == 0) d.step_y = 1;
	/* DirectPyramidFeatureExtractor.cpp:77-80 */
	if (d.step_x < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "DirectPyramidFeatureExtractor: stepX has to be greater than zero");
	if (d.step_y < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "DirectPyramidFeatureExtractor: stepY has to be greater than zero");
	if (wvm && (d.patch_width != wvm->dev.fsx || d.patch_height != wvm->dev.fsy))
		return fail(FDB_ERR_INVALID_ARGUMENT, "patch size differs from the WVM filter size");
	if (d.patch_width < 1 || d.patch_height < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "bad patch size");
	if (d.max_positives_per_frame <= 0) d.max_positives_per_frame = 4096;
	Plan probe;
	s = build_plan(d, 64, 64, &probe); /* validates the pyramid parameters (ImagePyramid.cpp:84-89) */
	if (s) return s;
	fdb_detector* det = new fdb_detector;
	det->ctx = ctx; det->desc = d; det->wvm = wvm; det->svm = svm;
	*out = det;
	return FDB_OK;
} FDB_API_CATCH

int fdb_detector_create_rvm(fdb_ctx* ctx, const fdb_detector_desc* desc, fdb_rvm* rvm, fdb_detector** out) try {
	if (!rvm) return fail(FDB_ERR_INVALID_ARGUMENT, "null rvm");
	return fdb_detector_create(ctx, desc, nullptr, static_cast<fdb_svm*>(rvm), out);
} FDB_API_CATCH

void fdb_detector_destroy(fdb_detector* det) {
	if (!det) return;
	cudaSetDevice(det->ctx->device);
	cudaStreamSynchronize(det->ctx->stream);
	release(det);
	if (det->d_bgr) cudaFree(det->d_bgr);
	if (det->d_gray) cudaFree(det->d_gray);
	delete det;
}

int fdb_detector_prepare(fdb_detector* det, int32_t width, int32_t height, int32_t max_batch) try {
	if (!det) return fail(FDB_ERR_INVALID_ARGUMENT, "null detector");
	int s = check_ctx(det->ctx); if (s) return s;
	if (max_batch < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "max_batch must be positive");
	CUDA_TRY(cudaStreamSynchronize(det->ctx->stream));
	release(det);
	s = build_plan(det->desc, width, height, &det->plan);
	if (s) return s;
	const Plan& plan = det->plan;
	if (plan.windows >= (int64_t)1 << 31) return fail(FDB_ERR_UNSUPPORTED, "too many windows per frame");
	det->max_batch = max_batch;
	{
		/* the SVM's input must be what its feature space produces (RbfKernel.hpp:33-38 throws on a mismatch) */
		fdb_feature_desc fd{};
		fd.kind = FDB_FEATURE_HQ64;
		if (det->has_feature) fd = det->fdesc;
		FeatureShape sh;
		s = feature_shape(fd, det->desc.patch_width, det->desc.patch_height, &sh); if (s) return s;
		if (det->svm && (det->svm->dev.dim != sh.dim || (det->svm->dev.sv_type == FDB_SV_F32) != (sh.is_float != 0)))
			return fail(FDB_ERR_INVALID_ARGUMENT, "RbfKernel: the SVM's support vectors do not match the feature space (type or dimension)");
		det->farena_bytes = 0;
		if (det->has_feature) { s = feature_build(fd, det->desc.patch_width, det->desc.patch_height, plan, &det->feat, &det->farena_bytes, det->owned); if (s) return s; }
	}
	/* chunking: big batches flow through the slots in quarters so that copies, kernels and host work overlap */
	det->chunk = max_batch >= 16 ? std::min((max_batch + 3) / 4, 64) : max_batch; /* at most 64 frames: candidate lists stay within the in-order copy */
	if (const char* e = std::getenv("FDB_CHUNK_FRAMES")) { /* tuning knob: frames per pipeline chunk */
		const int v = std::atoi(e);
		if (v > 0) det->chunk = std::min(v, (int)max_batch);
	}
	det->n_slots = std::min(PIPE_SLOTS, (max_batch + det->chunk - 1) / det->chunk);
	const int64_t cap64 = (int64_t)det->desc.max_positives_per_frame * det->chunk;
	det->cand_cap = (int)std::max<int64_t>(OPT_CAND, std::min<int64_t>(cap64, (int64_t)1 << 26));
	/* a late copy of a long list queues behind whatever kernels the device is running for the next chunk (measured: the host
	 * then waits for them) - so the copy that travels in stream order right behind stage 1 takes up to 64 K candidates (1 MB) */
	det->opt_cand = std::min(det->cand_cap, 65536);
	det->items_cap = det->cand_cap;

	s = build_pyramid_jobs(plan.images, plan.max_down, width, height, &det->jobs, det->owned); if (s) return s;
	CUDA_TRY(cudaEventCreateWithFlags(&det->ev_begin, cudaEventDisableTiming));
	for (int i = 0; i < det->n_slots; ++i) {
		Slot& sl = det->slots[i];
		CUDA_TRY(cudaStreamCreateWithFlags(&sl.st, cudaStreamNonBlocking));
		CUDA_TRY(cudaStreamCreateWithFlags(&sl.st_copy, cudaStreamNonBlocking));
		CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_stage1, cudaEventDisableTiming));
		CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_svm, cudaEventDisableTiming));
		s = dev_alloc(&sl.d_frames, (size_t)det->chunk * width * height, det->owned); if (s) return s;
		s = dev_alloc(&sl.d_arena, (size_t)det->chunk * (size_t)plan.arena_bytes, det->owned); if (s) return s;
		uint8_t* cbuf = nullptr;
		s = dev_alloc(&cbuf, FDB_NCOUNTERS * sizeof(int) + (size_t)det->cand_cap * sizeof(Candidate), det->owned); if (s) return s;
		sl.d_counters = reinterpret_cast<int*>(cbuf);
		sl.d_cand = reinterpret_cast<Candidate*>(cbuf + FDB_NCOUNTERS * sizeof(int));
		/* deep queue: room for 1/16 of the windows of a chunk (beyond that the detector leaves the fast path) */
		sl.deep.count = sl.d_counters + 1;
		sl.deep.next = sl.d_counters + 2;
		sl.deep.cap = (int)std::max<int64_t>(1024, std::min<int64_t>(plan.windows * det->chunk / 16 + 1024, (int64_t)1 << 24));
		s = dev_alloc(&sl.deep.rec, (size_t)sl.deep.cap, det->owned); if (s) return s;
		s = dev_alloc(&sl.deep.patch, (size_t)sl.deep.cap * (size_t)(det->wvm ? det->wvm->dev.nwords : 1), det->owned); if (s) return s;
		s = dev_alloc(&sl.d_items, (size_t)det->items_cap, det->owned); if (s) return s;
		s = dev_alloc(&sl.d_dist, (size_t)det->items_cap, det->owned); if (s) return s;
		if (det->has_feature && det->feat.kind != FDB_FEATURE_HQ64) {
			if (det->farena_bytes) { s = dev_alloc(&sl.d_farena, (size_t)det->chunk * (size_t)det->farena_bytes, det->owned); if (s) return s; }
			uint8_t* fb = nullptr;
			s = dev_alloc(&fb, (size_t)FEAT_BATCH * (size_t)det->feat.dim * 4, det->owned); if (s) return s;
			sl.d_feat = fb;
		}
		uint8_t* hbuf = nullptr;
		s = host_alloc(&hbuf, FDB_NCOUNTERS * sizeof(int) + (size_t)det->opt_cand * sizeof(Candidate), det->owned_host); if (s) return s;
		sl.h_counters = reinterpret_cast<int*>(hbuf);
		s = host_alloc(&sl.h_cand_big, (size_t)det->cand_cap, det->owned_host); if (s) return s;
		s = host_alloc(&sl.h_items, (size_t)det->items_cap, det->owned_host); if (s) return s;
		s = host_alloc(&sl.h_dist, (size_t)det->items_cap, det->owned_host); if (s) return s;
	}
	/* TMA descriptors for the group kernel's tiles: layer li of slot i is a 3-D u8 tensor {width, height, chunk} with
	 * strides {pitch, arena_bytes}; the box is one warp tile. When the driver entry point is missing the kernel stages
	 * tiles with plain loads. */
	det->use_tma = false;
	{
		const char* env = std::getenv("FDB_NO_TMA");
		if (!(env && env[0] == '1') && group_supported(det->desc.patch_width, det->desc.patch_height)) {
			std::vector<int> which;
			for (const PlanLayer& L : plan.layers) which.push_back(L.image);
			bool ok = true;
			for (int i = 0; i < det->n_slots && ok; ++i) {
				std::vector<CUtensorMap> maps;
				ok = encode_tile_maps(plan.images, which, det->slots[i].d_arena, plan.arena_bytes, det->chunk, &maps);
				if (ok) { s = upload(maps.data(), maps.size(), &det->slots[i].d_tmaps, det->owned); if (s) return s; }
			}
			det->use_tma = ok;
		}
	}
	s = dev_alloc(&det->d_layers, FDB_MAX_LAYERS, det->owned); if (s) return s;
	s = dev_alloc(&det->d_layers_roi, FDB_MAX_LAYERS, det->owned); if (s) return s;
	s = upload_layers(det, plan, det->d_layers, det->ctx->stream); if (s) return s;
	/* work items of the fast path: whole-image scan, step 1, supported window size, <= 4 grey values per filter */
	det->use_strips = det->wvm && det->desc.step_x == 1 && det->desc.step_y == 1 && det->wvm->dev.bfrag != nullptr
			&& group_supported(det->desc.patch_width, det->desc.patch_height) && det->wvm->dev.num_lin > WVM_KA
			&& det->wvm->dev.num_used > WVM_KA;
	std::vector<GroupItem> items;
	std::vector<GroupImage> gimages(plan.layers.size());
	for (size_t li = 0; li < plan.layers.size(); ++li) {
		const PyrImage& im = plan.images[plan.layers[li].image];
		gimages[li].offset = im.offset; gimages[li].width = im.width; gimages[li].height = im.height; gimages[li].pitch = im.pitch;
		gimages[li].tma_ok = det->use_tma && im.offset >= 0 ? 1 : 0;
	}
	if (det->use_strips)
		for (size_t li = 0; li < plan.layers.size(); ++li) {
			const int model0 = 0, first = (int)plan.layers[li].first_window;
			append_strip_items(plan.layers[li], (int)li, det->desc.patch_height, 1, &model0, &first, 1, &items);
		}
	det->n_gitems = (int)items.size();
	s = upload(items.data(), items.size(), &det->d_gitems, det->owned); if (s) return s;
	s = upload(gimages.data(), gimages.size(), &det->d_gimages, det->owned); if (s) return s;
	det->d_all_items = nullptr; det->d_all_dist = nullptr; det->h_all_dist = nullptr; det->d_all_level = nullptr; det->h_all_level = nullptr;
	if (!det->wvm) {
		/* `single` detector: every window of a frame is an SVM work item, canonical order */
		std::vector<SvmItem> all((size_t)plan.windows);
		size_t k = 0;
		for (size_t li = 0; li < plan.layers.size(); ++li) {
			const PlanLayer& L = plan.layers[li];
			for (int iy = 0; iy < L.windows_y; ++iy)
				for (int ix = 0; ix < L.windows_x; ++ix) {
					SvmItem it; it.frame = 0; it.layer = (int)li;
					it.x = L.begin_x + ix * det->desc.step_x; it.y = L.begin_y + iy * det->desc.step_y;
					all[k++] = it;
				}
		}
		s = upload(all.data(), all.size(), &det->d_all_items, det->owned); if (s) return s;
		s = dev_alloc(&det->d_all_dist, (size_t)std::max<int64_t>(plan.windows, 1), det->owned); if (s) return s;
		s = host_alloc(&det->h_all_dist, (size_t)std::max<int64_t>(plan.windows, 1), det->owned_host); if (s) return s;
		s = dev_alloc(&det->d_all_level, (size_t)std::max<int64_t>(plan.windows, 1), det->owned); if (s) return s;
		s = host_alloc(&det->h_all_level, (size_t)std::max<int64_t>(plan.windows, 1), det->owned_host); if (s) return s;
		det->d_sd_dist = nullptr;
		if (det->svm->has_dense && plan.windows > 0) {
			det->sd_pos_cap = (int)std::min<int64_t>(plan.windows * det->chunk, (int64_t)1 << 26); /* every window may be positive */
			s = dev_alloc(&det->d_sd_dist, (size_t)plan.windows * det->chunk, det->owned); if (s) return s;
			s = dev_alloc(&det->d_sd_count, 4, det->owned); if (s) return s;
			s = dev_alloc(&det->d_sd_pos, (size_t)det->sd_pos_cap, det->owned); if (s) return s;
			s = host_alloc(&det->h_sd_count, 4, det->owned_host); if (s) return s;
			s = host_alloc(&det->h_sd_pos, (size_t)det->sd_pos_cap, det->owned_host); if (s) return s;
		}
	}
	det->prepared = true;
	return FDB_OK;
} FDB_API_CATCH

int fdb_detector_layers(fdb_detector* det, fdb_layer_info* out, int32_t cap, int32_t* n_layers) try {
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
} FDB_API_CATCH

int64_t fdb_detector_windows_per_frame(fdb_detector* det) { return det && det->prepared ? det->plan.windows : -1; }
int64_t fdb_detector_pyramid_bytes(fdb_detector* det) {
	if (!det || !det->prepared) return -1;
	int64_t b = 0;
	for (const PyrImage& im : det->plan.images) if (im.kind != IMG_FRAME) b += (int64_t)im.pitch * im.height;
	return b;
}

int fdb_detect_batch(fdb_detector* det, const uint8_t* frames_host, int64_t pitch, int32_t n_frames, int32_t stage,
		fdb_window_score* dense_out, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) try {
	return detect_impl(det, frames_host, false, pitch, n_frames, stage, dense_out, false, detections_out, det_cap, n_detections);
} FDB_API_CATCH

int fdb_detect_batch_device(fdb_detector* det, const uint8_t* frames_device, int32_t n_frames, int32_t stage,
		fdb_window_score* dense_out_device, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) try {
	return detect_impl(det, frames_device, true, 0, n_frames, stage, dense_out_device, true, detections_out, det_cap, n_detections);
} FDB_API_CATCH

/* GrayscaleFilter::applyTo on the device for n frames of interleaved 8-bit BGR in host memory -> d_gray (contiguous) */
static int upload_bgr_as_gray(cudaStream_t st, const uint8_t* bgr_host, int64_t pitch, int W, int H, int n, uint8_t* d_bgr, uint8_t* d_gray) {
	if (pitch == 3ll * W) CUDA_TRY(cudaMemcpyAsync(d_bgr, bgr_host, (size_t)3 * W * H * n, cudaMemcpyHostToDevice, st));
	else CUDA_TRY(cudaMemcpy2DAsync(d_bgr, (size_t)3 * W, bgr_host, (size_t)pitch, (size_t)3 * W, (size_t)H * n, cudaMemcpyHostToDevice, st));
	launch_bgr2gray(st, d_bgr, d_gray, (int64_t)W * H * n);
	CUDA_TRY(cudaGetLastError());
	return FDB_OK;
}

int fdb_gray_from_bgr(fdb_ctx* ctx, const uint8_t* bgr_host, int64_t pitch, int32_t W, int32_t H, int32_t n_frames, uint8_t* gray_host) try {
	int s = check_ctx(ctx); if (s) return s;
	if (!bgr_host || !gray_host || W < 1 || H < 1 || n_frames < 0 || pitch < 3ll * W) return fail(FDB_ERR_INVALID_ARGUMENT, "bad BGR frame batch");
	if (n_frames == 0) return FDB_OK;
	const size_t px = (size_t)W * H * n_frames;
	uint8_t* d = nullptr;
	CUDA_TRY(cudaMalloc((void**)&d, 4 * px + 16));
	s = upload_bgr_as_gray(ctx->stream, bgr_host, pitch, W, H, n_frames, d + ((px + 15) & ~(size_t)15), d);
	ctx->launches += 1;
	cudaError_t e = s ? cudaSuccess : cudaMemcpyAsync(gray_host, d, px, cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	cudaFree(d);
	if (s) return s;
	if (e != cudaSuccess) return fail(FDB_ERR_CUDA, std::string("fdb_gray_from_bgr: ") + cudaGetErrorString(e));
	return FDB_OK;
} FDB_API_CATCH

int fdb_detect_batch_bgr(fdb_detector* det, const uint8_t* bgr_host, int64_t pitch, int32_t n_frames, int32_t stage,
		fdb_window_score* dense_out, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	const int W = det->plan.width, H = det->plan.height;
	if (n_frames < 0 || (n_frames > 0 && !bgr_host) || pitch < 3ll * W) return fail(FDB_ERR_INVALID_ARGUMENT, "bad BGR frame batch");
	const int64_t px = (int64_t)W * H * n_frames;
	if (px > det->bgr_cap_px) {
		cudaStreamSynchronize(det->ctx->stream);
		if (det->d_bgr) cudaFree(det->d_bgr);
		if (det->d_gray) cudaFree(det->d_gray);
		det->d_bgr = det->d_gray = nullptr; det->bgr_cap_px = 0;
		CUDA_TRY(cudaMalloc((void**)&det->d_bgr, (size_t)3 * px));
		CUDA_TRY(cudaMalloc((void**)&det->d_gray, (size_t)px));
		det->bgr_cap_px = px;
	}
	if (n_frames > 0) {
		s = upload_bgr_as_gray(det->ctx->stream, bgr_host, pitch, W, H, n_frames, det->d_bgr, det->d_gray); if (s) return s;
		det->ctx->launches += 1;
		CUDA_TRY(cudaStreamSynchronize(det->ctx->stream)); /* the pipeline slots run on their own streams */
	}
	return detect_impl(det, det->d_gray, true, 0, n_frames, stage, dense_out, false, detections_out, det_cap, n_detections);
} FDB_API_CATCH

int fdb_detect_enqueue_device(fdb_detector* det, const uint8_t* frames_device, int32_t n_frames, fdb_window_score* dense_out_device) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (!det->wvm) return fail(FDB_ERR_INVALID_ARGUMENT, "this entry point needs a detector with a WVM first stage");
	if (n_frames < 0 || n_frames > det->max_batch) return fail(FDB_ERR_INVALID_ARGUMENT, "n_frames exceeds the prepared batch");
	const Plan& plan = det->plan;
	for (int base = 0; base < n_frames; base += det->chunk) {
		const int n = std::min(det->chunk, n_frames - base);
		s = enqueue_stage1(det, det->slots[0], det->ctx->stream, frames_device + (int64_t)base * plan.width * plan.height, n, plan,
				det->d_layers, plan.windows, dense_out_device ? dense_out_device + (int64_t)base * plan.windows : nullptr, nullptr, true);
		if (s) return s;
	}
	return FDB_OK;
} FDB_API_CATCH

int fdb_detect_profile_device(fdb_detector* det, const uint8_t* frames_device, int32_t n_frames, double ms_out[6]) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (!det->wvm) return fail(FDB_ERR_INVALID_ARGUMENT, "this entry point needs a detector with a WVM first stage");
	if (n_frames < 0 || n_frames > det->max_batch || !ms_out) return fail(FDB_ERR_INVALID_ARGUMENT, "bad arguments");
	fdb_ctx* c = det->ctx;
	const Plan& plan = det->plan;
	for (int k = 0; k < 6; ++k) ms_out[k] = 0;
	for (int base = 0; base < n_frames; base += det->chunk) {
		const int n = std::min(det->chunk, n_frames - base);
		s = enqueue_stage1(det, det->slots[0], c->stream, frames_device + (int64_t)base * plan.width * plan.height, n, plan,
				det->d_layers, plan.windows, nullptr, nullptr, true, true);
		if (s) return s;
		CUDA_TRY(cudaEventSynchronize(c->ev[4]));
		float a = 0, b = 0, w = 0, d = 0, t = 0;
		CUDA_TRY(cudaEventElapsedTime(&a, c->ev[1], c->ev[2]));
		CUDA_TRY(cudaEventElapsedTime(&b, c->ev[2], c->ev[3]));
		CUDA_TRY(cudaEventElapsedTime(&w, c->ev[3], c->ev[5]));
		CUDA_TRY(cudaEventElapsedTime(&d, c->ev[5], c->ev[4]));
		CUDA_TRY(cudaEventElapsedTime(&t, c->ev[1], c->ev[4]));
		ms_out[0] += a; ms_out[1] += b; ms_out[2] += w; ms_out[3] += d; ms_out[4] += t; ms_out[5] += 1;
	}
	return FDB_OK;
} FDB_API_CATCH

int fdb_detect_roi(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, int32_t roi_x, int32_t roi_y, int32_t roi_w,
		int32_t roi_h, int32_t stage, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (!det->wvm) return fail(FDB_ERR_INVALID_ARGUMENT, "this entry point needs a detector with a WVM first stage");
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
	Slot& sl = det->slots[0];
	s = upload_layers(det, plan, det->d_layers_roi, st); if (s) return s;
	CUDA_TRY(cudaMemcpy2DAsync(sl.d_frames, (size_t)W, frame_host, (size_t)pitch, (size_t)W, (size_t)H, cudaMemcpyHostToDevice, st));
	std::fill(det->counts, det->counts + 5, 0);
	det->counts[0] = windows;
	sl.base = 0; sl.n = 1; sl.frames_dev = sl.d_frames;
	s = enqueue_stage1(det, sl, st, sl.d_frames, 1, plan, det->d_layers_roi, windows, nullptr, nullptr, true);
	if (s) return s;
	s = enqueue_fetch(det, sl, st); if (s) return s;
	std::vector<fdb_detection> dets;
	s = phase_a(det, sl, st, plan, det->d_layers_roi, stage); if (s) return s;
	s = phase_b(det, sl, plan, stage, is_roi, dets); if (s) return s;
	return copy_out(dets, detections_out, det_cap, n_detections);
} FDB_API_CATCH

int fdb_extract_patches(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, uint8_t* patches_out, int64_t cap_windows,
		int64_t* n_windows) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (!det->wvm) return fail(FDB_ERR_INVALID_ARGUMENT, "this entry point needs a detector with a WVM first stage");
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
	Slot& sl = det->slots[0];
	CUDA_TRY(cudaMemcpy2DAsync(sl.d_frames, (size_t)plan.width, frame_host, (size_t)pitch, (size_t)plan.width, (size_t)plan.height,
			cudaMemcpyHostToDevice, st));
	s = enqueue_stage1(det, sl, st, sl.d_frames, 1, plan, det->d_layers, plan.windows, nullptr, det->d_patches, false);
	if (s) return s;
	if (bytes) CUDA_TRY(cudaMemcpyAsync(patches_out, det->d_patches, (size_t)bytes, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	return FDB_OK;
} FDB_API_CATCH

int fdb_pyramid_layer(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, int32_t layer_index, uint8_t* out, int64_t cap) try {
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
	Slot& sl = det->slots[0];
	CUDA_TRY(cudaMemcpy2DAsync(sl.d_frames, (size_t)plan.width, frame_host, (size_t)pitch, (size_t)plan.width, (size_t)plan.height,
			cudaMemcpyHostToDevice, st));
	s = enqueue_stage1(det, sl, st, sl.d_frames, 1, plan, det->d_layers, 0, nullptr, nullptr, false);
	if (s) return s;
	const PyrImage& im = plan.images[(size_t)L->image];
	const uint8_t* src = im.kind == IMG_FRAME ? sl.d_frames : sl.d_arena + im.offset;
	CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)L->width, src, (size_t)im.pitch, (size_t)L->width, (size_t)L->height, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	return FDB_OK;
} FDB_API_CATCH

int fdb_feature_shape(const fdb_feature_desc* desc, int32_t patch_width, int32_t patch_height, int32_t* dim, int32_t* is_float) try {
	if (!desc) return fail(FDB_ERR_INVALID_ARGUMENT, "null feature descriptor");
	FeatureShape sh;
	int s = feature_shape(*desc, patch_width, patch_height, &sh); if (s) return s;
	if (dim) *dim = sh.dim;
	if (is_float) *is_float = sh.is_float;
	return FDB_OK;
} FDB_API_CATCH

int fdb_detector_set_feature(fdb_detector* det, const fdb_feature_desc* desc) try {
	if (!det || !desc) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	int s = check_ctx(det->ctx); if (s) return s;
	FeatureShape sh;
	s = feature_shape(*desc, det->desc.patch_width, det->desc.patch_height, &sh); if (s) return s;
	CUDA_TRY(cudaStreamSynchronize(det->ctx->stream));
	release(det); /* buffers depend on the feature space: prepare again */
	det->fdesc = *desc;
	det->has_feature = true;
	return FDB_OK;
} FDB_API_CATCH

int fdb_extract_features(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, const int32_t* layer_x_y, int64_t n, void* out) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (!frame_host || (n > 0 && (!layer_x_y || !out))) return fail(FDB_ERR_INVALID_ARGUMENT, "null buffer");
	if (!det->has_feature || det->feat.kind == FDB_FEATURE_HQ64)
		return fail(FDB_ERR_INVALID_ARGUMENT, "no feature space set (hq64 patches: fdb_extract_patches)");
	const Plan& plan = det->plan;
	if (pitch < plan.width) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	cudaStream_t st = det->ctx->stream;
	Slot& sl = det->slots[0];
	const int pw = det->desc.patch_width, ph = det->desc.patch_height;
	std::vector<SvmItem> items((size_t)n);
	for (int64_t i = 0; i < n; ++i) {
		int li = -1;
		for (size_t k = 0; k < plan.layers.size(); ++k) if (plan.layers[k].index == layer_x_y[3 * i]) li = (int)k;
		/* DirectPyramidFeatureExtractor.cpp:133-136: out-of-bounds windows yield no patch */
		if (li < 0) return fail(FDB_ERR_INVALID_ARGUMENT, "no such pyramid layer");
		const int x = layer_x_y[3 * i + 1], y = layer_x_y[3 * i + 2];
		if (x < 0 || y < 0 || x + pw > plan.layers[(size_t)li].width || y + ph > plan.layers[(size_t)li].height)
			return fail(FDB_ERR_INVALID_ARGUMENT, "window outside the pyramid layer");
		SvmItem it; it.frame = 0; it.layer = li; it.x = x; it.y = y;
		items[(size_t)i] = it;
	}
	CUDA_TRY(cudaMemcpy2DAsync(sl.d_frames, (size_t)plan.width, frame_host, (size_t)pitch, (size_t)plan.width, (size_t)plan.height,
			cudaMemcpyHostToDevice, st));
	sl.frames_dev = sl.d_frames; sl.base = 0; sl.n = 1;
	s = enqueue_stage1(det, sl, st, sl.d_frames, 1, plan, det->d_layers, 0, nullptr, nullptr, false);
	if (s) return s;
	if (det->feat.layer_channels) {
		launch_feature_layers(st, det->feat, sl.d_frames, plan.width, plan.height, 1, sl.d_arena, plan.arena_bytes, det->d_layers,
				sl.d_farena, det->farena_bytes);
		det->ctx->launches++;
	}
	const size_t vec_bytes = (size_t)det->feat.dim * (det->feat.is_float ? 4 : 1);
	const int batch = std::min<int>(FEAT_BATCH, det->items_cap);
	for (int64_t off = 0; off < n; off += batch) {
		const int m = (int)std::min<int64_t>(batch, n - off);
		CUDA_TRY(cudaMemcpyAsync(sl.d_items, items.data() + off, sizeof(SvmItem) * (size_t)m, cudaMemcpyHostToDevice, st));
		launch_feature_patches(st, det->feat, sl.d_frames, plan.width, plan.height, sl.d_arena, plan.arena_bytes, det->d_layers,
				sl.d_farena, det->farena_bytes, sl.d_items, m, sl.d_feat);
		det->ctx->launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpyAsync((uint8_t*)out + (size_t)off * vec_bytes, sl.d_feat, vec_bytes * (size_t)m, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	return FDB_OK;
} FDB_API_CATCH

int fdb_detect_face_features(fdb_detector* face, fdb_detector* const* features, int32_t n_features, const uint8_t* frame_host,
		int64_t pitch, fdb_detection* face_out, int64_t face_cap, int64_t* n_face, fdb_detection* feature_out,
		int64_t feature_cap_each, int64_t* n_feature) try {
	if (!face || !n_face || (n_features > 0 && (!features || !n_feature))) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	for (int i = 0; i < n_features; ++i) n_feature[i] = 0;
	/* ffpDetectApp.cpp:555-557: facePatches = detector->detect(img) */
	int s = fdb_detect_batch(face, frame_host, pitch, 1, face->svm ? FDB_STAGE_NMS : FDB_STAGE_WVM, nullptr, face_out, face_cap, n_face);
	if (s) return s;
	if (*n_face == 0) return FDB_OK; /* the reference indexes facePatches[0] unconditionally (ffpDetectApp.cpp:591): no face, no ROI */
	/* Patch::getBounds() of the first (most probable) face: Rect(x - width / 2, y - height / 2, width, height) */
	const fdb_detection& f = face_out[0];
	const int rx = f.center_x - f.width / 2, ry = f.center_y - f.height / 2;
	for (int i = 0; i < n_features; ++i) { /* ffpDetectApp.cpp:589-596: detector->detect(img, bounds) */
		fdb_detector* d = features[i];
		if (!d) return fail(FDB_ERR_INVALID_ARGUMENT, "null feature detector");
		s = fdb_detect_roi(d, frame_host, pitch, rx, ry, f.width, f.height, d->svm ? FDB_STAGE_NMS : FDB_STAGE_WVM,
				feature_out + (int64_t)i * feature_cap_each, feature_cap_each, &n_feature[i]);
		if (s) return s;
	}
	return FDB_OK;
} FDB_API_CATCH

int fdb_evaluate_samples(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, const int32_t* samples_xywh, int64_t n,
		int32_t max_svm_patches, uint8_t* target_out, double* weight_out) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (!det->wvm) return fail(FDB_ERR_INVALID_ARGUMENT, "fdb_evaluate_samples needs a detector with a WVM");
	if (det->has_feature && det->feat.kind != FDB_FEATURE_HQ64)
		return fail(FDB_ERR_UNSUPPORTED, "fdb_evaluate_samples: both classifiers work on the HistEq64 patch (WvmSvmModel.cpp:53-59)");
	if (n < 0 || !frame_host || (n > 0 && (!samples_xywh || !target_out || !weight_out))) return fail(FDB_ERR_INVALID_ARGUMENT, "null buffer");
	if (pitch < det->plan.width) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	return evaluate_samples(det, frame_host, pitch, samples_xywh, n, max_svm_patches, target_out, weight_out);
} FDB_API_CATCH

int fdb_detect_single(fdb_detector* det, const uint8_t* frames_host, int64_t pitch, int32_t n_frames, double* distance_out,
		fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (det->wvm || !det->svm) return fail(FDB_ERR_INVALID_ARGUMENT, "fdb_detect_single needs a detector created with an SVM only");
	if (n_frames < 0 || (n_frames > 0 && !frames_host)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad frame batch");
	if (pitch < det->plan.width) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	return detect_single(det, frames_host, false, pitch, n_frames, distance_out, detections_out, det_cap, n_detections);
} FDB_API_CATCH

/* SlidingWindowDetector::detect(image, roi) (SlidingWindowDetector.cpp:53-79) of a `single` detector: the windows
 * PyramidFeatureExtractor::extract(stepX, stepY, roi) visits (DirectPyramidFeatureExtractor.cpp:84-121), every one classified,
 * positives in extract order - what ffpDetectApp.cpp:591 calls for every feature detector inside the face box */
int fdb_detect_single_roi(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, int32_t roi_x, int32_t roi_y, int32_t roi_w,
		int32_t roi_h, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (det->wvm || !det->svm) return fail(FDB_ERR_INVALID_ARGUMENT, "fdb_detect_single_roi needs a detector created with an SVM or RVM only");
	if (!frame_host) return fail(FDB_ERR_INVALID_ARGUMENT, "null frame");
	Plan plan = det->plan;
	const int W = plan.width, H = plan.height;
	if (pitch < W) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	const int64_t windows = enumerate_windows(&plan, det->desc.patch_width, det->desc.patch_height, det->desc.step_x, det->desc.step_y,
			roi_x, roi_y, roi_w, roi_h);
	plan.windows = windows;
	std::fill(det->counts, det->counts + 5, 0);
	det->counts[0] = windows;
	std::vector<fdb_detection> dets;
	if (windows > 0) {
		cudaStream_t st = det->ctx->stream;
		Slot& sl = det->slots[0];
		std::vector<SvmItem> items((size_t)windows);
		size_t k = 0;
		for (size_t li = 0; li < plan.layers.size(); ++li) {
			const PlanLayer& L = plan.layers[li];
			for (int iy = 0; iy < L.windows_y; ++iy)
				for (int ix = 0; ix < L.windows_x; ++ix) {
					SvmItem it; it.frame = 0; it.layer = (int)li;
					it.x = L.begin_x + ix * det->desc.step_x; it.y = L.begin_y + iy * det->desc.step_y;
					items[k++] = it;
				}
		}
		CUDA_TRY(cudaMemcpy2DAsync(sl.d_frames, (size_t)W, frame_host, (size_t)pitch, (size_t)W, (size_t)H, cudaMemcpyHostToDevice, st));
		sl.frames_dev = sl.d_frames; sl.base = 0; sl.n = 1;
		s = enqueue_stage1(det, sl, st, sl.d_frames, 1, det->plan, det->d_layers, 0, nullptr, nullptr, false);
		if (s) return s;
		const bool is_rvm = det->svm->dev.rvm_filters > 0;
		/* d_all_items / d_all_dist hold a whole-image scan: a ROI never has more windows */
		CUDA_TRY(cudaMemcpyAsync(det->d_all_items, items.data(), sizeof(SvmItem) * (size_t)windows, cudaMemcpyHostToDevice, st));
		svm_stage(det, sl, st, det->plan, det->d_layers, det->d_all_items, (int)windows, det->d_all_dist, true, is_rvm ? det->d_all_level : nullptr);
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpyAsync(det->h_all_dist, det->d_all_dist, sizeof(double) * (size_t)windows, cudaMemcpyDeviceToHost, st));
		if (is_rvm) CUDA_TRY(cudaMemcpyAsync(det->h_all_level, det->d_all_level, sizeof(int) * (size_t)windows, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		for (int64_t w = 0; w < windows; ++w) {
			const double dist = det->h_all_dist[w];
			int level = -1;
			if (is_rvm) {
				level = det->h_all_level[w];
				if (!(level + 1 == det->svm->dev.rvm_filters && dist >= (double)det->svm->rvm_thresholds[(size_t)level])) continue;
			} else if (!(dist >= det->svm->dev.threshold)) continue;
			fdb_detection d;
			fill_detection(&d, plan, det->desc, 0, w);
			d.reserved = 0;
			d.wvm_level = level;
			d.wvm_fout = std::numeric_limits<float>::quiet_NaN();
			d.wvm_probability = std::numeric_limits<double>::quiet_NaN();
			d.svm_distance = dist;
			d.svm_probability = is_rvm ? rvm_probability(det->svm->logistic_a, det->svm->logistic_b, dist)
					: svm_probability(det->svm->logistic_a, det->svm->logistic_b, dist);
			d.probability = d.svm_probability;
			d.positive = 1;
			dets.push_back(d);
		}
		/* the whole-image item table of fdb_detect_single lives in d_all_items: restore it */
		std::vector<SvmItem> all((size_t)det->plan.windows);
		size_t q = 0;
		for (size_t li = 0; li < det->plan.layers.size(); ++li) {
			const PlanLayer& L = det->plan.layers[li];
			for (int iy = 0; iy < L.windows_y; ++iy)
				for (int ix = 0; ix < L.windows_x; ++ix) {
					SvmItem it; it.frame = 0; it.layer = (int)li;
					it.x = L.begin_x + ix * det->desc.step_x; it.y = L.begin_y + iy * det->desc.step_y;
					all[q++] = it;
				}
		}
		if (!all.empty()) CUDA_TRY(cudaMemcpy(det->d_all_items, all.data(), sizeof(SvmItem) * all.size(), cudaMemcpyHostToDevice));
	}
	det->counts[1] = det->counts[2] = det->counts[3] = det->counts[4] = (int64_t)dets.size();
	return copy_out(dets, detections_out, det_cap, n_detections);
} FDB_API_CATCH

/* PyramidFeatureExtractor::extract(layer, x, y) / extract(x, y, w, h) for a list of windows of one frame
 * (DirectPyramidFeatureExtractor.cpp:125-143): layer_x_y = n triples {pyramid layer index, window corner x, y inside the layer};
 * out = n vectors of the detector's patch filter (HistEq64 patches, or the feature space set with fdb_detector_set_feature);
 * valid_out[i] = 0 where the reference returns an empty pointer (no such layer, window not inside the layer image) - such
 * vectors are left untouched. */
int fdb_extract_windows(fdb_detector* det, const uint8_t* frame_host, int64_t pitch, const int32_t* layer_x_y, int64_t n, void* out,
		uint8_t* valid_out) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (!frame_host || n < 0 || (n > 0 && (!layer_x_y || !out || !valid_out))) return fail(FDB_ERR_INVALID_ARGUMENT, "null buffer");
	const Plan& plan = det->plan;
	if (pitch < plan.width) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	const int pw = det->desc.patch_width, ph = det->desc.patch_height;
	const bool hq = !det->has_feature || det->feat.kind == FDB_FEATURE_HQ64;
	const size_t vec_bytes = hq ? (size_t)pw * ph : (size_t)det->feat.dim * (det->feat.is_float ? 4 : 1);
	std::vector<SvmItem> items;
	std::vector<int64_t> where;
	for (int64_t i = 0; i < n; ++i) {
		valid_out[i] = 0;
		int li = -1;
		for (size_t k = 0; k < plan.layers.size(); ++k) if (plan.layers[k].index == layer_x_y[3 * i]) li = (int)k;
		if (li < 0) continue;
		const int x = layer_x_y[3 * i + 1], y = layer_x_y[3 * i + 2];
		if (x < 0 || y < 0 || x + pw > plan.layers[(size_t)li].width || y + ph > plan.layers[(size_t)li].height) continue; /* :135-136 */
		SvmItem it; it.frame = 0; it.layer = li; it.x = x; it.y = y;
		items.push_back(it); where.push_back(i);
		valid_out[i] = 1;
	}
	if (items.empty()) return FDB_OK;
	cudaStream_t st = det->ctx->stream;
	Slot& sl = det->slots[0];
	CUDA_TRY(cudaMemcpy2DAsync(sl.d_frames, (size_t)plan.width, frame_host, (size_t)pitch, (size_t)plan.width, (size_t)plan.height,
			cudaMemcpyHostToDevice, st));
	sl.frames_dev = sl.d_frames; sl.base = 0; sl.n = 1;
	s = enqueue_stage1(det, sl, st, sl.d_frames, 1, plan, det->d_layers, 0, nullptr, nullptr, false);
	if (s) return s;
	if (!hq && det->feat.layer_channels) {
		launch_feature_layers(st, det->feat, sl.d_frames, plan.width, plan.height, 1, sl.d_arena, plan.arena_bytes, det->d_layers,
				sl.d_farena, det->farena_bytes);
		det->ctx->launches++;
	}
	const int batch = std::min<int>(hq ? 4096 : FEAT_BATCH, det->items_cap);
	std::vector<void*> tmp;
	uint8_t* d_vec = nullptr;
	if (hq) { s = dev_alloc(&d_vec, (size_t)batch * vec_bytes, tmp); if (s) return s; }
	std::vector<uint8_t> host((size_t)batch * vec_bytes);
	for (size_t off = 0; off < items.size(); off += (size_t)batch) {
		const int m = (int)std::min<size_t>((size_t)batch, items.size() - off);
		cudaError_t e = cudaMemcpyAsync(sl.d_items, items.data() + off, sizeof(SvmItem) * (size_t)m, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess) {
			if (hq) launch_hq64_items(st, pw, ph, sl.d_frames, plan.width, plan.height, sl.d_arena, plan.arena_bytes, det->d_layers, sl.d_items, m, d_vec);
			else launch_feature_patches(st, det->feat, sl.d_frames, plan.width, plan.height, sl.d_arena, plan.arena_bytes, det->d_layers,
					sl.d_farena, det->farena_bytes, sl.d_items, m, sl.d_feat);
			det->ctx->launches++;
			e = cudaGetLastError();
		}
		if (e == cudaSuccess) e = cudaMemcpyAsync(host.data(), hq ? (const void*)d_vec : (const void*)sl.d_feat, vec_bytes * (size_t)m, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		if (e != cudaSuccess) { free_all(tmp); return fail(FDB_ERR_CUDA, std::string("fdb_extract_windows: ") + cudaGetErrorString(e)); }
		for (int k = 0; k < m; ++k) std::memcpy((uint8_t*)out + (size_t)where[off + (size_t)k] * vec_bytes, host.data() + (size_t)k * vec_bytes, vec_bytes);
	}
	free_all(tmp);
	return FDB_OK;
} FDB_API_CATCH

int fdb_detect_single_device(fdb_detector* det, const uint8_t* frames_device, int32_t n_frames, double* distance_device,
		fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) try {
	if (!det || !det->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector not prepared (call fdb_detector_prepare)");
	int s = check_ctx(det->ctx); if (s) return s;
	if (det->wvm || !det->svm) return fail(FDB_ERR_INVALID_ARGUMENT, "fdb_detect_single_device needs a detector created with an SVM only");
	if (n_frames < 0 || (n_frames > 0 && !frames_device)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad frame batch");
	if (!single_dense_usable(det)) { /* any other classifier / feature space: the per-window kernels; distances only to host memory */
		if (distance_device) return fail(FDB_ERR_UNSUPPORTED, "fdb_detect_single_device: distances stay on the device only for the tensor-core SVM (u8 RBF on HistEq64 patches)");
		return detect_single(det, frames_device, true, det->plan.width, n_frames, nullptr, detections_out, det_cap, n_detections);
	}
	return detect_single_dense(det, frames_device, true, det->plan.width, n_frames, distance_device, true, detections_out, det_cap, n_detections);
} FDB_API_CATCH

int fdb_detector_single_dense(fdb_detector* det) { return det && det->prepared && single_dense_usable(det) ? 1 : 0; }

int fdb_detector_single_dense_profile(fdb_detector* det, double* kernel_ms, int32_t* launches) try {
	if (!det || !kernel_ms || !launches) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*kernel_ms = det->sd_kernel_ms; *launches = det->sd_kernel_launches;
	return FDB_OK;
} FDB_API_CATCH

int fdb_detector_last_counts(fdb_detector* det, int64_t counts[5]) try {
	if (!det || !counts) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	std::memcpy(counts, det->counts, sizeof(det->counts));
	return FDB_OK;
} FDB_API_CATCH

} // extern "C"
