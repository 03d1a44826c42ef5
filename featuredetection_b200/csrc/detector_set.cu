/*
 * detector_set.cu - all detectors of an application over one frame batch.
 *
 * ffpDetectApp builds one detector per landmark cfg - each with its OWN ImagePyramid (ffpDetectApp.cpp:407,435) - and runs
 * every one of them on every frame (ffpDetectApp.cpp:548-596). The 15 cfgs use 4 distinct pyramid parameter sets and 5
 * window sizes; 12 of them scan the same four layers. A detector set gives the same results as running its members one
 * after the other (tests/test_detector_set.py) while
 *   - every distinct pyramid image is built once per frame (the union of the members' pyramid plans in one arena);
 *   - windows of the same layer and size are equalised once for a pack of models (wvm_group.cu), one launch per window size;
 *   - the survivors of all members finish their cascades in one launch of the deep kernel;
 *   - host post-processing (overlap elimination, SVM stage, grid NMS) runs per member exactly as in detector.cu.
 * Members that cannot take the group kernels (other window sizes or steps, models without early exits that overflow the
 * deep queue) run through their own pipeline inside the same call.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "detector_internal.h"

#if defined(__linux__)
#include <sched.h>
#endif

using namespace fdb;

namespace {

#define SET_SIDE_STREAMS 2
struct SetSlot {
	cudaStream_t st = nullptr;
	cudaStream_t side[SET_SIDE_STREAMS] = {nullptr, nullptr}; /* the window launches of a chunk are independent: run side by side, one */
	cudaEvent_t ev_fork = nullptr, ev_join[SET_SIDE_STREAMS] = {nullptr, nullptr}; /* kernel's CTAs fill the SMs another one drains */
	cudaEvent_t ev_s1_end = nullptr; /* the chunk's deep kernel is done: the next chunk's stage 1 may start */
	uint8_t* d_frames = nullptr;
	uint8_t* d_arena = nullptr;
	CUtensorMap* d_tmaps = nullptr; /* one per union image */
	int* d_cursors = nullptr;       /* work counters of the window launches */
	int n = 0, base = 0;
	const uint8_t* frames_dev = nullptr;
	bool busy = false;
};

struct HostTimer { /* adds the host wall clock of its scope to *dst (milliseconds) */
	double* dst; std::chrono::steady_clock::time_point t0;
	explicit HostTimer(double* d) : dst(d), t0(std::chrono::steady_clock::now()) {}
	~HostTimer() { *dst += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

/* host threads for the members' post-processing (overlap elimination, NMS): the pure-CPU parts of different members are
 * independent. run(n, f) calls f(0..n-1) on the pool and the calling thread and returns when all are done. */
class HostPool {
public:
	explicit HostPool(int threads) {
		for (int i = 1; i < threads; ++i) workers_.emplace_back([this] { loop(); });
	}
	~HostPool() {
		{ std::lock_guard<std::mutex> g(m_); stop_ = true; }
		cv_.notify_all();
		for (std::thread& t : workers_) t.join();
	}
	int threads() const { return (int)workers_.size() + 1; }
	void run(int n, const std::function<void(int)>& f) {
		if (n <= 0) return;
		if (workers_.empty() || n == 1) { for (int i = 0; i < n; ++i) f(i); return; }
		{
			std::lock_guard<std::mutex> g(m_);
			job_ = &f; n_ = n; next_ = 0; pending_ = n; ++generation_;
		}
		cv_.notify_all();
		work();
		std::unique_lock<std::mutex> g(m_);
		done_.wait(g, [this] { return pending_ == 0; });
		job_ = nullptr;
	}
private:
	void work() {
		for (;;) {
			int i;
			const std::function<void(int)>* f;
			{
				std::lock_guard<std::mutex> g(m_);
				if (!job_ || next_ >= n_) return;
				i = next_++; f = job_;
			}
			(*f)(i);
			std::lock_guard<std::mutex> g(m_);
			if (--pending_ == 0) done_.notify_all();
		}
	}
	void loop() {
		uint64_t seen = 0;
		for (;;) {
			{
				std::unique_lock<std::mutex> g(m_);
				cv_.wait(g, [&] { return stop_ || generation_ != seen; });
				if (stop_) return;
				seen = generation_;
			}
			work();
		}
	}
	std::vector<std::thread> workers_;
	std::mutex m_;
	std::condition_variable cv_, done_;
	const std::function<void(int)>* job_ = nullptr;
	int n_ = 0, next_ = 0, pending_ = 0;
	uint64_t generation_ = 0;
	bool stop_ = false;
};

int host_thread_count() {
	if (const char* e = std::getenv("FDB_HOST_THREADS")) { const int v = std::atoi(e); if (v > 0) return std::min(v, 64); }
	int n = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
	cpu_set_t set;
	if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set); /* a rank pinned to its share of the cores uses only those */
#endif
	return std::max(1, std::min(n, 16));
}

struct SetLaunch {  /* one wvm_group_kernel launch: all strips of one window size and pack width */
	int pw = 0, ph = 0, pack = 1;
	bool tc_ok = false;   /* every model of the launch has the tcgen05 operand */
	GroupItem* d_items = nullptr; int n_items = 0;
};

} // namespace

struct fdb_detector_set {
	fdb_ctx* ctx = nullptr;
	std::vector<fdb_detector*> dets;
	bool prepared = false;
	int W = 0, H = 0, max_batch = 0, chunk = 0, n_slots = 0;
	std::vector<PyrImage> images;       /* union of the members' pyramid images, dependency order */
	std::vector<std::vector<int>> umap; /* member d, plan image k -> union image */
	int64_t arena_bytes = 0;
	int max_down = 0;
	PyramidJobs jobs;
	GroupImage* d_images = nullptr;
	std::vector<DevLayer*> d_layers;    /* member d's layer table against the union arena */
	std::vector<SetLaunch> launches;
	std::vector<char> fast;             /* member d takes the group kernels */
	bool use_tma = false;
	SetSlot slots[PIPE_SLOTS];
	cudaEvent_t ev_begin = nullptr;
	std::vector<cudaEvent_t> trace_ev;   /* FDB_SET_TRACE: begin / end of stage 1 per chunk, timing enabled */
	bool trace = false;
	std::vector<void*> owned, owned_host;
	int64_t windows = 0;                /* per frame, all members */
	HostPool* pool = nullptr;           /* host threads of the members' post-processing */
	double host_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* last call, host wall clock: enqueue, phase A (overlap elimination, SVM launch), phase B, whole call, waiting for stage 1 */
};

namespace {

void set_release(fdb_detector_set* s) {
	for (SetSlot& sl : s->slots) {
		if (sl.st) { cudaStreamSynchronize(sl.st); cudaStreamDestroy(sl.st); }
		for (int k = 0; k < SET_SIDE_STREAMS; ++k) {
			if (sl.side[k]) { cudaStreamSynchronize(sl.side[k]); cudaStreamDestroy(sl.side[k]); sl.side[k] = nullptr; }
			if (sl.ev_join[k]) { cudaEventDestroy(sl.ev_join[k]); sl.ev_join[k] = nullptr; }
		}
		if (sl.ev_fork) { cudaEventDestroy(sl.ev_fork); sl.ev_fork = nullptr; }
		if (sl.ev_s1_end) { cudaEventDestroy(sl.ev_s1_end); sl.ev_s1_end = nullptr; }
		sl = SetSlot();
	}
	if (s->ev_begin) { cudaEventDestroy(s->ev_begin); s->ev_begin = nullptr; }
	delete s->pool; s->pool = nullptr;
	free_all(s->owned, &s->owned_host);
	s->images.clear(); s->umap.clear(); s->d_layers.clear(); s->launches.clear(); s->fast.clear();
	s->jobs = PyramidJobs();
	s->prepared = false;
}

/* the members' plans merged: an image is identified by how it is made (the frame, a resize target, the pyrDown of an image) */
void build_union(fdb_detector_set* s) {
	std::map<std::tuple<int, int, int, int>, int> index; /* (kind, source union image, width, height) -> union image */
	int64_t offset = 0;
	auto align_up = [](int64_t v, int64_t a) { return (v + a - 1) / a * a; };
	for (fdb_detector* det : s->dets) {
		std::vector<int> m(det->plan.images.size(), -1);
		for (size_t k = 0; k < det->plan.images.size(); ++k) {
			const PyrImage& im = det->plan.images[k];
			const int src = im.kind == IMG_PYRDOWN ? m[(size_t)im.src] : -1;
			const auto key = std::make_tuple(im.kind, src, im.width, im.height);
			auto it = index.find(key);
			if (it == index.end()) {
				PyrImage u = im;
				u.src = src; u.kept = false; u.layer_index = -1;
				if (u.kind != IMG_FRAME) {
					u.offset = offset;
					offset = align_up(offset + (int64_t)u.pitch * u.height, 128);
				}
				s->max_down = std::max(s->max_down, u.down);
				s->images.push_back(u);
				it = index.emplace(key, (int)s->images.size() - 1).first;
			}
			m[k] = it->second;
		}
		s->umap.push_back(m);
	}
	s->arena_bytes = std::max<int64_t>(align_up(offset, 128), 128);
}

GroupModel member_model(const fdb_detector_set* s, int d, const Slot& msl, fdb_window_score* dense) {
	const fdb_detector* det = s->dets[(size_t)d];
	GroupModel gm{};
	gm.m = det->wvm->dev;
	gm.m.step_x = gm.m.step_y = 1;
	gm.dense = dense;
	gm.windows_per_frame = (int)det->plan.windows;
	gm.cand_cap = det->cand_cap;
	gm.cand = msl.d_cand; gm.cand_count = msl.d_counters; gm.q = msl.deep;
	return gm;
}

/* work items of the window kernels for the members on the group path (s->fast): windows of the same image, size and grid are
 * equalised once for all members that scan them. Called at prepare and again when a member leaves the group path. */
int set_build_launches(fdb_detector_set* s) {
	const int nd = (int)s->dets.size();
	int r = FDB_OK;
	s->launches.clear();
	typedef std::tuple<int, int, int, int, int, int, int> GridKey; /* patch w, h, union image, begin x, y, windows x, y */
	std::map<GridKey, std::vector<std::pair<int, int>>> grids;     /* -> (member, layer) */
	for (int d = 0; d < nd; ++d) {
		if (!s->fast[(size_t)d]) continue;
		const fdb_detector* det = s->dets[(size_t)d];
		for (size_t li = 0; li < det->plan.layers.size(); ++li) {
			const PlanLayer& p = det->plan.layers[li];
			if (p.windows_x <= 0 || p.windows_y <= 0) continue;
			grids[GridKey(det->desc.patch_width, det->desc.patch_height, s->umap[(size_t)d][(size_t)p.image], p.begin_x, p.begin_y,
					p.windows_x, p.windows_y)].push_back(std::make_pair(d, (int)li));
		}
	}
	std::map<std::tuple<int, int, int, int>, std::vector<GroupItem>> by_launch; /* (patch w, h, pack, tcgen05 operand present) -> items */
	for (const auto& kv : grids) {
		const int pw = std::get<0>(kv.first), ph = std::get<1>(kv.first), image = std::get<2>(kv.first);
		const PlanLayer& p = s->dets[(size_t)kv.second[0].first]->plan.layers[(size_t)kv.second[0].second];
		for (int tc = 1; tc >= 0; --tc) { /* models with the tcgen05 operand pack by four, the others by two */
			std::vector<int> models, firsts;
			for (const auto& ml : kv.second) {
				if ((s->dets[(size_t)ml.first]->wvm->dev.btc != nullptr) != (tc == 1)) continue;
				models.push_back(ml.first);
				firsts.push_back((int)s->dets[(size_t)ml.first]->plan.layers[(size_t)ml.second].first_window);
			}
			if (models.empty()) continue;
			std::vector<GroupItem> items;
			append_strip_items(p, image, ph, (int)models.size(), models.data(), firsts.data(), group_max_pack(tc == 1), &items);
			/* kernel instances exist for packs of 1, 2 and GRP_MAX_PACK */
			for (const GroupItem& it : items) by_launch[std::make_tuple(pw, ph, it.nm > 2 ? GRP_MAX_PACK : it.nm, tc)].push_back(it);
		}
	}
	for (auto& kv : by_launch) {
		SetLaunch L;
		L.pw = std::get<0>(kv.first); L.ph = std::get<1>(kv.first); L.pack = std::get<2>(kv.first); L.tc_ok = std::get<3>(kv.first) != 0;
		/* pack-major order: consecutive units share the models' fragment tables (L1) */
		std::stable_sort(kv.second.begin(), kv.second.end(), [](const GroupItem& a, const GroupItem& b) { return a.model[0] < b.model[0]; });
		L.n_items = (int)kv.second.size();
		r = upload(kv.second.data(), kv.second.size(), &L.d_items, s->owned); if (r) return r;
		s->launches.push_back(L);
	}
	if (s->launches.size() > 16) return fail(FDB_ERR_UNSUPPORTED, "more than 16 window-size classes in a detector set");
	/* longest first: the launches run on a few streams side by side, the short ones fill the SMs the long ones drain */
	std::stable_sort(s->launches.begin(), s->launches.end(), [](const SetLaunch& a, const SetLaunch& b) {
		return (int64_t)a.n_items * a.ph * a.pack > (int64_t)b.n_items * b.ph * b.pack; });
	return FDB_OK;
}

/* pyramid + stage 1 of every fast member for the chunk of slot si; ev (optional): marks after resize, pyrDown, window kernels */
int set_enqueue(fdb_detector_set* s, int si, fdb_window_score* const* dense_dev, cudaEvent_t* marks) {
	SetSlot& ss = s->slots[si];
	fdb_ctx* c = s->ctx;
	cudaStream_t st = ss.st;
	const int nd = (int)s->dets.size();
	if (!s->launches.empty()) CUDA_TRY(cudaMemsetAsync(ss.d_cursors, 0, sizeof(int) * s->launches.size(), st));
	for (int d = 0; d < nd; ++d) if (s->fast[(size_t)d]) CUDA_TRY(cudaMemsetAsync(s->dets[(size_t)d]->slots[si].d_counters, 0, FDB_NCOUNTERS * sizeof(int), st));
	if (marks) CUDA_TRY(cudaEventRecord(marks[0], st));
	if (s->trace && !marks) {
		cudaEvent_t e0, e1;
		CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
		s->trace_ev.push_back(e0); s->trace_ev.push_back(e1);
		CUDA_TRY(cudaEventRecord(e0, st));
	}
	{ const int r = enqueue_pyramid(c, st, s->jobs, ss.frames_dev, s->W, s->H, ss.n, ss.d_arena, s->arena_bytes, marks ? marks[1] : nullptr); if (r) return r; }
	if (marks) CUDA_TRY(cudaEventRecord(marks[2], st));
	GroupArgs ga{};
	ga.n_frames = ss.n; ga.images = s->d_images; ga.tmaps = s->use_tma ? ss.d_tmaps : nullptr;
	ga.frames = ss.frames_dev; ga.W = s->W; ga.H = s->H; ga.arena = ss.d_arena; ga.arena_stride = s->arena_bytes;
	DeepArgs da{};
	da.images = s->d_images; da.frames = ss.frames_dev; da.W = s->W; da.H = s->H; da.arena = ss.d_arena; da.arena_stride = s->arena_bytes;
	std::vector<int> deep_of((size_t)nd, -1);
	for (int d = 0; d < nd; ++d) {
		if (!s->fast[(size_t)d]) continue;
		fdb_detector* det = s->dets[(size_t)d];
		fdb_window_score* dense = dense_dev && dense_dev[d] ? dense_dev[d] + (int64_t)ss.base * det->plan.windows : nullptr;
		ga.models[d] = member_model(s, d, det->slots[si], dense);
		da.models[da.n_models++] = ga.models[d];
	}
	/* the window launches only share atomic counters: launch k goes to stream k mod (1 + SET_SIDE_STREAMS), joined before the deep kernel */
	const bool fork = !marks && s->launches.size() > 1;
	if (fork) {
		CUDA_TRY(cudaEventRecord(ss.ev_fork, st));
		for (int k = 0; k < SET_SIDE_STREAMS; ++k) CUDA_TRY(cudaStreamWaitEvent(ss.side[k], ss.ev_fork, 0));
	}
	for (size_t k = 0; k < s->launches.size(); ++k) {
		const SetLaunch& L = s->launches[k];
		ga.items = L.d_items; ga.n_items = L.n_items; ga.cursor = ss.d_cursors + k;
		const int lane = fork ? (int)(k % (1 + SET_SIDE_STREAMS)) : 0;
		launch_wvm_group(lane == 0 ? st : ss.side[lane - 1], L.pw, L.ph, L.pack, ga, L.tc_ok);
		c->launches++;
	}
	if (fork)
		for (int k = 0; k < SET_SIDE_STREAMS; ++k) {
			CUDA_TRY(cudaEventRecord(ss.ev_join[k], ss.side[k]));
			CUDA_TRY(cudaStreamWaitEvent(st, ss.ev_join[k], 0));
		}
	if (marks) CUDA_TRY(cudaEventRecord(marks[3], st));
	if (da.n_models) { launch_wvm_deep_group(st, da); c->launches++; }
	if (marks) CUDA_TRY(cudaEventRecord(marks[4], st));
	if (s->trace && !marks) CUDA_TRY(cudaEventRecord(s->trace_ev.back(), st));
	CUDA_TRY(cudaGetLastError());
	return FDB_OK;
}

int set_pipeline(fdb_detector_set* s, const uint8_t* frames, bool frames_on_device, int64_t pitch, int32_t n_frames, int32_t stage,
		fdb_window_score* const* dense_dev, std::vector<std::vector<fdb_detection>>& results) {
	const int W = s->W, H = s->H, nd = (int)s->dets.size();
	for (int d = 0; d < nd; ++d) {
		fdb_detector* det = s->dets[(size_t)d];
		std::fill(det->counts, det->counts + 5, 0);
		det->counts[0] = det->plan.windows * n_frames;
	}
	CUDA_TRY(cudaEventRecord(s->ev_begin, s->ctx->stream));
	for (int i = 0; i < s->n_slots; ++i) CUDA_TRY(cudaStreamWaitEvent(s->slots[i].st, s->ev_begin, 0));
	/* chunk schedule: full chunks, the last one split into 1/2 + 1/4 + 1/8 + 1/8 - what follows the last chunk's stage 1 (its host
	 * phases and the SVM kernels) overlaps nothing, so it should be short */
	std::vector<std::pair<int, int>> chunks; /* (first frame, frames) */
	for (int base = 0; base < n_frames; base += s->chunk) chunks.push_back(std::make_pair(base, std::min(s->chunk, n_frames - base)));
	if (chunks.size() >= 2 && chunks.back().second >= 16) {
		std::pair<int, int> rest = chunks.back();
		chunks.pop_back();
		while (rest.second >= 16) { /* pieces of at least 8 frames: two groups of four for the tcgen05 kernel */
			const int half = rest.second / 2;
			chunks.push_back(std::make_pair(rest.first, half));
			rest = std::make_pair(rest.first + half, rest.second - half);
		}
		chunks.push_back(rest);
	}
	const int n_chunks = (int)chunks.size();
	int enq = 0, a_done = 0, retired = 0;
	auto do_a = [&]() -> int {
		const int si = a_done % s->n_slots;
		SetSlot& ss = s->slots[si];
		{ /* stage 1 of the chunk: every member's records are complete after the last member's event */
			HostTimer wait(&s->host_ms[4]);
			for (int d = nd - 1; d >= 0; --d) if (s->fast[(size_t)d]) { CUDA_TRY(cudaEventSynchronize(s->dets[(size_t)d]->slots[si].ev_stage1)); break; }
		}
		HostTimer timer(&s->host_ms[1]);
		std::vector<int> members;
		{
			HostTimer t5(&s->host_ms[5]);
			for (int d = 0; d < nd; ++d) {
				if (!s->fast[(size_t)d]) continue;
				fdb_detector* det = s->dets[(size_t)d];
				Slot& msl = det->slots[si];
				msl.n = ss.n; msl.base = ss.base; msl.frames_dev = ss.frames_dev;
				msl.arena = ss.d_arena; msl.arena_stride = s->arena_bytes;
				const int r = phase_a_fetch(det, msl, ss.st, s->d_layers[(size_t)d], 1);
				if (r) return r;
				members.push_back(d);
			}
			for (int d : members) { const int r = phase_a_fetch_wait(s->dets[(size_t)d]->slots[si], ss.st); if (r) return r; }
		}
		std::vector<int> status(members.size(), FDB_OK);
		{
			HostTimer t6(&s->host_ms[6]);
			s->pool->run((int)members.size(), [&](int k) {
				fdb_detector* det = s->dets[(size_t)members[(size_t)k]];
				status[(size_t)k] = phase_a_host(det, det->slots[si], det->plan, stage);
			});
		}
		HostTimer t7(&s->host_ms[7]);
		for (size_t k = 0; k < members.size(); ++k) {
			if (status[k]) return fail(status[k], "SVM work list overflow");
			fdb_detector* det = s->dets[(size_t)members[k]];
			const int r = phase_a_launch(det, det->slots[si], ss.st, det->plan, s->d_layers[(size_t)members[k]], stage);
			if (r) return r;
		}
		++a_done;
		return FDB_OK;
	};
	auto do_b = [&]() -> int {
		HostTimer timer(&s->host_ms[2]);
		const int si = retired % s->n_slots;
		std::vector<int> members;
		for (int d = 0; d < nd; ++d) if (s->fast[(size_t)d]) members.push_back(d);
		/* the slot stream is in order: the last member's SVM results are the last to arrive */
		if (!members.empty()) CUDA_TRY(cudaEventSynchronize(s->dets[(size_t)members.back()]->slots[si].ev_svm));
		s->pool->run((int)members.size(), [&](int k) {
			fdb_detector* det = s->dets[(size_t)members[(size_t)k]];
			phase_b_host(det, det->slots[si], det->plan, stage, false, results[(size_t)members[(size_t)k]]);
		});
		s->slots[si].busy = false;
		++retired;
		return FDB_OK;
	};
	int r = FDB_OK;
	auto enqueue_next = [&]() -> int {
		HostTimer timer(&s->host_ms[0]);
		const int si = enq % s->n_slots;
		SetSlot& ss = s->slots[si];
		ss.base = chunks[(size_t)enq].first;
		ss.n = chunks[(size_t)enq].second;
		ss.busy = true;
		if (frames_on_device) ss.frames_dev = frames + (int64_t)ss.base * W * H;
		else {
			if (pitch == W)
				CUDA_TRY(cudaMemcpyAsync(ss.d_frames, frames + (int64_t)ss.base * W * H, (size_t)W * H * ss.n, cudaMemcpyHostToDevice, ss.st));
			else
				CUDA_TRY(cudaMemcpy2DAsync(ss.d_frames, (size_t)W, frames + (int64_t)ss.base * pitch * H, (size_t)pitch, (size_t)W,
						(size_t)H * ss.n, cudaMemcpyHostToDevice, ss.st));
			ss.frames_dev = ss.d_frames;
		}
		/* stage 1 of consecutive chunks runs in order (the copy above does not wait): chunks are enqueued as far ahead as there are
		 * slots, and two chunks sharing the SMs would both finish late - the host phases of the first would start late */
		if (enq > 0) CUDA_TRY(cudaStreamWaitEvent(ss.st, s->slots[(enq - 1) % s->n_slots].ev_s1_end, 0));
		const int r2 = set_enqueue(s, si, dense_dev, nullptr); if (r2) return r2;
		CUDA_TRY(cudaEventRecord(ss.ev_s1_end, ss.st));
		for (int d = 0; d < nd; ++d) {
			if (!s->fast[(size_t)d]) continue;
			Slot& msl = s->dets[(size_t)d]->slots[si];
			CUDA_TRY(cudaMemcpyAsync(msl.h_counters, msl.d_counters, FDB_NCOUNTERS * sizeof(int) + (size_t)s->dets[(size_t)d]->opt_cand * sizeof(Candidate), cudaMemcpyDeviceToHost, ss.st));
			CUDA_TRY(cudaEventRecord(msl.ev_stage1, ss.st));
		}
		++enq;
		return FDB_OK;
	};
	for (;;) {
		/* the GPU first: every free slot gets the next chunk before the host turns to its own phases */
		while (enq < n_chunks && !s->slots[enq % s->n_slots].busy) { r = enqueue_next(); if (r) return r; }
		if (a_done < enq) {
			r = do_a(); if (r) return r;
			if (retired < a_done - 1) { r = do_b(); if (r) return r; } /* phase B of the chunk before: its SVM kernels have had a chunk's time */
		} else if (retired < a_done) {
			r = do_b(); if (r) return r;
		} else break;
	}
	for (int i = 0; i < s->n_slots; ++i) CUDA_TRY(cudaStreamSynchronize(s->slots[i].st));
	return FDB_OK;
}

int set_detect(fdb_detector_set* s, const uint8_t* frames, bool frames_on_device, int64_t pitch, int32_t n_frames, int32_t stage,
		fdb_window_score* const* dense_dev, fdb_detection* dets_out, int64_t det_cap, int64_t* n_dets) {
	if (!s || !s->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector set not prepared (call fdb_detector_set_prepare)");
	int r = check_ctx(s->ctx); if (r) return r;
	if (n_frames < 0 || (n_frames > 0 && !frames)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad frame batch");
	if (stage < FDB_STAGE_WVM || stage > FDB_STAGE_NMS) return fail(FDB_ERR_INVALID_ARGUMENT, "bad stage");
	if (!frames_on_device && pitch < s->W) return fail(FDB_ERR_INVALID_ARGUMENT, "pitch smaller than the frame width");
	const int nd = (int)s->dets.size();
	std::vector<std::vector<fdb_detection>> results((size_t)nd);
	std::fill(s->host_ms, s->host_ms + 8, 0.0);
	s->trace = std::getenv("FDB_SET_TRACE") != nullptr;
	HostTimer whole(&s->host_ms[3]);
	for (int attempt = 0; attempt < nd + 1; ++attempt) {
		for (auto& v : results) v.clear();
		r = set_pipeline(s, frames, frames_on_device, pitch, n_frames, stage, dense_dev, results);
		for (int i = 0; i < s->n_slots; ++i) { cudaStreamSynchronize(s->slots[i].st); s->slots[i].busy = false; }
		for (fdb_detector* det : s->dets) for (int i = 0; i < det->n_slots; ++i) det->slots[i].busy = false;
		if (r != STATUS_REDO) break;
		/* a member's deep queue overflowed (phase_a took it off the fast path): it runs alone from now on */
		for (int d = 0; d < nd; ++d) if (s->fast[(size_t)d] && !s->dets[(size_t)d]->use_strips) s->fast[(size_t)d] = 0;
		r = set_build_launches(s); if (r) return r; /* the packs must not name a member that left */
	}
	if (r) return r;
	if (s->trace && s->trace_ev.size() >= 2) { /* stage-1 intervals of the chunks on the device clock: gaps = the GPU waiting for the host */
		cudaEventSynchronize(s->trace_ev.back());
		std::fprintf(stderr, "fdb_detector_set: stage 1 per chunk [begin, end] ms:");
		for (size_t k = 0; k + 1 < s->trace_ev.size(); k += 2) {
			float b = 0.f, e = 0.f;
			cudaEventElapsedTime(&b, s->trace_ev[0], s->trace_ev[k]);
			cudaEventElapsedTime(&e, s->trace_ev[0], s->trace_ev[k + 1]);
			std::fprintf(stderr, " [%.1f, %.1f]", b, e);
		}
		std::fprintf(stderr, "\n");
	}
	for (cudaEvent_t e : s->trace_ev) cudaEventDestroy(e);
	s->trace_ev.clear();
	if (std::getenv("FDB_SET_TRACE"))
		std::fprintf(stderr, "fdb_detector_set: enqueue %.1f wait %.1f phaseA %.1f (fetch %.1f host %.1f on %d threads, launch %.1f) phaseB %.1f ms\n", s->host_ms[0],
				s->host_ms[4], s->host_ms[1], s->host_ms[5], s->host_ms[6], s->pool ? s->pool->threads() : 1, s->host_ms[7], s->host_ms[2]);
	/* members outside the group kernels: their own pipeline (own pyramid), same results */
	for (int d = 0; d < nd; ++d) {
		if (s->fast[(size_t)d]) continue;
		fdb_detector* det = s->dets[(size_t)d];
		const int64_t cap = std::max<int64_t>((int64_t)det->desc.max_positives_per_frame * n_frames, 1);
		std::vector<fdb_detection> tmp((size_t)cap);
		int64_t got = 0;
		fdb_window_score* dense = dense_dev ? dense_dev[d] : nullptr;
		r = detect_impl(det, frames, frames_on_device, pitch, n_frames, stage, dense, true, tmp.data(), cap, &got);
		if (r) return r;
		results[(size_t)d].assign(tmp.begin(), tmp.begin() + got);
	}
	std::vector<fdb_detection> all;
	for (int d = 0; d < nd; ++d)
		for (fdb_detection& x : results[(size_t)d]) { x.reserved = d; all.push_back(x); }
	return copy_out(all, dets_out, det_cap, n_dets);
}

} // namespace

extern "C" {

int fdb_detector_set_create(fdb_ctx* ctx, fdb_detector* const* detectors, int32_t n, fdb_detector_set** out) try {
	int r = check_ctx(ctx); if (r) return r;
	if (!detectors || !out || n < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "a detector set needs at least one detector");
	*out = nullptr;
	if (n > GRP_MAX_MODELS) return fail(FDB_ERR_UNSUPPORTED, "more than 16 detectors in a set");
	for (int i = 0; i < n; ++i) {
		if (!detectors[i] || detectors[i]->ctx != ctx) return fail(FDB_ERR_INVALID_ARGUMENT, "detector of another context (or null) in the set");
		if (!detectors[i]->wvm) return fail(FDB_ERR_INVALID_ARGUMENT, "a set holds cascade detectors (WVM first stage); `single` detectors run on their own");
		for (int j = 0; j < i; ++j) if (detectors[j] == detectors[i]) return fail(FDB_ERR_INVALID_ARGUMENT, "the same detector twice in a set");
	}
	fdb_detector_set* s = new fdb_detector_set;
	s->ctx = ctx;
	s->dets.assign(detectors, detectors + n);
	*out = s;
	return FDB_OK;
} FDB_API_CATCH

void fdb_detector_set_destroy(fdb_detector_set* s) {
	if (!s) return;
	cudaSetDevice(s->ctx->device);
	cudaStreamSynchronize(s->ctx->stream);
	set_release(s);
	delete s;
}

int fdb_detector_set_prepare(fdb_detector_set* s, int32_t width, int32_t height, int32_t max_batch) try {
	if (!s) return fail(FDB_ERR_INVALID_ARGUMENT, "null detector set");
	int r = check_ctx(s->ctx); if (r) return r;
	if (max_batch < 1) return fail(FDB_ERR_INVALID_ARGUMENT, "max_batch must be positive");
	CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
	set_release(s);
	const int nd = (int)s->dets.size();
	for (fdb_detector* det : s->dets) { r = fdb_detector_prepare(det, width, height, max_batch); if (r) return r; }
	s->W = width; s->H = height; s->max_batch = max_batch;
	s->chunk = s->dets[0]->chunk; s->n_slots = s->dets[0]->n_slots;
	s->windows = 0;
	for (fdb_detector* det : s->dets) s->windows += det->plan.windows;
	build_union(s);
	r = build_pyramid_jobs(s->images, s->max_down, width, height, &s->jobs, s->owned); if (r) return r;
	/* the set runs the fast members; a member prepared off the fast path (window size, step, model shape) runs alone */
	s->fast.assign((size_t)nd, 0);
	for (int d = 0; d < nd; ++d) s->fast[(size_t)d] = s->dets[(size_t)d]->use_strips ? 1 : 0;
	CUDA_TRY(cudaEventCreateWithFlags(&s->ev_begin, cudaEventDisableTiming));
	s->pool = new HostPool(std::min(host_thread_count(), nd));
	for (int i = 0; i < s->n_slots; ++i) {
		SetSlot& sl = s->slots[i];
		CUDA_TRY(cudaStreamCreateWithFlags(&sl.st, cudaStreamNonBlocking));
		CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming));
		CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_s1_end, cudaEventDisableTiming));
		for (int k = 0; k < SET_SIDE_STREAMS; ++k) {
			CUDA_TRY(cudaStreamCreateWithFlags(&sl.side[k], cudaStreamNonBlocking));
			CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_join[k], cudaEventDisableTiming));
		}
		r = dev_alloc(&sl.d_frames, (size_t)s->chunk * width * height, s->owned); if (r) return r;
		r = dev_alloc(&sl.d_arena, (size_t)s->chunk * (size_t)s->arena_bytes, s->owned); if (r) return r;
		r = dev_alloc(&sl.d_cursors, 64, s->owned); if (r) return r;
	}
	s->use_tma = false;
	{
		const char* env = std::getenv("FDB_NO_TMA");
		if (!(env && env[0] == '1')) {
			std::vector<int> which(s->images.size());
			for (size_t k = 0; k < which.size(); ++k) which[k] = (int)k;
			bool ok = true;
			for (int i = 0; i < s->n_slots && ok; ++i) {
				std::vector<CUtensorMap> maps;
				ok = encode_tile_maps(s->images, which, s->slots[i].d_arena, s->arena_bytes, s->chunk, &maps);
				if (ok) { r = upload(maps.data(), maps.size(), &s->slots[i].d_tmaps, s->owned); if (r) return r; }
			}
			s->use_tma = ok;
		}
	}
	std::vector<GroupImage> gim(s->images.size());
	for (size_t k = 0; k < gim.size(); ++k) {
		const PyrImage& im = s->images[k];
		gim[k].offset = im.kind == IMG_FRAME ? -1 : im.offset; gim[k].width = im.width; gim[k].height = im.height; gim[k].pitch = im.pitch;
		gim[k].tma_ok = s->use_tma && im.kind != IMG_FRAME ? 1 : 0;
	}
	r = upload(gim.data(), gim.size(), &s->d_images, s->owned); if (r) return r;
	/* members' layer tables against the union arena (SVM stage, feature layers) */
	s->d_layers.assign((size_t)nd, nullptr);
	for (int d = 0; d < nd; ++d) {
		const Plan& plan = s->dets[(size_t)d]->plan;
		std::vector<DevLayer> L(std::max<size_t>(plan.layers.size(), 1));
		for (size_t i = 0; i < plan.layers.size(); ++i) {
			const PlanLayer& p = plan.layers[i];
			const PyrImage& u = s->images[(size_t)s->umap[(size_t)d][(size_t)p.image]];
			L[i].offset = u.kind == IMG_FRAME ? -1 : u.offset;
			L[i].width = p.width; L[i].height = p.height; L[i].pitch = u.pitch;
			L[i].begin_x = p.begin_x; L[i].begin_y = p.begin_y; L[i].windows_x = p.windows_x; L[i].windows_y = p.windows_y;
			L[i].first_window = (int)p.first_window;
			L[i].tma_ok = 0;
		}
		r = upload(L.data(), L.size(), &s->d_layers[(size_t)d], s->owned); if (r) return r;
	}
	r = set_build_launches(s); if (r) return r;
	s->prepared = true;
	return FDB_OK;
} FDB_API_CATCH

int64_t fdb_detector_set_windows_per_frame(fdb_detector_set* s) { return s && s->prepared ? s->windows : -1; }

int fdb_detector_set_info(fdb_detector_set* s, int32_t* n_images, int64_t* pyramid_bytes, int32_t* n_window_launches, int32_t* n_fast) try {
	if (!s || !s->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector set not prepared");
	if (n_images) { int k = 0; for (const PyrImage& im : s->images) k += im.kind != IMG_FRAME; *n_images = k; }
	if (pyramid_bytes) { int64_t b = 0; for (const PyrImage& im : s->images) if (im.kind != IMG_FRAME) b += (int64_t)im.pitch * im.height; *pyramid_bytes = b; }
	if (n_window_launches) *n_window_launches = (int32_t)s->launches.size();
	if (n_fast) { int k = 0; for (char f : s->fast) k += f; *n_fast = k; }
	return FDB_OK;
} FDB_API_CATCH

int fdb_detector_set_last_host_ms(fdb_detector_set* s, double ms_out[8]) try {
	if (!s || !ms_out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	std::memcpy(ms_out, s->host_ms, sizeof(s->host_ms));
	return FDB_OK;
} FDB_API_CATCH

int fdb_detector_set_detect_batch(fdb_detector_set* s, const uint8_t* frames_host, int64_t pitch, int32_t n_frames, int32_t stage,
		fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) try {
	return set_detect(s, frames_host, false, pitch, n_frames, stage, nullptr, detections_out, det_cap, n_detections);
} FDB_API_CATCH

int fdb_detector_set_detect_batch_device(fdb_detector_set* s, const uint8_t* frames_device, int32_t n_frames, int32_t stage,
		fdb_window_score* const* dense_out_device, fdb_detection* detections_out, int64_t det_cap, int64_t* n_detections) try {
	return set_detect(s, frames_device, true, 0, n_frames, stage, dense_out_device, detections_out, det_cap, n_detections);
} FDB_API_CATCH

int fdb_detector_set_profile_device(fdb_detector_set* s, const uint8_t* frames_device, int32_t n_frames, double ms_out[6]) try {
	if (!s || !s->prepared) return fail(FDB_ERR_INVALID_ARGUMENT, "detector set not prepared");
	int r = check_ctx(s->ctx); if (r) return r;
	if (n_frames < 0 || n_frames > s->max_batch || !ms_out || !frames_device) return fail(FDB_ERR_INVALID_ARGUMENT, "bad arguments");
	cudaEvent_t ev[5];
	for (int k = 0; k < 5; ++k) CUDA_TRY(cudaEventCreate(&ev[k]));
	for (int k = 0; k < 6; ++k) ms_out[k] = 0;
	SetSlot& ss = s->slots[0];
	for (int base = 0; base < n_frames && r == FDB_OK; base += s->chunk) {
		ss.base = base; ss.n = std::min(s->chunk, n_frames - base);
		ss.frames_dev = frames_device + (int64_t)base * s->W * s->H;
		r = set_enqueue(s, 0, nullptr, ev);
		if (r) break;
		if (cudaEventSynchronize(ev[4]) != cudaSuccess) { r = fail(FDB_ERR_CUDA, "profile: event synchronize failed"); break; }
		float a = 0, b = 0, w = 0, d = 0, t = 0;
		cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]); cudaEventElapsedTime(&w, ev[2], ev[3]);
		cudaEventElapsedTime(&d, ev[3], ev[4]); cudaEventElapsedTime(&t, ev[0], ev[4]);
		ms_out[0] += a; ms_out[1] += b; ms_out[2] += w; ms_out[3] += d; ms_out[4] += t; ms_out[5] += (double)s->launches.size();
	}
	for (int k = 0; k < 5; ++k) cudaEventDestroy(ev[k]);
	return r;
} FDB_API_CATCH

} // extern "C"
