/*
 * hostpost.cpp - the small, inherently sequential tail of the five-stage cascade, on the host.
 *
 * Product code (C++).  After the GPU has classified every window, only a handful of candidates
 * per frame remain; the reference's greedy overlap elimination and block NMS are order-dependent
 * O(n^2) procedures on those few records, so they run here on the exact double-precision values:
 *   OverlapElimination::eliminate            libDetection/src/detection/OverlapElimination.cpp:44-105
 *   nonMaximaSuppression + NMS bookkeeping   libDetection/src/detection/FiveStageSlidingWindowDetector.cpp:143-184,276-311
 *   logistic functions                       ProbabilisticWvmClassifier.cpp:52, ProbabilisticSvmClassifier.cpp:54-58
 * std::sort ties are implementation-defined in the reference; here sorting is stable, i.e. ties
 * keep their incoming (canonical window) order.
 */
#include "fdb_internal.h"

#include <algorithm>
#include <cmath>
#include <map>

namespace fdb {

double wvm_probability(double a, double b, float fout) {
	const double d = fout;
	return 1.0f / (1.0f + std::exp(a + b * d));
}

double svm_probability(double a, double b, double distance) {
	const double fABp = a + b * distance;
	return fABp >= 0 ? std::exp(-fABp) / (1.0 + std::exp(-fABp)) : 1.0 / (1.0 + std::exp(fABp));
}

/* ProbabilisticRvmClassifier.cpp:62 */
double rvm_probability(double a, double b, double distance) {
	return 1.0f / (1.0f + std::exp(a + b * distance));
}

void stable_sort_desc(std::vector<fdb_detection>& v) {
	std::stable_sort(v.begin(), v.end(), [](const fdb_detection& a, const fdb_detection& b) {
		return a.probability > b.probability;
	});
}

void overlap_eliminate(std::vector<fdb_detection>& v, float dist, float ratio_in) {
	if (v.empty()) return;
	const float ratio = (ratio_in > 0.0f && ratio_in <= 1.0f) ? ratio_in : 0.0f;
	stable_sort_desc(v);
	const size_t n = v.size();
	std::vector<char> dead(n, 0);
	auto kills = [&](size_t acc, size_t pro) {
		const int wa = v[acc].width, wp = v[pro].width;
		const float d = dist <= 1.0 ? dist * (float)std::max(wa, wp) : dist;
		return std::abs(v[acc].center_x - v[pro].center_x) < d && std::abs(v[acc].center_y - v[pro].center_y) < d
				&& ((float)std::min(wa, wp) / (float)std::max(wa, wp)) > ratio;
	};
	if (dist > 1.0f && n > 64) {
		/* absolute distance (the cfg's dist 5.0): an accepted patch can only eliminate patches whose centre lies within
		 * `dist` pixels, so each one looks at the 3 x 3 neighbourhood of a grid of dist-sized cells instead of at every
		 * later candidate. Same pairs tested in the same greedy order as the reference's double loop => same survivors. */
		const int cell = std::max(1, (int)std::ceil(dist));
		int min_x = v[0].center_x, min_y = v[0].center_y, max_x = min_x, max_y = min_y;
		for (const fdb_detection& d : v) {
			min_x = std::min(min_x, d.center_x); max_x = std::max(max_x, d.center_x);
			min_y = std::min(min_y, d.center_y); max_y = std::max(max_y, d.center_y);
		}
		const int gw = (max_x - min_x) / cell + 1, gh = (max_y - min_y) / cell + 1;
		if ((int64_t)gw * gh <= (int64_t)1 << 22) {
			std::vector<int> head((size_t)gw * gh, -1), next(n, -1);
			for (size_t i = n; i-- > 0;) { /* lists in ascending candidate order */
				const size_t c = (size_t)((v[i].center_y - min_y) / cell) * gw + (size_t)((v[i].center_x - min_x) / cell);
				next[i] = head[c]; head[c] = (int)i;
			}
			for (size_t acc = 0; acc < n; ++acc) {
				if (dead[acc]) continue;
				const int cx = (v[acc].center_x - min_x) / cell, cy = (v[acc].center_y - min_y) / cell;
				for (int gy = std::max(cy - 1, 0); gy <= std::min(cy + 1, gh - 1); ++gy)
					for (int gx = std::max(cx - 1, 0); gx <= std::min(cx + 1, gw - 1); ++gx)
						for (int pro = head[(size_t)gy * gw + gx]; pro >= 0; pro = next[(size_t)pro])
							if ((size_t)pro > acc && !dead[(size_t)pro] && kills(acc, (size_t)pro)) dead[(size_t)pro] = 1;
			}
			size_t k = 0;
			for (size_t i = 0; i < n; ++i) if (!dead[i]) v[k++] = v[i];
			v.resize(k);
			return;
		}
	}
	for (size_t acc = 0; acc < n; ++acc) {
		if (dead[acc]) continue;
		for (size_t pro = acc + 1; pro < n; ++pro)
			if (!dead[pro] && kills(acc, pro)) dead[pro] = 1;
	}
	size_t k = 0;
	for (size_t i = 0; i < n; ++i)
		if (!dead[i]) v[k++] = v[i];
	v.resize(k);
}

namespace {

struct Peak { int x, y; float value; };

/* Sparse evaluation of nonMaximaSuppression(map, sz, dst, mask) where map is zero except at
 * `pts`: returns the maxima coordinates. use_mask selects points with value > 0.3f only. */
std::vector<Peak> sparse_block_nms(const std::vector<Peak>& pts, int M, int N, int sz, bool use_mask) {
	std::vector<Peak> maxima;
	std::vector<const Peak*> sel;
	for (const Peak& p : pts)
		if (!use_mask || p.value > 0.3f) sel.push_back(&p);
	/* group by block */
	std::map<std::pair<int, int>, std::vector<const Peak*>> blocks;
	for (const Peak* p : sel)
		blocks[{p->y / (sz + 1), p->x / (sz + 1)}].push_back(p);
	for (auto& kv : blocks) {
		const int m = kv.first.first * (sz + 1), n = kv.first.second * (sz + 1);
		/* first occurrence (row-major) of the block maximum; background cells are 0 */
		const Peak* best = nullptr;
		for (const Peak* p : kv.second)
			if (!best || p->value > best->value || (p->value == best->value && (p->y < best->y || (p->y == best->y && p->x < best->x))))
				best = p;
		if (!use_mask && !(best->value > 0.0f)) continue; /* an empty cell wins the block: 0 > vnmax is false */
		const double vcmax = best->value;
		const int cy = best->y, cx = best->x;
		const int in0 = std::max(cy - sz, 0), in1 = std::min(cy + sz + 1, M);
		const int jn0 = std::max(cx - sz, 0), jn1 = std::min(cx + sz + 1, N);
		const int b_y0 = m, b_y1 = std::min(m + sz + 1, in1), b_x0 = n, b_x1 = std::min(n + sz + 1, jn1);
		double vnmax = 0; /* minMaxLoc yields 0 when nothing is selected; background is 0 as well */
		bool any = false;
		for (const Peak* q : sel) {
			if (q->y < in0 || q->y >= in1 || q->x < jn0 || q->x >= jn1) continue;
			if (q->y >= b_y0 && q->y < b_y1 && q->x >= b_x0 && q->x < b_x1) continue;
			if (!any || q->value > vnmax) { vnmax = q->value; any = true; }
		}
		if (!use_mask && any && vnmax < 0) vnmax = 0; /* unmasked: zero background cells take part */
		if (vcmax > vnmax) maxima.push_back(*best);
	}
	std::sort(maxima.begin(), maxima.end(), [](const Peak& a, const Peak& b) { return a.y < b.y || (a.y == b.y && a.x < b.x); });
	return maxima;
}

} // namespace

void five_stage_nms(std::vector<fdb_detection>& v, int width, int height) {
	/* probability map: max over patches sharing a centre (float) */
	std::map<std::pair<int, int>, float> cell;
	for (const fdb_detection& d : v) {
		if (d.center_x < 0 || d.center_y < 0 || d.center_x >= width || d.center_y >= height) continue;
		auto it = cell.find({d.center_y, d.center_x});
		const float cur = it == cell.end() ? 0.0f : it->second;
		if (cur < d.probability) cell[{d.center_y, d.center_x}] = (float)d.probability;
	}
	std::vector<Peak> pts;
	for (auto& kv : cell) pts.push_back({kv.first.second, kv.first.first, kv.second});
	std::vector<Peak> maxima = sparse_block_nms(pts, height, width, 35, true);
	if (maxima.empty()) {
		maxima = sparse_block_nms(pts, height, width, 35, false);
		if (maxima.empty()) return; /* FiveStageSlidingWindowDetector.cpp:292-294: list returned as is */
	}
	stable_sort_desc(v);
	std::vector<fdb_detection> out;
	for (const Peak& p : maxima)
		for (const fdb_detection& d : v)
			if (d.center_x == p.x && d.center_y == p.y) { out.push_back(d); break; }
	v.swap(out);
	stable_sort_desc(v);
}

} // namespace fdb


/* detection::NonMaximumSuppression::eliminateRedundantDetections (libDetection/src/detection/NonMaximumSuppression.cpp:27-112),
 * the IoU suppression of the reference's newer detector family (AggregatedFeaturesDetector.cpp:104-112); host only.
 * Candidates are sorted by ascending score (:35-39; std::sort - equal scores keep their input order here); the best
 * remaining candidate opens a cluster and takes every candidate whose overlap with it exceeds the threshold (:48-58,
 * intersection over union on integer rectangles :60-64); each cluster yields its best member, or the (score-weighted)
 * mean box rounded half away from zero with the best score (:74-109). Output order: clusters by descending best score. */
extern "C" int fdb_non_maximum_suppression(float* scores, int32_t* rects_xywh, int64_t n, double overlap_threshold,
		int32_t maximum_type, int64_t* n_out) try {
	using fdb::fail;
	if (n < 0 || (n > 0 && (!scores || !rects_xywh)) || !n_out) return fail(FDB_ERR_INVALID_ARGUMENT, "bad detection list");
	if (maximum_type < FDB_NMS_MAX_SCORE || maximum_type > FDB_NMS_WEIGHTED_AVERAGE)
		return fail(FDB_ERR_INVALID_ARGUMENT, "NonMaximumSuppression: unsupported maximum type");
	*n_out = n;
	if (overlap_threshold == 1.0 || n == 0) return FDB_OK; /* :28-29 */
	struct Box { float score; int x, y, w, h; };
	std::vector<Box> pending((size_t)n);
	for (int64_t i = 0; i < n; ++i) {
		Box b = {scores[i], rects_xywh[4 * i], rects_xywh[4 * i + 1], rects_xywh[4 * i + 2], rects_xywh[4 * i + 3]};
		pending[(size_t)i] = b;
	}
	std::stable_sort(pending.begin(), pending.end(), [](const Box& a, const Box& b) { return a.score < b.score; });
	auto iou = [](const Box& a, const Box& b) {
		const int x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
		const int iw = std::min(a.x + a.w, b.x + b.w) - x1, ih = std::min(a.y + a.h, b.y + b.h) - y1;
		const double inter = (iw <= 0 || ih <= 0) ? 0.0 : (double)(iw * ih);     /* cv::Rect operator& and area() */
		const double uni = (double)(a.w * a.h) + (double)(b.w * b.h) - inter;
		return inter / uni;
	};
	int64_t out = 0;
	std::vector<Box> members, rest;
	while (!pending.empty()) {
		const Box best = pending.back();
		members.clear(); rest.clear();
		for (const Box& c : pending) (iou(best, c) <= overlap_threshold ? rest : members).push_back(c);
		std::reverse(members.begin(), members.end()); /* descending score: the best one first */
		pending.swap(rest);
		if (members.empty()) return fail(FDB_ERR_RUNTIME, "NonMaximumSuppression: a box does not overlap itself (empty rectangle)");
		Box r = members.front();
		if (maximum_type != FDB_NMS_MAX_SCORE) {
			double ws = 0, xs = 0, ys = 0, wsum = 0, hs = 0;
			for (const Box& m : members) {
				const double wgt = maximum_type == FDB_NMS_AVERAGE ? 1.0 : (double)m.score;
				ws += wgt;
				if (maximum_type == FDB_NMS_AVERAGE) { xs += m.x; ys += m.y; wsum += m.w; hs += m.h; }
				else { xs += wgt * m.x; ys += wgt * m.y; wsum += wgt * m.w; hs += wgt * m.h; }
			}
			const double den = maximum_type == FDB_NMS_AVERAGE ? (double)members.size() : ws;
			r.x = (int)std::round(xs / den); r.y = (int)std::round(ys / den);
			r.w = (int)std::round(wsum / den); r.h = (int)std::round(hs / den);
		}
		scores[out] = r.score;
		rects_xywh[4 * out] = r.x; rects_xywh[4 * out + 1] = r.y; rects_xywh[4 * out + 2] = r.w; rects_xywh[4 * out + 3] = r.h;
		++out;
	}
	*n_out = out;
	return FDB_OK;
} FDB_API_CATCH

/* AggregatedFeaturesDetector::getPositiveWindows for one layer (AggregatedFeaturesDetector.cpp:92-118) with
 * AggregatedFeaturesExtractor::computeBoundsInImagePixels (AggregatedFeaturesExtractor.cpp:121-128) and rescaleWindow:
 * every score-map position above the threshold becomes a box in image pixels. scale_x / scale_y = layer size / image size
 * (ImagePyramid.cpp:178-179,187-188). Appends to scores_out / rects_out (capacity cap); *n_out = number found (may exceed cap:
 * nothing past cap is written). Host only. */
extern "C" int fdb_aggdet_windows(const float* score_map, int32_t valid_rows, int32_t valid_cols, float threshold, int32_t kernel_rows,
		int32_t kernel_cols, int32_t cell_size, double scale_x, double scale_y, float width_scale, float height_scale,
		float* scores_out, int32_t* rects_xywh_out, int64_t cap, int64_t* n_out) try {
	using fdb::fail;
	if (!n_out || valid_rows < 0 || valid_cols < 0 || (valid_rows > 0 && valid_cols > 0 && !score_map) || cap < 0
			|| (cap > 0 && (!scores_out || !rects_xywh_out)))
		return fail(FDB_ERR_INVALID_ARGUMENT, "bad score map");
	if (kernel_rows < 1 || kernel_cols < 1 || cell_size < 1 || !(scale_x > 0) || !(scale_y > 0)) return fail(FDB_ERR_INVALID_ARGUMENT, "bad geometry");
	int64_t n = 0;
	for (int y = 0; y < valid_rows; ++y)
		for (int x = 0; x < valid_cols; ++x) {
			const float score = score_map[(size_t)y * valid_cols + x];
			if (!(score > threshold)) continue;
			if (n < cap) {
				/* bounds in layer cells (x, y, kernel) -> image pixels: round(value * cellSize / scale), half away from zero */
				const int bx = (int)std::round((double)(x * cell_size) / scale_x), by = (int)std::round((double)(y * cell_size) / scale_y);
				const int bw = (int)std::round((double)(kernel_cols * cell_size) / scale_x), bh = (int)std::round((double)(kernel_rows * cell_size) / scale_y);
				const int cx = bx + bw / 2, cy = by + bh / 2;                       /* Patch::computeCenter */
				const int rw = (int)(width_scale * bw), rh = (int)(height_scale * bh); /* Size(float, float) -> Size_<int> truncates */
				scores_out[n] = score;
				rects_xywh_out[4 * n] = cx - rw / 2; rects_xywh_out[4 * n + 1] = cy - rh / 2; /* Patch::computeBounds */
				rects_xywh_out[4 * n + 2] = rw; rects_xywh_out[4 * n + 3] = rh;
			}
			++n;
		}
	*n_out = n;
	return FDB_OK;
} FDB_API_CATCH

