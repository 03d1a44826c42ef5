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
	std::vector<char> dead(v.size(), 0);
	for (size_t acc = 0; acc < v.size(); ++acc) {
		if (dead[acc]) continue;
		for (size_t pro = acc + 1; pro < v.size(); ++pro) {
			if (dead[pro]) continue;
			const int wa = v[acc].width, wp = v[pro].width;
			const float d = dist <= 1.0 ? dist * (float)std::max(wa, wp) : dist;
			if (std::abs(v[acc].center_x - v[pro].center_x) < d && std::abs(v[acc].center_y - v[pro].center_y) < d
					&& ((float)std::min(wa, wp) / (float)std::max(wa, wp)) > ratio)
				dead[pro] = 1;
		}
	}
	size_t k = 0;
	for (size_t i = 0; i < v.size(); ++i)
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
