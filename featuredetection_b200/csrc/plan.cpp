/*
 * plan.cpp - host-side geometry of the image pyramid and of the window scan.
 *
 * Product code (C++): decides which pyramid images exist for a W x H frame, where they live in
 * the per-frame device arena, and which windows DirectPyramidFeatureExtractor::extract would
 * visit.  Semantics follow the reference:
 *   ImagePyramid::ImagePyramid(double,double,double)   libImageProcessing/src/imageprocessing/ImagePyramid.cpp:79-92
 *   ImagePyramid::createLayers(const Mat&)             ImagePyramid.cpp:170-198
 *   ImagePyramidLayer::getScaled / getOriginal         include/imageprocessing/ImagePyramidLayer.hpp:65-67,98-100
 *   DirectPyramidFeatureExtractor::extract             DirectPyramidFeatureExtractor.cpp:75-123
 * Only chains that end in at least one kept layer are materialised (the reference also builds
 * the others but nothing reads them).
 */
#include "fdb_internal.h"

#include <algorithm>
#include <cmath>

namespace fdb {

int cv_round(double v) {
	return (int)std::nearbyint(v); /* cvRound: round half to even */
}

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

int build_plan(const fdb_detector_desc& d, int width, int height, Plan* out) {
	if (width < 1 || height < 1)
		return fail(FDB_ERR_INVALID_ARGUMENT, "frame size must be positive");
	const double inc = d.incremental_scale_factor;
	if (inc <= 0 || inc >= 1)
		return fail(FDB_ERR_INVALID_ARGUMENT, "the incremental scale factor must be greater than zero and smaller than one");
	if (d.min_scale_factor <= 0)
		return fail(FDB_ERR_INVALID_ARGUMENT, "the minimum scale factor must be greater than zero");
	if (d.max_scale_factor > 1)
		return fail(FDB_ERR_INVALID_ARGUMENT, "the maximum scale factor must not exceed one");
	Plan p;
	p.width = width; p.height = height;
	p.octave_layer_count = (int)static_cast<size_t>(std::round(std::log(0.5) / std::log(inc)));
	if (p.octave_layer_count < 1)
		return fail(FDB_ERR_INVALID_ARGUMENT, "the number of layers per octave must be greater than zero");
	p.incremental_scale_factor = std::pow(0.5, 1. / p.octave_layer_count);
	p.min_scale = d.min_scale_factor; p.max_scale = d.max_scale_factor;

	int64_t offset = 0;
	for (int i = 0; i < p.octave_layer_count; ++i) {
		/* walk the chain once to see whether its last image is kept */
		std::vector<PyrImage> chain;
		double sf = std::pow(p.incremental_scale_factor, i);
		PyrImage top{};
		top.width = cv_round(width * sf); top.height = cv_round(height * sf);
		if (top.width < 1 || top.height < 1) continue; /* cv::resize would assert */
		top.kind = (top.width == width && top.height == height) ? IMG_FRAME : IMG_RESIZE;
		top.src = -1; top.octave = i; top.down = 0; top.scale = sf; top.level = 0;
		top.kept = sf <= p.max_scale && sf >= p.min_scale;
		top.layer_index = i;
		chain.push_back(top);
		int pw = top.width, ph = top.height;
		sf *= 0.5;
		for (int j = 1; sf >= p.min_scale && pw > 1; ++j, sf *= 0.5) {
			PyrImage im{};
			im.kind = IMG_PYRDOWN; im.octave = i; im.down = j; im.scale = sf; im.level = j;
			im.width = (pw + 1) / 2; im.height = (ph + 1) / 2;
			im.kept = sf <= p.max_scale;
			im.layer_index = i + j * p.octave_layer_count;
			chain.push_back(im);
			pw = im.width; ph = im.height;
		}
		if (!chain.back().kept) continue; /* nothing of this chain is ever read */
		int prev = -1;
		for (PyrImage& im : chain) {
			if (im.kind == IMG_PYRDOWN) im.src = prev;
			if (im.kind != IMG_FRAME) {
				im.pitch = (int)align_up(im.width, 16); /* 16-byte row pitch: TMA tensor maps need it, word stores like it */
				im.offset = offset;
				offset = align_up(offset + (int64_t)im.pitch * im.height, 128);
			} else {
				im.pitch = width;
				im.offset = -1;
			}
			p.max_down = std::max(p.max_down, im.down);
			p.images.push_back(im);
			prev = (int)p.images.size() - 1;
		}
	}
	p.arena_bytes = std::max<int64_t>(align_up(offset, 128), 128);
	for (size_t k = 0; k < p.images.size(); ++k) {
		const PyrImage& im = p.images[k];
		if (!im.kept) continue;
		PlanLayer L{};
		L.image = (int)k; L.index = im.layer_index; L.scale = im.scale;
		L.width = im.width; L.height = im.height;
		L.orig_patch_w = cv_round(d.patch_width / im.scale);
		L.orig_patch_h = cv_round(d.patch_height / im.scale);
		p.layers.push_back(L);
	}
	std::sort(p.layers.begin(), p.layers.end(), [](const PlanLayer& a, const PlanLayer& b) { return a.index < b.index; });
	if ((int)p.layers.size() > FDB_MAX_LAYERS)
		return fail(FDB_ERR_UNSUPPORTED, "more than 64 pyramid layers");
	p.windows = enumerate_windows(&p, d.patch_width, d.patch_height, d.step_x > 0 ? d.step_x : 1,
			d.step_y > 0 ? d.step_y : 1, 0, 0, 0, 0);
	*out = p;
	return FDB_OK;
}

int64_t enumerate_windows(Plan* plan, int patch_w, int patch_h, int step_x, int step_y,
		int roi_x, int roi_y, int roi_w, int roi_h) {
	/* DirectPyramidFeatureExtractor.cpp:84-92 */
	if (roi_x == 0 && roi_y == 0 && roi_w == 0 && roi_h == 0) {
		roi_w = plan->width; roi_h = plan->height;
	} else {
		int x = std::max(0, roi_x), y = std::max(0, roi_y);
		roi_w = std::min(plan->width, roi_w + x) - x;
		roi_h = std::min(plan->height, roi_h + y) - y;
		roi_x = x; roi_y = y;
	}
	int64_t total = 0;
	for (PlanLayer& L : plan->layers) {
		/* :110-114: strict '<' loop bounds against the SCALED roi end, theoretical scale */
		L.begin_x = cv_round(roi_x * L.scale); L.begin_y = cv_round(roi_y * L.scale);
		const int end_x = cv_round((roi_x + roi_w) * L.scale), end_y = cv_round((roi_y + roi_h) * L.scale);
		const int last_x = end_x - patch_w - 1, last_y = end_y - patch_h - 1;
		L.windows_x = last_x < L.begin_x ? 0 : (last_x - L.begin_x) / step_x + 1;
		L.windows_y = last_y < L.begin_y ? 0 : (last_y - L.begin_y) / step_y + 1;
		/* memory safety: a window never leaves the layer image (cannot trigger for cvRound'ed
		 * ends, which exceed the ceil-halved layer size by at most one) */
		while (L.windows_x > 0 && L.begin_x + (L.windows_x - 1) * step_x + patch_w > L.width) --L.windows_x;
		while (L.windows_y > 0 && L.begin_y + (L.windows_y - 1) * step_y + patch_h > L.height) --L.windows_y;
		if (L.windows_x == 0 || L.windows_y == 0) L.windows_x = L.windows_y = 0;
		L.first_window = total;
		total += (int64_t)L.windows_x * L.windows_y;
	}
	return total;
}

} // namespace fdb
