/*
 * ref_driver.cpp - thin C entry points over the reference's OWN, UNMODIFIED sources, compiled
 * from where they lie under /root/reference into oracle/_ref/libfdref.so (see oracle/Makefile).
 *
 * TEST INFRASTRUCTURE ONLY.  It pins oracle/fd_oracle.c (the restatement that travels to the GPU
 * box) and can serve as the "reference" CPU baseline.  No reference source is copied into this
 * repository: this file only *calls* the reference classes
 *   imageprocessing::HistEq64Filter            (libImageProcessing/.../HistEq64Filter.cpp)
 *   classification::WvmClassifier + IImg       (libClassification/.../WvmClassifier.cpp, IImg.cpp)
 *   classification::ProbabilisticWvmClassifier (.../ProbabilisticWvmClassifier.cpp)
 *   classification::SvmClassifier + RbfKernel  (.../SvmClassifier.cpp, RbfKernel.hpp)
 *   classification::ProbabilisticSvmClassifier (.../ProbabilisticSvmClassifier.cpp)
 *   detection::OverlapElimination              (libDetection/.../OverlapElimination.cpp)
 * Synthetic model state enters WvmClassifier through its protected members (a subclass), the way
 * WvmClassifier::loadFromMatlab (WvmClassifier.cpp:348-770) fills them.
 */
/* RvmClassifier keeps its model private and fills it only from MATLAB files (RvmClassifier.cpp:131-319, needs libmat): the
 * test driver reaches the members through the preprocessor; the class itself is compiled unmodified from the reference tree */
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>
#include "opencv2/core/core.hpp"
#include "boost/property_tree/ptree.hpp"
#define private public
#define protected public
#include "classification/RvmClassifier.hpp"
#undef protected
#undef private
#include "classification/ProbabilisticRvmClassifier.hpp"
#include "classification/WvmClassifier.hpp"
#include "classification/ProbabilisticWvmClassifier.hpp"
#include "classification/SvmClassifier.hpp"
#include "classification/ProbabilisticSvmClassifier.hpp"
#include "classification/RbfKernel.hpp"
#include "classification/PolynomialKernel.hpp"
#include "classification/HistogramIntersectionKernel.hpp"
#include "classification/LinearKernel.hpp"
#include "imageprocessing/HistEq64Filter.hpp"
#include "imageprocessing/Patch.hpp"
#include "detection/ClassifiedPatch.hpp"
#include "detection/OverlapElimination.hpp"
#include "detection/NonMaximumSuppression.hpp"
#include "imageprocessing/filtering/FhogFilter.hpp"
#include "imageprocessing/filtering/GradientHistogramFilter.hpp"

#include "fdb200.h"
#include "fd_oracle.h"

#include <chrono>

#include <algorithm>
#include <memory>
#include <vector>

using cv::Mat;
using std::make_shared;
using std::shared_ptr;
using std::vector;

namespace {

/* Fills the protected state exactly as the Matlab loader would (allocation with new[] because
 * ~WvmClassifier delete[]s everything, WvmClassifier.cpp:51-74). */
class SynWvm : public classification::WvmClassifier {
public:
	explicit SynWvm(const fdb_wvm_desc& d) {
		filter_size_x = d.filter_size_x;
		filter_size_y = d.filter_size_y;
		numLinFilters = d.num_lin_filters;
		numFiltersPerLevel = d.num_filters_per_level;
		numLevels = d.num_levels;
		basisParam = d.basis_param;
		const int n = numLinFilters;
		linFilters = new float*[n];
		hkWeights = new float*[n];
		lin_thresholds = new float[n];
		area = new Area*[n];
		app_rsv_convol = new double[n];
		size_t wofs = 0;
		int slot = 0;
		size_t rofs = 0;
		for (int f = 0; f < n; ++f) {
			linFilters[f] = new float[filter_size_x * filter_size_y](); /* never read by the evaluator */
			hkWeights[f] = new float[n]();
			for (int p = 0; p <= f; ++p) hkWeights[f][p] = d.hk_weights[wofs + p];
			wofs += (size_t)f + 1;
			lin_thresholds[f] = d.lin_thresholds[f];
			app_rsv_convol[f] = d.app_rsv_convol[f];
			const int cntval = d.area_cntval[f];
			vector<int> cr(cntval);
			for (int v = 0; v < cntval; ++v) cr[v] = v == 0 ? 0 : d.area_cntrec[slot + v];
			area[f] = new Area(cntval, cr.data());
			for (int v = 0; v < cntval; ++v) {
				area[f]->val[v] = d.area_val[slot + v];
				for (int r = 0; r < cr[v]; ++r) {
					const fdb_rect4& q = d.area_rec[rofs++];
					area[f]->rec[v][r].x1 = q.x1; area[f]->rec[v][r].y1 = q.y1;
					area[f]->rec[v][r].x2 = q.x2; area[f]->rec[v][r].y2 = q.y2;
				}
			}
			slot += cntval;
		}
		filter_output = new float[n];      /* WvmClassifier.cpp:759-760 */
		u_kernel_eval = new float[n];
		for (int f = 0; f < n; ++f)
			hierarchicalThresholdsFromFile.push_back(d.hierarchical_thresholds[f]);
		setLimitReliabilityFilter(d.limit_reliability_filter); /* :753 */
		setNumUsedFilters(d.num_used_filters);                 /* :762 */
	}
};

struct RefWvm {
	shared_ptr<SynWvm> wvm;
	shared_ptr<classification::ProbabilisticWvmClassifier> pwvm;
	int w, h;
};

struct RefSvm {
	shared_ptr<classification::SvmClassifier> svm;
	shared_ptr<classification::ProbabilisticSvmClassifier> psvm;
	int dim, type;
};

} // namespace

extern "C" {

/* HistEq64Filter::applyTo on a ROI view (non-continuous when pitch != w) */
void ref_hq64(const uint8_t* src, int pitch, int w, int h, uint8_t* dst) {
	static imageprocessing::HistEq64Filter filter;
	Mat roi(h, w, CV_8U, (void*)src, (size_t)pitch);
	Mat out(h, w, CV_8U, dst);
	filter.applyTo(roi, out);
	if (out.data != dst) std::memcpy(dst, out.data, (size_t)w * h);
}

void* ref_wvm_create(const fdb_wvm_desc* d) {
	RefWvm* r = new RefWvm;
	r->wvm = make_shared<SynWvm>(*d);
	r->pwvm = make_shared<classification::ProbabilisticWvmClassifier>(r->wvm, d->logistic_a, d->logistic_b);
	r->w = d->filter_size_x; r->h = d->filter_size_y;
	return r;
}
void ref_wvm_free(void* p) { delete (RefWvm*)p; }

/* WvmClassifier::computeHyperplaneDistance + ProbabilisticWvmClassifier::getProbability */
void ref_wvm_eval(void* p, const uint8_t* patch, int* level, float* fout, double* probability, int* positive) {
	RefWvm* r = (RefWvm*)p;
	Mat m(r->h, r->w, CV_8U, (void*)patch);
	std::pair<int, double> ld = r->wvm->computeHyperplaneDistance(m);
	*level = ld.first;
	*fout = (float)ld.second; /* exact: the value is a widened float */
	std::pair<bool, double> pr = r->pwvm->getProbability(ld);
	if (probability) *probability = pr.second;
	if (positive) *positive = pr.first ? 1 : 0;
}

static shared_ptr<classification::Kernel> ref_make_kernel(int kind, double gamma, double alpha, double constant, int degree) {
	if (kind == FDB_KERNEL_POLYNOMIAL) return make_shared<classification::PolynomialKernel>(alpha, constant, degree);
	if (kind == FDB_KERNEL_HIK) return make_shared<classification::HistogramIntersectionKernel>();
	if (kind == FDB_KERNEL_LINEAR) return make_shared<classification::LinearKernel>();
	return make_shared<classification::RbfKernel>(gamma);
}

struct RefRvm {
	shared_ptr<classification::RvmClassifier> rvm;
	shared_ptr<classification::ProbabilisticRvmClassifier> prvm;
	int dim, type;
};

void* ref_rvm_create(const fdb_rvm_desc* d) {
	RefRvm* r = new RefRvm;
	r->rvm = make_shared<classification::RvmClassifier>(ref_make_kernel(d->kernel, d->gamma, d->poly_alpha, d->poly_constant, d->poly_degree));
	for (int i = 0; i < d->num_filters; ++i) {
		Mat sv(1, d->dim, d->sv_type == FDB_SV_U8 ? CV_8U : CV_32F);
		const size_t bytes = (d->sv_type == FDB_SV_U8 ? 1 : 4) * (size_t)d->dim;
		std::memcpy(sv.data, (const uint8_t*)d->support_vectors + (size_t)i * bytes, bytes);
		r->rvm->supportVectors.push_back(sv);
		r->rvm->coefficients.push_back(vector<float>(d->coefficients + (size_t)i * (i + 1) / 2, d->coefficients + (size_t)i * (i + 1) / 2 + i + 1));
		r->rvm->hierarchicalThresholds.push_back(d->hierarchical_thresholds[i]);
	}
	r->rvm->bias = d->bias;
	r->rvm->setNumFiltersToUse((unsigned int)(d->num_filters_to_use < 0 ? 0 : d->num_filters_to_use));
	r->prvm = make_shared<classification::ProbabilisticRvmClassifier>(r->rvm, d->logistic_a, d->logistic_b);
	r->dim = d->dim; r->type = d->sv_type;
	return r;
}
void ref_rvm_free(void* p) { delete (RefRvm*)p; }

/* RvmClassifier::computeHyperplaneDistance + ProbabilisticRvmClassifier::getProbability(pair) */
int ref_rvm_eval(void* p, const void* x, double* distance, double* probability, int* positive) {
	RefRvm* r = (RefRvm*)p;
	Mat m(1, r->dim, r->type == FDB_SV_U8 ? CV_8U : CV_32F, (void*)x);
	std::pair<int, double> ld = r->rvm->computeHyperplaneDistance(m);
	*distance = ld.second;
	std::pair<bool, double> pr = r->prvm->getProbability(ld);
	if (probability) *probability = pr.second;
	if (positive) *positive = pr.first ? 1 : 0;
	return ld.first;
}

/* drawing helper referenced by FhogAggregationFilter::visualizeUnsignedHistograms only (GradientHistogramFilter.cpp needs
 * OpenCV's drawing API and is not compiled): never called by the oracle */
cv::Mat imageprocessing::filtering::GradientHistogramFilter::visualizeUnsignedHistograms(const cv::Mat&, int, int, int) {
	throw std::runtime_error("visualizeUnsignedHistograms is not part of the oracle");
}

/* imageprocessing::filtering::FhogFilter::applyTo (the reference's own FhogFilter.cpp / FhogAggregationFilter.cpp) */
int64_t ref_fhog(const uint8_t* image, int cols, int rows, int channels, int cell, int unsigned_bins, int interpolate_bins,
		int interpolate_cells, float alpha, float* out) {
	Mat img(rows, cols, CV_MAKETYPE(CV_8U, channels));
	std::memcpy(img.data, image, (size_t)rows * cols * channels);
	imageprocessing::filtering::FhogFilter filter(cell, unsigned_bins, interpolate_bins != 0, interpolate_cells != 0, alpha);
	Mat desc;
	filter.applyTo(img, desc);
	const size_t n = (size_t)desc.rows * desc.cols * desc.channels();
	std::memcpy(out, desc.data, sizeof(float) * n);
	return (int64_t)n;
}

/* detection::NonMaximumSuppression (the reference's own NonMaximumSuppression.cpp) on n scored boxes, in place */
int64_t ref_non_maximum_suppression(float* scores, int32_t* rects, int64_t n, double overlap_threshold, int maximum_type) {
	vector<detection::Detection> c;
	for (int64_t i = 0; i < n; ++i) c.push_back(detection::Detection{scores[i], cv::Rect(rects[4 * i], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3])});
	detection::NonMaximumSuppression nms(overlap_threshold, (detection::NonMaximumSuppression::MaximumType)maximum_type);
	vector<detection::Detection> r = nms.eliminateRedundantDetections(c);
	for (size_t i = 0; i < r.size(); ++i) {
		scores[i] = r[i].score;
		rects[4 * i] = r[i].bounds.x; rects[4 * i + 1] = r[i].bounds.y; rects[4 * i + 2] = r[i].bounds.width; rects[4 * i + 3] = r[i].bounds.height;
	}
	return (int64_t)r.size();
}

void* ref_svm_create(const fdb_svm_desc* d) {
	RefSvm* r = new RefSvm;
	r->svm = make_shared<classification::SvmClassifier>(ref_make_kernel(d->kernel, d->gamma, d->poly_alpha, d->poly_constant, d->poly_degree));
	vector<Mat> svs;
	for (int i = 0; i < d->num_sv; ++i) {
		if (d->sv_type == FDB_SV_U8) {
			Mat sv(1, d->dim, CV_8U);
			std::memcpy(sv.data, (const uint8_t*)d->support_vectors + (size_t)i * d->dim, (size_t)d->dim);
			svs.push_back(sv);
		} else {
			Mat sv(1, d->dim, CV_32F);
			std::memcpy(sv.data, (const float*)d->support_vectors + (size_t)i * d->dim, sizeof(float) * (size_t)d->dim);
			svs.push_back(sv);
		}
	}
	vector<float> coef(d->coefficients, d->coefficients + d->num_sv);
	r->svm->setSvmParameters(svs, coef, d->bias);
	r->svm->setThreshold(d->threshold);
	r->psvm = make_shared<classification::ProbabilisticSvmClassifier>(r->svm, d->logistic_a, d->logistic_b);
	r->dim = d->dim; r->type = d->sv_type;
	return r;
}
void ref_svm_free(void* p) { delete (RefSvm*)p; }

/* SvmClassifier::computeHyperplaneDistance + ProbabilisticSvmClassifier::getProbability */
void ref_svm_eval(void* p, const void* x, double* distance, double* probability, int* positive) {
	RefSvm* r = (RefSvm*)p;
	Mat m(1, r->dim, r->type == FDB_SV_U8 ? CV_8U : CV_32F, (void*)x);
	double d = r->svm->computeHyperplaneDistance(m);
	*distance = d;
	std::pair<bool, double> pr = r->psvm->getProbability(d);
	if (probability) *probability = pr.second;
	if (positive) *positive = pr.first ? 1 : 0;
}

/* ProbabilisticSvmClassifier::store (-> SvmClassifier::store) into a text file: the reference's own writer */
int ref_svm_store(void* p, const char* path) {
	RefSvm* r = (RefSvm*)p;
	std::ofstream file(path);
	if (!file) return -1;
	r->psvm->store(file);
	return 0;
}

/* OverlapElimination::eliminate on n candidates {center_x, center_y, width, probability};
 * writes the surviving candidates' input indices (in output order) to keep_out, returns count.
 * NOTE: std::sort is not stable; inputs with equal probabilities may come back in a different
 * order than the oracle's stable sort. */
int ref_overlap_eliminate(float dist, float ratio, int n, const int* cx, const int* cy, const int* width,
		const double* probability, int* keep_out) {
	detection::OverlapElimination oe(dist, ratio);
	vector<shared_ptr<detection::ClassifiedPatch>> in;
	for (int i = 0; i < n; ++i) {
		Mat data(1, 1, CV_32S);
		data.at<int>(0, 0) = i;
		auto patch = make_shared<imageprocessing::Patch>(cx[i], cy[i], width[i], width[i], data);
		in.push_back(make_shared<detection::ClassifiedPatch>(patch, true, probability[i]));
	}
	vector<shared_ptr<detection::ClassifiedPatch>> out = oe.eliminate(in);
	for (size_t i = 0; i < out.size(); ++i)
		keep_out[i] = out[i]->getPatch()->getData().at<int>(0, 0);
	return (int)out.size();
}

/* The reference arm of bench.py: one frame through the reference's OWN classes in the order of
 * SlidingWindowDetector::detect() (SlidingWindowDetector.cpp:87-98) and
 * FiveStageSlidingWindowDetector::detect (FiveStageSlidingWindowDetector.cpp:187-322):
 *   extract: Mat(layer, bounds) -> HistEq64Filter::applyTo -> make_shared<Patch>   (per window heap objects,
 *            DirectPyramidFeatureExtractor.cpp:113-118)
 *   classify: ProbabilisticWvmClassifier::getProbability per patch, keep positives
 *   OverlapElimination::eliminate, ProbabilisticSvmClassifier::classify on the survivors
 * The OpenCV-owned pyramid (cv::resize / cv::pyrDown) and the cv::minMaxLoc based grid NMS cannot be
 * compiled without OpenCV; those two steps come from the restatement in fd_oracle.c.
 * timing_out (NULL or 5 doubles, seconds): pyramid, extract+hq64, wvm, oe, svm+nms. */
int64_t ref_detect_frame_ex(const fdb_detector_desc* desc, void* wvm_h, void* svm_h, const fdo_features* feat, const uint8_t* frame,
		int width, int height, int stage, fdb_window_score* dense_out, int64_t* windows_out, int64_t* det_windows_out,
		int64_t det_cap, double* timing_out);

static int64_t ref_detect_on_pyramid(const fdb_detector_desc* desc, void* wvm_h, void* svm_h, const fdo_features* feat, fdo_pyramid* pyr,
		int width, int height, int stage, fdb_window_score* dense_out, int64_t* windows_out, int64_t* det_windows_out,
		int64_t det_cap, double* timing_out);

int64_t ref_detect_frame(const fdb_detector_desc* desc, void* wvm_h, void* svm_h, const uint8_t* frame, int width, int height,
		int stage, fdb_window_score* dense_out, int64_t* windows_out, int64_t* det_windows_out, int64_t det_cap, double* timing_out) {
	return ref_detect_frame_ex(desc, wvm_h, svm_h, nullptr, frame, width, height, stage, dense_out, windows_out, det_windows_out,
			det_cap, timing_out);
}

/* feat != NULL: the SVM stage classifies the window's vector in that feature space (the second extractor of a
 * two-feature cascade; its layer filters and patch filter come from fd_features.c, which is pinned bit for bit
 * against the reference's own filter classes - see ref_features.cpp) with the reference's own SvmClassifier. */
int64_t ref_detect_frame_ex(const fdb_detector_desc* desc, void* wvm_h, void* svm_h, const fdo_features* feat, const uint8_t* frame,
		int width, int height, int stage, fdb_window_score* dense_out, int64_t* windows_out, int64_t* det_windows_out,
		int64_t det_cap, double* timing_out) {
	typedef std::chrono::steady_clock clk;
	auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
	auto t0 = clk::now();
	fdo_pyramid* pyr = fdo_pyramid_build(frame, width, height, width, desc->incremental_scale_factor,
			desc->min_scale_factor, desc->max_scale_factor);
	if (!pyr) return -2;
	auto t1 = clk::now();
	const int64_t n = ref_detect_on_pyramid(desc, wvm_h, svm_h, feat, pyr, width, height, stage, dense_out, windows_out, det_windows_out,
			det_cap, timing_out);
	if (timing_out) timing_out[0] = secs(t0, t1);
	fdo_pyramid_free(pyr);
	return n;
}

/* the same on a pyramid the caller built (bench.py's reference arm builds it with cv2's resize / pyrDown - the SIMD code the real
 * reference links, bit-identical to fdo_pyramid_build by tests/test_oracle_golden.py - in the loop order of
 * ImagePyramid::createLayers, ImagePyramid.cpp:170-198). layers: n_layers kept layers sorted by index; not freed here. */
int64_t ref_detect_layers_ex(const fdb_detector_desc* desc, void* wvm_h, void* svm_h, const fdo_features* feat, const fdo_layer* layers,
		int n_layers, int width, int height, int stage, fdb_window_score* dense_out, int64_t* windows_out, int64_t* det_windows_out,
		int64_t det_cap, double* timing_out) {
	fdo_pyramid pyr;
	std::memset(&pyr, 0, sizeof(pyr));
	pyr.octave_layer_count = (int)(size_t)std::round(std::log(0.5) / std::log(desc->incremental_scale_factor));
	pyr.incremental_scale_factor = std::pow(0.5, 1. / pyr.octave_layer_count);
	pyr.min_scale_factor = desc->min_scale_factor; pyr.max_scale_factor = desc->max_scale_factor;
	pyr.image_width = width; pyr.image_height = height;
	pyr.n_layers = n_layers; pyr.layers = const_cast<fdo_layer*>(layers);
	const int64_t n = ref_detect_on_pyramid(desc, wvm_h, svm_h, feat, &pyr, width, height, stage, dense_out, windows_out, det_windows_out,
			det_cap, timing_out);
	if (timing_out) timing_out[0] = 0.0;
	return n;
}

static int64_t ref_detect_on_pyramid(const fdb_detector_desc* desc, void* wvm_h, void* svm_h, const fdo_features* feat, fdo_pyramid* pyr,
		int width, int height, int stage, fdb_window_score* dense_out, int64_t* windows_out, int64_t* det_windows_out,
		int64_t det_cap, double* timing_out) {
	typedef std::chrono::steady_clock clk;
	auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
	RefWvm* rw = (RefWvm*)wvm_h;
	RefSvm* rs = (RefSvm*)svm_h;
	static imageprocessing::HistEq64Filter hq64;
	auto t1 = clk::now();
	const int pw = desc->patch_width, ph = desc->patch_height;
	const int sx = desc->step_x > 0 ? desc->step_x : 1, sy = desc->step_y > 0 ? desc->step_y : 1;
	vector<fdb_layer_info> infos((size_t)std::max(pyr->n_layers, 1));
	const int64_t total = fdo_enumerate(pyr, pw, ph, sx, sy, 0, 0, 0, 0, infos.data(), pyr->n_layers);
	if (windows_out) *windows_out = total;
	vector<shared_ptr<imageprocessing::Patch>> patches;
	vector<int64_t> window_of;
	patches.reserve((size_t)total);
	for (int li = 0; li < pyr->n_layers; ++li) {
		const fdo_layer& L = pyr->layers[li];
		const fdb_layer_info& I = infos[(size_t)li];
		Mat image(L.height, L.width, CV_8U, L.data);
		for (int iy = 0; iy < I.windows_y; ++iy)
			for (int ix = 0; ix < I.windows_x; ++ix) {
				cv::Rect bounds(ix * sx, iy * sy, pw, ph);
				const int ox = cvRound(bounds.x / L.scale) + I.orig_patch_width / 2;
				const int oy = cvRound(bounds.y / L.scale) + I.orig_patch_height / 2;
				Mat data = hq64.applyTo(Mat(image, bounds));
				patches.push_back(make_shared<imageprocessing::Patch>(ox, oy, I.orig_patch_width, I.orig_patch_height, data));
			}
	}
	auto t2 = clk::now();
	vector<shared_ptr<detection::ClassifiedPatch>> classified;
	for (size_t i = 0; i < patches.size(); ++i) {
		std::pair<int, double> ld = rw->wvm->computeHyperplaneDistance(patches[i]->getData());
		std::pair<bool, double> res = rw->pwvm->getProbability(ld);
		if (dense_out) { dense_out[i].fout = (float)ld.second; dense_out[i].level = ld.first; }
		if (res.first) { classified.push_back(make_shared<detection::ClassifiedPatch>(patches[i], res)); window_of.push_back((int64_t)i); }
	}
	auto t3 = clk::now();
	/* remember the window index of each patch through the stages */
	std::vector<std::pair<const imageprocessing::Patch*, int64_t>> ids;
	for (size_t i = 0; i < classified.size(); ++i) ids.push_back({classified[i]->getPatch().get(), window_of[i]});
	if (stage >= FDB_STAGE_OE) {
		detection::OverlapElimination oe(desc->oe_dist, desc->oe_ratio);
		classified = oe.eliminate(classified);
	}
	auto t4 = clk::now();
	vector<shared_ptr<detection::ClassifiedPatch>> positives = classified;
	if (stage >= FDB_STAGE_SVM && rs) {
		positives.clear();
		vector<vector<uint8_t>> filtered; /* filtered pyramid layers, built on first use */
		for (const auto& p : classified) {
			Mat data = p->getPatch()->getData();
			if (feat) {
				int64_t win = -1;
				for (const auto& id : ids) if (id.first == p->getPatch().get()) win = id.second;
				int li = 0;
				while (li + 1 < pyr->n_layers && win >= infos[(size_t)li + 1].first_window) ++li;
				const int64_t local = win - infos[(size_t)li].first_window;
				const int wy = (int)(local / infos[(size_t)li].windows_x), wx = (int)(local % infos[(size_t)li].windows_x);
				const fdo_layer& L = pyr->layers[li];
				const int chn = fdo_features_layer_channels(feat);
				if (chn && filtered.empty()) filtered.resize((size_t)pyr->n_layers);
				if (chn && filtered[(size_t)li].empty()) {
					filtered[(size_t)li].resize((size_t)L.width * L.height * chn);
					fdo_features_filter_layer(feat, L.data, L.width, L.height, filtered[(size_t)li].data());
				}
				const bool is_float = fdo_features_is_float(feat) != 0;
				data = Mat(1, fdo_features_dim(feat), is_float ? CV_32F : CV_8U);
				fdo_features_patch(feat, chn ? filtered[(size_t)li].data() : L.data, L.width, wx * sx, wy * sy, data.data);
			}
			bool ok = rs->psvm->classify(data);
			if (ok) positives.push_back(make_shared<detection::ClassifiedPatch>(p->getPatch(), ok));
		}
	}
	int64_t n = 0;
	vector<fdb_detection> dets;
	for (const auto& p : positives) {
		fdb_detection d;
		std::memset(&d, 0, sizeof(d));
		d.center_x = p->getPatch()->getX(); d.center_y = p->getPatch()->getY();
		d.width = p->getPatch()->getWidth(); d.height = p->getPatch()->getHeight();
		d.probability = p->getProbability();
		for (const auto& id : ids) if (id.first == p->getPatch().get()) d.window = id.second;
		dets.push_back(d);
	}
	n = (int64_t)dets.size();
	if (stage >= FDB_STAGE_NMS && rs && n) n = fdo_five_stage_nms(dets.data(), n, width, height);
	auto t5 = clk::now();
	if (n > det_cap) n = -1;
	for (int64_t i = 0; i < n; ++i) det_windows_out[i] = dets[(size_t)i].window;
	if (timing_out) { timing_out[0] = 0.0; timing_out[1] = secs(t1, t2); timing_out[2] = secs(t2, t3); timing_out[3] = secs(t3, t4); timing_out[4] = secs(t4, t5); }
	return n;
}

} // extern "C"
