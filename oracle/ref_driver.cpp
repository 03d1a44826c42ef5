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
#include "classification/WvmClassifier.hpp"
#include "classification/ProbabilisticWvmClassifier.hpp"
#include "classification/SvmClassifier.hpp"
#include "classification/ProbabilisticSvmClassifier.hpp"
#include "classification/RbfKernel.hpp"
#include "imageprocessing/HistEq64Filter.hpp"
#include "imageprocessing/Patch.hpp"
#include "detection/ClassifiedPatch.hpp"
#include "detection/OverlapElimination.hpp"

#include "fdb200.h"

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

void* ref_svm_create(const fdb_svm_desc* d) {
	RefSvm* r = new RefSvm;
	r->svm = make_shared<classification::SvmClassifier>(make_shared<classification::RbfKernel>(d->gamma));
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

} // extern "C"
