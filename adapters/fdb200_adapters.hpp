/*
 * fdb200_adapters.hpp - C++ adapters that put the B200 path behind the reference's OWN interfaces.
 *
 * Compile this header inside the reference tree (it includes the reference's unchanged headers and
 * OpenCV) and link against libfdb200.so; see INTEGRATION.md.  It implements
 *   imageprocessing::PyramidFeatureExtractor   (libImageProcessing/include/imageprocessing/PyramidFeatureExtractor.hpp:52-118,
 *                                               FeatureExtractor.hpp:32-53)
 *   classification::ProbabilisticClassifier    (libClassification/include/classification/ProbabilisticClassifier.hpp:33,
 *                                               BinaryClassifier.hpp:32,42)
 *   detection::Detector                        (libDetection/include/detection/Detector.hpp:59-79)
 * on top of the C ABI of include/fdb200.h.  Conventions kept from the reference: shared_ptr
 * ownership, borrowed const cv::Mat& inputs, std::invalid_argument / std::runtime_error for
 * failures, an empty shared_ptr for an out-of-bounds single extraction
 * (DirectPyramidFeatureExtractor.cpp:135-136), objects not re-entrant.
 */
#ifndef FDB200_ADAPTERS_HPP_
#define FDB200_ADAPTERS_HPP_

#include "classification/ProbabilisticClassifier.hpp"
#include "detection/ClassifiedPatch.hpp"
#include "detection/Detector.hpp"
#include "imageprocessing/Patch.hpp"
#include "imageprocessing/PyramidFeatureExtractor.hpp"
#include "imageprocessing/VersionedImage.hpp"

#include "fdb200.h"

#include <algorithm>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fdb200 {

inline void check(int status) {
	if (status == FDB_OK) return;
	const std::string msg = std::string("fdb200: ") + fdb_last_error();
	if (status == FDB_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
	throw std::runtime_error(msg);
}

/* one CUDA device + stream; shared by the objects created from it */
class Context {
public:
	explicit Context(int device = -1) : handle(nullptr) { check(fdb_ctx_create(device, &handle)); }
	~Context() { fdb_ctx_destroy(handle); }
	Context(const Context&) = delete;
	Context& operator=(const Context&) = delete;
	fdb_ctx* get() const { return handle; }
private:
	fdb_ctx* handle;
};

/* GrayscaleFilter::applyTo (GrayscaleFilter.cpp:18-24): 1-channel frames are taken as they are (copy branch), 3-channel frames go
 * through cvtColor(CV_BGR2GRAY) - on the device (fdb_gray_from_bgr) */
inline cv::Mat grayOf(fdb_ctx* ctx, const cv::Mat& image) {
	if (image.depth() != CV_8U || (image.channels() != 1 && image.channels() != 3))
		throw std::invalid_argument("fdb200: frames must be 8-bit with 1 (gray) or 3 (BGR) channels");
	if (image.channels() == 1) return image;
	cv::Mat gray(image.rows, image.cols, CV_8UC1);
	check(fdb_gray_from_bgr(ctx, image.ptr<unsigned char>(0), (int64_t)image.step, image.cols, image.rows, 1, gray.ptr<unsigned char>(0)));
	return gray;
}

/* classification::ProbabilisticWvmClassifier replacement (ProbabilisticWvmClassifier.cpp:42-54) */
class B200ProbabilisticWvmClassifier : public classification::ProbabilisticClassifier {
public:
	B200ProbabilisticWvmClassifier(std::shared_ptr<Context> context, const fdb_wvm_desc& desc) :
			context(context), handle(nullptr), dim(desc.filter_size_x * desc.filter_size_y) {
		check(fdb_wvm_create(context->get(), &desc, &handle));
	}
	~B200ProbabilisticWvmClassifier() { fdb_wvm_destroy(handle); }

	bool classify(const cv::Mat& featureVector) const { return evaluate(featureVector).positive; }
	std::pair<bool, double> getConfidence(const cv::Mat& featureVector) const {
		const Result r = evaluate(featureVector); /* WvmClassifier::getConfidence(pair), WvmClassifier.cpp:84-89 */
		return std::make_pair(r.positive, r.positive ? (double)r.fout : -(double)r.fout);
	}
	std::pair<bool, double> getProbability(const cv::Mat& featureVector) const {
		const Result r = evaluate(featureVector);
		return std::make_pair(r.positive, r.probability);
	}
	/* WvmClassifier::computeHyperplaneDistance: index of the last filter and its output */
	std::pair<int, double> computeHyperplaneDistance(const cv::Mat& featureVector) const {
		const Result r = evaluate(featureVector);
		return std::make_pair(r.level, (double)r.fout);
	}
	void setLimitReliabilityFilter(float value) { check(fdb_wvm_set_limit_reliability_filter(handle, value)); }
	fdb_wvm* get() const { return handle; }

private:
	struct Result { int level; float fout; double probability; bool positive; };
	Result evaluate(const cv::Mat& v) const {
		if (v.depth() != CV_8U || !v.isContinuous() || (int)(v.total() * v.channels()) != dim)
			throw std::invalid_argument("fdb200: WVM feature vectors are continuous 8-bit patches of the filter size");
		int32_t level = 0; float fout = 0; double prob = 0; uint8_t pos = 0;
		check(fdb_wvm_get_probability(handle, v.ptr<uchar>(0), 1, &level, &fout, &prob, &pos));
		Result r = {level, fout, prob, pos != 0};
		return r;
	}
	std::shared_ptr<Context> context;
	fdb_wvm* handle;
	int dim;
};

/* classification::ProbabilisticSvmClassifier replacement (ProbabilisticSvmClassifier.cpp:42-58) */
class B200ProbabilisticSvmClassifier : public classification::ProbabilisticClassifier {
public:
	B200ProbabilisticSvmClassifier(std::shared_ptr<Context> context, const fdb_svm_desc& desc) :
			context(context), handle(nullptr), dim(desc.dim), type(desc.sv_type) {
		check(fdb_svm_create(context->get(), &desc, &handle));
	}
	~B200ProbabilisticSvmClassifier() { fdb_svm_destroy(handle); }

	bool classify(const cv::Mat& featureVector) const { return evaluate(featureVector).positive; }
	std::pair<bool, double> getConfidence(const cv::Mat& featureVector) const {
		const Result r = evaluate(featureVector); /* SvmClassifier::getConfidence, SvmClassifier.cpp:48-53 */
		return std::make_pair(r.positive, r.positive ? r.distance : -r.distance);
	}
	std::pair<bool, double> getProbability(const cv::Mat& featureVector) const {
		const Result r = evaluate(featureVector);
		return std::make_pair(r.positive, r.probability);
	}
	double computeHyperplaneDistance(const cv::Mat& featureVector) const { return evaluate(featureVector).distance; }
	void setThreshold(float threshold) { check(fdb_svm_set_threshold(handle, threshold)); }
	fdb_svm* get() const { return handle; }

private:
	struct Result { double distance, probability; bool positive; };
	Result evaluate(const cv::Mat& v) const {
		const int want = type == FDB_SV_U8 ? CV_8U : CV_32F;
		if (v.depth() != want || !v.isContinuous() || (int)(v.total() * v.channels()) != dim)
			throw std::invalid_argument("fdb200: SVM feature vector has the wrong type or length"); /* RbfKernel.hpp:33-38 */
		double d = 0, p = 0; uint8_t pos = 0;
		check(fdb_svm_get_probability(handle, v.ptr<uchar>(0), 1, &d, &p, &pos));
		Result r = {d, p, pos != 0};
		return r;
	}
	std::shared_ptr<Context> context;
	fdb_svm* handle;
	int dim, type;
};

/* classification::ProbabilisticRvmClassifier replacement (ProbabilisticRvmClassifier.cpp:42-64 over RvmClassifier.cpp:50-112) */
class B200ProbabilisticRvmClassifier : public classification::ProbabilisticClassifier {
public:
	B200ProbabilisticRvmClassifier(std::shared_ptr<Context> context, const fdb_rvm_desc& desc) :
			context(context), handle(nullptr), dim(desc.dim), type(desc.sv_type) {
		check(fdb_rvm_create(context->get(), &desc, &handle));
	}
	~B200ProbabilisticRvmClassifier() { fdb_rvm_destroy(handle); }

	bool classify(const cv::Mat& featureVector) const { return evaluate(featureVector).positive; }
	std::pair<bool, double> getConfidence(const cv::Mat& featureVector) const {
		const Result r = evaluate(featureVector); /* RvmClassifier::getConfidence(pair), RvmClassifier.cpp:59-64 */
		return std::make_pair(r.positive, r.positive ? r.distance : -r.distance);
	}
	std::pair<bool, double> getProbability(const cv::Mat& featureVector) const {
		const Result r = evaluate(featureVector);
		return std::make_pair(r.positive, r.probability);
	}
	/* RvmClassifier::computeHyperplaneDistance: (level reached, distance) */
	std::pair<int, double> computeHyperplaneDistance(const cv::Mat& featureVector) const {
		const Result r = evaluate(featureVector);
		return std::make_pair(r.level, r.distance);
	}
	void setNumFiltersToUse(unsigned int numFilters) { check(fdb_rvm_set_num_filters_to_use(handle, (int32_t)numFilters)); }
	fdb_rvm* get() const { return handle; }

private:
	struct Result { int level; double distance, probability; bool positive; };
	Result evaluate(const cv::Mat& v) const {
		const int want = type == FDB_SV_U8 ? CV_8U : CV_32F;
		if (v.depth() != want || !v.isContinuous() || (int)(v.total() * v.channels()) != dim)
			throw std::invalid_argument("fdb200: RVM feature vector has the wrong type or length");
		int32_t level = 0; double d = 0, p = 0; uint8_t pos = 0;
		check(fdb_rvm_get_probability(handle, v.ptr<uchar>(0), 1, &level, &d, &p, &pos));
		Result r = {level, d, p, pos != 0};
		return r;
	}
	std::shared_ptr<Context> context;
	fdb_rvm* handle;
	int dim, type;
};

/* detection::Detector replacement: FiveStageSlidingWindowDetector (stage = FDB_STAGE_NMS, svm != null)
 * or plain SlidingWindowDetector (stage = FDB_STAGE_WVM). It is also the PyramidFeatureExtractor of the
 * reference graph (ImagePyramid + GrayscaleFilter + DirectPyramidFeatureExtractor + HistEq64Filter). */
class B200SlidingWindowDetector : public detection::Detector, public imageprocessing::PyramidFeatureExtractor {
public:
	B200SlidingWindowDetector(std::shared_ptr<Context> context, const fdb_detector_desc& desc,
			std::shared_ptr<B200ProbabilisticWvmClassifier> wvm, std::shared_ptr<B200ProbabilisticSvmClassifier> svm,
			int stage = FDB_STAGE_NMS) :
			context(context), wvm(wvm), svm(svm), desc(desc), handle(nullptr), stage(svm ? stage : FDB_STAGE_WVM),
			width(0), height(0) {
		check(fdb_detector_create(context->get(), &desc, wvm->get(), svm ? svm->get() : nullptr, &handle));
	}
	~B200SlidingWindowDetector() { fdb_detector_destroy(handle); }
	fdb_detector* get() const { return handle; }
	const cv::Mat& currentFrame() const { return current; } /* gray frame of the last update() */

	/* ---- detection::Detector ---- */
	std::vector<std::shared_ptr<detection::ClassifiedPatch>> detect(const cv::Mat& image) {
		const cv::Mat gray = grayOf(context->get(), image);
		prepareFor(gray);
		return run(gray, cv::Rect(), false);
	}
	std::vector<std::shared_ptr<detection::ClassifiedPatch>> detect(const cv::Mat& image, const cv::Rect& roi) {
		const cv::Mat gray = grayOf(context->get(), image);
		prepareFor(gray);
		return run(gray, roi, true);
	}
	std::vector<std::shared_ptr<detection::ClassifiedPatch>> detect(std::shared_ptr<imageprocessing::VersionedImage> image) {
		return detect(image->getData());
	}

	/* ---- imageprocessing::FeatureExtractor / PyramidFeatureExtractor ---- */
	using imageprocessing::FeatureExtractor::update;
	void update(std::shared_ptr<imageprocessing::VersionedImage> image) {
		const cv::Mat gray = grayOf(context->get(), image->getData());
		prepareFor(gray);
		current = gray.clone();
		patches.clear();
	}
	std::shared_ptr<imageprocessing::Patch> extract(int x, int y, int w, int h) const {
		/* DirectPyramidFeatureExtractor.cpp:67-73: the layer whose patch width is closest to w */
		const double power = std::log((double)desc.patch_width / (double)w) / std::log(incrementalScale());
		const int index = (int)std::round(power); /* ImagePyramid.cpp:307-310 */
		const fdb_layer_info* L = findLayer(index);
		if (!L) return std::shared_ptr<imageprocessing::Patch>();
		return extractAt(*L, cvRound((x - w / 2) * L->scale), cvRound((y - h / 2) * L->scale));
	}
	std::vector<std::shared_ptr<imageprocessing::Patch>> extract(int stepX, int stepY, cv::Rect roi = cv::Rect(),
			int firstLayer = -1, int lastLayer = -1, int stepLayer = 1) const {
		if (stepX < 1) throw std::invalid_argument("DirectPyramidFeatureExtractor: stepX has to be greater than zero");
		if (stepY < 1) throw std::invalid_argument("DirectPyramidFeatureExtractor: stepY has to be greater than zero");
		if (stepLayer < 1) throw std::invalid_argument("DirectPyramidFeatureExtractor: stepLayer has to be greater than zero");
		if (current.empty()) return std::vector<std::shared_ptr<imageprocessing::Patch>>();
		std::vector<std::shared_ptr<imageprocessing::Patch>> out;
		const int dim = desc.patch_width * desc.patch_height;
		const bool whole = roi.x == 0 && roi.y == 0 && roi.width == 0 && roi.height == 0;
		if (whole && stepX == desc.step_x && stepY == desc.step_y && stepLayer == 1) {
			/* the detector's own scan: every patch in one call */
			ensurePatches();
			for (size_t li = 0; li < layers.size(); ++li) {
				const fdb_layer_info& L = layers[li];
				if ((firstLayer >= 0 && L.index < firstLayer) || (lastLayer >= 0 && L.index > lastLayer)) continue;
				for (int iy = 0; iy < L.windows_y; ++iy)
					for (int ix = 0; ix < L.windows_x; ++ix) {
						const int64_t w = L.first_window + (int64_t)iy * L.windows_x + ix;
						out.push_back(makePatch(L, ix * desc.step_x, iy * desc.step_y, &patches[(size_t)w * dim]));
					}
			}
			return out;
		}
		/* any other step / region / layer stride: the window list of DirectPyramidFeatureExtractor.cpp:84-121, patches
		 * from the device in one call */
		cv::Rect r = roi;
		if (whole) { r.x = 0; r.y = 0; r.width = width; r.height = height; }
		else { /* :87-92 */
			const int x0 = std::max(0, r.x), y0 = std::max(0, r.y);
			r.width = std::min(width, r.width + x0) - x0; r.height = std::min(height, r.height + y0) - y0;
			r.x = x0; r.y = y0;
		}
		std::vector<int32_t> lxy;
		std::vector<const fdb_layer_info*> owner;
		for (size_t li = 0; li < layers.size(); li += (size_t)stepLayer) { /* :99: the stride runs over the whole layer list */
			const fdb_layer_info& L = layers[li];
			if ((firstLayer >= 0 && L.index < firstLayer) || (lastLayer >= 0 && L.index > lastLayer)) continue;
			const int bx = cvRound(r.x * L.scale), by = cvRound(r.y * L.scale);             /* :110-111 */
			const int ex = cvRound((r.x + r.width) * L.scale), ey = cvRound((r.y + r.height) * L.scale);
			for (int y = by; y + desc.patch_height < ey; y += stepY)                         /* strict '<', :113-114 */
				for (int x = bx; x + desc.patch_width < ex; x += stepX) {
					lxy.push_back(L.index); lxy.push_back(x); lxy.push_back(y);
					owner.push_back(&L);
				}
		}
		const int64_t n = (int64_t)owner.size();
		if (n == 0) return out;
		std::vector<uint8_t> data((size_t)n * dim), valid((size_t)n);
		check(fdb_extract_windows(handle, current.ptr<uchar>(0), (int64_t)current.step, &lxy[0], n, &data[0], &valid[0]));
		for (int64_t i = 0; i < n; ++i)
			if (valid[(size_t)i]) out.push_back(makePatch(*owner[(size_t)i], lxy[3 * i + 1], lxy[3 * i + 2], &data[(size_t)i * dim]));
		return out;
	}
	std::shared_ptr<imageprocessing::Patch> extract(int layer, int x, int y) const {
		const fdb_layer_info* L = findLayer(layer);
		if (!L) return std::shared_ptr<imageprocessing::Patch>();
		return extractAt(*L, x - desc.patch_width / 2, y - desc.patch_height / 2); /* DirectPyramidFeatureExtractor.cpp:125-131 */
	}
	int getLayerIndex(int w, int /*h*/) const {
		const double power = std::log((double)desc.patch_width / (double)w) / std::log(incrementalScale());
		const fdb_layer_info* L = findLayer((int)std::round(power));
		return L ? L->index : -1;
	}
	double getMinScaleFactor() const { return desc.min_scale_factor; }
	double getMaxScaleFactor() const { return desc.max_scale_factor; }
	double getIncrementalScaleFactor() const { return incrementalScale(); }
	cv::Size getPatchSize() const { return cv::Size(desc.patch_width, desc.patch_height); }
	cv::Size getImageSize() const { return cv::Size(width, height); }
	std::vector<std::pair<int, double>> getLayerScales() const {
		std::vector<std::pair<int, double>> v;
		for (size_t i = 0; i < layers.size(); ++i) v.push_back(std::make_pair(layers[i].index, layers[i].scale));
		return v;
	}
	std::vector<cv::Size> getLayerSizes() const {
		std::vector<cv::Size> v;
		for (size_t i = 0; i < layers.size(); ++i) v.push_back(cv::Size(layers[i].width, layers[i].height));
		return v;
	}
	std::vector<cv::Size> getPatchSizes() const {
		std::vector<cv::Size> v;
		for (size_t i = 0; i < layers.size(); ++i) v.push_back(cv::Size(layers[i].orig_patch_width, layers[i].orig_patch_height));
		return v;
	}

private:
	double incrementalScale() const {
		/* ImagePyramid.cpp:90-91 */
		const double olc = std::floor(std::log(0.5) / std::log(desc.incremental_scale_factor) + 0.5);
		return std::pow(0.5, 1. / olc);
	}
	void prepareFor(const cv::Mat& image) {
		if (image.cols == width && image.rows == height) return;
		check(fdb_detector_prepare(handle, image.cols, image.rows, 1));
		width = image.cols; height = image.rows;
		int32_t n = 0;
		check(fdb_detector_layers(handle, nullptr, 0, &n));
		layers.resize((size_t)n);
		if (n) check(fdb_detector_layers(handle, &layers[0], n, &n));
		patches.clear();
	}
	const fdb_layer_info* findLayer(int index) const {
		for (size_t i = 0; i < layers.size(); ++i) if (layers[i].index == index) return &layers[i];
		return nullptr;
	}
	void ensurePatches() const {
		if (!patches.empty() || current.empty()) return;
		const int64_t nwin = fdb_detector_windows_per_frame(handle);
		patches.resize((size_t)nwin * desc.patch_width * desc.patch_height);
		int64_t got = 0;
		if (nwin) check(fdb_extract_patches(handle, current.ptr<uchar>(0), (int64_t)current.step, &patches[0], nwin, &got));
	}
	std::shared_ptr<imageprocessing::Patch> makePatch(const fdb_layer_info& L, int x, int y, const uint8_t* data) const {
		cv::Mat m(desc.patch_height, desc.patch_width, CV_8U);
		for (int r = 0; r < desc.patch_height; ++r)
			for (int c = 0; c < desc.patch_width; ++c) m.ptr<uchar>(r)[c] = data[r * desc.patch_width + c];
		/* DirectPyramidFeatureExtractor.cpp:115-118 */
		const int ox = cvRound(x / L.scale) + L.orig_patch_width / 2, oy = cvRound(y / L.scale) + L.orig_patch_height / 2;
		return std::make_shared<imageprocessing::Patch>(ox, oy, L.orig_patch_width, L.orig_patch_height, m);
	}
	std::shared_ptr<imageprocessing::Patch> extractAt(const fdb_layer_info& L, int x, int y) const {
		/* DirectPyramidFeatureExtractor.cpp:133-143: not inside the layer image => empty pointer */
		if (x < 0 || y < 0 || x + desc.patch_width > L.width || y + desc.patch_height > L.height || current.empty())
			return std::shared_ptr<imageprocessing::Patch>();
		const int dim = desc.patch_width * desc.patch_height;
		if (desc.step_x == 1 && desc.step_y == 1 && x < L.windows_x && y < L.windows_y) { /* inside the cached step-1 scan */
			ensurePatches();
			const int64_t w = L.first_window + (int64_t)y * L.windows_x + x;
			return makePatch(L, x, y, &patches[(size_t)w * dim]);
		}
		/* the last rows / columns of a layer (the scan's strict '<' bound leaves them out) and coarser scans: one window */
		const int32_t lxy[3] = {L.index, x, y};
		std::vector<uint8_t> data((size_t)dim);
		uint8_t valid = 0;
		check(fdb_extract_windows(handle, current.ptr<uchar>(0), (int64_t)current.step, lxy, 1, &data[0], &valid));
		return valid ? makePatch(L, x, y, &data[0]) : std::shared_ptr<imageprocessing::Patch>();
	}
	std::vector<std::shared_ptr<detection::ClassifiedPatch>> run(const cv::Mat& image, const cv::Rect& roi, bool useRoi) {
		std::vector<fdb_detection> dets(4096);
		int64_t n = 0;
		for (;;) {
			const int status = useRoi
					? fdb_detect_roi(handle, image.ptr<uchar>(0), (int64_t)image.step, roi.x, roi.y, roi.width, roi.height, stage, &dets[0], (int64_t)dets.size(), &n)
					: fdb_detect_batch(handle, image.ptr<uchar>(0), (int64_t)image.step, 1, stage, nullptr, &dets[0], (int64_t)dets.size(), &n);
			if (status == FDB_ERR_OVERFLOW && n > (int64_t)dets.size()) { dets.resize((size_t)n); continue; }
			check(status);
			break;
		}
		/* the reference hands the extractor's patch (the HistEq64 data) along with every ClassifiedPatch
		 * (SlidingWindowDetector.cpp:93-96; read again by FiveStageSlidingWindowDetector.cpp:258-261 and the trackers):
		 * the patches of the detections come back in one call */
		const int dim = desc.patch_width * desc.patch_height;
		std::vector<int32_t> lxy((size_t)n * 3);
		for (int64_t i = 0; i < n; ++i) { lxy[3 * i] = dets[(size_t)i].layer; lxy[3 * i + 1] = dets[(size_t)i].x; lxy[3 * i + 2] = dets[(size_t)i].y; }
		std::vector<uint8_t> data((size_t)n * dim), valid((size_t)n);
		if (n) check(fdb_extract_windows(handle, image.ptr<uchar>(0), (int64_t)image.step, &lxy[0], n, &data[0], &valid[0]));
		std::vector<std::shared_ptr<detection::ClassifiedPatch>> out;
		for (int64_t i = 0; i < n; ++i) {
			const fdb_detection& d = dets[(size_t)i];
			cv::Mat m(desc.patch_height, desc.patch_width, CV_8U);
			for (int r = 0; r < desc.patch_height; ++r)
				for (int c = 0; c < desc.patch_width; ++c) m.ptr<uchar>(r)[c] = data[(size_t)i * dim + r * desc.patch_width + c];
			std::shared_ptr<imageprocessing::Patch> patch = std::make_shared<imageprocessing::Patch>(d.center_x, d.center_y, d.width, d.height, m);
			out.push_back(std::make_shared<detection::ClassifiedPatch>(patch, d.positive != 0, d.probability));
		}
		return out;
	}

	std::shared_ptr<Context> context;
	std::shared_ptr<B200ProbabilisticWvmClassifier> wvm;
	std::shared_ptr<B200ProbabilisticSvmClassifier> svm;
	fdb_detector_desc desc;
	fdb_detector* handle;
	int stage;
	int width, height;
	std::vector<fdb_layer_info> layers;
	cv::Mat current;
	mutable std::vector<uint8_t> patches;
};

/* detection::Detector replacement for ffpDetectApp's `single` detectors (ffpDetectApp.cpp:427-500): a SlidingWindowDetector
 * whose classifier is a ProbabilisticSvmClassifier (psvm) or ProbabilisticRvmClassifier (prvm) - every window is classified
 * (SlidingWindowDetector.cpp:87-98). u8 RBF SVMs on 20x20 patches run on the tensor cores (csrc/svm_dense.cu). */
class B200SingleDetector : public detection::Detector {
public:
	B200SingleDetector(std::shared_ptr<Context> context, const fdb_detector_desc& desc, std::shared_ptr<B200ProbabilisticSvmClassifier> svm) :
			context(context), svm(svm), handle(nullptr), width(0), height(0) {
		check(fdb_detector_create(context->get(), &desc, nullptr, svm->get(), &handle));
	}
	B200SingleDetector(std::shared_ptr<Context> context, const fdb_detector_desc& desc, std::shared_ptr<B200ProbabilisticRvmClassifier> rvm) :
			context(context), rvm(rvm), handle(nullptr), width(0), height(0) {
		check(fdb_detector_create_rvm(context->get(), &desc, rvm->get(), &handle));
	}
	~B200SingleDetector() { fdb_detector_destroy(handle); }
	fdb_detector* get() const { return handle; }

	std::vector<std::shared_ptr<detection::ClassifiedPatch>> detect(const cv::Mat& image) {
		const cv::Mat gray = grayOf(context->get(), image);
		if (gray.cols != width || gray.rows != height) {
			check(fdb_detector_prepare(handle, gray.cols, gray.rows, 1));
			width = gray.cols; height = gray.rows;
		}
		const int64_t cap = fdb_detector_windows_per_frame(handle);
		std::vector<fdb_detection> dets((size_t)(cap > 0 ? cap : 1));
		int64_t n = 0;
		check(fdb_detect_single(handle, gray.ptr<uchar>(0), (int64_t)gray.step, 1, nullptr, &dets[0], (int64_t)dets.size(), &n));
		std::vector<std::shared_ptr<detection::ClassifiedPatch>> out;
		for (int64_t i = 0; i < n; ++i) {
			const fdb_detection& d = dets[(size_t)i];
			std::shared_ptr<imageprocessing::Patch> patch = std::make_shared<imageprocessing::Patch>(d.center_x, d.center_y, d.width, d.height, cv::Mat());
			out.push_back(std::make_shared<detection::ClassifiedPatch>(patch, d.positive != 0, d.probability));
		}
		return out;
	}
	/* SlidingWindowDetector::detect(image, roi) (SlidingWindowDetector.cpp:53-79): what ffpDetectApp.cpp:591 calls on every
	 * feature detector with the bounds of the face */
	std::vector<std::shared_ptr<detection::ClassifiedPatch>> detect(const cv::Mat& image, const cv::Rect& roi) {
		const cv::Mat gray = grayOf(context->get(), image);
		if (gray.cols != width || gray.rows != height) {
			check(fdb_detector_prepare(handle, gray.cols, gray.rows, 1));
			width = gray.cols; height = gray.rows;
		}
		const int64_t cap = fdb_detector_windows_per_frame(handle);
		std::vector<fdb_detection> dets((size_t)(cap > 0 ? cap : 1));
		int64_t n = 0;
		check(fdb_detect_single_roi(handle, gray.ptr<uchar>(0), (int64_t)gray.step, roi.x, roi.y, roi.width, roi.height, &dets[0],
				(int64_t)dets.size(), &n));
		std::vector<std::shared_ptr<detection::ClassifiedPatch>> out;
		for (int64_t i = 0; i < n; ++i) {
			const fdb_detection& d = dets[(size_t)i];
			std::shared_ptr<imageprocessing::Patch> patch = std::make_shared<imageprocessing::Patch>(d.center_x, d.center_y, d.width, d.height, cv::Mat());
			out.push_back(std::make_shared<detection::ClassifiedPatch>(patch, d.positive != 0, d.probability));
		}
		return out;
	}
	std::vector<std::shared_ptr<detection::ClassifiedPatch>> detect(std::shared_ptr<imageprocessing::VersionedImage> image) {
		return detect(image->getData());
	}

private:
	std::shared_ptr<Context> context;
	std::shared_ptr<B200ProbabilisticSvmClassifier> svm;
	std::shared_ptr<B200ProbabilisticRvmClassifier> rvm;
	fdb_detector* handle;
	int width, height;
};

} // namespace fdb200
#endif
