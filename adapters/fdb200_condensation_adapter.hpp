/*
 * fdb200_condensation_adapter.hpp - condensation::MeasurementModel on top of the fdb200 C ABI: the drop-in for
 * condensation::WvmSvmModel (libCondensation/include/condensation/WvmSvmModel.hpp:28-64,
 * libCondensation/src/condensation/WvmSvmModel.cpp:35-119), the tracker's sparse caller of the WVM -> SVM pair.
 * Compiles against the reference's unchanged headers (tests/test_adapters_compile.py).
 *
 *   auto det   = std::make_shared<fdb200::B200SlidingWindowDetector>(ctx, desc, wvm, svm);   // owns pyramid + classifiers
 *   auto model = std::make_shared<fdb200::B200WvmSvmModel>(det);
 *   model->evaluate(image, samples);            // one fdb_evaluate_samples call: all particles of the frame
 */
#ifndef FDB200_CONDENSATION_ADAPTER_HPP_
#define FDB200_CONDENSATION_ADAPTER_HPP_

#include "condensation/MeasurementModel.hpp"
#include "condensation/Sample.hpp"
#include "imageprocessing/VersionedImage.hpp"

#include "fdb200_adapters.hpp"

namespace fdb200 {

class B200WvmSvmModel : public condensation::MeasurementModel {
public:
	/* maxSvmPatches: the reference evaluates the SVM on the 8 most probable WVM positives (WvmSvmModel.cpp:96-99) */
	explicit B200WvmSvmModel(std::shared_ptr<B200SlidingWindowDetector> detector, int maxSvmPatches = 8) :
			detector(detector), maxSvmPatches(maxSvmPatches) {}

	void update(std::shared_ptr<imageprocessing::VersionedImage> image) { detector->update(image); }

	/* WvmSvmModel.cpp:44-72: one sample, no top-8 cut */
	void evaluate(condensation::Sample& sample) const {
		const int32_t xywh[4] = {sample.getX(), sample.getY(), sample.getWidth(), sample.getHeight()};
		uint8_t target = 0;
		double weight = 0;
		const cv::Mat& frame = detector->currentFrame();
		check(fdb_evaluate_samples(detector->get(), frame.ptr<uchar>(0), (int64_t)frame.step, xywh, 1, 0, &target, &weight));
		sample.setTarget(target != 0);
		sample.setWeight(weight);
	}

	/* WvmSvmModel.cpp:74-119: all samples of a frame */
	void evaluate(std::shared_ptr<imageprocessing::VersionedImage> image, std::vector<std::shared_ptr<condensation::Sample>>& samples) {
		update(image);
		const size_t n = samples.size();
		if (!n) return;
		std::vector<int32_t> xywh(4 * n);
		for (size_t i = 0; i < n; ++i) {
			xywh[4 * i] = samples[i]->getX(); xywh[4 * i + 1] = samples[i]->getY();
			xywh[4 * i + 2] = samples[i]->getWidth(); xywh[4 * i + 3] = samples[i]->getHeight();
		}
		std::vector<uint8_t> target(n);
		std::vector<double> weight(n);
		const cv::Mat& frame = detector->currentFrame();
		check(fdb_evaluate_samples(detector->get(), frame.ptr<uchar>(0), (int64_t)frame.step, &xywh[0], (int64_t)n, maxSvmPatches,
				&target[0], &weight[0]));
		for (size_t i = 0; i < n; ++i) {
			samples[i]->setTarget(target[i] != 0);
			samples[i]->setWeight(weight[i]);
		}
	}

private:
	std::shared_ptr<B200SlidingWindowDetector> detector;
	int maxSvmPatches;
};

} // namespace fdb200
#endif
