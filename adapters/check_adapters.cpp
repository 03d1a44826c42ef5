/* Compile check of the adapters against the reference's unchanged headers (tests/test_adapters_compile.py):
 * instantiates every adapter so that a missing pure-virtual override fails the build. */
#include "fdb200_adapters.hpp"
#include "fdb200_sdm_adapter.hpp"
#include "fdb200_condensation_adapter.hpp"

/* stand-in with the getters of superviseddescent::SdmLandmarkModel (SdmLandmarkModel.hpp:74-87; that header itself needs
 * OpenCV nonfree, Boost and libImageIO, which the compile check does not have) */
struct MockSdmModel {
	int getNumLandmarks() const { return 13; }
	int getNumCascadeSteps() const { return 1; }
	cv::Mat getMeanShape() const { return cv::Mat::zeros(26, 1, CV_32FC1); }
	cv::Mat getRegressorData(int) { return cv::Mat::zeros(13 * 279 + 1, 26, CV_32FC1); }
	std::string getDescriptorType(int) { return "vlhog-uoctti"; }
};

int main() {
	std::shared_ptr<fdb200::Context> ctx; /* not created: no GPU needed to type-check */
	fdb_wvm_desc wd = fdb_wvm_desc();
	fdb_svm_desc sd = fdb_svm_desc();
	fdb_detector_desc dd = fdb_detector_desc();
	if (ctx) {
		std::shared_ptr<fdb200::B200ProbabilisticWvmClassifier> wvm = std::make_shared<fdb200::B200ProbabilisticWvmClassifier>(ctx, wd);
		std::shared_ptr<fdb200::B200ProbabilisticSvmClassifier> svm = std::make_shared<fdb200::B200ProbabilisticSvmClassifier>(ctx, sd);
		fdb_rvm_desc rd = fdb_rvm_desc();
		std::shared_ptr<classification::ProbabilisticClassifier> c3 = std::make_shared<fdb200::B200ProbabilisticRvmClassifier>(ctx, rd);
		std::shared_ptr<detection::Detector> single1 = std::make_shared<fdb200::B200SingleDetector>(ctx, dd, svm);
		std::shared_ptr<detection::Detector> single2 = std::make_shared<fdb200::B200SingleDetector>(ctx, dd, std::make_shared<fdb200::B200ProbabilisticRvmClassifier>(ctx, rd));
		std::shared_ptr<classification::ProbabilisticClassifier> c1 = wvm, c2 = svm;
		std::shared_ptr<fdb200::B200SlidingWindowDetector> det = std::make_shared<fdb200::B200SlidingWindowDetector>(ctx, dd, wvm, svm);
		std::shared_ptr<detection::Detector> d = det;
		std::shared_ptr<imageprocessing::PyramidFeatureExtractor> e = det;
		cv::Mat frame(480, 640, CV_8U);
		d->detect(frame);
		e->update(frame);
		e->extract(1, 1);
		std::shared_ptr<condensation::MeasurementModel> mm = std::make_shared<fdb200::B200WvmSvmModel>(det);
		std::vector<std::shared_ptr<condensation::Sample>> particles(1, std::make_shared<condensation::Sample>(320, 240, 200));
		mm->evaluate(std::make_shared<imageprocessing::VersionedImage>(frame), particles);
		mm->evaluate(*particles[0]);
		fdb200::B200SdmLandmarkModelFitting fit(ctx->get(), MockSdmModel());
		cv::Mat shape = fit.alignRigid(MockSdmModel().getMeanShape(), cv::Rect(10, 10, 100, 100));
		shape = fit.optimize(shape, frame);
		fdb200::B200SdmLandmarkModelFitting fromFile(ctx->get(), std::string("model.txt"));
	}
	return 0;
}
