/* Executes the adapters of fdb200_adapters.hpp - the reference's own interfaces (Detector, PyramidFeatureExtractor,
 * ProbabilisticClassifier) on top of libfdb200.so - on a GPU and prints what they return; tests/test_gpu_adapters.py compares the
 * output with the oracle. Built against the reference's unchanged headers + the cv::Mat stand-in of oracle/shim by oracle/Makefile
 * (where /root/reference is mounted) into oracle/_ref/run_adapters, which travels to the GPU box like the other prebuilt checkers.
 * usage: run_adapters wvm.mat thresholds.mat svm.mat frame.raw W H inc min max pw ph */
#include "fdb200_adapters.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

using namespace fdb200;

static unsigned long long checksum(const cv::Mat& m) {
	unsigned long long h = 1469598103934665603ull;
	for (int r = 0; r < m.rows; ++r)
		for (int c = 0; c < m.cols; ++c) { h ^= m.ptr<uchar>(r)[c]; h *= 1099511628211ull; }
	return h;
}

static void print_dets(const char* tag, const std::vector<std::shared_ptr<detection::ClassifiedPatch>>& v) {
	std::printf("%s %zu\n", tag, v.size());
	for (size_t i = 0; i < v.size(); ++i) {
		const std::shared_ptr<imageprocessing::Patch> p = v[i]->getPatch();
		std::printf("%s_ITEM %d %d %d %d %.17g %d %d %d %llu\n", tag, p->getX(), p->getY(), p->getWidth(), p->getHeight(), v[i]->getProbability(),
				v[i]->isPositive() ? 1 : 0, p->getData().rows, p->getData().cols, p->getData().rows ? checksum(p->getData()) : 0ull);
	}
}

static void print_patches(const char* tag, const std::vector<std::shared_ptr<imageprocessing::Patch>>& v) {
	unsigned long long h = 0, g = 0;
	for (size_t i = 0; i < v.size(); ++i) {
		h = h * 31 + checksum(v[i]->getData());
		g = g * 31 + (unsigned long long)(v[i]->getX() * 7919 + v[i]->getY() * 104729 + v[i]->getWidth() * 13 + v[i]->getHeight());
	}
	std::printf("%s %zu %llu %llu\n", tag, v.size(), h, g);
}

int main(int argc, char** argv) {
	if (argc < 12) { std::fprintf(stderr, "usage: run_adapters wvm.mat thr.mat svm.mat frame.raw W H inc min max pw ph\n"); return 2; }
	try {
		const int W = std::atoi(argv[5]), H = std::atoi(argv[6]);
		cv::Mat frame(H, W, CV_8U);
		{ std::ifstream f(argv[4], std::ios::binary); f.read(reinterpret_cast<char*>(frame.ptr<uchar>(0)), (std::streamsize)W * H); if (!f) throw std::runtime_error("frame file too short"); }
		fdb_wvm_file* wf = nullptr; fdb_svm_file* sf = nullptr;
		check(fdb_wvm_file_load(argv[1], argv[2], &wf));
		check(fdb_svm_mat_load(argv[3], nullptr, &sf));
		fdb_wvm_desc wd = *fdb_wvm_file_desc(wf);
		fdb_svm_desc sd = *fdb_svm_file_desc(sf);
		sd.logistic_a = 0.00556; sd.logistic_b = -2.95;
		fdb_detector_desc dd = fdb_detector_desc();
		dd.incremental_scale_factor = std::atof(argv[7]); dd.min_scale_factor = std::atof(argv[8]); dd.max_scale_factor = std::atof(argv[9]);
		dd.patch_width = std::atoi(argv[10]); dd.patch_height = std::atoi(argv[11]);
		dd.step_x = dd.step_y = 1; dd.oe_dist = 5.0f; dd.oe_ratio = 0.0f;
		std::shared_ptr<Context> ctx = std::make_shared<Context>(0);
		std::shared_ptr<B200ProbabilisticWvmClassifier> wvm = std::make_shared<B200ProbabilisticWvmClassifier>(ctx, wd);
		std::shared_ptr<B200ProbabilisticSvmClassifier> svm = std::make_shared<B200ProbabilisticSvmClassifier>(ctx, sd);
		std::shared_ptr<B200SlidingWindowDetector> five = std::make_shared<B200SlidingWindowDetector>(ctx, dd, wvm, svm);
		std::shared_ptr<B200SlidingWindowDetector> sliding = std::make_shared<B200SlidingWindowDetector>(ctx, dd, wvm, std::shared_ptr<B200ProbabilisticSvmClassifier>());
		std::shared_ptr<detection::Detector> det = five;

		print_dets("FIVE", det->detect(frame));                                        /* FiveStageSlidingWindowDetector::detect(Mat) */
		print_dets("FIVE_ROI", det->detect(frame, cv::Rect(150, 100, 330, 300)));      /* ... detect(Mat, Rect) */
		print_dets("WVM", sliding->detect(frame));                                     /* SlidingWindowDetector::detect(Mat) */

		std::shared_ptr<imageprocessing::PyramidFeatureExtractor> ex = sliding;
		ex->update(frame);
		print_patches("EXTRACT_ALL", ex->extract(1, 1));
		print_patches("EXTRACT_STEP", ex->extract(2, 3, cv::Rect(100, 80, 300, 250), -1, -1, 2));
		const std::vector<std::pair<int, double>> scales = ex->getLayerScales();
		const std::vector<cv::Size> sizes = ex->getLayerSizes();
		print_patches("EXTRACT_LAYERS", ex->extract(4, 4, cv::Rect(), scales[1].first, scales[3].first, 1));
		std::printf("LAYERS %zu\n", scales.size());
		/* single windows: inside the scan, the last column / row of a layer (outside the scan's strict bound), out of bounds */
		const int li = 2;
		const int pw = dd.patch_width, ph = dd.patch_height;
		const int xs[4] = {pw / 2 + 3, sizes[li].width - pw + pw / 2, sizes[li].width - pw + pw / 2 + 1, pw / 2 - 1};
		const int ys[4] = {ph / 2 + 5, sizes[li].height - ph + ph / 2, ph / 2 + 5, ph / 2};
		for (int k = 0; k < 4; ++k) {
			std::shared_ptr<imageprocessing::Patch> p = ex->extract(scales[li].first, xs[k], ys[k]);
			if (!p) std::printf("SINGLE %d none\n", k);
			else std::printf("SINGLE %d %d %d %d %d %llu\n", k, p->getX(), p->getY(), p->getWidth(), p->getHeight(), checksum(p->getData()));
		}
		std::shared_ptr<imageprocessing::Patch> byBox = ex->extract(320, 240, 160, 160);
		if (!byBox) std::printf("BYBOX none\n");
		else std::printf("BYBOX %d %d %d %d %llu\n", byBox->getX(), byBox->getY(), byBox->getWidth(), byBox->getHeight(), checksum(byBox->getData()));
		if (byBox) {
			std::shared_ptr<classification::ProbabilisticClassifier> c1 = wvm, c2 = svm;
			const std::pair<bool, double> a = c1->getProbability(byBox->getData()), b = c2->getProbability(byBox->getData());
			std::printf("CLASSIFY %d %.17g %d %.17g\n", a.first ? 1 : 0, a.second, b.first ? 1 : 0, b.second);
		}
		/* `single` psvm detector (ffpDetectApp.cpp:427-500) on a crop: whole image and region of interest */
		cv::Mat crop(120, 160, CV_8U);
		for (int r = 0; r < 120; ++r) for (int c = 0; c < 160; ++c) crop.ptr<uchar>(r)[c] = frame.ptr<uchar>(r)[c];
		fdb_detector_desc sdsc = dd; sdsc.min_scale_factor = 0.2; sdsc.max_scale_factor = 0.5;
		std::shared_ptr<detection::Detector> single = std::make_shared<B200SingleDetector>(ctx, sdsc, svm);
		print_dets("SINGLE_DET", single->detect(crop));
		print_dets("SINGLE_ROI", single->detect(crop, cv::Rect(30, 20, 100, 90)));
		fdb_wvm_file_free(wf); fdb_svm_file_free(sf);
	} catch (const std::exception& e) {
		std::fprintf(stderr, "run_adapters: %s\n", e.what());
		return 1;
	}
	return 0;
}
