/*
 * ref_features.cpp - thin C entry points over the reference's OWN, UNMODIFIED feature-space filters,
 * compiled from where they lie under /root/reference into oracle/_ref/libfdref.so (oracle/Makefile):
 *   imageprocessing::GradientBinningFilter   (libImageProcessing/src/imageprocessing/GradientBinningFilter.cpp)
 *   imageprocessing::HistogramFilter         (.../HistogramFilter.cpp)
 *   imageprocessing::SpatialHistogramFilter  (.../SpatialHistogramFilter.cpp)
 *   imageprocessing::HogFilter               (.../HogFilter.cpp)
 *   imageprocessing::ExtendedHogFilter       (.../ExtendedHogFilter.cpp)
 *   imageprocessing::LbpFilter               (.../LbpFilter.cpp + LbpFilter.hpp operators)
 * TEST INFRASTRUCTURE ONLY: pins oracle/fd_features.c.  The filters that are thin wrappers over OpenCV
 * (GradientFilter -> cv::Sobel, HistogramEqualizationFilter -> cv::equalizeHist, WhiteningFilter -> cv::dft)
 * are pinned against cv2 instead (tests/golden/cv2_features.npz).
 */
#include "imageprocessing/GradientBinningFilter.hpp"
#include "imageprocessing/SpatialHistogramFilter.hpp"
#include "imageprocessing/HogFilter.hpp"
#include "imageprocessing/ExtendedHogFilter.hpp"
#include "imageprocessing/LbpFilter.hpp"

#include "fdb200.h"

#include <cstring>
#include <memory>

using cv::Mat;
using namespace imageprocessing;

extern "C" {

/* the binning look-up tables, read back by filtering an image that holds every (gx, gy) pair:
 * pixel (row gy, col gx) = {gx, gy}  ->  table index gx | gy << 8 = row * 256 + col */
void ref_gradient_bin_luts(int bins, int signed_gradients, uint8_t* one_bin, uint8_t* two_bin) {
	Mat codes(256, 256, CV_8UC2);
	for (int y = 0; y < 256; ++y)
		for (int x = 0; x < 256; ++x) { codes.ptr<cv::Vec2b>(y)[x][0] = (uchar)x; codes.ptr<cv::Vec2b>(y)[x][1] = (uchar)y; }
	if (one_bin) {
		std::unique_ptr<GradientBinningFilter> f(new GradientBinningFilter(bins, signed_gradients != 0, false));
		Mat out = f->applyTo(codes);
		std::memcpy(one_bin, out.data, 65536 * 2);
	}
	if (two_bin) {
		std::unique_ptr<GradientBinningFilter> f(new GradientBinningFilter(bins, signed_gradients != 0, true));
		Mat out = f->applyTo(codes);
		std::memcpy(two_bin, out.data, 65536 * 4);
	}
}

void ref_lbp(const uint8_t* src, int w, int h, int lbp_type, uint8_t* out) {
	LbpFilter::Type t = lbp_type == FDB_LBP8 ? LbpFilter::Type::LBP8 : lbp_type == FDB_LBP8_UNIFORM ? LbpFilter::Type::LBP8_UNIFORM
			: lbp_type == FDB_LBP4 ? LbpFilter::Type::LBP4 : LbpFilter::Type::LBP4_ROTATED;
	LbpFilter f(t);
	Mat image(h, w, CV_8U, (void*)src);
	Mat filtered = f.applyTo(image);
	for (int y = 0; y < h; ++y) std::memcpy(out + (size_t)y * w, filtered.ptr<uchar>(y), (size_t)w);
}

/* the patch filter of the hog / ehog / lbp feature spaces, chosen as AdaptiveTracking::createHogFilter /
 * createHistogramFilter do (AdaptiveTracking.cpp:241-253,270-292), on a rows x cols ROI (row pitch in bytes) of a
 * binned layer with `channels` bytes per pixel. Returns the number of floats written. */
int ref_patch_histogram(const fdb_feature_desc* d, int bins, const uint8_t* roi, int pitch, int rows, int cols, int channels,
		float* out, int cap) {
	Mat image(rows, cols, CV_8UC(channels), (void*)roi, (size_t)pitch);
	std::shared_ptr<ImageFilter> filter;
	if (d->kind == FDB_FEATURE_EHOG) {
		filter = std::make_shared<ExtendedHogFilter>(bins, d->cell_size, d->interpolate_cells != 0, d->signed_and_unsigned != 0, d->ehog_alpha);
	} else if (d->kind == FDB_FEATURE_HOG && !(d->block_size == 1 && !d->signed_and_unsigned)) {
		filter = std::make_shared<HogFilter>(bins, d->cell_size, d->block_size, d->interpolate_cells != 0, d->signed_and_unsigned != 0);
	} else {
		HistogramFilter::Normalization n = HistogramFilter::Normalization::NONE;
		switch (d->normalization) {
		case FDB_NORM_L2NORM: n = HistogramFilter::Normalization::L2NORM; break;
		case FDB_NORM_L2HYS: n = HistogramFilter::Normalization::L2HYS; break;
		case FDB_NORM_L1NORM: n = HistogramFilter::Normalization::L1NORM; break;
		case FDB_NORM_L1SQRT: n = HistogramFilter::Normalization::L1SQRT; break;
		default: break;
		}
		filter = std::make_shared<SpatialHistogramFilter>(bins, d->cell_size, d->block_size, d->interpolate_cells != 0,
				d->concatenate != 0, n);
	}
	Mat result = filter->applyTo(image);
	const int count = result.rows * result.cols * result.channels();
	if (count > cap) return -count;
	for (int y = 0; y < result.rows; ++y)
		std::memcpy(out + (size_t)y * result.cols * result.channels(), result.ptr<float>(y), sizeof(float) * (size_t)result.cols * result.channels());
	return count;
}

} // extern "C"
