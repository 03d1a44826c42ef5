/*
 * fdb200_sdm_adapter.hpp - superviseddescent::SdmLandmarkModelFitting on the B200 path.
 *
 * The reference's fitting class is concrete (libSupervisedDescent/include/superviseddescent/SdmLandmarkModel.hpp:142-260): it
 * holds an SdmLandmarkModel by value and offers
 *     cv::Mat alignRigid(cv::Mat modelShape, cv::Rect faceBox) const     (:156-192)
 *     cv::Mat optimize(cv::Mat modelShape, cv::Mat image)               (:199-256)
 * B200SdmLandmarkModelFitting has the same two calls with the same argument meaning (column-vector shapes, all x then all y; 8-bit
 * 1-channel image) and error behaviour (std::runtime_error for a shape that is not a column vector; an exception where the
 * reference's Mat::operator()(Rect) throws because a descriptor window leaves the image), plus the batched call the GPU wants.
 * It is built from any model object with the reference's getters (getNumLandmarks, getNumCascadeSteps, getMeanShape,
 * getRegressorData(level), getDescriptorType(level): SdmLandmarkModel.cpp:43-71), i.e. from an unchanged
 * superviseddescent::SdmLandmarkModel - or straight from the model file (SdmLandmarkModel::load, :130-232).
 */
#ifndef FDB200_SDM_ADAPTER_HPP_
#define FDB200_SDM_ADAPTER_HPP_

#include "opencv2/core/core.hpp"

#include "fdb200.h"

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace fdb200 {

class B200SdmLandmarkModelFitting {
public:
	/* `Model` = superviseddescent::SdmLandmarkModel (taken by value there; read once here and uploaded) */
	template <class Model>
	B200SdmLandmarkModelFitting(fdb_ctx* ctx, Model model) : handle(nullptr) {
		const int L = model.getNumLandmarks(), steps = model.getNumCascadeSteps();
		cv::Mat mean = model.getMeanShape(); /* column vector, 2 L x 1 (SdmLandmarkModel.cpp:53-56) */
		std::vector<float> meanv(2 * (size_t)L);
		for (int i = 0; i < 2 * L; ++i) meanv[i] = mean.rows == 1 ? mean.at<float>(0, i) : mean.at<float>(i, 0);
		std::vector<cv::Mat> regs;
		std::vector<const float*> ptrs;
		for (int s = 0; s < steps; ++s) {
			if (model.getDescriptorType(s) != "vlhog-uoctti")
				throw std::runtime_error("fdb200: only vlhog-uoctti cascade steps run on the GPU");
			cv::Mat R = model.getRegressorData(s);
			if (R.rows != L * 279 + 1 || R.cols != 2 * L || R.depth() != CV_32F)
				throw std::runtime_error("fdb200: regressor is not (279 L + 1) x 2 L float32 (adaptive vlhog-uoctti parameters only)");
			regs.push_back(R.isContinuous() ? R : R.clone());
			ptrs.push_back(regs.back().template ptr<float>(0));
		}
		fdb_sdm_desc d = fdb_sdm_desc();
		d.num_landmarks = L; d.num_cascade_steps = steps; d.mean_landmarks = meanv.data(); d.regressors = ptrs.data();
		check(fdb_sdm_create(ctx, &d, &handle));
	}
	/* SdmLandmarkModel::load(filename) + SdmLandmarkModelFitting(model) */
	B200SdmLandmarkModelFitting(fdb_ctx* ctx, const std::string& modelFile) : handle(nullptr) {
		fdb_sdm_file* f = nullptr;
		check(fdb_sdm_file_load(modelFile.c_str(), &f));
		const int status = fdb_sdm_create(ctx, fdb_sdm_file_desc(f), &handle);
		fdb_sdm_file_free(f);
		check(status);
	}
	~B200SdmLandmarkModelFitting() { fdb_sdm_destroy(handle); }
	B200SdmLandmarkModelFitting(const B200SdmLandmarkModelFitting&) = delete;
	B200SdmLandmarkModelFitting& operator=(const B200SdmLandmarkModelFitting&) = delete;

	int getNumLandmarks() const { return fdb_sdm_num_landmarks(handle); }
	int getNumCascadeSteps() const { return fdb_sdm_num_cascade_steps(handle); }

	/* SdmLandmarkModel.hpp:156-192: places the shape (in [-0.5, 0.5]^2) into the face box; like the reference it modifies the
	 * caller's buffer and returns a header on it */
	cv::Mat alignRigid(cv::Mat modelShape, cv::Rect faceBox) const {
		if (modelShape.cols != 1)
			throw std::runtime_error("The supplied model shape does not have one column (i.e. it doesn't seem to be a column-vector).");
		const int n = modelShape.rows / 2;
		const float ax = (float)faceBox.width, bx = (float)(0.5 * faceBox.width + faceBox.x); /* the folded MatExpr, see fdb_sdm_align_rigid */
		const float ay = (float)faceBox.height, by = (float)(0.5 * faceBox.height + faceBox.y);
		for (int i = 0; i < n; ++i) {
			volatile float tx = modelShape.at<float>(i, 0) * ax, ty = modelShape.at<float>(n + i, 0) * ay;
			modelShape.at<float>(i, 0) = tx + bx;
			modelShape.at<float>(n + i, 0) = ty + by;
		}
		return modelShape;
	}

	/* SdmLandmarkModel.hpp:199-256 for one face */
	cv::Mat optimize(cv::Mat modelShape, cv::Mat image) {
		std::vector<cv::Mat> shapes(1, modelShape), images(1, image);
		return optimize(shapes, images, std::vector<int>())[0];
	}

	/* the same for many faces at once: shape k lies in images[imageOfShape[k]] (empty: shape k in image k); all images must have
	 * one size. Throws for the first face whose descriptor window leaves the image, as the sequential reference loop would. */
	std::vector<cv::Mat> optimize(const std::vector<cv::Mat>& modelShapes, const std::vector<cv::Mat>& images, const std::vector<int>& imageOfShape) {
		const int L = getNumLandmarks();
		if (images.empty()) throw std::invalid_argument("fdb200: no image");
		const int W = images[0].cols, H = images[0].rows;
		std::vector<unsigned char> frames((size_t)W * H * images.size());
		for (size_t k = 0; k < images.size(); ++k) {
			const cv::Mat& im = images[k];
			if (im.depth() != CV_8U || im.channels() != 1 || im.cols != W || im.rows != H)
				throw std::invalid_argument("fdb200: images must be 8-bit 1-channel and of one size (convert with cv::cvtColor first)");
			for (int y = 0; y < H; ++y) std::copy(im.ptr<unsigned char>(y), im.ptr<unsigned char>(y) + W, frames.begin() + (k * H + y) * (size_t)W);
		}
		std::vector<float> shapes(modelShapes.size() * 2 * (size_t)L);
		for (size_t k = 0; k < modelShapes.size(); ++k) {
			if (modelShapes[k].cols != 1 || modelShapes[k].rows != 2 * L) throw std::runtime_error("fdb200: model shape is not a 2 L x 1 column vector");
			for (int i = 0; i < 2 * L; ++i) shapes[k * 2 * L + i] = modelShapes[k].at<float>(i, 0);
		}
		std::vector<int32_t> status(modelShapes.size(), 0), which(imageOfShape.begin(), imageOfShape.end());
		check(fdb_sdm_optimize_batch(handle, frames.data(), W, W, H, (int32_t)images.size(), which.empty() ? nullptr : which.data(),
				(int64_t)modelShapes.size(), shapes.data(), status.data(), nullptr));
		std::vector<cv::Mat> out;
		for (size_t k = 0; k < modelShapes.size(); ++k) {
			if (status[k] != 0) throw std::runtime_error("fdb200: descriptor window outside the image (cv::Mat::operator()(Rect) asserts in the reference)");
			cv::Mat m(2 * L, 1, CV_32FC1);
			for (int i = 0; i < 2 * L; ++i) m.at<float>(i, 0) = shapes[k * 2 * L + i];
			out.push_back(m);
		}
		return out;
	}

private:
	static void check(int status) {
		if (status == FDB_OK) return;
		const std::string msg = std::string("fdb200: ") + fdb_last_error();
		if (status == FDB_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
		throw std::runtime_error(msg);
	}
	fdb_sdm* handle;
};

} // namespace fdb200
#endif
