/*
 * Minimal stand-in for <opencv2/imgproc/imgproc.hpp>: the pieces of the filter-engine API that
 * imageprocessing::LbpFilter (LbpFilter.hpp / LbpFilter.cpp) needs to compile UNMODIFIED into
 * oracle/_ref.  FilterEngine::apply below hands a cv::BaseFilter the row pointers OpenCV would:
 * src[k] = row (k - anchor.y) of the image, extended by the border on the left/right, starting at
 * column -anchor.x.  Only non-separable 2-D filters on CV_8UC1 with BORDER_REPLICATE are supported.
 * Test infrastructure only. Written for this repository; contains no OpenCV code.
 */
#ifndef FDB_SHIM_OPENCV_IMGPROC_HPP
#define FDB_SHIM_OPENCV_IMGPROC_HPP

#include "opencv2/core/core.hpp"

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4 };

template<class T> class Ptr {
public:
	Ptr() {}
	Ptr(T* p) : p_(p) {}
	template<class U> Ptr(const Ptr<U>& o) : p_(o.shared()) {}
	T* operator->() const { return p_.get(); }
	T& operator*() const { return *p_; }
	bool empty() const { return !p_; }
	std::shared_ptr<T> shared() const { return p_; }
private:
	std::shared_ptr<T> p_;
};

class BaseFilter {
public:
	virtual ~BaseFilter() {}
	virtual void operator()(const uchar** src, uchar* dst, int dststep, int dstcount, int width, int cn) = 0;
	Size ksize;
	Point anchor;
};
class BaseRowFilter { public: virtual ~BaseRowFilter() {} };
class BaseColumnFilter { public: virtual ~BaseColumnFilter() {} };

class FilterEngine {
public:
	FilterEngine(const Ptr<BaseFilter>& filter2D, const Ptr<BaseRowFilter>&, const Ptr<BaseColumnFilter>&,
			int srcType, int dstType, int bufType, int rowBorderType) :
			filter2D_(filter2D), srcType_(srcType), dstType_(dstType), border_(rowBorderType) {
		(void)bufType;
	}
	void apply(const Mat& src, Mat& dst) {
		if (srcType_ != CV_8U || dstType_ != CV_8U || border_ != BORDER_REPLICATE || src.type() != CV_8U)
			throw std::runtime_error("shim FilterEngine: only CV_8UC1 with BORDER_REPLICATE");
		const int kw = filter2D_->ksize.width, kh = filter2D_->ksize.height;
		const int ax = filter2D_->anchor.x, ay = filter2D_->anchor.y;
		const int w = src.cols, h = src.rows, ew = w + kw - 1, eh = h + kh - 1;
		std::vector<uchar> ext((size_t)ew * eh);
		for (int y = 0; y < eh; ++y) {
			const int sy = std::min(std::max(y - ay, 0), h - 1);
			for (int x = 0; x < ew; ++x)
				ext[(size_t)y * ew + x] = src.ptr<uchar>(sy)[std::min(std::max(x - ax, 0), w - 1)];
		}
		std::vector<const uchar*> rows((size_t)eh);
		for (int y = 0; y < eh; ++y) rows[(size_t)y] = &ext[(size_t)y * ew];
		Mat out(h, w, CV_8U);
		(*filter2D_)(rows.data(), out.data, (int)out.step, h, w, 1);
		dst = out;
	}
private:
	Ptr<BaseFilter> filter2D_;
	int srcType_, dstType_, border_;
};

} // namespace cv

#endif
