/*
 * Minimal stand-in for <opencv2/core/core.hpp>, just enough of cv::Mat / Rect / Point / Size
 * for the reference's classification, HistEq64 and overlap-elimination sources to compile
 * UNMODIFIED into oracle/_ref (OpenCV itself is not installed in this image).
 * Test infrastructure only. Written for this repository; contains no OpenCV code.
 */
#ifndef FDB_SHIM_OPENCV_CORE_HPP
#define FDB_SHIM_OPENCV_CORE_HPP

#include <algorithm>
#include <functional>
#include <cmath>
#include <math.h> /* as the real opencv2/core/types_c.h does: with it, unqualified sqrt/pow/exp on float arguments
                   * resolve to the float overloads (libstdc++'s <math.h> and MSVC both export them globally) */
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH_MASK 7
#define CV_MAKETYPE(depth, cn) (((depth) & CV_MAT_DEPTH_MASK) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAKE_TYPE CV_MAKETYPE
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_8UC(n) CV_MAKETYPE(CV_8U, (n))
#define CV_8UC2 CV_MAKETYPE(CV_8U, 2)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_32FC(n) CV_MAKETYPE(CV_32F, (n))
#define CV_PI 3.1415926535897932384626433832795
#define CV_SCHARR -1

inline int cvRound(double v) { return (int)std::nearbyint(v); }
inline int cvFloor(double v) { return (int)std::floor(v); }

namespace cv {

template<class T> struct Point_ {
	T x, y;
	Point_() : x(0), y(0) {}
	Point_(T x, T y) : x(x), y(y) {}
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;

template<class T> struct Size_ {
	T width, height;
	Size_() : width(0), height(0) {}
	Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template<class T> struct Rect_ {
	T x, y, width, height;
	Rect_() : x(0), y(0), width(0), height(0) {}
	Rect_(T x, T y, T w, T h) : x(x), y(y), width(w), height(h) {}
	Point_<T> tl() const { return Point_<T>(x, y); }
	Point_<T> br() const { return Point_<T>(x + width, y + height); }
	Size_<T> size() const { return Size_<T>(width, height); }
	T area() const { return width * height; }
};
/* intersection (OpenCV: operator&= clamps to an empty Rect_() when the rectangles do not overlap) */
template<class T> inline Rect_<T> operator&(const Rect_<T>& a, const Rect_<T>& b) {
	const T x1 = a.x > b.x ? a.x : b.x, y1 = a.y > b.y ? a.y : b.y;
	const T w = (a.x + a.width < b.x + b.width ? a.x + a.width : b.x + b.width) - x1;
	const T h = (a.y + a.height < b.y + b.height ? a.y + a.height : b.y + b.height) - y1;
	return (w <= 0 || h <= 0) ? Rect_<T>() : Rect_<T>(x1, y1, w, h);
}
typedef Rect_<int> Rect;
#ifndef CV_32FC2
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#endif

template<class T, int N> struct Vec {
	T val[N];
	Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
	T& operator[](int i) { return val[i]; }
	const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 2> Vec2b;
typedef Vec<uchar, 4> Vec4b;
typedef Vec<float, 2> Vec2f;

/* saturate_cast<uchar>: cvRound (half to even) then clamp, as OpenCV defines it */
template<class T> inline T saturate_cast(double v);
template<> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : (i > 255 ? 255 : i)); }
template<class T> inline T saturate_cast(float v);
template<> inline uchar saturate_cast<uchar>(float v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : (i > 255 ? 255 : i)); }
template<class T> inline T saturate_cast(int v);
template<> inline uchar saturate_cast<uchar>(int v) { return (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

class Mat;
/* the two matrix expressions the histogram filters use: m / s (evaluated as convertTo(type, 1/s), i.e. a
 * multiplication by (float)(1/s)) and min(m, s); assigned INTO the destination's existing buffer when it has
 * the right size and type, as cv::MatExpr does */
struct MatExpr {
	const Mat* a;
	int op; /* 0: scale, 1: min */
	double s;
};

class Mat {
public:
	enum { CONTINUOUS_FLAG = 1 << 14, MAGIC = 0x42FF0000 };

	Mat() : flags(MAGIC), dims(0), rows(0), cols(0), data(nullptr), step(0) {}
	Mat(int r, int c, int type) : flags(MAGIC), dims(0), rows(0), cols(0), data(nullptr), step(0) { create(r, c, type); }
	Mat(int r, int c, int type, void* ext, size_t stepBytes = 0) :
			flags(MAGIC | (type & 0xFFF)), dims(2), rows(r), cols(c), data((uchar*)ext), step(0) {
		size_t minstep = (size_t)c * elemSize();
		step = stepBytes ? stepBytes : minstep;
		if (step == minstep || r == 1) flags |= CONTINUOUS_FLAG;
	}
	Mat(const Mat& m, const Rect& roi) :
			flags(m.flags & ~CONTINUOUS_FLAG), dims(2), rows(roi.height), cols(roi.width),
			data(m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize()), step(m.step), buffer(m.buffer) {
		if (roi.width == m.cols && (m.flags & CONTINUOUS_FLAG)) flags |= CONTINUOUS_FLAG;
		if (rows == 1) flags |= CONTINUOUS_FLAG;
	}

	void create(int r, int c, int type) {
		type &= 0xFFF;
		if (data && rows == r && cols == c && this->type() == type) return;
		flags = MAGIC | type | CONTINUOUS_FLAG;
		dims = 2; rows = r; cols = c;
		step = (size_t)c * elemSize();
		buffer = std::make_shared<std::vector<uchar>>((size_t)r * step + 16);
		data = buffer->data();
	}
	Mat clone() const {
		Mat m;
		if (empty()) return m;
		m.create(rows, cols, type());
		for (int y = 0; y < rows; ++y)
			std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
		return m;
	}
	void copyTo(Mat& dst) const { dst = clone(); }

	template<class T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
	template<class T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
	uchar* ptr(int y = 0) { return data + (size_t)y * step; }
	const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
	template<class T> T* ptr(int y, int x) { return (T*)(data + (size_t)y * step + (size_t)x * elemSize()); }
	template<class T> const T* ptr(int y, int x) const { return (const T*)(data + (size_t)y * step + (size_t)x * elemSize()); }
	static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, (size_t)r * m.step); return m; }
	inline Mat& operator=(const MatExpr& e);
	template<class T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
	template<class T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }

	bool isContinuous() const { return (flags & CONTINUOUS_FLAG) != 0; }
	bool empty() const { return data == nullptr || rows * cols == 0; }
	size_t total() const { return (size_t)rows * cols; }
	int type() const { return flags & 0xFFF; }
	int depth() const { return flags & CV_MAT_DEPTH_MASK; }
	int channels() const { return ((flags & 0xFFF) >> CV_CN_SHIFT) + 1; }
	size_t elemSize1() const {
		static const size_t s[8] = {1, 1, 2, 2, 4, 4, 8, 2};
		return s[depth()];
	}
	size_t elemSize() const { return elemSize1() * channels(); }
	Size size() const { return Size(cols, rows); }

	double dot(const Mat& m) const {
		double r = 0;
		size_t n = total() * channels();
		if (depth() == CV_8U) { const uchar* a = ptr<uchar>(); const uchar* b = m.ptr<uchar>(); for (size_t i = 0; i < n; ++i) r += (double)a[i] * b[i]; }
		else if (depth() == CV_32F) { /* OpenCV 2.4.3 dotProd_<float, double> with CV_ENABLE_UNROLLED: four products at a time */
			const float* a = ptr<float>(); const float* b = m.ptr<float>(); size_t i = 0;
			for (; i + 4 <= n; i += 4) r += (double)a[i] * b[i] + (double)a[i + 1] * b[i + 1] + (double)a[i + 2] * b[i + 2] + (double)a[i + 3] * b[i + 3];
			for (; i < n; ++i) r += (double)a[i] * b[i];
		}
		else if (depth() == CV_64F) { const double* a = ptr<double>(); const double* b = m.ptr<double>(); for (size_t i = 0; i < n; ++i) r += a[i] * b[i]; }
		else throw std::runtime_error("shim Mat::dot: unsupported depth");
		return r;
	}

	int flags;
	int dims;
	int rows, cols;
	uchar* data;
	size_t step;

private:
	std::shared_ptr<std::vector<uchar>> buffer;
};

inline Mat& Mat::operator=(const MatExpr& e) {
	const Mat src = *e.a; /* header copy keeps the source buffer alive if *this is the source */
	if (src.depth() != CV_32F) throw std::runtime_error("shim MatExpr: only CV_32F");
	create(src.rows, src.cols, src.type());
	const int n = src.cols * src.channels();
	const float fs = e.op == 0 ? (float)(1.0 / e.s) : (float)e.s;
	for (int y = 0; y < src.rows; ++y) {
		const float* a = src.ptr<float>(y);
		float* d = ptr<float>(y);
		if (e.op == 0) for (int i = 0; i < n; ++i) d[i] = a[i] * fs;
		else for (int i = 0; i < n; ++i) d[i] = a[i] < fs ? a[i] : fs;
	}
	return *this;
}

inline MatExpr operator/(const Mat& a, double s) { MatExpr e; e.a = &a; e.op = 0; e.s = s; return e; }
inline MatExpr min(const Mat& a, double s) { MatExpr e; e.a = &a; e.op = 1; e.s = s; return e; }

/* cv::norm for CV_32F arrays: double accumulation */
inline double norm(const Mat& m, int normType) {
	if (m.depth() != CV_32F) throw std::runtime_error("shim norm: only CV_32F");
	const int n = m.cols * m.channels();
	double s = 0;
	for (int y = 0; y < m.rows; ++y) {
		const float* a = m.ptr<float>(y);
		if (normType == NORM_L2) for (int i = 0; i < n; ++i) s += (double)a[i] * (double)a[i];
		else if (normType == NORM_L1) for (int i = 0; i < n; ++i) s += std::fabs((double)a[i]);
		else throw std::runtime_error("shim norm: unsupported type");
	}
	return normType == NORM_L2 ? std::sqrt(s) : s;
}

inline void sqrt(const Mat& src, Mat& dst) {
	const Mat a = src;
	dst.create(a.rows, a.cols, a.type());
	const int n = a.cols * a.channels();
	for (int y = 0; y < a.rows; ++y) for (int i = 0; i < n; ++i) dst.ptr<float>(y)[i] = std::sqrt(a.ptr<float>(y)[i]);
}

/* only referenced by code paths the oracle never runs (GradientOrientationFilter::extractMagnitude) */
inline void mixChannels(const std::vector<Mat>&, std::vector<Mat>, const std::vector<int>&) { throw std::runtime_error("shim: cv::mixChannels is not available"); }

} // namespace cv

#endif
