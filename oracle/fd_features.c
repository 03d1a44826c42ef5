/*
 * fd_features.c - CPU restatement of the reference's feature spaces (see fd_oracle.h):
 * the patch-filter chains and pyramid layer filters that turn a window into a feature vector.
 * TEST INFRASTRUCTURE ONLY - never linked into the product library.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math.  float/double evaluation order follows the
 * reference operation by operation.  OpenCV-owned primitives (Sobel, equalizeHist, dft, norm,
 * convertTo) are restated from the published OpenCV algorithms and pinned against cv2 4.13.0
 * (tests/golden/cv2_features.npz); the in-repo filters are pinned against the reference's own
 * sources compiled into oracle/_ref (tests/test_oracle_features.py).
 */
#include "fd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif

static int reflect101(int p, int len) {
	if (len == 1) return 0;
	while (p < 0 || p >= len) {
		if (p < 0) p = -p;
		else p = 2 * len - 2 - p;
	}
	return p;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* cv::saturate_cast<uchar>(double/float) = clamp(cvRound(v)) */
static uint8_t sat_u8_d(double v) { return (uint8_t)clampi((int)nearbyint(v), 0, 255); }
static uint8_t sat_u8_f(float v) { return (uint8_t)clampi((int)nearbyintf(v), 0, 255); }

/* ---------------------------------------------------------------------------------------------
 * GradientFilter::applyTo (GradientFilter.cpp:38-59) for CV_8U, blurKernelSize 0:
 * cv::Sobel(image, gx, CV_8U, 1, 0, k, scale, 127), cv::Sobel(image, gy, CV_8U, 0, 1, k, scale, 127),
 * scale = 1/2 (k = 1) or 1/8 (k = 3) (:20-27).  OpenCV: separable float filter, BORDER_REFLECT_101,
 * result saturate_cast<uchar>(sum * scale + delta); every intermediate is exact in float32.
 * out: interleaved {gx, gy} (cv::merge, :58).
 * ------------------------------------------------------------------------------------------- */
void fdo_gradient_u8(const uint8_t* src, int w, int h, int pitch, int ksize, uint8_t* out_xy) {
	for (int y = 0; y < h; ++y) {
		const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * pitch;
		const uint8_t* r1 = src + (size_t)y * pitch;
		const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * pitch;
		for (int x = 0; x < w; ++x) {
			const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
			float gx, gy;
			if (ksize == 1) {
				gx = (float)(r1[xp] - r1[xm]) * 0.5f + 127.f;
				gy = (float)(r2[x] - r0[x]) * 0.5f + 127.f;
			} else { /* ksize 3: deriv [-1 0 1], smooth [1 2 1] * 1/8 */
				const int dx = (r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]);
				const int dy = (r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]);
				gx = (float)dx * 0.125f + 127.f;
				gy = (float)dy * 0.125f + 127.f;
			}
			out_xy[2 * ((size_t)y * w + x)] = sat_u8_f(gx);
			out_xy[2 * ((size_t)y * w + x) + 1] = sat_u8_f(gy);
		}
	}
}

/* ---------------------------------------------------------------------------------------------
 * GradientBinningFilter::GradientBinningFilter (GradientBinningFilter.cpp:18-59): look-up tables
 * indexed by gx | gy << 8 (little-endian union of {uchar x, y}, :21-26,76).
 * one_bin: [65536][2] {bin, weight}; two_bin: [65536][4] {bin0, w0, bin1, w1}
 * ------------------------------------------------------------------------------------------- */
void fdo_gradient_bin_luts(int bins, int signed_gradients, uint8_t* one_bin, uint8_t* two_bin) {
	for (int x = 0; x < 256; ++x) {
		double gradientX = ((double)x - 127) / 255;
		for (int y = 0; y < 256; ++y) {
			double gradientY = ((double)y - 127) / 255;
			double direction = atan2(gradientY, gradientX);
			double magnitude = sqrt(gradientX * gradientX + gradientY * gradientY);
			double bin;
			if (signed_gradients) {
				direction += M_PI;
				bin = direction * (unsigned)bins / (2 * M_PI);
			} else {
				if (direction < 0) direction += M_PI;
				bin = direction * (unsigned)bins / M_PI;
			}
			const int idx = x | (y << 8);
			if (one_bin) {
				one_bin[2 * idx] = (uint8_t)((uint8_t)round(bin) % bins);
				one_bin[2 * idx + 1] = sat_u8_d(255 * magnitude);
			}
			if (two_bin) {
				two_bin[4 * idx] = (uint8_t)((uint8_t)floor(bin) % bins);
				two_bin[4 * idx + 2] = (uint8_t)((uint8_t)ceil(bin) % bins);
				two_bin[4 * idx + 3] = sat_u8_d(255 * magnitude * (bin - floor(bin)));
				two_bin[4 * idx + 1] = sat_u8_d(255 * magnitude - two_bin[4 * idx + 3]);
			}
		}
	}
}

/* ---------------------------------------------------------------------------------------------
 * LbpFilter::applyTo (LbpFilter.cpp:56-85) for CV_8U with the 3x3 operators of LbpFilter.hpp:88-178,
 * FilterEngine border BORDER_REPLICATE (:67-69); uniform mapping LbpFilter.cpp:20-44
 * ------------------------------------------------------------------------------------------- */
static int lbp_is_uniform(uint8_t code) {
	int transitions = 0;
	int previousBit = (code >> 7) & 1;
	for (int pos = 0; pos < 8; ++pos) {
		int currentBit = (code >> pos) & 1;
		if (previousBit != currentBit) { transitions++; previousBit = currentBit; }
	}
	return transitions <= 2;
}

void fdo_lbp_map(int lbp_type, uint8_t map[256]) {
	for (int i = 0; i < 256; ++i) map[i] = (uint8_t)i;
	if (lbp_type == FDB_LBP8_UNIFORM) {
		int emptyIndex = 1;
		for (int i = 0; i < 256; ++i) map[i] = lbp_is_uniform((uint8_t)i) ? (uint8_t)emptyIndex++ : 0;
	}
}

int fdo_lbp_bins(int lbp_type) {
	return lbp_type == FDB_LBP8 ? 256 : (lbp_type == FDB_LBP8_UNIFORM ? 59 : 16);
}

void fdo_lbp_u8(const uint8_t* src, int w, int h, int pitch, int lbp_type, uint8_t* out) {
	uint8_t map[256];
	fdo_lbp_map(lbp_type, map);
	for (int y = 0; y < h; ++y) {
		const uint8_t* prev = src + (size_t)clampi(y - 1, 0, h - 1) * pitch;
		const uint8_t* curr = src + (size_t)y * pitch;
		const uint8_t* next = src + (size_t)clampi(y + 1, 0, h - 1) * pitch;
		for (int x = 0; x < w; ++x) {
			const int xm = clampi(x - 1, 0, w - 1), xp = clampi(x + 1, 0, w - 1);
			const int c = curr[x];
			int code = 0;
			if (lbp_type == FDB_LBP8 || lbp_type == FDB_LBP8_UNIFORM) {
				code |= (prev[xm] > c) << 7; code |= (prev[x] > c) << 6; code |= (prev[xp] > c) << 5;
				code |= (curr[xp] > c) << 4; code |= (next[xp] > c) << 3; code |= (next[x] > c) << 2;
				code |= (next[xm] > c) << 1; code |= (curr[xm] > c) << 0;
			} else if (lbp_type == FDB_LBP4) {
				code |= (prev[x] > c) << 3; code |= (curr[xp] > c) << 2; code |= (next[x] > c) << 1; code |= (curr[xm] > c) << 0;
			} else {
				code |= (prev[xm] > c) << 3; code |= (prev[xp] > c) << 2; code |= (next[xp] > c) << 1; code |= (next[xm] > c) << 0;
			}
			out[(size_t)y * w + x] = map[code];
		}
	}
}

/* ---------------------------------------------------------------------------------------------
 * cv::equalizeHist for CV_8UC1 (HistogramEqualizationFilter.cpp:17-20): published OpenCV algorithm
 * ------------------------------------------------------------------------------------------- */
void fdo_equalize_hist_u8(const uint8_t* src, int pitch, int w, int h, uint8_t* dst) {
	int hist[256];
	memset(hist, 0, sizeof(hist));
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) hist[src[(size_t)y * pitch + x]]++;
	const int total = w * h;
	int i = 0;
	while (!hist[i]) ++i;
	uint8_t lut[256];
	if (hist[i] == total) {
		for (int k = 0; k < 256; ++k) lut[k] = (uint8_t)i;
	} else {
		const float scale = 255.f / (float)(total - hist[i]);
		int sum = 0;
		memset(lut, 0, sizeof(lut));
		for (lut[i++] = 0; i < 256; ++i) {
			sum += hist[i];
			lut[i] = sat_u8_f((float)sum * scale);
		}
	}
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) dst[(size_t)y * w + x] = lut[src[(size_t)y * pitch + x]];
}

/* ---------------------------------------------------------------------------------------------
 * WhiteningFilter (WhiteningFilter.cpp:20-81)
 * ------------------------------------------------------------------------------------------- */
/* getFilter (:62-81): float32 arithmetic; pow(rho / cutoff, 4) and exp are evaluated in double
 * (pow(float, int) promotes to double in C++11) and multiplied into the float entry */
void fdo_whitening_filter(int w, int h, float alpha, float cutoff, float* filter) {
	const float nyquistFrequency = 0.5f;
	for (int row = 0; row < h; ++row)
		for (int col = 0; col < w; ++col) {
			int shiftedRow = (row + h / 2) % h;
			int shiftedCol = (col + w / 2) % w;
			float fx = -nyquistFrequency + shiftedCol * (2 * nyquistFrequency) / (w - 1);
			float fy = -nyquistFrequency + shiftedRow * (2 * nyquistFrequency) / (h - 1);
			float rho = sqrtf(fx * fx + fy * fy);
			float v = powf(rho, alpha);
			if (cutoff > 0) v = (float)((double)v * exp(-pow((double)(rho / cutoff), 4)));
			filter[row * w + col] = v;
		}
}

/* applyTo (:20-56) for CV_8U.  cv::dft is OpenCV's float32 transform; here the transforms are evaluated in
 * double (direct O(n^2) sums with exact-argument twiddles), the closest value to what any float32 FFT
 * approximates.  Semantics pinned against cv2 4.13: forward DFT_SCALE | DFT_COMPLEX_OUTPUT = full spectrum / (w*h);
 * DFT_INVERSE | DFT_REAL_OUTPUT on the complex spectrum = Hermitian inverse that reads columns 0..w/2 only
 * (complex inverse along columns, then complex-to-real along rows ignoring the imaginary parts of columns 0
 * and w/2).  Output: saturate_cast<uchar>(v + 127) (convertTo(CV_8U, 1, 127), :49). */
void fdo_whitening_u8(const uint8_t* src, int pitch, int w, int h, const float* filter, uint8_t* dst, float* real_out) {
	const int hw = w / 2 + 1;
	double* cw = (double*)malloc(sizeof(double) * 2 * (size_t)(w + h));
	double* sw = cw + w; double* ch = sw + w; double* sh = ch + h;
	for (int k = 0; k < w; ++k) { cw[k] = cos(2 * M_PI * k / w); sw[k] = sin(2 * M_PI * k / w); }
	for (int k = 0; k < h; ++k) { ch[k] = cos(2 * M_PI * k / h); sh[k] = sin(2 * M_PI * k / h); }
	double* a_re = (double*)calloc((size_t)h * hw * 4, sizeof(double));
	double* a_im = a_re + (size_t)h * hw;
	double* b_re = a_im + (size_t)h * hw;
	double* b_im = b_re + (size_t)h * hw;
	/* forward rows: A[y][u] = sum_x f[y][x] e^{-2 pi i u x / w}, u = 0..w/2 */
	for (int y = 0; y < h; ++y)
		for (int u = 0; u < hw; ++u) {
			double re = 0, im = 0;
			for (int x = 0; x < w; ++x) {
				const int k = (u * x) % w;
				const double f = (double)(float)src[(size_t)y * pitch + x]; /* convertTo(CV_32F), :25 */
				re += f * cw[k]; im -= f * sw[k];
			}
			a_re[y * hw + u] = re; a_im[y * hw + u] = im;
		}
	/* forward columns + DFT_SCALE, then the whitening filter (:38-44, float spectrum * float filter) */
	const double inv_n = 1.0 / ((double)w * h);
	for (int v = 0; v < h; ++v)
		for (int u = 0; u < hw; ++u) {
			double re = 0, im = 0;
			for (int y = 0; y < h; ++y) {
				const int k = (v * y) % h;
				re += a_re[y * hw + u] * ch[k] + a_im[y * hw + u] * sh[k];
				im += a_im[y * hw + u] * ch[k] - a_re[y * hw + u] * sh[k];
			}
			const float fre = (float)(re * inv_n), fim = (float)(im * inv_n); /* the spectrum is stored as float32 */
			b_re[v * hw + u] = (double)(fre * filter[v * w + u]);
			b_im[v * hw + u] = (double)(fim * filter[v * w + u]);
		}
	/* inverse columns: C[y][u] = sum_v B[v][u] e^{+2 pi i v y / h} */
	for (int y = 0; y < h; ++y)
		for (int u = 0; u < hw; ++u) {
			double re = 0, im = 0;
			for (int v = 0; v < h; ++v) {
				const int k = (v * y) % h;
				re += b_re[v * hw + u] * ch[k] - b_im[v * hw + u] * sh[k];
				im += b_im[v * hw + u] * ch[k] + b_re[v * hw + u] * sh[k];
			}
			a_re[y * hw + u] = re; a_im[y * hw + u] = im;
		}
	/* inverse rows, complex-to-real: f[y][x] = C0 + (-1)^x C_{w/2} + 2 sum_{0<u<w/2} Re(C_u e^{+2 pi i u x / w}) */
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) {
			double s = a_re[y * hw];
			for (int u = 1; u < hw; ++u) {
				const int k = (u * x) % w;
				if (2 * u == w) s += a_re[y * hw + u] * cw[k];
				else s += 2 * (a_re[y * hw + u] * cw[k] - a_im[y * hw + u] * sw[k]);
			}
			const float f = (float)s;
			if (real_out) real_out[y * w + x] = f;
			dst[y * w + x] = sat_u8_f(f + 127.f);
		}
	free(a_re); free(cw);
}

/* ---------------------------------------------------------------------------------------------
 * HistogramFilter (HistogramFilter.cpp): cell histograms, normalisation
 * ------------------------------------------------------------------------------------------- */
typedef struct cache_entry { int index1, index2; float weight1, weight2; } cache_entry;

/* createCache (:199-220) */
static void create_cache(cache_entry* cache, int size, int count) {
	for (int matIndex = 0; matIndex < size; ++matIndex) {
		cache_entry entry;
		double realIndex = (double)count * ((double)matIndex + 0.5) / (double)size - 0.5;
		entry.index1 = (int)floor(realIndex);
		entry.index2 = entry.index1 + 1;
		entry.weight2 = (float)(realIndex - entry.index1);
		entry.weight1 = 1.f - entry.weight2;
		if (entry.index1 < 0) { entry.index1 = entry.index2; entry.weight1 = 0; }
		else if (entry.index2 >= count) { entry.index2 = entry.index1; entry.weight2 = 0; }
		cache[matIndex] = entry;
	}
}

/* createCellHistograms (:23-197); image: rows x cols x channels u8 with row pitch (bytes) */
static void create_cell_histograms(const uint8_t* image, int pitch, int rows, int cols, int channels,
		float* histograms, int binCount, int rowCount, int columnCount, int interpolate) {
	memset(histograms, 0, sizeof(float) * (size_t)rowCount * columnCount * binCount);
	const float factor = 1.f / 255.f;
	if (interpolate) {
		cache_entry* rowCache = (cache_entry*)malloc(sizeof(cache_entry) * (size_t)(rows + cols));
		cache_entry* colCache = rowCache + rows;
		create_cache(rowCache, rows, rowCount);
		create_cache(colCache, cols, columnCount);
		for (int imageRow = 0; imageRow < rows; ++imageRow) {
			const uint8_t* rowValues = image + (size_t)imageRow * pitch;
			const int rowIndex0 = rowCache[imageRow].index1, rowIndex1 = rowCache[imageRow].index2;
			const float rowWeight1 = rowCache[imageRow].weight2, rowWeight0 = rowCache[imageRow].weight1;
			for (int imageCol = 0; imageCol < cols; ++imageCol) {
				const int colIndex0 = colCache[imageCol].index1, colIndex1 = colCache[imageCol].index2;
				const float colWeight1 = colCache[imageCol].weight2, colWeight0 = colCache[imageCol].weight1;
				const uint8_t* px = rowValues + (size_t)imageCol * channels;
				const int nb = channels == 4 ? 2 : 1;
				for (int corner = 0; corner < 4; ++corner) {
					const int ri = corner < 2 ? rowIndex0 : rowIndex1, ci = (corner & 1) ? colIndex1 : colIndex0;
					const float rw = corner < 2 ? rowWeight0 : rowWeight1, cwt = (corner & 1) ? colWeight1 : colWeight0;
					const int ok = (corner < 2 ? rowIndex0 >= 0 : rowIndex1 < rowCount)
							&& ((corner & 1) ? colIndex1 < columnCount : colIndex0 >= 0);
					if (!ok) continue;
					float* hv = histograms + ((size_t)ri * columnCount + ci) * binCount;
					if (channels == 1) hv[px[0]] += rw * cwt;
					else
						for (int b = 0; b < nb; ++b) {
							const float weight = factor * px[2 * b + 1];
							hv[px[2 * b]] += weight * rw * cwt;
						}
				}
			}
		}
		free(rowCache);
	} else {
		float* hv = histograms;
		for (int cellRow = 0; cellRow < rowCount; ++cellRow)
			for (int cellCol = 0; cellCol < columnCount; ++cellCol) {
				const int startRow = (cellRow * rows) / rowCount, startCol = (cellCol * cols) / columnCount;
				const int endRow = ((cellRow + 1) * rows) / rowCount, endCol = ((cellCol + 1) * cols) / columnCount;
				for (int imageRow = startRow; imageRow < endRow; ++imageRow) {
					const uint8_t* rowValues = image + (size_t)imageRow * pitch;
					for (int imageCol = startCol; imageCol < endCol; ++imageCol) {
						const uint8_t* px = rowValues + (size_t)imageCol * channels;
						if (channels == 1) hv[px[0]]++;
						else if (channels == 2) hv[px[0]] += factor * px[1];
						else { hv[px[0]] += factor * px[1]; hv[px[2]] += factor * px[3]; }
					}
				}
				hv += binCount;
			}
	}
}

/* cv::norm(NORM_L2 / NORM_L1) of a float array: double accumulation */
static double norm_l2(const float* v, int n) {
	double s = 0;
	for (int i = 0; i < n; ++i) s += (double)v[i] * (double)v[i];
	return sqrt(s);
}
static double norm_l1(const float* v, int n) {
	double s = 0;
	for (int i = 0; i < n; ++i) s += fabs((double)v[i]);
	return s;
}
/* Mat / double -> convertTo(type, 1/s): float32 multiply by (float)(1/s) */
static void div_scalar(float* v, int n, double s) {
	const float scale = (float)(1.0 / s);
	for (int i = 0; i < n; ++i) v[i] = v[i] * scale;
}
static const float hist_eps = 1e-4f; /* HistogramFilter::eps (:19) */

/* HistogramFilter::normalize (:222-252) */
static void normalize_hist(float* v, int n, int normalization) {
	switch (normalization) {
	case FDB_NORM_L2NORM: { float norm = (float)norm_l2(v, n); div_scalar(v, n, (double)(norm + hist_eps)); break; }
	case FDB_NORM_L2HYS: {
		float norm = (float)norm_l2(v, n); div_scalar(v, n, (double)(norm + hist_eps));
		for (int i = 0; i < n; ++i) v[i] = v[i] < (float)0.2 ? v[i] : (float)0.2;
		norm = (float)norm_l2(v, n); div_scalar(v, n, (double)(norm + hist_eps));
		break; }
	case FDB_NORM_L1NORM: { float norm = (float)norm_l1(v, n); div_scalar(v, n, (double)(norm + hist_eps)); break; }
	case FDB_NORM_L1SQRT: {
		float norm = (float)norm_l1(v, n); div_scalar(v, n, (double)(norm + hist_eps));
		for (int i = 0; i < n; ++i) v[i] = sqrtf(v[i]);
		break; }
	default: break;
	}
}

/* ---------------------------------------------------------------------------------------------
 * prepared feature space
 * ------------------------------------------------------------------------------------------- */
struct fdo_features {
	fdb_feature_desc d;
	int pw, ph, dim, is_float, layer_channels, bins;
	int cell_rows, cell_cols, use_hog_filter;
	uint8_t* lut;       /* binning LUT of the layer filter */
	float* whi_filter;
};

int fdo_features_dim(const fdo_features* f) { return f->dim; }
int fdo_features_is_float(const fdo_features* f) { return f->is_float; }
int fdo_features_layer_channels(const fdo_features* f) { return f->layer_channels; }

fdo_features* fdo_features_create(const fdb_feature_desc* d, int patch_w, int patch_h) {
	fdo_features* f = (fdo_features*)calloc(1, sizeof(fdo_features));
	f->d = *d; f->pw = patch_w; f->ph = patch_h;
	switch (d->kind) {
	case FDB_FEATURE_HQ64: case FDB_FEATURE_GRAY: case FDB_FEATURE_HISTEQ:
		f->dim = patch_w * patch_h; f->is_float = 0; f->layer_channels = 0; break;
	case FDB_FEATURE_WHI:
		f->dim = patch_w * patch_h; f->is_float = 1; f->layer_channels = 0;
		f->whi_filter = (float*)malloc(sizeof(float) * (size_t)f->dim);
		fdo_whitening_filter(patch_w, patch_h, d->whi_alpha, d->whi_cutoff, f->whi_filter);
		break;
	case FDB_FEATURE_HOG: case FDB_FEATURE_EHOG: case FDB_FEATURE_LBP: {
		f->is_float = 1;
		/* cvRound(rows / cellHeight) (SpatialHistogramFilter.cpp:57-58, HogFilter.cpp:59-60, ExtendedHogFilter.cpp:55-56) */
		f->cell_rows = fdo_cvround((double)patch_h / (double)d->cell_size);
		f->cell_cols = fdo_cvround((double)patch_w / (double)d->cell_size);
		if (d->kind == FDB_FEATURE_LBP) {
			f->layer_channels = 1; f->bins = fdo_lbp_bins(d->lbp_type);
		} else {
			f->bins = d->bins;
			f->layer_channels = d->interpolate_bins ? 4 : 2;
			f->lut = (uint8_t*)malloc((size_t)65536 * f->layer_channels);
			fdo_gradient_bin_luts(d->bins, d->signed_gradients, d->interpolate_bins ? NULL : f->lut, d->interpolate_bins ? f->lut : NULL);
		}
		const int half = f->bins / 2;
		if (d->kind == FDB_FEATURE_EHOG) {
			f->dim = f->cell_rows * f->cell_cols * (f->bins + (d->signed_and_unsigned ? half : 0) + 4);
		} else {
			/* AdaptiveTracking::createHogFilter (AdaptiveTracking.cpp:241-253) */
			f->use_hog_filter = d->kind == FDB_FEATURE_HOG && !(d->block_size == 1 && !d->signed_and_unsigned);
			const int br = f->cell_rows - d->block_size + 1, bc = f->cell_cols - d->block_size + 1;
			if (f->use_hog_filter)
				f->dim = br * bc * d->block_size * d->block_size * (f->bins + (d->signed_and_unsigned ? half : 0));
			else if (d->block_size == 1)
				f->dim = f->cell_rows * f->cell_cols * f->bins;
			else
				f->dim = br * bc * (d->concatenate ? d->block_size * d->block_size * f->bins : f->bins);
		}
		break; }
	default: free(f); return NULL;
	}
	return f;
}

void fdo_features_free(fdo_features* f) {
	if (!f) return;
	free(f->lut); free(f->whi_filter); free(f);
}

/* ImagePyramid layer filters (ImagePyramid.cpp:182,191: layerFilter->applyTo(scaledImage)):
 * gray layer -> filtered layer with fdo_features_layer_channels() bytes per pixel */
void fdo_features_filter_layer(const fdo_features* f, const uint8_t* gray, int w, int h, uint8_t* out) {
	if (f->layer_channels == 0) { memcpy(out, gray, (size_t)w * h); return; }
	if (f->d.kind == FDB_FEATURE_LBP) { fdo_lbp_u8(gray, w, h, w, f->d.lbp_type, out); return; }
	uint8_t* g = (uint8_t*)malloc((size_t)w * h * 2);
	fdo_gradient_u8(gray, w, h, w, f->d.gradient_kernel, g);
	const int ch = f->layer_channels;
	for (size_t i = 0; i < (size_t)w * h; ++i) { /* GradientBinningFilter::applyTo (:66-92) */
		const int idx = g[2 * i] | (g[2 * i + 1] << 8);
		memcpy(out + i * ch, f->lut + (size_t)idx * ch, (size_t)ch);
	}
	free(g);
}

/* SpatialHistogramFilter::createBlockHistograms (SpatialHistogramFilter.cpp:69-94) */
static void spatial_blocks(const fdo_features* f, const float* cells, float* out) {
	const int bs = f->d.block_size, bins = f->bins, concat = f->d.concatenate;
	const int size = concat ? bs * bs * bins : bins;
	const int br = f->cell_rows - bs + 1, bc = f->cell_cols - bs + 1;
	memset(out, 0, sizeof(float) * (size_t)br * bc * size);
	float* values = out;
	for (int blockRow = 0; blockRow < br; ++blockRow)
		for (int blockCol = 0; blockCol < bc; ++blockCol) {
			float* block = values;
			for (int cellRow = blockRow; cellRow < blockRow + bs; ++cellRow)
				for (int cellCol = blockCol; cellCol < blockCol + bs; ++cellCol) {
					const float* c = cells + ((size_t)cellRow * f->cell_cols + cellCol) * bins;
					for (int bin = 0; bin < bins; ++bin) values[bin] += c[bin];
					if (concat) values += bins;
				}
			if (!concat) values += bins;
			normalize_hist(block, size, f->d.normalization);
		}
}

/* HogFilter::computeCellEnergies + createBlockHistograms (HogFilter.cpp:69-122) */
static void hog_blocks(const fdo_features* f, const float* cells, float* out) {
	const int bs = f->d.block_size, bins = f->bins, half = bins / 2, su = f->d.signed_and_unsigned;
	const int ncell = f->cell_rows * f->cell_cols;
	float* energies = (float*)malloc(sizeof(float) * (size_t)ncell);
	for (int cellIndex = 0; cellIndex < ncell; ++cellIndex) {
		const float* c = cells + (size_t)cellIndex * bins;
		float energy = 0;
		if (su) {
			for (int b = 0; b < half; ++b) { float u = c[b] + c[half + b]; energy += u * u; }
		} else {
			for (int b = 0; b < bins; ++b) energy += c[b] * c[b];
		}
		energies[cellIndex] = energy;
	}
	const int br = f->cell_rows - bs + 1, bc = f->cell_cols - bs + 1;
	float* values = out;
	for (int blockRow = 0; blockRow < br; ++blockRow)
		for (int blockCol = 0; blockCol < bc; ++blockCol) {
			float energy = 0;
			for (int cellRow = blockRow; cellRow < blockRow + bs; ++cellRow)
				for (int cellCol = blockCol; cellCol < blockCol + bs; ++cellCol)
					energy += energies[cellRow * f->cell_cols + cellCol];
			float normalizer = 1.f / sqrtf(energy + hist_eps);
			for (int cellRow = blockRow; cellRow < blockRow + bs; ++cellRow)
				for (int cellCol = blockCol; cellCol < blockCol + bs; ++cellCol) {
					const float* c = cells + ((size_t)cellRow * f->cell_cols + cellCol) * bins;
					for (int b = 0; b < bins; ++b) values[b] = normalizer * c[b];
					values += bins;
					if (su) {
						for (int b = 0; b < half; ++b) values[b] = normalizer * (c[b] + c[half + b]);
						values += half;
					}
				}
		}
	free(energies);
}

/* ExtendedHogFilter::createDescriptors (ExtendedHogFilter.cpp:65-209) */
static void ehog_descriptors(const fdo_features* f, const float* cells, float* out) {
	const int bins = f->bins, half = bins / 2, su = f->d.signed_and_unsigned;
	const int R = f->cell_rows, Cc = f->cell_cols;
	const float alpha = f->d.ehog_alpha;
	float* energies = (float*)calloc((size_t)R * Cc, sizeof(float));
	for (int cellIndex = 0; cellIndex < R * Cc; ++cellIndex) {
		const float* c = cells + (size_t)cellIndex * bins;
		if (su) for (int b = 0; b < half; ++b) { float sum = c[b] + c[b + half]; energies[cellIndex] += sum * sum; }
		else for (int b = 0; b < bins; ++b) energies[cellIndex] += c[b] * c[b];
	}
	float* values = out;
	for (int cellRow = 0; cellRow < R; ++cellRow)
		for (int cellCol = 0; cellCol < Cc; ++cellCol) {
			const float* c = cells + ((size_t)cellRow * Cc + cellCol) * bins;
			const int r1 = cellRow, r0 = cellRow - 1 > 0 ? cellRow - 1 : 0, r2 = cellRow + 1 < R - 1 ? cellRow + 1 : R - 1;
			const int c1 = cellCol, c0 = cellCol - 1 > 0 ? cellCol - 1 : 0, c2 = cellCol + 1 < Cc - 1 ? cellCol + 1 : Cc - 1;
#define E(r, cc) energies[(r) * Cc + (cc)]
			float n1 = 1.f / sqrtf(E(r0, c0) + E(r0, c1) + E(r1, c0) + E(r1, c1) + hist_eps);
			float n2 = 1.f / sqrtf(E(r0, c1) + E(r0, c2) + E(r1, c1) + E(r1, c2) + hist_eps);
			float n3 = 1.f / sqrtf(E(r1, c0) + E(r1, c1) + E(r2, c0) + E(r2, c1) + hist_eps);
			float n4 = 1.f / sqrtf(E(r1, c1) + E(r1, c2) + E(r2, c1) + E(r2, c2) + hist_eps);
#undef E
			float t1 = 0, t2 = 0, t3 = 0, t4 = 0;
			for (int b = 0; b < bins; ++b) {
				float h1 = fminf(alpha, c[b] * n1), h2 = fminf(alpha, c[b] * n2);
				float h3 = fminf(alpha, c[b] * n3), h4 = fminf(alpha, c[b] * n4);
				values[b] = (float)(0.5 * (h1 + h2 + h3 + h4));
				t1 += h1; t2 += h2; t3 += h3; t4 += h4;
			}
			values += bins;
			if (su) {
				for (int b = 0; b < half; ++b) {
					float sum = c[b] + c[b + half];
					float h1 = fminf(alpha, sum * n1), h2 = fminf(alpha, sum * n2);
					float h3 = fminf(alpha, sum * n3), h4 = fminf(alpha, sum * n4);
					values[b] = (float)(0.5 * (h1 + h2 + h3 + h4));
				}
				values += half;
			}
			values[0] = (float)(0.2357 * t1); values[1] = (float)(0.2357 * t2);
			values[2] = (float)(0.2357 * t3); values[3] = (float)(0.2357 * t4);
			values += 4;
		}
	free(energies);
}

/* The patch filter chain on the window whose top-left corner is (x, y) of a filtered layer
 * (layer_w pixels per row, fdo_features_layer_channels() bytes per pixel; raw gray when 0).
 * out: dim elements, u8 or float32. */
void fdo_features_patch(const fdo_features* f, const uint8_t* layer, int layer_w, int x, int y, void* out) {
	const int pw = f->pw, ph = f->ph;
	const int ch = f->layer_channels ? f->layer_channels : 1;
	const int pitch = layer_w * ch;
	const uint8_t* roi = layer + (size_t)y * pitch + (size_t)x * ch;
	switch (f->d.kind) {
	case FDB_FEATURE_HQ64: fdo_hq64(roi, pitch, pw, ph, (uint8_t*)out); return;
	case FDB_FEATURE_GRAY:
		for (int r = 0; r < ph; ++r) memcpy((uint8_t*)out + (size_t)r * pw, roi + (size_t)r * pitch, (size_t)pw);
		return;
	case FDB_FEATURE_HISTEQ: fdo_equalize_hist_u8(roi, pitch, pw, ph, (uint8_t*)out); return;
	case FDB_FEATURE_WHI: {
		const int n = pw * ph;
		uint8_t* a = (uint8_t*)malloc((size_t)n * 2);
		uint8_t* b = a + n;
		fdo_whitening_u8(roi, pitch, pw, ph, f->whi_filter, a, NULL);
		fdo_equalize_hist_u8(a, pw, pw, ph, b);
		float* v = (float*)out;
		/* ConversionFilter(CV_32F, 1/127.5, -1) (ConversionFilter.cpp:16-19): float32 src * (float)alpha + (float)beta */
		const float alpha = (float)(1.0 / 127.5), beta = (float)-1.0;
		for (int i = 0; i < n; ++i) v[i] = (float)b[i] * alpha + beta;
		/* UnitNormFilter::normalize (UnitNormFilter.cpp:35-38): double norm, eps 1e-4f */
		const double norm = norm_l2(v, n);
		div_scalar(v, n, norm + hist_eps);
		free(a);
		return; }
	default: break;
	}
	float* cells = (float*)malloc(sizeof(float) * (size_t)f->cell_rows * f->cell_cols * f->bins);
	create_cell_histograms(roi, pitch, ph, pw, f->layer_channels, cells, f->bins, f->cell_rows, f->cell_cols,
			f->d.interpolate_cells);
	if (f->d.kind == FDB_FEATURE_EHOG) ehog_descriptors(f, cells, (float*)out);
	else if (f->use_hog_filter) hog_blocks(f, cells, (float*)out);
	else if (f->d.block_size == 1) { /* SpatialHistogramFilter.cpp:59-61 */
		memcpy(out, cells, sizeof(float) * (size_t)f->dim);
		normalize_hist((float*)out, f->dim, f->d.normalization);
	} else spatial_blocks(f, cells, (float*)out);
	free(cells);
}
