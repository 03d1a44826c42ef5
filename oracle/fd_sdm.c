/*
 * fd_sdm.c - CPU restatement of the supervised-descent landmark regressor (BASELINE configs[4]):
 *   SdmLandmarkModelFitting::alignRigid / optimize   libSupervisedDescent/include/superviseddescent/SdmLandmarkModel.hpp:156-256
 *   VlHogDescriptorExtractor::getDescriptors         libSupervisedDescent/include/superviseddescent/DescriptorExtractor.hpp:106-219
 *   vl_hog_new / vl_hog_put_image / vl_hog_extract   libSupervisedDescent/src/superviseddescent/hog.c:174-215,595-727,857-1063
 *                                                    (vendored VLFeat HOG, UoCTTI variant)
 *   SdmLandmarkModel::load                           libSupervisedDescent/src/superviseddescent/SdmLandmarkModel.cpp:130-232
 * TEST INFRASTRUCTURE ONLY - never linked into the product library.
 *
 * Pinned: the HOG against the reference's own hog.c compiled into oracle/_ref (bit-exact,
 * tests/test_oracle_sdm.py); cv::resize on CV_32F and cv::gemm against cv2 4.13 (tests/golden/sdm.npz);
 * the whole fit against the in-repo model detect-landmarks/share/models/SDM_Model_HOG_Zhenhua_22072014.txt.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math.
 */
#include "fd_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.141592653589793
#endif

/* cv::resize(src, dst, Size(dw, dh)) for CV_32FC1, INTER_LINEAR (DescriptorExtractor.hpp:182): float coefficient
 * tables, horizontal pass S[sx]*a0 + S[sx+1]*a1, vertical pass S0*b0 + S1*b1, all float32 */
void fdo_resize_linear_f32(const float* src, int sw, int sh, float* dst, int dw, int dh) {
	if (sw == dw && sh == dh) { memcpy(dst, src, sizeof(float) * (size_t)sw * sh); return; }
	if (sw == 2 * dw && sh == 2 * dh) { /* exact 2x decimation: cv::resize switches INTER_LINEAR to the INTER_AREA fast path */
		for (int dy = 0; dy < dh; ++dy)
			for (int dx = 0; dx < dw; ++dx) {
				const float* S = src + (size_t)(2 * dy) * sw + 2 * dx;
				dst[(size_t)dy * dw + dx] = (((S[0] + S[1]) + S[sw]) + S[sw + 1]) * 0.25f;
			}
		return;
	}
	const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
	int* xofs = (int*)malloc(sizeof(int) * (size_t)(dw + dh));
	int* yofs = xofs + dw;
	float* alpha = (float*)malloc(sizeof(float) * 2 * (size_t)(dw + dh));
	float* beta = alpha + 2 * dw;
	for (int dx = 0; dx < dw; ++dx) {
		float fx = (float)((dx + 0.5) * scale_x - 0.5);
		int sx = (int)floorf(fx);
		fx -= sx;
		if (sx < 0) { fx = 0; sx = 0; }
		if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
		xofs[dx] = sx; alpha[2 * dx] = 1.f - fx; alpha[2 * dx + 1] = fx;
	}
	for (int dy = 0; dy < dh; ++dy) {
		float fy = (float)((dy + 0.5) * scale_y - 0.5);
		int sy = (int)floorf(fy);
		fy -= sy;
		yofs[dy] = sy; beta[2 * dy] = 1.f - fy; beta[2 * dy + 1] = fy;
	}
	float* rows = (float*)malloc(sizeof(float) * 2 * (size_t)dw);
	for (int dy = 0; dy < dh; ++dy) {
		for (int k = 0; k < 2; ++k) {
			int sy = yofs[dy] + k;
			sy = sy < 0 ? 0 : (sy >= sh ? sh - 1 : sy);
			const float* S = src + (size_t)sy * sw;
			for (int dx = 0; dx < dw; ++dx) {
				const int sx = xofs[dx];
				const int sx1 = sx + 1 < sw ? sx + 1 : sx; /* coefficient is 0 there */
				rows[k * dw + dx] = S[sx] * alpha[2 * dx] + S[sx1] * alpha[2 * dx + 1];
			}
		}
		for (int dx = 0; dx < dw; ++dx) dst[(size_t)dy * dw + dx] = rows[dx] * beta[2 * dy] + rows[dw + dx] * beta[2 * dy + 1];
	}
	free(rows); free(alpha); free(xofs);
}

/* vl_hog_new(VlHogVariantUoctti, numOrientations, transposed = false) + vl_hog_put_image(image, width, height, 1, cellSize)
 * + vl_hog_extract (hog.c:174-215, 595-727, 857-1063).  features: [3 * no + 4][hogHeight][hogWidth] floats.
 * Returns hogWidth (== hogHeight for square input) through hw / hh. */
void fdo_vlhog_uoctti(const float* image, int width, int height, int cellSize, int no, float* features, int* hw_out, int* hh_out) {
	const int hogWidth = (width + cellSize / 2) / cellSize, hogHeight = (height + cellSize / 2) / cellSize; /* :541-542 */
	const int hogStride = hogWidth * hogHeight;
	float* ox = (float*)malloc(sizeof(float) * 2 * (size_t)no);
	float* oy = ox + no;
	for (int o = 0; o < no; ++o) { /* :193-202 */
		double angle = o * M_PI / no;
		ox[o] = (float)cos(angle); oy[o] = (float)sin(angle);
	}
	float* hog = (float*)calloc((size_t)hogStride * no * 2, sizeof(float));
	float* hogNorm = (float*)calloc((size_t)hogStride, sizeof(float));
#define at(x, y, k) (hog[(x) + (y) * hogWidth + (k) * hogStride])
	for (int y = 1; y < height - 1; ++y)
		for (int x = 1; x < width - 1; ++x) { /* :617-725 */
			const float* iter = image + y * width + x;
			float gradx = *(iter + 1) - *(iter - 1);
			float grady = *(iter + width) - *(iter - width);
			float grad2 = gradx * gradx + grady * grady;
			if (!(grad2 > 0)) { gradx = 0; grady = 0; grad2 = 0; } /* :640-644: a channel is taken only if grad2_ > grad2 (= 0) */
			float grad = sqrtf(grad2);
			gradx /= (grad > 1e-10 ? grad : 1e-10);
			grady /= (grad > 1e-10 ? grad : 1e-10);
			float w0 = 0, w1 = 0;
			int b0 = -1, b1 = -1;
			for (int k = 0; k < no; ++k) { /* :656-672 */
				float score = gradx * ox[k] + grady * oy[k];
				int bin = k;
				if (score < 0) { score = -score; bin += no; }
				if (score > w0) { b1 = b0; w1 = w0; b0 = bin; w0 = score; }
				else if (score > w1) { b1 = bin; w1 = score; }
			}
			(void)b1; (void)w1;                     /* useBilinearOrientationAssigment is false (:185): weight 1, one bin (:679-682) */
			if (b0 < 0) continue;                   /* :694 */
			float hx = (x + 0.5) / cellSize - 0.5;  /* :697-708 */
			float hy = (y + 0.5) / cellSize - 0.5;
			int binx = (int)floorf(hx), biny = (int)floorf(hy);
			float wx2 = hx - binx, wy2 = hy - biny;
			float wx1 = 1.0 - wx2, wy1 = 1.0 - wy2;
			if (binx >= 0 && biny >= 0) at(binx, biny, b0) += grad * wx1 * wy1;
			if (binx < hogWidth - 1 && biny >= 0) at(binx + 1, biny, b0) += grad * wx2 * wy1;
			if (binx < hogWidth - 1 && biny < hogHeight - 1) at(binx + 1, biny + 1, b0) += grad * wx2 * wy2;
			if (binx >= 0 && biny < hogHeight - 1) at(binx, biny + 1, b0) += grad * wx1 * wy2;
		}
	/* squared L2 norm of the undirected histogram of every cell (:879-894) */
	for (int k = 0; k < no; ++k)
		for (int c = 0; c < hogStride; ++c) {
			float h = hog[c + k * hogStride] + hog[c + (k + no) * hogStride];
			hogNorm[c] += h * h;
		}
#define atNorm(x, y) (hogNorm[(x) + (y) * hogWidth])
	for (int y = 0; y < hogHeight; ++y)
		for (int x = 0; x < hogWidth; ++x) { /* :930-1060 */
			const int xm = x - 1 > 0 ? x - 1 : 0, xp = x + 1 < hogWidth - 1 ? x + 1 : hogWidth - 1;
			const int ym = y - 1 > 0 ? y - 1 : 0, yp = y + 1 < hogHeight - 1 ? y + 1 : hogHeight - 1;
			double norm1 = atNorm(xm, ym), norm2 = atNorm(x, ym), norm3 = atNorm(xp, ym);
			double norm4 = atNorm(xm, y), norm5 = atNorm(x, y), norm6 = atNorm(xp, y);
			double norm7 = atNorm(xm, yp), norm8 = atNorm(x, yp), norm9 = atNorm(xp, yp);
			double factor1 = 1.0 / sqrt(norm1 + norm2 + norm4 + norm5 + 1e-4);
			double factor2 = 1.0 / sqrt(norm2 + norm3 + norm5 + norm6 + 1e-4);
			double factor3 = 1.0 / sqrt(norm4 + norm5 + norm7 + norm8 + 1e-4);
			double factor4 = 1.0 / sqrt(norm5 + norm6 + norm8 + norm9 + 1e-4);
			double t1 = 0, t2 = 0, t3 = 0, t4 = 0;
			const float* iter = hog + x + hogWidth * y;
			float* oiter = features + x + hogWidth * y;
			for (int k = 0; k < no; ++k) {
				double ha = iter[hogStride * k], hb = iter[hogStride * (k + no)];
				double ha1 = factor1 * ha, ha2 = factor2 * ha, ha3 = factor3 * ha, ha4 = factor4 * ha;
				double hb1 = factor1 * hb, hb2 = factor2 * hb, hb3 = factor3 * hb, hb4 = factor4 * hb;
				double hc1 = ha1 + hb1, hc2 = ha2 + hb2, hc3 = ha3 + hb3, hc4 = ha4 + hb4;
#define MIN02(v) ((v) < 0.2 ? (v) : 0.2)
				ha1 = MIN02(ha1); ha2 = MIN02(ha2); ha3 = MIN02(ha3); ha4 = MIN02(ha4);
				hb1 = MIN02(hb1); hb2 = MIN02(hb2); hb3 = MIN02(hb3); hb4 = MIN02(hb4);
				hc1 = MIN02(hc1); hc2 = MIN02(hc2); hc3 = MIN02(hc3); hc4 = MIN02(hc4);
#undef MIN02
				t1 += hc1; t2 += hc2; t3 += hc3; t4 += hc4;
				oiter[0] = (float)(0.5 * (ha1 + ha2 + ha3 + ha4));
				oiter[hogStride * no] = (float)(0.5 * (hb1 + hb2 + hb3 + hb4));
				oiter[2 * hogStride * no] = (float)(0.5 * (hc1 + hc2 + hc3 + hc4));
				oiter += hogStride;
			}
			oiter += 2 * hogStride * no;
			const float tex = 1.0f / sqrtf(18.0f);
			oiter[0] = (float)(tex * t1); oiter[hogStride] = (float)(tex * t2);
			oiter[2 * hogStride] = (float)(tex * t3); oiter[3 * hogStride] = (float)(tex * t4);
		}
#undef at
#undef atNorm
	if (hw_out) *hw_out = hogWidth;
	if (hh_out) *hh_out = hogHeight;
	free(hog); free(hogNorm); free(ox);
}

#ifdef FDO_USE_REF_HOG
int ref_vlhog_uoctti(const float* image, int width, int height, int cell_size, int num_orientations, float* features, int* hw_out, int* hh_out);
#endif

#define SDM_PATCH 30 /* adaptive: 3 cells of 10 px (DescriptorExtractor.hpp:140-144) */
#define SDM_CELL 10
#define SDM_BINS 9
#define SDM_DESC (3 * 3 * (3 * SDM_BINS + 4)) /* 279 */

/* VlHogDescriptorExtractor::getDescriptors, adaptive branch (DescriptorExtractor.hpp:106-219) on a 1-channel u8 image.
 * out: L x 279. Returns 0, or -1 when a region of interest falls outside the (extended) image - the reference's
 * Mat::operator() would throw there. */
int fdo_sdm_descriptors(const uint8_t* image, int cols, int rows, int pitch, const float* pts_xy, int L, int windowSizeHalf, float* out) {
	const int pwh = windowSizeHalf, side = 2 * pwh;
	if (side < 4) return -1; /* vl_hog asserts width > 3 only after the resize; a degenerate window cannot be cropped */
	float* roi = (float*)malloc(sizeof(float) * ((size_t)side * side + SDM_PATCH * SDM_PATCH + (size_t)SDM_DESC));
	float* patch = roi + (size_t)side * side;
	float* feat = patch + SDM_PATCH * SDM_PATCH;
	int rc = 0;
	for (int i = 0; i < L && rc == 0; ++i) {
		const int x = fdo_cvround(pts_xy[2 * i]), y = fdo_cvround(pts_xy[2 * i + 1]); /* :157-158 */
		int rx = x - pwh, ry = y - pwh, bl = 0, bt = 0, br = 0, bb = 0;
		if (x - pwh < 0 || y - pwh < 0 || x + pwh >= cols || y + pwh >= rows) { /* :161-172 */
			bl = (x - pwh) < 0 ? abs(x - pwh) : 0;
			bt = (y - pwh) < 0 ? abs(y - pwh) : 0;
			br = (x + pwh) >= cols ? abs(cols - (x + pwh)) : 0;
			bb = (y + pwh) >= rows ? abs(rows - (y + pwh)) : 0;
			rx = (x - pwh) + bl;
			ry = (y - pwh) + br; /* sic: the reference adds borderRight to the row offset (:169) */
		}
		const int ecols = cols + bl + br, erows = rows + bt + bb;
		if (rx < 0 || ry < 0 || rx + side > ecols || ry + side > erows) { rc = -1; break; }
		for (int r = 0; r < side; ++r)
			for (int c = 0; c < side; ++c) {
				const int sx = rx + c - bl, sy = ry + r - bt; /* extended image = black canvas around the image */
				roi[(size_t)r * side + c] = (sx >= 0 && sy >= 0 && sx < cols && sy < rows) ? (float)image[(size_t)sy * pitch + sx] : 0.f;
			}
		fdo_resize_linear_f32(roi, side, side, patch, SDM_PATCH, SDM_PATCH); /* :181-183 */
		int hw, hh;
#ifdef FDO_USE_REF_HOG /* oracle/_ref build: the reference's own hog.c (through ref_sdm.c) computes the descriptor */
		ref_vlhog_uoctti(patch, SDM_PATCH, SDM_PATCH, SDM_CELL, SDM_BINS, feat, &hw, &hh);
#else
		fdo_vlhog_uoctti(patch, SDM_PATCH, SDM_PATCH, SDM_CELL, SDM_BINS, feat, &hw, &hh);
#endif
		/* :196-204: per dimension the hh x ww plane is transposed and flattened -> index j*ww*hh + x*hh + y */
		float* o = out + (size_t)i * SDM_DESC;
		for (int j = 0; j < 3 * SDM_BINS + 4; ++j)
			for (int xx = 0; xx < hw; ++xx)
				for (int yy = 0; yy < hh; ++yy)
					o[j * hw * hh + xx * hh + yy] = feat[j * hw * hh + yy * hw + xx];
	}
	free(roi);
	return rc;
}

struct fdo_sdm {
	int L, steps, rows; /* rows = L * 279 + 1 */
	float* mean;        /* 2L: all x, then all y */
	float** R;          /* [steps][rows * 2L] */
};

fdo_sdm* fdo_sdm_create(int num_landmarks, int num_steps, const float* mean, const float* const* regressors) {
	fdo_sdm* m = (fdo_sdm*)calloc(1, sizeof(fdo_sdm));
	m->L = num_landmarks; m->steps = num_steps; m->rows = num_landmarks * SDM_DESC + 1;
	m->mean = (float*)malloc(sizeof(float) * 2 * (size_t)m->L);
	memcpy(m->mean, mean, sizeof(float) * 2 * (size_t)m->L);
	m->R = (float**)calloc((size_t)num_steps, sizeof(float*));
	for (int s = 0; s < num_steps; ++s) {
		const size_t n = (size_t)m->rows * 2 * m->L;
		m->R[s] = (float*)malloc(sizeof(float) * n);
		memcpy(m->R[s], regressors[s], sizeof(float) * n);
	}
	return m;
}

void fdo_sdm_free(fdo_sdm* m) {
	if (!m) return;
	for (int s = 0; s < m->steps; ++s) free(m->R[s]);
	free(m->R); free(m->mean); free(m);
}

int fdo_sdm_num_landmarks(const fdo_sdm* m) { return m->L; }
int fdo_sdm_num_steps(const fdo_sdm* m) { return m->steps; }
const float* fdo_sdm_mean(const fdo_sdm* m) { return m->mean; }
const float* fdo_sdm_regressor(const fdo_sdm* m, int step) { return m->R[step]; }

/* SdmLandmarkModel::load (SdmLandmarkModel.cpp:130-232) for models whose steps are all "vlhog-uoctti" with adaptive
 * parameters (empty descriptorParameters), e.g. detect-landmarks/share/models/SDM_Model_HOG_Zhenhua_22072014.txt */
fdo_sdm* fdo_sdm_load(const char* path) {
	FILE* fp = fopen(path, "r");
	if (!fp) return NULL;
	char line[1 << 12], word[64], type[64];
	int L = 0, steps = 0;
	fdo_sdm* m = NULL;
	if (!fgets(line, sizeof line, fp)) goto fail;                          /* description */
	if (fscanf(fp, "%63s %d", word, &L) != 2 || L < 13) goto fail;         /* numLandmarks n (optimize reads landmarks 8..12) */
	for (int i = 0; i < L; ++i) if (fscanf(fp, "%63s", word) != 1) goto fail; /* identifiers */
	m = (fdo_sdm*)calloc(1, sizeof(fdo_sdm));
	m->L = L; m->rows = L * SDM_DESC + 1;
	m->mean = (float*)malloc(sizeof(float) * 2 * (size_t)L);
	for (int i = 0; i < 2 * L; ++i) if (fscanf(fp, "%f", &m->mean[i]) != 1) goto fail;
	if (fscanf(fp, "%63s %d", word, &steps) != 2 || steps < 1) goto fail;  /* numCascadeSteps n */
	m->steps = steps;
	m->R = (float**)calloc((size_t)steps, sizeof(float*));
	for (int s = 0; s < steps; ++s) {
		int idx, rows, cols;
		if (fscanf(fp, " cascadeStep %d rows %d cols %d", &idx, &rows, &cols) != 3) goto fail;
		if (rows != m->rows || cols != 2 * L) goto fail;
		if (fscanf(fp, " descriptorType %63s", type) != 1 || strcmp(type, "vlhog-uoctti") != 0) goto fail;
		if (fscanf(fp, " descriptorPostprocessing %63s", word) != 1) goto fail;
		if (fscanf(fp, " descriptorParameters") != 0) goto fail;
		if (!fgets(line, sizeof line, fp)) goto fail;                      /* rest of the line: must be empty (adaptive) */
		for (const char* c = line; *c; ++c) if (*c != ' ' && *c != '\r' && *c != '\n') goto fail;
		m->R[s] = (float*)malloc(sizeof(float) * (size_t)rows * cols);
		for (size_t k = 0; k < (size_t)rows * cols; ++k) if (fscanf(fp, "%f", &m->R[s][k]) != 1) goto fail;
	}
	fclose(fp);
	return m;
fail:
	fclose(fp);
	fdo_sdm_free(m);
	return NULL;
}

/* SdmLandmarkModelFitting::alignRigid (SdmLandmarkModel.hpp:156-192) with modelShape = mean:
 * `xCoords = (xCoords + 0.5f) * faceBox.width + faceBox.x` is a cv::MatExpr. OpenCV folds it (core/src/matop.cpp, MatOp_AddEx:
 * operator+(Mat, Scalar) -> {alpha 1, s 0.5}; multiply(w) -> {alpha w, s 0.5 w}; add(x) -> {alpha w, s 0.5 w + x}, all in double)
 * and assigns through Mat::convertTo(CV_32F, alpha, s), whose CV_32F kernel (core/src/convert.cpp cvtScale_<float, float, float>)
 * computes saturate_cast<float>(src * (float)alpha + (float)s): ONE multiply and ONE add in float32, not the three operations
 * the source line suggests. 0.5 w + x is a half-integer, exact in float. */
void fdo_sdm_align_rigid(const fdo_sdm* m, int fx, int fy, int fw, int fh, float* shape) {
	const float ax = (float)fw, bx = (float)(0.5 * fw + fx), ay = (float)fh, by = (float)(0.5 * fh + fy);
	for (int i = 0; i < m->L; ++i) {
		shape[i] = m->mean[i] * ax + bx;
		shape[m->L + i] = m->mean[m->L + i] * ay + by;
	}
}

/* one cascade step's geometry (SdmLandmarkModel.hpp:212-229): eye-mouth distance and the HOG window half size */
void fdo_sdm_window(const float* shape, int L, int step, int num_steps, float* distance_out, int* window_half_out) {
	const float a1x = (shape[8] + shape[9]) / 2.0f, a1y = (shape[8 + L] + shape[9 + L]) / 2.0f;
	const float a2x = (shape[11] + shape[12]) / 2.0f, a2y = (shape[11 + L] + shape[12 + L]) / 2.0f;
	const float dx = a1x - a2x, dy = a1y - a2y;
	const float d = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy); /* cv::norm(Vec2f): double accumulation */
	float windowSize = d / 2.0f;
	float windowSizeHalf = windowSize / 2;
	windowSizeHalf = (float)round(windowSizeHalf * (1 / (1 + exp((double)((step + 1) - num_steps)))));
	const int NUM_CELL = 3;
	*window_half_out = (int)windowSizeHalf + NUM_CELL - ((int)windowSizeHalf % NUM_CELL);
	*distance_out = d;
}

/* SdmLandmarkModelFitting::optimize (SdmLandmarkModel.hpp:199-256) on a 1-channel u8 image; shape: 2L floats in/out.
 * features_out: NULL or [steps][L * 279] (the feature rows, for stage-wise comparison). Returns 0 or -1 (see above). */
int fdo_sdm_optimize(const fdo_sdm* m, const uint8_t* image, int cols, int rows, int pitch, float* shape, float* features_out) {
	const int L = m->L, K = L * SDM_DESC, N = 2 * L;
	float* pts = (float*)malloc(sizeof(float) * ((size_t)N + K));
	float* feat = pts + N;
	double* acc = (double*)malloc(sizeof(double) * (size_t)N);
	int rc = 0;
	for (int step = 0; step < m->steps && rc == 0; ++step) {
		for (int i = 0; i < L; ++i) { pts[2 * i] = shape[i]; pts[2 * i + 1] = shape[i + L]; }
		float d; int wsh;
		fdo_sdm_window(shape, L, step, m->steps, &d, &wsh);
		rc = fdo_sdm_descriptors(image, cols, rows, pitch, pts, L, wsh, feat);
		if (rc) break;
		if (features_out) memcpy(features_out + (size_t)step * K, feat, sizeof(float) * (size_t)K);
		/* deltaShape = features * R[0:rows-1] + R[rows-1] (:241): cv::gemm on CV_32F accumulates in double */
		const float* R = m->R[step];
		for (int j = 0; j < N; ++j) acc[j] = 0;
		for (int k = 0; k < K; ++k) {
			const double a = feat[k];
			const float* Rk = R + (size_t)k * N;
			for (int j = 0; j < N; ++j) acc[j] += a * (double)Rk[j];
		}
		for (int j = 0; j < N; ++j) {
			const float delta = (float)(acc[j] + (double)R[(size_t)K * N + j]);
			shape[j] = shape[j] + delta * d; /* :243: modelShape + deltaShape.t() * dynamicFaceSizeDistance */
		}
	}
	free(acc); free(pts);
	return rc;
}
