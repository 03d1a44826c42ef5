/*
 * fd_oracle.c - CPU restatement of the reference's sliding-window hot path (see fd_oracle.h).
 * TEST INFRASTRUCTURE ONLY - never linked into the product library.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math  (no FMA contraction: the reference is plain
 * x86-64 SSE2 code; float/double evaluation order below follows it operation by operation).
 */
#include "fd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---------------------------------------------------------------------------------------------
 * OpenCV-owned primitives (restated from the published OpenCV algorithms; pinned against
 * cv2 4.13.0 by tests/test_oracle_primitives.py)
 * ------------------------------------------------------------------------------------------- */

int fdo_cvround(double v) {
	/* nearbyint in the default rounding mode = round half to even = cvtsd2si */
	return (int)nearbyint(v);
}

static int round_half_even_f(float v) {
	return (int)nearbyintf(v);
}

/* OpenCV resize, INTER_LINEAR, 8UC1: fixed point with INTER_RESIZE_COEF_BITS = 11.
 * Horizontal pass keeps 8+11 bits, vertical pass is the FixedPtCast form
 * ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2.  Same-size = copy; exact 2x2 decimation
 * takes OpenCV's INTER_AREA fast path. */
void fdo_resize_linear_u8(const uint8_t* src, int sw, int sh, int spitch,
		uint8_t* dst, int dw, int dh) {
	if (dw == sw && dh == sh) {
		for (int y = 0; y < sh; ++y)
			memcpy(dst + (size_t)y * dw, src + (size_t)y * spitch, (size_t)sw);
		return;
	}
	double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
	double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
	if (sw == 2 * dw && sh == 2 * dh) { /* iscale == 2 exactly: area-fast path */
		for (int y = 0; y < dh; ++y) {
			const uint8_t* s0 = src + (size_t)(2 * y) * spitch;
			const uint8_t* s1 = s0 + spitch;
			for (int x = 0; x < dw; ++x)
				dst[(size_t)y * dw + x] = (uint8_t)((s0[2 * x] + s0[2 * x + 1] + s1[2 * x] + s1[2 * x + 1] + 2) >> 2);
		}
		return;
	}
	int* xofs = (int*)malloc(sizeof(int) * (size_t)dw);
	short* ialpha = (short*)malloc(sizeof(short) * 2 * (size_t)dw);
	int* yofs = (int*)malloc(sizeof(int) * (size_t)dh);
	short* ibeta = (short*)malloc(sizeof(short) * 2 * (size_t)dh);
	for (int dx = 0; dx < dw; ++dx) {
		float fx = (float)((dx + 0.5) * scale_x - 0.5);
		int sx = (int)floorf(fx);
		fx -= sx;
		if (sx < 0) { fx = 0; sx = 0; }
		if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
		xofs[dx] = sx;
		ialpha[2 * dx] = (short)round_half_even_f((1.f - fx) * 2048.f);
		ialpha[2 * dx + 1] = (short)round_half_even_f(fx * 2048.f);
	}
	for (int dy = 0; dy < dh; ++dy) {
		float fy = (float)((dy + 0.5) * scale_y - 0.5);
		int sy = (int)floorf(fy);
		fy -= sy;
		/* vertical: OpenCV does not zero fy at the borders, it clips the row indices */
		yofs[dy] = sy;
		ibeta[2 * dy] = (short)round_half_even_f((1.f - fy) * 2048.f);
		ibeta[2 * dy + 1] = (short)round_half_even_f(fy * 2048.f);
	}
	int* row0 = (int*)malloc(sizeof(int) * (size_t)dw);
	int* row1 = (int*)malloc(sizeof(int) * (size_t)dw);
	for (int dy = 0; dy < dh; ++dy) {
		int sy0 = yofs[dy];
		int y0 = sy0 < 0 ? 0 : (sy0 >= sh ? sh - 1 : sy0);
		int y1 = sy0 + 1 < 0 ? 0 : (sy0 + 1 >= sh ? sh - 1 : sy0 + 1);
		const uint8_t* s0 = src + (size_t)y0 * spitch;
		const uint8_t* s1 = src + (size_t)y1 * spitch;
		for (int dx = 0; dx < dw; ++dx) {
			int sx = xofs[dx];
			int sx1 = sx + 1 < sw ? sx + 1 : sx; /* coefficient is 0 there */
			int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
			row0[dx] = s0[sx] * a0 + s0[sx1] * a1;
			row1[dx] = s1[sx] * a0 + s1[sx1] * a1;
		}
		int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
		for (int dx = 0; dx < dw; ++dx)
			dst[(size_t)dy * dw + dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
	}
	free(row0); free(row1); free(xofs); free(ialpha); free(yofs); free(ibeta);
}

static int reflect101(int p, int len) {
	if (len == 1)
		return 0;
	while (p < 0 || p >= len) {
		if (p < 0)
			p = -p;
		else
			p = 2 * (len - 1) - p;
	}
	return p;
}

/* OpenCV pyrDown, 8U: separable [1 4 6 4 1]/16 twice, BORDER_REFLECT_101,
 * dst = (sum + 128) >> 8 */
void fdo_pyrdown_u8(const uint8_t* src, int sw, int sh, uint8_t* dst) {
	int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
	static const int k[5] = {1, 4, 6, 4, 1};
	int* rows = (int*)malloc(sizeof(int) * (size_t)dw * 5);
	for (int y = 0; y < dh; ++y) {
		for (int t = 0; t < 5; ++t) {
			int sy = reflect101(2 * y - 2 + t, sh);
			const uint8_t* s = src + (size_t)sy * sw;
			int* r = rows + (size_t)t * dw;
			for (int x = 0; x < dw; ++x) {
				int acc = 0;
				for (int u = 0; u < 5; ++u)
					acc += k[u] * s[reflect101(2 * x - 2 + u, sw)];
				r[x] = acc;
			}
		}
		for (int x = 0; x < dw; ++x) {
			int acc = 0;
			for (int t = 0; t < 5; ++t)
				acc += k[t] * rows[(size_t)t * dw + x];
			dst[(size_t)y * dw + x] = (uint8_t)((acc + 128) >> 8);
		}
	}
	free(rows);
}

/* ---------------------------------------------------------------------------------------------
 * ImagePyramid (ImagePyramid.cpp:79-92, 170-198)
 * ------------------------------------------------------------------------------------------- */

static int layer_cmp(const void* a, const void* b) {
	const fdo_layer* la = (const fdo_layer*)a;
	const fdo_layer* lb = (const fdo_layer*)b;
	return (la->index > lb->index) - (la->index < lb->index);
}

fdo_pyramid* fdo_pyramid_build(const uint8_t* gray, int width, int height, int pitch,
		double incremental_scale_factor, double min_scale_factor, double max_scale_factor) {
	/* ImagePyramid.cpp:84-91: argument checks, octaveLayerCount and the re-derived increment */
	if (incremental_scale_factor <= 0 || incremental_scale_factor >= 1) return NULL;
	if (min_scale_factor <= 0) return NULL;
	if (max_scale_factor > 1) return NULL;
	fdo_pyramid* p = (fdo_pyramid*)calloc(1, sizeof(fdo_pyramid));
	p->octave_layer_count = (int)(size_t)round(log(0.5) / log(incremental_scale_factor));
	p->incremental_scale_factor = pow(0.5, 1. / p->octave_layer_count);
	p->min_scale_factor = min_scale_factor;
	p->max_scale_factor = max_scale_factor;
	p->image_width = width;
	p->image_height = height;
	int cap = 0;
	/* ImagePyramid.cpp:173-194 */
	for (int i = 0; i < p->octave_layer_count; ++i) {
		double scaleFactor = pow(p->incremental_scale_factor, i);
		int w = fdo_cvround(width * scaleFactor), h = fdo_cvround(height * scaleFactor);
		uint8_t* scaled = (uint8_t*)malloc((size_t)w * h);
		fdo_resize_linear_u8(gray, width, height, pitch, scaled, w, h);
		p->n_resize++; p->px_resize += (int64_t)w * h;
		int keep_prev = 0;
		if (scaleFactor <= max_scale_factor && scaleFactor >= min_scale_factor) {
			if (p->n_layers == cap) { cap = cap ? 2 * cap : 16; p->layers = (fdo_layer*)realloc(p->layers, sizeof(fdo_layer) * cap); }
			fdo_layer L = {i, scaleFactor, w, h, scaled};
			p->layers[p->n_layers++] = L;
			keep_prev = 1;
		}
		uint8_t* prev = scaled;
		int pw = w, ph = h;
		scaleFactor *= 0.5;
		for (int j = 1; scaleFactor >= min_scale_factor && pw > 1; ++j, scaleFactor *= 0.5) {
			int nw = (pw + 1) / 2, nh = (ph + 1) / 2;
			uint8_t* down = (uint8_t*)malloc((size_t)nw * nh);
			fdo_pyrdown_u8(prev, pw, ph, down);
			p->n_pyrdown++; p->px_pyrdown += (int64_t)nw * nh;
			int keep = 0;
			if (scaleFactor <= max_scale_factor) {
				if (p->n_layers == cap) { cap = cap ? 2 * cap : 16; p->layers = (fdo_layer*)realloc(p->layers, sizeof(fdo_layer) * cap); }
				fdo_layer L = {i + j * p->octave_layer_count, scaleFactor, nw, nh, down};
				p->layers[p->n_layers++] = L;
				keep = 1;
			}
			if (!keep_prev) free(prev);
			prev = down; pw = nw; ph = nh; keep_prev = keep;
		}
		if (!keep_prev) free(prev);
	}
	if (p->n_layers)
		qsort(p->layers, (size_t)p->n_layers, sizeof(fdo_layer), layer_cmp);
	return p;
}

void fdo_pyramid_free(fdo_pyramid* p) {
	if (!p) return;
	for (int i = 0; i < p->n_layers; ++i) free(p->layers[i].data);
	free(p->layers);
	free(p);
}

/* ---------------------------------------------------------------------------------------------
 * HistEq64Filter::applyTo (HistEq64Filter.cpp:32-125)
 * ------------------------------------------------------------------------------------------- */
void fdo_hq64(const uint8_t* src, int pitch, int w, int h, uint8_t* dst) {
	float stretchFactor = 255.0f / (float)(w * h);      /* :34 */
	float pdf_bins[64];
	for (int i = 0; i < 64; i++) pdf_bins[i] = 0.0f;      /* :48-50 */
	for (int z = 0; z < h; z++) {                         /* :59-67, LUTbin[v] = v>>2 (:14-25) */
		const uint8_t* row = src + (size_t)z * pitch;
		for (int i = 0; i < w; i++)
			pdf_bins[row[i] >> 2] = pdf_bins[row[i] >> 2] + 1;
	}
	for (int i = 0; i < 64; i++)                          /* :70-74 */
		if (pdf_bins[i] != 0) pdf_bins[i] = pdf_bins[i] * stretchFactor;
	float cdf_BINS[64];                                   /* :77-81 sequential float cumsum */
	cdf_BINS[0] = pdf_bins[0];
	for (int i = 1; i < 64; i++) cdf_BINS[i] = cdf_BINS[i - 1] + pdf_bins[i];
	uint8_t eq[64];
	for (int i = 0; i < 64; i++)                          /* :84-87,97: (uchar)floor(LUTeq + 0.5), 0.5 is double */
		eq[i] = (uint8_t)floor((double)cdf_BINS[i] + 0.5);
	for (int z = 0; z < h; z++) {
		const uint8_t* row = src + (size_t)z * pitch;
		for (int i = 0; i < w; i++) dst[(size_t)z * w + i] = eq[row[i] >> 2];
	}
}

/* ---------------------------------------------------------------------------------------------
 * WVM (WvmClassifier.cpp:91-149,165-181,191-346; IImg.cpp:26-65)
 * ------------------------------------------------------------------------------------------- */
struct fdo_wvm {
	int fsx, fsy, num_lin, per_level, num_levels, num_used;
	float basis_param;
	float* lin_thresholds;
	float* hk_weights;   /* packed triangle */
	double* app_rsv_convol;
	float* thresholds;   /* hierarchicalThresholds (file + limitReliabilityFilter) */
	int* cntval;
	int* val_off;        /* [num_lin+1] */
	double* val;
	int* cntrec;         /* per value slot */
	int* rec_off;        /* per value slot (+1) */
	fdb_rect4* rec;
	double logistic_a, logistic_b;
};

static void* dup_mem(const void* p, size_t n) {
	void* q = malloc(n ? n : 1);
	if (n) memcpy(q, p, n);
	return q;
}

fdo_wvm* fdo_wvm_create(const fdb_wvm_desc* d) {
	fdo_wvm* m = (fdo_wvm*)calloc(1, sizeof(fdo_wvm));
	int n = d->num_lin_filters;
	m->fsx = d->filter_size_x; m->fsy = d->filter_size_y;
	m->num_lin = n; m->per_level = d->num_filters_per_level; m->num_levels = d->num_levels;
	/* setNumUsedFilters, WvmClassifier.cpp:151-158 */
	m->num_used = (d->num_used_filters > n || d->num_used_filters == 0) ? n : d->num_used_filters;
	m->basis_param = d->basis_param;
	m->lin_thresholds = (float*)dup_mem(d->lin_thresholds, sizeof(float) * n);
	m->hk_weights = (float*)dup_mem(d->hk_weights, sizeof(float) * ((size_t)n * (n + 1) / 2));
	m->app_rsv_convol = (double*)dup_mem(d->app_rsv_convol, sizeof(double) * n);
	m->thresholds = (float*)malloc(sizeof(float) * n);
	/* setLimitReliabilityFilter, WvmClassifier.cpp:165-181 */
	for (int i = 0; i < n; ++i)
		m->thresholds[i] = d->limit_reliability_filter != 0.0f
				? d->hierarchical_thresholds[i] + d->limit_reliability_filter : d->hierarchical_thresholds[i];
	m->cntval = (int*)dup_mem(d->area_cntval, sizeof(int) * n);
	m->val_off = (int*)malloc(sizeof(int) * (n + 1));
	m->val_off[0] = 0;
	for (int f = 0; f < n; ++f) m->val_off[f + 1] = m->val_off[f] + d->area_cntval[f];
	int nval = m->val_off[n];
	m->val = (double*)dup_mem(d->area_val, sizeof(double) * nval);
	m->cntrec = (int*)dup_mem(d->area_cntrec, sizeof(int) * nval);
	m->rec_off = (int*)malloc(sizeof(int) * (nval + 1));
	m->rec_off[0] = 0;
	for (int f = 0; f < n; ++f)
		for (int v = 0; v < d->area_cntval[f]; ++v) {
			int s = m->val_off[f] + v;
			m->rec_off[s + 1] = m->rec_off[s] + (v == 0 ? 0 : d->area_cntrec[s]);
		}
	m->rec = (fdb_rect4*)dup_mem(d->area_rec, sizeof(fdb_rect4) * (size_t)m->rec_off[nval]);
	m->logistic_a = d->logistic_a; m->logistic_b = d->logistic_b;
	return m;
}

void fdo_wvm_free(fdo_wvm* m) {
	if (!m) return;
	free(m->lin_thresholds); free(m->hk_weights); free(m->app_rsv_convol); free(m->thresholds);
	free(m->cntval); free(m->val_off); free(m->val); free(m->cntrec); free(m->rec_off); free(m->rec);
	free(m);
}

/* IImg::calIImgPatch (IImg.cpp:26-65): float row sums, column-wise float accumulation */
static void cal_iimg_patch(const uint8_t* in, int w, int h, int sqr, float* data) {
	float rowsum = 0;
	for (int c = 0; c < w; c++) {
		rowsum += sqr ? (float)(in[c] * in[c]) : (float)in[c];
		data[c] = rowsum;
	}
	long z = w, zb = 0;
	for (int r = 1; r < h; r++) {
		rowsum = 0;
		for (int c = 0; c < w; c++) {
			rowsum += sqr ? (float)(in[z + c] * in[z + c]) : (float)in[z + c];
			data[z + c] = data[zb + c] + rowsum;
		}
		z += w; zb += w;
	}
}

/* linEvalWvmHisteq64 (WvmClassifier.cpp:191-346) */
static float lin_eval(const fdo_wvm* m, int level, int n, float* hk_kernel_eval, float* u_kernel_eval,
		const float* iimg_x, const float* iimg_xx) {
	const float* this_weight = m->hk_weights + (size_t)level * (level + 1) / 2;
	float res = -m->lin_thresholds[level];                                  /* :201 */
	const int w = m->fsx;
	const int dr = (m->fsy - 1) * w + (w - 1);                              /* :241 */
	double norm_new = iimg_xx[dr];                                          /* :255 */
	float sumv0 = iimg_x[dr];
	double sum_xp = 0.0;
	int base = m->val_off[level];
	int cntval = m->cntval[level];
	for (int v = 1; v < cntval; v++) {                                      /* :277 */
		float sumv = 0;
		const fdb_rect4* rec = m->rec + m->rec_off[base + v];
		int cnt = m->cntrec[base + v];
		for (int r = 0; r < cnt; r++) {
			int ax1 = rec[r].x1 - 1, ax2 = rec[r].x2, ay1 = rec[r].y1;      /* :291 (fx = fy = 0) */
			int ay1w = (ay1 - 1) * w, ay2w = rec[r].y2 * w;                 /* :292 */
			if (ax1 + 1 > 0 && ay1 > 0)                                     /* :293-301 */
				sumv += iimg_x[ay2w + ax2] - iimg_x[ay1w + ax2] - iimg_x[ay2w + ax1] + iimg_x[ay1w + ax1];
			else if (ax1 + 1 > 0)
				sumv += iimg_x[ay2w + ax2] - iimg_x[ay2w + ax1];
			else if (ay1 > 0)
				sumv += iimg_x[ay2w + ax2] - iimg_x[ay1w + ax2];
			else
				sumv += iimg_x[ay2w + ax2];
		}
		sumv0 -= sumv;                                                      /* :308 */
		sum_xp += sumv * m->val[base + v];                                  /* :309 float*double */
	}
	sum_xp += sumv0 * m->val[base + 0];                                     /* :312 */
	sum_xp += u_kernel_eval[n];                                             /* :313 */
	u_kernel_eval[n] = (float)sum_xp;                                       /* :314 truncating store */
	norm_new -= 2 * sum_xp;                                                 /* :316 */
	norm_new += m->app_rsv_convol[level];                                   /* :322 */
	double norm = norm_new;
	hk_kernel_eval[level] = (float)(exp(-m->basis_param * norm));           /* :333 (-float)*double */
	for (int p = 0; p <= level; ++p)                                        /* :340-341 float mul, float add */
		res += this_weight[p] * hk_kernel_eval[p];
	return res;
}

#define FDO_MAX_PATCH 4096
#define FDO_MAX_FILTERS 4096

static void wvm_eval_impl(const fdo_wvm* m, const uint8_t* patch, int* level_out, float* fout_out, float* all_levels) {
	float iimg_x[FDO_MAX_PATCH], iimg_xx[FDO_MAX_PATCH];
	float hk[FDO_MAX_FILTERS];
	float u[FDO_MAX_FILTERS];
	cal_iimg_patch(patch, m->fsx, m->fsy, 0, iimg_x);                       /* :124-127 */
	cal_iimg_patch(patch, m->fsx, m->fsy, 1, iimg_xx);
	for (int n = 0; n < m->per_level; n++) u[n] = 0.0f;                     /* :129-131 */
	int filter_level = -1;
	float fout = 0.0;
	if (all_levels) {
		for (filter_level = 0; filter_level < m->num_used; ++filter_level)
			all_levels[filter_level] = lin_eval(m, filter_level, filter_level % m->per_level, hk, u, iimg_x, iimg_xx);
		return;
	}
	do {                                                                    /* :134-138 */
		filter_level++;
		fout = lin_eval(m, filter_level, filter_level % m->per_level, hk, u, iimg_x, iimg_xx);
	} while (fout >= m->thresholds[filter_level] && filter_level + 1 < m->num_used);
	*level_out = filter_level;
	*fout_out = fout;
}

void fdo_wvm_eval(const fdo_wvm* m, const uint8_t* patch, int* level_out, float* fout_out) {
	wvm_eval_impl(m, patch, level_out, fout_out, NULL);
}

void fdo_wvm_eval_all_levels(const fdo_wvm* m, const uint8_t* patch, float* fout_per_level) {
	int l; float f;
	wvm_eval_impl(m, patch, &l, &f, fout_per_level);
}

int fdo_wvm_classify(const fdo_wvm* m, int level, float fout) {
	/* WvmClassifier.cpp:91-98; fout travels as double, thresholds are float */
	double d = fout;
	return level + 1 == m->num_lin && d >= m->thresholds[level];
}

double fdo_wvm_probability(const fdo_wvm* m, float fout) {
	/* ProbabilisticWvmClassifier.cpp:52 */
	double d = fout;
	return 1.0f / (1.0f + exp(m->logistic_a + m->logistic_b * d));
}

/* ---------------------------------------------------------------------------------------------
 * SVM (SvmClassifier.cpp:44-60; RbfKernel.hpp:32-40,78-108; ProbabilisticSvmClassifier.cpp:54-58)
 * ------------------------------------------------------------------------------------------- */
struct fdo_svm {
	int kernel; double poly_alpha, poly_constant; int poly_degree;
	double gamma;
	int num_sv, dim, sv_type;
	void* sv;
	float* coef;
	float bias, threshold;
	double logistic_a, logistic_b;
};

fdo_svm* fdo_svm_create(const fdb_svm_desc* d) {
	fdo_svm* s = (fdo_svm*)calloc(1, sizeof(fdo_svm));
	s->kernel = d->kernel; s->poly_alpha = d->poly_alpha; s->poly_constant = d->poly_constant; s->poly_degree = d->poly_degree;
	s->gamma = d->gamma; s->num_sv = d->num_sv; s->dim = d->dim; s->sv_type = d->sv_type;
	size_t es = d->sv_type == FDB_SV_U8 ? 1 : 4;
	s->sv = dup_mem(d->support_vectors, es * (size_t)d->num_sv * d->dim);
	s->coef = (float*)dup_mem(d->coefficients, sizeof(float) * d->num_sv);
	s->bias = d->bias; s->threshold = d->threshold;
	s->logistic_a = d->logistic_a; s->logistic_b = d->logistic_b;
	return s;
}

void fdo_svm_free(fdo_svm* s) {
	if (!s) return;
	free(s->sv); free(s->coef); free(s);
}

/* cv::Mat::dot (OpenCV 2.4.3 modules/core/src/matmul.cpp dotProd_): CV_8U sums integer products (exact); CV_32F is
 * dotProd_<float, double>: float64 products, four added left to right, then to the running sum (CV_ENABLE_UNROLLED) */
static double mat_dot(const void* a, const void* b, int n, int sv_type) {
	if (sv_type == FDB_SV_U8) {
		const uint8_t* l = (const uint8_t*)a; const uint8_t* r = (const uint8_t*)b;
		long long sum = 0;
		for (int k = 0; k < n; ++k) sum += (int)l[k] * (int)r[k];
		return (double)sum;
	}
	const float* l = (const float*)a; const float* r = (const float*)b;
	double result = 0;
	int i = 0;
	for (; i <= n - 4; i += 4)
		result += (double)l[i] * r[i] + (double)l[i + 1] * r[i + 1] + (double)l[i + 2] * r[i + 2] + (double)l[i + 3] * r[i + 3];
	for (; i < n; ++i) result += (double)l[i] * r[i];
	return result;
}

/* Kernel::compute of the four kernels (RbfKernel.hpp:32-40,78-108; PolynomialKernel.hpp:38-40,62-70;
 * HistogramIntersectionKernel.hpp:31-39,59-83; LinearKernel.hpp:27-29) */
double fdo_kernel_value(int kernel, double gamma, double alpha, double constant, int degree, const void* x, const void* y, int dim, int sv_type) {
	if (kernel == FDB_KERNEL_RBF) {
		double ssd;
		if (sv_type == FDB_SV_U8) {                                         /* RbfKernel.hpp:78-88 */
			const uint8_t* l = (const uint8_t*)x; const uint8_t* r = (const uint8_t*)y;
			int sum = 0;
			for (int k = 0; k < dim; ++k) { int diff = l[k] - r[k]; sum += diff * diff; }
			ssd = sum;
		} else {                                                            /* RbfKernel.hpp:97-108 */
			const float* l = (const float*)x; const float* r = (const float*)y;
			float sum = 0;
			for (int k = 0; k < dim; ++k) { float diff = l[k] - r[k]; sum += diff * diff; }
			ssd = sum;
		}
		return exp(-gamma * ssd);                                           /* RbfKernel.hpp:39 */
	}
	if (kernel == FDB_KERNEL_HIK) {
		if (sv_type == FDB_SV_U8) {                                         /* HistogramIntersectionKernel.hpp:59-67 */
			const uint8_t* l = (const uint8_t*)x; const uint8_t* r = (const uint8_t*)y;
			int sum = 0;
			for (int k = 0; k < dim; ++k) sum += l[k] < r[k] ? l[k] : r[k];
			return sum;
		}
		const float* l = (const float*)x; const float* r = (const float*)y; /* HistogramIntersectionKernel.hpp:72-80 */
		float sum = 0;
		for (int k = 0; k < dim; ++k) sum += r[k] < l[k] ? r[k] : l[k];     /* std::min(l, r) */
		return sum;
	}
	const double dot = mat_dot(x, y, dim, sv_type);
	if (kernel == FDB_KERNEL_LINEAR) return dot;                            /* LinearKernel.hpp:28 */
	double tmp = alpha * dot + constant, ret = 1.0;                         /* PolynomialKernel.hpp:39,62-70 */
	for (int t = degree; t > 0; t /= 2) {
		if (t % 2 == 1) ret *= tmp;
		tmp = tmp * tmp;
	}
	return ret;
}

double fdo_svm_distance(const fdo_svm* s, const void* x) {
	double distance = -s->bias;                                             /* SvmClassifier.cpp:56 */
	const size_t es = s->sv_type == FDB_SV_U8 ? 1 : 4;
	for (int i = 0; i < s->num_sv; ++i)                                     /* SvmClassifier.cpp:58 */
		distance += s->coef[i] * fdo_kernel_value(s->kernel, s->gamma, s->poly_alpha, s->poly_constant, s->poly_degree, x,
				(const uint8_t*)s->sv + (size_t)i * s->dim * es, s->dim, s->sv_type);
	return distance;
}

int fdo_svm_classify(const fdo_svm* s, double distance) {
	return distance >= s->threshold;
}

double fdo_svm_probability(const fdo_svm* s, double distance) {
	double fABp = s->logistic_a + s->logistic_b * distance;
	return fABp >= 0 ? exp(-fABp) / (1.0 + exp(-fABp)) : 1.0 / (1.0 + exp(fABp));
}

/* ---------------------------------------------------------------------------------------------
 * RVM cascade (RvmClassifier.cpp:66-112; ProbabilisticRvmClassifier.cpp:52-64)
 * ------------------------------------------------------------------------------------------- */
struct fdo_rvm {
	int kernel; double gamma, poly_alpha, poly_constant; int poly_degree;
	int num_filters, use, dim, sv_type;
	void* sv;
	float* coef;  /* packed lower triangle */
	float* thr;
	float bias;
	double logistic_a, logistic_b;
};

fdo_rvm* fdo_rvm_create(const fdb_rvm_desc* d) {
	fdo_rvm* r = (fdo_rvm*)calloc(1, sizeof(fdo_rvm));
	r->kernel = d->kernel; r->gamma = d->gamma; r->poly_alpha = d->poly_alpha; r->poly_constant = d->poly_constant; r->poly_degree = d->poly_degree;
	r->num_filters = d->num_filters; r->dim = d->dim; r->sv_type = d->sv_type;
	r->use = (d->num_filters_to_use <= 0 || d->num_filters_to_use > d->num_filters) ? d->num_filters : d->num_filters_to_use; /* RvmClassifier.cpp:119-126 */
	size_t es = d->sv_type == FDB_SV_U8 ? 1 : 4;
	r->sv = dup_mem(d->support_vectors, es * (size_t)d->num_filters * d->dim);
	r->coef = (float*)dup_mem(d->coefficients, sizeof(float) * (size_t)d->num_filters * (d->num_filters + 1) / 2);
	r->thr = (float*)dup_mem(d->hierarchical_thresholds, sizeof(float) * d->num_filters);
	r->bias = d->bias; r->logistic_a = d->logistic_a; r->logistic_b = d->logistic_b;
	return r;
}

void fdo_rvm_free(fdo_rvm* r) {
	if (!r) return;
	free(r->sv); free(r->coef); free(r->thr); free(r);
}

/* RvmClassifier::computeHyperplaneDistance (RvmClassifier.cpp:75-85) with computeHyperplaneDistanceCached (:94-112) as the
 * reference runs it: the cache vector is created with numFiltersToUse elements, so level 0 takes the full-sum branch
 * (-bias + c[0][0] k_0) and leaves one element; every later level finds size == level and adds c[l][l] k_l to the previous
 * distance. Returns the level, *distance the last distance. */
int fdo_rvm_eval(const fdo_rvm* r, const void* x, double* distance) {
	const size_t es = r->sv_type == FDB_SV_U8 ? 1 : 4;
	int level = -1;
	double d = 0;
	size_t cache_size = (size_t)r->use;              /* vector<double> filterEvalCache(numFiltersToUse) */
	double cache_back = 0;
	do {
		++level;
		if (cache_size == (size_t)level && level != 0) {
			d = cache_back;
			d += r->coef[(size_t)level * (level + 1) / 2 + level] * fdo_kernel_value(r->kernel, r->gamma, r->poly_alpha, r->poly_constant,
					r->poly_degree, x, (const uint8_t*)r->sv + (size_t)level * r->dim * es, r->dim, r->sv_type);
			cache_size++;
		} else {
			d = -r->bias;
			for (int i = 0; i <= level; ++i)
				d += r->coef[(size_t)level * (level + 1) / 2 + i] * fdo_kernel_value(r->kernel, r->gamma, r->poly_alpha, r->poly_constant,
						r->poly_degree, x, (const uint8_t*)r->sv + (size_t)i * r->dim * es, r->dim, r->sv_type);
			cache_size = 1;
		}
		cache_back = d;
	} while (d >= r->thr[level] && level + 1 < r->use);
	*distance = d;
	return level;
}

int fdo_rvm_classify(const fdo_rvm* r, int level, double distance) { /* RvmClassifier.cpp:66-73 */
	return level + 1 == r->use && distance >= r->thr[level];
}

double fdo_rvm_probability(const fdo_rvm* r, double distance) { /* ProbabilisticRvmClassifier.cpp:62 */
	return 1.0f / (1.0f + exp(r->logistic_a + r->logistic_b * distance));
}

/* ---------------------------------------------------------------------------------------------
 * Window enumeration (DirectPyramidFeatureExtractor.cpp:75-123; ImagePyramidLayer.hpp:65-67,98-100)
 * ------------------------------------------------------------------------------------------- */
static void clamp_roi(int W, int H, int* rx, int* ry, int* rw, int* rh) {
	if (*rx == 0 && *ry == 0 && *rw == 0 && *rh == 0) { *rw = W; *rh = H; return; }   /* :84-86 */
	int x = *rx > 0 ? *rx : 0, y = *ry > 0 ? *ry : 0;                                   /* :88-91 */
	int w = (W < *rw + x ? W : *rw + x) - x;
	int h = (H < *rh + y ? H : *rh + y) - y;
	*rx = x; *ry = y; *rw = w; *rh = h;
}

static int count_steps(int begin, int patch, int end, int step) {
	/* number of v = begin, begin+step, ... with v + patch < end */
	int last = end - patch - 1; /* largest admissible v */
	if (last < begin) return 0;
	return (last - begin) / step + 1;
}

int64_t fdo_enumerate(const fdo_pyramid* p, int patch_w, int patch_h, int step_x, int step_y,
		int roi_x, int roi_y, int roi_w, int roi_h, fdb_layer_info* out, int cap) {
	clamp_roi(p->image_width, p->image_height, &roi_x, &roi_y, &roi_w, &roi_h);
	int64_t total = 0;
	for (int i = 0; i < p->n_layers && i < cap; ++i) {
		const fdo_layer* L = &p->layers[i];
		fdb_layer_info* o = &out[i];
		o->index = L->index; o->scale = L->scale; o->width = L->width; o->height = L->height;
		o->orig_patch_width = fdo_cvround(patch_w / L->scale);
		o->orig_patch_height = fdo_cvround(patch_h / L->scale);
		int bx = fdo_cvround(roi_x * L->scale), by = fdo_cvround(roi_y * L->scale);
		int ex = fdo_cvround((roi_x + roi_w) * L->scale), ey = fdo_cvround((roi_y + roi_h) * L->scale);
		o->windows_x = count_steps(bx, patch_w, ex, step_x);
		o->windows_y = count_steps(by, patch_h, ey, step_y);
		if (o->windows_x <= 0 || o->windows_y <= 0) { o->windows_x = o->windows_y = 0; }
		o->first_window = total;
		total += (int64_t)o->windows_x * o->windows_y;
	}
	return total;
}

/* ---------------------------------------------------------------------------------------------
 * Detection stages
 * ------------------------------------------------------------------------------------------- */

/* stable descending insertion/merge sort on ClassifiedPatch::probability
 * (std::sort in the reference: OverlapElimination.cpp:62, FiveStageSlidingWindowDetector.cpp:298,311) */
static void stable_sort_desc(fdb_detection* a, int64_t n) {
	if (n < 2) return;
	fdb_detection* tmp = (fdb_detection*)malloc(sizeof(fdb_detection) * (size_t)n);
	for (int64_t width = 1; width < n; width *= 2) {
		for (int64_t lo = 0; lo < n; lo += 2 * width) {
			int64_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
			int64_t i = lo, j = mid, k = lo;
			while (i < mid && j < hi) tmp[k++] = (a[j].probability > a[i].probability) ? a[j++] : a[i++];
			while (i < mid) tmp[k++] = a[i++];
			while (j < hi) tmp[k++] = a[j++];
		}
		memcpy(a, tmp, sizeof(fdb_detection) * (size_t)n);
	}
	free(tmp);
}

/* OverlapElimination::eliminate (OverlapElimination.cpp:44-105) in place; returns new count */
static int64_t overlap_eliminate(fdb_detection* c, int64_t n, float dist_in, float ratio_in) {
	if (n == 0) return 0;
	float dist = dist_in;
	float ratio = (ratio_in > 0.0f && ratio_in <= 1.0f) ? ratio_in : 0.0f;
	stable_sort_desc(c, n);
	for (int64_t acc = 0; acc < n; ++acc) {
		int64_t k = acc + 1;
		for (int64_t pro = acc + 1; pro < n; ++pro) {
			int wa = c[acc].width, wp = c[pro].width;
			float d;
			if (dist <= 1.0) d = dist * (float)(wa > wp ? wa : wp);
			else d = dist;
			int dx = abs(c[acc].center_x - c[pro].center_x), dy = abs(c[acc].center_y - c[pro].center_y);
			int mn = wa < wp ? wa : wp, mx = wa > wp ? wa : wp;
			if ((dx < d) && (dy < d) && (((float)mn / (float)mx) > ratio))
				continue; /* erased */
			c[k++] = c[pro];
		}
		n = k;
	}
	return n;
}

/* cv::minMaxLoc(src(rows, cols), 0, &maxVal, 0, &maxLoc, mask) on a dense float map: first
 * occurrence of the maximum among selected elements; maxVal = 0, loc = (-1,-1) when nothing
 * is selected. sel(y, x) decides selection. */
typedef struct { const float* map; const uint8_t* mask; int W, H; int by0, by1, bx0, bx1; int use_block; } nms_sel;

static void masked_max(const nms_sel* s, int y0, int y1, int x0, int x1, double* vmax, int* ly, int* lx) {
	int found = 0; float best = 0; int bx = -1, by = -1;
	for (int y = y0; y < y1; ++y)
		for (int x = x0; x < x1; ++x) {
			int sel = 1;
			if (s->mask && !s->mask[(size_t)y * s->W + x]) sel = 0;
			if (s->use_block && y >= s->by0 && y < s->by1 && x >= s->bx0 && x < s->bx1) sel = 0;
			if (!sel) continue;
			float v = s->map[(size_t)y * s->W + x];
			if (!found || v > best) { found = 1; best = v; by = y; bx = x; }
		}
	if (!found) { *vmax = 0; *ly = y0 - 1; *lx = x0 - 1; }
	else { *vmax = best; *ly = by; *lx = bx; }
}

/* nonMaximaSuppression (FiveStageSlidingWindowDetector.cpp:143-184); dst is H*W bytes, zeroed here */
static void non_maxima_suppression(const float* src, int M, int N, int sz, uint8_t* dst, const uint8_t* mask) {
	memset(dst, 0, (size_t)M * N);
	for (int m = 0; m < M; m += sz + 1)
		for (int n = 0; n < N; n += sz + 1) {
			int ic0 = m, ic1 = m + sz + 1 < M ? m + sz + 1 : M;
			int jc0 = n, jc1 = n + sz + 1 < N ? n + sz + 1 : N;
			nms_sel s = {src, mask, N, M, 0, 0, 0, 0, 0};
			double vcmax, vnmax; int cy, cx, ny, nx;
			masked_max(&s, ic0, ic1, jc0, jc1, &vcmax, &cy, &cx);
			int in0 = cy - sz > 0 ? cy - sz : 0, in1 = cy + sz + 1 < M ? cy + sz + 1 : M;
			int jn0 = cx - sz > 0 ? cx - sz : 0, jn1 = cx + sz + 1 < N ? cx + sz + 1 : N;
			/* blockmask: zero over the rows/cols of the block inside the neighbourhood window */
			int iis0 = ic0 - in0, iis1 = (ic0 - in0 + sz + 1 < in1 - in0) ? ic0 - in0 + sz + 1 : in1 - in0;
			int jis0 = jc0 - jn0, jis1 = (jc0 - jn0 + sz + 1 < jn1 - jn0) ? jc0 - jn0 + sz + 1 : jn1 - jn0;
			s.use_block = 1;
			s.by0 = in0 + iis0; s.by1 = in0 + iis1; s.bx0 = jn0 + jis0; s.bx1 = jn0 + jis1;
			masked_max(&s, in0, in1, jn0, jn1, &vnmax, &ny, &nx);
			if (vcmax > vnmax) dst[(size_t)cy * N + cx] = 255;
		}
}

/* grid NMS + bookkeeping of FiveStageSlidingWindowDetector::detect(Mat)
 * (FiveStageSlidingWindowDetector.cpp:276-311) on one frame's SVM-positive patches, in place */
int64_t fdo_five_stage_nms(fdb_detection* det_out, int64_t n, int width, int height) {
	float* map = (float*)calloc((size_t)width * height, sizeof(float));
	uint8_t* mask = (uint8_t*)calloc((size_t)width * height, 1);
	uint8_t* maxima = (uint8_t*)malloc((size_t)width * height);
	for (int64_t i = 0; i < n; ++i) {
		int px = det_out[i].center_x, py = det_out[i].center_y;
		if (px < 0 || py < 0 || px >= width || py >= height) continue; /* UB in the reference */
		if (map[(size_t)py * width + px] < det_out[i].probability)
			map[(size_t)py * width + px] = (float)det_out[i].probability;
	}
	for (size_t i = 0; i < (size_t)width * height; ++i) mask[i] = map[i] > 0.3f ? 255 : 0;
	non_maxima_suppression(map, height, width, 35, maxima, mask);
	int64_t nz = 0;
	for (size_t i = 0; i < (size_t)width * height; ++i) nz += maxima[i] != 0;
	int skip = 0;
	if (nz == 0) {
		non_maxima_suppression(map, height, width, 35, maxima, NULL);
		for (size_t i = 0; i < (size_t)width * height; ++i) nz += maxima[i] != 0;
		if (nz == 0) skip = 1; /* :292-294 returns svmPatchesPositive as is */
	}
	if (!skip) {
		stable_sort_desc(det_out, n);                               /* :298 */
		fdb_detection* sel = (fdb_detection*)malloc(sizeof(fdb_detection) * (size_t)(nz ? nz : 1));
		int64_t m = 0;
		for (int y = 0; y < height; ++y)                            /* findNonZero order: row-major */
			for (int x = 0; x < width; ++x) {
				if (!maxima[(size_t)y * width + x]) continue;
				for (int64_t i = 0; i < n; ++i)
					if (det_out[i].center_x == x && det_out[i].center_y == y) { sel[m++] = det_out[i]; break; }
			}
		memcpy(det_out, sel, sizeof(fdb_detection) * (size_t)m);
		free(sel);
		n = m;
		stable_sort_desc(det_out, n);                               /* :311 */
	}
	free(map); free(mask); free(maxima);
	return n;
}

static void fill_geometry(fdb_detection* d, const fdb_layer_info* L, int frame, int x, int y, int64_t window) {
	/* DirectPyramidFeatureExtractor.cpp:115-118 */
	d->frame = frame; d->layer = L->index; d->x = x; d->y = y;
	d->width = L->orig_patch_width; d->height = L->orig_patch_height;
	d->center_x = fdo_cvround(x / L->scale) + L->orig_patch_width / 2;
	d->center_y = fdo_cvround(y / L->scale) + L->orig_patch_height / 2;
	d->window = window;
}

int64_t fdo_detect_frame(const fdb_detector_desc* desc, const fdo_wvm* wvm, const fdo_svm* svm,
		const uint8_t* frame, int width, int height, int pitch, int frame_index,
		int roi_x, int roi_y, int roi_w, int roi_h, int stage,
		fdb_window_score* dense_out, uint8_t* patches_out,
		fdb_detection* det_out, int64_t det_cap, int64_t counts_out[5], double* timing_out) {
	return fdo_detect_frame_ex(desc, wvm, svm, NULL, frame, width, height, pitch, frame_index, roi_x, roi_y, roi_w, roi_h,
			stage, dense_out, patches_out, NULL, det_out, det_cap, counts_out, timing_out);
}

/* filtered copies of all pyramid layers (layer filters of the feature space), or NULL when it has none */
static uint8_t** filter_layers(const fdo_pyramid* pyr, const fdo_features* f) {
	if (!f || fdo_features_layer_channels(f) == 0) return NULL;
	uint8_t** out = (uint8_t**)calloc((size_t)pyr->n_layers, sizeof(uint8_t*));
	for (int li = 0; li < pyr->n_layers; ++li) {
		const fdo_layer* L = &pyr->layers[li];
		out[li] = (uint8_t*)malloc((size_t)L->width * L->height * fdo_features_layer_channels(f));
		fdo_features_filter_layer(f, L->data, L->width, L->height, out[li]);
	}
	return out;
}

static void free_layers(uint8_t** fl, int n) {
	if (!fl) return;
	for (int i = 0; i < n; ++i) free(fl[i]);
	free(fl);
}

int fdo_extract_features(const fdb_detector_desc* desc, const fdo_features* f, const uint8_t* frame, int width, int height,
		int pitch, const int32_t* layer_x_y, int64_t n, void* out) {
	fdo_pyramid* pyr = fdo_pyramid_build(frame, width, height, pitch, desc->incremental_scale_factor,
			desc->min_scale_factor, desc->max_scale_factor);
	if (!pyr) return -2;
	uint8_t** fl = filter_layers(pyr, f);
	const size_t es = fdo_features_is_float(f) ? 4 : 1;
	int rc = 0;
	for (int64_t i = 0; i < n && rc == 0; ++i) {
		int li = -1;
		for (int k = 0; k < pyr->n_layers; ++k) if (pyr->layers[k].index == layer_x_y[3 * i]) li = k;
		if (li < 0) { rc = -1; break; }
		fdo_features_patch(f, fl ? fl[li] : pyr->layers[li].data, pyr->layers[li].width, layer_x_y[3 * i + 1], layer_x_y[3 * i + 2],
				(uint8_t*)out + (size_t)i * fdo_features_dim(f) * es);
	}
	free_layers(fl, pyr->n_layers);
	fdo_pyramid_free(pyr);
	return rc;
}

int64_t fdo_detect_frame_ex(const fdb_detector_desc* desc, const fdo_wvm* wvm, const fdo_svm* svm, const fdo_features* svm_features,
		const uint8_t* frame, int width, int height, int pitch, int frame_index,
		int roi_x, int roi_y, int roi_w, int roi_h, int stage,
		fdb_window_score* dense_out, uint8_t* patches_out, double* svm_dense_out,
		fdb_detection* det_out, int64_t det_cap, int64_t counts_out[5], double* timing_out) {
	double t0 = now_s();
	int is_roi = !(roi_x == 0 && roi_y == 0 && roi_w == 0 && roi_h == 0);
	fdo_pyramid* pyr = fdo_pyramid_build(frame, width, height, pitch, desc->incremental_scale_factor,
			desc->min_scale_factor, desc->max_scale_factor);
	if (!pyr) return -2;
	double t1 = now_s();
	const int pw = desc->patch_width, ph = desc->patch_height;
	const int sx = desc->step_x > 0 ? desc->step_x : 1, sy = desc->step_y > 0 ? desc->step_y : 1;
	fdb_layer_info* infos = (fdb_layer_info*)calloc((size_t)(pyr->n_layers ? pyr->n_layers : 1), sizeof(fdb_layer_info));
	int64_t total = fdo_enumerate(pyr, pw, ph, sx, sy, roi_x, roi_y, roi_w, roi_h, infos, pyr->n_layers);
	int rx = roi_x, ry = roi_y, rw = roi_w, rh = roi_h;
	clamp_roi(width, height, &rx, &ry, &rw, &rh);
	int64_t n = 0, counts[5] = {total, 0, 0, 0, 0};
	double t_hq = 0, t_wvm = 0;
	uint8_t* patch = (uint8_t*)malloc((size_t)pw * ph);
	int overflow = 0;
	uint8_t** flayers = filter_layers(pyr, svm_features);
	void* fvec = svm_features ? malloc((size_t)fdo_features_dim(svm_features) * 4) : NULL;
	/* stage 1: SlidingWindowDetector::detect() (SlidingWindowDetector.cpp:87-98) */
	for (int li = 0; li < pyr->n_layers; ++li) {
		const fdo_layer* L = &pyr->layers[li];
		const fdb_layer_info* I = &infos[li];
		int bx = fdo_cvround(rx * L->scale), by = fdo_cvround(ry * L->scale);
		for (int iy = 0; iy < I->windows_y; ++iy)
			for (int ix = 0; ix < I->windows_x; ++ix) {
				int x = bx + ix * sx, y = by + iy * sy;
				int64_t w = I->first_window + (int64_t)iy * I->windows_x + ix;
				if (!wvm) {
					/* `single` detector with a psvm classifier (ffpDetectApp.cpp:427-500): feature chain, then
					 * ProbabilisticSvmClassifier::getProbability (ProbabilisticSvmClassifier.cpp:50-58) */
					if (svm_features) fdo_features_patch(svm_features, flayers ? flayers[li] : L->data, L->width, x, y, fvec);
					else fdo_hq64(L->data + (size_t)y * L->width + x, L->width, pw, ph, patch);
					const double dist = fdo_svm_distance(svm, svm_features ? fvec : (void*)patch);
					if (svm_dense_out) svm_dense_out[w] = dist;
					if (fdo_svm_classify(svm, dist)) {
						if (n >= det_cap) { overflow = 1; continue; }
						fdb_detection* d = &det_out[n++];
						memset(d, 0, sizeof(*d));
						fill_geometry(d, I, frame_index, x, y, w);
						d->wvm_level = -1; d->wvm_fout = NAN; d->wvm_probability = NAN;
						d->svm_distance = dist; d->svm_probability = fdo_svm_probability(svm, dist);
						d->probability = d->svm_probability;
						d->positive = 1;
					}
					continue;
				}
				double a = timing_out ? now_s() : 0;
				fdo_hq64(L->data + (size_t)y * L->width + x, L->width, pw, ph, patch);
				double b = timing_out ? now_s() : 0;
				int level; float fout;
				fdo_wvm_eval(wvm, patch, &level, &fout);
				if (timing_out) { double c = now_s(); t_hq += b - a; t_wvm += c - b; }
				if (dense_out) { dense_out[w].fout = fout; dense_out[w].level = level; }
				if (patches_out) memcpy(patches_out + (size_t)w * pw * ph, patch, (size_t)pw * ph);
				if (fdo_wvm_classify(wvm, level, fout)) {
					if (n >= det_cap) { overflow = 1; continue; }
					fdb_detection* d = &det_out[n++];
					memset(d, 0, sizeof(*d));
					fill_geometry(d, I, frame_index, x, y, w);
					d->wvm_level = level; d->wvm_fout = fout;
					d->wvm_probability = fdo_wvm_probability(wvm, fout);
					d->svm_distance = NAN; d->svm_probability = NAN;
					d->probability = d->wvm_probability;
					d->positive = 1;
				}
			}
	}
	free(patch);
	counts[1] = n;
	double t2 = now_s();
	double t_oe = 0, t_svm = 0;
	if (!wvm) { counts[2] = counts[3] = counts[4] = n; stage = 0; }
	if (!overflow && stage >= FDB_STAGE_OE) {
		n = overlap_eliminate(det_out, n, desc->oe_dist, desc->oe_ratio);
		counts[2] = n;
		t_oe = now_s() - t2;
	}
	double t3 = now_s();
	if (!overflow && stage >= FDB_STAGE_SVM && svm) {
		/* FiveStageSlidingWindowDetector.cpp:258-270: classify -> ClassifiedPatch(patch, bool) => probability 0.5 */
		uint8_t* p2 = (uint8_t*)malloc((size_t)pw * ph);
		int64_t k = 0;
		for (int64_t i = 0; i < n; ++i) {
			fdb_detection d = det_out[i];
			const fdo_layer* L = NULL;
			for (int li = 0; li < pyr->n_layers; ++li) if (pyr->layers[li].index == d.layer) L = &pyr->layers[li];
			int lidx = (int)(L - pyr->layers);
			if (svm_features) fdo_features_patch(svm_features, flayers ? flayers[lidx] : L->data, L->width, d.x, d.y, fvec);
			else fdo_hq64(L->data + (size_t)d.y * L->width + d.x, L->width, pw, ph, p2);
			d.svm_distance = fdo_svm_distance(svm, svm_features ? fvec : (void*)p2);
			d.svm_probability = fdo_svm_probability(svm, d.svm_distance);
			d.positive = fdo_svm_classify(svm, d.svm_distance);
			d.probability = 0.5;
			if (d.positive) det_out[k++] = d;
		}
		free(p2);
		n = k;
		counts[3] = n;
		if (stage >= FDB_STAGE_NMS && !is_roi) {
			n = fdo_five_stage_nms(det_out, n, width, height);
			counts[4] = n;
		} else {
			stable_sort_desc(det_out, n);                                   /* ROI variant :366 */
			counts[4] = n;
		}
		t_svm = now_s() - t3;
	}
	if (counts_out) memcpy(counts_out, counts, sizeof(counts));
	if (timing_out) {
		timing_out[0] = t1 - t0; timing_out[1] = t_hq; timing_out[2] = t_wvm;
		timing_out[3] = t_oe; timing_out[4] = t_svm;
	}
	free(infos);
	free(fvec);
	free_layers(flayers, pyr->n_layers);
	fdo_pyramid_free(pyr);
	return overflow ? -1 : n;
}

/* GrayscaleFilter::applyTo (GrayscaleFilter.cpp:18-24), 3-channel branch: cv::cvtColor(image, filtered, CV_BGR2GRAY).
 * OpenCV 2.4.3 (pinned by the reference's CMake; imgproc/src/color.cpp, RGB2Gray<uchar>: yuv_shift 14, B2Y 1868, G2Y 9617,
 * R2Y 4899, rounding offset 1 << 13 folded into the R table). PARITY UNPINNED: OpenCV 2.4.3 is not installable here and
 * cv2 4.13 uses 15-bit coefficients (differs by 1 in ~0.3 % of the pixels); tests/test_oracle_golden.py checks the formula
 * against that bound only. */
void fdo_bgr_to_gray(const uint8_t* bgr, int width, int height, int pitch, uint8_t* gray) {
	for (int y = 0; y < height; ++y) {
		const uint8_t* s = bgr + (size_t)y * pitch;
		uint8_t* d = gray + (size_t)y * width;
		for (int x = 0; x < width; ++x, s += 3) d[x] = (uint8_t)((s[0] * 1868 + s[1] * 9617 + s[2] * 4899 + (1 << 13)) >> 14);
	}
}

/* ---------------------------------------------------------------------------------------------
 * detection::NonMaximumSuppression (libDetection/src/detection/NonMaximumSuppression.cpp:27-112)
 * ------------------------------------------------------------------------------------------- */
typedef struct { float score; int x, y, w, h; long long order; } nms_box;

static int nms_cmp(const void* a, const void* b) { /* ascending score (:36-38); ties by input order (std::sort leaves them open) */
	const nms_box* p = (const nms_box*)a; const nms_box* q = (const nms_box*)b;
	if (p->score < q->score) return -1;
	if (p->score > q->score) return 1;
	return p->order < q->order ? -1 : (p->order > q->order ? 1 : 0);
}

static double nms_overlap(const nms_box* a, const nms_box* b) { /* :60-64 with cv::Rect operator& / area() */
	int x1 = a->x > b->x ? a->x : b->x, y1 = a->y > b->y ? a->y : b->y;
	int x2 = a->x + a->w < b->x + b->w ? a->x + a->w : b->x + b->w, y2 = a->y + a->h < b->y + b->h ? a->y + a->h : b->y + b->h;
	int iw = x2 - x1, ih = y2 - y1;
	if (iw <= 0 || ih <= 0) { iw = 0; ih = 0; }
	double intersectionArea = iw * ih;
	double unionArea = a->w * a->h + b->w * b->h - intersectionArea;
	return intersectionArea / unionArea;
}

int64_t fdo_non_maximum_suppression(float* scores, int32_t* rects, int64_t n, double overlap_threshold, int maximum_type) {
	if (overlap_threshold == 1.0 || n <= 0) return n;                       /* :28-29 */
	nms_box* cand = (nms_box*)malloc(sizeof(nms_box) * (size_t)n);
	nms_box* cluster = (nms_box*)malloc(sizeof(nms_box) * (size_t)n);
	for (int64_t i = 0; i < n; ++i) {
		cand[i].score = scores[i]; cand[i].x = rects[4 * i]; cand[i].y = rects[4 * i + 1]; cand[i].w = rects[4 * i + 2]; cand[i].h = rects[4 * i + 3];
		cand[i].order = i;
	}
	qsort(cand, (size_t)n, sizeof(nms_box), nms_cmp);                       /* sortByScore */
	int64_t m = n, out = 0;
	while (m > 0) {                                                        /* cluster(): :41-46 */
		const nms_box detection = cand[m - 1];
		int64_t keep = 0, nc = 0;
		for (int64_t i = 0; i < m; ++i) {                                   /* stable_partition (:50-52) */
			if (nms_overlap(&detection, &cand[i]) <= overlap_threshold) cand[keep++] = cand[i];
			else cluster[nc++] = cand[i];
		}
		m = keep;
		if (nc == 0) break;                                                 /* an empty box never overlaps itself: the reference would loop forever */
		/* the cluster in reversed (descending score) order (:54): member k is cluster[nc - 1 - k] */
		nms_box r = cluster[nc - 1];                                        /* MAX_SCORE: cluster.front() */
		if (maximum_type == 1 || maximum_type == 2) {
			double weightSum = 0, xSum = 0, ySum = 0, wSum = 0, hSum = 0;
			for (int64_t k = 0; k < nc; ++k) {
				const nms_box* e = &cluster[nc - 1 - k];
				double weight = maximum_type == 1 ? 1.0 : e->score;
				weightSum += weight;
				if (maximum_type == 1) { xSum += e->x; ySum += e->y; wSum += e->w; hSum += e->h; }
				else { xSum += weight * e->x; ySum += weight * e->y; wSum += weight * e->w; hSum += weight * e->h; }
			}
			double den = maximum_type == 1 ? (double)nc : weightSum;
			r.x = (int)round(xSum / den); r.y = (int)round(ySum / den); r.w = (int)round(wSum / den); r.h = (int)round(hSum / den);
		}
		scores[out] = r.score; rects[4 * out] = r.x; rects[4 * out + 1] = r.y; rects[4 * out + 2] = r.w; rects[4 * out + 3] = r.h;
		++out;
	}
	free(cand); free(cluster);
	return out;
}
