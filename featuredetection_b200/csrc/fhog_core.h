/*
 * fhog_core.h - the arithmetic of the FHOG layer filter (the feature map of
 * detection::AggregatedFeaturesDetector), written once as host/device functions:
 *   - fhog.cu wraps them into CUDA kernels (thread = (cell, bin) for the histograms, thread = cell for the descriptors);
 *   - tests/test_fhog_host_emulation.py compiles the same functions with g++ and checks them bit for bit against the
 *     oracle (which is pinned against the reference's own FhogFilter / FhogAggregationFilter sources).
 *   - aggdet.cu (the batched detector path) stages the per-pixel LUT entries of a tile once and replays them per cell
 *     (fhog_cell_histogram): also emulated on the host by that test, tile by tile as the kernel does.
 *
 * Reference: FhogFilter.cpp:20-122, FhogFilter.hpp:112-208 (signed histograms), FhogAggregationFilter.cpp:43-150
 * (energies, normalisers, descriptor). float32 additions are not associative: a histogram bin of a cell receives its
 * contributions in raster order of the pixels (FhogFilter.hpp:118-123), which fhog_signed_bin() replays.
 */
#ifndef FDB_FHOG_CORE_H_
#define FDB_FHOG_CORE_H_

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FHOG_HD __host__ __device__ __forceinline__
#else
#define FHOG_HD static inline
#endif

/* products and sums that must not be contracted into FMAs (the reference is plain SSE2 code) */
#if defined(__CUDA_ARCH__)
#define FHOG_MUL(a, b) __fmul_rn((a), (b))
#define FHOG_ADD(a, b) __fadd_rn((a), (b))
#define FHOG_DMUL(a, b) __dmul_rn((a), (b))
#else
#define FHOG_MUL(a, b) ((a) * (b))   /* host builds use -ffp-contract=off */
#define FHOG_ADD(a, b) ((a) + (b))
#define FHOG_DMUL(a, b) ((a) * (b))
#endif

/* FhogFilter::LutEntry (FhogFilter.hpp:73-81): bin indices and weights of one (dx, dy) gradient code */
struct FhogLutEntry {
	int32_t index1, index2;
	float weight1, weight2;
	float magnitude;
};

/* FhogFilter::Coefficients of a pixel row / column (computeInterpolationCoefficents, FhogFilter.cpp:69-92) */
struct FhogCoef {
	int32_t index1, index2;
	float weight1, weight2;
};

FHOG_HD FhogCoef fhog_pixel_coef(int pixel, int cell, int size_cells, int interpolate_cells) {
	FhogCoef c;
	if (interpolate_cells) {
		const float real = (pixel + 0.5f) / cell - 0.5f;
		int i1 = (int)floorf(real), i2 = i1 + 1;
		float w2 = real - i1, w1 = i2 - real;
		if (i1 < 0) { i1 = i2; w1 = 0; }
		else if (i2 >= size_cells) { i2 = i1; w2 = 0; }
		c.index1 = i1; c.index2 = i2; c.weight1 = w1; c.weight2 = w2;
	} else {
		c.index1 = pixel / cell; c.index2 = -1; c.weight1 = 1; c.weight2 = 0;
	}
	return c;
}

/* FhogFilter::getBinCoefficients<true / false> (FhogFilter.hpp:127-168): the LUT entry of pixel (r, c) */
FHOG_HD const FhogLutEntry* fhog_pixel_entry(const FhogLutEntry* lut, const uint8_t* image, int pitch, int rows, int cols,
		int channels, int r, int c) {
	const int pr = r - 1 < 0 ? 0 : r - 1, nr = r + 1 > rows - 1 ? rows - 1 : r + 1;
	const int pc = c - 1 < 0 ? 0 : c - 1, nc = c + 1 > cols - 1 ? cols - 1 : c + 1;
	if (channels == 1) {
		const int dx = image[r * pitch + nc] - image[r * pitch + pc] + 256;
		const int dy = image[nr * pitch + c] - image[pr * pitch + c] + 256;
		return lut + (dy * 512 + dx);
	}
	const FhogLutEntry* e[3];
	for (int k = 0; k < 3; ++k) {
		const int dx = image[r * pitch + nc * 3 + k] - image[r * pitch + pc * 3 + k] + 256;
		const int dy = image[nr * pitch + c * 3 + k] - image[pr * pitch + c * 3 + k] + 256;
		e[k] = lut + (dy * 512 + dx);
	}
	if (e[0]->magnitude > e[1]->magnitude) return e[0]->magnitude > e[2]->magnitude ? e[0] : e[2];
	return e[1]->magnitude > e[2]->magnitude ? e[1] : e[2];
}

/* one bin of the signed histogram of cell (cr, cc): every pixel that contributes to the cell, in raster order
 * (addToSignedHistograms, FhogFilter.hpp:170-207). With interpolation a pixel row touches the cells index1 and index2 of
 * its coefficients; rows [r_lo, r_hi) / columns [c_lo, c_hi) bound the pixels that can touch this cell. */
FHOG_HD float fhog_signed_bin(const FhogLutEntry* lut, const uint8_t* image, int pitch, int rows, int cols, int channels,
		int cell, int crow, int ccol, int interpolate_bins, int interpolate_cells, int cr, int cc, int bin) {
	const int rows_used = crow * cell, cols_used = ccol * cell;
	int r_lo = cr * cell, r_hi = r_lo + cell, c_lo = cc * cell, c_hi = c_lo + cell;
	if (interpolate_cells) { /* pixels up to one cell away interpolate into this cell */
		r_lo -= cell; r_hi += cell; c_lo -= cell; c_hi += cell;
	}
	if (r_lo < 0) r_lo = 0;
	if (c_lo < 0) c_lo = 0;
	if (r_hi > rows_used) r_hi = rows_used;
	if (c_hi > cols_used) c_hi = cols_used;
	float acc = 0.f;
	for (int r = r_lo; r < r_hi; ++r) {
		const FhogCoef R = fhog_pixel_coef(r, cell, crow, interpolate_cells);
		/* the weight(s) of this pixel row for cell row cr; index1 == index2 at the borders: both statements of the
		 * reference hit the same cell, first with weight1 then with weight2 */
		const int hit1 = R.index1 == cr, hit2 = interpolate_cells && R.index2 == cr;
		if (!hit1 && !hit2) continue;
		for (int c = c_lo; c < c_hi; ++c) {
			const FhogCoef Cc = fhog_pixel_coef(c, cell, ccol, interpolate_cells);
			const int chit1 = Cc.index1 == cc, chit2 = interpolate_cells && Cc.index2 == cc;
			if (!chit1 && !chit2) continue;
			const FhogLutEntry* e = fhog_pixel_entry(lut, image, pitch, rows, cols, channels, r, c);
			float bw; /* this pixel's weight for `bin`: weight1 if bin == index1, weight2 if bin == index2 (never both) */
			if (e->index1 == bin) bw = e->weight1;
			else if (interpolate_bins && e->index2 == bin) bw = e->weight2;
			else continue;
			if (!interpolate_cells) { acc = FHOG_ADD(acc, bw); continue; }
			/* statement order of the reference for one pixel: (row1, col1), (row1, col2), (row2, col1), (row2, col2) */
			if (hit1 && chit1) acc = FHOG_ADD(acc, FHOG_MUL(FHOG_MUL(bw, R.weight1), Cc.weight1));
			if (hit1 && chit2) acc = FHOG_ADD(acc, FHOG_MUL(FHOG_MUL(bw, R.weight1), Cc.weight2));
			if (hit2 && chit1) acc = FHOG_ADD(acc, FHOG_MUL(FHOG_MUL(bw, R.weight2), Cc.weight1));
			if (hit2 && chit2) acc = FHOG_ADD(acc, FHOG_MUL(FHOG_MUL(bw, R.weight2), Cc.weight2));
		}
	}
	return acc;
}

/* the LUT entry of a pixel in the compact form the batched kernel (aggdet.cu) keeps in shared memory */
struct FhogPix {
	uint8_t i1, i2;
	uint16_t valid;
	float w1, w2;
};

/* ALL signed bins of cell (cr, cc) in one walk over the pixels that feed the cell, in raster order - the same additions, per
 * bin in the same order, as fhog_signed_bin() performs bin by bin (addToSignedHistograms, FhogFilter.hpp:170-207).
 * entries: the FhogPix of pixel (r, c) at entries[(r - r0) * stride + (c - c0)] for every pixel the walk touches;
 * h: 2 * unsigned_bins floats, zeroed by the caller. */
FHOG_HD void fhog_cell_histogram(const FhogPix* entries, int r0, int c0, int stride, int cell, int crow, int ccol,
		int interpolate_bins, int interpolate_cells, int cr, int cc, float* h) {
	const int rows_used = crow * cell, cols_used = ccol * cell;
	const int halo = interpolate_cells ? cell : 0;
	int r_lo = cr * cell - halo, r_hi = cr * cell + cell + halo, c_lo = cc * cell - halo, c_hi = cc * cell + cell + halo;
	if (r_lo < 0) r_lo = 0;
	if (c_lo < 0) c_lo = 0;
	if (r_hi > rows_used) r_hi = rows_used;
	if (c_hi > cols_used) c_hi = cols_used;
	for (int r = r_lo; r < r_hi; ++r) {
		const FhogCoef R = fhog_pixel_coef(r, cell, crow, interpolate_cells);
		const int hit1 = R.index1 == cr, hit2 = interpolate_cells && R.index2 == cr;
		if (!hit1 && !hit2) continue;
		const FhogPix* row = entries + (r - r0) * stride - c0;
		for (int c = c_lo; c < c_hi; ++c) {
			const FhogCoef Cc = fhog_pixel_coef(c, cell, ccol, interpolate_cells);
			const int chit1 = Cc.index1 == cc, chit2 = interpolate_cells && Cc.index2 == cc;
			if (!chit1 && !chit2) continue;
			const FhogPix e = row[c];
			for (int k = 0; k < 2; ++k) {
				if (k == 1 && !interpolate_bins) break;
				const int bin = k == 0 ? e.i1 : e.i2;
				const float bw = k == 0 ? e.w1 : e.w2;
				float acc = h[bin];
				if (!interpolate_cells) acc = FHOG_ADD(acc, bw);
				else { /* statement order of the reference for one pixel: (row1, col1), (row1, col2), (row2, col1), (row2, col2) */
					if (hit1 && chit1) acc = FHOG_ADD(acc, FHOG_MUL(FHOG_MUL(bw, R.weight1), Cc.weight1));
					if (hit1 && chit2) acc = FHOG_ADD(acc, FHOG_MUL(FHOG_MUL(bw, R.weight1), Cc.weight2));
					if (hit2 && chit1) acc = FHOG_ADD(acc, FHOG_MUL(FHOG_MUL(bw, R.weight2), Cc.weight1));
					if (hit2 && chit2) acc = FHOG_ADD(acc, FHOG_MUL(FHOG_MUL(bw, R.weight2), Cc.weight2));
				}
				h[bin] = acc;
			}
		}
	}
}

/* FhogAggregationFilter::computeGradientEnergy (FhogAggregationFilter.cpp:60-68) of one cell's signed histogram */
FHOG_HD float fhog_energy(const float* signed_hist, int unsigned_bins) {
	float energy = 0.f;
	for (int bin = 0; bin < unsigned_bins; ++bin) {
		const float u = FHOG_ADD(signed_hist[bin], signed_hist[bin + unsigned_bins]);
		energy = FHOG_ADD(energy, FHOG_MUL(u, u));
	}
	return energy;
}

/* computeNormalizers + computeDescriptor (FhogAggregationFilter.cpp:82-150) of cell (r, c): hist = signed histogram of
 * the cell (2 * unsigned_bins floats), energies = [crow][ccol], out = 3 * unsigned_bins + 4 floats (may not alias hist) */
FHOG_HD void fhog_descriptor(const float* hist, const float* energies, int crow, int ccol, int r, int c, int unsigned_bins,
		float alpha, float* out) {
	const int signed_bins = 2 * unsigned_bins;
	const int pr = r - 1 < 0 ? 0 : r - 1, nr = r + 1 > crow - 1 ? crow - 1 : r + 1;
	const int pc = c - 1 < 0 ? 0 : c - 1, nc = c + 1 > ccol - 1 ? ccol - 1 : c + 1;
	const float eps = 1e-4f;
#define FHOG_E(rr, cc2) energies[(rr) * ccol + (cc2)]
#define FHOG_N(a, b, c2, d) (1.f / sqrtf(FHOG_ADD(FHOG_ADD(FHOG_ADD(FHOG_ADD((a), (b)), (c2)), (d)), eps)))
	const float n0 = FHOG_N(FHOG_E(pr, pc), FHOG_E(pr, c), FHOG_E(r, pc), FHOG_E(r, c));
	const float n1 = FHOG_N(FHOG_E(pr, c), FHOG_E(pr, nc), FHOG_E(r, c), FHOG_E(r, nc));
	const float n2 = FHOG_N(FHOG_E(r, pc), FHOG_E(r, c), FHOG_E(nr, pc), FHOG_E(nr, c));
	const float n3 = FHOG_N(FHOG_E(r, c), FHOG_E(r, nc), FHOG_E(nr, c), FHOG_E(nr, nc));
#undef FHOG_N
#undef FHOG_E
	float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
	for (int bin = 0; bin < unsigned_bins; ++bin) { /* contrast-insensitive features */
		const float u = FHOG_ADD(hist[bin], hist[bin + unsigned_bins]);
		const float v0 = fminf(alpha, FHOG_MUL(n0, u)), v1 = fminf(alpha, FHOG_MUL(n1, u));
		const float v2 = fminf(alpha, FHOG_MUL(n2, u)), v3 = fminf(alpha, FHOG_MUL(n3, u));
		out[signed_bins + bin] = (float)FHOG_DMUL(0.5, (double)FHOG_ADD(FHOG_ADD(FHOG_ADD(v0, v1), v2), v3));
	}
	for (int bin = 0; bin < signed_bins; ++bin) { /* contrast-sensitive features */
		const float s = hist[bin];
		const float v0 = fminf(alpha, FHOG_MUL(n0, s)), v1 = fminf(alpha, FHOG_MUL(n1, s));
		const float v2 = fminf(alpha, FHOG_MUL(n2, s)), v3 = fminf(alpha, FHOG_MUL(n3, s));
		out[bin] = (float)FHOG_DMUL(0.5, (double)FHOG_ADD(FHOG_ADD(FHOG_ADD(v0, v1), v2), v3));
		e0 = FHOG_ADD(e0, v0); e1 = FHOG_ADD(e1, v1); e2 = FHOG_ADD(e2, v2); e3 = FHOG_ADD(e3, v3);
	}
	out[signed_bins + unsigned_bins] = (float)FHOG_DMUL(0.2357, (double)e0); /* energy (texture) features */
	out[signed_bins + unsigned_bins + 1] = (float)FHOG_DMUL(0.2357, (double)e1);
	out[signed_bins + unsigned_bins + 2] = (float)FHOG_DMUL(0.2357, (double)e2);
	out[signed_bins + unsigned_bins + 3] = (float)FHOG_DMUL(0.2357, (double)e3);
}

/* ConvolutionFilter::applyTo as AggregatedFeaturesDetector configures it (ConvolutionFilter.cpp:31-49,
 * AggregatedFeaturesDetector.cpp:60-65): score(y, x) = -bias + sum over channels of the correlation of channel c with the
 * kernel's channel c, anchor (0, 0). feat: [rows][cols][D] float32, weights: [kh][kw][D]. The order of the float32 sums is
 * channel by channel, then kernel rows, then kernel columns (the order of the oracle's restatement; OpenCV's own order inside
 * cv::filter2D is not reproducible - it uses a DFT for kernels of >= 50 elements - so parity with the reference is 1e-4). Only
 * positions with the whole window inside the layer are scored (AggregatedFeaturesDetector.cpp:95-98). */
FHOG_HD float aggdet_score(const float* feat, int cols, int D, const float* weights, int kh, int kw, float bias, int y, int x) {
	float score = -bias;
	for (int c = 0; c < D; ++c) {
		float tmp = 0.f;
		for (int i = 0; i < kh; ++i)
			for (int j = 0; j < kw; ++j)
				tmp = FHOG_ADD(tmp, FHOG_MUL(feat[((size_t)(y + i) * cols + (x + j)) * D + c], weights[((size_t)i * kw + j) * D + c]));
		score = FHOG_ADD(score, tmp);
	}
	return score;
}

#endif
