/*
 * fd_fhog.c - CPU restatement of the reference's FHOG layer filter (TEST INFRASTRUCTURE: only tests/, smoke() and the CPU
 * legs of bench.py may use anything under oracle/). Groundwork for SURVEY.md 8(f) rank 2 (AggregatedFeaturesDetector): the
 * feature map the linear SVM of that detector is convolved with.
 *
 *   FhogFilter::applyTo / createGradientLut / computeSignedHistograms / getBinCoefficients / addToSignedHistograms
 *       libImageProcessing/src/imageprocessing/filtering/FhogFilter.cpp:20-122
 *       libImageProcessing/include/imageprocessing/filtering/FhogFilter.hpp:112-208
 *   GradientOrientationFilter::computeOrientation   filtering/GradientOrientationFilter.cpp:137-144 (full orientations)
 *   GradientMagnitudeFilter::computeMagnitude        filtering/GradientMagnitudeFilter.cpp:80-86
 *   FhogAggregationFilter::computeDescriptors        filtering/FhogAggregationFilter.cpp:43-150 (eps 1e-4, 0.5 and 0.2357)
 *
 * Pinned: bit-identical to those sources compiled unmodified into oracle/_ref (tests/test_fhog_oracle.py).
 * All arithmetic is float32 in the reference's order; the in-place call computeDescriptors(descriptors, descriptors, ..)
 * of FhogFilter.cpp:66 is reproduced (the signed histogram of a cell is overwritten bin by bin).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "fd_oracle.h"

typedef struct { int index1, index2; float weight1, weight2; } fhog_coef;
typedef struct { fhog_coef bins; float magnitude; } fhog_lut_entry;

#define FHOG_TWO_PI ((float)(2 * M_PI))

static float fhog_orientation(float gx, float gy) { /* GradientOrientationFilter.cpp:137-144 with half == false */
	float orientation = atan2f(gy, gx);
	if (orientation < 0) orientation += FHOG_TWO_PI;
	return orientation;
}

static fhog_lut_entry* fhog_build_lut(int signed_bins, int interpolate_bins) { /* FhogFilter.cpp:36-57 */
	fhog_lut_entry* lut = (fhog_lut_entry*)calloc(512 * 512, sizeof(fhog_lut_entry));
	const float value2bin = signed_bins / FHOG_TWO_PI; /* FhogFilter.cpp:27 */
	for (int cx = 1; cx < 512; ++cx) {
		float gx = (cx - 256) / (255.0f * 2.0f);
		for (int cy = 1; cy < 512; ++cy) {
			float gy = (cy - 256) / (255.0f * 2.0f);
			fhog_lut_entry e;
			memset(&e, 0, sizeof e);
			e.magnitude = sqrtf(gx * gx + gy * gy);
			float orientation = fhog_orientation(gx, gy);
			if (interpolate_bins) { /* computeInterpolatedBins, FhogFilter.cpp:108-118 */
				const float bin = orientation * value2bin;
				e.bins.index1 = (int)bin;
				e.bins.index2 = e.bins.index1 + 1;
				if (e.bins.index2 == signed_bins) e.bins.index2 = 0;
				e.bins.weight2 = e.magnitude * (bin - e.bins.index1);
				e.bins.weight1 = e.magnitude - e.bins.weight2;
			} else { /* computeBin, :100-105 */
				int bin = (int)(orientation * value2bin + 0.5f);
				if (bin == signed_bins) bin = 0;
				e.bins.index1 = bin;
				e.bins.weight1 = e.magnitude;
			}
			lut[cy * 512 + cx] = e;
		}
	}
	return lut;
}

static void fhog_interp(fhog_coef* c, int size_px, int size_cells, int cell, int interpolate_cells) { /* FhogFilter.cpp:69-92 */
	for (int pixel = 0; pixel < size_px; ++pixel) {
		if (interpolate_cells) {
			float real = (pixel + 0.5f) / cell - 0.5f;
			int i1 = (int)floorf(real), i2 = i1 + 1;
			float w2 = real - i1, w1 = i2 - real;
			if (i1 < 0) { i1 = i2; w1 = 0; }
			else if (i2 >= size_cells) { i2 = i1; w2 = 0; }
			c[pixel].index1 = i1; c[pixel].index2 = i2; c[pixel].weight1 = w1; c[pixel].weight2 = w2;
		} else {
			c[pixel].index1 = pixel / cell; c[pixel].index2 = -1; c[pixel].weight1 = 1; c[pixel].weight2 = 0;
		}
	}
}

/* image: rows x cols x channels (1 or 3) u8, row pitch = cols * channels. out: (rows / cell) x (cols / cell) x
 * (3 * unsigned_bins + 4) float32. Returns the number of floats written, or -1. */
int64_t fdo_fhog(const uint8_t* image, int cols, int rows, int channels, int cell, int unsigned_bins, int interpolate_bins,
		int interpolate_cells, float alpha, float* out) {
	if (!image || !out || cell < 1 || unsigned_bins < 1 || (channels != 1 && channels != 3) || !(alpha > 0)) return -1;
	const int signed_bins = 2 * unsigned_bins, D = signed_bins + unsigned_bins + 4;
	const int crow = rows / cell, ccol = cols / cell;
	fhog_lut_entry* lut = fhog_build_lut(signed_bins, interpolate_bins);
	memset(out, 0, sizeof(float) * (size_t)crow * ccol * D);
	fhog_coef* rc = (fhog_coef*)malloc(sizeof(fhog_coef) * (size_t)(crow * cell + 1));
	fhog_coef* cc = (fhog_coef*)malloc(sizeof(fhog_coef) * (size_t)(ccol * cell + 1));
	fhog_interp(rc, crow * cell, crow, cell, interpolate_cells);
	fhog_interp(cc, ccol * cell, ccol, cell, interpolate_cells);
	const int pitch = cols * channels;
	for (int r = 0; r < crow * cell; ++r) { /* computeSignedHistograms, FhogFilter.hpp:112-125 */
		const int pr = r - 1 < 0 ? 0 : r - 1, nr = r + 1 > rows - 1 ? rows - 1 : r + 1;
		for (int c = 0; c < ccol * cell; ++c) {
			const int pc = c - 1 < 0 ? 0 : c - 1, nc = c + 1 > cols - 1 ? cols - 1 : c + 1;
			fhog_coef b;
			if (channels == 1) { /* getBinCoefficients<true>, :127-136 */
				int dx = image[r * pitch + nc] - image[r * pitch + pc] + 256;
				int dy = image[nr * pitch + c] - image[pr * pitch + c] + 256;
				b = lut[dy * 512 + dx].bins;
			} else { /* getBinCoefficients<false>, :138-168: the channel with the strongest gradient */
				int idx[3];
				for (int k = 0; k < 3; ++k) {
					int dx = image[r * pitch + nc * 3 + k] - image[r * pitch + pc * 3 + k] + 256;
					int dy = image[nr * pitch + c * 3 + k] - image[pr * pitch + c * 3 + k] + 256;
					idx[k] = dy * 512 + dx;
				}
				if (lut[idx[0]].magnitude > lut[idx[1]].magnitude)
					b = lut[idx[0]].magnitude > lut[idx[2]].magnitude ? lut[idx[0]].bins : lut[idx[2]].bins;
				else
					b = lut[idx[1]].magnitude > lut[idx[2]].magnitude ? lut[idx[1]].bins : lut[idx[2]].bins;
			}
			const fhog_coef R = rc[r], C = cc[c];
			if (interpolate_cells) { /* addToSignedHistograms, :170-198 */
				float* h11 = out + ((size_t)R.index1 * ccol + C.index1) * D;
				float* h12 = out + ((size_t)R.index1 * ccol + C.index2) * D;
				float* h21 = out + ((size_t)R.index2 * ccol + C.index1) * D;
				float* h22 = out + ((size_t)R.index2 * ccol + C.index2) * D;
				h11[b.index1] += b.weight1 * R.weight1 * C.weight1;
				if (interpolate_bins) h11[b.index2] += b.weight2 * R.weight1 * C.weight1;
				h12[b.index1] += b.weight1 * R.weight1 * C.weight2;
				if (interpolate_bins) h12[b.index2] += b.weight2 * R.weight1 * C.weight2;
				h21[b.index1] += b.weight1 * R.weight2 * C.weight1;
				if (interpolate_bins) h21[b.index2] += b.weight2 * R.weight2 * C.weight1;
				h22[b.index1] += b.weight1 * R.weight2 * C.weight2;
				if (interpolate_bins) h22[b.index2] += b.weight2 * R.weight2 * C.weight2;
			} else { /* :199-207 */
				float* h = out + ((size_t)R.index1 * ccol + C.index1) * D;
				h[b.index1] += b.weight1;
				if (interpolate_bins) h[b.index2] += b.weight2;
			}
		}
	}
	/* FhogAggregationFilter::computeDescriptors (FhogAggregationFilter.cpp:43-150), in place */
	float* energies = (float*)malloc(sizeof(float) * (size_t)(crow * ccol + 1));
	for (int i = 0; i < crow * ccol; ++i) { /* computeGradientEnergy, :60-68 */
		const float* h = out + (size_t)i * D;
		float energy = 0;
		for (int bin = 0; bin < unsigned_bins; ++bin) { float u = h[bin] + h[bin + unsigned_bins]; energy += u * u; }
		energies[i] = energy;
	}
	const float eps = 1e-4f; /* FhogAggregationFilter.cpp:20 */
	for (int r = 0; r < crow; ++r)
		for (int c = 0; c < ccol; ++c) {
			const int pr = r - 1 < 0 ? 0 : r - 1, nr = r + 1 > crow - 1 ? crow - 1 : r + 1;
			const int pc = c - 1 < 0 ? 0 : c - 1, nc = c + 1 > ccol - 1 ? ccol - 1 : c + 1;
#define E(rr, cc2) energies[(rr) * ccol + (cc2)]
			const float n[4] = { /* computeNormalizers, :82-103 */
				1.f / sqrtf(E(pr, pc) + E(pr, c) + E(r, pc) + E(r, c) + eps),
				1.f / sqrtf(E(pr, c) + E(pr, nc) + E(r, c) + E(r, nc) + eps),
				1.f / sqrtf(E(r, pc) + E(r, c) + E(nr, pc) + E(nr, c) + eps),
				1.f / sqrtf(E(r, c) + E(r, nc) + E(nr, c) + E(nr, nc) + eps)};
#undef E
			float* d = out + ((size_t)r * ccol + c) * D; /* descriptor and signed histogram share the storage */
			float energy[4] = {0, 0, 0, 0};
			for (int bin = 0; bin < unsigned_bins; ++bin) { /* contrast-insensitive features, :129-135 */
				float u = d[bin] + d[bin + unsigned_bins];
				float v0 = fminf(alpha, n[0] * u), v1 = fminf(alpha, n[1] * u), v2 = fminf(alpha, n[2] * u), v3 = fminf(alpha, n[3] * u);
				d[signed_bins + bin] = (float)(0.5 * (v0 + v1 + v2 + v3));
			}
			for (int bin = 0; bin < signed_bins; ++bin) { /* contrast-sensitive features, :137-141 */
				float s = d[bin];
				float v0 = fminf(alpha, n[0] * s), v1 = fminf(alpha, n[1] * s), v2 = fminf(alpha, n[2] * s), v3 = fminf(alpha, n[3] * s);
				d[bin] = (float)(0.5 * (v0 + v1 + v2 + v3));
				energy[0] += v0; energy[1] += v1; energy[2] += v2; energy[3] += v3;
			}
			for (int k = 0; k < 4; ++k) d[signed_bins + unsigned_bins + k] = (float)(0.2357 * energy[k]); /* :143-146 */
		}
	free(energies); free(rc); free(cc); free(lut);
	return (int64_t)crow * ccol * D;
}
