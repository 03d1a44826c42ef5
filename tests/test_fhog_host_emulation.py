"""The FHOG kernels (csrc/fhog.cu) could not be run on a B200 in this round; their arithmetic lives in host/device functions
(csrc/fhog_core.h) that this test compiles with g++ (-ffp-contract=off, as the kernels use _rn intrinsics) and drives in the
same decomposition as the kernels - one call per (cell, bin) for the histograms, one per cell for the descriptors - against
the oracle, which is pinned to the reference's own FhogFilter sources. Bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from featuredetection_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
#include <cmath>
#include <vector>
#include <cstring>
#include "fhog_core.h"

/* the host-side table builder of fhog.cu, restated here through the same formulas (FhogFilter.cpp:36-57) */
static void build_lut(int unsigned_bins, int interpolate_bins, std::vector<FhogLutEntry>& lut) {
	const int signed_bins = 2 * unsigned_bins;
	const float two_pi = static_cast<float>(2 * M_PI);
	const float value2bin = signed_bins / two_pi;
	lut.assign(512 * 512, FhogLutEntry{0, 0, 0.f, 0.f, 0.f});
	for (int cx = 1; cx < 512; ++cx) {
		const float gx = (cx - 256) / (255.0f * 2.0f);
		for (int cy = 1; cy < 512; ++cy) {
			const float gy = (cy - 256) / (255.0f * 2.0f);
			FhogLutEntry e{0, 0, 0.f, 0.f, 0.f};
			e.magnitude = std::sqrt(gx * gx + gy * gy);
			float orientation = std::atan2(gy, gx);
			if (orientation < 0) orientation += two_pi;
			if (interpolate_bins) {
				const float bin = orientation * value2bin;
				e.index1 = static_cast<int>(bin); e.index2 = e.index1 + 1;
				if (e.index2 == signed_bins) e.index2 = 0;
				e.weight2 = e.magnitude * (bin - e.index1); e.weight1 = e.magnitude - e.weight2;
			} else {
				int bin = static_cast<int>(orientation * value2bin + 0.5f);
				if (bin == signed_bins) bin = 0;
				e.index1 = bin; e.weight1 = e.magnitude;
			}
			lut[(size_t)cy * 512 + cx] = e;
		}
	}
}

extern "C" long long emulate_fhog(const unsigned char* image, int cols, int rows, int channels, int cell, int unsigned_bins,
		int interpolate_bins, int interpolate_cells, float alpha, float* out) {
	std::vector<FhogLutEntry> lut;
	build_lut(unsigned_bins, interpolate_bins, lut);
	const int crow = rows / cell, ccol = cols / cell, signed_bins = 2 * unsigned_bins, D = 3 * unsigned_bins + 4;
	const int pitch = cols * channels;
	std::vector<float> hist((size_t)crow * ccol * signed_bins), energies((size_t)crow * ccol);
	for (long long idx = 0; idx < (long long)crow * ccol * signed_bins; ++idx) { /* fhog_hist_kernel: thread = (cell, bin) */
		const int bin = (int)(idx % signed_bins), cellidx = (int)(idx / signed_bins);
		const int cr = cellidx / ccol, cc = cellidx - cr * ccol;
		hist[idx] = fhog_signed_bin(lut.data(), image, pitch, rows, cols, channels, cell, crow, ccol, interpolate_bins, interpolate_cells, cr, cc, bin);
	}
	for (int i = 0; i < crow * ccol; ++i) energies[i] = fhog_energy(&hist[(size_t)i * signed_bins], unsigned_bins); /* fhog_energy_kernel */
	for (int i = 0; i < crow * ccol; ++i) /* fhog_desc_kernel: thread = cell */
		fhog_descriptor(&hist[(size_t)i * signed_bins], energies.data(), crow, ccol, i / ccol, i % ccol, unsigned_bins, alpha, out + (size_t)i * D);
	return (long long)crow * ccol * D;
}

/* aggdet_hist_kernel (csrc/aggdet.cu), tile by tile: phase 1 = one FhogPix per pixel of the tile + halo, phase 2 = thread per cell
 * (fhog_cell_histogram); then the energies and descriptors as aggdet_desc_kernel computes them */
extern "C" long long emulate_fhog_tiled(const unsigned char* image, int cols, int rows, int cell, int unsigned_bins,
		int interpolate_bins, int interpolate_cells, float alpha, int tc, float* out) {
	std::vector<FhogLutEntry> lut;
	build_lut(unsigned_bins, interpolate_bins, lut);
	const int crow = rows / cell, ccol = cols / cell, sb = 2 * unsigned_bins, D = 3 * unsigned_bins + 4;
	const int rows_used = crow * cell, cols_used = ccol * cell;
	std::vector<float> hist((size_t)crow * ccol * sb), energies((size_t)crow * ccol);
	const int halo = interpolate_cells ? cell : 0, region = tc * cell + 2 * halo;
	std::vector<FhogPix> px((size_t)region * region);
	for (int ty = 0; ty * tc < crow; ++ty)
		for (int tx = 0; tx * tc < ccol; ++tx) {
			const int cr0 = ty * tc, cc0 = tx * tc, pr0 = cr0 * cell - halo, pc0 = cc0 * cell - halo;
			for (int i = 0; i < region * region; ++i) {
				const int r = pr0 + i / region, c = pc0 + i % region;
				FhogPix e; e.i1 = 0; e.i2 = 0; e.valid = 0; e.w1 = 0.f; e.w2 = 0.f;
				if (r >= 0 && c >= 0 && r < rows_used && c < cols_used) {
					const FhogLutEntry* q = fhog_pixel_entry(lut.data(), image, cols, rows, cols, 1, r, c);
					e.i1 = (uint8_t)q->index1; e.i2 = (uint8_t)q->index2; e.w1 = q->weight1; e.w2 = q->weight2; e.valid = 1;
				}
				px[i] = e;
			}
			for (int t = 0; t < tc * tc; ++t) {
				const int cr = cr0 + t / tc, cc = cc0 + t % tc;
				if (cr >= crow || cc >= ccol) continue;
				std::vector<float> h(sb + 1, 0.f);
				fhog_cell_histogram(px.data(), pr0, pc0, region, cell, crow, ccol, interpolate_bins, interpolate_cells, cr, cc, h.data());
				for (int b = 0; b < sb; ++b) hist[((size_t)cr * ccol + cc) * sb + b] = h[b];
				energies[(size_t)cr * ccol + cc] = fhog_energy(h.data(), unsigned_bins);
			}
		}
	for (int i = 0; i < crow * ccol; ++i)
		fhog_descriptor(&hist[(size_t)i * sb], energies.data(), crow, ccol, i / ccol, i % ccol, unsigned_bins, alpha, out + (size_t)i * D);
	return (long long)crow * ccol * D;
}

extern "C" void emulate_score_map(const float* feat, int rows, int cols, int D, const float* weights, int kh, int kw, float bias, float* scores) {
	const int vh = rows - kh + 1, vw = cols - kw + 1;
	for (int i = 0; i < vh * vw; ++i) scores[i] = aggdet_score(feat, cols, D, weights, kh, kw, bias, i / vw, i % vw); /* aggdet_score_kernel */
}
'''

ARGS = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]


@pytest.fixture(scope="module")
def emu(tmp_path_factory, built):
    d = tmp_path_factory.mktemp("fhog_emu")
    src = d / "emu.cpp"
    src.write_text(DRIVER)
    so = d / "libfhogemu.so"
    r = subprocess.run(["g++", "-std=c++11", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", str(so), str(src),
                        "-I" + os.path.join(ROOT, "featuredetection_b200", "csrc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(so))
    lib.emulate_fhog.restype = C.c_longlong
    lib.emulate_fhog.argtypes = ARGS
    lib.emulate_fhog_tiled.restype = C.c_longlong
    lib.emulate_fhog_tiled.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
    lib.emulate_score_map.restype = None
    lib.emulate_score_map.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
    return lib


@pytest.mark.parametrize("cell,bins,ib,ic,alpha", [(4, 9, True, True, 0.2), (8, 9, False, True, 0.2), (4, 6, True, False, 0.2),
                                                   (5, 9, False, False, 0.5), (6, 8, True, True, 1.0), (3, 9, False, True, 0.2)])
def test_kernel_arithmetic_equals_the_pinned_oracle(emu, cell, bins, ib, ic, alpha):
    from oracle import fdoracle as fo
    gray = syn.synthetic_frame(3)[:131, :203]
    rng = np.random.default_rng(1)
    bgr = np.stack([gray, np.roll(gray, 3, 1), rng.integers(0, 256, gray.shape, dtype=np.uint8)], axis=2)
    flat = np.full((40, 48), 77, np.uint8)
    for img in (gray, bgr, flat, np.ascontiguousarray(gray[:cell, :cell * 2]), np.ascontiguousarray(gray[:2 * cell + 1, :cell])):
        img = np.ascontiguousarray(img)
        rows, cols = img.shape[:2]
        ch = 1 if img.ndim == 2 else 3
        want = fo.fhog(img, cell, bins, ib, ic, alpha)
        got = np.full_like(want, np.nan)
        n = emu.emulate_fhog(img.ctypes.data, cols, rows, ch, cell, bins, int(ib), int(ic), alpha, got.ctypes.data)
        assert n == want.size
        assert np.array_equal(got, want), (img.shape, float(np.nanmax(np.abs(got - want))))


@pytest.mark.parametrize("cell,bins,ib,ic,tc", [(4, 9, False, True, 8), (8, 9, True, True, 6), (4, 6, True, False, 8), (5, 9, False, False, 3),
                                                 (6, 8, True, True, 2), (3, 9, False, True, 1)])
def test_batched_kernel_decomposition_equals_the_pinned_oracle(emu, cell, bins, ib, ic, tc):
    """the tile + halo staging and the per-cell raster walk of aggdet_hist_kernel (fhog_cell_histogram), emulated on the host"""
    from oracle import fdoracle as fo
    gray = np.ascontiguousarray(syn.synthetic_frame(3)[:131, :203])
    flat = np.full((40, 48), 77, np.uint8)
    for img in (gray, flat, np.ascontiguousarray(gray[:cell, :cell * 2]), np.ascontiguousarray(gray[:2 * cell + 1, :cell])):
        img = np.ascontiguousarray(img)
        rows, cols = img.shape
        want = fo.fhog(img, cell, bins, ib, ic, 0.2)
        got = np.full_like(want, np.nan)
        n = emu.emulate_fhog_tiled(img.ctypes.data, cols, rows, cell, bins, int(ib), int(ic), 0.2, tc, got.ctypes.data)
        assert n == want.size
        assert np.array_equal(got, want), (img.shape, float(np.nanmax(np.abs(got - want))))


def test_score_map_arithmetic_equals_the_oracle_restatement(emu):
    """aggdet_score (thread = score-map position) against the score maps of oracle/fdoracle.py:aggregated_features_detect"""
    from oracle import fdoracle as fo
    frame = np.ascontiguousarray(syn.synthetic_frame(6)[:160, :200])
    rng = np.random.default_rng(4)
    w = rng.normal(0, 0.1, (5, 6, 31)).astype(np.float32)
    rects, scores, maps = fo.aggregated_features_detect(frame, w, bias=0.25, threshold=1e9, cell=4, octave_layer_count=3, want_scores=True)
    inc = 0.5 ** (1.0 / 3)
    checked = 0
    # the same layers the restatement used (its scale limits are recomputed here through its own public pieces)
    mn = (6 * 4) / float(int(160 / ((5 * 4) / (6 * 4.0))) if (5 * 4) / (6 * 4.0) > 160 / 200.0 else 200)
    import math
    min_scale = math.pow(inc, int(math.log(mn) / math.log(inc)))
    _, layers = fo.pyramid(frame, inc, min_scale, 1.0)
    assert len(layers) == len(maps)
    for (_, _, img), want in zip(layers, maps):
        feat = fo.fhog(img, 4)
        rows, cols, D = feat.shape
        if want.size == 0:
            continue
        got = np.full_like(want, np.nan)
        emu.emulate_score_map(feat.ctypes.data, rows, cols, D, w.ctypes.data, 5, 6, 0.25, got.ctypes.data)
        assert np.array_equal(got, want), float(np.nanmax(np.abs(got - want)))
        checked += 1
    assert checked >= 3


@pytest.mark.gpu
@pytest.mark.parametrize("cell,bins,ib,ic,alpha", [(4, 9, True, True, 0.2), (8, 9, False, False, 0.2)])
def test_fdb_fhog_on_the_gpu(cell, bins, ib, ic, alpha):
    """fdb_fhog (the kernels) against the oracle, bit for bit (first run on a B200: round 2, gpurun_out/r2a_fhog.log)"""
    from oracle import fdoracle as fo
    from featuredetection_b200 import capi
    from featuredetection_b200.detector import Context
    ctx = Context(0)
    gray = np.ascontiguousarray(syn.synthetic_frame(3)[:131, :203])
    bgr = np.ascontiguousarray(np.stack([gray, np.roll(gray, 3, 1), np.roll(gray, 5, 0)], axis=2))
    for img in (gray, bgr):
        rows, cols = img.shape[:2]
        ch = 1 if img.ndim == 2 else 3
        want = fo.fhog(img, cell, bins, ib, ic, alpha)
        got = np.full_like(want, np.nan)
        capi.check(ctx.lib, ctx.lib.fdb_fhog(ctx.h, img.ctypes.data, cols * ch, cols, rows, ch, cell, bins, int(ib), int(ic), alpha, got.ctypes.data))
        assert np.array_equal(got, want)


@pytest.mark.gpu
def test_fdb_fhog_score_map_on_the_gpu():
    from oracle import fdoracle as fo
    from featuredetection_b200 import capi
    from featuredetection_b200.detector import Context
    ctx = Context(0)
    img = np.ascontiguousarray(syn.synthetic_frame(6)[:160, :200])
    rng = np.random.default_rng(4)
    w = rng.normal(0, 0.1, (5, 6, 31)).astype(np.float32)
    feat = fo.fhog(img, 4)
    rows, cols, D = feat.shape
    want = np.zeros((rows - 4, cols - 5), np.float32)
    for y in range(want.shape[0]):
        for x in range(want.shape[1]):
            s = np.float32(-0.25)
            for c in range(D):
                t = np.float32(0)
                for i in range(5):
                    for j in range(6):
                        t = np.float32(t + np.float32(feat[y + i, x + j, c] * w[i, j, c]))
                s = np.float32(s + t)
            want[y, x] = s
    got = np.full_like(want, np.nan)
    capi.check(ctx.lib, ctx.lib.fdb_fhog_score_map(ctx.h, img.ctypes.data, 200, 200, 160, 1, 4, 9, 0, 1, 0.2, w.ctypes.data, 5, 6, 0.25, got.ctypes.data))
    assert np.array_equal(got, want)
