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


@pytest.mark.gpu_unverified
@pytest.mark.skipif(os.environ.get("FDB_RUN_UNVERIFIED") != "1", reason="fhog.cu has not run on a B200 yet: set FDB_RUN_UNVERIFIED=1 on a GPU box")
@pytest.mark.parametrize("cell,bins,ib,ic,alpha", [(4, 9, True, True, 0.2), (8, 9, False, False, 0.2)])
def test_fdb_fhog_on_the_gpu(cell, bins, ib, ic, alpha):
    """first thing to run next round: fdb_fhog (the kernels) against the oracle, bit for bit"""
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
