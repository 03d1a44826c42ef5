"""The C++ adapters (adapters/fdb200_adapters.hpp) must compile against the reference's UNCHANGED
interface headers: PyramidFeatureExtractor, ProbabilisticClassifier, Detector, condensation::MeasurementModel. Needs /root/reference
(present where the driver runs the CPU suite, absent on the GPU box -> skipped there)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "libDetection")), reason="reference tree not mounted")
def test_adapters_compile_against_reference_headers(tmp_path):
    cmd = ["g++", "-std=c++11", "-c", "-o", str(tmp_path / "check.o"),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "adapters"), "-I" + os.path.join(ROOT, "oracle", "shim"),
           "-I" + os.path.join(REF, "libClassification", "include"), "-I" + os.path.join(REF, "libImageProcessing", "include"),
           "-I" + os.path.join(REF, "libDetection", "include"), "-I" + os.path.join(REF, "libCondensation", "include"),
           "-Wno-deprecated-declarations",
           os.path.join(ROOT, "adapters", "check_adapters.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_header_is_plain_c(tmp_path):
    """include/fdb200.h is a C ABI: it must compile as C11 with no C++ or CUDA types."""
    src = tmp_path / "t.c"
    src.write_text('#include "fdb200.h"\nint main(void) { fdb_window_score s; s.level = FDB_ABI_VERSION; return s.level == 0; }\n')
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-c", "-o", str(tmp_path / "t.o"), "-I" + os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
