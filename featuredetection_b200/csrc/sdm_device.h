/*
 * sdm_device.h - device-side model of the supervised-descent regressor (sdm.cu) and its launchers.
 */
#ifndef FDB_SDM_DEVICE_H_
#define FDB_SDM_DEVICE_H_

#include <cuda_runtime.h>

#include <cstdint>

#define SDM_MAX_STEPS 8

namespace fdb {

struct DevSdm {              /* passed to the kernels by value */
	int L, steps, K, N;      /* landmarks, cascade steps, K = 279 L feature length, N = 2 L */
	const float* R[SDM_MAX_STEPS]; /* [K + 1][N] float32, row K = bias (SdmLandmarkModel.hpp:241) */
	double step_factor[SDM_MAX_STEPS]; /* 1 / (1 + exp(step + 1 - steps)) (SdmLandmarkModel.hpp:226) */
	float ox[9], oy[9];      /* vl_hog orientation vectors (hog.c:193-202) */
	int bin_of[30];          /* floor((x + 0.5) / 10 - 0.5) for the 30-px patch (hog.c:697-708) */
	float w1_of[30], w2_of[30];
	float tex;               /* 1 / sqrt(18) */
};

void sdm_fill_tables(DevSdm* m, int L, int steps);
void launch_sdm_hog(cudaStream_t st, const DevSdm& m, const uint8_t* frames, int W, int H, const int* face_frame, const float* shapes,
		int step, const float* pts_xy, int window_half, int n_faces, float* features, int* status);
void launch_sdm_gemm(cudaStream_t st, const DevSdm& m, int step, const float* features, int n_faces, float* delta);
void launch_sdm_update(cudaStream_t st, const DevSdm& m, int step, const float* delta, float* shapes, const int* status, int n_faces);

} // namespace fdb
#endif
