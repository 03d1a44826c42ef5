/*
 * svm.cu - stage 3 of the cascade on the GPU: RBF support vector machine (sm_100a).
 *
 * One CTA evaluates one feature vector against all support vectors:
 *   SvmClassifier::computeHyperplaneDistance      libClassification/src/classification/SvmClassifier.cpp:55-60
 *   RbfKernel::compute / sum of squared differences libClassification/include/classification/RbfKernel.hpp:32-40,78-108
 * For windows of a frame the HistEq64 patch (HistEq64Filter.cpp:32-125) is rebuilt in shared
 * memory from the pyramid layer (the stage-1 kernel does not store 400-byte patches per window).
 *
 * Exactness: u8 SSD is an exact integer (vabsdiffu4 + dp4a); float SSD keeps the reference's
 * sequential float32 order; the kernel value exp(-gamma*ssd) and the sum over support vectors
 * are double precision, accumulated by ONE thread in support-vector order (double addition is
 * not associative, and classify() compares the sum against a threshold).
 * Support vectors are stored transposed ([word][sv]) so that the 256 threads of a CTA, each
 * owning one support vector, read consecutive words.
 */
#include <cuda_runtime.h>
#include <cstdint>

#include "fdb_internal.h"
#include "wvm_device.h"

namespace fdb {

#define SVM_THREADS 256
#define SVM_CHUNK 2048

template <int MODE> /* 0: window of a frame (u8, hq64 built here), 1: given u8 vectors, 2: given f32 vectors */
__global__ void __launch_bounds__(SVM_THREADS) svm_kernel(const DevSvm s, int patch_w, int patch_h,
		const uint8_t* __restrict__ frames, int W, int H, const uint8_t* __restrict__ arena, int64_t arena_stride,
		const DevLayer* __restrict__ layers, const SvmItem* __restrict__ items,
		const void* __restrict__ vectors, double* __restrict__ distance_out, int* __restrict__ level_out) {
	extern __shared__ __align__(16) unsigned char svm_smem[];
	double* s_prod = reinterpret_cast<double*>(svm_smem);                    /* [SVM_CHUNK] */
	uint32_t* s_x = reinterpret_cast<uint32_t*>(svm_smem + sizeof(double) * SVM_CHUNK); /* [nwords] or float[dim] */
	__shared__ uint32_t s_hist[64];
	__shared__ uint8_t s_eq[64];
	const int tid = threadIdx.x;
	const int item = blockIdx.x;

	if (MODE == 0) {
		const SvmItem it = items[item];
		const DevLayer L = layers[it.layer];
		const uint8_t* img = (L.offset < 0 ? frames + (int64_t)it.frame * W * H
				: arena + (int64_t)it.frame * arena_stride + L.offset) + (int64_t)it.y * L.pitch + it.x;
		const int npix = patch_w * patch_h;
		if (tid < 64) s_hist[tid] = 0;
		for (int i = tid; i < s.nwords; i += SVM_THREADS) s_x[i] = 0;
		__syncthreads();
		for (int i = tid; i < npix; i += SVM_THREADS) {
			const int r = i / patch_w, c = i - r * patch_w;
			atomicAdd(&s_hist[img[(int64_t)r * L.pitch + c] >> 2], 1u);
		}
		__syncthreads();
		if (tid == 0) { /* sequential float cumsum, HistEq64Filter.cpp:70-87,97 */
			const float stretch = __fdiv_rn(255.0f, (float)npix);
			float cdf = 0.f;
			for (int b = 0; b < 64; ++b) {
				cdf = __fadd_rn(cdf, __fmul_rn((float)s_hist[b], stretch));
				const float fl = floorf(cdf);
				s_eq[b] = (uint8_t)((int)fl + (__fsub_rn(cdf, fl) >= 0.5f ? 1 : 0));
			}
		}
		__syncthreads();
		for (int i = tid; i < npix; i += SVM_THREADS) {
			const int r = i / patch_w, c = i - r * patch_w;
			const uint32_t e = s_eq[img[(int64_t)r * L.pitch + c] >> 2];
			atomicOr(&s_x[i >> 2], e << (8 * (i & 3)));
		}
	} else if (MODE == 1) {
		const uint8_t* v = reinterpret_cast<const uint8_t*>(vectors) + (int64_t)item * s.dim;
		for (int i = tid; i < s.nwords; i += SVM_THREADS) {
			uint32_t w = 0;
			for (int k = 0; k < 4; ++k) if (4 * i + k < s.dim) w |= (uint32_t)v[4 * i + k] << (8 * k);
			s_x[i] = w;
		}
	} else {
		const float* v = reinterpret_cast<const float*>(vectors) + (int64_t)item * s.dim;
		float* xf = reinterpret_cast<float*>(s_x);
		for (int i = tid; i < s.dim; i += SVM_THREADS) xf[i] = v[i];
	}
	__syncthreads();

	/* RvmClassifier (s.rvm_filters > 0): the same kernel values feed the cascade of RvmClassifier::computeHyperplaneDistance
	 * (RvmClassifier.cpp:75-85) through computeHyperplaneDistanceCached (:94-112): level 0 is -bias + c[0][0] k_0, level l
	 * adds c[l][l] k_l to the distance of level l - 1, until a level's distance falls below its threshold */
	const bool rvm = s.rvm_filters > 0;
	const int n_vec = rvm ? s.rvm_filters : s.num_sv;
	double distance = -(double)s.bias; /* SvmClassifier.cpp:56 / RvmClassifier.cpp:104: double distance = -bias */
	int level = -1;
	bool done = false;
	for (int base = 0; base < n_vec; base += SVM_CHUNK) {
		const int cnt = min(SVM_CHUNK, n_vec - base);
		for (int i = tid; i < cnt; i += SVM_THREADS) {
			const int sv = base + i;
			double kv;
			if (MODE != 2) {
				const uint32_t* __restrict__ col = s.sv_words + sv;
				if (s.kernel == FDB_KERNEL_RBF) {                             /* RbfKernel.hpp:39,78-88 */
					uint32_t acc = 0;
					for (int j = 0; j < s.nwords; ++j) {
						const uint32_t d = __vabsdiffu4(s_x[j], col[(size_t)j * s.num_sv]);
						acc = __dp4a(d, d, acc);
					}
					kv = exp(__dmul_rn(-s.gamma, (double)(int)acc));
				} else if (s.kernel == FDB_KERNEL_HIK) {                      /* HistogramIntersectionKernel.hpp:59-67: int sum */
					uint32_t acc = 0;
					for (int j = 0; j < s.nwords; ++j) acc = __dp4a(__vminu4(s_x[j], col[(size_t)j * s.num_sv]), 0x01010101u, acc);
					kv = (double)(int)acc;
				} else {                                                      /* cv::Mat::dot on CV_8U: an exact integer */
					uint32_t acc = 0;
					for (int j = 0; j < s.nwords; ++j) acc = __dp4a(s_x[j], col[(size_t)j * s.num_sv], acc);
					kv = (double)(int)acc;
				}
			} else {
				const float* xf = reinterpret_cast<const float*>(s_x);
				const float* __restrict__ col = s.sv_f32 + sv;
				if (s.kernel == FDB_KERNEL_RBF) {                             /* RbfKernel.hpp:97-108: float32, sequential */
					float sum = 0.f;
					for (int k = 0; k < s.dim; ++k) {
						const float diff = __fsub_rn(xf[k], col[(size_t)k * s.num_sv]);
						sum = __fadd_rn(sum, __fmul_rn(diff, diff));
					}
					kv = exp(__dmul_rn(-s.gamma, (double)sum));
				} else if (s.kernel == FDB_KERNEL_HIK) {                      /* HistogramIntersectionKernel.hpp:72-80: float32 sum */
					float sum = 0.f;
					for (int k = 0; k < s.dim; ++k) sum = __fadd_rn(sum, fminf(xf[k], col[(size_t)k * s.num_sv]));
					kv = (double)sum;
				} else {
					/* cv::Mat::dot on CV_32F (OpenCV 2.4.3 dotProd_<float, double>): float64 products, four at a time */
					double r = 0.0;
					int k = 0;
					for (; k <= s.dim - 4; k += 4) {
						double q = __dmul_rn((double)xf[k], (double)col[(size_t)k * s.num_sv]);
						q = __dadd_rn(q, __dmul_rn((double)xf[k + 1], (double)col[(size_t)(k + 1) * s.num_sv]));
						q = __dadd_rn(q, __dmul_rn((double)xf[k + 2], (double)col[(size_t)(k + 2) * s.num_sv]));
						q = __dadd_rn(q, __dmul_rn((double)xf[k + 3], (double)col[(size_t)(k + 3) * s.num_sv]));
						r = __dadd_rn(r, q);
					}
					for (; k < s.dim; ++k) r = __dadd_rn(r, __dmul_rn((double)xf[k], (double)col[(size_t)k * s.num_sv]));
					kv = r;
				}
			}
			if (s.kernel == FDB_KERNEL_POLYNOMIAL) {                          /* PolynomialKernel.hpp:38-40,62-70 */
				double tmp = __dadd_rn(__dmul_rn(s.poly_alpha, kv), s.poly_constant), ret = 1.0;
				for (int t = s.poly_degree; t > 0; t /= 2) {
					if (t % 2 == 1) ret = __dmul_rn(ret, tmp);
					tmp = __dmul_rn(tmp, tmp);
				}
				kv = ret;
			}
			s_prod[i] = __dmul_rn((double)s.coef[sv], kv);                   /* SvmClassifier.cpp:58 / RvmClassifier.cpp:98 (coef = c[l][l]) */
		}
		__syncthreads();
		if (tid == 0) {
			if (!rvm) {
				for (int i = 0; i < cnt; ++i) distance = __dadd_rn(distance, s_prod[i]);
			} else {
				for (int i = 0; i < cnt && !done; ++i) {
					distance = __dadd_rn(distance, s_prod[i]);
					level = base + i;
					done = !(distance >= (double)s.rvm_thresholds[level] && level + 1 < s.rvm_filters); /* RvmClassifier.cpp:84 */
				}
			}
		}
		__syncthreads();
	}
	if (tid == 0) {
		distance_out[item] = distance;
		if (rvm && level_out) level_out[item] = level;
	}
}

/* RBF SVM on float32 feature vectors, SVMB_W vectors per CTA: svm_kernel<2> streams the whole model (num_sv x dim floats,
 * 1.6 MB for 1024 vectors of 400) from L2 once per classified vector, which bounds the `single` detector in a feature space
 * (every window of every frame is classified) by L2 bandwidth. Here a thread owns one support vector of the chunk and keeps
 * SVMB_W sums of squared differences, so every model element that arrives is used for SVMB_W windows. Arithmetic and order are
 * those of svm_kernel<2>: float32 sequential SSD per (vector, support vector) pair (RbfKernel.hpp:97-108), exp in double,
 * coefficient product, then the double sum over the support vectors in their order (SvmClassifier.cpp:55-60). */
#define SVMB_W 8
__global__ void __launch_bounds__(SVM_THREADS) svm_f32_rbf_block_kernel(const DevSvm s, const float* __restrict__ vectors, int n,
		double* __restrict__ distance_out) {
	extern __shared__ __align__(16) unsigned char svmb_smem[];
	double* s_prod = reinterpret_cast<double*>(svmb_smem);                                   /* [SVMB_W][SVM_THREADS] */
	float* s_x = reinterpret_cast<float*>(svmb_smem + sizeof(double) * SVMB_W * SVM_THREADS);  /* [dim][SVMB_W] */
	const int tid = threadIdx.x;
	const int v0 = blockIdx.x * SVMB_W;
	const int nw = min(SVMB_W, n - v0);
	for (int i = tid; i < s.dim * SVMB_W; i += SVM_THREADS) {
		const int k = i / SVMB_W, w = i - k * SVMB_W;
		s_x[i] = w < nw ? vectors[(int64_t)(v0 + w) * s.dim + k] : 0.f;
	}
	__syncthreads();
	double distance = -(double)s.bias; /* thread w < nw: the hyperplane distance of vector v0 + w */
	for (int base = 0; base < s.num_sv; base += SVM_THREADS) {
		const int cnt = min(SVM_THREADS, s.num_sv - base);
		if (tid < cnt) {
			const int sv = base + tid;
			const float* __restrict__ col = s.sv_f32 + sv;
			float sum[SVMB_W];
#pragma unroll
			for (int w = 0; w < SVMB_W; ++w) sum[w] = 0.f;
			for (int k = 0; k < s.dim; ++k) {
				const float c = col[(size_t)k * s.num_sv];
				const float4 xa = *reinterpret_cast<const float4*>(s_x + k * SVMB_W), xb = *reinterpret_cast<const float4*>(s_x + k * SVMB_W + 4);
				const float x[SVMB_W] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
				for (int w = 0; w < SVMB_W; ++w) {
					const float diff = __fsub_rn(x[w], c);
					sum[w] = __fadd_rn(sum[w], __fmul_rn(diff, diff));
				}
			}
			const double coef = (double)s.coef[sv];
#pragma unroll
			for (int w = 0; w < SVMB_W; ++w) s_prod[w * SVM_THREADS + tid] = __dmul_rn(coef, exp(__dmul_rn(-s.gamma, (double)sum[w])));
		}
		__syncthreads();
		if (tid < nw) for (int i = 0; i < cnt; ++i) distance = __dadd_rn(distance, s_prod[tid * SVM_THREADS + i]);
		__syncthreads();
	}
	if (tid < nw) distance_out[v0 + tid] = distance;
}

/* RBF SVM on HistEq64 windows of frames, SVMU_W windows per CTA: svm_kernel<0> streams the whole model (num_sv x nwords words,
 * 590 KB for 1024 vectors of 24 x 24) from L2 once per stage-1 survivor - 177 GB per 256-frame step of the 15 landmark detectors.
 * Here warp w equalises window w of the block (HistEq64Filter.cpp:32-125, the sequential float32 cumulative histogram on its lane
 * 0), then a thread owns one support vector of the chunk and keeps SVMU_W integer sums of squared differences, so every model
 * word that arrives is used for SVMU_W windows. Arithmetic and order are those of svm_kernel<0>: exact integer SSD
 * (RbfKernel.hpp:39,78-88), exp in double, coefficient product, the double sum over the support vectors in their order
 * (SvmClassifier.cpp:55-60) - the distances are bit-identical. */
#define SVMU_W 8
static_assert(SVMU_W * 32 == SVM_THREADS, "one warp per window of the block");
__global__ void __launch_bounds__(SVM_THREADS) svm_u8_rbf_block_kernel(const DevSvm s, int patch_w, int patch_h,
		const uint8_t* __restrict__ frames, int W, int H, const uint8_t* __restrict__ arena, int64_t arena_stride,
		const DevLayer* __restrict__ layers, const SvmItem* __restrict__ items, int n_items, double* __restrict__ distance_out) {
	extern __shared__ __align__(16) unsigned char svmu_smem[];
	double* s_prod = reinterpret_cast<double*>(svmu_smem);                                        /* [SVMU_W][SVM_THREADS] */
	uint32_t* s_x = reinterpret_cast<uint32_t*>(svmu_smem + sizeof(double) * SVMU_W * SVM_THREADS); /* [nwords][SVMU_W] */
	__shared__ uint32_t s_hist[SVMU_W][64];
	__shared__ uint8_t s_eq[SVMU_W][64];
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const int v0 = blockIdx.x * SVMU_W;
	const int nw = min(SVMU_W, n_items - v0);
	for (int i = tid; i < s.nwords * SVMU_W; i += SVM_THREADS) s_x[i] = 0;
	s_hist[w][lane] = 0; s_hist[w][lane + 32] = 0;
	__syncthreads();
	if (w < nw) { /* warp w: window v0 + w */
		const SvmItem it = items[v0 + w];
		const DevLayer L = layers[it.layer];
		const uint8_t* img = (L.offset < 0 ? frames + (int64_t)it.frame * W * H
				: arena + (int64_t)it.frame * arena_stride + L.offset) + (int64_t)it.y * L.pitch + it.x;
		const int npix = patch_w * patch_h;
		for (int i = lane; i < npix; i += 32) {
			const int r = i / patch_w, c = i - r * patch_w;
			atomicAdd(&s_hist[w][img[(int64_t)r * L.pitch + c] >> 2], 1u);
		}
		__syncwarp();
		if (lane == 0) { /* sequential float cumsum, HistEq64Filter.cpp:70-87,97 */
			const float stretch = __fdiv_rn(255.0f, (float)npix);
			float cdf = 0.f;
			for (int b = 0; b < 64; ++b) {
				cdf = __fadd_rn(cdf, __fmul_rn((float)s_hist[w][b], stretch));
				const float fl = floorf(cdf);
				s_eq[w][b] = (uint8_t)((int)fl + (__fsub_rn(cdf, fl) >= 0.5f ? 1 : 0));
			}
		}
		__syncwarp();
		for (int i = lane; i < npix; i += 32) {
			const int r = i / patch_w, c = i - r * patch_w;
			const uint32_t e = s_eq[w][img[(int64_t)r * L.pitch + c] >> 2];
			atomicOr(&s_x[(i >> 2) * SVMU_W + w], e << (8 * (i & 3)));
		}
	}
	__syncthreads();
	double distance = -(double)s.bias; /* thread t < nw: the hyperplane distance of window v0 + t (SvmClassifier.cpp:56) */
	for (int base = 0; base < s.num_sv; base += SVM_THREADS) {
		const int cnt = min(SVM_THREADS, s.num_sv - base);
		if (tid < cnt) {
			const int sv = base + tid;
			const uint32_t* __restrict__ col = s.sv_words + sv;
			uint32_t acc[SVMU_W];
#pragma unroll
			for (int k = 0; k < SVMU_W; ++k) acc[k] = 0;
			for (int j = 0; j < s.nwords; ++j) {
				const uint32_t c = col[(size_t)j * s.num_sv];
				const uint4 xa = *reinterpret_cast<const uint4*>(s_x + j * SVMU_W), xb = *reinterpret_cast<const uint4*>(s_x + j * SVMU_W + 4);
				const uint32_t x[SVMU_W] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
				for (int k = 0; k < SVMU_W; ++k) {
					const uint32_t d = __vabsdiffu4(x[k], c);
					acc[k] = __dp4a(d, d, acc[k]);
				}
			}
			const double coef = (double)s.coef[sv];
#pragma unroll
			for (int k = 0; k < SVMU_W; ++k)
				s_prod[k * SVM_THREADS + tid] = __dmul_rn(coef, exp(__dmul_rn(-s.gamma, (double)(int)acc[k])));      /* SvmClassifier.cpp:58 */
		}
		__syncthreads();
		if (tid < nw) for (int i = 0; i < cnt; ++i) distance = __dadd_rn(distance, s_prod[tid * SVM_THREADS + i]);
		__syncthreads();
	}
	if (tid < nw) distance_out[v0 + tid] = distance;
}

static size_t svmu_smem_bytes(const DevSvm& s) { return sizeof(double) * SVMU_W * SVM_THREADS + sizeof(uint32_t) * (size_t)s.nwords * SVMU_W; }

static size_t svmb_smem_bytes(const DevSvm& s) { return sizeof(double) * SVMB_W * SVM_THREADS + sizeof(float) * (size_t)s.dim * SVMB_W; }

/* HistEq64 patches (HistEq64Filter.cpp:32-125) of a list of windows -> [n][patch_w * patch_h] u8: the patch data of
 * DirectPyramidFeatureExtractor::extract(x, y, width, height) (DirectPyramidFeatureExtractor.cpp:67-73,133-147) for
 * sparse callers (condensation::WvmSvmModel). One CTA of 128 threads per window. */
__global__ void __launch_bounds__(128) hq64_items_kernel(int patch_w, int patch_h, const uint8_t* __restrict__ frames, int W, int H,
		const uint8_t* __restrict__ arena, int64_t arena_stride, const DevLayer* __restrict__ layers,
		const SvmItem* __restrict__ items, uint8_t* __restrict__ out) {
	__shared__ uint32_t s_hist[64];
	__shared__ uint8_t s_eq[64];
	const int tid = threadIdx.x;
	const SvmItem it = items[blockIdx.x];
	const DevLayer L = layers[it.layer];
	const uint8_t* img = (L.offset < 0 ? frames + (int64_t)it.frame * W * H
			: arena + (int64_t)it.frame * arena_stride + L.offset) + (int64_t)it.y * L.pitch + it.x;
	const int npix = patch_w * patch_h;
	if (tid < 64) s_hist[tid] = 0;
	__syncthreads();
	for (int i = tid; i < npix; i += 128) {
		const int r = i / patch_w, c = i - r * patch_w;
		atomicAdd(&s_hist[img[(int64_t)r * L.pitch + c] >> 2], 1u);
	}
	__syncthreads();
	if (tid == 0) { /* sequential float cumsum, HistEq64Filter.cpp:70-87,97 */
		const float stretch = __fdiv_rn(255.0f, (float)npix);
		float cdf = 0.f;
		for (int b = 0; b < 64; ++b) {
			cdf = __fadd_rn(cdf, __fmul_rn((float)s_hist[b], stretch));
			const float fl = floorf(cdf);
			s_eq[b] = (uint8_t)((int)fl + (__fsub_rn(cdf, fl) >= 0.5f ? 1 : 0));
		}
	}
	__syncthreads();
	uint8_t* dst = out + (int64_t)blockIdx.x * npix;
	for (int i = tid; i < npix; i += 128) {
		const int r = i / patch_w, c = i - r * patch_w;
		dst[i] = s_eq[img[(int64_t)r * L.pitch + c] >> 2];
	}
}

void launch_hq64_items(cudaStream_t st, int patch_w, int patch_h, const uint8_t* frames, int W, int H, const uint8_t* arena,
		int64_t arena_stride, const DevLayer* layers, const SvmItem* items, int n_items, uint8_t* out) {
	if (n_items == 0) return;
	hq64_items_kernel<<<(unsigned)n_items, 128, 0, st>>>(patch_w, patch_h, frames, W, H, arena, arena_stride, layers, items, out);
}

static size_t svm_smem_bytes(const DevSvm& s) {
	size_t xbytes = s.sv_type == FDB_SV_F32 ? sizeof(float) * (size_t)s.dim : sizeof(uint32_t) * (size_t)s.nwords;
	return sizeof(double) * SVM_CHUNK + ((xbytes + 15) / 16) * 16;
}

int svm_configure() {
	cudaError_t e = cudaFuncSetAttribute(svm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(svm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(svm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(svm_f32_rbf_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(svm_u8_rbf_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
	return (int)e;
}

void launch_svm_windows(cudaStream_t st, const DevSvm& s, int patch_w, int patch_h, const uint8_t* frames, int W, int H,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, const SvmItem* items, int n_items,
		double* distance_out, int* level_out) {
	if (n_items == 0) return;
	if (s.sv_type != FDB_SV_F32 && s.kernel == FDB_KERNEL_RBF && s.rvm_filters == 0 && !level_out && n_items >= 2 * SVMU_W
			&& svmu_smem_bytes(s) <= 128 * 1024) {
		svm_u8_rbf_block_kernel<<<(unsigned)((n_items + SVMU_W - 1) / SVMU_W), SVM_THREADS, svmu_smem_bytes(s), st>>>(s, patch_w, patch_h,
				frames, W, H, arena, arena_stride, layers, items, n_items, distance_out);
		return;
	}
	svm_kernel<0><<<(unsigned)n_items, SVM_THREADS, svm_smem_bytes(s), st>>>(s, patch_w, patch_h, frames, W, H, arena,
			arena_stride, layers, items, nullptr, distance_out, level_out);
}

void launch_svm_vectors(cudaStream_t st, const DevSvm& s, const void* vectors, int n, double* distance_out, int* level_out) {
	if (n == 0) return;
	if (s.sv_type == FDB_SV_F32 && s.kernel == FDB_KERNEL_RBF && s.rvm_filters == 0 && n >= 2 * SVMB_W && svmb_smem_bytes(s) <= 128 * 1024) {
		svm_f32_rbf_block_kernel<<<(unsigned)((n + SVMB_W - 1) / SVMB_W), SVM_THREADS, svmb_smem_bytes(s), st>>>(s, reinterpret_cast<const float*>(vectors), n, distance_out);
		return;
	}
	if (s.sv_type == FDB_SV_F32)
		svm_kernel<2><<<(unsigned)n, SVM_THREADS, svm_smem_bytes(s), st>>>(s, 0, 0, nullptr, 0, 0, nullptr, 0, nullptr,
				nullptr, vectors, distance_out, level_out);
	else
		svm_kernel<1><<<(unsigned)n, SVM_THREADS, svm_smem_bytes(s), st>>>(s, 0, 0, nullptr, 0, 0, nullptr, 0, nullptr,
				nullptr, vectors, distance_out, level_out);
}

} // namespace fdb
