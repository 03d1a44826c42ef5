/*
 * api_types.h - private object definitions and small CUDA helpers shared by api.cu (contexts,
 * classifiers) and detector.cu (the detector pipeline).
 */
#ifndef FDB_API_TYPES_H_
#define FDB_API_TYPES_H_

#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "fdb_internal.h"
#include "wvm_device.h"

#define CUDA_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
	return fdb::fail(FDB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } while (0)

struct fdb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	int64_t launches = 0;
	cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; /* stopwatch + per-kernel profile marks */
};

struct fdb_wvm {
	fdb_ctx* ctx = nullptr;
	fdb::DevWvm dev{};
	std::vector<void*> owned;
	std::vector<float> thresholds_from_file;
	float* d_thresholds = nullptr;
	float limit = 0.f;
	double logistic_a = 0, logistic_b = 0;
	std::vector<float> thresholds; /* host copy incl. limit */
};

struct fdb_svm {
	fdb_ctx* ctx = nullptr;
	fdb::DevSvm dev{};
	bool has_dense = false;     /* tensor-core form available (u8 support vectors, see svm_dense.cu) */
	fdb::DevSvmDense dense{};
	std::vector<void*> owned;
	double logistic_a = 0, logistic_b = 0;
	std::vector<float> rvm_thresholds; /* host copy (RVM only) */
};

/* RvmClassifier + ProbabilisticRvmClassifier: the same device form as the SVM (support vectors, kernel) plus the cascade */
struct fdb_rvm : fdb_svm {};

namespace fdb {

inline int check_ctx(fdb_ctx* c) {
	if (!c) return fail(FDB_ERR_INVALID_ARGUMENT, "null context");
	CUDA_TRY(cudaSetDevice(c->device));
	return FDB_OK;
}

inline void free_all(std::vector<void*>& dev, std::vector<void*>* host = nullptr) {
	for (void* p : dev) cudaFree(p);
	dev.clear();
	if (host) { for (void* p : *host) cudaFreeHost(p); host->clear(); }
}

template <class T>
int upload(const T* host, size_t n, T** dev, std::vector<void*>& owned) {
	*dev = nullptr;
	void* p = nullptr;
	CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)));
	owned.push_back(p);
	if (n) CUDA_TRY(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
	*dev = (T*)p;
	return FDB_OK;
}

template <class T>
int dev_alloc(T** out, size_t n, std::vector<void*>& owned) {
	void* p = nullptr;
	CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)));
	owned.push_back(p);
	*out = (T*)p;
	return FDB_OK;
}

template <class T>
int host_alloc(T** out, size_t n, std::vector<void*>& owned) {
	void* p = nullptr;
	CUDA_TRY(cudaMallocHost(&p, std::max<size_t>(n * sizeof(T), 16)));
	owned.push_back(p);
	*out = (T*)p;
	return FDB_OK;
}

} // namespace fdb
#endif
