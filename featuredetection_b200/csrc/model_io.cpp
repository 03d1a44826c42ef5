/*
 * model_io.cpp - host-side reader of the reference's own SVM text container.
 *
 * Product code (C++). Restates the stream format of
 *   SvmClassifier::store / SvmClassifier::load(std::ifstream&)              libClassification/src/classification/SvmClassifier.cpp:68-158
 *   loadSupportVectors<T>                                                   libClassification/include/classification/SvmClassifier.hpp:151-163
 *   ProbabilisticSvmClassifier::store / load(std::ifstream&)  ("Logistic a b")   ProbabilisticSvmClassifier.cpp:65-78
 * so that a model written by the reference (e.g. by the trackers' TrainableProbabilisticSvmClassifier) loads
 * into fdb_svm_create. Only the RBF kernel is evaluated on the GPU; other kernel lines are parsed and
 * reported as FDB_ERR_UNSUPPORTED. Like the reference, CV_8U support vectors are stored as raw
 * characters (operator<< on uchar), i.e. the loader reads one non-whitespace character per element.
 */
#include <fstream>
#include <string>
#include <vector>

#include "fdb_internal.h"

struct fdb_svm_file {
	fdb_svm_desc desc;
	std::vector<float> coefficients;
	std::vector<uint8_t> sv_u8;
	std::vector<float> sv_f32;
	int rows = 0, cols = 0, channels = 0, depth = 0;
};

using namespace fdb;

extern "C" {

int fdb_svm_file_load(const char* path, fdb_svm_file** out) {
	if (!path || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	std::ifstream file(path);
	if (!file) return fail(FDB_ERR_RUNTIME, "SvmClassifier: Cannot read from stream");
	fdb_svm_file* f = new fdb_svm_file;
	fdb_svm_desc& d = f->desc;
	d = fdb_svm_desc();
	std::string tmp, kernelType;
	file >> tmp >> kernelType; /* "Kernel" <type> */
	bool supported = false;
	if (kernelType == "RBF") {
		file >> d.gamma;
		d.kernel = FDB_KERNEL_RBF;
		supported = true;
	} else if (kernelType == "Polynomial") {
		int degree; double constant, scale;
		file >> degree >> constant >> scale;
	} else if (kernelType != "Linear" && kernelType != "HIK") {
		delete f;
		return fail(FDB_ERR_RUNTIME, "SvmClassifier: Invalid kernel type: " + kernelType);
	}
	file >> tmp >> d.bias; /* "Bias" */
	size_t count = 0;
	file >> tmp >> count;   /* "Coefficients" */
	if (!file || count == 0 || count > (1u << 24)) { delete f; return fail(FDB_ERR_RUNTIME, "SvmClassifier: Invalid classifier file"); }
	f->coefficients.resize(count);
	for (size_t i = 0; i < count; ++i) file >> f->coefficients[i];
	file >> tmp >> count >> f->rows >> f->cols >> f->channels >> f->depth; /* "SupportVectors" */
	if (!file || count != f->coefficients.size() || f->rows < 1 || f->cols < 1 || f->channels < 1) {
		delete f; return fail(FDB_ERR_RUNTIME, "SvmClassifier: Invalid classifier file");
	}
	const size_t dim = (size_t)f->rows * f->cols * f->channels;
	if (f->depth == 0) { /* CV_8U: raw characters */
		f->sv_u8.resize(count * dim);
		for (size_t i = 0; i < count * dim; ++i) { unsigned char c; file >> c; f->sv_u8[i] = c; }
		d.sv_type = FDB_SV_U8; d.support_vectors = f->sv_u8.data();
	} else if (f->depth == 5) { /* CV_32F */
		f->sv_f32.resize(count * dim);
		for (size_t i = 0; i < count * dim; ++i) file >> f->sv_f32[i];
		d.sv_type = FDB_SV_F32; d.support_vectors = f->sv_f32.data();
	} else {
		delete f;
		return fail(FDB_ERR_UNSUPPORTED, "SvmClassifier: only CV_8U and CV_32F support vectors are evaluated on the GPU");
	}
	if (!file) { delete f; return fail(FDB_ERR_RUNTIME, "SvmClassifier: Invalid classifier file"); }
	d.num_sv = (int32_t)count; d.dim = (int32_t)dim;
	d.coefficients = f->coefficients.data();
	d.threshold = 0.0f;                        /* VectorMachineClassifier default */
	d.logistic_a = 0.00556; d.logistic_b = -2.95; /* ProbabilisticSvmClassifier.hpp:36 defaults */
	if (file >> tmp && tmp == "Logistic") file >> d.logistic_a >> d.logistic_b;
	if (!supported) { delete f; return fail(FDB_ERR_UNSUPPORTED, "SvmClassifier: kernel " + kernelType + " is not evaluated on the GPU (RBF only)"); }
	*out = f;
	return FDB_OK;
}

const fdb_svm_desc* fdb_svm_file_desc(const fdb_svm_file* f) { return f ? &f->desc : nullptr; }

void fdb_svm_file_free(fdb_svm_file* f) { delete f; }

} // extern "C"
