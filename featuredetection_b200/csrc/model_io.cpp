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
#include <cstdlib>
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

/* ------------------------------------------------------------------------------------------------
 * SdmLandmarkModel::load (libSupervisedDescent/src/superviseddescent/SdmLandmarkModel.cpp:130-232): line-oriented text -
 * description, "numLandmarks n", n identifiers, 2n mean coordinates (all x, then all y), "numCascadeSteps s", and per step
 * "cascadeStep i rows r cols c", "descriptorType t", "descriptorPostprocessing p", "descriptorParameters ..." followed by
 * r lines of c floats. Only "vlhog-uoctti" with empty (face-size adaptive) parameters runs on the GPU.
 * ---------------------------------------------------------------------------------------------- */
struct fdb_sdm_file {
	fdb_sdm_desc desc;
	std::vector<float> mean;
	std::vector<std::vector<float>> regressors;
	std::vector<const float*> regressor_ptrs;
	std::vector<std::string> identifiers;
};

namespace {

bool sdm_getline(std::ifstream& file, std::string& line) {
	if (!std::getline(file, line)) return false;
	while (!line.empty() && line.back() == '\r') line.pop_back(); /* boost::trim_right_if(line, is_any_of("\r")) */
	return true;
}

std::vector<std::string> sdm_split(const std::string& line) { /* boost::split(..., is_any_of(" ")): empty tokens are kept */
	std::vector<std::string> out(1);
	for (char c : line) { if (c == ' ') out.emplace_back(); else out.back().push_back(c); }
	return out;
}

} // namespace

extern "C" {

int fdb_sdm_file_load(const char* path, fdb_sdm_file** out) {
	if (!path || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	std::ifstream file(path);
	if (!file.is_open()) return fail(FDB_ERR_RUNTIME, std::string("Given SDM model file could not be opened: ") + path);
	fdb_sdm_file* f = new fdb_sdm_file;
	auto bad = [&](const std::string& why) { delete f; return fail(FDB_ERR_RUNTIME, "SdmLandmarkModel::load: " + why); };
	try {
		std::string line;
		std::vector<std::string> tok;
		if (!sdm_getline(file, line)) return bad("empty file");          /* description */
		if (!sdm_getline(file, line)) return bad("missing numLandmarks");
		tok = sdm_split(line);
		if (tok.size() < 2) return bad("bad numLandmarks line");
		const int L = std::stoi(tok[1]);
		if (L < 1 || L > 4096) return bad("bad numLandmarks");
		for (int i = 0; i < L; ++i) { if (!sdm_getline(file, line)) return bad("truncated identifiers"); f->identifiers.push_back(line); }
		f->mean.resize(2 * (size_t)L);
		for (int i = 0; i < 2 * L; ++i) { if (!sdm_getline(file, line)) return bad("truncated mean"); f->mean[i] = std::stof(line); }
		if (!sdm_getline(file, line)) return bad("missing numCascadeSteps");
		tok = sdm_split(line);
		if (tok.size() < 2) return bad("bad numCascadeSteps line");
		const int steps = std::stoi(tok[1]);
		if (steps < 1 || steps > 64) return bad("bad numCascadeSteps");
		for (int s = 0; s < steps; ++s) {
			if (!sdm_getline(file, line)) return bad("missing cascadeStep header");
			tok = sdm_split(line);
			if (tok.size() < 6) return bad("bad cascadeStep header");
			const int rows = std::stoi(tok[3]), cols = std::stoi(tok[5]);
			if (!sdm_getline(file, line)) return bad("missing descriptorType");
			tok = sdm_split(line);
			const std::string type = tok.size() > 1 ? tok[1] : std::string();
			if (!sdm_getline(file, line)) return bad("missing descriptorPostprocessing");
			if (!sdm_getline(file, line)) return bad("missing descriptorParameters");
			tok = sdm_split(line);
			if (type != "vlhog-uoctti") {
				delete f;
				if (type == "OpenCVSift" || type == "vlhog-dt") return fail(FDB_ERR_UNSUPPORTED, "descriptorType " + type + " is not evaluated on the GPU (vlhog-uoctti only)");
				return fail(FDB_ERR_RUNTIME, "descriptorType does not match 'OpenCVSift', 'vlhog-dt' or 'vlhog-uoctti'.");
			}
			if (tok.size() == 7) { delete f; return fail(FDB_ERR_UNSUPPORTED, "vlhog-uoctti with fixed numCells/cellSize/numBins is not evaluated on the GPU (adaptive parameters only)"); }
			if (tok.size() != 2) return bad("descriptorParameters must either be empty (=face-size adaptive parameters) or contain numCells, cellSize and numBins.");
			if (rows != L * 279 + 1 || cols != 2 * L) return bad("regressor size does not match numLandmarks * 279 + 1 x 2 * numLandmarks");
			std::vector<float> R((size_t)rows * cols);
			for (int j = 0; j < rows; ++j) {
				if (!sdm_getline(file, line)) return bad("truncated regressor data");
				const char* p = line.c_str();
				for (int c = 0; c < cols; ++c) {
					char* end = nullptr;
					R[(size_t)j * cols + c] = std::strtof(p, &end);
					if (end == p) return bad("bad float in regressor data");
					p = end;
				}
			}
			f->regressors.push_back(std::move(R));
		}
		for (auto& r : f->regressors) f->regressor_ptrs.push_back(r.data());
		f->desc.num_landmarks = L;
		f->desc.num_cascade_steps = steps;
		f->desc.mean_landmarks = f->mean.data();
		f->desc.regressors = f->regressor_ptrs.data();
	} catch (const std::exception& e) { /* boost::bad_lexical_cast in the reference */
		return bad(std::string("bad number: ") + e.what());
	}
	*out = f;
	return FDB_OK;
}

const fdb_sdm_desc* fdb_sdm_file_desc(const fdb_sdm_file* f) { return f ? &f->desc : nullptr; }
void fdb_sdm_file_free(fdb_sdm_file* f) { delete f; }

} // extern "C"
