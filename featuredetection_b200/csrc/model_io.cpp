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
#include <memory>
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

int fdb_svm_file_load(const char* path, fdb_svm_file** out) try {
	if (!path || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	std::ifstream file(path);
	if (!file) return fail(FDB_ERR_RUNTIME, "SvmClassifier: Cannot read from stream");
	std::unique_ptr<fdb_svm_file> f(new fdb_svm_file);
	fdb_svm_desc& d = f->desc;
	d = fdb_svm_desc();
	std::string tmp, kernelType;
	file >> tmp >> kernelType; /* "Kernel" <type> */
	if (kernelType == "RBF") {
		file >> d.gamma;
		d.kernel = FDB_KERNEL_RBF;
	} else if (kernelType == "Polynomial") { /* SvmClassifier.cpp:76-77: degree, constant, alpha */
		file >> d.poly_degree >> d.poly_constant >> d.poly_alpha;
		d.kernel = FDB_KERNEL_POLYNOMIAL;
	} else if (kernelType == "Linear") {
		d.kernel = FDB_KERNEL_LINEAR;
	} else if (kernelType == "HIK") {
		d.kernel = FDB_KERNEL_HIK;
	} else {
		return fail(FDB_ERR_RUNTIME, "SvmClassifier: Invalid kernel type: " + kernelType);
	}
	file >> tmp >> d.bias; /* "Bias" */
	size_t count = 0;
	file >> tmp >> count;   /* "Coefficients" */
	if (!file || count == 0 || count > (1u << 24)) return fail(FDB_ERR_RUNTIME, "SvmClassifier: Invalid classifier file");
	f->coefficients.resize(count);
	for (size_t i = 0; i < count; ++i) file >> f->coefficients[i];
	file >> tmp >> count >> f->rows >> f->cols >> f->channels >> f->depth; /* "SupportVectors" */
	if (!file || count != f->coefficients.size() || f->rows < 1 || f->cols < 1 || f->channels < 1)
		return fail(FDB_ERR_RUNTIME, "SvmClassifier: Invalid classifier file");
	/* the sizes come from the file: bound them before they size an allocation (fdb_svm_create's own limit on a vector is
	 * 96 KB of float32 = 24 576 elements; the element count of all vectors must fit the stream that follows) */
	const size_t max_dim = 96 * 1024 / 4;
	if ((size_t)f->rows > max_dim || (size_t)f->cols > max_dim || (size_t)f->channels > max_dim ||
			(size_t)f->rows * f->cols > max_dim || (size_t)f->rows * f->cols * f->channels > max_dim)
		return fail(FDB_ERR_RUNTIME, "SvmClassifier: Invalid classifier file (support vector of more than 24576 elements)");
	const size_t dim = (size_t)f->rows * f->cols * f->channels;
	if (count > ((size_t)1 << 31) / dim) return fail(FDB_ERR_RUNTIME, "SvmClassifier: Invalid classifier file (support vectors exceed 2^31 elements)");
	if (f->depth == 0) { /* CV_8U: raw characters */
		f->sv_u8.resize(count * dim);
		for (size_t i = 0; i < count * dim && file; ++i) { unsigned char c; file >> c; f->sv_u8[i] = c; }
		d.sv_type = FDB_SV_U8; d.support_vectors = f->sv_u8.data();
	} else if (f->depth == 5) { /* CV_32F */
		f->sv_f32.resize(count * dim);
		for (size_t i = 0; i < count * dim && file; ++i) file >> f->sv_f32[i];
		d.sv_type = FDB_SV_F32; d.support_vectors = f->sv_f32.data();
	} else {
		return fail(FDB_ERR_UNSUPPORTED, "SvmClassifier: only CV_8U and CV_32F support vectors are evaluated on the GPU");
	}
	if (!file) return fail(FDB_ERR_RUNTIME, "SvmClassifier: Invalid classifier file");
	d.num_sv = (int32_t)count; d.dim = (int32_t)dim;
	d.coefficients = f->coefficients.data();
	d.threshold = 0.0f;                        /* VectorMachineClassifier default */
	d.logistic_a = 0.00556; d.logistic_b = -2.95; /* ProbabilisticSvmClassifier.hpp:36 defaults */
	if (file >> tmp && tmp == "Logistic") file >> d.logistic_a >> d.logistic_b;
	*out = f.release();
	return FDB_OK;
} FDB_API_CATCH

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

int fdb_sdm_file_load(const char* path, fdb_sdm_file** out) try {
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
} FDB_API_CATCH

const fdb_sdm_desc* fdb_sdm_file_desc(const fdb_sdm_file* f) { return f ? &f->desc : nullptr; }
void fdb_sdm_file_free(fdb_sdm_file* f) { delete f; }

} // extern "C"

/* ------------------------------------------------------------------------------------------------
 * MATLAB classifier files (the format every ffpDetectApp .cfg points at: classifierFile / thresholdsFile)
 *   WvmClassifier::loadFromMatlab                            libClassification/src/classification/WvmClassifier.cpp:348-770
 *   ProbabilisticWvmClassifier::loadSigmoidParamsFromMatlab  ProbabilisticWvmClassifier.cpp:95-137
 *   SvmClassifier::loadFromMatlab                            SvmClassifier.cpp:240-335
 *   ProbabilisticSvmClassifier::loadSigmoidParamsFromMatlab  ProbabilisticSvmClassifier.cpp:114-162
 * restated over matfile.cpp (the reference goes through MATLAB's libmat). The unit conversions are the loader's:
 * grey values x 255 (:644), app_rsv_convol x 65025 (:696), basisParam / 65025 (:555), support vectors (uchar)(255 v)
 * (SvmClassifier.cpp:308), posterior vectors are {B, A} (ProbabilisticWvmClassifier.cpp:123-124).
 * Where the reference would run on with uninitialised memory (a missing weight_hk%d or param_nonlin1) this loader fails.
 * ---------------------------------------------------------------------------------------------- */
#include "matfile.h"

struct fdb_wvm_file {
	fdb_wvm_desc desc;
	std::vector<float> lin_thresholds, hk_weights, hierarchical_thresholds;
	std::vector<double> app_rsv_convol, area_val;
	std::vector<int32_t> area_cntval, area_cntrec;
	std::vector<fdb_rect4> area_rec;
};

struct fdb_rvm_file {
	fdb_rvm_desc desc;
	std::vector<float> support_vectors, coefficients, hierarchical_thresholds;
};

namespace {

const MatArray* mat_var(const MatFile& f, const std::string& name) { return f.get(name); }

/* the two dimensions the loaders read (mxGetDimensions()[0], [1]) */
int dim0(const MatArray& a) { return a.dims.size() > 0 ? a.dims[0] : 0; }
int dim1(const MatArray& a) { return a.dims.size() > 1 ? a.dims[1] : 0; }

int load_posterior(const char* path, const char* var, bool required, double* A, double* B) {
	MatFile f; std::string err;
	if (!mat_read(path, &f, &err)) return fail(required ? FDB_ERR_INVALID_ARGUMENT : FDB_ERR_RUNTIME, "Unable to open the thresholds/logistic file: " + err);
	const MatArray* p = mat_var(f, var);
	*A = 0; *B = 0;
	if (!p || dim1(*p) != 2 || p->real.size() < 2) {
		if (required) return fail(FDB_ERR_RUNTIME, std::string("Unable to find the vector ") + var + " (of size 2). If you don't want probabilistic output, don't use a probabilistic classifier.");
		return FDB_OK; /* ProbabilisticSvmClassifier.cpp:139-151: warning, A = B = 0 */
	}
	*B = p->real[0]; *A = p->real[1];
	return FDB_OK;
}

} // namespace

extern "C" {

int fdb_wvm_file_load(const char* classifier_path, const char* thresholds_path, fdb_wvm_file** out) try {
	if (!classifier_path || !thresholds_path || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	MatFile mf; std::string err;
	if (!mat_read(classifier_path, &mf, &err)) return fail(FDB_ERR_INVALID_ARGUMENT, "WvmClassifier: Could not open the provided classifier filename: " + err);
	const MatArray* a = mat_var(mf, "num_hk");
	if (!a || a->real.empty()) return fail(FDB_ERR_RUNTIME, "WvmClassifier: There is a no num_hk in the classifier file.");
	const int nfilter = (int)a->real[0];
	if (nfilter < 1 || nfilter > 4096) return fail(FDB_ERR_RUNTIME, "WvmClassifier: bad num_hk");
	if (mat_var(mf, "wrvm")) return fail(FDB_ERR_RUNTIME, "WvmClassifier: Reading all wvm filters at once using the structure 'wrvm' is not (yet) supported.");
	a = mat_var(mf, "support_hk1");
	if (!a) return fail(FDB_ERR_RUNTIME, "WvmClassifier: Unable to find the matrix 'support_hk1' in the classifier file.");
	if (a->dims.size() != 2) return fail(FDB_ERR_RUNTIME, "WvmClassifier: The matrix 'support_hk' in the classifier file should have 2 dimensions.");
	std::unique_ptr<fdb_wvm_file> f(new fdb_wvm_file);
	fdb_wvm_desc& d = f->desc;
	d = fdb_wvm_desc();
	d.filter_size_y = dim0(*a); d.filter_size_x = dim1(*a); /* :500-505 */
	d.num_lin_filters = nfilter;
	f->hk_weights.assign((size_t)nfilter * (nfilter + 1) / 2, 0.f);
	for (int i = 0; i < nfilter; ++i) {
		const std::string idx = std::to_string(i + 1);
		a = mat_var(mf, "support_hk" + idx); /* the vectors themselves are not used by the rectangle evaluator */
		if (!a) return fail(FDB_ERR_RUNTIME, "WvmClassifier: Unable to find the matrix 'support_hk" + idx + "' in the classifier file.");
		if (a->dims.size() != 2) return fail(FDB_ERR_RUNTIME, "WvmClassifier: The matrix 'filter" + idx + "' in the classifier file should have 2 dimensions.");
		a = mat_var(mf, "weight_hk" + idx);
		if (!a) return fail(FDB_ERR_RUNTIME, "WvmClassifier: Unable to find the matrix 'weight_hk" + idx + "' (the reference would continue with uninitialised weights)");
		if (dim1(*a) != i + 1 && dim0(*a) != i + 1)
			return fail(FDB_ERR_RUNTIME, "WvmClassifier: The matrix weight_hk" + idx + " in the classifier file should have a dimensions 1x" + idx + " or " + idx + "x1");
		if (a->real.size() < (size_t)i + 1) /* right dims, but a cell / struct / sparse / empty array: no numeric data to read */
			return fail(FDB_ERR_RUNTIME, "WvmClassifier: The matrix weight_hk" + idx + " in the classifier file is not a numeric array of " + idx + " elements");
		for (int j = 0; j <= i; ++j) f->hk_weights[(size_t)i * (i + 1) / 2 + j] = (float)a->real[j];
	}
	a = mat_var(mf, "param_nonlin1_rvm");
	if (!a) a = mat_var(mf, "param_nonlin1");
	if (!a || a->real.size() < 3) return fail(FDB_ERR_RUNTIME, "WvmClassifier: neither 'param_nonlin1_rvm' nor 'param_nonlin1' found (the reference would continue with an uninitialised bias)");
	const float bias = (float)a->real[0];
	d.basis_param = (float)(a->real[2] / 65025.0);
	f->lin_thresholds.assign((size_t)nfilter, bias);
	a = mat_var(mf, "num_hk_wvm");
	if (!a || a->real.empty()) return fail(FDB_ERR_RUNTIME, "WvmClassifier: Variable 'num_hk_wvm' not found in classifier file.");
	d.num_filters_per_level = (int)a->real[0];
	a = mat_var(mf, "num_lev_wvm");
	if (!a || a->real.empty()) return fail(FDB_ERR_RUNTIME, "WvmClassifier: Variable 'num_lev_wvm' not found in classifier file.");
	d.num_levels = (int)a->real[0];
	a = mat_var(mf, "area");
	if (!a || !a->is_struct()) return fail(FDB_ERR_RUNTIME, "WvmClassifier: 'area' not found (right *.mat/kernel?)");
	const int nHK = dim1(*a);
	if (nHK != nfilter || d.num_filters_per_level * d.num_levels != nHK)
		return fail(FDB_ERR_RUNTIME, "WvmClassifier: Variable 'area' in the classifier file has wrong dimensions:" + std::to_string(nHK) + "(==" + std::to_string(nfilter) + ")");
	for (int h = 0; h < nHK; ++h) {
		const MatArray* mval = a->field(h, "val_u");
		if (!mval) return fail(FDB_ERR_RUNTIME, "WvmClassifier: 'val_u' not found (WVM: 'val_u', else: 'val', right *.mat/kernel?)");
		const int cntval = dim1(*mval);
		const MatArray* mcnt = a->field(h, "cntrec_u");
		const MatArray* mrec = a->field(h, "crec");
		if (!mcnt || (int)mcnt->real.size() < cntval || (int)mval->real.size() < cntval) return fail(FDB_ERR_RUNTIME, "WvmClassifier: 'cntrec_u' missing or shorter than 'val_u'");
		f->area_cntval.push_back(cntval);
		for (int v = 0; v < cntval; ++v) {
			f->area_val.push_back(mval->real[v] * 255.0F); /* :644 */
			f->area_cntrec.push_back((int)mcnt->real[v]);
		}
		for (int v = 1; v < cntval; ++v) /* descriptor order (f, v, r), v >= 1: the evaluator never reads the rectangles of v == 0 (:277) */
			for (int r = 0; r < (int)mcnt->real[v]; ++r) {
				if (!mrec || !mrec->is_struct()) return fail(FDB_ERR_RUNTIME, "WvmClassifier: 'crec' not found in 'area'");
				const int64_t e = (int64_t)r * cntval + v; /* :653 */
				const MatArray* x1 = mrec->field(e, "x1"); const MatArray* y1 = mrec->field(e, "y1");
				const MatArray* x2 = mrec->field(e, "x2"); const MatArray* y2 = mrec->field(e, "y2");
				if (!x1 || !y1 || !x2 || !y2 || x1->real.empty() || y1->real.empty() || x2->real.empty() || y2->real.empty())
					return fail(FDB_ERR_RUNTIME, "WvmClassifier: rectangle " + std::to_string(r) + " of grey value " + std::to_string(v) + " of filter " + std::to_string(h + 1) + " is missing in 'crec'");
				fdb_rect4 rc;
				rc.x1 = (int)x1->real[0]; rc.y1 = (int)y1->real[0]; rc.x2 = (int)x2->real[0]; rc.y2 = (int)y2->real[0];
				f->area_rec.push_back(rc);
			}
	}
	a = mat_var(mf, "app_rsv_convol");
	if (!a) return fail(FDB_ERR_RUNTIME, "WvmClassifier: 'app_rsv_convol' not found.");
	if (dim1(*a) != nfilter || (int)a->real.size() < nfilter) return fail(FDB_ERR_RUNTIME, "WvmClassifier: 'app_rsv_convol' not right dim:" + std::to_string(dim1(*a)) + " (==" + std::to_string(nfilter) + ")");
	for (int h = 0; h < nfilter; ++h) f->app_rsv_convol.push_back(a->real[h] * 65025.0);

	MatFile tf;
	if (!mat_read(thresholds_path, &tf, &err)) return fail(FDB_ERR_RUNTIME, "WvmClassifier: Unable to open the thresholds file (wrong format?):" + err);
	a = mat_var(tf, "hierar_thresh");
	if (!a) return fail(FDB_ERR_RUNTIME, "WvmClassifier: Unable to find the matrix hierar_thresh in the thresholds file.");
	for (int o = 0; o < dim1(*a) && o < (int)a->real.size(); ++o) f->hierarchical_thresholds.push_back((float)a->real[o]);
	if ((int)f->hierarchical_thresholds.size() != nfilter)
		return fail(FDB_ERR_RUNTIME, "WvmClassifier: Something seems to be wrong, hierarchicalThresholdsFromFile.size() != numLinFilters; " +
				std::to_string(f->hierarchical_thresholds.size()) + "!=" + std::to_string(nfilter));
	a = mat_var(tf, "posterior_wrvm");
	if (!a) return fail(FDB_ERR_RUNTIME, "ProbabilisticWvmClassifier: Unable to find the vector posterior_wrvm. If you don't want probabilistic output, don't use a probabilistic classifier.");
	if (dim1(*a) != 2 || a->real.size() < 2) return fail(FDB_ERR_RUNTIME, "ProbabilisticWvmClassifier: Size of vector posterior_wrvm !=2. If you don't want probabilistic output, don't use a probabilistic classifier.");
	d.logistic_b = a->real[0]; d.logistic_a = a->real[1];

	d.num_used_filters = 280;          /* :358; clamped to num_lin_filters by setNumUsedFilters (:151-158) */
	d.limit_reliability_filter = 0.f;  /* :363; the cfg's "threshold" is applied by the caller (ProbabilisticWvmClassifier.cpp:82) */
	d.lin_thresholds = f->lin_thresholds.data();
	d.hk_weights = f->hk_weights.data();
	d.app_rsv_convol = f->app_rsv_convol.data();
	d.hierarchical_thresholds = f->hierarchical_thresholds.data();
	d.area_cntval = f->area_cntval.data();
	d.area_val = f->area_val.data();
	d.area_cntrec = f->area_cntrec.data();
	d.area_rec = f->area_rec.data();
	*out = f.release();
	return FDB_OK;
} FDB_API_CATCH

const fdb_wvm_desc* fdb_wvm_file_desc(const fdb_wvm_file* f) { return f ? &f->desc : nullptr; }
void fdb_wvm_file_free(fdb_wvm_file* f) { delete f; }

/* RvmClassifier::loadFromMatlab (RvmClassifier.cpp:141-319) + ProbabilisticRvmClassifier::loadSigmoidParamsFromMatlab
 * (ProbabilisticRvmClassifier.cpp:92-125): num_hk, param_nonlin1_rvm | param_nonlin1 {bias, kernel type, basis parameter
 * (/ 65025), power, divisor}, support_hk%d as CV_32F vectors in row-major order (no grey-value scaling), weight_hk%d (i + 1
 * coefficients of level i), hierar_thresh and posterior_wrvm = {B, A} from the thresholds file; setNumFiltersToUse(num_hk). */
int fdb_rvm_file_load(const char* classifier_path, const char* thresholds_path, fdb_rvm_file** out) try {
	if (!classifier_path || !thresholds_path || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	MatFile mf; std::string err;
	if (!mat_read(classifier_path, &mf, &err)) return fail(FDB_ERR_INVALID_ARGUMENT, "RvmClassifier: Could not open the provided classifier filename: " + err);
	const MatArray* a = mat_var(mf, "num_hk");
	if (!a || a->real.empty()) return fail(FDB_ERR_RUNTIME, "RvmClassifier: There is a no num_hk in the classifier file.");
	const int nfilter = (int)a->real[0];
	if (nfilter < 1 || nfilter > 4096) return fail(FDB_ERR_RUNTIME, "RvmClassifier: bad num_hk");
	a = mat_var(mf, "param_nonlin1_rvm");
	if (!a) a = mat_var(mf, "param_nonlin1");
	if (!a || a->real.size() < 5) return fail(FDB_ERR_RUNTIME, "RvmClassifier: Could not find kernel parameters and bias.");
	std::unique_ptr<fdb_rvm_file> f(new fdb_rvm_file);
	fdb_rvm_desc& d = f->desc;
	d = fdb_rvm_desc();
	d.bias = (float)a->real[0];
	const int nonLinType = (int)a->real[1];
	const float basisParam = (float)(a->real[2] / 65025.0);
	const int polyPower = (int)a->real[3];
	const float divisor = (float)a->real[4];
	if (nonLinType == 1) { /* PolynomialKernel(1 / divisor, basisParam / divisor, polyPower), float arithmetic */
		d.kernel = FDB_KERNEL_POLYNOMIAL;
		d.poly_alpha = 1 / divisor; d.poly_constant = basisParam / divisor; d.poly_degree = polyPower;
	} else if (nonLinType == 2) {
		d.kernel = FDB_KERNEL_RBF;
		d.gamma = basisParam;
	} else return fail(FDB_ERR_RUNTIME, "RvmClassifier: Unsupported kernel type. Currently, only polynomial and RBF kernels are supported.");
	a = mat_var(mf, "support_hk1");
	if (!a) return fail(FDB_ERR_RUNTIME, "RvmClassifier: Unable to find the matrix 'support_hk1' in the classifier file.");
	if (a->dims.size() != 2) return fail(FDB_ERR_RUNTIME, "RvmClassifier: The matrix 'support_hk1' in the classifier file should have 2 dimensions.");
	const int h = dim0(*a), w = dim1(*a);
	d.num_filters = nfilter; d.dim = w * h; d.sv_type = FDB_SV_F32;
	f->support_vectors.assign((size_t)nfilter * w * h, 0.f);
	f->coefficients.assign((size_t)nfilter * (nfilter + 1) / 2, 0.f);
	int n_weights = 0;
	for (int i = 0; i < nfilter; ++i) {
		const std::string idx = std::to_string(i + 1);
		a = mat_var(mf, "support_hk" + idx);
		if (!a) return fail(FDB_ERR_RUNTIME, "RvmClassifier: Unable to find the matrix 'support_hk" + idx + "' in the classifier file.");
		if (a->dims.size() != 2) return fail(FDB_ERR_RUNTIME, "RvmClassifier: The matrix 'support_hk" + idx + "' in the classifier file should have 2 dimensions.");
		if (a->real.size() < (size_t)w * h) return fail(FDB_ERR_RUNTIME, "RvmClassifier: support_hk" + idx + " is smaller than support_hk1");
		float* values = &f->support_vectors[(size_t)i * w * h];
		size_t k = 0;
		for (int x = 0; x < w; ++x)      /* column-major order (ML-convention), RvmClassifier.cpp:252-254 */
			for (int y = 0; y < h; ++y) values[(size_t)y * w + x] = (float)a->real[k++];
		a = mat_var(mf, "weight_hk" + idx);
		if (a) { /* a missing weight_hk is skipped by the reference and caught by the size check at the end */
			if (dim1(*a) != i + 1 && dim0(*a) != i + 1)
				return fail(FDB_ERR_RUNTIME, "RvmClassifier: The matrix weight_hk" + idx + " in the classifier file should have a dimensions 1x" + idx + " or " + idx + "x1");
			if (a->real.size() < (size_t)i + 1)
				return fail(FDB_ERR_RUNTIME, "RvmClassifier: The matrix weight_hk" + idx + " in the classifier file is not a numeric array of " + idx + " elements");
			for (int j = 0; j <= i; ++j) f->coefficients[(size_t)i * (i + 1) / 2 + j] = (float)a->real[j];
			++n_weights;
		}
	}
	MatFile tf;
	if (!mat_read(thresholds_path, &tf, &err)) return fail(FDB_ERR_RUNTIME, "RvmClassifier: Unable to open the thresholds file (wrong format?):" + err);
	a = mat_var(tf, "hierar_thresh");
	if (!a) return fail(FDB_ERR_RUNTIME, "RvmClassifier: Unable to find the matrix hierar_thresh in the thresholds file.");
	for (int o = 0; o < dim1(*a) && (size_t)o < a->real.size(); ++o) f->hierarchical_thresholds.push_back((float)a->real[o]);
	if ((int)f->hierarchical_thresholds.size() != n_weights || n_weights != nfilter)
		return fail(FDB_ERR_RUNTIME, "RvmClassifier: Something seems to be wrong, hierarchicalThresholds.size() != coefficients.size(): "
				+ std::to_string(f->hierarchical_thresholds.size()) + "!=" + std::to_string(n_weights));
	a = mat_var(tf, "posterior_wrvm");
	if (!a) return fail(FDB_ERR_RUNTIME, "ProbabilisticRvmClassifier: Unable to find the vector posterior_wrvm. If you don't want probabilistic output, don't use a probabilistic classifier.");
	if (dim1(*a) != 2 || a->real.size() < 2) return fail(FDB_ERR_RUNTIME, "ProbabilisticRvmClassifier: Size of vector posterior_wrvm !=2. If you don't want probabilistic output, don't use a probabilistic classifier.");
	d.logistic_b = a->real[0]; d.logistic_a = a->real[1];
	d.num_filters_to_use = nfilter;
	d.support_vectors = f->support_vectors.data();
	d.coefficients = f->coefficients.data();
	d.hierarchical_thresholds = f->hierarchical_thresholds.data();
	*out = f.release();
	return FDB_OK;
} FDB_API_CATCH

const fdb_rvm_desc* fdb_rvm_file_desc(const fdb_rvm_file* f) { return f ? &f->desc : nullptr; }

void fdb_rvm_file_free(fdb_rvm_file* f) { delete f; }

int fdb_svm_mat_load(const char* classifier_path, const char* logistic_path, fdb_svm_file** out) try {
	if (!classifier_path || !out) return fail(FDB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	MatFile mf; std::string err;
	if (!mat_read(classifier_path, &mf, &err)) return fail(FDB_ERR_INVALID_ARGUMENT, "SvmClassifier: Could not open the provided classifier filename: " + err);
	const MatArray* a = mat_var(mf, "param_nonlin1");
	if (!a || a->real.size() < 5) return fail(FDB_ERR_RUNTIME, "SvmClassifier: There is a no param_nonlin1 in the classifier file.");
	std::unique_ptr<fdb_svm_file> f(new fdb_svm_file);
	fdb_svm_desc& d = f->desc;
	d = fdb_svm_desc();
	d.bias = (float)a->real[0];
	const int nonLinType = (int)a->real[1];
	const float basisParam = (float)(a->real[2] / 65025.0);
	if (nonLinType == 1) { /* SvmClassifier.cpp:270-272: PolynomialKernel(1 / divisor, basisParam / divisor, polyPower), float arithmetic */
		const int polyPower = (int)a->real[3];
		const float divisor = (float)a->real[4];
		d.kernel = FDB_KERNEL_POLYNOMIAL;
		d.poly_alpha = 1 / divisor; d.poly_constant = basisParam / divisor; d.poly_degree = polyPower;
	} else if (nonLinType == 2) {
		d.kernel = FDB_KERNEL_RBF;
		d.gamma = basisParam; /* RbfKernel(double gamma) receives the float */
	} else return fail(FDB_ERR_RUNTIME, "SvmClassifier: Unsupported kernel type. Currently, only polynomial and RBF kernels are supported.");
	a = mat_var(mf, "support_nonlin1");
	if (!a) return fail(FDB_ERR_RUNTIME, "SvmClassifier: There is a nonlinear SVM in the file, but the matrix support_nonlin1 is lacking.");
	if (a->dims.size() != 3) return fail(FDB_ERR_RUNTIME, "SvmClassifier: The matrix support_nonlin1 in the file should have 3 dimensions.");
	const int h = a->dims[0], w = a->dims[1], nsv = a->dims[2];
	if (h < 1 || w < 1 || nsv < 1 || (int64_t)h * w > 96 * 1024 / 4 || a->real.size() / ((size_t)h * w) < (size_t)nsv)
		return fail(FDB_ERR_RUNTIME, "SvmClassifier: The matrix support_nonlin1 is not a numeric h x w x n array (or a vector exceeds 24576 elements)");
	f->sv_u8.resize((size_t)nsv * w * h);
	size_t k = 0;
	for (int sv = 0; sv < nsv; ++sv)
		for (int x = 0; x < w; ++x)      /* column-major order (ML-convention), SvmClassifier.cpp:303-309 */
			for (int y = 0; y < h; ++y)
				f->sv_u8[(size_t)sv * w * h + (size_t)y * w + x] = static_cast<uint8_t>(255.0 * a->real[k++]);
	a = mat_var(mf, "weight_nonlin1");
	if (!a || (int)a->real.size() < nsv) return fail(FDB_ERR_RUNTIME, "SvmClassifier: There is a nonlinear SVM in the file but the matrix threshold_nonlin is lacking.");
	for (int sv = 0; sv < nsv; ++sv) f->coefficients.push_back(static_cast<float>(a->real[sv]));
	f->rows = h; f->cols = w; f->channels = 1; f->depth = 0;
	d.num_sv = nsv; d.dim = w * h; d.sv_type = FDB_SV_U8;
	d.support_vectors = f->sv_u8.data();
	d.coefficients = f->coefficients.data();
	d.threshold = 0.f; /* the cfg's "threshold" is applied by the caller (ProbabilisticSvmClassifier.cpp:100) */
	if (logistic_path) {
		const int s = load_posterior(logistic_path, "posterior_svm", false, &d.logistic_a, &d.logistic_b);
		if (s) return s;
	}
	*out = f.release();
	return FDB_OK;
} FDB_API_CATCH

} // extern "C"
