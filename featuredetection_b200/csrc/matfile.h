/*
 * matfile.h - minimal reader of MATLAB Level-5 MAT-files (the container of the reference's classifier models).
 * The reference reads them through MATLAB's own libmat (matOpen / matGetVariable / mxGetPr / mxGetField, e.g.
 * WvmClassifier.cpp:365-372), a proprietary dependency that is absent here; the file format itself is published
 * ("MAT-File Format", MathWorks) and is what this reader implements: little-endian v5 files, miCOMPRESSED elements
 * (zlib), numeric arrays of any storage type (returned as double, which is what mxGetPr yields for the double-class
 * variables of these models), struct arrays and cell arrays.  v7.3 (HDF5) files are rejected.
 */
#ifndef FDB_MATFILE_H_
#define FDB_MATFILE_H_

#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace fdb {

struct MatArray {
	int cls = 0;                        /* mxDOUBLE_CLASS = 6, mxSTRUCT_CLASS = 2, mxCELL_CLASS = 1, ... */
	std::vector<int> dims;
	std::vector<double> real;           /* numeric / char classes: column-major elements */
	std::vector<std::string> fields;    /* struct classes */
	std::vector<MatArray> children;     /* struct: [element * fields.size() + field]; cell: [element] */
	int64_t numel() const { int64_t n = dims.empty() ? 0 : 1; for (int d : dims) n *= d; return n; }
	bool is_struct() const { return cls == 2; }
	/* mxGetField(array, index, name): nullptr when the field does not exist */
	const MatArray* field(int64_t index, const char* name) const;
};

struct MatFile {
	std::map<std::string, MatArray> vars;
	/* matGetVariable: nullptr when absent */
	const MatArray* get(const std::string& name) const { auto it = vars.find(name); return it == vars.end() ? nullptr : &it->second; }
};

/* matOpen + read of every variable. Returns false and sets `error` when the file cannot be opened or parsed. */
bool mat_read(const std::string& path, MatFile* out, std::string* error);

} // namespace fdb
#endif
