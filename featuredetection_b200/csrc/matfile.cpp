/*
 * matfile.cpp - MATLAB Level-5 MAT-file reader (see matfile.h). Product code (host, C++).
 */
#include "matfile.h"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace fdb {

namespace {

enum { miINT8 = 1, miUINT8 = 2, miINT16 = 3, miUINT16 = 4, miINT32 = 5, miUINT32 = 6, miSINGLE = 7, miDOUBLE = 9, miINT64 = 12, miUINT64 = 13,
       miMATRIX = 14, miCOMPRESSED = 15, miUTF8 = 16, miUTF16 = 17, miUTF32 = 18 };
enum { mxCELL = 1, mxSTRUCT = 2, mxOBJECT = 3, mxCHAR = 4, mxSPARSE = 5, mxDOUBLE = 6, mxUINT64 = 15 };

struct Cursor {
	const uint8_t* p;
	size_t n, pos;
	void need(size_t k) const { if (pos + k > n) throw std::runtime_error("truncated MAT-file element"); }
	uint32_t u32() { need(4); uint32_t v; std::memcpy(&v, p + pos, 4); pos += 4; return v; }
};

struct Element { uint32_t type; const uint8_t* data; size_t size; };

/* one data element (tag + data); advances past the padding to the next 8-byte boundary */
Element next_element(Cursor& c) {
	const uint32_t w = c.u32();
	Element e;
	if (w >> 16) { /* small data element: type and size share the first word, up to 4 data bytes follow */
		e.type = w & 0xffffu; e.size = w >> 16;
		if (e.size > 4) throw std::runtime_error("bad small data element");
		c.need(4);
		e.data = c.p + c.pos; c.pos += 4;
		return e;
	}
	e.type = w; e.size = c.u32();
	c.need(e.size);
	e.data = c.p + c.pos;
	c.pos += e.size;
	if (e.type != miCOMPRESSED) c.pos = (c.pos + 7) & ~(size_t)7;
	if (c.pos > c.n) c.pos = c.n;
	return e;
}

size_t type_size(uint32_t t) {
	switch (t) {
	case miINT8: case miUINT8: case miUTF8: return 1;
	case miINT16: case miUINT16: case miUTF16: return 2;
	case miINT32: case miUINT32: case miSINGLE: case miUTF32: return 4;
	case miDOUBLE: case miINT64: case miUINT64: return 8;
	default: throw std::runtime_error("unsupported MAT-file storage type " + std::to_string(t));
	}
}

void to_double(const Element& e, std::vector<double>* out) {
	const size_t sz = type_size(e.type), n = e.size / sz;
	out->resize(n);
	for (size_t i = 0; i < n; ++i) {
		const uint8_t* q = e.data + i * sz;
		double v;
		switch (e.type) {
		case miINT8: v = *(const int8_t*)q; break;
		case miUINT8: case miUTF8: v = *q; break;
		case miINT16: { int16_t t; std::memcpy(&t, q, 2); v = t; break; }
		case miUINT16: case miUTF16: { uint16_t t; std::memcpy(&t, q, 2); v = t; break; }
		case miINT32: { int32_t t; std::memcpy(&t, q, 4); v = t; break; }
		case miUINT32: case miUTF32: { uint32_t t; std::memcpy(&t, q, 4); v = t; break; }
		case miSINGLE: { float t; std::memcpy(&t, q, 4); v = t; break; }
		case miDOUBLE: { std::memcpy(&v, q, 8); break; }
		case miINT64: { int64_t t; std::memcpy(&t, q, 8); v = (double)t; break; }
		default: { uint64_t t; std::memcpy(&t, q, 8); v = (double)t; break; }
		}
		(*out)[i] = v;
	}
}

void parse_matrix(const uint8_t* data, size_t size, MatArray* a, std::string* name, int depth) {
	if (depth > 16) throw std::runtime_error("MAT-file nesting too deep");
	*a = MatArray();
	if (size == 0) return; /* empty placeholder (e.g. an unset struct field) */
	Cursor c{data, size, 0};
	const Element flags = next_element(c);
	if (flags.type != miUINT32 || flags.size < 8) throw std::runtime_error("bad array flags");
	uint32_t f0; std::memcpy(&f0, flags.data, 4);
	a->cls = (int)(f0 & 0xffu);
	const bool is_complex = (f0 & 0x0800u) != 0;
	const Element dims = next_element(c);
	if (dims.type != miINT32) throw std::runtime_error("bad dimensions element");
	for (size_t i = 0; i + 4 <= dims.size; i += 4) { int32_t d; std::memcpy(&d, dims.data + i, 4); if (d < 0) throw std::runtime_error("negative dimension"); a->dims.push_back(d); }
	const Element nm = next_element(c);
	if (name) name->assign((const char*)nm.data, nm.size);
	const int64_t numel = a->numel();
	if (a->cls == mxSTRUCT || a->cls == mxOBJECT) {
		if (a->cls == mxOBJECT) next_element(c); /* class name */
		const Element flen = next_element(c);
		int32_t len = 0;
		if (flen.size >= 4) std::memcpy(&len, flen.data, 4);
		if (len <= 0 || len > 256) throw std::runtime_error("bad struct field name length");
		const Element fnames = next_element(c);
		const size_t nf = fnames.size / (size_t)len;
		for (size_t i = 0; i < nf; ++i) {
			const char* s = (const char*)fnames.data + i * len;
			a->fields.emplace_back(s, strnlen(s, (size_t)len));
		}
		a->children.resize((size_t)numel * nf);
		for (size_t i = 0; i < a->children.size(); ++i) {
			const Element e = next_element(c);
			if (e.type != miMATRIX) throw std::runtime_error("struct field is not a matrix element");
			parse_matrix(e.data, e.size, &a->children[i], nullptr, depth + 1);
		}
		a->cls = mxSTRUCT;
	} else if (a->cls == mxCELL) {
		a->children.resize((size_t)numel);
		for (size_t i = 0; i < a->children.size(); ++i) {
			const Element e = next_element(c);
			if (e.type != miMATRIX) throw std::runtime_error("cell is not a matrix element");
			parse_matrix(e.data, e.size, &a->children[i], nullptr, depth + 1);
		}
	} else if (a->cls == mxCHAR || (a->cls >= mxDOUBLE && a->cls <= mxUINT64)) {
		if (numel > 0) {
			const Element re = next_element(c);
			to_double(re, &a->real);
			if ((int64_t)a->real.size() < numel) throw std::runtime_error("numeric array shorter than its dimensions");
			a->real.resize((size_t)numel);
		}
		(void)is_complex; /* the imaginary part, if any, is ignored like mxGetPr does */
	} /* sparse / function handles: kept as an empty array of that class */
}

void inflate_all(const uint8_t* src, size_t n, std::vector<uint8_t>* out) {
	z_stream zs;
	std::memset(&zs, 0, sizeof zs);
	if (inflateInit(&zs) != Z_OK) throw std::runtime_error("zlib inflateInit failed");
	zs.next_in = const_cast<Bytef*>(src);
	zs.avail_in = (uInt)n;
	out->assign(std::max<size_t>(4 * n, 1 << 12), 0);
	size_t have = 0;
	for (;;) {
		if (have == out->size()) out->resize(out->size() * 2);
		zs.next_out = out->data() + have;
		zs.avail_out = (uInt)std::min<size_t>(out->size() - have, 1u << 30);
		const size_t before = zs.avail_out;
		const int rc = inflate(&zs, Z_NO_FLUSH);
		have += before - zs.avail_out;
		if (rc == Z_STREAM_END) break;
		if (rc != Z_OK && rc != Z_BUF_ERROR) { inflateEnd(&zs); throw std::runtime_error("zlib inflate failed (corrupt compressed MAT-file element)"); }
		if (rc == Z_BUF_ERROR && zs.avail_in == 0 && zs.avail_out != 0) { inflateEnd(&zs); throw std::runtime_error("truncated compressed MAT-file element"); }
	}
	inflateEnd(&zs);
	out->resize(have);
}

} // namespace

const MatArray* MatArray::field(int64_t index, const char* name) const {
	if (cls != 2 || index < 0 || index >= numel()) return nullptr;
	for (size_t f = 0; f < fields.size(); ++f)
		if (fields[f] == name) return &children[(size_t)index * fields.size() + f];
	return nullptr;
}

bool mat_read(const std::string& path, MatFile* out, std::string* error) {
	out->vars.clear();
	FILE* fp = std::fopen(path.c_str(), "rb");
	if (!fp) { if (error) *error = "cannot open " + path; return false; }
	std::vector<uint8_t> buf;
	{
		uint8_t tmp[1 << 16];
		size_t k;
		while ((k = std::fread(tmp, 1, sizeof tmp, fp)) > 0) buf.insert(buf.end(), tmp, tmp + k);
		std::fclose(fp);
	}
	try {
		if (buf.size() >= 8 && std::memcmp(buf.data(), "\x89HDF\r\n\x1a\n", 8) == 0) throw std::runtime_error("HDF5 container (MAT-file v7.3) is not supported; save with -v7 or -v6");
		if (buf.size() < 128) throw std::runtime_error("not a MAT-file (shorter than the 128-byte header)");
		if (std::memcmp(buf.data(), "MATLAB 5.0 MAT-file", 19) != 0) throw std::runtime_error("not a Level-5 MAT-file (missing 'MATLAB 5.0 MAT-file' header)");
		if (!(buf[126] == 'I' && buf[127] == 'M')) throw std::runtime_error("big-endian MAT-files are not supported");
		Cursor c{buf.data(), buf.size(), 128};
		std::vector<uint8_t> inflated;
		while (c.pos + 8 <= c.n) {
			const Element e = next_element(c);
			const uint8_t* data = e.data;
			size_t size = e.size;
			if (e.type == miCOMPRESSED) {
				inflate_all(e.data, e.size, &inflated);
				Cursor ci{inflated.data(), inflated.size(), 0};
				const Element inner = next_element(ci);
				if (inner.type != miMATRIX) continue;
				data = inner.data; size = inner.size;
			} else if (e.type != miMATRIX) continue;
			MatArray a;
			std::string name;
			parse_matrix(data, size, &a, &name, 0);
			if (!name.empty()) out->vars[name] = std::move(a);
		}
	} catch (const std::exception& ex) {
		if (error) *error = path + ": " + ex.what();
		out->vars.clear();
		return false;
	}
	return true;
}

} // namespace fdb
