/* Minimal stand-in for boost::filesystem::path (compile-only). */
#ifndef FDB_SHIM_BOOST_FS_PATH_HPP
#define FDB_SHIM_BOOST_FS_PATH_HPP
#include <string>
namespace boost { namespace filesystem {
class path {
public:
	path() {}
	path(const std::string& s) : s(s) {}
	path(const char* c) : s(c) {}
	path extension() const { size_t p = s.rfind('.'); return p == std::string::npos ? path() : path(s.substr(p)); }
	const std::string& string() const { return s; }
	bool operator==(const char* o) const { return s == o; }
	bool operator==(const path& o) const { return s == o.s; }
private:
	std::string s;
};
}}
#endif
