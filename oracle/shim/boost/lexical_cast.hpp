/* Minimal stand-in for boost::lexical_cast (Boost is not installed). Test infrastructure only. */
#ifndef FDB_SHIM_BOOST_LEXICAL_CAST_HPP
#define FDB_SHIM_BOOST_LEXICAL_CAST_HPP
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
template<class Target, class Source>
Target lexical_cast(const Source& s) {
	std::stringstream ss;
	ss << s;
	Target t;
	if (!(ss >> t)) throw std::runtime_error("shim lexical_cast failed");
	return t;
}
}
#endif
