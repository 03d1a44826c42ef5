/* Minimal stand-in for boost::property_tree::ptree: only lets the reference's
 * Probabilistic*Classifier::load(ptree) definitions compile; never called by the oracle. */
#ifndef FDB_SHIM_BOOST_PTREE_HPP
#define FDB_SHIM_BOOST_PTREE_HPP
#include <stdexcept>
#include <string>
namespace boost { namespace property_tree {
class ptree {
public:
	template<class T> T get(const std::string& key) const { throw std::runtime_error("shim ptree: no key " + key); }
	template<class T> T get(const std::string&, const T& def) const { return def; }
};
}}
#endif
