/* Minimal stand-in for boost::indirect_iterator / make_indirect_iterator: a random-access
 * iterator adaptor that dereferences twice (used by OverlapElimination.cpp:62 to std::sort the
 * pointed-to ClassifiedPatch objects). Test infrastructure only. */
#ifndef FDB_SHIM_BOOST_INDIRECT_ITERATOR_HPP
#define FDB_SHIM_BOOST_INDIRECT_ITERATOR_HPP
#include <iterator>
#include <memory>
namespace boost {
template<class It>
class indirect_iterator {
	typedef typename std::iterator_traits<It>::value_type pointer_like;
public:
	typedef std::random_access_iterator_tag iterator_category;
	typedef typename std::pointer_traits<pointer_like>::element_type value_type;
	typedef typename std::iterator_traits<It>::difference_type difference_type;
	typedef value_type* pointer;
	typedef value_type& reference;
	indirect_iterator() {}
	explicit indirect_iterator(It it) : it(it) {}
	reference operator*() const { return **it; }
	pointer operator->() const { return &**it; }
	reference operator[](difference_type n) const { return **(it + n); }
	indirect_iterator& operator++() { ++it; return *this; }
	indirect_iterator operator++(int) { indirect_iterator t(*this); ++it; return t; }
	indirect_iterator& operator--() { --it; return *this; }
	indirect_iterator operator--(int) { indirect_iterator t(*this); --it; return t; }
	indirect_iterator& operator+=(difference_type n) { it += n; return *this; }
	indirect_iterator& operator-=(difference_type n) { it -= n; return *this; }
	indirect_iterator operator+(difference_type n) const { return indirect_iterator(it + n); }
	indirect_iterator operator-(difference_type n) const { return indirect_iterator(it - n); }
	difference_type operator-(const indirect_iterator& o) const { return it - o.it; }
	bool operator==(const indirect_iterator& o) const { return it == o.it; }
	bool operator!=(const indirect_iterator& o) const { return it != o.it; }
	bool operator<(const indirect_iterator& o) const { return it < o.it; }
	bool operator>(const indirect_iterator& o) const { return it > o.it; }
	bool operator<=(const indirect_iterator& o) const { return it <= o.it; }
	bool operator>=(const indirect_iterator& o) const { return it >= o.it; }
private:
	It it;
};
template<class It>
indirect_iterator<It> operator+(typename indirect_iterator<It>::difference_type n, const indirect_iterator<It>& i) { return i + n; }
template<class It>
indirect_iterator<It> make_indirect_iterator(It it) { return indirect_iterator<It>(it); }
}
#endif
