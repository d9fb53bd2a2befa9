// Minimal stand-in for <boost/range/iterator_range.hpp> -- TEST INFRASTRUCTURE ONLY.
#ifndef ADSB_SHIM_BOOST_ITERATOR_RANGE_HPP
#define ADSB_SHIM_BOOST_ITERATOR_RANGE_HPP

#include <iterator>

namespace boost {

template <typename It>
class iterator_range {
    It b_{}, e_{};

public:
    using iterator = It;
    using const_iterator = It;

    iterator_range() = default;
    iterator_range(It b, It e) : b_{b}, e_{e} { }

    It begin() const { return b_; }
    It end() const { return e_; }
    bool empty() const { return b_ == e_; }

    friend bool operator==(const iterator_range& a, const iterator_range& b) {
        return a.b_ == b.b_ && a.e_ == b.e_;
    }
    friend bool operator!=(const iterator_range& a, const iterator_range& b) { return !(a == b); }
};

template <typename It>
iterator_range<It> make_iterator_range(It b, It e) { return {b, e}; }

template <typename R>
auto begin(const R& r) -> decltype(r.begin()) { return r.begin(); }

template <typename R>
auto end(const R& r) -> decltype(r.end()) { return r.end(); }

}  // namespace boost

#endif
