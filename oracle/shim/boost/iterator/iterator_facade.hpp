// Minimal stand-in for <boost/iterator/iterator_facade.hpp> -- TEST INFRASTRUCTURE ONLY.
// Written for this repo so that the unmodified reference headers under /root/reference
// compile without Boost (not installed in this image).  Only the subset used by
// include/ads/util/iter/product.hpp (reference) is provided.
#ifndef ADSB_SHIM_BOOST_ITERATOR_FACADE_HPP
#define ADSB_SHIM_BOOST_ITERATOR_FACADE_HPP

#include <cstddef>
#include <iterator>

namespace boost {

struct forward_traversal_tag { };

class iterator_core_access {
public:
    template <typename It>
    static void increment(It& it) { it.increment(); }

    template <typename It>
    static auto dereference(const It& it) -> decltype(it.dereference()) { return it.dereference(); }

    template <typename It>
    static bool equal(const It& a, const It& b) { return a.equal(b); }
};

template <typename Derived, typename Value, typename Traversal, typename Reference = Value&,
          typename Difference = std::ptrdiff_t>
class iterator_facade {
public:
    using value_type = Value;
    using reference = Reference;
    using pointer = Value*;
    using difference_type = Difference;
    using iterator_category = std::forward_iterator_tag;

    Reference operator*() const { return iterator_core_access::dereference(self()); }

    Derived& operator++() {
        iterator_core_access::increment(self());
        return self();
    }

    Derived operator++(int) {
        Derived tmp = self();
        iterator_core_access::increment(self());
        return tmp;
    }

    friend bool operator==(const Derived& a, const Derived& b) {
        return iterator_core_access::equal(a, b);
    }

    friend bool operator!=(const Derived& a, const Derived& b) {
        return !iterator_core_access::equal(a, b);
    }

private:
    Derived& self() { return static_cast<Derived&>(*this); }
    const Derived& self() const { return static_cast<const Derived&>(*this); }
};

}  // namespace boost

#endif
