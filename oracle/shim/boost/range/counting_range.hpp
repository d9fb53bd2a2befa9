// Minimal stand-in for <boost/range/counting_range.hpp> -- TEST INFRASTRUCTURE ONLY.
#ifndef ADSB_SHIM_BOOST_COUNTING_RANGE_HPP
#define ADSB_SHIM_BOOST_COUNTING_RANGE_HPP

#include "boost/iterator/counting_iterator.hpp"
#include "boost/range/iterator_range.hpp"

namespace boost {

template <typename T>
iterator_range<counting_iterator<T>> counting_range(T a, T b) {
    return {counting_iterator<T>(a), counting_iterator<T>(b)};
}

}  // namespace boost

#endif
