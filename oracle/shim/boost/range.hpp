// Minimal stand-in for <boost/range.hpp> -- TEST INFRASTRUCTURE ONLY.
#ifndef ADSB_SHIM_BOOST_RANGE_HPP
#define ADSB_SHIM_BOOST_RANGE_HPP
#include "boost/range/counting_range.hpp"
#include "boost/range/iterator_range.hpp"
#endif
