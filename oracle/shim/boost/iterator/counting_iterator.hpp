// Minimal stand-in for <boost/iterator/counting_iterator.hpp> -- TEST INFRASTRUCTURE ONLY.
#ifndef ADSB_SHIM_BOOST_COUNTING_ITERATOR_HPP
#define ADSB_SHIM_BOOST_COUNTING_ITERATOR_HPP

#include <cstddef>
#include <iterator>

namespace boost {

template <typename T>
class counting_iterator {
    T v_{};

public:
    using value_type = T;
    using reference = T;
    using pointer = const T*;
    using difference_type = std::ptrdiff_t;
    using iterator_category = std::random_access_iterator_tag;

    counting_iterator() = default;
    explicit counting_iterator(T v) : v_{v} { }

    T operator*() const { return v_; }
    counting_iterator& operator++() { ++v_; return *this; }
    counting_iterator operator++(int) { auto t = *this; ++v_; return t; }
    counting_iterator& operator--() { --v_; return *this; }
    counting_iterator& operator+=(difference_type d) { v_ += static_cast<T>(d); return *this; }
    counting_iterator operator+(difference_type d) const { return counting_iterator(v_ + static_cast<T>(d)); }
    difference_type operator-(const counting_iterator& o) const { return v_ - o.v_; }
    T operator[](difference_type d) const { return v_ + static_cast<T>(d); }

    bool operator==(const counting_iterator& o) const { return v_ == o.v_; }
    bool operator!=(const counting_iterator& o) const { return v_ != o.v_; }
    bool operator<(const counting_iterator& o) const { return v_ < o.v_; }
};

}  // namespace boost

#endif
