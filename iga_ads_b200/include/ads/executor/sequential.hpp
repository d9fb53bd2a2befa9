// ads/executor/sequential.hpp -- for_each / synchronized (include/ads/executor/sequential.hpp:15-24) for host
// lambdas that examples may still run (set-up, diagnostics).  The ADS step itself runs no host lambdas.
#ifndef ADSB_ADS_EXECUTOR_SEQUENTIAL_HPP
#define ADSB_ADS_EXECUTOR_SEQUENTIAL_HPP

namespace ads {

class sequential_executor {
public:
    template <typename Range, typename Fun>
    void for_each(const Range& range, Fun&& fun) const {
        for (auto&& item : range) fun(item);
    }
    template <typename Fun>
    void synchronized(Fun&& fun) const {
        fun();
    }
};

}  // namespace ads

#endif
