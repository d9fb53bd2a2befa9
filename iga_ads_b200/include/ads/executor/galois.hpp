// ads/executor/galois.hpp -- galois_executor (include/ads/executor/galois.hpp:18-44) for host lambdas: the
// reference runs its element loops through galois::do_all with a SimpleLock for `synchronized`.  Galois is not a
// dependency here; the same contract -- for_each over a range on `threads` workers, synchronized(fun) under one
// lock -- is kept with std::thread, so an example's own host-side element loop compiles and runs unchanged.
// (The ADS step itself runs no host lambdas: the device forms replace these loops.)
#ifndef ADSB_ADS_EXECUTOR_GALOIS_HPP
#define ADSB_ADS_EXECUTOR_GALOIS_HPP

#include <algorithm>
#include <iterator>
#include <mutex>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

namespace ads {

class galois_executor {
    int threads_;
    mutable std::mutex lock_;

public:
    explicit galois_executor(int threads) : threads_{std::max(threads, 1)} { }

    template <typename Fun>
    void synchronized(Fun&& fun) const {
        std::lock_guard<std::mutex> guard{lock_};
        fun();
    }

    template <typename Range, typename Fun>
    void for_each(const Range& range, Fun&& fun) const {
        using std::begin;
        using std::end;
        if (threads_ == 1) {
            for (auto&& item : range) fun(item);
            return;
        }
        using item_t = std::decay_t<decltype(*begin(range))>;
        const std::vector<item_t> items(begin(range), end(range));
        const std::size_t n = items.size();
        std::vector<std::thread> pool;
        for (int t = 0; t < threads_; ++t)
            pool.emplace_back([&, t] {
                for (std::size_t i = n * t / threads_, hi = n * (t + 1) / threads_; i < hi; ++i) fun(items[i]);
            });
        for (auto& th : pool) th.join();
    }
};

}  // namespace ads

#endif
