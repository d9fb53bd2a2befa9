// Stand-in for the reference's include/ads/executor/galois.hpp -- TEST INFRASTRUCTURE ONLY.
// Galois is not installed, so `galois_executor` here runs the element loop the way the
// reference's own sequential_executor does (include/ads/executor/sequential.hpp:15-24),
// or, when constructed with threads > 1, splits the element range over std::threads with
// one global mutex for `synchronized` -- the same contract as the reference's
// galois::do_all + SimpleLock (include/ads/executor/galois.hpp:28-44).  Shadowing works
// because this directory precedes /root/reference/include on the include path.
#ifndef ADSB_SHIM_ADS_EXECUTOR_GALOIS_HPP
#define ADSB_SHIM_ADS_EXECUTOR_GALOIS_HPP

#include <algorithm>
#include <iterator>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include <galois/Timer.h>

namespace ads {

class galois_executor {
    int threads_;
    mutable std::mutex lock_;

public:
    static int& thread_override() {
        static int n = 0;  // 0: honour the constructor argument; >0: force this many threads
        return n;
    }

    explicit galois_executor(int threads) : threads_{threads} { }

    template <typename Fun>
    void synchronized(Fun fun) const {
        std::lock_guard<std::mutex> guard{lock_};
        fun();
    }

    template <typename Range, typename Fun>
    void for_each(Range range, Fun&& fun) const {
        using std::begin;
        using std::end;
        int nt = thread_override() > 0 ? thread_override() : 1;
        if (nt <= 1) {
            std::for_each(begin(range), end(range), std::forward<Fun>(fun));
            return;
        }
        using value_t = decltype(*begin(range));
        std::vector<std::decay_t<value_t>> items;
        for (auto it = begin(range); it != end(range); ++it) items.push_back(*it);
        std::vector<std::thread> pool;
        const std::size_t n = items.size();
        for (int t = 0; t < nt; ++t) {
            pool.emplace_back([&, t] {
                std::size_t lo = n * t / nt, hi = n * (t + 1) / nt;
                for (std::size_t i = lo; i < hi; ++i) fun(items[i]);
            });
        }
        for (auto& th : pool) th.join();
    }
};

}  // namespace ads

#endif
