// Stand-in for <galois/Timer.h> -- TEST INFRASTRUCTURE ONLY (Galois is not installed).
// The reference examples wrap compute_rhs in a galois::StatTimer; this one keeps the
// same start/stop/get surface on std::chrono so the oracle driver can read the
// reference's own "integration" time.
#ifndef ADSB_SHIM_GALOIS_TIMER_H
#define ADSB_SHIM_GALOIS_TIMER_H

#include <chrono>
#include <cstdint>

namespace galois {

class StatTimer {
    using clock = std::chrono::steady_clock;
    clock::time_point t0_{};
    std::uint64_t total_us_ = 0;

public:
    explicit StatTimer(const char* /*name*/ = nullptr) { }
    void start() { t0_ = clock::now(); }
    void stop() {
        total_us_ += static_cast<std::uint64_t>(
            std::chrono::duration_cast<std::chrono::microseconds>(clock::now() - t0_).count());
    }
    // Galois reports milliseconds
    std::uint64_t get() const { return total_us_ / 1000; }
    std::uint64_t get_usec() const { return total_us_; }
};

}  // namespace galois

#endif
