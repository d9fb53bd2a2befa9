// ads/device.hpp -- the only place the C++ host layer touches libadsb200's C ABI state: one device
// context per simulation, error translation, and the bookkeeping that lets ads::lin::tensor objects
// keep a device mirror (managed buffer id + validity flags).
#ifndef ADSB_ADS_DEVICE_HPP
#define ADSB_ADS_DEVICE_HPP

#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "adsb200.h"

namespace ads::device {

inline void check(int rc) {
    if (rc < 0) throw std::runtime_error(std::string("libadsb200: ") + adsb_last_error());
}

// Owns an adsb_ctx; hands out managed buffer ids and remembers which factors sit in which slot.
class context {
public:
    context(int ndim, const int* n) {
        check(adsb_create(ndim, n, nullptr, nullptr, 0, &h_));
    }
    ~context() { adsb_destroy(h_); }
    context(const context&) = delete;
    context& operator=(const context&) = delete;

    adsb_ctx* handle() const { return h_; }

    // managed buffer ids are recycled: a tensor returns its id when it is destroyed or re-shaped
    int new_buffer() {
        if (!free_bufs_.empty()) {
            const int b = free_bufs_.back();
            free_bufs_.pop_back();
            return b;
        }
        if (next_buf_ >= ADSB_MAX_BUFFERS) throw std::runtime_error("libadsb200: out of managed buffers");
        return next_buf_++;
    }
    void free_buffer(int b) noexcept {
        try {
            free_bufs_.push_back(b);
        } catch (...) {  // out of memory while recycling an id: leak the id, never throw from a destructor
        }
    }

    // slot holding this factor on `axis`, uploading it first if its content is new
    int factor_slot(int axis, int n, int kl, int ku, int ldab, const double* ab, const int* ipiv) {
        std::uint64_t h = 1469598103934665603ull;
        auto mix = [&h](const void* p, std::size_t bytes) {
            const unsigned char* c = static_cast<const unsigned char*>(p);
            for (std::size_t i = 0; i < bytes; ++i) h = (h ^ c[i]) * 1099511628211ull;
        };
        mix(ab, sizeof(double) * static_cast<std::size_t>(n) * ldab);
        mix(ipiv, sizeof(int) * static_cast<std::size_t>(n));
        auto& known = slots_[axis];
        for (std::size_t s = 0; s < known.size(); ++s)
            if (known[s] == h) return static_cast<int>(s);
        int slot = static_cast<int>(known.size());
        if (slot >= ADSB_MAX_SLOTS) {  // recycle the oldest non-primary slot
            slot = 1 + (evict_[axis]++ % (ADSB_MAX_SLOTS - 1));
            known[slot] = h;
        } else {
            known.push_back(h);
        }
        check(adsb_set_axis_factor(h_, axis, slot, n, kl, ku, ldab, ab, ipiv));
        return slot;
    }

    // which line-factor table (generalised ADS) the context holds on `axis`
    bool line_factors_current(int axis, const void* owner, unsigned long long version) const {
        return line_owner_[axis] == owner && line_version_[axis] == version;
    }
    void line_factors_uploaded(int axis, const void* owner, unsigned long long version) {
        line_owner_[axis] = owner;
        line_version_[axis] = version;
    }

private:
    adsb_ctx* h_ = nullptr;
    std::array<const void*, 3> line_owner_{};
    std::array<unsigned long long, 3> line_version_{};
    int next_buf_ = 0;
    std::vector<int> free_bufs_;
    std::array<std::vector<std::uint64_t>, 3> slots_;
    std::array<int, 3> evict_{};
};

// Device mirror of one tensor.  host_valid / dev_valid say which copy is current.
struct mirror {
    std::shared_ptr<context> ctx;
    int buf = -1;
    bool host_valid = true;
    bool dev_valid = false;
};

}  // namespace ads::device

#endif
