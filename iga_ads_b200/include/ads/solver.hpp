// ads/solver.hpp -- ads_solve(rhs, buffer, dims...): the alternating-direction solve
// (reference: include/ads/solver.hpp:35-41,:148-160,:200-226).  The reference does, per axis,
// dgbtrs on the contiguous lines followed by a cyclic rotation of the tensor into `buffer`; here all
// sweeps run on the device against the one canonical layout (adsb_solve), `buffer` is not needed.
// The generalised variant (one dimension "special": a different matrix per line, solver.hpp:56-96,:170-195) takes an
// ads::line_factors object in the place of that dimension's dim_data -- the factorised matrices the reference's
// callable would apply line by line (examples/maxwell/maxwell_ads.hpp:139-163) -- and runs adsb_solve_special.
#ifndef ADSB_ADS_SOLVER_HPP
#define ADSB_ADS_SOLVER_HPP

#include <algorithm>
#include <memory>
#include <stdexcept>
#include <vector>

#include "ads/lin/tensor.hpp"
#include "ads/simulation/dimension.hpp"

namespace ads {

// One factorised band matrix per line of a special dimension.  Lines are numbered over the other axes in the
// tensor's own order (first index fastest): special x: l = iy + ny*iz; y: l = ix + nx*iz; z: l = ix + nx*iy
// (2-D: the other index).  set_matrix factorises a copy of M (lin::factorize semantics) into the table.
class line_factors {
public:
    line_factors(int kl, int ku, int n, long long lines)
    : kl{kl}, ku{ku}, n{n}, lines{lines}, ld_{2 * kl + ku + 1}
    , ab_(static_cast<std::size_t>(lines) * n * ld_), ipiv_(static_cast<std::size_t>(lines) * n) { }

    void set_matrix(long long line, const lin::band_matrix& M) {
        if (M.cols != n || M.kl != kl || M.ku != ku || M.column_size() != ld_) throw std::runtime_error("line_factors: matrix shape");
        double* dst = ab_.data() + static_cast<std::size_t>(line) * n * ld_;
        std::copy(M.full_buffer(), M.full_buffer() + static_cast<std::size_t>(n) * ld_, dst);
        device::check(adsb_band_factorize(n, kl, ku, dst, ld_, ipiv_.data() + static_cast<std::size_t>(line) * n));
        ++version_;
    }
    const double* factors() const { return ab_.data(); }
    const int* pivots() const { return ipiv_.data(); }
    unsigned long long version() const { return version_; }

    const int kl, ku, n;
    const long long lines;

private:
    int ld_;
    std::vector<double> ab_;
    std::vector<int> ipiv_;
    unsigned long long version_ = 0;
};

namespace detail {

// generalised ADS: `special` replaces the dim_data of axis `special_axis`; dims[special_axis] is ignored
template <std::size_t Rank>
void device_solve_special(lin::tensor<double, Rank>& rhs, int special_axis, const line_factors& special,
                          const dim_data* const (&dims)[Rank]) {
    static_assert(Rank == 2 || Rank == 3, "generalised ADS: 2-D or 3-D");
    if (!rhs.attached()) {
        int n[3] = {1, 1, 1};
        for (std::size_t d = 0; d < Rank; ++d) n[d] = rhs.size(static_cast<int>(d));
        rhs.attach(std::make_shared<device::context>(static_cast<int>(Rank), n));
    }
    auto& dev = *rhs.context();
    int slots[3] = {0, 0, 0};
    for (std::size_t d = 0; d < Rank; ++d) {
        if (static_cast<int>(d) == special_axis) continue;
        const auto& M = dims[d]->M;
        slots[d] = dev.factor_slot(static_cast<int>(d), M.cols, M.kl, M.ku, M.column_size(), M.full_buffer(), dims[d]->ctx.pivot());
    }
    if (!dev.line_factors_current(special_axis, &special, special.version())) {
        device::check(adsb_set_line_factors(dev.handle(), special_axis, special.kl, special.ku, special.factors(), special.pivots()));
        dev.line_factors_uploaded(special_axis, &special, special.version());
    }
    rhs.to_device();
    device::check(adsb_solve_special(dev.handle(), rhs.device_buffer(), special_axis, slots));
    rhs.device_written();
}

template <std::size_t Rank>
void device_solve(lin::tensor<double, Rank>& rhs, const dim_data* const (&dims)[Rank]) {
    if (!rhs.attached()) {
        int n[3] = {1, 1, 1};
        for (std::size_t d = 0; d < Rank; ++d) n[d] = rhs.size(static_cast<int>(d));
        rhs.attach(std::make_shared<device::context>(Rank < 2 ? 2 : static_cast<int>(Rank), n));
    }
    auto& dev = *rhs.context();
    int slots[3] = {0, 0, 0};
    for (std::size_t d = 0; d < Rank; ++d) {
        const auto& M = dims[d]->M;
        slots[d] = dev.factor_slot(static_cast<int>(d), M.cols, M.kl, M.ku, M.column_size(), M.full_buffer(),
                                   dims[d]->ctx.pivot());
    }
    rhs.to_device();
    if (Rank == 1)
        device::check(adsb_sweep(dev.handle(), rhs.device_buffer(), 0, slots[0]));
    else
        device::check(adsb_solve(dev.handle(), rhs.device_buffer(), slots));
    rhs.device_written();
}

}  // namespace detail

inline void ads_solve(lin::tensor<double, 1>& rhs, const dim_data& dim) {
    const dim_data* const dims[1] = {&dim};
    detail::device_solve(rhs, dims);
}

inline void ads_solve(lin::tensor<double, 2>& rhs, lin::tensor<double, 2>& /*buffer*/, const dim_data& dx, const dim_data& dy) {
    const dim_data* const dims[2] = {&dx, &dy};
    detail::device_solve(rhs, dims);
}

inline void ads_solve(lin::tensor<double, 3>& rhs, lin::tensor<double, 3>& /*buffer*/, const dim_data& dx, const dim_data& dy,
                      const dim_data& dz) {
    const dim_data* const dims[3] = {&dx, &dy, &dz};
    detail::device_solve(rhs, dims);
}

// generalised ADS, 3-D: the line_factors argument marks the special dimension (solver.hpp:222-226 with a callable)
inline void ads_solve(lin::tensor<double, 3>& rhs, lin::tensor<double, 3>& /*buffer*/, const line_factors& sx, const dim_data& dy,
                      const dim_data& dz) {
    const dim_data* const dims[3] = {nullptr, &dy, &dz};
    detail::device_solve_special(rhs, 0, sx, dims);
}
inline void ads_solve(lin::tensor<double, 3>& rhs, lin::tensor<double, 3>& /*buffer*/, const dim_data& dx, const line_factors& sy,
                      const dim_data& dz) {
    const dim_data* const dims[3] = {&dx, nullptr, &dz};
    detail::device_solve_special(rhs, 1, sy, dims);
}
inline void ads_solve(lin::tensor<double, 3>& rhs, lin::tensor<double, 3>& /*buffer*/, const dim_data& dx, const dim_data& dy,
                      const line_factors& sz) {
    const dim_data* const dims[3] = {&dx, &dy, nullptr};
    detail::device_solve_special(rhs, 2, sz, dims);
}
// 2-D
inline void ads_solve(lin::tensor<double, 2>& rhs, lin::tensor<double, 2>& /*buffer*/, const line_factors& sx, const dim_data& dy) {
    const dim_data* const dims[2] = {nullptr, &dy};
    detail::device_solve_special(rhs, 0, sx, dims);
}
inline void ads_solve(lin::tensor<double, 2>& rhs, lin::tensor<double, 2>& /*buffer*/, const dim_data& dx, const line_factors& sy) {
    const dim_data* const dims[2] = {&dx, nullptr};
    detail::device_solve_special(rhs, 1, sy, dims);
}

}  // namespace ads

#endif
