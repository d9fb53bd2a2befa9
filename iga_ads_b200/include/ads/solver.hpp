// ads/solver.hpp -- ads_solve(rhs, buffer, dims...): the alternating-direction solve
// (reference: include/ads/solver.hpp:35-41,:148-160,:200-226).  The reference does, per axis,
// dgbtrs on the contiguous lines followed by a cyclic rotation of the tensor into `buffer`; here all
// sweeps run on the device against the one canonical layout (adsb_solve), `buffer` is not needed.
// Only the standard variant (every dimension a dim_data) is provided.
#ifndef ADSB_ADS_SOLVER_HPP
#define ADSB_ADS_SOLVER_HPP

#include <memory>

#include "ads/lin/tensor.hpp"
#include "ads/simulation/dimension.hpp"

namespace ads {

namespace detail {

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

}  // namespace ads

#endif
