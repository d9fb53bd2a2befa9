// Stand-in for the reference's include/ads/output_manager.hpp -- TEST INFRASTRUCTURE ONLY.
// The real header needs Boost.Format and writes gnuplot/VTK files; output is outside the
// hot path, so the oracle build swallows it.
#ifndef ADSB_SHIM_ADS_OUTPUT_MANAGER_HPP
#define ADSB_SHIM_ADS_OUTPUT_MANAGER_HPP

#include <cstddef>
#include <fstream>

#include <boost/format.hpp>

namespace ads {

template <std::size_t Dim>
struct output_manager {
    template <typename... Args>
    explicit output_manager(Args&&...) { }

    template <typename... Args>
    void to_file(Args&&...) { }
};

}  // namespace ads

#endif
