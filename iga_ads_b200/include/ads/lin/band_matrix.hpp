// ads/lin/band_matrix.hpp, solver_ctx, factorize, solve_with_factorized -- LAPACK general-band
// storage with factor workspace exactly as the reference (include/ads/lin/band_matrix.hpp:18-79,
// solver_ctx.hpp:14-30, band_solve.hpp:16-31).  factorize runs on the host (adsb_band_factorize =
// dgbtrf semantics); the substitutions run on the device.
#ifndef ADSB_ADS_LIN_BAND_MATRIX_HPP
#define ADSB_ADS_LIN_BAND_MATRIX_HPP

#include <algorithm>
#include <vector>

#include "ads/device.hpp"
#include "ads/lin/tensor.hpp"

namespace ads::lin {

class band_matrix {
public:
    int kl, ku, rows, cols, row_offset;

    band_matrix(int kl, int ku, int n) : band_matrix{kl, ku, n, n, kl} { }
    band_matrix(int kl, int ku, int rows, int cols, int row_offset = 0)
    : kl{kl}, ku{ku}, rows{rows}, cols{cols}, row_offset{row_offset}
    , data_(static_cast<std::size_t>(column_size()) * cols) { }

    int column_size() const { return row_offset + kl + ku + 1; }  // ldab
    double& operator()(int i, int j) { return data_[static_cast<std::size_t>(j) * column_size() + row_offset + ku + i - j]; }
    double operator()(int i, int j) const { return data_[static_cast<std::size_t>(j) * column_size() + row_offset + ku + i - j]; }
    double* full_buffer() { return data_.data(); }
    const double* full_buffer() const { return data_.data(); }
    void zero() { std::fill(data_.begin(), data_.end(), 0.0); }

private:
    std::vector<double> data_;
};

struct solver_ctx {
    std::vector<int> pivot_vector;
    int info = 0;
    int lda;
    solver_ctx(int n, int lda) : pivot_vector(n), lda{lda} { }
    explicit solver_ctx(const band_matrix& a) : solver_ctx{a.cols, a.column_size()} { }
    int* pivot() { return pivot_vector.data(); }
    const int* pivot() const { return pivot_vector.data(); }
};

inline void factorize(band_matrix& a, solver_ctx& ctx) {
    const int rc = adsb_band_factorize(a.cols, a.kl, a.ku, a.full_buffer(), a.column_size(), ctx.pivot());
    ctx.info = rc == ADSB_ESINGULAR ? 1 : 0;  // like the reference, a singular factor is recorded, not thrown
    if (rc < 0 && rc != ADSB_ESINGULAR) device::check(rc);
}

// Solve along the first index of `rhs` for all its lines (dgbtrs with nrhs = size / size(0)).
template <std::size_t Rank>
void solve_with_factorized(const band_matrix& a, tensor<double, Rank>& rhs, solver_ctx& ctx) {
    int n[3] = {rhs.size(0), rhs.size() / rhs.size(0), 1};
    auto dev = std::make_shared<device::context>(2, n);
    device::check(adsb_set_axis_factor(dev->handle(), 0, 0, a.cols, a.kl, a.ku, a.column_size(), a.full_buffer(), ctx.pivot()));
    const int buf = dev->new_buffer();
    device::check(adsb_upload(dev->handle(), buf, rhs.data()));
    device::check(adsb_sweep(dev->handle(), buf, 0, 0));
    device::check(adsb_download(dev->handle(), buf, rhs.data()));
}

template <std::size_t Rank>
void solve(band_matrix& a, tensor<double, Rank>& rhs) {
    solver_ctx ctx{a};
    factorize(a, ctx);
    solve_with_factorized(a, rhs, ctx);
}

}  // namespace ads::lin

#endif
