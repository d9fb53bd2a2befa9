// ads/basis_data.hpp -- per-axis quadrature tables (reference: include/ads/basis_data.hpp:21-86,
// src/ads/basis_data.cpp:63-114): b[e][q][d][i], x[e][q], w[q], J[e].  Stored flat (the layout the
// device wants) with accessor functions instead of the reference's double**** pointers.
#ifndef ADSB_ADS_BASIS_DATA_HPP
#define ADSB_ADS_BASIS_DATA_HPP

#include <vector>

#include "ads/bspline/bspline.hpp"

namespace ads {

struct basis_data {
    int degree = 0, elements = 0, dofs = 0, derivatives = 0, quad_order = 0;
    std::vector<int> first_dofs;
    std::vector<double> b_flat, x_flat, w_, J_;

    basis_data() = default;
    basis_data(const bspline::basis& B, int derivatives, int quad_order, double a, double b)
    : degree{B.degree}, elements{B.elements()}, dofs{B.dofs()}, derivatives{derivatives}, quad_order{quad_order}
    , first_dofs(elements), b_flat(static_cast<std::size_t>(elements) * quad_order * (derivatives + 1) * (degree + 1))
    , x_flat(static_cast<std::size_t>(elements) * quad_order), w_(quad_order), J_(elements) {
        device::check(adsb_basis_tables(degree, elements, a, b, quad_order, derivatives, b_flat.data(), x_flat.data(),
                                        w_.data(), J_.data(), first_dofs.data()));
    }

    int first_dof(int e) const { return first_dofs[e]; }
    int last_dof(int e) const { return first_dofs[e] + degree; }
    int dofs_per_element() const { return degree + 1; }
    double b(int e, int q, int d, int i) const {
        return b_flat[((static_cast<std::size_t>(e) * quad_order + q) * (derivatives + 1) + d) * (degree + 1) + i];
    }
    double x(int e, int q) const { return x_flat[static_cast<std::size_t>(e) * quad_order + q]; }
    double w(int q) const { return w_[q]; }
    double J(int e) const { return J_[e]; }
};

}  // namespace ads

#endif
