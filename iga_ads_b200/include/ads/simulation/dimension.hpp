// ads/simulation/dimension.hpp -- one axis: basis + Gram matrix + quadrature tables + LU context
// (reference: include/ads/simulation/dimension.hpp:20-59, src/ads/simulation/dimension.cpp:8-29).
#ifndef ADSB_ADS_SIMULATION_DIMENSION_HPP
#define ADSB_ADS_SIMULATION_DIMENSION_HPP

#include <algorithm>

#include "ads/basis_data.hpp"
#include "ads/lin/band_matrix.hpp"
#include "ads/simulation/config.hpp"

namespace ads {

struct dim_data {  // include/ads/solver.hpp:17-20
    const lin::band_matrix& M;
    lin::solver_ctx& ctx;
};

// 1-D quadrature matrices in band storage (src/ads/form_matrix.cpp:8-60); kind 0 Gram, 1 stiffness,
// 2 advection, 3 Gram + h * stiffness (examples/implicit/implicit.hpp:46-64)
inline void form_matrix_1d(lin::band_matrix& M, int kind, int p, int elements, double a, double b, double h = 0.0) {
    device::check(adsb_matrix_1d(kind, p, elements, a, b, h, 0, M.full_buffer()));
}

struct dimension {
    int p;
    int elements;
    double a;
    double b;
    bspline::basis B;
    lin::band_matrix M;
    basis_data basis;
    lin::solver_ctx ctx;

    dimension(const dim_config& config, int derivatives)
    : p{config.p}, elements{config.elements}, a{config.a}, b{config.b}
    , B{bspline::create_basis(a, b, p, elements, config.repeated_nodes)}
    , M{p, p, B.dofs()}
    , basis{B, derivatives, config.quad_order, a, b}
    , ctx{M} {
        form_matrix_1d(M, 0, p, elements, a, b);  // gram_matrix_1d(M, basis)
    }

    int dofs() const { return B.dofs(); }
    dim_data data() { return {M, ctx}; }

    void fix_dof(int k) {  // src/ads/simulation/dimension.cpp:23-29
        const int last = dofs() - 1;
        for (int i = std::max(k - p, 0); i <= std::min(k + p, last); ++i) M(k, i) = 0;
        M(k, k) = 1;
    }
    void fix_left() { fix_dof(0); }
    void fix_right() { fix_dof(dofs() - 1); }
    void factorize_matrix() { lin::factorize(M, ctx); }
};

}  // namespace ads

#endif
