// ads/bspline/bspline.hpp -- knot vectors and basis evaluation (reference: include/ads/bspline/bspline.hpp,
// src/ads/bspline/bspline.cpp:26-43,:61-81,:102-160), forwarded to libadsb200's host entry points.
#ifndef ADSB_ADS_BSPLINE_HPP
#define ADSB_ADS_BSPLINE_HPP

#include <vector>

#include "ads/device.hpp"

namespace ads::bspline {

struct basis {
    std::vector<double> knot;
    int degree = 0;

    int knot_size() const { return static_cast<int>(knot.size()); }
    int dofs() const { return knot_size() - degree - 1; }
    int elements() const { return dofs() - degree; }
    double begin() const { return knot.front(); }
    double end() const { return knot.back(); }
};

inline basis create_basis(double a, double b, int p, int elements, int repeated_nodes = 0) {
    if (repeated_nodes != 0) throw std::runtime_error("libadsb200: repeated knots are outside the ADS-step path");
    basis B;
    B.degree = p;
    B.knot.resize(elements + 2 * p + 1);
    device::check(adsb_knots(p, elements, a, b, B.knot.data()));
    return B;
}

inline int find_span(double x, const basis& b) { return adsb_find_span(x, b.knot.data(), b.knot_size(), b.degree); }

// out[d][i], d = 0..ders, i = 0..p
inline void eval_basis_with_derivatives(int span, double x, const basis& b, double* out, int ders) {
    device::check(adsb_basis_ders(span, x, b.knot.data(), b.degree, ders, out));
}

}  // namespace ads::bspline

#endif
