// generalised_ads -- ads_solve with a special dimension (include/ads/solver.hpp:56-96,:170-195): every line along y
// has its own matrix M + h(ix, iz) S, the shape of examples/maxwell/maxwell_ads.hpp:139-163,:189-203; x and z use the
// Gram factors.  Prints a checksum of the solution of a fixed right-hand side.
//     generalised_ads [p] [elements]
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "ads/simulation.hpp"

int main(int argc, char* argv[]) {
    const int p = argc > 1 ? std::atoi(argv[1]) : 2;
    const int n = argc > 2 ? std::atoi(argv[2]) : 10;
    ads::dim_config cfg{p, n};
    ads::dimension x{cfg, 1}, y{cfg, 1}, z{cfg, 1};
    x.factorize_matrix();
    z.factorize_matrix();
    const int nd = x.dofs();
    ads::line_factors By{p, p, nd, static_cast<long long>(nd) * nd};
    for (int iz = 0; iz < nd; ++iz)
        for (int ix = 0; ix < nd; ++ix) {
            ads::lin::band_matrix K{p, p, nd};
            const double h = 1e-3 * (1 + ix + 2 * iz) / nd;   // a coefficient that differs from line to line
            ads::form_matrix_1d(K, 3, p, n, 0.0, 1.0, h);
            By.set_matrix(ix + static_cast<long long>(nd) * iz, K);
        }
    ads::lin::tensor<double, 3> rhs{{nd, nd, nd}}, buffer{{nd, nd, nd}};
    for (int k = 0; k < nd; ++k)
        for (int j = 0; j < nd; ++j)
            for (int i = 0; i < nd; ++i) rhs(i, j, k) = std::sin(0.3 * i) + 0.5 * std::cos(0.2 * j) + 0.1 * k;
    ads::ads_solve(rhs, buffer, x.data(), By, z.data());
    double sum = 0;
    for (int i = 0; i < rhs.size(); ++i) sum += rhs.data()[i];
    std::printf("generalised_ads p=%d n=%d: sum(x) = %.12e\n", p, n, sum);
}
