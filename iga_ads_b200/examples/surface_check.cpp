// surface_check -- exercises the parts of the C++17 host layer the examples do not: value semantics of
// lin::tensor with a device mirror (a copy taken while the device copy is the newer one must hold the current
// values), buffer-id recycling, device projection of an arbitrary callable, norms and output sampling.
//     surface_check [elements]      prints "surface_check OK" and a few numbers (exit code 0), or throws
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

#include "ads/simulation.hpp"

namespace {

struct probe : ads::simulation_3d {
    using Base = ads::simulation_3d;
    vector_type u, v;
    explicit probe(const ads::config_3d& c) : Base{c}, u{shape()}, v{shape()} { }
    void step(int, double) override { }

    void require(bool ok, const char* what) {
        if (!ok) throw std::runtime_error(std::string("surface_check: ") + what);
    }

    void run_checks() {
        prepare_matrices();
        const double pi = 3.14159265358979323846;
        auto f = [pi](double x, double y, double z) { return std::sin(pi * x) * std::sin(pi * y) * std::sin(pi * z) + 0.25 * x; };
        projection(u, f);      // device: adsb_project_values
        solve(u);              // u_h = L2 projection of f
        const double e = errorL2(u, f), n = normL2(u);
        std::cout << "L2 error of the projection = " << e << ", |u_h| = " << n << "\n";
        require(e < 2e-3 * n, "L2 projection error too large");
        // copy while the device copy is the newer one: the copy must hold the solved values, not stale host data
        vector_type snapshot = u;
        require(std::abs(snapshot(3, 4, 5) - u(3, 4, 5)) == 0.0, "copy taken from a device-resident tensor is stale");
        // the source moves on; the snapshot must not follow it
        compute_rhs(ads::make_form(1.0, {1e-3, 1e-3, 1e-3}), u, v);
        solve(v);
        using std::swap;
        swap(u, v);
        require(snapshot(3, 4, 5) != u(3, 4, 5), "snapshot aliases the tensor it was copied from");
        // temporaries attach and release managed buffers: more of them than the context has buffers
        for (int k = 0; k < 3 * ADSB_MAX_BUFFERS; ++k) {
            vector_type tmp = snapshot;
            solve(tmp);
        }
        const auto vals = sample(u, 8);
        require(vals.size() == 9u * 9u * 9u, "sample size");
        double s = 0;
        for (double x : vals) s += x;
        std::cout << "sum of 9^3 samples = " << s << ", H1 norm = " << normH1(u) << "\n";
        require(std::isfinite(s), "samples not finite");
    }
};

}  // namespace

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 12;
    ads::dim_config dim{2, n};
    ads::config_3d c{dim, dim, dim, ads::timesteps_config{1, 1e-3}, 1};
    probe p{c};
    p.run_checks();
    std::cout << "surface_check OK\n";
    return 0;
}
