// element_loop_check -- a reference-style compute_rhs() written exactly as examples/scalability/test3d.hpp:66-95
// writes it (executor.for_each over elements(), element_rhs(), eval_fun / eval_basis / grad_dot,
// executor.synchronized + update_global_rhs), compiled against these headers and run ON THE HOST: the part of the
// class surface an unchanged example needs for its own element loops.  No GPU is touched.
//     element_loop_check [p] [elements] [threads]     prints sum(rhs) and |rhs|_2 for the synthetic input
//     u_prev(i, j, k) = sin(0.3 i) + 0.5 cos(0.2 j) + 0.1 k   (tests compare with the oracle on the same input)
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "ads/simulation.hpp"

namespace {

class scalability_host : public ads::simulation_3d {
    using Base = ads::simulation_3d;
    vector_type u, u_prev;
    ads::galois_executor executor;

public:
    scalability_host(const ads::config_3d& c, int threads) : Base{c}, u{shape()}, u_prev{shape()}, executor{threads} {
        for (int k = 0; k < u_prev.size(2); ++k)
            for (int j = 0; j < u_prev.size(1); ++j)
                for (int i = 0; i < u_prev.size(0); ++i) u_prev(i, j, k) = std::sin(0.3 * i) + 0.5 * std::cos(0.2 * j) + 0.1 * k;
    }
    void step(int, double) override { }

    double forcing(double x, double y, double z) const {  // test3d.hpp:58-64
        const double pi = 3.14159265358979323846;
        double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
        double r = std::sqrt(dx * dx + dy * dy + dz * dz);
        return std::exp(-r) + 1 + std::cos(pi * x) * std::cos(pi * y) * std::cos(pi * z);
    }

    const vector_type& compute_rhs() {  // test3d.hpp:66-95
        auto& rhs = u;
        zero(rhs);
        executor.for_each(elements(), [&](index_type e) {
            auto U = element_rhs();
            double J = jacobian(e);
            for (auto q : quad_points()) {
                double w = weight(q);
                auto x = point(e, q);
                value_type u = eval_fun(u_prev, e, q);
                for (auto a : dofs_on_element(e)) {
                    auto aa = dof_global_to_local(e, a);
                    value_type v = eval_basis(e, q, a);
                    double gradient_prod = grad_dot(u, v);
                    double val = u.val * v.val - steps.dt * (gradient_prod - forcing(x[0], x[1], x[2]));
                    U(aa[0], aa[1], aa[2]) += val * w * J;
                }
            }
            executor.synchronized([&]() { update_global_rhs(rhs, U, e); });
        });
        return rhs;
    }
};

}  // namespace

int main(int argc, char* argv[]) {
    const int p = argc > 1 ? std::atoi(argv[1]) : 2;
    const int n = argc > 2 ? std::atoi(argv[2]) : 6;
    const int threads = argc > 3 ? std::atoi(argv[3]) : 1;
    ads::dim_config dim{p, n};
    ads::config_3d c{dim, dim, dim, ads::timesteps_config{1, 1e-6}, 1};
    scalability_host sim{c, threads};
    const auto& rhs = sim.compute_rhs();
    // the same data through a tensor_view
    auto view = ads::lin::as_tensor(rhs.data(), rhs.sizes());
    double sum = 0, sq = 0;
    for (int k = 0; k < view.size(2); ++k)
        for (int j = 0; j < view.size(1); ++j)
            for (int i = 0; i < view.size(0); ++i) {
                sum += view(i, j, k);
                sq += view(i, j, k) * view(i, j, k);
            }
    std::printf("element loop p=%d n=%d threads=%d: sum(rhs) = %.15e  |rhs|_2 = %.15e\n", p, n, threads, sum, std::sqrt(sq));
}
