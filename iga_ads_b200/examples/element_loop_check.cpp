// element_loop_check -- a host-side compute_rhs() in the reference's idiom (the structure of
// examples/scalability/test3d.hpp:66-95: executor.for_each over elements(), element_rhs(), eval_fun / eval_basis /
// grad_dot, executor.synchronized + update_global_rhs), compiled against these headers and run ON THE HOST: the
// part of the class surface an example needs to keep its own element loops.  No GPU is touched.
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

    // The element loop in the reference's idiom: every element integrates into its own local tensor, the executor
    // serialises the scatter.  Integrand of the scalability example: u v - dt (grad u . grad v - f).
    const vector_type& compute_rhs() {
        vector_type& out = u;
        zero(out);
        const double dt = steps.dt;
        auto integrate_element = [&](index_type elem) {
            vector_type local = element_rhs();
            const double jac = jacobian(elem);
            for (auto qp : quad_points()) {
                const double wj = weight(qp) * jac;
                const auto pt = point(elem, qp);
                const double f = forcing(pt[0], pt[1], pt[2]);
                const value_type prev = eval_fun(u_prev, elem, qp);
                for (auto dof : dofs_on_element(elem)) {
                    const value_type test = eval_basis(elem, qp, dof);
                    const index_type at = dof_global_to_local(elem, dof);
                    local(at[0], at[1], at[2]) += (prev.val * test.val - dt * (grad_dot(prev, test) - f)) * wj;
                }
            }
            executor.synchronized([&] { update_global_rhs(out, local, elem); });
        };
        executor.for_each(elements(), integrate_element);
        return out;
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
