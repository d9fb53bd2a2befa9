// implicit_2d -- examples/implicit/implicit.hpp + main.cpp: Peaceman-Rachford-style splitting, two half
// steps per step, K = M + dt/2 S implicit along one axis (implicit.hpp:46-64,:77-82,:97-111).
#include <cstdio>
#include <cstdlib>

#include "ads/simulation.hpp"

namespace ads::problems {

class implicit_2d : public simulation_2d {
    using Base = simulation_2d;
    vector_type u, u_prev;
    lin::band_matrix Kx, Ky;
    lin::solver_ctx Kx_ctx, Ky_ctx;  // the reference factorises K into x.ctx; separate contexts keep both pivot vectors

public:
    explicit implicit_2d(const config_2d& config)
    : Base{config}, u{shape()}, u_prev{shape()}, Kx{x.p, x.p, x.dofs()}, Ky{y.p, y.p, y.dofs()}, Kx_ctx{Kx}, Ky_ctx{Ky} { }

    double init_state(double px, double py) {
        double dx = px - 0.5, dy = py - 0.5;
        double r2 = std::min(12 * (dx * dx + dy * dy), 1.0);
        return (r2 - 1) * (r2 - 1) * (r2 + 1) * (r2 + 1);
    }
    const vector_type& solution() const { return u; }

private:
    void before() override {
        prepare_matrices();
        const double h = 0.5 * steps.dt;
        form_matrix_1d(Kx, 3, x.p, x.elements, x.a, x.b, h);
        form_matrix_1d(Ky, 3, y.p, y.elements, y.a, y.b, h);
        lin::factorize(Kx, Kx_ctx);
        lin::factorize(Ky, Ky_ctx);
        projection(u, [this](double a, double b) { return init_state(a, b); });
        solve(u);
    }

    void step(int /*iter*/, double /*t*/) override {
        using std::swap;
        const double h = 0.5 * steps.dt;
        swap(u, u_prev);
        Base::compute_rhs(make_form(1.0, {0.0, h, 0.0}), u_prev, u);   // compute_rhs_1: explicit in y
        ads_solve(u, buffer, dim_data{Kx, Kx_ctx}, y.data());
        swap(u, u_prev);
        Base::compute_rhs(make_form(1.0, {h, 0.0, 0.0}), u_prev, u);   // compute_rhs_2: explicit in x
        ads_solve(u, buffer, x.data(), dim_data{Ky, Ky_ctx});
    }
};

}  // namespace ads::problems

int main(int argc, char* argv[]) {
    const int p = argc > 1 ? std::atoi(argv[1]) : 2;
    const int n = argc > 2 ? std::atoi(argv[2]) : 40;
    const int nsteps = argc > 3 ? std::atoi(argv[3]) : 10;
    const double dt = argc > 4 ? std::atof(argv[4]) : 1e-2;
    ads::dim_config dim{p, n};
    ads::config_2d c{dim, dim, ads::timesteps_config{nsteps, dt}, 1};
    ads::problems::implicit_2d sim{c};
    sim.run();
    const auto& u = sim.solution();
    double sum = 0;
    for (int i = 0; i < u.size(); ++i) sum += u.data()[i];
    std::printf("implicit_2d p=%d n=%d steps=%d dt=%g: sum(u) = %.14f\n", p, n, nsteps, dt, sum);
}
