// implicit_3d -- the 3-D analogue of examples/implicit/implicit.hpp (BASELINE.json configs[3]; the reference ships the
// 2-D class, SURVEY.md 3.5 defines the extension): three sub-steps per step, sub-step d implicit along axis d with
// K_d = M_d + (dt/3) S_d and explicit along the other two (implicit.hpp:46-64,:77-82,:97-111 per axis).
//     implicit_3d [p] [elements] [steps] [dt]
#include <cstdio>
#include <cstdlib>

#include "ads/simulation.hpp"

namespace ads::problems {

class implicit_3d : public simulation_3d {
    using Base = simulation_3d;
    vector_type u, u_prev;
    lin::band_matrix Kx, Ky, Kz;
    lin::solver_ctx Kx_ctx, Ky_ctx, Kz_ctx;

public:
    explicit implicit_3d(const config_3d& config)
    : Base{config}, u{shape()}, u_prev{shape()}, Kx{x.p, x.p, x.dofs()}, Ky{y.p, y.p, y.dofs()}, Kz{z.p, z.p, z.dofs()}
    , Kx_ctx{Kx}, Ky_ctx{Ky}, Kz_ctx{Kz} { }

    double init_state(double px, double py, double pz) {  // implicit.hpp:38-43 with the third coordinate
        double dx = px - 0.5, dy = py - 0.5, dz = pz - 0.5;
        double r2 = std::min(12 * (dx * dx + dy * dy + dz * dz), 1.0);
        return (r2 - 1) * (r2 - 1) * (r2 + 1) * (r2 + 1);
    }
    const vector_type& solution() const { return u; }

private:
    void before() override {
        prepare_matrices();
        const double tau = steps.dt / 3.0;
        form_matrix_1d(Kx, 3, x.p, x.elements, x.a, x.b, tau);
        form_matrix_1d(Ky, 3, y.p, y.elements, y.a, y.b, tau);
        form_matrix_1d(Kz, 3, z.p, z.elements, z.a, z.b, tau);
        lin::factorize(Kx, Kx_ctx);
        lin::factorize(Ky, Ky_ctx);
        lin::factorize(Kz, Kz_ctx);
        projection(u, [this](double a, double b, double c) { return init_state(a, b, c); });
        solve(u);
    }

    void step(int /*iter*/, double /*t*/) override {
        using std::swap;
        const double tau = steps.dt / 3.0;
        swap(u, u_prev);
        Base::compute_rhs(make_form(1.0, {0.0, tau, tau}), u_prev, u);   // explicit in y, z
        ads_solve(u, buffer, dim_data{Kx, Kx_ctx}, y.data(), z.data());
        swap(u, u_prev);
        Base::compute_rhs(make_form(1.0, {tau, 0.0, tau}), u_prev, u);   // explicit in x, z
        ads_solve(u, buffer, x.data(), dim_data{Ky, Ky_ctx}, z.data());
        swap(u, u_prev);
        Base::compute_rhs(make_form(1.0, {tau, tau, 0.0}), u_prev, u);   // explicit in x, y
        ads_solve(u, buffer, x.data(), y.data(), dim_data{Kz, Kz_ctx});
    }
};

}  // namespace ads::problems

int main(int argc, char* argv[]) {
    const int p = argc > 1 ? std::atoi(argv[1]) : 3;
    const int n = argc > 2 ? std::atoi(argv[2]) : 8;
    const int nsteps = argc > 3 ? std::atoi(argv[3]) : 2;
    const double dt = argc > 4 ? std::atof(argv[4]) : 1e-2;
    ads::dim_config dim{p, n};
    ads::config_3d c{dim, dim, dim, ads::timesteps_config{nsteps, dt}, 1};
    ads::problems::implicit_3d sim{c};
    sim.run();
    const auto& u = sim.solution();
    double sum = 0;
    for (int i = 0; i < u.size(); ++i) sum += u.data()[i];
    std::printf("implicit_3d p=%d n=%d steps=%d dt=%g: sum(u) = %.14f\n", p, n, nsteps, dt, sum);
}
