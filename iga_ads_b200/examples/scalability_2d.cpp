// scalability_2d -- examples/scalability/test2d.hpp + main.cpp: x.fix_left(), u = 1 then solve, forcing term
// evaluated at the quadrature points (test2d.hpp:49-54) -> the general quadrature kernel.
//     scalability_2d [p] [elements] [steps]
#include <cstdio>
#include <cstdlib>

#include "ads/simulation.hpp"

namespace ads::problems {

class scalability_2d : public simulation_2d {
    using Base = simulation_2d;
    vector_type u, u_prev;

public:
    explicit scalability_2d(const config_2d& config) : Base{config}, u{shape()}, u_prev{shape()} { }
    const vector_type& solution() const { return u; }

private:
    void before() override {
        x.fix_left();
        prepare_matrices();
        for (int i = 0; i < u.size(); ++i) u.data()[i] = 1.0;
        solve(u);
    }
    void before_step(int /*iter*/, double /*t*/) override {
        using std::swap;
        swap(u, u_prev);
    }
    void step(int /*iter*/, double /*t*/) override {
        const double dt = steps.dt;
        Base::compute_rhs(make_form(1.0, {dt, dt, 0.0}, ADSB_RHS_QUADRATURE, dt, -1, 1), u_prev, u);
        solve(u);
    }
};

}  // namespace ads::problems

int main(int argc, char* argv[]) {
    const int p = argc > 1 ? std::atoi(argv[1]) : 3;
    const int n = argc > 2 ? std::atoi(argv[2]) : 16;
    const int nsteps = argc > 3 ? std::atoi(argv[3]) : 3;
    ads::dim_config dim{p, n};
    ads::config_2d c{dim, dim, ads::timesteps_config{nsteps, 1e-6}, 1};
    ads::problems::scalability_2d sim{c};
    sim.run();
    const auto& u = sim.solution();
    double sum = 0;
    for (int i = 0; i < u.size(); ++i) sum += u.data()[i];
    std::printf("scalability_2d p=%d n=%d steps=%d: sum(u) = %.14f\n", p, n, nsteps, sum);
}
