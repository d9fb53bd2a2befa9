// heat_3d -- the reference example (examples/heat/heat_3d.hpp) on top of libadsb200: same class,
// same constructor, same before()/before_step()/step() structure.  Only compute_rhs() differs: the
// element loop of heat_3d.hpp:49-67 is replaced by the device form  (u, v) - dt (grad u, grad v).
#ifndef ADSB_EXAMPLES_HEAT_3D_HPP
#define ADSB_EXAMPLES_HEAT_3D_HPP

#include <algorithm>

#include "ads/simulation.hpp"

namespace ads::problems {

class heat_3d : public simulation_3d {
private:
    using Base = simulation_3d;
    vector_type u, u_prev;
    int method;

public:
    explicit heat_3d(const config_3d& config, int method = ADSB_RHS_COLLAPSED)
    : Base{config}, u{shape()}, u_prev{shape()}, method{method} { }

    double init_state(double x, double y, double z) {
        double dx = x - 0.5;
        double dy = y - 0.5;
        double dz = z - 0.5;
        double r2 = std::min(8 * (dx * dx + dy * dy + dz * dz), 1.0);
        return (r2 - 1) * (r2 - 1) * (r2 + 1) * (r2 + 1);
    };

    const vector_type& solution() const { return u; }

private:
    void before() override {
        prepare_matrices();

        auto init = [this](double x, double y, double z) { return init_state(x, y, z); };
        projection(u, init);
        solve(u);
    }

    void before_step(int /*iter*/, double /*t*/) override {
        using std::swap;
        swap(u, u_prev);
    }

    void step(int /*iter*/, double /*t*/) override {
        compute_rhs();
        solve(u);
    }

    void compute_rhs() {
        auto& rhs = u;
        const double dt = steps.dt;
        Base::compute_rhs(make_form(1.0, {dt, dt, dt}, method), u_prev, rhs);
    }
};

}  // namespace ads::problems

#endif
