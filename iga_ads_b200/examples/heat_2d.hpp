// heat_2d -- examples/heat/heat_2d.hpp: x.fix_left(), Dirichlet row v(0, i) = proj(sin pi y)(i)
// written into the right-hand side before every solve (heat_2d.hpp:40-52), zero initial state.
#ifndef ADSB_EXAMPLES_HEAT_2D_HPP
#define ADSB_EXAMPLES_HEAT_2D_HPP

#include <cmath>

#include "ads/simulation.hpp"

namespace ads::problems {

class heat_2d : public simulation_2d {
private:
    using Base = simulation_2d;
    vector_type u, u_prev;
    std::vector<double> row;  // 1-D projection of sin(pi y)

public:
    explicit heat_2d(const config_2d& config) : Base{config}, u{shape()}, u_prev{shape()}, row(y.dofs()) { }

    const vector_type& solution() const { return u; }

private:
    void prepare() {
        x.fix_left();
        prepare_matrices();
        // compute_projection(buf, y.basis, sin(pi y))  (include/ads/projection.hpp:12-36)
        const auto& bd = y.basis;
        for (int e = 0; e < bd.elements; ++e)
            for (int q = 0; q < bd.quad_order; ++q)
                for (int a = 0; a <= bd.degree; ++a)
                    row[bd.first_dof(e) + a] += std::sin(bd.x(e, q) * M_PI) * bd.b(e, q, 0, a) * bd.w(q) * bd.J(e);
    }

    void apply_bc(vector_type& v) {
        on_device(v);
        v.to_device();
        device::check(adsb_set_plane(v.context()->handle(), v.device_buffer(), 0, 0, row.data()));
        v.device_written();
    }

    void before() override {
        prepare();
        zero(u);
        apply_bc(u);
        solve(u);
    }

    void before_step(int /*iter*/, double /*t*/) override {
        using std::swap;
        swap(u, u_prev);
    }

    void step(int /*iter*/, double /*t*/) override {
        const double dt = steps.dt;
        Base::compute_rhs(make_form(1.0, {dt, dt, 0.0}), u_prev, u);
        apply_bc(u);
        solve(u);
    }
};

}  // namespace ads::problems

#endif
