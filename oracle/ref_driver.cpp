// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A C-ABI driver around the UNMODIFIED reference sources under /root/reference.  It is
// compiled by oracle/Makefile into oracle/_ref/libads_ref.so together with the reference's
// own src/ads/*.cpp; nothing from the reference is copied into this repository.  The
// reference's example classes (examples/heat/heat_3d.hpp, heat_2d.hpp,
// examples/implicit/implicit.hpp, examples/scalability/test{2,3}d.hpp) are #included from
// where they lie; their private members are reached with the `#define private public`
// trick so the driver can inject an initial coefficient tensor and call
// prepare_matrices()/before_step()/step()/compute_rhs() exactly as simulation_base::run()
// (src/ads/simulation/simulation_base.cpp:11-20) would.
//
// Galois / Boost / output_manager are replaced by the stand-ins in oracle/shim (the
// element loop then runs through the sequential path, or through a std::thread pool that
// keeps the reference's for_each + synchronized contract when threads > 1).
//
// Uses: (1) pin the C restatement in oracle/ads_oracle.c, (2) generate tests/golden/*.npz
// (tests/golden/make_golden.py), (3) the `cpu_baseline` / `--impl reference` timings.

// Every standard header the reference touches must be included BEFORE the access hack.
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <memory>
#include <mutex>
#include <limits>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <string_view>
#include <thread>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include <boost/format.hpp>
#include <boost/iterator/iterator_facade.hpp>
#include <boost/range.hpp>
#include <boost/range/counting_range.hpp>
#include <galois/Timer.h>

#define private public
#define protected public
#include "ads/executor/galois.hpp"
#include "ads/quad/gauss.hpp"
#include "ads/simulation.hpp"
#include "flow/flow.hpp"
#include "heat/heat_2d.hpp"
#include "heat/heat_3d.hpp"
#include "implicit/implicit.hpp"
#include "scalability/test2d.hpp"
#include "scalability/test3d.hpp"
#undef private
#undef protected

namespace {

using clk = std::chrono::steady_clock;

double seconds_since(clk::time_point t0) {
    return std::chrono::duration<double>(clk::now() - t0).count();
}

struct cout_silencer {
    std::ostringstream sink;  // must be constructed before its rdbuf is installed
    std::streambuf* old;
    cout_silencer() : old{std::cout.rdbuf(sink.rdbuf())} { }
    ~cout_silencer() { std::cout.rdbuf(old); }
};

// 3-D extension of examples/implicit/implicit.hpp (the reference ships only the 2-D class;
// SURVEY.md section 3.5 defines the 3-D analogue): three sub-steps per step, sub-step d is
// implicit along axis d with K_d = M_d + tau*S_d, tau = dt/3, explicit in the others.
// Built exclusively from reference pieces (simulation_3d helpers, ads_solve, lin::factorize).
class implicit_3d_ref : public ads::simulation_3d {
public:
    using Base = ads::simulation_3d;
    vector_type u, u_prev;
    ads::galois_executor executor{1};
    ads::lin::band_matrix Kx, Ky, Kz;
    // separate pivot storage per K (the 2-D reference reuses x.ctx, implicit.hpp:77-82)
    ads::lin::solver_ctx cx, cy, cz;
    double tau;

    explicit implicit_3d_ref(const ads::config_3d& c)
    : Base{c}
    , u{shape()}
    , u_prev{shape()}
    , Kx{x.p, x.p, x.B.dofs()}
    , Ky{y.p, y.p, y.B.dofs()}
    , Kz{z.p, z.p, z.B.dofs()}
    , cx{Kx}
    , cy{Ky}
    , cz{Kz}
    , tau{c.steps.dt / 3.0} {
        matrix(Kx, x.basis, tau);
        matrix(Ky, y.basis, tau);
        matrix(Kz, z.basis, tau);
    }

    static void matrix(ads::lin::band_matrix& K, const ads::basis_data& d, double h) {
        for (ads::element_id e = 0; e < d.elements; ++e) {
            for (int q = 0; q < d.quad_order; ++q) {
                int first = d.first_dof(e);
                int last = d.last_dof(e);
                for (int a = 0; a + first <= last; ++a) {
                    for (int b = 0; b + first <= last; ++b) {
                        auto va = d.b[e][q][0][a], vb = d.b[e][q][0][b];
                        auto da = d.b[e][q][1][a], db = d.b[e][q][1][b];
                        K(a + first, b + first) += (va * vb + h * da * db) * d.w[q] * d.J[e];
                    }
                }
            }
        }
    }

    void prepare() {
        Base::prepare_matrices();
        ads::lin::factorize(Kx, cx);
        ads::lin::factorize(Ky, cy);
        ads::lin::factorize(Kz, cz);
    }

    void compute_rhs_dir(int d) {
        auto& rhs = u;
        zero(rhs);
        executor.for_each(elements(), [&](index_type e) {
            auto U = element_rhs();
            double J = jacobian(e);
            for (auto q : quad_points()) {
                double w = weight(q);
                value_type uu = eval_fun(u_prev, e, q);
                for (auto a : dofs_on_element(e)) {
                    auto aa = dof_global_to_local(e, a);
                    value_type v = eval_basis(e, q, a);
                    double g = 0;
                    if (d != 0) g += uu.dx * v.dx;
                    if (d != 1) g += uu.dy * v.dy;
                    if (d != 2) g += uu.dz * v.dz;
                    double val = uu.val * v.val - tau * g;
                    U(aa[0], aa[1], aa[2]) += val * w * J;
                }
            }
            executor.synchronized([&]() { update_global_rhs(rhs, U, e); });
        });
    }

    void solve_dir(int d) {
        using ads::dim_data;
        if (d == 0) ads::ads_solve(u, buffer, dim_data{Kx, cx}, y.data(), z.data());
        if (d == 1) ads::ads_solve(u, buffer, x.data(), dim_data{Ky, cy}, z.data());
        if (d == 2) ads::ads_solve(u, buffer, x.data(), y.data(), dim_data{Kz, cz});
    }

    void do_step() {
        for (int d = 0; d < 3; ++d) {
            if (d > 0) {
                using std::swap;
                swap(u, u_prev);
            }
            compute_rhs_dir(d);
            solve_dir(d);
        }
    }
};

template <typename Tensor>
void copy_in(Tensor& t, const double* src) {
    std::copy(src, src + t.size(), t.data());
}

template <typename Tensor>
void copy_out(const Tensor& t, double* dst) {
    std::copy(t.data(), t.data() + t.size(), dst);
}

// default shipped initial state: the example's own before()
struct call_before {
    template <typename Sim>
    void operator()(Sim& s) const { s.before(); }
};

// Drive a reference simulation object the way simulation_base::run() does, but starting
// from a caller-supplied coefficient tensor (init_mode 0) or the shipped before() (1).
//   stage 0: nsteps full steps; returns u
//   stage 1: one compute_rhs from u_prev := input; returns rhs (held in u)
template <typename Sim, typename RhsFn, typename InitFn>
int drive(Sim& sim, int init_mode, int stage, int nsteps, double dt, double* u, double* timings,
          RhsFn rhs_only, InitFn shipped_init) {
    cout_silencer quiet;
    auto t_setup = clk::now();
    if (init_mode == 1) {
        shipped_init(sim);
    } else {
        sim.prepare_matrices();
        copy_in(sim.u, u);
    }
    if (timings) timings[2] = seconds_since(t_setup);
    if (stage == 1) {
        using std::swap;
        swap(sim.u, sim.u_prev);
        auto t0 = clk::now();
        rhs_only(sim);
        if (timings) timings[1] = seconds_since(t0);
        copy_out(sim.u, u);
        return 0;
    }
    auto t0 = clk::now();
    for (int i = 0; i < nsteps; ++i) {
        double t = i * dt;
        sim.before_step(i, t);
        sim.step(i, t);
        sim.after_step(i, t);
    }
    if (timings) timings[0] = seconds_since(t0);
    copy_out(sim.u, u);
    return 0;
}

ads::lin::band_matrix make_band(int n, int kl, int ku, const double* ab) {
    ads::lin::band_matrix M{kl, ku, n};
    std::copy(ab, ab + static_cast<std::size_t>(M.column_size()) * n, M.full_buffer());
    return M;
}

}  // namespace

extern "C" {

int ref_set_threads(int n) {
    ads::galois_executor::thread_override() = n;
    return 0;
}

// include/ads/quad/gauss.hpp:14-15
int ref_gauss(int q, double* xs, double* ws) {
    if (q < 2 || q > 64) return -1;
    for (int i = 0; i < q; ++i) {
        xs[i] = ads::quad::gauss::Xs[q][i];
        ws[i] = ads::quad::gauss::Ws[q][i];
    }
    return 0;
}

int ref_dofs(int p, int elements) {
    ads::dim_config c{p, elements};
    ads::dimension d{c, 1};
    return d.dofs();
}

// dimension + basis_data tables (src/ads/simulation/dimension.cpp:8-21, src/ads/basis_data.cpp:63-114)
//   b_flat[e][k][d][i], xq[e][k], w[k], J[e], first_dof[e], knots[elements+2p+1]
int ref_basis_tables(int p, int elements, double a, double b, int quad_order, int ders,
                     double* b_flat, double* xq, double* w, double* J, int* first_dof,
                     double* knots) {
    ads::dim_config c{p, elements, a, b, quad_order, 0};
    ads::dimension d{c, ders};
    const auto& bd = d.basis;
    int q = bd.quad_order;
    for (int e = 0; e < bd.elements; ++e) {
        for (int k = 0; k < q; ++k) {
            for (int dd = 0; dd <= ders; ++dd)
                for (int i = 0; i <= p; ++i)
                    b_flat[((static_cast<std::size_t>(e) * q + k) * (ders + 1) + dd) * (p + 1) + i] =
                        bd.b[e][k][dd][i];
            xq[e * q + k] = bd.x[e][k];
        }
        J[e] = bd.J[e];
        first_dof[e] = bd.first_dof(e);
    }
    for (int k = 0; k < q; ++k) w[k] = bd.w[k];
    if (knots) std::copy(d.B.knot.begin(), d.B.knot.end(), knots);
    return 0;
}

// kind 0: gram (form_matrix.cpp:8-24, as built by dimension), 1: stiffness (:26-42),
// 2: advection (:44-60), 3: M + h*S "implicit" matrix (examples/implicit/implicit.hpp:46-64 uses
// 0.5*h; the caller passes the already-scaled h).  fix: bit0 = fix_left, bit1 = fix_right
// (dimension.cpp:23-29).  ab is the full LAPACK buffer, ldab = 3p+1.
int ref_matrix_1d(int kind, int p, int elements, double a, double b, double h, int fix, double* ab) {
    ads::dim_config c{p, elements, a, b};
    ads::dimension d{c, 1};
    if (kind != 0) {
        d.M.zero();
        if (kind == 1) ads::stiffness_matrix_1d(d.M, d.basis);
        if (kind == 2) ads::advection_matrix_1d(d.M, d.basis);
        if (kind == 3) implicit_3d_ref::matrix(d.M, d.basis, h);
    }
    if (fix & 1) d.fix_left();
    if (fix & 2) d.fix_right();
    std::size_t len = static_cast<std::size_t>(d.M.column_size()) * d.M.cols;
    std::copy(d.M.full_buffer(), d.M.full_buffer() + len, ab);
    return 0;
}

// lin::factorize (include/ads/lin/band_solve.hpp:16-18) -> LAPACK dgbtrf_
int ref_factorize(int n, int kl, int ku, double* ab, int* ipiv) {
    auto M = make_band(n, kl, ku, ab);
    ads::lin::solver_ctx ctx{M};
    ads::lin::factorize(M, ctx);
    std::copy(M.full_buffer(), M.full_buffer() + static_cast<std::size_t>(M.column_size()) * n, ab);
    std::copy(ctx.pivot_vector.begin(), ctx.pivot_vector.end(), ipiv);
    return ctx.info;
}

// lin::solve_with_factorized (band_solve.hpp:26-31) -> LAPACK dgbtrs_
int ref_solve_factorized(int n, int kl, int ku, const double* ab, const int* ipiv, double* b,
                         int nrhs) {
    auto M = make_band(n, kl, ku, ab);
    ads::lin::solver_ctx ctx{M};
    std::copy(ipiv, ipiv + n, ctx.pivot_vector.begin());
    ads::lin::solve_with_factorized(M, b, ctx, nrhs);
    return ctx.info;
}

// lin::cyclic_transpose (include/ads/lin/tensor/cyclic_transpose.hpp:54-63), rank 2 or 3
int ref_cyclic_transpose(int ndim, const int* n, const double* in, double* out) {
    if (ndim == 3) {
        auto a = ads::lin::as_tensor(const_cast<double*>(in), n[0], n[1], n[2]);
        ads::lin::cyclic_transpose(a, out);
    } else if (ndim == 2) {
        auto a = ads::lin::as_tensor(const_cast<double*>(in), n[0], n[1]);
        ads::lin::cyclic_transpose(a, out);
    } else {
        return -1;
    }
    return 0;
}

// ads_solve (include/ads/solver.hpp:222-226) with already-factorised 1-D matrices.
int ref_ads_solve(int ndim, const int* n, const int* kl, const int* ku, const double* const* ab,
                  const int* const* ipiv, double* rhs) {
    std::vector<ads::lin::band_matrix> Ms;
    std::vector<ads::lin::solver_ctx> ctxs;
    for (int d = 0; d < ndim; ++d) {
        Ms.push_back(make_band(n[d], kl[d], ku[d], ab[d]));
        ctxs.emplace_back(Ms.back());
        std::copy(ipiv[d], ipiv[d] + n[d], ctxs.back().pivot_vector.begin());
    }
    using ads::dim_data;
    if (ndim == 1) {
        ads::lin::tensor<double, 1> t{{n[0]}};
        copy_in(t, rhs);
        ads::ads_solve(t, dim_data{Ms[0], ctxs[0]});
        copy_out(t, rhs);
    } else if (ndim == 2) {
        ads::lin::tensor<double, 2> t{{n[0], n[1]}}, buf{{n[0], n[1]}};
        copy_in(t, rhs);
        ads::ads_solve(t, buf, dim_data{Ms[0], ctxs[0]}, dim_data{Ms[1], ctxs[1]});
        copy_out(t, rhs);
    } else if (ndim == 3) {
        ads::lin::tensor<double, 3> t{{n[0], n[1], n[2]}}, buf{{n[0], n[1], n[2]}};
        copy_in(t, rhs);
        ads::ads_solve(t, buf, dim_data{Ms[0], ctxs[0]}, dim_data{Ms[1], ctxs[1]},
                       dim_data{Ms[2], ctxs[2]});
        copy_out(t, rhs);
    } else {
        return -1;
    }
    return 0;
}

// examples/heat/heat_3d.hpp (shipped class, sequential loop, eval_fun inside the dof loop)
int ref_heat3d(int p, int n, double dt, int nsteps, int init_mode, int stage, double* u,
               double* timings) {
    ads::dim_config dim{p, n};
    ads::config_3d c{dim, dim, dim, ads::timesteps_config{nsteps, dt}, 1};
    ads::problems::heat_3d sim{c};
    return drive(sim, init_mode, stage, nsteps, dt, u, timings,
                 [](ads::problems::heat_3d& s) { s.compute_rhs(); }, call_before{});
}

// examples/heat/heat_2d.hpp (fix_left + Dirichlet row overwrite before each solve)
int ref_heat2d(int p, int n, double dt, int nsteps, int init_mode, int stage, double* u,
               double* timings) {
    ads::dim_config dim{p, n};
    ads::config_2d c{dim, dim, ads::timesteps_config{nsteps, dt}, 1};
    ads::problems::heat_2d sim{c};
    return drive(sim, init_mode, stage, nsteps, dt, u, timings,
                 [](ads::problems::heat_2d& s) { s.compute_rhs(); }, call_before{});
}

// examples/implicit/implicit.hpp; stage 1 -> compute_rhs_1, stage 2 -> compute_rhs_2
int ref_implicit2d(int p, int n, double dt, int nsteps, int init_mode, int stage, double* u,
                   double* timings) {
    ads::dim_config dim{p, n};
    ads::config_2d c{dim, dim, ads::timesteps_config{nsteps, dt}, 1};
    ads::implicit_2d sim{c, 1 << 30};
    // implicit_2d::before() (implicit.hpp:84-95) minus its energy()/L2norm() printout
    auto init = [](ads::implicit_2d& s) {
        s.prepare_matrices();
        s.projection(s.u, [&s](double x, double y) { return s.init_state(x, y); });
        s.solve(s.u);
    };
    if (stage == 2) {
        return drive(sim, init_mode, 1, nsteps, dt, u, timings,
                     [](ads::implicit_2d& s) { s.compute_rhs_2(); }, init);
    }
    return drive(sim, init_mode, stage, nsteps, dt, u, timings,
                 [](ads::implicit_2d& s) { s.compute_rhs_1(); }, init);
}

// examples/scalability/test3d.hpp (fix_left, forcing, hoisted eval_fun, element_rhs scatter)
int ref_scalability3d(int p, int n, double dt, int nsteps, int init_mode, int stage, double* u,
                      double* timings) {
    ads::dim_config dim{p, n};
    ads::config_3d c{dim, dim, dim, ads::timesteps_config{nsteps, dt}, 1};
    ads::problems::scalability_3d sim{c, 1};
    int rc = drive(sim, init_mode, stage, nsteps, dt, u, timings,
                   [](ads::problems::scalability_3d& s) { s.compute_rhs(); }, call_before{});
    if (timings) timings[1] = static_cast<double>(sim.integration_timer.get_usec()) * 1e-6;
    return rc;
}

// examples/scalability/test2d.hpp
int ref_scalability2d(int p, int n, double dt, int nsteps, int init_mode, int stage, double* u,
                      double* timings) {
    ads::dim_config dim{p, n};
    ads::config_2d c{dim, dim, ads::timesteps_config{nsteps, dt}, 1};
    ads::problems::scalability_2d sim{c, 1};
    int rc = drive(sim, init_mode, stage, nsteps, dt, u, timings,
                   [](ads::problems::scalability_2d& s) { s.compute_rhs(); }, call_before{});
    if (timings) timings[1] = static_cast<double>(sim.integration_timer.get_usec()) * 1e-6;
    return rc;
}

// examples/flow/flow.hpp: nonlinear pointwise form with the permeability tabulated at the quadrature points.
// kq_out (may be NULL) receives that table, x fastest: kq[(ex*q+kx) + nqx*((ey*q+ky) + nqy*(ez*q+kz))].
int ref_flow(int p, int n, double dt, int nsteps, int init_mode, int stage, double* u, double* kq_out,
             double* timings) {
    ads::dim_config dim{p, n};
    ads::config_3d c{dim, dim, dim, ads::timesteps_config{nsteps, dt}, 1};
    ads::problems::flow sim{c};
    {
        cout_silencer quiet;
        sim.fill_permeability_map();
    }
    if (kq_out) {
        const int q = p + 1;
        const std::size_t nq = static_cast<std::size_t>(n) * q;
        for (int ez = 0; ez < n; ++ez)
            for (int ey = 0; ey < n; ++ey)
                for (int ex = 0; ex < n; ++ex)
                    for (int kz = 0; kz < q; ++kz)
                        for (int ky = 0; ky < q; ++ky)
                            for (int kx = 0; kx < q; ++kx)
                                kq_out[(ex * q + kx) + nq * ((ey * q + ky) + nq * (static_cast<std::size_t>(ez) * q + kz))] =
                                    sim.kq(ex, ey, ez, kx, ky, kz);
    }
    if (stage == 1 || init_mode == 1)
        return drive(sim, init_mode, stage, stage == 1 ? nsteps : 0, dt, u, timings,
                     [](ads::problems::flow& s) { s.compute_rhs(0.0); }, call_before{}) ||
               (stage == 1 ? 0 : [&] {
                   copy_in(sim.u, u);
                   for (int i = 0; i < nsteps; ++i) {
                       sim.before_step(i, i * dt);
                       sim.step(i, i * dt);
                   }
                   copy_out(sim.u, u);
                   return 0;
               }());
    // after_step() only prints the energy and writes files (flow.hpp:116-123); its iostream formatting is
    // left out of the step loop here
    sim.prepare_matrices();
    copy_in(sim.u, u);
    for (int i = 0; i < nsteps; ++i) {
        sim.before_step(i, i * dt);
        sim.step(i, i * dt);
    }
    copy_out(sim.u, u);
    return 0;
}

// 3-D implicit extension (see implicit_3d_ref above); stage 1..3 -> compute_rhs_dir(stage-1)
int ref_implicit3d(int p, int n, double dt, int nsteps, int init_mode, int stage, double* u,
                   double* timings) {
    ads::dim_config dim{p, n};
    ads::config_3d c{dim, dim, dim, ads::timesteps_config{nsteps, dt}, 1};
    implicit_3d_ref sim{c};
    cout_silencer quiet;
    sim.prepare();
    if (init_mode == 1) {
        auto init = [](double x, double y, double z) {
            double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
            double r2 = std::min(12 * (dx * dx + dy * dy + dz * dz), 1.0);
            return (r2 - 1) * (r2 - 1) * (r2 + 1) * (r2 + 1);
        };
        sim.projection(sim.u, init);
        sim.solve(sim.u);
    } else {
        copy_in(sim.u, u);
    }
    if (stage >= 1) {
        using std::swap;
        swap(sim.u, sim.u_prev);
        sim.compute_rhs_dir(stage - 1);
        copy_out(sim.u, u);
        return 0;
    }
    auto t0 = clk::now();
    for (int i = 0; i < nsteps; ++i) {
        using std::swap;
        swap(sim.u, sim.u_prev);
        sim.do_step();
    }
    if (timings) timings[0] = seconds_since(t0);
    copy_out(sim.u, u);
    return 0;
}

// Split timing of the reference hot path for the CPU baseline: nsteps x {compute_rhs, solve}
// on the scalability_3d class without forcing cost removed (it is the reference's own
// benchmark harness, examples/scalability/main.cpp:8-32).  timings: [0]=total, [1]=rhs, [3]=solve
int ref_time_heat3d_hoisted(int p, int n, double dt, int nsteps, double* u, double* timings) {
    ads::dim_config dim{p, n};
    ads::config_3d c{dim, dim, dim, ads::timesteps_config{nsteps, dt}, 1};
    ads::problems::scalability_3d sim{c, 1};
    cout_silencer quiet;
    sim.prepare_matrices();
    copy_in(sim.u, u);
    double t_rhs = 0, t_solve = 0;
    auto t0 = clk::now();
    for (int i = 0; i < nsteps; ++i) {
        sim.before_step(i, i * dt);
        auto a = clk::now();
        sim.compute_rhs();
        t_rhs += seconds_since(a);
        a = clk::now();
        sim.solve(sim.u);
        t_solve += seconds_since(a);
    }
    timings[0] = seconds_since(t0);
    timings[1] = t_rhs;
    timings[3] = t_solve;
    copy_out(sim.u, u);
    return 0;
}

}  // extern "C"
