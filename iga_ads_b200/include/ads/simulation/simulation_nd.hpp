// ads/simulation/simulation_nd.hpp -- simulation_2d / simulation_3d: the toolkit user problems
// inherit from (reference: include/ads/simulation/simulation_2d.hpp:24-140, simulation_3d.hpp:23-150,
// src/ads/simulation/simulation_3d.cpp:8-22).  Same member names and meaning; what changed:
//   * solve(v) runs adsb_solve on the device mirror of v (no rotation buffer needed);
//   * compute_rhs(form, u_prev, rhs) evaluates one of the device right-hand sides instead of a host
//     element loop; the per-element helpers (elements(), quad_points(), eval_basis, eval_fun, ...)
//     are still there for set-up code and diagnostics written the reference's way.
#ifndef ADSB_ADS_SIMULATION_ND_HPP
#define ADSB_ADS_SIMULATION_ND_HPP

#include <array>
#include <memory>
#include <vector>

#include "ads/lin/tensor.hpp"
#include "ads/simulation/dimension.hpp"
#include "ads/simulation/simulation_base.hpp"
#include "ads/solver.hpp"

namespace ads {

// value + gradient (include/ads/util/function_value/function_value_3d.hpp:9-55)
struct function_value_2d {
    double val = 0, dx = 0, dy = 0;
};
struct function_value_3d {
    double val = 0, dx = 0, dy = 0, dz = 0;
};

// The constant-coefficient form  alpha (u, v) - sum_k beta_k (d_k u, d_k v) + gamma F  of adsb_form.
inline adsb_form make_form(double alpha, std::array<double, 3> beta, int method = ADSB_RHS_COLLAPSED, double gamma = 0.0,
                           int forcing_buf = -1, int source = 0) {
    adsb_form f{};
    f.alpha = alpha;
    for (int d = 0; d < 3; ++d) f.beta[d] = beta[d];
    f.gamma = gamma;
    f.forcing_buf = forcing_buf;
    f.method = method;
    f.source = source;
    return f;
}

// lexicographic index range, first index slowest (include/ads/util/iter/product.hpp:80-150)
template <std::size_t N>
std::vector<std::array<int, N>> index_product(const std::array<int, N>& lo, const std::array<int, N>& hi) {
    std::vector<std::array<int, N>> out;
    std::array<int, N> i = lo;
    for (std::size_t d = 0; d < N; ++d)
        if (lo[d] >= hi[d]) return out;
    while (true) {
        out.push_back(i);
        std::size_t d = N;
        while (d-- > 0) {
            if (++i[d] < hi[d]) break;
            i[d] = lo[d];
            if (d == 0) return out;
        }
    }
}

template <std::size_t D>
class simulation_nd : public simulation_base {
public:
    using vector_type = lin::tensor<double, D>;
    using index_type = std::array<int, D>;
    using value_type = std::conditional_t<D == 2, function_value_2d, function_value_3d>;

protected:
    std::vector<dimension> dims_;
    vector_type buffer;
    std::shared_ptr<device::context> dev_;

    explicit simulation_nd(std::vector<dimension> dims, const timesteps_config& steps)
    : simulation_base{steps}, dims_{std::move(dims)}, buffer{shape_of(dims_)} { }

    static std::array<int, D> shape_of(const std::vector<dimension>& dims) {
        std::array<int, D> s{};
        for (std::size_t d = 0; d < D; ++d) s[d] = dims[d].dofs();
        return s;
    }

    device::context& dev() {
        if (!dev_) {
            int n[3] = {1, 1, 1};
            for (std::size_t d = 0; d < D; ++d) n[d] = dims_[d].dofs();
            dev_ = std::make_shared<device::context>(static_cast<int>(D), n);
            for (std::size_t d = 0; d < D; ++d) {
                const auto& bd = dims_[d].basis;
                device::check(adsb_set_axis_tables(dev_->handle(), static_cast<int>(d), bd.degree, bd.elements, bd.quad_order,
                                                   bd.derivatives, bd.b_flat.data(), bd.x_flat.data(), bd.w_.data(),
                                                   bd.J_.data(), bd.first_dofs.data()));
            }
        }
        return *dev_;
    }
    void on_device(vector_type& v) {
        dev();
        v.attach(dev_);
    }

public:
    std::array<int, D> shape() const { return shape_of(dims_); }

    void prepare_matrices() {
        for (auto& d : dims_) d.factorize_matrix();
    }

    // ads_solve(rhs, buffer, x.data(), y.data()[, z.data()])  (simulation_3d.hpp:41)
    void solve(vector_type& rhs) {
        on_device(rhs);
        if constexpr (D == 2)
            ads_solve(rhs, buffer, dims_[0].data(), dims_[1].data());
        else
            ads_solve(rhs, buffer, dims_[0].data(), dims_[1].data(), dims_[2].data());
    }

    // rhs <- form(u_prev) on the device; replaces the element loop of the examples' compute_rhs()
    void compute_rhs(const adsb_form& form, vector_type& u_prev, vector_type& rhs) {
        on_device(u_prev);
        on_device(rhs);
        u_prev.to_device();
        device::check(adsb_compute_rhs(dev().handle(), &form, u_prev.device_buffer(), rhs.device_buffer()));
        rhs.device_written();
    }

    // L2-projection right-hand side  u_a = sum_e sum_q f(x_q) B_a(x_q) w J  (include/ads/projection.hpp:12-153).
    // f is a host callable, so it is evaluated on the host -- once per quadrature point, in slabs of z elements
    // that keep the table of values below ~256 MB -- and the sum itself runs on the device
    // (adsb_project_values; one slab reproduces the reference's summation order exactly).  No list of element
    // index tuples is built (the reference's elements() range at 512^3 would be 1.6 GB as a vector).
    template <typename Function>
    void projection(vector_type& v, Function f) {
        on_device(v);
        std::size_t nq[3] = {1, 1, 1};
        for (std::size_t d = 0; d < D; ++d) nq[d] = static_cast<std::size_t>(dims_[d].elements) * dims_[d].basis.quad_order;
        const int qz = D == 3 ? dims_[D - 1].basis.quad_order : 1;
        const int nez = D == 3 ? dims_[D - 1].elements : 1;
        const std::size_t per_elem = nq[0] * nq[1] * qz;
        const int slab = static_cast<int>(std::max<std::size_t>(1, std::min<std::size_t>(nez, (std::size_t{1} << 25) / per_elem)));
        std::vector<double> tab(per_elem * slab);
        for (int e0 = 0; e0 < nez; e0 += slab) {
            const int cnt = std::min(slab, nez - e0);
            for (int kz = 0; kz < cnt * qz; ++kz) {
                const double z = D == 3 ? dims_[D - 1].basis.x_flat[static_cast<std::size_t>(e0) * qz + kz] : 0.0;
                for (std::size_t ky = 0; ky < nq[1]; ++ky) {
                    const double y = dims_[1].basis.x_flat[ky];
                    double* row = tab.data() + (static_cast<std::size_t>(kz) * nq[1] + ky) * nq[0];
                    for (std::size_t kx = 0; kx < nq[0]; ++kx) {
                        const double x = dims_[0].basis.x_flat[kx];
                        if constexpr (D == 2)
                            row[kx] = f(x, y);
                        else
                            row[kx] = f(x, y, z);
                    }
                }
            }
            device::check(adsb_project_values(dev().handle(), v.device_buffer(), e0, cnt, tab.data(), e0 > 0 ? 1 : 0));
        }
        v.device_written();
    }

    // basic_simulation_Nd::normL2 / normH1 (include/ads/simulation/basic_simulation_3d.hpp:314-330) on the device
    double normL2(vector_type& u) { return norm_impl(u, 0, 0, 0.0, nullptr)[0]; }
    double normH1(vector_type& u) { return norm_impl(u, 1, 0, 0.0, nullptr)[0]; }
    // errorL2 against a host callable (basic_simulation_3d.hpp:380-398): f is tabulated at the quadrature points
    template <typename Function>
    double errorL2(vector_type& u, Function f) {
        std::size_t nq[3] = {1, 1, 1};
        for (std::size_t d = 0; d < D; ++d) nq[d] = static_cast<std::size_t>(dims_[d].elements) * dims_[d].basis.quad_order;
        std::vector<double> tab(nq[0] * nq[1] * nq[2]);
        for (std::size_t kz = 0; kz < nq[2]; ++kz)
            for (std::size_t ky = 0; ky < nq[1]; ++ky)
                for (std::size_t kx = 0; kx < nq[0]; ++kx) {
                    const double x = dims_[0].basis.x_flat[kx], y = dims_[1].basis.x_flat[ky];
                    double& t = tab[kx + nq[0] * (ky + nq[1] * kz)];
                    if constexpr (D == 2)
                        t = f(x, y);
                    else
                        t = f(x, y, dims_[2].basis.x_flat[kz]);
                }
        return norm_impl(u, 0, 2, 0.0, tab.data())[0];
    }
    // {error, norm of the reference} against the validation solution sin(pi x) sin(pi y) [sin(pi z)] exp(-d pi^2 t)
    // (examples/validation/validation.hpp:45-55,:121-129); h1: H1 instead of L2
    std::array<double, 2> error_validation(vector_type& u, double t, bool h1) { return norm_impl(u, h1 ? 1 : 0, 1, t, nullptr); }

    // output_manager<D>::evaluate (include/ads/output_manager.hpp:66-73,:101-118): the spline on intervals + 1
    // points per axis; values come back first index fastest, the order the reference's writers print them in
    std::vector<double> sample(vector_type& u, int intervals) {
        on_device(u);
        u.to_device();
        int npts[3] = {1, 1, 1};
        std::vector<std::vector<double>> pts(D), kn(D);
        const double* pp[3] = {nullptr, nullptr, nullptr};
        const double* kp[3] = {nullptr, nullptr, nullptr};
        std::size_t total = 1;
        for (std::size_t d = 0; d < D; ++d) {
            npts[d] = intervals + 1;
            pts[d].resize(npts[d]);
            for (int i = 0; i <= intervals; ++i) {  // ads::linspace -> lerp(i, n, a, b)
                const double t = static_cast<double>(i) / static_cast<double>(intervals);
                pts[d][i] = (1 - t) * dims_[d].a + t * dims_[d].b;
            }
            kn[d].resize(dims_[d].elements + 2 * dims_[d].p + 1);
            device::check(adsb_knots(dims_[d].p, dims_[d].elements, dims_[d].a, dims_[d].b, kn[d].data()));
            pp[d] = pts[d].data();
            kp[d] = kn[d].data();
            total *= static_cast<std::size_t>(npts[d]);
        }
        std::vector<double> out(total);
        device::check(adsb_sample(dev().handle(), u.device_buffer(), npts, pp, kp, out.data()));
        return out;
    }

private:
    std::array<double, 2> norm_impl(vector_type& u, int kind, int ref, double t, const double* tab) {
        on_device(u);
        u.to_device();
        std::array<double, 2> out{};
        device::check(adsb_norm(dev().handle(), u.device_buffer(), kind, ref, t, tab, out.data()));
        return out;
    }

public:
    // ---- per-element helpers (simulation_3d.hpp:64-136, simulation_2d.hpp:61-131)
    std::vector<index_type> elements() const {
        index_type lo{}, hi{};
        for (std::size_t d = 0; d < D; ++d) hi[d] = dims_[d].elements;
        return index_product<D>(lo, hi);
    }
    std::vector<index_type> quad_points() const {
        index_type lo{}, hi{};
        for (std::size_t d = 0; d < D; ++d) hi[d] = dims_[d].basis.quad_order;
        return index_product<D>(lo, hi);
    }
    std::vector<index_type> dofs_on_element(index_type e) const {
        index_type lo{}, hi{};
        for (std::size_t d = 0; d < D; ++d) {
            lo[d] = dims_[d].basis.first_dof(e[d]);
            hi[d] = dims_[d].basis.last_dof(e[d]) + 1;
        }
        return index_product<D>(lo, hi);
    }
    double jacobian(index_type e) const {
        double J = 1;
        for (std::size_t d = 0; d < D; ++d) J *= dims_[d].basis.J(e[d]);
        return J;
    }
    double weight(index_type q) const {
        double w = 1;
        for (std::size_t d = 0; d < D; ++d) w *= dims_[d].basis.w(q[d]);
        return w;
    }
    std::array<double, D> point(index_type e, index_type q) const {
        std::array<double, D> x{};
        for (std::size_t d = 0; d < D; ++d) x[d] = dims_[d].basis.x(e[d], q[d]);
        return x;
    }
    index_type dof_global_to_local(index_type e, index_type a) const {
        index_type loc{};
        for (std::size_t d = 0; d < D; ++d) loc[d] = a[d] - dims_[d].basis.first_dof(e[d]);
        return loc;
    }
    value_type eval_basis(index_type e, index_type q, index_type a) const {
        const index_type loc = dof_global_to_local(e, a);
        double B[D], dB[D];
        for (std::size_t d = 0; d < D; ++d) {
            B[d] = dims_[d].basis.b(e[d], q[d], 0, loc[d]);
            dB[d] = dims_[d].basis.b(e[d], q[d], 1, loc[d]);
        }
        value_type v;
        if constexpr (D == 2) {
            v.val = B[0] * B[1];
            v.dx = dB[0] * B[1];
            v.dy = B[0] * dB[1];
        } else {
            v.val = B[0] * B[1] * B[2];
            v.dx = dB[0] * B[1] * B[2];
            v.dy = B[0] * dB[1] * B[2];
            v.dz = B[0] * B[1] * dB[2];
        }
        return v;
    }
    value_type eval_fun(const vector_type& v, index_type e, index_type q) const {
        value_type u;
        for (auto b : dofs_on_element(e)) {
            const double c = v_at(v, b);
            const value_type B = eval_basis(e, q, b);
            u.val += c * B.val;
            u.dx += c * B.dx;
            u.dy += c * B.dy;
            if constexpr (D == 3) u.dz += c * B.dz;
        }
        return u;
    }
    // element-local right-hand side and its scatter (simulation_3d.hpp:138-145, simulation_2d.hpp:133-140): what an
    // example's own host-side compute_rhs() loop uses -- zero(rhs); for e: U = element_rhs(); ...;
    // update_global_rhs(rhs, U, e).  Such a loop compiles and runs unchanged against these headers (on the host;
    // the tensor's device copy is refreshed on the next device call); the device forms replace it for speed.
    index_type local_shape() const {
        index_type s{};
        for (std::size_t d = 0; d < D; ++d) s[d] = dims_[d].basis.dofs_per_element();
        return s;
    }
    vector_type element_rhs() const { return vector_type{local_shape()}; }
    void update_global_rhs(vector_type& global, const vector_type& local, index_type e) const {
        for (auto a : dofs_on_element(e)) v_at(global, a) += v_at(local, dof_global_to_local(e, a));
    }

    double grad_dot(const value_type& a, const value_type& b) const {
        if constexpr (D == 2)
            return a.dx * b.dx + a.dy * b.dy;
        else
            return a.dx * b.dx + a.dy * b.dy + a.dz * b.dz;
    }

private:
    static double& v_at(vector_type& v, index_type a) {
        if constexpr (D == 2)
            return v(a[0], a[1]);
        else
            return v(a[0], a[1], a[2]);
    }
    static double v_at(const vector_type& v, index_type a) {
        if constexpr (D == 2)
            return v(a[0], a[1]);
        else
            return v(a[0], a[1], a[2]);
    }
};

class simulation_2d : public simulation_nd<2> {
public:
    dimension &x, &y;
    explicit simulation_2d(const config_2d& c)
    : simulation_nd<2>{{dimension{c.x, c.derivatives}, dimension{c.y, c.derivatives}}, c.steps}, x{dims_[0]}, y{dims_[1]} { }
};

class simulation_3d : public simulation_nd<3> {
public:
    dimension &x, &y, &z;
    explicit simulation_3d(const config_3d& c)
    : simulation_nd<3>{{dimension{c.x, c.derivatives}, dimension{c.y, c.derivatives}, dimension{c.z, c.derivatives}}, c.steps}
    , x{dims_[0]}, y{dims_[1]}, z{dims_[2]} { }
};

}  // namespace ads

#endif
