// ads/slabs.hpp -- a 3-D ADS simulation sharded into z-slabs over the GPUs of one box, driven from one C++17
// process (adsb_slabs_* of libadsb200; csrc/slab_host.cpp).  The reference has no counterpart (single address
// space, simulation_base::run: src/ads/simulation/simulation_base.cpp:11-20); the class takes what a
// simulation_3d already holds -- its three ads::dimension objects with factorised matrices -- and a sub-step
// program (right-hand-side form + factor slots per sub-step), and advances the state on `world` ranks:
// devices {0, 1, ...} when the box has that many GPUs, otherwise `world` virtual ranks sharing device 0.
#ifndef ADSB_ADS_SLABS_HPP
#define ADSB_ADS_SLABS_HPP

#include <array>
#include <vector>

#include "ads/device.hpp"
#include "ads/lin/tensor.hpp"
#include "ads/simulation/dimension.hpp"

namespace ads {

class slab_cluster {
public:
    // devices.size() ranks; rank r runs on devices[r] (all equal: virtual ranks on that device)
    slab_cluster(dimension& x, dimension& y, dimension& z, const std::vector<int>& devices) {
        dimension* d[3] = {&x, &y, &z};
        int n[3];
        for (int a = 0; a < 3; ++a) n[a] = d[a]->dofs();
        device::check(adsb_slabs_create(static_cast<int>(devices.size()), devices.data(), n, &h_));
        for (int a = 0; a < 3; ++a) {
            const basis_data& b = d[a]->basis;
            device::check(adsb_slabs_set_axis_tables(h_, a, b.degree, b.elements, b.quad_order, b.derivatives, b.b_flat.data(),
                                                     b.x_flat.data(), b.w_.data(), b.J_.data(), b.first_dofs.data()));
            set_factor(a, 0, d[a]->M, d[a]->ctx);  // slot 0: the dimension's own (factorised) matrix
        }
    }
    ~slab_cluster() { adsb_slabs_destroy(h_); }
    slab_cluster(const slab_cluster&) = delete;
    slab_cluster& operator=(const slab_cluster&) = delete;

    // another factorised matrix of `axis` (K = M + h S of an implicit sub-step) in `slot`
    void set_factor(int axis, int slot, const lin::band_matrix& M, const lin::solver_ctx& ctx) {
        device::check(adsb_slabs_set_axis_factor(h_, axis, slot, M.cols, M.kl, M.ku, M.column_size(), M.full_buffer(), ctx.pivot()));
    }
    // the sub-steps of one time step; fixes the slab bounds and builds every rank
    void commit(const std::vector<adsb_substep>& program) {
        device::check(adsb_slabs_commit(h_, program.data(), static_cast<int>(program.size())));
    }
    void set_state(lin::tensor<double, 3>& u) { device::check(adsb_slabs_upload(h_, u.data())); }
    void get_state(lin::tensor<double, 3>& u) {
        device::check(adsb_slabs_download(h_, u.data()));
    }
    void advance(int steps) {
        device::check(adsb_slabs_step(h_, steps));
        device::check(adsb_slabs_synchronize(h_));
    }
    std::vector<int> bounds() const {
        int info[4];
        device::check(adsb_slabs_info(h_, nullptr, info));
        std::vector<int> b(info[0] + 1);
        device::check(adsb_slabs_info(h_, b.data(), info));
        return b;
    }
    bool virtual_ranks() const {
        int info[4];
        device::check(adsb_slabs_info(h_, nullptr, info));
        return info[1] != 0;
    }

private:
    adsb_slabs* h_ = nullptr;
};

}  // namespace ads

#endif
