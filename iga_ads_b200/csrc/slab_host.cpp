// slab_host.cpp -- the z-slab sharded ADS step driven from ONE C++17 host process (adsb_slabs_*).
//
// The reference is single-address-space C++ (simulation_base::run, src/ads/simulation/simulation_base.cpp:11-20);
// this is the same loop for several GPUs of one box without any Python or torch in the process: rank r owns the
// planes [bounds[r], bounds[r+1]) on device devices[r], neighbouring devices map each other's memory
// (cudaDeviceEnablePeerAccess), and per sub-step every rank runs
//     right-hand side (p halo planes of both neighbours) -> x sweep -> y sweep -> fused distributed z sweep
// exactly as iga_ads_b200/slab.py does with one process per GPU.  Because all streams belong to
// one process, the end-of-sub-step neighbour barrier is a pair of cudaStreamWaitEvent's instead of the flag kernel.
// When all ranks name the same device they share it as concurrent streams with an equal share of the SMs each
// ("virtual ranks": the fused sweeps of neighbouring ranks wait for one another inside the kernel, so they must be
// resident together) -- the single-GPU test vehicle of this file.
// Built on the public C ABI only (adsb_create, adsb_rhs_view, adsb_sweep_view, adsb_dist_sweep_view, ...).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "adsb200.h"
#include "internal.hpp"

namespace {

struct HostFactor {
    int n, kl, ku, ldab;
    std::vector<double> ab;
    std::vector<int> ipiv;
};
struct HostTables {
    bool set = false;
    int p, elements, q, ders;
    std::vector<double> b, xq, w, J;
    std::vector<int> first;
};

struct Rank {
    int device = 0;
    adsb_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;      // recorded after the rank's last kernel of a sub-step
    int z0 = 0, cz = 0;
    double* H[2] = {nullptr, nullptr};  // state buffers with p halo planes on both sides
    double* dseg = nullptr;
    double* xst = nullptr;
    int* err = nullptr;
    bool forcing = false;
};

}  // namespace

struct adsb_slabs {
    int world = 0;
    int ng[3] = {1, 1, 1};
    bool virtual_ranks = false;
    HostTables tab[3];
    std::map<std::pair<int, int>, HostFactor> fac;
    std::vector<Rank> r;
    std::vector<int> bounds;
    std::vector<adsb_substep> prog;
    bool committed = false;
    int p = 0, KL = 0, KD = 0, nl = 64, lag = 4, cur = 0;
    long long pitch = 0, plane = 0, lines = 0;
    long long launches = 0;
};

using adsb::fail;

namespace {

int cuda_fail(cudaError_t e, const char* what) {
    return fail(ADSB_ENODEVICE, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                            \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)
#define OK(call)                  \
    do {                          \
        int rc_ = (call);         \
        if (rc_ < 0) return rc_;  \
    } while (0)

adsb_view view_of(const adsb_slabs* s, int nx, int planes) {
    adsb_view v{};
    v.n[0] = nx;
    v.n[1] = s->ng[1];
    v.n[2] = planes;
    v.s[0] = 1;
    v.s[1] = s->pitch;
    v.s[2] = s->plane;
    return v;
}

double* interior(const adsb_slabs* s, const Rank& R, int k) { return R.H[k] + (long long) s->p * s->plane; }

// boundary planes of state buffer k -> the neighbours' halo regions (set-up only; inside a step the fused sweep
// stores them itself)
int publish(adsb_slabs* s, int k) {
    const long long pl = s->plane;
    const int p = s->p;
    for (int q = 0; q < s->world; ++q) {
        Rank& R = s->r[q];
        CU(cudaSetDevice(R.device));
        if (q > 0) {
            const Rank& L = s->r[q - 1];
            CU(cudaMemcpyAsync(L.H[k] + (long long) (p + L.cz) * pl, R.H[k] + (long long) p * pl, sizeof(double) * p * pl,
                               cudaMemcpyDefault, R.stream));
        }
        if (q + 1 < s->world) {
            const Rank& N = s->r[q + 1];
            CU(cudaMemcpyAsync(N.H[k], R.H[k] + (long long) R.cz * pl, sizeof(double) * p * pl, cudaMemcpyDefault, R.stream));
        }
    }
    for (auto& R : s->r) {
        CU(cudaSetDevice(R.device));
        CU(cudaStreamSynchronize(R.stream));
    }
    return ADSB_OK;
}

}  // namespace

extern "C" {

int adsb_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

int adsb_slabs_create(int world, const int* devices, const int* n_global, adsb_slabs** out) {
    if (!out || !devices || !n_global || world < 1 || world > 64) return fail(ADSB_EINVAL, "slabs_create: bad argument");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(ADSB_ENODEVICE, "slabs_create: no CUDA device (libadsb200 has no CPU fallback)");
    auto* s = new adsb_slabs;
    s->world = world;
    for (int d = 0; d < 3; ++d) s->ng[d] = n_global[d];
    s->r.resize(world);
    s->virtual_ranks = world > 1;
    for (int q = 0; q < world; ++q) {
        if (devices[q] < 0 || devices[q] >= count) {
            delete s;
            return fail(ADSB_EINVAL, "slabs_create: bad device ordinal");
        }
        s->r[q].device = devices[q];
        if (devices[q] != devices[0]) s->virtual_ranks = false;
    }
    if (!s->virtual_ranks)
        for (int q = 0; q < world; ++q)
            for (int t = q + 1; t < world; ++t)
                if (devices[q] == devices[t]) {
                    delete s;
                    return fail(ADSB_EINVAL, "slabs_create: ranks must be on distinct devices, or all on one (virtual ranks)");
                }
    *out = s;
    return ADSB_OK;
}

int adsb_slabs_destroy(adsb_slabs* s) {
    if (!s) return ADSB_OK;
    for (auto& R : s->r) {
        cudaSetDevice(R.device);
        if (R.stream) cudaStreamSynchronize(R.stream);
        if (R.ctx) adsb_destroy(R.ctx);
        cudaFree(R.H[0]);
        cudaFree(R.H[1]);
        cudaFree(R.dseg);
        cudaFree(R.xst);
        cudaFree(R.err);
        if (R.done) cudaEventDestroy(R.done);
        if (R.stream) cudaStreamDestroy(R.stream);
    }
    delete s;
    return ADSB_OK;
}

int adsb_slabs_set_axis_tables(adsb_slabs* s, int axis, int p, int elements, int q, int ders, const double* b_flat,
                               const double* xq, const double* w, const double* J, const int* first_dof) {
    if (!s || axis < 0 || axis > 2 || !b_flat || !xq || !w || !J || !first_dof) return fail(ADSB_EINVAL, "slabs_set_axis_tables: bad argument");
    if (s->committed) return fail(ADSB_ESTATE, "slabs: already committed");
    HostTables& t = s->tab[axis];
    t.set = true;
    t.p = p;
    t.elements = elements;
    t.q = q;
    t.ders = ders;
    t.b.assign(b_flat, b_flat + (size_t) elements * q * (ders + 1) * (p + 1));
    t.xq.assign(xq, xq + (size_t) elements * q);
    t.w.assign(w, w + q);
    t.J.assign(J, J + elements);
    t.first.assign(first_dof, first_dof + elements);
    return ADSB_OK;
}

int adsb_slabs_set_axis_factor(adsb_slabs* s, int axis, int slot, int n, int kl, int ku, int ldab, const double* ab,
                               const int* ipiv) {
    if (!s || axis < 0 || axis > 2 || slot < 0 || slot >= ADSB_MAX_SLOTS || !ab || !ipiv) return fail(ADSB_EINVAL, "slabs_set_axis_factor: bad argument");
    if (s->committed) return fail(ADSB_ESTATE, "slabs: already committed");
    if (n != s->ng[axis]) return fail(ADSB_EINVAL, "slabs_set_axis_factor: n != n_global[axis]");
    HostFactor f{n, kl, ku, ldab, std::vector<double>(ab, ab + (size_t) ldab * n), std::vector<int>(ipiv, ipiv + n)};
    s->fac[{axis, slot}] = std::move(f);
    return ADSB_OK;
}

int adsb_slabs_commit(adsb_slabs* s, const adsb_substep* prog, int nsub) {
    if (!s || !prog || nsub < 1) return fail(ADSB_EINVAL, "slabs_commit: bad argument");
    if (s->committed) return fail(ADSB_ESTATE, "slabs: already committed");
    for (int d = 0; d < 3; ++d)
        if (!s->tab[d].set) return fail(ADSB_ESTATE, "slabs_commit: axis tables missing");
    s->prog.assign(prog, prog + nsub);
    s->p = s->tab[2].p;
    const int W = s->world, p = s->p, nx = s->ng[0], ny = s->ng[1], nz = s->ng[2];
    s->pitch = nx + (nx & 1);
    s->plane = s->pitch * ny;
    s->lines = s->plane;
    // slabs = segments of the z lines of the first z factor in the program
    std::vector<int> zslots;
    for (const auto& sub : s->prog)
        if (std::find(zslots.begin(), zslots.end(), sub.slots[2]) == zslots.end()) zslots.push_back(sub.slots[2]);
    for (const auto& sub : s->prog)
        for (int d = 0; d < 3; ++d)
            if (!s->fac.count({d, sub.slots[d]})) return fail(ADSB_ESTATE, "slabs_commit: a factor named by the program is missing");
    const HostFactor& z0f = s->fac[{2, zslots[0]}];
    s->bounds.assign(W + 1, 0);
    if (W > 1)
        OK(adsb_segment_bounds(nz, z0f.kl, z0f.ipiv.data(), W, 1, s->bounds.data()));
    else
        s->bounds[1] = nz;
    int rows = 0;
    for (int q = 0; q < W; ++q) rows = std::max(rows, s->bounds[q + 1] - s->bounds[q]);
    for (int q = 0; q < W; ++q)
        if (s->bounds[q + 1] - s->bounds[q] < std::max(p, 1)) return fail(ADSB_EINVAL, "slabs_commit: slabs thinner than the spline degree");
    // peer access between neighbours
    if (!s->virtual_ranks)
        for (int q = 0; q + 1 < W; ++q) {
            const int a = s->r[q].device, b = s->r[q + 1].device;
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, a, b));
            if (!can) return fail(ADSB_ENODEVICE, "slabs_commit: neighbouring devices cannot map each other's memory");
            CU(cudaSetDevice(a));
            cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
            CU(cudaSetDevice(b));
            e = cudaDeviceEnablePeerAccess(a, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->r[0].device));
    bool any_forcing = false;
    for (const auto& sub : s->prog) any_forcing = any_forcing || sub.form.gamma != 0.0;
    for (int q = 0; q < W; ++q) {
        Rank& R = s->r[q];
        R.z0 = s->bounds[q];
        R.cz = s->bounds[q + 1] - s->bounds[q];
        CU(cudaSetDevice(R.device));
        CU(cudaStreamCreateWithFlags(&R.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&R.done, cudaEventDisableTiming));
        const int lo[3] = {0, 0, R.z0}, cnt[3] = {nx, ny, R.cz};
        OK(adsb_create(3, s->ng, lo, cnt, R.device, &R.ctx));
        OK(adsb_set_stream(R.ctx, R.stream));
        if (s->virtual_ranks) OK(adsb_set_sm_limit(R.ctx, std::max(1, sms / W)));
        for (int d = 0; d < 3; ++d) {
            const HostTables& t = s->tab[d];
            OK(adsb_set_axis_tables(R.ctx, d, t.p, t.elements, t.q, t.ders, t.b.data(), t.xq.data(), t.w.data(), t.J.data(), t.first.data()));
        }
        for (const auto& kv : s->fac) {
            const HostFactor& f = kv.second;
            OK(adsb_set_axis_factor(R.ctx, kv.first.first, kv.first.second, f.n, f.kl, f.ku, f.ldab, f.ab.data(), f.ipiv.data()));
        }
        for (int slot : zslots) {
            if (W > 1) OK(adsb_set_axis_segments(R.ctx, 2, slot, W, s->bounds.data(), q, 1));
            int info[8] = {};
            if (W > 1) {
                OK(adsb_segment_info(R.ctx, 2, slot, info));
                if (info[2] != 1 || info[3] != 1)
                    return fail(ADSB_EINVAL, "slabs_commit: the z factor couples more than neighbouring slabs at this slab "
                                             "thickness (chain depth > 1): not handled by the C++ slab host");
                s->KL = std::max(s->KL, info[0]);
                s->KD = std::max(s->KD, info[1]);
            }
        }
        const size_t nh = (size_t) (rows + 2 * p) * s->plane;
        for (int k = 0; k < 2; ++k) {
            CU(cudaMalloc((void**) &R.H[k], nh * sizeof(double)));
            CU(cudaMemsetAsync(R.H[k], 0, nh * sizeof(double), R.stream));
        }
        CU(cudaMalloc((void**) &R.err, sizeof(int)));
        CU(cudaMemsetAsync(R.err, 0, sizeof(int), R.stream));
        if (any_forcing) {
            OK(adsb_load_tensor(R.ctx, 1, 0, 2 /* FORCING */));
            R.forcing = true;
        }
    }
    if (W > 1) {
        // lines per tile of the fused sweep: as slab.py picks it, then what every rank's shared memory allows
        const int sc = (rows + 17) / 18;
        s->nl = 64;
        while (s->nl > 16 && (s->nl / 2) * sc > 288) s->nl /= 2;
        for (int q = 0; q < W; ++q)
            for (int slot : zslots) {
                const adsb_view v = view_of(s, (int) s->pitch, s->r[q].cz);
                while (s->nl >= 16) {
                    const int ok = adsb_dist_sweep_check(s->r[q].ctx, 2, slot, q, &v, s->nl, s->lag);
                    if (ok < 0) return ok;
                    if (ok) break;
                    s->nl /= 2;
                }
                if (s->nl < 16) return fail(ADSB_EINVAL, "slabs_commit: the fused distributed sweep does not fit this slab shape");
            }
        for (auto& R : s->r) {
            CU(cudaSetDevice(R.device));
            const size_t nd = (size_t) W * s->KL * s->lines, nxs = (size_t) W * s->KD * s->lines;
            CU(cudaMalloc((void**) &R.dseg, nd * sizeof(double)));
            CU(cudaMalloc((void**) &R.xst, nxs * sizeof(double)));
            // every word holds the sentinel until the neighbour's store replaces it
            std::vector<unsigned> fill(std::max(nd, nxs) * 2, ADSB_DIST_SENTINEL_WORD);
            CU(cudaMemcpy(R.dseg, fill.data(), nd * sizeof(double), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(R.xst, fill.data(), nxs * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    for (auto& R : s->r) {
        CU(cudaSetDevice(R.device));
        CU(cudaStreamSynchronize(R.stream));
    }
    s->committed = true;
    return ADSB_OK;
}

int adsb_slabs_upload(adsb_slabs* s, const double* host) {
    if (!s || !host || !s->committed) return fail(ADSB_ESTATE, "slabs_upload: not committed");
    const int nx = s->ng[0], ny = s->ng[1];
    for (auto& R : s->r) {
        CU(cudaSetDevice(R.device));
        CU(cudaMemcpy2DAsync(interior(s, R, s->cur), s->pitch * sizeof(double), host + (size_t) R.z0 * ny * nx, nx * sizeof(double),
                             nx * sizeof(double), (size_t) ny * R.cz, cudaMemcpyHostToDevice, R.stream));
    }
    for (auto& R : s->r) {
        CU(cudaSetDevice(R.device));
        CU(cudaStreamSynchronize(R.stream));
    }
    return publish(s, s->cur);
}

int adsb_slabs_download(adsb_slabs* s, double* host) {
    if (!s || !host || !s->committed) return fail(ADSB_ESTATE, "slabs_download: not committed");
    const int nx = s->ng[0], ny = s->ng[1];
    for (auto& R : s->r) {
        CU(cudaSetDevice(R.device));
        CU(cudaMemcpy2DAsync(host + (size_t) R.z0 * ny * nx, nx * sizeof(double), interior(s, R, s->cur), s->pitch * sizeof(double),
                             nx * sizeof(double), (size_t) ny * R.cz, cudaMemcpyDeviceToHost, R.stream));
    }
    for (auto& R : s->r) {
        CU(cudaSetDevice(R.device));
        CU(cudaStreamSynchronize(R.stream));
    }
    return ADSB_OK;
}

int adsb_slabs_step(adsb_slabs* s, int nsteps) {
    if (!s || !s->committed) return fail(ADSB_ESTATE, "slabs_step: not committed");
    const int W = s->world, p = s->p, nz = s->ng[2];
    const long long pl = s->plane;
    for (int it = 0; it < nsteps; ++it)
        for (const adsb_substep& sub : s->prog) {
            const int cur = s->cur, nxt = 1 - cur;
            // every rank launches its whole sub-step; the fused sweeps of neighbouring ranks meet inside the kernels
            for (int q = 0; q < W; ++q) {
                Rank& R = s->r[q];
                CU(cudaSetDevice(R.device));
                // neighbour barrier of the previous sub-step: their halo stores have landed, they are done with my
                // boundary values
                if (q > 0) CU(cudaStreamWaitEvent(R.stream, s->r[q - 1].done, 0));
                if (q + 1 < W) CU(cudaStreamWaitEvent(R.stream, s->r[q + 1].done, 0));
            }
            for (int q = 0; q < W; ++q) {
                Rank& R = s->r[q];
                CU(cudaSetDevice(R.device));
                const int lo = std::max(0, R.z0 - p), hi = std::min(nz, R.z0 + R.cz + p);
                const double* in = R.H[cur] + (long long) (lo - R.z0 + p) * pl;
                double* out = interior(s, R, nxt);
                const adsb_view vin = view_of(s, s->ng[0], hi - lo), v = view_of(s, s->ng[0], R.cz);
                const int in_lo[3] = {0, 0, lo}, out_lo[3] = {0, 0, R.z0};
                const double* forcing = (R.forcing && sub.form.gamma != 0.0) ? adsb_device_ptr(R.ctx, 2) : nullptr;
                OK(adsb_rhs_view(R.ctx, &sub.form, in, &vin, in_lo, forcing, out, &v, out_lo));
                OK(adsb_sweep_view(R.ctx, 0, sub.slots[0], out, &v, nullptr, out, &v, nullptr));
                OK(adsb_sweep_view(R.ctx, 1, sub.slots[1], out, &v, nullptr, out, &v, nullptr));
                if (W == 1) {
                    OK(adsb_sweep_view(R.ctx, 2, sub.slots[2], out, &v, nullptr, out, &v, nullptr));
                } else {
                    adsb_dist_args a{};
                    a.rank = q;
                    a.nranks = W;
                    a.nl = s->nl;
                    a.lag = s->lag;
                    a.dseg_local = R.dseg;
                    a.x_local = R.xst;
                    a.dseg_next = q + 1 < W ? s->r[q + 1].dseg : nullptr;
                    a.x_prev = q > 0 ? s->r[q - 1].xst : nullptr;
                    a.error_flag = R.err;
                    if (p > 0) {  // my first p planes are the upper halo of rank q-1, my last p planes the lower halo of q+1
                        if (q > 0) a.halo_prev = s->r[q - 1].H[nxt] + (long long) (p + s->r[q - 1].cz) * pl;
                        if (q + 1 < W) a.halo_next = s->r[q + 1].H[nxt];
                        a.halo_planes = p;
                    }
                    const adsb_view vl = view_of(s, (int) s->pitch, R.cz);
                    OK(adsb_dist_sweep_view(R.ctx, 2, sub.slots[2], out, &vl, &a));
                }
                s->launches += 4;
            }
            for (int q = 0; q < W; ++q) {
                Rank& R = s->r[q];
                CU(cudaSetDevice(R.device));
                CU(cudaEventRecord(R.done, R.stream));
            }
            s->cur = nxt;
        }
    return ADSB_OK;
}

int adsb_slabs_synchronize(adsb_slabs* s) {
    if (!s || !s->committed) return fail(ADSB_ESTATE, "slabs_synchronize: not committed");
    int bad = 0;
    for (auto& R : s->r) {
        CU(cudaSetDevice(R.device));
        CU(cudaStreamSynchronize(R.stream));
        int e = 0;
        CU(cudaMemcpy(&e, R.err, sizeof(int), cudaMemcpyDeviceToHost));
        bad |= e;
    }
    if (bad) return fail(ADSB_ESTATE, "slabs: a wait of the fused distributed sweep timed out");
    return ADSB_OK;
}

int adsb_slabs_info(adsb_slabs* s, int* bounds, int* info4) {
    if (!s || !s->committed) return fail(ADSB_ESTATE, "slabs_info: not committed");
    if (bounds) std::copy(s->bounds.begin(), s->bounds.end(), bounds);
    if (info4) {
        info4[0] = s->world;
        info4[1] = s->virtual_ranks ? 1 : 0;
        info4[2] = s->nl;
        info4[3] = (int) s->launches;
    }
    return ADSB_OK;
}

}  // extern "C"
