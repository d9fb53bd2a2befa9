// kernels_quadrhs.cu -- K1 (general pointwise forms): launcher of the brick quadrature kernel (quadbrick.cuh;
// instantiated per form in quadbrick_linear.cu / quadbrick_plain.cu / quadbrick_flow.cu) and the init kernel that
// starts the output as gamma * F or zero.  method ADSB_RHS_QUADRATURE of adsb_compute_rhs, adsb_compute_rhs_pointwise.
#include "quadbrick.cuh"

namespace adsb {

// ---- brick kernel (quadbrick.cuh): the shipped path of ADSB_RHS_QUADRATURE and adsb_compute_rhs_pointwise
namespace qb {
__global__ void init_box_kernel(double* y, const double* x, double a, long long n0, long long s1, long long s2) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n0) return;
    const long long o = i + blockIdx.y * s1 + blockIdx.z * s2;
    y[o] = x ? a * x[o] : 0.0;
}
}  // namespace qb

int launch_rhs_brick(int ndim, const QuadAxes& A, const RhsGeom& g, const PointFormArgs& f, const int elo[3],
                     const int en[3], const double* coef, cudaStream_t st, int* nlaunch) {
    const int p = A.p[0];
    for (int d = 0; d < ndim; ++d)
        if (A.p[d] != p || A.q[d] != p + 1 || A.st[d] != 2 * (p + 1)) return (int) cudaErrorInvalidValue;
    {
        const int n0 = g.out_n[0], n1 = g.out_n[1], n2 = ndim == 3 ? g.out_n[2] : 1;
        if (g.so[0] != 1) return (int) cudaErrorInvalidValue;
        dim3 block(128, 1, 1), grid((n0 + 127) / 128, n1, n2);
        qb::init_box_kernel<<<grid, block, 0, st>>>(g.out, g.gamma != 0.0 ? g.forcing : nullptr, g.gamma, n0, g.so[1],
                                                    ndim == 3 ? g.so[2] : 0);
        if (nlaunch) ++*nlaunch;
    }
    if (f.kind == 1) {
        if (ndim != 3 || !coef) return (int) cudaErrorInvalidValue;
        qb::FormFlow form{f.par[0], f.par[1]};
        return qb::launch_brick_form<qb::FormFlow>(ndim, A, g, form, elo, en, coef, g.max_sms, st, nlaunch);
    }
    if (f.kind != 0) return (int) cudaErrorInvalidValue;
    if (f.source && f.plain) {
        qb::FormLinear<true> form{f.alpha, {f.beta[0], f.beta[1], f.beta[2]}, {f.adv[0], f.adv[1], f.adv[2]}, f.gamma, f.source};
        return qb::launch_brick_form<qb::FormLinear<true>>(ndim, A, g, form, elo, en, nullptr, g.max_sms, st, nlaunch);
    }
    qb::FormLinear<false> form{f.alpha, {f.beta[0], f.beta[1], f.beta[2]}, {f.adv[0], f.adv[1], f.adv[2]}, f.gamma, f.source};
    return qb::launch_brick_form<qb::FormLinear<false>>(ndim, A, g, form, elo, en, nullptr, g.max_sms, st, nlaunch);
}

}  // namespace adsb
