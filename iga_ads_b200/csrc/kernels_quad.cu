// kernels_quad.cu -- quadrature of built-in pointwise sources against the B-spline basis, FP64.
//
//   project_kernel      F_a = sum_{e in supp(a)} sum_q f(x_q) B_a(x_q) w_q J_e
//                       = compute_projection (include/ads/projection.hpp:60-107).  One thread per
//                       DOF gathers its own (element, point) contributions in the reference's
//                       scatter order (elements lexicographic, x slowest; then points; the term is
//                       ((f*B)*w)*J without contraction), so the sum is reproduced bit for bit.
//   element_source_kernel + box_sum_kernel
//                       F_a = sum_{e in supp(a)} sum_q f(x_q) w_q J_e   (no test function): the
//                       load the scalability example adds to every DOF of an element
//                       (examples/scalability/test3d.hpp:86-88, test2d.hpp).
#include "kernels.cuh"

namespace adsb {

namespace {

constexpr double PI = 3.14159265358979323846;

template <int SRC>
__device__ __forceinline__ double source(double x, double y, double z, bool d3) {
    if (SRC == 0) {  // examples/heat/heat_3d.hpp:22-28
        const double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
        const double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        const double r2 = fmin(__dmul_rn(8.0, s), 1.0);
        const double a = r2 - 1, b = r2 + 1;
        return __dmul_rn(__dmul_rn(__dmul_rn(a, a), b), b);
    } else if (SRC == 1) {  // examples/implicit/implicit.hpp:38-43 (and its 3-D twin)
        const double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
        double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (d3) s = __dadd_rn(s, __dmul_rn(dz, dz));
        const double r2 = fmin(__dmul_rn(12.0, s), 1.0);
        const double a = r2 - 1, b = r2 + 1;
        return __dmul_rn(__dmul_rn(__dmul_rn(a, a), b), b);
    } else if (SRC == 2) {
        return 1.0;
    } else {  // examples/scalability/test3d.hpp:58-64, test2d.hpp:49-54
        const double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
        if (d3) {
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            return exp(-r) + 1 + cos(PI * x) * cos(PI * y) * cos(PI * z);
        }
        const double r = sqrt(dx * dx + dy * dy);
        return exp(-r) + 1 + cos(PI * x) * cos(PI * y);
    }
}

template <int SRC>
__global__ void project_kernel(const QuadAxes A, double* out, int lo0, int lo1, int lo2, int n0, int n1, int n2,
                               long long pitch0) {
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i1 = blockIdx.y, i2 = blockIdx.z;
    if (i0 >= n0) return;
    const bool d3 = A.ndim == 3;
    const int a0 = lo0 + i0, a1 = lo1 + i1, a2 = lo2 + i2;
    const int p0 = A.p[0], p1 = A.p[1], p2 = d3 ? A.p[2] : 0;
    const int q0 = A.q[0], q1 = A.q[1], q2 = d3 ? A.q[2] : 1;

    const int st0 = A.st[0], st1 = A.st[1], st2 = d3 ? A.st[2] : 0;
    double acc = 0.0;
    for (int e0 = max(a0 - p0, 0); e0 <= min(a0, A.ne[0] - 1); ++e0)
        for (int e1 = max(a1 - p1, 0); e1 <= min(a1, A.ne[1] - 1); ++e1)
            for (int e2 = d3 ? max(a2 - p2, 0) : 0; e2 <= (d3 ? min(a2, A.ne[2] - 1) : 0); ++e2) {
                const double J = d3 ? __dmul_rn(__dmul_rn(A.J[0][e0], A.J[1][e1]), A.J[2][e2])
                                    : __dmul_rn(A.J[0][e0], A.J[1][e1]);
                for (int k0 = 0; k0 < q0; ++k0)
                    for (int k1 = 0; k1 < q1; ++k1)
                        for (int k2 = 0; k2 < q2; ++k2) {
                            const double w = d3 ? __dmul_rn(__dmul_rn(A.w[0][k0], A.w[1][k1]), A.w[2][k2])
                                                : __dmul_rn(A.w[0][k0], A.w[1][k1]);
                            const double x = A.xq[0][e0 * q0 + k0], y = A.xq[1][e1 * q1 + k1];
                            const double z = d3 ? A.xq[2][e2 * q2 + k2] : 0.0;
                            double B = A.bt[0][(e0 * q0 + k0) * st0 + (a0 - e0)];
                            B = __dmul_rn(B, A.bt[1][(e1 * q1 + k1) * st1 + (a1 - e1)]);
                            if (d3) B = __dmul_rn(B, A.bt[2][(e2 * q2 + k2) * st2 + (a2 - e2)]);
                            const double f = source<SRC>(x, y, z, d3);
                            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(f, B), w), J));
                        }
            }
    out[i0 + pitch0 * (i1 + (long long) n1 * i2)] = acc;
}

// The same projection with f tabulated by the caller at the quadrature points of the z-element slab
// [ez_lo, ez_lo + ez_cnt) (2-D: the whole domain): tab[(e0*q0+k0) + nq0*((e1*q1+k1) + nq1*((e2-ez_lo)*q2+k2))].
// accumulate: add to `out` (the caller walks the slabs in ascending order); a single slab covering all z
// elements reproduces the reference's summation order exactly.
__global__ void project_tab_kernel(const QuadAxes A, const double* __restrict__ tab, int ez_lo, int ez_cnt, int accumulate,
                                   double* out, int n0, int n1, int n2, long long pitch0) {
    const int a0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int a1 = blockIdx.y, a2 = blockIdx.z;
    if (a0 >= n0) return;
    const bool d3 = A.ndim == 3;
    const int p0 = A.p[0], p1 = A.p[1], p2 = d3 ? A.p[2] : 0;
    const int q0 = A.q[0], q1 = A.q[1], q2 = d3 ? A.q[2] : 1;
    const int st0 = A.st[0], st1 = A.st[1], st2 = d3 ? A.st[2] : 0;
    const long long nq0 = (long long) A.ne[0] * q0, nq1 = (long long) A.ne[1] * q1;
    const int e2a = d3 ? max(max(a2 - p2, 0), ez_lo) : 0;
    const int e2b = d3 ? min(min(a2, A.ne[2] - 1), ez_lo + ez_cnt - 1) : 0;
    double acc = 0.0;
    for (int e0 = max(a0 - p0, 0); e0 <= min(a0, A.ne[0] - 1); ++e0)
        for (int e1 = max(a1 - p1, 0); e1 <= min(a1, A.ne[1] - 1); ++e1)
            for (int e2 = e2a; e2 <= e2b; ++e2) {
                const double J = d3 ? __dmul_rn(__dmul_rn(A.J[0][e0], A.J[1][e1]), A.J[2][e2])
                                    : __dmul_rn(A.J[0][e0], A.J[1][e1]);
                for (int k0 = 0; k0 < q0; ++k0)
                    for (int k1 = 0; k1 < q1; ++k1)
                        for (int k2 = 0; k2 < q2; ++k2) {
                            const double w = d3 ? __dmul_rn(__dmul_rn(A.w[0][k0], A.w[1][k1]), A.w[2][k2])
                                                : __dmul_rn(A.w[0][k0], A.w[1][k1]);
                            double B = A.bt[0][(e0 * q0 + k0) * st0 + (a0 - e0)];
                            B = __dmul_rn(B, A.bt[1][(e1 * q1 + k1) * st1 + (a1 - e1)]);
                            if (d3) B = __dmul_rn(B, A.bt[2][(e2 * q2 + k2) * st2 + (a2 - e2)]);
                            const double f = tab[(e0 * q0 + k0) + nq0 * ((e1 * q1 + k1) + nq1 * (long long) ((e2 - ez_lo) * q2 + k2))];
                            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(f, B), w), J));
                        }
            }
    double* o = out + a0 + pitch0 * (a1 + (long long) n1 * a2);
    *o = accumulate ? __dadd_rn(*o, acc) : acc;
}

// G[e] = sum_q f(x_q) w J over the element box [elo, elo+en)
template <int SRC>
__global__ void element_source_kernel(const QuadAxes A, double* G, int elo0, int elo1, int elo2, int en0, int en1,
                                      int en2) {
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i1 = blockIdx.y, i2 = blockIdx.z;
    if (i0 >= en0) return;
    const bool d3 = A.ndim == 3;
    const int e0 = elo0 + i0, e1 = elo1 + i1, e2 = elo2 + i2;
    const int q0 = A.q[0], q1 = A.q[1], q2 = d3 ? A.q[2] : 1;
    const double J = d3 ? A.J[0][e0] * A.J[1][e1] * A.J[2][e2] : A.J[0][e0] * A.J[1][e1];
    double acc = 0.0;
    for (int k0 = 0; k0 < q0; ++k0)
        for (int k1 = 0; k1 < q1; ++k1)
            for (int k2 = 0; k2 < q2; ++k2) {
                const double w = d3 ? A.w[0][k0] * A.w[1][k1] * A.w[2][k2] : A.w[0][k0] * A.w[1][k1];
                const double x = A.xq[0][e0 * q0 + k0], y = A.xq[1][e1 * q1 + k1];
                const double z = d3 ? A.xq[2][e2 * q2 + k2] : 0.0;
                acc += source<SRC>(x, y, z, d3) * w * J;
            }
    G[i0 + (long long) en0 * (i1 + (long long) en1 * i2)] = acc;
}

__global__ void box_sum_kernel(const QuadAxes A, const double* G, double* out, int elo0, int elo1, int elo2, int en0,
                               int en1, int lo0, int lo1, int lo2, int n0, int n1, int n2, long long pitch0) {
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i1 = blockIdx.y, i2 = blockIdx.z;
    if (i0 >= n0) return;
    const bool d3 = A.ndim == 3;
    const int a0 = lo0 + i0, a1 = lo1 + i1, a2 = lo2 + i2;
    const int p0 = A.p[0], p1 = A.p[1], p2 = d3 ? A.p[2] : 0;
    double acc = 0.0;
    for (int e0 = max(a0 - p0, 0); e0 <= min(a0, A.ne[0] - 1); ++e0)
        for (int e1 = max(a1 - p1, 0); e1 <= min(a1, A.ne[1] - 1); ++e1)
            for (int e2 = d3 ? max(a2 - p2, 0) : 0; e2 <= (d3 ? min(a2, A.ne[2] - 1) : 0); ++e2)
                acc += G[(e0 - elo0) + (long long) en0 * ((e1 - elo1) + (long long) en1 * (e2 - elo2))];
    out[i0 + pitch0 * (i1 + (long long) n1 * i2)] = acc;
}

}  // namespace

int launch_project(int src, const QuadAxes& A, double* out, const int lo[3], const int n[3], cudaStream_t st,
                   long long pitch0) {
    if (pitch0 <= 0) pitch0 = n[0];
    dim3 block(128, 1, 1), grid((n[0] + 127) / 128, n[1], n[2]);
    switch (src) {
    case 0: project_kernel<0><<<grid, block, 0, st>>>(A, out, lo[0], lo[1], lo[2], n[0], n[1], n[2], pitch0); break;
    case 1: project_kernel<1><<<grid, block, 0, st>>>(A, out, lo[0], lo[1], lo[2], n[0], n[1], n[2], pitch0); break;
    case 2: project_kernel<2><<<grid, block, 0, st>>>(A, out, lo[0], lo[1], lo[2], n[0], n[1], n[2], pitch0); break;
    case 3: project_kernel<3><<<grid, block, 0, st>>>(A, out, lo[0], lo[1], lo[2], n[0], n[1], n[2], pitch0); break;
    default: return (int) cudaErrorInvalidValue;
    }
    return (int) cudaGetLastError();
}

int launch_project_tab(const QuadAxes& A, const double* tab, int ez_lo, int ez_cnt, int accumulate, double* out,
                       const int n[3], cudaStream_t st, long long pitch0) {
    if (pitch0 <= 0) pitch0 = n[0];
    dim3 block(128, 1, 1), grid((n[0] + 127) / 128, n[1], n[2]);
    project_tab_kernel<<<grid, block, 0, st>>>(A, tab, ez_lo, ez_cnt, accumulate, out, n[0], n[1], n[2], pitch0);
    return (int) cudaGetLastError();
}

int launch_element_source(int src, const QuadAxes& A, double* G, const int elo[3], const int en[3], cudaStream_t st) {
    dim3 block(128, 1, 1), grid((en[0] + 127) / 128, en[1], en[2]);
    switch (src) {
    case 0: element_source_kernel<0><<<grid, block, 0, st>>>(A, G, elo[0], elo[1], elo[2], en[0], en[1], en[2]); break;
    case 1: element_source_kernel<1><<<grid, block, 0, st>>>(A, G, elo[0], elo[1], elo[2], en[0], en[1], en[2]); break;
    case 2: element_source_kernel<2><<<grid, block, 0, st>>>(A, G, elo[0], elo[1], elo[2], en[0], en[1], en[2]); break;
    case 3: element_source_kernel<3><<<grid, block, 0, st>>>(A, G, elo[0], elo[1], elo[2], en[0], en[1], en[2]); break;
    default: return (int) cudaErrorInvalidValue;
    }
    return (int) cudaGetLastError();
}

int launch_box_sum(const QuadAxes& A, const double* G, double* out, const int elo[3], const int en[3],
                   const int lo[3], const int n[3], cudaStream_t st, long long pitch0) {
    if (pitch0 <= 0) pitch0 = n[0];
    dim3 block(128, 1, 1), grid((n[0] + 127) / 128, n[1], n[2]);
    box_sum_kernel<<<grid, block, 0, st>>>(A, G, out, elo[0], elo[1], elo[2], en[0], en[1], lo[0], lo[1], lo[2], n[0],
                                           n[1], n[2], pitch0);
    return (int) cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------
// Output sampling: the spline at a tensor-product grid of points (output_manager<2/3>::write / evaluate,
// include/ads/output_manager.hpp:66-73,:101-118 with bspline::eval, include/ads/bspline/eval.hpp:161-192).
// The host finds the span of every point and evaluates the p+1 non-zero basis functions per axis
// (find_span / eval_basis, O(points * p^2)); a thread per output point does the (p+1)^d contraction in the
// reference's loop order (ix outermost) and association ((u * bx) * by) * bz, without contraction to FMA.
namespace {
struct SampleAxes {
    int ndim;
    int npts[3], p[3];
    const int* first[3];    // [npts] first DOF of the point's span (span - p)
    const double* val[3];   // [npts][p+1]
};
__global__ void sample_kernel(const SampleAxes A, const double* __restrict__ u, long long s1, long long s2,
                              double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= A.npts[0]) return;
    const bool d3 = A.ndim == 3;
    const int p0 = A.p[0], p1 = A.p[1], p2 = d3 ? A.p[2] : 0;
    const double* bx = A.val[0] + (size_t) i * (p0 + 1);
    const double* by = A.val[1] + (size_t) j * (p1 + 1);
    const double* bz = d3 ? A.val[2] + (size_t) k * (p2 + 1) : nullptr;
    const double* base = u + A.first[0][i] + s1 * A.first[1][j] + (d3 ? s2 * A.first[2][k] : 0);
    double value = 0.0;
    for (int ix = 0; ix <= p0; ++ix)
        for (int iy = 0; iy <= p1; ++iy)
            for (int iz = 0; iz <= p2; ++iz) {
                double t = __dmul_rn(__dmul_rn(base[ix + s1 * iy + s2 * iz], bx[ix]), by[iy]);
                if (d3) t = __dmul_rn(t, bz[iz]);
                value = __dadd_rn(value, t);
            }
    out[i + (long long) A.npts[0] * (j + (long long) A.npts[1] * k)] = value;
}
}  // namespace

int launch_sample(int ndim, const int* npts, const int* p, const int* const* first, const double* const* val,
                  const double* u, long long s1, long long s2, double* out, cudaStream_t st) {
    SampleAxes A{};
    A.ndim = ndim;
    for (int d = 0; d < 3; ++d) {
        A.npts[d] = d < ndim ? npts[d] : 1;
        A.p[d] = d < ndim ? p[d] : 0;
        A.first[d] = d < ndim ? first[d] : nullptr;
        A.val[d] = d < ndim ? val[d] : nullptr;
    }
    if (A.npts[1] > 65535 || A.npts[2] > 65535) return (int) cudaErrorInvalidValue;
    dim3 block(128), grid((A.npts[0] + 127) / 128, A.npts[1], A.npts[2]);
    sample_kernel<<<grid, block, 0, st>>>(A, u, s1, ndim == 3 ? s2 : 0, out);
    return (int) cudaGetLastError();
}

}  // namespace adsb
