// kernels_norm.cu -- norms and errors of the spline solution by element quadrature, FP64.
//
// Replaces basic_simulation_2d/3d::norm / error (include/ads/simulation/basic_simulation_3d.hpp:281-398,
// used by examples/validation/validation.hpp:121-129):
//     val = sum_e sum_q  N( u_h(x_q) - ref(x_q) ) w_q J_e,      N = L2: d.val^2,   H1: d.val^2 + |grad d|^2
// with u_h and grad u_h interpolated from the coefficient tensor through the per-axis tables (the same
// eval_fun as the right-hand side, include/ads/simulation/simulation_3d.hpp:120-128).  One thread per element;
// the partial sums are reduced in a fixed order (block tree, then one block over the block results), so the
// result is deterministic.  A diagnostic, not part of the step: ~ (p+1)^d q^d FMA per element.
#include "kernels.cuh"

namespace adsb {

namespace {

constexpr double PI = 3.14159265358979323846;

struct Val4 {
    double v, dx, dy, dz;
};

// REF 0: none (norm of u_h); 1: the validation solution sin(pi x) sin(pi y) [sin(pi z)] exp(-d pi^2 t)
// (examples/validation/validation.hpp:45-55 and its 3-D twin); 2: tabulated values at the quadrature points
template <int REF>
__device__ __forceinline__ Val4 reference(double x, double y, double z, bool d3, double t, double tab) {
    Val4 r{0, 0, 0, 0};
    if (REF == 1) {
        const double sc = exp(-(d3 ? 3.0 : 2.0) * PI * PI * t);
        const double sx = sin(PI * x), sy = sin(PI * y), sz = d3 ? sin(PI * z) : 1.0;
        const double cx = cos(PI * x), cy = cos(PI * y), cz = d3 ? cos(PI * z) : 0.0;
        r.v = sc * sx * sy * sz;
        r.dx = sc * PI * cx * sy * sz;
        r.dy = sc * PI * sx * cy * sz;
        r.dz = d3 ? sc * PI * sx * sy * cz : 0.0;
    } else if (REF == 2) {
        r.v = tab;
    }
    return r;
}

constexpr int NORM_THREADS = 128;

// partial[2 * block + {0, 1}] = sum over the block's elements of N(u_h - ref), N(ref)
template <int REF>
__global__ void __launch_bounds__(NORM_THREADS)
    norm_kernel(const QuadAxes A, const double* __restrict__ u, long long s1, long long s2, int h1, double t,
                const double* __restrict__ tab, double* __restrict__ partial) {
    const bool d3 = A.ndim == 3;
    const int e0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int e1 = blockIdx.y, e2 = blockIdx.z;
    const int p0 = A.p[0], p1 = A.p[1], p2 = d3 ? A.p[2] : 0;
    const int q0 = A.q[0], q1 = A.q[1], q2 = d3 ? A.q[2] : 1;
    const int st0 = A.st[0], st1 = A.st[1], st2 = d3 ? A.st[2] : 0;
    double err = 0.0, refn = 0.0;
    if (e0 < A.ne[0]) {
        const double J = d3 ? A.J[0][e0] * A.J[1][e1] * A.J[2][e2] : A.J[0][e0] * A.J[1][e1];
        const double* base = u + e0 + s1 * e1 + s2 * e2;  // first DOF of the element (first_dof[e] = e)
        for (int k2 = 0; k2 < q2; ++k2)
            for (int k1 = 0; k1 < q1; ++k1)
                for (int k0 = 0; k0 < q0; ++k0) {
                    const double* b0 = A.bt[0] + (size_t) (e0 * q0 + k0) * st0;
                    const double* b1 = A.bt[1] + (size_t) (e1 * q1 + k1) * st1;
                    const double* b2 = d3 ? A.bt[2] + (size_t) (e2 * q2 + k2) * st2 : nullptr;
                    Val4 uu{0, 0, 0, 0};
                    for (int i2 = 0; i2 <= p2; ++i2) {
                        const double B2 = d3 ? b2[i2] : 1.0, D2 = d3 ? b2[p2 + 1 + i2] : 0.0;
                        for (int i1 = 0; i1 <= p1; ++i1) {
                            const double B1 = b1[i1], D1 = b1[p1 + 1 + i1];
                            double sv = 0.0, sd = 0.0;  // sum over i0 of c * B0 and c * B0'
                            const double* row = base + s1 * i1 + s2 * i2;
                            for (int i0 = 0; i0 <= p0; ++i0) {
                                const double c = row[i0];
                                sv = fma(c, b0[i0], sv);
                                sd = fma(c, b0[p0 + 1 + i0], sd);
                            }
                            uu.v = fma(sv, B1 * B2, uu.v);
                            uu.dx = fma(sd, B1 * B2, uu.dx);
                            uu.dy = fma(sv, D1 * B2, uu.dy);
                            uu.dz = fma(sv, B1 * D2, uu.dz);
                        }
                    }
                    const double w = d3 ? A.w[0][k0] * A.w[1][k1] * A.w[2][k2] : A.w[0][k0] * A.w[1][k1];
                    const double x = A.xq[0][e0 * q0 + k0], y = A.xq[1][e1 * q1 + k1];
                    const double z = d3 ? A.xq[2][e2 * q2 + k2] : 0.0;
                    double tv = 0.0;
                    if (REF == 2) {
                        const long long n0 = (long long) A.ne[0] * q0, n1 = (long long) A.ne[1] * q1;
                        tv = tab[(e0 * q0 + k0) + n0 * ((e1 * q1 + k1) + n1 * (long long) (e2 * q2 + k2))];
                    }
                    const Val4 r = reference<REF>(x, y, z, d3, t, tv);
                    const double dv = uu.v - r.v, ddx = uu.dx - r.dx, ddy = uu.dy - r.dy, ddz = uu.dz - r.dz;
                    double ne = dv * dv, nr = r.v * r.v;
                    if (h1) {
                        ne += ddx * ddx + ddy * ddy + ddz * ddz;
                        nr += r.dx * r.dx + r.dy * r.dy + r.dz * r.dz;
                    }
                    err += ne * w * J;
                    refn += nr * w * J;
                }
    }
    __shared__ double se[NORM_THREADS], sr[NORM_THREADS];
    se[threadIdx.x] = err;
    sr[threadIdx.x] = refn;
    __syncthreads();
    for (int h = NORM_THREADS / 2; h > 0; h >>= 1) {
        if ((int) threadIdx.x < h) {
            se[threadIdx.x] += se[threadIdx.x + h];
            sr[threadIdx.x] += sr[threadIdx.x + h];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long blk = blockIdx.x + (long long) gridDim.x * (blockIdx.y + (long long) gridDim.y * blockIdx.z);
        partial[2 * blk] = se[0];
        partial[2 * blk + 1] = sr[0];
    }
}

// out[0] = sum partial[2k], out[1] = sum partial[2k+1]: one block, fixed order
__global__ void __launch_bounds__(256) norm_finish_kernel(const double* __restrict__ partial, long long nblocks, double* out) {
    __shared__ double se[256], sr[256];
    double a = 0.0, b = 0.0;
    for (long long k = threadIdx.x; k < nblocks; k += 256) {
        a += partial[2 * k];
        b += partial[2 * k + 1];
    }
    se[threadIdx.x] = a;
    sr[threadIdx.x] = b;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
        if ((int) threadIdx.x < h) {
            se[threadIdx.x] += se[threadIdx.x + h];
            sr[threadIdx.x] += sr[threadIdx.x + h];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = se[0];
        out[1] = sr[0];
    }
}

}  // namespace

long long norm_partial_doubles(const QuadAxes& A) {
    const long long bx = (A.ne[0] + NORM_THREADS - 1) / NORM_THREADS;
    return 2 * bx * A.ne[1] * (A.ndim == 3 ? A.ne[2] : 1);
}

// out (device, 2 doubles): sum of N(u_h - ref) w J and of N(ref) w J over all elements
int launch_norm(const QuadAxes& A, const double* u, long long s1, long long s2, int h1, int ref, double t,
                const double* tab, double* partial, double* out, cudaStream_t st) {
    const int ne2 = A.ndim == 3 ? A.ne[2] : 1;
    if (A.ne[1] > 65535 || ne2 > 65535) return (int) cudaErrorInvalidValue;
    dim3 block(NORM_THREADS), grid((A.ne[0] + NORM_THREADS - 1) / NORM_THREADS, A.ne[1], ne2);
    switch (ref) {
    case 0: norm_kernel<0><<<grid, block, 0, st>>>(A, u, s1, s2, h1, t, tab, partial); break;
    case 1: norm_kernel<1><<<grid, block, 0, st>>>(A, u, s1, s2, h1, t, tab, partial); break;
    case 2: norm_kernel<2><<<grid, block, 0, st>>>(A, u, s1, s2, h1, t, tab, partial); break;
    default: return (int) cudaErrorInvalidValue;
    }
    norm_finish_kernel<<<1, 256, 0, st>>>(partial, (long long) grid.x * grid.y * grid.z, out);
    return (int) cudaGetLastError();
}

}  // namespace adsb
