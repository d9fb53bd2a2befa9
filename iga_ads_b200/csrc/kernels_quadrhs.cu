// kernels_quadrhs.cu -- K1 (general form): right-hand side by element-wise Gauss quadrature with sum
// factorisation, FP64, sm_100a.  method ADSB_RHS_QUADRATURE of adsb_compute_rhs.
//
// This is the kernel for forms that are evaluated POINTWISE at the quadrature points -- what the
// reference's compute_rhs() lambdas do (examples/scalability/test3d.hpp:66-95 is the model:
// u = eval_fun(u_prev, e, q) once per point, then the loop over test functions) -- as opposed to
// the pre-integrated operator of kernels_rhs.cu, which only covers constant-coefficient forms.
// Per element (one thread each, lanes along x so that every global access of a warp is a
// contiguous row segment):
//   interpolate   u, du/dx, du/dy, du/dz at the (p+1)^d Gauss points from the (p+1)^d local
//                 coefficients, one axis at a time (sum factorisation, simulation_3d.hpp:120-128
//                 computes the same values point by point);
//   pointwise     k0 = (alpha u) wJ, k_d = -beta_d du/dx_d wJ, plus gamma f(x_q) wJ for a built-in
//                 source (added WITHOUT the test function factor, exactly as test3d.hpp:86-88);
//   integrate     r_a = sum_q k0 B_a + k_x dB_a/dx + ... by the transposed factorisation;
//   scatter       r_a is added to the global tensor (update_global_rhs, simulation_3d.hpp:140-145)
//                 with red.global.add.f64; the caller zeroes the tensor first, as the reference does.
// The per-axis tables b[e][q][d][i] are the reference's own (basis_data, src/ads/basis_data.cpp:63-114)
// staged in shared memory per CTA.  ~1.7 kFMA per element at p=2: the bound is the FP64 pipe.
// The sum over elements is not ordered (atomics): results vary in the last bits from run to run.
#include "quadbrick.cuh"

namespace adsb {

namespace {

constexpr double PI = 3.14159265358979323846;

__device__ __forceinline__ double source_value(int src, bool d3, double x, double y, double z) {
    if (src == 1) {  // examples/scalability/test3d.hpp:58-64, test2d.hpp:49-54
        const double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
        if (d3) return exp(-sqrt(dx * dx + dy * dy + dz * dz)) + 1 + cos(PI * x) * cos(PI * y) * cos(PI * z);
        return exp(-sqrt(dx * dx + dy * dy)) + 1 + cos(PI * x) * cos(PI * y);
    }
    return 0.0;
}

constexpr int QEX = 32, QEY = 4;  // elements per CTA along x and y

// PZ = P (3-D) or 0 (2-D: one point, B = 1, dB = 0 along z)
template <int P, bool D3>
__global__ void __launch_bounds__(QEX* QEY, 2)
    quad_rhs_kernel(const QuadAxes A, const RhsGeom g, int source, int e0x, int e0y, int e0z, int enx, int eny) {
    constexpr int N1 = P + 1, Q = P + 1;
    constexpr int PZ = D3 ? P : 0, NZ = PZ + 1, QZ = D3 ? Q : 1;
    constexpr int ST = 2 * N1;  // doubles per quadrature point in the tables: values | derivatives
    __shared__ double sBx[QEX][Q][ST], sBy[QEY][Q][ST], sBz[QZ][ST];
    __shared__ double sWJx[QEX][Q], sWJy[QEY][Q], sWJz[QZ], sXx[QEX][Q], sXy[QEY][Q], sXz[QZ];

    const int lx = threadIdx.x, ly = threadIdx.y;
    const int tid = ly * QEX + lx;
    const int ex = e0x + blockIdx.x * QEX + lx, ey = e0y + blockIdx.y * QEY + ly, ez = D3 ? e0z + blockIdx.z : 0;
    const int exc = min(ex, A.ne[0] - 1), eyc = min(ey, A.ne[1] - 1);

    // stage this CTA's table slices
    for (int i = tid; i < QEX * Q * ST; i += QEX * QEY) {
        const int e = min(e0x + (int) blockIdx.x * QEX + i / (Q * ST), A.ne[0] - 1);
        (&sBx[0][0][0])[i] = A.bt[0][(size_t) e * Q * ST + i % (Q * ST)];
    }
    for (int i = tid; i < QEY * Q * ST; i += QEX * QEY) {
        const int e = min(e0y + (int) blockIdx.y * QEY + i / (Q * ST), A.ne[1] - 1);
        (&sBy[0][0][0])[i] = A.bt[1][(size_t) e * Q * ST + i % (Q * ST)];
    }
    for (int i = tid; i < QZ * ST; i += QEX * QEY)
        (&sBz[0][0])[i] = D3 ? A.bt[2][(size_t) ez * Q * ST + i] : (i == 0 ? 1.0 : 0.0);
    for (int i = tid; i < QEX * Q; i += QEX * QEY) {
        const int e = min(e0x + (int) blockIdx.x * QEX + i / Q, A.ne[0] - 1);
        (&sWJx[0][0])[i] = A.w[0][i % Q] * A.J[0][e];
        (&sXx[0][0])[i] = A.xq[0][e * Q + i % Q];
    }
    for (int i = tid; i < QEY * Q; i += QEX * QEY) {
        const int e = min(e0y + (int) blockIdx.y * QEY + i / Q, A.ne[1] - 1);
        (&sWJy[0][0])[i] = A.w[1][i % Q] * A.J[1][e];
        (&sXy[0][0])[i] = A.xq[1][e * Q + i % Q];
    }
    for (int i = tid; i < QZ; i += QEX * QEY) {
        sWJz[i] = D3 ? A.w[2][i] * A.J[2][ez] : 1.0;
        sXz[i] = D3 ? A.xq[2][ez * Q + i] : 0.0;
    }
    __syncthreads();
    if (ex >= e0x + enx || ey >= e0y + eny) return;
    (void) exc;
    (void) eyc;

    // local coefficients: DOF (ex + ax, ey + ay, ez + az); the caller guarantees `in` covers them
    const double* cin = g.in + (long long) (ex - g.in_lo[0]) * g.si[0] + (long long) (ey - g.in_lo[1]) * g.si[1] +
                        (D3 ? (long long) (ez - g.in_lo[2]) * g.si[2] : 0);
    double out[N1][N1][NZ];
#pragma unroll
    for (int a = 0; a < N1; ++a)
#pragma unroll
        for (int b = 0; b < N1; ++b)
#pragma unroll
            for (int c = 0; c < NZ; ++c) out[a][b][c] = 0.0;
    double plain = 0.0;  // sum_q gamma f(x_q) w J: added to every local DOF

#pragma unroll 1
    for (int qx = 0; qx < Q; ++qx) {
        // x stage: t1v = sum_ax B c, t1d = sum_ax dB c
        double bx[N1], dbx[N1];
#pragma unroll
        for (int i = 0; i < N1; ++i) {
            bx[i] = sBx[lx][qx][i];
            dbx[i] = sBx[lx][qx][N1 + i];
        }
        double t1v[N1][NZ], t1d[N1][NZ];
#pragma unroll
        for (int b = 0; b < N1; ++b)
#pragma unroll
            for (int c = 0; c < NZ; ++c) {
                double v = 0.0, d = 0.0;
#pragma unroll
                for (int a = 0; a < N1; ++a) {
                    const double cv = __ldg(cin + a * g.si[0] + b * g.si[1] + (D3 ? c * g.si[2] : 0));
                    v = fma(bx[a], cv, v);
                    d = fma(dbx[a], cv, d);
                }
                t1v[b][c] = v;
                t1d[b][c] = d;
            }
        double s1a[N1][NZ], s1b[N1][NZ];
#pragma unroll
        for (int b = 0; b < N1; ++b)
#pragma unroll
            for (int c = 0; c < NZ; ++c) s1a[b][c] = s1b[b][c] = 0.0;
        const double wjx = sWJx[lx][qx], px = sXx[lx][qx];
#pragma unroll 1
        for (int qy = 0; qy < Q; ++qy) {
            double by[N1], dby[N1];
#pragma unroll
            for (int b = 0; b < N1; ++b) {
                by[b] = sBy[ly][qy][b];
                dby[b] = sBy[ly][qy][N1 + b];
            }
            double t2v[NZ], t2x[NZ], t2y[NZ];
#pragma unroll
            for (int c = 0; c < NZ; ++c) {
                double v = 0.0, x = 0.0, y = 0.0;
#pragma unroll
                for (int b = 0; b < N1; ++b) {
                    v = fma(by[b], t1v[b][c], v);
                    x = fma(by[b], t1d[b][c], x);
                    y = fma(dby[b], t1v[b][c], y);
                }
                t2v[c] = v;
                t2x[c] = x;
                t2y[c] = y;
            }
            double s2v[NZ], s2x[NZ], s2y[NZ];
#pragma unroll
            for (int c = 0; c < NZ; ++c) s2v[c] = s2x[c] = s2y[c] = 0.0;
            const double wjxy = wjx * sWJy[ly][qy], py = sXy[ly][qy];
#pragma unroll
            for (int qz = 0; qz < QZ; ++qz) {
                double u = 0.0, ux = 0.0, uy = 0.0, uz = 0.0;
#pragma unroll
                for (int c = 0; c < NZ; ++c) {
                    const double bz = sBz[qz][c], dbz = D3 ? sBz[qz][N1 + c] : 0.0;
                    u = fma(bz, t2v[c], u);
                    ux = fma(bz, t2x[c], ux);
                    uy = fma(bz, t2y[c], uy);
                    if (D3) uz = fma(dbz, t2v[c], uz);
                }
                // ---- the pointwise form
                const double wJ = wjxy * sWJz[qz];
                const double k0 = g.alpha * u * wJ;
                const double k1 = -g.beta[0] * ux * wJ, k2 = -g.beta[1] * uy * wJ, k3 = D3 ? -g.beta[2] * uz * wJ : 0.0;
                if (source) plain = fma(g.gamma * source_value(source, D3, px, py, sXz[qz]), wJ, plain);
                // ---- z stage of the integration
#pragma unroll
                for (int c = 0; c < NZ; ++c) {
                    const double bz = sBz[qz][c], dbz = D3 ? sBz[qz][N1 + c] : 0.0;
                    s2v[c] = fma(bz, k0, s2v[c]);
                    if (D3) s2v[c] = fma(dbz, k3, s2v[c]);
                    s2x[c] = fma(bz, k1, s2x[c]);
                    s2y[c] = fma(bz, k2, s2y[c]);
                }
            }
#pragma unroll
            for (int b = 0; b < N1; ++b)
#pragma unroll
                for (int c = 0; c < NZ; ++c) {
                    s1a[b][c] = fma(by[b], s2v[c], s1a[b][c]);
                    s1a[b][c] = fma(dby[b], s2y[c], s1a[b][c]);
                    s1b[b][c] = fma(by[b], s2x[c], s1b[b][c]);
                }
        }
#pragma unroll
        for (int a = 0; a < N1; ++a)
#pragma unroll
            for (int b = 0; b < N1; ++b)
#pragma unroll
                for (int c = 0; c < NZ; ++c) {
                    out[a][b][c] = fma(bx[a], s1a[b][c], out[a][b][c]);
                    out[a][b][c] = fma(dbx[a], s1b[b][c], out[a][b][c]);
                }
    }

    // scatter into the DOFs of the out box this context owns
#pragma unroll
    for (int c = 0; c < NZ; ++c) {
        const int gz = ez + c;
        if (D3 && (gz < g.out_lo[2] || gz >= g.out_lo[2] + g.out_n[2])) continue;
#pragma unroll
        for (int b = 0; b < N1; ++b) {
            const int gy = ey + b;
            if (gy < g.out_lo[1] || gy >= g.out_lo[1] + g.out_n[1]) continue;
            double* row = g.out + (long long) (gy - g.out_lo[1]) * g.so[1] + (D3 ? (long long) (gz - g.out_lo[2]) * g.so[2] : 0);
#pragma unroll
            for (int a = 0; a < N1; ++a) {
                const int gx = ex + a;
                if (gx >= g.out_lo[0] && gx < g.out_lo[0] + g.out_n[0])
                    atomicAdd(row + (long long) (gx - g.out_lo[0]) * g.so[0], out[a][b][c] + plain);
            }
        }
    }
}

__global__ void axpy_kernel(double* y, const double* x, double a, long long n0, long long s1, long long s2, int n1, int n2) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n0) return;
    const long long o = i + blockIdx.y * s1 + blockIdx.z * s2;
    y[o] = fma(a, x[o], y[o]);
    (void) n1;
    (void) n2;
}

__global__ void zero_box_kernel(double* y, long long n0, long long s1, long long s2) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i < n0) y[i + blockIdx.y * s1 + blockIdx.z * s2] = 0.0;
}

template <int P>
int launch_q(int ndim, const QuadAxes& A, const RhsGeom& g, int source, const int elo[3], const int en[3], cudaStream_t st) {
    dim3 block(QEX, QEY, 1);
    dim3 grid((en[0] + QEX - 1) / QEX, (en[1] + QEY - 1) / QEY, ndim == 3 ? en[2] : 1);
    if (ndim == 3)
        quad_rhs_kernel<P, true><<<grid, block, 0, st>>>(A, g, source, elo[0], elo[1], elo[2], en[0], en[1]);
    else
        quad_rhs_kernel<P, false><<<grid, block, 0, st>>>(A, g, source, elo[0], elo[1], 0, en[0], en[1]);
    return (int) cudaGetLastError();
}

}  // namespace

// `out` must be zero on entry (the reference's zero(rhs)).  Elements [elo, elo+en) are integrated;
// only DOFs inside the out box receive contributions.
int launch_rhs_quadrature(int ndim, const QuadAxes& A, const RhsGeom& g, int source, const int elo[3], const int en[3],
                          cudaStream_t st) {
    const int p = A.p[0];
    for (int d = 0; d < ndim; ++d)
        if (A.p[d] != p || A.q[d] != p + 1 || A.st[d] != 2 * (p + 1)) return (int) cudaErrorInvalidValue;
    switch (p) {
    case 1: return launch_q<1>(ndim, A, g, source, elo, en, st);
    case 2: return launch_q<2>(ndim, A, g, source, elo, en, st);
    case 3: return launch_q<3>(ndim, A, g, source, elo, en, st);
    case 4: return launch_q<4>(ndim, A, g, source, elo, en, st);
    case 5: return launch_q<5>(ndim, A, g, source, elo, en, st);
    default: return (int) cudaErrorInvalidValue;
    }
}

int launch_zero_box(double* y, const int n[3], const long long s[3], cudaStream_t st) {
    dim3 block(128, 1, 1), grid((n[0] + 127) / 128, n[1], n[2]);
    zero_box_kernel<<<grid, block, 0, st>>>(y, n[0], s[1], s[2]);
    return (int) cudaGetLastError();
}

int launch_axpy_box(double* y, const double* x, double a, const int n[3], const long long s[3], cudaStream_t st) {
    dim3 block(128, 1, 1), grid((n[0] + 127) / 128, n[1], n[2]);
    axpy_kernel<<<grid, block, 0, st>>>(y, x, a, n[0], s[1], s[2], n[1], n[2]);
    return (int) cudaGetLastError();
}

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
