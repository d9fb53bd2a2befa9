// kernels_rhs.cu -- K1 (collapsed form): right-hand side of the ADS step as pre-integrated sum
// factorisation, FP64, sm_100a.
//
// Replaces the compute_rhs() bodies of the examples on the path
// (examples/heat/heat_3d.hpp:49-67, heat_2d.hpp:80-106, implicit/implicit.hpp:132-182,
// scalability/test3d.hpp:66-95):
//     rhs_a = sum_e sum_q [ alpha u v_a - sum_k beta_k d_k u d_k v_a ] w_q J_e  (+ gamma F_a)
// with u = sum_b c_b B_b.  The quadrature is a tensor product and the coefficients are constants,
// so the sum over (e, q) factorises EXACTLY (not up to quadrature error) into the reference's own
// 1-D quadrature matrices M_d = sum w J B B, S_d = sum w J B' B' (src/ads/form_matrix.cpp:8-42):
//     rhs = (Kx (x) My (x) Mz) c - beta_y (Mx (x) Sy (x) Mz) c - beta_z (Mx (x) My (x) Sz) c,
//     Kx = alpha Mx - beta_x Sx
// evaluated axis by axis:  x: P = Kx c, Q = Mx c;  y: G = My P - beta_y Sy Q, H = -beta_z My Q;
// z: rhs = Mz G + Sz H.   7 band products of width 2p+1 per DOF, 16 B of HBM per DOF.
//
// Tiling: a CTA owns a TX x TY column of DOFs and marches along z.  Per z-plane the raw tile
// (with a p-wide halo) is staged in shared memory, the x and y products run out of shared memory
// with their coefficient rows held in registers (a thread keeps its x and its y for the whole
// march), and the z product runs on a register window of the last 2p+1 planes.  Every DOF is
// written exactly once by its owner (no atomics, deterministic).
#include "kernels.cuh"

namespace adsb {

namespace {

constexpr int TX = 32;
constexpr int TYB = 8;  // threads along y

template <int P, int NPT, bool D3>
__global__ void __launch_bounds__(TX* TYB) rhs_collapsed_kernel(const RhsOps ops, const RhsGeom g, int zseg) {
    constexpr int W = 2 * P + 1;
    constexpr int TY = TYB * NPT;
    constexpr int UW = TX + 2 * P;         // raw tile width
    constexpr int UH = TY + 2 * P;         // raw tile height
    __shared__ double U[UH][UW + 1];
    __shared__ double Pf[UH][TX];
    __shared__ double Qf[UH][TX];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * TX + tx;
    const int x0 = g.out_lo[0] + blockIdx.x * TX;
    const int y0 = g.out_lo[1] + blockIdx.y * TY;
    const int gx = x0 + tx;
    const int nx = ops.n[0], ny = ops.n[1];
    const bool xin = gx < g.out_lo[0] + g.out_n[0];

    // coefficient rows of this thread's x (clamped: out-of-box lanes compute garbage nobody stores)
    double kx[W], mx[W];
    {
        const int gxc = min(gx, nx - 1);
#pragma unroll
        for (int m = 0; m < W; ++m) {
            const double a = ops.Mx[gxc * W + m], s = ops.Sx[gxc * W + m];
            mx[m] = a;
            kx[m] = g.alpha * a - g.beta[0] * s;
        }
    }
    double my[NPT][W], sy[NPT][W];
    bool yin[NPT];
#pragma unroll
    for (int r = 0; r < NPT; ++r) {
        const int gy = y0 + ty + r * TYB;
        yin[r] = gy < g.out_lo[1] + g.out_n[1];
        const int gyc = min(gy, ny - 1);
#pragma unroll
        for (int m = 0; m < W; ++m) {
            my[r][m] = ops.My[gyc * W + m];
            sy[r][m] = -g.beta[1] * ops.Sy[gyc * W + m];
        }
    }

    double Gw[NPT][D3 ? W : 1], Hw[NPT][D3 ? W : 1];
#pragma unroll
    for (int r = 0; r < NPT; ++r)
#pragma unroll
        for (int m = 0; m < (D3 ? W : 1); ++m) Gw[r][m] = Hw[r][m] = 0.0;

    const int nz = D3 ? ops.n[2] : 1;
    const int zs = D3 ? g.out_lo[2] + blockIdx.z * zseg : 0;
    const int ze = D3 ? min(zs + zseg, g.out_lo[2] + g.out_n[2]) : 1;
    const int kb = D3 ? zs - P : 0, ke = D3 ? ze + P : 1;

    for (int kin = kb; kin < ke; ++kin) {
        const bool plane_ok = kin >= 0 && kin < nz;  // uniform
        double Gn[NPT], Hn[NPT];
#pragma unroll
        for (int r = 0; r < NPT; ++r) Gn[r] = Hn[r] = 0.0;
        if (plane_ok) {
            const double* src = g.in + (long long) (kin - g.in_lo[2]) * g.si[2];
            for (int idx = tid; idx < UH * UW; idx += TX * TYB) {
                const int r = idx / UW, c = idx - r * UW;
                const int yy = y0 - P + r, xx = x0 - P + c;
                double val = 0.0;
                if (xx >= 0 && xx < nx && yy >= 0 && yy < ny)
                    val = src[(long long) (xx - g.in_lo[0]) * g.si[0] + (long long) (yy - g.in_lo[1]) * g.si[1]];
                U[r][c] = val;
            }
            __syncthreads();
            for (int r = ty; r < UH; r += TYB) {
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int m = 0; m < W; ++m) {
                    const double uv = U[r][tx + m];
                    a = fma(kx[m], uv, a);
                    b = fma(mx[m], uv, b);
                }
                Pf[r][tx] = a;
                Qf[r][tx] = b;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
                const int yl = ty + r * TYB;
                double gg = 0.0, hh = 0.0;
#pragma unroll
                for (int m = 0; m < W; ++m) {
                    const double pv = Pf[yl + m][tx], qv = Qf[yl + m][tx];
                    gg = fma(my[r][m], pv, gg);
                    gg = fma(sy[r][m], qv, gg);
                    hh = fma(my[r][m], qv, hh);
                }
                Gn[r] = gg;
                Hn[r] = -g.beta[2] * hh;
            }
        }
        if (D3) {
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
#pragma unroll
                for (int m = 0; m < W - 1; ++m) {
                    Gw[r][m] = Gw[r][m + 1];
                    Hw[r][m] = Hw[r][m + 1];
                }
                Gw[r][W - 1] = Gn[r];
                Hw[r][W - 1] = Hn[r];
            }
            const int kout = kin - P;
            if (kout >= zs && kout < ze) {
                double mz[W], sz[W];
#pragma unroll
                for (int m = 0; m < W; ++m) {
                    mz[m] = ops.Mz[kout * W + m];
                    sz[m] = ops.Sz[kout * W + m];
                }
#pragma unroll
                for (int r = 0; r < NPT; ++r) {
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < W; ++m) {
                        acc = fma(mz[m], Gw[r][m], acc);
                        acc = fma(sz[m], Hw[r][m], acc);
                    }
                    if (xin && yin[r]) {
                        const int gy = y0 + ty + r * TYB;
                        const long long o = (long long) (gx - g.out_lo[0]) * g.so[0] +
                                            (long long) (gy - g.out_lo[1]) * g.so[1] +
                                            (long long) (kout - g.out_lo[2]) * g.so[2];
                        if (g.forcing) acc = fma(g.gamma, g.forcing[o], acc);
                        g.out[o] = acc;
                    }
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
                if (xin && yin[r]) {
                    const int gy = y0 + ty + r * TYB;
                    const long long o = (long long) (gx - g.out_lo[0]) * g.so[0] +
                                        (long long) (gy - g.out_lo[1]) * g.so[1];
                    double acc = Gn[r];
                    if (g.forcing) acc = fma(g.gamma, g.forcing[o], acc);
                    g.out[o] = acc;
                }
            }
        }
    }
}

template <int P, int NPT>
int launch_p(int ndim, const RhsOps& ops, const RhsGeom& g, cudaStream_t st) {
    constexpr int TY = TYB * NPT;
    dim3 block(TX, TYB, 1);
    if (ndim == 3) {
        // z segments: enough CTAs to fill the machine a few times over, at most 2P/zseg overhead
        const int tiles = ((g.out_n[0] + TX - 1) / TX) * ((g.out_n[1] + TY - 1) / TY);
        int nseg = (148 * 8 + tiles - 1) / tiles;
        int zseg = (g.out_n[2] + nseg - 1) / nseg;
        if (zseg < 16 * P) zseg = 16 * P;
        if (zseg > g.out_n[2]) zseg = g.out_n[2];
        nseg = (g.out_n[2] + zseg - 1) / zseg;
        dim3 grid((g.out_n[0] + TX - 1) / TX, (g.out_n[1] + TY - 1) / TY, nseg);
        rhs_collapsed_kernel<P, NPT, true><<<grid, block, 0, st>>>(ops, g, zseg);
    } else {
        dim3 grid((g.out_n[0] + TX - 1) / TX, (g.out_n[1] + TY - 1) / TY, 1);
        rhs_collapsed_kernel<P, NPT, false><<<grid, block, 0, st>>>(ops, g, 1);
    }
    return (int) cudaGetLastError();
}

__global__ void set_plane_kernel(double* t, long long sa, long long sb, int na, int nb, const double* values) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a < na && b < nb) t[a * sa + b * sb] = values[a + (long long) b * na];
}

}  // namespace

int launch_rhs_collapsed(int ndim, const RhsOps& ops, const RhsGeom& g, cudaStream_t st) {
    const int p = ops.p[0];
    if (ops.p[1] != p || (ndim == 3 && ops.p[2] != p)) return (int) cudaErrorInvalidValue;
    switch (p) {
    case 1: return launch_p<1, 2>(ndim, ops, g, st);
    case 2: return launch_p<2, 2>(ndim, ops, g, st);
    case 3: return launch_p<3, 1>(ndim, ops, g, st);
    case 4: return launch_p<4, 1>(ndim, ops, g, st);
    case 5: return launch_p<5, 1>(ndim, ops, g, st);
    default: return (int) cudaErrorInvalidValue;
    }
}

int launch_set_plane(double* t, const long long s[3], const int n[3], int axis, int idx,
                     const double* values, cudaStream_t st) {
    int a = (axis + 1) % 3, b = (axis + 2) % 3;
    if (a > b) { int tmp = a; a = b; b = tmp; }
    dim3 block(128, 1, 1), grid((n[a] + 127) / 128, n[b], 1);
    set_plane_kernel<<<grid, block, 0, st>>>(t + idx * s[axis], s[a], s[b], n[a], n[b], values);
    return (int) cudaGetLastError();
}

}  // namespace adsb
