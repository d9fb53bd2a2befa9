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

// One CTA: TX x (TYB*NPT) DOFs, marching along z.  NPT vertically adjacent outputs per thread.
template <int P, int NPT, bool D3>
__global__ void __launch_bounds__(TX* TYB, 2) rhs_collapsed_kernel(const RhsOps ops, const RhsGeom g, int zseg) {
    constexpr int W = 2 * P + 1;
    constexpr int TY = TYB * NPT;
    constexpr int UW = TX + 2 * P;         // raw tile width
    constexpr int UH = TY + 2 * P;         // raw tile height
    constexpr int NTH = TX * TYB;
    constexpr int NPF = (UH * UW + NTH - 1) / NTH;  // raw-tile elements per thread
    constexpr int NXR = (UH + TYB - 1) / TYB;       // x-pass rows per thread
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

    // where this thread's raw-tile elements live (fixed for the whole march); -1: outside the domain
    int pf_off[NPF];
    int pf_sm[NPF];
#pragma unroll
    for (int k = 0; k < NPF; ++k) {
        const int idx = tid + k * NTH;
        const int r = idx / UW, c = idx - r * UW;
        const int yy = y0 - P + r, xx = x0 - P + c;
        const bool ok = idx < UH * UW && xx >= 0 && xx < nx && yy >= 0 && yy < ny;
        pf_off[k] = ok ? (int) ((xx - g.in_lo[0]) * g.si[0] + (yy - g.in_lo[1]) * g.si[1]) : -1;
        pf_sm[k] = idx < UH * UW ? r * (UW + 1) + c : -1;
    }

    // coefficient rows of this thread's x and y's (clamped: out-of-box lanes compute garbage nobody stores)
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
        const int gy = y0 + ty * NPT + r;
        yin[r] = gy < g.out_lo[1] + g.out_n[1];
        const int gyc = min(gy, ny - 1);
#pragma unroll
        for (int m = 0; m < W; ++m) {
            my[r][m] = ops.My[gyc * W + m];
            sy[r][m] = -g.beta[1] * ops.Sy[gyc * W + m];
        }
    }

    // acc[r][d]: partial sums of the outputs at planes kin-2P+d .. (scatter form of the z product)
    double acc[NPT][D3 ? W : 1];
#pragma unroll
    for (int r = 0; r < NPT; ++r)
#pragma unroll
        for (int m = 0; m < (D3 ? W : 1); ++m) acc[r][m] = 0.0;

    const int nz = D3 ? ops.n[2] : 1;
    const int zs = D3 ? g.out_lo[2] + blockIdx.z * zseg : 0;
    const int ze = D3 ? min(zs + zseg, g.out_lo[2] + g.out_n[2]) : 1;
    const int kb = D3 ? zs - P : 0, ke = D3 ? ze + P : 1;
    double* Uflat = &U[0][0];

    double pf[NPF];
    auto prefetch = [&](int k) {
        const bool ok = k >= 0 && k < nz && k < ke;
        const double* src = g.in + (long long) (k - g.in_lo[2]) * g.si[2];
#pragma unroll
        for (int q = 0; q < NPF; ++q) pf[q] = (ok && pf_off[q] >= 0) ? __ldg(src + pf_off[q]) : 0.0;
    };
    prefetch(kb);

    for (int kin = kb; kin < ke; ++kin) {
        const bool plane_ok = kin >= 0 && kin < nz;  // uniform
        double Gn[NPT], Hn[NPT];
#pragma unroll
        for (int r = 0; r < NPT; ++r) Gn[r] = Hn[r] = 0.0;
        if (plane_ok) {
#pragma unroll
            for (int q = 0; q < NPF; ++q)
                if (pf_sm[q] >= 0) Uflat[pf_sm[q]] = pf[q];
            __syncthreads();
            prefetch(kin + 1);
#pragma unroll
            for (int q = 0; q < NXR; ++q) {
                const int r = ty + q * TYB;
                if (r < UH) {
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
            }
            __syncthreads();
            double pc[NPT + 2 * P], qc[NPT + 2 * P];
#pragma unroll
            for (int m = 0; m < NPT + 2 * P; ++m) {
                pc[m] = Pf[ty * NPT + m][tx];
                qc[m] = Qf[ty * NPT + m][tx];
            }
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
                double gg = 0.0, hh = 0.0;
#pragma unroll
                for (int m = 0; m < W; ++m) {
                    gg = fma(my[r][m], pc[r + m], gg);
                    gg = fma(sy[r][m], qc[r + m], gg);
                    hh = fma(my[r][m], qc[r + m], hh);
                }
                Gn[r] = gg;
                Hn[r] = -g.beta[2] * hh;
            }
        } else {
            prefetch(kin + 1);
        }
        if (D3) {
            if (plane_ok) {
                // column kin of Mz / Sz: entry d multiplies into output plane kin - P + d ... kin + P
                double cm[W], cs[W];
#pragma unroll
                for (int d = 0; d < W; ++d) {
                    cm[d] = ops.MzT[kin * W + d];
                    cs[d] = ops.SzT[kin * W + d];
                }
#pragma unroll
                for (int r = 0; r < NPT; ++r)
#pragma unroll
                    for (int d = 0; d < W; ++d) {
                        acc[r][d] = fma(cm[d], Gn[r], acc[r][d]);
                        acc[r][d] = fma(cs[d], Hn[r], acc[r][d]);
                    }
            }
            const int kout = kin - P;  // complete now: every input plane <= kout + P has been added
            if (kout >= zs && kout < ze) {
#pragma unroll
                for (int r = 0; r < NPT; ++r) {
                    if (xin && yin[r]) {
                        const int gy = y0 + ty * NPT + r;
                        const long long o = (long long) (gx - g.out_lo[0]) * g.so[0] +
                                            (long long) (gy - g.out_lo[1]) * g.so[1] +
                                            (long long) (kout - g.out_lo[2]) * g.so[2];
                        double val = acc[r][0];
                        if (g.forcing) val = fma(g.gamma, g.forcing[o], val);
                        __stcs(g.out + o, val);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
#pragma unroll
                for (int d = 0; d < W - 1; ++d) acc[r][d] = acc[r][d + 1];
                acc[r][W - 1] = 0.0;
            }
        } else {
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
                if (xin && yin[r]) {
                    const int gy = y0 + ty * NPT + r;
                    const long long o = (long long) (gx - g.out_lo[0]) * g.so[0] +
                                        (long long) (gy - g.out_lo[1]) * g.so[1];
                    double val = Gn[r];
                    if (g.forcing) val = fma(g.gamma, g.forcing[o], val);
                    __stcs(g.out + o, val);
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
