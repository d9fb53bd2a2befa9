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
// Kernel (v3).  A CTA of 8 warps owns a 64 x (8*NPT) tile of DOFs and marches along z.  A lane owns
// two adjacent x (all shared-memory and global accesses are 128-bit), a thread NPT rows of them.
//   * raw planes (tile + halo) arrive through a 3-stage cp.async ring (LDGSTS, zero-filled outside
//     the domain), issued two planes ahead -- no registers are tied up by loads in flight;
//   * x product: 2P+2 inputs from shared memory -> P, Q for two x, written to a double-buffered
//     shared tile (coefficient rows of the lane's two x live in registers for the whole march);
//   * y product out of that tile (coefficient rows broadcast from shared memory);
//   * z product as a scatter onto 2P+1 partial output planes held in registers; the plane loop is
//     unrolled 2P+1 times so the rotation of that window is static register renaming.
// One __syncthreads per plane.  Every DOF is written exactly once by its owner (no atomics).
#include <cstdint>
#include <cstdlib>

#include "kernels.cuh"

namespace adsb {

namespace {

constexpr int TXV = 64;   // tile width in DOFs (32 lanes x 2)
constexpr int NSTAGE = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

template <int BYTES>
__device__ __forceinline__ void cp_async(void* dst, const void* src, bool valid) {
    const int sz = valid ? BYTES : 0;  // src-size 0: the destination is zero-filled, nothing is read
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int P, int NPT, int NWARP>
struct RhsTile {
    static constexpr int W = 2 * P + 1;
    static constexpr int WP = W + 1;            // coefficient rows padded to an even length
    static constexpr int PH = P + (P & 1);      // x halo rounded up to even: 16 B aligned windows
    static constexpr int TY = NWARP * NPT;
    static constexpr int UH = TY + 2 * P;       // rows of the raw / P / Q tiles
    static constexpr int RW = TXV + 2 * PH;     // raw row length in doubles
    static constexpr int NTH = NWARP * 32;
    static constexpr int RAW = UH * RW;         // doubles per raw stage
    static constexpr int PQ = UH * TXV;         // doubles per P (or Q) tile
    static constexpr int YC = TY * 2 * WP;      // y coefficient rows
    static constexpr int SMEM_DOUBLES = NSTAGE * RAW + 4 * PQ + YC;
};

// VEC: x rows are 16 B aligned on both sides (even extents / strides / offsets) -> 128-bit global access
__device__ __forceinline__ double2 lds2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts2(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void cp_async_u32(uint32_t dst, const void* src, int bytes, bool vec) {
    if (vec)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

// YREG: the y coefficient rows of the thread's NPT rows live in registers (else: broadcast from shared)
template <int P, int NPT, int NWARP, int MINB, bool YREG, bool D3, bool VEC>
__global__ void __launch_bounds__(NWARP * 32, MINB) rhs_collapsed_kernel(const RhsOps ops, const RhsGeom g, int zseg) {
    using T = RhsTile<P, NPT, NWARP>;
    constexpr int W = T::W, WP = T::WP, PH = T::PH, TY = T::TY, UH = T::UH, RW = T::RW, NTH = T::NTH;
    constexpr int CB = VEC ? 2 : 1;                       // doubles per cp.async
    constexpr int NCH = (UH * RW / CB + NTH - 1) / NTH;   // chunks per thread per plane
    constexpr int NXF = UH / NWARP;                       // x-pass rows every warp does
    constexpr int NXP = UH - NXF * NWARP;                 // warps [0, NXP) do one more
    constexpr uint32_t RAWB = T::RAW * 8, PQB = T::PQ * 8;
    extern __shared__ __align__(16) double smem[];
    const uint32_t raw_u = smem_u32(smem);                // [NSTAGE][UH][RW]
    const uint32_t pf_u = raw_u + NSTAGE * RAWB;          // [2][UH][TXV]
    const uint32_t qf_u = pf_u + 2 * PQB;                 // [2][UH][TXV]
    double* yc = smem + NSTAGE * T::RAW + 4 * T::PQ;      // [TY][2][WP]: My row | -beta_y Sy row
    const uint32_t yc_u = qf_u + 2 * PQB;

    const int lane = threadIdx.x, w = threadIdx.y;
    const int tid = w * 32 + lane;
    const int x0 = g.out_lo[0] + blockIdx.x * TXV;
    const int y0 = g.out_lo[1] + blockIdx.y * TY;
    const int nx = ops.n[0], ny = ops.n[1];
    const int nz = D3 ? ops.n[2] : 1;

    // readable box = domain intersected with the box `in` covers
    const int rx0 = max(0, g.in_lo[0]), rx1 = min(nx, g.in_lo[0] + g.in_n[0]);
    const int ry0 = max(0, g.in_lo[1]), ry1 = min(ny, g.in_lo[1] + g.in_n[1]);
    const int rz0 = D3 ? max(0, g.in_lo[2]) : 0, rz1 = D3 ? min(nz, g.in_lo[2] + g.in_n[2]) : 1;

    const int zs = D3 ? g.out_lo[2] + blockIdx.z * zseg : 0;
    const int ze = D3 ? min(zs + zseg, g.out_lo[2] + g.out_n[2]) : 1;
    const int kb = D3 ? zs - P : 0;
    const int NP = D3 ? ze - zs + 2 * P : 1;              // input planes kb .. kb + NP - 1
    const int pl0 = max(rz0, kb), pl1 = min(rz1, kb + NP);  // planes that are really read

    // this thread's cp.async chunks: fixed (row, column); the plane base pointer marches
    int ch_off[NCH];       // element offset inside a plane, -1: outside the readable box
    uint32_t ch_dst[NCH];  // shared address in stage 0, ~0: no such chunk
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int idx = tid + k * NTH;
        const int r = idx / (RW / CB), c = (idx - r * (RW / CB)) * CB;
        const int yy = y0 - P + r, xx = x0 - PH + c;
        const bool in_tile = idx < UH * RW / CB;
        const bool ok = in_tile && xx >= rx0 && xx + CB <= rx1 && yy >= ry0 && yy < ry1;
        ch_off[k] = ok ? (int) ((xx - g.in_lo[0]) + (long long) (yy - g.in_lo[1]) * g.si[1]) : -1;
        ch_dst[k] = in_tile ? raw_u + (uint32_t) (r * RW + c) * 8 : 0xffffffffu;
    }
    const long long zstep = D3 ? g.si[2] : 0;
    const double* pl_src = g.in + (D3 ? (long long) (kb - g.in_lo[2]) * g.si[2] : 0);  // plane k_issue
    int k_issue = kb;  // plane the next issue() loads
    uint32_t st_issue = 0;
    auto issue = [&]() {
        const bool pl = k_issue >= pl0 && k_issue < pl1;
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
            if (ch_dst[q] != 0xffffffffu) {
                const bool v = pl && ch_off[q] >= 0;
                cp_async_u32(ch_dst[q] + st_issue, v ? pl_src + ch_off[q] : g.in, v ? 8 * CB : 0, VEC);
            }
        }
        cp_async_commit();
        ++k_issue;
        pl_src += zstep;
        st_issue = (st_issue == (NSTAGE - 1) * RAWB) ? 0 : st_issue + RAWB;
    };

    // coefficient rows of this lane's two x (clamped: out-of-box lanes compute values nobody stores)
    double kx[2][W], mx[2][W];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int gxc = min(x0 + 2 * lane + b, nx - 1);
#pragma unroll
        for (int m = 0; m < W; ++m) {
            const double a = ops.Mx[gxc * W + m], s = ops.Sx[gxc * W + m];
            mx[b][m] = a;
            kx[b][m] = g.alpha * a - g.beta[0] * s;
        }
    }
    double myr[YREG ? NPT : 1][W], syr[YREG ? NPT : 1][W];
    if (YREG) {
#pragma unroll
        for (int r = 0; r < NPT; ++r) {
            const int gyc = min(y0 + w * NPT + r, ny - 1);
#pragma unroll
            for (int m = 0; m < W; ++m) {
                myr[r][m] = ops.My[gyc * W + m];
                syr[r][m] = -g.beta[1] * ops.Sy[gyc * W + m];
            }
        }
    } else {
        for (int i = tid; i < TY * W; i += NTH) {
            const int r = i / W, m = i - r * W;
            const int gyc = min(y0 + r, ny - 1);
            yc[r * 2 * WP + m] = ops.My[gyc * W + m];
            yc[r * 2 * WP + WP + m] = -g.beta[1] * ops.Sy[gyc * W + m];
        }
    }

    // x product of one row of a staged plane: 2P+2 inputs -> P, Q for the lane's two x
    const uint32_t xs_u = raw_u + (uint32_t) (w * RW + 2 * lane) * 8;   // row w of stage 0
    const uint32_t xp_u = pf_u + (uint32_t) (w * TXV + 2 * lane) * 8;   // row w of P tile 0
    auto xrow = [&](uint32_t src, uint32_t dst) {
        double win[2 * PH + 2];
#pragma unroll
        for (int i = 0; i < 2 * PH + 2; i += 2) {
            const double2 t = lds2(src + i * 8);
            win[i] = t.x;
            win[i + 1] = t.y;
        }
        double a[2], b[2];
#pragma unroll
        for (int xb = 0; xb < 2; ++xb) {
            a[xb] = kx[xb][0] * win[xb + PH - P];
            b[xb] = mx[xb][0] * win[xb + PH - P];
#pragma unroll
            for (int m = 1; m < W; ++m) {
                a[xb] = fma(kx[xb][m], win[xb + PH - P + m], a[xb]);
                b[xb] = fma(mx[xb][m], win[xb + PH - P + m], b[xb]);
            }
        }
        sts2(dst, a[0], a[1]);
        sts2(dst + 2 * PQB, b[0], b[1]);  // Q tiles follow the two P tiles
    };
    auto xpass = [&](uint32_t st_off, uint32_t pq_off) {
#pragma unroll
        for (int q = 0; q < NXF; ++q) xrow(xs_u + st_off + q * NWARP * RW * 8, xp_u + pq_off + q * NWARP * TXV * 8);
        if (NXP > 0 && w < NXP) xrow(xs_u + st_off + NXF * NWARP * RW * 8, xp_u + pq_off + NXF * NWARP * TXV * 8);
    };

    // y product for this thread's NPT rows x 2 columns out of a P/Q tile
    const uint32_t yp_u = pf_u + (uint32_t) (w * NPT * TXV + 2 * lane) * 8;
    const uint32_t ycr_u = yc_u + (uint32_t) (w * NPT * 2 * WP) * 8;
    auto ypass = [&](uint32_t pq_off, double (&G)[NPT][2], double (&H)[NPT][2]) {
        double pc[NPT + 2 * P][2], qc[NPT + 2 * P][2];
#pragma unroll
        for (int m = 0; m < NPT + 2 * P; ++m) {
            const double2 a = lds2(yp_u + pq_off + m * TXV * 8);
            const double2 b = lds2(yp_u + pq_off + 2 * PQB + m * TXV * 8);
            pc[m][0] = a.x; pc[m][1] = a.y;
            qc[m][0] = b.x; qc[m][1] = b.y;
        }
#pragma unroll
        for (int r = 0; r < NPT; ++r) {
            double cy[2 * WP];
            if (YREG) {
#pragma unroll
                for (int m = 0; m < W; ++m) {
                    cy[m] = myr[r][m];
                    cy[WP + m] = syr[r][m];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 2 * WP; i += 2) {
                    const double2 t = lds2(ycr_u + (r * 2 * WP + i) * 8);
                    cy[i] = t.x;
                    cy[i + 1] = t.y;
                }
            }
#pragma unroll
            for (int xb = 0; xb < 2; ++xb) {
                double gg = cy[0] * pc[r][xb], hh = D3 ? cy[0] * qc[r][xb] : 0.0;
                gg = fma(cy[WP], qc[r][xb], gg);
#pragma unroll
                for (int m = 1; m < W; ++m) {
                    gg = fma(cy[m], pc[r + m][xb], gg);
                    gg = fma(cy[WP + m], qc[r + m][xb], gg);
                    if (D3) hh = fma(cy[m], qc[r + m][xb], hh);
                }
                G[r][xb] = gg;
                H[r][xb] = D3 ? -g.beta[2] * hh : 0.0;
            }
        }
    };

    // outputs of this thread: NPT row pairs; offsets inside a plane + marching plane pointers
    const int gx = x0 + 2 * lane;
    const int xend = g.out_lo[0] + g.out_n[0];
    int o_row[NPT];  // -1: nothing to store
#pragma unroll
    for (int r = 0; r < NPT; ++r) {
        const int gy = y0 + w * NPT + r;
        const bool ok = gy < g.out_lo[1] + g.out_n[1] && gx < xend;
        o_row[r] = ok ? (int) ((gx - g.out_lo[0]) + (long long) (gy - g.out_lo[1]) * g.so[1]) : -1;
    }
    const bool second = gx + 1 < xend;
    const long long ostep = D3 ? g.so[2] : 0;
    double* o_pl = g.out + (D3 ? (long long) (zs - g.out_lo[2]) * g.so[2] : 0);
    const double* f_pl = g.forcing ? g.forcing + (D3 ? (long long) (zs - g.out_lo[2]) * g.so[2] : 0) : nullptr;
    auto store = [&](int r, double v0, double v1) {
        if (o_row[r] >= 0) {
            double* o = o_pl + o_row[r];
            if (VEC) {  // gx even and xend even: the pair is inside or outside as a whole
                if (f_pl) {
                    const double2 f = __ldg(reinterpret_cast<const double2*>(f_pl + o_row[r]));
                    v0 = fma(g.gamma, f.x, v0);
                    v1 = fma(g.gamma, f.y, v1);
                }
                __stcs(reinterpret_cast<double2*>(o), make_double2(v0, v1));
            } else {
                if (f_pl) v0 = fma(g.gamma, __ldg(f_pl + o_row[r]), v0);
                __stcs(o, v0);
                if (second) {
                    if (f_pl) v1 = fma(g.gamma, __ldg(f_pl + o_row[r] + 1), v1);
                    __stcs(o + 1, v1);
                }
            }
        }
    };
    auto next_out_plane = [&]() {
        o_pl += ostep;
        if (f_pl) f_pl += ostep;
    };

    if (!D3) {
        issue();
        cp_async_wait<0>();
        __syncthreads();
        xpass(0, 0);
        __syncthreads();
        double G[NPT][2], H[NPT][2];
        ypass(0, G, H);
#pragma unroll
        for (int r = 0; r < NPT; ++r) store(r, G[r][0], G[r][1]);
        return;
    }

    // acc[r][xb][s]: partial sums of output planes; slot (d + rot) % W holds plane q - P + d while
    // input plane q is being scattered
    double acc[NPT][2][W];
#pragma unroll
    for (int r = 0; r < NPT; ++r)
#pragma unroll
        for (int xb = 0; xb < 2; ++xb)
#pragma unroll
            for (int s = 0; s < W; ++s) acc[r][xb][s] = 0.0;

    // the z column table carries P zero rows on both sides, so planes outside the domain need no clamp
    const double* zr = ops.MSzT + (long long) (kb + P) * 2 * WP;

    // prologue: planes 0 .. NSTAGE-1 in flight, x product of plane 0
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) issue();
    cp_async_wait<NSTAGE - 1>();
    __syncthreads();
    xpass(0, 0);
    uint32_t st_off = RAWB, pq_off = 0;  // stage of plane q+1, P/Q tile of plane q

    for (int q0 = 0; q0 < NP; q0 += W) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const int q = q0 + j;  // y/z products of plane q (rotation j), x product of plane q+1
            if (q < NP) {
                cp_async_wait<NSTAGE - 2>();
                __syncthreads();
                issue();  // plane q + NSTAGE into the stage plane q was read from (free since the barrier)
                double G[NPT][2], H[NPT][2];
                ypass(pq_off, G, H);
                double cz[2 * WP];
#pragma unroll
                for (int i = 0; i < 2 * WP; i += 2) {
                    const double2 t = __ldg(reinterpret_cast<const double2*>(zr + i));
                    cz[i] = t.x;
                    cz[i + 1] = t.y;
                }
                zr += 2 * WP;
#pragma unroll
                for (int r = 0; r < NPT; ++r)
#pragma unroll
                    for (int xb = 0; xb < 2; ++xb)
#pragma unroll
                        for (int d = 0; d < W; ++d) {
                            const int s = (d + j) % W;
                            acc[r][xb][s] = fma(cz[d], G[r][xb], acc[r][xb][s]);
                            acc[r][xb][s] = fma(cz[WP + d], H[r][xb], acc[r][xb][s]);
                        }
                if (q >= 2 * P) {  // output plane zs + q - 2P is complete
#pragma unroll
                    for (int r = 0; r < NPT; ++r) store(r, acc[r][0][j], acc[r][1][j]);
                    next_out_plane();
                }
#pragma unroll
                for (int r = 0; r < NPT; ++r) {
                    acc[r][0][j] = 0.0;
                    acc[r][1][j] = 0.0;
                }
                pq_off ^= PQB;
                if (q + 1 < NP) xpass(st_off, pq_off);
                st_off = (st_off == (NSTAGE - 1) * RAWB) ? 0 : st_off + RAWB;
            }
        }
    }
    cp_async_wait<0>();
}

// Direct evaluation of the same operator for a narrow box of outputs (the x remainder of the tiling):
// one thread per (x, y) column and z segment, products in the same x -> y -> z order; the thread marches
// along z with the 2P+1 partial output planes in registers, so every input plane is read once per
// segment (+2P halo planes) instead of 2P+1 times.  Tiny share of the work.
constexpr int EDGE_ZSEG = 16;
template <int P, bool D3>
__global__ void rhs_edge_kernel(const RhsOps ops, const RhsGeom g) {
    constexpr int W = 2 * P + 1, WP = W + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int bw = g.out_n[0];
    const int xi = idx % bw, yi = idx / bw;
    if (yi >= g.out_n[1]) return;
    const int nx = ops.n[0], ny = ops.n[1], nz = D3 ? ops.n[2] : 1;
    const int gx = g.out_lo[0] + xi, gy = g.out_lo[1] + yi;
    const int rx0 = max(0, g.in_lo[0]), rx1 = min(nx, g.in_lo[0] + g.in_n[0]);
    const int ry0 = max(0, g.in_lo[1]), ry1 = min(ny, g.in_lo[1] + g.in_n[1]);
    const int rz0 = D3 ? max(0, g.in_lo[2]) : 0, rz1 = D3 ? min(nz, g.in_lo[2] + g.in_n[2]) : 1;
    const int zs = D3 ? g.out_lo[2] + blockIdx.y * EDGE_ZSEG : 0;
    const int ze = D3 ? min(zs + EDGE_ZSEG, g.out_lo[2] + g.out_n[2]) : 1;
    double kx[W], mx[W], my[W], sy[W];
#pragma unroll
    for (int m = 0; m < W; ++m) {
        const double a = ops.Mx[gx * W + m], s = ops.Sx[gx * W + m];
        mx[m] = a;
        kx[m] = g.alpha * a - g.beta[0] * s;
        my[m] = ops.My[gy * W + m];
        sy[m] = -g.beta[1] * ops.Sy[gy * W + m];
    }
    double acc[W];  // acc[d]: partial sum of output plane k - P + d while input plane k is scattered
#pragma unroll
    for (int d = 0; d < W; ++d) acc[d] = 0.0;
    long long o = (long long) xi + (long long) yi * g.so[1] + (D3 ? (long long) (zs - g.out_lo[2]) * g.so[2] : 0);
    for (int k = (D3 ? zs - P : 0); k < (D3 ? ze + P : 1); ++k) {
        double G = 0.0, H = 0.0;
        if (k >= rz0 && k < rz1) {
#pragma unroll
            for (int dy = -P; dy <= P; ++dy) {
                const int yy = gy + dy;
                if (yy < ry0 || yy >= ry1) continue;
                const double* row = g.in + (long long) (yy - g.in_lo[1]) * g.si[1] +
                                    (D3 ? (long long) (k - g.in_lo[2]) * g.si[2] : 0) - g.in_lo[0];
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int m = 0; m < W; ++m) {
                    const int xx = gx - P + m;
                    const double uv = (xx >= rx0 && xx < rx1) ? __ldg(row + xx) : 0.0;
                    a = fma(kx[m], uv, a);
                    b = fma(mx[m], uv, b);
                }
                G = fma(my[dy + P], a, G);
                G = fma(sy[dy + P], b, G);
                H = fma(my[dy + P], b, H);
            }
        }
        if (!D3) {
            acc[0] = G;
        } else {
            const double* zc = ops.MSzT + (long long) (k + P) * 2 * WP;  // column k: A(k-P .. k+P, k)
            const double Hs = -g.beta[2] * H;
#pragma unroll
            for (int d = 0; d < W; ++d) {
                acc[d] = fma(zc[d], G, acc[d]);
                acc[d] = fma(zc[WP + d], Hs, acc[d]);
            }
        }
        if (!D3 || k - P >= zs) {  // output plane k - P is complete
            double v = acc[0];
            if (g.forcing) v = fma(g.gamma, g.forcing[o], v);
            g.out[o] = v;
            o += D3 ? g.so[2] : 0;
        }
#pragma unroll
        for (int d = 0; d + 1 < W; ++d) acc[d] = acc[d + 1];
        acc[W - 1] = 0.0;
    }
}

int sm_count() {
    static int n = [] {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return n;
}

template <int P, int NPT, int NWARP, int MINB, bool YREG>
int launch_p(int ndim, const RhsOps& ops, const RhsGeom& g0, cudaStream_t st, bool try_tma, int* nlaunch,
             const RhsSide* side) {
    using T = RhsTile<P, NPT, NWARP>;
    if (g0.si[0] != 1 || g0.so[0] != 1) return (int) cudaErrorInvalidValue;  // x must be contiguous
    // x remainder of the 64-wide tiling: a nearly empty tile column would cost a full one, so a
    // narrow remainder goes to the direct kernel instead
    RhsGeom g = g0;
    const int rem = g0.out_n[0] % TXV;
    const bool split = rem > 0 && rem <= 16 && g0.out_n[0] > TXV;
    if (split) g.out_n[0] = g0.out_n[0] - rem;
    auto even = [](long long v) { return (v & 1) == 0; };
    bool vec = even(g.si[1]) && even(g.so[1]) && even(g.in_lo[0]) && even(g.out_lo[0]) && even(g.out_n[0]) &&
               even(g.in_n[0]) && even(ops.n[0]) && ((uintptr_t) g.in % 16 == 0) && ((uintptr_t) g.out % 16 == 0) &&
               (!g.forcing || (uintptr_t) g.forcing % 16 == 0);
    if (ndim == 3) vec = vec && even(g.si[2]) && even(g.so[2]);
    const int smem = T::SMEM_DOUBLES * (int) sizeof(double);
    dim3 block(32, NWARP, 1);
    const int tx = (g.out_n[0] + TXV - 1) / TXV, ty = (g.out_n[1] + T::TY - 1) / T::TY;
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute((const void*) kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int) e;
        int nseg = 1, zseg = 1;
        if (ndim == 3) {
            // z segments: every segment re-reads 2P halo planes and pays a pipeline prologue; choose the
            // count that minimises (waves of resident CTAs) x (planes per CTA)
            int per_sm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T::NTH, smem);
            const int sms = (g.max_sms > 0 && g.max_sms < sm_count()) ? g.max_sms : sm_count();
            const long long slots = (long long) sms * (per_sm > 0 ? per_sm : 1);
            const long long tiles = (long long) tx * ty;
            long long best = -1;
            for (int ns = 1; ns <= 64 && ns <= g.out_n[2]; ++ns) {
                const int zs = (g.out_n[2] + ns - 1) / ns;
                if (ns > 1 && zs < 8 * P) break;
                const long long waves = (tiles * ((g.out_n[2] + zs - 1) / zs) + slots - 1) / slots;
                const long long cost = waves * (zs + 2 * P + 4);
                if (best < 0 || cost < best) {
                    best = cost;
                    zseg = zs;
                }
            }
            nseg = (g.out_n[2] + zseg - 1) / zseg;
        }
        dim3 grid(tx, ty, nseg);
        kern<<<grid, block, smem, st>>>(ops, g, zseg);
        return (int) cudaGetLastError();
    };
    int rc = -1;
    const bool fork = split && side && side->side && ndim == 3 && try_tma;
    if (fork && cudaEventRecord(side->fork, st) != cudaSuccess) return (int) cudaGetLastError();
    if (ndim == 3 && try_tma) rc = launch_rhs_tma(ops, g, st);  // -1: not eligible
    if (ndim == 2 && try_tma) rc = launch_rhs2d_march(ops, g, st);  // -1: not eligible / not enabled
    const bool used_tma = rc != -1;
    if (rc != -1)
        ;
    else if (ndim == 3)
        rc = vec ? go(rhs_collapsed_kernel<P, NPT, NWARP, MINB, YREG, true, true>)
                 : go(rhs_collapsed_kernel<P, NPT, NWARP, MINB, YREG, true, false>);
    else
        rc = vec ? go(rhs_collapsed_kernel<P, NPT, NWARP, MINB, YREG, false, true>)
                 : go(rhs_collapsed_kernel<P, NPT, NWARP, MINB, YREG, false, false>);
    if (nlaunch) *nlaunch = split ? 2 : 1;
    if (rc != (int) cudaSuccess || !split) return rc;
    RhsGeom e = g0;
    e.out_lo[0] = g0.out_lo[0] + g.out_n[0];
    e.out_n[0] = rem;
    e.out = g0.out + g.out_n[0];
    if (g0.forcing) e.forcing = g0.forcing + g.out_n[0];
    const long long cols = (long long) rem * e.out_n[1];
    dim3 eb(128, 1, 1), eg((unsigned) ((cols + 127) / 128), ndim == 3 ? (e.out_n[2] + EDGE_ZSEG - 1) / EDGE_ZSEG : 1, 1);
    // the remainder kernel may run next to the main kernel on the caller's side stream (fork / join)
    const bool forked = fork && used_tma;  // (fork implies ndim == 3)
    cudaStream_t es = forked ? side->side : st;
    if (forked) {
        cudaError_t ce = cudaStreamWaitEvent(side->side, side->fork, 0);
        if (ce != cudaSuccess) return (int) ce;
    }
    rc = -1;
    if (used_tma && ndim == 3) rc = launch_rhs_tma(ops, e, es, true);  // same kernel, 8-pair-wide warps; -1: not eligible
    if (rc == -1) {
        if (ndim == 3)
            rhs_edge_kernel<P, true><<<eg, eb, 0, es>>>(ops, e);
        else
            rhs_edge_kernel<P, false><<<eg, eb, 0, es>>>(ops, e);
        rc = (int) cudaGetLastError();
    }
    if (forked) {
        cudaError_t ce = cudaEventRecord(side->join, side->side);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(st, side->join, 0);
        if (ce != cudaSuccess) return (int) ce;
    }
    return rc;
}

// values: dense plane [b][a], stored either densely (vnx huge) or as the first elements of a managed tensor
// (dense index i sits at i % vnx + vpitch * (i / vnx): x rows of vnx doubles, vpitch apart)
__global__ void set_plane_kernel(double* t, long long sa, long long sb, int na, int nb, const double* values, int vnx,
                                 long long vpitch) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a < na && b < nb) {
        const long long i = a + (long long) b * na;
        t[a * sa + b * sb] = values[i % vnx + vpitch * (i / vnx)];
    }
}

}  // namespace

int launch_rhs_collapsed(int ndim, const RhsOps& ops, const RhsGeom& g, cudaStream_t st, int* nlaunch,
                         const RhsSide* side) {
    const int p = ops.p[0];
    if (ops.p[1] != p || (ndim == 3 && ops.p[2] != p)) return (int) cudaErrorInvalidValue;
    // ADSB_RHS_VARIANT: 0 = TMA-fed kernel where eligible (kernels_rhs_tma.cu), else the cp.async kernel;
    // 10 = cp.async kernel only; 11 / 12 = its other launch shapes (kept for measurements)
    static const int variant = [] {
        const char* e = getenv("ADSB_RHS_VARIANT");
        return e ? atoi(e) : 0;
    }();
    const bool tma = variant < 10;
    switch (p) {
    case 1: return launch_p<1, 2, 8, 2, false>(ndim, ops, g, st, tma, nlaunch, side);
    case 2:
        // cp.async kernel, measured at 514^3 on B200: 0.93 ms (8 warps, 1 CTA/SM, y rows in registers), 1.02
        // (variant 11: 2 CTAs/SM, y rows broadcast from shared), 1.11 (variant 12: 12 warps)
        if (variant == 11) return launch_p<2, 2, 8, 2, false>(ndim, ops, g, st, false, nlaunch, side);
        if (variant == 12) return launch_p<2, 2, 12, 1, true>(ndim, ops, g, st, false, nlaunch, side);
        return launch_p<2, 2, 8, 1, true>(ndim, ops, g, st, tma, nlaunch, side);
    case 3: return launch_p<3, 1, 8, 2, false>(ndim, ops, g, st, tma, nlaunch, side);
    case 4: return launch_p<4, 1, 8, 1, false>(ndim, ops, g, st, tma, nlaunch, side);
    case 5: return launch_p<5, 1, 8, 1, false>(ndim, ops, g, st, tma, nlaunch, side);
    default: return (int) cudaErrorInvalidValue;
    }
}

int launch_set_plane(double* t, const long long s[3], const int n[3], int axis, int idx,
                     const double* values, cudaStream_t st, int vnx, long long vpitch) {
    if (vnx <= 0) {  // dense values
        vnx = 0x7fffffff;
        vpitch = 0;
    }
    int a = (axis + 1) % 3, b = (axis + 2) % 3;
    if (a > b) { int tmp = a; a = b; b = tmp; }
    dim3 block(128, 1, 1), grid((n[a] + 127) / 128, n[b], 1);
    set_plane_kernel<<<grid, block, 0, st>>>(t + idx * s[axis], s[a], s[b], n[a], n[b], values, vnx, vpitch);
    return (int) cudaGetLastError();
}

}  // namespace adsb
