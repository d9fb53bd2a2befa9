// kernels_rhs_tma.cu -- K1 (collapsed form), TMA-fed 3-D variant, FP64, sm_100a.
//
// Same operator and the same x -> y -> z product order as kernels_rhs.cu (see there for the algebra;
// replaces examples/heat/heat_3d.hpp:49-67, implicit/implicit.hpp:132-182, scalability/test3d.hpp:66-95).
// kernels_rhs.cu is bound by the shared-memory pipe (1.16 LSU wavefronts per DOF, profiles/r1b_*):
// every plane goes global -> shared (LDGSTS) -> x product -> shared -> y product with a (NPT+2p)/NPT
// re-read.  This kernel moves fewer bytes through that pipe:
//   * raw planes (tile + halo) are fetched by the TMA engine: one cp.async.bulk.tensor per plane and CTA
//     (SASS UTMALDG), issued by one thread, completion on an mbarrier; the halo outside the array is
//     zero-filled by the engine -- no per-thread address arithmetic, predicates or LSU wavefronts;
//   * a warp does the x product of exactly the NPT rows it owns in the y product and keeps those
//     P/Q values in registers; only the 2p neighbouring rows come back out of shared memory
//     (y re-read 2p/NPT instead of (NPT+2p)/NPT), the 2p halo rows of the tile are shared out one per warp;
//   * y and z coefficient rows are broadcast reads of small shared tables (z table staged once per
//     CTA instead of a global load per plane);
//   * small CTAs (NWARP warps, several per SM) so that the shared-memory-heavy x phase of one CTA
//     overlaps the FP64-heavy y/z phase of another.
// z product as in kernels_rhs.cu: scatter onto 2p+1 partial output planes held in registers, plane
// loop unrolled 2p+1 times so the window rotation is register renaming.  One __syncthreads per plane.
#include <cuda.h>

#include <cstdint>
#include <cstdlib>

#include "kernels.cuh"
#include "tma.cuh"

namespace adsb {

namespace {

// XP = x pairs per tile row: a warp is XP lanes wide and 32/XP row groups tall.  XP = 32 is the work
// horse (tile 64 DOFs wide); XP = 8 serves the narrow x remainder of the 64-wide tiling (<= 16 columns).
template <int P, int NPT, int NWARP, int NSTAGE, int XP>
struct Cfg {
    static_assert(XP == 32 || XP == 8, "lane layouts with conflict-free 128-bit shared accesses");
    static constexpr int W = 2 * P + 1;
    static constexpr int PH = P + (P & 1);   // x halo rounded up to even: 16 B aligned windows
    static constexpr int TXV = 2 * XP;       // tile width in DOFs
    static constexpr int RG = 32 / XP;       // row groups per warp
    static constexpr int TY = NWARP * RG * NPT;  // output rows per tile
    static constexpr int UH = TY + 2 * P;    // rows of the raw / P / Q tiles
    static constexpr int RW = TXV + 2 * PH;  // raw row length in doubles
    static constexpr int NTH = NWARP * 32;
    static constexpr int RAW_BYTES = UH * RW * 8;                       // one TMA box
    static constexpr int RAW_STRIDE = (RAW_BYTES + 127) / 128 * 128;    // stage pitch (TMA destination: 128 B aligned)
    static constexpr int PQ_BYTES = UH * TXV * 8;                       // one P (or Q) tile
    static constexpr int YC_BYTES = TY * W * 16;                        // [TY][W] (My, -beta_y Sy)
    static constexpr int FIXED_BYTES = NSTAGE * RAW_STRIDE + 4 * PQ_BYTES + YC_BYTES + NSTAGE * 8 + 8 + 128;
    static constexpr int NHALO = (2 * P * XP + NTH - 1) / NTH;          // halo-row x products a thread may have to do
};

__device__ __forceinline__ double2 lds2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts2(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void tma_load_box3(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

template <int P, int NPT, int NWARP, int NSTAGE, int MINB, bool FORCING, int XP>
__global__ void __launch_bounds__(NWARP * 32, MINB)
    rhs_tma_kernel(const __grid_constant__ CUtensorMap tmap, const RhsOps ops, const RhsGeom g, int zseg, int zfirst) {
    using C = Cfg<P, NPT, NWARP, NSTAGE, XP>;
    constexpr int W = C::W, PH = C::PH, TY = C::TY, UH = C::UH, RW = C::RW, NTH = C::NTH, TXV = C::TXV;
    constexpr uint32_t RAWS = C::RAW_STRIDE, PQB = C::PQ_BYTES;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base_u = (smem_addr(smem_raw) + 127u) & ~127u;
    unsigned char* base_p = smem_raw + (base_u - smem_addr(smem_raw));
    const uint32_t raw_u = base_u;                     // [NSTAGE][UH][RW]
    const uint32_t pq_u = raw_u + NSTAGE * RAWS;       // [2 buffers][P | Q][UH][TXV]
    const uint32_t yc_u = pq_u + 4 * PQB;              // [TY][W] double2
    const uint32_t bar_u = yc_u + C::YC_BYTES;         // NSTAGE mbarriers
    const uint32_t zt_u = bar_u + NSTAGE * 8 + ((NSTAGE & 1) ? 8 : 0);  // [NP][W] double2 (Mz, -beta_z Sz) columns
    double2* yc = reinterpret_cast<double2*>(base_p + (yc_u - base_u));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base_p + (bar_u - base_u));
    double2* zt = reinterpret_cast<double2*>(base_p + (zt_u - base_u));

    const int tid = threadIdx.x;
    const int lx = tid % XP;   // x pair of this thread inside the tile row
    const int rg = tid / XP;   // its row group: rows rg*NPT .. rg*NPT + NPT-1 of the tile
    const int x0 = g.out_lo[0] + blockIdx.x * TXV;
    const int y0 = g.out_lo[1] + blockIdx.y * TY;
    const int nx = ops.n[0], ny = ops.n[1];
    // z segments: the first one has zfirst planes, the others zseg (see launch_cfg)
    const int zs = g.out_lo[2] + (blockIdx.z == 0 ? 0 : zfirst + ((int) blockIdx.z - 1) * zseg);
    const int ze = blockIdx.z == 0 ? g.out_lo[2] + min(zfirst, g.out_n[2]) : min(zs + zseg, g.out_lo[2] + g.out_n[2]);
    const int kb = zs - P;
    const int NP = ze - zs + 2 * P;  // input planes kb .. kb + NP - 1

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // y rows of the tile and z columns of the planes this CTA marches through
    for (int i = tid; i < TY * W; i += NTH) {
        const int r = i / W, m = i - r * W;
        const int gyc = min(y0 + r, ny - 1);
        yc[i] = make_double2(ops.My[gyc * W + m], -g.beta[1] * ops.Sy[gyc * W + m]);
    }
    {
        // the z column table carries P zero rows on both sides, so planes outside the domain need no clamp
        const double* zr = ops.MSzT + (long long) (kb + P) * 2 * (W + 1);
        for (int i = tid; i < NP * W; i += NTH) {
            const int q = i / W, d = i - q * W;
            zt[i] = make_double2(zr[q * 2 * (W + 1) + d], -g.beta[2] * zr[q * 2 * (W + 1) + (W + 1) + d]);
        }
    }
    // coefficient rows of this lane's two x (clamped: out-of-box lanes compute values nobody stores)
    double kx[2][W], mx[2][W];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int gxc = min(x0 + 2 * lx + b, nx - 1);
#pragma unroll
        for (int m = 0; m < W; ++m) {
            const double a = ops.Mx[gxc * W + m], s = ops.Sx[gxc * W + m];
            mx[b][m] = a;
            kx[b][m] = g.alpha * a - g.beta[0] * s;
        }
    }
    __syncthreads();
    pdl_wait();  // everything above read constant tables only; the tensors belong to the previous kernel until here

    // TMA box of plane k: coordinates relative to the box `in` covers; everything outside is zero-filled
    const int c0 = x0 - PH - g.in_lo[0], c1 = y0 - P - g.in_lo[1], c2b = kb - g.in_lo[2];
    auto issue = [&](int q, int stage) {
        mbar_expect_tx_u32(bar_u + stage * 8, C::RAW_BYTES);
        tma_load_box3(raw_u + stage * RAWS, &tmap, c0, c1, c2b + q, bar_u + stage * 8);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s)
            if (s < NP) issue(s, s);
    }

    // x product of one loaded window: 2*PH+2 inputs -> P, Q for the lane's two x
    auto xmul = [&](const double (&win)[2 * PH + 2], double (&a)[2], double (&b)[2]) {
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
    };
    const uint32_t lane_raw = (uint32_t) (2 * lx) * 8;  // column offset of the thread's window in a raw row
    const uint32_t lane_pq = (uint32_t) (2 * lx) * 8;   // ... of its pair in a P/Q row
    const int jown = P + rg * NPT;                      // first own row (tile row index, 0 = y0 - P)

    // outputs of this thread: NPT row pairs; offsets inside a plane + marching plane pointers
    const int gx = x0 + 2 * lx;
    const int xend = g.out_lo[0] + g.out_n[0];
    int o_row[NPT];  // -1: nothing to store
#pragma unroll
    for (int r = 0; r < NPT; ++r) {
        const int gy = y0 + rg * NPT + r;
        const bool ok = gy < g.out_lo[1] + g.out_n[1] && gx < xend;
        o_row[r] = ok ? (int) ((gx - g.out_lo[0]) + (long long) (gy - g.out_lo[1]) * g.so[1]) : -1;
    }
    double* o_pl = g.out + (long long) (zs - g.out_lo[2]) * g.so[2];
    const double* f_pl = FORCING ? g.forcing + (long long) (zs - g.out_lo[2]) * g.so[2] : nullptr;

    // acc[r][xb][s]: partial sums of output planes; slot (d + rot) % W holds plane q - P + d while
    // input plane q is being scattered
    double acc[NPT][2][W];
#pragma unroll
    for (int r = 0; r < NPT; ++r)
#pragma unroll
        for (int xb = 0; xb < 2; ++xb)
#pragma unroll
            for (int s = 0; s < W; ++s) acc[r][xb][s] = 0.0;

    int stage = 0;
    uint32_t parity = 0, pqbuf = 0;
    uint32_t zrow_u = zt_u;
    for (int q0 = 0; q0 < NP; q0 += W) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const int q = q0 + j;  // plane q: x, y products and z scatter with rotation j
            if (q < NP) {
                mbar_wait_u32(bar_u + stage * 8, parity);
                const uint32_t raw_s = raw_u + stage * RAWS + lane_raw;
                const uint32_t pq_s = pq_u + pqbuf * 2 * PQB + lane_pq;
                // x product: own rows stay in registers; rows a neighbouring warp needs go to shared.  The
                // window of the next row is loaded before the current row is multiplied out (the stores
                // in between would otherwise pin every load behind them).
                double pr[NPT][2], qr[NPT][2];
                {
                    // rows of this thread: own rows, then its share of the 2P halo rows of the tile (P above, P
                    // below; halo task t = row t / XP, pair t % XP = lx)
                    constexpr int NXR = NPT + C::NHALO;
                    auto row_of = [&](int i) -> int {    // tile row of the i-th of them
                        if (i < NPT) return jown + i;
                        const int h = min((tid + (i - NPT) * NTH) / XP, 2 * P - 1);  // surplus threads shadow the last halo row
                        return h < P ? h : TY + h;
                    };
                    double win[2][2 * PH + 2];
                    auto load_win = [&](int i, double (&wv)[2 * PH + 2]) {
                        const uint32_t src = raw_s + (uint32_t) (row_of(i) * RW) * 8;
#pragma unroll
                        for (int c = 0; c < 2 * PH + 2; c += 2) {
                            const double2 t = lds2(src + c * 8);
                            wv[c] = t.x;
                            wv[c + 1] = t.y;
                        }
                    };
                    load_win(0, win[0]);
#pragma unroll
                    for (int i = 0; i < NXR; ++i) {
                        if (i + 1 < NXR) load_win(i + 1, win[(i + 1) & 1]);
                        double a[2], b[2];
                        xmul(win[i & 1], a, b);
                        const uint32_t dst = pq_s + (uint32_t) (row_of(i) * TXV) * 8;
                        if (i < NPT) {
                            pr[i][0] = a[0]; pr[i][1] = a[1];
                            qr[i][0] = b[0]; qr[i][1] = b[1];
                            if (i < P || i >= NPT - P) {
                                sts2(dst, a[0], a[1]);
                                sts2(dst + PQB, b[0], b[1]);
                            }
                        } else if (tid + (i - NPT) * NTH < 2 * P * XP) {
                            sts2(dst, a[0], a[1]);
                            sts2(dst + PQB, b[0], b[1]);
                        }
                    }
                }
                __syncthreads();
                // the raw stage is free again: fetch plane q + NSTAGE into it
                if (tid == 0 && q + NSTAGE < NP) issue(q + NSTAGE, stage);

                // y product, input row by input row (each feeds up to 2P+1 output rows: independent chains).
                // Extended column e = 0 .. NPT+2P-1 is tile row rg*NPT + e; e in [P, P+NPT) is an own row
                // (registers), the others are read back from shared.  Own rows go first: their operands
                // are there while the halo loads are still in flight.
                double G[NPT][2], H[NPT][2];
#pragma unroll
                for (int r = 0; r < NPT; ++r) G[r][0] = G[r][1] = H[r][0] = H[r][1] = 0.0;
                double hp[2 * P][2], hq[2 * P][2];
#pragma unroll
                for (int i = 0; i < 2 * P; ++i) {
                    const int e = i < P ? i : NPT + i;
                    const double2 a = lds2(pq_s + (uint32_t) ((rg * NPT + e) * TXV) * 8);
                    const double2 b = lds2(pq_s + PQB + (uint32_t) ((rg * NPT + e) * TXV) * 8);
                    hp[i][0] = a.x; hp[i][1] = a.y;
                    hq[i][0] = b.x; hq[i][1] = b.y;
                }
                auto yrow = [&](int e, const double (&pe)[2], const double (&qe)[2]) {
                    double2 cy[W];
#pragma unroll
                    for (int m = 0; m < W; ++m) {
                        const int r = e - m;
                        if (r >= 0 && r < NPT) cy[m] = lds2(yc_u + (uint32_t) (((rg * NPT + r) * W + m) * 16));
                    }
#pragma unroll
                    for (int m = 0; m < W; ++m) {
                        const int r = e - m;
                        if (r >= 0 && r < NPT) {
#pragma unroll
                            for (int xb = 0; xb < 2; ++xb) {
                                G[r][xb] = fma(cy[m].x, pe[xb], G[r][xb]);
                                H[r][xb] = fma(cy[m].x, qe[xb], H[r][xb]);
                                G[r][xb] = fma(cy[m].y, qe[xb], G[r][xb]);
                            }
                        }
                    }
                };
#pragma unroll
                for (int r = 0; r < NPT; ++r) yrow(P + r, pr[r], qr[r]);
#pragma unroll
                for (int i = 0; i < 2 * P; ++i) yrow(i < P ? i : NPT + i, hp[i], hq[i]);
                // z product: scatter plane q onto the 2P+1 output planes it feeds
#pragma unroll
                for (int d = 0; d < W; ++d) {
                    const double2 cz = lds2(zrow_u + d * 16);
                    const int s = (d + j) % W;
#pragma unroll
                    for (int r = 0; r < NPT; ++r)
#pragma unroll
                        for (int xb = 0; xb < 2; ++xb) {
                            acc[r][xb][s] = fma(cz.x, G[r][xb], acc[r][xb][s]);
                            acc[r][xb][s] = fma(cz.y, H[r][xb], acc[r][xb][s]);
                        }
                }
                zrow_u += W * 16;
                if (q >= 2 * P) {  // output plane zs + q - 2P is complete
#pragma unroll
                    for (int r = 0; r < NPT; ++r) {
                        if (o_row[r] >= 0) {
                            double v0 = acc[r][0][j], v1 = acc[r][1][j];
                            if (FORCING) {
                                const double2 f = __ldg(reinterpret_cast<const double2*>(f_pl + o_row[r]));
                                v0 = fma(g.gamma, f.x, v0);
                                v1 = fma(g.gamma, f.y, v1);
                            }
                            __stcs(reinterpret_cast<double2*>(o_pl + o_row[r]), make_double2(v0, v1));
                        }
                    }
                    o_pl += g.so[2];
                    if (FORCING) f_pl += g.so[2];
                }
#pragma unroll
                for (int r = 0; r < NPT; ++r) {
                    acc[r][0][j] = 0.0;
                    acc[r][1][j] = 0.0;
                }
                pqbuf ^= 1;
                if (++stage == NSTAGE) {
                    stage = 0;
                    parity ^= 1;
                }
            }
        }
    }
}

int sm_count_tma() {
    static int n = [] {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return n;
}

template <int P, int NPT, int NWARP, int NSTAGE, int MINB, int XP = 32>
int launch_cfg(const RhsOps& ops, const RhsGeom& g, cudaStream_t st) {
    using C = Cfg<P, NPT, NWARP, NSTAGE, XP>;
    constexpr int TXV = C::TXV;
    auto kern = g.forcing ? rhs_tma_kernel<P, NPT, NWARP, NSTAGE, MINB, true, XP>
                          : rhs_tma_kernel<P, NPT, NWARP, NSTAGE, MINB, false, XP>;
    CUtensorMap map;
    const unsigned long long dims[3] = {(unsigned long long) g.in_n[0], (unsigned long long) g.in_n[1],
                                        (unsigned long long) g.in_n[2]};
    const unsigned long long strides[2] = {(unsigned long long) g.si[1], (unsigned long long) g.si[2]};
    const unsigned box[3] = {(unsigned) C::RW, (unsigned) C::UH, 1u};
    if (!encode_tensor_map3(&map, g.in, dims, strides, box)) return -1;

    const int tx = (g.out_n[0] + TXV - 1) / TXV, ty = (g.out_n[1] + C::TY - 1) / C::TY;
    const int smem_budget = 227 * 1024 / MINB - 1024;
    const int zcap = (smem_budget - C::FIXED_BYTES) / (C::W * 16) - 2 * P;  // planes whose z columns fit in shared
    if (zcap < 8 * P) return -1;
    // z segments: every segment re-reads 2P halo planes and pays a pipeline prologue; choose the count
    // that minimises (waves of resident CTAs) x (planes per CTA)
    const int sms = (g.max_sms > 0 && g.max_sms < sm_count_tma()) ? g.max_sms : sm_count_tma();
    const long long slots = (long long) sms * MINB;
    const long long tiles = (long long) tx * ty;
    long long best = -1;
    int zseg = 1;
    for (int ns = 1; ns <= 256 && ns <= g.out_n[2]; ++ns) {
        const int zs = (g.out_n[2] + ns - 1) / ns;
        if (zs > zcap) continue;
        if (ns > 1 && zs < 8 * P) break;
        const long long waves = (tiles * ((g.out_n[2] + zs - 1) / zs) + slots - 1) / slots;
        const long long cost = waves * (zs + 2 * P + 4);
        if (best < 0 || cost < best) {
            best = cost;
            zseg = zs;
        }
    }
    if (best < 0) return -1;
    int zfirst = zseg;
    int nseg = (g.out_n[2] + zseg - 1) / zseg;
    // One column cannot fill the device twice over (tiles < slots < 2 tiles): cut every column into a LONG
    // and a SHORT segment.  CTAs are dispatched in block order, so the `tiles` long segments start first
    // and the slots left over work through the short ones, j = ceil(tiles / spare) in a row each, while
    // the long ones run; lengths chosen so both finish together:  long + c = j (short + c).
    if (tiles < slots && slots < 2 * tiles) {
        const long long spare = slots - tiles;
        const long long j = (tiles + spare - 1) / spare;
        const int c = 2 * P + 4;
        const int zshort = (int) ((g.out_n[2] + c - c * j) / (j + 1));
        const int zlong = g.out_n[2] - zshort;
        if (zshort >= 8 * P && zlong <= zcap && (long long) (zlong + c) < best) {
            zfirst = zlong;
            zseg = zshort;
            nseg = 2;
        }
    }
    const int smem = C::FIXED_BYTES + ((zfirst > zseg ? zfirst : zseg) + 2 * P) * C::W * 16;
    cudaError_t e = cudaFuncSetAttribute((const void*) kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int) e;
    dim3 grid(tx, ty, nseg), block(C::NTH, 1, 1);
    if (cudaError_t e = launch_ex(kern, grid, block, smem, st, true, map, ops, g, zseg, zfirst); e != cudaSuccess) return (int) e;
    return (int) cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// 2-D variant (ADSB_RHS2D_MARCH=0 switches back to the cp.async kernel):
//     rhs = (Kx (x) My) c - beta_y (Mx (x) Sy) c,   Kx = alpha Mx - beta_x Sx
// (examples/heat/heat_2d.hpp:80-106, implicit/implicit.hpp:132-182).  The cp.async kernel does one 64 x 8 tile
// per CTA with no pipelining (0.41 ms at 4099^2, p=3, against a 0.04 ms roofline; this kernel: 0.093 ms, and
// 0.057 instead of 0.162 ms at 4096^2, p=2).  Here a CTA owns a strip of
// 64*NWARP x values and MARCHES along y: row blocks of W = 2p+1 rows arrive by TMA (ring of NSTAGE), a thread
// does the x product of its two x out of a 128-bit shared window and scatters the row onto 2p+1 partial
// output rows in registers (rotation = position of the row in its block: compile time).  No data moves
// between threads, so the only barrier is the one that frees a ring slot.
template <int P, int NWARP, int NSTAGE, bool FORCING>
__global__ void __launch_bounds__(NWARP * 32)
    rhs2d_march_kernel(const __grid_constant__ CUtensorMap tmap, const RhsOps ops, const RhsGeom g, int yseg) {
    constexpr int W = 2 * P + 1, PH = P + (P & 1), NTH = NWARP * 32, TXW = 64 * NWARP, RW = TXW + 2 * PH;
    constexpr uint32_t STAGE_BYTES = RW * W * 8, STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base_u = (smem_addr(smem_raw) + 127u) & ~127u;
    unsigned char* base_p = smem_raw + (base_u - smem_addr(smem_raw));
    const uint32_t raw_u = base_u;                          // [NSTAGE][W rows][RW]
    const uint32_t bar_u = raw_u + NSTAGE * STAGE_STRIDE;   // NSTAGE mbarriers
    const uint32_t yt_u = bar_u + NSTAGE * 8 + ((NSTAGE & 1) ? 8 : 0);  // [rows][W] double2 (My, -beta_y Sy) columns
    uint64_t* bars = reinterpret_cast<uint64_t*>(base_p + (bar_u - base_u));
    double2* yt = reinterpret_cast<double2*>(base_p + (yt_u - base_u));

    const int tid = threadIdx.x;
    const int x0 = g.out_lo[0] + blockIdx.x * TXW;
    const int nx = ops.n[0];
    const int ys = g.out_lo[1] + blockIdx.y * yseg;
    const int ye = min(ys + yseg, g.out_lo[1] + g.out_n[1]);
    const int kb = ys - P;
    const int NR = ye - ys + 2 * P;          // input rows kb .. kb + NR - 1
    const int NB = (NR + W - 1) / W;         // row blocks

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    {   // the column table carries P zero rows on both sides; rows past NR (tail of the last block) read as zero
        const double* yr = ops.MSzT + (long long) (kb + P) * 2 * (W + 1);
        for (int i = tid; i < NB * W * W; i += NTH) {
            const int q = i / W, d = i - q * W;
            yt[i] = q < NR ? make_double2(yr[q * 2 * (W + 1) + d], -g.beta[1] * yr[q * 2 * (W + 1) + (W + 1) + d])
                           : make_double2(0.0, 0.0);
        }
    }
    double kx[2][W], mx[2][W];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int gxc = min(x0 + 2 * tid + b, nx - 1);
#pragma unroll
        for (int m = 0; m < W; ++m) {
            const double a = ops.Mx[gxc * W + m], sv = ops.Sx[gxc * W + m];
            mx[b][m] = a;
            kx[b][m] = g.alpha * a - g.beta[0] * sv;
        }
    }
    __syncthreads();

    const int c0 = x0 - PH - g.in_lo[0], c1b = kb - g.in_lo[1];
    auto issue = [&](int blk, int stage) {  // rows kb + blk*W .. +W-1 (outside the array: zero-filled)
        mbar_expect_tx_u32(bar_u + stage * 8, STAGE_BYTES);
        tma_load_box3(raw_u + stage * STAGE_STRIDE, &tmap, c0, c1b + blk * W, 0, bar_u + stage * 8);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s)
            if (s < NB) issue(s, s);
    }

    const int gx = x0 + 2 * tid;
    const bool live = gx < g.out_lo[0] + g.out_n[0];
    double* o_row = g.out + (gx - g.out_lo[0]) + (long long) (ys - g.out_lo[1]) * g.so[1];
    const double* f_row = FORCING ? g.forcing + (gx - g.out_lo[0]) + (long long) (ys - g.out_lo[1]) * g.so[1] : nullptr;

    double acc[2][W];
#pragma unroll
    for (int xb = 0; xb < 2; ++xb)
#pragma unroll
        for (int sl = 0; sl < W; ++sl) acc[xb][sl] = 0.0;

    int stage = 0;
    uint32_t parity = 0, yrow_u = yt_u;
    for (int blk = 0; blk < NB; ++blk) {
        mbar_wait_u32(bar_u + stage * 8, parity);
        const uint32_t rs = raw_u + stage * STAGE_STRIDE + (uint32_t) (2 * tid) * 8;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const int q = blk * W + j;  // input row kb + q, rotation j
            double win[2 * PH + 2];
#pragma unroll
            for (int c = 0; c < 2 * PH + 2; c += 2) {
                const double2 t = lds2(rs + (uint32_t) (j * RW + c) * 8);
                win[c] = t.x;
                win[c + 1] = t.y;
            }
            double pv[2], qv[2];
#pragma unroll
            for (int xb = 0; xb < 2; ++xb) {
                pv[xb] = kx[xb][0] * win[xb + PH - P];
                qv[xb] = mx[xb][0] * win[xb + PH - P];
#pragma unroll
                for (int m = 1; m < W; ++m) {
                    pv[xb] = fma(kx[xb][m], win[xb + PH - P + m], pv[xb]);
                    qv[xb] = fma(mx[xb][m], win[xb + PH - P + m], qv[xb]);
                }
            }
#pragma unroll
            for (int d = 0; d < W; ++d) {
                const double2 cy = lds2(yrow_u + (uint32_t) ((j * W + d) * 16));
                const int sl = (d + j) % W;
#pragma unroll
                for (int xb = 0; xb < 2; ++xb) {
                    acc[xb][sl] = fma(cy.x, pv[xb], acc[xb][sl]);
                    acc[xb][sl] = fma(cy.y, qv[xb], acc[xb][sl]);
                }
            }
            if (q >= 2 * P && q < NR) {  // output row ys + q - 2P is complete
                if (live) {
                    double v0 = acc[0][j], v1 = acc[1][j];
                    if (FORCING) {
                        const double2 f = __ldg(reinterpret_cast<const double2*>(f_row));
                        v0 = fma(g.gamma, f.x, v0);
                        v1 = fma(g.gamma, f.y, v1);
                    }
                    __stcs(reinterpret_cast<double2*>(o_row), make_double2(v0, v1));
                }
                o_row += g.so[1];
                if (FORCING) f_row += g.so[1];
            }
            acc[0][j] = 0.0;
            acc[1][j] = 0.0;
        }
        yrow_u += W * W * 16;
        __syncthreads();  // every thread is done with this ring slot
        if (tid == 0 && blk + NSTAGE < NB) issue(blk + NSTAGE, stage);
        if (++stage == NSTAGE) {
            stage = 0;
            parity ^= 1;
        }
    }
}

template <int P, int NWARP, int NSTAGE>
int launch_2d_cfg(const RhsOps& ops, const RhsGeom& g, cudaStream_t st) {
    constexpr int W = 2 * P + 1, PH = P + (P & 1), TXW = 64 * NWARP, RW = TXW + 2 * PH;
    constexpr int STAGE_STRIDE = (RW * W * 8 + 127) / 128 * 128;
    auto kern = g.forcing ? rhs2d_march_kernel<P, NWARP, NSTAGE, true> : rhs2d_march_kernel<P, NWARP, NSTAGE, false>;
    CUtensorMap map;
    const unsigned long long dims[3] = {(unsigned long long) g.in_n[0], (unsigned long long) g.in_n[1], 1ull};
    const unsigned long long strides[2] = {(unsigned long long) g.si[1], (unsigned long long) g.si[1] * g.in_n[1]};
    const unsigned box[3] = {(unsigned) RW, (unsigned) W, 1u};
    if (RW > 256 || !encode_tensor_map3(&map, g.in, dims, strides, box)) return -1;
    const int tx = (g.out_n[0] + TXW - 1) / TXW;
    const int fixed = NSTAGE * STAGE_STRIDE + NSTAGE * 8 + 8 + 128;
    // y segments: ~3 CTAs per SM worth of strips, each at least 16 p rows long and short enough for its table
    const int cap = (64 * 1024 - fixed) / (W * 16) - 2 * P - W;
    const long long slots = 3ll * ((g.max_sms > 0 && g.max_sms < sm_count_tma()) ? g.max_sms : sm_count_tma());
    int nseg = (int) ((slots + tx - 1) / tx);
    if (nseg < 1) nseg = 1;
    int yseg = (g.out_n[1] + nseg - 1) / nseg;
    if (yseg < 16 * P) yseg = 16 * P < g.out_n[1] ? 16 * P : g.out_n[1];
    if (yseg > cap) yseg = cap;
    if (yseg < 1) return -1;
    nseg = (g.out_n[1] + yseg - 1) / yseg;
    const int nb = (yseg + 2 * P + W - 1) / W;
    const int smem = fixed + nb * W * W * 16;
    cudaError_t e = cudaFuncSetAttribute((const void*) kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int) e;
    dim3 grid(tx, nseg, 1), block(NWARP * 32, 1, 1);
    kern<<<grid, block, smem, st>>>(map, ops, g, yseg);
    return (int) cudaGetLastError();
}

}  // namespace

// 2-D marching kernel.  0: launched; -1: not eligible / not enabled; else a cudaError_t.
int launch_rhs2d_march(const RhsOps& ops, const RhsGeom& g, cudaStream_t st) {
    static const bool enabled = [] {
        const char* e = getenv("ADSB_RHS2D_MARCH");
        return !e || atoi(e) != 0;
    }();
    if (!enabled || !ops.MSzT) return -1;
    const int p = ops.p[0];
    if (ops.p[1] != p) return -1;
    auto even = [](long long v) { return (v & 1) == 0; };
    if (g.si[0] != 1 || g.so[0] != 1 || !even(g.si[1]) || !even(g.so[1])) return -1;
    if (!even(g.out_n[0]) && !(g.out_lo[0] + g.out_n[0] == ops.n[0] && g.so[1] > g.out_n[0])) return -1;
    if ((uintptr_t) g.in % 16 || (uintptr_t) g.out % 16 || (g.forcing && (uintptr_t) g.forcing % 16)) return -1;
    for (int d = 0; d < 2; ++d)
        if (g.in_lo[d] < 0 || g.in_lo[d] + g.in_n[d] > ops.n[d]) return -1;
    switch (p) {
    case 2: return launch_2d_cfg<2, 3, 4>(ops, g, st);   // strips of 192 x (raw row 196 doubles <= TMA box limit 256)
    case 3: return launch_2d_cfg<3, 3, 4>(ops, g, st);
    default: return -1;
    }
}

// 0: launched; -1: this problem is not eligible (caller uses the cp.async kernel); else a cudaError_t.
// narrow: the output box is at most 16 columns wide (x remainder of the 64-wide tiling): 8-pair-wide warps.
int launch_rhs_tma(const RhsOps& ops, const RhsGeom& g, cudaStream_t st, bool narrow) {
    const int p = ops.p[0];
    if (ops.p[1] != p || ops.p[2] != p) return -1;
    auto even = [](long long v) { return (v & 1) == 0; };
    if (g.si[0] != 1 || g.so[0] != 1) return -1;
    // TMA: 16 B aligned base and strides; pair stores: even output extents / strides
    if (!even(g.si[1]) || !even(g.si[2]) || !even(g.so[1]) || !even(g.so[2])) return -1;
    // an odd row length is fine when the last pair's second element is a pad column: the box ends at the
    // domain edge and the output rows are pitched wider than the box (the managed tensors' layout)
    if (!even(g.out_n[0]) && !(g.out_lo[0] + g.out_n[0] == ops.n[0] && g.so[1] > g.out_n[0])) return -1;
    if ((uintptr_t) g.in % 16 || (uintptr_t) g.out % 16 || (g.forcing && (uintptr_t) g.forcing % 16)) return -1;
    for (int d = 0; d < 3; ++d)  // outside `in` reads as zero: only right when `in` does not stick out of the domain
        if (g.in_lo[d] < 0 || g.in_lo[d] + g.in_n[d] > ops.n[d]) return -1;
    static const int variant = [] {
        const char* e = getenv("ADSB_RHS_TMA_VARIANT");
        return e ? atoi(e) : 0;
    }();
    if (narrow) {
        if (g.out_n[0] > 16) return -1;
        switch (p) {
        case 2: return launch_cfg<2, 4, 4, 3, 2, 8>(ops, g, st);
        case 3: return launch_cfg<3, 2, 6, 3, 1, 8>(ops, g, st);
        default: return -1;
        }
    }
    switch (p) {
    case 2:
        if (variant == 1) return launch_cfg<2, 4, 4, 4, 2>(ops, g, st);
        if (variant == 2) return launch_cfg<2, 4, 8, 3, 1>(ops, g, st);
        if (variant == 3) return launch_cfg<2, 2, 4, 3, 3>(ops, g, st);
        if (variant == 4) return launch_cfg<2, 2, 8, 3, 2>(ops, g, st);
        return launch_cfg<2, 4, 4, 3, 2>(ops, g, st);
    case 3:
        if (variant == 1) return launch_cfg<3, 2, 6, 3, 2>(ops, g, st);
        return launch_cfg<3, 2, 6, 3, 1>(ops, g, st);
    case 4:
        if (variant == 1) return launch_cfg<4, 2, 6, 3, 1>(ops, g, st);
        return launch_cfg<4, 1, 8, 3, 1>(ops, g, st);
    case 5:
        if (variant == 1) return launch_cfg<5, 2, 4, 3, 1>(ops, g, st);
        return launch_cfg<5, 1, 8, 3, 1>(ops, g, st);
    default: return -1;
    }
}

}  // namespace adsb
