// kernels_seg.cu -- segmented substitution: the boundary-state kernels and the correction pass.
//
// A sweep line is cut into segments (host_setup.cpp, build_segment_plan).  Pass A solves every segment on
// its own with the ordinary sweep kernels (kernels_sweep_tile.cu) and leaves xhat.  The exact dgbtrs
// result (include/ads/lin/band_solve.hpp:21-31) then needs, per line and segment, KL forward values din
// and KD backward values tin:
//   seg_dseg_kernel     Dseg_s = E_s * xhat_s[last KL rows]                     (what segment s sends right)
//   seg_din_kernel      din_s  = sum_d Wf[s][d] Dseg_{s-d};  X_s = xhat_s[first KD rows] + XiF_s din_s
//   seg_tin_kernel      tin_s  = sum_d Vb[s][d] X_{s+d}                         (only when the chain is deeper than 1)
//   seg_correct_*       x_j    = xhat_j + Psi(j,:) tin_s + Xi(j,:) din_s        (pass B, 16 B/DOF, streaming)
// In a slab-sharded run a segment is one GPU's slab: Dseg and X are the only data that cross NVLink
// (KL + KD doubles per line instead of the slab itself); the kernels simply store them through peer
// pointers into the neighbours' state arrays.  State arrays are [S][K][L]: segment, component, line.
#include <cstdint>
#include <type_traits>

#include "kernels.cuh"
#include "tma.cuh"

namespace adsb {

namespace {

constexpr int SEG_MAX_DST = 8;

struct DstList {
    double* p[SEG_MAX_DST];
    int n;
};

__device__ __forceinline__ long long line_off(int l0, int l1, long long s0, long long s1) {
    return (long long) l0 * s0 + (long long) l1 * s1;
}

// One thread per (line, segment).
template <int KL>
__global__ void __launch_bounds__(256) seg_dseg_kernel(const SegDev T, const SegGeom G, DstList dst) {
    pdl_wait();
    const int l0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int l1 = blockIdx.y;
    const int s = G.s_lo + blockIdx.z;
    if (l0 >= G.L0) return;
    const long long L = (long long) G.L0 * G.L1, line = l0 + (long long) G.L0 * l1;
    const int b = T.bounds[s + 1];
    const double* x = G.in + line_off(l0, l1, G.s0_in, G.s1_in) + (long long) (b - KL - G.row_base) * G.sj_in;
    double xl[KL];
#pragma unroll
    for (int m = 0; m < KL; ++m) xl[m] = __ldg(x + (long long) m * G.sj_in);
    const double* E = T.E + (size_t) s * KL * KL;
#pragma unroll
    for (int k = 0; k < KL; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < KL; ++m) acc = fma(__ldg(E + k * KL + m), xl[m], acc);
        for (int i = 0; i < dst.n; ++i) dst.p[i][((size_t) s * KL + k) * L + line] = acc;
    }
}

template <int KL, int KD>
__global__ void __launch_bounds__(256)
    seg_din_kernel(const SegDev T, const SegGeom G, const double* __restrict__ dseg, double* __restrict__ din, DstList xdst) {
    pdl_wait();
    const int l0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int l1 = blockIdx.y;
    const int s = G.s_lo + blockIdx.z;
    if (l0 >= G.L0) return;
    const long long L = (long long) G.L0 * G.L1, line = l0 + (long long) G.L0 * l1;
    double d[KL];
#pragma unroll
    for (int k = 0; k < KL; ++k) d[k] = 0.0;
    for (int dd = 1; dd <= T.DF && s - dd >= 0; ++dd) {
        const double* W = T.Wf + ((size_t) s * T.DF + dd - 1) * KL * KL;
        double e[KL];
#pragma unroll
        for (int q = 0; q < KL; ++q) e[q] = dseg[((size_t) (s - dd) * KL + q) * L + line];
#pragma unroll
        for (int k = 0; k < KL; ++k)
#pragma unroll
            for (int q = 0; q < KL; ++q) d[k] = fma(__ldg(W + k * KL + q), e[q], d[k]);
    }
#pragma unroll
    for (int k = 0; k < KL; ++k) din[((size_t) s * KL + k) * L + line] = d[k];
    const int a = T.bounds[s];
    const double* x = G.in + line_off(l0, l1, G.s0_in, G.s1_in) + (long long) (a - G.row_base) * G.sj_in;
    const double* XiF = T.XiF + (size_t) s * KD * KL;
#pragma unroll
    for (int i = 0; i < KD; ++i) {
        double acc = __ldg(x + (long long) i * G.sj_in);
#pragma unroll
        for (int q = 0; q < KL; ++q) acc = fma(__ldg(XiF + i * KL + q), d[q], acc);
        for (int t = 0; t < xdst.n; ++t) xdst.p[t][((size_t) s * KD + i) * L + line] = acc;
    }
}

template <int KD>
__global__ void __launch_bounds__(256)
    seg_tin_kernel(const SegDev T, int s_lo, long long L, const double* __restrict__ X, double* __restrict__ tin) {
    pdl_wait();
    const long long line = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int s = s_lo + blockIdx.y;
    if (line >= L) return;
    double t[KD];
#pragma unroll
    for (int k = 0; k < KD; ++k) t[k] = 0.0;
    for (int dd = 1; dd <= T.DB && s + dd < T.S; ++dd) {
        const double* V = T.Vb + ((size_t) s * T.DB + dd - 1) * KD * KD;
        double e[KD];
#pragma unroll
        for (int q = 0; q < KD; ++q) e[q] = X[((size_t) (s + dd) * KD + q) * L + line];
#pragma unroll
        for (int k = 0; k < KD; ++k)
#pragma unroll
            for (int q = 0; q < KD; ++q) t[k] = fma(__ldg(V + k * KD + q), e[q], t[k]);
    }
#pragma unroll
    for (int k = 0; k < KD; ++k) tin[((size_t) s * KD + k) * L + line] = t[k];
}

// Pass B, lanes along l0 (unit stride): a thread owns one line and walks RC rows of one segment.
// tin_is_x: the backward chain has depth 1, tin_s is X_{s+1} itself (no seg_tin_kernel ran).
template <int KL, int KD, int RC>
__global__ void __launch_bounds__(128)
    seg_correct_strided(const SegDev T, const SegGeom G, const double* __restrict__ din, const double* __restrict__ tin,
                        int tin_is_x, int chunks_per_seg) {
    pdl_wait();
    constexpr int KC = KD + KL;
    const int l0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int l1 = blockIdx.y;
    const int s = G.s_lo + blockIdx.z / chunks_per_seg, ck = blockIdx.z % chunks_per_seg;
    const int a = T.bounds[s], b = T.bounds[s + 1];
    const int j0 = a + ck * RC, j1 = min(b, j0 + RC);
    if (l0 >= G.L0 || j0 >= j1) return;
    const long long L = (long long) G.L0 * G.L1, line = l0 + (long long) G.L0 * l1;
    double st[KC];  // tin | din
    const int ts = tin_is_x ? s + 1 : s;
#pragma unroll
    for (int k = 0; k < KD; ++k) st[k] = (tin_is_x && s + 1 >= T.S) ? 0.0 : __ldg(tin + ((size_t) ts * KD + k) * L + line);
#pragma unroll
    for (int k = 0; k < KL; ++k) st[KD + k] = __ldg(din + ((size_t) s * KL + k) * L + line);
    const double* src = G.in + line_off(l0, l1, G.s0_in, G.s1_in) + (long long) (j0 - G.row_base) * G.sj_in;
    double* dst = G.out + line_off(l0, l1, G.s0_out, G.s1_out) + (long long) (j0 - G.row_base) * G.sj_out;
    const double* cf = T.cf + (size_t) j0 * KC;
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
        double acc = __ldcs(src);
#pragma unroll
        for (int k = 0; k < KC; ++k) acc = fma(__ldg(cf + k), st[k], acc);
        __stcs(dst, acc);
        src += G.sj_in;
        dst += G.sj_out;
        cf += KC;
    }
}

// Pass B when the lines of the view are contiguous in memory (l0 unit stride, l1 stride L0: the z sweep of an
// x-fastest tensor): the lines are numbered flat, a thread owns TWO neighbouring lines (128-bit accesses) and
// walks RC rows.  Blocks are ordered rows-fastest and from the LAST lines backwards: pass A finished with
// those lines, so on a slab that is not much larger than the 126 MB L2 most of pass B's reads hit L2.
template <int KL, int KD, int RC>
__global__ void __launch_bounds__(128)
    seg_correct_flat(const SegDev T, const SegGeom G, const double* __restrict__ din, const double* __restrict__ tin,
                     int tin_is_x, int chunks_per_seg) {
    pdl_wait();
    constexpr int KC = KD + KL;
    const long long L = (long long) G.L0 * G.L1;
    const long long line = 2 * ((long long) (gridDim.y - 1 - blockIdx.y) * blockDim.x + threadIdx.x);
    const int s = G.s_lo + blockIdx.x / chunks_per_seg, ck = blockIdx.x % chunks_per_seg;
    const int a = T.bounds[s], b = T.bounds[s + 1];
    const int j0 = a + ck * RC, j1 = min(b, j0 + RC);
    if (line >= L || j0 >= j1) return;
    double2 st[KC];  // tin | din of the two lines
    const int ts = tin_is_x ? s + 1 : s;
    const bool no_t = tin_is_x && s + 1 >= T.S;
#pragma unroll
    for (int k = 0; k < KD; ++k)
        st[k] = no_t ? make_double2(0.0, 0.0) : __ldg(reinterpret_cast<const double2*>(tin + ((size_t) ts * KD + k) * L + line));
#pragma unroll
    for (int k = 0; k < KL; ++k) st[KD + k] = __ldg(reinterpret_cast<const double2*>(din + ((size_t) s * KL + k) * L + line));
    const double* src = G.in + line + (long long) (j0 - G.row_base) * G.sj_in;
    double* dst = G.out + line + (long long) (j0 - G.row_base) * G.sj_out;
    const double* cf = T.cf + (size_t) j0 * KC;
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
        double2 acc = __ldcs(reinterpret_cast<const double2*>(src));
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const double c = __ldg(cf + k);
            acc.x = fma(c, st[k].x, acc.x);
            acc.y = fma(c, st[k].y, acc.y);
        }
        __stcs(reinterpret_cast<double2*>(dst), acc);
        src += G.sj_in;
        dst += G.sj_out;
        cf += KC;
    }
}

// Pass B, sweep axis contiguous: a thread owns one row j of a segment and walks LB lines.
template <int KL, int KD, int LB>
__global__ void __launch_bounds__(128)
    seg_correct_contig(const SegDev T, const SegGeom G, const double* __restrict__ din, const double* __restrict__ tin,
                       int tin_is_x, int chunks_per_seg) {
    pdl_wait();
    constexpr int KC = KD + KL;
    const int s = G.s_lo + blockIdx.x / chunks_per_seg, ck = blockIdx.x % chunks_per_seg;
    const int a = T.bounds[s], b = T.bounds[s + 1];
    const int j = a + ck * (int) blockDim.x + (int) threadIdx.x;
    if (j >= b) return;
    const long long L = (long long) G.L0 * G.L1;
    double c[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) c[k] = __ldg(T.cf + (size_t) j * KC + k);
    const int ts = tin_is_x ? s + 1 : s;
    const bool no_t = tin_is_x && s + 1 >= T.S;
    const long long lbase = (long long) blockIdx.y * LB;
    for (int q = 0; q < LB; ++q) {
        const long long line = lbase + q;
        if (line >= L) break;
        const int l0 = (int) (line % G.L0), l1 = (int) (line / G.L0);
        const long long o = (long long) (j - G.row_base);
        double acc = __ldcs(G.in + line_off(l0, l1, G.s0_in, G.s1_in) + o * G.sj_in);
#pragma unroll
        for (int k = 0; k < KD; ++k)
            if (!no_t) acc = fma(c[k], __ldg(tin + ((size_t) ts * KD + k) * L + line), acc);
#pragma unroll
        for (int k = 0; k < KL; ++k) acc = fma(c[KD + k], __ldg(din + ((size_t) s * KL + k) * L + line), acc);
        __stcs(G.out + line_off(l0, l1, G.s0_out, G.s1_out) + o * G.sj_out, acc);
    }
}

template <typename F>
int dispatch(int KL, int KD, F&& f) {
    // (KL, KD) = (P, P) without row interchanges, (P, 2P) with (same variants as the sweep kernels)
#define ADSB_SEG_CASE(P)                                                         \
    if (KL == P && KD == P) return f(std::integral_constant<int, P>{}, std::integral_constant<int, P>{}); \
    if (KL == P && KD == 2 * P) return f(std::integral_constant<int, P>{}, std::integral_constant<int, 2 * P>{});
    ADSB_SEG_CASE(1)
    ADSB_SEG_CASE(2)
    ADSB_SEG_CASE(3)
    ADSB_SEG_CASE(4)
    ADSB_SEG_CASE(5)
#undef ADSB_SEG_CASE
    return (int) cudaErrorInvalidValue;
}

DstList make_dst(double* const* dst, int ndst) {
    DstList d{};
    d.n = ndst < SEG_MAX_DST ? ndst : SEG_MAX_DST;
    for (int i = 0; i < d.n; ++i) d.p[i] = dst[i];
    return d;
}

}  // namespace

int launch_seg_dseg(const SegDev& T, const SegGeom& G, double* const* dst, int ndst, cudaStream_t st) {
    if (ndst < 1 || ndst > SEG_MAX_DST || G.s_hi <= G.s_lo) return (int) cudaErrorInvalidValue;
    const DstList d = make_dst(dst, ndst);
    return dispatch(T.KL, T.KD, [&](auto kl, auto) {
        constexpr int KL = decltype(kl)::value;
        dim3 block(128), grid((G.L0 + 127) / 128, G.L1, G.s_hi - G.s_lo);
        return (int) launch_ex(seg_dseg_kernel<KL>, grid, block, 0, st, true, T, G, d);
    });
}

int launch_seg_din(const SegDev& T, const SegGeom& G, const double* dseg, double* din, double* const* xdst, int ndst,
                   cudaStream_t st) {
    if (ndst < 1 || ndst > SEG_MAX_DST || G.s_hi <= G.s_lo) return (int) cudaErrorInvalidValue;
    const DstList d = make_dst(xdst, ndst);
    return dispatch(T.KL, T.KD, [&](auto kl, auto kd) {
        constexpr int KL = decltype(kl)::value, KD = decltype(kd)::value;
        dim3 block(128), grid((G.L0 + 127) / 128, G.L1, G.s_hi - G.s_lo);
        return (int) launch_ex(seg_din_kernel<KL, KD>, grid, block, 0, st, true, T, G, dseg, din, d);
    });
}

int launch_seg_tin(const SegDev& T, int s_lo, int s_hi, long long L, const double* X, double* tin, cudaStream_t st) {
    if (s_hi <= s_lo) return (int) cudaErrorInvalidValue;
    return dispatch(T.KL, T.KD, [&](auto, auto kd) {
        constexpr int KD = decltype(kd)::value;
        dim3 block(256), grid((unsigned) ((L + 255) / 256), s_hi - s_lo);
        return (int) launch_ex(seg_tin_kernel<KD>, grid, block, 0, st, true, T, s_lo, L, X, tin);
    });
}

int launch_seg_correct(const SegDev& T, const SegGeom& G, const double* din, const double* tin, int tin_is_x,
                       int max_rows, cudaStream_t st) {
    if (G.s_hi <= G.s_lo) return (int) cudaErrorInvalidValue;
    return dispatch(T.KL, T.KD, [&](auto kl, auto kd) {
        constexpr int KL = decltype(kl)::value, KD = decltype(kd)::value;
        if (G.sj_in == 1 && G.sj_out == 1) {
            constexpr int LB = 16;
            const int cps = (max_rows + 127) / 128;
            const long long L = (long long) G.L0 * G.L1;
            dim3 block(128), grid(cps * (G.s_hi - G.s_lo), (unsigned) ((L + LB - 1) / LB));
            return (int) launch_ex(seg_correct_contig<KL, KD, LB>, grid, block, 0, st, true, T, G, din, tin, tin_is_x, cps);
        } else {
            constexpr int RC = 16;
            const int cps = (max_rows + RC - 1) / RC;
            const long long L = (long long) G.L0 * G.L1;
            auto al16 = [](const void* p) { return ((uintptr_t) p & 15) == 0; };
            const bool flat = G.s0_in == 1 && G.s0_out == 1 && (G.L1 == 1 || (G.s1_in == G.L0 && G.s1_out == G.L0)) &&
                              L % 2 == 0 && G.sj_in % 2 == 0 && G.sj_out % 2 == 0 && al16(G.in) && al16(G.out) &&
                              al16(din) && al16(tin) && (L + 255) / 256 <= 65535;
            if (flat) {
                dim3 block(128), grid(cps * (G.s_hi - G.s_lo), (unsigned) ((L + 255) / 256));
                return (int) launch_ex(seg_correct_flat<KL, KD, RC>, grid, block, 0, st, true, T, G, din, tin, tin_is_x, cps);
            } else {
                dim3 block(128), grid((G.L0 + 127) / 128, G.L1, cps * (G.s_hi - G.s_lo));
                return (int) launch_ex(seg_correct_strided<KL, KD, RC>, grid, block, 0, st, true, T, G, din, tin, tin_is_x, cps);
            }
        }
        return (int) cudaGetLastError();
    });
}

}  // namespace adsb
