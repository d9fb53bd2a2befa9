// quadbrick.cuh -- K1 (general pointwise forms): right-hand side by Gauss quadrature, sum-factorised over a
// BRICK of elements per CTA, FP64, sm_100a.  method ADSB_RHS_QUADRATURE of adsb_compute_rhs and
// adsb_compute_rhs_pointwise.
//
// What it replaces: the element loop of every compute_rhs() of the reference -- zero(rhs); for e: for q:
// u = eval_fun(u_prev, e, q); for a: U(a) += form(u, v_a) w J; update_global_rhs (examples/scalability/
// test3d.hpp:66-95 is the model, examples/flow/flow.hpp:74-101 a nonlinear instance; eval_fun / eval_basis:
// include/ads/simulation/simulation_3d.hpp:64-145).  The pointwise form is a template functor.
//
// The element loop costs 9 m^4 FMA per element (m = p + 1) when every element interpolates for itself.  Here a
// CTA owns EX x EY element columns and marches along z, and the three 1-D stages are shared between
// neighbouring elements (2 m^2 + 3 m^3 + 4 m^4 FMA per DOF and direction):
//   per DOF plane c   X stage  coefficients (a, b) -> (gx, b):  V = Bx c, D = Bx' c           shared memory
//                     Y stage  (gx, b) -> (gx, gy):  w = By V, wx = By D, wy = By' V          into REGISTERS
//   per element ez    Z stage  every thread owns PT point columns (gx, gy) and keeps the last m planes of
//                     (w, wx, wy) in registers: u, du/dx, du/dy, du/dz at the q Gauss points of the element;
//                     the pointwise form; Z^T stage onto m partial planes (t0, t1, t2), also registers --
//                     the 8 m^4 part of the work never leaves the register file;
//   per finished DOF plane    Y^T (partials per element row through shared memory, summed in a fixed order),
//                     X^T, and one read-modify-write of the plane's (EX + p)(EY + p) DOFs in global memory.
// Neighbouring bricks share their p-wide rims, so the bricks are launched in COLOURS (2 per axis when a brick
// is at least p elements wide; z is cut into segments, two colours as well): bricks of one launch touch
// disjoint DOFs and add without atomics, the colours run in stream order -> the sum is deterministic.  The
// output starts as gamma * F (or zero) from a small init kernel.
// Tables b[e][q][d][i], x[e][q], w[q], J[e] are the reference's own (src/ads/basis_data.cpp:63-114), staged in
// shared memory per CTA; the bound is the FP64 pipe.
#ifndef ADSB_QUADBRICK_CUH
#define ADSB_QUADBRICK_CUH

#include <algorithm>

#include "kernels.cuh"

namespace adsb {
namespace qb {

constexpr double PI = 3.14159265358979323846;

__device__ __forceinline__ double source_value(int src, bool d3, double x, double y, double z) {
    if (src == 1) {  // examples/scalability/test3d.hpp:58-64, test2d.hpp:49-54
        const double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
        if (d3) return exp(-sqrt(dx * dx + dy * dy + dz * dz)) + 1 + cos(PI * x) * cos(PI * y) * cos(PI * z);
        return exp(-sqrt(dx * dx + dy * dy)) + 1 + cos(PI * x) * cos(PI * y);
    }
    if (src == 2) {  // examples/flow/flow.hpp:124-128
        const double pi2 = 2 * PI;
        return 1 + sin(pi2 * x) * sin(pi2 * y) * sin(pi2 * z);
    }
    return 0.0;
}

struct PointIn {
    double u, ux, uy, uz;  // u_prev and its gradient at the point
    double x, y, z;        // the point
    double coef;           // the caller's coefficient table at the point (forms with USES_COEF)
    double col, zf;        // the form's own column_value(x, y) and plane_value(z) (forms with SEPARABLE): factors of
                           // a separable coefficient, evaluated once per point column / once per z point
};
struct PointOut {
    double k0, k1, k2, k3;  // the integrand is k0 v + k1 dv/dx + k2 dv/dy + k3 dv/dz   (before the factor w J)
    double kp;              // PLAIN forms: a term added to every DOF of the element without the test function
};

// rhs = alpha u v - sum_k beta_k d_k u d_k v - (adv . grad u) v + gamma f(x) [v]
// (heat_3d.hpp:49-67, heat_2d.hpp:80-106, implicit.hpp:132-182, scalability test3d.hpp:66-95 with PLAIN: its
// forcing enters without the test function, test3d.hpp:86-88; the advection term is pollution_3d.hpp's)
template <bool PLAIN_>
struct FormLinear {
    static constexpr bool PLAIN = PLAIN_, USES_COEF = false, SEPARABLE = false;
    double alpha, beta[3], adv[3], gamma;
    int source;
    __device__ __forceinline__ void operator()(const PointIn& in, PointOut& o, bool d3) const {
        o.k0 = alpha * in.u - (adv[0] * in.ux + adv[1] * in.uy + adv[2] * in.uz);
        o.k1 = -beta[0] * in.ux;
        o.k2 = -beta[1] * in.uy;
        o.k3 = -beta[2] * in.uz;
        o.kp = 0.0;
        if (source) {
            const double f = gamma * source_value(source, d3, in.x, in.y, in.z);
            if (PLAIN)
                o.kp = f;
            else
                o.k0 += f;
        }
    }
};

// examples/flow/flow.hpp:74-101:  (u v + dt (-k(x) exp(mi u) grad u . grad v + h(x) v)) w J, k tabulated at the
// quadrature points (fill_permeability_map, flow.hpp:53-60)
struct FormFlow {
    static constexpr bool PLAIN = false, USES_COEF = true, SEPARABLE = true;
    double dt, mi;
    // the forcing 1 + sin(2 pi x) sin(2 pi y) sin(2 pi z) (flow.hpp:124-128) in its separable factors
    __device__ __forceinline__ double column_value(double x, double y) const { return sin(2 * PI * x) * sin(2 * PI * y); }
    __device__ __forceinline__ double plane_value(double z) const { return sin(2 * PI * z); }
    __device__ __forceinline__ void operator()(const PointIn& in, PointOut& o, bool) const {
        const double h = 1 + in.col * in.zf;
        const double e = -dt * in.coef * exp(mi * in.u);
        o.k0 = in.u + dt * h;
        o.k1 = e * in.ux;
        o.k2 = e * in.uy;
        o.k3 = e * in.uz;
        o.kp = 0.0;
    }
};

struct BrickGrid {
    int elo[3], en[3];               // elements to integrate: [elo, elo + en)
    int ntx, nty, nseg, seg_len;     // bricks along x, y; z segments of seg_len elements
    int cx, cy, cz, ncx, ncy, ncz;   // this launch's colour and the colour counts
    const double* coef;              // per-point coefficient table of the whole domain, x fastest, or nullptr
    long long cq1, cq2;              // its row and plane strides (points)
};

// brick shape per degree: EX x EY element columns, PT point columns per thread (PT divides q = p + 1).  The
// register file is split over the four SM sub-partitions, so what counts is warps per sub-partition:
// <= 8 warps leave 255 registers per thread, 9 ... 12 warps 168, 13 ... 16 warps 128.
template <int P> struct BrickCfg;
template <> struct BrickCfg<1> { static constexpr int EX = 16, EY = 8, PT = 2; };
template <> struct BrickCfg<2> { static constexpr int EX = 14, EY = 6, PT = 3; };  // 252 threads = 8 warps: 255 registers
template <> struct BrickCfg<3> { static constexpr int EX = 8, EY = 4, PT = 2; };
template <> struct BrickCfg<4> { static constexpr int EX = 4, EY = 4, PT = 1; };
template <> struct BrickCfg<5> { static constexpr int EX = 4, EY = 3, PT = 2; };   // 216 threads

template <int P, bool D3, bool PLAIN>
struct BrickDims {
    using C = BrickCfg<P>;
    static constexpr int EX = C::EX, EY = C::EY, PT = C::PT;
    static constexpr int M = P + 1, Q = P + 1, NG = Q / PT, GX = EX * Q, GY = EY * Q, DXn = EX + P, DYn = EY + P;
    static constexpr int MZ = D3 ? M : 1, QZ = D3 ? Q : 1, PZ = D3 ? P : 0;
    static constexpr int NR = EY * NG, NT = GX * NR, NC = DXn * DYn;
    static_assert(Q % PT == 0, "PT must divide the number of quadrature points");
    static_assert(NC <= NT && NT <= 1024, "brick shape");
    // shared-memory carve-up (doubles)
    static constexpr int oBx = 0;                        // [2][M][GX]
    static constexpr int oBy = oBx + 2 * M * GX;         // [EY][Q][2][M]
    static constexpr int oBz = oBy + EY * Q * 2 * M;     // [2][QZ][2][MZ]
    static constexpr int oWx = oBz + 2 * QZ * 2 * MZ;    // [GX]  w J      (the z tables are double-buffered)
    static constexpr int oXx = oWx + GX;                 // [GX]  point coordinates
    static constexpr int oWy = oXx + GX;                 // [GY]
    static constexpr int oXy = oWy + GY;
    static constexpr int oWz = oXy + GY;                 // [2][QZ]
    static constexpr int oXz = oWz + 2 * QZ;
    static constexpr int oZf = oXz + 2 * QZ;             // [2][QZ]  the form's plane_value(z)
    static constexpr int oC = oZf + 2 * QZ;              // [DYn][DXn]  coefficient plane
    static constexpr int oV = oC + NC;                   // [DYn][GX]
    static constexpr int oD = oV + DYn * GX;
    static constexpr int oP0 = oD + DYn * GX;            // [NR][M][GX]
    static constexpr int oP1 = oP0 + NR * M * GX;
    static constexpr int oPp = oP1 + NR * M * GX;        // [NR][GX]
    static constexpr int oS0 = oPp + (PLAIN ? NR * GX : 0);  // [DYn][GX]
    static constexpr int oS1 = oS0 + DYn * GX;
    static constexpr int oSp = oS1 + DYn * GX;
    static constexpr int total = oSp + (PLAIN ? DYn * GX : 0);
};

template <int P, bool D3, class Form>
__global__ void __launch_bounds__((BrickDims<P, D3, Form::PLAIN>::NT), 1)
    quad_brick_kernel(const QuadAxes A, const RhsGeom g, const Form form, const BrickGrid G) {
    using B = BrickDims<P, D3, Form::PLAIN>;
    constexpr int EX = B::EX, EY = B::EY, PT = B::PT, M = B::M, Q = B::Q, NG = B::NG, GX = B::GX, GY = B::GY;
    constexpr int DXn = B::DXn, DYn = B::DYn, MZ = B::MZ, QZ = B::QZ, PZ = B::PZ, NT = B::NT, NC = B::NC;
    constexpr bool PLAIN = Form::PLAIN;
    extern __shared__ double sm[];
    double* const sBx = sm + B::oBx;
    double* const sBy = sm + B::oBy;
    double* const sBz = sm + B::oBz;
    double* const sWx = sm + B::oWx;
    double* const sXx = sm + B::oXx;
    double* const sWy = sm + B::oWy;
    double* const sXy = sm + B::oXy;
    double* const sWz = sm + B::oWz;
    double* const sXz = sm + B::oXz;
    double* const sZf = sm + B::oZf;
    double* const sC = sm + B::oC;
    double* const sV = sm + B::oV;
    double* const sD = sm + B::oD;
    double* const sP0 = sm + B::oP0;
    double* const sP1 = sm + B::oP1;
    double* const sPp = sm + B::oPp;
    double* const sS0 = sm + B::oS0;
    double* const sS1 = sm + B::oS1;
    double* const sSp = sm + B::oSp;

    const int tid = threadIdx.x;
    const int tx = G.cx + G.ncx * (int) blockIdx.x, ty = G.cy + G.ncy * (int) blockIdx.y, ts = G.cz + G.ncz * (int) blockIdx.z;
    if (tx >= G.ntx || ty >= G.nty || ts >= G.nseg) return;
    const int ex0 = G.elo[0] + tx * EX, ey0 = G.elo[1] + ty * EY;
    const int ex_end = G.elo[0] + G.en[0], ey_end = G.elo[1] + G.en[1];
    const int ez0 = D3 ? G.elo[2] + ts * G.seg_len : 0;
    const int nez = D3 ? min(G.seg_len, G.elo[2] + G.en[2] - ez0) : 1;

    // ---- the brick's table slices; elements beyond the range get zero tables and zero weights
    for (int i = tid; i < 2 * M * GX; i += NT) {
        const int d = i / (M * GX), r = i % (M * GX), j = r / GX, gx = r % GX;
        const int ex = ex0 + gx / Q, qx = gx % Q;
        sBx[i] = ex < ex_end ? A.bt[0][((size_t) ex * Q + qx) * 2 * M + d * M + j] : 0.0;
    }
    for (int i = tid; i < EY * Q * 2 * M; i += NT) {
        const int ey = ey0 + i / (Q * 2 * M);
        sBy[i] = ey < ey_end ? A.bt[1][(size_t) ey * Q * 2 * M + i % (Q * 2 * M)] : 0.0;
    }
    for (int i = tid; i < GX; i += NT) {
        const int ex = ex0 + i / Q;
        const bool ok = ex < ex_end;
        sWx[i] = ok ? A.w[0][i % Q] * A.J[0][ex] : 0.0;
        sXx[i] = ok ? A.xq[0][ex * Q + i % Q] : 0.0;
    }
    for (int i = tid; i < GY; i += NT) {
        const int ey = ey0 + i / Q;
        const bool ok = ey < ey_end;
        sWy[i] = ok ? A.w[1][i % Q] * A.J[1][ey] : 0.0;
        sXy[i] = ok ? A.xq[1][ey * Q + i % Q] : 0.0;
    }
    __syncthreads();

    // ---- this thread's point columns: gx, element row ey, points qy = s PT .. s PT + PT - 1 of it
    const int gx = tid % GX, r = tid / GX, ey = r / NG, s = r % NG;
    const bool col_ok = (ex0 + gx / Q) < ex_end && (ey0 + ey) < ey_end;
    double wJxy[PT], colv[Form::SEPARABLE ? PT : 1];
#pragma unroll
    for (int t = 0; t < PT; ++t) {
        wJxy[t] = sWx[gx] * sWy[ey * Q + s * PT + t];
        if constexpr (Form::SEPARABLE) colv[t] = form.column_value(sXx[gx], sXy[ey * Q + s * PT + t]);
    }
    const double* const byT = sBy + (ey * Q + s * PT) * 2 * M;  // [t][d][j]
    const long long coef_col = G.coef ? (long long) (ex0 * Q + gx) + G.cq1 * (long long) ((ey0 + ey) * Q + s * PT) : 0;

    double Ww[MZ][PT], Wx[MZ][PT], Wy[MZ][PT];  // the last MZ planes of By V, By D, By' V at the thread's columns
    double T0[MZ][PT], T1[MZ][PT], T2[MZ][PT];  // partial DOF planes: sum_gz Bz k0 + Bz' k3, Bz k1, Bz k2
    double Tp[PLAIN ? MZ : 1][PT];
#pragma unroll
    for (int j = 0; j < MZ; ++j)
#pragma unroll
        for (int t = 0; t < PT; ++t) {
            Ww[j][t] = Wx[j][t] = Wy[j][t] = 0.0;
            T0[j][t] = T1[j][t] = T2[j][t] = 0.0;
            if (PLAIN) Tp[j][t] = 0.0;
        }

    // ---- the DOF column (a, b) = (tid % DXn, tid / DXn) this thread loads and, later, adds to
    const int ca = tid % DXn, cb = tid / DXn;
    const int ga = ex0 + ca, gb = ey0 + cb;
    const bool in_xy = tid < NC && ga >= g.in_lo[0] && ga < g.in_lo[0] + g.in_n[0] && gb >= g.in_lo[1] && gb < g.in_lo[1] + g.in_n[1];
    const bool out_xy = tid < NC && ga >= g.out_lo[0] && ga < g.out_lo[0] + g.out_n[0] && gb >= g.out_lo[1] && gb < g.out_lo[1] + g.out_n[1];
    const double* const src_col = g.in + (long long) (ga - g.in_lo[0]) * g.si[0] + (long long) (gb - g.in_lo[1]) * g.si[1];
    double* const dst_col = g.out + (long long) (ga - g.out_lo[0]) * g.so[0] + (long long) (gb - g.out_lo[1]) * g.so[1];
    auto load_c = [&](int c) -> double {  // coefficient of DOF plane c, zero outside the input box
        if (!in_xy) return 0.0;
        if (D3 && (c < g.in_lo[2] || c >= g.in_lo[2] + g.in_n[2])) return 0.0;
        return __ldg(src_col + (D3 ? (long long) (c - g.in_lo[2]) * g.si[2] : 0));
    };
    auto out_ptr = [&](int c) -> double* {  // DOF plane c of the output, nullptr outside the out box
        if (!out_xy) return nullptr;
        if (D3 && (c < g.out_lo[2] || c >= g.out_lo[2] + g.out_n[2])) return nullptr;
        return dst_col + (D3 ? (long long) (c - g.out_lo[2]) * g.so[2] : 0);
    };

    // Iteration k:  plane ez0 + k enters (X, Y stages), element ez0 + k - PZ is integrated (Z, form, Z^T), the
    // partial plane ez0 + k - PZ leaves (Y^T partials).  Its reduction (Y^T sums, X^T, read-modify-write) runs
    // one iteration later, next to the X stage and at the head of the register phase, so an iteration has two
    // barriers and the global read-modify-write latency hides behind the X stage.
    const int niter = nez + 2 * PZ;
    double creg = load_c(ez0);
#pragma unroll 1
    for (int k = 0; k <= niter; ++k) {
        const bool live = k < niter;
        const bool interp = k < nez + PZ;           // DOF plane ez0 + k enters
        const int e = k - PZ;                       // element ez0 + e is integrated
        const bool elem = live && e >= 0 && e < nez;
        const bool emit = live && k >= PZ;          // partial plane ez0 + k - PZ leaves
        const bool drain = k >= 1 && k - 1 >= PZ;   // the plane that left in iteration k - 1 is reduced now
        double* const sBzk = sBz + (k & 1) * (QZ * 2 * MZ);
        double* const sWzk = sWz + (k & 1) * QZ;
        double* const sXzk = sXz + (k & 1) * QZ;
        double* const sZfk = sZf + (k & 1) * QZ;
        if (live && tid < NC) sC[tid] = creg;
        if (elem) {
            if (D3) {
                for (int i = tid; i < QZ * 2 * MZ; i += NT) sBzk[i] = A.bt[2][(size_t) (ez0 + e) * Q * 2 * M + i];
                for (int i = tid; i < QZ; i += NT) {
                    sWzk[i] = A.w[2][i] * A.J[2][ez0 + e];
                    sXzk[i] = A.xq[2][(ez0 + e) * Q + i];
                    if constexpr (Form::SEPARABLE) sZfk[i] = form.plane_value(A.xq[2][(ez0 + e) * Q + i]);
                }
            } else if (tid == 0) {
                sBzk[0] = 1.0;  // value
                sBzk[1] = 0.0;  // derivative
                sWzk[0] = 1.0;
                sXzk[0] = 0.0;
                if constexpr (Form::SEPARABLE) sZfk[0] = form.plane_value(0.0);
            }
        }
        double* const dst = drain ? out_ptr(ez0 + k - 1 - PZ) : nullptr;
        const double old = dst ? *dst : 0.0;  // in flight across the X stage
        __syncthreads();
        if (interp && k + 1 < nez + PZ) creg = load_c(ez0 + k + 1);  // in flight during the whole iteration

        // ---- X stage: V(gx, b) = sum_i Bx(gx, i) c(ex + i, b), D likewise with the derivatives
        if (interp) {
            for (int j = tid; j < GX * DYn; j += NT) {
                const int b = j / GX, x = j % GX, ex = x / Q;
                double v = 0.0, d = 0.0;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const double cv = sC[b * DXn + ex + i];
                    v = fma(sBx[i * GX + x], cv, v);
                    d = fma(sBx[(M + i) * GX + x], cv, d);
                }
                sV[j] = v;
                sD[j] = d;
            }
        }
        // ---- Y^T of the previous plane: S(gx, b) = sum over the element rows b - p .. b and their point groups
        if (drain) {
            for (int j = tid; j < GX * DYn; j += NT) {
                const int b = j / GX, x = j % GX;
                double s0 = 0.0, s1 = 0.0, sp = 0.0;
#pragma unroll
                for (int i = 0; i <= P; ++i) {
                    const int el = b - i;
                    if (el >= 0 && el < EY) {
#pragma unroll
                        for (int gq = 0; gq < NG; ++gq) {
                            const int rr = el * NG + gq;
                            s0 += sP0[(rr * M + i) * GX + x];
                            s1 += sP1[(rr * M + i) * GX + x];
                            if (PLAIN) sp += sPp[rr * GX + x];
                        }
                    }
                }
                sS0[j] = s0;
                sS1[j] = s1;
                if (PLAIN) sSp[j] = sp;
            }
        }
        __syncthreads();

        // ---- X^T of the previous plane and its read-modify-write (bricks of one launch touch disjoint DOFs)
        if (dst) {
            double v0 = 0.0, v1 = 0.0, vp = 0.0;
#pragma unroll
            for (int i = 0; i <= P; ++i) {
                const int el = ca - i;
                if (el >= 0 && el < EX) {
#pragma unroll
                    for (int qx = 0; qx < Q; ++qx) {
                        const int x = el * Q + qx;
                        v0 = fma(sBx[i * GX + x], sS0[cb * GX + x], v0);
                        v1 = fma(sBx[(M + i) * GX + x], sS1[cb * GX + x], v1);
                        if (PLAIN) vp += sSp[cb * GX + x];
                    }
                }
            }
            *dst = old + ((v0 + v1) + vp);
        }
        if (!live) break;

        // ---- Y stage into the register window
        if (interp) {
#pragma unroll
            for (int j = 0; j + 1 < MZ; ++j)
#pragma unroll
                for (int t = 0; t < PT; ++t) {
                    Ww[j][t] = Ww[j + 1][t];
                    Wx[j][t] = Wx[j + 1][t];
                    Wy[j][t] = Wy[j + 1][t];
                }
            double v[M], d[M];
#pragma unroll
            for (int j = 0; j < M; ++j) {
                v[j] = sV[(ey + j) * GX + gx];
                d[j] = sD[(ey + j) * GX + gx];
            }
#pragma unroll
            for (int t = 0; t < PT; ++t) {
                double w = 0.0, wx = 0.0, wy = 0.0;
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    const double by = byT[(t * 2 + 0) * M + j], dby = byT[(t * 2 + 1) * M + j];
                    w = fma(by, v[j], w);
                    wx = fma(by, d[j], wx);
                    wy = fma(dby, v[j], wy);
                }
                Ww[MZ - 1][t] = w;
                Wx[MZ - 1][t] = wx;
                Wy[MZ - 1][t] = wy;
            }
        }

        // ---- Z stage, the pointwise form, Z^T stage: registers only
        if (elem) {
#pragma unroll
            for (int qz = 0; qz < QZ; ++qz) {
                double bz[MZ], dbz[MZ];
#pragma unroll
                for (int j = 0; j < MZ; ++j) {
                    bz[j] = sBzk[(qz * 2 + 0) * MZ + j];
                    dbz[j] = sBzk[(qz * 2 + 1) * MZ + j];
                }
                const double wz = sWzk[qz];
#pragma unroll
                for (int t = 0; t < PT; ++t) {
                    PointIn in;
                    in.u = in.ux = in.uy = in.uz = 0.0;
#pragma unroll
                    for (int j = 0; j < MZ; ++j) {
                        in.u = fma(bz[j], Ww[j][t], in.u);
                        in.ux = fma(bz[j], Wx[j][t], in.ux);
                        in.uy = fma(bz[j], Wy[j][t], in.uy);
                        if (D3) in.uz = fma(dbz[j], Ww[j][t], in.uz);
                    }
                    in.x = sXx[gx];
                    in.y = sXy[ey * Q + s * PT + t];
                    in.z = sXzk[qz];
                    in.coef = in.col = in.zf = 0.0;
                    if constexpr (Form::SEPARABLE) {
                        in.col = colv[t];
                        in.zf = sZfk[qz];
                    }
                    if (Form::USES_COEF) {
                        if (col_ok && G.coef)
                            in.coef = __ldg(G.coef + coef_col + G.cq1 * t + (D3 ? G.cq2 * (long long) ((ez0 + e) * Q + qz) : 0));
                    }
                    PointOut o;
                    form(in, o, D3);
                    const double wJ = wJxy[t] * wz;
                    const double k0 = o.k0 * wJ, k1 = o.k1 * wJ, k2 = o.k2 * wJ, k3 = o.k3 * wJ;
#pragma unroll
                    for (int j = 0; j < MZ; ++j) {
                        T0[j][t] = fma(bz[j], k0, T0[j][t]);
                        if (D3) T0[j][t] = fma(dbz[j], k3, T0[j][t]);
                        T1[j][t] = fma(bz[j], k1, T1[j][t]);
                        T2[j][t] = fma(bz[j], k2, T2[j][t]);
                        if (PLAIN) Tp[j][t] = fma(o.kp, wJ, Tp[j][t]);
                    }
                }
            }
        }

        // ---- the oldest partial plane is complete: Y^T partial sums of this thread's points
        if (emit) {
#pragma unroll
            for (int j = 0; j < M; ++j) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int t = 0; t < PT; ++t) {
                    const double by = byT[(t * 2 + 0) * M + j], dby = byT[(t * 2 + 1) * M + j];
                    s0 = fma(by, T0[0][t], s0);
                    s0 = fma(dby, T2[0][t], s0);
                    s1 = fma(by, T1[0][t], s1);
                }
                sP0[(r * M + j) * GX + gx] = s0;
                sP1[(r * M + j) * GX + gx] = s1;
            }
            if (PLAIN) {
                double sp = 0.0;
#pragma unroll
                for (int t = 0; t < PT; ++t) sp += Tp[0][t];
                sPp[r * GX + gx] = sp;
            }
#pragma unroll
            for (int j = 0; j + 1 < MZ; ++j)
#pragma unroll
                for (int t = 0; t < PT; ++t) {
                    T0[j][t] = T0[j + 1][t];
                    T1[j][t] = T1[j + 1][t];
                    T2[j][t] = T2[j + 1][t];
                    if (PLAIN) Tp[j][t] = Tp[j + 1][t];
                }
#pragma unroll
            for (int t = 0; t < PT; ++t) {
                T0[MZ - 1][t] = T1[MZ - 1][t] = T2[MZ - 1][t] = 0.0;
                if (PLAIN) Tp[MZ - 1][t] = 0.0;
            }
        }
    }
}

__global__ void init_box_kernel(double* y, const double* x, double a, long long n0, long long s1, long long s2);

// out = gamma * forcing (or zero) over the out box of g, then every colour of bricks in stream order.
// *nlaunch += kernels launched.  Returns a cudaError_t as int.
template <int P, bool D3, class Form>
int launch_brick(const QuadAxes& A, const RhsGeom& g, const Form& form, const int elo[3], const int en[3],
                 const double* coef, int max_sms, cudaStream_t st, int* nlaunch) {
    using B = BrickDims<P, D3, Form::PLAIN>;
    const size_t smem = (size_t) B::total * sizeof(double);
    auto kern = quad_brick_kernel<P, D3, Form>;
    if (smem > 48 * 1024) {  // per device, so not cached in a static
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return (int) e;
    }
    BrickGrid G{};
    for (int d = 0; d < 3; ++d) {
        G.elo[d] = elo[d];
        G.en[d] = en[d];
    }
    G.ntx = (en[0] + B::EX - 1) / B::EX;
    G.nty = (en[1] + B::EY - 1) / B::EY;
    G.ncx = G.ntx > 1 ? 1 + (P + B::EX - 1) / B::EX : 1;
    G.ncy = G.nty > 1 ? 1 + (P + B::EY - 1) / B::EY : 1;
    G.nseg = 1;
    G.seg_len = D3 ? en[2] : 1;
    if (D3) {
        // enough bricks per launch for ~8 waves of one CTA per SM; segments of at least max(32, p) elements
        const int sms = max_sms > 0 ? max_sms : 148;
        const long long per_colour = (long long) ((G.ntx + G.ncx - 1) / G.ncx) * ((G.nty + G.ncy - 1) / G.ncy);
        long long want = (8LL * sms + per_colour - 1) / per_colour;  // segments per z colour
        int nseg = (int) std::min<long long>(2 * want, std::max(1, en[2] / std::max(32, P)));
        if (nseg > 1 && (nseg & 1)) ++nseg;
        if (nseg > 1) {
            G.seg_len = (en[2] + nseg - 1) / nseg;
            G.nseg = (en[2] + G.seg_len - 1) / G.seg_len;
        }
    }
    G.ncz = G.nseg > 1 ? 1 + (P + G.seg_len - 1) / G.seg_len : 1;
    G.coef = coef;
    G.cq1 = (long long) A.ne[0] * B::Q;
    G.cq2 = G.cq1 * A.ne[1] * B::Q;
    for (int cz = 0; cz < G.ncz; ++cz)
        for (int cy = 0; cy < G.ncy; ++cy)
            for (int cx = 0; cx < G.ncx; ++cx) {
                const int nx = (G.ntx - cx + G.ncx - 1) / G.ncx, ny = (G.nty - cy + G.ncy - 1) / G.ncy,
                          nz = (G.nseg - cz + G.ncz - 1) / G.ncz;
                if (nx <= 0 || ny <= 0 || nz <= 0) continue;
                G.cx = cx;
                G.cy = cy;
                G.cz = cz;
                kern<<<dim3(nx, ny, nz), B::NT, smem, st>>>(A, g, form, G);
                if (nlaunch) ++*nlaunch;
            }
    return (int) cudaGetLastError();
}

// one translation unit per form instantiates its degrees
template <class Form>
int launch_brick_form(int ndim, const QuadAxes& A, const RhsGeom& g, const Form& form, const int elo[3], const int en[3],
                      const double* coef, int max_sms, cudaStream_t st, int* nlaunch);

#define ADSB_BRICK_DECLARE(FORM)                                                                                   \
    template <>                                                                                                    \
    int launch_brick_form<FORM>(int ndim, const QuadAxes& A, const RhsGeom& g, const FORM& form, const int elo[3], \
                                const int en[3], const double* coef, int max_sms, cudaStream_t st, int* nlaunch);
ADSB_BRICK_DECLARE(FormLinear<false>)
ADSB_BRICK_DECLARE(FormLinear<true>)
ADSB_BRICK_DECLARE(FormFlow)

#define ADSB_BRICK_DISPATCH(FORM, MAXP)                                                                             \
    template <>                                                                                                     \
    int launch_brick_form<FORM>(int ndim, const QuadAxes& A, const RhsGeom& g, const FORM& form, const int elo[3],  \
                                const int en[3], const double* coef, int max_sms, cudaStream_t st, int* nlaunch) {  \
        const int p = A.p[0];                                                                                       \
        if (p > MAXP) return (int) cudaErrorInvalidValue;                                                           \
        const bool d3 = ndim == 3;                                                                                  \
        switch (p) {                                                                                                \
        case 1: return d3 ? launch_brick<1, true, FORM>(A, g, form, elo, en, coef, max_sms, st, nlaunch)             \
                          : launch_brick<1, false, FORM>(A, g, form, elo, en, coef, max_sms, st, nlaunch);           \
        case 2: return d3 ? launch_brick<2, true, FORM>(A, g, form, elo, en, coef, max_sms, st, nlaunch)             \
                          : launch_brick<2, false, FORM>(A, g, form, elo, en, coef, max_sms, st, nlaunch);           \
        case 3: return d3 ? launch_brick<3, true, FORM>(A, g, form, elo, en, coef, max_sms, st, nlaunch)             \
                          : launch_brick<3, false, FORM>(A, g, form, elo, en, coef, max_sms, st, nlaunch);           \
        case 4: return d3 ? launch_brick<(MAXP >= 4 ? 4 : 1), true, FORM>(A, g, form, elo, en, coef, max_sms, st, nlaunch)  \
                          : launch_brick<(MAXP >= 4 ? 4 : 1), false, FORM>(A, g, form, elo, en, coef, max_sms, st, nlaunch); \
        case 5: return d3 ? launch_brick<(MAXP >= 5 ? 5 : 1), true, FORM>(A, g, form, elo, en, coef, max_sms, st, nlaunch)  \
                          : launch_brick<(MAXP >= 5 ? 5 : 1), false, FORM>(A, g, form, elo, en, coef, max_sms, st, nlaunch); \
        default: return (int) cudaErrorInvalidValue;                                                                \
        }                                                                                                           \
    }

}  // namespace qb
}  // namespace adsb

#endif
