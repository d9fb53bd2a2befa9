// sweep_core.cuh -- the chunk-parallel banded substitution on register-resident chunks, shared by the
// sweep kernels (kernels_sweep.cu: operands straight from global memory; kernels_sweep_tile.cu:
// persistent CTAs fed by the TMA engine).  See kernels_sweep.cu for the algorithm.
#ifndef ADSB_SWEEP_CORE_CUH
#define ADSB_SWEEP_CORE_CUH

#include <cstdint>

#include "kernels.cuh"
#include "tma.cuh"

namespace adsb {

// CS: the coefficient tables were copied to shared memory (plain loads); else read-only global loads
template <bool CS>
__device__ __forceinline__ double ldc(const double* p) {
    return CS ? *p : __ldg(p);
}
template <int N, bool CS>
__device__ __forceinline__ void ldrec(const double* __restrict__ p, double* o) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
        const double2 t = CS ? *reinterpret_cast<const double2*>(p + k) : __ldg(reinterpret_cast<const double2*>(p + k));
        o[k] = t.x;
        o[k + 1] = t.y;
    }
    if (N & 1) o[N - 1] = ldc<CS>(p + N - 1);
}
// CTA-wide barrier, or a named barrier over the first `nsync` threads (consumer warps of a
// warp-specialised kernel)
__device__ __forceinline__ void sweep_sync(int nsync) {
    if (nsync == 0)
        __syncthreads();
    else
        asm volatile("bar.sync 1, %0;" ::"r"(nsync) : "memory");
}

// v[r][0 .. CH+KL): chunk c of RL lines (original data; rows past the end of the line are zero).
// On return v[r][0 .. CH) holds the solution.  fst / bst: shared scratch [SC][KL][NL], [SC][KD][NL]
// with NL = NLt * RL lines per CTA; line r of lane tx is r * NLt + tx.  Contains CTA-wide barriers.
template <int KL, int KD, bool PIV, int CH, int RL, bool CS = false>
__device__ __forceinline__ void sweep_core(const SweepFactor& F, double (&v)[RL][CH + KL], double* fst, double* bst,
                                           int c, int tx, int NLt, int SC, int nsync = 0) {
    constexpr int LF = sweep_pitch(KL), LB = sweep_pitch(KD + 1), LC = sweep_pitch(KD + KL);
    constexpr int RF = sweep_rec(KL), RB = sweep_rec(KD + 1), RC = sweep_rec(KD + KL);  // doubles loaded per record
    constexpr int MD = SWEEP_MAX_DEPTH_DEV;
    const int NL = NLt * RL;
    const int j0 = c * CH;
    // ---------------------------------------------------------------- F1: local forward
    double dl[RL][KL];
#pragma unroll
    for (int r = 0; r < RL; ++r)
#pragma unroll
        for (int k = 0; k < KL; ++k) dl[r][k] = v[r][CH + k];
    {
        const double* cf = F.cfF + (long long) j0 * LF;
        const int* pv = F.pv + j0;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            double L[RF];
            ldrec<RF, CS>(cf + i * LF, L);
            int t = 0;
            if (PIV) t = __ldg(pv + i);
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                if (PIV) {
#pragma unroll
                    for (int q = 1; q <= KL; ++q) {
                        if (t == q) {
                            const double tmp = v[r][i];
                            v[r][i] = v[r][i + q];
                            v[r][i + q] = tmp;
                        }
                    }
                }
#pragma unroll
                for (int q = 1; q <= KL; ++q) v[r][i + q] = fma(-L[q - 1], v[r][i], v[r][i + q]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RL; ++r)
#pragma unroll
        for (int k = 0; k < KL; ++k) {
            dl[r][k] = v[r][CH + k] - dl[r][k];
            fst[(c * KL + k) * NL + r * NLt + tx] = dl[r][k];
        }

    // ---------------------------------------------------------------- B1: local backward (zero right state)
    {
        const double* cf = F.cfB + (long long) j0 * LB;
#pragma unroll
        for (int i = CH - 1; i >= 0; --i) {
            double Ub[RB];
            ldrec<RB, CS>(cf + i * LB, Ub);
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                double acc = v[r][i];
#pragma unroll
                for (int k = KD; k >= 1; --k)
                    if (i + k < CH) acc = fma(-Ub[k - 1], v[r][i + k], acc);
                v[r][i] = acc * Ub[KD];
            }
        }
    }
    sweep_sync(nsync);

    // ---------------------------------------------------------------- S1: forward states, X_c
    double dlt[RL][KL];
    if (F.seq) {
        if (c == 0) {
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                const int ln = r * NLt + tx;
                double d[KL];
#pragma unroll
                for (int k = 0; k < KL; ++k) d[k] = 0.0;
                for (int cc = 0; cc < SC - 1; ++cc) {
                    const double* T = F.T + cc * KL * KL;
                    double nd[KL];
#pragma unroll
                    for (int k = 0; k < KL; ++k) {
                        double acc = fst[(cc * KL + k) * NL + ln];
#pragma unroll
                        for (int q = 0; q < KL; ++q) acc = fma(ldc<CS>(T + k * KL + q), d[q], acc);
                        nd[k] = acc;
                    }
#pragma unroll
                    for (int k = 0; k < KL; ++k) {
                        d[k] = nd[k];
                        fst[(cc * KL + k) * NL + ln] = nd[k];  // now delta_{cc+1}
                    }
                }
            }
        }
        sweep_sync(nsync);
#pragma unroll
        for (int r = 0; r < RL; ++r)
#pragma unroll
            for (int k = 0; k < KL; ++k) dlt[r][k] = (c > 0) ? fst[((c - 1) * KL + k) * NL + r * NLt + tx] : 0.0;
    } else {
#pragma unroll
        for (int r = 0; r < RL; ++r)
#pragma unroll
            for (int k = 0; k < KL; ++k) dlt[r][k] = (c > 0) ? fst[((c - 1) * KL + k) * NL + r * NLt + tx] : 0.0;
        for (int d = 2; d <= F.DF && c - d >= 0; ++d) {
            const double* W = F.W + ((long long) c * (MD - 1) + d - 2) * KL * KL;
            double w[KL * KL];
#pragma unroll
            for (int q = 0; q < KL * KL; ++q) w[q] = ldc<CS>(W + q);
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                double e[KL];
#pragma unroll
                for (int q = 0; q < KL; ++q) e[q] = fst[((c - d) * KL + q) * NL + r * NLt + tx];
#pragma unroll
                for (int k = 0; k < KL; ++k)
#pragma unroll
                    for (int q = 0; q < KL; ++q) dlt[r][k] = fma(w[k * KL + q], e[q], dlt[r][k]);
            }
        }
    }
    {
        const double* cf = F.cfC + (long long) j0 * LC;
#pragma unroll
        for (int i = 0; i < KD; ++i) {
            double xi[KL];
#pragma unroll
            for (int q = 0; q < KL; ++q) xi[q] = ldc<CS>(cf + i * LC + KD + q);
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                double acc = v[r][i];
#pragma unroll
                for (int q = 0; q < KL; ++q) acc = fma(xi[q], dlt[r][q], acc);
                bst[(c * KD + i) * NL + r * NLt + tx] = acc;
            }
        }
    }
    sweep_sync(nsync);

    // ---------------------------------------------------------------- S2: backward states
    double tt[RL][KD];
    if (F.seq) {
        if (c == 0) {
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                const int ln = r * NLt + tx;
                double t[KD];
#pragma unroll
                for (int k = 0; k < KD; ++k) t[k] = 0.0;
                for (int cc = SC - 1; cc >= 1; --cc) {
                    const double* Rm = F.Rm + cc * KD * KD;
                    double nt[KD];
#pragma unroll
                    for (int i = 0; i < KD; ++i) {
                        double acc = bst[(cc * KD + i) * NL + ln];
#pragma unroll
                        for (int k = 0; k < KD; ++k) acc = fma(ldc<CS>(Rm + i * KD + k), t[k], acc);
                        nt[i] = acc;
                    }
#pragma unroll
                    for (int i = 0; i < KD; ++i) {
                        t[i] = nt[i];
                        bst[(cc * KD + i) * NL + ln] = nt[i];  // now t_{cc-1}
                    }
                }
            }
        }
        sweep_sync(nsync);
#pragma unroll
        for (int r = 0; r < RL; ++r)
#pragma unroll
            for (int k = 0; k < KD; ++k) tt[r][k] = (c + 1 < SC) ? bst[((c + 1) * KD + k) * NL + r * NLt + tx] : 0.0;
    } else {
#pragma unroll
        for (int r = 0; r < RL; ++r)
#pragma unroll
            for (int k = 0; k < KD; ++k) tt[r][k] = (c + 1 < SC) ? bst[((c + 1) * KD + k) * NL + r * NLt + tx] : 0.0;
        for (int d = 2; d <= F.DB && c + d < SC; ++d) {
            const double* V = F.V + ((long long) c * (MD - 1) + d - 2) * KD * KD;
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                double e[KD];
#pragma unroll
                for (int q = 0; q < KD; ++q) e[q] = bst[((c + d) * KD + q) * NL + r * NLt + tx];
#pragma unroll
                for (int k = 0; k < KD; ++k)
#pragma unroll
                    for (int q = 0; q < KD; ++q) tt[r][k] = fma(ldc<CS>(V + k * KD + q), e[q], tt[r][k]);
            }
        }
    }

    // ---------------------------------------------------------------- B3: correct x
    {
        const double* cf = F.cfC + (long long) j0 * LC;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            double C[RC];
            ldrec<RC, CS>(cf + i * LC, C);
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                double acc = v[r][i];
#pragma unroll
                for (int k = 0; k < KD; ++k) acc = fma(C[k], tt[r][k], acc);
#pragma unroll
                for (int q = 0; q < KL; ++q) acc = fma(C[KD + q], dlt[r][q], acc);
                v[r][i] = acc;
            }
        }
    }

}

}  // namespace adsb

#endif
