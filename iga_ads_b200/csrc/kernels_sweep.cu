// kernels_sweep.cu -- K2: batched banded forward/back substitution along one axis, FP64, sm_100a.
//
// Replaces lin::solve_with_factorized -> dgbtrs_('N') (include/ads/lin/band_solve.hpp:21-31) and
// the lin::cyclic_transpose that follows it in ads_solve (include/ads/solver.hpp:35-41,
// include/ads/lin/tensor/cyclic_transpose.hpp:54-63).  The tensor is never rotated in HBM: every
// sweep reads and writes the canonical (x fastest) layout, 16 B per DOF.
//
// Parallel scheme.  A CTA owns NL lines; each line of n unknowns is cut into SC chunks of CH
// columns; one thread owns R adjacent chunks of one line and keeps them in registers (R independent
// recurrences per thread hide the FP64 / table-load latency).  All lines share one factor, so the
// response of a chunk to its incoming recurrence states is tabulated once on the host
// (build_sweep_plan, host_setup.cpp).  Per thread:
//   F1  pivoted unit-lower forward recurrence over each chunk from the ORIGINAL data    (KL fma/col)
//   B1  upper back substitution over each chunk with zero state on the right          (KD fma + mul)
//   --  publish the forward out-states Delta_c                                          (barrier)
//   S1  delta_c = sum_d W_{c,d} Delta_{c-d};  X_c = xfirst_c + Xi_c delta_c             (barrier)
//   S2  t_c = sum_d V_{c,d} X_{c+d}
//   B3  x = x_local + Psi t_c + Xi delta_c; store                                       (KD+KL fma/col)
// which is algebraically the sequential dgbtrs recurrence (same factor, same pivots); only the
// association of a few additions differs (last-bit level, checked by the parity tests).  When the
// state responses do not decay fast enough for the finite-depth sums (plan.seq), S1/S2 chain the
// states sequentially instead (one thread per line).
//
// Memory.  STRIDED sweeps (y, z): lanes run along x, so every global access of a half-warp is a
// 128 B row segment; chunks go straight from HBM to registers and back.  CONTIG sweep (x): the NL
// lines of a CTA are staged through shared memory -- moved by the TMA engine as bulk copies
// (cp.async.bulk + mbarrier) when rows are 16 B aligned -- so the global side is read and written in
// whole contiguous rows while each thread picks its chunk out of shared memory.
#include <cstdint>

#include "kernels.cuh"

namespace adsb {

namespace {

template <int N>
__device__ __forceinline__ void ldrec(const double* __restrict__ p, double* o) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p + k));
        o[k] = t.x;
        o[k + 1] = t.y;
    }
    if (N & 1) o[N - 1] = __ldg(p + N - 1);
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes)
                 : "memory");
}

// RL lines per thread share every coefficient load; NLt = blockDim.x lanes; a CTA owns NLt*RL lines.
template <int KL, int KD, bool PIV, int CH, int RL, bool CONTIG>
__global__ void __launch_bounds__(512, 1) sweep_kernel(const SweepFactor F, const SweepGeom G) {
    extern __shared__ __align__(16) double smem[];
    constexpr int LF = KL + (KL & 1), LB = (KD + 1) + ((KD + 1) & 1), LC = (KD + KL) + ((KD + KL) & 1);
    constexpr int MD = SWEEP_MAX_DEPTH_DEV;
    const int NLt = blockDim.x, SC = blockDim.y, NL = NLt * RL;
    const int tx = threadIdx.x, c = threadIdx.y;  // c: chunk of this thread
    const int tid = c * NLt + tx, nthr = NLt * SC;
    const int lbase = blockIdx.x * NL;
    const int m = blockIdx.y;
    const int n = F.n;
    const int j0 = c * CH;
    const bool full = (j0 + CH + KL <= n) && (lbase + NL <= G.L0);
    bool act[RL];
#pragma unroll
    for (int r = 0; r < RL; ++r) act[r] = lbase + r * NLt + tx < G.L0;

    double* fst = smem;                   // [SC][KL][NL] forward states
    double* bst = fst + SC * KL * NL;     // [SC][KD][NL] backward states
    double* tile = smem + ((SC * (KL + KD) * NL + 1) & ~1);  // CONTIG only: [NL][pitch], 16 B aligned
    __shared__ uint64_t bar;

    double v[RL][CH + KL];

    // ---------------------------------------------------------------- load
    if (CONTIG) {
        const int lines = min(NL, G.L0 - lbase);
        const double* src = G.in + (long long) lbase * G.s0_in + (long long) m * G.s1_in;
        if (G.bulk) {
            if (tid == 0) mbar_init(&bar, 1);
            __syncthreads();
            if (tid < 32) {
                if (tid == 0) mbar_expect_tx(&bar, (uint32_t) (lines * n * 8));
                __syncwarp();
                for (int ln = tid; ln < lines; ln += 32)
                    bulk_g2s(tile + ln * G.pitch, src + ln * G.s0_in, (uint32_t) (n * 8), &bar);
            }
            mbar_wait(&bar, 0);
        } else {
            for (int ln = 0; ln < lines; ++ln) {
                const double* row = src + ln * G.s0_in;
                double* dst = tile + ln * G.pitch;
#pragma unroll 4
                for (int j = tid; j < n; j += nthr) dst[j] = __ldcs(row + j);
            }
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < RL; ++r) {
            const double* mine = tile + (r * NLt + tx) * G.pitch + j0;
            if (CH % 2 == 0 && KL % 2 == 0) {  // 16 B aligned: pitch and j0 are even
#pragma unroll
                for (int i = 0; i < CH + KL; i += 2) {
                    double2 t = make_double2(0.0, 0.0);
                    if (full || (act[r] && j0 + i < n)) t = *reinterpret_cast<const double2*>(mine + i);
                    v[r][i] = t.x;
                    v[r][i + 1] = (full || j0 + i + 1 < n) ? t.y : 0.0;
                }
            } else {
#pragma unroll
                for (int i = 0; i < CH + KL; ++i) v[r][i] = (full || (act[r] && j0 + i < n)) ? mine[i] : 0.0;
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < RL; ++r) {
            const double* src = G.in + (long long) (lbase + r * NLt + tx) * G.s0_in + (long long) m * G.s1_in;
            if (G.off_in) {
#pragma unroll
                for (int i = 0; i < CH + KL; ++i) {
                    const int j = j0 + i;
                    v[r][i] = (act[r] && j < n) ? __ldcs(src + G.off_in[j]) : 0.0;
                }
            } else if (full) {
                const double* p = src + (long long) j0 * G.sj_in;
#pragma unroll
                for (int i = 0; i < CH + KL; ++i) {
                    v[r][i] = __ldcs(p);
                    p += G.sj_in;
                }
            } else {
#pragma unroll
                for (int i = 0; i < CH + KL; ++i) {
                    const int j = j0 + i;
                    v[r][i] = (act[r] && j < n) ? __ldcs(src + j * G.sj_in) : 0.0;
                }
            }
        }
    }

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
            double L[LF];
            ldrec<LF>(cf + i * LF, L);
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
            double Ub[LB];
            ldrec<LB>(cf + i * LB, Ub);
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
    __syncthreads();

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
                        for (int q = 0; q < KL; ++q) acc = fma(__ldg(T + k * KL + q), d[q], acc);
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
        __syncthreads();
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
            for (int q = 0; q < KL * KL; ++q) w[q] = __ldg(W + q);
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
            for (int q = 0; q < KL; ++q) xi[q] = __ldg(cf + i * LC + KD + q);
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                double acc = v[r][i];
#pragma unroll
                for (int q = 0; q < KL; ++q) acc = fma(xi[q], dlt[r][q], acc);
                bst[(c * KD + i) * NL + r * NLt + tx] = acc;
            }
        }
    }
    __syncthreads();

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
                        for (int k = 0; k < KD; ++k) acc = fma(__ldg(Rm + i * KD + k), t[k], acc);
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
        __syncthreads();
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
                    for (int q = 0; q < KD; ++q) tt[r][k] = fma(__ldg(V + k * KD + q), e[q], tt[r][k]);
            }
        }
    }

    // ---------------------------------------------------------------- B3: correct x
    {
        const double* cf = F.cfC + (long long) j0 * LC;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            double C[LC];
            ldrec<LC>(cf + i * LC, C);
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

    // ---------------------------------------------------------------- store
    if (CONTIG) {
#pragma unroll
        for (int r = 0; r < RL; ++r) {
            double* mine = tile + (r * NLt + tx) * G.pitch + j0;
            if (CH % 2 == 0) {
#pragma unroll
                for (int i = 0; i < CH; i += 2) {
                    if (full || (act[r] && j0 + i + 1 < n))
                        *reinterpret_cast<double2*>(mine + i) = make_double2(v[r][i], v[r][i + 1]);
                    else if (act[r] && j0 + i < n)
                        mine[i] = v[r][i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < CH; ++i)
                    if (full || (act[r] && j0 + i < n)) mine[i] = v[r][i];
            }
        }
        const int lines = min(NL, G.L0 - lbase);
        double* dstb = G.out + (long long) lbase * G.s0_out + (long long) m * G.s1_out;
        if (G.bulk) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid < 32) {
                for (int ln = tid; ln < lines; ln += 32)
                    bulk_s2g(dstb + ln * G.s0_out, tile + ln * G.pitch, (uint32_t) (n * 8));
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else {
            __syncthreads();
            for (int ln = 0; ln < lines; ++ln) {
                double* row = dstb + ln * G.s0_out;
                const double* srow = tile + ln * G.pitch;
#pragma unroll 4
                for (int j = tid; j < n; j += nthr) __stcs(row + j, srow[j]);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < RL; ++r) {
            double* dst = G.out + (long long) (lbase + r * NLt + tx) * G.s0_out + (long long) m * G.s1_out;
            if (G.off_out) {
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int j = j0 + i;
                    if (act[r] && j < n) __stcs(dst + G.off_out[j], v[r][i]);
                }
            } else if (full) {
                double* p = dst + (long long) j0 * G.sj_out;
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    __stcs(p, v[r][i]);
                    p += G.sj_out;
                }
            } else {
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int j = j0 + i;
                    if (act[r] && j < n) __stcs(dst + j * G.sj_out, v[r][i]);
                }
            }
        }
    }
}

using kern_t = void (*)(const SweepFactor, const SweepGeom);

template <int P, bool PIV>
kern_t pick_mode(bool contig) {
    constexpr int KD = PIV ? 2 * P : P;
    return contig ? (kern_t) sweep_kernel<P, KD, PIV, SWEEP_CH, SWEEP_RL, true>
                  : (kern_t) sweep_kernel<P, KD, PIV, SWEEP_CH, SWEEP_RL, false>;
}

kern_t pick(int KL, bool piv, bool contig) {
    switch (KL) {
    case 1: return piv ? pick_mode<1, true>(contig) : pick_mode<1, false>(contig);
    case 2: return piv ? pick_mode<2, true>(contig) : pick_mode<2, false>(contig);
    case 3: return piv ? pick_mode<3, true>(contig) : pick_mode<3, false>(contig);
    case 4: return piv ? pick_mode<4, true>(contig) : pick_mode<4, false>(contig);
    case 5: return piv ? pick_mode<5, true>(contig) : pick_mode<5, false>(contig);
    default: return nullptr;
    }
}

}  // namespace

int sweep_smem_bytes(const SweepFactor& F, bool contig, int NLt, int pitch) {
    const int NL = NLt * SWEEP_RL;
    long long d = (long long) F.SC * (F.KL + F.KD) * NL + (contig ? (long long) NL * pitch + 4 : 0);
    return (int) (d * sizeof(double));
}

int launch_sweep(const SweepFactor& F, const SweepGeom& G, bool contig, int NLt, cudaStream_t st) {
    kern_t k = pick(F.KL, F.piv != 0, contig);
    if (!k) return (int) cudaErrorInvalidValue;
    if (NLt * F.SC > 512) return (int) cudaErrorInvalidConfiguration;
    const int smem = sweep_smem_bytes(F, contig, NLt, G.pitch);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute((const void*) k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int) e;
    }
    dim3 block(NLt, F.SC, 1);
    const int NL = NLt * SWEEP_RL;
    dim3 grid((G.L0 + NL - 1) / NL, G.L1, 1);
    k<<<grid, block, smem, st>>>(F, G);
    return (int) cudaGetLastError();
}

}  // namespace adsb
