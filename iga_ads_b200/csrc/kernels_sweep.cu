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
#include "sweep_core.cuh"

namespace adsb {

namespace {

// RL lines per thread share every coefficient load; NLt = blockDim.x lanes; a CTA owns NLt*RL lines.
template <int KL, int KD, bool PIV, int CH, int RL, bool CONTIG>
__global__ void __launch_bounds__(512, 1) sweep_kernel(const SweepFactor F, const SweepGeom G) {
    extern __shared__ __align__(16) double smem[];
    constexpr int LF = sweep_pitch(KL), LB = sweep_pitch(KD + 1), LC = sweep_pitch(KD + KL);
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
            // one flat loop over (line, column): every thread has many independent loads in flight
            const int total = lines * n;
#pragma unroll 8
            for (int idx = tid; idx < total; idx += nthr) {
                const int ln = idx / n, j = idx - ln * n;
                tile[ln * G.pitch + j] = __ldcs(src + ln * G.s0_in + j);
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

    sweep_core<KL, KD, PIV, CH, RL>(F, v, fst, bst, c, tx, NLt, SC);

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
            const int total = lines * n;
#pragma unroll 8
            for (int idx = tid; idx < total; idx += nthr) {
                const int ln = idx / n, j = idx - ln * n;
                __stcs(dstb + ln * G.s0_out + j, tile[ln * G.pitch + j]);
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
