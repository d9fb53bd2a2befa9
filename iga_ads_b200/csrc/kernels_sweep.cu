// kernels_sweep.cu -- K2: batched banded forward/back substitution along one axis, FP64, sm_100a.
//
// Replaces lin::solve_with_factorized -> dgbtrs_('N') (include/ads/lin/band_solve.hpp:21-31) and
// the lin::cyclic_transpose that follows it in ads_solve (include/ads/solver.hpp:35-41,
// include/ads/lin/tensor/cyclic_transpose.hpp:54-63).  The tensor is never rotated in HBM: every
// sweep reads and writes the canonical (x fastest) layout, 16 B per DOF.
//
// Parallel scheme.  A CTA owns NL lines; each line of n unknowns is cut into S chunks of CH
// columns, one thread per (line, chunk), the chunk held in registers.  All lines share one factor,
// so the response of a chunk to its incoming recurrence state is a table built once on the host
// (build_sweep_plan).  Per thread:
//   F1  pivoted unit-lower forward recurrence over its chunk from the ORIGINAL data     (KL fma/col)
//   F2  one thread per line chains the KL-wide partial-update states across chunks      (S steps)
//   B1  y = y_local + Phi*state; upper back substitution over its chunk, zero on the right (KL+KD fma)
//   B2  one thread per line chains the KD-wide states right to left                     (S steps)
//   B3  x = x_local + Psi*state; store                                                  (KD fma/col)
// which is algebraically the sequential dgbtrs recurrence (same factor, same pivots); only the
// association of a few additions differs (last-bit level, checked by the parity tests).
//
// Memory.  STRIDED sweeps (y, z): lanes run along x, so every global access of a half-warp is a
// 128 B row segment; the chunk goes straight from HBM to registers and back.  CONTIG sweep (x):
// the NL lines of a CTA are staged through shared memory with an odd pitch so that the global side
// is read and written in full contiguous rows and the per-thread side is bank-conflict free.
#include <cstdio>

#include "kernels.cuh"

namespace adsb {

namespace {

template <int KL, int KD, bool PIV, int CH, bool CONTIG>
__global__ void __launch_bounds__(512, 1) sweep_kernel(const SweepFactor F, const SweepGeom G) {
    extern __shared__ double smem[];
    const int NL = blockDim.x, S = blockDim.y;
    const int tx = threadIdx.x, s = threadIdx.y;
    const int l = blockIdx.x * NL + tx;
    const int m = blockIdx.y;
    const bool active = l < G.L0;
    const int n = F.n;
    const int j0 = s * CH;

    double* fst = smem;                  // [S][KL][NL] forward states
    double* bst = fst + S * KL * NL;     // [S][KD][NL] backward states
    double* tile = bst + S * KD * NL;    // CONTIG only: [NL][pitch]

    double v[CH + KL];

    // ---------------------------------------------------------------- load
    if (CONTIG) {
        const int tid = s * NL + tx, nthr = NL * S;
        const int lines = min(NL, G.L0 - blockIdx.x * NL);
        const double* src = G.in + (long long) blockIdx.x * NL * G.s0_in + (long long) m * G.s1_in;
        for (int ln = 0; ln < lines; ++ln) {
            const double* row = src + ln * G.s0_in;
            double* dst = tile + ln * G.pitch;
            for (int j = tid; j < n; j += nthr) dst[j] = row[j];
        }
        __syncthreads();
        const double* mine = tile + tx * G.pitch;
#pragma unroll
        for (int i = 0; i < CH + KL; ++i) {
            const int j = j0 + i;
            v[i] = (active && j < n) ? mine[j] : 0.0;
        }
    } else {
        const double* src = G.in + (long long) l * G.s0_in + (long long) m * G.s1_in;
        if (G.off_in) {
#pragma unroll
            for (int i = 0; i < CH + KL; ++i) {
                const int j = j0 + i;
                v[i] = (active && j < n) ? src[G.off_in[j]] : 0.0;
            }
        } else {
#pragma unroll
            for (int i = 0; i < CH + KL; ++i) {
                const int j = j0 + i;
                v[i] = (active && j < n) ? src[j * G.sj_in] : 0.0;
            }
        }
    }

    // ---------------------------------------------------------------- F1: local forward
    double o[KL];
#pragma unroll
    for (int r = 0; r < KL; ++r) o[r] = v[CH + r];
    {
        const double* Lm = F.Lm + (long long) j0 * KL;
        const int* pv = F.pv + j0;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (PIV) {
                const int t = pv[i];
#pragma unroll
                for (int r = 1; r <= KL; ++r) {
                    if (t == r) {
                        const double tmp = v[i];
                        v[i] = v[i + r];
                        v[i + r] = tmp;
                    }
                }
            }
#pragma unroll
            for (int r = 1; r <= KL; ++r) v[i + r] = fma(-Lm[i * KL + r - 1], v[i], v[i + r]);
        }
    }
    if (s < S - 1) {
#pragma unroll
        for (int r = 0; r < KL; ++r) fst[(s * KL + r) * NL + tx] = v[CH + r] - o[r];
    }
    __syncthreads();

    // ---------------------------------------------------------------- F2: chain forward states
    if (s == 0) {
        double d[KL];
#pragma unroll
        for (int r = 0; r < KL; ++r) d[r] = 0.0;
        for (int sp = 0; sp < S - 1; ++sp) {
            const double* T = F.T + sp * KL * KL;
            double nd[KL];
#pragma unroll
            for (int r = 0; r < KL; ++r) {
                double acc = fst[(sp * KL + r) * NL + tx];
#pragma unroll
                for (int c = 0; c < KL; ++c) acc = fma(T[r * KL + c], d[c], acc);
                nd[r] = acc;
            }
#pragma unroll
            for (int r = 0; r < KL; ++r) {
                d[r] = nd[r];
                fst[(sp * KL + r) * NL + tx] = nd[r];  // now the incoming state of chunk sp+1
            }
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- B1: correct y, local backward
    {
        double d[KL];
#pragma unroll
        for (int r = 0; r < KL; ++r) d[r] = (s > 0) ? fst[((s - 1) * KL + r) * NL + tx] : 0.0;
        const double* Phi = F.Phi + (long long) j0 * KL;
        const double* Ut = F.Ut + (long long) j0 * KD;
        const double* rinv = F.rinv + j0;
#pragma unroll
        for (int i = CH - 1; i >= 0; --i) {
            double acc = v[i];
#pragma unroll
            for (int r = 0; r < KL; ++r) acc = fma(Phi[i * KL + r], d[r], acc);
#pragma unroll
            for (int k = KD; k >= 1; --k)
                if (i + k < CH) acc = fma(-Ut[i * KD + k - 1], v[i + k], acc);
            v[i] = acc * rinv[i];
        }
    }
    if (s > 0) {
#pragma unroll
        for (int k = 0; k < KD; ++k) bst[(s * KD + k) * NL + tx] = v[k];
    }
    __syncthreads();

    // ---------------------------------------------------------------- B2: chain backward states
    if (s == 0) {
        double t[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) t[k] = 0.0;
        for (int sp = S - 1; sp >= 1; --sp) {
            const double* Psi = F.Psi + (long long) sp * CH * KD;
            double nt[KD];
#pragma unroll
            for (int i = 0; i < KD; ++i) {
                double acc = bst[(sp * KD + i) * NL + tx];
#pragma unroll
                for (int k = 0; k < KD; ++k) acc = fma(Psi[i * KD + k], t[k], acc);
                nt[i] = acc;
            }
#pragma unroll
            for (int i = 0; i < KD; ++i) {
                t[i] = nt[i];
                bst[(sp * KD + i) * NL + tx] = nt[i];  // now the incoming state of chunk sp-1
            }
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- B3: correct x, store
    {
        double t[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) t[k] = (s < S - 1) ? bst[((s + 1) * KD + k) * NL + tx] : 0.0;
        const double* Psi = F.Psi + (long long) j0 * KD;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            double acc = v[i];
#pragma unroll
            for (int k = 0; k < KD; ++k) acc = fma(Psi[i * KD + k], t[k], acc);
            v[i] = acc;
        }
    }
    if (CONTIG) {
        double* mine = tile + tx * G.pitch;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int j = j0 + i;
            if (active && j < n) mine[j] = v[i];
        }
        __syncthreads();
        const int tid = s * NL + tx, nthr = NL * S;
        const int lines = min(NL, G.L0 - blockIdx.x * NL);
        double* dstb = G.out + (long long) blockIdx.x * NL * G.s0_out + (long long) m * G.s1_out;
        for (int ln = 0; ln < lines; ++ln) {
            double* row = dstb + ln * G.s0_out;
            const double* srow = tile + ln * G.pitch;
            for (int j = tid; j < n; j += nthr) row[j] = srow[j];
        }
    } else {
        double* dst = G.out + (long long) l * G.s0_out + (long long) m * G.s1_out;
        if (G.off_out) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int j = j0 + i;
                if (active && j < n) dst[G.off_out[j]] = v[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int j = j0 + i;
                if (active && j < n) dst[j * G.sj_out] = v[i];
            }
        }
    }
}

using kern_t = void (*)(const SweepFactor, const SweepGeom);

template <int P, bool PIV>
kern_t pick_mode(bool contig) {
    constexpr int KD = PIV ? 2 * P : P;
    return contig ? (kern_t) sweep_kernel<P, KD, PIV, SWEEP_CH, true>
                  : (kern_t) sweep_kernel<P, KD, PIV, SWEEP_CH, false>;
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

int sweep_smem_bytes(const SweepFactor& F, bool contig, int NL, int pitch) {
    long long d = (long long) F.S * (F.KL + F.KD) * NL + (contig ? (long long) NL * pitch : 0);
    return (int) (d * sizeof(double));
}

int launch_sweep(const SweepFactor& F, const SweepGeom& G, bool contig, int NL, cudaStream_t st) {
    kern_t k = pick(F.KL, F.piv != 0, contig);
    if (!k) return (int) cudaErrorInvalidValue;
    if (NL * F.S > 512) return (int) cudaErrorInvalidConfiguration;
    const int smem = sweep_smem_bytes(F, contig, NL, G.pitch);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute((const void*) k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int) e;
    }
    dim3 block(NL, F.S, 1);
    dim3 grid((G.L0 + NL - 1) / NL, G.L1, 1);
    k<<<grid, block, smem, st>>>(F, G);
    return (int) cudaGetLastError();
}

}  // namespace adsb
