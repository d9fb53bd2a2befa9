// kernels_lines.cu -- the "special dimension" of the generalised ADS: every line along one axis is solved with
// its OWN factorised band matrix (include/ads/solver.hpp:56-96,:170-195; the caller's callable there is
// examples/maxwell/maxwell_ads.hpp:139-163: one dgbtrs per line with the matrix of that line).
//
// One thread per line runs LAPACK's dgbtrs('N') recurrence (the semantics of include/ads/lin/band_solve.hpp:21-31)
// with a register window: forward elimination keeps rows j .. j+kl (the rows a pivot swap or a multiplier can
// touch), back substitution keeps the kl+ku rows above the current one.  Every coefficient and every pivot index
// is used once, so the factor table -- (2kl+ku+1) doubles + one int per DOF -- is the traffic; it is stored
// [column j][band row r][line] with the line index fastest and lanes run along the lines, so factor reads are
// coalesced for every axis.  The tensor itself is read and written in place: coalesced for the y and z axes
// (lines numbered x fastest); for the x axis a warp's 32 lines are 32 different rows, so its tiles go through
// shared memory (a 32 x 33 transpose per 32 columns).
#include "kernels.cuh"

namespace adsb {

namespace {

constexpr int LT = 32;  // lines per block along the fast line index (one warp); blockDim.y warps take more lines

// element j of line `line`: base + j * sj, where base = (line % L0) * s0 + (line / L0) * s1
template <int KL, int KU>
__global__ void __launch_bounds__(128)
    line_sweep_kernel(double* __restrict__ t, const double* __restrict__ ab, const int* __restrict__ ipiv, int n, long long lines,
                      int L0, long long s0, long long s1, long long sj) {
    constexpr int KD = KL + KU, LD = 2 * KL + KU + 1;
    const long long line = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= lines) return;
    double* const b = t + (line % L0) * s0 + (line / L0) * s1;
    const double* const A = ab + line;   // A[(j * LD + r) * lines]
    const int* const pv = ipiv + line;   // pv[j * lines], 0-based row index

    // ---- L y = P b
    double w[KL + 1];
#pragma unroll
    for (int i = 0; i <= KL; ++i) w[i] = i < n ? b[i * sj] : 0.0;
    for (int j = 0; j < n; ++j) {
        if (j < n - 1) {
            const int l = pv[(long long) j * lines] - j;  // 0 .. KL
#pragma unroll
            for (int i = 1; i <= KL; ++i)
                if (i == l) {
                    const double tmp = w[0];
                    w[0] = w[i];
                    w[i] = tmp;
                }
#pragma unroll
            for (int i = 1; i <= KL; ++i)
                if (j + i < n) w[i] -= w[0] * A[((long long) j * LD + KD + i) * lines];
        }
        b[j * sj] = w[0];
#pragma unroll
        for (int i = 0; i < KL; ++i) w[i] = w[i + 1];
        w[KL] = (j + KL + 1 < n) ? b[(long long) (j + KL + 1) * sj] : 0.0;
    }
    // ---- U x = y, column oriented (dtbsv 'U','N','N'): x_j = y_j / U(j,j); y_{j-k} -= x_j U(j-k, j)
    double v[KD + 1];
#pragma unroll
    for (int k = 0; k <= KD; ++k) v[k] = (n - 1 - k >= 0) ? b[(long long) (n - 1 - k) * sj] : 0.0;
    for (int j = n - 1; j >= 0; --j) {
        const double x = v[0] / A[((long long) j * LD + KD) * lines];
#pragma unroll
        for (int k = 1; k <= KD; ++k)
            if (j - k >= 0) v[k] -= x * A[((long long) j * LD + KD - k) * lines];
        b[j * sj] = x;
#pragma unroll
        for (int k = 0; k < KD; ++k) v[k] = v[k + 1];
        v[KD] = (j - KD - 1 >= 0) ? b[(long long) (j - KD - 1) * sj] : 0.0;
    }
}

// x axis: the n elements of a line are contiguous and the lines are `pitch` apart.  A warp owns 32 lines; the
// tile [32 lines][32 columns] moves between global and shared memory with the lanes along the columns, the
// recurrence reads it with the lanes along the lines (pitch 33: no bank conflicts).
template <int KL, int KU>
__global__ void __launch_bounds__(32)
    line_sweep_x_kernel(double* __restrict__ t, const double* __restrict__ ab, const int* __restrict__ ipiv, int n, long long lines,
                        long long pitch, double* __restrict__ scratch) {
    // the whole line set of the warp lives in a global scratch slab [n][32] (line fastest) between the passes;
    // shared memory only transposes
    constexpr int KD = KL + KU, LD = 2 * KL + KU + 1;
    __shared__ double tile[LT][LT + 1];
    const int lane = threadIdx.x;
    const long long line0 = (long long) blockIdx.x * LT;
    const long long line = line0 + lane;
    const bool live = line < lines;
    double* const slab = scratch + (long long) blockIdx.x * n * LT;  // [j][lane]
    // gather: rows line0 .. line0+31, 32 columns at a time
    for (int c0 = 0; c0 < n; c0 += LT) {
        for (int r = 0; r < LT; ++r) {
            const long long ln = line0 + r;
            tile[r][lane] = (ln < lines && c0 + lane < n) ? t[ln * pitch + c0 + lane] : 0.0;
        }
        __syncwarp();
        for (int c = 0; c < LT && c0 + c < n; ++c) slab[(long long) (c0 + c) * LT + lane] = tile[lane][c];
        __syncwarp();
    }
    if (live) {
        const double* const A = ab + line;
        const int* const pv = ipiv + line;
        double* const b = slab + lane;
        const long long sj = LT;
        double w[KL + 1];
#pragma unroll
        for (int i = 0; i <= KL; ++i) w[i] = i < n ? b[i * sj] : 0.0;
        for (int j = 0; j < n; ++j) {
            if (j < n - 1) {
                const int l = pv[(long long) j * lines] - j;
#pragma unroll
                for (int i = 1; i <= KL; ++i)
                    if (i == l) {
                        const double tmp = w[0];
                        w[0] = w[i];
                        w[i] = tmp;
                    }
#pragma unroll
                for (int i = 1; i <= KL; ++i)
                    if (j + i < n) w[i] -= w[0] * A[((long long) j * LD + KD + i) * lines];
            }
            b[j * sj] = w[0];
#pragma unroll
            for (int i = 0; i < KL; ++i) w[i] = w[i + 1];
            w[KL] = (j + KL + 1 < n) ? b[(long long) (j + KL + 1) * sj] : 0.0;
        }
        double v[KD + 1];
#pragma unroll
        for (int k = 0; k <= KD; ++k) v[k] = (n - 1 - k >= 0) ? b[(long long) (n - 1 - k) * sj] : 0.0;
        for (int j = n - 1; j >= 0; --j) {
            const double x = v[0] / A[((long long) j * LD + KD) * lines];
#pragma unroll
            for (int k = 1; k <= KD; ++k)
                if (j - k >= 0) v[k] -= x * A[((long long) j * LD + KD - k) * lines];
            b[j * sj] = x;
#pragma unroll
            for (int k = 0; k < KD; ++k) v[k] = v[k + 1];
            v[KD] = (j - KD - 1 >= 0) ? b[(long long) (j - KD - 1) * sj] : 0.0;
        }
    }
    __syncwarp();
    // scatter back
    for (int c0 = 0; c0 < n; c0 += LT) {
        for (int c = 0; c < LT && c0 + c < n; ++c) tile[lane][c] = slab[(long long) (c0 + c) * LT + lane];
        __syncwarp();
        for (int r = 0; r < LT; ++r) {
            const long long ln = line0 + r;
            if (ln < lines && c0 + lane < n) t[ln * pitch + c0 + lane] = tile[r][lane];
        }
        __syncwarp();
    }
}

template <int P>
int launch_p(double* t, const double* ab, const int* ipiv, int n, long long lines, int axis, int L0, long long s0,
             long long s1, long long sj, double* scratch, cudaStream_t st) {
    if (axis == 0) {
        const unsigned blocks = (unsigned) ((lines + LT - 1) / LT);
        line_sweep_x_kernel<P, P><<<blocks, LT, 0, st>>>(t, ab, ipiv, n, lines, s0, scratch);
    } else {
        const unsigned blocks = (unsigned) ((lines + 127) / 128);
        line_sweep_kernel<P, P><<<blocks, 128, 0, st>>>(t, ab, ipiv, n, lines, L0, s0, s1, sj);
    }
    return (int) cudaGetLastError();
}

}  // namespace

// In-place solve of `lines` lines of n elements, line `l` with the factor ab[(j*LD + r)*lines + l] (LAPACK band
// storage, LD = 2kl+ku+1, kl = ku = p) and 0-based pivot rows ipiv[j*lines + l].  axis 0: lines are contiguous
// rows, s0 apart (scratch: ceil(lines/32)*32*n doubles); otherwise element j of line l sits at
// (l % L0)*s0 + (l / L0)*s1 + j*sj.  Returns a cudaError_t as int.
int launch_line_sweep(int p, double* t, const double* ab, const int* ipiv, int n, long long lines, int axis, int L0,
                      long long s0, long long s1, long long sj, double* scratch, cudaStream_t st) {
    switch (p) {
    case 1: return launch_p<1>(t, ab, ipiv, n, lines, axis, L0, s0, s1, sj, scratch, st);
    case 2: return launch_p<2>(t, ab, ipiv, n, lines, axis, L0, s0, s1, sj, scratch, st);
    case 3: return launch_p<3>(t, ab, ipiv, n, lines, axis, L0, s0, s1, sj, scratch, st);
    case 4: return launch_p<4>(t, ab, ipiv, n, lines, axis, L0, s0, s1, sj, scratch, st);
    case 5: return launch_p<5>(t, ab, ipiv, n, lines, axis, L0, s0, s1, sj, scratch, st);
    default: return (int) cudaErrorInvalidValue;
    }
}

}  // namespace adsb
