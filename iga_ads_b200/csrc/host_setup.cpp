// host_setup.cpp -- the once-per-simulation host work of the ADS step, C++17, no CUDA.
//
// Written from scratch; numerically it follows the reference algorithms step for step so the
// tables and factors that get uploaded are the ones the reference computes:
//   Gauss rule            include/ads/quad/gauss.hpp:14-15 (hard-coded 20-digit literals there;
//                         recomputed here in long double and rounded)
//   knot vector / spans   src/ads/bspline/bspline.cpp:26-43, :61-81
//   basis + derivatives   src/ads/bspline/bspline.cpp:102-160 (NURBS book A2.3)
//   quadrature tables     src/ads/basis_data.cpp:63-114
//   1-D matrices          src/ads/form_matrix.cpp:8-60, examples/implicit/implicit.hpp:46-64
//   fix_dof               src/ads/simulation/dimension.cpp:23-29
//   band LU               include/ads/lin/band_solve.hpp:16-18 -> LAPACK dgbtrf_ (unblocked DGBTF2
//                         is what LAPACK runs for these bandwidths)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "adsb200.h"
#include "internal.hpp"
#include "kernels.cuh"

namespace adsb {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

namespace {

inline double mix(double t, double a, double b) { return (1 - t) * a + t * b; }
inline double mix(int i, int n, double a, double b) {
    return mix(static_cast<double>(i) / static_cast<double>(n), a, b);
}

// Legendre P_q and its derivative at t
void legendre(int q, long double t, long double& val, long double& der) {
    long double pm = 1, pc = t;
    for (int k = 2; k <= q; ++k) {
        long double pn = ((2 * k - 1) * t * pc - (k - 1) * pm) / k;
        pm = pc;
        pc = pn;
    }
    val = pc;
    der = q * (t * pc - pm) / (t * t - 1);
}

}  // namespace

int gauss_rule(int q, double* x, double* w) {
    if (q < 2 || q > 64) return fail(ADSB_EINVAL, "gauss: q must be in 2..64");
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int i = 0; i < q; ++i) {
        long double t = cosl(pi * (i + 0.75L) / (q + 0.5L));  // i-th root from the right
        long double v, d;
        for (int it = 0; it < 100; ++it) {
            legendre(q, t, v, d);
            long double step = v / d;
            t -= step;
            if (fabsl(step) < 1e-21L) break;
        }
        if (2 * i + 1 == q) t = 0;
        legendre(q, t, v, d);
        x[q - 1 - i] = static_cast<double>(t);
        w[q - 1 - i] = static_cast<double>(2 / ((1 - t * t) * d * d));
    }
    return ADSB_OK;
}

int make_knots(int p, int elements, double a, double b, double* knot) {
    if (p < 1 || elements < 1) return fail(ADSB_EINVAL, "knots: need p >= 1, elements >= 1");
    const int size = elements + 2 * p + 1;
    for (int i = 0; i <= p; ++i) {
        knot[i] = a;
        knot[size - 1 - i] = b;
    }
    for (int i = 1; i < elements; ++i) knot[p + i] = mix(i, elements, a, b);
    return size;
}

int find_span(double x, const double* knot, int knot_size, int p) {
    int lo = p, hi = knot_size - p - 1;
    if (x >= knot[hi]) return hi - 1;
    if (x <= knot[lo]) return lo;
    int mid = (lo + hi) / 2;
    while (x < knot[mid] || x >= knot[mid + 1]) {
        (x < knot[mid] ? hi : lo) = mid;
        mid = (lo + hi) / 2;
    }
    return mid;
}

// out[d*(p+1) + r], d = 0..ders
void basis_ders(int span, double x, const double* knot, int p, int ders, double* out) {
    constexpr int M = ADSB_MAX_P + 2;
    double ndu[M][M], a[2][M], left[M], right[M];
    ndu[0][0] = 1;
    for (int j = 1; j <= p; ++j) {
        left[j] = x - knot[span + 1 - j];
        right[j] = knot[span + j] - x;
        double saved = 0;
        for (int r = 0; r < j; ++r) {
            ndu[j][r] = right[r + 1] + left[j - r];
            double tmp = ndu[r][j - 1] / ndu[j][r];
            ndu[r][j] = saved + right[r + 1] * tmp;
            saved = left[j - r] * tmp;
        }
        ndu[j][j] = saved;
    }
    for (int j = 0; j <= p; ++j) out[j] = ndu[j][p];
    for (int r = 0; r <= p; ++r) {
        int s1 = 0, s2 = 1;
        a[0][0] = 1;
        for (int k = 1; k <= ders; ++k) {
            double d = 0;
            const int rk = r - k, pk = p - k;
            if (r >= k) {
                a[s2][0] = a[s1][0] / ndu[pk + 1][rk];
                d = a[s2][0] * ndu[rk][pk];
            }
            const int j1 = rk >= -1 ? 1 : -rk;
            const int j2 = r - 1 <= pk ? k - 1 : p - r;
            for (int j = j1; j <= j2; ++j) {
                a[s2][j] = (a[s1][j] - a[s1][j - 1]) / ndu[pk + 1][rk + j];
                d += a[s2][j] * ndu[rk + j][pk];
            }
            if (r <= pk) {
                a[s2][k] = -a[s1][k - 1] / ndu[pk + 1][r];
                d += a[s2][k] * ndu[r][pk];
            }
            out[k * (p + 1) + r] = d;
            std::swap(s1, s2);
        }
    }
    int f = p;
    for (int k = 1; k <= ders; ++k) {
        for (int j = 0; j <= p; ++j) out[k * (p + 1) + j] *= f;
        f *= (p - k);
    }
}

int basis_tables(int p, int elements, double a, double b, int q, int ders, double* bt, double* xq,
                 double* w, double* J, int* first_dof) {
    if (p < 1 || p > ADSB_MAX_P) return fail(ADSB_EINVAL, "basis_tables: p out of range");
    if (ders < 0 || ders > p) return fail(ADSB_EINVAL, "basis_tables: ders out of range");
    if (elements < 1) return fail(ADSB_EINVAL, "basis_tables: elements < 1");
    std::vector<double> knot(elements + 2 * p + 1), gx(q), gw(q);
    int ks = make_knots(p, elements, a, b, knot.data());
    if (ks < 0) return ks;
    if (int rc = gauss_rule(q, gx.data(), gw.data())) return rc;
    for (int k = 0; k < q; ++k) w[k] = gw[k];
    // element end points are the distinct knots (elem_division = 1 => lerp(k, 1, x1, x2))
    for (int e = 0; e < elements; ++e) {
        // elem_division = 1: lerp(0, 1, x1, x2) == x1 and lerp(1, 1, x1, x2) == x2 exactly
        const double x1 = knot[p + e], x2 = knot[p + e + 1];
        J[e] = 0.5 * (x2 - x1);
        first_dof[e] = e;
        for (int k = 0; k < q; ++k) {
            double t = 0.5 * (gx[k] + 1);
            double xx = mix(t, x1, x2);
            xq[e * q + k] = xx;
            int span = find_span(xx, knot.data(), ks, p);
            basis_ders(span, xx, knot.data(), p, ders,
                       bt + (static_cast<size_t>(e) * q + k) * (ders + 1) * (p + 1));
        }
    }
    return ADSB_OK;
}

// band layout helpers: A(i,j) at ab[j*ldab + kl+ku+i-j], kl = ku = p, ldab = 3p+1
int matrix_from_tables(int kind, double h, int p, int elements, int q, int ders, const double* bt,
                       const double* w, const double* J, double* ab) {
    if ((kind == 1 || kind == 2 || kind == 3) && ders < 1)
        return fail(ADSB_EINVAL, "matrix_1d: derivative tables needed");
    const int n = elements + p, ldab = 3 * p + 1;
    std::fill(ab, ab + static_cast<size_t>(ldab) * n, 0.0);
    auto at = [&](int i, int j) -> double& { return ab[static_cast<size_t>(j) * ldab + 2 * p + i - j]; };
    const int m = p + 1;
    for (int e = 0; e < elements; ++e) {
        for (int k = 0; k < q; ++k) {
            const double* val = bt + (static_cast<size_t>(e) * q + k) * (ders + 1) * m;
            const double* der = ders >= 1 ? val + m : nullptr;
            for (int r = 0; r < m; ++r) {
                for (int c = 0; c < m; ++c) {
                    double& dst = at(e + r, e + c);
                    switch (kind) {
                    case 0: dst += val[r] * val[c] * w[k] * J[e]; break;
                    case 1: dst += der[r] * der[c] * w[k] * J[e]; break;
                    case 2: dst += val[r] * der[c] * w[k] * J[e]; break;
                    default: dst += (val[r] * val[c] + h * der[r] * der[c]) * w[k] * J[e]; break;
                    }
                }
            }
        }
    }
    return ADSB_OK;
}

void fix_dof(int k, int p, int n, double* ab) {
    const int ldab = 3 * p + 1, last = n - 1;
    auto at = [&](int i, int j) -> double& { return ab[static_cast<size_t>(j) * ldab + 2 * p + i - j]; };
    for (int i = std::max(k - p, 0); i <= std::min(k + p, last); ++i) at(k, i) = 0;
    at(k, k) = 1;
}

int matrix_1d(int kind, int p, int elements, double a, double b, double h, int fix, double* ab) {
    if (kind < 0 || kind > 3) return fail(ADSB_EINVAL, "matrix_1d: kind must be 0..3");
    if (p < 1 || p > ADSB_MAX_P || elements < 1) return fail(ADSB_EINVAL, "matrix_1d: bad p/elements");
    const int q = p + 1, ders = 1, m = p + 1;
    std::vector<double> bt(static_cast<size_t>(elements) * q * (ders + 1) * m), xq(elements * q), w(q),
        J(elements);
    std::vector<int> fd(elements);
    if (int rc = basis_tables(p, elements, a, b, q, ders, bt.data(), xq.data(), w.data(), J.data(), fd.data()))
        return rc;
    if (int rc = matrix_from_tables(kind, h, p, elements, q, ders, bt.data(), w.data(), J.data(), ab))
        return rc;
    if (fix & 1) fix_dof(0, p, elements + p, ab);
    if (fix & 2) fix_dof(elements + p - 1, p, elements + p, ab);
    return ADSB_OK;
}

// Unblocked banded LU with partial pivoting on the LAPACK layout (DGBTF2 semantics).
int band_factorize(int n, int kl, int ku, double* ab, int ldab, int* ipiv) {
    if (n < 1 || kl < 0 || ku < 0 || ldab < 2 * kl + ku + 1)
        return fail(ADSB_EINVAL, "band_factorize: bad dimensions");
    const int kv = kl + ku;
    auto A = [&](int r, int c) -> double& { return ab[static_cast<size_t>(c) * ldab + r]; };
    for (int j = ku + 1; j < std::min(kv, n); ++j)
        for (int i = kv - j; i < kl; ++i) A(i, j) = 0;
    int ju = 0, info = 0;
    for (int j = 0; j < n; ++j) {
        if (j + kv < n)
            for (int i = 0; i < kl; ++i) A(i, j + kv) = 0;
        const int km = std::min(kl, n - 1 - j);
        int jp = 0;
        double best = std::fabs(A(kv, j));
        for (int i = 1; i <= km; ++i) {
            double v = std::fabs(A(kv + i, j));
            if (v > best) {
                best = v;
                jp = i;
            }
        }
        ipiv[j] = j + jp + 1;
        if (A(kv + jp, j) != 0) {
            ju = std::max(ju, std::min(j + ku + jp, n - 1));
            if (jp != 0)
                for (int c = j; c <= ju; ++c) std::swap(A(kv + jp - (c - j), c), A(kv - (c - j), c));
            if (km > 0) {
                const double r = 1.0 / A(kv, j);
                for (int i = 1; i <= km; ++i) A(kv + i, j) *= r;
                for (int c = j + 1; c <= ju; ++c) {
                    const double t = A(kv - (c - j), c);
                    if (t != 0)
                        for (int i = 1; i <= km; ++i) A(kv + i - (c - j), c) -= A(kv + i, j) * t;
                }
            }
        } else if (info == 0) {
            info = j + 1;
        }
    }
    if (info) return fail(ADSB_ESINGULAR, "band_factorize: zero pivot at column " + std::to_string(info));
    return ADSB_OK;
}

// ---------------------------------------------------------------------------------------------
// A dgbtrf factor WITH row interchanges widens U to kl + ku super-diagonals and, for the Gram matrices of
// degree >= 4, lengthens the chunk chains of the substitution kernel until they no longer fit its depth.  The
// 1-D matrices on this path (Gram, Gram + h * stiffness, with fix_left / fix_right rows) are symmetric
// positive definite up to those rows, so Gaussian elimination WITHOUT interchanges is backward stable for
// them.  refactor_without_pivoting rebuilds A = P^T L U from the caller's factor (every entry a sum of at most
// kl + 1 products: an O(eps) perturbation of A, like dgbtrf's own), eliminates again without interchanges and
// reports whether that was safe: every multiplier at most 4 in magnitude, no pivot below 1e-8 max|A|.  The
// solve then is the same linear system with a different, equally stable elimination order; results agree with
// dgbtrs to rounding (tests: every pivoting factor of the golden set and of the oracle comparisons).
// ---------------------------------------------------------------------------------------------
bool refactor_without_pivoting(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv,
                               std::vector<double>& out) {
    const int kd = kl + ku;
    const int M = kl + kd;            // margin: a row i is kept on columns [i - M, i + M]
    const int W = 2 * M + 1;
    std::vector<double> X(static_cast<size_t>(n) * W, 0.0);
    auto at = [&](int i, int c) -> double& { return X[static_cast<size_t>(i) * W + (c - i + M)]; };
    auto Aat = [&](int r, int c) { return ab[static_cast<size_t>(c) * ldab + r]; };
    for (int c = 0; c < n; ++c)
        for (int k = 0; k <= std::min(kd, c); ++k) at(c - k, c) = Aat(kd - k, c);
    std::vector<double> tmp(W);
    for (int j = n - 2; j >= 0; --j) {
        const int lm = std::min(kl, n - 1 - j);
        const int chi = std::min(n - 1, j + kd);
        for (int i = 1; i <= lm; ++i) {
            const double l = Aat(kd + i, j);
            if (l != 0.0)
                for (int c = j; c <= chi; ++c) at(j + i, c) += l * at(j, c);
        }
        const int pj = ipiv[j] - 1;
        if (pj < j || pj > j + lm) return false;
        if (pj != j) {
            // rows j and pj change places; their entries lie on columns [j - kl, j + kd + kl] at most
            const int clo = std::max(0, j - kl), chi2 = std::min(n - 1, j + kd + kl);
            for (int c = clo; c <= chi2; ++c) {
                const bool in_j = std::abs(c - j) <= M, in_p = std::abs(c - pj) <= M;
                const double vj = in_j ? at(j, c) : 0.0, vp = in_p ? at(pj, c) : 0.0;
                if ((!in_j && vp != 0.0) || (!in_p && vj != 0.0)) return false;
                if (in_j) at(j, c) = vp;
                if (in_p) at(pj, c) = vj;
            }
        }
    }
    double amax = 0.0;
    for (int i = 0; i < n; ++i)
        for (int c = std::max(0, i - M); c <= std::min(n - 1, i + M); ++c) amax = std::max(amax, std::fabs(at(i, c)));
    if (!(amax > 0.0)) return false;
    for (int i = 0; i < n; ++i)
        for (int c = std::max(0, i - M); c <= std::min(n - 1, i + M); ++c)
            if ((c < i - kl || c > i + ku) && std::fabs(at(i, c)) > 1e-13 * amax) return false;  // not a (kl, ku) band
    // elimination without interchanges, in place on the band rows
    out.assign(static_cast<size_t>(n) * ldab, 0.0);
    for (int j = 0; j < n; ++j) {
        const double piv = at(j, j);
        if (!(std::fabs(piv) >= 1e-8 * amax)) return false;
        const int lm = std::min(kl, n - 1 - j), cu = std::min(n - 1, j + ku);
        for (int i = 1; i <= lm; ++i) {
            const double l = at(j + i, j) / piv;
            if (!(std::fabs(l) <= 4.0)) return false;
            out[static_cast<size_t>(j) * ldab + kd + i] = l;
            for (int c = j + 1; c <= cu; ++c) at(j + i, c) -= l * at(j, c);
        }
        for (int k = 0; k <= std::min(ku, j); ++k) out[static_cast<size_t>(j) * ldab + kd - k] = at(j - k, j);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Plan for the chunk-parallel substitution kernel (kernels_sweep.cu).  A line of n unknowns is cut
// into SC chunks of CH columns.  Each chunk runs the pivoted forward recurrence and then the back
// substitution from ZERO incoming states (local pass); the exact dgbtrs result is recovered as
//     x = x_local + Xi * delta_c + Psi * t_c
// delta_c : true forward state entering chunk c (KL partially updated rows),
// t_c     : true first KD unknowns right of chunk c,
// chained across chunks by  delta_{c+1} = Delta_c + T_c delta_c   and   t_{c-1} = X_c + R_c t_c,
// X_c = xfirst_local_c + Xi_first_c delta_c.  All response tables depend only on the factor, so
// they are tabulated here once.  Because the responses decay, the chains are evaluated to a finite
// depth D in parallel (delta_c = sum_{d<=D} W_{c,d} Delta_{c-d}); the depth is chosen so that the
// dropped products are below 1e-18 -- if that needs more than MAX_DEPTH terms the kernel falls
// back to the sequential chain (seq = 1).
// ---------------------------------------------------------------------------------------------
namespace {
void matmul(int K, const double* A, const double* B, double* C) {  // C = A*B, K x K row-major
    for (int r = 0; r < K; ++r)
        for (int c = 0; c < K; ++c) {
            double acc = 0;
            for (int m = 0; m < K; ++m) acc += A[r * K + m] * B[m * K + c];
            C[r * K + c] = acc;
        }
}
double maxabs(const double* A, int count) {
    double m = 0;
    for (int i = 0; i < count; ++i) m = std::max(m, std::fabs(A[i]));
    return m;
}
}  // namespace

int build_sweep_plan(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, int ch, int group,
                     SweepPlan& P, bool force_piv) {
    const int kd = kl + ku;
    bool piv = force_piv;  // a slab of a factor with row interchanges uses the interchange variant even if its own
                           // columns have none, so that every slab runs the same kernel and state widths
    for (int j = 0; j < n; ++j) {
        int t = ipiv[j] - 1 - j;
        if (t < 0 || t > kl || j + t >= n) return fail(ADSB_EINVAL, "factor: bad pivot vector");
        if (t) piv = true;
    }
    auto Aat = [&](int r, int c) { return ab[static_cast<size_t>(c) * ldab + r]; };
    int kd_eff = 0;  // widest non-zero super-diagonal of U
    for (int c = 0; c < n; ++c)
        for (int k = 1; k <= std::min(kd, c); ++k)
            if (Aat(kd - k, c) != 0) kd_eff = std::max(kd_eff, k);
    for (int j = 0; j < n; ++j)
        if (Aat(kd, j) == 0) return fail(ADSB_ESINGULAR, "factor: zero diagonal in U at column " + std::to_string(j + 1));
    // template variant: (KL, KD) = (P, P) without pivoting, (P, 2P) with
    int var = piv ? std::max({kl, (kd_eff + 1) / 2, 1}) : std::max({kl, kd_eff, 1});
    if (var > 5) return fail(ADSB_EINVAL, "factor: bandwidth beyond the compiled kernel variants (p <= 5)");
    const int KL = var, KD = piv ? 2 * var : var;
    if (ch < KD) return fail(ADSB_EINVAL, "factor: chunk shorter than the band");
    const int ST = ((n + ch - 1) / ch + group - 1) / group;  // threads per line
    const int SC = ST * group;                                // chunks per line (some may be empty)
    const size_t rows = static_cast<size_t>(SC) * ch + KL + KD;
    P.n = n; P.KL = KL; P.KD = KD; P.piv = piv ? 1 : 0; P.CH = ch; P.R = group; P.SC = SC; P.ST = ST;
    P.LF = sweep_pitch(KL); P.LB = sweep_pitch(KD + 1); P.LC = sweep_pitch(KD + KL);
    P.rows = static_cast<int>(rows);
    std::vector<double> Lm(rows * KL, 0.0), Ut(rows * KD, 0.0), rinv(rows, 0.0), Phi(rows * KL, 0.0),
        Psi(rows * KD, 0.0), Xi(rows * KL, 0.0), T(static_cast<size_t>(SC) * KL * KL, 0.0),
        Rm(static_cast<size_t>(SC) * KD * KD, 0.0);
    P.pv.assign(rows, 0);
    for (int j = 0; j < n; ++j) {
        const int lm = std::min(kl, n - 1 - j);
        for (int i = 0; i < lm; ++i) Lm[static_cast<size_t>(j) * KL + i] = Aat(kd + 1 + i, j);
        P.pv[j] = ipiv[j] - 1 - j;
        for (int k = 1; k <= kd_eff && j + k < n; ++k) Ut[static_cast<size_t>(j) * KD + k - 1] = Aat(kd - k, j + k);
        rinv[j] = 1.0 / Aat(kd, j);
    }
    std::vector<double> win(ch + KL), xs(ch + KD);
    for (int c = 0; c < SC; ++c) {
        const int j0 = c * ch;
        // forward response to a unit perturbation of window row r at the chunk start
        for (int r = 0; r < KL; ++r) {
            std::fill(win.begin(), win.end(), 0.0);
            win[r] = 1.0;
            for (int i = 0; i < ch; ++i) {
                const size_t j = static_cast<size_t>(j0) + i;
                const int t = P.pv[j];
                if (t) std::swap(win[i], win[i + t]);
                for (int m = 1; m <= KL; ++m) win[i + m] = std::fma(-Lm[j * KL + m - 1], win[i], win[i + m]);
                Phi[j * KL + r] = win[i];
            }
            for (int m = 0; m < KL; ++m) T[(static_cast<size_t>(c) * KL + m) * KL + r] = win[ch + m];
            // its image under the local back substitution (zero state on the right)
            std::fill(xs.begin(), xs.end(), 0.0);
            for (int i = ch - 1; i >= 0; --i) {
                const size_t j = static_cast<size_t>(j0) + i;
                double acc = Phi[j * KL + r];
                for (int m = KD; m >= 1; --m) acc = std::fma(-Ut[j * KD + m - 1], xs[i + m], acc);
                xs[i] = acc * rinv[j];
                Xi[j * KL + r] = xs[i];
            }
        }
        // backward response to a unit value of the k-th unknown right of the chunk
        for (int k = 0; k < KD; ++k) {
            std::fill(xs.begin(), xs.end(), 0.0);
            xs[ch + k] = 1.0;
            for (int i = ch - 1; i >= 0; --i) {
                const size_t j = static_cast<size_t>(j0) + i;
                double acc = 0;
                for (int m = KD; m >= 1; --m) acc = std::fma(-Ut[j * KD + m - 1], xs[i + m], acc);
                xs[i] = acc * rinv[j];
                Psi[j * KD + k] = xs[i];
            }
            for (int i = 0; i < KD; ++i) Rm[(static_cast<size_t>(c) * KD + i) * KD + k] = xs[i];
        }
    }
    // chained products and the depth at which they vanish
    const int MD = SWEEP_MAX_DEPTH;
    P.W.assign(static_cast<size_t>(SC) * (MD - 1) * KL * KL, 0.0);
    P.V.assign(static_cast<size_t>(SC) * (MD - 1) * KD * KD, 0.0);
    int DF = 1, DB = 1;
    // growing responses (stiffness-dominated K = M + hS far from diagonal dominance): explicit
    // products of the transfer matrices would cancel catastrophically -> chain sequentially
    bool seq = maxabs(T.data(), static_cast<int>(T.size())) > 1.0 || maxabs(Rm.data(), static_cast<int>(Rm.size())) > 1.0;
    std::vector<double> cur(std::max(KL * KL, KD * KD)), nxt(cur.size());
    const double tiny = 1e-18;  // dropped chain terms: a hundredth of the unit round-off relative to the data
    for (int c = 0; c < SC; ++c) {
        // forward: delta_c = Delta_{c-1} + T_{c-1} Delta_{c-2} + T_{c-1} T_{c-2} Delta_{c-3} + ...
        if (c >= 2) {
            std::copy(&T[static_cast<size_t>(c - 1) * KL * KL], &T[static_cast<size_t>(c - 1) * KL * KL] + KL * KL, cur.begin());
            for (int d = 2; c - d >= 0; ++d) {
                if (maxabs(cur.data(), KL * KL) < tiny) break;
                if (d > MD) { seq = true; break; }
                std::copy(cur.begin(), cur.begin() + KL * KL, &P.W[(static_cast<size_t>(c) * (MD - 1) + d - 2) * KL * KL]);
                DF = std::max(DF, d);
                if (c - d - 1 < 0) break;
                matmul(KL, cur.data(), &T[static_cast<size_t>(c - d) * KL * KL], nxt.data());
                std::swap(cur, nxt);
            }
        }
        // backward: t_c = X_{c+1} + R_{c+1} X_{c+2} + R_{c+1} R_{c+2} X_{c+3} + ...
        if (c + 2 < SC) {
            std::copy(&Rm[static_cast<size_t>(c + 1) * KD * KD], &Rm[static_cast<size_t>(c + 1) * KD * KD] + KD * KD, cur.begin());
            for (int d = 2; c + d < SC; ++d) {
                if (maxabs(cur.data(), KD * KD) < tiny) break;
                if (d > MD) { seq = true; break; }
                std::copy(cur.begin(), cur.begin() + KD * KD, &P.V[(static_cast<size_t>(c) * (MD - 1) + d - 2) * KD * KD]);
                DB = std::max(DB, d);
                if (c + d + 1 >= SC) break;
                matmul(KD, cur.data(), &Rm[static_cast<size_t>(c + d) * KD * KD], nxt.data());
                std::swap(cur, nxt);
            }
        }
    }
    P.DF = DF; P.DB = DB; P.seq = seq ? 1 : 0;
    // packed per-column records
    P.cfF.assign(rows * P.LF, 0.0);
    P.cfB.assign(rows * P.LB, 0.0);
    P.cfC.assign(rows * P.LC, 0.0);
    for (size_t j = 0; j < rows; ++j) {
        for (int m = 0; m < KL; ++m) P.cfF[j * P.LF + m] = Lm[j * KL + m];
        for (int k = 0; k < KD; ++k) P.cfB[j * P.LB + k] = Ut[j * KD + k];
        P.cfB[j * P.LB + KD] = rinv[j];
        for (int k = 0; k < KD; ++k) P.cfC[j * P.LC + k] = Psi[j * KD + k];
        for (int m = 0; m < KL; ++m) P.cfC[j * P.LC + KD + m] = Xi[j * KL + m];
    }
    P.T = std::move(T);
    P.Rm = std::move(Rm);
    return ADSB_OK;
}


// ---------------------------------------------------------------------------------------------
// Segmented substitution.  With L, U the (row-interchanged) band factors and a segment s = rows
// [a, b) whose boundary no interchange crosses, the sequential dgbtrs recurrence splits exactly into
//     xhat_s  = solve of the segment alone (its own columns of the factor, zero incoming states)
//     din_s   = forward updates that the columns left of a apply to rows a .. a+KL-1
//     tin_s   = true unknowns b .. b+KD-1
//     x_s     = xhat_s + Xi_s din_s + Psi_s tin_s
// chained by  din_{s+1} = E_s xhat_s[last KL] + T_s din_s   and   tin_{s-1} = xhat_s[first KD] + Xi_s[first KD] din_s
// + Psi_s[first KD] tin_s.  E, T, Xi, Psi depend only on the factor; the chains are expanded to the depth at
// which the products of T (of Psi[first KD]) drop below `tol`.
// ---------------------------------------------------------------------------------------------
namespace {
struct BandRows {  // per column j: multipliers, U row (super-diagonals), diagonal, pivot offset
    int n, kl, kd;
    std::vector<double> Lm, Ut, diag;
    std::vector<int> pv;
};
int unpack_factor(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, BandRows& B) {
    const int kd = kl + ku;
    B.n = n; B.kl = kl; B.kd = kd;
    B.Lm.assign(static_cast<size_t>(n) * std::max(kl, 1), 0.0);
    B.Ut.assign(static_cast<size_t>(n) * std::max(kd, 1), 0.0);
    B.diag.assign(n, 0.0);
    B.pv.assign(n, 0);
    auto Aat = [&](int r, int c) { return ab[static_cast<size_t>(c) * ldab + r]; };
    for (int j = 0; j < n; ++j) {
        const int t = ipiv[j] - 1 - j;
        if (t < 0 || t > kl || j + t >= n) return fail(ADSB_EINVAL, "factor: bad pivot vector");
        B.pv[j] = t;
        const int lm = std::min(kl, n - 1 - j);
        for (int i = 0; i < lm; ++i) B.Lm[static_cast<size_t>(j) * kl + i] = Aat(kd + 1 + i, j);
        for (int k = 1; k <= kd && j + k < n; ++k) B.Ut[static_cast<size_t>(j) * kd + k - 1] = Aat(kd - k, j + k);
        B.diag[j] = Aat(kd, j);
        if (B.diag[j] == 0) return fail(ADSB_ESINGULAR, "factor: zero diagonal in U at column " + std::to_string(j + 1));
    }
    return ADSB_OK;
}
bool clean_cut(const BandRows& B, int a) {  // no interchange of a column left of a reaches row a or beyond
    for (int j = std::max(0, a - B.kl); j < a; ++j)
        if (j + B.pv[j] >= a) return false;
    return true;
}
void matmul_rect(int R, int K, int C, const double* A, const double* Bm, double* Cm) {  // (R x K) * (K x C)
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) {
            double acc = 0;
            for (int m = 0; m < K; ++m) acc += A[r * K + m] * Bm[m * C + c];
            Cm[r * C + c] = acc;
        }
}
}  // namespace

int pick_segment_bounds(int n, int kl, const int* ipiv, int S, int align, int min_rows, int* bounds) {
    if (S < 1 || n < 1) return fail(ADSB_EINVAL, "segments: bad count");
    BandRows B;
    B.n = n; B.kl = kl; B.kd = 0;
    B.pv.assign(n, 0);
    for (int j = 0; j < n; ++j) {
        const int t = ipiv[j] - 1 - j;
        if (t < 0 || t > kl || j + t >= n) return fail(ADSB_EINVAL, "factor: bad pivot vector");
        B.pv[j] = t;
    }
    if (align < 1) align = 1;
    bounds[0] = 0;
    bounds[S] = n;
    for (int s = 1; s < S; ++s) {
        // target: balanced cut, rounded to the alignment; then the nearest clean row (alternating search)
        const long long ideal = static_cast<long long>(n) * s / S;
        int target = static_cast<int>((ideal + align / 2) / align * align);
        int found = -1;
        auto usable = [&](int a) { return a > bounds[s - 1] && a < n && clean_cut(B, a); };
        for (int d = 0; d <= 2 && found < 0 && align > 1; ++d)  // aligned candidates first
            for (int sgn = -1; sgn <= 1 && found < 0; sgn += 2)
                if (usable(target + sgn * d * align)) found = target + sgn * d * align;
        for (int d = 0; d <= 4 * kl + align && found < 0; ++d)
            for (int sgn = -1; sgn <= 1 && found < 0; sgn += 2)
                if (usable(target + sgn * d)) found = target + sgn * d;
        if (found < 0) return fail(ADSB_EINVAL, "segments: no cut free of row interchanges near row " + std::to_string(target));
        bounds[s] = found;
    }
    for (int s = 0; s < S; ++s)
        if (bounds[s + 1] - bounds[s] < min_rows) return fail(ADSB_EINVAL, "segments: a segment is shorter than the band");
    return ADSB_OK;
}

int build_segment_plan(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, int S, const int* bounds,
                       double tol, SegPlan& P) {
    BandRows B;
    if (int rc = unpack_factor(n, kl, ku, ldab, ab, ipiv, B)) return rc;
    const int kd = B.kd;
    bool piv = false;
    int kd_eff = 0;
    for (int j = 0; j < n; ++j) {
        if (B.pv[j]) piv = true;
        for (int k = 1; k <= kd; ++k)
            if (B.Ut[static_cast<size_t>(j) * kd + k - 1] != 0) kd_eff = std::max(kd_eff, k);
    }
    // same (KL, KD) variant as build_sweep_plan picks for the whole factor
    const int var = piv ? std::max({kl, (kd_eff + 1) / 2, 1}) : std::max({kl, kd_eff, 1});
    const int KL = var, KD = piv ? 2 * var : var;
    if (S < 1 || bounds[0] != 0 || bounds[S] != n) return fail(ADSB_EINVAL, "segments: bounds must run from 0 to n");
    for (int s = 0; s < S; ++s) {
        if (bounds[s + 1] - bounds[s] < std::max(KL, KD)) return fail(ADSB_EINVAL, "segments: a segment is shorter than the band");
        if (s && !clean_cut(B, bounds[s])) return fail(ADSB_EINVAL, "segments: a row interchange crosses a segment boundary");
    }
    P = SegPlan{};
    P.n = n; P.KL = KL; P.KD = KD; P.piv = piv ? 1 : 0; P.S = S;
    P.bounds.assign(bounds, bounds + S + 1);
    const int KC = KD + KL;
    P.cf.assign(static_cast<size_t>(n) * KC, 0.0);
    P.E.assign(static_cast<size_t>(S) * KL * KL, 0.0);
    P.XiF.assign(static_cast<size_t>(S) * KD * KL, 0.0);
    std::vector<double> T(static_cast<size_t>(S) * KL * KL, 0.0), R(static_cast<size_t>(S) * KD * KD, 0.0);
    auto Lm = [&](int j, int m) { return m <= kl ? B.Lm[static_cast<size_t>(j) * kl + m - 1] : 0.0; };  // m = 1..
    auto Ut = [&](int j, int m) { return m <= kd ? B.Ut[static_cast<size_t>(j) * kd + m - 1] : 0.0; };
    for (int s = 0; s < S; ++s) {
        const int a = bounds[s], b = bounds[s + 1], ns = b - a;
        std::vector<double> win(ns + KL), xs(ns + KD), phi(ns);
        for (int r = 0; r < KL; ++r) {  // unit forward in-state on row a + r
            std::fill(win.begin(), win.end(), 0.0);
            win[r] = 1.0;
            for (int i = 0; i < ns; ++i) {
                const int j = a + i, t = B.pv[j];
                if (t) std::swap(win[i], win[i + t]);
                for (int m = 1; m <= KL; ++m)
                    if (j + m < n) win[i + m] = std::fma(-Lm(j, m), win[i], win[i + m]);
                phi[i] = win[i];
            }
            for (int m = 0; m < KL; ++m) T[(static_cast<size_t>(s) * KL + m) * KL + r] = win[ns + m];
            std::fill(xs.begin(), xs.end(), 0.0);
            for (int i = ns - 1; i >= 0; --i) {
                const int j = a + i;
                double acc = phi[i];
                for (int m = KD; m >= 1; --m)
                    if (i + m < ns) acc = std::fma(-Ut(j, m), xs[i + m], acc);
                xs[i] = acc / B.diag[j];
                P.cf[static_cast<size_t>(j) * KC + KD + r] = xs[i];
            }
            for (int i = 0; i < KD; ++i) P.XiF[(static_cast<size_t>(s) * KD + i) * KL + r] = xs[i];
        }
        for (int k = 0; k < KD; ++k) {  // unit value of unknown b + k
            std::fill(xs.begin(), xs.end(), 0.0);
            xs[ns + k] = 1.0;
            for (int i = ns - 1; i >= 0; --i) {
                const int j = a + i;
                double acc = 0;
                for (int m = KD; m >= 1; --m) acc = std::fma(-Ut(j, m), xs[i + m], acc);
                xs[i] = acc / B.diag[j];
                P.cf[static_cast<size_t>(j) * KC + k] = xs[i];
            }
            for (int i = 0; i < KD; ++i) R[(static_cast<size_t>(s) * KD + i) * KD + k] = xs[i];
        }
        // E_s = (updates of rows b .. b+KL-1 by the last KL columns) * (U restricted to the last KL rows)
        std::vector<double> Lout(KL * KL, 0.0), Ul(KL * KL, 0.0);
        for (int ii = 0; ii < KL; ++ii) {
            const int j = b - KL + ii;
            Ul[ii * KL + ii] = B.diag[j];
            for (int m = 1; ii + m < KL; ++m) Ul[ii * KL + ii + m] = Ut(j, m);
            for (int k = 0; k < KL; ++k) {
                const int m = b + k - j;
                if (m >= 1 && m <= KL && b + k < n) Lout[k * KL + ii] = -Lm(j, m);
            }
        }
        matmul_rect(KL, KL, KL, Lout.data(), Ul.data(), &P.E[static_cast<size_t>(s) * KL * KL]);
    }
    if (maxabs(T.data(), static_cast<int>(T.size())) > 1.0 || maxabs(R.data(), static_cast<int>(R.size())) > 1.0)
        return fail(ADSB_EINVAL, "segments: the factor's boundary responses grow (not diagonally dominant enough)");
    // chain products: din_s = Dseg_{s-1} + T_{s-1} Dseg_{s-2} + T_{s-1} T_{s-2} Dseg_{s-3} + ...
    auto chain = [&](int K, const std::vector<double>& M, bool forward, std::vector<double>& out) {
        const int Dmax = std::max(1, S - 1);
        std::vector<std::vector<double>> prod(static_cast<size_t>(S) * Dmax);
        int depth = 1;
        std::vector<double> cur(K * K), nxt(K * K);
        for (int s = 0; s < S; ++s) {
            std::fill(cur.begin(), cur.end(), 0.0);
            for (int i = 0; i < K; ++i) cur[i * K + i] = 1.0;
            for (int d = 1; d <= Dmax; ++d) {
                const int src = forward ? s - d : s + d;  // segment whose state term d multiplies
                if (src < 0 || src >= S) break;
                if (d > 1 && maxabs(cur.data(), K * K) < tol) break;
                prod[static_cast<size_t>(s) * Dmax + d - 1] = cur;
                depth = std::max(depth, d);
                // next: multiply by the transfer of segment `src` (forward: on the right; backward: on the right too)
                matmul(K, cur.data(), &M[static_cast<size_t>(src) * K * K], nxt.data());
                std::swap(cur, nxt);
            }
        }
        out.assign(static_cast<size_t>(S) * depth * K * K, 0.0);
        for (int s = 0; s < S; ++s)
            for (int d = 0; d < depth; ++d) {
                const auto& m = prod[static_cast<size_t>(s) * Dmax + d];
                if (!m.empty()) std::copy(m.begin(), m.end(), &out[(static_cast<size_t>(s) * depth + d) * K * K]);
            }
        return depth;
    };
    P.DF = chain(KL, T, true, P.Wf);
    P.DB = chain(KD, R, false, P.Vb);
    return ADSB_OK;
}

}  // namespace adsb

// ------------------------------------- C ABI (host part) --------------------------------------
extern "C" {

int adsb_abi_version(void) { return ADSB_ABI_VERSION; }
const char* adsb_last_error(void) { return adsb::g_last_error.c_str(); }

int adsb_gauss(int q, double* x, double* w) { return adsb::gauss_rule(q, x, w); }

int adsb_knots(int p, int elements, double a, double b, double* knots) {
    return adsb::make_knots(p, elements, a, b, knots);
}

int adsb_find_span(double x, const double* knots, int knot_size, int p) {
    return adsb::find_span(x, knots, knot_size, p);
}

int adsb_basis_ders(int span, double x, const double* knots, int p, int ders, double* out) {
    if (p < 1 || p > ADSB_MAX_P || ders < 0) return adsb::fail(ADSB_EINVAL, "basis_ders: bad p/ders");
    adsb::basis_ders(span, x, knots, p, ders, out);
    return ADSB_OK;
}

int adsb_basis_tables(int p, int elements, double a, double b, int q, int ders, double* b_flat,
                      double* xq, double* w, double* J, int* first_dof) {
    return adsb::basis_tables(p, elements, a, b, q, ders, b_flat, xq, w, J, first_dof);
}

int adsb_matrix_1d(int kind, int p, int elements, double a, double b, double h, int fix, double* ab) {
    return adsb::matrix_1d(kind, p, elements, a, b, h, fix, ab);
}

int adsb_band_factorize(int n, int kl, int ku, double* ab, int ldab, int* ipiv) {
    return adsb::band_factorize(n, kl, ku, ab, ldab, ipiv);
}

int adsb_segment_bounds(int n, int kl, const int* ipiv, int nseg, int align, int* bounds) {
    if (!ipiv || !bounds) return adsb::fail(ADSB_EINVAL, "segment_bounds: null argument");
    return adsb::pick_segment_bounds(n, kl, ipiv, nseg, align, 1, bounds);
}

int adsb_segment_plan(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, int nseg, const int* bounds,
                      double tol, int* dims, double* E, double* Wf, double* Vb, double* XiF, double* cf) {
    if (!ab || !ipiv || !bounds) return adsb::fail(ADSB_EINVAL, "segment_plan: null argument");
    adsb::SegPlan P;
    if (int rc = adsb::build_segment_plan(n, kl, ku, ldab, ab, ipiv, nseg, bounds, tol > 0 ? tol : 1e-20, P)) return rc;
    if (dims) {
        const int d[8] = {P.KL, P.KD, P.piv, P.S, P.DF, P.DB, P.n, 0};
        std::copy(d, d + 8, dims);
    }
    auto put = [](double* dst, const std::vector<double>& v) {
        if (dst) std::copy(v.begin(), v.end(), dst);
    };
    put(E, P.E);
    put(Wf, P.Wf);
    put(Vb, P.Vb);
    put(XiF, P.XiF);
    put(cf, P.cf);
    return ADSB_OK;
}

}  // extern "C"
