/* oracle/ads_oracle.c -- TEST INFRASTRUCTURE ONLY.  See ads_oracle.h for the rules.
 *
 * A plain-C restatement of the reference algorithm for the ADS time step.  Every function
 * cites the reference file:line it follows (paths relative to /root/reference).  Operation
 * order is kept the same as the reference inside a quadrature point / a matrix column so that
 * differences against the compiled reference stay at the last-bit level; the band solve is
 * third-party arithmetic in the reference (LAPACK dgbtrf_/dgbtrs_, unpinned version, call
 * sites include/ads/lin/band_solve.hpp:17,:29) and is restated here from the published
 * LAPACK 3.9 unblocked algorithms DGBTF2 / DGBTRS(+DTBSV).
 *
 * Parity status: PINNED (tests/test_oracle.py: reference KATs + golden vectors produced by the
 * compiled unmodified reference + live comparison when oracle/_ref/libads_ref.so loads).
 */
#define _GNU_SOURCE
#include "ads_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define MAXP 8
#define MAXQ 16

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/* ---- Gauss-Legendre rule -------------------------------------------------------------------
 * The reference hard-codes 20-digit literals (include/ads/quad/gauss.hpp, n = 2..64).  They are
 * the classical Gauss-Legendre nodes/weights, recomputed here by Newton iteration on P_q in
 * long double and rounded to double; tests check bit-equality with the reference table. */
int orc_gauss(int q, double* x, double* w) {
    if (q < 2 || q > 64) return -1;
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < q; ++i) {
        /* i-th root counted from the right; store ascending */
        long double t = cosl(pi * ((long double) i + 0.75L) / ((long double) q + 0.5L));
        long double dp = 1;
        for (int it = 0; it < 100; ++it) {
            long double p0 = 1, p1 = t;
            for (int k = 2; k <= q; ++k) {
                long double pk = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k;
                p0 = p1;
                p1 = pk;
            }
            dp = q * (t * p1 - p0) / (t * t - 1);
            long double dt = p1 / dp;
            t -= dt;
            if (fabsl(dt) < 1e-21L) break;
        }
        /* derivative at the converged root */
        {
            long double p0 = 1, p1 = t;
            for (int k = 2; k <= q; ++k) {
                long double pk = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k;
                p0 = p1;
                p1 = pk;
            }
            dp = q * (t * p1 - p0) / (t * t - 1);
        }
        int idx = q - 1 - i;
        if (2 * i + 1 == q) t = 0;
        x[idx] = (double) t;
        w[idx] = (double) (2 / ((1 - t * t) * dp * dp));
    }
    return 0;
}

/* ---- B-spline basis ------------------------------------------------------------------------ */

/* include/ads/util.hpp:16-25 */
static double lerp_t(double t, double a, double b) { return (1 - t) * a + t * b; }
static double lerp_i(int i, int n, double a, double b) { return lerp_t((double) i / (double) n, a, b); }

/* src/ads/bspline/bspline.cpp:26-43 (the 5-argument create_basis with repeated_nodes = 0, which is
 * what dimension::bspline_basis calls, include/ads/simulation/dimension.hpp:53-56) */
int orc_knots(int p, int elements, double a, double b, double* knot) {
    int points = elements + 1;
    int knot_size = 2 * (p + 1) + (points - 2);
    for (int i = 0; i <= p; ++i) {
        knot[i] = a;
        knot[knot_size - i - 1] = b;
    }
    for (int i = 1; i < points - 1; ++i) knot[p + 1 + (i - 1)] = lerp_i(i, elements, a, b);
    return knot_size;
}

/* src/ads/bspline/bspline.cpp:61-81 */
int orc_find_span(double x, const double* knot, int knot_size, int p) {
    int low = p;
    int high = knot_size - p - 1;
    if (x >= knot[high]) return high - 1;
    if (x <= knot[low]) return low;
    int idx = (low + high) / 2;
    while (x < knot[idx] || x >= knot[idx + 1]) {
        if (x < knot[idx]) high = idx; else low = idx;
        idx = (low + high) / 2;
    }
    return idx;
}

/* src/ads/bspline/bspline.cpp:102-160 (NURBS book A2.3) */
void orc_basis_ders(int i, double x, const double* knot, int p, int der, double* out) {
    double ndu[MAXP + 1][MAXP + 1], a[2][MAXP + 1], left[MAXP + 2], right[MAXP + 2];
    ndu[0][0] = 1;
    for (int j = 1; j <= p; ++j) {
        left[j] = x - knot[i + 1 - j];
        right[j] = knot[i + j] - x;
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
        for (int k = 1; k <= der; ++k) {
            double d = 0;
            int rk = r - k, pk = p - k;
            if (r >= k) {
                a[s2][0] = a[s1][0] / ndu[pk + 1][rk];
                d = a[s2][0] * ndu[rk][pk];
            }
            int j1 = (rk >= -1) ? 1 : -rk;
            int j2 = (r - 1 <= pk) ? k - 1 : p - r;
            for (int j = j1; j <= j2; ++j) {
                a[s2][j] = (a[s1][j] - a[s1][j - 1]) / ndu[pk + 1][rk + j];
                d += a[s2][j] * ndu[rk + j][pk];
            }
            if (r <= pk) {
                a[s2][k] = -a[s1][k - 1] / ndu[pk + 1][r];
                d += a[s2][k] * ndu[r][pk];
            }
            out[k * (p + 1) + r] = d;
            int t = s1; s1 = s2; s2 = t;
        }
    }
    int r = p;
    for (int k = 1; k <= der; ++k) {
        for (int j = 0; j <= p; ++j) out[k * (p + 1) + j] *= r;
        r *= (p - k);
    }
}

/* src/ads/basis_data.cpp:63-114 (elem_division = 1) + bspline.cpp:162-173 (first_nonzero_dofs) */
int orc_basis_tables(int p, int elements, double a, double b, int q, int ders, double* b_flat,
                     double* xq, double* w, double* J, int* first_dof) {
    if (p > MAXP || q > MAXQ) return -1;
    int ks = elements + 2 * p + 1;
    double* knot = (double*) malloc(sizeof(double) * (size_t) ks);
    double* points = (double*) malloc(sizeof(double) * (size_t) (elements + 1));
    double gx[64], gw[64];
    orc_knots(p, elements, a, b, knot);
    orc_gauss(q, gx, gw);
    /* basis::points = distinct knots (include/ads/bspline/bspline.hpp:24-31) */
    int np = 0;
    points[np++] = knot[0];
    for (int i = 1; i < ks; ++i)
        if (knot[i] != knot[i - 1]) points[np++] = knot[i];
    /* basis_data.cpp:88-94 with elem_division = 1: lerp(0,1,x1,x2), lerp(1,1,x1,x2) */
    double* pts = (double*) malloc(sizeof(double) * (size_t) (elements + 1));
    for (int e = 0; e < elements; ++e) {
        double x1 = points[e], x2 = points[e + 1];
        pts[e] = lerp_i(0, 1, x1, x2);
        pts[e + 1] = lerp_i(1, 1, x1, x2);
    }
    int ne = 0;
    for (int i = p; i + 1 < ks - p; ++i)
        if (knot[i] != knot[i + 1]) first_dof[ne++] = i - p;
    for (int k = 0; k < q; ++k) w[k] = gw[k];
    for (int e = 0; e < elements; ++e) {
        double x1 = pts[e], x2 = pts[e + 1];
        J[e] = 0.5 * (x2 - x1);
        for (int k = 0; k < q; ++k) {
            double t = 0.5 * (gx[k] + 1);
            xq[e * q + k] = lerp_t(t, x1, x2);
        }
        for (int k = 0; k < q; ++k) {
            double x = xq[e * q + k];
            int span = orc_find_span(x, knot, ks, p);
            orc_basis_ders(span, x, knot, p, ders,
                           b_flat + ((size_t) e * q + k) * (size_t) ((ders + 1) * (p + 1)));
        }
    }
    free(knot); free(points); free(pts);
    return 0;
}

/* ---- 1-D matrices in LAPACK band storage ---------------------------------------------------- */

/* include/ads/lin/band_matrix.hpp:31-40,:69-73: band_matrix(kl,ku,n) => row_offset = kl,
 * column_size = 2kl+ku+1, A(i,j) at data[j*ldab + kl+ku+i-j] */
#define AB(ab, ldab, kl, ku, i, j) (ab)[(size_t) (j) * (ldab) + (kl) + (ku) + (i) - (j)]

/* src/ads/form_matrix.cpp:8-60, examples/implicit/implicit.hpp:46-64, dimension.cpp:23-29 */
int orc_matrix_1d(int kind, int p, int elements, double a, double b, double h, int fix, double* ab) {
    int q = p + 1, n = elements + p, ldab = 3 * p + 1;
    size_t tb = (size_t) elements * q * 2 * (p + 1);
    double* bt = (double*) malloc(sizeof(double) * tb);
    double* xq = (double*) malloc(sizeof(double) * (size_t) elements * q);
    double* J = (double*) malloc(sizeof(double) * (size_t) elements);
    int* fd = (int*) malloc(sizeof(int) * (size_t) elements);
    double w[MAXQ];
    orc_basis_tables(p, elements, a, b, q, 1, bt, xq, w, J, fd);
    memset(ab, 0, sizeof(double) * (size_t) ldab * n);
    for (int e = 0; e < elements; ++e) {
        for (int k = 0; k < q; ++k) {
            const double* B = bt + ((size_t) e * q + k) * 2 * (p + 1);
            const double* dB = B + (p + 1);
            int first = fd[e];
            for (int ia = 0; ia <= p; ++ia) {
                for (int ib = 0; ib <= p; ++ib) {
                    double va = B[ia], vb = B[ib], da = dB[ia], db = dB[ib];
                    double* dst = &AB(ab, ldab, p, p, ia + first, ib + first);
                    if (kind == 0) *dst += va * vb * w[k] * J[e];
                    else if (kind == 1) *dst += da * db * w[k] * J[e];
                    else if (kind == 2) *dst += va * db * w[k] * J[e];
                    else *dst += (va * vb + h * da * db) * w[k] * J[e];
                }
            }
        }
    }
    int last = n - 1;
    for (int side = 0; side < 2; ++side) {
        if (!(fix & (1 << side))) continue;
        int kk = side == 0 ? 0 : last;
        int lo = kk - p < 0 ? 0 : kk - p, hi = kk + p > last ? last : kk + p;
        for (int i = lo; i <= hi; ++i) AB(ab, ldab, p, p, kk, i) = 0;
        AB(ab, ldab, p, p, kk, kk) = 1;
    }
    free(bt); free(xq); free(J); free(fd);
    return 0;
}

/* ---- LAPACK 3.9 DGBTF2 (unblocked; DGBTRF uses it whenever kl < 32) ------------------------- */
int orc_dgbtrf(int n, int kl, int ku, double* ab, int ldab, int* ipiv) {
    int kv = ku + kl, info = 0;
#define A1(i, j) ab[(size_t) ((j) -1) * ldab + ((i) -1)] /* 1-based band row i, column j */
    for (int j = ku + 2; j <= (kv < n ? kv : n); ++j)
        for (int i = kv - j + 2; i <= kl; ++i) A1(i, j) = 0;
    int ju = 1;
    for (int j = 1; j <= n; ++j) {
        if (j + kv <= n)
            for (int i = 1; i <= kl; ++i) A1(i, j + kv) = 0;
        int km = kl < n - j ? kl : n - j;
        int jp = 1; /* idamax over km+1 entries starting at (kv+1, j): first maximum */
        double best = fabs(A1(kv + 1, j));
        for (int i = 2; i <= km + 1; ++i) {
            double v = fabs(A1(kv + i, j));
            if (v > best) { best = v; jp = i; }
        }
        ipiv[j - 1] = jp + j - 1;
        if (A1(kv + jp, j) != 0) {
            int c = j + ku + jp - 1;
            if (c > n) c = n;
            if (c > ju) ju = c;
            if (jp != 1) { /* dswap(ju-j+1, A(kv+jp,j), ldab-1, A(kv+1,j), ldab-1) */
                for (int t = 0; t < ju - j + 1; ++t) {
                    double* x = &ab[(size_t) (j - 1) * ldab + (kv + jp - 1) + (size_t) t * (ldab - 1)];
                    double* y = &ab[(size_t) (j - 1) * ldab + (kv + 1 - 1) + (size_t) t * (ldab - 1)];
                    double tmp = *x; *x = *y; *y = tmp;
                }
            }
            if (km > 0) {
                double r = 1.0 / A1(kv + 1, j);
                for (int i = 1; i <= km; ++i) A1(kv + 1 + i, j) *= r;
                if (ju > j) { /* dger(km, ju-j, -1, l, 1, U(j, j+1..ju), ldab-1, A(j+1.., j+1..ju), ldab-1) */
                    for (int c = j + 1; c <= ju; ++c) {
                        double uc = A1(kv + 1 + j - c, c); /* full-matrix element (j, c) */
                        if (uc != 0) {
                            double t = -uc;
                            for (int i = 1; i <= km; ++i) A1(kv + 1 + j + i - c, c) += A1(kv + 1 + i, j) * t;
                        }
                    }
                }
            }
        } else if (info == 0) {
            info = j;
        }
    }
#undef A1
    return info;
}

/* ---- LAPACK DGBTRS('N') = pivoted L sweep (dswap+dger per column) then DTBSV('U','N','N') ---- */
int orc_dgbtrs(int n, int kl, int ku, int nrhs, const double* ab, int ldab, const int* ipiv,
               double* b, int ldb) {
    int kd = ku + kl; /* 0-based row of the diagonal */
    for (int r = 0; r < nrhs; ++r) {
        double* x = b + (size_t) r * ldb;
        if (kl > 0) {
            for (int j = 0; j < n - 1; ++j) {
                int lm = kl < n - 1 - j ? kl : n - 1 - j;
                int l = ipiv[j] - 1;
                if (l != j) { double t = x[l]; x[l] = x[j]; x[j] = t; }
                double xj = x[j];
                const double* col = ab + (size_t) j * ldab + kd + 1;
                double t = -xj; /* dger, alpha = -1 */
                for (int i = 0; i < lm; ++i) x[j + 1 + i] += col[i] * t;
            }
        }
        for (int j = n - 1; j >= 0; --j) {
            if (x[j] != 0) {
                const double* col = ab + (size_t) j * ldab;
                x[j] = x[j] / col[kd];
                double t = x[j];
                int lo = j - kd > 0 ? j - kd : 0;
                for (int i = j - 1; i >= lo; --i) x[i] -= t * col[kd + i - j];
            }
        }
    }
    return 0;
}

void orc_band_matvec(int n, int kl, int ku, const double* ab, int ldab, const double* x, double* y) {
    for (int i = 0; i < n; ++i) {
        double s = 0;
        int lo = i - kl > 0 ? i - kl : 0, hi = i + ku < n - 1 ? i + ku : n - 1;
        for (int j = lo; j <= hi; ++j) s += AB(ab, ldab, kl, ku, i, j) * x[j];
        y[i] = s;
    }
}

/* ---- tensor rotation + ADS solve ------------------------------------------------------------ */

/* include/ads/lin/tensor/cyclic_transpose.hpp:19-63: out(i1,..,i_{R-1},i0) = in(i0,..,i_{R-1}),
 * column-major (first index fastest, include/ads/util/multi_array/ordering/reverse.hpp:28-31) */
void orc_cyclic_transpose(int ndim, const int* n, const double* in, double* out) {
    if (ndim == 2) {
        for (int i1 = 0; i1 < n[1]; ++i1)
            for (int i0 = 0; i0 < n[0]; ++i0)
                out[i1 + (size_t) n[1] * i0] = in[i0 + (size_t) n[0] * i1];
    } else if (ndim == 3) {
        for (int i2 = 0; i2 < n[2]; ++i2)
            for (int i1 = 0; i1 < n[1]; ++i1)
                for (int i0 = 0; i0 < n[0]; ++i0)
                    out[i1 + (size_t) n[1] * (i2 + (size_t) n[2] * i0)] =
                        in[i0 + (size_t) n[0] * (i1 + (size_t) n[1] * i2)];
    } else {
        memcpy(out, in, sizeof(double) * (size_t) n[0]);
    }
}

/* include/ads/solver.hpp:35-41 (per axis: solve_with_factorized, then rotate), :148-166 */
int orc_ads_solve(int ndim, const int* n, const int* kl, const int* ku, const double* const* ab,
                  const int* const* ipiv, double* rhs, double* buf) {
    size_t N = 1;
    int sz[3];
    for (int d = 0; d < ndim; ++d) { N *= (size_t) n[d]; sz[d] = n[d]; }
    if (ndim == 1) return orc_dgbtrs(n[0], kl[0], ku[0], 1, ab[0], 2 * kl[0] + ku[0] + 1, ipiv[0], rhs, n[0]);
    double* cur = rhs;
    double* other = buf;
    for (int d = 0; d < ndim; ++d) {
        int nrhs = (int) (N / (size_t) sz[0]);
        orc_dgbtrs(sz[0], kl[d], ku[d], nrhs, ab[d], 2 * kl[d] + ku[d] + 1, ipiv[d], cur, sz[0]);
        orc_cyclic_transpose(ndim, sz, cur, other);
        int f = sz[0];
        for (int k = 0; k + 1 < ndim; ++k) sz[k] = sz[k + 1];
        sz[ndim - 1] = f;
        double* t = cur; cur = other; other = t;
    }
    /* the reference swaps the tensor objects when ndim is odd (solver.hpp:152-159); the caller
     * of this restatement always wants the answer in `rhs` */
    if (cur != rhs) memcpy(rhs, cur, sizeof(double) * N);
    return 0;
}

/* ---- element quadrature RHS ----------------------------------------------------------------- */

typedef struct {
    int p, q, elements, n;
    double *b, *xq, *J;
    double w[MAXQ];
    int* fd;
} axis_t;

static void axis_init(axis_t* a, int p, int elements) {
    a->p = p; a->q = p + 1; a->elements = elements; a->n = elements + p;
    a->b = (double*) malloc(sizeof(double) * (size_t) elements * a->q * 2 * (p + 1));
    a->xq = (double*) malloc(sizeof(double) * (size_t) elements * a->q);
    a->J = (double*) malloc(sizeof(double) * (size_t) elements);
    a->fd = (int*) malloc(sizeof(int) * (size_t) elements);
    orc_basis_tables(p, elements, 0.0, 1.0, a->q, 1, a->b, a->xq, a->w, a->J, a->fd);
}

static void axis_free(axis_t* a) { free(a->b); free(a->xq); free(a->J); free(a->fd); }

#define BV(ax, e, k, d, i) (ax)->b[(((size_t) (e) * (ax)->q + (k)) * 2 + (d)) * ((ax)->p + 1) + (i)]

/* examples/scalability/test3d.hpp:58-64 and test2d.hpp:49-54 */
static double forcing3(double x, double y, double z) {
    double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
    double r = sqrt(dx * dx + dy * dy + dz * dz);
    return exp(-r) + 1 + cos(M_PI * x) * cos(M_PI * y) * cos(M_PI * z);
}
static double forcing2(double x, double y) {
    double dx = x - 0.5, dy = y - 0.5;
    double r = sqrt(dx * dx + dy * dy);
    return exp(-r) + 1 + cos(M_PI * x) * cos(M_PI * y);
}

/* 3-D: examples/heat/heat_3d.hpp:49-67, examples/scalability/test3d.hpp:66-95 with the helpers
 * of include/ads/simulation/simulation_3d.hpp:83-128,:138-145 */
static void rhs3(const axis_t* X, const axis_t* Y, const axis_t* Z, const orc_form* f,
                 const double* up, double* rhs) {
    int p = X->p, q = X->q, m = p + 1;
    size_t nx = (size_t) X->n, ny = (size_t) Y->n, N = nx * ny * (size_t) Z->n;
    memset(rhs, 0, sizeof(double) * N);
    double U[(MAXP + 1) * (MAXP + 1) * (MAXP + 1)];
    for (int ex = 0; ex < X->elements; ++ex)
    for (int ey = 0; ey < Y->elements; ++ey)
    for (int ez = 0; ez < Z->elements; ++ez) {
        double J = X->J[ex] * Y->J[ey] * Z->J[ez];
        int fx = X->fd[ex], fy = Y->fd[ey], fz = Z->fd[ez];
        if (f->scatter) memset(U, 0, sizeof(double) * (size_t) (m * m * m));
        for (int kx = 0; kx < q; ++kx)
        for (int ky = 0; ky < q; ++ky)
        for (int kz = 0; kz < q; ++kz) {
            double w = X->w[kx] * Y->w[ky] * Z->w[kz];
            /* eval_fun (simulation_3d.hpp:120-128) */
            double uv = 0, ux = 0, uy = 0, uz = 0;
            for (int ax = 0; ax < m; ++ax)
            for (int ay = 0; ay < m; ++ay)
            for (int az = 0; az < m; ++az) {
                double c = up[(fx + ax) + nx * ((fy + ay) + ny * (size_t) (fz + az))];
                double B1 = BV(X, ex, kx, 0, ax), B2 = BV(Y, ey, ky, 0, ay), B3 = BV(Z, ez, kz, 0, az);
                double d1 = BV(X, ex, kx, 1, ax), d2 = BV(Y, ey, ky, 1, ay), d3 = BV(Z, ez, kz, 1, az);
                uv += c * (B1 * B2 * B3);
                ux += c * (d1 * B2 * B3);
                uy += c * (B1 * d2 * B3);
                uz += c * (B1 * B2 * d3);
            }
            double fv = 0;
            if (f->forcing) fv = forcing3(X->xq[ex * q + kx], Y->xq[ey * q + ky], Z->xq[ez * q + kz]);
            for (int ax = 0; ax < m; ++ax)
            for (int ay = 0; ay < m; ++ay)
            for (int az = 0; az < m; ++az) {
                double B1 = BV(X, ex, kx, 0, ax), B2 = BV(Y, ey, ky, 0, ay), B3 = BV(Z, ez, kz, 0, az);
                double d1 = BV(X, ex, kx, 1, ax), d2 = BV(Y, ey, ky, 1, ay), d3 = BV(Z, ez, kz, 1, az);
                double v = B1 * B2 * B3, vx = d1 * B2 * B3, vy = B1 * d2 * B3, vz = B1 * B2 * d3;
                double gp;
                if (f->grad_mask == 7) gp = ux * vx + uy * vy + uz * vz;
                else {
                    gp = 0;
                    if (f->grad_mask & 1) gp += ux * vx;
                    if (f->grad_mask & 2) gp += uy * vy;
                    if (f->grad_mask & 4) gp += uz * vz;
                }
                double val = f->forcing ? uv * v - f->tau * (gp - fv) : uv * v - f->tau * gp;
                if (f->scatter) U[ax + m * (ay + m * az)] += val * w * J;
                else rhs[(fx + ax) + nx * ((fy + ay) + ny * (size_t) (fz + az))] += val * w * J;
            }
        }
        if (f->scatter)
            for (int ax = 0; ax < m; ++ax)
            for (int ay = 0; ay < m; ++ay)
            for (int az = 0; az < m; ++az)
                rhs[(fx + ax) + nx * ((fy + ay) + ny * (size_t) (fz + az))] += U[ax + m * (ay + m * az)];
    }
}

/* 2-D: examples/heat/heat_2d.hpp:80-106, examples/implicit/implicit.hpp:132-182,
 * examples/scalability/test2d.hpp:56-85; helpers include/ads/simulation/simulation_2d.hpp:88-140 */
static void rhs2(const axis_t* X, const axis_t* Y, const orc_form* f, const double* up, double* rhs) {
    int p = X->p, q = X->q, m = p + 1;
    size_t nx = (size_t) X->n, N = nx * (size_t) Y->n;
    memset(rhs, 0, sizeof(double) * N);
    double U[(MAXP + 1) * (MAXP + 1)];
    for (int ex = 0; ex < X->elements; ++ex)
    for (int ey = 0; ey < Y->elements; ++ey) {
        double J = X->J[ex] * Y->J[ey];
        int fx = X->fd[ex], fy = Y->fd[ey];
        memset(U, 0, sizeof(double) * (size_t) (m * m));
        for (int kx = 0; kx < q; ++kx)
        for (int ky = 0; ky < q; ++ky) {
            double w = X->w[kx] * Y->w[ky];
            double uv = 0, ux = 0, uy = 0;
            for (int ax = 0; ax < m; ++ax)
            for (int ay = 0; ay < m; ++ay) {
                double c = up[(fx + ax) + nx * (size_t) (fy + ay)];
                double B1 = BV(X, ex, kx, 0, ax), B2 = BV(Y, ey, ky, 0, ay);
                double d1 = BV(X, ex, kx, 1, ax), d2 = BV(Y, ey, ky, 1, ay);
                uv += c * (B1 * B2);
                ux += c * (d1 * B2);
                uy += c * (B1 * d2);
            }
            double fv = 0;
            if (f->forcing) fv = forcing2(X->xq[ex * q + kx], Y->xq[ey * q + ky]);
            for (int ax = 0; ax < m; ++ax)
            for (int ay = 0; ay < m; ++ay) {
                double B1 = BV(X, ex, kx, 0, ax), B2 = BV(Y, ey, ky, 0, ay);
                double d1 = BV(X, ex, kx, 1, ax), d2 = BV(Y, ey, ky, 1, ay);
                double v = B1 * B2, vx = d1 * B2, vy = B1 * d2;
                double gp;
                if (f->grad_mask == 3) gp = ux * vx + uy * vy;
                else if (f->grad_mask == 1) gp = ux * vx;
                else if (f->grad_mask == 2) gp = uy * vy;
                else gp = 0;
                double val = f->forcing ? uv * v - f->tau * (gp - fv) : uv * v - f->tau * gp;
                U[ax + m * ay] += val * w * J;
            }
        }
        for (int ax = 0; ax < m; ++ax)
        for (int ay = 0; ay < m; ++ay)
            rhs[(fx + ax) + nx * (size_t) (fy + ay)] += U[ax + m * ay];
    }
}

int orc_compute_rhs(const orc_grid* g, const orc_form* f, const double* u_prev, double* rhs) {
    axis_t A;
    axis_init(&A, g->p, g->elements);
    if (g->ndim == 3) rhs3(&A, &A, &A, f, u_prev, rhs);
    else rhs2(&A, &A, f, u_prev, rhs);
    axis_free(&A);
    return 0;
}

/* ---- L2 projection (include/ads/projection.hpp:12-153) -------------------------------------- */

static double init_heat3d(double x, double y, double z) { /* examples/heat/heat_3d.hpp:22-28 */
    double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
    double r2 = fmin(8 * (dx * dx + dy * dy + dz * dz), 1.0);
    return (r2 - 1) * (r2 - 1) * (r2 + 1) * (r2 + 1);
}
static double init_implicit3d(double x, double y, double z) { /* 3-D twin of implicit.hpp:38-43 */
    double dx = x - 0.5, dy = y - 0.5, dz = z - 0.5;
    double r2 = fmin(12 * (dx * dx + dy * dy + dz * dz), 1.0);
    return (r2 - 1) * (r2 - 1) * (r2 + 1) * (r2 + 1);
}
static double init_implicit2d(double x, double y) { /* examples/implicit/implicit.hpp:38-43 */
    double dx = x - 0.5, dy = y - 0.5;
    double r2 = fmin(12 * (dx * dx + dy * dy), 1.0);
    return (r2 - 1) * (r2 - 1) * (r2 + 1) * (r2 + 1);
}
static double init_zero2d(double x, double y) { (void) x; (void) y; return 0; } /* heat_2d.hpp:32-38 */

static void project3(const axis_t* X, const axis_t* Y, const axis_t* Z,
                     double (*fn)(double, double, double), double* u) {
    int q = X->q, m = X->p + 1;
    size_t nx = (size_t) X->n, ny = (size_t) Y->n;
    memset(u, 0, sizeof(double) * nx * ny * (size_t) Z->n);
    for (int ex = 0; ex < X->elements; ++ex)
    for (int ey = 0; ey < Y->elements; ++ey)
    for (int ez = 0; ez < Z->elements; ++ez) {
        double J = 1; J *= X->J[ex]; J *= Y->J[ey]; J *= Z->J[ez];
        for (int kx = 0; kx < q; ++kx)
        for (int ky = 0; ky < q; ++ky)
        for (int kz = 0; kz < q; ++kz) {
            double w = 1; w *= X->w[kx]; w *= Y->w[ky]; w *= Z->w[kz];
            double fv = fn(X->xq[ex * q + kx], Y->xq[ey * q + ky], Z->xq[ez * q + kz]);
            for (int ax = 0; ax < m; ++ax)
            for (int ay = 0; ay < m; ++ay)
            for (int az = 0; az < m; ++az) {
                double B = 1;
                B *= BV(X, ex, kx, 0, ax); B *= BV(Y, ey, ky, 0, ay); B *= BV(Z, ez, kz, 0, az);
                u[(X->fd[ex] + ax) + nx * ((Y->fd[ey] + ay) + ny * (size_t) (Z->fd[ez] + az))] += fv * B * w * J;
            }
        }
    }
}

static void project2(const axis_t* X, const axis_t* Y, double (*fn)(double, double), double* u) {
    int q = X->q, m = X->p + 1;
    size_t nx = (size_t) X->n;
    memset(u, 0, sizeof(double) * nx * (size_t) Y->n);
    for (int ex = 0; ex < X->elements; ++ex)
    for (int ey = 0; ey < Y->elements; ++ey) {
        double J = 1; J *= X->J[ex]; J *= Y->J[ey];
        for (int kx = 0; kx < q; ++kx)
        for (int ky = 0; ky < q; ++ky) {
            double w = 1; w *= X->w[kx]; w *= Y->w[ky];
            double fv = fn(X->xq[ex * q + kx], Y->xq[ey * q + ky]);
            for (int ax = 0; ax < m; ++ax)
            for (int ay = 0; ay < m; ++ay) {
                double B = 1; B *= BV(X, ex, kx, 0, ax); B *= BV(Y, ey, ky, 0, ay);
                u[(X->fd[ex] + ax) + nx * (size_t) (Y->fd[ey] + ay)] += fv * B * w * J;
            }
        }
    }
}

static void project1_sin(const axis_t* Y, double* buf) { /* heat_2d.hpp:41-42 */
    int q = Y->q, m = Y->p + 1;
    memset(buf, 0, sizeof(double) * (size_t) Y->n);
    for (int e = 0; e < Y->elements; ++e) {
        double J = 1; J *= Y->J[e];
        for (int k = 0; k < q; ++k) {
            double w = 1; w *= Y->w[k];
            double fv = sin(Y->xq[e * q + k] * M_PI);
            for (int a = 0; a < m; ++a) {
                double B = 1; B *= BV(Y, e, k, 0, a);
                buf[Y->fd[e] + a] += fv * B * w * J;
            }
        }
    }
}

/* ---- whole problems ------------------------------------------------------------------------- */

typedef struct {
    int n, kl, ku, ldab;
    double* ab;
    int* ipiv;
} fac_t;

static void fac_make(fac_t* F, int kind, int p, int elements, double h, int fix, int* shared_ipiv) {
    F->n = elements + p; F->kl = p; F->ku = p; F->ldab = 3 * p + 1;
    F->ab = (double*) malloc(sizeof(double) * (size_t) F->ldab * F->n);
    F->ipiv = shared_ipiv ? shared_ipiv : (int*) malloc(sizeof(int) * (size_t) F->n);
    orc_matrix_1d(kind, p, elements, 0.0, 1.0, h, fix, F->ab);
    orc_dgbtrf(F->n, F->kl, F->ku, F->ab, F->ldab, F->ipiv);
}

static void solve_nd(int ndim, const fac_t* const* F, double* u, double* buf, double* tsolve) {
    int n[3], kl[3], ku[3];
    const double* ab[3];
    const int* ip[3];
    for (int d = 0; d < ndim; ++d) {
        n[d] = F[d]->n; kl[d] = F[d]->kl; ku[d] = F[d]->ku; ab[d] = F[d]->ab; ip[d] = F[d]->ipiv;
    }
    double t0 = now_s();
    orc_ads_solve(ndim, n, kl, ku, ab, ip, u, buf);
    if (tsolve) *tsolve += now_s() - t0;
}

int orc_run(int problem, int p, int elements, double dt, int nsteps, int init_mode, int stage,
            double* u, double* timings) {
    int ndim = (problem == 0 || problem == 3 || problem == 5) ? 3 : 2;
    axis_t A;
    axis_init(&A, p, elements);
    size_t n = (size_t) A.n, N = ndim == 3 ? n * n * n : n * n;
    double* up = (double*) malloc(sizeof(double) * N);
    double* buf = (double*) malloc(sizeof(double) * N);
    double t_rhs = 0, t_solve = 0;

    /* matrices (prepare_matrices of each example) */
    int fix_x = (problem == 1 || problem == 3 || problem == 4) ? 1 : 0; /* x.fix_left() */
    fac_t Mx, M, Kx, Ky, Kz;
    memset(&Kx, 0, sizeof Kx); memset(&Ky, 0, sizeof Ky); memset(&Kz, 0, sizeof Kz);
    fac_make(&Mx, 0, p, elements, 0, fix_x, NULL);
    fac_make(&M, 0, p, elements, 0, 0, NULL);
    if (problem == 2) {
        /* implicit.hpp:77-82: Kx/Ky are factorised INTO x.ctx / y.ctx, overwriting M's pivots */
        fac_make(&Kx, 3, p, elements, 0.5 * dt, 0, Mx.ipiv);
        fac_make(&Ky, 3, p, elements, 0.5 * dt, 0, M.ipiv);
    }
    if (problem == 5) {
        fac_make(&Kx, 3, p, elements, dt / 3.0, 0, NULL);
        fac_make(&Ky, 3, p, elements, dt / 3.0, 0, NULL);
        fac_make(&Kz, 3, p, elements, dt / 3.0, 0, NULL);
    }
    const fac_t* FM[3] = {&Mx, &M, &M};
    double* dirichlet = NULL;
    if (problem == 1) {
        dirichlet = (double*) malloc(sizeof(double) * n);
        project1_sin(&A, dirichlet);
    }

#define SOLVE_M()                                                                  \
    do {                                                                           \
        if (problem == 1) for (size_t i = 0; i < n; ++i) u[0 + n * i] = dirichlet[i]; \
        solve_nd(ndim, FM, u, buf, &t_solve);                                      \
    } while (0)

    /* before() */
    if (init_mode == 1) {
        if (problem == 0) { project3(&A, &A, &A, init_heat3d, u); SOLVE_M(); }
        else if (problem == 1) { project2(&A, &A, init_zero2d, u); SOLVE_M(); }
        else if (problem == 2) { project2(&A, &A, init_implicit2d, u); SOLVE_M(); }
        else if (problem == 5) { project3(&A, &A, &A, init_implicit3d, u); SOLVE_M(); }
        else { for (size_t i = 0; i < N; ++i) u[i] = 1; SOLVE_M(); }
    }

    orc_form f;
    f.tau = dt; f.grad_mask = ndim == 3 ? 7 : 3; f.forcing = 0; f.scatter = 1;
    if (problem == 0) f.scatter = 0;
    if (problem == 3 || problem == 4) f.forcing = 1;
    if (problem == 2) f.tau = 0.5 * dt;
    if (problem == 5) f.tau = dt / 3.0;

    int nsub = problem == 2 ? 2 : problem == 5 ? 3 : 1;
    double t0 = now_s();
    int steps = stage >= 1 ? 1 : nsteps;
    for (int it = 0; it < steps; ++it) {
        for (int s = 0; s < nsub; ++s) {
            if (stage >= 1 && s != stage - 1) continue;
            memcpy(up, u, sizeof(double) * N); /* swap(u, u_prev): u_prev := u, u gets overwritten */
            if (problem == 2) f.grad_mask = s == 0 ? 2 : 1;            /* implicit.hpp:148,:174 */
            if (problem == 5) f.grad_mask = 7 & ~(1 << s);
            double a = now_s();
            if (ndim == 3) rhs3(&A, &A, &A, &f, up, u); else rhs2(&A, &A, &f, up, u);
            t_rhs += now_s() - a;
            if (stage >= 1) break;
            if (problem == 2) {
                const fac_t* F1[2] = {&Kx, &M};
                const fac_t* F2[2] = {&Mx, &Ky};
                solve_nd(2, s == 0 ? F1 : F2, u, buf, &t_solve);
            } else if (problem == 5) {
                const fac_t* F[3] = {&Mx, &M, &M};
                if (s == 0) F[0] = &Kx;
                if (s == 1) F[1] = &Ky;
                if (s == 2) F[2] = &Kz;
                solve_nd(3, F, u, buf, &t_solve);
            } else {
                SOLVE_M();
            }
        }
    }
    if (timings) { timings[0] = now_s() - t0; timings[1] = t_rhs; timings[3] = t_solve; }
#undef SOLVE_M
    free(up); free(buf); free(dirichlet);
    free(Mx.ab); free(Mx.ipiv); free(M.ab); free(M.ipiv);
    free(Kx.ab); free(Ky.ab); free(Kz.ab);
    if (problem == 5) { free(Kx.ipiv); free(Ky.ipiv); free(Kz.ipiv); }
    axis_free(&A);
    return 0;
}

int orc_project_init(int problem, int p, int elements, double* u) {
    axis_t A;
    axis_init(&A, p, elements);
    if (problem == 0) project3(&A, &A, &A, init_heat3d, u);
    else if (problem == 5) project3(&A, &A, &A, init_implicit3d, u);
    else if (problem == 2) project2(&A, &A, init_implicit2d, u);
    else { axis_free(&A); return -1; }
    axis_free(&A);
    return 0;
}
