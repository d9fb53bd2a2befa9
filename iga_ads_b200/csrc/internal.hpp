// internal.hpp -- declarations shared by the translation units of libadsb200.so (not installed).
#ifndef ADSB_INTERNAL_HPP
#define ADSB_INTERNAL_HPP

#include <string>
#include <vector>

namespace adsb {

int fail(int code, const std::string& msg);
extern thread_local std::string g_last_error;  // what adsb_last_error() returns

int gauss_rule(int q, double* x, double* w);
int make_knots(int p, int elements, double a, double b, double* knot);
int find_span(double x, const double* knot, int knot_size, int p);
void basis_ders(int span, double x, const double* knot, int p, int ders, double* out);
int basis_tables(int p, int elements, double a, double b, int q, int ders, double* bt, double* xq,
                 double* w, double* J, int* first_dof);
int matrix_from_tables(int kind, double h, int p, int elements, int q, int ders, const double* bt,
                       const double* w, const double* J, double* ab);
int matrix_1d(int kind, int p, int elements, double a, double b, double h, int fix, double* ab);
int band_factorize(int n, int kl, int ku, double* ab, int ldab, int* ipiv);
// Same matrix, eliminated again without row interchanges (host_setup.cpp); false: not safe, keep the pivoted factor
bool refactor_without_pivoting(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, std::vector<double>& out);

constexpr int SWEEP_MAX_DEPTH = 6;  // longest parallel state chain; beyond it the kernel chains sequentially

// Host image of one factor prepared for the chunk-parallel sweep kernel (see host_setup.cpp).
struct SweepPlan {
    int n = 0, KL = 0, KD = 0, piv = 0, CH = 0, R = 0, SC = 0, ST = 0, rows = 0;
    int LF = 0, LB = 0, LC = 0;  // doubles per column in cfF / cfB / cfC (even, 16 B aligned records)
    int DF = 1, DB = 1, seq = 0; // chain depths (forward / backward); seq: use the sequential chain
    std::vector<int> pv;         // [rows]      pivot row offset of column j (0..KL)
    std::vector<double> cfF;     // [rows][LF]  multipliers of column j (rows j+1..j+KL)
    std::vector<double> cfB;     // [rows][LB]  U(j, j+1..j+KD), then 1 / U(j,j)
    std::vector<double> cfC;     // [rows][LC]  Psi(j, 0..KD) (response to t), then Xi(j, 0..KL) (to delta)
    std::vector<double> T;       // [SC][KL][KL] forward state transfer across chunk c
    std::vector<double> Rm;      // [SC][KD][KD] first KD rows of Psi of chunk c
    std::vector<double> W;       // [SC][MAX_DEPTH-1][KL][KL]  W_{c,d}, d = 2..  (products of T)
    std::vector<double> V;       // [SC][MAX_DEPTH-1][KD][KD]  V_{c,d}, d = 2..  (products of Rm)
};
int build_sweep_plan(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, int ch, int group,
                     SweepPlan& plan, bool force_piv = false);

// Segmented substitution (host_setup.cpp): a line is cut into S segments [bounds[s], bounds[s+1]); every
// segment is solved on its own (the same factor restricted to the segment, zero incoming states) and the
// exact dgbtrs result is recovered from KL + KD boundary values per segment and line.  Segments are the
// slabs of a sharded run (one per GPU) or the pieces of a line too long for one CTA.
struct SegPlan {
    int n = 0, KL = 0, KD = 0, piv = 0, S = 0;
    int DF = 1, DB = 1;        // chain depths: din_s uses segments s-1 .. s-DF, tin_s uses s+1 .. s+DB
    std::vector<int> bounds;   // [S+1]
    std::vector<double> E;     // [S][KL][KL]      Dseg_s = E_s * xhat_s[last KL rows]
    std::vector<double> Wf;    // [S][DF][KL][KL]  din_s = sum_d Wf[s][d-1] * Dseg_{s-d}   (Wf[s][0] = I)
    std::vector<double> Vb;    // [S][DB][KD][KD]  tin_s = sum_d Vb[s][d-1] * X_{s+d}      (Vb[s][0] = I)
    std::vector<double> XiF;   // [S][KD][KL]      X_s = xhat_s[first KD rows] + XiF_s * din_s
    std::vector<double> cf;    // [n][KD+KL]       x_j = xhat_j + Psi(j,:) tin_s + Xi(j,:) din_s
};
// Balanced segment boundaries that no row interchange of the factor crosses (and, when align > 1, that are
// multiples of `align` where possible).  Returns ADSB_EINVAL when no such cut exists near a target.
int pick_segment_bounds(int n, int kl, const int* ipiv, int S, int align, int min_rows, int* bounds);
int build_segment_plan(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, int S, const int* bounds,
                       double tol, SegPlan& plan);

}  // namespace adsb

#endif
