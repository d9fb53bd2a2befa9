// internal.hpp -- declarations shared by the translation units of libadsb200.so (not installed).
#ifndef ADSB_INTERNAL_HPP
#define ADSB_INTERNAL_HPP

#include <string>
#include <vector>

namespace adsb {

int fail(int code, const std::string& msg);

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

// Host image of one factor prepared for the chunk-parallel sweep kernel (see host_setup.cpp).
struct SweepPlan {
    int n = 0, KL = 0, KD = 0, piv = 0, CH = 0, S = 0;
    std::vector<double> Lm;    // [n][KL]   multipliers of column j (rows j+1..j+KL), zero padded
    std::vector<int> pv;       // [n]       pivot row offset of column j (0..KL)
    std::vector<double> Ut;    // [n][KD]   U(j, j+1..j+KD), zero padded
    std::vector<double> rinv;  // [n]       1 / U(j,j)
    std::vector<double> Phi;   // [n][KL]   forward response of row j to the chunk's incoming state
    std::vector<double> Psi;   // [n][KD]   backward response of row j to the chunk's incoming state
    std::vector<double> T;     // [S][KL][KL] forward state transfer across a whole chunk
};
int build_sweep_plan(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, int ch,
                     SweepPlan& plan);

}  // namespace adsb

#endif
