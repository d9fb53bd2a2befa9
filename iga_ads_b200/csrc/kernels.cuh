// kernels.cuh -- device-side argument blocks and launcher prototypes (internal to libadsb200.so).
#ifndef ADSB_KERNELS_CUH
#define ADSB_KERNELS_CUH

#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace adsb {

// ADSB_PDL=0 switches programmatic dependent launch off (every kernel then starts after its predecessor ended)
inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("ADSB_PDL");
        return !e || atoi(e) != 0;
    }();
    return on;
}
// Launch with the programmatic-stream-serialization attribute (see tma.cuh: pdl_wait / pdl_launch).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                             Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

constexpr int SWEEP_CH = 18;  // columns per chunk (one chunk per thread, held in registers)
constexpr int SWEEP_CH_DIST = 12;  // shorter chunks for the fused distributed sweep (kernels_sweep_dist.cu)
constexpr int SWEEP_RL = 2;   // lines per thread: they share every coefficient load and interleave for ILP
constexpr int SWEEP_MAX_DEPTH_DEV = 6;
// Coefficient records (multipliers, U rows, Psi|Xi) of one column: `sweep_rec` doubles are loaded (even: 128-bit
// loads), records are `sweep_pitch` doubles apart.  (A wider pitch that keeps the 4 chunks a warp touches off
// each other's banks was measured: no gain, and its 17 KB of shared memory cost a ring slot.)
__host__ __device__ constexpr int sweep_rec(int len) { return len + (len & 1); }
__host__ __device__ constexpr int sweep_pitch(int len) { return sweep_rec(len); }
constexpr int SWEEP_MAX_BOXES = 32;      // TMA boxes per tile and direction (tile-streaming sweep kernel)  // == SWEEP_MAX_DEPTH of internal.hpp

// One factorised band matrix, prepared by build_sweep_plan (host_setup.cpp).
struct SweepFactor {
    const double* cfF;  // [rows][LF]  multipliers
    const double* cfB;  // [rows][LB]  U rows + reciprocal diagonal
    const double* cfC;  // [rows][LC]  Psi | Xi
    const int* pv;      // [rows]
    const double* T;    // [SC][KL][KL]
    const double* Rm;   // [SC][KD][KD]
    const double* W;    // [SC][MAX_DEPTH-1][KL][KL]
    const double* V;    // [SC][MAX_DEPTH-1][KD][KD]
    int n, ST, SC, KL, KD, piv, DF, DB, seq;
    // the seven tables above are pieces of ONE allocation (each padded to an even number of doubles), so a
    // persistent CTA stages them with a single bulk copy: blob = cfF, blob_doubles in all, off[i] = start of
    // cfB, cfC, T, Rm, W, V inside it
    int blob_doubles;
    int off[6];
};

// Lines of one sweep.  A line is addressed as base(l0, l1) + off(j):
//   base = l0*s0 + l1*s1, off(j) = off[j] if off != nullptr else j*sj.
// STRIDED mode wants s0 == 1 (lanes run along l0); CONTIG mode wants sj == 1.
struct SweepGeom {
    const double* in;
    double* out;
    const long long* off_in;
    const long long* off_out;
    long long sj_in, sj_out;
    long long s0_in, s0_out;
    long long s1_in, s1_out;
    int L0, L1;
    int pitch;  // CONTIG: shared-memory row pitch in doubles (even)
    int bulk;   // CONTIG: rows are 16 B aligned on both sides -> TMA bulk copies
    int max_ctas;  // persistent kernels: at most this many CTAs (0: one per SM)
    int reverse;   // tile kernel: walk the tiles from the last one backwards (start on what the previous kernel left in L2)
    int pad_ok;    // CONTIG, odd n: every line is followed by a pad element inside the allocation (managed
                   // tensors, in place): bulk copies may move n+1 doubles so that they stay 16 B multiples
};

// Tile-streaming variant (kernels_sweep_tile.cu): persistent CTAs, TMA-fed shared-memory ring.
struct SweepTileGeom {
    const double* in;
    double* out;
    long long s0_in, s1_in, s0_out, s1_out;  // CONTIG: element strides of the two line indices
    int L0, L1;        // lines: l0 in [0, L0) (x for the strided sweeps), l1 in [0, L1)
    int nb0, ntiles;   // tiles of 16 lines along l0; nb0 * L1 tiles in all
    int pitch;         // CONTIG: shared row pitch of a line (doubles)
    int ncopy;         // CONTIG: doubles moved per line (n, or n+1 when n is odd and the line has a pad element)
    // STRIDED: a tile is moved as boxes of whole rows, one tensor map per box (exact extents, so a
    // box never spills over its neighbour): maps[0 .. nbox_in) load, maps[nbox_in .. +nbox_out) store
    const void* maps;  // CUtensorMap array in global memory
    int nbox_in, nbox_out;
    int row0_in[SWEEP_MAX_BOXES], row0_out[SWEEP_MAX_BOXES];  // first tile row of each box
    int load_bytes;    // sum of the load boxes (mbarrier transaction count)
    int tile_doubles;  // shared-memory doubles per ring slot
    int nbuf;          // ring depth
    long long s_row;   // element stride of the rows (fused distributed sweep: halo stores address planes directly)
    int reverse;       // tiles in descending order
};
// 0: launched; -1: not eligible (caller falls back to launch_sweep); otherwise a cudaError_t.
// off_in_h / off_out_h: optional host row-offset tables (see adsb_sweep_view); they must be piecewise
// linear with few pieces (the block layouts of an all-to-all are).
int launch_sweep_tile(const SweepFactor& F, const SweepGeom& G, bool contig, const long long* off_in_h,
                      const long long* off_out_h, cudaStream_t st);

// nseg segments of the same lines in one launch (kernels_sweep_tile.cu); same return convention
int launch_sweep_tile_multi(int nseg, const SweepFactor* Fs, const SweepGeom* Gs, bool contig, cudaStream_t st);
int sweep_strided_maps(const SweepGeom& G, int n, int NL, const long long* off_in_h, const long long* off_out_h,
                       cudaStream_t st, SweepTileGeom& T);

// returns cudaError_t as int
int launch_sweep(const SweepFactor& F, const SweepGeom& G, bool contig, int NLt, cudaStream_t st);
int sweep_smem_bytes(const SweepFactor& F, bool contig, int NLt, int pitch);

// Collapsed right-hand side: rhs = sum over Kronecker terms of 1-D band operators applied to u.
// op[d] points at coefficient rows [n_d][2p+1] (row i holds A(i, i-p .. i+p), zero outside).
struct RhsOps {
    const double* Mx;  // Gram
    const double* Sx;  // stiffness
    const double* My;
    const double* Sy;
    const double* MSzT;  // column table [n+2p][2][2p+2]: row k+p holds Mz(k-p .. k+p, k), pad, Sz(k-p .. k+p, k), pad
                         // (what input plane k feeds into output planes k-p .. k+p); p zero rows on both sides
    int p[3];
    int n[3];  // global extents
};
struct RhsGeom {
    const double* in;
    double* out;
    const double* forcing;  // same layout as out, or nullptr
    long long si[3], so[3];
    int in_lo[3], in_n[3];    // box covered by `in` in global indices
    int out_lo[3], out_n[3];  // box to produce
    double alpha, beta[3], gamma;
    int max_sms;  // launch heuristics assume this many SMs (0: the whole device)
};
// Optional second stream + two events: the narrow x-remainder kernel is forked onto `side` so it runs
// next to the main kernel (it fills the CTA slots the main grid leaves free) and joined back into st.
struct RhsSide {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
// *nlaunch (optional) receives the number of kernels launched
int launch_rhs_collapsed(int ndim, const RhsOps& ops, const RhsGeom& G, cudaStream_t st, int* nlaunch = nullptr,
                         const RhsSide* side = nullptr);
// TMA-fed 3-D variant (kernels_rhs_tma.cu).  0: launched; -1: not eligible; otherwise a cudaError_t.
int launch_rhs_tma(const RhsOps& ops, const RhsGeom& G, cudaStream_t st, bool narrow = false);
// 2-D y-marching variant (ADSB_RHS2D_MARCH=0 disables it); ops.MSzT must hold the column table of axis 1.
int launch_rhs2d_march(const RhsOps& ops, const RhsGeom& G, cudaStream_t st);
// Encode a 3-D FP64 tiled tensor map (element strides of dims 1 and 2; dim 0 is contiguous) into *map
// (a CUtensorMap, 128 bytes, 64 B aligned).  False when the driver entry point is missing or refuses.
bool encode_tensor_map3(void* map, const double* base, const unsigned long long dims[3],
                        const unsigned long long strides[2], const unsigned box[3]);

// Per-axis quadrature tables on the device (layout of adsb_basis_tables, ders = 1).
struct QuadAxes {
    int ndim;
    int p[3], q[3], ne[3], st[3];  // st = (ders+1)*(p+1) doubles per quadrature point
    const double* bt[3];  // [ne][q][2][p+1]
    const double* xq[3];  // [ne][q]
    const double* w[3];   // [q]
    const double* J[3];   // [ne]
};
// pitch0: doubles between consecutive x rows of `out` (0: dense, n[0])
int launch_project(int src, const QuadAxes& A, double* out, const int lo[3], const int n[3], cudaStream_t st,
                   long long pitch0 = 0);
// f tabulated at the quadrature points of the z-element slab [ez_lo, ez_lo + ez_cnt) (device array)
int launch_project_tab(const QuadAxes& A, const double* tab, int ez_lo, int ez_cnt, int accumulate, double* out,
                       const int n[3], cudaStream_t st, long long pitch0 = 0);
int launch_element_source(int src, const QuadAxes& A, double* G, const int elo[3], const int en[3], cudaStream_t st);
int launch_box_sum(const QuadAxes& A, const double* G, double* out, const int elo[3], const int en[3],
                   const int lo[3], const int n[3], cudaStream_t st, long long pitch0 = 0);

// Pointwise form of the brick quadrature kernel (quadbrick.cuh); mirrors adsb_point_form.
struct PointFormArgs {
    int kind;  // 0 linear (+ advection, built-in source), 1 flow (nonlinear, tabulated coefficient)
    double alpha, beta[3], adv[3], gamma;
    int source, plain;
    double par[4];
};
// out box of g = g.gamma * g.forcing (or zero) + the quadrature sums of the elements [elo, elo + en); deterministic.
// coef: per-point coefficient table of the whole domain (x fastest) for the forms that use one.
int launch_rhs_brick(int ndim, const QuadAxes& A, const RhsGeom& g, const PointFormArgs& f, const int elo[3],
                     const int en[3], const double* coef, cudaStream_t st, int* nlaunch);

// Per-line factors (kernels_lines.cu): the special dimension of the generalised ADS.
int launch_line_sweep(int p, double* t, const double* ab, const int* ipiv, int n, long long lines, int axis, int L0,
                      long long s0, long long s1, long long sj, double* scratch, cudaStream_t st);

// ---- segmented substitution (kernels_seg.cu; tables: build_segment_plan in host_setup.cpp)
struct SegDev {
    int n, KL, KD, S, DF, DB;
    const int* bounds;   // [S+1]
    const double* E;     // [S][KL][KL]
    const double* Wf;    // [S][DF][KL][KL]
    const double* Vb;    // [S][DB][KD][KD]
    const double* XiF;   // [S][KD][KL]
    const double* cf;    // [n][KD+KL]  Psi | Xi
};
// Lines of a view: element (row j, line (l0, l1)) at (j - row_base)*sj + l0*s0 + l1*s1; state arrays are
// [S][K][L0*L1] with line = l0 + L0*l1.  row_base: global row of the view's first row (a slab's first plane).
struct SegGeom {
    const double* in;
    double* out;
    long long sj_in, sj_out, s0_in, s0_out, s1_in, s1_out;
    int L0, L1;
    int row_base;
    int s_lo, s_hi;  // segments [s_lo, s_hi)
};
int launch_seg_dseg(const SegDev& T, const SegGeom& G, double* const* dst, int ndst, cudaStream_t st);
int launch_seg_din(const SegDev& T, const SegGeom& G, const double* dseg, double* din, double* const* xdst, int ndst,
                   cudaStream_t st);
int launch_seg_tin(const SegDev& T, int s_lo, int s_hi, long long L, const double* X, double* tin, cudaStream_t st);
// max_rows: longest segment among [s_lo, s_hi); tin_is_x: chain depth 1, `tin` is the X array itself
int launch_seg_correct(const SegDev& T, const SegGeom& G, const double* din, const double* tin, int tin_is_x,
                       int max_rows, cudaStream_t st);

// Fused distributed sweep (kernels_sweep_dist.cu): one rank's arguments.  State arrays are [S][K][lines] and
// hold ADSB_DIST_SENTINEL_BITS in every word between sweeps (the exchange protocol polls the data itself).
constexpr unsigned long long ADSB_DIST_SENTINEL_BITS = 0x7FF7A5A57FF7A5A5ull;  // a signalling NaN, both halves equal
struct SweepDistArgs {
    int rank, row_base, lag;
    double* dseg_local;
    double* x_local;
    double* dseg_next;  // state array of rank + 1 (peer pointer), nullptr on the last rank
    double* x_prev;     // state array of rank - 1, nullptr on the first rank
    int* error_flag;
    // optional: the first / last halo_planes rows of the corrected slab are also stored to these planes (same
    // line layout as the slab): the neighbours' halo regions
    double* halo_prev;
    double* halo_next;
    int halo_planes;
};
int launch_neighbor_barrier(unsigned long long* mine, unsigned long long* prev, unsigned long long* next, int* err,
                            cudaStream_t st);
// dry_run: only report eligibility (0 / -1), launch nothing
int launch_sweep_dist(const SweepFactor& F, int CH, const SegDev& T, const SweepGeom& G, const SweepDistArgs& D, int NL,
                      cudaStream_t st, bool dry_run = false);

// Spline values at a tensor-product grid of points (kernels_quad.cu): per axis the first DOF of every point's
// span and its p+1 basis values (device arrays); out[i + n0*(j + n1*k)].
int launch_sample(int ndim, const int* npts, const int* p, const int* const* first, const double* const* val,
                  const double* u, long long s1, long long s2, double* out, cudaStream_t st);

// Norms / errors by element quadrature (kernels_norm.cu).  partial: norm_partial_doubles() doubles of scratch;
// out: 2 doubles on the device (sum of N(u_h - ref) w J, sum of N(ref) w J).  h1: add the gradient terms;
// ref: 0 none, 1 validation solution at time t, 2 tabulated values `tab` at the quadrature points (x fastest).
long long norm_partial_doubles(const QuadAxes& A);
int launch_norm(const QuadAxes& A, const double* u, long long s1, long long s2, int h1, int ref, double t,
                const double* tab, double* partial, double* out, cudaStream_t st);

// vnx > 0: `values` are the first elements of a tensor with x rows of vnx doubles, vpitch apart (else dense)
int launch_set_plane(double* t, const long long s[3], const int n[3], int axis, int idx,
                     const double* values, cudaStream_t st, int vnx = 0, long long vpitch = 0);

}  // namespace adsb

#endif
