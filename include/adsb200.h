/* adsb200.h -- C ABI of libadsb200.so: the B200-native ADS time-step hot path.
 *
 * The reference (marcinlos/iga-ads) has no FFI boundary for this path: it is a C++17 template
 * library whose "interface" is the class surface the examples inherit from, plus three LAPACK
 * symbols.  This header is the boundary a maintainer would bind instead; each entry point cites
 * the reference interface it replaces (paths relative to the reference tree).  The C++17 host
 * layer in iga_ads_b200/include/ads/ (same names as the reference: ads::dimension,
 * ads::simulation_2d/3d, ads::lin::tensor, ads::ads_solve ...) and the Python mirror in
 * iga_ads_b200/ call nothing but these functions.
 *
 * Conventions
 *   - plain pointers and sizes only; opaque handle adsb_ctx; no C++/torch types
 *   - every function returns 0 on success, a negative code on failure, with a message in
 *     adsb_last_error() (thread-local).  The reference ignores LAPACK `info`
 *     (include/ads/lin/band_solve.hpp:17,:29); we surface it (singular factor => ADSB_ESINGULAR).
 *   - tensors are column-major (first index fastest), exactly as ads::lin::tensor
 *     (include/ads/util/multi_array/ordering/reverse.hpp:28-31)
 *   - band matrices use the LAPACK general-band layout with factor workspace of
 *     ads::lin::band_matrix(kl, ku, n): ldab = 2*kl+ku+1, A(i,j) at ab[j*ldab + kl+ku+i-j]
 *     (include/ads/lin/band_matrix.hpp:31-40,:69-73)
 *   - there is NO CPU fallback: device entry points fail with ADSB_ENODEVICE without a GPU.
 */
#ifndef ADSB200_H
#define ADSB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ADSB_ABI_VERSION 2

#define ADSB_OK 0
#define ADSB_EINVAL (-1)     /* bad argument */
#define ADSB_ENODEVICE (-2)  /* no CUDA device / CUDA call failed */
#define ADSB_ESINGULAR (-3)  /* zero pivot in a band factor (LAPACK info > 0) */
#define ADSB_ESTATE (-4)     /* tables / factors / buffers not set up for this call */
#define ADSB_ENOMEM (-5)

#define ADSB_MAX_P 7         /* spline degree limit of the device kernels */
#define ADSB_MAX_SLOTS 4     /* factor slots per axis (M, K, ...) */
#define ADSB_MAX_BUFFERS 8   /* managed coefficient tensors per context */

typedef struct adsb_ctx adsb_ctx;

int adsb_abi_version(void);
const char* adsb_last_error(void);

/* ======================= host-side setup (pure CPU, O(n p^2 q)) ===========================
 * Runs once per simulation; stays on the host exactly as in the reference. */

/* Gauss-Legendre nodes (ascending) and weights on [-1,1], 2 <= q <= 64.
 * Replaces quad::gauss::Xs/Ws (include/ads/quad/gauss.hpp:14-15). */
int adsb_gauss(int q, double* x, double* w);

/* Clamped uniform knot vector; returns its size elements+2p+1 (or <0).
 * Replaces bspline::create_basis (src/ads/bspline/bspline.cpp:26-43, repeated_nodes = 0). */
int adsb_knots(int p, int elements, double a, double b, double* knots);

/* Replaces bspline::find_span (src/ads/bspline/bspline.cpp:61-81); returns the span. */
int adsb_find_span(double x, const double* knots, int knot_size, int p);

/* Values and derivatives of the p+1 non-zero B-splines at x: out[d*(p+1)+i], d = 0..ders.
 * Replaces bspline::eval_basis_with_derivatives (src/ads/bspline/bspline.cpp:102-160). */
int adsb_basis_ders(int span, double x, const double* knots, int p, int ders, double* out);

/* Per-axis quadrature tables, flat: b[e][k][d][i] (elements x q x (ders+1) x (p+1)), xq[e][k],
 * w[k], J[e], first_dof[e].  Replaces ads::basis_data (src/ads/basis_data.cpp:63-114). */
int adsb_basis_tables(int p, int elements, double a, double b, int q, int ders, double* b_flat,
                      double* xq, double* w, double* J, int* first_dof);

/* 1-D quadrature matrices in band layout (ldab = 3p+1, n = elements+p columns), accumulated in
 * the reference's loop order.  kind: 0 Gram, 1 stiffness, 2 advection
 * (src/ads/form_matrix.cpp:8-60), 3 Gram + h*stiffness (examples/implicit/implicit.hpp:46-64,
 * the caller passes the already scaled h).  fix: bit0 fix_left, bit1 fix_right
 * (src/ads/simulation/dimension.cpp:23-29). */
int adsb_matrix_1d(int kind, int p, int elements, double a, double b, double h, int fix,
                   double* ab);

/* Banded LU with partial pivoting, in place, 1-based ipiv.  Replaces lin::factorize -> dgbtrf_
 * (include/ads/lin/band_solve.hpp:16-18).  Returns ADSB_ESINGULAR for an exactly zero pivot. */
int adsb_band_factorize(int n, int kl, int ku, double* ab, int ldab, int* ipiv);

/* ================================ device context ===========================================
 * One context per GPU (one process per GPU).  It owns the device copies of the per-axis tables,
 * operators and factors, and (optionally) managed coefficient tensors.
 *
 * n_global[d] are the DOF counts of the whole problem; lo[d]/cnt[d] the box this context owns
 * (lo = 0, cnt = n_global for a single GPU; a z-slab or y-slab for a sharded run).  ndim 2 or 3. */
int adsb_create(int ndim, const int* n_global, const int* lo, const int* cnt, int device,
                adsb_ctx** out);
int adsb_destroy(adsb_ctx* ctx);

/* CUDA stream (cudaStream_t passed as void*) all later calls enqueue on; NULL = default stream. */
int adsb_set_stream(adsb_ctx* ctx, void* cuda_stream);
int adsb_synchronize(adsb_ctx* ctx);
/* Cap the number of SMs the persistent sweep kernels occupy and the right-hand-side launch heuristics
 * assume (0 = the whole device).  Lets a caller run two contexts on two streams side by side, e.g. a
 * sweep whose stores travel over NVLink next to the compute of the following slab chunk.  No reference
 * counterpart (the reference is single-threaded host code). */
int adsb_set_sm_limit(adsb_ctx* ctx, int sms);
/* Strided device-to-device copy on the context's stream (cudaMemcpy2DAsync, copy engines): `height` rows
 * of `width_bytes`, row pitches in bytes.  dst may be peer memory mapped into this process -- the slab
 * exchange of a sharded step moves its blocks with this while the SMs work on the next chunk. */
int adsb_copy2d(adsb_ctx* ctx, void* dst, long long dst_pitch_bytes, const void* src, long long src_pitch_bytes,
                long long width_bytes, long long height);

/* Upload one axis' quadrature tables (layout of adsb_basis_tables); the library derives the 1-D
 * Gram / stiffness / advection operators from them by the same quadrature.
 * Replaces the basis_data member of ads::dimension (include/ads/simulation/dimension.hpp:20-59)
 * as seen by eval_basis / eval_fun (include/ads/simulation/simulation_3d.hpp:98-128). */
int adsb_set_axis_tables(adsb_ctx* ctx, int axis, int p, int elements, int q, int ders,
                         const double* b_flat, const double* xq, const double* w, const double* J,
                         const int* first_dof);

/* Upload a factorised band matrix (output of adsb_band_factorize / dgbtrf_) into `slot` of `axis`.
 * Replaces ads::dim_data{M, ctx} (include/ads/solver.hpp:17-20).  kl, ku <= 5.  What the sweeps do with it:
 * a factor with row interchanges is eliminated again without them when that is stable (adsb_band_unpivot);
 * lines of more than 576 rows are cut into segments by the library (adsb_set_axis_segments semantics); a factor
 * whose segments cannot be cut (boundary responses that grow: strongly stiffness-dominated matrices) keeps the
 * unsegmented register-path kernel, which handles lines of up to 4608 rows -- longer lines with such a factor
 * make adsb_sweep / adsb_solve fail with ADSB_EINVAL (the reference has no such limit). */
int adsb_set_axis_factor(adsb_ctx* ctx, int axis, int slot, int n, int kl, int ku, int ldab,
                         const double* ab, const int* ipiv);

/* ---- managed coefficient tensors (cnt[0]*cnt[1]*cnt[2] doubles each, ids 0..ADSB_MAX_BUFFERS-1)
 * Device mirrors of lin::tensor objects (include/ads/lin/tensor/tensor.hpp:14-50). */
int adsb_upload(adsb_ctx* ctx, int buf, const double* host);
int adsb_download(adsb_ctx* ctx, int buf, double* host);
/* The same copies enqueued on `cuda_stream` (cudaStream_t as void*) without waiting: pinned `host` memory,
 * the caller orders the stream against the context's own (events) -- used to overlap the next input's
 * upload and the previous result's download with a running step.  The first use of a buffer allocates it. */
int adsb_upload_async(adsb_ctx* ctx, int buf, const double* host, void* cuda_stream);
int adsb_download_async(adsb_ctx* ctx, int buf, double* host, void* cuda_stream);
int adsb_swap(adsb_ctx* ctx, int buf_a, int buf_b);        /* std::swap(u, u_prev) */
int adsb_zero(adsb_ctx* ctx, int buf);                      /* zero(rhs), tensor.hpp:44-50 */
int adsb_bind(adsb_ctx* ctx, int buf, double* device_ptr);  /* adopt caller-owned device memory */
double* adsb_device_ptr(adsb_ctx* ctx, int buf);
/* Device layout of the managed tensors: the reference's index order (x fastest) with every x row padded to
 * adsb_row_pitch() doubles (cnt[0] rounded up to even, so rows are 16 B aligned for the TMA-fed kernels);
 * element (i, j, k) sits at i + pitch * (j + cnt[1] * k).  adsb_upload / adsb_download convert from / to the
 * dense host layout; memory handed to adsb_bind must follow the padded layout. */
long long adsb_row_pitch(adsb_ctx* ctx);

/* Overwrite the hyper-plane index `idx` of `axis` with `values` (host, product of the other
 * extents doubles).  Replaces the Dirichlet overwrite `v(0,i) = buf(i)` of
 * examples/heat/heat_2d.hpp:40-47. */
int adsb_set_plane(adsb_ctx* ctx, int buf, int axis, int idx, const double* values);

/* ---- right-hand side:  rhs_a = sum_e sum_q [ alpha*u*v_a - sum_k beta[k]*d_k u*d_k v_a ] w J
 *                                 + gamma * F_a
 * the form of every compute_rhs() on the path (examples/heat/heat_3d.hpp:49-67,
 * heat_2d.hpp:80-106, implicit/implicit.hpp:132-182, scalability/test3d.hpp:66-95).
 * method ADSB_RHS_COLLAPSED applies the exactly pre-integrated 1-D quadrature operators
 * (sum factorisation carried through the quadrature sums; HBM-bound);
 * method ADSB_RHS_QUADRATURE evaluates u and grad u at every Gauss point and integrates against
 * the test functions by sum factorisation (general pointwise forms; FP64-bound, deterministic): the brick
 * kernel of adsb_compute_rhs_pointwise with the linear form; it replaces zero(rhs) + the element loop +
 * update_global_rhs (include/ads/simulation/simulation_3d.hpp:140-145). */
#define ADSB_RHS_COLLAPSED 0
#define ADSB_RHS_QUADRATURE 1
typedef struct {
    double alpha;    /* coefficient of u*v */
    double beta[3];  /* coefficient of d_k u * d_k v (pass +dt for "- dt * grad.grad") */
    double gamma;    /* coefficient of the load tensor in `forcing_buf` (0: none) */
    int forcing_buf; /* managed buffer id holding F, or -1 */
    int method;      /* ADSB_RHS_COLLAPSED / ADSB_RHS_QUADRATURE */
    int source;      /* ADSB_RHS_QUADRATURE only: built-in source f evaluated at the quadrature points and
                        added as gamma*f(x_q)*w*J to every DOF of the element -- without the test function,
                        as examples/scalability/test3d.hpp:86-88 does; 0 none, 1 the scalability forcing */
} adsb_form;
int adsb_compute_rhs(adsb_ctx* ctx, const adsb_form* form, int src_buf, int dst_buf);

/* ---- generalised ADS: one "special" axis whose lines each have their own band matrix
 * Replaces ads_solve(rhs, buf, dims...) with a callable among the dims (include/ads/solver.hpp:56-96,:170-195; the
 * callable of examples/maxwell/maxwell_ads.hpp:139-163 solves line (i, j) with its own factorised matrix).
 * adsb_set_line_factors: ab_lines[l][j][r] is the factor of line l exactly as adsb_band_factorize / dgbtrf_ leaves
 * it (n columns of 2kl+ku+1 rows), ipiv_lines[l][j] its 1-based pivots; kl = ku = p <= 5.  Lines are numbered over
 * the other axes in the tensor's own order (x fastest): axis 0: l = iy + ny*iz; axis 1: l = ix + nx*iz; axis 2:
 * l = ix + nx*iy.  adsb_solve_special solves the special axis first (one dgbtrs recurrence per line, a thread per
 * line), then the remaining axes with the factor slots slots[d], as the reference does. */
int adsb_set_line_factors(adsb_ctx* ctx, int axis, int kl, int ku, const double* ab_lines, const int* ipiv_lines);
int adsb_solve_special(adsb_ctx* ctx, int buf, int special_axis, const int* slots);

/* ---- general pointwise forms: the element loop of compute_rhs() with an arbitrary integrand
 *     rhs_a = [forcing_scale * F_a] + sum_e sum_q ( k0 v_a + k1 d_x v_a + k2 d_y v_a + k3 d_z v_a ) w J
 * where k0..k3 are functions of the point, u_prev and grad u_prev there (include/ads/simulation/simulation_3d.hpp:
 * 64-145: eval_fun / eval_basis / update_global_rhs; the model loop is examples/scalability/test3d.hpp:66-95).
 * Evaluated by Gauss quadrature with sum factorisation over bricks of elements, FP64-bound, deterministic
 * (csrc/quadbrick.cuh); the integrand is a template functor of the kernel, selected by `kind`:
 *   ADSB_POINT_LINEAR  k0 = alpha u - adv . grad u + gamma f(x),  k_d = -beta[d] d_d u     (every form of
 *                      adsb_form, plus the advection term of examples/pollution/pollution_3d.hpp);
 *                      source: 0 none, 1 the scalability forcing (test3d.hpp:58-64), 2 the flow forcing
 *                      (examples/flow/flow.hpp:124-128); source_plain 1: gamma f w J goes to every DOF of the
 *                      element WITHOUT the test function, as test3d.hpp:86-88 does
 *   ADSB_POINT_FLOW    examples/flow/flow.hpp:74-101 (nonlinear): k0 = u + dt h(x), k_d = -dt k(x) exp(mi u) d_d u
 *                      with par[0] = dt, par[1] = mi, h = source 2 and k the table of adsb_set_point_coefficient
 *                      (flow.hpp:53-60 fill_permeability_map); 3-D, p <= 3
 * The context must own the whole domain, same degree on every axis, quad_order = p + 1, derivatives = 1. */
#define ADSB_POINT_LINEAR 0
#define ADSB_POINT_FLOW 1
typedef struct {
    int kind;
    double alpha;
    double beta[3];
    double adv[3];
    double gamma;
    int source;
    int source_plain;
    double par[4];
    int forcing_buf;       /* managed buffer holding a load tensor F, or -1 */
    double forcing_scale;  /* its coefficient */
} adsb_point_form;
/* values (host): one double per quadrature point of the domain, x fastest:
 * values[(ex*q+kx) + nqx*((ey*q+ky) + nqy*(ez*q+kz))], nq_d = elements_d * q_d */
int adsb_set_point_coefficient(adsb_ctx* ctx, const double* values);
int adsb_compute_rhs_pointwise(adsb_ctx* ctx, const adsb_point_form* form, int src_buf, int dst_buf);

/* Load tensor of a built-in source: F_a = sum_{e in supp(a)} sum_q f(x_q) [v_a(x_q)] w J.
 * source 1: the scalability forcing (examples/scalability/test3d.hpp:58-64, test2d.hpp:49-54).
 * with_test_function 0 reproduces the reference form, which adds dt*f(x_q)*w*J to every DOF of
 * the element without the factor v_a (test3d.hpp:86-88). */
int adsb_load_tensor(adsb_ctx* ctx, int source, int with_test_function, int dst_buf);

/* L2-projection right-hand side of a built-in initial state (include/ads/projection.hpp:12-153):
 * state 0 heat_3d bump (examples/heat/heat_3d.hpp:22-28), 1 implicit 2-D/3-D bump
 * (examples/implicit/implicit.hpp:38-43), 2 constant one. */
int adsb_project_init(adsb_ctx* ctx, int state, int dst_buf);

/* ---- output sampling: the spline on a tensor-product grid of points
 * Replaces output_manager<2/3>::write / evaluate (include/ads/output_manager.hpp:66-73,:101-118), i.e.
 * bspline::eval at every point (include/ads/bspline/eval.hpp:161-192): points[d] are npts[d] coordinates along
 * axis d (the reference uses linspace(a, b, intervals): intervals + 1 points), knots[d] the axis' knot vector
 * (adsb_knots).  out (host) receives npts[0]*npts[1][*npts[2]] values, first index fastest -- the order the
 * reference's writers print them in (include/ads/output/vtk.hpp:45-82, gnuplot.hpp:45-56).  The spans and basis
 * values are found on the host exactly as the reference does, the contraction runs on the device in the
 * reference's order.  The context must own the whole domain; the call synchronises. */
int adsb_sample(adsb_ctx* ctx, int buf, const int* npts, const double* const* points, const double* const* knots,
                double* out);

/* ---- norms and errors of the spline solution by element quadrature (a diagnostic outside the step)
 * Replaces basic_simulation_2d/3d::normL2 / normH1 / errorL2 / errorH1
 * (include/ads/simulation/basic_simulation_3d.hpp:281-398; basic_simulation_2d.hpp likewise):
 *   out2[0] = sqrt( sum_e sum_q N(u_h(x_q) - ref(x_q)) w J ),   out2[1] = sqrt( sum_e sum_q N(ref(x_q)) w J )
 * kind 0: N = val^2 (L2); 1: val^2 + |grad|^2 (H1).  ref 0: none (out2[0] is the norm of u_h); 1: the
 * validation solution sin(pi x) sin(pi y) [sin(pi z)] exp(-d pi^2 t) (examples/validation/validation.hpp:45-55);
 * 2: values tabulated by the caller at the quadrature points, x fastest: ref_values[(ex*q+kx) + nqx*((ey*q+ky) +
 * nqy*(ez*q+kz))] (host memory; L2 only).  The context must own the whole domain; the call synchronises. */
int adsb_norm(adsb_ctx* ctx, int buf, int kind, int ref, double t, const double* ref_values, double* out2);

/* The same projection for ANY function: the caller tabulates f at the quadrature points (host memory, x fastest:
 * values[(ex*q+kx) + nqx*((ey*q+ky) + nqy*((ez-ez_lo)*q+kz))], nq = elements*q per axis) of the z-element slab
 * [ez_lo, ez_lo+ez_cnt) (ignored in 2-D; ez_cnt <= 0: all z elements) -- a slab at a time keeps the table small (the whole 512^3 problem
 * would need 29 GB).  accumulate != 0 adds to dst (walk the slabs in ascending order); one call over all z
 * elements reproduces the reference's summation order exactly (include/ads/projection.hpp:60-107), slab-wise
 * accumulation differs by rounding only.  The call synchronises. */
int adsb_project_values(adsb_ctx* ctx, int dst_buf, int ez_lo, int ez_cnt, const double* values, int accumulate);

/* ---- ADS solve: one batched banded forward/back substitution per axis, in place.
 * Replaces ads::ads_solve(rhs, buffer, dims...) (include/ads/solver.hpp:35-41,:148-160,:222-226)
 * i.e. 3 x { lin::solve_with_factorized -> dgbtrs_ (include/ads/lin/band_solve.hpp:21-31) +
 * lin::cyclic_transpose (include/ads/lin/tensor/cyclic_transpose.hpp:54-63) }.  The tensor never
 * moves: each sweep reads and writes the canonical layout.  slots[d] selects the factor of axis d. */
int adsb_solve(adsb_ctx* ctx, int buf, const int* slots);

/* One axis only (lin::solve_with_factorized on the axis-`axis` lines of the tensor). */
int adsb_sweep(adsb_ctx* ctx, int buf, int axis, int slot);

/* ---- whole steps resident on the device (simulation_base::run's loop body,
 * src/ads/simulation/simulation_base.cpp:11-20 with step() of the examples):
 *   repeat nsteps: for each sub-step s: swap(u, u_prev); rhs(form[s]) -> u; [planes]; solve(slots[s]) */
typedef struct {
    adsb_form form;
    int slots[3];
    int fix_axis;     /* -1, or axis whose plane 0 is overwritten before the solve (heat_2d) */
    int fix_buf;      /* managed buffer holding the plane values in its first entries */
} adsb_substep;
int adsb_step(adsb_ctx* ctx, int u_buf, int uprev_buf, const adsb_substep* sub, int nsub, int nsteps);

/* Device time (ms) spent in the stages since the last call: [0] rhs, [1..3] sweep per axis,
 * [4] other.  Only collected when enabled (costs event records). */
int adsb_enable_timing(adsb_ctx* ctx, int on);
int adsb_stage_times(adsb_ctx* ctx, double* ms5);

/* Number of kernels this context has launched so far. */
long long adsb_launch_count(adsb_ctx* ctx);

/* ============== pointer-level entry points (sharded runs; caller owns the memory) ===========
 * A view is extents n[3] and element strides s[3] (doubles); unused axes have n = 1. */
typedef struct {
    int n[3];
    long long s[3];
} adsb_view;

/* Sweep along `axis` over all lines of the view.  row_off_in/out (may be NULL) give, for each
 * index j along the sweep axis, the element offset of that hyper-plane instead of j*s[axis]
 * (host arrays of n[axis] entries): lets the sweep read / write the block layout of an all-to-all. */
int adsb_sweep_view(adsb_ctx* ctx, int axis, int slot, const double* in, const adsb_view* vin,
                    const long long* row_off_in, double* out, const adsb_view* vout,
                    const long long* row_off_out);

/* Collapsed RHS on a box: `in` covers [in_lo, in_lo+vin.n) in global DOF indices (with whatever
 * halo the caller exchanged), `out` covers [out_lo, out_lo+vout.n).  Input outside the global
 * domain is never read; input inside the domain but outside the `in` box is an error. */
int adsb_rhs_view(adsb_ctx* ctx, const adsb_form* form, const double* in, const adsb_view* vin,
                  const int* in_lo, const double* forcing, double* out, const adsb_view* vout,
                  const int* out_lo);

/* The factor adsb_set_axis_factor actually sweeps with.  A dgbtrf factor with row interchanges is eliminated
 * again WITHOUT interchanges when that is stable (the matrices of this path are symmetric positive definite up to
 * their fix_left / fix_right rows): U keeps ku super-diagonals instead of kl + ku and the chunk chains of the
 * substitution kernel stay short.  ab_out (n columns of ldab rows, dgbtrf layout, identity pivots) receives it;
 * returns 0, or 1 when the elimination would not be safe (a multiplier above 4, a pivot below 1e-8 max|A|) and the
 * caller's interchanges are kept.  Same linear system as lin::solve_with_factorized (include/ads/lin/band_solve.hpp:
 * 21-31), results agree with dgbtrs to rounding.  ADSB_UNPIVOT=0 in the environment switches the step off. */
int adsb_band_unpivot(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, double* ab_out);

/* ---- introspection (used by the CPU test-suite to check the substitution plan without a GPU)
 * Builds the chunked-substitution plan of a factor exactly as adsb_set_axis_factor does and copies
 * it out.  dims[16] = {KL, KD, piv, CH, R, SC, ST, rows, LF, LB, LC, DF, DB, seq, MAX_DEPTH, 0};
 * arrays: pv[rows], cfF[rows*LF], cfB[rows*LB], cfC[rows*LC], T[SC*KL*KL], Rm[SC*KD*KD],
 * W[SC*(MAX_DEPTH-1)*KL*KL], V[SC*(MAX_DEPTH-1)*KD*KD]; any output pointer may be NULL. */
int adsb_sweep_plan(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, int* dims,
                    int* pv, double* cfF, double* cfB, double* cfC, double* T, double* Rm, double* W,
                    double* V);

/* ================== segmented substitution (long lines, slab-sharded sweeps) =================
 * dgbtrs over a line cut into segments [bounds[s], bounds[s+1]): every segment is solved alone with the
 * factor's own columns (pass A: an ordinary sweep with the segment's factor), then corrected with KL + KD
 * boundary values per segment and line (pass B).  Algebraically the recurrence of
 * lin::solve_with_factorized -> dgbtrs_ (include/ads/lin/band_solve.hpp:21-31) with the same factor and
 * pivots; segments are the z-slabs of a sharded run (one per GPU: only the boundary values cross NVLink,
 * no transpose) or the pieces of a line too long for one CTA (heat_2d 4096^2). */

/* Balanced cuts that no row interchange of the factor crosses, multiples of `align` where possible;
 * bounds[nseg+1].  Host only. */
int adsb_segment_bounds(int n, int kl, const int* ipiv, int nseg, int align, int* bounds);

/* Introspection (CPU tests): the segment tables exactly as adsb_set_axis_segments builds them.
 * dims[8] = {KL, KD, piv, S, DF, DB, n, 0}; E[S*KL*KL], Wf[S*DF*KL*KL], Vb[S*DB*KD*KD], XiF[S*KD*KL],
 * cf[n*(KD+KL)] (Psi | Xi per row); any output may be NULL.  tol <= 0: default chain cut-off 1e-20. */
int adsb_segment_plan(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, int nseg, const int* bounds,
                      double tol, int* dims, double* E, double* Wf, double* Vb, double* XiF, double* cf);

/* Cut the factor already uploaded to (axis, slot) into nseg segments; builds and uploads the segment tables
 * and the pass-A factors of segments [local_lo, local_lo + local_cnt) (a sharded rank: its own slab only).
 * Fails with ADSB_EINVAL when a cut crosses a row interchange, a segment is shorter than the band, or the
 * factor's boundary responses grow (then use the unsegmented sweep / the transposing exchange). */
int adsb_set_axis_segments(adsb_ctx* ctx, int axis, int slot, int nseg, const int* bounds, int local_lo, int local_cnt);
/* info8 = {KL, KD, DF, DB, S, local_lo, local_cnt, n}: state widths (doubles per line and segment boundary:
 * KL forward, KD backward) and chain depths (how many segments to the left / right a segment depends on). */
int adsb_segment_info(adsb_ctx* ctx, int axis, int slot, int* info8);

/* Pass A: dgbtrs of segment `seg` alone; the view spans exactly the rows of that segment along `axis`. */
int adsb_seg_sweep_view(adsb_ctx* ctx, int axis, int slot, int seg, const double* in, const adsb_view* vin, double* out,
                        const adsb_view* vout);
/* Boundary states.  The view holds the rows [row_base, row_base + vin->n[axis]) of the line (all of it on one
 * GPU; a rank's slab in a sharded run).  State arrays are [S][K][lines] doubles, lines = product of the two
 * other extents (lower-stride axis fastest); dst / xdst list every array the values are stored to -- the
 * local one and, through peer pointers, those of the ranks that depend on this segment.
 *   dseg:  Dseg[s] = E_s * xhat_s[last KL rows]                                   s in [s_lo, s_hi)
 *   din:   din[s] = sum_{d=1..DF} Wf[s][d] Dseg[s-d];  X[s] = xhat_s[first KD rows] + XiF_s din[s]
 *   tin:   tin[s] = sum_{d=1..DB} Vb[s][d] X[s+d]       (skip when DB == 1: pass X to adsb_seg_correct_view)
 *   correct (pass B):  out = xhat + Psi tin[s] + Xi din[s]; in == out allowed. */
/* The same, fused: pass A, the boundary exchange with the two neighbouring ranks and pass B in ONE persistent
 * kernel per rank, software-pipelined tile by tile (csrc/kernels_sweep_dist.cu).  The slab is read from HBM
 * once and written once; the boundary values travel as 8-byte peer stores that validate themselves: every word
 * of the state arrays holds a sentinel (the 32-bit pattern ADSB_DIST_SENTINEL_WORD twice: a signalling NaN)
 * until the neighbour's store replaces it; the receiver polls the word, takes it and puts the sentinel back.
 * The caller fills both state arrays with the sentinel ONCE (e.g. a 32-bit memset of the word) and separates
 * consecutive sweeps on the same arrays by a barrier over the ranks (the end-of-step halo barrier does).
 * Every rank must make the same sequence of calls with the same nl / lag / SM limit; chain depths DF = DB = 1
 * only (ADSB_ESTATE otherwise: use the separate entry points above).  In place on `data` (the rows of segment
 * `rank`).  state arrays: [S][K][lines] as above. */
#define ADSB_DIST_SENTINEL_WORD 0x7FF7A5A5u
typedef struct {
    int rank, nranks;
    int nl;                          /* lines per tile: 16, 32 or 64 */
    int lag;                         /* tiles between pass A and the exchange stages (0: default 4) */
    double* dseg_local;              /* own state arrays, sentinel-filled */
    double* x_local;
    double* dseg_next;               /* state array of rank + 1 (peer pointer; ignored on the last rank) */
    double* x_prev;                  /* state array of rank - 1 (ignored on the first rank) */
    int* error_flag;                 /* device int, set to 1 when a poll timed out (~30 s); may be NULL */
    /* optional halo publishing (the right-hand side of the next step needs p planes of each neighbour): the
     * first halo_planes planes of the finished slab are also stored at halo_prev, the last ones at halo_next
     * (peer pointers into the neighbours' state buffers, same plane layout as `data`; NULL: skip) */
    double* halo_prev;
    double* halo_next;
    int halo_planes;
} adsb_dist_args;
int adsb_dist_sweep_view(adsb_ctx* ctx, int axis, int slot, double* data, const adsb_view* view,
                         const adsb_dist_args* args);
/* 1 when adsb_dist_sweep_view can run this slab shape with nl lines per tile and the given lag (shared memory,
 * kernel variant, chain depths), 0 when not (use the separate entry points), < 0 on bad arguments.  Host only:
 * lets every rank of a run agree on the path before anything is launched. */
int adsb_dist_sweep_check(adsb_ctx* ctx, int axis, int slot, int rank, const adsb_view* view, int nl, int lag);

/* Barrier with the two neighbouring ranks on the context's stream (a one-thread kernel): every dependency of
 * the slab-sharded step -- halo planes, re-use of the boundary-value arrays -- is between neighbours, so the
 * ranks need no all-to-all barrier.  flags_*: 3 64-bit words per rank in peer-mapped, zero-initialised memory
 * ([0] written by the previous rank, [1] by the next, [2] the rank's own counter, which makes the call safe
 * to replay from a CUDA graph); flags_prev / flags_next are the neighbours' arrays (NULL at the ends). */
int adsb_neighbor_barrier(adsb_ctx* ctx, unsigned long long* flags_local, unsigned long long* flags_prev,
                          unsigned long long* flags_next, int* error_flag);

int adsb_seg_dseg_view(adsb_ctx* ctx, int axis, int slot, int s_lo, int s_hi, int row_base, const double* xhat,
                       const adsb_view* vin, double* const* dst, int ndst);
int adsb_seg_din_view(adsb_ctx* ctx, int axis, int slot, int s_lo, int s_hi, int row_base, const double* xhat,
                      const adsb_view* vin, const double* dseg, double* din, double* const* xdst, int ndst);
int adsb_seg_tin(adsb_ctx* ctx, int axis, int slot, int s_lo, int s_hi, long long lines, const double* X, double* tin);
int adsb_seg_correct_view(adsb_ctx* ctx, int axis, int slot, int s_lo, int s_hi, int row_base, const double* in,
                          const adsb_view* vin, double* out, const adsb_view* vout, const double* din,
                          const double* tin_or_x);

/* ================== the z-slab sharded step from one host process (csrc/slab_host.cpp) ==================
 * `world` ranks; rank r owns a slab of z planes on devices[r].  Distinct devices: neighbours map each other's
 * memory (peer access over NVLink) and every rank runs right-hand side, x / y sweeps and the fused distributed
 * z sweep on its own stream, the end-of-sub-step neighbour barrier being stream-event waits.  All ranks on ONE
 * device: "virtual ranks" sharing it as concurrent streams with world-th of the SMs each (single-GPU testing).
 * The reference's counterpart is simulation_base::run (src/ads/simulation/simulation_base.cpp:11-20) in one
 * address space; iga_ads_b200/slab.py is the same step with one process per GPU.
 * Order of calls: create; set_axis_tables x3 and set_axis_factor for every (axis, slot) the program names (same
 * arguments as the adsb_ctx entry points, replicated to every rank); commit(program) -- picks the slab bounds
 * (adsb_segment_bounds of the first z factor), builds the ranks, fails with ADSB_EINVAL when the z factor couples
 * more than neighbouring slabs (chain depth > 1) or the fused sweep does not fit; upload / step / download.
 * upload / download: the whole tensor, dense, x fastest (host memory).  step enqueues nsteps time steps and
 * returns; synchronize waits and reports a timed-out wait of the fused sweep.  Forms with gamma != 0 use the
 * scalability load tensor (adsb_load_tensor(ctx, 1, 0, .)) of every slab. */
typedef struct adsb_slabs adsb_slabs;
int adsb_device_count(void);  /* CUDA devices visible to the process (0: none) */
int adsb_slabs_create(int world, const int* devices, const int* n_global, adsb_slabs** out);
int adsb_slabs_destroy(adsb_slabs* s);
int adsb_slabs_set_axis_tables(adsb_slabs* s, int axis, int p, int elements, int q, int ders, const double* b_flat,
                               const double* xq, const double* w, const double* J, const int* first_dof);
int adsb_slabs_set_axis_factor(adsb_slabs* s, int axis, int slot, int n, int kl, int ku, int ldab, const double* ab,
                               const int* ipiv);
int adsb_slabs_commit(adsb_slabs* s, const adsb_substep* program, int nsub);
int adsb_slabs_upload(adsb_slabs* s, const double* host);
int adsb_slabs_download(adsb_slabs* s, double* host);
int adsb_slabs_step(adsb_slabs* s, int nsteps);
int adsb_slabs_synchronize(adsb_slabs* s);
/* bounds[world + 1] (may be NULL); info4 = {world, virtual ranks (0/1), lines per tile of the fused sweep, kernel launches so far} */
int adsb_slabs_info(adsb_slabs* s, int* bounds, int* info4);

#ifdef __cplusplus
}
#endif

#endif /* ADSB200_H */
