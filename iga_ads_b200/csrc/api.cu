// api.cu -- device context and the C ABI of libadsb200.so (see include/adsb200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "adsb200.h"
#include "internal.hpp"
#include "kernels.cuh"

using namespace adsb;

namespace {

struct SegSet {  // adsb_set_axis_segments: tables + the pass-A factor of every local segment
    bool set = false;
    SegDev dev{};
    std::vector<int> bounds;
    int local_lo = 0, local_cnt = 0;
    std::vector<SweepFactor> local;  // [local_cnt]
    std::vector<SweepFactor> local_dist;  // the same factors with chunks of SWEEP_CH_DIST columns (fused kernel), or empty
    std::vector<void*> allocs;
};

struct DevFactor {
    bool set = false;
    SweepFactor f{};
    std::vector<void*> allocs;
    // host copy of the dgbtrf output (segments are cut from it later)
    int n = 0, kl = 0, ku = 0, ldab = 0;
    std::vector<double> ab;
    std::vector<int> ipiv;
    SegSet seg;
    bool auto_seg = false;  // segments were cut by the library because the line is too long for one CTA
};

struct AxisData {
    bool tables = false;
    int p = 0, elements = 0, q = 0, ders = 0, n = 0;
    std::vector<double> bt, xq, w, J;  // host copies (layout of adsb_basis_tables)
    double* d_M = nullptr;             // [n][2p+1] Gram rows
    double* d_S = nullptr;             // [n][2p+1] stiffness rows
    double* d_MST = nullptr;           // [n+2p][2][2p+2] Gram | stiffness columns A(k-p..k+p, k) in row k+p
    double* d_bt = nullptr;            // device copy of bt
    double* d_xq = nullptr;
    double* d_wJ = nullptr;            // [elements][q] w[k]*J[e]
    double* d_w = nullptr;
    double* d_J = nullptr;
    DevFactor fac[ADSB_MAX_SLOTS];
};

struct TimedSpan {
    int stage;
    cudaEvent_t a, b;
};

}  // namespace

struct adsb_ctx {
    int ndim = 0;
    int ng[3] = {1, 1, 1}, lo[3] = {0, 0, 0}, cnt[3] = {1, 1, 1};
    int device = 0;
    cudaStream_t stream = nullptr;
    AxisData ax[3];
    double* buf[ADSB_MAX_BUFFERS] = {};
    bool owned[ADSB_MAX_BUFFERS] = {};
    std::map<std::vector<long long>, long long*> off_cache;
    bool timing = false;
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> free_events;
    double acc_ms[5] = {};
    long long launches = 0;
    int sm_limit = 0;  // adsb_set_sm_limit
    RhsSide side;      // second stream + events for the x-remainder kernel of the right-hand side (lazy)
    // lines too long for one CTA (> 576 rows) are swept in segments (kernels_seg.cu): pass A of the segments
    // runs on a few streams side by side; boundary-state scratch is shared by all axes
    std::vector<cudaStream_t> seg_streams;
    std::vector<cudaEvent_t> seg_events;  // [0] fork, [1 + i] join of stream i
    double* seg_scratch = nullptr;
    size_t seg_scratch_doubles = 0;
    double* point_coef = nullptr;  // adsb_set_point_coefficient: coefficient table at the quadrature points
    // generalised ADS: per-line factors of one axis (adsb_set_line_factors), device layout [j][r][line]
    struct LineFactors {
        double* ab = nullptr;
        int* ipiv = nullptr;
        double* scratch = nullptr;  // x axis only: the transposed line set
        int p = 0, n = 0;
        long long lines = 0;
    } line_fac[3];
    // Managed tensors keep the reference's index order (x fastest) but pad every x row to an EVEN number of
    // doubles: rows, planes and the tensor itself then start 16 B aligned, which is what the TMA-fed
    // kernels need (n = elements + p is odd for every odd degree).  The pad column is never read as data.
    long long pitch0() const { return cnt[0] + (cnt[0] & 1); }
    size_t local_size() const { return (size_t) pitch0() * cnt[1] * cnt[2]; }   // doubles allocated
    size_t rows() const { return (size_t) cnt[1] * cnt[2]; }
};

namespace {

int cuda_fail(cudaError_t e, const char* what) {
    return fail(ADSB_ENODEVICE, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                              \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);   \
    } while (0)

int select_device(adsb_ctx* c) {
    CU(cudaSetDevice(c->device));
    return ADSB_OK;
}

template <typename T>
int upload_vec(const std::vector<T>& h, size_t padded, T** d, std::vector<void*>* track) {
    std::vector<T> tmp(h);
    tmp.resize(std::max(padded, h.size()), T{});
    CU(cudaMalloc((void**) d, tmp.size() * sizeof(T)));
    if (track) track->push_back(*d);
    CU(cudaMemcpy(*d, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ADSB_OK;
}

void free_segments(SegSet& g) {
    for (void* p : g.allocs) cudaFree(p);
    g = SegSet{};
}

void free_factor(DevFactor& f) {
    for (void* p : f.allocs) cudaFree(p);
    f.allocs.clear();
    f.set = false;
    free_segments(f.seg);
}

int upload_plan(const SweepPlan& P, SweepFactor& out, std::vector<void*>& allocs) {
    // one blob: cfF | cfB | cfC | T | Rm | W | V, every piece padded to an even count (16 B granules)
    const std::vector<double>* parts[7] = {&P.cfF, &P.cfB, &P.cfC, &P.T, &P.Rm, &P.W, &P.V};
    std::vector<double> blob;
    int off[7];
    for (int i = 0; i < 7; ++i) {
        off[i] = (int) blob.size();
        blob.insert(blob.end(), parts[i]->begin(), parts[i]->end());
        if (blob.size() & 1) blob.push_back(0.0);
    }
    double* d = nullptr;
    int* pv = nullptr;
    if (int rc = upload_vec(blob, 0, &d, &allocs)) return rc;
    if (int rc = upload_vec(P.pv, 0, &pv, &allocs)) return rc;
    out = SweepFactor{d + off[0], d + off[1], d + off[2], pv, d + off[3], d + off[4], d + off[5], d + off[6],
                      P.n, P.ST, P.SC, P.KL, P.KD, P.piv, P.DF, P.DB, P.seq, (int) blob.size(),
                      {off[1], off[2], off[3], off[4], off[5], off[6]}};
    return ADSB_OK;
}

int ensure_buf(adsb_ctx* c, int b) {
    if (b < 0 || b >= ADSB_MAX_BUFFERS) return fail(ADSB_EINVAL, "buffer id out of range");
    if (!c->buf[b]) {
        CU(cudaMalloc((void**) &c->buf[b], c->local_size() * sizeof(double)));
        CU(cudaMemsetAsync(c->buf[b], 0, c->local_size() * sizeof(double), c->stream));  // finite pad column
        c->owned[b] = true;
    }
    return ADSB_OK;
}


// segment tables + pass-A factors of the local segments for an uploaded factor
int set_segments_impl(adsb_ctx* c, DevFactor& D, int nseg, const int* bounds, int local_lo, int local_cnt) {
    SegPlan P;
    // chain cut-off 1e-17: dropped products change the result by less than a tenth of the unit round-off
    if (int rc = build_segment_plan(D.n, D.kl, D.ku, D.ldab, D.ab.data(), D.ipiv.data(), nseg, bounds, 1e-17, P)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    free_segments(D.seg);
    SegSet& g = D.seg;
    int* d_bounds;
    double *E, *Wf, *Vb, *XiF, *cf;
    if (int rc = upload_vec(P.bounds, 0, &d_bounds, &g.allocs)) return rc;
    if (int rc = upload_vec(P.E, 0, &E, &g.allocs)) return rc;
    if (int rc = upload_vec(P.Wf, 0, &Wf, &g.allocs)) return rc;
    if (int rc = upload_vec(P.Vb, 0, &Vb, &g.allocs)) return rc;
    if (int rc = upload_vec(P.XiF, 0, &XiF, &g.allocs)) return rc;
    if (int rc = upload_vec(P.cf, 0, &cf, &g.allocs)) return rc;
    g.dev = SegDev{P.n, P.KL, P.KD, P.S, P.DF, P.DB, d_bounds, E, Wf, Vb, XiF, cf};
    g.bounds = P.bounds;
    g.local_lo = local_lo;
    g.local_cnt = local_cnt;
    g.local.resize(local_cnt);
    for (int i = 0; i < local_cnt; ++i) {
        // pass-A factor of segment s: the factor's own columns [a, b), pivots renumbered
        const int a = P.bounds[local_lo + i], b = P.bounds[local_lo + i + 1];
        std::vector<int> ip(D.ipiv.begin() + a, D.ipiv.begin() + b);
        for (int& v : ip) v -= a;
        SweepPlan L;
        if (int rc = build_sweep_plan(b - a, D.kl, D.ku, D.ldab, D.ab.data() + (size_t) a * D.ldab, ip.data(), SWEEP_CH, 1, L,
                                      P.piv != 0))
            return rc;
        if (int rc = upload_plan(L, g.local[i], g.allocs)) return rc;
        // the fused distributed sweep prefers shorter chunks (more threads per tile); only slabs use it
        if (local_cnt == 1 && L.KD <= SWEEP_CH_DIST) {
            SweepPlan Ld;
            if (build_sweep_plan(b - a, D.kl, D.ku, D.ldab, D.ab.data() + (size_t) a * D.ldab, ip.data(), SWEEP_CH_DIST, 1, Ld,
                                 P.piv != 0) == ADSB_OK && Ld.KL == L.KL && Ld.KD == L.KD) {
                g.local_dist.resize(1);
                if (int rc = upload_plan(Ld, g.local_dist[0], g.allocs)) return rc;
            }
        }
    }
    g.set = true;
    return ADSB_OK;
}

int device_sms() {
    static int sms = [] {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return sms;
}

// Lines too long for one CTA of the tile kernel: segments of <= 520 rows (29 chunks), cut where no row
// interchange crosses.  Silently leaves the factor unsegmented when that is impossible (growing boundary
// responses ...): the register-path kernel then handles the line as before.
void try_auto_segments(adsb_ctx* c, DevFactor& D) {
    D.auto_seg = false;
    static const bool on = [] {
        const char* e = getenv("ADSB_AUTO_SEGMENTS");
        return !e || atoi(e) != 0;
    }();
    if (!on || D.f.SC <= 32) return;
    // chunks per segment: as many as the tile kernel's shared memory holds (two ring slots of 16 lines, the
    // factor tables of the segment, the chunk states) -- 29 chunks = 522 rows for p <= 2, fewer for wider bands
    const int KL = D.f.KL, KD = D.f.KD;
    int mc = 29;
    for (; mc >= 6; --mc) {
        const size_t rows = (size_t) mc * SWEEP_CH + KL + KD;
        const size_t fixed = (size_t) mc * (KL + KD) * 16 + rows * (sweep_pitch(KL) + sweep_pitch(KD + 1) + sweep_pitch(KD + KL)) +
                             (size_t) mc * (KL * KL + KD * KD) * SWEEP_MAX_DEPTH_DEV;
        const size_t tile = ((size_t) mc * SWEEP_CH + KL + 8) * 16;
        if ((fixed + 2 * tile) * 8 + 1024 <= 222 * 1024) break;
    }
    if (mc < 6) return;
    const int seg_rows = mc * SWEEP_CH - 2;
    const int S = (D.n + seg_rows - 1) / seg_rows;
    std::vector<int> bounds(S + 1);
    const std::string keep = g_last_error;
    if (pick_segment_bounds(D.n, D.kl, D.ipiv.data(), S, 2, std::max(D.kl + D.ku, 2), bounds.data()) == ADSB_OK) {
        bool even = true;
        for (int s = 1; s < S; ++s) even = even && (bounds[s] % 2 == 0);
        if (even && set_segments_impl(c, D, S, bounds.data(), 0, S) == ADSB_OK) D.auto_seg = true;
    }
    if (!D.auto_seg) {
        free_segments(D.seg);
        g_last_error = keep;
    }
}

cudaEvent_t get_event(adsb_ctx* c) {
    if (!c->free_events.empty()) {
        cudaEvent_t e = c->free_events.back();
        c->free_events.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct StageTimer {
    adsb_ctx* c;
    TimedSpan s{};
    bool on;
    StageTimer(adsb_ctx* ctx, int stage) : c(ctx), on(ctx->timing) {
        if (on) {
            s.stage = stage;
            s.a = get_event(c);
            s.b = get_event(c);
            cudaEventRecord(s.a, c->stream);
        }
    }
    ~StageTimer() {
        if (on) {
            cudaEventRecord(s.b, c->stream);
            c->spans.push_back(s);
        }
    }
};

int sweep_max_threads() {  // tuning knob (threads per CTA of the sweep kernel)
    static int v = [] {
        const char* e = getenv("ADSB_SWEEP_THREADS");
        int t = e ? atoi(e) : 512;
        return t < 32 ? 32 : (t > 512 ? 512 : t);
    }();
    return v;
}

int pick_nl(int SC) {  // lanes per CTA (each thread carries SWEEP_RL lines)
    int nl = 32;
    while (nl > 2 && nl * SC > sweep_max_threads()) nl >>= 1;
    return nl;
}

int get_offsets(adsb_ctx* c, const long long* host, int n, const long long** dev) {
    *dev = nullptr;
    if (!host) return ADSB_OK;
    std::vector<long long> key(host, host + n);
    auto it = c->off_cache.find(key);
    if (it == c->off_cache.end()) {
        long long* d = nullptr;
        CU(cudaMalloc((void**) &d, n * sizeof(long long)));
        CU(cudaMemcpy(d, host, n * sizeof(long long), cudaMemcpyHostToDevice));
        it = c->off_cache.emplace(std::move(key), d).first;
    }
    *dev = it->second;
    return ADSB_OK;
}

int sweep_segmented(adsb_ctx* c, int axis, int slot, const double* in, const adsb_view& vi, double* out,
                    const adsb_view& vo, bool managed);

// Lines that follow one another in memory (unit stride along l0, l1 stride = L0: the z lines of an x-fastest
// tensor without row padding): number them flat, so tiles of NL lines never end half empty at a row end.
void flatten_lines(SweepGeom& G) {
    if (G.L1 > 1 && G.s0_in == 1 && G.s0_out == 1 && G.s1_in == G.L0 && G.s1_out == G.L0 &&
        (long long) G.L0 * G.L1 < (1ll << 30)) {
        G.L0 *= G.L1;
        G.L1 = 1;
    }
}

// sweep along `axis` of a view; see adsb_sweep_view
int sweep_impl(adsb_ctx* c, int axis, int slot, const double* in, const adsb_view& vi,
               const long long* off_in_h, double* out, const adsb_view& vo, const long long* off_out_h,
               bool managed = false, const SweepFactor* Fsel = nullptr) {
    if (axis < 0 || axis >= c->ndim) return fail(ADSB_EINVAL, "sweep: bad axis");
    if (slot < 0 || slot >= ADSB_MAX_SLOTS || !c->ax[axis].fac[slot].set)
        return fail(ADSB_ESTATE, "sweep: no factor uploaded for this axis/slot");
    const SweepFactor& F = Fsel ? *Fsel : c->ax[axis].fac[slot].f;
    if (!Fsel && c->ax[axis].fac[slot].auto_seg && !off_in_h && !off_out_h && vi.n[axis] == F.n)
        return sweep_segmented(c, axis, slot, in, vi, out, vo, managed);
    for (int d = 0; d < 3; ++d)
        if (vi.n[d] != vo.n[d]) return fail(ADSB_EINVAL, "sweep: in/out extents differ");
    if (vi.n[axis] != F.n) return fail(ADSB_EINVAL, "sweep: the view does not span the whole axis");
    const int a = (axis + 1) % 3, b = (axis + 2) % 3;
    const long long* off_in = nullptr;
    const long long* off_out = nullptr;
    if (int rc = get_offsets(c, off_in_h, F.n, &off_in)) return rc;
    if (int rc = get_offsets(c, off_out_h, F.n, &off_out)) return rc;
    SweepGeom G{};
    G.in = in;
    G.out = out;
    G.off_in = off_in;
    G.off_out = off_out;
    G.sj_in = vi.s[axis];
    G.sj_out = vo.s[axis];
    const bool contig = !off_in && !off_out && vi.s[axis] == 1 && vo.s[axis] == 1 && F.n > 1;
    int l0, l1;
    if (contig) {
        // lines enumerated along the perpendicular axis with the smaller stride first
        l0 = (vi.s[a] <= vi.s[b]) ? a : b;
    } else {
        // lanes along the perpendicular axis with unit stride (or the smaller one)
        l0 = (vi.s[a] <= vi.s[b]) ? a : b;
        if (vi.n[l0] == 1) l0 = (l0 == a) ? b : a;
    }
    l1 = (l0 == a) ? b : a;
    G.L0 = vi.n[l0];
    G.L1 = vi.n[l1];
    G.s0_in = vi.s[l0];
    G.s0_out = vo.s[l0];
    G.s1_in = vi.s[l1];
    G.s1_out = vo.s[l1];
    G.max_ctas = c->sm_limit;
    // The x sweep follows the right-hand side, which writes its planes in ascending z: walking the x tiles from
    // the last plane backwards starts on what is still in L2 (it matters when a tensor is not much larger than
    // the 126 MB L2: 256^3 problems, the slabs of an 8-GPU run); the y sweep then walks forwards again.
    static const bool reverse_x = [] {
        const char* e = getenv("ADSB_SWEEP_REVERSE_X");
        return !e || atoi(e) != 0;
    }();
    G.reverse = (reverse_x && axis == 0 && c->ndim == 3) ? 1 : 0;
    if (!contig && !off_in && !off_out) flatten_lines(G);
    // managed tensors pad every x row: an in-place x sweep of an odd-length line may move the pad along
    G.pad_ok = managed && in == out && axis == 0 && vi.s[1] > vi.n[0];
    G.pitch = F.n + (F.n & 1);
    if (G.pitch % 4 == 0) G.pitch += 2;  // pitch = 2 (mod 4): conflict-free 128-bit column access
    G.bulk = 0;
    static const bool use_tile = [] {
        const char* e = getenv("ADSB_SWEEP_TILE");
        return !e || atoi(e) != 0;
    }();
    if (use_tile) {
        StageTimer t(c, 1 + axis);
        const int rc = launch_sweep_tile(F, G, contig, off_in_h, off_out_h, c->stream);
        if (rc == 0) {
            c->launches++;
            return ADSB_OK;
        }
        if (rc > 0) return cuda_fail((cudaError_t) rc, "sweep tile kernel launch");
    }
    int NL = pick_nl(F.SC);
    if (contig) {
        while (NL > 1 && sweep_smem_bytes(F, true, NL, G.pitch) > 200 * 1024) NL >>= 1;
        if (sweep_smem_bytes(F, true, NL, G.pitch) > 220 * 1024)
            return fail(ADSB_EINVAL, "sweep: line too long for the shared-memory staged x sweep");
        const bool aligned = ((uintptr_t) in % 16 == 0) && ((uintptr_t) out % 16 == 0) && F.n % 2 == 0 &&
                             G.s0_in % 2 == 0 && G.s1_in % 2 == 0 && G.s0_out % 2 == 0 && G.s1_out % 2 == 0;
        G.bulk = aligned ? 1 : 0;
    }
    if (NL * F.SC > 512) return fail(ADSB_EINVAL, "sweep: axis too long for the compiled chunking");
    if (G.L1 > 65535) return fail(ADSB_EINVAL, "sweep: outer extent beyond grid limits");
    StageTimer t(c, 1 + axis);
    cudaError_t e = (cudaError_t) launch_sweep(F, G, contig, NL, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "sweep kernel launch");
    c->launches++;
    return ADSB_OK;
}

adsb_view local_view(const adsb_ctx* c) {
    adsb_view v{};
    for (int d = 0; d < 3; ++d) v.n[d] = c->cnt[d];
    v.s[0] = 1;
    v.s[1] = c->pitch0();
    v.s[2] = c->pitch0() * c->cnt[1];
    return v;
}

int quad_axes(adsb_ctx* c, QuadAxes& A) {
    A = QuadAxes{};
    A.ndim = c->ndim;
    for (int d = 0; d < c->ndim; ++d) {
        const AxisData& a = c->ax[d];
        if (!a.tables) return fail(ADSB_ESTATE, "axis tables not uploaded");
        A.p[d] = a.p;
        A.q[d] = a.q;
        A.ne[d] = a.elements;
        A.st[d] = (a.ders + 1) * (a.p + 1);
        A.bt[d] = a.d_bt;
        A.xq[d] = a.d_xq;
        A.w[d] = a.d_w;
        A.J[d] = a.d_J;
    }
    return ADSB_OK;
}


// General quadrature path (brick kernel): out = gamma * F + the quadrature sums of every element touching the box.
int rhs_quadrature_impl(adsb_ctx* c, const adsb_form& f, const double* in, const adsb_view& vi, const int* in_lo,
                        const double* forcing, double* out, const adsb_view& vo, const int* out_lo) {
    QuadAxes A;
    if (int rc = quad_axes(c, A)) return rc;
    RhsGeom g{};
    g.in = in;
    g.out = out;
    g.forcing = nullptr;
    int elo[3] = {0, 0, 0}, en[3] = {1, 1, 1};
    for (int d = 0; d < 3; ++d) {
        g.si[d] = vi.s[d];
        g.so[d] = vo.s[d];
        g.in_lo[d] = in_lo[d];
        g.in_n[d] = vi.n[d];
        g.out_lo[d] = out_lo[d];
        g.out_n[d] = vo.n[d];
        g.beta[d] = d < c->ndim ? f.beta[d] : 0.0;
        if (d < c->ndim) {
            const int p = c->ax[d].p;
            if (c->ax[d].q != p + 1 || c->ax[d].ders != 1)
                return fail(ADSB_EINVAL, "compute_rhs(quadrature): needs quad_order = p + 1 and derivatives = 1");
            elo[d] = std::max(0, out_lo[d] - p);
            const int ehi = std::min(c->ax[d].elements - 1, out_lo[d] + vo.n[d] - 1);
            en[d] = ehi - elo[d] + 1;
            if (in_lo[d] > elo[d] || in_lo[d] + vi.n[d] < ehi + p + 1)
                return fail(ADSB_EINVAL, "compute_rhs: input box lacks the p-wide halo of the output box");
        }
    }
    g.alpha = f.alpha;
    g.gamma = f.gamma;
    if (f.source < 0 || f.source > 1) return fail(ADSB_EINVAL, "compute_rhs: unknown source");
    StageTimer t(c, 0);
    // brick kernel: out = gamma F (load tensor, if any) + quadrature sums; a built-in source without a load
    // tensor is evaluated at the points and enters without the test function (test3d.hpp:86-88)
    PointFormArgs pf{};
    pf.kind = 0;
    pf.alpha = f.alpha;
    for (int d = 0; d < 3; ++d) pf.beta[d] = g.beta[d];
    pf.gamma = f.gamma;
    pf.source = (!forcing && f.gamma != 0.0) ? f.source : 0;
    pf.plain = 1;
    g.forcing = forcing;
    g.max_sms = c->sm_limit;
    int nl = 0;
    cudaError_t e = (cudaError_t) launch_rhs_brick(c->ndim, A, g, pf, elo, en, nullptr, c->stream, &nl);
    c->launches += nl;
    if (e != cudaSuccess) return cuda_fail(e, "quadrature rhs kernels");
    return ADSB_OK;
}

int rhs_impl(adsb_ctx* c, const adsb_form& f, const double* in, const adsb_view& vi, const int* in_lo,
             const double* forcing, double* out, const adsb_view& vo, const int* out_lo) {
    if (f.method != ADSB_RHS_COLLAPSED && f.method != ADSB_RHS_QUADRATURE)
        return fail(ADSB_EINVAL, "compute_rhs: unknown method");
    for (int d = 0; d < c->ndim; ++d)
        if (!c->ax[d].tables) return fail(ADSB_ESTATE, "compute_rhs: axis tables not uploaded");
    if (f.method == ADSB_RHS_QUADRATURE) return rhs_quadrature_impl(c, f, in, vi, in_lo, forcing, out, vo, out_lo);
    RhsOps ops{};
    ops.Mx = c->ax[0].d_M;
    ops.Sx = c->ax[0].d_S;
    ops.My = c->ax[1].d_M;
    ops.Sy = c->ax[1].d_S;
    ops.MSzT = c->ax[c->ndim - 1].d_MST;  // column table of the marching axis (z in 3-D, y in 2-D)
    RhsGeom g{};
    g.in = in;
    g.out = out;
    g.forcing = (f.gamma != 0.0) ? forcing : nullptr;
    for (int d = 0; d < 3; ++d) {
        ops.p[d] = d < c->ndim ? c->ax[d].p : 0;
        ops.n[d] = c->ng[d];
        g.si[d] = vi.s[d];
        g.so[d] = vo.s[d];
        g.in_lo[d] = in_lo[d];
        g.in_n[d] = vi.n[d];
        g.out_lo[d] = out_lo[d];
        g.out_n[d] = vo.n[d];
        g.beta[d] = d < c->ndim ? f.beta[d] : 0.0;
        if (d < c->ndim) {
            // the input box must cover the output box widened by p, clipped to the domain
            const int need_lo = std::max(0, out_lo[d] - c->ax[d].p);
            const int need_hi = std::min(c->ng[d], out_lo[d] + vo.n[d] + c->ax[d].p);
            if (in_lo[d] > need_lo || in_lo[d] + vi.n[d] < need_hi)
                return fail(ADSB_EINVAL, "compute_rhs: input box lacks the p-wide halo of the output box");
        }
    }
    g.alpha = f.alpha;
    g.gamma = f.gamma;
    g.max_sms = c->sm_limit;
    StageTimer t(c, 0);
    int nl = 1;
    if (!c->side.side) {
        CU(cudaStreamCreateWithFlags(&c->side.side, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->side.fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->side.join, cudaEventDisableTiming));
    }
    cudaError_t e = (cudaError_t) launch_rhs_collapsed(c->ndim, ops, g, c->stream, &nl, &c->side);
    if (e != cudaSuccess) return cuda_fail(e, "rhs kernel launch");
    c->launches += nl;
    return ADSB_OK;
}

int rows_from_band(int n, int p, const std::vector<double>& ab, std::vector<double>& rows) {
    const int W = 2 * p + 1, ldab = 3 * p + 1;
    rows.assign((size_t) n * W, 0.0);
    for (int i = 0; i < n; ++i)
        for (int m = 0; m < W; ++m) {
            const int j = i - p + m;
            if (j >= 0 && j < n) rows[(size_t) i * W + m] = ab[(size_t) j * ldab + 2 * p + i - j];
        }
    return ADSB_OK;
}

// column table: out[k][d] = A(k - p + d, k)
int cols_from_band(int n, int p, const std::vector<double>& ab, std::vector<double>& cols) {
    const int W = 2 * p + 1, ldab = 3 * p + 1;
    cols.assign((size_t) n * W, 0.0);
    for (int k = 0; k < n; ++k)
        for (int d = 0; d < W; ++d) {
            const int i = k - p + d;
            if (i >= 0 && i < n) cols[(size_t) k * W + d] = ab[(size_t) k * ldab + 2 * p + i - k];
        }
    return ADSB_OK;
}

}  // namespace

extern "C" {

int adsb_create(int ndim, const int* n_global, const int* lo, const int* cnt, int device, adsb_ctx** out) {
    if (!out || !n_global || (ndim != 2 && ndim != 3)) return fail(ADSB_EINVAL, "create: ndim must be 2 or 3");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ADSB_ENODEVICE, "create: no CUDA device (libadsb200 has no CPU fallback)");
    if (device < 0 || device >= count) return fail(ADSB_EINVAL, "create: bad device ordinal");
    auto* c = new adsb_ctx;
    c->ndim = ndim;
    c->device = device;
    for (int d = 0; d < ndim; ++d) {
        c->ng[d] = n_global[d];
        c->lo[d] = lo ? lo[d] : 0;
        c->cnt[d] = cnt ? cnt[d] : n_global[d];
        if (c->ng[d] < 1 || c->lo[d] < 0 || c->cnt[d] < 1 || c->lo[d] + c->cnt[d] > c->ng[d]) {
            delete c;
            return fail(ADSB_EINVAL, "create: bad extents");
        }
    }
    if (int rc = select_device(c)) {
        delete c;
        return rc;
    }
    *out = c;
    return ADSB_OK;
}

int adsb_destroy(adsb_ctx* c) {
    if (!c) return ADSB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& a : c->ax) {
        cudaFree(a.d_M);
        cudaFree(a.d_S);
        cudaFree(a.d_MST);
        cudaFree(a.d_bt);
        cudaFree(a.d_xq);
        cudaFree(a.d_wJ);
        cudaFree(a.d_w);
        cudaFree(a.d_J);
        for (auto& f : a.fac) free_factor(f);
    }
    for (int b = 0; b < ADSB_MAX_BUFFERS; ++b)
        if (c->owned[b]) cudaFree(c->buf[b]);
    for (auto& kv : c->off_cache) cudaFree(kv.second);
    for (auto& s : c->spans) {
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    for (auto e : c->free_events) cudaEventDestroy(e);
    for (auto st : c->seg_streams) cudaStreamDestroy(st);
    for (auto e : c->seg_events) cudaEventDestroy(e);
    cudaFree(c->seg_scratch);
    cudaFree(c->point_coef);
    for (auto& lf : c->line_fac) {
        cudaFree(lf.ab);
        cudaFree(lf.ipiv);
        cudaFree(lf.scratch);
    }
    if (c->side.side) {
        cudaStreamDestroy(c->side.side);
        cudaEventDestroy(c->side.fork);
        cudaEventDestroy(c->side.join);
    }
    delete c;
    return ADSB_OK;
}

int adsb_set_sm_limit(adsb_ctx* c, int sms) {
    if (!c || sms < 0) return fail(ADSB_EINVAL, "set_sm_limit: bad argument");
    c->sm_limit = sms;
    return ADSB_OK;
}

int adsb_copy2d(adsb_ctx* c, void* dst, long long dpitch, const void* src, long long spitch, long long width,
                long long height) {
    if (!c || !dst || !src || width < 0 || height < 0 || dpitch < width || spitch < width)
        return fail(ADSB_EINVAL, "copy2d: bad argument");
    if (int rc = select_device(c)) return rc;
    if (width == 0 || height == 0) return ADSB_OK;
    CU(cudaMemcpy2DAsync(dst, (size_t) dpitch, src, (size_t) spitch, (size_t) width, (size_t) height, cudaMemcpyDefault,
                         c->stream));
    return ADSB_OK;
}

int adsb_set_stream(adsb_ctx* c, void* s) {
    if (!c) return fail(ADSB_EINVAL, "null context");
    c->stream = (cudaStream_t) s;
    return ADSB_OK;
}

int adsb_synchronize(adsb_ctx* c) {
    if (!c) return fail(ADSB_EINVAL, "null context");
    CU(cudaStreamSynchronize(c->stream));
    return ADSB_OK;
}

int adsb_set_axis_tables(adsb_ctx* c, int axis, int p, int elements, int q, int ders, const double* b_flat,
                         const double* xq, const double* w, const double* J, const int* first_dof) {
    if (!c || axis < 0 || axis >= c->ndim) return fail(ADSB_EINVAL, "set_axis_tables: bad axis");
    if (p < 1 || p > 5) return fail(ADSB_EINVAL, "set_axis_tables: device kernels are built for 1 <= p <= 5");
    if (elements + p != c->ng[axis]) return fail(ADSB_EINVAL, "set_axis_tables: elements + p != n_global[axis]");
    if (ders < 1) return fail(ADSB_EINVAL, "set_axis_tables: first derivatives are required");
    if (!b_flat || !xq || !w || !J || !first_dof || q < 1 || elements < 1)
        return fail(ADSB_EINVAL, "set_axis_tables: null table / bad sizes");
    for (int e = 0; e < elements; ++e)
        if (first_dof[e] != e) return fail(ADSB_EINVAL, "set_axis_tables: repeated knots are not supported");
    if (int rc = select_device(c)) return rc;
    AxisData& a = c->ax[axis];
    a.p = p;
    a.elements = elements;
    a.q = q;
    a.ders = ders;
    a.n = elements + p;
    const size_t nb = (size_t) elements * q * (ders + 1) * (p + 1);
    a.bt.assign(b_flat, b_flat + nb);
    a.xq.assign(xq, xq + (size_t) elements * q);
    a.w.assign(w, w + q);
    a.J.assign(J, J + elements);
    std::vector<double> ab((size_t) (3 * p + 1) * a.n), rows;
    CU(cudaStreamSynchronize(c->stream));  // kernels in flight may still read the tables replaced below
    cudaFree(a.d_M);
    cudaFree(a.d_S);
    cudaFree(a.d_MST);
    a.d_MST = nullptr;
    cudaFree(a.d_bt);
    cudaFree(a.d_xq);
    cudaFree(a.d_wJ);
    cudaFree(a.d_w);
    cudaFree(a.d_J);
    a.d_M = a.d_S = a.d_bt = a.d_xq = a.d_wJ = a.d_w = a.d_J = nullptr;
    if (int rc = matrix_from_tables(0, 0.0, p, elements, q, ders, b_flat, w, J, ab.data())) return rc;
    const int W = 2 * p + 1, WP = W + 1;
    std::vector<double> mst((size_t) (a.n + 2 * p) * 2 * WP, 0.0);  // p zero rows on both sides
    rows_from_band(a.n, p, ab, rows);
    if (int rc = upload_vec(rows, 0, &a.d_M, nullptr)) return rc;
    cols_from_band(a.n, p, ab, rows);
    for (int k = 0; k < a.n; ++k)
        for (int d = 0; d < W; ++d) mst[(size_t) (k + p) * 2 * WP + d] = rows[(size_t) k * W + d];
    if (int rc = matrix_from_tables(1, 0.0, p, elements, q, ders, b_flat, w, J, ab.data())) return rc;
    rows_from_band(a.n, p, ab, rows);
    if (int rc = upload_vec(rows, 0, &a.d_S, nullptr)) return rc;
    cols_from_band(a.n, p, ab, rows);
    for (int k = 0; k < a.n; ++k)
        for (int d = 0; d < W; ++d) mst[(size_t) (k + p) * 2 * WP + WP + d] = rows[(size_t) k * W + d];
    if (int rc = upload_vec(mst, 0, &a.d_MST, nullptr)) return rc;
    if (int rc = upload_vec(a.bt, 0, &a.d_bt, nullptr)) return rc;
    if (int rc = upload_vec(a.xq, 0, &a.d_xq, nullptr)) return rc;
    std::vector<double> wJ((size_t) elements * q);
    for (int e = 0; e < elements; ++e)
        for (int k = 0; k < q; ++k) wJ[(size_t) e * q + k] = w[k] * J[e];
    if (int rc = upload_vec(wJ, 0, &a.d_wJ, nullptr)) return rc;
    if (int rc = upload_vec(a.w, 0, &a.d_w, nullptr)) return rc;
    if (int rc = upload_vec(a.J, 0, &a.d_J, nullptr)) return rc;
    a.tables = true;
    return ADSB_OK;
}

int adsb_set_axis_factor(adsb_ctx* c, int axis, int slot, int n, int kl, int ku, int ldab, const double* ab,
                         const int* ipiv) {
    if (!c || axis < 0 || axis >= c->ndim) return fail(ADSB_EINVAL, "set_axis_factor: bad axis");
    if (slot < 0 || slot >= ADSB_MAX_SLOTS) return fail(ADSB_EINVAL, "set_axis_factor: bad slot");
    if (n != c->ng[axis]) return fail(ADSB_EINVAL, "set_axis_factor: n != n_global[axis]");
    if (kl < 0 || ku < 0 || ldab < 2 * kl + ku + 1) return fail(ADSB_EINVAL, "set_axis_factor: bad band shape");
    if (int rc = select_device(c)) return rc;
    static_assert(SWEEP_MAX_DEPTH == SWEEP_MAX_DEPTH_DEV, "depth constants out of sync");
    // factors with row interchanges are eliminated again without them when that is stable (host_setup.cpp:
    // refactor_without_pivoting); ADSB_UNPIVOT=0 keeps the caller's interchanges
    static const bool unpivot = [] {
        const char* e = getenv("ADSB_UNPIVOT");
        return !e || atoi(e) != 0;
    }();
    std::vector<double> ab2;
    std::vector<int> ipiv2;
    bool has_piv = false;
    for (int j = 0; j < n; ++j) has_piv = has_piv || ipiv[j] != j + 1;
    if (unpivot && has_piv && refactor_without_pivoting(n, kl, ku, ldab, ab, ipiv, ab2)) {
        ipiv2.resize(n);
        for (int j = 0; j < n; ++j) ipiv2[j] = j + 1;
        ab = ab2.data();
        ipiv = ipiv2.data();
    }
    SweepPlan P;
    if (int rc = build_sweep_plan(n, kl, ku, ldab, ab, ipiv, SWEEP_CH, 1, P)) return rc;
    DevFactor& D = c->ax[axis].fac[slot];
    CU(cudaStreamSynchronize(c->stream));
    free_factor(D);
    if (int rc = upload_plan(P, D.f, D.allocs)) return rc;
    D.n = n;
    D.kl = kl;
    D.ku = ku;
    D.ldab = ldab;
    D.ab.assign(ab, ab + (size_t) ldab * n);
    D.ipiv.assign(ipiv, ipiv + n);
    D.set = true;
    try_auto_segments(c, D);
    D.set = true;
    return ADSB_OK;
}

int adsb_upload(adsb_ctx* c, int b, const double* host) {
    if (!c || !host) return fail(ADSB_EINVAL, "upload: null argument");
    if (int rc = select_device(c)) return rc;
    if (int rc = ensure_buf(c, b)) return rc;
    if (c->pitch0() == c->cnt[0])
        CU(cudaMemcpyAsync(c->buf[b], host, c->local_size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    else
        CU(cudaMemcpy2DAsync(c->buf[b], c->pitch0() * sizeof(double), host, c->cnt[0] * sizeof(double),
                             c->cnt[0] * sizeof(double), c->rows(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return ADSB_OK;
}

int adsb_download(adsb_ctx* c, int b, double* host) {
    if (!c || !host) return fail(ADSB_EINVAL, "download: null argument");
    if (b < 0 || b >= ADSB_MAX_BUFFERS || !c->buf[b]) return fail(ADSB_ESTATE, "download: buffer not allocated");
    if (int rc = select_device(c)) return rc;
    if (c->pitch0() == c->cnt[0])
        CU(cudaMemcpyAsync(host, c->buf[b], c->local_size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    else
        CU(cudaMemcpy2DAsync(host, c->cnt[0] * sizeof(double), c->buf[b], c->pitch0() * sizeof(double),
                             c->cnt[0] * sizeof(double), c->rows(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return ADSB_OK;
}

// Same copies without the trailing synchronise, on a caller-supplied stream (cudaStream_t as void*): lets a
// caller overlap the upload of the next input and the download of the previous result with a step that runs
// on the context's own stream.  The caller orders the streams (events) and keeps `host` pinned and alive.
int adsb_upload_async(adsb_ctx* c, int b, const double* host, void* stream) {
    if (!c || !host) return fail(ADSB_EINVAL, "upload_async: null argument");
    if (int rc = select_device(c)) return rc;
    if (int rc = ensure_buf(c, b)) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    if (c->pitch0() == c->cnt[0])
        CU(cudaMemcpyAsync(c->buf[b], host, c->local_size() * sizeof(double), cudaMemcpyHostToDevice, st));
    else
        CU(cudaMemcpy2DAsync(c->buf[b], c->pitch0() * sizeof(double), host, c->cnt[0] * sizeof(double),
                             c->cnt[0] * sizeof(double), c->rows(), cudaMemcpyHostToDevice, st));
    return ADSB_OK;
}

int adsb_download_async(adsb_ctx* c, int b, double* host, void* stream) {
    if (!c || !host) return fail(ADSB_EINVAL, "download_async: null argument");
    if (b < 0 || b >= ADSB_MAX_BUFFERS || !c->buf[b]) return fail(ADSB_ESTATE, "download_async: buffer not allocated");
    if (int rc = select_device(c)) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    if (c->pitch0() == c->cnt[0])
        CU(cudaMemcpyAsync(host, c->buf[b], c->local_size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    else
        CU(cudaMemcpy2DAsync(host, c->cnt[0] * sizeof(double), c->buf[b], c->pitch0() * sizeof(double),
                             c->cnt[0] * sizeof(double), c->rows(), cudaMemcpyDeviceToHost, st));
    return ADSB_OK;
}

int adsb_swap(adsb_ctx* c, int a, int b) {
    if (!c || a < 0 || b < 0 || a >= ADSB_MAX_BUFFERS || b >= ADSB_MAX_BUFFERS)
        return fail(ADSB_EINVAL, "swap: bad buffer id");
    std::swap(c->buf[a], c->buf[b]);
    std::swap(c->owned[a], c->owned[b]);
    return ADSB_OK;
}

int adsb_zero(adsb_ctx* c, int b) {
    if (!c) return fail(ADSB_EINVAL, "null context");
    if (int rc = select_device(c)) return rc;
    if (int rc = ensure_buf(c, b)) return rc;
    CU(cudaMemsetAsync(c->buf[b], 0, c->local_size() * sizeof(double), c->stream));
    return ADSB_OK;
}

int adsb_bind(adsb_ctx* c, int b, double* p) {
    if (!c || b < 0 || b >= ADSB_MAX_BUFFERS) return fail(ADSB_EINVAL, "bind: bad buffer id");
    if (c->owned[b]) cudaFree(c->buf[b]);
    c->buf[b] = p;
    c->owned[b] = false;
    return ADSB_OK;
}

long long adsb_row_pitch(adsb_ctx* c) { return c ? c->pitch0() : 0; }

double* adsb_device_ptr(adsb_ctx* c, int b) {
    if (!c || b < 0 || b >= ADSB_MAX_BUFFERS) return nullptr;
    if (select_device(c) || ensure_buf(c, b)) return nullptr;
    return c->buf[b];
}

int adsb_set_plane(adsb_ctx* c, int b, int axis, int idx, const double* values) {
    if (!c || axis < 0 || axis >= c->ndim || idx < 0 || idx >= c->cnt[axis])
        return fail(ADSB_EINVAL, "set_plane: bad axis/index");
    if (b < 0 || b >= ADSB_MAX_BUFFERS || !c->buf[b]) return fail(ADSB_ESTATE, "set_plane: buffer not allocated");
    if (int rc = select_device(c)) return rc;
    adsb_view v = local_view(c);
    size_t count = (size_t) c->cnt[0] * c->cnt[1] * c->cnt[2] / c->cnt[axis];
    if (!values) return fail(ADSB_EINVAL, "set_plane: null values");
    double* d = nullptr;
    CU(cudaMalloc((void**) &d, count * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, values, count * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = (cudaError_t) launch_set_plane(c->buf[b], v.s, v.n, axis, idx, d, c->stream);
    c->launches++;
    const cudaError_t e2 = cudaStreamSynchronize(c->stream);
    cudaFree(d);  // on every path
    if (e != cudaSuccess) return cuda_fail(e, "set_plane kernel");
    if (e2 != cudaSuccess) return cuda_fail(e2, "set_plane kernel");
    return ADSB_OK;
}

int adsb_compute_rhs(adsb_ctx* c, const adsb_form* f, int src, int dst) {
    if (!c || !f) return fail(ADSB_EINVAL, "compute_rhs: null argument");
    if (src == dst) return fail(ADSB_EINVAL, "compute_rhs: src and dst must differ");
    if (src < 0 || src >= ADSB_MAX_BUFFERS || !c->buf[src]) return fail(ADSB_ESTATE, "compute_rhs: src not allocated");
    if (int rc = select_device(c)) return rc;
    if (int rc = ensure_buf(c, dst)) return rc;
    const double* forcing = nullptr;
    const bool pointwise_source = f->method == ADSB_RHS_QUADRATURE && f->source != 0 && f->forcing_buf < 0;
    if (f->gamma != 0.0 && !pointwise_source) {
        if (f->forcing_buf < 0 || f->forcing_buf >= ADSB_MAX_BUFFERS || !c->buf[f->forcing_buf])
            return fail(ADSB_ESTATE, "compute_rhs: forcing buffer not allocated");
        forcing = c->buf[f->forcing_buf];
    }
    adsb_view v = local_view(c);
    return rhs_impl(c, *f, c->buf[src], v, c->lo, forcing, c->buf[dst], v, c->lo);
}

// ---- general pointwise forms (brick quadrature kernel, quadbrick.cuh)
int adsb_set_point_coefficient(adsb_ctx* c, const double* values) {
    if (!c || !values) return fail(ADSB_EINVAL, "set_point_coefficient: null argument");
    if (int rc = select_device(c)) return rc;
    size_t count = 1;
    for (int d = 0; d < c->ndim; ++d) {
        if (!c->ax[d].tables) return fail(ADSB_ESTATE, "set_point_coefficient: axis tables not uploaded");
        if (c->cnt[d] != c->ng[d]) return fail(ADSB_ESTATE, "set_point_coefficient: the context must own the whole domain");
        count *= (size_t) c->ax[d].elements * c->ax[d].q;
    }
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(c->point_coef);
    c->point_coef = nullptr;
    CU(cudaMalloc((void**) &c->point_coef, count * sizeof(double)));
    CU(cudaMemcpy(c->point_coef, values, count * sizeof(double), cudaMemcpyHostToDevice));
    return ADSB_OK;
}

int adsb_compute_rhs_pointwise(adsb_ctx* c, const adsb_point_form* f, int src, int dst) {
    if (!c || !f) return fail(ADSB_EINVAL, "compute_rhs_pointwise: null argument");
    if (src == dst) return fail(ADSB_EINVAL, "compute_rhs_pointwise: src and dst must differ");
    if (src < 0 || src >= ADSB_MAX_BUFFERS || !c->buf[src]) return fail(ADSB_ESTATE, "compute_rhs_pointwise: src not allocated");
    if (f->kind != ADSB_POINT_LINEAR && f->kind != ADSB_POINT_FLOW) return fail(ADSB_EINVAL, "compute_rhs_pointwise: unknown form");
    if (f->source < 0 || f->source > 2) return fail(ADSB_EINVAL, "compute_rhs_pointwise: unknown source");
    if (f->kind == ADSB_POINT_FLOW) {
        if (c->ndim != 3) return fail(ADSB_EINVAL, "compute_rhs_pointwise: the flow form is 3-D");
        if (c->ax[0].p > 3) return fail(ADSB_EINVAL, "compute_rhs_pointwise: the flow form is built for p <= 3");
        if (!c->point_coef) return fail(ADSB_ESTATE, "compute_rhs_pointwise: no coefficient table (adsb_set_point_coefficient)");
    }
    if (int rc = select_device(c)) return rc;
    if (int rc = ensure_buf(c, dst)) return rc;
    QuadAxes A;
    if (int rc = quad_axes(c, A)) return rc;
    const adsb_view v = local_view(c);
    RhsGeom g{};
    g.in = c->buf[src];
    g.out = c->buf[dst];
    int elo[3] = {0, 0, 0}, en[3] = {1, 1, 1};
    for (int d = 0; d < 3; ++d) {
        g.si[d] = g.so[d] = v.s[d];
        g.in_lo[d] = g.out_lo[d] = c->lo[d];
        g.in_n[d] = g.out_n[d] = v.n[d];
        if (d < c->ndim) {
            if (c->cnt[d] != c->ng[d]) return fail(ADSB_ESTATE, "compute_rhs_pointwise: the context must own the whole domain");
            if (c->ax[d].q != c->ax[d].p + 1 || c->ax[d].ders != 1 || c->ax[d].p != c->ax[0].p)
                return fail(ADSB_EINVAL, "compute_rhs_pointwise: needs the same p on every axis, quad_order = p + 1, derivatives = 1");
            en[d] = c->ax[d].elements;
        }
    }
    if (f->forcing_buf >= 0) {
        if (f->forcing_buf >= ADSB_MAX_BUFFERS || !c->buf[f->forcing_buf]) return fail(ADSB_ESTATE, "compute_rhs_pointwise: forcing buffer not allocated");
        g.forcing = c->buf[f->forcing_buf];
        g.gamma = f->forcing_scale;
    }
    g.max_sms = c->sm_limit;
    PointFormArgs pf{};
    pf.kind = f->kind;
    pf.alpha = f->alpha;
    for (int d = 0; d < 3; ++d) {
        pf.beta[d] = d < c->ndim ? f->beta[d] : 0.0;
        pf.adv[d] = d < c->ndim ? f->adv[d] : 0.0;
    }
    pf.gamma = f->gamma;
    pf.source = f->gamma != 0.0 ? f->source : 0;
    pf.plain = f->source_plain;
    for (int i = 0; i < 4; ++i) pf.par[i] = f->par[i];
    StageTimer t(c, 0);
    int nl = 0;
    cudaError_t e = (cudaError_t) launch_rhs_brick(c->ndim, A, g, pf, elo, en, c->point_coef, c->stream, &nl);
    c->launches += nl;
    if (e != cudaSuccess) return cuda_fail(e, "pointwise rhs kernels");
    return ADSB_OK;
}

// ---- generalised ADS: one axis with a different factor per line (include/ads/solver.hpp:56-96,:170-195)
int adsb_set_line_factors(adsb_ctx* c, int axis, int kl, int ku, const double* ab_lines, const int* ipiv_lines) {
    if (!c || !ab_lines || !ipiv_lines) return fail(ADSB_EINVAL, "set_line_factors: null argument");
    if (axis < 0 || axis >= c->ndim) return fail(ADSB_EINVAL, "set_line_factors: bad axis");
    if (kl != ku || kl < 1 || kl > 5) return fail(ADSB_EINVAL, "set_line_factors: needs kl = ku = p, 1 <= p <= 5");
    for (int d = 0; d < c->ndim; ++d)
        if (c->cnt[d] != c->ng[d]) return fail(ADSB_ESTATE, "set_line_factors: the context must own the whole domain");
    if (int rc = select_device(c)) return rc;
    const int n = c->ng[axis], ld = 2 * kl + ku + 1;
    long long lines = 1;
    for (int d = 0; d < c->ndim; ++d)
        if (d != axis) lines *= c->ng[d];
    // [line][j][r] -> [j][r][line]; pivots 1-based -> 0-based rows
    std::vector<double> ab((size_t) lines * n * ld);
    std::vector<int> pv((size_t) lines * n);
    for (long long l = 0; l < lines; ++l)
        for (int j = 0; j < n; ++j) {
            for (int r = 0; r < ld; ++r) ab[((size_t) j * ld + r) * lines + l] = ab_lines[((size_t) l * n + j) * ld + r];
            const int pr = ipiv_lines[(size_t) l * n + j] - 1;
            if (pr < j || pr > std::min(n - 1, j + kl)) return fail(ADSB_EINVAL, "set_line_factors: pivot index out of range");
            pv[(size_t) j * lines + l] = pr;
        }
    auto& lf = c->line_fac[axis];
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(lf.ab);
    cudaFree(lf.ipiv);
    cudaFree(lf.scratch);
    lf = adsb_ctx::LineFactors{};
    CU(cudaMalloc((void**) &lf.ab, ab.size() * sizeof(double)));
    CU(cudaMalloc((void**) &lf.ipiv, pv.size() * sizeof(int)));
    CU(cudaMemcpy(lf.ab, ab.data(), ab.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(lf.ipiv, pv.data(), pv.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (axis == 0) CU(cudaMalloc((void**) &lf.scratch, (size_t) ((lines + 31) / 32) * 32 * n * sizeof(double)));
    lf.p = kl;
    lf.n = n;
    lf.lines = lines;
    return ADSB_OK;
}

int adsb_solve_special(adsb_ctx* c, int b, int special_axis, const int* slots) {
    if (!c) return fail(ADSB_EINVAL, "null context");
    if (b < 0 || b >= ADSB_MAX_BUFFERS || !c->buf[b]) return fail(ADSB_ESTATE, "solve_special: buffer not allocated");
    if (special_axis < 0 || special_axis >= c->ndim) return fail(ADSB_EINVAL, "solve_special: bad axis");
    const auto& lf = c->line_fac[special_axis];
    if (!lf.ab) return fail(ADSB_ESTATE, "solve_special: no line factors on this axis (adsb_set_line_factors)");
    if (int rc = select_device(c)) return rc;
    const long long pitch = c->pitch0(), ny = c->cnt[1];
    {
        StageTimer t(c, 1 + special_axis);
        int L0 = c->cnt[0];
        long long s0 = 1, s1 = 0, sj = 0;
        if (special_axis == 0) {
            s0 = pitch;  // rows are the lines
        } else if (special_axis == 1) {
            s1 = pitch * ny;
            sj = pitch;
        } else {
            s1 = pitch;
            sj = pitch * ny;
        }
        cudaError_t e = (cudaError_t) launch_line_sweep(lf.p, c->buf[b], lf.ab, lf.ipiv, lf.n, lf.lines, special_axis, L0, s0, s1,
                                                        sj, lf.scratch, c->stream);
        c->launches++;
        if (e != cudaSuccess) return cuda_fail(e, "line sweep kernel");
    }
    // the other axes, in order, with their ordinary factors (the reference solves the special dimension first)
    for (int d = 0; d < c->ndim; ++d) {
        if (d == special_axis) continue;
        if (int rc = adsb_sweep(c, b, d, slots ? slots[d] : 0)) return rc;
    }
    return ADSB_OK;
}

int adsb_rhs_view(adsb_ctx* c, const adsb_form* f, const double* in, const adsb_view* vin, const int* in_lo,
                  const double* forcing, double* out, const adsb_view* vout, const int* out_lo) {
    if (!c || !f || !in || !out || !vin || !vout || !in_lo || !out_lo)
        return fail(ADSB_EINVAL, "rhs_view: null argument");
    if (int rc = select_device(c)) return rc;
    return rhs_impl(c, *f, in, *vin, in_lo, forcing, out, *vout, out_lo);
}

int adsb_sweep_view(adsb_ctx* c, int axis, int slot, const double* in, const adsb_view* vin,
                    const long long* row_off_in, double* out, const adsb_view* vout, const long long* row_off_out) {
    if (!c || !in || !out || !vin || !vout) return fail(ADSB_EINVAL, "sweep_view: null argument");
    if (int rc = select_device(c)) return rc;
    return sweep_impl(c, axis, slot, in, *vin, row_off_in, out, *vout, row_off_out);
}

int adsb_sweep(adsb_ctx* c, int b, int axis, int slot) {
    if (!c) return fail(ADSB_EINVAL, "null context");
    if (b < 0 || b >= ADSB_MAX_BUFFERS || !c->buf[b]) return fail(ADSB_ESTATE, "sweep: buffer not allocated");
    if (axis < 0 || axis >= c->ndim) return fail(ADSB_EINVAL, "sweep: bad axis");
    if (c->cnt[axis] != c->ng[axis]) return fail(ADSB_ESTATE, "sweep: this context does not own whole lines of the axis");
    if (int rc = select_device(c)) return rc;
    adsb_view v = local_view(c);
    return sweep_impl(c, axis, slot, c->buf[b], v, nullptr, c->buf[b], v, nullptr, true);
}

int adsb_solve(adsb_ctx* c, int b, const int* slots) {
    if (!c) return fail(ADSB_EINVAL, "null context");
    for (int d = 0; d < c->ndim; ++d)
        if (int rc = adsb_sweep(c, b, d, slots ? slots[d] : 0)) return rc;
    return ADSB_OK;
}

int adsb_step(adsb_ctx* c, int u, int up, const adsb_substep* sub, int nsub, int nsteps) {
    if (!c || !sub || nsub < 1 || nsteps < 0) return fail(ADSB_EINVAL, "step: bad arguments");
    for (int it = 0; it < nsteps; ++it) {
        for (int s = 0; s < nsub; ++s) {
            if (int rc = adsb_swap(c, u, up)) return rc;
            if (int rc = adsb_compute_rhs(c, &sub[s].form, up, u)) return rc;
            if (sub[s].fix_axis >= 0) {
                const int ax = sub[s].fix_axis;
                if (ax >= c->ndim || sub[s].fix_buf < 0 || sub[s].fix_buf >= ADSB_MAX_BUFFERS || !c->buf[sub[s].fix_buf])
                    return fail(ADSB_ESTATE, "step: bad fix_axis / fix_buf");
                adsb_view v = local_view(c);
                // the plane values are the first elements (host order) of the managed tensor fix_buf
                cudaError_t e = (cudaError_t) launch_set_plane(c->buf[u], v.s, v.n, ax, 0, c->buf[sub[s].fix_buf], c->stream,
                                                               c->cnt[0], c->pitch0());
                if (e != cudaSuccess) return cuda_fail(e, "set_plane kernel");
                c->launches++;
            }
            if (int rc = adsb_solve(c, u, sub[s].slots)) return rc;
        }
    }
    return ADSB_OK;
}

int adsb_enable_timing(adsb_ctx* c, int on) {
    if (!c) return fail(ADSB_EINVAL, "null context");
    c->timing = on != 0;
    return ADSB_OK;
}

int adsb_stage_times(adsb_ctx* c, double* ms5) {
    if (!c || !ms5) return fail(ADSB_EINVAL, "stage_times: null argument");
    if (int rc = select_device(c)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    for (auto& s : c->spans) {
        float ms = 0;
        cudaEventElapsedTime(&ms, s.a, s.b);
        c->acc_ms[s.stage] += ms;
        c->free_events.push_back(s.a);
        c->free_events.push_back(s.b);
    }
    c->spans.clear();
    for (int i = 0; i < 5; ++i) {
        ms5[i] = c->acc_ms[i];
        c->acc_ms[i] = 0;
    }
    return ADSB_OK;
}

long long adsb_launch_count(adsb_ctx* c) { return c ? c->launches : 0; }

int adsb_load_tensor(adsb_ctx* c, int source, int with_test_function, int dst) {
    if (!c) return fail(ADSB_EINVAL, "null context");
    if (source != 1) return fail(ADSB_EINVAL, "load_tensor: unknown source");
    if (int rc = select_device(c)) return rc;
    if (int rc = ensure_buf(c, dst)) return rc;
    QuadAxes A;
    if (int rc = quad_axes(c, A)) return rc;
    StageTimer t(c, 4);
    if (with_test_function) {
        cudaError_t e = (cudaError_t) launch_project(3, A, c->buf[dst], c->lo, c->cnt, c->stream, c->pitch0());
        if (e != cudaSuccess) return cuda_fail(e, "project kernel");
        c->launches++;
        return ADSB_OK;
    }
    int elo[3] = {0, 0, 0}, en[3] = {1, 1, 1};
    size_t count = 1;
    for (int d = 0; d < c->ndim; ++d) {
        elo[d] = std::max(0, c->lo[d] - c->ax[d].p);
        const int ehi = std::min(c->ax[d].elements - 1, c->lo[d] + c->cnt[d] - 1);
        en[d] = ehi - elo[d] + 1;
        count *= (size_t) en[d];
    }
    double* G = nullptr;
    CU(cudaMalloc((void**) &G, count * sizeof(double)));
    cudaError_t e = (cudaError_t) launch_element_source(3, A, G, elo, en, c->stream);
    if (e == cudaSuccess)
        e = (cudaError_t) launch_box_sum(A, G, c->buf[dst], elo, en, c->lo, c->cnt, c->stream, c->pitch0());
    c->launches += 2;
    cudaError_t e2 = cudaStreamSynchronize(c->stream);
    cudaFree(G);
    if (e != cudaSuccess) return cuda_fail(e, "load tensor kernels");
    if (e2 != cudaSuccess) return cuda_fail(e2, "load tensor kernels");
    return ADSB_OK;
}

int adsb_project_init(adsb_ctx* c, int state, int dst) {
    if (!c) return fail(ADSB_EINVAL, "null context");
    if (state < 0 || state > 2) return fail(ADSB_EINVAL, "project_init: unknown state");
    if (int rc = select_device(c)) return rc;
    if (int rc = ensure_buf(c, dst)) return rc;
    QuadAxes A;
    if (int rc = quad_axes(c, A)) return rc;
    StageTimer t(c, 4);
    cudaError_t e = (cudaError_t) launch_project(state, A, c->buf[dst], c->lo, c->cnt, c->stream, c->pitch0());
    if (e != cudaSuccess) return cuda_fail(e, "project kernel");
    c->launches++;
    return ADSB_OK;
}

int adsb_sample(adsb_ctx* c, int b, const int* npts, const double* const* points, const double* const* knots,
                double* out) {
    if (!c || !npts || !points || !knots || !out) return fail(ADSB_EINVAL, "sample: null argument");
    if (b < 0 || b >= ADSB_MAX_BUFFERS || !c->buf[b]) return fail(ADSB_ESTATE, "sample: buffer not allocated");
    for (int d = 0; d < c->ndim; ++d) {
        if (c->cnt[d] != c->ng[d]) return fail(ADSB_ESTATE, "sample: the context must own the whole domain");
        if (!c->ax[d].tables) return fail(ADSB_ESTATE, "sample: axis tables not uploaded");
        if (npts[d] < 1 || !points[d] || !knots[d]) return fail(ADSB_EINVAL, "sample: bad points");
    }
    if (int rc = select_device(c)) return rc;
    int p[3] = {0, 0, 0};
    int* d_first[3] = {nullptr, nullptr, nullptr};
    double* d_val[3] = {nullptr, nullptr, nullptr};
    std::vector<void*> tmp;
    auto cleanup = [&] {
        for (void* q : tmp) cudaFree(q);
    };
    size_t total = 1;
    for (int d = 0; d < c->ndim; ++d) {
        const AxisData& ax = c->ax[d];
        p[d] = ax.p;
        const int ks = ax.elements + 2 * ax.p + 1;
        std::vector<int> first(npts[d]);
        std::vector<double> val((size_t) npts[d] * (ax.p + 1));
        for (int i = 0; i < npts[d]; ++i) {
            // bspline::eval: span of the point, the p+1 non-zero functions start at DOF span - p
            const int span = find_span(points[d][i], knots[d], ks, ax.p);
            first[i] = span - ax.p;
            basis_ders(span, points[d][i], knots[d], ax.p, 0, &val[(size_t) i * (ax.p + 1)]);
        }
        if (int rc = upload_vec(first, 0, &d_first[d], &tmp)) { cleanup(); return rc; }
        if (int rc = upload_vec(val, 0, &d_val[d], &tmp)) { cleanup(); return rc; }
        total *= (size_t) npts[d];
    }
    double* d_out = nullptr;
    if (cudaMalloc((void**) &d_out, total * sizeof(double)) != cudaSuccess) {
        cleanup();
        return fail(ADSB_ENOMEM, "sample: out of device memory");
    }
    tmp.push_back(d_out);
    cudaError_t e;
    {
        StageTimer tm(c, 4);
        e = (cudaError_t) launch_sample(c->ndim, npts, p, d_first, d_val, c->buf[b], c->pitch0(), c->pitch0() * c->cnt[1], d_out,
                                        c->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, total * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cleanup();
    if (e != cudaSuccess) return cuda_fail(e, "sample kernel");
    c->launches++;
    return ADSB_OK;
}

int adsb_norm(adsb_ctx* c, int b, int kind, int ref, double t, const double* ref_values, double* out2) {
    if (!c || !out2) return fail(ADSB_EINVAL, "norm: null argument");
    if (b < 0 || b >= ADSB_MAX_BUFFERS || !c->buf[b]) return fail(ADSB_ESTATE, "norm: buffer not allocated");
    if (kind < 0 || kind > 1 || ref < 0 || ref > 2) return fail(ADSB_EINVAL, "norm: kind is 0 (L2) or 1 (H1), ref 0..2");
    if (ref == 2 && (!ref_values || kind != 0)) return fail(ADSB_EINVAL, "norm: tabulated reference values give the L2 error only");
    for (int d = 0; d < c->ndim; ++d)
        if (c->cnt[d] != c->ng[d]) return fail(ADSB_ESTATE, "norm: the context must own the whole domain");
    if (int rc = select_device(c)) return rc;
    QuadAxes A;
    if (int rc = quad_axes(c, A)) return rc;
    const long long np = norm_partial_doubles(A);
    size_t ntab = 0;
    if (ref == 2) {
        ntab = 1;
        for (int d = 0; d < c->ndim; ++d) ntab *= (size_t) A.ne[d] * A.q[d];
    }
    double* scratch = nullptr;
    CU(cudaMalloc((void**) &scratch, ((size_t) np + 2 + ntab) * sizeof(double)));
    double* d_out = scratch + np;
    double* d_tab = ntab ? d_out + 2 : nullptr;
    cudaError_t e = cudaSuccess;
    if (ntab) e = cudaMemcpyAsync(d_tab, ref_values, ntab * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    {
        StageTimer tm(c, 4);
        if (e == cudaSuccess)
            e = (cudaError_t) launch_norm(A, c->buf[b], c->pitch0(), c->pitch0() * c->cnt[1], kind, ref, t, d_tab, scratch, d_out,
                                          c->stream);
    }
    double h[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(scratch);
    if (e != cudaSuccess) return cuda_fail(e, "norm kernels");
    c->launches += 2;
    out2[0] = std::sqrt(h[0]);
    out2[1] = std::sqrt(h[1]);
    return ADSB_OK;
}

int adsb_project_values(adsb_ctx* c, int dst, int ez_lo, int ez_cnt, const double* values, int accumulate) {
    if (!c || !values) return fail(ADSB_EINVAL, "project_values: null argument");
    for (int d = 0; d < c->ndim; ++d)
        if (c->cnt[d] != c->ng[d]) return fail(ADSB_ESTATE, "project_values: the context must own the whole domain");
    if (int rc = select_device(c)) return rc;
    if (int rc = ensure_buf(c, dst)) return rc;
    QuadAxes A;
    if (int rc = quad_axes(c, A)) return rc;
    const bool d3 = c->ndim == 3;
    if (!d3) {
        ez_lo = 0;
        ez_cnt = 1;
    } else if (ez_cnt <= 0) {  // the whole domain
        ez_lo = 0;
        ez_cnt = A.ne[2];
    }
    if (ez_lo < 0 || ez_cnt < 1 || (d3 && ez_lo + ez_cnt > A.ne[2])) return fail(ADSB_EINVAL, "project_values: bad element slab");
    const size_t count = (size_t) A.ne[0] * A.q[0] * A.ne[1] * A.q[1] * (d3 ? (size_t) ez_cnt * A.q[2] : 1);
    double* d_tab = nullptr;
    CU(cudaMalloc((void**) &d_tab, count * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d_tab, values, count * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    {
        StageTimer t(c, 4);
        if (e == cudaSuccess)
            e = (cudaError_t) launch_project_tab(A, d_tab, ez_lo, ez_cnt, accumulate, c->buf[dst], c->cnt, c->stream, c->pitch0());
    }
    const cudaError_t e2 = cudaStreamSynchronize(c->stream);
    cudaFree(d_tab);
    if (e != cudaSuccess) return cuda_fail(e, "project_values kernel");
    if (e2 != cudaSuccess) return cuda_fail(e2, "project_values kernel");
    c->launches++;
    return ADSB_OK;
}

int adsb_band_unpivot(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, double* ab_out) {
    if (!ab || !ipiv || !ab_out || n < 1 || kl < 0 || ku < 0 || ldab < 2 * kl + ku + 1)
        return fail(ADSB_EINVAL, "band_unpivot: bad argument");
    std::vector<double> out;
    if (!refactor_without_pivoting(n, kl, ku, ldab, ab, ipiv, out)) return 1;
    std::copy(out.begin(), out.end(), ab_out);
    return ADSB_OK;
}

int adsb_sweep_plan(int n, int kl, int ku, int ldab, const double* ab, const int* ipiv, int* dims, int* pv,
                    double* cfF, double* cfB, double* cfC, double* T, double* Rm, double* W, double* V) {
    SweepPlan P;
    if (int rc = build_sweep_plan(n, kl, ku, ldab, ab, ipiv, SWEEP_CH, 1, P)) return rc;
    if (dims) {
        const int d[16] = {P.KL, P.KD, P.piv, P.CH, P.R, P.SC, P.ST, P.rows, P.LF, P.LB, P.LC, P.DF, P.DB, P.seq,
                           SWEEP_MAX_DEPTH, 0};
        std::copy(d, d + 16, dims);
    }
    auto put = [](auto* dst, const auto& v) {
        if (dst) std::copy(v.begin(), v.end(), dst);
    };
    put(pv, P.pv);
    put(cfF, P.cfF);
    put(cfB, P.cfB);
    put(cfC, P.cfC);
    put(T, P.T);
    put(Rm, P.Rm);
    put(W, P.W);
    put(V, P.V);
    return ADSB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ segmented substitution
namespace {

int seg_of(adsb_ctx* c, int axis, int slot, SegSet** out) {
    if (!c || axis < 0 || axis >= c->ndim) return fail(ADSB_EINVAL, "segments: bad axis");
    if (slot < 0 || slot >= ADSB_MAX_SLOTS || !c->ax[axis].fac[slot].set)
        return fail(ADSB_ESTATE, "segments: no factor uploaded for this axis/slot");
    SegSet& g = c->ax[axis].fac[slot].seg;
    if (!g.set) return fail(ADSB_ESTATE, "segments: adsb_set_axis_segments was not called for this axis/slot");
    *out = &g;
    return ADSB_OK;
}

// lines of a view perpendicular to `axis`: l0 = the other axis with the smaller stride
int seg_geom(const adsb_ctx* c, int axis, const double* in, const adsb_view& vi, double* out, const adsb_view& vo,
             int row_base, int s_lo, int s_hi, SegGeom& G) {
    const int a = (axis + 1) % 3, b = (axis + 2) % 3;
    int l0 = (vi.s[a] <= vi.s[b]) ? a : b;
    if (vi.n[l0] == 1 && vi.n[l0 == a ? b : a] > 1 && vi.s[axis] != 1) l0 = (l0 == a) ? b : a;
    const int l1 = (l0 == a) ? b : a;
    for (int d = 0; d < 3; ++d)
        if (vi.n[d] != vo.n[d]) return fail(ADSB_EINVAL, "segments: in/out extents differ");
    G = SegGeom{};
    G.in = in;
    G.out = out;
    G.sj_in = vi.s[axis];
    G.sj_out = vo.s[axis];
    G.s0_in = vi.s[l0];
    G.s0_out = vo.s[l0];
    G.s1_in = vi.s[l1];
    G.s1_out = vo.s[l1];
    G.L0 = vi.n[l0];
    G.L1 = vi.n[l1];
    G.row_base = row_base;
    G.s_lo = s_lo;
    G.s_hi = s_hi;
    // contiguous lines (the z lines of an x-fastest tensor): number them flat, no idle lanes at row ends
    if (G.s0_in == 1 && G.s0_out == 1 && G.s1_in == G.L0 && G.s1_out == G.L0 && (long long) G.L0 * G.L1 < (1ll << 30)) {
        G.L0 *= G.L1;
        G.L1 = 1;
    }
    if (G.L1 > 65535) return fail(ADSB_EINVAL, "segments: outer extent beyond grid limits");
    (void) c;
    return ADSB_OK;
}

int seg_check_rows(const SegSet& g, int s_lo, int s_hi, int row_base, int rows) {
    if (s_lo < 0 || s_hi > g.dev.S || s_lo >= s_hi) return fail(ADSB_EINVAL, "segments: bad segment range");
    if (g.bounds[s_lo] < row_base || g.bounds[s_hi] > row_base + rows)
        return fail(ADSB_EINVAL, "segments: the view does not hold the rows of these segments");
    return ADSB_OK;
}

}  // namespace

namespace {

// All segments of every line on this GPU (lines longer than one CTA can hold): pass A of the segments side by
// side on a few streams (each a persistent kernel on its share of the SMs), then the boundary kernels and pass B.
int sweep_segmented(adsb_ctx* c, int axis, int slot, const double* in, const adsb_view& vi, double* out,
                    const adsb_view& vo, bool managed) {
    DevFactor& D = c->ax[axis].fac[slot];
    SegSet& g = D.seg;
    const int S = g.dev.S, KL = g.dev.KL, KD = g.dev.KD;
    SegGeom G;
    if (int rc = seg_geom(c, axis, out, vo, out, vo, 0, 0, S, G)) return rc;
    const size_t L = (size_t) G.L0 * G.L1;
    const size_t need = (size_t) S * (2 * KL + 2 * KD) * L;
    if (c->seg_scratch_doubles < need) {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(c->seg_scratch);
        c->seg_scratch = nullptr;
        c->seg_scratch_doubles = 0;
        CU(cudaMalloc((void**) &c->seg_scratch, need * sizeof(double)));
        c->seg_scratch_doubles = need;
    }
    double* dseg = c->seg_scratch;
    double* xst = dseg + (size_t) S * KL * L;
    double* din = xst + (size_t) S * KD * L;
    double* tin = din + (size_t) S * KL * L;
    const int ns = std::min(S, 8);
    while ((int) c->seg_streams.size() < ns) {
        cudaStream_t st;
        CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        c->seg_streams.push_back(st);
    }
    while ((int) c->seg_events.size() < ns + 1) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->seg_events.push_back(e);
    }
    StageTimer t(c, 1 + axis);
    cudaStream_t main_stream = c->stream;
    // pass A, preferably ONE launch: the CTAs with blockIdx.y = s sweep segment s (csrc/kernels_sweep_tile.cu)
    bool launched = false;
    {
        const int ax1 = (axis + 1) % 3, ax2 = (axis + 2) % 3;
        const bool contig = vi.s[axis] == 1 && vo.s[axis] == 1;
        int l0 = (vi.s[ax1] <= vi.s[ax2]) ? ax1 : ax2;
        if (!contig && vi.n[l0] == 1) l0 = (l0 == ax1) ? ax2 : ax1;
        const int l1 = (l0 == ax1) ? ax2 : ax1;
        std::vector<SweepGeom> Gs(S);
        for (int s = 0; s < S; ++s) {
            const int a0 = g.bounds[s];
            SweepGeom& Q = Gs[s];
            Q = SweepGeom{};
            Q.in = in + (long long) a0 * vi.s[axis];
            Q.out = out + (long long) a0 * vo.s[axis];
            Q.sj_in = vi.s[axis];
            Q.sj_out = vo.s[axis];
            Q.L0 = vi.n[l0];
            Q.L1 = vi.n[l1];
            Q.s0_in = vi.s[l0];
            Q.s0_out = vo.s[l0];
            Q.s1_in = vi.s[l1];
            Q.s1_out = vo.s[l1];
            Q.max_ctas = c->sm_limit;
            if (!contig) flatten_lines(Q);
            Q.pad_ok = managed && in == out && axis == 0 && vi.s[1] > vi.n[0] && s == S - 1;  // only the last segment ends at the pad
        }
        static const bool multi_on = [] {
            const char* e = getenv("ADSB_SEG_ONE_LAUNCH");
            return !e || atoi(e) != 0;
        }();
        const int rcm = multi_on ? launch_sweep_tile_multi(S, g.local.data(), Gs.data(), contig, main_stream) : -1;
        if (rcm > 0) return cuda_fail((cudaError_t) rcm, "segmented sweep (one launch)");
        launched = rcm == 0;
        if (launched) c->launches++;
    }
    const int keep_limit = c->sm_limit;
    const bool keep_timing = c->timing;
    int rc = ADSB_OK;
    if (!launched) {
        // fall-back: one launch per segment, side by side on a few streams
        const int sms = (keep_limit > 0 && keep_limit < device_sms()) ? keep_limit : device_sms();
        CU(cudaEventRecord(c->seg_events[0], main_stream));
        c->timing = false;
        c->sm_limit = std::max(1, sms / ns);
        for (int s = 0; s < S && rc == ADSB_OK; ++s) {
            cudaStream_t st = c->seg_streams[s % ns];
            if (s < ns) CU(cudaStreamWaitEvent(st, c->seg_events[0], 0));
            const int a = g.bounds[s], rows = g.bounds[s + 1] - a;
            adsb_view svi = vi, svo = vo;
            svi.n[axis] = svo.n[axis] = rows;
            c->stream = st;
            rc = sweep_impl(c, axis, slot, in + (long long) a * vi.s[axis], svi, nullptr, out + (long long) a * vo.s[axis], svo,
                            nullptr, managed, &g.local[s]);
        }
        c->stream = main_stream;
        c->sm_limit = keep_limit;
        c->timing = keep_timing;
        if (rc != ADSB_OK) return rc;
        for (int i = 0; i < ns; ++i) {
            CU(cudaEventRecord(c->seg_events[1 + i], c->seg_streams[i]));
            CU(cudaStreamWaitEvent(main_stream, c->seg_events[1 + i], 0));
        }
    }
    double* dst1[1] = {dseg};
    double* dst2[1] = {xst};
    cudaError_t e = (cudaError_t) launch_seg_dseg(g.dev, G, dst1, 1, main_stream);
    if (e == cudaSuccess) e = (cudaError_t) launch_seg_din(g.dev, G, dseg, din, dst2, 1, main_stream);
    const double* back = xst;
    if (e == cudaSuccess && g.dev.DB > 1) {
        e = (cudaError_t) launch_seg_tin(g.dev, 0, S, (long long) L, xst, tin, main_stream);
        back = tin;
        c->launches++;
    }
    int max_rows = 0;
    for (int s = 0; s < S; ++s) max_rows = std::max(max_rows, g.bounds[s + 1] - g.bounds[s]);
    if (e == cudaSuccess) e = (cudaError_t) launch_seg_correct(g.dev, G, din, back, g.dev.DB == 1 ? 1 : 0, max_rows, main_stream);
    if (e != cudaSuccess) return cuda_fail(e, "segmented sweep kernels");
    c->launches += 3;
    return ADSB_OK;
}

}  // namespace

extern "C" {

int adsb_set_axis_segments(adsb_ctx* c, int axis, int slot, int nseg, const int* bounds, int local_lo, int local_cnt) {
    if (!c || axis < 0 || axis >= c->ndim) return fail(ADSB_EINVAL, "set_axis_segments: bad axis");
    if (slot < 0 || slot >= ADSB_MAX_SLOTS || !c->ax[axis].fac[slot].set)
        return fail(ADSB_ESTATE, "set_axis_segments: upload the factor first (adsb_set_axis_factor)");
    if (!bounds || nseg < 1 || local_lo < 0 || local_cnt < 0 || local_lo + local_cnt > nseg)
        return fail(ADSB_EINVAL, "set_axis_segments: bad segment arguments");
    if (int rc = select_device(c)) return rc;
    c->ax[axis].fac[slot].auto_seg = false;
    return set_segments_impl(c, c->ax[axis].fac[slot], nseg, bounds, local_lo, local_cnt);
}

int adsb_segment_info(adsb_ctx* c, int axis, int slot, int* info8) {
    SegSet* g;
    if (int rc = seg_of(c, axis, slot, &g)) return rc;
    if (!info8) return fail(ADSB_EINVAL, "segment_info: null argument");
    const int v[8] = {g->dev.KL, g->dev.KD, g->dev.DF, g->dev.DB, g->dev.S, g->local_lo, g->local_cnt, g->dev.n};
    std::copy(v, v + 8, info8);
    return ADSB_OK;
}

int adsb_seg_sweep_view(adsb_ctx* c, int axis, int slot, int seg, const double* in, const adsb_view* vin, double* out,
                        const adsb_view* vout) {
    SegSet* g;
    if (int rc = seg_of(c, axis, slot, &g)) return rc;
    if (!in || !out || !vin || !vout) return fail(ADSB_EINVAL, "seg_sweep_view: null argument");
    if (seg < g->local_lo || seg >= g->local_lo + g->local_cnt)
        return fail(ADSB_EINVAL, "seg_sweep_view: segment is not local to this context");
    if (int rc = select_device(c)) return rc;
    return sweep_impl(c, axis, slot, in, *vin, nullptr, out, *vout, nullptr, false, &g->local[seg - g->local_lo]);
}

static bool dist_short_chunks() {  // ADSB_DIST_SHORT_CHUNKS=0: keep 18-column chunks in the fused sweep
    static const bool on = [] {
        const char* e = getenv("ADSB_DIST_SHORT_CHUNKS");
        return !e || atoi(e) != 0;
    }();
    return on;
}

int adsb_dist_sweep_check(adsb_ctx* c, int axis, int slot, int rank, const adsb_view* v, int nl, int lag) {
    SegSet* g;
    if (int rc = seg_of(c, axis, slot, &g)) return rc;
    if (!v || rank < g->local_lo || rank >= g->local_lo + g->local_cnt) return fail(ADSB_EINVAL, "dist_sweep_check: bad argument");
    const SweepFactor& F = g->local[rank - g->local_lo];
    if (v->n[axis] != F.n || v->s[axis] == 1) return 0;
    const int a = (axis + 1) % 3, b = (axis + 2) % 3;
    const int l0 = (v->s[a] <= v->s[b]) ? a : b, l1 = (l0 == a) ? b : a;
    SweepGeom G{};
    G.sj_in = G.sj_out = v->s[axis];
    G.L0 = v->n[l0];
    G.L1 = v->n[l1];
    G.s0_in = G.s0_out = v->s[l0];
    G.s1_in = G.s1_out = v->s[l1];
    if (G.s0_in != 1 || (G.sj_in & 1) || (G.s1_in & 1)) return 0;
    flatten_lines(G);
    SweepDistArgs D{};
    D.lag = lag > 0 ? lag : 4;
    if (dist_short_chunks() && !g->local_dist.empty() &&
        launch_sweep_dist(g->local_dist[0], SWEEP_CH_DIST, g->dev, G, D, nl, nullptr, true) == 0)
        return 1;
    return launch_sweep_dist(F, SWEEP_CH, g->dev, G, D, nl, nullptr, true) == 0 ? 1 : 0;
}

int adsb_dist_sweep_view(adsb_ctx* c, int axis, int slot, double* data, const adsb_view* v, const adsb_dist_args* d) {
    SegSet* g;
    if (int rc = seg_of(c, axis, slot, &g)) return rc;
    if (!data || !v || !d) return fail(ADSB_EINVAL, "dist_sweep_view: null argument");
    static_assert(ADSB_DIST_SENTINEL_WORD == (unsigned) (ADSB_DIST_SENTINEL_BITS >> 32) &&
                      ADSB_DIST_SENTINEL_WORD == (unsigned) (ADSB_DIST_SENTINEL_BITS & 0xffffffffu),
                  "sentinel constants out of sync");
    if (d->rank < g->local_lo || d->rank >= g->local_lo + g->local_cnt || d->nranks != g->dev.S)
        return fail(ADSB_EINVAL, "dist_sweep_view: rank is not a local segment of this context");
    if (!d->dseg_local || !d->x_local || (d->rank > 0 && !d->x_prev) || (d->rank + 1 < d->nranks && !d->dseg_next))
        return fail(ADSB_EINVAL, "dist_sweep_view: missing state arrays");
    if (int rc = select_device(c)) return rc;
    const SweepFactor& F = g->local[d->rank - g->local_lo];
    if (v->n[axis] != F.n) return fail(ADSB_EINVAL, "dist_sweep_view: the view does not span the slab");
    if (v->s[axis] == 1) return fail(ADSB_EINVAL, "dist_sweep_view: the sharded axis must not be the contiguous one");
    const int a = (axis + 1) % 3, b = (axis + 2) % 3;
    const int l0 = (v->s[a] <= v->s[b]) ? a : b, l1 = (l0 == a) ? b : a;
    SweepGeom G{};
    G.in = data;
    G.out = data;
    G.sj_in = G.sj_out = v->s[axis];
    G.L0 = v->n[l0];
    G.L1 = v->n[l1];
    G.s0_in = G.s0_out = v->s[l0];
    G.s1_in = G.s1_out = v->s[l1];
    G.max_ctas = c->sm_limit;
    flatten_lines(G);
    SweepDistArgs D{};
    D.rank = d->rank;
    D.row_base = g->bounds[d->rank];
    D.lag = d->lag > 0 ? d->lag : 4;
    D.dseg_local = d->dseg_local;
    D.x_local = d->x_local;
    D.dseg_next = d->rank + 1 < d->nranks ? d->dseg_next : nullptr;
    D.x_prev = d->rank > 0 ? d->x_prev : nullptr;
    D.error_flag = d->error_flag;
    D.halo_prev = d->rank > 0 ? d->halo_prev : nullptr;
    D.halo_next = d->rank + 1 < d->nranks ? d->halo_next : nullptr;
    D.halo_planes = d->halo_planes;
    if ((D.halo_prev && (uintptr_t) D.halo_prev % 16) || (D.halo_next && (uintptr_t) D.halo_next % 16))
        return fail(ADSB_EINVAL, "dist_sweep_view: halo pointers must be 16 B aligned");
    StageTimer t(c, 1 + axis);
    int rc = -1;
    if (dist_short_chunks() && !g->local_dist.empty())
        rc = launch_sweep_dist(g->local_dist[0], SWEEP_CH_DIST, g->dev, G, D, d->nl, c->stream);
    if (rc == -1) rc = launch_sweep_dist(F, SWEEP_CH, g->dev, G, D, d->nl, c->stream);
    if (rc == -1) return fail(ADSB_ESTATE, "dist_sweep_view: this factor / view is not eligible for the fused kernel");
    if (rc != 0) return cuda_fail((cudaError_t) rc, "distributed sweep kernel launch");
    c->launches++;
    return ADSB_OK;
}

int adsb_neighbor_barrier(adsb_ctx* c, unsigned long long* flags_local, unsigned long long* flags_prev,
                          unsigned long long* flags_next, int* error_flag) {
    if (!c || !flags_local) return fail(ADSB_EINVAL, "neighbor_barrier: null argument");
    if (int rc = select_device(c)) return rc;
    cudaError_t e = (cudaError_t) launch_neighbor_barrier(flags_local, flags_prev, flags_next, error_flag, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "neighbour barrier kernel");
    c->launches++;
    return ADSB_OK;
}

int adsb_seg_dseg_view(adsb_ctx* c, int axis, int slot, int s_lo, int s_hi, int row_base, const double* xhat,
                       const adsb_view* vin, double* const* dst, int ndst) {
    SegSet* g;
    if (int rc = seg_of(c, axis, slot, &g)) return rc;
    if (!xhat || !vin || !dst) return fail(ADSB_EINVAL, "seg_dseg_view: null argument");
    if (int rc = seg_check_rows(*g, s_lo, s_hi, row_base, vin->n[axis])) return rc;
    if (int rc = select_device(c)) return rc;
    SegGeom G;
    if (int rc = seg_geom(c, axis, xhat, *vin, nullptr, *vin, row_base, s_lo, s_hi, G)) return rc;
    StageTimer t(c, 1 + axis);
    cudaError_t e = (cudaError_t) launch_seg_dseg(g->dev, G, dst, ndst, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "seg_dseg kernel");
    c->launches++;
    return ADSB_OK;
}

int adsb_seg_din_view(adsb_ctx* c, int axis, int slot, int s_lo, int s_hi, int row_base, const double* xhat,
                      const adsb_view* vin, const double* dseg, double* din, double* const* xdst, int ndst) {
    SegSet* g;
    if (int rc = seg_of(c, axis, slot, &g)) return rc;
    if (!xhat || !vin || !dseg || !din || !xdst) return fail(ADSB_EINVAL, "seg_din_view: null argument");
    if (int rc = seg_check_rows(*g, s_lo, s_hi, row_base, vin->n[axis])) return rc;
    if (int rc = select_device(c)) return rc;
    SegGeom G;
    if (int rc = seg_geom(c, axis, xhat, *vin, nullptr, *vin, row_base, s_lo, s_hi, G)) return rc;
    StageTimer t(c, 1 + axis);
    cudaError_t e = (cudaError_t) launch_seg_din(g->dev, G, dseg, din, xdst, ndst, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "seg_din kernel");
    c->launches++;
    return ADSB_OK;
}

int adsb_seg_tin(adsb_ctx* c, int axis, int slot, int s_lo, int s_hi, long long lines, const double* X, double* tin) {
    SegSet* g;
    if (int rc = seg_of(c, axis, slot, &g)) return rc;
    if (!X || !tin || lines < 1 || s_lo < 0 || s_hi > g->dev.S || s_lo >= s_hi)
        return fail(ADSB_EINVAL, "seg_tin: bad argument");
    if (int rc = select_device(c)) return rc;
    StageTimer t(c, 1 + axis);
    cudaError_t e = (cudaError_t) launch_seg_tin(g->dev, s_lo, s_hi, lines, X, tin, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "seg_tin kernel");
    c->launches++;
    return ADSB_OK;
}

int adsb_seg_correct_view(adsb_ctx* c, int axis, int slot, int s_lo, int s_hi, int row_base, const double* in,
                          const adsb_view* vin, double* out, const adsb_view* vout, const double* din,
                          const double* tin_or_x) {
    SegSet* g;
    if (int rc = seg_of(c, axis, slot, &g)) return rc;
    if (!in || !out || !vin || !vout || !din || !tin_or_x) return fail(ADSB_EINVAL, "seg_correct_view: null argument");
    if (int rc = seg_check_rows(*g, s_lo, s_hi, row_base, vin->n[axis])) return rc;
    if (int rc = select_device(c)) return rc;
    SegGeom G;
    if (int rc = seg_geom(c, axis, in, *vin, out, *vout, row_base, s_lo, s_hi, G)) return rc;
    int max_rows = 0;
    for (int s = s_lo; s < s_hi; ++s) max_rows = std::max(max_rows, g->bounds[s + 1] - g->bounds[s]);
    StageTimer t(c, 1 + axis);
    cudaError_t e = (cudaError_t) launch_seg_correct(g->dev, G, din, tin_or_x, g->dev.DB == 1 ? 1 : 0, max_rows, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "seg_correct kernel");
    c->launches++;
    return ADSB_OK;
}

}  // extern "C"
