// kernels_sweep_tile.cu -- K2, TMA-fed variant: persistent CTAs stream tiles of lines through a
// shared-memory ring so that loads, the substitution and stores of neighbouring tiles overlap.
//
// Same algorithm and arithmetic as kernels_sweep.cu (sweep_core.cuh); only the data movement
// differs.  One CTA per SM loops over tiles of NL = 16 lines:
//   STRIDED sweeps (y, z; lanes along x): a tile is n rows x 16 x-values.  The TMA engine moves it
//     with 3-D tensor-map copies (cp.async.bulk.tensor, boxes of 16 x <=256 rows; SASS UTMALDG /
//     UTMASTG); rows past the end of the line are zero-filled on load and clipped on store.
//   CONTIG sweep (x): a tile is 16 whole lines, one bulk copy (cp.async.bulk, UBLKCP) per line.
// Ring of NBUF tiles: while the threads solve tile i out of shared memory (in place), tiles i+1 and
// i+2 are landing and tile i-1 is draining.  A thread reads / writes its chunk with 128-bit shared
// accesses at compile-time offsets -- no per-element address arithmetic, no LSU queue limits on
// the bytes in flight.
#include <cuda.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "kernels.cuh"
#include "sweep_core.cuh"

namespace adsb {

namespace {

constexpr int MAX_NBUF = 3;

// Block = ncons consumer threads (NLt lanes x chunks, padded to whole warps) + one producer warp.
template <int KL, int KD, bool PIV, int CH, int NL, bool CONTIG>
__device__ __forceinline__ void sweep_tile_body(const SweepFactor& F0, const SweepTileGeom& G) {
    constexpr int RL = SWEEP_RL, NLt = NL / RL;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int SC = F0.SC;
    const int ncons = (int) blockDim.x - 32;
    const int tid = threadIdx.x;
    const int n = F0.n;
    const int tile_doubles = G.tile_doubles;
    double* tiles = reinterpret_cast<double*>(smem_raw);
    double* fst = tiles + (size_t) G.nbuf * tile_doubles;  // [SC][KL][NL]
    double* bst = fst + SC * KL * NL;                      // [SC][KD][NL]
    double* s_tab = bst + SC * KD * NL;                    // the factor's tables, same layout as F0's blob
    double* s_cfF = s_tab;
    double* s_cfB = s_tab + F0.off[0];
    double* s_cfC = s_tab + F0.off[1];
    double* s_T = s_tab + F0.off[2];
    double* s_Rm = s_tab + F0.off[3];
    double* s_W = s_tab + F0.off[4];
    double* s_V = s_tab + F0.off[5];
    uint64_t* full = reinterpret_cast<uint64_t*>(s_tab + F0.blob_doubles);
    uint64_t* done = full + MAX_NBUF;
    uint64_t* tabbar = done + MAX_NBUF;

    pdl_launch();  // every CTA of this grid is resident (one per SM): the next kernel's CTAs may queue up behind them
    if (tid == 0) {
        for (int b = 0; b < G.nbuf; ++b) {
            mbar_init(&full[b], 1);
            mbar_init(&done[b], 1);
        }
        mbar_init(tabbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int my_count = (G.ntiles - (int) blockIdx.x + (int) gridDim.x - 1) / (int) gridDim.x;

    if (tid < ncons) {
        // consumers set the CTA up while the producer's first loads are already in flight:
        // STRIDED: rows [n, rows_needed) of every ring slot are never written by the TMA and must read as
        // zero (chunks read past the line end); CONTIG reads are guarded, nothing to clear
        if (!CONTIG) {
            const int pad0 = n * NL, pad = tile_doubles - pad0;
            for (int i = tid; i < G.nbuf * pad; i += ncons) tiles[(size_t) (i / pad) * tile_doubles + pad0 + i % pad] = 0.0;
        }
        sweep_sync(ncons);
        mbar_wait(tabbar, 0);  // the factor's tables (one bulk copy issued by the producer) have landed
    }

    if (tid >= ncons) {
        // ------------------------------------------------------------------ producer warp
        if (tid != ncons) return;
        // the factor's tables first: every tile of this persistent CTA uses them (one bulk copy, UBLKCP)
        mbar_expect_tx(tabbar, (uint32_t) (F0.blob_doubles * 8));
        bulk_g2s(s_tab, F0.cfF, (uint32_t) (F0.blob_doubles * 8), tabbar);
        pdl_wait();  // the tensor itself: only once the previous kernel of the stream has finished with it
        auto issue_load = [&](int i) {
            const int t = G.reverse ? G.ntiles - 1 - ((int) blockIdx.x + i * (int) gridDim.x) : (int) blockIdx.x + i * (int) gridDim.x;
            const int bx = t % G.nb0, m = t / G.nb0;
            const int b = i % G.nbuf;
            double* dst = tiles + (size_t) b * tile_doubles;
            if (CONTIG) {
                const int lines = min(NL, G.L0 - bx * NL);
                mbar_expect_tx(&full[b], (uint32_t) (lines * G.ncopy * 8));
                const double* src = G.in + (long long) (bx * NL) * G.s0_in + (long long) m * G.s1_in;
                for (int ln = 0; ln < lines; ++ln)
                    bulk_g2s(dst + ln * G.pitch, src + ln * G.s0_in, (uint32_t) (G.ncopy * 8), &full[b]);
            } else {
                const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(G.maps);
                mbar_expect_tx(&full[b], (uint32_t) G.load_bytes);
                for (int k = 0; k < G.nbox_in; ++k) tma_load_3d(dst + G.row0_in[k] * NL, maps + k, bx * NL, 0, m, &full[b]);
            }
        };
        auto issue_store = [&](int i) {
            const int t = G.reverse ? G.ntiles - 1 - ((int) blockIdx.x + i * (int) gridDim.x) : (int) blockIdx.x + i * (int) gridDim.x;
            const int bx = t % G.nb0, m = t / G.nb0;
            const int b = i % G.nbuf;
            const double* src = tiles + (size_t) b * tile_doubles;
            if (CONTIG) {
                const int lines = min(NL, G.L0 - bx * NL);
                double* dst = G.out + (long long) (bx * NL) * G.s0_out + (long long) m * G.s1_out;
                for (int ln = 0; ln < lines; ++ln)
                    bulk_s2g(dst + ln * G.s0_out, src + ln * G.pitch, (uint32_t) (G.ncopy * 8));
            } else {
                const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(G.maps) + G.nbox_in;
                for (int k = 0; k < G.nbox_out; ++k) tma_store_3d(maps + k, bx * NL, 0, m, src + G.row0_out[k] * NL);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        };
        for (int i = 0; i < G.nbuf - 1 && i < my_count; ++i) issue_load(i);
        for (int j = 0; j < my_count; ++j) {
            if (j + G.nbuf - 1 < my_count) {
                // ring slot of tile j-1: reusable once its store has read shared memory
                if (j >= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                issue_load(j + G.nbuf - 1);
            }
            mbar_wait(&done[j % G.nbuf], (uint32_t) ((j / G.nbuf) & 1));
            issue_store(j);
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    SweepFactor F = F0;
    F.cfF = s_cfF;
    F.cfB = s_cfB;
    F.cfC = s_cfC;
    F.T = s_T;
    F.Rm = s_Rm;
    F.W = s_W;
    F.V = s_V;
    const int tx = tid % NLt;
    const int c = min(tid / NLt, SC - 1);  // padding threads shadow the last chunk (identical values)
    const int j0 = c * CH;

    for (int i = 0; i < my_count; ++i) {
        const int b = i % G.nbuf;
        double* tile = tiles + (size_t) b * tile_doubles;
        mbar_wait(&full[b], (uint32_t) ((i / G.nbuf) & 1));

        double v[RL][CH + KL];
        bool act[RL];
        if (CONTIG) {
            const int t = G.reverse ? G.ntiles - 1 - ((int) blockIdx.x + i * (int) gridDim.x) : (int) blockIdx.x + i * (int) gridDim.x;
            const int lbase = (t % G.nb0) * NL;
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                act[r] = lbase + r * NLt + tx < G.L0;
                const double* mine = tile + (r * NLt + tx) * G.pitch + j0;
#pragma unroll
                for (int q = 0; q < CH + KL; q += 2) {  // CH even on this path; pitch and j0 even
                    double2 t2 = make_double2(0.0, 0.0);
                    if (act[r] && j0 + q < n) t2 = *reinterpret_cast<const double2*>(mine + q);
                    v[r][q] = t2.x;
                    if (q + 1 < CH + KL) v[r][q + 1] = (j0 + q + 1 < n) ? t2.y : 0.0;  // CH + KL is odd for odd KL
                }
            }
        } else {
            const double* mine = tile + (size_t) j0 * NL + 2 * tx;
#pragma unroll
            for (int q = 0; q < CH + KL; ++q) {
                const double2 t2 = *reinterpret_cast<const double2*>(mine + q * NL);
                v[0][q] = t2.x;
                v[1][q] = t2.y;
            }
        }

        sweep_core<KL, KD, PIV, CH, RL, true>(F, v, fst, bst, c, tx, NLt, SC, ncons);

        if (CONTIG) {
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                double* mine = tile + (r * NLt + tx) * G.pitch + j0;
#pragma unroll
                for (int q = 0; q < CH; q += 2) {
                    if (act[r] && j0 + q + 1 < n)
                        *reinterpret_cast<double2*>(mine + q) = make_double2(v[r][q], v[r][q + 1]);
                    else if (act[r] && j0 + q < n)
                        mine[q] = v[r][q];
                }
            }
        } else {
            double* mine = tile + (size_t) j0 * NL + 2 * tx;
#pragma unroll
            for (int q = 0; q < CH; ++q)
                if (j0 + q < n) *reinterpret_cast<double2*>(mine + q * NL) = make_double2(v[0][q], v[1][q]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        sweep_sync(ncons);
        if (tid == 0) mbar_arrive(&done[b]);
    }
}

template <int KL, int KD, bool PIV, int CH, int NL, bool CONTIG>
__global__ void __launch_bounds__(288, 1) sweep_tile_kernel(const SweepFactor F0, const SweepTileGeom G) {
    sweep_tile_body<KL, KD, PIV, CH, NL, CONTIG>(F0, G);
}

// Several segments of the same lines in ONE launch (lines too long for one CTA, api.cu: sweep_segmented): the
// CTAs with blockIdx.y = s work on segment s -- its own factor and its own tile geometry, read once from
// global memory into shared memory.
struct SweepSegArgs {
    SweepFactor F;
    SweepTileGeom G;
};
template <int KL, int KD, bool PIV, int CH, int NL, bool CONTIG>
__global__ void __launch_bounds__(288, 1) sweep_tile_multi_kernel(const SweepSegArgs* __restrict__ segs) {
    __shared__ SweepSegArgs a;
    static_assert(sizeof(SweepSegArgs) % 4 == 0, "copied word by word");
    const int* src = reinterpret_cast<const int*>(segs + blockIdx.y);
    int* dst = reinterpret_cast<int*>(&a);
    for (int i = threadIdx.x; i < (int) (sizeof(SweepSegArgs) / 4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    sweep_tile_body<KL, KD, PIV, CH, NL, CONTIG>(a.F, a.G);
}

using tile_kern_t = void (*)(const SweepFactor, const SweepTileGeom);
using tile_multi_kern_t = void (*)(const SweepSegArgs*);

template <int P, bool PIV>
tile_multi_kern_t pick_multi_mode(bool contig) {  // long lines: always 16 lines per tile
    constexpr int KD = PIV ? 2 * P : P;
    return contig ? (tile_multi_kern_t) sweep_tile_multi_kernel<P, KD, PIV, SWEEP_CH, 16, true>
                  : (tile_multi_kern_t) sweep_tile_multi_kernel<P, KD, PIV, SWEEP_CH, 16, false>;
}
tile_multi_kern_t pick_multi(int KL, bool piv, bool contig) {
    switch (KL) {
    case 1: return piv ? pick_multi_mode<1, true>(contig) : pick_multi_mode<1, false>(contig);
    case 2: return piv ? pick_multi_mode<2, true>(contig) : pick_multi_mode<2, false>(contig);
    case 3: return piv ? pick_multi_mode<3, true>(contig) : pick_multi_mode<3, false>(contig);
    case 4: return piv ? pick_multi_mode<4, true>(contig) : pick_multi_mode<4, false>(contig);
    case 5: return piv ? pick_multi_mode<5, true>(contig) : pick_multi_mode<5, false>(contig);
    default: return nullptr;
    }
}

template <int P, bool PIV, int NL>
tile_kern_t pick_mode(bool contig) {
    constexpr int KD = PIV ? 2 * P : P;
    return contig ? (tile_kern_t) sweep_tile_kernel<P, KD, PIV, SWEEP_CH, NL, true>
                  : (tile_kern_t) sweep_tile_kernel<P, KD, PIV, SWEEP_CH, NL, false>;
}

template <int NL>
tile_kern_t pick(int KL, bool piv, bool contig) {
    switch (KL) {
    case 1: return piv ? pick_mode<1, true, NL>(contig) : pick_mode<1, false, NL>(contig);
    case 2: return piv ? pick_mode<2, true, NL>(contig) : pick_mode<2, false, NL>(contig);
    case 3: return piv ? pick_mode<3, true, NL>(contig) : pick_mode<3, false, NL>(contig);
    case 4: return piv ? pick_mode<4, true, NL>(contig) : pick_mode<4, false, NL>(contig);
    case 5: return piv ? pick_mode<5, true, NL>(contig) : pick_mode<5, false, NL>(contig);
    default: return nullptr;
    }
}

typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_fn_t encode_fn() {
    static encode_fn_t fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (encode_fn_t) p;
    }();
    return fn;
}

// 3-D map over (x, rows of one box, other axis) with element strides (1, sj, s1); box NL x rows x 1
bool make_map(CUtensorMap* m, const double* base, int L0, int rows, int L1, long long sj, long long s1, int NL) {
    encode_fn_t enc = encode_fn();
    if (!enc) return false;
    if (sj <= 0) {
        if (rows > 1) return false;  // rows running backwards or on top of each other: not a tensor map; the
                                     // caller falls back to the register-path kernel
        sj = 2;                      // single row: any valid stride
    }
    const cuuint64_t dims[3] = {(cuuint64_t) L0, (cuuint64_t) rows, (cuuint64_t) (L1 > 0 ? L1 : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t) sj * 8, (cuuint64_t) (L1 > 1 ? s1 : sj * rows) * 8};
    const cuuint32_t box[3] = {(cuuint32_t) NL, (cuuint32_t) rows, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*) base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
           CUDA_SUCCESS;
}

struct Box {
    int row0, rows;
    long long off, stride;  // element offset of its first row, row stride
};

// Cut rows 0..n-1 (offset off[j], or j*sj) into boxes of <= 256 rows with constant stride.
bool cut_boxes(int n, const long long* off, long long sj, std::vector<Box>& out) {
    out.clear();
    int j = 0;
    while (j < n) {
        int e = j + 1;
        long long stride = sj;
        if (off) {
            stride = e < n ? off[e] - off[j] : 0;
            while (e < n && off[e] - off[e - 1] == stride) ++e;
        } else {
            e = n;
        }
        // [j, e) is linear; split evenly into pieces of at most 256 rows
        const int len = e - j, pieces = (len + 255) / 256, per = (len + pieces - 1) / pieces;
        for (int r = j; r < e; r += per) {
            Box b{r, std::min(per, e - r), off ? off[r] : (long long) r * sj, stride};
            out.push_back(b);
        }
        j = e;
    }
    return (int) out.size() <= SWEEP_MAX_BOXES;
}

std::mutex g_map_mutex;
std::map<std::vector<long long>, void*> g_map_cache;  // encoded maps already resident on the device

}  // namespace

bool encode_tensor_map3(void* map, const double* base, const unsigned long long dims[3],
                        const unsigned long long strides[2], const unsigned box[3]) {
    encode_fn_t enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t d[3] = {dims[0], dims[1], dims[2]};
    const cuuint64_t s[2] = {strides[0] * 8, strides[1] * 8};
    const cuuint32_t b[3] = {box[0], box[1], box[2]};
    const cuuint32_t es[3] = {1, 1, 1};
    return enc(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*) base, d, s, b, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tensor maps of the STRIDED tiles (NL x-values by n rows, one 3-D map per linear piece of the row-offset
// tables, boxes of <= 256 rows), cached on the device; fills T.maps, T.nbox_*, T.row0_*, T.load_bytes.
// 0: done; -1: this geometry cannot be moved by the TMA engine; otherwise a cudaError_t.
int sweep_strided_maps(const SweepGeom& G, int n, int NL, const long long* off_in_h, const long long* off_out_h,
                       cudaStream_t st, SweepTileGeom& T) {
    auto even = [](long long v) { return (v & 1) == 0; };
    const bool ptr_ok = ((uintptr_t) G.in % 16 == 0) && ((uintptr_t) G.out % 16 == 0);
    if (!ptr_ok || G.s0_in != 1 || G.s0_out != 1 || (G.L1 > 1 && (!even(G.s1_in) || !even(G.s1_out)))) return -1;
    std::vector<Box> bin, bout;
    if (!cut_boxes(n, off_in_h, G.sj_in, bin) || !cut_boxes(n, off_out_h, G.sj_out, bout)) return -1;
    // a box of several rows needs a positive, even row stride (a descending or repeating offset table cannot be
    // one tensor map: the register-path kernel takes such sweeps)
    for (const auto& b : bin)
        if (!even(b.off) || (b.rows > 1 && (!even(b.stride) || b.stride <= 0))) return -1;
    for (const auto& b : bout)
        if (!even(b.off) || (b.rows > 1 && (!even(b.stride) || b.stride <= 0))) return -1;
    T.nbox_in = (int) bin.size();
    T.nbox_out = (int) bout.size();
    T.load_bytes = 0;
    std::vector<long long> key = {(long long) (uintptr_t) G.in, (long long) (uintptr_t) G.out, G.L0, G.L1, n, NL,
                                  G.s1_in, G.s1_out};
    for (size_t k = 0; k < bin.size(); ++k) {
        T.row0_in[k] = bin[k].row0;
        T.load_bytes += bin[k].rows * NL * 8;
        key.insert(key.end(), {bin[k].row0, bin[k].rows, bin[k].off, bin[k].stride});
    }
    for (size_t k = 0; k < bout.size(); ++k) {
        T.row0_out[k] = bout[k].row0;
        key.insert(key.end(), {-1 - bout[k].row0, bout[k].rows, bout[k].off, bout[k].stride});
    }
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it == g_map_cache.end()) {
        std::vector<CUtensorMap> maps(bin.size() + bout.size());
        for (size_t k = 0; k < bin.size(); ++k)
            if (!make_map(&maps[k], G.in + bin[k].off, G.L0, bin[k].rows, G.L1, bin[k].stride, G.s1_in, NL)) return -1;
        for (size_t k = 0; k < bout.size(); ++k)
            if (!make_map(&maps[bin.size() + k], G.out + bout[k].off, G.L0, bout[k].rows, G.L1, bout[k].stride, G.s1_out, NL))
                return -1;
        void* d = nullptr;
        if (cudaMalloc(&d, maps.size() * sizeof(CUtensorMap)) != cudaSuccess) return -1;
        cudaError_t e = cudaMemcpyAsync(d, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // `maps` is a local
        if (e != cudaSuccess) return (int) e;
        it = g_map_cache.emplace(std::move(key), d).first;
    }
    T.maps = it->second;
    return 0;
}

namespace {

struct TilePrep {
    SweepTileGeom T{};
    size_t smem = 0;
    int NL = 0, ncons = 0;
};

// Geometry, tensor maps and shared-memory budget of one tile sweep.  0: ready; -1: not eligible; else cudaError_t.
int prepare_sweep_tile(const SweepFactor& F, const SweepGeom& G, bool contig, const long long* off_in_h,
                       const long long* off_out_h, cudaStream_t st, int force_NL, TilePrep& P) {
    // lines per tile: as many lanes as keep the CTA near 256 consumer threads (lanes x chunks): 16 lines for
    // 17..32 chunks (n <= 576), 32 for 9..16, 64 for 5..8, 128 for <= 4 chunks (the thin slabs of a sharded z
    // sweep); wider tiles also move longer rows (128 B .. 1 KB).  ADSB_SWEEP_NL=12|16|32|64|128 pins it.
    static const int NL_env = [] {
        const char* e = getenv("ADSB_SWEEP_NL");
        const int v = e ? atoi(e) : 0;
        return (v == 12 || v == 16 || v == 32 || v == 64 || v == 128) ? v : 0;
    }();
    int NL = force_NL ? force_NL : NL_env;
    if (!NL) {
        // shared memory of a CTA with nl lines per tile and two ring slots (same arithmetic as below)
        auto fits = [&](int nl) {
            const size_t fixed = ((size_t) F.SC * (F.KL + F.KD) * nl + (size_t) F.blob_doubles) * 8 + 256;
            const size_t tile = ((size_t) F.SC * SWEEP_CH + F.KL + 18) * nl * 8;
            return fixed + 2 * tile <= 226 * 1024;
        };
        NL = 16;
        const int cap = contig ? 32 : 128;  // whole lines per tile: more of them only costs shared memory
        while (NL < cap && NL * F.SC <= 256 && fits(2 * NL)) NL *= 2;
    }
    const int NLt = NL / SWEEP_RL;
    if (contig && (off_in_h || off_out_h)) return -1;
    const int ncons = (NLt * F.SC + 31) / 32 * 32;
    if (ncons > 256) return -1;
    if (contig && SWEEP_CH % 2) return -1;  // the 128-bit chunk path needs an even CH
    auto even = [](long long v) { return (v & 1) == 0; };
    const bool ptr_ok = ((uintptr_t) G.in % 16 == 0) && ((uintptr_t) G.out % 16 == 0);
    SweepTileGeom& T = P.T;
    T = SweepTileGeom{};
    T.in = G.in;
    T.out = G.out;
    T.L0 = G.L0;
    T.L1 = G.L1;
    T.s0_in = G.s0_in;
    T.s1_in = G.s1_in;
    T.s0_out = G.s0_out;
    T.s1_out = G.s1_out;
    T.nb0 = (G.L0 + NL - 1) / NL;
    T.ntiles = T.nb0 * G.L1;
    T.reverse = G.reverse;
    const int rows_needed = F.SC * SWEEP_CH + F.KL;
    if (contig) {
        if (!ptr_ok || !even(G.s0_in) || !even(G.s1_in) || !even(G.s0_out) || !even(G.s1_out)) return -1;
        if (!even(F.n) && !G.pad_ok) return -1;  // bulk copies move multiples of 16 B
        T.ncopy = F.n + (F.n & 1);
        int pitch = rows_needed + (rows_needed & 1);
        while (pitch % 16 != 2) pitch += 2;  // lanes 16 B apart modulo 128 B: conflict-free 128-bit chunk access
        T.pitch = pitch;
        T.tile_doubles = NL * pitch;
    } else {
        if (int rc = sweep_strided_maps(G, F.n, NL, off_in_h, off_out_h, st, T)) return rc;
        T.tile_doubles = rows_needed * NL;
    }
    T.tile_doubles = (T.tile_doubles + 15) & ~15;  // 128 B granules
    const size_t fixed_doubles = (size_t) F.SC * (F.KL + F.KD) * NL + (size_t) F.blob_doubles;
    const size_t fixed_bytes = fixed_doubles * 8 + (2 * MAX_NBUF + 1) * 8 + 64;
    if ((uintptr_t) F.cfF % 16 != 0 || F.blob_doubles % 2 != 0) return -1;
    const size_t budget = 226 * 1024;
    if (fixed_bytes + 2 * (size_t) T.tile_doubles * 8 > budget) return -1;
    int nbuf = (int) ((budget - fixed_bytes) / ((size_t) T.tile_doubles * 8));
    if (nbuf > MAX_NBUF) nbuf = MAX_NBUF;
    T.nbuf = nbuf;
    P.smem = (size_t) nbuf * T.tile_doubles * 8 + fixed_bytes;
    P.NL = NL;
    P.ncons = ncons;
    return 0;
}

int device_sm_count() {
    static int sms = [] {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return sms;
}

std::map<std::vector<long long>, void*> g_multi_cache;  // SweepSegArgs arrays already resident on the device

}  // namespace

// Returns cudaSuccess (0) when the tile kernel was launched, -1 when this sweep is not eligible (the
// caller then uses the register-path kernel), or a CUDA error.
int launch_sweep_tile(const SweepFactor& F, const SweepGeom& G, bool contig, const long long* off_in_h,
                      const long long* off_out_h, cudaStream_t st) {
    TilePrep P;
    if (int rc = prepare_sweep_tile(F, G, contig, off_in_h, off_out_h, st, 0, P)) return rc;
    const int NL = P.NL;
    tile_kern_t k = NL == 12 ? pick<12>(F.KL, F.piv != 0, contig)
                    : NL == 32 ? pick<32>(F.KL, F.piv != 0, contig)
                    : NL == 64 ? pick<64>(F.KL, F.piv != 0, contig)
                    : NL == 128 ? pick<128>(F.KL, F.piv != 0, contig) : pick<16>(F.KL, F.piv != 0, contig);
    if (!k) return -1;
    cudaError_t e = cudaFuncSetAttribute((const void*) k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) P.smem);
    if (e != cudaSuccess) return (int) e;
    const int sms = device_sm_count();
    dim3 block(P.ncons + 32, 1, 1);
    const int cap = (G.max_ctas > 0 && G.max_ctas < sms) ? G.max_ctas : sms;
    dim3 grid(P.T.ntiles < cap ? P.T.ntiles : cap, 1, 1);
    return (int) launch_ex(k, grid, block, P.smem, st, true, F, P.T);
}

// nseg segments of the same lines (Fs[s] on the rows Gs[s] addresses) in one launch, the SMs shared out evenly.
// 0: launched; -1: not eligible (the caller launches segment by segment); else cudaError_t.
int launch_sweep_tile_multi(int nseg, const SweepFactor* Fs, const SweepGeom* Gs, bool contig, cudaStream_t st) {
    if (nseg < 1) return -1;
    std::vector<SweepSegArgs> args(nseg);
    size_t smem = 0;
    int ncons = 0, ntiles = 0;
    for (int s = 0; s < nseg; ++s) {
        TilePrep P;
        if (int rc = prepare_sweep_tile(Fs[s], Gs[s], contig, nullptr, nullptr, st, 16, P)) return rc;
        if (Fs[s].KL != Fs[0].KL || Fs[s].KD != Fs[0].KD || Fs[s].piv != Fs[0].piv) return -1;  // one kernel variant
        args[s].F = Fs[s];
        args[s].G = P.T;
        smem = std::max(smem, P.smem);
        ncons = std::max(ncons, P.ncons);
        ntiles = std::max(ntiles, P.T.ntiles);
    }
    tile_multi_kern_t k = pick_multi(Fs[0].KL, Fs[0].piv != 0, contig);
    if (!k) return -1;
    // the argument array lives on the device; one per distinct set of operands (a time loop alternates two)
    std::vector<long long> key;
    for (int s = 0; s < nseg; ++s)
        key.insert(key.end(), {(long long) (uintptr_t) Gs[s].in, (long long) (uintptr_t) Gs[s].out, (long long) (uintptr_t) Fs[s].cfF,
                               (long long) (uintptr_t) args[s].G.maps, Gs[s].L0, Gs[s].L1, Gs[s].s0_in, Gs[s].s1_in, Gs[s].s0_out,
                               Gs[s].s1_out, Gs[s].sj_in, Gs[s].sj_out, (long long) contig});
    void* d_args = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        auto it = g_multi_cache.find(key);
        if (it == g_multi_cache.end()) {
            if (cudaMalloc(&d_args, args.size() * sizeof(SweepSegArgs)) != cudaSuccess) return -1;
            cudaError_t e = cudaMemcpyAsync(d_args, args.data(), args.size() * sizeof(SweepSegArgs), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // `args` is a local
            if (e != cudaSuccess) return (int) e;
            it = g_multi_cache.emplace(std::move(key), d_args).first;
        }
        d_args = it->second;
    }
    cudaError_t e = cudaFuncSetAttribute((const void*) k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    const int sms = device_sm_count();
    const int cap = (Gs[0].max_ctas > 0 && Gs[0].max_ctas < sms) ? Gs[0].max_ctas : sms;
    const int per = std::max(1, std::min(ntiles, cap / nseg));
    dim3 block(ncons + 32, 1, 1), grid(per, nseg, 1);
    return (int) launch_ex(k, grid, block, smem, st, true, (const SweepSegArgs*) d_args);
}

}  // namespace adsb
