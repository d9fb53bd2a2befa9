// kernels_sweep_dist.cu -- K2 for a slab-sharded axis: the banded substitution along lines that are cut across
// GPUs, in ONE pass over the slab, fused with its boundary exchange over NVLink.
//
// Every rank runs this kernel on its own slab (one segment of every line; tables: build_segment_plan).  The
// algebra is the segmented substitution of kernels_seg.cu -- pass A (the slab's own columns of the factor),
// the KL forward / KD backward boundary values per line, pass B (x = xhat + Xi din + Psi tin) -- i.e. the
// recurrence of lin::solve_with_factorized -> dgbtrs_ (include/ads/lin/band_solve.hpp:21-31) with the same
// factor and pivots.  What is new is the schedule: the three stages run software-pipelined inside one
// persistent kernel, tile by tile, and talk to the neighbouring GPUs through peer pointers:
//
//   iteration i of a CTA (tile sequence identical on all ranks, CTA b pairs with CTA b of the neighbours):
//     A (tile i)       TMA load, local substitution in shared memory, TMA store of xhat in place;
//                      Dseg = E * xhat[last KL rows] stored into the NEXT rank's state array
//     B1 (tile i-K)    din <- the PREVIOUS rank's Dseg (polled, see below); X = xhat[first KD rows] + XiF din
//                      stored into the previous rank's state array
//     B2 (tile i-2K)   tin <- the NEXT rank's X (polled); the tile comes back by TMA (an L2 hit: it was stored
//                      2K tiles ago), x += Psi tin + Xi din in shared memory, TMA store in place
//
// so the NVLink latency (a few microseconds) hides behind K tiles of work, the slab is read from HBM once and
// written once (16 B/DOF: the intermediate xhat lives in L2 -- 2K tiles per CTA, tens of MB in all), and no
// host-side barrier separates the stages.  Warp roles: group A (pass A), group B (boundary values and pass B,
// so the latency of polling the neighbours overlaps pass A of later tiles), two producer warps (one per ring).
//
// Exchange protocol: the boundary values validate themselves.  Every word of the state arrays holds a sentinel
// (one particular signalling-NaN bit pattern, ADSB_DIST_SENTINEL) until the neighbour's 8-byte store replaces
// it; the receiver polls the word it needs with volatile loads, takes the value and puts the sentinel back
// for the next sweep.  No flags, no fences, no epochs: a first version that published per-tile flags behind
// __threadfence_system() spent 20 us per tile waiting for the fences to drain the SM's TMA traffic.
#include <cuda.h>

#include <cstdint>
#include <cstdlib>

#include "kernels.cuh"
#include "sweep_core.cuh"

namespace adsb {

namespace {

constexpr int DIST_NBUF = 2;   // slots of the pass-B ring
constexpr int DIST_NBUF_A = 3; // at most this many slots in the pass-A ring (G.nbuf: 3 when shared memory allows)

// Boundary values validate themselves: a word of the inbox holds the sentinel until the neighbour's store has
// replaced it.  peek() reads it (possibly still the sentinel), take() spins until the value is there and puts the
// sentinel back; it gives up after ~30 s (sets *err; the result is then garbage but nothing hangs) -- long enough
// for a neighbour whose host thread is late with its launch, short enough to end a run whose peer has died.
__device__ __forceinline__ unsigned long long peek_value(const double* slot) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(slot) : "memory");
    return v;
}
__device__ __forceinline__ double take_value(double* slot, unsigned long long v, int* err) {
    if (v == ADSB_DIST_SENTINEL_BITS) {
        const long long t0 = clock64();
        unsigned ns = 32;
        do {
            __nanosleep(ns);
            if (ns < 512) ns *= 2;
            v = peek_value(slot);
            // one time-out ends all waiting of this launch (and of the following ones until the caller clears *err)
            if (clock64() - t0 > 60000000000ll || (err && *reinterpret_cast<volatile int*>(err))) {
                if (err) atomicExch(err, 1);
                break;
            }
        } while (v == ADSB_DIST_SENTINEL_BITS);
    }
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot), "l"(ADSB_DIST_SENTINEL_BITS) : "memory");
    return __longlong_as_double((long long) v);
}

constexpr int DIST_NB = 128;  // threads of the B group (boundary values + pass B)

__device__ __forceinline__ void group_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Thread roles: [0, ncons) group A (pass A: lanes x chunks, as in sweep_tile_kernel), [ncons, ncons + 128) group B
// (boundary values and pass B), then two producer warps (ring A, ring D).  The groups meet only through
// mbarriers: group A hands the first KD rows of every finished tile to group B through a small ring.
template <int KL, int KD, bool PIV, int CH, int NL>
__global__ void __launch_bounds__(384, 1)
    sweep_dist_kernel(const SweepFactor F0, const SegDev T, const SweepTileGeom G, const SweepDistArgs D) {
    constexpr int RL = SWEEP_RL, NLt = NL / RL, KC = KD + KL, NB = DIST_NB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int SC = F0.SC, n = F0.n;
    const int ncons = (int) blockDim.x - NB - 64;
    const int tid = threadIdx.x;
    const int tile_doubles = G.tile_doubles;
    const int K = D.lag;
    const int r = D.rank, S = T.S;
    const int nbA = G.nbuf;
    double* ringA = reinterpret_cast<double*>(smem_raw);
    double* ringD = ringA + (size_t) nbA * tile_doubles;
    double* fst = ringD + (size_t) DIST_NBUF * tile_doubles;  // [SC][KL][NL]
    double* bst = fst + SC * KL * NL;                         // [SC][KD][NL]
    double* s_tab = bst + SC * KD * NL;                       // factor tables (blob layout)
    double* s_cf = s_tab + F0.blob_doubles;                   // [n][KC]   Psi | Xi of the slab's rows
    double* s_E = s_cf + ((n * KC + 1) & ~1);                 // [KL][KL]
    double* s_XiF = s_E + ((KL * KL + 1) & ~1);               // [KD][KL]
    double* s_xf = s_XiF + ((KD * KL + 1) & ~1);              // [K+1][KD][NL]  xhat first rows: group A -> group B
    double* s_din = s_xf + (K + 1) * KD * NL;                 // [K+1][KL][NL]  din of the tiles between B1 and B2
    double* s_tin = s_din + (K + 1) * KL * NL;                // [KD][NL]
    uint64_t* fullA = reinterpret_cast<uint64_t*>(s_tin + KD * NL);
    uint64_t* doneA = fullA + DIST_NBUF_A;
    uint64_t* fullD = doneA + DIST_NBUF_A;
    uint64_t* doneD = fullD + DIST_NBUF;
    uint64_t* tabbar = doneD + DIST_NBUF;
    uint64_t* xf_full = tabbar + 1;        // [K+1]
    uint64_t* xf_free = xf_full + (K + 1); // [K+1]
    volatile int* stored = reinterpret_cast<volatile int*>(xf_free + (K + 1));  // tiles whose pass-A store is complete

    pdl_launch();
    if (tid == 0) {
        for (int b = 0; b < DIST_NBUF_A; ++b) {
            mbar_init(&fullA[b], 1);
            mbar_init(&doneA[b], 1);
        }
        for (int b = 0; b < DIST_NBUF; ++b) {
            mbar_init(&fullD[b], 1);
            mbar_init(&doneD[b], 1);
        }
        for (int b = 0; b <= K; ++b) {
            mbar_init(&xf_full[b], 1);
            mbar_init(&xf_free[b], 1);
        }
        mbar_init(tabbar, 1);
        *stored = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // pad rows of both rings (rows past the slab read as zero) and the segment tables of this slab
    {
        const int nthr = (int) blockDim.x;
        const int pad0 = n * NL, pad = tile_doubles - pad0;
        for (int i = tid; i < (nbA + DIST_NBUF) * pad; i += nthr) ringA[(size_t) (i / pad) * tile_doubles + pad0 + i % pad] = 0.0;
        const int a = D.row_base;
        for (int i = tid; i < n * KC; i += nthr) s_cf[i] = T.cf[(size_t) a * KC + i];
        for (int i = tid; i < KL * KL; i += nthr) s_E[i] = T.E[(size_t) r * KL * KL + i];
        for (int i = tid; i < KD * KL; i += nthr) s_XiF[i] = T.XiF[(size_t) r * KD * KL + i];
    }
    __syncthreads();

    const int my_count = (G.ntiles - (int) blockIdx.x + (int) gridDim.x - 1) / (int) gridDim.x;
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(G.maps);
    // the lines are numbered flat (launch_sweep_dist insists on contiguous lines, L1 == 1): tile i of this CTA
    // starts at line (blockIdx.x + i * gridDim.x) * NL -- no divisions anywhere in the loops
    auto tile_of = [&](int i, int& bx, int& m) {
        bx = blockIdx.x + i * gridDim.x;
        m = 0;
    };
    const long long L = (long long) G.L0 * G.L1;

    if (tid >= ncons + NB) {
        const int pt = tid - ncons - NB;
        if (pt == 0) {
            // ---------------------------------------------------------------- producer of ring A (pass A)
            mbar_expect_tx(tabbar, (uint32_t) (F0.blob_doubles * 8));
            bulk_g2s(s_tab, F0.cfF, (uint32_t) (F0.blob_doubles * 8), tabbar);
            pdl_wait();
            auto load = [&](int i) {
                int bx, m;
                tile_of(i, bx, m);
                const int b = i % nbA;
                double* dst = ringA + (size_t) b * tile_doubles;
                mbar_expect_tx(&fullA[b], (uint32_t) G.load_bytes);
                for (int k = 0; k < G.nbox_in; ++k) tma_load_3d(dst + G.row0_in[k] * NL, maps + k, bx * NL, 0, m, &fullA[b]);
            };
            auto store = [&](int i) {
                int bx, m;
                tile_of(i, bx, m);
                const double* src = ringA + (size_t) (i % nbA) * tile_doubles;
                for (int k = 0; k < G.nbox_out; ++k)
                    tma_store_3d(maps + G.nbox_in + k, bx * NL, 0, m, src + G.row0_out[k] * NL);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            };
            for (int i = 0; i < nbA - 1 && i < my_count; ++i) load(i);
            for (int j = 0; j < my_count; ++j) {
                if (j + nbA - 1 < my_count) {
                    // the slot of tile j-1 is reloaded: its store must have read shared memory (with three slots
                    // the store before it, so the newest store may still be draining)
                    if (j >= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    load(j + nbA - 1);
                }
                mbar_wait(&doneA[j % nbA], (uint32_t) ((j / nbA) & 1));
                store(j);
                // all but the two most recent stores have reached memory: tiles 0 .. j-2 may be read back
                asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
                if (j >= 1) *stored = j - 1;
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __threadfence_block();
            *stored = my_count;
        } else if (pt == 32) {
            // ---------------------------------------------------------------- producer of ring D (pass B)
            pdl_wait();
            auto load = [&](int i) {
                while (*stored < i + 1) __nanosleep(40);  // pass A's store of this tile is complete
                int bx, m;
                tile_of(i, bx, m);
                const int b = i % DIST_NBUF;
                double* dst = ringD + (size_t) b * tile_doubles;
                mbar_expect_tx(&fullD[b], (uint32_t) G.load_bytes);
                // pass A stored through the `out` maps: read the tile back through the same ones
                for (int k = 0; k < G.nbox_out; ++k)
                    tma_load_3d(dst + G.row0_out[k] * NL, maps + G.nbox_in + k, bx * NL, 0, m, &fullD[b]);
            };
            auto store = [&](int i) {
                int bx, m;
                tile_of(i, bx, m);
                const double* src = ringD + (size_t) (i % DIST_NBUF) * tile_doubles;
                for (int k = 0; k < G.nbox_out; ++k)
                    tma_store_3d(maps + G.nbox_in + k, bx * NL, 0, m, src + G.row0_out[k] * NL);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            };
            for (int i = 0; i < DIST_NBUF - 1 && i < my_count; ++i) load(i);
            for (int j = 0; j < my_count; ++j) {
                if (j + DIST_NBUF - 1 < my_count) {
                    if (j >= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    load(j + DIST_NBUF - 1);
                }
                mbar_wait(&doneD[j % DIST_NBUF], (uint32_t) ((j / DIST_NBUF) & 1));
                store(j);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }

    if (tid >= ncons) {
        // ====================================================================== group B
        const int t = tid - ncons;
        const int lp = t % NLt, rg = t / NLt;   // lane pair and first row of this thread in pass B
        constexpr int RG = NB / NLt;            // rows advance by RG
        constexpr int NV1 = (KL * NL + NB - 1) / NB, NV2 = (KD * NL + NB - 1) / NB;  // polled words per thread
        // the words this thread polls for tile j: Dseg of the previous rank (stage B1), X of the next rank (B2)
        auto slot1 = [&](int j, int v) -> double* {
            const int i = t + v * NB;
            if (r == 0 || j < 0 || j >= my_count || i >= KL * NL) return nullptr;
            const long long l0 = (long long) (blockIdx.x + j * gridDim.x) * NL;
            const int kk = i / NL, ln = i % NL;
            if (l0 + ln >= G.L0) return nullptr;
            return D.dseg_local + ((size_t) (r - 1) * KL + kk) * L + l0 + ln;
        };
        auto slot2 = [&](int j, int v) -> double* {
            const int i = t + v * NB;
            if (r + 1 >= S || j < 0 || j >= my_count || i >= KD * NL) return nullptr;
            const long long l0 = (long long) (blockIdx.x + j * gridDim.x) * NL;
            const int ii = i / NL, ln = i % NL;
            if (l0 + ln >= G.L0) return nullptr;
            return D.x_local + ((size_t) (r + 1) * KD + ii) * L + l0 + ln;
        };
        // the loads of iteration jj + 1 are issued during iteration jj, so their latency (and, mostly, the wait
        // for the neighbour) is off the critical path
        unsigned long long pre1[NV1], pre2[NV2];
#pragma unroll
        for (int v = 0; v < NV1; ++v) {
            double* a = slot1(0, v);
            pre1[v] = a ? peek_value(a) : 0ull;
        }
#pragma unroll
        for (int v = 0; v < NV2; ++v) {
            double* a = slot2(-K, v);
            pre2[v] = a ? peek_value(a) : 0ull;
        }
        int s1 = 0, use1 = 0;  // hand-over ring slot of tile j1 (= j1 % (K+1)) and how often it has been used
        for (int jj = 0; jj < my_count + K; ++jj) {
            const int j1 = jj, j2 = jj - K;
            const int s2 = (s1 + 1 > K) ? 0 : s1 + 1;  // slot of tile j2 = j1 - K  ==  (j1 + 1) mod (K+1)
            const bool do1 = j1 < my_count, do2 = j2 >= 0;
            int bx1 = 0, m1 = 0;
            if (do1) tile_of(j1, bx1, m1);
            const long long line1 = (long long) bx1 * NL + (long long) m1 * G.L0;
            const int lanes1 = min(NL, G.L0 - bx1 * NL);
            double* dn1 = s_din + (size_t) s1 * KL * NL;
            // ---- the boundary values both stages need (zero where there is no neighbour / no line)
#pragma unroll
            for (int v = 0; v < NV1; ++v) {
                const int i = t + v * NB;
                if (do1 && i < KL * NL) {
                    double* a = slot1(j1, v);
                    dn1[i] = a ? take_value(a, pre1[v], D.error_flag) : 0.0;
                }
                double* nx = slot1(j1 + 1, v);
                pre1[v] = nx ? peek_value(nx) : 0ull;
            }
#pragma unroll
            for (int v = 0; v < NV2; ++v) {
                const int i = t + v * NB;
                if (do2 && i < KD * NL) {
                    double* a = slot2(j2, v);
                    s_tin[i] = a ? take_value(a, pre2[v], D.error_flag) : 0.0;
                }
                double* nx = slot2(j2 + 1, v);
                pre2[v] = nx ? peek_value(nx) : 0ull;
            }
            if (do1) mbar_wait(&xf_full[s1], (uint32_t) (use1 & 1));
            if (do2) mbar_wait(&fullD[j2 % DIST_NBUF], (uint32_t) ((j2 / DIST_NBUF) & 1));
            group_sync(2, NB);
            // ---- B1: X = xhat[first KD rows] + XiF din  -> previous rank
            if (do1 && D.x_prev) {
                const double* xf = s_xf + (size_t) s1 * KD * NL;
                for (int i = t; i < KD * NL; i += NB) {
                    const int ii = i / NL, ln = i % NL;
                    if (ln < lanes1) {
                        double acc = xf[i];
#pragma unroll
                        for (int q = 0; q < KL; ++q) acc = fma(s_XiF[ii * KL + q], dn1[q * NL + ln], acc);
                        D.x_prev[((size_t) r * KD + ii) * L + line1 + ln] = acc;
                    }
                }
            }
            // ---- B2: x = xhat + Psi tin + Xi din on the tile that came back
            if (do2) {
                const int b = j2 % DIST_NBUF;
                double* tile = ringD + (size_t) b * tile_doubles;
                int bx2, m2;
                tile_of(j2, bx2, m2);
                const int lanes2 = min(NL, G.L0 - bx2 * NL);
                const long long hoff2 = (long long) bx2 * NL + (long long) m2 * G.s1_out;  // offset of the tile's lines in a plane
                const double* dn2 = s_din + (size_t) s2 * KL * NL;
                double2 st[KC];
#pragma unroll
                for (int k = 0; k < KD; ++k) st[k] = *reinterpret_cast<const double2*>(s_tin + k * NL + 2 * lp);
#pragma unroll
                for (int k = 0; k < KL; ++k) st[KD + k] = *reinterpret_cast<const double2*>(dn2 + k * NL + 2 * lp);
                for (int row = rg; row < n; row += RG) {
                    double2 acc = *reinterpret_cast<const double2*>(tile + (size_t) row * NL + 2 * lp);
                    const double* cf = s_cf + (size_t) row * KC;
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        acc.x = fma(cf[k], st[k].x, acc.x);
                        acc.y = fma(cf[k], st[k].y, acc.y);
                    }
                    *reinterpret_cast<double2*>(tile + (size_t) row * NL + 2 * lp) = acc;
                }
                // the slab's first / last planes are the neighbours' halo planes of the next right-hand side: the
                // thread that corrected such a row stores it straight into their state buffers (peer stores; the
                // step's barrier orders them)
                if ((D.halo_prev || D.halo_next) && 2 * lp < lanes2) {
                    const int hp = D.halo_planes;
                    for (int h = 0; h < 2 * hp; ++h) {
                        const int row = h < hp ? h : n - 2 * hp + h;
                        double* dst = h < hp ? D.halo_prev : D.halo_next;
                        if (row % RG == rg && dst)
                            *reinterpret_cast<double2*>(dst + (long long) (h < hp ? h : h - hp) * G.s_row + hoff2 + 2 * lp) =
                                *reinterpret_cast<const double2*>(tile + (size_t) row * NL + 2 * lp);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            group_sync(2, NB);
            if (t == 0) {
                if (do1) mbar_arrive(&xf_free[s1]);
                if (do2) mbar_arrive(&doneD[j2 % DIST_NBUF]);
            }
            if (++s1 > K) {
                s1 = 0;
                ++use1;
            }
        }
        return;
    }

    // ========================================================================== group A
    mbar_wait(tabbar, 0);
    SweepFactor F = F0;
    F.cfF = s_tab;
    F.cfB = s_tab + F0.off[0];
    F.cfC = s_tab + F0.off[1];
    F.T = s_tab + F0.off[2];
    F.Rm = s_tab + F0.off[3];
    F.W = s_tab + F0.off[4];
    F.V = s_tab + F0.off[5];
    const int tx = tid % NLt;
    const int c = min(tid / NLt, SC - 1);  // padding threads shadow the last chunk (identical values)
    const int j0 = c * CH;
    int xs = 0, use = 0, b = 0, useA = 0;  // hand-over ring slot / use count; pass-A ring slot / use count
    for (int it = 0; it < my_count; ++it) {
        double* tile = ringA + (size_t) b * tile_doubles;
        mbar_wait(&fullA[b], (uint32_t) (useA & 1));
        double v[RL][CH + KL];
        {
            const double* mine = tile + (size_t) j0 * NL + 2 * tx;
#pragma unroll
            for (int q = 0; q < CH + KL; ++q) {
                const double2 t2 = *reinterpret_cast<const double2*>(mine + q * NL);
                v[0][q] = t2.x;
                v[1][q] = t2.y;
            }
        }
        sweep_core<KL, KD, PIV, CH, RL, true>(F, v, fst, bst, c, tx, NLt, SC, ncons);
        {
            double* mine = tile + (size_t) j0 * NL + 2 * tx;
#pragma unroll
            for (int q = 0; q < CH; ++q)
                if (j0 + q < n) *reinterpret_cast<double2*>(mine + q * NL) = make_double2(v[0][q], v[1][q]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // slot of the hand-over ring: free once group B has used its previous content
        if (use > 0) mbar_wait(&xf_free[xs], (uint32_t) ((use - 1) & 1));
        sweep_sync(ncons);
        // the tile is final for pass A: its first KD rows go to group B, Dseg to the next rank
        int bx, m;
        tile_of(it, bx, m);
        const long long line0 = (long long) bx * NL + (long long) m * G.L0;
        const int lanes = min(NL, G.L0 - bx * NL);
        double* xf = s_xf + (size_t) xs * KD * NL;
        for (int i = tid; i < KD * NL; i += ncons) xf[i] = tile[i];  // rows 0 .. KD-1 are the first KD*NL doubles
        if (D.dseg_next) {
            for (int i = tid; i < KL * NL; i += ncons) {
                const int kk = i / NL, ln = i % NL;
                if (ln < lanes) {
                    double acc = 0.0;
#pragma unroll
                    for (int mm = 0; mm < KL; ++mm) acc = fma(s_E[kk * KL + mm], tile[(size_t) (n - KL + mm) * NL + ln], acc);
                    D.dseg_next[((size_t) r * KL + kk) * L + line0 + ln] = acc;
                }
            }
        }
        sweep_sync(ncons);
        if (tid == 0) {
            mbar_arrive(&doneA[b]);
            mbar_arrive(&xf_full[xs]);
        }
        if (++xs > K) {
            xs = 0;
            ++use;
        }
        if (++b >= nbA) {
            b = 0;
            ++useA;
        }
    }
}

// Barrier with the two neighbouring ranks only (every dependency of the slab-sharded step is between
// neighbours): bump a device-side counter, publish it to the neighbours' flag words, wait for theirs.
// flags (symmetric memory, zero-initialised): [0] written by the previous rank, [1] by the next, [2] own counter.
__global__ void neighbor_barrier_kernel(unsigned long long* mine, unsigned long long* prev, unsigned long long* next,
                                        int* err) {
    if (threadIdx.x != 0) return;
    const unsigned long long c = mine[2] + 1;
    mine[2] = c;
    __threadfence_system();
    if (prev) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(prev + 1), "l"(c) : "memory");
    if (next) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(next), "l"(c) : "memory");
    const long long t0 = clock64();
    for (int side = 0; side < 2; ++side) {
        if (!(side == 0 ? prev : next)) continue;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + side) : "memory");
            if (clock64() - t0 > 60000000000ll) {
                if (err) atomicExch(err, 1);
                return;
            }
        } while (v < c);
    }
}

using dist_kern_t = void (*)(const SweepFactor, const SegDev, const SweepTileGeom, const SweepDistArgs);

template <int P, bool PIV, int CH>
dist_kern_t pick_nl(int NL) {
    constexpr int KD = PIV ? 2 * P : P;
    if constexpr (KD > CH) return nullptr;
    else switch (NL) {
        case 16: return (dist_kern_t) sweep_dist_kernel<P, KD, PIV, CH, 16>;
        case 32: return (dist_kern_t) sweep_dist_kernel<P, KD, PIV, CH, 32>;
        case 64: return (dist_kern_t) sweep_dist_kernel<P, KD, PIV, CH, 64>;
        default: return nullptr;
        }
}

template <int CH>
dist_kern_t pick_ch(int KL, bool piv, int NL) {
    switch (KL) {
    case 1: return piv ? pick_nl<1, true, CH>(NL) : pick_nl<1, false, CH>(NL);
    case 2: return piv ? pick_nl<2, true, CH>(NL) : pick_nl<2, false, CH>(NL);
    case 3: return piv ? pick_nl<3, true, CH>(NL) : pick_nl<3, false, CH>(NL);
    case 4: return piv ? pick_nl<4, true, CH>(NL) : pick_nl<4, false, CH>(NL);
    case 5: return piv ? pick_nl<5, true, CH>(NL) : pick_nl<5, false, CH>(NL);
    default: return nullptr;
    }
}

// chunk length of the slab's factor plan: 18 columns as everywhere, or 12 -- shorter chunks mean more chunks per
// line, i.e. more threads in the pass-A group of a CTA whose tiles are only 16 .. 64 lines wide
dist_kern_t pick(int KL, bool piv, int NL, int CH) {
    return CH == SWEEP_CH ? pick_ch<SWEEP_CH>(KL, piv, NL) : CH == SWEEP_CH_DIST ? pick_ch<SWEEP_CH_DIST>(KL, piv, NL) : nullptr;
}

}  // namespace

int launch_neighbor_barrier(unsigned long long* mine, unsigned long long* prev, unsigned long long* next, int* err,
                            cudaStream_t st) {
    neighbor_barrier_kernel<<<1, 32, 0, st>>>(mine, prev, next, err);
    return (int) cudaGetLastError();
}

// 0: launched; -1: not eligible (the caller runs pass A / boundary kernels / pass B separately); else cudaError_t
int launch_sweep_dist(const SweepFactor& F, int CH, const SegDev& T, const SweepGeom& G, const SweepDistArgs& D, int NL,
                      cudaStream_t st, bool dry_run) {
    if (T.DF != 1 || T.DB != 1) return -1;                   // neighbours only
    if (T.KL != F.KL || T.KD != F.KD) return -1;             // the slab's own factor needs the same kernel variant
    if (G.in != G.out || G.sj_in != G.sj_out || G.s1_in != G.s1_out) return -1;  // in place
    if (NL != 16 && NL != 32 && NL != 64) return -1;
    if (G.L1 != 1) return -1;  // contiguous lines, numbered flat (the z lines of an x-fastest tensor without row gaps)
    const int NLt = NL / SWEEP_RL;
    const int ncons = (NLt * F.SC + 31) / 32 * 32;
    if (ncons > 192 || NLt > DIST_NB || DIST_NB % NLt != 0 || D.lag < 1 || D.lag > 16) return -1;
    if ((uintptr_t) F.cfF % 16 != 0 || F.blob_doubles % 2 != 0) return -1;
    SweepTileGeom Tg{};
    Tg.in = G.in;
    Tg.out = G.out;
    Tg.L0 = G.L0;
    Tg.L1 = G.L1;
    Tg.s0_in = Tg.s0_out = G.s0_in;
    Tg.s1_in = Tg.s1_out = G.s1_in;
    Tg.nb0 = (G.L0 + NL - 1) / NL;
    Tg.ntiles = Tg.nb0 * G.L1;
    Tg.s_row = G.sj_out;
    if ((D.halo_prev || D.halo_next) && (D.halo_planes < 1 || D.halo_planes > F.n || (G.L0 & 1))) return -1;
    if (!dry_run)
        if (int rc = sweep_strided_maps(G, F.n, NL, nullptr, nullptr, st, Tg)) return rc;
    const int rows_needed = F.SC * CH + F.KL;
    Tg.tile_doubles = (rows_needed * NL + 15) & ~15;
    const int KC = T.KD + T.KL, K = D.lag;
    const size_t rest = (size_t) F.SC * (F.KL + F.KD) * NL + F.blob_doubles + ((F.n * KC + 1) & ~1) + ((T.KL * T.KL + 1) & ~1) +
                        ((T.KD * T.KL + 1) & ~1) + (size_t) (K + 1) * T.KD * NL + (size_t) (K + 1) * T.KL * NL + (size_t) T.KD * NL;
    const size_t fixed_b = rest * 8 + (2 * DIST_NBUF_A + 2 * DIST_NBUF + 1 + 2 * (K + 1)) * 8 + 64;
    Tg.nbuf = (fixed_b + (size_t) (DIST_NBUF_A + DIST_NBUF) * Tg.tile_doubles * 8 <= 226 * 1024) ? DIST_NBUF_A : 2;
    const size_t doubles = (size_t) (Tg.nbuf + DIST_NBUF) * Tg.tile_doubles + (size_t) F.SC * (F.KL + F.KD) * NL + F.blob_doubles +
                           ((F.n * KC + 1) & ~1) + ((T.KL * T.KL + 1) & ~1) + ((T.KD * T.KL + 1) & ~1) +
                           (size_t) (K + 1) * T.KD * NL + (size_t) (K + 1) * T.KL * NL + (size_t) T.KD * NL;
    const size_t smem = doubles * 8 + (2 * DIST_NBUF_A + 2 * DIST_NBUF + 1 + 2 * (K + 1)) * 8 + 64;
    if (smem > 226 * 1024) return -1;
    dist_kern_t k = pick(F.KL, F.piv != 0, NL, CH);
    if (!k) return -1;
    if (dry_run) return 0;
    cudaError_t e = cudaFuncSetAttribute((const void*) k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    static int sms = [] {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    int cap = (G.max_ctas > 0 && G.max_ctas < sms) ? G.max_ctas : sms;
    dim3 block(ncons + DIST_NB + 64, 1, 1), grid(Tg.ntiles < cap ? Tg.ntiles : cap, 1, 1);
    return (int) launch_ex(k, grid, block, smem, st, true, F, T, Tg, D);
}

}  // namespace adsb
