// K2 + K3: the `--fast` ("quasi-fast folding") search step for sm_100a.
//
// Replaces hendrics/efsearch.py:876-925 (search_with_qffa_step):
//   K2  _fast_phase / _fast_phase_fdot / _fast_phase_fddot (efsearch.py:854-873) fused with
//       histogram2d -> _hist2d_numba_seq (hendrics/base.py:667-693): one streaming pass over the
//       events builds the (time slice x phase bin) profiles of up to QF_CMAX centre frequencies.
//       HBM-bound: 8 B/event per pass; phases and per-event bins never leave the SM.
//   K3  _fast_step / shift_and_sum (efsearch.py:817-851, 350-363) + _z_n_fast_cached
//       (hendrics/ffa.py:59-104): for every (L, Q) bin-shift pair, roll each time slice by
//       rint(b*L + b^2*Q) mod nbin, sum, and reduce the summed profile to Z^2_n.
//
// Arithmetic follows the reference operation by operation (separately rounded fp64 products and
// sums, half-to-even rint, truncating bin indices, right edges excluded), so the 2-D profiles are
// bit-exact and the statistics differ only by summation order.
#include <algorithm>
#include <cmath>
#include <vector>

#include "hx_common.cuh"

namespace hx {

constexpr int QF_CMAX = 8;      // centres fused per pass over the events
constexpr int QF_WROWS = 4;     // time slices kept in the sliding shared-memory window (power of two)
constexpr int QF_THREADS = 256;
constexpr int QF_TILE = 1024;   // events per pipeline stage (8 KB)
constexpr int QF_STAGES = 4;    // pipeline depth: HBM latency x bandwidth needs ~90 KB in flight per SM
constexpr int QF_HDR = 64;      // QF_STAGES mbarriers + QF_STAGES "warps done" counters

struct QffaCentres {
    double f[QF_CMAX], fd[QF_CMAX], fdd[QF_CMAX];
};

struct QffaHistArgs {
    const double *t;
    long long n_ev;
    long long ev_per_cta;   // contiguous events per CTA, multiple of QF_TILE
    double t_first;         // ranges[1][0]
    double delta0, delta1;  // 1/((1-0)/nbin), 1/((t_last-t_first)/nprof)  (host, reference formula)
    double t_safe;          // FAST: events with |t| >= t_safe have |phase| >= 0.5 for every centre
    int nbin, nprof;
    int logk;        // log2 of the bank-spreading replication of the shared histogram
    uint32_t *prof;  // [C][nprof][nbin]
    QffaCentres c;   // FAST: pre-scaled by nbin * 2^(logk+2) (exact): floor() is the byte offset of the bin
    QffaCentres raw;
    double *unsafe_t;         // FAST: events with |t| < t_safe are appended here for the generic list kernel
    unsigned int *unsafe_n;   //       (count; capacity unsafe_cap, overflow => host falls back)
    unsigned int unsafe_cap;
};

// Phase bin of one event for one centre, reference operation order (efsearch.py:854-873):
// MODE 0: t*f ; 1: + 0.5*t*t*fd ; 2: + (1/6)*t*t*t*fdd.  Returns -1 when the reference drops the event
// (x rounded to 1.0 => i == nbin fails `i < bins[0]`, base.py:674).
template <int MODE>
__device__ __forceinline__ int qffa_bin_generic(double t, double h, double s6, double f, double fd, double fdd,
                                                double delta0, int nbin) {
    double ph = __dmul_rn(t, f);
    if (MODE >= 1) ph = __dadd_rn(ph, __dmul_rn(h, fd));
    if (MODE >= 2) ph = __dadd_rn(ph, __dmul_rn(s6, fdd));
    const double fl = __dsub_rn(__dadd_rd(ph, HX_MAGIC), HX_MAGIC);  // floor(ph), exact
    const double x = __dsub_rn(ph, fl);                              // phases - floor(phases)
    const double fi = __dmul_rn(x, delta0);                          // (x - 0) * delta0
    const int i = __double2loint(__dadd_rd(fi, HX_MAGIC));           // int(i), i >= 0
    return i < nbin ? i : -1;
}

// FAST (power-of-two nbin, |phase| >= 0.5): phase - floor(phase) is exact there, delta0 == nbin, so
// int(frac(ph)*nbin) == floor(nbin*ph) mod nbin.  Constants are pre-scaled by nbin*2^(logk+2), so the
// masked low word of floor() is directly the byte offset of the bin inside a (centre, row) slot.
template <int MODE>
__device__ __forceinline__ uint32_t qffa_binoff_fast(double t, double h, double s6, double fN, double fdN,
                                                     double fddN, uint32_t mask) {
    double ph = __dmul_rn(t, fN);
    if (MODE >= 1) ph = __dadd_rn(ph, __dmul_rn(h, fdN));
    if (MODE >= 2) ph = __dadd_rn(ph, __dmul_rn(s6, fddN));
    return (uint32_t)__double2loint(__dadd_rd(ph, HX_MAGIC)) & mask;
}

// Persistent CTAs: each owns a contiguous range of the (sorted) events, streamed through a 2-stage
// cp.async.bulk + mbarrier pipeline.  Counts go to a sliding window of QF_WROWS time slices in shared
// memory (slot = row % QF_WROWS, k bank-spread copies per bin); because the events are sorted, a row is
// complete as soon as a stage starts beyond it, and is then flushed to global memory and recycled.
// FAST: only events with |t| >= t_safe are binned here (5 instructions per event*centre); the few
// others are appended to a list for qffa_list_kernel, which replays the reference operation by operation.
template <int C, int MODE, bool FAST>
__global__ void __launch_bounds__(QF_THREADS) qffa_hist_kernel(const QffaHistArgs a) {
    extern __shared__ __align__(16) unsigned char qsmem[];
    // layout: QF_STAGES mbarriers | QF_STAGES done counters | tiles QF_STAGES x QF_TILE x 8 B | window
    const uint32_t sbase = smem_u32(qsmem);
    int *done = reinterpret_cast<int *>(qsmem + QF_STAGES * 8);
    const double *tiles = reinterpret_cast<const double *>(qsmem + QF_HDR);
    uint32_t *sh = reinterpret_cast<uint32_t *>(qsmem + QF_HDR + QF_STAGES * QF_TILE * 8);  // [C][QF_WROWS][nbin][k]
    const uint32_t sh_addr = sbase + QF_HDR + QF_STAGES * QF_TILE * 8;
    const int nbin = a.nbin, nprof = a.nprof, logk = a.logk;
    const int rsize = nbin << logk;      // words of one (centre, row slot)
    const int wsize = QF_WROWS * rsize;  // words of one centre's window
    const int tid = threadIdx.x, lane = tid & 31;
    const int nwarps = QF_THREADS / 32;
    for (int i = tid; i < C * wsize; i += QF_THREADS) sh[i] = 0u;

    const long long e0 = (long long)blockIdx.x * a.ev_per_cta;
    const long long e1 = min(a.n_ev, e0 + a.ev_per_cta);
    const int ntiles = e1 > e0 ? (int)((e1 - e0 + QF_TILE - 1) / QF_TILE) : 0;
    const uint32_t sub = (uint32_t)tid & ((1u << logk) - 1u);
    const uint32_t binmask = (uint32_t)(nbin - 1) << (logk + 2);  // FAST: byte offset of the bin in a slot
    const uint32_t cstride = (uint32_t)wsize * 4u;                // bytes between centres
    int slot_shift = 0;                                           // FAST (nbin power of two): log2(bytes of a slot)
    while ((1u << slot_shift) < (uint32_t)rsize * 4u) slot_shift++;

    auto issue = [&](int it) {
        const long long s = e0 + (long long)it * QF_TILE;
        const int cnt = (int)min((long long)QF_TILE, e1 - s);
        const uint32_t bytes = (uint32_t)cnt * 8u, bulk = bytes & ~15u;
        const int st = it % QF_STAGES;
        const uint32_t dst = sbase + QF_HDR + (uint32_t)st * QF_TILE * 8;
        const uint32_t bar = sbase + (uint32_t)st * 8;
        const char *src = reinterpret_cast<const char *>(a.t) + s * 8;
        if (bytes != bulk)
            *reinterpret_cast<double *>(qsmem + QF_HDR + (size_t)st * QF_TILE * 8 + bulk) =
                *reinterpret_cast<const double *>(src + bulk);
        if (bulk) {
            mbar_expect_tx(bar, bulk);
            bulk_g2s(dst, src, bulk, bar);
        } else {
            mbar_arrive(bar);
        }
    };
    if (tid == 0) {
        for (int st = 0; st < QF_STAGES; st++) {
            mbar_init(sbase + (uint32_t)st * 8, 1);
            done[st] = 0;
        }
        fence_mbar_init();
        fence_proxy_async();
        for (int st = 0; st < QF_STAGES; st++)
            if (ntiles > st) issue(st);
    }
    __syncthreads();

    // row of an event time; -1 when outside [0, nprof) (right edge excluded, base.py:674)
    auto row_of = [&](double t) -> int {
        const double jf = __dmul_rn(__dsub_rn(t, a.t_first), a.delta1);
        return (jf >= 0.0 && jf < (double)nprof) ? __double2loint(__dadd_rd(jf, HX_MAGIC)) : -1;
    };
    // flush rows [r0, r1) of the window to global memory and clear their slots
    auto flush_rows = [&](int r0, int r1) {
        for (int r = r0; r < r1; r++) {
            const int slot = r & (QF_WROWS - 1);
            for (int i = tid; i < C * nbin; i += QF_THREADS) {
                const int c = i / nbin, b = i - c * nbin;
                uint32_t *p = sh + c * wsize + slot * rsize + (b << logk);
                uint32_t v = 0;
                for (int s = 0; s < (1 << logk); s++) v += p[s], p[s] = 0u;
                if (v && r >= 0 && r < nprof) atomicAdd(&a.prof[((size_t)c * nprof + r) * nbin + b], v);
            }
        }
    };

    int j_base = 0;  // first row of the window (rows j_base .. j_base + QF_WROWS - 1 are resident)
    for (int it = 0; it < ntiles; it++) {
        const int buf = it % QF_STAGES;
        mbar_wait(sbase + (uint32_t)buf * 8, (uint32_t)(it / QF_STAGES) & 1u);
        const double *tp = tiles + (size_t)buf * QF_TILE;
        const int cnt = (int)min((long long)QF_TILE, e1 - (e0 + (long long)it * QF_TILE));
        // slide the window: every row before the first row of this stage is complete (sorted events)
        int j_first = row_of(tp[0]);
        if (j_first < 0) j_first = (it == 0) ? 0 : j_base;
        if (it == 0) {
            j_base = j_first;
        } else if (j_first > j_base) {
            __syncthreads();  // all increments of the previous stage have been issued by every warp
            flush_rows(j_base, min(j_first, j_base + QF_WROWS));
            j_base = j_first;
            __syncthreads();
        }
#pragma unroll 2
        for (int e = tid; e < cnt; e += QF_THREADS) {
            const double t = tp[e];
            // all predicates first, one (almost always taken) branch: keeps the loop body straight-line
            const double jf = __dmul_rn(__dsub_rn(t, a.t_first), a.delta1);
            const bool valid = jf >= 0.0 && jf < (double)nprof;  // right edge excluded (base.py:674)
            const int j = __double2loint(__dadd_rd(jf, HX_MAGIC));
            const bool in_win = (unsigned)(j - j_base) < (unsigned)QF_WROWS;
            const int slot = j & (QF_WROWS - 1);
            double tssq = 0.0, h = 0.0, s6 = 0.0;
            if (MODE >= 1) {
                tssq = __dmul_rn(t, t);
                h = __dmul_rn(0.5, tssq);  // == (0.5*t)*t bit for bit
            }
            if (MODE >= 2) s6 = __dmul_rn(__dmul_rn(1.0 / 6, tssq), t);
            if (FAST) {
                const bool safe = fabs(t) >= a.t_safe;
                if (valid && in_win && safe) {
                    const uint32_t evoff = sh_addr + ((uint32_t)slot << slot_shift) + sub * 4u;
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        const uint32_t off = qffa_binoff_fast<MODE>(t, h, s6, a.c.f[c], a.c.fd[c], a.c.fdd[c], binmask);
                        red_shared_inc(off + evoff + (uint32_t)c * cstride);
                    }
                } else if (valid && !safe) {  // rare: defer to the generic list kernel
                    const unsigned int pos = atomicAdd(a.unsafe_n, 1u);
                    if (pos < a.unsafe_cap) a.unsafe_t[pos] = t;
                } else if (valid) {  // data gap or unsorted input: correct, just slower
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        const uint32_t off = qffa_binoff_fast<MODE>(t, h, s6, a.c.f[c], a.c.fd[c], a.c.fdd[c], binmask);
                        atomicAdd(&a.prof[((size_t)c * nprof + j) * nbin + (off >> (logk + 2))], 1u);
                    }
                }
            } else if (valid) {
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const int i = qffa_bin_generic<MODE>(t, h, s6, a.raw.f[c], a.raw.fd[c], a.raw.fdd[c], a.delta0, nbin);
                    if (i < 0) continue;
                    if (in_win)
                        atomicAdd(&sh[c * wsize + slot * rsize + ((i << logk) | sub)], 1u);
                    else
                        atomicAdd(&a.prof[((size_t)c * nprof + j) * nbin + i], 1u);
                }
            }
        }
        // stage hand-back: the last warp to finish re-arms the buffer
        __syncwarp();
        if (lane == 0) {
            const int prev = atomicAdd(&done[buf], 1);
            if (prev == nwarps - 1) {
                done[buf] = 0;
                if (it + QF_STAGES < ntiles) {
                    fence_proxy_async();
                    issue(it + QF_STAGES);
                }
            }
        }
    }
    __syncthreads();
    flush_rows(j_base, j_base + QF_WROWS);
}

// The events the FAST kernel deferred (|t| < t_safe): reference arithmetic, global atomics.
template <int MODE>
__global__ void __launch_bounds__(256) qffa_list_kernel(const QffaHistArgs a, int C) {
    if (*a.unsafe_n > a.unsafe_cap && blockIdx.x == 0 && threadIdx.x == 0) a.unsafe_n[1] = 1u;  // overflow flag
    const unsigned int n = min(*a.unsafe_n, a.unsafe_cap);
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const double t = a.unsafe_t[e];
        const double jf = __dmul_rn(__dsub_rn(t, a.t_first), a.delta1);
        if (!(jf >= 0.0 && jf < (double)a.nprof)) continue;
        const int j = __double2loint(__dadd_rd(jf, HX_MAGIC));
        double tssq = 0.0, h = 0.0, s6 = 0.0;
        if (MODE >= 1) {
            tssq = __dmul_rn(t, t);
            h = __dmul_rn(0.5, tssq);
        }
        if (MODE >= 2) s6 = __dmul_rn(__dmul_rn(1.0 / 6, tssq), t);
        for (int c = 0; c < C; c++) {
            const int i = qffa_bin_generic<MODE>(t, h, s6, a.raw.f[c], a.raw.fd[c], a.raw.fdd[c], a.delta0, a.nbin);
            if (i >= 0) atomicAdd(&a.prof[((size_t)c * a.nprof + j) * a.nbin + i], 1u);
        }
    }
}


// ================================================================================================
// K2-lane: the `--fast` 2-D profiles with lane = centre frequency (the K1 layout applied to K2).
//
// qffa_hist_kernel (lane = event) is bound by bank conflicts of its shared-memory atomics: 32 photons
// of a warp hit 32 random phase bins.  Here every lane owns one centre frequency (R per thread) and a
// private [bin][lane] histogram column, every increment is a conflict-free ATOMS, and the events --
// sorted in time -- are walked one time slice ("row") after the other: the row boundaries are found
// once per call by binary search (the row index fl(fl(t - t0) * delta1) is monotone in t), the inner
// loop is K1's straight-line loop, and a column is drained into prof[row] when its row ends.
// Photons with |t| < t_safe, for which `phase - floor(phase)` is not exact, are folded like the others
// and then corrected by qffa_fix_kernel, which re-bins them with the reference's operation sequence.
// Requires power-of-two nbin and sorted times; otherwise the lane = event kernel above is used.
// ================================================================================================
struct QLaneArgs {
    const double *t;
    long long n_ev;
    const long long *row_start;  // [nprof + 1] first event of every row; row_start[nprof] = first dropped event
    const double *f, *fd, *fdd;  // [n_c] centre constants of this batch (device)
    int n_c;
    int nbin, nprof;
    long long ev_per_cta;  // events per CTA along grid.y (even), over [0, row_start[nprof])
    double scale;      // nbin * 128
    uint32_t binmask;  // (nbin-1) << 7, packed: (nbin/2-1) << 7
    uint32_t topbit;
    int pack_shift;
    int flush_events;  // packed: drain at least every this many events
    uint32_t *prof;    // [ceil(n_c/32)][nprof][nbin][32]   (COEF = false)
    int warps;
    uint32_t hist_off, tile_off;
    int tile_ev;
    int *errflag;
    // COEF = true: a finished time slice leaves the SM as its harmonic coefficients, not as its nbin counts
    double *coef;                 // [n_c][nprof][nharm][2]  (Cr_h, Sr_h), accumulated with fp64 atomics
    unsigned long long *rowcnt;   // [n_c][nprof] events of the slice
    const double *tab;            // [2 * nbin] cos | sin (global); staged at extra_off
    int nharm;                    // <= QL_NHMAX
    uint32_t extra_off;           // the staged tables
    int rows_per_cta;             // > 0: CTA y folds the time slices [y * rows_per_cta, (y + 1) * rows_per_cta) and nothing else
    // streamed upload (hx_qffa_search from a host array): the events arrive chunk by chunk over PCIe WHILE this kernel
    // runs; *progress = number of chunks of `chunk_events` events already in place (written by the copy stream after
    // each chunk).  A CTA waits until its whole event range has arrived.  NULL: everything is resident.
    const int *progress;
    long long chunk_events;
};
constexpr int QL_NHMAX = 4;

template <int R, int MODE, bool PACKED, bool COEF = false>
__global__ void __launch_bounds__(512, 1) qffa_lane_kernel(const QLaneArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int UE = 4;
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t blk = (uint32_t)a.nbin << (PACKED ? 6 : 7);
    const uint32_t hist = sbase + a.hist_off;
    const uint32_t tileb = sbase + a.tile_off;
    const int TILE_EV = a.tile_ev;
    const uint32_t TILE_BYTES = (uint32_t)(TILE_EV + 16) * 8u;
    int *done = reinterpret_cast<int *>(smem + 16);
    const int nbin = a.nbin, nprof = a.nprof, nwarps = a.warps;
    if ((hist & (blk - 1)) != 0) {
        if (tid == 0) atomicExch(a.errflag, 1);
        return;
    }
    for (uint32_t o = tid * 4; o < (uint32_t)nwarps * R * blk; o += blockDim.x * 4) sts_u32(hist + o, 0u);
    double *tsm = reinterpret_cast<double *>(smem + a.extra_off);   // COEF: cos | sin tables
    if (COEF)
        for (int i = tid; i < 2 * nbin; i += blockDim.x) tsm[i] = a.tab[i];

    // centre constants: slot = 32 centres
    const long long slot0 = ((long long)blockIdx.x * nwarps + warp) * R;
    double fN[R], fdN[MODE >= 1 ? R : 1], fddN[MODE >= 2 ? R : 1];
    uint32_t cb[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const long long c = (slot0 + r) * 32 + lane;
        const bool valid = c < a.n_c;
        fN[r] = valid ? __dmul_rn(a.f[c], a.scale) : 0.0;
        if (MODE >= 1) fdN[r] = valid ? __dmul_rn(a.fd[c], a.scale) : 0.0;
        if (MODE >= 2) fddN[r] = valid ? __dmul_rn(a.fdd[c], a.scale) : 0.0;
        cb[r] = hist + (uint32_t)(warp * R + r) * blk + (uint32_t)lane * 4u;
    }

    // events of this CTA (a contiguous chunk; rows may be shared with the neighbouring CTAs: drains are atomic)
    const long long n_valid = a.row_start[nprof];
    long long ev_lo = min(n_valid, (long long)blockIdx.y * a.ev_per_cta);
    long long ev_hi = min(n_valid, ev_lo + a.ev_per_cta);
    if (a.rows_per_cta > 0) {  // whole time slices per CTA: every slice is folded by exactly one CTA
        const int j_lo = (int)min((long long)nprof, (long long)blockIdx.y * a.rows_per_cta);
        ev_lo = a.row_start[j_lo], ev_hi = a.row_start[min(nprof, j_lo + a.rows_per_cta)];
    }
    if (a.progress != nullptr && ev_hi > ev_lo) {
        if (tid == 0) {
            const int need = (int)((ev_hi + a.chunk_events - 1) / a.chunk_events);
            const volatile int *prog = a.progress;
            unsigned long long t0 = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while (*prog < need) {
                __nanosleep(2000);
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) {  // 20 s without the chunk: the host side failed; give up loudly
                    atomicExch(a.errflag, 2);
                    break;
                }
            }
            __threadfence();
            fence_proxy_async();  // the copy engine's writes, seen by this thread, before the bulk copies that read them
        }
        __syncthreads();
        if (*reinterpret_cast<volatile int *>(a.errflag) == 2) return;
    }
    int j0;  // row of the first event: last j with row_start[j] <= ev_lo
    {
        int lo = 0, hi = nprof;  // invariant: row_start[lo] <= ev_lo, answer in [lo, hi)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.row_start[mid] <= ev_lo) lo = mid;
            else hi = mid;
        }
        j0 = lo;
    }
    const int j1 = nprof;
    const long long es = ev_lo & ~1LL;  // 16-byte aligned stream start (a slice boundary may be odd: one event is skipped)
    const int ntiles = ev_hi > es ? (int)((ev_hi - es + TILE_EV - 1) / TILE_EV) : 0;
    const uint32_t bar[2] = {sbase, sbase + 8};
    auto issue = [&](int it) {
        const long long s = es + (long long)it * TILE_EV;
        const int cnt = (int)min((long long)TILE_EV, ev_hi - s);
        const uint32_t bytes = (uint32_t)cnt * 8u, bulk = bytes & ~15u;
        const uint32_t dst = tileb + (uint32_t)(it & 1) * TILE_BYTES;
        const char *src = reinterpret_cast<const char *>(a.t) + s * 8;
        if (bytes != bulk)
            *reinterpret_cast<double *>(smem + a.tile_off + (uint32_t)(it & 1) * TILE_BYTES + bulk) =
                *reinterpret_cast<const double *>(src + bulk);
        if (bulk) {
            mbar_expect_tx(bar[it & 1], bulk);
            bulk_g2s(dst, src, bulk, bar[it & 1]);
        } else {
            mbar_arrive(bar[it & 1]);
        }
    };
    if (tid == 0) {
        mbar_init(bar[0], 1);
        mbar_init(bar[1], 1);
        done[0] = 0, done[1] = 0;
        fence_mbar_init();
        fence_proxy_async();
        if (ntiles > 0) issue(0);
        if (ntiles > 1) issue(1);
    }
    __syncthreads();

    const uint32_t mask = a.binmask, topbit = a.topbit;
    // drain this thread's columns into prof[slot][row][bin][lane] (coalesced across the warp) and clear them
    const long long n_slots = ((long long)a.n_c + 31) >> 5;
    auto drain = [&](int row) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (slot0 + r >= n_slots) continue;  // padding slot of the last CTA: nothing to keep
            uint32_t *dst = a.prof + (((size_t)(slot0 + r) * nprof + row) * nbin) * 32 + lane;
            if (PACKED) {
                const int half = nbin >> 1;
                const uint32_t lowmask = topbit - 1u;
                for (int w = 0; w < half; w++) {
                    const uint32_t ad = cb[r] + ((uint32_t)w << 7);
                    const uint32_t v = lds_u32(ad);
                    if (v) {
                        sts_u32(ad, 0u);
                        const uint32_t up = v >> a.pack_shift, lowc = (v & lowmask) - up;
                        if (lowc) atomicAdd(dst + (size_t)w * 32, lowc);
                        if (up) atomicAdd(dst + (size_t)(w + half) * 32, up);
                    }
                }
            } else {
                for (int b = 0; b < nbin; b++) {
                    const uint32_t ad = cb[r] + ((uint32_t)b << 7);
                    const uint32_t v = lds_u32(ad);
                    if (v) {
                        sts_u32(ad, 0u);
                        atomicAdd(dst + (size_t)b * 32, v);
                    }
                }
            }
        }
    };
    // COEF: add the harmonic coefficients of this thread's columns (a partial DFT of the counts of time slice `row`: a DFT
    // is linear, so the packed counters can be emptied as often as they need) to coef[centre][row] and clear the columns.
    // One thread's successive adds to an address are ordered; with snap_rows no other CTA touches the slice, so the
    // result is reproducible bit for bit.  Otherwise (slices larger than a CTA's range) the adds of different CTAs
    // interleave: correct, reproducible to rounding.
    auto accumulate = [&](int row) {
        const double *tc = tsm, *ts = tsm + nbin;
        const int nh = a.nharm;
#pragma unroll
        for (int r = 0; r < R; r++) {
            double c[QL_NHMAX] = {0.0, 0.0, 0.0, 0.0}, sn[QL_NHMAX] = {0.0, 0.0, 0.0, 0.0}, cnt = 0.0;
            if (PACKED) {
                // bins w and w + nbin/2 share a word; harmonic k of bin w + nbin/2 is (-1)^k times that of bin w
                const int half = nbin >> 1;
                const uint32_t lowmask = topbit - 1u;
                for (int w = 0; w < half; w++) {
                    const uint32_t ad = cb[r] + ((uint32_t)w << 7);
                    const uint32_t v = lds_u32(ad);
                    if (v) {
                        sts_u32(ad, 0u);
                        const uint32_t up = v >> a.pack_shift, lowc = (v & lowmask) - up;
                        const double dsum = (double)(lowc + up), ddif = (double)lowc - (double)up;
                        cnt += dsum;
#pragma unroll
                        for (int h = 0; h < QL_NHMAX; h++)
                            if (h < nh) {
                                const int idx = ((h + 1) * w) & (nbin - 1);
                                const double d = (h & 1) ? dsum : ddif;  // harmonic k = h + 1
                                c[h] = __fma_rn(tc[idx], d, c[h]);
                                sn[h] = __fma_rn(ts[idx], d, sn[h]);
                            }
                    }
                }
            } else {
                for (int b = 0; b < nbin; b++) {
                    const uint32_t ad = cb[r] + ((uint32_t)b << 7);
                    const uint32_t v = lds_u32(ad);
                    if (v) {
                        sts_u32(ad, 0u);
                        const double d = (double)v;
                        cnt += d;
#pragma unroll
                        for (int h = 0; h < QL_NHMAX; h++)
                            if (h < nh) {
                                const int idx = ((h + 1) * b) & (nbin - 1);
                                c[h] = __fma_rn(tc[idx], d, c[h]);
                                sn[h] = __fma_rn(ts[idx], d, sn[h]);
                            }
                    }
                }
            }
            const long long cc = (slot0 + r) * 32 + lane;
            if (cc < a.n_c && cnt != 0.0) {
                double *dst = a.coef + ((size_t)cc * nprof + row) * nh * 2;
#pragma unroll
                for (int h = 0; h < QL_NHMAX; h++)
                    if (h < nh) atomicAdd(dst + 2 * h, c[h]), atomicAdd(dst + 2 * h + 1, sn[h]);
                atomicAdd(a.rowcnt + (size_t)cc * nprof + row, (unsigned long long)cnt);
            }
        }
    };
    auto one = [&](double t) {
        double h = 0.0, s6 = 0.0;
        if (MODE >= 1) {
            const double tssq = __dmul_rn(t, t);
            h = __dmul_rn(0.5, tssq);
            if (MODE >= 2) s6 = __dmul_rn(__dmul_rn(1.0 / 6, tssq), t);
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            double x = __dmul_rn(t, fN[r]);
            if (MODE >= 1) x = __dadd_rn(x, __dmul_rn(h, fdN[r]));
            if (MODE >= 2) x = __dadd_rn(x, __dmul_rn(s6, fddN[r]));
            const uint32_t lo = (uint32_t)__double2loint(__dadd_rd(x, HX_MAGIC));
            if (PACKED) red_shared_add_u32((lo & mask) | cb[r], (lo & topbit) | 1u);
            else red_shared_inc((lo & mask) | cb[r]);
        }
    };
    // bin address (and packed increment) of event t for centre r
    auto addr = [&](double t, int r, uint32_t &inc) -> uint32_t {
        double x = __dmul_rn(t, fN[r]);
        if (MODE >= 1) {
            const double tssq = __dmul_rn(t, t);
            x = __dadd_rn(x, __dmul_rn(__dmul_rn(0.5, tssq), fdN[r]));
            if (MODE >= 2) x = __dadd_rn(x, __dmul_rn(__dmul_rn(__dmul_rn(1.0 / 6, tssq), t), fddN[r]));
        }
        const uint32_t lo = (uint32_t)__double2loint(__dadd_rd(x, HX_MAGIC));
        inc = PACKED ? ((lo & topbit) | 1u) : 1u;
        return (lo & mask) | cb[r];
    };
    // plain events [b, e) of a stage.  K1's skewed software pipeline: the bin addresses of group g + 1 are computed (FP64
    // pipe) while the increments of group g are issued (shared-memory pipe), so the two interleave finely.
    auto span = [&](const unsigned char *tp, int b, int e) {
        int i = b;
        if ((i & 1) && i < e) one(*reinterpret_cast<const double *>(tp + (size_t)i * 8)), i++;
        // measured on B200 (profiles/r02_sweep_qlane_fused_skew.txt): the skew pays for R <= 2 (+10 % at nbin = 64, +6 % at
        // nbin = 128 R = 2) and costs registers for R = 4 (-20 % at nbin = 16), which keeps the straight loop
        constexpr bool SKEW = R <= 2;
        const int ngroups = SKEW ? (e - i) / UE : 0;
        if (!SKEW && i + UE <= e) {
            double2 v0 = *reinterpret_cast<const double2 *>(tp + (size_t)i * 8);
            double2 v1 = *reinterpret_cast<const double2 *>(tp + (size_t)i * 8 + 16);
            for (; i + UE <= e; i += UE) {
                // next group first (may read the 16-event pad after the stage; unused then)
                const double2 n0 = *reinterpret_cast<const double2 *>(tp + (size_t)i * 8 + 32);
                const double2 n1 = *reinterpret_cast<const double2 *>(tp + (size_t)i * 8 + 48);
                one(v0.x), one(v0.y), one(v1.x), one(v1.y);
                v0 = n0, v1 = n1;
            }
        }
        if (ngroups > 0) {
            uint32_t ad_cur[UE * R], inc_cur[UE * R], ad_nxt[UE * R], inc_nxt[UE * R];
            double tc[UE], tn[UE];
            auto load4 = [&](double *dst, int at) {  // may read the 16-event pad after the stage; unused then
                const double2 v0 = *reinterpret_cast<const double2 *>(tp + (size_t)at * 8);
                const double2 v1 = *reinterpret_cast<const double2 *>(tp + (size_t)at * 8 + 16);
                dst[0] = v0.x, dst[1] = v0.y, dst[2] = v1.x, dst[3] = v1.y;
            };
            load4(tc, i);
            load4(tn, i + UE);
#pragma unroll
            for (int k = 0; k < UE * R; k++) ad_cur[k] = addr(tc[k / R], k % R, inc_cur[k]);
            for (int g = 0; g < ngroups; g++) {
#pragma unroll
                for (int u = 0; u < UE; u++) tc[u] = tn[u];
                load4(tn, i + (g + 2) * UE);
#pragma unroll
                for (int k = 0; k < UE * R; k++) {
                    ad_nxt[k] = addr(tc[k / R], k % R, inc_nxt[k]);  // group g + 1 (garbage after the last group: never issued)
                    if (PACKED) red_shared_add_u32(ad_cur[k], inc_cur[k]);
                    else red_shared_inc(ad_cur[k]);
                }
#pragma unroll
                for (int k = 0; k < UE * R; k++) ad_cur[k] = ad_nxt[k], inc_cur[k] = inc_nxt[k];
            }
            i += ngroups * UE;
        }
        for (; i < e; i++) one(*reinterpret_cast<const double *>(tp + (size_t)i * 8));
    };

    int row = j0;                      // current row
    long long row_end = a.row_start[min(j0 + 1, nprof)];
    int since_drain = 0;
    for (int it = 0; it < ntiles; it++) {
        const int buf = it & 1;
        mbar_wait(bar[buf], (uint32_t)(it >> 1) & 1u);
        const long long s = es + (long long)it * TILE_EV;
        const int cnt = (int)min((long long)TILE_EV, ev_hi - s);
        const unsigned char *tp = smem + a.tile_off + (uint32_t)buf * TILE_BYTES;
        int e = (int)max(0LL, ev_lo - s);  // only the first stage can start before ev_lo
        while (e < cnt) {
            while (s + e >= row_end && row < j1) {  // finished row(s): drain and advance (warp-uniform)
                if (since_drain) {
                    if (COEF) accumulate(row);
                    else drain(row);
                    since_drain = 0;
                }
                row++;
                row_end = a.row_start[min(row + 1, nprof)];
            }
            int e_end = (int)min((long long)cnt, row_end - s);
            if (PACKED) e_end = min(e_end, e + a.flush_events - since_drain);
            span(tp, e, e_end);
            since_drain += e_end - e;
            e = e_end;
            if (PACKED && since_drain >= a.flush_events) {
                if (COEF) accumulate(row);
                else drain(row);
                since_drain = 0;
            }
        }
        __syncwarp();
        if (lane == 0) {
            const int prev = atomicAdd(&done[buf], 1);
            if (prev == nwarps - 1) {
                done[buf] = 0;
                if (it + 2 < ntiles) {
                    fence_proxy_async();
                    issue(it + 2);
                }
            }
        }
    }
    if (since_drain) {
        if (COEF) accumulate(min(row, nprof - 1));
        else drain(min(row, nprof - 1));
    }
}

// Re-bin the photons with |t| < t_safe: remove the fast-path bin, add the reference's bin (if any).
template <int MODE, bool COEF = false>
__global__ void __launch_bounds__(128) qffa_fix_kernel(const QLaneArgs a, const long long *__restrict__ urange,
                                                       double t_first, double delta0, double delta1) {
    const long long u0 = urange[0], u1 = urange[1];
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= a.n_c) return;
    const double f = a.f[c], fd = MODE >= 1 ? a.fd[c] : 0.0, fdd = MODE >= 2 ? a.fdd[c] : 0.0;
    const double fN = __dmul_rn(f, a.scale), fdN = __dmul_rn(fd, a.scale), fddN = __dmul_rn(fdd, a.scale);
    for (long long e = u0 + blockIdx.x; e < u1; e += gridDim.x) {
        const double t = a.t[e];
        const double jf = __dmul_rn(__dsub_rn(t, t_first), delta1);
        if (!(jf >= 0.0 && jf < (double)a.nprof)) continue;  // was never folded (beyond row_start[nprof])
        const int j = __double2loint(__dadd_rd(jf, HX_MAGIC));
        double h = 0.0, s6 = 0.0;
        if (MODE >= 1) {
            const double tssq = __dmul_rn(t, t);
            h = __dmul_rn(0.5, tssq);
            if (MODE >= 2) s6 = __dmul_rn(__dmul_rn(1.0 / 6, tssq), t);
        }
        double x = __dmul_rn(t, fN);
        if (MODE >= 1) x = __dadd_rn(x, __dmul_rn(h, fdN));
        if (MODE >= 2) x = __dadd_rn(x, __dmul_rn(s6, fddN));
        const int fast_bin = (int)(((uint32_t)__double2loint(__dadd_rd(x, HX_MAGIC)) >> 7) & (uint32_t)(a.nbin - 1));
        const int ref_bin = qffa_bin_generic<MODE>(t, h, s6, f, fd, fdd, delta0, a.nbin);
        if (ref_bin == fast_bin) continue;
        if (COEF) {  // move one count between bins in the harmonic domain: subtract the fast bin's rotation, add the right one
            double *dst = a.coef + ((size_t)c * a.nprof + j) * a.nharm * 2;
            for (int hh = 0; hh < a.nharm; hh++) {
                const int i_f = ((hh + 1) * fast_bin) & (a.nbin - 1);
                double dc = -a.tab[i_f], ds = -a.tab[a.nbin + i_f];
                if (ref_bin >= 0) {
                    const int i_r = ((hh + 1) * ref_bin) & (a.nbin - 1);
                    dc += a.tab[i_r], ds += a.tab[a.nbin + i_r];
                }
                atomicAdd(dst + 2 * hh, dc);
                atomicAdd(dst + 2 * hh + 1, ds);
            }
            if (ref_bin < 0) atomicAdd(a.rowcnt + (size_t)c * a.nprof + j, ~0ull);  // the reference drops this event
            continue;
        }
        uint32_t *p = a.prof + (((size_t)(c >> 5) * a.nprof + j) * a.nbin) * 32 + (c & 31);
        atomicSub(p + (size_t)fast_bin * 32, 1u);
        if (ref_bin >= 0) atomicAdd(p + (size_t)ref_bin * 32, 1u);
    }
}

// COEF path: events folded per centre = sum of its slices' counts
__global__ void qffa_tot_kernel(const unsigned long long *__restrict__ rowcnt, int n_c, int nprof,
                                unsigned long long *__restrict__ tot) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= n_c) return;
    unsigned long long sum = 0;
    for (int k = lane; k < nprof; k += 32) sum += rowcnt[(size_t)c * nprof + k];
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) tot[c] = sum;
}

// One pre-pass for the lane path: sortedness of the times (all threads), the first event of every time slice and the
// [u0, u1) range of events with |t| < t_safe (binary searches by the first nprof + 3 threads).
__global__ void qffa_prepass_kernel(const double *__restrict__ t, long long n, double t_first, double delta1, int nprof,
                                    double t_safe, long long *__restrict__ row_start, long long *__restrict__ urange,
                                    int *flag) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    for (long long i = gid; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = t[i];
        if (!(v >= t_first) || (i + 1 < n && t[i + 1] < v)) bad = true;
    }
    if (bad) atomicExch(flag, 1);
    if (row_start == nullptr) return;  // sortedness only (the boundaries came from the host: streamed upload)
    if (gid <= nprof) {
        const int j = (int)gid;  // smallest e with row(t[e]) >= j; row nprof = "dropped"
        long long lo = 0, hi = n;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            const double jf = __dmul_rn(__dsub_rn(t[mid], t_first), delta1);
            const bool before = (j == nprof) ? (jf < (double)nprof) : (jf < (double)j);
            if (before) lo = mid + 1;
            else hi = mid;
        }
        row_start[j] = lo;
    } else if (gid <= nprof + 2) {
        const bool first = gid == nprof + 1;  // u0: first t > -t_safe ; u1: first t >= t_safe
        const double bound = first ? -t_safe : t_safe;
        long long lo = 0, hi = n;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            const bool before = first ? (t[mid] <= bound) : (t[mid] < bound);
            if (before) lo = mid + 1;
            else hi = mid;
        }
        urange[first ? 0 : 1] = lo;
    }
}

// prof_lane [slot][row][bin][32] -> prof [c][row][bin]  (the layout of the public `profiles` output and of K3)
__global__ void qffa_unlane_kernel(const uint32_t *__restrict__ src, int n_c, int nprof, int nbin,
                                   uint32_t *__restrict__ dst) {
    const size_t n = (size_t)n_c * nprof * nbin;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / ((size_t)nprof * nbin), rb = i - c * ((size_t)nprof * nbin);
        dst[i] = src[(((c >> 5) * nprof * (size_t)nbin) + rb) * 32 + (c & 31)];
    }
}

// K3, harmonic-domain formulation.  The reference rolls every time slice k by s_k bins, sums the
// rolled slices (shift_and_sum, efsearch.py:350-363) and takes C_h = sum_b splat[b] cos(2 pi h b/nbin),
// S_h likewise (ffa.py:98-102).  A circular shift is a rotation of each harmonic:
//     sum_b row[(b - s) mod nbin] cos(h th_b) = Cr_h cos(h th_s) - Sr_h sin(h th_s)
//     sum_b row[(b - s) mod nbin] sin(h th_b) = Sr_h cos(h th_s) + Cr_h sin(h th_s)
// so the nbin-long adds per (slice, shift pair) collapse to nharm complex rotations of per-slice
// coefficients (Cr_h, Sr_h) computed once.  Same table values, same real numbers; only the fp64
// summation order differs from the reference (parity tolerance 1e-9, observed ~1e-14).

// (a) per (centre, slice): Cr_h, Sr_h for h = 1..nharm, and the total counts of each centre
__global__ void __launch_bounds__(256) qffa_rowdft_kernel(const uint32_t *__restrict__ prof, int nprof, int nbin,
                                                          int nharm, const double *__restrict__ tab,
                                                          double *__restrict__ coef /*[C][nprof][nharm][2]*/,
                                                          unsigned long long *__restrict__ tot /*[C]*/,
                                                          unsigned long long *__restrict__ rowcnt /*[C][nprof] or null*/) {
    extern __shared__ double tsm[];  // cos | sin
    for (int i = threadIdx.x; i < 2 * nbin; i += blockDim.x) tsm[i] = tab[i];
    __syncthreads();
    const int c = blockIdx.y, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nprof) return;
    const uint32_t *p = prof + ((size_t)c * nprof + row) * nbin;
    const bool p2 = (nbin & (nbin - 1)) == 0;
    unsigned long long cnt = 0;
    for (int b = lane; b < nbin; b += 32) cnt += p[b];
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0 && cnt) atomicAdd(&tot[c], cnt);
    if (lane == 0 && rowcnt) rowcnt[(size_t)c * nprof + row] = cnt;
    for (int h = 1; h <= nharm; h++) {
        double cr = 0.0, sr = 0.0;
        for (int b = lane; b < nbin; b += 32) {
            const int idx = p2 ? ((b * h) & (nbin - 1)) : (int)(((long long)b * h) % nbin);
            const double v = (double)p[b];
            cr += tsm[idx] * v;
            sr += tsm[nbin + idx] * v;
        }
        for (int o = 16; o; o >>= 1) {
            cr += __shfl_xor_sync(0xffffffffu, cr, o);
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
        }
        if (lane == 0) {
            double *d = coef + (((size_t)c * nprof + row) * nharm + (h - 1)) * 2;
            d[0] = cr, d[1] = sr;
        }
    }
}

// (b) one CTA per (centre, (L, Q) pair): rotate and sum the slice coefficients, reduce to Z^2_n
__global__ void __launch_bounds__(256) qffa_step_kernel(const double *__restrict__ coef,
                                                        const unsigned long long *__restrict__ tot, int nprof,
                                                        int nbin, const double *__restrict__ L,
                                                        const double *__restrict__ Q, int n_lq,
                                                        const double *__restrict__ base_shift,
                                                        const double *__restrict__ tab, int nharm,
                                                        double *__restrict__ stats) {
    extern __shared__ double ssm[];  // cos | sin | reduction scratch [2*8]
    double *red = ssm + 2 * nbin;
    for (int i = threadIdx.x; i < 2 * nbin; i += blockDim.x) ssm[i] = tab[i];
    __syncthreads();
    const int c = blockIdx.y, lq = blockIdx.x;
    const double l = L[lq], q = Q[lq];
    const bool p2 = (nbin & (nbin - 1)) == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double res = 0.0;
    for (int h = 1; h <= nharm; h++) {
        double cs = 0.0, sn = 0.0;
        for (int k = threadIdx.x; k < nprof; k += blockDim.x) {
            const double b = base_shift[k];
            const double tot_shift = __dadd_rn(__dmul_rn(b, l), __dmul_rn(__dmul_rn(b, b), q));
            const double r = __dsub_rn(__dadd_rn(tot_shift, HX_MAGIC), HX_MAGIC);  // np.rint: half to even
            int s = (int)r;
            s = p2 ? (s & (nbin - 1)) : ((s % nbin) + nbin) % nbin;  // np.mod -> [0, nbin)
            const int idx = p2 ? ((s * h) & (nbin - 1)) : (int)(((long long)s * h) % nbin);
            const double ct = ssm[idx], st = ssm[nbin + idx];
            const double *d = coef + (((size_t)c * nprof + k) * nharm + (h - 1)) * 2;
            const double cr = d[0], sr = d[1];
            cs += cr * ct - sr * st;
            sn += sr * ct + cr * st;
        }
        for (int o = 16; o; o >>= 1) {
            cs += __shfl_xor_sync(0xffffffffu, cs, o);
            sn += __shfl_xor_sync(0xffffffffu, sn, o);
        }
        __syncthreads();
        if (lane == 0) red[warp] = cs, red[8 + warp] = sn;
        __syncthreads();
        double ct = 0.0, st = 0.0;
        for (int w = 0; w < nw; w++) ct += red[w], st += red[8 + w];
        res += ct * ct + st * st;
    }
    if (threadIdx.x == 0) stats[(size_t)c * n_lq + lq] = 2.0 / (double)tot[c] * res;
}


// (a') the same coefficients straight from the lane layout prof[slot][row][bin][32]: one warp per
// (slot, row), lane = centre, so every load is one coalesced 128-byte line and nothing is reduced
// across lanes.  Replaces qffa_unlane_kernel + qffa_rowdft_kernel on the lane = centre path.
__global__ void __launch_bounds__(256) qffa_rowdft_lane_kernel(const uint32_t *__restrict__ prof, int n_c, int nprof,
                                                               int nbin, int nharm, const double *__restrict__ tab,
                                                               double *__restrict__ coef /*[C][nprof][nharm][2]*/,
                                                               unsigned long long *__restrict__ tot /*[C]*/,
                                                               unsigned long long *__restrict__ rowcnt) {
    extern __shared__ double tsm[];  // cos | sin
    for (int i = threadIdx.x; i < 2 * nbin; i += blockDim.x) tsm[i] = tab[i];
    __syncthreads();
    const int slot = blockIdx.y, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int c = slot * 32 + lane;
    if (row >= nprof) return;
    const uint32_t *p = prof + (((size_t)slot * nprof + row) * nbin) * 32 + lane;
    const bool p2 = (nbin & (nbin - 1)) == 0;
    unsigned long long cnt = 0;
    for (int b = 0; b < nbin; b++) cnt += p[(size_t)b * 32];
    if (c < n_c) {
        if (cnt) atomicAdd(&tot[c], cnt);
        if (rowcnt) rowcnt[(size_t)c * nprof + row] = cnt;
    }
    for (int h = 1; h <= nharm; h++) {
        double cr = 0.0, sr = 0.0;
        int idx = 0;
        for (int b = 0; b < nbin; b++) {
            const double v = (double)p[(size_t)b * 32];
            cr += tsm[idx] * v;
            sr += tsm[nbin + idx] * v;
            idx += h;
            if (p2) idx &= nbin - 1;
            else idx %= nbin;
        }
        if (c < n_c) {
            double *d = coef + (((size_t)c * nprof + row) * nharm + (h - 1)) * 2;
            d[0] = cr, d[1] = sr;
        }
    }
}

// (b') thread = (L, Q) pair: no reduction over threads at all for up to 64 pairs per CTA; the remaining
// threads of the CTA split the time slices.  NH harmonics are accumulated in registers.
template <int NH>
__global__ void __launch_bounds__(256) qffa_step2_kernel(const double *__restrict__ coef,
                                                         const unsigned long long *__restrict__ tot, int nprof,
                                                         int nbin, const double *__restrict__ L,
                                                         const double *__restrict__ Q, int n_lq, int pairs_per_cta,
                                                         const double *__restrict__ base_shift,
                                                         const double *__restrict__ tab, double *__restrict__ stats) {
    extern __shared__ __align__(16) double ssm[];  // (cos, sin) pairs [nbin] | partial sums [groups][pairs][2*NH]
    double2 *rot = reinterpret_cast<double2 *>(ssm);
    double *part = ssm + 2 * nbin;
    for (int i = threadIdx.x; i < nbin; i += blockDim.x) rot[i] = make_double2(tab[i], tab[nbin + i]);
    __syncthreads();
    const int c = blockIdx.y;
    const int groups = blockDim.x / pairs_per_cta;
    const int pl = threadIdx.x % pairs_per_cta, g = threadIdx.x / pairs_per_cta;
    const int lq = blockIdx.x * pairs_per_cta + pl;
    const bool active = lq < n_lq && g < groups;
    const bool p2 = (nbin & (nbin - 1)) == 0;
    double cs[NH], sn[NH];
#pragma unroll
    for (int h = 0; h < NH; h++) cs[h] = 0.0, sn[h] = 0.0;
    if (active) {
        const double l = L[lq], q = Q[lq];
        for (int k = g; k < nprof; k += groups) {
            const double b = base_shift[k];
            const double tot_shift = __dadd_rn(__dmul_rn(b, l), __dmul_rn(__dmul_rn(b, b), q));
            const double r = __dsub_rn(__dadd_rn(tot_shift, HX_MAGIC), HX_MAGIC);  // np.rint: half to even
            int s = (int)r;
            s = p2 ? (s & (nbin - 1)) : ((s % nbin) + nbin) % nbin;  // np.mod -> [0, nbin)
            const double2 *d = reinterpret_cast<const double2 *>(coef + ((size_t)c * nprof + k) * NH * 2);
            // the shift is per thread, so this lookup is a random (bank-conflicted) access: ONE 16-byte load for the
            // first harmonic's rotation; the rotations of harmonics 2..NH follow by angle addition (4 FP64 each)
            // instead of 2 more conflicted loads each -- the kernel was bound by those loads
            const double2 w1 = rot[s];
            double ct = w1.x, st = w1.y;
#pragma unroll
            for (int h = 0; h < NH; h++) {
                if (h > 0) {
                    const double cn = __fma_rn(ct, w1.x, -__dmul_rn(st, w1.y));
                    st = __fma_rn(st, w1.x, __dmul_rn(ct, w1.y));
                    ct = cn;
                }
                const double2 v = d[h];  // (Cr, Sr): the same address for every pair of a group: broadcast
                cs[h] += v.x * ct - v.y * st;
                sn[h] += v.y * ct + v.x * st;
            }
        }
    }
    if (g < groups) {
#pragma unroll
        for (int h = 0; h < NH; h++) {
            part[((size_t)g * pairs_per_cta + pl) * 2 * NH + 2 * h] = cs[h];
            part[((size_t)g * pairs_per_cta + pl) * 2 * NH + 2 * h + 1] = sn[h];
        }
    }
    __syncthreads();
    if (g == 0 && lq < n_lq) {
        double res = 0.0;
#pragma unroll
        for (int h = 0; h < NH; h++) {
            double a = 0.0, bsum = 0.0;
            for (int gg = 0; gg < groups; gg++) {
                a += part[((size_t)gg * pairs_per_cta + pl) * 2 * NH + 2 * h];
                bsum += part[((size_t)gg * pairs_per_cta + pl) * 2 * NH + 2 * h + 1];
            }
            res += a * a + bsum * bsum;
        }
        stats[(size_t)c * n_lq + lq] = 2.0 / (double)tot[c] * res;
    }
}

// Transient search (hendrics/efsearch.py:407-465, _average_and_z_sub_search): for every averaging level
// l (groups of 2^l consecutive time slices) the Z^2_n of the summed sub-profiles, written to all rows of
// the group.  Sums of profiles are sums of their harmonic coefficients, so no profile is touched again.
__global__ void __launch_bounds__(128) qffa_transient_kernel(const double *__restrict__ coef,
                                                             const unsigned long long *__restrict__ rowcnt,
                                                             int nprof, int nrows, int nharm, int n_levels,
                                                             double *__restrict__ results /*[C][n_levels][nrows]*/) {
    // nrows <= nprof: the reference keeps `profiles[:int(2 ** np.log2(nprof))]`, which drops the last time slice for
    // about 40 % of the non power-of-two nprof (2 ** log2(10) = 9.999...); coef / rowcnt keep their nprof stride.
    const int c = blockIdx.z, level = blockIdx.y;
    const int n_ave = 1 << level;
    const int groups = nrows / n_ave;  // int(profiles.shape[0] / n_ave_i)
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    double *out = results + ((size_t)c * n_levels + level) * nrows;
    if (g >= groups) {
        // rows not covered by any group keep the reference's initial 0
        if (g == groups)
            for (int r = groups * n_ave; r < nrows; r++) out[r] = 0.0;
        return;
    }
    unsigned long long cnt = 0;
    for (int r = g * n_ave; r < (g + 1) * n_ave; r++) cnt += rowcnt[(size_t)c * nprof + r];
    double z = 0.0;
    if (cnt != 0) {  // `if np.sum(new_profiles) == 0: continue`
        double res = 0.0;
        for (int h = 0; h < nharm; h++) {
            double cs = 0.0, sn = 0.0;
            for (int r = g * n_ave; r < (g + 1) * n_ave; r++) {
                const double *d = coef + (((size_t)c * nprof + r) * nharm + h) * 2;
                cs += d[0], sn += d[1];
            }
            res += cs * cs + sn * sn;
        }
        z = 2.0 / (double)cnt * res;
    }
    for (int r = g * n_ave; r < (g + 1) * n_ave; r++) out[r] = z;
}

typedef void (*qhist_fn_t)(const QffaHistArgs);
template <int C, bool FAST>
static qhist_fn_t pick_qhist_c(int mode) {
    return mode == 0 ? qffa_hist_kernel<C, 0, FAST> : mode == 1 ? qffa_hist_kernel<C, 1, FAST> : qffa_hist_kernel<C, 2, FAST>;
}
template <bool FAST>
static qhist_fn_t pick_qhist_f(int C, int mode) {
    switch (C) {
        case 1: return pick_qhist_c<1, FAST>(mode);
        case 2: return pick_qhist_c<2, FAST>(mode);
        case 3: return pick_qhist_c<3, FAST>(mode);
        case 4: return pick_qhist_c<4, FAST>(mode);
        case 5: return pick_qhist_c<5, FAST>(mode);
        case 6: return pick_qhist_c<6, FAST>(mode);
        case 7: return pick_qhist_c<7, FAST>(mode);
        default: return pick_qhist_c<8, FAST>(mode);
    }
}
static qhist_fn_t pick_qhist(int C, int mode, bool fast) {
    return fast ? pick_qhist_f<true>(C, mode) : pick_qhist_f<false>(C, mode);
}

static int g_qffa_cmax = QF_CMAX;  // tuning hooks
static int g_qffa_fast = 1;
static int g_qffa_logk = -1;
// CUDA events around the K2 launches of the last call (bench.py's live roofline measurement)
constexpr int QF_EVPOOL = 256;
static cudaEvent_t g_qev[2 * QF_EVPOOL];
static bool g_qev_ok = false;
static int g_qev_used = 0;
static long long g_qev_batches = 0;
static double g_q_last_ms = -1.0;


// ---- K2-lane host path -----------------------------------------------------------------------
typedef void (*qlane_fn_t)(const QLaneArgs);
template <int R, bool PACKED, bool COEF>
static qlane_fn_t pick_qlane_m(int mode) {
    return mode == 0 ? qffa_lane_kernel<R, 0, PACKED, COEF> : mode == 1 ? qffa_lane_kernel<R, 1, PACKED, COEF>
                                                                      : qffa_lane_kernel<R, 2, PACKED, COEF>;
}
template <bool COEF>
static qlane_fn_t pick_qlane_c(int R, int mode, bool packed) {
    if (packed) return R == 4 ? pick_qlane_m<4, true, COEF>(mode) : R == 2 ? pick_qlane_m<2, true, COEF>(mode) : pick_qlane_m<1, true, COEF>(mode);
    return R == 4 ? pick_qlane_m<4, false, COEF>(mode) : R == 2 ? pick_qlane_m<2, false, COEF>(mode) : pick_qlane_m<1, false, COEF>(mode);
}
static qlane_fn_t pick_qlane(int R, int mode, bool packed, bool coef) {
    return coef ? pick_qlane_c<true>(R, mode, packed) : pick_qlane_c<false>(R, mode, packed);
}

static int g_qffa_lane = 1;  // tuning hook: 0 forces the lane = event kernel

constexpr int QL_TILE = 1024;

// returns HX_OK, an error (< 0), or 1 when the inputs are not eligible (caller falls back to the lane = event path)
static int qffa_lane_path(const DeviceInfo &di, const double *d_t, int64_t n_ev, double t_first, double t_last,
                          const double *mean_f, const double *mean_fd, const double *mean_fdd, int64_t n_centres,
                          int nbin, int nprof, double delta0, double delta1, const double *d_L, const double *d_Q,
                          int64_t n_lq, const double *d_base, const double *d_tab, int nharm, double *d_stats,
                          uint32_t *d_profiles, cudaStream_t st, int transient_levels, int transient_rows,
                          const double *h_t = nullptr, bool *uploaded = nullptr) {
    // h_t != NULL: the events are still in HOST memory (hx_qffa_search); d_t is an empty device buffer that this function
    // fills -- streamed in chunks WHILE K2 folds the chunks that have arrived, when the fused path applies -- and
    // *uploaded tells the caller whether that happened (a `return 1` before it leaves the upload to the caller).
    const bool p2 = (nbin & (nbin - 1)) == 0;
    if (!g_qffa_lane || !g_qffa_fast || !p2 || nbin < 8 || nbin > 512 || n_centres < 1) return 1;
    int mode = 0;
    double fmin = INFINITY, fmax = 0.0, fdabs = 0.0, fddabs = 0.0;
    for (int64_t c = 0; c < n_centres; c++) {
        if (mean_fdd[c] != 0) mode = 2;
        else if (mean_fd[c] != 0) mode = std::max(mode, 1);
        fmin = std::min(fmin, mean_f[c]);
        fmax = std::max(fmax, std::fabs(mean_f[c]));
        fdabs = std::max(fdabs, std::fabs(mean_fd[c]));
        fddabs = std::max(fddabs, std::fabs(mean_fdd[c]));
    }
    if (!(fmin > 0) || !std::isfinite(fmin)) return 1;
    const double tmax = std::max(std::fabs(t_first), std::fabs(t_last));
    const double t_safe = 1.0 / fmin;
    auto g = [&](double x) { return x * fmin - 0.5 * x * x * fdabs - x * x * x * fddabs / 6; };
    const double scale = (double)nbin * 128.0;
    const double phmax = tmax * fmax + 0.5 * tmax * tmax * fdabs + tmax * tmax * tmax * fddabs / 6;
    if (!(g(t_safe) >= 0.75 && (tmax <= t_safe || g(tmax) >= 0.75) && phmax * scale < 1125899906842624.0)) return 1;

    // ---- once per call: sortedness, row boundaries, the |t| < t_safe range, centre constants on the device
    void *d_misc = nullptr, *d_cent = nullptr, *d_lane = nullptr, *d_plain = nullptr, *d_coef = nullptr;
    struct Free {
        void *&p;
        cudaStream_t s;
        ~Free() {
            if (p) cudaFreeAsync(p, s);
        }
    } f1{d_misc, st}, f2{d_cent, st}, f3{d_lane, st}, f4{d_plain, st}, f5{d_coef, st};
    const size_t misc_bytes = (size_t)(nprof + 1 + 2) * 8 + 16;  // row_start | urange | flags[2] | upload progress
    HX_CUDA_TRY(cudaMallocAsync(&d_misc, misc_bytes, st));
    HX_CUDA_TRY(cudaMemsetAsync(d_misc, 0, misc_bytes, st));
    long long *d_rowstart = reinterpret_cast<long long *>(d_misc), *d_urange = d_rowstart + nprof + 1;
    int *d_flags = reinterpret_cast<int *>(d_urange + 2);  // [0] unsorted, [1] layout error
    // streamed upload: fused single-batch searches on large event lists
    bool streamed = false;
    if (h_t) {
        bool ok = d_profiles == nullptr && nharm >= 1 && nharm <= QL_NHMAX && n_ev >= (1 << 22) &&
                  (size_t)n_centres * nprof * (2 * nharm + 1) * 8 <= ((size_t)2048 << 20);
        if (const char *e = getenv("HX_QFFA_FUSE")) ok = ok && atoi(e) != 0;
        if (const char *e = getenv("HX_QFFA_STREAM")) ok = ok && atoi(e) != 0;  // A-B hook
        streamed = ok;
        if (!streamed) {  // plain upload, then everything as for resident events
            HX_CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(d_t), h_t, (size_t)n_ev * 8, cudaMemcpyHostToDevice, st));
            if (uploaded) *uploaded = true;
        }
    }
    cudaStream_t cs = nullptr;  // copy stream of the streamed upload
    cudaEvent_t ev_sync = nullptr;
    struct StreamGuard {
        cudaStream_t &s;
        cudaEvent_t &e;
        ~StreamGuard() {
            if (s) cudaStreamSynchronize(s);  // no copy into the caller's buffer may outlive this call
            if (e) cudaEventDestroy(e);
            if (s) cudaStreamDestroy(s);
        }
    } stream_guard{cs, ev_sync};
    long long chunk_events = 0;
    int n_chunks = 0;
    std::vector<int> h_prog;
    int *d_prog = nullptr;
    bool copies_issued = false;
    auto issue_copies = [&]() -> int {  // chunk c, then "c + 1 chunks are in place", on the copy stream
        for (int c = 0; c < n_chunks; c++) {
            const long long off = (long long)c * chunk_events, cnt = std::min<long long>(chunk_events, n_ev - off);
            HX_CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(d_t) + off, h_t + off, (size_t)cnt * 8, cudaMemcpyHostToDevice, cs));
            HX_CUDA_TRY(cudaMemcpyAsync(d_prog, &h_prog[c], sizeof(int), cudaMemcpyHostToDevice, cs));
        }
        HX_CUDA_TRY(cudaEventRecord(ev_sync, cs));
        copies_issued = true;
        return HX_OK;
    };
    bool pinned = false;
    int h_flags[2] = {0, 0};
    long long h_urange[2] = {0, 0};
    if (!streamed) {
        const long long want = std::max<long long>((n_ev + 255) / 256, (nprof + 3 + 255) / 256);
        qffa_prepass_kernel<<<(unsigned)std::min<long long>(std::max<long long>(want, 1), 2368), 256, 0, st>>>(
            d_t, n_ev, t_first, delta1, nprof, t_safe, d_rowstart, d_urange, d_flags);
        HX_CHECK_LAUNCH("qffa_prepass_kernel");
        HX_CUDA_TRY(cudaMemcpyAsync(h_flags, d_flags, 8, cudaMemcpyDeviceToHost, st));
        HX_CUDA_TRY(cudaMemcpyAsync(h_urange, d_urange, 16, cudaMemcpyDeviceToHost, st));
        HX_CUDA_TRY(cudaStreamSynchronize(st));
        if (h_flags[0]) return 1;  // unsorted input: the lane = event kernel copes with it
    } else {
        if (uploaded) *uploaded = true;
        HX_CUDA_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        HX_CUDA_TRY(cudaEventCreateWithFlags(&ev_sync, cudaEventDisableTiming));
        n_chunks = (int)std::min<long long>(16, std::max<long long>(2, n_ev / (1 << 22)));  // >= 32 MB per chunk
        chunk_events = ((n_ev + n_chunks - 1) / n_chunks + 1) & ~1LL;
        n_chunks = (int)((n_ev + chunk_events - 1) / chunk_events);
        h_prog.resize(n_chunks);
        for (int c = 0; c < n_chunks; c++) h_prog[c] = c + 1;
        d_prog = d_flags + 2;  // zeroed by the memset of d_misc above
        // the copies may only start once that memset has run
        HX_CUDA_TRY(cudaEventRecord(ev_sync, st));
        HX_CUDA_TRY(cudaStreamWaitEvent(cs, ev_sync, 0));
        cudaPointerAttributes pa{};
        pinned = cudaPointerGetAttributes(&pa, h_t) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();  // an unregistered host pointer is not an error here
        // page-locked source: the copies are asynchronous, start them now and search the slice boundaries meanwhile;
        // pageable source: cudaMemcpyAsync blocks the host, so K2 is launched first (below) and then fed
        if (pinned) {
            int rcc = issue_copies();
            if (rcc) return rcc;
        }
        // slice boundaries and the |t| < t_safe range by binary search on the HOST array, with the kernels' arithmetic
        // (assumes sorted times; sortedness is verified on the device once everything is up, see below)
        std::vector<long long> rs((size_t)nprof + 3);
        auto before_row = [&](long long e, int j) {
            const double jf = (h_t[e] - t_first) * delta1;
            return j == nprof ? jf < (double)nprof : jf < (double)j;
        };
        long long prev = 0;
        for (int j = 0; j <= nprof; j++) {  // galloping from the previous boundary: the probes stay local
            long long lo = prev, hi = n_ev, stepg = 1024;
            while (lo + stepg < n_ev && before_row(lo + stepg, j)) lo += stepg, stepg *= 2;
            hi = std::min<long long>(n_ev, lo + stepg);
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (before_row(mid, j)) lo = mid + 1;
                else hi = mid;
            }
            rs[j] = prev = lo;
        }
        for (int k = 0; k < 2; k++) {  // u0: first t > -t_safe ; u1: first t >= t_safe
            long long lo = 0, hi = n_ev;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                const bool before = k == 0 ? (h_t[mid] <= -t_safe) : (h_t[mid] < t_safe);
                if (before) lo = mid + 1;
                else hi = mid;
            }
            rs[(size_t)nprof + 1 + k] = h_urange[k] = lo;
        }
        HX_CUDA_TRY(cudaMemcpyAsync(d_rowstart, rs.data(), rs.size() * 8, cudaMemcpyHostToDevice, st));
        HX_CUDA_TRY(cudaStreamSynchronize(st));  // rs is a local: the copy must have read it before it goes away
    }
    const long long n_unsafe = std::max<long long>(0, h_urange[1] - h_urange[0]);

    std::vector<double> hc(3 * (size_t)n_centres);
    for (int64_t c = 0; c < n_centres; c++) hc[c] = mean_f[c], hc[n_centres + c] = mean_fd[c], hc[2 * n_centres + c] = mean_fdd[c];
    HX_CUDA_TRY(cudaMallocAsync(&d_cent, hc.size() * 8, st));
    HX_CUDA_TRY(cudaMemcpyAsync(d_cent, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice, st));
    const double *d_f = reinterpret_cast<double *>(d_cent), *d_fd = d_f + n_centres, *d_fdd = d_fd + n_centres;

    // ---- kernel configuration (same footprint classes as K1)
    const bool packed = nbin >= 128;
    const int foot = packed ? nbin / 2 : nbin;
    // (centres per thread, warps per CTA) measured on B200 (tools/sweep_qlane.py, profiles/r01_sweep_qlane.txt)
    int R = foot <= 32 ? 4 : foot <= 64 ? 2 : 1;
    if (mode >= 1 && R == 4 && foot > 16) R = 2;
    int W = foot <= 16 ? 2 : 4;
    if (packed && foot == 64) R = 2, W = 4;  // nbin = 128 (the --fast default): 12 resident warps, half the event loads per unit
    // batches of equal size (multiples of 32 centres), bounded by 2 GiB of profile scratch (the whole C4 search -- 1252
    // centres, 1.25 GB -- is one batch: 38.2 -> 36.2 ms per search against three 512 MiB batches, profiles/r01_sweep_qbatch.txt)
    // Fused path (no `profiles` output, nharm <= 4): a finished time slice is reduced in the SM to its 2 * nharm harmonic
    // coefficients; the [slot][row][bin][32] count scratch (1 MB per centre at nbin = 128, nprof = 2048) does not exist.
    bool coef_mode = d_profiles == nullptr && nharm >= 1 && nharm <= QL_NHMAX;
    if (const char *e = getenv("HX_QFFA_FUSE")) coef_mode = coef_mode && atoi(e) != 0;  // tuning / A-B hook
    const size_t per_centre = coef_mode ? (size_t)nprof * (2 * nharm + 1) * 8 : (size_t)nprof * nbin * 4;
    size_t batch_bytes = (size_t)2048 << 20;
    if (const char *e = getenv("HX_QFFA_BATCH_MIB")) {  // tuning hook
        const long long mib = atoll(e);
        if (mib > 0) batch_bytes = (size_t)mib << 20;
    }
    const int64_t cap = std::max<int64_t>(32, (int64_t)(batch_bytes / per_centre) / 32 * 32);
    const int64_t n_batches = (n_centres + cap - 1) / cap;
    const int cb_max = (int)std::min<int64_t>(((n_centres + n_batches - 1) / n_batches + 31) / 32 * 32, n_centres + 31);
    const int slots_max = (cb_max + 31) / 32;
    while (R > 1 && R > slots_max) R /= 2;  // small batches: no idle trial slots
    {
        // CTA columns: as few as possible, each with just enough warps for its share of the slots
        const int gx = (slots_max + R * W - 1) / (R * W);
        W = std::max(1, (slots_max + R * gx - 1) / (R * gx));
        if (W == 3) W = 4;  // 1, 2 or a multiple of 4 warps: 3- and 6-warp CTAs load the SM's four schedulers unevenly
        else if (W > 4) W = (W + 3) / 4 * 4;
    }
    if (const char *e = getenv("HX_QLANE_RW")) {  // tuning hook: "R,W"
        int r = 0, w = 0;
        if (sscanf(e, "%d,%d", &r, &w) == 2 && (r == 1 || r == 2 || r == 4) && w >= 1 && w <= 16) R = r, W = w;
    }
    const uint32_t blk = (uint32_t)nbin * (packed ? 64u : 128u);
    uint32_t hist_off, tile_off, total, extra_off = 0;
    int tile_ev = QL_TILE;
    for (;;) {
        const uint32_t hist_bytes = (uint32_t)W * R * blk;
        const uint32_t hist_abs = (di.dyn_smem_base + 32 + blk - 1) / blk * blk;
        hist_off = hist_abs - di.dyn_smem_base;
        const uint32_t gap = hist_off - 32;
        // largest stage that fits in the alignment gap in front of the histograms (fused path: together with the tables)
        const uint32_t tab_bytes = coef_mode ? 2u * (uint32_t)nbin * 8u : 0u;
        int fit = 0;
        for (int t = QL_TILE; t >= 256; t -= 64)
            if (2u * (uint32_t)(t + 16) * 8u + tab_bytes <= gap) {
                fit = t;
                break;
            }
        tile_ev = fit ? fit : QL_TILE;
        const uint32_t tile_region = 2u * (uint32_t)(tile_ev + 16) * 8u;
        tile_off = fit ? 32u : hist_off + hist_bytes;
        total = std::max(hist_off + hist_bytes, tile_off + tile_region + tab_bytes);
        extra_off = tile_off + tile_region;  // the tables sit right behind the stages
        if (total <= (uint32_t)di.smem_optin) break;
        if (W > 1) W /= 2;
        else if (R > 1) R /= 2;
        else return 1;
    }
    const int ctas_per_sm = std::max(1, std::min(8, (int)(di.smem_per_sm / (total + di.reserved_smem))));
    qlane_fn_t fn = pick_qlane(R, mode, packed, coef_mode);
    HX_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));

    const size_t lane_elems = (size_t)slots_max * nprof * nbin * 32;
    const size_t prof_elems = (size_t)nprof * nbin;
    if (!coef_mode) HX_CUDA_TRY(cudaMallocAsync(&d_lane, lane_elems * 4, st));
    const size_t coef_elems = (size_t)cb_max * nprof * nharm * 2;
    HX_CUDA_TRY(cudaMallocAsync(&d_coef, coef_elems * 8 + (size_t)cb_max * 8 + (size_t)cb_max * nprof * 8, st));
    unsigned long long *d_tot = reinterpret_cast<unsigned long long *>(reinterpret_cast<double *>(d_coef) + coef_elems);
    unsigned long long *d_rowcnt = d_tot + cb_max;
    const size_t step_smem = (2 * (size_t)nbin + 16) * sizeof(double);

    g_qev_used = 0, g_qev_batches = 0, g_q_last_ms = -1.0;
    for (int64_t c0 = 0; c0 < n_centres; c0 += cb_max) {
        const int C = (int)std::min<int64_t>(cb_max, n_centres - c0);
        const int slots = (C + 31) / 32;
        if (coef_mode) HX_CUDA_TRY(cudaMemsetAsync(d_coef, 0, coef_elems * 8 + (size_t)cb_max * 8 + (size_t)cb_max * nprof * 8, st));
        else HX_CUDA_TRY(cudaMemsetAsync(d_lane, 0, (size_t)slots * nprof * nbin * 32 * 4, st));
        QLaneArgs a{};
        a.t = d_t, a.n_ev = n_ev, a.row_start = d_rowstart;
        a.f = d_f + c0, a.fd = d_fd + c0, a.fdd = d_fdd + c0, a.n_c = C;
        a.nbin = nbin, a.nprof = nprof, a.scale = scale;
        int log2nbin = 0;
        while ((1 << log2nbin) < nbin) log2nbin++;
        a.pack_shift = 6 + log2nbin, a.topbit = 1u << a.pack_shift;
        a.binmask = packed ? (uint32_t)(nbin / 2 - 1) << 7 : (uint32_t)(nbin - 1) << 7;
        a.flush_events = (int)((1u << a.pack_shift) - 1u);
        a.prof = reinterpret_cast<uint32_t *>(d_lane);
        a.warps = W, a.hist_off = hist_off, a.tile_off = tile_off, a.tile_ev = tile_ev, a.errflag = d_flags + 1;
        a.coef = reinterpret_cast<double *>(d_coef), a.rowcnt = d_rowcnt, a.tab = d_tab, a.nharm = nharm, a.extra_off = extra_off;
        a.rows_per_cta = 0;
        a.progress = streamed ? d_prog : nullptr, a.chunk_events = chunk_events;
        const int grid_x = (slots + R * W - 1) / (R * W);
        const long long want_y = std::max<long long>(1, 4LL * di.sm_count * ctas_per_sm / grid_x);
        long long per = std::max<long long>((n_ev + want_y - 1) / want_y, 4LL * tile_ev);
        per = (per + 1) & ~1LL;
        a.ev_per_cta = per;
        const int grid_y = (int)std::min<long long>(65535, (n_ev + per - 1) / per);
        if ((long long)grid_y * per < n_ev) a.ev_per_cta = ((n_ev + grid_y - 1) / grid_y + 1) & ~1LL;
        // Fused path with many more time slices than CTAs along y: give every CTA WHOLE slices, so that each slice's
        // coefficients come from one CTA only (bit-reproducible sums, no contended atomics).  Slices span equal times, so
        // for a steady source the CTAs stay balanced; k slices per CTA is chosen for well-filled waves.
        int grid_y_eff = grid_y;
        if (coef_mode && nprof >= 2 * want_y) {
            const long long slots = (long long)di.sm_count * ctas_per_sm;
            int best_k = 0;
            double best_eff = -1.0;
            const int k0 = (int)std::max<long long>(1, nprof / want_y);
            for (int k = std::max(1, k0 - 1); k <= k0 + 2; k++) {
                const double waves = (double)((nprof + k - 1) / k) * grid_x / (double)slots;
                const double eff = waves / std::ceil(waves) * (1.0 - 0.04 / std::max(waves, 1.0));
                if (eff > best_eff) best_eff = eff, best_k = k;
            }
            a.rows_per_cta = best_k;
            grid_y_eff = (nprof + best_k - 1) / best_k;
        }
        const bool timed = g_qev_ok && g_qev_used < QF_EVPOOL;
        if (timed) HX_CUDA_TRY(cudaEventRecord(g_qev[2 * g_qev_used], st));
        fn<<<dim3((unsigned)grid_x, (unsigned)grid_y_eff), W * 32, total, st>>>(a);
        HX_CHECK_LAUNCH("qffa_lane_kernel");
        if (timed) {
            HX_CUDA_TRY(cudaEventRecord(g_qev[2 * g_qev_used + 1], st));
            g_qev_used++;
        }
        g_qev_batches++;
        if (streamed) {
            if (!copies_issued) {  // pageable source: the kernel is already waiting for its chunks; feed it
                int rcc = issue_copies();
                if (rcc) return rcc;
            }
            // everything is up once the copy stream gets here: verify, on the device, the sortedness the host-side
            // boundaries relied on (the result is discarded and the caller falls back if it does not hold)
            HX_CUDA_TRY(cudaStreamWaitEvent(st, ev_sync, 0));
            qffa_prepass_kernel<<<(unsigned)std::min<long long>(std::max<long long>((n_ev + 255) / 256, 1), 2368), 256, 0, st>>>(
                d_t, n_ev, t_first, delta1, nprof, t_safe, nullptr, nullptr, d_flags);
            HX_CHECK_LAUNCH("qffa_prepass_kernel");
        }
        if (n_unsafe > 0) {
            dim3 gf((unsigned)std::min<long long>(n_unsafe, 256), (unsigned)((C + 127) / 128));
            if (coef_mode) gf.x = 1;  // one thread per centre walks all the (few) photons: its adds are ordered => reproducible
            if (coef_mode) {
                if (mode == 0) qffa_fix_kernel<0, true><<<gf, 128, 0, st>>>(a, d_urange, t_first, delta0, delta1);
                else if (mode == 1) qffa_fix_kernel<1, true><<<gf, 128, 0, st>>>(a, d_urange, t_first, delta0, delta1);
                else qffa_fix_kernel<2, true><<<gf, 128, 0, st>>>(a, d_urange, t_first, delta0, delta1);
            } else {
                if (mode == 0) qffa_fix_kernel<0><<<gf, 128, 0, st>>>(a, d_urange, t_first, delta0, delta1);
                else if (mode == 1) qffa_fix_kernel<1><<<gf, 128, 0, st>>>(a, d_urange, t_first, delta0, delta1);
                else qffa_fix_kernel<2><<<gf, 128, 0, st>>>(a, d_urange, t_first, delta0, delta1);
            }
            HX_CHECK_LAUNCH("qffa_fix_kernel");
        }
        if (d_profiles) {
            qffa_unlane_kernel<<<di.sm_count * 8, 256, 0, st>>>(reinterpret_cast<uint32_t *>(d_lane), C, nprof, nbin,
                                                                 d_profiles + (size_t)c0 * prof_elems);
            HX_CHECK_LAUNCH("qffa_unlane_kernel");
        }
        if (coef_mode) {
            qffa_tot_kernel<<<(unsigned)((C + 7) / 8), 256, 0, st>>>(d_rowcnt, C, nprof, d_tot);
            HX_CHECK_LAUNCH("qffa_tot_kernel");
        } else {
            HX_CUDA_TRY(cudaMemsetAsync(d_tot, 0, (size_t)cb_max * 8, st));
            dim3 g1((unsigned)((nprof + 7) / 8), (unsigned)slots);
            qffa_rowdft_lane_kernel<<<g1, 256, 2 * (size_t)nbin * sizeof(double), st>>>(
                reinterpret_cast<uint32_t *>(d_lane), C, nprof, nbin, nharm, d_tab, reinterpret_cast<double *>(d_coef), d_tot,
                transient_levels > 0 ? d_rowcnt : nullptr);
            HX_CHECK_LAUNCH("qffa_rowdft_lane_kernel");
        }
        if (transient_levels > 0) {
            dim3 g3((unsigned)((nprof + 1 + 127) / 128), (unsigned)transient_levels, (unsigned)C);
            qffa_transient_kernel<<<g3, 128, 0, st>>>(reinterpret_cast<double *>(d_coef), d_rowcnt, nprof, transient_rows,
                                                       nharm, transient_levels,
                                                       d_stats + (size_t)c0 * transient_levels * transient_rows);
            HX_CHECK_LAUNCH("qffa_transient_kernel");
        } else if (nharm <= 4) {
            const int ppc = (int)std::min<int64_t>(64, n_lq);  // pairs per CTA; 256/ppc groups of time slices
            const int groups = 256 / ppc;
            const size_t sm2 = (2 * (size_t)nbin + (size_t)groups * ppc * 2 * nharm) * sizeof(double);
            dim3 g2((unsigned)((n_lq + ppc - 1) / ppc), (unsigned)C);
            double *coefp = reinterpret_cast<double *>(d_coef), *outp = d_stats + (size_t)c0 * n_lq;
            switch (nharm) {
                case 1: qffa_step2_kernel<1><<<g2, 256, sm2, st>>>(coefp, d_tot, nprof, nbin, d_L, d_Q, (int)n_lq, ppc, d_base, d_tab, outp); break;
                case 2: qffa_step2_kernel<2><<<g2, 256, sm2, st>>>(coefp, d_tot, nprof, nbin, d_L, d_Q, (int)n_lq, ppc, d_base, d_tab, outp); break;
                case 3: qffa_step2_kernel<3><<<g2, 256, sm2, st>>>(coefp, d_tot, nprof, nbin, d_L, d_Q, (int)n_lq, ppc, d_base, d_tab, outp); break;
                default: qffa_step2_kernel<4><<<g2, 256, sm2, st>>>(coefp, d_tot, nprof, nbin, d_L, d_Q, (int)n_lq, ppc, d_base, d_tab, outp); break;
            }
            HX_CHECK_LAUNCH("qffa_step2_kernel");
        } else {
            dim3 g2((unsigned)n_lq, (unsigned)C);
            qffa_step_kernel<<<g2, 256, step_smem, st>>>(reinterpret_cast<double *>(d_coef), d_tot, nprof, nbin, d_L, d_Q,
                                                          (int)n_lq, d_base, d_tab, nharm, d_stats + (size_t)c0 * n_lq);
            HX_CHECK_LAUNCH("qffa_step_kernel");
        }
    }
    HX_CUDA_TRY(cudaMemcpyAsync(h_flags, d_flags, 8, cudaMemcpyDeviceToHost, st));
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    if (h_flags[1] == 2) {
        set_error("qffa_lane_kernel: the streamed upload of the events stalled");
        return HX_ERR_CUDA;
    }
    if (h_flags[1]) {
        set_error("qffa_lane_kernel: shared-memory layout mismatch");
        return HX_ERR_INTERNAL;
    }
    if (streamed && h_flags[0]) return 1;  // unsorted after all: the events are resident now, the caller falls back
    if (g_qev_used > 0) {
        double sum = 0.0;
        for (int i = 0; i < g_qev_used; i++) {
            float fms = 0.f;
            if (cudaEventElapsedTime(&fms, g_qev[2 * i], g_qev[2 * i + 1]) == cudaSuccess) sum += fms;
        }
        g_q_last_ms = sum * (double)g_qev_batches / (double)g_qev_used;
    }
    return HX_OK;
}

// transient_levels > 0: instead of the (L, Q) shift search, write the sub-interval statistics of the
// transient search, d_stats[n_centres][transient_levels][transient_rows]
int qffa_search_dev_impl(const double *d_t, int64_t n_ev, double t_first, double t_last, const double *mean_f,
                         const double *mean_fd, const double *mean_fdd, int64_t n_centres, int nbin, int nprof,
                         const double *L, const double *Q, int64_t n_lq, int nharm, double *d_stats,
                         uint32_t *d_profiles, cudaStream_t st, int transient_levels = 0, int transient_rows = 0,
                         const double *h_t = nullptr) {
    std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
    if (!d_t || !mean_f || !mean_fd || !mean_fdd || !L || !Q || !d_stats) {
        set_error("hx_qffa_search: null pointer argument");
        return HX_ERR_ARG;
    }
    if (n_ev < 1 || n_centres < 0 || n_lq < 1 || nbin < 2 || nbin > HX_MAX_NBIN || nprof < 1 || nprof > (1 << 20) ||
        nharm < 1 || nharm > HX_MAX_NHARM || n_lq > 65535) {
        set_error("hx_qffa_search: bad sizes (n_ev=%lld n_centres=%lld n_lq=%lld nbin=%d nprof=%d nharm=%d)",
                  (long long)n_ev, (long long)n_centres, (long long)n_lq, nbin, nprof, nharm);
        return HX_ERR_ARG;
    }
    if (n_centres == 0) return HX_OK;
    int device;
    HX_CUDA_TRY(cudaGetDevice(&device));
    DeviceInfo di;
    {
        int rc0 = get_device_info(device, &di);
        if (rc0) return rc0;
    }
    if ((reinterpret_cast<uintptr_t>(d_t) & 15) != 0) {
        set_error("hx_qffa_search_dev: d_times_c must be 16-byte aligned");
        return HX_ERR_ARG;
    }
    const int cmax = std::max(1, std::min(g_qffa_cmax, QF_CMAX));
    const bool p2 = (nbin & (nbin - 1)) == 0;
    // host-side constants, reference formulas
    const double delta0 = 1 / ((1.0 - 0.0) / nbin);
    const double delta1 = 1 / ((t_last - t_first) / nprof);
    std::vector<double> host((size_t)nprof + 2 * (size_t)nbin + 2 * (size_t)n_lq);
    double *h_base = host.data(), *h_tab = h_base + nprof, *h_L = h_tab + 2 * nbin, *h_Q = h_L + n_lq;
    {
        // np.linspace(-1, 1, nprof)
        const double step = nprof > 1 ? 2.0 / (nprof - 1) : 0.0;
        for (int k = 0; k < nprof; k++) h_base[k] = -1.0 + k * step;
        if (nprof > 1) h_base[nprof - 1] = 1.0;
        // 2*pi*arange(0, 1, 1/nbin)   (efsearch.py:819)
        const double ps = 1.0 / nbin;
        for (int j = 0; j < nbin; j++) {
            double tw = (2 * M_PI) * (j * ps);
            h_tab[j] = cos(tw);
            h_tab[nbin + j] = sin(tw);
        }
        for (int64_t i = 0; i < n_lq; i++) h_L[i] = L[i], h_Q[i] = Q[i];
    }
    constexpr unsigned int UNSAFE_CAP = 1u << 20;
    void *d_const = nullptr, *d_scratch = nullptr, *d_unsafe = nullptr;
    bool used_fast = false;
    HX_CUDA_TRY(cudaMallocAsync(&d_const, host.size() * sizeof(double), st));
    struct Free {
        void *&p;
        cudaStream_t s;
        ~Free() {
            if (p) cudaFreeAsync(p, s);
        }
    } f1{d_const, st}, f2{d_scratch, st}, f3{d_unsafe, st};
    HX_CUDA_TRY(cudaMallocAsync(&d_unsafe, (size_t)UNSAFE_CAP * 8 + 16, st));
    HX_CUDA_TRY(cudaMemsetAsync(reinterpret_cast<char *>(d_unsafe) + (size_t)UNSAFE_CAP * 8, 0, 16, st));
    HX_CUDA_TRY(cudaMemcpyAsync(d_const, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    const double *d_base = reinterpret_cast<double *>(d_const), *d_tab = d_base + nprof, *d_L = d_tab + 2 * nbin,
                 *d_Q = d_L + n_lq;
    if (!g_qev_ok) {
        bool ok = true;
        for (int i = 0; i < 2 * QF_EVPOOL && ok; i++) ok = cudaEventCreate(&g_qev[i]) == cudaSuccess;
        g_qev_ok = ok;
    }
    {
        // preferred path: lane = centre (conflict-free atomics); falls through when not eligible
        bool uploaded = false;
        int rl = qffa_lane_path(di, d_t, n_ev, t_first, t_last, mean_f, mean_fd, mean_fdd, n_centres, nbin, nprof,
                                delta0, delta1, d_L, d_Q, n_lq, d_base, d_tab, nharm, d_stats, d_profiles, st,
                                transient_levels, transient_rows, h_t, &uploaded);
        if (rl <= 0) return rl;
        if (h_t && !uploaded)  // the lane = centre path declined before touching the events: plain upload for the fallback
            HX_CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(d_t), h_t, (size_t)n_ev * 8, cudaMemcpyHostToDevice, st));
    }
    const size_t prof_elems = (size_t)nprof * nbin;
    if (!d_profiles) HX_CUDA_TRY(cudaMallocAsync(&d_scratch, prof_elems * cmax * sizeof(uint32_t), st));

    const size_t step_smem = (2 * (size_t)nbin + 16) * sizeof(double);
    const size_t coef_elems = (size_t)cmax * nprof * nharm * 2;
    void *d_coef = nullptr;
    Free f4{d_coef, st};
    HX_CUDA_TRY(cudaMallocAsync(&d_coef, coef_elems * sizeof(double) + (size_t)QF_CMAX * 8 + (size_t)cmax * nprof * 8, st));
    unsigned long long *d_tot = reinterpret_cast<unsigned long long *>(reinterpret_cast<double *>(d_coef) + coef_elems);
    unsigned long long *d_rowcnt = d_tot + QF_CMAX;

    if (!g_qev_ok) {
        bool ok = true;
        for (int i = 0; i < 2 * QF_EVPOOL && ok; i++) ok = cudaEventCreate(&g_qev[i]) == cudaSuccess;
        g_qev_ok = ok;
    }
    g_qev_used = 0, g_qev_batches = 0, g_q_last_ms = -1.0;
    for (int64_t c0 = 0; c0 < n_centres; c0 += cmax) {
        const int C = (int)std::min<int64_t>(cmax, n_centres - c0);
        uint32_t *prof = d_profiles ? d_profiles + (size_t)c0 * prof_elems : reinterpret_cast<uint32_t *>(d_scratch);
        HX_CUDA_TRY(cudaMemsetAsync(prof, 0, prof_elems * C * sizeof(uint32_t), st));
        QffaHistArgs a{};
        a.t = d_t, a.n_ev = n_ev, a.t_first = t_first, a.delta0 = delta0, a.delta1 = delta1;
        a.nbin = nbin, a.nprof = nprof, a.prof = prof;
        int mode = 0;
        double fmin = INFINITY, fmax = 0.0, fdabs = 0.0, fddabs = 0.0;
        for (int c = 0; c < C; c++) {
            a.raw.f[c] = mean_f[c0 + c], a.raw.fd[c] = mean_fd[c0 + c], a.raw.fdd[c] = mean_fdd[c0 + c];
            // efsearch.py:892-897 picks the formula per step; adding exact-zero terms is value-neutral
            if (a.raw.fdd[c] != 0) mode = 2;
            else if (a.raw.fd[c] != 0) mode = std::max(mode, 1);
            fmin = std::min(fmin, a.raw.f[c]);
            fmax = std::max(fmax, std::fabs(a.raw.f[c]));
            fdabs = std::max(fdabs, std::fabs(a.raw.fd[c]));
            fddabs = std::max(fddabs, std::fabs(a.raw.fdd[c]));
        }
        // FAST path: for |t| in [t_safe, tmax], |phase| >= g(|t|) = |t| fmin - t^2 |fd|/2 - |t|^3 |fdd|/6;
        // g is concave for t > 0, so g >= 0.5 on the interval iff it holds at both ends.
        bool fast = p2 && fmin > 0 && std::isfinite(fmin);
        if (fast) {
            const double tmax = std::max(std::fabs(t_first), std::fabs(t_last));
            auto g = [&](double x) { return x * fmin - 0.5 * x * x * fdabs - x * x * x * fddabs / 6; };
            a.t_safe = 1.0 / fmin;
            const double phmax = tmax * fmax + 0.5 * tmax * tmax * fdabs + tmax * tmax * tmax * fddabs / 6;
            fast = g(a.t_safe) >= 0.75 && (tmax <= a.t_safe || g(tmax) >= 0.75) && phmax * nbin < 1125899906842624.0;
        }
        if (g_qffa_fast == 0) fast = false;
        int logk = 3;  // 8 bank-spread copies while the window stays <= 64 KB
        while (logk > 0 && ((size_t)C * QF_WROWS * nbin * sizeof(uint32_t) << logk) > 64 * 1024) logk--;
        if (g_qffa_logk >= 0) logk = std::min(logk, g_qffa_logk);
        a.logk = logk;
        {
            const double sc = (double)nbin * (double)(1 << (logk + 2));  // power of two: exact
            for (int c = 0; c < C; c++)
                a.c.f[c] = a.raw.f[c] * sc, a.c.fd[c] = a.raw.fd[c] * sc, a.c.fdd[c] = a.raw.fdd[c] * sc;
            if (fast) {
                const double tmax = std::max(std::fabs(t_first), std::fabs(t_last));
                const double phmax = tmax * fmax + 0.5 * tmax * tmax * fdabs + tmax * tmax * tmax * fddabs / 6;
                if (!(phmax * sc < 1125899906842624.0)) fast = false;
            }
        }
        a.unsafe_t = reinterpret_cast<double *>(d_unsafe);
        a.unsafe_n = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(d_unsafe) + (size_t)UNSAFE_CAP * 8);
        a.unsafe_cap = UNSAFE_CAP;
        if (fast) HX_CUDA_TRY(cudaMemsetAsync(a.unsafe_n, 0, sizeof(unsigned int), st));
        const size_t hsmem = QF_HDR + QF_STAGES * (size_t)QF_TILE * 8 + ((size_t)C * QF_WROWS * nbin * sizeof(uint32_t) << logk);
        if (hsmem > (size_t)di.smem_optin) {
            set_error("hx_qffa_search: nbin=%d needs %zu B of shared memory", nbin, hsmem);
            return HX_ERR_ARG;
        }
        // persistent CTAs: as many as fit, each with a contiguous, stage-aligned range of events
        const int ctas_per_sm = std::max(1, std::min(8, (int)(di.smem_per_sm / (hsmem + di.reserved_smem))));
        long long n_cta = std::min<long long>((long long)di.sm_count * ctas_per_sm, (n_ev + QF_TILE - 1) / QF_TILE);
        long long per = (n_ev + n_cta - 1) / n_cta;
        per = (per + QF_TILE - 1) / QF_TILE * QF_TILE;
        n_cta = (n_ev + per - 1) / per;
        a.ev_per_cta = per;
        const unsigned grid = (unsigned)n_cta;
        qhist_fn_t fn = pick_qhist(C, mode, fast);
        HX_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
        const bool timed = g_qev_ok && g_qev_used < QF_EVPOOL;
        if (timed) HX_CUDA_TRY(cudaEventRecord(g_qev[2 * g_qev_used], st));
        fn<<<grid, QF_THREADS, hsmem, st>>>(a);
        HX_CHECK_LAUNCH("qffa_hist_kernel");
        if (timed) {
            HX_CUDA_TRY(cudaEventRecord(g_qev[2 * g_qev_used + 1], st));
            g_qev_used++;
        }
        g_qev_batches++;
        if (fast) {
            if (mode == 0) qffa_list_kernel<0><<<64, 256, 0, st>>>(a, C);
            else if (mode == 1) qffa_list_kernel<1><<<64, 256, 0, st>>>(a, C);
            else qffa_list_kernel<2><<<64, 256, 0, st>>>(a, C);
            HX_CHECK_LAUNCH("qffa_list_kernel");
            used_fast = true;
        }
        HX_CUDA_TRY(cudaMemsetAsync(d_tot, 0, (size_t)QF_CMAX * 8, st));
        dim3 g1((unsigned)((nprof + 7) / 8), (unsigned)C);
        qffa_rowdft_kernel<<<g1, 256, 2 * (size_t)nbin * sizeof(double), st>>>(
            prof, nprof, nbin, nharm, d_tab, reinterpret_cast<double *>(d_coef), d_tot,
            transient_levels > 0 ? d_rowcnt : nullptr);
        HX_CHECK_LAUNCH("qffa_rowdft_kernel");
        if (transient_levels > 0) {
            dim3 g3((unsigned)((nprof + 1 + 127) / 128), (unsigned)transient_levels, (unsigned)C);
            qffa_transient_kernel<<<g3, 128, 0, st>>>(reinterpret_cast<double *>(d_coef), d_rowcnt, nprof, transient_rows,
                                                       nharm, transient_levels,
                                                       d_stats + (size_t)c0 * transient_levels * transient_rows);
            HX_CHECK_LAUNCH("qffa_transient_kernel");
        } else {
            dim3 g2((unsigned)n_lq, (unsigned)C);
            qffa_step_kernel<<<g2, 256, step_smem, st>>>(reinterpret_cast<double *>(d_coef), d_tot, nprof, nbin, d_L,
                                                          d_Q, (int)n_lq, d_base, d_tab, nharm,
                                                          d_stats + (size_t)c0 * n_lq);
            HX_CHECK_LAUNCH("qffa_step_kernel");
        }
    }
    unsigned int flags_h[2] = {0, 0};
    HX_CUDA_TRY(cudaMemcpyAsync(flags_h, reinterpret_cast<char *>(d_unsafe) + (size_t)UNSAFE_CAP * 8, 8,
                                cudaMemcpyDeviceToHost, st));
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    if (g_qev_used > 0) {
        double sum = 0.0;
        for (int i = 0; i < g_qev_used; i++) {
            float f = 0.f;
            if (cudaEventElapsedTime(&f, g_qev[2 * i], g_qev[2 * i + 1]) == cudaSuccess) sum += f;
        }
        g_q_last_ms = sum * (double)g_qev_batches / (double)g_qev_used;  // scaled if more batches than pooled events
    }
    if (used_fast && flags_h[1] != 0) {
        // more than UNSAFE_CAP events inside |t| < 1/fmin: redo everything on the generic path
        const int saved = g_qffa_fast;
        g_qffa_fast = 0;
        int rc = qffa_search_dev_impl(d_t, n_ev, t_first, t_last, mean_f, mean_fd, mean_fdd, n_centres, nbin, nprof, L,
                                      Q, n_lq, nharm, d_stats, d_profiles, st, transient_levels, transient_rows);
        g_qffa_fast = saved;
        return rc;
    }
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_qffa_last_kernel_ms(double *ms, int64_t *launches) {
    if (!ms || hx::g_q_last_ms < 0) {
        hx::set_error("hx_qffa_last_kernel_ms: no timed --fast call yet");
        return HX_ERR_ARG;
    }
    *ms = hx::g_q_last_ms;
    if (launches) *launches = hx::g_qev_batches;
    return HX_OK;
}

int hx_qffa_set_tuning(int cmax) {
    // cmax in [1, 8]: centres per pass; +100 additionally disables the power-of-two fast phase path
    // +1000: additionally force the lane = event kernel (disable the lane = centre path)
    hx::g_qffa_lane = cmax >= 1000 ? 0 : 1;
    if (cmax >= 1000) cmax -= 1000;
    hx::g_qffa_fast = cmax >= 100 ? 0 : 1;
    if (cmax >= 100) cmax -= 100;
    hx::g_qffa_cmax = (cmax >= 1 && cmax <= hx::QF_CMAX) ? cmax : hx::QF_CMAX;
    return HX_OK;
}

int hx_qffa_search_dev(const double *d_times_c, int64_t n_ev, double t_first, double t_last, const double *mean_f,
                       const double *mean_fd, const double *mean_fdd, int64_t n_centres, int nbin, int nprof,
                       const double *L, const double *Q, int64_t n_lq, int nharm, double *d_stats,
                       uint32_t *d_profiles, void *stream) {
    return hx::qffa_search_dev_impl(d_times_c, n_ev, t_first, t_last, mean_f, mean_fd, mean_fdd, n_centres, nbin,
                                    nprof, L, Q, n_lq, nharm, d_stats, d_profiles,
                                    reinterpret_cast<cudaStream_t>(stream));
}

int hx_transient_search_dev(const double *d_times_c, int64_t n_ev, double t_first, double t_last,
                            const double *mean_f, const double *mean_fd, int64_t n_centres, int nbin, int nprof,
                            int nprof_used, int nharm, int n_levels, double *d_stats, void *stream) {
    if (n_levels < 1 || n_levels > 30 || n_centres < 0 || !mean_f || !mean_fd || nprof_used < 1 || nprof_used > nprof) {
        hx::set_error("hx_transient_search_dev: bad arguments");
        return HX_ERR_ARG;
    }
    std::vector<double> zeros((size_t)std::max<int64_t>(n_centres, 1), 0.0);
    const double one_lq[1] = {0.0};
    return hx::qffa_search_dev_impl(d_times_c, n_ev, t_first, t_last, mean_f, mean_fd, zeros.data(), n_centres, nbin,
                                    nprof, one_lq, one_lq, 1, nharm, d_stats, nullptr,
                                    reinterpret_cast<cudaStream_t>(stream), n_levels, nprof_used);
}

int hx_qffa_search(const double *times_c, int64_t n_ev, const double *mean_f, const double *mean_fd,
                   const double *mean_fdd, int64_t n_centres, int nbin, int nprof, const double *L, const double *Q,
                   int64_t n_lq, int nharm, double *stats, uint32_t *profiles, int device) {
    if (!times_c || !stats || n_ev < 1 || n_centres < 0 || n_lq < 1) {
        hx::set_error("hx_qffa_search: bad arguments");
        return HX_ERR_ARG;
    }
    hx::DeviceGuard dev_guard;  // declared before the device buffers: they are freed before the device is restored
    HX_CUDA_TRY(dev_guard.enter(device));
    // the events are uploaded by the search itself: in chunks, while K2 already folds the chunks that have arrived
    // (fused path, large lists), else in one copy
    hx::AsyncBuf dt(0), ds(0), dp(0);
    HX_CUDA_TRY(dt.alloc((size_t)n_ev * 8));
    HX_CUDA_TRY(ds.alloc((size_t)n_centres * n_lq * 8));
    const size_t pe = (size_t)n_centres * nprof * nbin;
    if (profiles) HX_CUDA_TRY(dp.alloc(pe * 4));
    int rc = hx::qffa_search_dev_impl(dt.as<double>(), n_ev, times_c[0], times_c[n_ev - 1], mean_f, mean_fd, mean_fdd,
                                      n_centres, nbin, nprof, L, Q, n_lq, nharm, ds.as<double>(),
                                      profiles ? dp.as<uint32_t>() : nullptr, 0, 0, 0, times_c);
    if (rc) return rc;
    HX_CUDA_TRY(cudaMemcpy(stats, ds.p, (size_t)n_centres * n_lq * 8, cudaMemcpyDeviceToHost));
    if (profiles) HX_CUDA_TRY(cudaMemcpy(profiles, dp.p, pe * 4, cudaMemcpyDeviceToHost));
    return HX_OK;
}

}  // extern "C"
