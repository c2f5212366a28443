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

// Transient search (hendrics/efsearch.py:407-465, _average_and_z_sub_search): for every averaging level
// l (groups of 2^l consecutive time slices) the Z^2_n of the summed sub-profiles, written to all rows of
// the group.  Sums of profiles are sums of their harmonic coefficients, so no profile is touched again.
__global__ void __launch_bounds__(128) qffa_transient_kernel(const double *__restrict__ coef,
                                                             const unsigned long long *__restrict__ rowcnt,
                                                             int nprof, int nharm, int n_levels,
                                                             double *__restrict__ results /*[C][n_levels][nprof]*/) {
    const int c = blockIdx.z, level = blockIdx.y;
    const int n_ave = 1 << level;
    const int groups = nprof / n_ave;  // int(profiles.shape[0] / n_ave_i)
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    double *out = results + ((size_t)c * n_levels + level) * nprof;
    if (g >= groups) {
        // rows not covered by any group keep the reference's initial 0
        if (g == groups)
            for (int r = groups * n_ave; r < nprof; r++) out[r] = 0.0;
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

// transient_levels > 0: instead of the (L, Q) shift search, write the sub-interval statistics of the
// transient search, d_stats[n_centres][transient_levels][nprof]
int qffa_search_dev_impl(const double *d_t, int64_t n_ev, double t_first, double t_last, const double *mean_f,
                         const double *mean_fd, const double *mean_fdd, int64_t n_centres, int nbin, int nprof,
                         const double *L, const double *Q, int64_t n_lq, int nharm, double *d_stats,
                         uint32_t *d_profiles, cudaStream_t st, int transient_levels = 0) {
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
            qffa_transient_kernel<<<g3, 128, 0, st>>>(reinterpret_cast<double *>(d_coef), d_rowcnt, nprof, nharm,
                                                       transient_levels,
                                                       d_stats + (size_t)c0 * transient_levels * nprof);
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
                                      Q, n_lq, nharm, d_stats, d_profiles, st, transient_levels);
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
                            int nharm, int n_levels, double *d_stats, void *stream) {
    if (n_levels < 1 || n_levels > 30 || n_centres < 0 || !mean_f || !mean_fd) {
        hx::set_error("hx_transient_search_dev: bad arguments");
        return HX_ERR_ARG;
    }
    std::vector<double> zeros((size_t)std::max<int64_t>(n_centres, 1), 0.0);
    const double one_lq[1] = {0.0};
    return hx::qffa_search_dev_impl(d_times_c, n_ev, t_first, t_last, mean_f, mean_fd, zeros.data(), n_centres, nbin,
                                    nprof, one_lq, one_lq, 1, nharm, d_stats, nullptr,
                                    reinterpret_cast<cudaStream_t>(stream), n_levels);
}

int hx_qffa_search(const double *times_c, int64_t n_ev, const double *mean_f, const double *mean_fd,
                   const double *mean_fdd, int64_t n_centres, int nbin, int nprof, const double *L, const double *Q,
                   int64_t n_lq, int nharm, double *stats, uint32_t *profiles, int device) {
    if (!times_c || !stats || n_ev < 1 || n_centres < 0 || n_lq < 1) {
        hx::set_error("hx_qffa_search: bad arguments");
        return HX_ERR_ARG;
    }
    HX_CUDA_TRY(cudaSetDevice(device));
    hx::DevBuf dt, ds, dp;
    HX_CUDA_TRY(dt.alloc((size_t)n_ev * 8));
    HX_CUDA_TRY(ds.alloc((size_t)n_centres * n_lq * 8));
    const size_t pe = (size_t)n_centres * nprof * nbin;
    if (profiles) HX_CUDA_TRY(dp.alloc(pe * 4));
    HX_CUDA_TRY(cudaMemcpy(dt.p, times_c, (size_t)n_ev * 8, cudaMemcpyHostToDevice));
    int rc = hx::qffa_search_dev_impl(dt.as<double>(), n_ev, times_c[0], times_c[n_ev - 1], mean_f, mean_fd, mean_fdd,
                                      n_centres, nbin, nprof, L, Q, n_lq, nharm, ds.as<double>(),
                                      profiles ? dp.as<uint32_t>() : nullptr, 0);
    if (rc) return rc;
    HX_CUDA_TRY(cudaMemcpy(stats, ds.p, (size_t)n_centres * n_lq * 8, cudaMemcpyDeviceToHost));
    if (profiles) HX_CUDA_TRY(cudaMemcpy(profiles, dp.p, pe * 4, cudaMemcpyDeviceToHost));
    return HX_OK;
}

}  // extern "C"
