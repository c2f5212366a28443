// Fast folding algorithm (HENzsearch / HENefsearch --ffa) for sm_100a.
//
// Replaces hendrics/ffa.py:291-349 (`ffa_step`, `_ffa`, `ffa`), hendrics/ffa.py:352-363 (`_quick_rebin`) and the 1-D
// histogram of hendrics/base.py:621-630 that `search_with_ffa` (hendrics/efsearch.py:1092-1138) feeds them with.
//
// The reference zero-pads the binned light curve to ntables x P (P = the trial period in bins, ntables a power of
// two), and runs log2(ntables) steps; step s replaces table n of every aligned group of 2^(s+1) tables by
//     A[g + l/2] + roll(A[g + l/2 + 2^s], -((l + 1)/2)),        g = n - l,  l = n mod 2^(s+1)          (ffa.py:234-308)
// and finally takes Z^2_n of every table.  Done step by step that is log2(ntables) read + write passes over HBM.
//
// Design (DESIGN.md "K5"): the butterfly is blocked.  A pass covers k consecutive steps s0 .. s0+k-1.  The 2^k output
// tables  c*2^k + m  (m < 2^k) of a pass depend on exactly the 2^k input tables  G + q*2^s0 + h  (q < 2^k) with
// G = (c >> s0) << (s0 + k), h = c mod 2^s0; on that working set, local step sigma is the reference's step sigma with every
// shift increased by h*2^sigma.  So one CTA gathers its 2^k tables into shared memory (coalesced row reads, zero beyond the
// raw light curve: the padded copy is never materialised), runs k steps between two shared-memory buffers and writes 2^k
// contiguous tables -- or, on the last pass, only their Z^2_n: the final tables never reach HBM.  HBM traffic is
// 8 B per table element and PASS instead of per step (log2(ntables) = 20: 3 passes of 7, 7, 6 steps).  Within a step a warp
// folds output tables 2m and 2m+1 together: they read the same two input tables with shifts one bin apart.
// Counts are uint32: the reference adds integer-valued doubles, which is exact, so integer sums are the same numbers.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "hx_common.cuh"

namespace hx {

// shared-memory row stride (in counts): a multiple of 4 (16-byte aligned rows whatever P) that is 4 mod 8, so that the 8
// tables x 4 consecutive bins an FP64 MMA A-fragment reads (epilogue) fall into 32 distinct banks
__host__ __device__ __forceinline__ int ffa_row_stride(int P) {
    int ps = (P + 3) & ~3;
    if ((ps & 7) != 4) ps += 4;
    return ps;
}

struct FfaPassArgs {
    const uint32_t *src;  // first pass: the raw light curve (n_src valid bins, zero beyond); later: ntables x P tables
    long long n_src;
    uint32_t *dst;        // ntables x P tables; null on the last pass
    long long ntables;
    int P, s0, k, nharm;
    int vec4;             // P % 4 == 0 and 16-byte aligned arrays: 128-bit copies
    const double *tab;    // [2][nharm][P]: cos, sin values harmonic h multiplies bin j with (last pass)
    double *stats;        // [ntables] (last pass)
};

// asynchronous global -> shared copies with zero fill of the bytes beyond src_bytes (SASS LDGSTS)
__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Z^2_n of one table (ffa.py:59-104 on the gathered tables); all lanes of the warp return the value
__device__ __forceinline__ double ffa_row_stat(const uint32_t *row, int P, int nharm, const double *tab, int lane) {
    double tot = 0.0;
    for (int j = lane; j < P; j += 32) tot += (double)row[j];
    tot = warp_sum(tot);
    double res = 0.0;
    for (int h = 0; h < nharm; h++) {
        const double *ct = tab + (size_t)h * P, *st = tab + (size_t)(nharm + h) * P;
        double c = 0.0, s = 0.0;
        for (int j = lane; j < P; j += 32) {
            const double p = (double)row[j];
            c += ct[j] * p;
            s += st[j] * p;
        }
        c = warp_sum(c), s = warp_sum(s);
        res += c * c + s * s;
    }
    return 2.0 / tot * res;  // tot = all the counts of the light curve; an empty one gives inf * 0 = NaN
}

// Persistent CTAs: tile t+1 (the 2^k tables of the next working set) is prefetched with asynchronous copies (LDGSTS)
// into the second load buffer while tile t is folded.
// Shared memory: two load buffers + one work buffer of 2^k tables at a row stride of Ps = ffa_row_stride(P) counts, + the
// harmonic tables on the last pass.
// Last pass: Z^2_n of the 2^k final tables straight from shared memory.  tables x [1, cos_1, sin_1, ..] is a small dense
// contraction (2^k x P x (1 + 2 nharm)), so for nharm <= 3 it runs on the FP64 tensor cores: one warp per 8 tables,
// mma.sync.m8n8k4.f64 over the bins, the harmonic table pre-arranged in B-fragment order (first version: a warp per table
// with three passes over the row and 50 shuffles -- 45 % of the last pass).
// Fold step: 4 bins per lane and trip, all loads of a trip in flight before the first add, 32-bit shared-window
// addresses.  (The first version -- one bin per lane and trip, generic pointers, a 64-bit modulo per table -- was bound by
// instruction issue at 1 warp-instruction per bin and step; ncu history in DESIGN.md.)
__global__ void __launch_bounds__(512) ffa_pass_kernel(const FfaPassArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P = a.P, k = a.k, s0 = a.s0;
    const int Ps = ffa_row_stride(P), nvec = (P + 3) >> 2;
    const int nrow = 1 << k;
    const uint32_t tile_bytes = (uint32_t)nrow * (uint32_t)Ps * 4u;
    const uint32_t s_ld0 = smem_u32(smem_raw);  // load buffers: s_ld0, s_ld0 + tile_bytes; work buffer behind them
    const uint32_t s_work = s_ld0 + 2u * tile_bytes;
    double *stab = reinterpret_cast<double *>(smem_raw + 3 * (size_t)tile_bytes + 512);  // last pass only
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool last = a.dst == nullptr;
    const long long ntiles = a.ntables >> k;

    const int step32 = 32 % P;  // rotated index advance between the 4 bins of a lane (32, or 32 mod P for tiny periods)
    // bins covered by 4-bins-per-lane trips: a remainder of more than 64 bins still takes one (partly idle; its loads may
    // run up to 96 words past the row, into the slack), a shorter one goes bin by bin
    const int Pmain = (P & 127) > 64 ? ((P + 127) & ~127) : (P & ~127);

    // MMA epilogue: stab[kq][lane] = T[4 kq + lane % 4][lane / 4], T[j] = (1, cos_1 j, sin_1 j, cos_2 j, ...) padded with
    // zeros to 8 columns and to a multiple of 4 bins; else the plain [2][nharm][P] tables
    const bool mma_stat = last && a.nharm <= 3 && nrow >= 8;
    if (mma_stat) {
        const int nk = (P + 3) >> 2;
        for (int i = threadIdx.x; i < nk * 32; i += blockDim.x) {
            const int j = 4 * (i >> 5) + (i & 3), n = (i & 31) >> 2;
            double v = 0.0;
            if (j < P) {
                if (n == 0) v = 1.0;
                else if (n <= 2 * a.nharm) v = a.tab[(size_t)(((n - 1) & 1) * a.nharm + ((n - 1) >> 1)) * P + j];
            }
            stab[i] = v;
        }
    } else if (last) {
        for (int i = threadIdx.x; i < 2 * a.nharm * P; i += blockDim.x) stab[i] = a.tab[i];
    }

    // gather of tile c: local table q <- table G + q*2^s0 + h; bins beyond the raw light curve are zero-filled by the
    // copy itself (src-size 0)
    auto issue_gather = [&](long long c, uint32_t sb) {
        const long long h = c & ((1LL << s0) - 1);
        const long long G = (c >> s0) << (s0 + k);
        if (a.vec4) {  // P % 4 == 0: rows are 16-byte aligned in HBM too
            for (int q = warp; q < nrow; q += nwarps) {
                const long long base = (G + ((long long)q << s0) + h) * P;
                const uint32_t d = sb + (uint32_t)q * (uint32_t)Ps * 4u;
                for (int j4 = lane; j4 < nvec; j4 += 32) {
                    const long long idx = base + 4LL * j4;
                    const long long left = a.n_src - idx;
                    const uint32_t nb = left >= 4 ? 16u : left > 0 ? (uint32_t)left * 4u : 0u;
                    cp_async16_zfill(d + 16u * (uint32_t)j4, a.src + (nb ? idx : 0), nb);
                }
            }
        } else {
            for (int q = warp; q < nrow; q += nwarps) {
                const long long base = (G + ((long long)q << s0) + h) * P;
                const uint32_t d = sb + (uint32_t)q * (uint32_t)Ps * 4u;
                for (int j = lane; j < P; j += 32) {
                    const long long idx = base + j;
                    const bool ok = idx < a.n_src;
                    cp_async4_zfill(d + 4u * (uint32_t)j, a.src + (ok ? idx : 0), ok ? 4u : 0u);
                }
            }
        }
        cp_async_commit();
    };

    long long c = blockIdx.x;
    if (c < ntiles) issue_gather(c, s_ld0);
    for (int it = 0; c < ntiles; it++, c += gridDim.x) {
        cp_async_wait_all();
        __syncthreads();  // tile c has landed; every thread is done with the previous tile (its buffers are free)
        if (c + gridDim.x < ntiles) issue_gather(c + gridDim.x, s_ld0 + (uint32_t)((it + 1) & 1) * tile_bytes);
        // hs = (h * 2^sg) mod P, doubled from step to step: the only 64-bit division is this one per tile
        uint32_t hs = (uint32_t)((c & ((1LL << s0) - 1)) % P);

        uint32_t s_in = s_ld0 + (uint32_t)(it & 1) * tile_bytes, s_out = s_work;
        for (int sg = 0; sg < k; sg++) {
            // Output tables 2m and 2m+1 of a step read the SAME two input tables, with shifts that differ by one bin
            // (l even: start l/2, shift l/2; l+1: start l/2, shift l/2 + 1), so a warp folds them as a pair: per 4 bins and
            // lane 4 + 4 + 4 loads and 8 stores instead of 16 + 8, one row set-up instead of two.
            for (int r = 2 * warp; r < nrow; r += 2 * nwarps) {
                const int mu = r & ((2 << sg) - 1);  // even
                const int i1 = (r - mu) + (mu >> 1), i2 = i1 + (1 << sg);
                uint32_t t = hs + (uint32_t)(mu >> 1);  // shift of table r, reduced mod P; table r+1: one more
                if (t >= (uint32_t)P) {
                    t -= (uint32_t)P;
                    if (t >= (uint32_t)P) t %= (uint32_t)P;
                }
                const int sh = (int)t;
                const uint32_t a1 = s_in + (uint32_t)i1 * (uint32_t)Ps * 4u, a2 = s_in + (uint32_t)i2 * (uint32_t)Ps * 4u;
                const uint32_t ao = s_out + (uint32_t)r * (uint32_t)Ps * 4u, ao1 = ao + (uint32_t)Ps * 4u;
                // 4 bins per lane and trip at a lane stride of one word (bins j0, j0+32, j0+64, j0+96): every LDS / STS of
                // the warp touches 32 consecutive words -- conflict-free.  (128-bit accesses for the unrotated table with
                // four 16-byte-strided LDS.32 for the rotated one cost twice the wavefronts: 4-way bank conflicts, ncu.)
                for (int j0 = lane; j0 < Pmain; j0 += 128) {
                    int e = j0 + sh;
                    if (e >= P) e -= P;
                    uint32_t x[4], y0[4], y1[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        x[u] = lds_u32(a1 + (uint32_t)((j0 + 32 * u) << 2));
                        y0[u] = lds_u32(a2 + (uint32_t)(e << 2));
                        const int e1 = (e + 1 == P) ? 0 : e + 1;
                        y1[u] = lds_u32(a2 + (uint32_t)(e1 << 2));
                        e += step32;
                        if (e >= P) e -= P;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (j0 + 32 * u < P) {
                            sts_u32(ao + (uint32_t)((j0 + 32 * u) << 2), x[u] + y0[u]);
                            sts_u32(ao1 + (uint32_t)((j0 + 32 * u) << 2), x[u] + y1[u]);
                        }
                }
                // the last P mod 128 bins, one per lane and trip (a 4-bin trip for them would be mostly idle lanes:
                // P = 129 ran 55 % slower than P = 128 that way)
                for (int j = Pmain + lane; j < P; j += 32) {
                    int e = j + sh;
                    if (e >= P) e -= P;
                    const int e1 = (e + 1 == P) ? 0 : e + 1;
                    const uint32_t xv = lds_u32(a1 + (uint32_t)(j << 2));
                    sts_u32(ao + (uint32_t)(j << 2), xv + lds_u32(a2 + (uint32_t)(e << 2)));
                    sts_u32(ao1 + (uint32_t)(j << 2), xv + lds_u32(a2 + (uint32_t)(e1 << 2)));
                }
            }
            __syncthreads();
            const uint32_t tmp = s_in;
            s_in = s_out, s_out = tmp;
            hs <<= 1;
            if (hs >= (uint32_t)P) hs -= (uint32_t)P;
        }

        // s_in now holds tables c*2^k .. c*2^k + 2^k - 1 of the next level
        const uint32_t *in = reinterpret_cast<const uint32_t *>(smem_raw + (s_in - s_ld0));
        if (mma_stat) {
            const int nk = (P + 3) >> 2;
            const int grp = lane >> 2, tig = lane & 3;
            for (int r0 = warp * 8; r0 < nrow; r0 += nwarps * 8) {
                const uint32_t arow = s_in + (uint32_t)(r0 + grp) * (uint32_t)Ps * 4u + (uint32_t)tig * 4u;
                double d0 = 0.0, d1 = 0.0;  // D[grp][2 tig], D[grp][2 tig + 1]
                for (int kq = 0; kq < nk; kq++) {
                    // bins past P multiply zero table rows (the counts there are stale but finite)
                    const double av = __uint2double_rn(lds_u32(arow + (uint32_t)kq * 16u));
                    const double bv = stab[kq * 32 + lane];
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(d0), "+d"(d1)
                                 : "d"(av), "d"(bv));
                }
                // column 0 = total counts (lane tig 0, d0); every other column enters as its square
                double sq = (tig == 0 ? 0.0 : d0 * d0) + d1 * d1;
                sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                const double tot = __shfl_sync(0xffffffffu, d0, lane & ~3);
                if (tig == 0) a.stats[c * nrow + r0 + grp] = 2.0 / tot * sq;
            }
        } else if (last) {
            for (int r = warp; r < nrow; r += nwarps) {
                const double z = ffa_row_stat(in + (size_t)r * Ps, P, a.nharm, stab, lane);
                if (lane == 0) a.stats[c * nrow + r] = z;
            }
        } else {
            uint32_t *d = a.dst + (size_t)c * nrow * P;  // 2^k contiguous tables (row stride P in HBM)
            if (a.vec4) {  // P % 4 == 0: 128-bit copies, row by row (the shared-memory rows are padded)
                for (int q = warp; q < nrow; q += nwarps) {
                    const uint4 *s4 = reinterpret_cast<const uint4 *>(in + (size_t)q * Ps);
                    uint4 *d4 = reinterpret_cast<uint4 *>(d + (size_t)q * P);
                    for (int j4 = lane; j4 < nvec; j4 += 32) d4[j4] = s4[j4];
                }
            } else {
                for (int q = warp; q < nrow; q += nwarps)
                    for (int j = lane; j < P; j += 32) d[(size_t)q * P + j] = in[(size_t)q * Ps + j];
            }
        }
    }
}

// ---- fallback for trial periods too long for shared memory: one step per launch through HBM ---------------------
__global__ void __launch_bounds__(256) ffa_step_global_kernel(const uint32_t *__restrict__ src, long long n_src,
                                                              uint32_t *__restrict__ dst, long long ntables, int P, int step) {
    const int lane = threadIdx.x & 31;
    const long long warp_g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long sp = 2LL << step;
    for (long long n = warp_g; n < ntables; n += nw) {
        const long long l = n & (sp - 1), g = n - l;
        const long long i1 = g + (l >> 1), i2 = i1 + (1LL << step);
        const int sh = (int)(((l + 1) >> 1) % P);
        for (int j = lane; j < P; j += 32) {
            int jj = j + sh;
            if (jj >= P) jj -= P;
            const long long x1 = i1 * P + j, x2 = i2 * P + jj;
            dst[n * P + j] = (x1 < n_src ? src[x1] : 0u) + (x2 < n_src ? src[x2] : 0u);
        }
    }
}

__global__ void __launch_bounds__(256) ffa_stat_global_kernel(const uint32_t *__restrict__ tables, long long n_src,
                                                              long long ntables, int P, int nharm,
                                                              const double *__restrict__ tab, double *__restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const long long warp_g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long n = warp_g; n < ntables; n += nw) {
        const double z = ffa_row_stat(tables + (size_t)n * P, P, nharm, tab, lane);
        if (lane == 0) stats[n] = z;
    }
    (void)n_src;
}

// ---- 1-D histogram and rebinning ----------------------------------------------------------------------------------
// hendrics/base.py:621-630: delta = 1/((hi-lo)/bins); i = (x-lo)*delta; counted iff 0 <= i < bins at index int(i)
// x0: the caller's re-zeroing `(times - t0).astype(double)` (efsearch.py:1131), one more separately rounded subtraction
__global__ void __launch_bounds__(256) hist1d_kernel(const double *__restrict__ x, long long n, double x0, double lo,
                                                     double delta, long long bins, uint32_t *__restrict__ H) {
    const double fb = (double)bins;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const double i = __dmul_rn(__dsub_rn(__dsub_rn(x[t], x0), lo), delta);
        if (i >= 0.0 && i < fb) atomicAdd(H + (long long)i, 1u);
    }
}

// ffa.py:352-363: out[i] = sum(in[i*r .. i*r + r - 1])
__global__ void __launch_bounds__(256) rebin_kernel(const uint32_t *__restrict__ in, long long n_out, long long r,
                                                    uint32_t *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (long long)gridDim.x * blockDim.x) {
        uint32_t s = 0;
        for (long long q = 0; q < r; q++) s += in[i * r + q];
        out[i] = s;
    }
}

static cudaEvent_t g_ffa_ev0[64], g_ffa_ev1[64];
static bool g_ffa_ev_ok[64] = {false};
static int g_ffa_dev = -1, g_ffa_passes = 0, g_ffa_steps = 0, g_ffa_k = 0;
static int g_ffa_k_override = 0;  // tuning hook: steps per pass (0: heuristic; -1: force the per-step HBM fallback)

static size_t pass_smem(int k, int P, int nharm, bool last) {
    // 2 load buffers + 1 work buffer at a row stride of ceil4(P) counts (+ 512 bytes of slack), the harmonic tables on
    // the last pass
    return (size_t)3 * ((size_t)1 << k) * ffa_row_stride(P) * 4 + 512 +
           (last ? std::max((size_t)2 * nharm * P * 8, (size_t)((P + 3) / 4) * 32 * 8) : 0);
}

int ffa_dev_impl(const uint32_t *d_counts, int64_t n_raw, int64_t bin_period, int nharm, const double *tab,
                 double *d_stats, int64_t ntables, cudaStream_t st) {
    std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
    if (!d_counts || !tab || !d_stats) {
        set_error("hx_ffa: null pointer argument");
        return HX_ERR_ARG;
    }
    int S = 0;
    while ((1LL << S) < ntables) S++;
    if (n_raw < 0 || bin_period < 2 || bin_period > (1 << 24) || nharm < 1 || nharm > HX_MAX_NHARM || ntables < 2 ||
        (1LL << S) != ntables || ntables * bin_period < n_raw || ntables > (1LL << 40) / bin_period) {
        set_error("hx_ffa: bad sizes (n_raw=%lld bin_period=%lld nharm=%d ntables=%lld; ntables must be a power of two "
                  ">= 2 with ntables*bin_period >= n_raw)",
                  (long long)n_raw, (long long)bin_period, nharm, (long long)ntables);
        return HX_ERR_ARG;
    }
    const int P = (int)bin_period;
    int device;
    HX_CUDA_TRY(cudaGetDevice(&device));
    DeviceInfo di;
    int rc = get_device_info(device, &di);
    if (rc) return rc;

    AsyncBuf b_tab(st), b_a(st), b_b(st);
    HX_CUDA_TRY(b_tab.alloc((size_t)2 * nharm * P * sizeof(double)));
    HX_CUDA_TRY(cudaMemcpyAsync(b_tab.p, tab, (size_t)2 * nharm * P * sizeof(double), cudaMemcpyHostToDevice, st));

    // steps per pass: as few passes over HBM as shared memory allows (k_fit steps fit in one CTA), split evenly -- measured
    // on B200 (tools/quick_ffa.py): 3 passes of 7 steps at one 512-thread CTA per SM beat 4 passes of 5 at four CTAs per SM
    int k_fit = 0;
    for (int k = 1; k <= S && k <= 12; k++)
        if (pass_smem(k, P, nharm, true) <= (size_t)di.smem_optin) k_fit = k;
    int k_pref = k_fit;
    if (k_fit > 0) {
        const int npass_min = (S + k_fit - 1) / k_fit;
        k_pref = (S + npass_min - 1) / npass_min;
    }
    if (const char *e = getenv("HX_FFA_K")) g_ffa_k_override = atoi(e);
    int k_use = g_ffa_k_override > 0 ? std::min(g_ffa_k_override, k_fit) : k_pref;
    const bool fallback = g_ffa_k_override < 0 || k_use < 1;
    if (!g_ffa_ev_ok[device]) {
        HX_CUDA_TRY(cudaEventCreate(&g_ffa_ev0[device]));
        HX_CUDA_TRY(cudaEventCreate(&g_ffa_ev1[device]));
        g_ffa_ev_ok[device] = true;
    }
    const size_t level_bytes = (size_t)ntables * P * sizeof(uint32_t);
    HX_CUDA_TRY(cudaEventRecord(g_ffa_ev0[device], st));
    int passes = 0;
    if (!fallback) {
        const int npass = (S + k_use - 1) / k_use;
        if (npass > 1) HX_CUDA_TRY(b_a.alloc(level_bytes));
        if (npass > 2) HX_CUDA_TRY(b_b.alloc(level_bytes));
        HX_CUDA_TRY(cudaFuncSetAttribute(ffa_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin));
        // several small CTAs per SM overlap each other's gather / fold / write-out phases: ask for the full shared-memory
        // carve-out (the default heuristic sizes it for the first CTA only)
        HX_CUDA_TRY(cudaFuncSetAttribute(ffa_pass_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         (int)cudaSharedmemCarveoutMaxShared));
        const uint32_t *src = d_counts;
        long long n_src = n_raw;
        int s0 = 0;
        for (int p = 0; p < npass; p++) {
            const int k = (S - s0 + (npass - p) - 1) / (npass - p);  // balanced split of the remaining steps
            const bool last = p == npass - 1;
            FfaPassArgs a{};
            a.src = src, a.n_src = n_src, a.ntables = ntables, a.P = P, a.s0 = s0, a.k = k, a.nharm = nharm;
            a.tab = b_tab.as<double>(), a.stats = d_stats;
            a.vec4 = (P % 4 == 0) && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
            a.dst = last ? nullptr : ((p & 1) ? b_b.as<uint32_t>() : b_a.as<uint32_t>());
            int threads = k >= 6 ? 512 : std::min(256, std::max(32, 32 << std::min(k, 3)));
            if (const char *e = getenv("HX_FFA_THREADS")) threads = std::max(32, std::min(512, atoi(e) / 32 * 32));  // tuning hook
            const size_t smem = pass_smem(k, P, nharm, last);
            int per_sm = 0;
            HX_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ffa_pass_kernel, threads, smem));
            const long long ntiles = ntables >> k;
            const unsigned grid = (unsigned)std::min<long long>(ntiles, (long long)di.sm_count * std::max(per_sm, 1));
            if (getenv("HX_FFA_DEBUG"))
                fprintf(stderr, "[hx_ffa] pass %d: s0=%d k=%d threads=%d smem=%zu B, %d CTAs/SM, grid %u for %lld tiles\n", p, s0, k,
                        threads, smem, per_sm, grid, ntiles);
            ffa_pass_kernel<<<grid, threads, smem, st>>>(a);
            HX_CHECK_LAUNCH("ffa_pass_kernel");
            src = a.dst, n_src = ntables * (long long)P, s0 += k;
            passes++;
        }
        g_ffa_k = k_use;
    } else {
        HX_CUDA_TRY(b_a.alloc(level_bytes));
        HX_CUDA_TRY(b_b.alloc(level_bytes));
        const uint32_t *src = d_counts;
        long long n_src = n_raw;
        const unsigned grid = (unsigned)std::min<long long>((ntables * 32 + 255) / 256, (long long)di.sm_count * 16);
        for (int s = 0; s < S; s++) {
            uint32_t *dst = (s & 1) ? b_b.as<uint32_t>() : b_a.as<uint32_t>();
            ffa_step_global_kernel<<<grid, 256, 0, st>>>(src, n_src, dst, ntables, P, s);
            HX_CHECK_LAUNCH("ffa_step_global_kernel");
            src = dst, n_src = ntables * (long long)P;
            passes++;
        }
        ffa_stat_global_kernel<<<grid, 256, 0, st>>>(src, n_src, ntables, P, nharm, b_tab.as<double>(), d_stats);
        HX_CHECK_LAUNCH("ffa_stat_global_kernel");
        g_ffa_k = 0;
    }
    HX_CUDA_TRY(cudaEventRecord(g_ffa_ev1[device], st));
    g_ffa_dev = device, g_ffa_passes = passes, g_ffa_steps = S;
    HX_CUDA_TRY(cudaStreamSynchronize(st));
    return HX_OK;
}

}  // namespace hx

extern "C" {

int hx_hist1d_dev(const double *d_x, int64_t n, double x0, double lo, double hi, int64_t bins, uint32_t *d_counts,
                  void *stream) {
    std::lock_guard<std::recursive_mutex> api_lock(hx::api_mutex());
    if (!d_x || !d_counts || n < 0 || bins < 1 || !(hi > lo)) {
        hx::set_error("hx_hist1d_dev: bad arguments (n=%lld bins=%lld range=[%g, %g])", (long long)n, (long long)bins, lo, hi);
        return HX_ERR_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    HX_CUDA_TRY(cudaMemsetAsync(d_counts, 0, (size_t)bins * sizeof(uint32_t), st));
    if (n > 0) {
        const double delta = 1.0 / ((hi - lo) / (double)bins);
        hx::hist1d_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, st>>>(d_x, n, x0, lo, delta,
                                                                                                   bins, d_counts);
        HX_CHECK_LAUNCH("hist1d_kernel");
    }
    return HX_OK;
}

int hx_rebin_u32_dev(const uint32_t *d_in, int64_t n_in, int64_t rebin, uint32_t *d_out, void *stream) {
    std::lock_guard<std::recursive_mutex> api_lock(hx::api_mutex());
    if (!d_in || !d_out || n_in < 0 || rebin < 1) {
        hx::set_error("hx_rebin_u32_dev: bad arguments");
        return HX_ERR_ARG;
    }
    const long long n_out = n_in / rebin;
    if (n_out == 0) return HX_OK;
    hx::rebin_kernel<<<(unsigned)std::min<long long>((n_out + 255) / 256, 148 * 16), 256, 0,
                       reinterpret_cast<cudaStream_t>(stream)>>>(d_in, n_out, rebin, d_out);
    HX_CHECK_LAUNCH("rebin_kernel");
    return HX_OK;
}

int hx_ffa_dev(const uint32_t *d_counts, int64_t n_raw, int64_t bin_period, int nharm, const double *tab, double *d_stats,
               int64_t ntables, void *stream) {
    return hx::ffa_dev_impl(d_counts, n_raw, bin_period, nharm, tab, d_stats, ntables, reinterpret_cast<cudaStream_t>(stream));
}

int hx_ffa(const uint32_t *counts, int64_t n_raw, int64_t bin_period, int nharm, const double *tab, double *stats,
           int64_t ntables, int device) {
    if (!counts || !tab || !stats || n_raw < 0 || ntables < 1) {
        hx::set_error("hx_ffa: bad arguments");
        return HX_ERR_ARG;
    }
    hx::DeviceGuard dev_guard;  // declared before the device buffers: they are freed before the device is restored
    HX_CUDA_TRY(dev_guard.enter(device));
    hx::DevBuf dc, ds;
    HX_CUDA_TRY(dc.alloc((size_t)n_raw * sizeof(uint32_t)));
    HX_CUDA_TRY(ds.alloc((size_t)ntables * sizeof(double)));
    HX_CUDA_TRY(cudaMemcpy(dc.p, counts, (size_t)n_raw * sizeof(uint32_t), cudaMemcpyHostToDevice));
    int rc = hx::ffa_dev_impl(dc.as<uint32_t>(), n_raw, bin_period, nharm, tab, ds.as<double>(), ntables, 0);
    if (rc) return rc;
    HX_CUDA_TRY(cudaMemcpy(stats, ds.p, (size_t)ntables * sizeof(double), cudaMemcpyDeviceToHost));
    return HX_OK;
}

int hx_ffa_last_kernel_ms(double *ms, int64_t *passes, int64_t *steps, int64_t *steps_per_pass) {
    if (!ms || hx::g_ffa_dev < 0) {
        hx::set_error("hx_ffa_last_kernel_ms: no FFA has been run");
        return HX_ERR_ARG;
    }
    float f = 0.f;
    HX_CUDA_TRY(cudaEventSynchronize(hx::g_ffa_ev1[hx::g_ffa_dev]));
    HX_CUDA_TRY(cudaEventElapsedTime(&f, hx::g_ffa_ev0[hx::g_ffa_dev], hx::g_ffa_ev1[hx::g_ffa_dev]));
    *ms = f;
    if (passes) *passes = hx::g_ffa_passes;
    if (steps) *steps = hx::g_ffa_steps;
    if (steps_per_pass) *steps_per_pass = hx::g_ffa_k;
    return HX_OK;
}

int hx_ffa_set_tuning(int steps_per_pass) {
    hx::g_ffa_k_override = steps_per_pass;
    return HX_OK;
}

}  // extern "C"
