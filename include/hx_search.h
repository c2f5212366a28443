/*
 * hx_search.h -- C ABI of libhxsearch.so, the B200 (sm_100a) pulsar periodicity-search library.
 *
 * Drop-in boundary for ONE hot path of StingraySoftware/HENDRICS: the (f, fdot) trial-grid
 * Z^2_n / epoch-folding search behind HENzsearch / HENefsearch, including `--fast`.
 * The reference is pure Python and has no FFI for this path; each entry
 * point below cites the reference interface (relative to the HENDRICS tree) that a maintainer
 * would rebind to it -- the ctypes stubs are in INTEGRATION.md and hendrics_b200/_cabi.py.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every buffer; inputs are never modified;
 *   - every function returns HX_OK (0) or a negative HX_ERR_* code; hx_last_error() returns a
 *     thread-local human-readable message for the last failure;
 *   - functions without a `_dev` suffix take HOST pointers and perform H2D copy, kernels and D2H
 *     copy synchronously on `device`;
 *   - `_dev` functions take DEVICE pointers for the bulk data (resident on the current device),
 *     run on `stream` (a cudaStream_t cast to void*, NULL = default stream) and may synchronize
 *     that stream before returning (they read back a few bytes of status);
 *   - there is NO CPU fallback: without a usable CUDA device every compute call fails with
 *     HX_ERR_CUDA.
 */
#ifndef HX_SEARCH_H
#define HX_SEARCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HX_OK 0
#define HX_ERR_ARG (-1)    /* invalid argument (null pointer, nbin out of range, ...) */
#define HX_ERR_CUDA (-2)   /* CUDA runtime / launch failure, or no device */
#define HX_ERR_NOMEM (-3)  /* device or host allocation failed */
#define HX_ERR_RANGE (-4)  /* |phase| too large for fp64 folding (>= 2^51 cycles) */
#define HX_ERR_INTERNAL (-5)

#define HX_MAX_NBIN 1024 /* phase bins per profile, 2 <= nbin <= HX_MAX_NBIN */
#define HX_MAX_NHARM 64

/* flags for hx_fold_search* */
#define HX_FOLD_DEFAULT 0u
#define HX_FOLD_FORCE_GENERIC 1u /* disable the power-of-two-nbin fast path (testing) */
#define HX_FOLD_FORCE_SCRATCH 2u /* disable the fused epilogue (testing) */
#define HX_FOLD_EXACT_PHASE 4u   /* NOT the parity path: phase mod 1 from error-free products (accurate to ~1 ulp
                                    of 1.0 for any |t*f|), for long baselines where the reference's fp64 phase
                                    loses log2(t*f) bits; validated against exact rational arithmetic */

/* ---- library / device ------------------------------------------------------------------ */

/* Version of this ABI (major*1000 + minor). */
int hx_version(void);
/* Message for the last error raised on the calling thread ("" if none). */
const char *hx_last_error(void);
/* Number of CUDA devices visible (0 and HX_ERR_CUDA when there is none). */
int hx_device_count(int *count);
/* Number of this library's kernel launches so far in this process (bench evidence). */
int64_t hx_launch_count(void);

/* ---- K1: exact fold search ---------------------------------------------------------------
 * Replaces the inner loop of stingray.pulse.search._folding_search as driven by
 * z_n_search / epoch_folding_search (imported at hendrics/efsearch.py:15-18, called at
 * hendrics/efsearch.py:1179-1187): for every trial (fdots[i], freqs[j]) fold all events into an
 * nbin-bin phase histogram (phase = t*f + 0.5*t*t*fd, separately rounded fp64, no FMA) and
 * reduce it to
 *     nharm == 0 : epoch-folding chi^2   = sum_j (p_j - mean)^2 / mean
 *     nharm >= 1 : binned Z^2_nharm      (formula of hendrics/ffa.py:139-157)
 *
 *   times    [n_ev]        event times, already re-zeroed and segment-selected by the caller
 *   freqs    [n_f], fdots [n_fd]
 *   stats    [n_fd * n_f]  row-major (fdot, freq); overwritten
 *   profiles [n_fd * n_f * nbin] uint32 counts, or NULL when not wanted
 */
int hx_fold_search(const double *times, int64_t n_ev, const double *freqs, int64_t n_f, const double *fdots,
                   int64_t n_fd, int nbin, int nharm, unsigned flags, double *stats, uint32_t *profiles,
                   int device);

/* Same on device-resident buffers.  `d_times` must be 16-byte aligned. */
int hx_fold_search_dev(const double *d_times, int64_t n_ev, const double *d_freqs, int64_t n_f,
                       const double *d_fdots, int64_t n_fd, int nbin, int nharm, unsigned flags, double *d_stats,
                       uint32_t *d_profiles, void *stream);

/* ---- K1win: the exact fold search on many small time windows, one launch --------------------------
 * Replaces (a) the window loop of dyn_folding_search (hendrics/efsearch.py:1221-1231; run by default by the CLI for
 * observations longer than --dynstep, :1912-1922) and (b) the segment loop of stingray's _folding_search as driven by
 * `--segment-size` (hendrics/efsearch.py:1658-1663): per window the same per-trial fold + statistic as hx_fold_search.
 *   d_times [n_ev]              event times, SORTED ascending, device
 *   win_start, win_stop [n_win] HOST arrays: window w holds the events with win_start[w] <= x < win_stop[w], where
 *     mode 0 (segments of ONE search): x = fl(t - origin), the times re-zeroed once; a window with no event or whose
 *             events span < 0.2 * segment_size is skipped, as upstream skips it (status 0);
 *     mode 1 (independent searches):   x = t (raw times); the window is then searched like upstream searches a whole
 *             event list: re-zeroed to ITS first event, events at its last time dropped (half-open single segment),
 *             skipped when the rest spans < 0.2 of its length (status 0: upstream returns 0/0 = NaN); an empty or
 *             zero-length window makes upstream raise, for which dyn_folding_search stores zeros (status 2).
 *   d_stats [n_win * n_fd * n_f] statistics per window (NaN for status 0, 0 for status 2)
 *   d_sum   [n_fd * n_f] or NULL: sum over the folded windows in window order (upstream's `stats += stat`)
 *   h_status[n_win]             HOST array of the status codes above (1 = folded).  Synchronizes `stream`.
 * Meant for windows too small to fill the GPU one by one (reference operation sequence, events read from global
 * memory); large windows are better served by one hx_fold_search_dev call each. */
int hx_fold_windows_dev(const double *d_times, int64_t n_ev, const double *win_start, const double *win_stop,
                        int64_t n_win, int mode, double origin, double segment_size, const double *d_freqs, int64_t n_f,
                        const double *d_fdots, int64_t n_fd, int nbin, int nharm, double *d_stats, double *d_sum,
                        int *h_status, void *stream);

/* ---- K1w: weighted / exposure-corrected fold search ------------------------------------------
 * Replaces the SLOW branch of stingray.pulse.search._folding_search (use_times=True ->
 * stingray.pulse.pulsar.fold_events per trial), which z_n_search / epoch_folding_search take when
 * `weights` is an array (light-curve input, hendrics/efsearch.py:1155-1157) or `expocorr` is set
 * (`--expocorr`, hendrics/efsearch.py:1164-1167).  Per trial: phase = t*f + (0.5*t*t)*fd minus its floor,
 * weighted histogram on the edges linspace(0, 1, nbin + 1) (np.histogram: searchsorted semantics, last
 * bin closed), optional division by the normalised per-bin exposure of the GTIs
 * (stingray.pulse.pulsar.phase_exposure), then chi^2 (nharm == 0) or binned Z^2_nharm.
 *
 *   times    [n_ev]   event times (or light-curve bin times) of ONE segment, as upstream passes them
 *   weights  [n_ev]   or NULL: every event weighs `unit_weight`
 *   gti      [n_gti*2] HOST array of (start, stop) pairs in the frame of `times`; n_gti == 0: no exposure
 *            correction.  start_time / stop_time: times[0] and times[-1] + |times[1] - times[0]|
 *   stats    [n_fd * n_f]; profiles [n_fd * n_f * nbin] fp64 (exposure-corrected) or NULL
 */
int hx_wfold_search(const double *times, const double *weights, double unit_weight, int64_t n_ev, const double *freqs,
                    int64_t n_f, const double *fdots, int64_t n_fd, int nbin, int nharm, const double *gti,
                    int64_t n_gti, double start_time, double stop_time, double *stats, double *profiles, int device);
int hx_wfold_search_dev(const double *d_times, const double *d_weights, double unit_weight, int64_t n_ev,
                        const double *d_freqs, int64_t n_f, const double *d_fdots, int64_t n_fd, int nbin, int nharm,
                        const double *gti, int64_t n_gti, double start_time, double stop_time, double *d_stats,
                        double *d_profiles, void *stream);
/* Device time of the wfold_kernel launches of the most recent hx_wfold_search* call and the histogram cell type it
 * chose: 0 = fp64 cells (arbitrary weights), 1 = uint32 cells (integer weights), 2 = uint32 cells (no weight array);
 * n_events_folded = events actually streamed (zero-weight entries are dropped when they are more than a quarter). */
int hx_wfold_last_kernel_ms(double *ms, int *cell_kind, int64_t *n_events_folded);

/* ---- K2 + K3: one centre step of the `--fast` quasi-fast-folding search ------------------
 * Replaces hendrics/efsearch.py:876-925 (search_with_qffa_step) = _fast_phase* (:854-873) +
 * histogram2d (hendrics/base.py:667-693) + _fast_step (:817-851) for `n_centres` centre
 * frequencies in one pass over the events.
 *
 *   times_c   [n_ev]  times minus the mid-time, fp64, sorted ascending
 *   mean_f/mean_fd/mean_fdd [n_centres]
 *   L, Q      [n_lq]  bin-shift pairs (the meshgrid of efsearch.py:921, flattened row-major)
 *   stats     [n_centres * n_lq]
 *   profiles  [n_centres * nprof * nbin] uint32 2-D profiles (time slice, phase bin), or NULL
 *
 * With profiles == NULL and nharm <= 4 (sorted times, power-of-two nbin) the 2-D profiles never leave the SM:
 * every finished time slice is reduced to its 2 * nharm harmonic coefficients there (the only thing _fast_step
 * needs of it), so the call needs 8 * nprof * (2 * nharm + 1) bytes of scratch per centre instead of
 * 4 * nprof * nbin.  Statistics are the same to fp64 summation order (1e-12); with profiles != NULL the
 * counts are materialised and bit-exact.
 * hx_qffa_search (host array in) uploads the events itself; on that fused path and for >= 2^22 events it does
 * so in chunks on a copy stream WHILE the kernel folds the chunks already in place (page-locked `times_c` gives
 * the best overlap; pageable works).  The times must be sorted for this (verified on the device; unsorted
 * input is still handled, by the slower lane = event kernel).
 */
int hx_qffa_search(const double *times_c, int64_t n_ev, const double *mean_f, const double *mean_fd,
                   const double *mean_fdd, int64_t n_centres, int nbin, int nprof, const double *L, const double *Q,
                   int64_t n_lq, int nharm, double *stats, uint32_t *profiles, int device);

/* Device variant: only the event times and the outputs are device pointers; the per-centre
 * scalars and the shift grids are small HOST arrays.  t_first/t_last are times_c[0] and
 * times_c[n_ev-1] (the histogram range of efsearch.py:902).  Synchronizes `stream`. */
int hx_qffa_search_dev(const double *d_times_c, int64_t n_ev, double t_first, double t_last, const double *mean_f,
                       const double *mean_fd, const double *mean_fdd, int64_t n_centres, int nbin, int nprof,
                       const double *L, const double *Q, int64_t n_lq, int nharm, double *d_stats,
                       uint32_t *d_profiles, void *stream);

/* ---- transient search (widening) ----------------------------------------
 * Replaces hendrics/efsearch.py:468-485 (_transient_search_step) = _fast_phase_fdot + histogram2d +
 * _average_and_z_sub_search (:407-465) for `n_centres` centre frequencies: for every averaging level
 * l < n_levels (groups of 2^l consecutive time slices) the Z^2_nharm of each group's summed profile.
 *   d_stats [n_centres * n_levels * nprof_used]   (centre, level, time slice)
 * The events are histogrammed into nprof time slices; the averaging uses the first nprof_used of them:
 * nprof_used = int(2 ** np.log2(nprof)) and n_levels = int(np.log2(nprof_used)) in the reference (:436-441) -- the
 * floating-point round trip drops the last slice for some non power-of-two nprof (10 -> 9, 100 -> 99), which the
 * caller reproduces with numpy.  Same conventions as hx_qffa_search_dev. */
int hx_transient_search_dev(const double *d_times_c, int64_t n_ev, double t_first, double t_last,
                            const double *mean_f, const double *mean_fd, int64_t n_centres, int nbin, int nprof,
                            int nprof_used, int nharm, int n_levels, double *d_stats, void *stream);

/* ---- orbital-demodulation fold (widening) -------------------------------
 * Replaces the T0 loop of hendrics/efsearch.py:294-302 (folding_orbital_search): for every trial epoch of ascending
 * node T0[k], demodulate the event times twice (:295-296), then run the fast branch of
 * epoch_folding_search / z_n_search on them for the single trial (freq, fdot) (:297).
 *   times [n_ev] (as folding_orbital_search computes them: events.time - gti[0, 0]);  stats [n_t0];
 *   profiles [n_t0 * nbin] uint32 or NULL.  nharm == 0: epoch-folding chi^2, else Z^2_nharm.
 * sin() is CUDA's, not numpy's; measured: profiles identical to the reference arithmetic on all test cases
 * (profiles/r02_orbital_diff.txt; a last-bit difference of sin almost never reaches a bin edge). */
int hx_orbital_search(const double *times, int64_t n_ev, double freq, double fdot, double X, double Porb,
                      const double *T0, int64_t n_t0, int nbin, int nharm, double *stats, uint32_t *profiles,
                      int device);
int hx_orbital_search_dev(const double *d_times, int64_t n_ev, double freq, double fdot, double X, double Porb,
                          const double *d_T0, int64_t n_t0, int nbin, int nharm, double *d_stats, uint32_t *d_profiles,
                          void *stream);

/* ---- K4: unbinned Z^2_n (extension; oracle formula hendrics/efsearch.py:366-404) ---------
 * stats[i*n_f + j] = 2/N * sum_{k=1..nharm} [(sum cos 2 pi k phi)^2 + (sum sin 2 pi k phi)^2]
 * with phi = frac(t*f + 0.5*t*t*fd) per event, the phase rounded like the reference's fp64 phase
 * (_pulse_phase_fast), one sin/cos per event*trial and the higher harmonics by angle addition.
 * Limits: 1 <= nharm <= 8 (per-trial sums live in registers); |phase| < 2^48 cycles (HX_ERR_RANGE);
 * non-finite inputs give HX_ERR_ARG; d_times 16-byte aligned.  Results are bit-reproducible from run
 * to run (no floating-point atomics). */
int hx_zn_unbinned(const double *times, int64_t n_ev, const double *freqs, int64_t n_f, const double *fdots,
                   int64_t n_fd, int nharm, double *stats, int device);
int hx_zn_unbinned_dev(const double *d_times, int64_t n_ev, const double *d_freqs, int64_t n_f,
                       const double *d_fdots, int64_t n_fd, int nharm, double *d_stats, void *stream);
/* measurement only: duration (ms) of the last K4 launch on the current device, and the FP64
 * instructions it issues per event*trial */
int hx_zn_last_kernel_ms(double *ms, int *fp64_ops_per_unit);

/* ---- top-k candidates ----------------------------------------------------------------------
 * Replaces the argpartition branch of hendrics/base.py:1140-1146 (find_peaks_in_image):
 * the k largest entries of stats[n], value descending, ties by ascending flat index.
 * idx/val have room for k entries; *n_out = min(k, n).
 */
int hx_topk(const double *stats, int64_t n, int k, int64_t *idx, double *val, int64_t *n_out, int device);
int hx_topk_dev(const double *d_stats, int64_t n, int k, int64_t *h_idx, double *h_val, int64_t *n_out,
                void *stream); /* synchronizes `stream`; idx/val are HOST buffers */

/* ---- fast folding algorithm (--ffa) ---------------------------------------------------------
 * hx_hist1d_dev replaces hendrics/base.py:621-660 (`histogram` / `hist1d_numba_seq`, unweighted) as called by
 * `search_with_ffa` (hendrics/efsearch.py:1131-1136): with v = x - x0 (the caller's re-zeroing `times - t0`, a separately
 * rounded fp64 subtraction; pass x0 = 0 for none), counts[int((v - lo) * delta)] += 1 for 0 <= (v - lo) * delta < bins,
 * delta = 1 / ((hi - lo) / bins) -- right edge excluded.  d_counts[bins] is overwritten.
 * hx_rebin_u32_dev replaces `_quick_rebin` (hendrics/ffa.py:352-363): d_out[i] = sum(d_in[i*rebin .. i*rebin+rebin-1]),
 * i < n_in / rebin.
 * hx_ffa[_dev] replaces `ffa` / `_ffa` / `ffa_step` (hendrics/ffa.py:291-349): the light curve counts[n_raw] is
 * zero-padded to ntables x bin_period (ntables = 2^ceil(log2(n_raw / bin_period + 1)), computed by the caller like
 * ffa.py:340), folded in log2(ntables) steps and stats[i] = Z^2_nharm of table i (trial period bin_period +
 * i / (ntables - 1) bins); an empty light curve gives NaN (the reference's numba code raises ZeroDivisionError).  tab is a HOST array [2][nharm][bin_period]:
 * the cos (tab[0]) and sin (tab[1]) value harmonic h+1 multiplies bin j with, i.e. the stride-(h+1) slices of the
 * reference's tiled tables (ffa.py:98-102, 320-328; the caller builds them with numpy exactly as the reference does,
 * including the off-by-one length of np.arange(0, 1, 1/bin_period) for some periods).  Counts are uint32 (the
 * reference adds integer-valued doubles: same numbers).  The _dev variants take device pointers and a CUDA stream;
 * hx_ffa_dev synchronizes the stream.
 */
int hx_hist1d_dev(const double *d_x, int64_t n, double x0, double lo, double hi, int64_t bins, uint32_t *d_counts,
                  void *stream);
int hx_rebin_u32_dev(const uint32_t *d_in, int64_t n_in, int64_t rebin, uint32_t *d_out, void *stream);
int hx_ffa(const uint32_t *counts, int64_t n_raw, int64_t bin_period, int nharm, const double *tab, double *stats,
           int64_t ntables, int device);
int hx_ffa_dev(const uint32_t *d_counts, int64_t n_raw, int64_t bin_period, int nharm, const double *tab,
               double *d_stats, int64_t ntables, void *stream);

/* ---- measurement helpers --------------------------------------------------------------------
 * FP64 pipe issue-rate microbenchmarks used as the roofline denominator:
 *   kind 0: dependent-free DFMA chain      -> *ops_per_s = FP64 instructions/s (1 per DFMA)
 *   kind 1: DMUL + DADD pairs (no FMA)     -> FP64 instructions/s
 *   kind 2: shared-memory LDS+IADD+STS read-modify-write, conflict-free -> RMW/s
 *   kind 3: shared-memory atomic add, conflict-free                     -> atomics/s
 *   kind 4: the (f, fdot) inner-loop mix (2 DMUL + 2 DADD + LOP3 + ATOMS), no other memory traffic
 *           -> units/s: an empirical ceiling for K1 in event*trials/s
 *   kind 5: the f-only mix (DMUL + DADD + LOP3 + ATOMS) -> units/s
 *   kind 6: F2I.S64.F64.FLOOR issue rate -> conversions/s
 *   kinds 7..10: the kind-4 mix with 2, 4, 6, 8 of every 8 floors taken through F2I (XU) instead of the magic add
 *   kinds 11, 12: K1's (f, fdot) loop at K1's occupancy (12 warps/SM) with the fdot term computed per thread (11) or
 *           read from a per-CTA shared-memory tile (12) -> units/s   (experiments recorded in DESIGN.md)
 *   kinds 13, 14: K1's f-only loop with the events broadcast from shared memory (13) or __constant__ memory (14)
 *   kinds 15, 16: as 11, 12 with 4 trials per thread at 6 warps/SM
 * Runs on `device`, blocks, returns the best of `reps` timings.
 */
int hx_microbench(int kind, int reps, double *ops_per_s, int device);

/* Device time (CUDA events on the launching stream) and launch shape of the most recent K1
 * fold_kernel launch of this process. */
int hx_fold_last_kernel_ms(double *ms, int64_t *grid, int64_t *block, int64_t *smem_bytes);

/* Instantiation of that launch: trials per thread R, histogram mode (0 generic, 1 power-of-two, 2 packed, 3 exact
 * phase), whether the R trials of a thread share their fdot (rowfd) and whether the fdot term comes from the per-CTA
 * tile (m2t); *fp64_per_unit = FP64 instructions the inner loop issues per event*trial (-1 for the generic / exact
 * modes, whose count depends on the data), *smem_wavefronts_per_unit = shared-memory data-pipe wavefronts per
 * event*trial and warp (one per increment + the broadcast event loads).  bench.py's executed-work figures. */
int hx_fold_last_kernel_shape(int *R, int *mode, int *rowfd, int *m2t, double *fp64_per_unit,
                              double *smem_wavefronts_per_unit);

/* Summed device time of the K2 (qffa_hist_kernel) launches of the most recent hx_qffa_search* call
 * and their number. */
int hx_qffa_last_kernel_ms(double *ms, int64_t *launches);

/* Device time of the most recent hx_ffa* call (all passes + statistics), the number of passes over HBM, of folding
 * steps, and the steps fused per pass (0: per-step fallback). */
int hx_ffa_last_kernel_ms(double *ms, int64_t *passes, int64_t *steps, int64_t *steps_per_pass);

/* ---- tuning hooks (tests / benchmarks only; 0 restores the built-in heuristics) ------------- */
int hx_ffa_set_tuning(int steps_per_pass); /* > 0: steps fused per pass; -1: one step per launch through HBM */
int hx_fold_set_tuning(int trials_per_thread, int warps_per_cta);
int hx_qffa_set_tuning(int centres_per_pass);
/* force_generic bit 0: never use the pre-scaled fast bin path; bit 1: fp64 cells even for integer weights;
 * bit 2: never drop zero-weight entries */
int hx_wfold_set_tuning(int trials_per_thread, int events_per_group, int warps_per_cta, int force_generic);
int hx_fold_set_packing(int mode); /* 0: never pack counters, 1: whenever possible, other: automatic */
int hx_fold_set_generic_fast(int on); /* 0: non power-of-two nbin always replays the reference's operation sequence */

#ifdef __cplusplus
}
#endif
#endif /* HX_SEARCH_H */
