/*
 * CPU ORACLE (plain C restatement) for the HENDRICS pulsar periodicity-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY: built by oracle/Makefile into oracle/libhxoracle.so and loaded
 * only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  The product never links or calls it.
 *
 * PARITY STATUS: parity unpinned at the stingray boundary (see oracle/hx_oracle.py header);
 * this file is itself checked bit-for-bit against oracle/hx_oracle.py in tests/test_oracle.py.
 *
 * Follows (relative to /root/reference):
 *   stingray.pulse.search._pulse_phase_fast/_profile_fast/_folding_search  [UPSTREAM-RECALLED,
 *       SURVEY.md Appendix A; imported at hendrics/efsearch.py:15-18]
 *   stingray.pulse.pulsar.ef_profile_stat / z_n_binned_events  == hendrics/ffa.py:139-157
 *   hendrics/efsearch.py:854-873  _fast_phase*
 *   hendrics/base.py:667-677      _hist2d_numba_seq
 *   hendrics/efsearch.py:350-363  shift_and_sum
 *   hendrics/efsearch.py:817-851  _fast_step
 *   hendrics/ffa.py:59-104        _z_n_fast_cached
 *
 * Build flags matter: -ffp-contract=off (the reference never fuses multiply-add) and no
 * -ffast-math.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- minimal pthread "parallel for" with a shared atomic work counter (libgomp is not in
 * this image) ------------------------------------------------------------------------- */
typedef void (*ora_body_fn)(int64_t begin, int64_t end, void *ctx);
typedef struct {
    ora_body_fn body;
    void *ctx;
    int64_t n, chunk;
    int64_t next; /* accessed with __atomic builtins */
} ora_pf_t;

static void *ora_pf_worker(void *arg) {
    ora_pf_t *pf = (ora_pf_t *)arg;
    for (;;) {
        int64_t b = __atomic_fetch_add(&pf->next, pf->chunk, __ATOMIC_RELAXED);
        if (b >= pf->n) break;
        int64_t e = b + pf->chunk < pf->n ? b + pf->chunk : pf->n;
        pf->body(b, e, pf->ctx);
    }
    return NULL;
}

int ora_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void ora_parallel_for(int64_t n, int64_t chunk, int nthreads, ora_body_fn body, void *ctx) {
    if (nthreads <= 0) nthreads = ora_max_threads();
    if (nthreads > 256) nthreads = 256;
    ora_pf_t pf = {body, ctx, n, chunk < 1 ? 1 : chunk, 0};
    if (nthreads == 1) {
        ora_pf_worker(&pf);
        return;
    }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; i++)
        if (pthread_create(&th[started], NULL, ora_pf_worker, &pf) == 0) started++;
    ora_pf_worker(&pf);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
}

static void ora_tables(int nbin, double *ctab, double *stab) {
    /* hendrics/ffa.py:40,51: 2*pi * arange(0, 1, 1/nbin) */
    double step = 1.0 / nbin;
    for (int j = 0; j < nbin; j++) {
        double tw = (M_PI * 2) * (j * step);
        ctab[j] = cos(tw);
        stab[j] = sin(tw);
    }
}

/* stingray ef_profile_stat (nharm == 0) or z_n_binned_events (nharm >= 1) */
static double ora_stat(const int64_t *prof, int nbin, int nharm, const double *ctab, const double *stab) {
    if (nharm == 0) {
        double mean = 0.0;
        for (int j = 0; j < nbin; j++) mean += (double)prof[j];
        mean /= nbin;
        double chi = 0.0;
        for (int j = 0; j < nbin; j++) {
            double d = (double)prof[j] - mean;
            chi += d * d / mean;
        }
        return chi;
    }
    double tot = 0.0;
    for (int j = 0; j < nbin; j++) tot += (double)prof[j];
    double res = 0.0;
    for (int k = 1; k <= nharm; k++) {
        double c = 0.0, s = 0.0;
        for (int j = 0; j < nbin; j++) {
            int idx = (int)(((int64_t)j * k) % nbin);
            c += ctab[idx] * (double)prof[j];
            s += stab[idx] * (double)prof[j];
        }
        res += c * c + s * s;
    }
    return 2.0 / tot * res;
}

/* _pulse_phase_fast + _profile_fast for one trial */
static void ora_fold_one(const double *ts, int64_t n, double f, double fd, int nbin, int64_t *prof) {
    memset(prof, 0, sizeof(int64_t) * (size_t)nbin);
    for (int64_t i = 0; i < n; i++) {
        double t = ts[i];
        double ph = t * f + 0.5 * t * t * fd; /* ((0.5*t)*t)*fd, separately rounded */
        ph -= floor(ph);
        int64_t b = (int64_t)floor(ph * nbin);
        if (b >= nbin) b = nbin - 1; /* A4 */
        if (b < 0) b = 0;
        prof[b] += 1;
    }
}

typedef struct {
    const double *ts;
    int64_t n_ev;
    const double *freqs;
    int64_t n_f;
    const double *fdots;
    int nbin, nharm;
    double *stats;
    int64_t *profiles;
    const double *ctab, *stab;
} ora_fold_ctx_t;

static void ora_fold_body(int64_t begin, int64_t end, void *vctx) {
    ora_fold_ctx_t *c = (ora_fold_ctx_t *)vctx;
    int64_t *prof = (int64_t *)malloc(sizeof(int64_t) * (size_t)c->nbin);
    for (int64_t g = begin; g < end; g++) {
        int64_t i = g / c->n_f, j = g - i * c->n_f;
        ora_fold_one(c->ts, c->n_ev, c->freqs[j], c->fdots[i], c->nbin, prof);
        c->stats[g] += ora_stat(prof, c->nbin, c->nharm, c->ctab, c->stab);
        if (c->profiles)
            for (int b = 0; b < c->nbin; b++) c->profiles[g * c->nbin + b] += prof[b];
    }
    free(prof);
}

/* stats[n_fd][n_f] += stat(profile(trial)); profiles (optional) [n_fd][n_f][nbin] += profile */
int ora_fold_search(const double *ts, int64_t n_ev, const double *freqs, int64_t n_f, const double *fdots,
                    int64_t n_fd, int nbin, int nharm, double *stats, int64_t *profiles, int nthreads) {
    double *ctab = (double *)malloc(sizeof(double) * (size_t)nbin);
    double *stab = (double *)malloc(sizeof(double) * (size_t)nbin);
    if (!ctab || !stab) return -1;
    ora_tables(nbin, ctab, stab);
    int64_t ntr = n_f * n_fd;
    ora_fold_ctx_t ctx = {ts, n_ev, freqs, n_f, fdots, nbin, nharm, stats, profiles, ctab, stab};
    ora_parallel_for(ntr, 4, nthreads, ora_fold_body, &ctx);
    free(ctab);
    free(stab);
    return 0;
}

/* ---- --fast path -------------------------------------------------------------------- */

/* _fast_phase / _fast_phase_fdot / _fast_phase_fddot, selected like search_with_qffa_step
 * (efsearch.py:892-897) */
static inline double ora_fast_phase(double t, double f, double fd, double fdd) {
    double ph;
    if (fdd != 0) {
        double tssq = t * t;
        ph = t * f + 0.5 * tssq * fd + (1.0 / 6) * tssq * t * fdd;
    } else if (fd != 0) {
        ph = t * f + 0.5 * t * t * fd;
    } else {
        ph = t * f;
    }
    return ph - floor(ph);
}

/* phases + hist2d_numba_seq(...).T : profiles[nprof][nbin] (uint64), serial like the reference */
int ora_qffa_profiles(const double *ts, int64_t n_ev, double mean_f, double mean_fd, double mean_fdd, int nbin,
                      int nprof, uint64_t *profiles) {
    memset(profiles, 0, sizeof(uint64_t) * (size_t)nbin * (size_t)nprof);
    double r00 = 0.0, r01 = 1.0, r10 = ts[0], r11 = ts[n_ev - 1];
    double delta0 = 1 / ((r01 - r00) / nbin);
    double delta1 = 1 / ((r11 - r10) / nprof);
    for (int64_t t = 0; t < n_ev; t++) {
        double x = ora_fast_phase(ts[t], mean_f, mean_fd, mean_fdd);
        double i = (x - r00) * delta0;
        double j = (ts[t] - r10) * delta1;
        if (0 <= i && i < nbin && 0 <= j && j < nprof) profiles[(int64_t)j * nbin + (int64_t)i] += 1;
    }
    return 0;
}

typedef struct {
    const uint64_t *profiles;
    int nprof, nbin;
    const double *L, *Q;
    int nharm;
    double *stats;
    const double *ctab, *stab, *base;
} ora_step_ctx_t;

static void ora_step_body(int64_t begin, int64_t end, void *vctx) {
    ora_step_ctx_t *x = (ora_step_ctx_t *)vctx;
    int nbin = x->nbin;
    double *splat = (double *)malloc(sizeof(double) * (size_t)nbin);
    for (int64_t c = begin; c < end; c++) {
        for (int b = 0; b < nbin; b++) splat[b] = 0.0;
        for (int k = 0; k < x->nprof; k++) {
            double total_shift = x->base[k] * x->L[c] + (x->base[k] * x->base[k]) * x->Q[c];
            double r = rint(total_shift); /* np.rint: half to even */
            double m = fmod(r, (double)nbin);
            if (m < 0) m += nbin; /* np.mod: result in [0, nbin) */
            int s = (int)m;
            const uint64_t *row = x->profiles + (int64_t)k * nbin;
            /* splat[b] += repeated[k, nbin - s + b] == row[(b - s) mod nbin] */
            for (int b = 0; b < nbin; b++) {
                int src = b - s;
                if (src < 0) src += nbin;
                splat[b] += (double)row[src];
            }
        }
        double tot = 0.0;
        for (int b = 0; b < nbin; b++) tot += splat[b];
        double res = 0.0;
        for (int k = 1; k <= x->nharm; k++) {
            double cc = 0.0, ss = 0.0;
            for (int b = 0; b < nbin; b++) {
                int idx = (int)(((int64_t)b * k) % nbin);
                cc += x->ctab[idx] * splat[b];
                ss += x->stab[idx] * splat[b];
            }
            res += cc * cc + ss * ss;
        }
        x->stats[c] = 2.0 / tot * res;
    }
    free(splat);
}

/* _fast_step on profiles[nprof][nbin]; L,Q [nL*nQ] row-major; stats [nL*nQ] */
int ora_fast_step(const uint64_t *profiles, int nprof, int nbin, const double *L, const double *Q, int64_t nLQ,
                  int nharm, double *stats, int nthreads) {
    double *ctab = (double *)malloc(sizeof(double) * (size_t)nbin);
    double *stab = (double *)malloc(sizeof(double) * (size_t)nbin);
    double *base = (double *)malloc(sizeof(double) * (size_t)nprof);
    if (!ctab || !stab || !base) return -1;
    /* _fast_step builds its table from np.arange(0, 1, 1/nbin) too (efsearch.py:819) */
    ora_tables(nbin, ctab, stab);
    /* np.linspace(-1, 1, nprof): start + i*step, last element forced to stop */
    {
        double step = (nprof > 1) ? 2.0 / (nprof - 1) : 0.0;
        for (int k = 0; k < nprof; k++) base[k] = -1.0 + k * step;
        if (nprof > 1) base[nprof - 1] = 1.0;
    }
    ora_step_ctx_t ctx = {profiles, nprof, nbin, L, Q, nharm, stats, ctab, stab, base};
    ora_parallel_for(nLQ, 1, nthreads, ora_step_body, &ctx);
    free(ctab);
    free(stab);
    free(base);
    return 0;
}
