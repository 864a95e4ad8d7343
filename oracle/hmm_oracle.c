/*
 * hmm_oracle.c -- CPU restatement of the reference's Viterbi and forward recurrences.
 *
 * TEST INFRASTRUCTURE ONLY (the parity checker).  Not linked into, imported by or used as
 * a fallback of the product (advntr_b200 / libadvhmm.so).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) outputs of the unmodified reference engine compiled into oracle/_ref
 *       (oracle/build_ref.py) on seeded models/reads, and
 *   (2) the committed golden vectors in tests/golden/ (generated from that engine by
 *       tests/golden/make_golden.py), incl. the SURVEY.md section 8c sanity constants.
 *
 * The inputs are the baked arrays of pomegranate's bake() (hmm.pyx:844-1123): emitting
 * states first, silent states after `silent_start` in topological order, in-edge CSR in the
 * reference's in-edge order, emis[l*K + code] = emission log-prob + log state weight.
 *
 * Build: gcc -O2 -fPIC -shared -o oracle/libhmm_oracle.so oracle/hmm_oracle.c -lm
 *        (-ffast-math must NOT be used: the op order is the contract)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t n_states, silent_start, start_index, end_index, finite, n_symbols;
    const int32_t* in_off;
    const int32_t* in_src;
    const double* in_logp;
    const double* emis;
} oracle_model;

/* hmm.pyx:1970-2136 (_viterbi).  `path` needs n + n_states entries (hmm.pyx:1954).
 * Returns the log-probability; *path_len = number of states on the path, or -1 when the
 * read is impossible (hmm.pyx:2100-2105, viterbi() then returns (-inf, None), :1967). */
double oracle_viterbi(const oracle_model* M, const uint8_t* seq, int32_t n,
                      int32_t* path, int32_t* path_len)
{
    const int m = M->n_states, S = M->silent_start, K = M->n_symbols;
    const int32_t* off = M->in_off;
    const int32_t* src = M->in_src;
    const double* w = M->in_logp;
    size_t cells = (size_t)(n + 1) * (size_t)m;
    double* v = (double*)calloc(cells, sizeof(double));
    int32_t* tbx = (int32_t*)calloc(cells, sizeof(int32_t));   /* row of the predecessor  */
    int32_t* tby = (int32_t*)calloc(cells, sizeof(int32_t));   /* state of the predecessor */
    int i, l, k;
    *path_len = -1;

    /* row 0: only the start state is live (hmm.pyx:1999-2001) ... */
    for (l = 0; l < m; ++l) v[l] = -INFINITY;
    v[M->start_index] = 0.0;
    /* ... then silent states reachable through silent edges (hmm.pyx:2003-2023) */
    for (l = S; l < m; ++l) {
        if (l == M->start_index) continue;
        for (k = off[l]; k < off[l + 1]; ++k) {
            int ki = src[k];
            if (ki < S || ki >= l) continue;
            double cand = v[ki] + w[k];
            if (cand > v[l]) { v[l] = cand; tbx[l] = 0; tby[l] = ki; }
        }
    }

    for (i = 0; i < n; ++i) {
        const double* prev = v + (size_t)i * m;
        double* cur = v + (size_t)(i + 1) * m;
        int32_t* cx = tbx + (size_t)(i + 1) * m;
        int32_t* cy = tby + (size_t)(i + 1) * m;
        int sym = seq[i];
        /* emitting states consume symbol i from row i (hmm.pyx:2026-2042):
         * candidate = (v + t) + e, evaluated left to right, first strict maximum wins */
        for (l = 0; l < S; ++l) {
            double e = M->emis[(size_t)l * K + sym];
            cur[l] = -INFINITY;
            for (k = off[l]; k < off[l + 1]; ++k) {
                double cand = prev[src[k]] + w[k] + e;
                if (cand > cur[l]) { cur[l] = cand; cx[l] = i; cy[l] = src[k]; }
            }
        }
        /* silent states, pass 1: emitting sources of the same row (hmm.pyx:2044-2063) */
        for (l = S; l < m; ++l) {
            cur[l] = -INFINITY;
            for (k = off[l]; k < off[l + 1]; ++k) {
                int ki = src[k];
                if (ki >= S) continue;
                double cand = cur[ki] + w[k];
                if (cand > cur[l]) { cur[l] = cand; cx[l] = i + 1; cy[l] = ki; }
            }
        }
        /* silent states, pass 2: earlier silent states of the same row (hmm.pyx:2065-2083) */
        for (l = S; l < m; ++l) {
            for (k = off[l]; k < off[l + 1]; ++k) {
                int ki = src[k];
                if (ki < S || ki >= l) continue;
                double cand = cur[ki] + w[k];
                if (cand > cur[l]) { cur[l] = cand; cx[l] = i + 1; cy[l] = ki; }
            }
        }
    }

    /* termination (hmm.pyx:2089-2098) */
    double logp;
    int end;
    if (M->finite == 1) {
        end = M->end_index;
        logp = v[(size_t)n * m + end];
    } else {
        end = -1;
        logp = -INFINITY;
        for (l = 0; l < m; ++l)
            if (v[(size_t)n * m + l] > logp) { logp = v[(size_t)n * m + l]; end = l; }
    }

    if (logp != -INFINITY) {
        /* traceback to (0, start), then reverse (hmm.pyx:2107-2130) */
        int px = n, py = end, len = 0;
        while (px != 0 || py != M->start_index) {
            path[len++] = py;
            size_t c = (size_t)px * m + py;
            px = tbx[c];
            py = tby[c];
        }
        path[len++] = py;
        for (i = 0; i < len / 2; ++i) {
            int32_t t = path[i]; path[i] = path[len - 1 - i]; path[len - 1 - i] = t;
        }
        *path_len = len;
    }
    free(v); free(tbx); free(tby);
    return logp;
}

/* fp32 twin of oracle_viterbi: the same recurrence with every table rounded to float once and
 * all arithmetic in float (same operation order).  Checker of the engine's optional ADVHMM_FP32
 * mode; the reference itself has no fp32 path. */
double oracle_viterbi_f32(const oracle_model* M, const uint8_t* seq, int32_t n,
                          int32_t* path, int32_t* path_len)
{
    const int m = M->n_states, S = M->silent_start, K = M->n_symbols;
    const int32_t* off = M->in_off;
    const int32_t* src = M->in_src;
    const int E = off[m];
    float* w = (float*)malloc(sizeof(float) * (size_t)(E > 0 ? E : 1));
    float* em = (float*)malloc(sizeof(float) * (size_t)(S * K > 0 ? S * K : 1));
    size_t cells = (size_t)(n + 1) * (size_t)m;
    float* v = (float*)calloc(cells, sizeof(float));
    int32_t* tbx = (int32_t*)calloc(cells, sizeof(int32_t));
    int32_t* tby = (int32_t*)calloc(cells, sizeof(int32_t));
    int i, l, k;
    for (k = 0; k < E; ++k) w[k] = (float)M->in_logp[k];
    for (k = 0; k < S * K; ++k) em[k] = (float)M->emis[k];
    *path_len = -1;
    for (l = 0; l < m; ++l) v[l] = -INFINITY;
    v[M->start_index] = 0.0f;
    for (l = S; l < m; ++l) {
        if (l == M->start_index) continue;
        for (k = off[l]; k < off[l + 1]; ++k) {
            int ki = src[k];
            if (ki < S || ki >= l) continue;
            float cand = v[ki] + w[k];
            if (cand > v[l]) { v[l] = cand; tbx[l] = 0; tby[l] = ki; }
        }
    }
    for (i = 0; i < n; ++i) {
        const float* prev = v + (size_t)i * m;
        float* cur = v + (size_t)(i + 1) * m;
        int32_t* cx = tbx + (size_t)(i + 1) * m;
        int32_t* cy = tby + (size_t)(i + 1) * m;
        int sym = seq[i];
        for (l = 0; l < S; ++l) {
            float e = em[(size_t)l * K + sym];
            cur[l] = -INFINITY;
            for (k = off[l]; k < off[l + 1]; ++k) {
                float t = prev[src[k]] + w[k];
                float cand = t + e;
                if (cand > cur[l]) { cur[l] = cand; cx[l] = i; cy[l] = src[k]; }
            }
        }
        for (l = S; l < m; ++l) {
            cur[l] = -INFINITY;
            for (k = off[l]; k < off[l + 1]; ++k) {
                int ki = src[k];
                if (ki >= S) continue;
                float cand = cur[ki] + w[k];
                if (cand > cur[l]) { cur[l] = cand; cx[l] = i + 1; cy[l] = ki; }
            }
        }
        for (l = S; l < m; ++l) {
            for (k = off[l]; k < off[l + 1]; ++k) {
                int ki = src[k];
                if (ki < S || ki >= l) continue;
                float cand = cur[ki] + w[k];
                if (cand > cur[l]) { cur[l] = cand; cx[l] = i + 1; cy[l] = ki; }
            }
        }
    }
    float logp = v[(size_t)n * m + M->end_index];
    int end = M->end_index;
    if (logp != -INFINITY) {
        int px = n, py = end, len = 0;
        while (px != 0 || py != M->start_index) {
            path[len++] = py;
            size_t c = (size_t)px * m + py;
            px = tbx[c];
            py = tby[c];
        }
        path[len++] = py;
        for (i = 0; i < len / 2; ++i) {
            int32_t t = path[i]; path[i] = path[len - 1 - i]; path[len - 1 - i] = t;
        }
        *path_len = len;
    }
    free(v); free(tbx); free(tby); free(w); free(em);
    return (double)logp;
}

/* utils.pyx:72-90 (pair_lse) */
static double pair_lse(double x, double y)
{
    if (x == INFINITY || y == INFINITY) return INFINITY;
    if (x == -INFINITY) return y;
    if (y == -INFINITY) return x;
    if (x > y) return x + log(exp(y - x) + 1);
    return y + log(exp(x - y) + 1);
}

/* hmm.pyx:1371-1484 (_forward) + :1300-1313 (_vl_log_probability).  Two rows suffice. */
double oracle_log_probability(const oracle_model* M, const uint8_t* seq, int32_t n)
{
    const int m = M->n_states, S = M->silent_start, K = M->n_symbols;
    const int32_t* off = M->in_off;
    const int32_t* src = M->in_src;
    const double* w = M->in_logp;
    double* a = (double*)malloc(sizeof(double) * m);
    double* b = (double*)malloc(sizeof(double) * m);
    int i, l, k;
    for (l = 0; l < m; ++l) a[l] = -INFINITY;
    a[M->start_index] = 0.0;
    for (l = S; l < m; ++l) {                      /* hmm.pyx:1402-1424 */
        if (l == M->start_index) continue;
        double acc = -INFINITY;
        for (k = off[l]; k < off[l + 1]; ++k) {
            int ki = src[k];
            if (ki < S || ki >= l) continue;
            acc = pair_lse(acc, a[ki] + w[k]);
        }
        a[l] = acc;
    }
    for (i = 0; i < n; ++i) {
        int sym = seq[i];
        for (l = 0; l < S; ++l) {                  /* hmm.pyx:1427-1444: e added after the sum */
            double acc = -INFINITY;
            for (k = off[l]; k < off[l + 1]; ++k)
                acc = pair_lse(acc, a[src[k]] + w[k]);
            b[l] = acc + M->emis[(size_t)l * K + sym];
        }
        for (l = S; l < m; ++l) {                  /* hmm.pyx:1446-1461 */
            double acc = -INFINITY;
            for (k = off[l]; k < off[l + 1]; ++k) {
                int ki = src[k];
                if (ki >= S) continue;
                acc = pair_lse(acc, b[ki] + w[k]);
            }
            b[l] = acc;
        }
        for (l = S; l < m; ++l) {                  /* hmm.pyx:1463-1480 */
            double acc = -INFINITY;
            for (k = off[l]; k < off[l + 1]; ++k) {
                int ki = src[k];
                if (ki < S || ki >= l) continue;
                acc = pair_lse(acc, b[ki] + w[k]);
            }
            b[l] = pair_lse(b[l], acc);
        }
        double* t = a; a = b; b = t;
    }
    double logp;
    if (M->finite == 1) {
        logp = a[M->end_index];
    } else {
        logp = -INFINITY;
        for (l = 0; l < S; ++l) logp = pair_lse(logp, a[l]);
    }
    free(a); free(b);
    return logp;
}

/* Batch drivers used by the tests and by bench.py's cpu_baseline leg ("port"). */
void oracle_viterbi_batch(const oracle_model* M, const uint8_t* seqs, const int64_t* seq_off,
                          int32_t n_reads, double* logp, int32_t* path_len,
                          int32_t* paths, int64_t path_stride)
{
    for (int r = 0; r < n_reads; ++r) {
        int32_t n = (int32_t)(seq_off[r + 1] - seq_off[r]);
        int32_t* p = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + M->n_states));
        int32_t len;
        logp[r] = oracle_viterbi(M, seqs + seq_off[r], n, p, &len);
        path_len[r] = len;
        if (paths && len > 0)
            memcpy(paths + (size_t)r * path_stride, p,
                   sizeof(int32_t) * (size_t)(len < path_stride ? len : path_stride));
        free(p);
    }
}

void oracle_viterbi_f32_batch(const oracle_model* M, const uint8_t* seqs, const int64_t* seq_off,
                              int32_t n_reads, double* logp, int32_t* path_len,
                              int32_t* paths, int64_t path_stride)
{
    for (int r = 0; r < n_reads; ++r) {
        int32_t n = (int32_t)(seq_off[r + 1] - seq_off[r]);
        int32_t* p = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + M->n_states));
        int32_t len;
        logp[r] = oracle_viterbi_f32(M, seqs + seq_off[r], n, p, &len);
        path_len[r] = len;
        if (paths && len > 0)
            memcpy(paths + (size_t)r * path_stride, p,
                   sizeof(int32_t) * (size_t)(len < path_stride ? len : path_stride));
        free(p);
    }
}

void oracle_log_probability_batch(const oracle_model* M, const uint8_t* seqs,
                                  const int64_t* seq_off, int32_t n_reads, double* logp)
{
    for (int r = 0; r < n_reads; ++r)
        logp[r] = oracle_log_probability(M, seqs + seq_off[r],
                                         (int32_t)(seq_off[r + 1] - seq_off[r]));
}
