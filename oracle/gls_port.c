/* ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C restatement of the search half of the gnngls hot path.  Each function
 * cites the reference lines it follows (paths relative to /root/reference).
 * Pinned bit-for-bit against golden vectors produced by the reference's own
 * Python (oracle/make_golden.py -> tests/golden/, tests/test_oracle_golden.py).
 *
 * Build: see oracle/Makefile.  MUST be compiled with -ffp-contract=off: every
 * fp64 expression below relies on one IEEE rounding per source-level operation,
 * in the reference's left-to-right association.
 *
 * Tours are int32[n+1] with tour[0] == tour[n] == depot.  D is row-major
 * double[n*n].
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AT(M, n, a, b) ((M)[(size_t)(a) * (size_t)(n) + (size_t)(b)])

/* numpy.isclose(0, d) with default rtol=1e-5, atol=1e-8 (operators.py:42,65,118,141):
 * |0 - d| <= atol + rtol*|d| */
static int is_close_to_zero(double d) {
    double ad = fabs(d);
    double rhs = 1e-5 * ad;
    rhs = 1e-8 + rhs;
    return ad <= rhs;
}

/* gnngls/__init__.py:17-21 — sequential sum starting from integer 0 */
double glsp_tour_cost(const double *D, int n, const int *tour) {
    double c = 0.0;
    for (int p = 0; p < n; ++p) c += AT(D, n, tour[p], tour[p + 1]);
    return c;
}

/* operators.py:14-29 */
static double two_opt_delta(const int *t, const double *D, int n, int i, int j) {
    if (i == j) return 0.0;
    if (j < i) { int s = i; i = j; j = s; }
    int a = t[i], b = t[i - 1], c = t[j], d = t[j - 1];
    double x = AT(D, n, a, c) + AT(D, n, b, d);
    x = x - AT(D, n, a, b);
    x = x - AT(D, n, c, d);
    return x;
}

/* operators.py:6-11 — reverse positions i..j-1 */
static void two_opt_apply(const int *t, int n, int i, int j, int *out) {
    if (j < i) { int s = i; i = j; j = s; }
    for (int p = 0; p <= n; ++p) out[p] = t[p];
    if (i == j) return;
    for (int p = i; p < j; ++p) out[p] = t[j - 1 - (p - i)];
}

/* operators.py:83-103 */
static double relocate_delta(const int *t, const double *D, int n, int i, int j) {
    if (i == j) return 0.0;
    int a = t[i - 1], b = t[i], c = t[i + 1], d, e;
    if (i < j) { d = t[j]; e = t[j + 1]; } else { d = t[j - 1]; e = t[j]; }
    double x = -AT(D, n, a, b);
    x = x - AT(D, n, b, c);
    x = x + AT(D, n, a, c);
    x = x - AT(D, n, d, e);
    x = x + AT(D, n, d, b);
    x = x + AT(D, n, b, e);
    return x;
}

/* operators.py:76-80 — pop(i) then insert(j) */
static void relocate_apply(const int *t, int n, int i, int j, int *out) {
    int node = t[i], q = 0;
    int tmp_len = n; /* after pop the list has n entries */
    for (int p = 0; p <= n; ++p) if (p != i) out[q++] = t[p];
    (void)tmp_len;
    for (int p = n; p > j; --p) out[p] = out[p - 1];
    out[j] = node;
}

/* All four scans share the acceptance rule of operators.py:41-46 */
#define CONSIDER(delta_expr)                                              \
    do {                                                                  \
        double dl = (delta_expr);                                         \
        if (dl < best && !is_close_to_zero(dl)) {                         \
            best = dl; bi = i; bj = j; found = 1;                         \
            if (first_improvement) goto done;                             \
        }                                                                 \
    } while (0)

/* operators.py:32-50 */
int glsp_two_opt_a2a(const int *tour, const double *D, int n, int first_improvement,
                     double *delta, int *mi, int *mj, int *out) {
    double best = 0.0; int bi = 0, bj = 0, found = 0;
    for (int i = 1; i <= n - 1; ++i)
        for (int j = i + 1; j <= n - 1; ++j) {
            if (abs(i - j) < 2) continue;
            CONSIDER(two_opt_delta(tour, D, n, i, j));
        }
done:
    *delta = found ? best : 0.0; *mi = bi; *mj = bj;
    if (found) two_opt_apply(tour, n, bi, bj, out); else memcpy(out, tour, sizeof(int) * (n + 1));
    return found;
}

/* operators.py:53-73 */
int glsp_two_opt_o2a(const int *tour, const double *D, int n, int i, int first_improvement,
                     double *delta, int *mi, int *mj, int *out) {
    double best = 0.0; int bi = 0, bj = 0, found = 0;
    if (!(i > 0 && i < n)) return -1;
    for (int j = 1; j <= n - 1; ++j) {
        if (abs(i - j) < 2) continue;
        CONSIDER(two_opt_delta(tour, D, n, i, j));
    }
done:
    *delta = found ? best : 0.0; *mi = bi; *mj = bj;
    if (found) two_opt_apply(tour, n, bi, bj, out); else memcpy(out, tour, sizeof(int) * (n + 1));
    return found;
}

/* operators.py:106-126 */
int glsp_relocate_o2a(const int *tour, const double *D, int n, int i, int first_improvement,
                      double *delta, int *mi, int *mj, int *out) {
    double best = 0.0; int bi = 0, bj = 0, found = 0;
    if (!(i > 0 && i < n)) return -1;
    for (int j = 1; j <= n - 1; ++j) {
        if (i == j) continue;
        CONSIDER(relocate_delta(tour, D, n, i, j));
    }
done:
    *delta = found ? best : 0.0; *mi = bi; *mj = bj;
    if (found) relocate_apply(tour, n, bi, bj, out); else memcpy(out, tour, sizeof(int) * (n + 1));
    return found;
}

/* operators.py:129-147 */
int glsp_relocate_a2a(const int *tour, const double *D, int n, int first_improvement,
                      double *delta, int *mi, int *mj, int *out) {
    double best = 0.0; int bi = 0, bj = 0, found = 0;
    for (int i = 1; i <= n - 1; ++i)
        for (int j = 1; j <= n - 1; ++j) {
            if (j == i) continue;
            if (i - j == 1) continue;
            CONSIDER(relocate_delta(tour, D, n, i, j));
        }
done:
    *delta = found ? best : 0.0; *mi = bi; *mj = bj;
    if (found) relocate_apply(tour, n, bi, bj, out); else memcpy(out, tour, sizeof(int) * (n + 1));
    return found;
}

typedef struct { double *ev; int n_ev, max_ev; long moves_2opt, moves_reloc, sweeps; } evlog_t;

static void log_cost(evlog_t *lg, double c) {
    if (lg && lg->ev && lg->n_ev < lg->max_ev) lg->ev[lg->n_ev] = c;
    if (lg) lg->n_ev++;
}

/* algorithms.py:111-132.  tour updated in place; scratch is int[n+1]. */
static void local_search_impl(int *tour, double *cost, const double *D, int n, int first_improvement,
                              int *scratch, evlog_t *lg) {
    int improved = 1;
    while (improved) {
        improved = 0;
        for (int op = 0; op < 2; ++op) {
            double delta; int mi, mj, found;
            if (op == 0) found = glsp_two_opt_a2a(tour, D, n, first_improvement, &delta, &mi, &mj, scratch);
            else         found = glsp_relocate_a2a(tour, D, n, first_improvement, &delta, &mi, &mj, scratch);
            if (lg) lg->sweeps++;
            if (found && delta < 0) {
                improved = 1;
                *cost = *cost + delta;
                memcpy(tour, scratch, sizeof(int) * (n + 1));
                log_cost(lg, *cost);
            }
        }
    }
}

int glsp_local_search(int *tour, double *cost, const double *D, int n, int first_improvement,
                      double *events, int *n_events, int max_events) {
    int *scratch = (int *)malloc(sizeof(int) * (n + 1));
    evlog_t lg = {events, 0, max_events, 0, 0, 0};
    local_search_impl(tour, cost, D, n, first_improvement, scratch, &lg);
    *n_events = lg.n_ev;
    free(scratch);
    return 0;
}

/* algorithms.py:9-18 with K_n adjacency in ascending node order (generate_instances.py:31-33);
 * Python's min() keeps the first minimum. */
void glsp_nearest_neighbor(const double *W, int n, int depot, int *tour) {
    char *used = (char *)calloc(n, 1);
    tour[0] = depot; used[depot] = 1;
    for (int len = 1; len < n; ++len) {
        int i = tour[len - 1], bj = -1; double bw = 0.0;
        for (int j = 0; j < n; ++j) {
            if (j == i || used[j]) continue;
            double w = AT(W, n, i, j);
            if (bj < 0 || w < bw) { bj = j; bw = w; }
        }
        tour[len] = bj; used[bj] = 1;
    }
    tour[n] = depot;
    free(used);
}

/* algorithms.py:135-195 with the wall-clock test at :146 replaced by a fixed count of
 * outer iterations (n_iters).  guides: double[n_guides][n][n] (symmetric).  pen (optional
 * out): double[n*n] final penalties. */
int glsp_guided_local_search(const double *D, const double *guides, int n_guides, int n,
                             const int *init_tour, double init_cost, int n_iters, int perturbation_moves,
                             int first_improvement, int *best_tour, double *best_cost,
                             double *events, int *n_events, int max_events, double *pen_out,
                             long *counters /* optional [3]: sweeps, o2a evals, accepted perturbation moves */) {
    size_t nn = (size_t)n * n;
    double k = 0.1 * init_cost;                       /* :137  0.1 * init_cost / len(G.nodes) */
    k = k / (double)n;
    double *pen = (double *)calloc(nn, sizeof(double)); /* :138 */
    double *Dg = (double *)malloc(nn * sizeof(double));
    int *cur = (int *)malloc(sizeof(int) * (n + 1));
    int *scratch = (int *)malloc(sizeof(int) * (n + 1));
    evlog_t lg = {events, 0, max_events, 0, 0, 0};
    long o2a = 0, pmoves = 0;
    memcpy(cur, init_tour, sizeof(int) * (n + 1));
    double cur_cost = init_cost;

    local_search_impl(cur, &cur_cost, D, n, first_improvement, scratch, &lg);   /* :142 */
    memcpy(best_tour, cur, sizeof(int) * (n + 1));                              /* :143 */
    *best_cost = cur_cost;

    for (int it = 0; it < n_iters; ++it) {                                      /* :146 (fixed count) */
        const double *guide = guides + (size_t)(it % n_guides) * nn;            /* :147 */
        int moves = 0;
        while (moves < perturbation_moves) {                                    /* :151 */
            double max_util = 0.0; int me = -1;
            for (int p = 0; p < n; ++p) {                                       /* :155-159 */
                int u = cur[p], v = cur[p + 1];
                double util = AT(guide, n, u, v) / (1.0 + AT(pen, n, u, v));
                if (util > max_util || me < 0) { max_util = util; me = p; }
            }
            int eu = cur[me], ev = cur[me + 1];
            AT(pen, n, eu, ev) += 1.0;                                          /* :161 */
            if (eu != ev) AT(pen, n, ev, eu) = AT(pen, n, eu, ev);
            for (size_t q = 0; q < nn; ++q) {                                   /* :163-164 */
                double kp = k * pen[q];
                Dg[q] = D[q] + kp;
            }
            int ends[2] = {eu, ev};
            for (int s = 0; s < 2; ++s) {                                       /* :167 */
                int node = ends[s];
                if (node == 0) continue;                                        /* :168 */
                int i = 0;
                while (cur[i] != node) ++i;                                     /* :169 list.index */
                for (int op = 0; op < 2; ++op) {                                /* :171 */
                    double delta; int mi, mj, found;
                    if (op == 0) found = glsp_two_opt_o2a(cur, Dg, n, i, first_improvement, &delta, &mi, &mj, scratch);
                    else         found = glsp_relocate_o2a(cur, Dg, n, i, first_improvement, &delta, &mi, &mj, scratch);
                    ++o2a;
                    if (found < 0) { free(pen); free(Dg); free(cur); free(scratch); return -1; }
                    if (found && delta < 0) {                                   /* :175-183 */
                        memcpy(cur, scratch, sizeof(int) * (n + 1));
                        cur_cost = glsp_tour_cost(D, n, cur);
                        log_cost(&lg, cur_cost);
                        moves += 1; ++pmoves;                                   /* :185 */
                    }
                }
            }
        }
        local_search_impl(cur, &cur_cost, D, n, first_improvement, scratch, &lg);  /* :188 */
        if (cur_cost < *best_cost) {                                               /* :190-191 */
            memcpy(best_tour, cur, sizeof(int) * (n + 1));
            *best_cost = cur_cost;
        }
    }
    *n_events = lg.n_ev;
    if (pen_out) memcpy(pen_out, pen, nn * sizeof(double));
    if (counters) { counters[0] = lg.sweeps; counters[1] = o2a; counters[2] = pmoves; }
    free(pen); free(Dg); free(cur); free(scratch);
    return 0;
}

/* scripts/test.py:79-83 — float32 regret (already inverse-scaled) widened to fp64 and clamped
 * at 0, scattered to a symmetric n x n matrix in line-graph node order (i<j lexicographic). */
void glsp_regret_matrix(const float *regret, int n, double *W) {
    size_t v = 0;
    for (int i = 0; i < n; ++i) {
        AT(W, n, i, i) = 0.0;
        for (int j = i + 1; j < n; ++j, ++v) {
            double r = (double)regret[v];
            if (!(r > 0.0)) r = 0.0;     /* np.maximum(r, 0) */
            AT(W, n, i, j) = r; AT(W, n, j, i) = r;
        }
    }
}

/* ---- batch driver (CPU baseline): test.py:79-95 per instance, instances across threads ---- */
typedef struct {
    const double *D; const float *regret; int B, n, n_iters, perturbation_moves;
    int *best_tours; double *best_costs; long *counters; int tid, nthreads;
} batch_job_t;

static void *batch_worker(void *arg) {
    batch_job_t *jb = (batch_job_t *)arg;
    int n = jb->n; size_t nn = (size_t)n * n; size_t N = (size_t)n * (n - 1) / 2;
    double *W = (double *)malloc(nn * sizeof(double));
    int *init = (int *)malloc(sizeof(int) * (n + 1));
    for (int b = jb->tid; b < jb->B; b += jb->nthreads) {
        const double *D = jb->D + (size_t)b * nn;
        const double *guide = D; int nev = 0;
        if (jb->regret) { glsp_regret_matrix(jb->regret + (size_t)b * N, n, W); guide = W; }
        glsp_nearest_neighbor(guide, n, 0, init);                     /* test.py:85/88 */
        double c0 = glsp_tour_cost(D, n, init);                       /* test.py:90 */
        glsp_guided_local_search(D, guide, 1, n, init, c0, jb->n_iters, jb->perturbation_moves, 0,
                                 jb->best_tours + (size_t)b * (n + 1), jb->best_costs + b,
                                 NULL, &nev, 0, NULL, jb->counters ? jb->counters + 3 * (size_t)b : NULL);
    }
    free(W); free(init);
    return NULL;
}

int glsp_pipeline_batch(const double *D, const float *regret, int B, int n, int n_iters,
                        int perturbation_moves, int nthreads, int *best_tours, double *best_costs,
                        long *counters) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    batch_job_t *jobs = (batch_job_t *)malloc(sizeof(batch_job_t) * nthreads);
    for (int t = 0; t < nthreads; ++t) {
        batch_job_t j = {D, regret, B, n, n_iters, perturbation_moves, best_tours, best_costs, counters, t, nthreads};
        jobs[t] = j;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
