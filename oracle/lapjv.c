/* ORACLE / TEST INFRASTRUCTURE ONLY - never linked into the product library.
 *
 * Dense Jonker-Volgenant linear assignment as performed by the third-party
 * `lap` package (gatagat/lap 0.4.x, requirements/runtime.txt:2 - not vendored
 * in the reference tree).  Call sites in the reference:
 *   mmtrack/models/trackers/ocsort_tracker_disparity.py:260-261, :312-313
 *   lap.lapjv(dists, extend_cost=True, cost_limit=1 - match_iou_thr)
 * Restated from the published algorithm (Jonker & Volgenant 1987, in the
 * phase order and scan order of lap's dense solver): column reduction with a
 * reverse column sweep, reduction transfer, two passes of augmenting row
 * reduction, then shortest-path augmentation with READY/SCAN/TODO column
 * lists.  Scan orders decide results only under exact cost ties.
 */
#include <stdlib.h>
#include <string.h>

#define BIG 1.0e300

typedef struct {
    int n;
    const double *c;   /* n x n row-major */
    int *x, *y;        /* row->col, col->row */
    double *v;         /* column potentials */
} jv_t;

#define COST(p, i, j) ((p)->c[(size_t)(i) * (p)->n + (j)])

/* phase 1+2: returns number of free rows, listed ascending in free_rows */
static int column_reduction_transfer(jv_t *p, int *free_rows)
{
    const int n = p->n;
    int i, j, nfree = 0;
    char *unique = (char *)malloc((size_t)n);
    for (i = 0; i < n; i++) { p->x[i] = -1; p->v[i] = BIG; p->y[i] = 0; unique[i] = 1; }
    for (i = 0; i < n; i++)
        for (j = 0; j < n; j++) {
            double c = COST(p, i, j);
            if (c < p->v[j]) { p->v[j] = c; p->y[j] = i; }
        }
    for (j = n - 1; j >= 0; j--) {
        i = p->y[j];
        if (p->x[i] < 0) p->x[i] = j;
        else { unique[i] = 0; p->y[j] = -1; }
    }
    for (i = 0; i < n; i++) {
        if (p->x[i] < 0) free_rows[nfree++] = i;
        else if (unique[i]) {
            int j1 = p->x[i], j2;
            double m = BIG;
            for (j2 = 0; j2 < n; j2++) {
                double r;
                if (j2 == j1) continue;
                r = COST(p, i, j2) - p->v[j2];
                if (r < m) m = r;
            }
            p->v[j1] -= m;
        }
    }
    free(unique);
    return nfree;
}

/* phase 3: one pass of augmenting row reduction */
static int augmenting_row_reduction(jv_t *p, int nfree, int *free_rows)
{
    const int n = p->n;
    int cur = 0, next_free = 0;
    unsigned long long visits = 0;
    while (cur < nfree) {
        int j, j1 = 0, j2 = -1, i0, fr;
        double u1, u2 = BIG, v1_new;
        int lowers;
        visits++;
        fr = free_rows[cur++];
        u1 = COST(p, fr, 0) - p->v[0];
        for (j = 1; j < n; j++) {
            double r = COST(p, fr, j) - p->v[j];
            if (r < u2) {
                if (r >= u1) { u2 = r; j2 = j; }
                else { u2 = u1; u1 = r; j2 = j1; j1 = j; }
            }
        }
        i0 = p->y[j1];
        v1_new = p->v[j1] - (u2 - u1);
        lowers = v1_new < p->v[j1];
        if (visits < (unsigned long long)cur * (unsigned long long)n) {
            if (lowers) p->v[j1] = v1_new;
            else if (i0 >= 0 && j2 >= 0) { j1 = j2; i0 = p->y[j2]; }
            if (i0 >= 0) {
                if (lowers) free_rows[--cur] = i0;
                else free_rows[next_free++] = i0;
            }
        } else if (i0 >= 0) {
            free_rows[next_free++] = i0;
        }
        p->x[fr] = j1;
        p->y[j1] = fr;
    }
    return next_free;
}

/* phase 4 helpers: cols[0..lo) READY, [lo..hi) SCAN, [hi..n) TODO */
static int collect_minima(int n, int lo, const double *d, int *cols)
{
    int hi = lo + 1, k;
    double mind = d[cols[lo]];
    for (k = hi; k < n; k++) {
        int j = cols[k];
        if (d[j] <= mind) {
            if (d[j] < mind) { hi = lo; mind = d[j]; }
            cols[k] = cols[hi];
            cols[hi++] = j;
        }
    }
    return hi;
}

static int relax_from_scan(jv_t *p, int *plo, int *phi, double *d, int *cols, int *pred)
{
    const int n = p->n;
    int lo = *plo, hi = *phi, k;
    while (lo != hi) {
        int j = cols[lo++];
        int i = p->y[j];
        double mind = d[j];
        double h = COST(p, i, j) - p->v[j] - mind;
        for (k = hi; k < n; k++) {
            double r;
            j = cols[k];
            r = COST(p, i, j) - p->v[j] - h;
            if (r < d[j]) {
                d[j] = r;
                pred[j] = i;
                if (r == mind) {
                    if (p->y[j] < 0) return j;
                    cols[k] = cols[hi];
                    cols[hi++] = j;
                }
            }
        }
    }
    *plo = lo; *phi = hi;
    return -1;
}

static int shortest_path(jv_t *p, int start, int *pred, int *cols, double *d)
{
    const int n = p->n;
    int lo = 0, hi = 0, n_ready = 0, final_j = -1, k;
    double mind;
    for (k = 0; k < n; k++) { cols[k] = k; pred[k] = start; d[k] = COST(p, start, k) - p->v[k]; }
    while (final_j == -1) {
        if (lo == hi) {
            n_ready = lo;
            hi = collect_minima(n, lo, d, cols);
            for (k = lo; k < hi; k++)
                if (p->y[cols[k]] < 0) final_j = cols[k];
        }
        if (final_j == -1)
            final_j = relax_from_scan(p, &lo, &hi, d, cols, pred);
    }
    mind = d[cols[lo]];
    for (k = 0; k < n_ready; k++) { int j = cols[k]; p->v[j] += d[j] - mind; }
    return final_j;
}

static void augment(jv_t *p, int nfree, const int *free_rows)
{
    const int n = p->n;
    int *pred = (int *)malloc(sizeof(int) * n);
    int *cols = (int *)malloc(sizeof(int) * n);
    double *d = (double *)malloc(sizeof(double) * n);
    int f;
    for (f = 0; f < nfree; f++) {
        int i = -1, j = shortest_path(p, free_rows[f], pred, cols, d), guard = 0;
        while (i != free_rows[f] && guard++ <= n) {
            int t;
            i = pred[j];
            p->y[j] = i;
            t = j; j = p->x[i]; p->x[i] = t;
        }
    }
    free(pred); free(cols); free(d);
}

/* Solve a square dense problem; x[i] = column of row i, y[j] = row of col j */
int oracle_lapjv_square(int n, const double *cost, int *x, int *y)
{
    jv_t p;
    int *free_rows, nfree, pass = 0;
    if (n <= 0) return 0;
    p.n = n; p.c = cost; p.x = x; p.y = y;
    p.v = (double *)malloc(sizeof(double) * n);
    free_rows = (int *)malloc(sizeof(int) * n);
    nfree = column_reduction_transfer(&p, free_rows);
    while (nfree > 0 && pass < 2) { nfree = augmenting_row_reduction(&p, nfree, free_rows); pass++; }
    if (nfree > 0) augment(&p, nfree, free_rows);
    free(p.v); free(free_rows);
    return 0;
}

/* lap.lapjv(cost[R x C], extend_cost=True, cost_limit=limit): extended
 * (R+C)^2 matrix filled with limit/2, zero lower-right block; x (len R) and
 * y (len C) use -1 for "unmatched". */
int oracle_lapjv_extended(int R, int C, const double *cost, double limit, int *x, int *y)
{
    const int n = R + C;
    double *ext;
    int *xe, *ye, i, j;
    if (n == 0) return 0;
    ext = (double *)malloc(sizeof(double) * (size_t)n * n);
    xe = (int *)malloc(sizeof(int) * n);
    ye = (int *)malloc(sizeof(int) * n);
    for (i = 0; i < n; i++)
        for (j = 0; j < n; j++)
            ext[(size_t)i * n + j] = (i < R && j < C) ? cost[(size_t)i * C + j]
                                    : (i >= R && j >= C) ? 0.0 : limit / 2.0;
    oracle_lapjv_square(n, ext, xe, ye);
    for (i = 0; i < R; i++) x[i] = xe[i] >= C ? -1 : xe[i];
    for (j = 0; j < C; j++) y[j] = ye[j] >= R ? -1 : ye[j];
    free(ext); free(xe); free(ye);
    return 0;
}
