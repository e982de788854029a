/*
 * oracle/lapjv.c -- TEST INFRASTRUCTURE ONLY (parity checker, never shipped, never timed as the product).
 *
 * CPU restatement of the dense Jonker-Volgenant linear-assignment solver that the reference calls as
 * `lap.lapjv` (third-party `lapx` 0.9.4, pinned in /root/reference/uv.lock:2522-2523; call sites
 * boxmot/trackers/association/matching.py:36 and boxmot/trackers/association/association.py:23).
 * lapx is NOT vendored under /root/reference and is not installed in this image, so this file restates the
 * published algorithm (R. Jonker, A. Volgenant, "A shortest augmenting path algorithm for dense and sparse
 * linear assignment problems", Computing 38, 1987): column reduction + reduction transfer, two passes of
 * augmenting row reduction, then shortest-augmenting-path augmentation for the remaining free rows.
 *
 * PARITY UNPINNED at the lapx boundary (no lapx binary, no golden vectors in the reference tests). It is
 * pinned instead by (a) brute force for n <= 8, (b) scipy.optimize.linear_sum_assignment on the same
 * (extended) matrix, see tests/test_oracle_lap.py.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/liboracle_lapjv.so oracle/lapjv.c
 */
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define BIG DBL_MAX

/* Phase 1: every column picks its cheapest row (scanning rows in ascending order so later rows win only
 * when strictly cheaper); columns are then claimed from the last column down, so each row keeps the
 * highest-index column that elected it.  Rows elected exactly once get their column price lowered by
 * the second-best reduced cost (reduction transfer). */
static int column_reduction(int n, const double *c, int *free_rows, int *x, int *y, double *v)
{
    int *once = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) { x[i] = -1; v[i] = BIG; y[i] = 0; once[i] = 1; }
    for (int i = 0; i < n; ++i) {
        const double *ci = c + (size_t)i * n;
        for (int j = 0; j < n; ++j) {
            if (ci[j] < v[j]) { v[j] = ci[j]; y[j] = i; }
        }
    }
    for (int j = n - 1; j >= 0; --j) {
        int i = y[j];
        if (x[i] < 0) {
            x[i] = j;
        } else {
            once[i] = 0;
            y[j] = -1;
        }
    }
    int n_free = 0;
    for (int i = 0; i < n; ++i) {
        if (x[i] < 0) {
            free_rows[n_free++] = i;
        } else if (once[i]) {
            int j = x[i];
            const double *ci = c + (size_t)i * n;
            double second = BIG;
            for (int k = 0; k < n; ++k) {
                if (k == j) continue;
                double r = ci[k] - v[k];
                if (r < second) second = r;
            }
            v[j] -= second;
        }
    }
    free(once);
    return n_free;
}

/* Phase 2: augmenting row reduction.  Each free row looks at its two cheapest reduced columns; it takes the
 * cheapest, lowering that column's price by the gap when the gap is positive, and evicts the previous owner
 * (which is retried immediately if the price moved, or deferred otherwise). */
static int row_reduction(int n, const double *c, int n_free, int *free_rows, int *x, int *y, double *v)
{
    int cur = 0, kept = 0;
    long long rounds = 0;
    while (cur < n_free) {
        ++rounds;
        int fi = free_rows[cur++];
        const double *ci = c + (size_t)fi * n;
        int j1 = 0, j2 = -1;
        double m1 = ci[0] - v[0], m2 = BIG;
        for (int j = 1; j < n; ++j) {
            double r = ci[j] - v[j];
            if (r < m2) {
                if (r >= m1) { m2 = r; j2 = j; }
                else { m2 = m1; m1 = r; j2 = j1; j1 = j; }
            }
        }
        int i0 = y[j1];
        double lowered = v[j1] - (m2 - m1);
        int moves = lowered < v[j1];
        if (rounds < (long long)cur * n) {
            if (moves) {
                v[j1] = lowered;
            } else if (i0 >= 0 && j2 >= 0) {
                j1 = j2;
                i0 = y[j2];
            }
            if (i0 >= 0) {
                if (moves) free_rows[--cur] = i0;
                else free_rows[kept++] = i0;
            }
        } else {
            if (i0 >= 0) free_rows[kept++] = i0;
        }
        x[fi] = j1;
        y[j1] = fi;
    }
    return kept;
}

/* Phase 3 helper: Dijkstra-like search over columns from `start`; returns the free column that ends the
 * cheapest alternating path and updates column prices of the scanned set. */
static int shortest_path(int n, const double *c, int start, const int *y, double *v, int *pred,
                         int *cols, double *d)
{
    int lo = 0, hi = 0, n_ready = 0, final_j = -1, band = 0;
    for (int j = 0; j < n; ++j) {
        cols[j] = j;
        pred[j] = start;
        d[j] = c[(size_t)start * n + j] - v[j];
    }
    while (final_j == -1) {
        if (lo == hi) {
            /* collect the columns at the current minimum distance into cols[lo..hi) */
            n_ready = lo;
            band = lo;
            double mind = d[cols[lo]];
            hi = lo + 1;
            for (int k = hi; k < n; ++k) {
                int j = cols[k];
                double dj = d[j];
                if (dj <= mind) {
                    if (dj < mind) { hi = lo; mind = dj; }
                    cols[k] = cols[hi];
                    cols[hi++] = j;
                }
            }
            for (int k = lo; k < hi; ++k) {
                if (y[cols[k]] < 0) final_j = cols[k];
            }
        }
        if (final_j == -1) {
            /* scan: relax through the owner of every column in the ready band */
            while (lo != hi && final_j == -1) {
                int j = cols[lo++];
                int i = y[j];
                double mind = d[j];
                const double *ci = c + (size_t)i * n;
                double h = ci[j] - v[j] - mind;
                for (int k = hi; k < n; ++k) {
                    int jj = cols[k];
                    double r = ci[jj] - v[jj] - h;
                    if (r < d[jj]) {
                        d[jj] = r;
                        pred[jj] = i;
                        if (r == mind) {
                            if (y[jj] < 0) { final_j = jj; break; }
                            cols[k] = cols[hi];
                            cols[hi++] = jj;
                        }
                    }
                }
            }
        }
    }
    {
        /* price update for every column that was fully scanned before the last band; all members of the
         * last band share the final distance */
        double mind = d[cols[band]];
        for (int k = 0; k < n_ready; ++k) {
            int j = cols[k];
            v[j] += d[j] - mind;
        }
    }
    return final_j;
}

static void augment_all(int n, const double *c, int n_free, const int *free_rows, int *x, int *y, double *v)
{
    int *pred = (int *)malloc(sizeof(int) * (size_t)n);
    int *cols = (int *)malloc(sizeof(int) * (size_t)n);
    double *d = (double *)malloc(sizeof(double) * (size_t)n);
    for (int f = 0; f < n_free; ++f) {
        int start = free_rows[f];
        int j = shortest_path(n, c, start, y, v, pred, cols, d);
        int i = -1;
        while (i != start) {
            i = pred[j];
            y[j] = i;
            int prev = x[i];
            x[i] = j;
            j = prev;
        }
    }
    free(pred); free(cols); free(d);
}

/* Solve the square n x n problem (row-major cost).  x[i] = column of row i, y[j] = row of column j. */
int oracle_lapjv_square(int n, const double *cost, int *x, int *y)
{
    if (n <= 0) return 0;
    int *free_rows = (int *)malloc(sizeof(int) * (size_t)n);
    double *v = (double *)malloc(sizeof(double) * (size_t)n);
    int n_free = column_reduction(n, cost, free_rows, x, y, v);
    for (int pass = 0; pass < 2 && n_free > 0; ++pass)
        n_free = row_reduction(n, cost, n_free, free_rows, x, y, v);
    if (n_free > 0) augment_all(n, cost, n_free, free_rows, x, y, v);
    free(free_rows); free(v);
    return 0;
}
