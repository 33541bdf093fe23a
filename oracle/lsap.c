/* Oracle: rectangular linear sum assignment, restated in plain C.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- built into oracle/_ref/liblsap_ref.so by
 * oracle/Makefile; never linked into the product library.
 *
 * Third-party algorithm: scipy 1.18.1 scipy.optimize.linear_sum_assignment (Crouse's shortest
 * augmenting path, "On implementing 2D rectangular assignment algorithms", IEEE TAES 2016), called by
 * the reference at deep_sort/sort/linear_assignment.py:56.  scipy's C++ is not under /root/reference,
 * so the published algorithm is restated here, including the details that decide WHICH optimal
 * assignment is returned when costs tie:
 *   - rows are inserted in index order; if nc < nr the problem is transposed first;
 *   - the set of unscanned columns is the array remaining[] = {nc-1, ..., 0}, scanned in array order,
 *     and a scanned column is removed by swapping the last element into its slot;
 *   - a column replaces the current minimum if its reduced path cost is strictly lower, or equal and
 *     the column is unassigned;
 *   - duals and path costs are double; the float32 costs convert exactly.
 * Pinned by differential fuzzing against the installed scipy (tests/test_oracle_thirdparty.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int augmenting_path(int nc, const double *cost, const double *u, const double *v, int *path,
                           const int *row4col, double *spc, int i, char *SR, char *SC, int *remaining,
                           double *p_min_val) {
    double min_val = 0;
    int num_remaining = nc;
    for (int it = 0; it < nc; it++) remaining[it] = nc - it - 1;
    memset(SC, 0, (size_t)nc);
    for (int j = 0; j < nc; j++) spc[j] = INFINITY;
    int sink = -1;
    while (sink == -1) {
        int index = -1;
        double lowest = INFINITY;
        SR[i] = 1;
        for (int it = 0; it < num_remaining; it++) {
            int j = remaining[it];
            double r = min_val + cost[(size_t)i * nc + j] - u[i] - v[j];
            if (r < spc[j]) { path[j] = i; spc[j] = r; }
            if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
        }
        min_val = lowest;
        if (min_val == INFINITY) return -1;
        int j = remaining[index];
        if (row4col[j] == -1) sink = j; else i = row4col[j];
        SC[j] = 1;
        remaining[index] = remaining[--num_remaining];
    }
    *p_min_val = min_val;
    return sink;
}

/* cost: nr x nc row-major float32.  Writes k = min(nr,nc) pairs (row_ind ascending) and returns k,
 * or -1 if infeasible. */
int lsap_ref_solve(int nr, int nc, const float *cost_f32, int *row_ind, int *col_ind) {
    if (nr == 0 || nc == 0) return 0;
    int transpose = nc < nr;
    int R = transpose ? nc : nr, C = transpose ? nr : nc;
    double *cost = (double *)malloc(sizeof(double) * (size_t)R * C);
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++)
            cost[(size_t)i * C + j] = transpose ? (double)cost_f32[(size_t)j * nc + i] : (double)cost_f32[(size_t)i * nc + j];
    double *u = (double *)calloc((size_t)R, sizeof(double)), *v = (double *)calloc((size_t)C, sizeof(double));
    double *spc = (double *)malloc(sizeof(double) * (size_t)C);
    int *path = (int *)malloc(sizeof(int) * (size_t)C), *col4row = (int *)malloc(sizeof(int) * (size_t)R);
    int *row4col = (int *)malloc(sizeof(int) * (size_t)C), *remaining = (int *)malloc(sizeof(int) * (size_t)C);
    char *SR = (char *)malloc((size_t)R), *SC = (char *)malloc((size_t)C);
    for (int j = 0; j < C; j++) { path[j] = -1; row4col[j] = -1; }
    for (int i = 0; i < R; i++) col4row[i] = -1;
    int ok = 1;
    for (int cur = 0; cur < R && ok; cur++) {
        double min_val;
        memset(SR, 0, (size_t)R);
        int sink = augmenting_path(C, cost, u, v, path, row4col, spc, cur, SR, SC, remaining, &min_val);
        if (sink < 0) { ok = 0; break; }
        u[cur] += min_val;
        for (int i = 0; i < R; i++)
            if (SR[i] && i != cur) u[i] += min_val - spc[col4row[i]];
        for (int j = 0; j < C; j++)
            if (SC[j]) v[j] -= min_val - spc[j];
        int j = sink;
        for (;;) {
            int i = path[j];
            row4col[j] = i;
            int t = col4row[i]; col4row[i] = j; j = t;
            if (i == cur) break;
        }
    }
    int k = -1;
    if (ok) {
        k = R;
        if (!transpose) {
            for (int i = 0; i < R; i++) { row_ind[i] = i; col_ind[i] = col4row[i]; }
        } else {                       /* pairs sorted by original row = ascending col4row value */
            int n = 0;
            for (int j = 0; j < C; j++)
                if (row4col[j] != -1) { row_ind[n] = j; col_ind[n] = row4col[j]; n++; }
        }
    }
    free(cost); free(u); free(v); free(spc); free(path); free(col4row); free(row4col); free(remaining);
    free(SR); free(SC);
    return k;
}
