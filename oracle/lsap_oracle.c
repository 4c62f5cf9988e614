/*
 * lsap_oracle.c -- CPU restatement of the assignment step of the reference's Hungarian matcher.
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by
 * the product path.
 *
 * The reference calls scipy.optimize.linear_sum_assignment on the CPU copy of the cost matrix
 * (/root/reference/models/GroundingDINO/matcher.py:104-110).  scipy is a third-party dependency that is not
 * vendored in /root/reference and not version-pinned by it (requirements.txt:3 says just `scipy`); the version in
 * this image is 1.18.1.  Its solver is the rectangular shortest-augmenting-path algorithm of D. F. Crouse, "On
 * implementing 2D rectangular assignment algorithms", IEEE TAES 52(4), 2016 (scipy/optimize/rectangular_lsap),
 * restated here from the published algorithm including scipy's documented conventions:
 *   - if there are more rows than columns the problem is transposed, and the answer is reported sorted by row;
 *   - the set of unscanned columns is an array filled in REVERSE order and compacted by moving the last element
 *     into the freed slot; among equal shortest-path costs a column that is still unassigned wins (the last one
 *     met in array order), otherwise the first column met.
 * Parity is pinned by tests/test_lsap_oracle_cpu.py against scipy itself (random, tie-heavy small-integer,
 * constant, rectangular both ways, infeasible); the GPU kernel (csrc/hm_lsap.cu) is then compared with this file
 * and with scipy.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* cost: nr x nc row-major doubles (finite or +inf).  a, b: min(nr, nc) outputs (row indices ascending, their
 * columns).  Returns 0, -1 (infeasible) or -2 (NaN / -inf entry), like scipy's error classes. */
int lsap_solve(int64_t nr, int64_t nc, const double *cost_in, int64_t *a, int64_t *b) {
  if (nr == 0 || nc == 0) return 0;
  const int transpose = nc < nr;
  double *cost = (double *)malloc(sizeof(double) * (size_t)(nr * nc));
  if (transpose) {
    for (int64_t i = 0; i < nr; ++i)
      for (int64_t j = 0; j < nc; ++j) cost[j * nr + i] = cost_in[i * nc + j];
    const int64_t t = nr; nr = nc; nc = t;
  } else {
    memcpy(cost, cost_in, sizeof(double) * (size_t)(nr * nc));
  }
  for (int64_t k = 0; k < nr * nc; ++k)
    if (isnan(cost[k]) || cost[k] == -INFINITY) { free(cost); return -2; }

  double *u = (double *)calloc((size_t)nr, sizeof(double)), *v = (double *)calloc((size_t)nc, sizeof(double));
  double *spc = (double *)malloc(sizeof(double) * (size_t)nc);
  int64_t *path = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc);
  int64_t *col4row = (int64_t *)malloc(sizeof(int64_t) * (size_t)nr);
  int64_t *row4col = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc);
  int64_t *remaining = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc);
  char *SR = (char *)malloc((size_t)nr), *SC = (char *)malloc((size_t)nc);
  for (int64_t i = 0; i < nr; ++i) col4row[i] = -1;
  for (int64_t j = 0; j < nc; ++j) { row4col[j] = -1; path[j] = -1; }
  int rc = 0;

  for (int64_t cur = 0; cur < nr && rc == 0; ++cur) {
    /* shortest augmenting path from row `cur` */
    double min_val = 0.0;
    int64_t n_rem = nc, i = cur, sink = -1;
    for (int64_t it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
    memset(SR, 0, (size_t)nr);
    memset(SC, 0, (size_t)nc);
    for (int64_t j = 0; j < nc; ++j) spc[j] = INFINITY;
    while (sink == -1) {
      int64_t index = -1;
      double lowest = INFINITY;
      SR[i] = 1;
      for (int64_t it = 0; it < n_rem; ++it) {
        const int64_t j = remaining[it];
        const double r = min_val + cost[i * nc + j] - u[i] - v[j];
        if (r < spc[j]) { path[j] = i; spc[j] = r; }
        if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
      }
      min_val = lowest;
      if (min_val == INFINITY) { rc = -1; break; }
      const int64_t j = remaining[index];
      if (row4col[j] == -1) sink = j; else i = row4col[j];
      SC[j] = 1;
      remaining[index] = remaining[--n_rem];
    }
    if (rc) break;
    /* dual update */
    u[cur] += min_val;
    for (int64_t r = 0; r < nr; ++r)
      if (SR[r] && r != cur) u[r] += min_val - spc[col4row[r]];
    for (int64_t j = 0; j < nc; ++j)
      if (SC[j]) v[j] -= min_val - spc[j];
    /* augment */
    int64_t j = sink;
    for (;;) {
      const int64_t r = path[j];
      row4col[j] = r;
      const int64_t t = col4row[r]; col4row[r] = j; j = t;
      if (r == cur) break;
    }
  }

  if (rc == 0) {
    if (transpose) {     /* rows of the caller are our columns: report them ascending */
      int64_t k = 0;
      for (int64_t j = 0; j < nc; ++j)
        if (row4col[j] != -1) { a[k] = j; b[k] = row4col[j]; ++k; }
    } else {
      for (int64_t r = 0; r < nr; ++r) { a[r] = r; b[r] = col4row[r]; }
    }
  }
  free(cost); free(u); free(v); free(spc); free(path); free(col4row); free(row4col); free(remaining); free(SR); free(SC);
  return rc;
}
