/*
 * oracle/nm3.h -- TEST INFRASTRUCTURE (CPU oracle).  Not part of the product path.
 *
 * Bounded Nelder-Mead simplex minimiser, written for this repository.
 *
 * Why it exists: the reference refines a patch with nlopt 2.6.1 LN_BOBYQA
 * (/root/reference/source/pmvs/optim.cpp:621-644).  nlopt is NOT vendored in the
 * reference tree (lib/CMakeLists.txt:19-25 clones it at configure time) and is absent
 * from this machine, and BASELINE.json's north_star replaces it with a batched
 * Nelder-Mead.  This header is the single written-down definition of that
 * Nelder-Mead.  It is used by
 *   - oracle/shim/nlopt.hpp  (stand-in that lets the reference's own optim.cpp link),
 *   - oracle/pmvs_oracle.c   (CPU restatement of the hot path),
 * and the CUDA kernel (cmvs-pmvs_b200/csrc/refine.cuh) implements the same steps.
 * PARITY UNPINNED: nothing in the reference pins optimiser iterates.
 *
 * Definition (n fixed small; all arithmetic in double):
 *   simplex    x_0 = clamp(start), x_i = x_0 + step*e_i (or x_0 - step*e_i if that
 *              would leave the box), f evaluated in that order: x_0, x_1, ..., x_n
 *   order      vertices kept sorted by f ascending, insertion sort, stable
 *              (a new point goes AFTER existing points of equal f)
 *   iteration  c  = centroid of the n best
 *              xr = clamp(c + (c - x_worst));              fr = f(xr)
 *              fr <  f_best           : xe = clamp(c + 2 (c - x_worst)); fe = f(xe)
 *                                       accept xe if fe < fr else xr
 *              f_best <= fr < f_2ndworst : accept xr
 *              otherwise              : xc = c + 0.5 (xr - c)     if fr < f_worst (outside)
 *                                       xc = c + 0.5 (x_worst - c) otherwise      (inside)
 *                                       fc = f(xc); accept xc if fc < min(fr, f_worst)
 *                                       else shrink: x_i = x_best + 0.5 (x_i - x_best), i = 1..n,
 *                                       re-evaluated in order, then re-sorted
 *   stop       after the initial simplex and after every iteration:
 *              size = max_i max_j |x_i[j] - x_best[j]|;  size <= xtol  -> NM3_XTOL_REACHED
 *              the evaluation budget is checked BEFORE each evaluation:
 *              evals == maxeval -> NM3_MAXEVAL_REACHED (best point so far is returned)
 *   result     x_best, f_best, number of evaluations
 */
#ifndef PMVS_ORACLE_NM3_H
#define PMVS_ORACLE_NM3_H

#define NM3_MAXN 4
#define NM3_XTOL_REACHED 4    /* same numeric values as nlopt's result enum */
#define NM3_MAXEVAL_REACHED 5

typedef double (*nm3_func)(unsigned n, const double* x, void* data);

static inline double nm3_clampd(double v, double lo, double hi) {
  return v < lo ? lo : (v > hi ? hi : v);
}

/* Returns NM3_XTOL_REACHED or NM3_MAXEVAL_REACHED.  x: in = start, out = best. */
static inline int nm3_minimize(unsigned n, nm3_func f, void* data, const double* lb, const double* ub,
                               double* x, double* fmin, double step, double xtol, int maxeval,
                               int* nevals) {
  double p[NM3_MAXN + 1][NM3_MAXN];
  double fv[NM3_MAXN + 1];
  int cnt = 0;
  int ret = NM3_XTOL_REACHED;
  unsigned i, j, k;

#define NM3_EVAL(dst, pt)                          \
  do {                                             \
    if (cnt >= maxeval) {                          \
      ret = NM3_MAXEVAL_REACHED;                   \
      goto done;                                   \
    }                                              \
    (dst) = f(n, (pt), data);                      \
    ++cnt;                                         \
  } while (0)

  /* insertion of vertex k into the sorted prefix [0, k) */
#define NM3_INSERT(k_)                                                   \
  do {                                                                   \
    double tx[NM3_MAXN];                                                 \
    double tf = fv[(k_)];                                                \
    unsigned q = (k_);                                                   \
    for (j = 0; j < n; ++j) tx[j] = p[(k_)][j];                          \
    while (q > 0 && tf < fv[q - 1]) {                                    \
      for (j = 0; j < n; ++j) p[q][j] = p[q - 1][j];                     \
      fv[q] = fv[q - 1];                                                 \
      --q;                                                               \
    }                                                                    \
    for (j = 0; j < n; ++j) p[q][j] = tx[j];                             \
    fv[q] = tf;                                                          \
  } while (0)

  for (j = 0; j < n; ++j) p[0][j] = nm3_clampd(x[j], lb[j], ub[j]);
  for (i = 1; i <= n; ++i) {
    for (j = 0; j < n; ++j) p[i][j] = p[0][j];
    if (p[0][i - 1] + step > ub[i - 1])
      p[i][i - 1] = p[0][i - 1] - step;
    else
      p[i][i - 1] = p[0][i - 1] + step;
  }
  /* fv defaults so that an early MAXEVAL exit still returns something sane */
  for (i = 0; i <= n; ++i) fv[i] = 1.0e300;
  for (i = 0; i <= n; ++i) {
    NM3_EVAL(fv[i], p[i]);
    NM3_INSERT(i);
  }

  for (;;) {
    double c[NM3_MAXN], xr[NM3_MAXN], xe[NM3_MAXN], xc[NM3_MAXN];
    double fr, fe, fc, size = 0.0;

    for (i = 1; i <= n; ++i)
      for (j = 0; j < n; ++j) {
        double d = p[i][j] - p[0][j];
        if (d < 0) d = -d;
        if (d > size) size = d;
      }
    if (size <= xtol) {
      ret = NM3_XTOL_REACHED;
      break;
    }

    for (j = 0; j < n; ++j) {
      double s = 0.0;
      for (i = 0; i < n; ++i) s += p[i][j];
      c[j] = s / (double)n;
    }
    for (j = 0; j < n; ++j) xr[j] = nm3_clampd(c[j] + (c[j] - p[n][j]), lb[j], ub[j]);
    NM3_EVAL(fr, xr);

    if (fr < fv[0]) {
      for (j = 0; j < n; ++j) xe[j] = nm3_clampd(c[j] + 2.0 * (c[j] - p[n][j]), lb[j], ub[j]);
      NM3_EVAL(fe, xe);
      if (fe < fr) {
        for (j = 0; j < n; ++j) p[n][j] = xe[j];
        fv[n] = fe;
      } else {
        for (j = 0; j < n; ++j) p[n][j] = xr[j];
        fv[n] = fr;
      }
      NM3_INSERT(n);
    } else if (fr < fv[n - 1]) {
      for (j = 0; j < n; ++j) p[n][j] = xr[j];
      fv[n] = fr;
      NM3_INSERT(n);
    } else {
      double fref;
      if (fr < fv[n]) {
        for (j = 0; j < n; ++j) xc[j] = c[j] + 0.5 * (xr[j] - c[j]);
        fref = fr;
      } else {
        for (j = 0; j < n; ++j) xc[j] = c[j] + 0.5 * (p[n][j] - c[j]);
        fref = fv[n];
      }
      NM3_EVAL(fc, xc);
      if (fc < fref) {
        for (j = 0; j < n; ++j) p[n][j] = xc[j];
        fv[n] = fc;
        NM3_INSERT(n);
      } else {
        for (i = 1; i <= n; ++i) {
          for (j = 0; j < n; ++j) p[i][j] = p[0][j] + 0.5 * (p[i][j] - p[0][j]);
          NM3_EVAL(fv[i], p[i]);
        }
        for (k = 1; k <= n; ++k) NM3_INSERT(k);
      }
    }
  }

done:
  /* On a MAXEVAL exit in the middle of a shrink the array may be unsorted: pick the min. */
  k = 0;
  for (i = 1; i <= n; ++i)
    if (fv[i] < fv[k]) k = i;
  for (j = 0; j < n; ++j) x[j] = p[k][j];
  *fmin = fv[k];
  if (nevals) *nevals = cnt;
  return ret;
#undef NM3_EVAL
#undef NM3_INSERT
}

#endif
