/*
 * C restatement of the reference's hot loops (TEST INFRASTRUCTURE / CPU BASELINE ONLY;
 * see oracle/__init__.py).  Compiled with gcc by oracle/build.py into
 * oracle/_build/liboracle.so.  Paths cited are relative to
 * /root/reference/src/cogent3.
 *
 *   c3o_inner                 numpy.inner(child_plh, psub)
 *                             evolve/likelihood_calculation.py:86
 *   c3o_sum_input_likelihoods numba sum_input_likelihoods
 *                             evolve/likelihood_tree_numba.py:7-25
 *   c3o_root_lh               numpy.inner(plh_root, mprobs)
 *                             evolve/likelihood_calculation.py:147-148
 *   c3o_log_sum_across_sites  numba get_log_sum_across_sites
 *                             evolve/likelihood_tree_numba.py:36-42
 *   c3o_fused_node            the same arithmetic as inner + sum_input_likelihoods for
 *                             one node, fused and OpenMP-parallel over patterns: a
 *                             STRONGER CPU baseline than the reference (which is
 *                             single-threaded apart from BLAS), used by bench.py so the
 *                             GPU is compared with all host cores.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

/* out[p, i] = sum_j child[p, j] * P[i, j] */
API void c3o_inner(const double* child, const double* P, double* out, int64_t rows, int M) {
  for (int64_t p = 0; p < rows; ++p) {
    const double* x = child + p * M;
    double* y = out + p * M;
    for (int i = 0; i < M; ++i) {
      const double* Pi = P + (int64_t)i * M;
      double s = 0.0;
      for (int j = 0; j < M; ++j) s += x[j] * Pi[j];
      y[i] = s;
    }
  }
}

/* result[p, m] = prod_c lik_c[index[c, p], m]; child 0 assigns, the others multiply */
API void c3o_sum_input_likelihoods(const int64_t* child_indexes, double* result,
                                   const double* const* likelihoods, int C, int64_t rows, int M) {
  for (int c = 0; c < C; ++c) {
    const int64_t* index = child_indexes + (int64_t)c * rows;
    const double* plhs = likelihoods[c];
    if (c == 0) {
      for (int64_t p = 0; p < rows; ++p) {
        const double* src = plhs + index[p] * M;
        double* dst = result + p * M;
        for (int m = 0; m < M; ++m) dst[m] = src[m];
      }
    } else {
      for (int64_t p = 0; p < rows; ++p) {
        const double* src = plhs + index[p] * M;
        double* dst = result + p * M;
        for (int m = 0; m < M; ++m) dst[m] *= src[m];
      }
    }
  }
}

API void c3o_root_lh(const double* plh, const double* pi, double* lh, int64_t rows, int M) {
  for (int64_t p = 0; p < rows; ++p) {
    double s = 0.0;
    for (int m = 0; m < M; ++m) s += plh[p * M + m] * pi[m];
    lh[p] = s;
  }
}

API double c3o_log_sum_across_sites(const double* lhs, const double* counts, int64_t rows) {
  double res = 0.0;
  for (int64_t i = 0; i < rows; ++i) res += log(lhs[i]) * counts[i];
  return res;
}

/* fused per-node update, all bins of one node, OpenMP over patterns.
 * child c: rows src[c] ([rows_c][K][M] layout like the GPU engine), gather index idx[c]
 * (int32 [rows]), P[c] ([K][M][M]) or NULL when the child rows are already multiplied. */
API void c3o_fused_node(double* out, int64_t rows, int K, int M, int C, const double* const* src,
                        const int32_t* const* idx, const double* const* P, int n_threads) {
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < rows; ++p) {
    for (int b = 0; b < K; ++b) {
      double* y = out + (p * K + b) * M;
      for (int c = 0; c < C; ++c) {
        const double* x = src[c] + ((int64_t)idx[c][p] * K + b) * M;
        if (P[c]) {
          const double* Pb = P[c] + (int64_t)b * M * M;
          for (int i = 0; i < M; ++i) {
            double s = 0.0;
            for (int j = 0; j < M; ++j) s += x[j] * Pb[i * M + j];
            if (c == 0) y[i] = s; else y[i] *= s;
          }
        } else {
          for (int i = 0; i < M; ++i) {
            if (c == 0) y[i] = x[i]; else y[i] *= x[i];
          }
        }
      }
    }
  }
}

/* lnL = sum_p log(sum_b bprob_b * lh[p, b]) * counts[p]  (sequential, like the reference) */
API double c3o_mix_log_sum(const double* plh_root, const double* pi, const double* bprobs,
                           const double* counts, int64_t rows, int K, int M) {
  double res = 0.0;
  for (int64_t p = 0; p < rows; ++p) {
    double mix = 0.0;
    for (int b = 0; b < K; ++b) {
      const double* x = plh_root + (p * K + b) * M;
      double s = 0.0;
      for (int m = 0; m < M; ++m) s += x[m] * pi[m];
      mix = (K == 1) ? s : mix + s * bprobs[b];
    }
    res += log(mix) * counts[p];
  }
  return res;
}

API int c3o_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
