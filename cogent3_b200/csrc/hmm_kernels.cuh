// Phylo-HMM reduction over the SITES.
//
// With sites_independent=False cogent3 reduces the per-bin root likelihoods with a hidden Markov model over
// "patches" of bins (PatchSiteDistribution / SiteHmm, evolve/likelihood_calculation.py:197-235,260-296):
//     plhs[site][patch] = sum_{bins b of the patch} weight_b * lh_b[index[site]]      (get_weighted_sum_lhs)
//     state <- (switch . state) * plhs[site]   for every site in order, rescaled by BASE = 2**100 whenever
//     max(state) < 1;   lnL = log(sum(state)) + exponent * log(BASE)                   (log_dot_reduce,
//                                                                  evolve/likelihood_tree.py:168-176)
// -- a Python loop over every alignment column in the reference.  The recursion is a product of P x P matrices
// M_s = diag(plhs[s]) . switch, and matrix products are associative: every thread multiplies the matrices of a
// contiguous run of sites (carrying a power-of-two exponent, exact), a block combines its threads' products in
// site order, and one thread folds the block products into the stationary vector.  Same value as the
// sequential loop up to the rounding of a different multiplication order (~1e-15 relative); powers of two are
// exact in both.  P = 2 (cogent3's PatchSiteDistribution always makes two patches).
#pragma once
#include "c3_types.cuh"

namespace c3 {

struct Hmm2 {
  double a00, a01, a10, a11;  // the product so far, times 2^-e
  int e;
};

// out = later . earlier
__device__ __forceinline__ Hmm2 hmm_combine(const Hmm2& later, const Hmm2& earlier) {
  Hmm2 r;
  r.a00 = later.a00 * earlier.a00 + later.a01 * earlier.a10;
  r.a01 = later.a00 * earlier.a01 + later.a01 * earlier.a11;
  r.a10 = later.a10 * earlier.a00 + later.a11 * earlier.a10;
  r.a11 = later.a10 * earlier.a01 + later.a11 * earlier.a11;
  r.e = later.e + earlier.e;
  const double m = fmax(fmax(r.a00, r.a01), fmax(r.a10, r.a11));
  if (m > 0.0 && m < __longlong_as_double(0x7ff0000000000000ll)) {
    const int k = ilogb(m);
    if (k < -64 || k > 64) {
      r.a00 = scalbn(r.a00, -k); r.a01 = scalbn(r.a01, -k); r.a10 = scalbn(r.a10, -k); r.a11 = scalbn(r.a11, -k);
      r.e += k;
    }
  }
  return r;
}

constexpr int HMM_THREADS = 256;

struct HmmArgs {
  const double* lh;          // [rows][K] root likelihoods per pattern and bin
  const int32_t* index;      // [n_sites] pattern of every site (LikelihoodTreeEdge.index)
  const int32_t* root_exp;   // [rows] per-pattern exponents when rescaling is active, else nullptr
  int64_t n_sites;
  int K;
  int32_t patch_of_bin[64];
  double weight_of_bin[64];
  double sw[4];              // switch matrix, row-major: state_i <- sum_j sw[i][j] state_j
  double patch_probs[2];
};

// one product per block: block_out[b] = M_{last site of b} ... M_{first site of b}
__global__ void __launch_bounds__(HMM_THREADS) hmm_chunk_kernel(const HmmArgs a, Hmm2* __restrict__ block_out) {
  __shared__ Hmm2 s[HMM_THREADS];
  const int64_t per_thread = (a.n_sites + (int64_t)gridDim.x * HMM_THREADS - 1) / ((int64_t)gridDim.x * HMM_THREADS);
  const int64_t t_global = (int64_t)blockIdx.x * HMM_THREADS + threadIdx.x;
  const int64_t lo = t_global * per_thread;
  const int64_t hi = (lo + per_thread < a.n_sites) ? lo + per_thread : a.n_sites;
  Hmm2 acc{1.0, 0.0, 0.0, 1.0, 0};
  for (int64_t site = lo; site < hi; ++site) {
    const int32_t p = a.index[site];
    double v0 = 0.0, v1 = 0.0;
    for (int b = 0; b < a.K; ++b) {
      const double t = __dmul_rn(a.lh[(int64_t)p * a.K + b], a.weight_of_bin[b]);  // temp = lh * weight
      if (a.patch_of_bin[b] == 0) v0 = __dadd_rn(v0, t); else v1 = __dadd_rn(v1, t);  // result[patch] += temp
    }
    Hmm2 m;
    m.a00 = v0 * a.sw[0]; m.a01 = v0 * a.sw[1];
    m.a10 = v1 * a.sw[2]; m.a11 = v1 * a.sw[3];
    m.e = a.root_exp ? a.root_exp[p] : 0;
    acc = hmm_combine(m, acc);
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  // ordered tree: thread t holds the product of threads [t, t + stride) -- later runs multiply from the left
  for (int stride = 1; stride < HMM_THREADS; stride <<= 1) {
    if ((threadIdx.x & (2 * stride - 1)) == 0) s[threadIdx.x] = hmm_combine(s[threadIdx.x + stride], s[threadIdx.x]);
    __syncthreads();
  }
  if (threadIdx.x == 0) block_out[blockIdx.x] = s[0];
}

// lnL = log(sum(M_all . patch_probs)) + e ln 2
__global__ void hmm_final_kernel(const HmmArgs a, const Hmm2* __restrict__ blocks, int n_blocks, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s0 = a.patch_probs[0], s1 = a.patch_probs[1];
  long long e = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const Hmm2 m = blocks[b];
    const double t0 = m.a00 * s0 + m.a01 * s1;
    const double t1 = m.a10 * s0 + m.a11 * s1;
    s0 = t0; s1 = t1;
    e += m.e;
    const double mx = fmax(s0, s1);
    if (mx > 0.0 && mx < __longlong_as_double(0x7ff0000000000000ll)) {
      const int k = ilogb(mx);
      if (k < -64 || k > 64) { s0 = scalbn(s0, -k); s1 = scalbn(s1, -k); e += k; }
    }
  }
  out[0] = log(s0 + s1) + (double)e * 0.6931471805599453094;
}

}  // namespace c3
