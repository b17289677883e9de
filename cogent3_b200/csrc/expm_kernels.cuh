// Batched P(t) = expm(Q t): one CTA per dirty (edge, rate class), all of them in
// one launch per exponentiator kind.
//
//   eigen : P = max(Re(evT . diag(exp(t roots)) . evI^T), 0)
//           -- EigenExponentiator.__call__, maths/matrix_exponentiation.py:35-40.
//           The eigen-decomposition itself stays on the host (numpy.linalg.eig /
//           inv, :132-146) so that roots / vectors are bit-identical to the
//           reference's; see DESIGN.md.
//   Pade  : scaling-and-squaring with the reference's data-dependent order
//           -- PadeExponentiator.__call__, maths/matrix_exponentiation.py:92-129.
//
// For a tip edge the CTA also refreshes the pre-multiplied tip table
// tip_inner[u][b][i] = sum_j P_b[i][j] table[u][j], which is numpy.inner(tip
// rows, psub) (evolve/likelihood_calculation.py:86) on the <= U+1 distinct tip
// rows -- the "column gather of P" of SURVEY §7 step 6.
#pragma once
#include "c3_types.cuh"

namespace c3 {

// numpy.maximum(result, 0.0) (maths/matrix_exponentiation.py:40) PROPAGATES NaN; fmax(NaN, 0) would
// return 0 and turn a NaN rate matrix / distance into a finite likelihood
__device__ __forceinline__ double clamp0(double v) { return v < 0.0 ? 0.0 : v; }


// tip_state[u] >= 0: row u is the indicator of that one state (an unambiguous symbol), so its inner
// product with the rows of P is just a column of P -- the same bits as the sum (every other term is
// fma(0, x, s) = s).  Only ambiguity codes and the gap row run the sum.  (ncu on the 61-state
// launch: the sums over all rows were more than half of the kernel's instructions.)
// one entry of tip_inner for a row that is not one-hot (ambiguity codes, the gap row); shared with the chain
// kernel's fused expm so that both produce the same bits
__device__ __forceinline__ double tip_inner_entry(const double* __restrict__ row, const double* sP, int i, int M, int MP) {
  // the j loop starts at i: consecutive lanes (consecutive i) then read different shared
  // memory banks of sP (row stride MP is a multiple of the bank count); the rows are 0/1
  // indicators, so the rotated summation order only matters for ambiguity codes (~1 ulp)
  double s = 0.0;
  int j = (MP > 8) ? i : 0;
  for (int n = 0; n < M; ++n) {
    s = fma(row[j], sP[i * MP + j], s);
    j = (j + 1 == M) ? 0 : j + 1;
  }
  return s;
}

// one entry of P = max(evT . diag(exp(t roots)) . evI^T, 0) in the real eigen form for small alphabets (k summed in
// order, explicit roundings: the same bits wherever it is inlined -- expm_eigen_kernel and the chain kernel)
__device__ __forceinline__ double expm_real_entry(const double* __restrict__ evT_row, const double* __restrict__ evI_row,
                                                  const double* er, int M) {
  double acc = 0.0;
  for (int k = 0; k < M; ++k) acc = __fma_rn(__dmul_rn(evT_row[k], er[k]), evI_row[k], acc);
  return clamp0(acc);
}

__device__ __forceinline__ void tip_inner_update(const EdgeDesc& ed, const double* sP, int bin,
                                                 int M, int MP, int K) {
  if (ed.tip_table == nullptr) return;
  const int64_t total = ed.tip_rows * M;
  for (int64_t o = threadIdx.x; o < total; o += blockDim.x) {
    const int64_t u = o / M;
    const int i = (int)(o - u * M);
    const int state = ed.tip_state[u];
    double s;
    if (state >= 0) {
      s = sP[i * MP + state];
    } else {
      s = tip_inner_entry(ed.tip_table + u * MP, sP, i, M, MP);
    }
    ed.tip_inner[(u * K + bin) * MP + i] = s;
  }
}

// slot layout (doubles): roots[2M] | evT[2 M M] | evI[2 M M]   (complex interleaved; the
// real variant uses the first M / M M entries of each region)  -- or Q[M M] for Pade.
__global__ void expm_eigen_kernel(const ExpmJob* __restrict__ jobs, const double* __restrict__ slots,
                                  int64_t slot_stride, const EdgeDesc* __restrict__ edges, int M,
                                  int MP, int K) {
  extern __shared__ double smem[];
  double* sP = smem;              // [MP*MP]
  double* s_er = sP + MP * MP;    // [M] exp(t re) cos(t im)
  double* s_ei = s_er + M;        // [M] exp(t re) sin(t im)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // see pdl_trigger()
  const ExpmJob job = jobs[blockIdx.x];
  const EdgeDesc ed = edges[job.node];
  double* Pout = ed.P + (int64_t)job.bin * MP * MP;
  const int MM = M * M;

  if (job.mode == EXPM_FIXED) {
    for (int o = threadIdx.x; o < MP * MP; o += blockDim.x) sP[o] = Pout[o];
    __syncthreads();
    tip_inner_update(ed, sP, job.bin, M, MP, K);
    return;
  }
  const double* slot = slots + (int64_t)job.slot * slot_stride;
  const double t = job.t;
  if (job.mode == EXPM_TN93) {
    // closed form for the F81 / HKY85 / TN93 family, alphabet order T, C, A, G: restates calc_TN93_P
    // (evolve/solved_models_numba.py:9-51) expression by expression (PredefinedNucleotide.calc_psub_matrix,
    // evolve/solved_models.py:43-49, is F81, HKY85 or TN93 when passed 0, 1 or 2 parameters)
    if (threadIdx.x < 16) {
      const double* pi = slot;
      const double alpha[2] = {slot[4], slot[5]};
      const double pi_star[2] = {pi[0] + pi[1], pi[2] + pi[3]};
      const double mu[2] = {alpha[0] * pi_star[0] + 1.0 * pi_star[1], 1.0 * pi_star[0] + alpha[1] * pi_star[1]};
      double scale_factor = 0.0;
      for (int motif = 0; motif < 4; ++motif) {
        const int i = motif / 2, other = 1 - i;
        scale_factor += (alpha[i] * pi[2 * i + 1 - motif % 2] + pi_star[other]) * pi[motif];
      }
      const double time = t / scale_factor;
      const double e_beta_t = exp(-time);
      const double transversion = 1 - e_beta_t;
      const int row = threadIdx.x / 4, column = threadIdx.x % 4;
      const int i = row / 2, j = column / 2, other = 1 - i;
      const double e_mu_t = exp(-mu[i] * time);
      const double transition = 1 + (pi_star[other] * e_beta_t - e_mu_t) / pi_star[i];
      double p = (i == j) ? transition : transversion;
      p *= pi[column];
      if (row == column) p += e_mu_t;
      sP[row * MP + column] = p;
      Pout[row * MP + column] = p;
    }
    __syncthreads();
    tip_inner_update(ed, sP, job.bin, M, MP, K);
    return;
  }
  // Large alphabets (codons): stage W = evT * exp(t roots) and evI in shared memory so the
  // M^3 contraction reads shared memory (odd row stride: conflict free) instead of
  // uncoalesced global memory.  sW / sI live behind sP in the dynamic allocation.
  const bool staged = M > 8;
  const int LD = M | 1;  // odd stride
  double* sW = s_ei + M;             // [M*LD] (real) or [2*M*LD] (complex: re, im planes)
  if (job.mode == EXPM_EIGEN_COMPLEX) {
    const double* roots = slot;
    const double* evT = slot + 2 * M;
    const double* evI = evT + 2 * MM;
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
      const double mag = exp(t * roots[2 * k]);
      double s, c;
      sincos(t * roots[2 * k + 1], &s, &c);
      s_er[k] = mag * c;
      s_ei[k] = mag * s;
    }
    __syncthreads();
    if (staged) {
      double* sWr = sW;
      double* sWi = sW + M * LD;
      double* sIr = sWi + M * LD;
      double* sIi = sIr + M * LD;
      for (int o = threadIdx.x; o < MM; o += blockDim.x) {
        const int i = o / M, k = o - i * M;
        const double ar = evT[2 * o], ai = evT[2 * o + 1];
        const double er = s_er[k], ei = s_ei[k];
        sWr[i * LD + k] = ar * er - ai * ei;  // w = evT[i,k] * exp_roots[k]
        sWi[i * LD + k] = ar * ei + ai * er;
        sIr[i * LD + k] = evI[2 * o];
        sIi[i * LD + k] = evI[2 * o + 1];
      }
      __syncthreads();
      // 2 x 2 register tiles: each shared-memory operand feeds two outputs (the contraction is bound
      // by its shared-memory loads); every element still sums k = 0..M-1 in order
      const int T = (M + 1) / 2;
      for (int o = threadIdx.x; o < T * T; o += blockDim.x) {
        // rows (i0, i0 + T) x columns (j0, j0 + T): consecutive lanes read consecutive rows of evI,
        // LD (odd) doubles apart -- conflict free; adjacent pairs (stride 2 LD) were 2-way conflicts
        const int i0 = o / T, j0 = o % T;
        const int i1 = min(i0 + T, M - 1), j1 = min(j0 + T, M - 1);
        double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
        for (int k = 0; k < M; ++k) {
          const double w0r = sWr[i0 * LD + k], w0i = sWi[i0 * LD + k];
          const double w1r = sWr[i1 * LD + k], w1i = sWi[i1 * LD + k];
          const double v0r = sIr[j0 * LD + k], v0i = sIi[j0 * LD + k];
          const double v1r = sIr[j1 * LD + k], v1i = sIi[j1 * LD + k];
          a00 += w0r * v0r - w0i * v0i;
          a01 += w0r * v1r - w0i * v1i;
          a10 += w1r * v0r - w1i * v0i;
          a11 += w1r * v1r - w1i * v1i;
        }
        const double r[2][2] = {{clamp0(a00), clamp0(a01)}, {clamp0(a10), clamp0(a11)}};
        for (int di = 0; di < 2; ++di)
          for (int dj = 0; dj < 2; ++dj)
            if (i0 + di * T < M && j0 + dj * T < M) {
              sP[(i0 + di * T) * MP + j0 + dj * T] = r[di][dj];
              Pout[(i0 + di * T) * MP + j0 + dj * T] = r[di][dj];
            }
      }
    } else {
      for (int o = threadIdx.x; o < MM; o += blockDim.x) {
        const int i = o / M, j = o - i * M;
        double acc = 0.0;
        for (int k = 0; k < M; ++k) {
          const double ar = evT[2 * (i * M + k)], ai = evT[2 * (i * M + k) + 1];
          const double er = s_er[k], ei = s_ei[k];
          // w = evT[i,k] * exp_roots[k]   (numpy: self.evT * exp_roots)
          const double wr = ar * er - ai * ei;
          const double wi = ar * ei + ai * er;
          const double br = evI[2 * (j * M + k)], bi = evI[2 * (j * M + k) + 1];
          // Re(w * evI[j,k])             (numpy.inner(...).real)
          acc += wr * br - wi * bi;
        }
        acc = clamp0(acc);
        sP[i * MP + j] = acc;
        Pout[i * MP + j] = acc;
      }
    }
  } else {
    const double* roots = slot;
    const double* evT = slot + 2 * M;
    const double* evI = evT + 2 * MM;
    for (int k = threadIdx.x; k < M; k += blockDim.x) s_er[k] = exp(t * roots[k]);
    __syncthreads();
    if (staged) {
      double* sI = sW + M * LD;
      for (int o = threadIdx.x; o < MM; o += blockDim.x) {
        const int i = o / M, k = o - i * M;
        sW[i * LD + k] = evT[o] * s_er[k];
        sI[i * LD + k] = evI[o];
      }
      __syncthreads();
      const int T = (M + 1) / 2;  // 2 x 2 register tiles, see the complex branch
      for (int o = threadIdx.x; o < T * T; o += blockDim.x) {
        // rows (i0, i0 + T) x columns (j0, j0 + T): consecutive lanes read consecutive rows of evI,
        // LD (odd) doubles apart -- conflict free; adjacent pairs (stride 2 LD) were 2-way conflicts
        const int i0 = o / T, j0 = o % T;
        const int i1 = min(i0 + T, M - 1), j1 = min(j0 + T, M - 1);
        double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
#pragma unroll 4
        for (int k = 0; k < M; ++k) {
          const double w0 = sW[i0 * LD + k], w1 = sW[i1 * LD + k];
          const double v0 = sI[j0 * LD + k], v1 = sI[j1 * LD + k];
          a00 += w0 * v0;
          a01 += w0 * v1;
          a10 += w1 * v0;
          a11 += w1 * v1;
        }
        const double r[2][2] = {{clamp0(a00), clamp0(a01)}, {clamp0(a10), clamp0(a11)}};
        for (int di = 0; di < 2; ++di)
          for (int dj = 0; dj < 2; ++dj)
            if (i0 + di * T < M && j0 + dj * T < M) {
              sP[(i0 + di * T) * MP + j0 + dj * T] = r[di][dj];
              Pout[(i0 + di * T) * MP + j0 + dj * T] = r[di][dj];
            }
      }
    } else {
      for (int o = threadIdx.x; o < MM; o += blockDim.x) {
        const int i = o / M, j = o - i * M;
        const double acc = expm_real_entry(evT + i * M, evI + j * M, s_er, M);
        sP[i * MP + j] = acc;
        Pout[i * MP + j] = acc;
      }
    }
  }
  // zero the padding of sP so the tip table never sees garbage
  for (int o = threadIdx.x; o < MP * MP; o += blockDim.x) {
    const int i = o / MP, j = o - i * MP;
    if (i >= M || j >= M) sP[o] = 0.0;
  }
  __syncthreads();
  tip_inner_update(ed, sP, job.bin, M, MP, K);
}

// C = A . B for M x M matrices with row stride MP held in shared memory
__device__ __forceinline__ void smem_matmul(double* __restrict__ C, const double* __restrict__ A,
                                            const double* __restrict__ B, int M, int MP) {
  for (int o = threadIdx.x; o < M * M; o += blockDim.x) {
    const int i = o / M, j = o - i * M;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s = fma(A[i * MP + k], B[k * MP + j], s);
    C[i * MP + j] = s;
  }
}

__global__ void expm_pade_kernel(const ExpmJob* __restrict__ jobs, const double* __restrict__ slots,
                                 int64_t slot_stride, const EdgeDesc* __restrict__ edges, int M,
                                 int MP, int K) {
  extern __shared__ double smem[];
  const int S = MP * MP;
  double* A = smem;
  double* X = A + S;
  double* T = X + S;
  double* N = T + S;
  double* D = N + S;
  __shared__ double s_red[64];
  __shared__ int s_j, s_q, s_piv;
  __shared__ double s_norm;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // see pdl_trigger()
  const ExpmJob job = jobs[blockIdx.x];
  const EdgeDesc ed = edges[job.node];
  double* Pout = ed.P + (int64_t)job.bin * S;
  const double* Q = slots + (int64_t)job.slot * slot_stride;
  const double t = job.t;
  const int tid = threadIdx.x;

  for (int o = tid; o < S; o += blockDim.x) {
    const int i = o / MP, j = o - i * MP;
    A[o] = (i < M && j < M) ? Q[i * M + j] * t : 0.0;  // A = self.Q * t  (:94)
  }
  __syncthreads();
  // norm = max row sum of |A|  (:97)
  if (tid < M) {
    double s = 0.0;
    for (int j = 0; j < M; ++j) s += fabs(A[tid * MP + j]);
    s_red[tid] = s;
  }
  __syncthreads();
  if (tid == 0) {
    double norm = s_red[0];
    for (int i = 1; i < M; ++i) norm = fmax(norm, s_red[i]);
    // j = int(floor(log(max(norm, 0.5)) / log(2.0))) + 1   (:98)
    const int j = (int)floor(log(fmax(norm, 0.5)) / log(2.0)) + 1;
    // order q: while e > 1e-12 ...  (:103-109)
    double e = 1.0, qf = 1.0;
    int q = 0;
    const double scaled = norm / exp2((double)j);
    while (e > 1e-12) {
      q += 1;
      const double q2 = 2.0 * q;
      qf *= ((double)q * q) / (q2 * (q2 - 1) * q2 * (q2 + 1));
      e = 8 * pow(scaled, 2.0 * q) * qf;
    }
    s_j = j;
    s_q = q;
    s_norm = norm;
  }
  __syncthreads();
  const int jj = s_j, q = s_q;
  const double inv_scale = exp2((double)jj);
  for (int o = tid; o < S; o += blockDim.x) A[o] = A[o] / inv_scale;  // A = A / 2.0**j (:99)
  __syncthreads();
  double c = 0.5;
  for (int o = tid; o < S; o += blockDim.x) {
    const int i = o / MP, j = o - i * MP;
    const double id = (i == j && i < M) ? 1.0 : 0.0;
    X[o] = A[o];
    N[o] = id + c * A[o];
    D[o] = id - c * A[o];
  }
  __syncthreads();
  for (int k = 2; k <= q; ++k) {
    c = c * (q - k + 1) / (double)(k * (2 * q - k + 1));
    smem_matmul(T, A, X, M, MP);  // X = dot(A, X)
    __syncthreads();
    double* tmp = X;
    X = T;
    T = tmp;
    const bool plus = (k % 2) == 0;
    for (int o = tid; o < M * M; o += blockDim.x) {
      const int i = o / M, j = o - i * M;
      const double cX = c * X[i * MP + j];
      N[i * MP + j] += cX;
      D[i * MP + j] = plus ? D[i * MP + j] + cX : D[i * MP + j] - cX;
    }
    __syncthreads();
  }
  // F = solve(D, N): LU with partial pivoting (what numpy.linalg.solve -> LAPACK dgesv does)
  for (int k = 0; k < M; ++k) {
    if (tid == 0) {
      int piv = k;
      double best = fabs(D[k * MP + k]);
      for (int i = k + 1; i < M; ++i) {
        const double v = fabs(D[i * MP + k]);
        if (v > best) { best = v; piv = i; }
      }
      s_piv = piv;
    }
    __syncthreads();
    const int piv = s_piv;
    if (piv != k) {
      for (int j = tid; j < 2 * M; j += blockDim.x) {
        double* Mx = j < M ? D : N;
        const int jc = j < M ? j : j - M;
        const double a = Mx[k * MP + jc];
        Mx[k * MP + jc] = Mx[piv * MP + jc];
        Mx[piv * MP + jc] = a;
      }
      __syncthreads();
    }
    const double pivot = D[k * MP + k];
    for (int i = k + 1 + tid; i < M; i += blockDim.x) D[i * MP + k] = D[i * MP + k] / pivot;
    __syncthreads();
    const int rows = M - k - 1;
    for (int o = tid; o < rows * (2 * M); o += blockDim.x) {
      const int i = k + 1 + o / (2 * M);
      const int j = o % (2 * M);
      const double l = D[i * MP + k];
      if (j < M) {
        if (j > k) D[i * MP + j] -= l * D[k * MP + j];
      } else {
        N[i * MP + (j - M)] -= l * N[k * MP + (j - M)];
      }
    }
    __syncthreads();
  }
  // back substitution, one right-hand-side column per thread
  for (int j = tid; j < M; j += blockDim.x) {
    for (int i = M - 1; i >= 0; --i) {
      double s = N[i * MP + j];
      for (int m = i + 1; m < M; ++m) s -= D[i * MP + m] * N[m * MP + j];
      N[i * MP + j] = s / D[i * MP + i];
    }
  }
  __syncthreads();
  double* F = N;
  double* G = D;
  for (int k = 1; k <= jj; ++k) {  // F = dot(F, F), j times (:127-128)
    smem_matmul(G, F, F, M, MP);
    __syncthreads();
    double* tmp = F;
    F = G;
    G = tmp;
  }
  // the Pade path is NOT clamped at zero by the reference (only the eigen path is)
  for (int o = tid; o < S; o += blockDim.x) {
    const int i = o / MP, j = o - i * MP;
    const double v = (i < M && j < M) ? F[o] : 0.0;
    A[o] = v;
    if (i < M && j < M) Pout[o] = v;
  }
  __syncthreads();
  tip_inner_update(ed, A, job.bin, M, MP, K);
}

}  // namespace c3
