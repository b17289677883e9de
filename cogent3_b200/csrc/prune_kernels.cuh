// Felsenstein pruning over compressed site patterns, one launch per tree level.
//
// For every dirty internal node n of the level and every pattern row p, bin b:
//     plh_n[p, b, :] = prod_c  ( P_c,b . plh_c[idx_c[p], b, :] )
// i.e. the reference's per-edge numpy.inner(child_plh, psub)
// (evolve/likelihood_calculation.py:86) FUSED with the gather-product
// sum_input_likelihoods (evolve/likelihood_tree_numba.py:7-25): the [P_child x M]
// "inner" intermediates are never materialised; each child row is read once
// (gathered through idx, 4 B per parent row), each parent row written once.
// Children are multiplied in tree order like the reference (child 0 assigns, the
// others multiply).  The fixed-motif mask of
// PartialLikelihoodProductDefnFixedMotif (likelihood_calculation.py:50-59) and, at
// the root, lh = numpy.inner(plh_root, pi) (:147-148) are fused in.
//
// Layout: partial rows are [row][bin][MP] doubles, so the K rate classes of one
// pattern are contiguous (K*M*8 = 128 B for GTR+G4: one gathered row = one cache
// line, the gather index is read once for all bins).
//
// PRODUCTION kernels (this file):
//   prune4_reg2_kernel     M = 4, K in {1, 2, 4}, 2 or 3 children: warp-private register pipeline with two
//                          strips in flight; HBM-bound (the streaming levels and the root)
//   prune4_small_kernel    M = 4: consecutive SMALL levels in one cluster launch (+ the reduction when the
//                          root is in it) -- real-data sizes are latency bound
//   prune4_kernel          M = 4, any K / any number of children: one thread per (row, bin)
//   prune_dmma_edge_kernel any M <= 64 padded to MP: edge-major FP64 tensor-core (DMMA) contraction
//   lnl_reduce_kernel      site-weighted reduction (+ the fused sum over ranks)
//   rescale_rows_kernel, root_exp_kernel, compose_map_kernel   per-pattern underflow rescaling
// Superseded families and measured-and-dropped experiments live in lab_kernels.cuh (included at the end).
#pragma once
#include "c3_types.cuh"

namespace c3 {

struct __align__(32) d4 {
  double x, y, z, w;
};

__device__ __forceinline__ d4 ld256_nc(const double* p) {
  d4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
               : "l"(p));
  return r;
}
// plain (coherent) 256-bit load: the rows may have been written earlier in this launch
__device__ __forceinline__ d4 ld256(const double* p) {
  d4 r;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st256(double* p, const d4& r) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.x), "d"(r.y), "d"(r.z),
               "d"(r.w)
               : "memory");
}

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor drains; pdl_wait() blocks until the predecessor
// has completed and its writes are visible, pdl_trigger() lets the successor start early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ int find_work(const int32_t* __restrict__ prefix, int n_work, int t) {
  int lo = 0, hi = n_work;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (prefix[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------
// 4-state kernel
// ---------------------------------------------------------------------------------------
constexpr int P4_MAXC = 4;      // children whose P is staged in shared memory
constexpr int P4_MAXK = 8;
constexpr int P4_PSTRIDE = 18;  // 16 + pad: bins land in different banks

__device__ __forceinline__ d4 matvec4(const double* __restrict__ P, const d4& x) {
  d4 y;
  y.x = fma(x.w, P[3], fma(x.z, P[2], fma(x.y, P[1], x.x * P[0])));
  y.y = fma(x.w, P[7], fma(x.z, P[6], fma(x.y, P[5], x.x * P[4])));
  y.z = fma(x.w, P[11], fma(x.z, P[10], fma(x.y, P[9], x.x * P[8])));
  y.w = fma(x.w, P[15], fma(x.z, P[14], fma(x.y, P[13], x.x * P[12])));
  return y;
}

template <int NC, int UNROLL, bool ROOT>
__device__ __forceinline__ void prune4_tile(const PruneArgs& a, const NodeDesc& nd,
                                            const ChildRef* __restrict__ s_child,
                                            const double* __restrict__ sP, int64_t row0, int rl,
                                            int b, int K, int rows_per_pass) {
  int64_t p[UNROLL];
  bool ok[UNROLL];
  int32_t r[UNROLL][NC];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    p[u] = row0 + (int64_t)u * rows_per_pass + rl;
    ok[u] = p[u] < nd.n_rows;
    const int64_t pp = ok[u] ? p[u] : nd.n_rows - 1;
#pragma unroll
    for (int c = 0; c < NC; ++c) r[u][c] = __ldg(s_child[c].idx + pp);
  }
  d4 x[UNROLL][NC];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u)
#pragma unroll
    for (int c = 0; c < NC; ++c)
      x[u][c] = ld256_nc(s_child[c].src + ((int64_t)r[u][c] * K + b) * 4);
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    d4 acc = matvec4(sP + (0 * K + b) * P4_PSTRIDE, x[u][0]);
#pragma unroll
    for (int c = 1; c < NC; ++c) {
      const d4 y = matvec4(sP + (c * K + b) * P4_PSTRIDE, x[u][c]);
      acc.x *= y.x; acc.y *= y.y; acc.z *= y.z; acc.w *= y.w;
    }
    if (nd.fixed_motif >= 0) {
      if (nd.fixed_motif != 0) acc.x = 0.0;
      if (nd.fixed_motif != 1) acc.y = 0.0;
      if (nd.fixed_motif != 2) acc.z = 0.0;
      if (nd.fixed_motif != 3) acc.w = 0.0;
    }
    if (ok[u]) {
      if (!ROOT || nd.out != nullptr) st256(nd.out + (p[u] * K + b) * 4, acc);  // (root rows: see c3_engine::store_root)
      if (ROOT) {
        // lh = numpy.inner(plh_root, pi)
        const double lh = fma(acc.w, a.pi[3], fma(acc.z, a.pi[2], fma(acc.y, a.pi[1], acc.x * a.pi[0])));
        a.lh_out[p[u] * K + b] = lh;
      }
    }
  }
}

// generic child count (polytomies): children handled one after the other
template <int UNROLL, bool ROOT>
__device__ __forceinline__ void prune4_tile_generic(const PruneArgs& a, const NodeDesc& nd,
                                                    const double* __restrict__ sP, bool p_in_smem,
                                                    int64_t row0, int rl, int b, int K,
                                                    int rows_per_pass) {
#pragma unroll 1
  for (int u = 0; u < UNROLL; ++u) {
    const int64_t p = row0 + (int64_t)u * rows_per_pass + rl;
    if (p >= nd.n_rows) continue;
    d4 acc;
#pragma unroll 1
    for (int c = 0; c < nd.n_children; ++c) {
      const ChildRef ch = a.childs[nd.child_begin + c];
      const int32_t r = __ldg(ch.idx + p);
      const d4 x = ld256_nc(ch.src + ((int64_t)r * K + b) * 4);
      d4 y;
      if (p_in_smem) {
        y = matvec4(sP + (c * K + b) * P4_PSTRIDE, x);
      } else if (ch.P != nullptr) {
        y = matvec4(ch.P + b * 16, x);
      } else {
        y = x;
      }
      if (c == 0) acc = y;
      else { acc.x *= y.x; acc.y *= y.y; acc.z *= y.z; acc.w *= y.w; }
    }
    if (nd.fixed_motif >= 0) {
      if (nd.fixed_motif != 0) acc.x = 0.0;
      if (nd.fixed_motif != 1) acc.y = 0.0;
      if (nd.fixed_motif != 2) acc.z = 0.0;
      if (nd.fixed_motif != 3) acc.w = 0.0;
    }
    if (!ROOT || nd.out != nullptr) st256(nd.out + (p * K + b) * 4, acc);
    if (ROOT) {
      const double lh = fma(acc.w, a.pi[3], fma(acc.z, a.pi[2], fma(acc.y, a.pi[1], acc.x * a.pi[0])));
      a.lh_out[p * K + b] = lh;
    }
  }
}

template <int KT, bool ROOT>
__global__ void __launch_bounds__(PRUNE4_THREADS, PRUNE4_MIN_BLOCKS) prune4_kernel(const PruneArgs a) {
  constexpr int UNROLL = PRUNE4_UNROLL;
  const int K = KT > 0 ? KT : a.K;
  const int rows_per_pass = PRUNE4_THREADS / K;
  const int rows_per_tile = rows_per_pass * UNROLL;
  __shared__ ChildRef s_child[P4_MAXC];
  __shared__ double sP[P4_MAXC * P4_MAXK * P4_PSTRIDE];
  const int tid = threadIdx.x;
  const int b = tid % K;
  const int rl = tid / K;
  const bool active = rl < rows_per_pass;

  for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
    const int w = find_work(a.tile_prefix, a.n_work, t);
    const NodeDesc nd = a.nodes[a.work_node[w]];
    const int64_t row0 = (int64_t)(t - a.tile_prefix[w]) * rows_per_tile;
    const int nc = nd.n_children;
    const bool p_in_smem = nc <= P4_MAXC && K <= P4_MAXK;
    __syncthreads();
    if (p_in_smem) {
      if (tid < nc) s_child[tid] = a.childs[nd.child_begin + tid];
      for (int o = tid; o < nc * K * 16; o += PRUNE4_THREADS) {
        const int c = o / (K * 16);
        const int rem = o - c * (K * 16);
        const int bb = rem >> 4, e = rem & 15;
        const double* P = a.childs[nd.child_begin + c].P;
        // pre-multiplied children (tips) use the identity: 1*x + 0*... is exact
        sP[(c * K + bb) * P4_PSTRIDE + e] = P ? P[bb * 16 + e] : ((e % 5 == 0) ? 1.0 : 0.0);
      }
    }
    __syncthreads();
    if (!active) continue;
    if (p_in_smem && nc == 2) {
      prune4_tile<2, UNROLL, ROOT>(a, nd, s_child, sP, row0, rl, b, K, rows_per_pass);
    } else {
      prune4_tile_generic<UNROLL, ROOT>(a, nd, sP, p_in_smem, row0, rl, b, K, rows_per_pass);
    }
  }
}

constexpr int P4S_U = 4;         // rows per lane per strip
constexpr int P4S_PSTRIDE = 18;  // doubles per P matrix in shared memory (16 + pad)

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async16_cg(unsigned smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
// .ca: allocate in L1.  Small child tables (tips, shallow nodes) are gathered from by every
// row of a large parent; with .cg all SMs hammer the same few L2 lines (measured 6x slower).
__device__ __forceinline__ void cp_async16_ca(unsigned smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Descriptors of the launch's work items, staged once per CTA in shared memory: the
// per-strip node lookup, the child pointers and the P staging then never wait on global
// memory (measured: ~20 us of dependent-load latency per launch before this).
constexpr int P4S_MAXW = 64;  // work items (dirty nodes) per strip launch; the host splits longer lists

template <int NC>
struct WorkDesc {
  double* out;
  int64_t n_rows;
  const double* src[NC];
  const int32_t* idx[NC];
  const double* P[NC];
  int32_t fixed_motif;
  int32_t pad;
};

template <int NC>
struct StripSmem {
  WorkDesc<NC> work[P4S_MAXW];
  int32_t prefix[P4S_MAXW + 1];
};

// developer timeline (PruneArgs::tl, C3_TIMELINE=1): when does a launch enter, start gathering, lose its first
// warp, lose its last one -- in one time base across the launches of an UN-profiled (graph-replayed) step
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tl_first(unsigned long long* tl, int slot) {
  if (tl != nullptr) atomicMin(tl + slot, tl_now());
}
__device__ __forceinline__ void tl_end(unsigned long long* tl) {
  if (tl != nullptr) {
    const unsigned long long t = tl_now();
    atomicMin(tl + 2, t);
    atomicMax(tl + 3, t);
  }
}

// ---------------------------------------------------------------------------------------
// 4-state kernel, register pipeline with TWO strips in flight ("regpipe2", binary nodes) -- the
// production kernel for the bulk of a tree.
// ncu on regpipe: DRAM 57 %, issue slots 26 %, 8 warps per SM, the stall samples sit on the first
// use of the gathered rows: with one strip (8 KB per warp, 64 KB per SM) in flight against ~2 us
// of loaded latency, Little's law gives the 4.8 TB/s it reaches.
// Each warp keeps two row buffers and two gather-index sets and walks its strips two at a time:
// while strip s is computed from buffer A, strip s+GW is in flight into B; A is refilled IN
// PLACE, child by child, with strip s+2GW as soon as the FMAs that consume it are issued, using
// indexes that were requested two steps earlier (a first attempt kept ONE index set, i.e. one
// step of lead: with the shorter steps the refill loads then stalled on their own addresses in
// the middle of the FMA block and the kernel was 9 % slower than regpipe).  Largest C2 level:
// 4.78 -> 5.29 TB/s; C5 levels 5.7-6.0 TB/s (the measured copy bandwidth is 6.55).  248 registers.
// ---------------------------------------------------------------------------------------
// Instantiated as <K, 2 children, U = 4 rows per lane, 8 warps> (248 registers) and, for the
// trifurcating root, <K, 3, 2, 12> (strips of 64 units, 12 warps per SM instead of regpipe's 6).
template <int K, int NC, int U, int WARPS, bool ROOT, bool DYN = false>
__global__ void __launch_bounds__(WARPS * 32, 1) prune4_reg2_kernel(const PruneArgs a) {
  constexpr int ROWS = 32 * U / K;  // pattern rows per strip
  __shared__ StripSmem<NC> sd;
  __shared__ __align__(16) double sPall[WARPS][NC * K * P4S_PSTRIDE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sP = sPall[warp];

  pdl_trigger();
  if (threadIdx.x == 0) tl_first(a.tl, 0);
  for (int i = threadIdx.x; i <= a.n_work; i += blockDim.x) sd.prefix[i] = a.tile_prefix[i];
  for (int i = threadIdx.x; i < a.n_work; i += blockDim.x) {
    const NodeDesc nd = a.nodes[a.work_node[i]];
    WorkDesc<NC>& wd = sd.work[i];
    wd.out = nd.out;
    wd.n_rows = nd.n_rows;
    wd.fixed_motif = nd.fixed_motif;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const ChildRef ch = a.childs[nd.child_begin + c];
      wd.src[c] = ch.src;
      wd.idx[c] = ch.idx;
      wd.P[c] = ch.P;
    }
  }
  __syncthreads();

  const int gw = blockIdx.x * WARPS + warp;
  const int GW = gridDim.x * WARPS;
  const int total = a.total_tiles;
  const int b = lane % K;
  const int slot = lane / K;
  // lean root: the rate classes of a row sit in K neighbouring lanes -- mix them here (the reduction's
  // arithmetic: products added in class order starting from 0.0) and write 8 bytes per row instead of 8 K
  const bool lean = ROOT && K > 1 && a.mix_out != nullptr;
  double bp[K];
#pragma unroll
  for (int j = 0; j < K; ++j) bp[j] = lean ? a.bprobs[j] : 0.0;

  int i_w = 0, g_w = 0, c_w = 0, p_w = -1;
  auto load_idx = [&](int32_t(&ix)[NC][U], int s) {
    while (s >= sd.prefix[i_w + 1]) ++i_w;
    const WorkDesc<NC>& wd = sd.work[i_w];
    const int64_t row0 = (int64_t)(s - sd.prefix[i_w]) * ROWS;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t p = row0 + u * (32 / K) + slot;
      if (p >= wd.n_rows) p = wd.n_rows - 1;
#pragma unroll
      for (int c = 0; c < NC; ++c) ix[c][u] = __ldg(wd.idx[c] + p);
    }
  };
  const double* gsrc[NC];
  auto gather_seek = [&](int s) {
    while (s >= sd.prefix[g_w + 1]) ++g_w;
#pragma unroll
    for (int c = 0; c < NC; ++c) gsrc[c] = sd.work[g_w].src[c];
  };

  // compute strip s from x; refill x with strip s2 (= s + 2 GW; indexes ix, requested two steps ago);
  // then request the indexes of strip s4 (= s + 4 GW) into ix
  auto step = [&](d4(&x)[NC][U], int32_t(&ix)[NC][U], int s, int s2, int s4) {
    while (s >= sd.prefix[c_w + 1]) ++c_w;
    const WorkDesc<NC>& wd = sd.work[c_w];
    if (c_w != p_w) {
      __syncwarp();
      for (int o = lane; o < NC * K * 16; o += 32) {
        const int c = o / (K * 16);
        const int rem = o - c * (K * 16);
        const double* P = wd.P[c];
        sP[(c * K + (rem >> 4)) * P4S_PSTRIDE + (rem & 15)] =
            P ? P[rem] : (((rem & 15) % 5 == 0) ? 1.0 : 0.0);
      }
      p_w = c_w;
      __syncwarp();
    }
    const bool refill = s2 < total;
    if (refill) gather_seek(s2);
    d4 acc[U];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double* Pc = sP + (c * K + b) * P4S_PSTRIDE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 p01 = *reinterpret_cast<const double2*>(Pc + 4 * i);
        const double2 p23 = *reinterpret_cast<const double2*>(Pc + 4 * i + 2);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double y =
              fma(x[c][u].w, p23.y, fma(x[c][u].z, p23.x, fma(x[c][u].y, p01.y, x[c][u].x * p01.x)));
          double& t = (i == 0) ? acc[u].x : (i == 1) ? acc[u].y : (i == 2) ? acc[u].z : acc[u].w;
          t = (c == 0) ? y : t * y;
        }
      }
      if (refill) {
#pragma unroll
        for (int u = 0; u < U; ++u) x[c][u] = ld256_nc(gsrc[c] + ((int64_t)ix[c][u] * K + b) * 4);
      }
    }
    if (s4 < total) load_idx(ix, s4);
    const int64_t node_row0 = (int64_t)(s - sd.prefix[c_w]) * ROWS;
    const int c_fixed = wd.fixed_motif;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = node_row0 + u * (32 / K) + slot;
      if (c_fixed >= 0) {
        if (c_fixed != 0) acc[u].x = 0.0;
        if (c_fixed != 1) acc[u].y = 0.0;
        if (c_fixed != 2) acc[u].z = 0.0;
        if (c_fixed != 3) acc[u].w = 0.0;
      }
      if (p < wd.n_rows && (!ROOT || wd.out != nullptr)) st256(wd.out + (p * K + b) * 4, acc[u]);
      if (ROOT) {
        const double lh =
            fma(acc[u].w, a.pi[3], fma(acc[u].z, a.pi[2], fma(acc[u].y, a.pi[1], acc[u].x * a.pi[0])));
        if (!lean) {
          if (p < wd.n_rows) a.lh_out[p * K + b] = lh;
        } else {
          double m = 0.0;
#pragma unroll
          for (int j = 0; j < K; ++j)
            m = __dadd_rn(m, __dmul_rn(__shfl_sync(0xffffffffu, lh, slot * K + j), bp[j]));
          if (b == 0 && p < wd.n_rows) a.mix_out[p] = m;
        }
      }
    }
  };

  d4 xa[NC][U], xb[NC][U];
  int32_t ia[NC][U], ib[NC][U];
  // prologue: strips gw, gw+GW into the buffers; indexes of gw+2GW, gw+3GW into the index sets
  if (gw < total) load_idx(ia, gw);  // static data: may overlap the previous launch
  if (gw + GW < total) load_idx(ib, gw + GW);
  pdl_wait();
  if (threadIdx.x == 0) tl_first(a.tl, 1);
  if (gw < total) {
    gather_seek(gw);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int u = 0; u < U; ++u) xa[c][u] = ld256_nc(gsrc[c] + ((int64_t)ia[c][u] * K + b) * 4);
  }
  if (gw + GW < total) {
    gather_seek(gw + GW);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int u = 0; u < U; ++u) xb[c][u] = ld256_nc(gsrc[c] + ((int64_t)ib[c][u] * K + b) * 4);
  }
  if (gw + 2 * GW < total) load_idx(ia, gw + 2 * GW);
  if (gw + 3 * GW < total) load_idx(ib, gw + 3 * GW);

  if (!DYN) {
#pragma unroll 1
    for (int s = gw; s < total; s += 2 * GW) {
      step(xa, ia, s, s + 2 * GW, s + 4 * GW);
      if (s + GW < total) step(xb, ib, s + GW, s + 3 * GW, s + 5 * GW);
    }
  } else {
    // DYNAMIC strips (C3_DYN=1): the first five strips of a warp are the static ones (gw + k GW), every further
    // one is claimed from a counter (a.strip_ctr, zeroed with the parameter block) one step before its indexes
    // are requested -- five steps before it is computed.  A warp on a slow SM then simply claims fewer strips: the
    // 4-9 us between the first and the last warp leaving a launch (lesson 17a) are what this is after.  Claims
    // grow monotonically, so the forward cursors keep working; per-strip arithmetic is untouched: bit-identical.
    int w0 = gw, w1 = gw + GW, w2 = gw + 2 * GW, w3 = gw + 3 * GW, w4 = gw + 4 * GW;
    int claim = 0;
    auto claim_next = [&]() {
      if (lane == 0) claim = atomicAdd(a.strip_ctr, 1);
    };
    auto rotate = [&]() {
      w0 = w1; w1 = w2; w2 = w3; w3 = w4;
      // (a negative sum after an overflow would restart at strip 0: clamp)
      const int c = __shfl_sync(0xffffffffu, claim, 0);
      w4 = (c < total) ? 5 * GW + c : total;
    };
#pragma unroll 1
    while (w0 < total) {
      claim_next();
      step(xa, ia, w0, w2, w4);
      rotate();
      if (w0 >= total) break;
      claim_next();
      step(xb, ib, w0, w2, w4);
      rotate();
    }
  }
  if (lane == 0) tl_end(a.tl);
}

__device__ __forceinline__ unsigned ld_relaxed_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------------
// 4-state kernel for SMALL levels: several consecutive tree levels in ONE launch.
// (Measured and dropped: the whole brca1 evaluation -- expm, 11 levels, reduction -- in one
// single-CTA launch with block barriers: 71 us vs 75 us for the three launches; one SM has to
// issue all ~190k warp instructions, which costs what the cluster barriers cost here.)
//
// Real alignments (brca1: 2763 root patterns, 11 levels) and the shallow levels of big
// ones are latency bound: a launch per level costs ~8 us of dependent-load latency for a
// few thousand rows.  Here one thread-block cluster (8 CTAs = 8 SMs) walks the levels in
// order, one (row, bin) unit per thread per step, with a cluster barrier between levels
// (barrier.cluster release/acquire makes the previous level's rows visible).  Arithmetic
// is identical to the strip kernel (same FMA order), so the results are bit-identical
// whichever kernel a node is assigned to.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

constexpr int SMALL_MAX_P = 192;  // P matrices (work item x staged child x bin) kept in shared memory (the whole
                                  // struct stays near 52 KB: a larger carve-out costs the overlap with the neighbouring launches)
struct SmallSmem {
  SmallLevel levels[SMALL_MAX_LEVELS];
  int32_t prefix[SMALL_MAX_WORK + SMALL_MAX_LEVELS];
  NodeDesc nodes[SMALL_MAX_WORK];
  ChildRef childs[SMALL_MAX_WORK * SMALL_STAGED_CHILDREN];
  double P[SMALL_MAX_P * 16];
};

// One (row, bin) unit of a small level with everything staged: the gather indexes of ALL children
// are requested first, then all their rows, then the arithmetic with P from shared memory -- the
// dependent chain per unit is two global latencies, not three per child.  Same arithmetic
// (matvec4, children multiplied in order) as every other 4-state kernel: bit-identical.
__device__ __forceinline__ void small_unit_staged(const PruneArgs& a, const NodeDesc& nd,
                                                  const ChildRef* __restrict__ ch, const double* __restrict__ sP,
                                                  int64_t p, int b, int K) {
  const int nc = nd.n_children;  // <= SMALL_STAGED_CHILDREN (checked by the caller)
  // the first two children together (1024 threads per CTA leave 64 registers: a third row would spill)
  const int32_t r0 = ch[0].idx[p];
  const int32_t r1 = (nc > 1) ? ch[1].idx[p] : 0;
  const d4 x0 = ld256(ch[0].src + ((int64_t)r0 * K + b) * 4);
  d4 x1 = x0;
  if (nc > 1) x1 = ld256(ch[1].src + ((int64_t)r1 * K + b) * 4);  // in flight together with x0
  d4 acc = (ch[0].P != nullptr) ? matvec4(sP + (0 * K + b) * 16, x0) : x0;
  if (nc > 1) {
    const d4 y = (ch[1].P != nullptr) ? matvec4(sP + (1 * K + b) * 16, x1) : x1;
    acc.x *= y.x; acc.y *= y.y; acc.z *= y.z; acc.w *= y.w;
  }
  for (int c = 2; c < nc; ++c) {
    const int32_t r = ch[c].idx[p];
    const d4 x = ld256(ch[c].src + ((int64_t)r * K + b) * 4);
    const d4 y = (ch[c].P != nullptr) ? matvec4(sP + (c * K + b) * 16, x) : x;
    acc.x *= y.x; acc.y *= y.y; acc.z *= y.z; acc.w *= y.w;
  }
  if (nd.fixed_motif >= 0) {
    if (nd.fixed_motif != 0) acc.x = 0.0;
    if (nd.fixed_motif != 1) acc.y = 0.0;
    if (nd.fixed_motif != 2) acc.z = 0.0;
    if (nd.fixed_motif != 3) acc.w = 0.0;
  }
  if (nd.out != nullptr) st256(nd.out + (p * K + b) * 4, acc);
  if (nd.is_root) {
    const double lh = fma(acc.w, a.pi[3], fma(acc.z, a.pi[2], fma(acc.y, a.pi[1], acc.x * a.pi[0])));
    a.lh_out[p * K + b] = lh;
  }
}

// ---- pieces of the site-weighted reduction shared by lnl_reduce_kernel and the small-levels kernel ----
__device__ __forceinline__ double block_sum(double v, double* s_warp) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  double total = 0.0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) total += s_warp[i];
  }
  return total;  // valid on thread 0
}

// one thread's share of a reduction tile: rows base + u * REDUCE_THREADS + t  (t < REDUCE_THREADS)
__device__ __forceinline__ double reduce_tile_partial(const double* __restrict__ lh, const double* __restrict__ bprobs,
                                                      const double* __restrict__ counts, int64_t n_rows, int K,
                                                      int tile, int t, const int32_t* __restrict__ root_exp,
                                                      const double* __restrict__ mixed = nullptr) {
  constexpr int U = REDUCE_ROWS_PER_TILE / REDUCE_THREADS;
  const int64_t base = (int64_t)tile * REDUCE_ROWS_PER_TILE;
  // all loads of the thread's rows first (one 256-bit / 128-bit load per row for K = 4 / 2), then the logs
  double mix[U], cnt[U];
  int32_t ex[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t p = base + u * REDUCE_THREADS + t;
    mix[u] = 1.0;
    cnt[u] = 0.0;
    ex[u] = 0;
    if (p < n_rows) {
      if (mixed != nullptr) {
        mix[u] = mixed[p];  // (lean root: the root launch mixed the classes, same arithmetic)
      } else if (K == 1) {
        mix[u] = lh[p];
      } else if (K == 4) {
        const d4 v = ld256(lh + p * 4);
        mix[u] = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(v.x, bprobs[0])), __dmul_rn(v.y, bprobs[1])),
                                     __dmul_rn(v.z, bprobs[2])), __dmul_rn(v.w, bprobs[3]));
      } else if (K == 2) {
        const double2 v = *reinterpret_cast<const double2*>(lh + p * 2);
        mix[u] = __dadd_rn(__dadd_rn(0.0, __dmul_rn(v.x, bprobs[0])), __dmul_rn(v.y, bprobs[1]));
      } else {
        double m = 0.0;
        for (int b = 0; b < K; ++b) m = __dadd_rn(m, __dmul_rn(lh[p * K + b], bprobs[b]));
        mix[u] = m;
      }
      cnt[u] = counts[p];
      if (root_exp != nullptr) ex[u] = root_exp[p];
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    // rescaled rows: lh = mix * 2^root_exp[p]  (exact powers of two carried as integers)
    const double lg = (root_exp != nullptr) ? fma((double)ex[u], 0.6931471805599453094, log(mix[u])) : log(mix[u]);
    acc += lg * cnt[u];
  }
  return acc;
}

// the reduction of a small problem done by the small-levels kernel itself (no separate launch)
constexpr int SMALL_REDUCE_MAX_TILES = 8;
__device__ __forceinline__ void publish_result(const ResultSink& r, double total) {
  *r.d_lnl = total;
  if (r.d_lnl2) *r.d_lnl2 = total;
  if (r.h_lnl) {
    // three words into mapped host memory: the value, the sequence number the host spins on, and a check word
    // (value bits ^ sequence).  The host accepts the triple only when it is consistent, so nothing has to order
    // the three stores: no system-scope fence (two PCIe round trips, ~4 us of every evaluation) on this path
    const unsigned long long s = *r.seq_ctr + 1;
    *r.seq_ctr = s;
    r.h_lnl[0] = total;
    r.h_seq[0] = s;
    r.h_seq[1] = (unsigned long long)__double_as_longlong(total) ^ s;
  }
}

struct ReduceTail {
  const double* bprobs;
  const double* counts;
  int64_t n_rows;
  int n_tiles;  // 0: no fused reduction
  ResultSink sink;
  double* tile_sums = nullptr;  // small-levels launch: scratch for the per-tile sums (global memory)
};

// n_work_total / n_prefix_total: sizes of the launch's work and prefix arrays.  When they
// fit, every descriptor the units need is staged in shared memory first, so the
// dependent-load chain per level is just  gather index -> child row  (P and the
// descriptors no longer sit on it).
//
// GRID = true ("medium levels"): the same walk by the WHOLE GPU -- one CTA per SM, cooperative launch, a
// grid barrier (arrive on a counter in device memory, spin until everybody has) between levels.  The
// levels between the small ones and the streaming ones (C2: 1 MB and 15 MB, after two tiny ones) each cost
// a launch of their own ~11-15 us whatever they hold; here a level costs a barrier (~2 us) plus its two
// dependent global latencies.  `bar` is zeroed by the host before the launch.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (ld_acquire_gpu_u32(bar) < target) __nanosleep(32);
  }
  __syncthreads();
}

template <bool GRID>
__global__ void __launch_bounds__(SMALL_THREADS, 1)
prune4_small_kernel(const PruneArgs a, const SmallLevel* __restrict__ levels, int n_levels,
                    int n_work_total, int n_prefix_total, const ReduceTail tail, unsigned* __restrict__ bar) {
  extern __shared__ __align__(16) unsigned char small_raw[];
  SmallSmem& sm = *reinterpret_cast<SmallSmem*>(small_raw);
  pdl_trigger();
  if (threadIdx.x == 0) tl_first(a.tl, 0);
  const int K = a.K;
  const unsigned nthreads = (GRID ? gridDim.x : cluster_nctarank()) * SMALL_THREADS;
  const unsigned tid = (GRID ? blockIdx.x : cluster_ctarank()) * SMALL_THREADS + threadIdx.x;
  bool staged = n_work_total <= SMALL_MAX_WORK && n_levels <= SMALL_MAX_LEVELS;
  if (staged) {
    for (int i = threadIdx.x; i < n_levels; i += SMALL_THREADS) sm.levels[i] = levels[i];
    for (int i = threadIdx.x; i < n_prefix_total; i += SMALL_THREADS) sm.prefix[i] = a.tile_prefix[i];
    for (int i = threadIdx.x; i < n_work_total; i += SMALL_THREADS) {
      const NodeDesc nd = a.nodes[a.work_node[i]];
      sm.nodes[i] = nd;
      for (int c = 0; c < nd.n_children && c < SMALL_STAGED_CHILDREN; ++c)
        sm.childs[i * SMALL_STAGED_CHILDREN + c] = a.childs[nd.child_begin + c];
    }
    __syncthreads();
  }
  pdl_wait();  // P matrices / tip tables of the expm launch, rows of earlier levels
  if (threadIdx.x == 0) tl_first(a.tl, 1);
  // the P matrices of every staged child, per bin, into shared memory
  const bool p_staged = staged && (int64_t)n_work_total * SMALL_STAGED_CHILDREN * K <= SMALL_MAX_P;
  if (p_staged) {
    const int per_item = SMALL_STAGED_CHILDREN * K * 16;
    for (int o = threadIdx.x; o < n_work_total * per_item; o += SMALL_THREADS) {
      const int i = o / per_item, rem = o - i * per_item;
      const int c = rem / (K * 16), e16 = rem - c * (K * 16);  // e16 = bin * 16 + element
      const double* P = (c < sm.nodes[i].n_children) ? sm.childs[i * SMALL_STAGED_CHILDREN + c].P : nullptr;
      if (P != nullptr) sm.P[o] = P[e16];
    }
    __syncthreads();
  }
  for (int li = 0; li < n_levels; ++li) {
    const SmallLevel L = staged ? sm.levels[li] : levels[li];
    const int32_t* prefix = (staged ? sm.prefix : a.tile_prefix) + L.prefix_off;
    for (int u = (int)tid; u < L.total_units; u += (int)nthreads) {
      const int w = find_work(prefix, L.n_work, u);
      const int wi = L.work_off + w;
      const NodeDesc nd = staged ? sm.nodes[wi] : a.nodes[a.work_node[wi]];
      const int local = u - prefix[w];
      const int64_t p = local / K;
      const int b = local - (int)p * K;
      if (p_staged && nd.n_children <= SMALL_STAGED_CHILDREN) {
        small_unit_staged(a, nd, sm.childs + wi * SMALL_STAGED_CHILDREN,
                          sm.P + (int64_t)wi * SMALL_STAGED_CHILDREN * K * 16, p, b, K);
        continue;
      }
      d4 acc;
      for (int c = 0; c < nd.n_children; ++c) {
        const ChildRef ch = (staged && c < SMALL_STAGED_CHILDREN) ? sm.childs[wi * SMALL_STAGED_CHILDREN + c]
                                                                  : a.childs[nd.child_begin + c];
        const int32_t r = ch.idx[p];
        const d4 x = ld256(ch.src + ((int64_t)r * K + b) * 4);
        d4 y;
        if (ch.P != nullptr) {
          y = matvec4(ch.P + b * 16, x);
        } else {
          y = x;  // pre-multiplied rows; same value as the strip kernel's identity product
        }
        if (c == 0) acc = y;
        else { acc.x *= y.x; acc.y *= y.y; acc.z *= y.z; acc.w *= y.w; }
      }
      if (nd.fixed_motif >= 0) {
        if (nd.fixed_motif != 0) acc.x = 0.0;
        if (nd.fixed_motif != 1) acc.y = 0.0;
        if (nd.fixed_motif != 2) acc.z = 0.0;
        if (nd.fixed_motif != 3) acc.w = 0.0;
      }
      if (nd.out != nullptr) st256(nd.out + (p * K + b) * 4, acc);
      if (nd.is_root) {
        const double lh = fma(acc.w, a.pi[3], fma(acc.z, a.pi[2], fma(acc.y, a.pi[1], acc.x * a.pi[0])));
        a.lh_out[p * K + b] = lh;
      }
    }
    if (li + 1 < n_levels) {
      if (GRID) grid_barrier(bar, (unsigned)(li + 1) * gridDim.x);
      else cluster_barrier();
    }
  }
  if (threadIdx.x == 0) tl_end(a.tl);
  if (!GRID && tail.n_tiles > 0) {
    // the root was in this launch and the problem is small: the site-weighted reduction right here
    // (same tiles, same per-thread rows, same summation trees as lnl_reduce_kernel: same bits), by
    // the cluster's first CTA once every CTA's lh values are there
    // -- the tiles dealt to the cluster's CTAs (one 1024-row tile each at brca1's size, instead of the first CTA
    // taking them one after the other), then the first CTA adds the tile sums in index order
    cluster_barrier();
    __shared__ double s_warp[SMALL_THREADS / 32];
    const int t = (int)threadIdx.x;
    const int rank = (int)cluster_ctarank(), n_rank = (int)cluster_nctarank();
    for (int tile = rank; tile < tail.n_tiles; tile += n_rank) {
      const double acc = (t < REDUCE_THREADS)
                             ? reduce_tile_partial(a.lh_out, tail.bprobs, tail.counts, tail.n_rows, K, tile, t, nullptr)
                             : 0.0;
      const double total = block_sum(acc, s_warp);  // (the warps past REDUCE_THREADS add +0.0)
      if (t == 0) tail.tile_sums[tile] = total;
    }
    cluster_barrier();  // (release / acquire at cluster scope: the tile sums are visible to the first CTA)
    if (rank != 0) return;
    double acc = 0.0;
    if (t < REDUCE_THREADS)
      for (int i = t; i < tail.n_tiles; i += REDUCE_THREADS) acc += __ldcg(tail.tile_sums + i);
    const double total = block_sum(acc, s_warp);
    if (t == 0) publish_result(tail.sink, total);
  }
}

// ---------------------------------------------------------------------------------------
// 4-state "chain" kernel: a SINGLE-BRANCH update of a small problem (the optimiser's inner loop: one
// branch length moves, the dirty nodes are the path from that branch to the root) in one launch with NO
// barrier between the levels.
//
// The small-levels launch walks the path level by level: per level a cluster barrier plus two dependent
// global latencies (gather index -> child row) -- 41 us for brca1's 11 levels.  But a pattern row of the root
// only needs ONE row of every node on the path, and which one is static: the composition of the gather
// tables from the root down (map_n[p] = row of node n that root row p looks at, built once per engine by
// compose_map_kernel, n_nodes x root_rows int32).  So the thread of (root row p, bin b) walks the whole
// path by itself: at every path node it gathers the rows of the CLEAN children at map_child[p] (all
// addresses are known up front: every index of the walk is requested before the first row), multiplies
// the row it carries in registers by the P of the branch it came up, stores the node's row at map_n[p]
// (root rows that share a lower row compute -- and store -- the same bits) and moves on; at the root it
// writes lh.  The last CTA to finish reduces (same tiles and summation trees as lnl_reduce_kernel).
// Arithmetic per row is that of every other 4-state kernel (matvec4, children multiplied in order):
// bit-identical to the level launches.
// ---------------------------------------------------------------------------------------
constexpr int CHAIN_MAX_NODES = 16;  // path length (tree depth) one launch walks
constexpr int CHAIN_THREADS = REDUCE_THREADS;
constexpr int CHAIN_MAX_CHILDREN = 3;

struct ChainSmem {
  NodeDesc nodes[CHAIN_MAX_NODES];
  ChildRef childs[CHAIN_MAX_NODES * CHAIN_MAX_CHILDREN];
  const int32_t* map_own[CHAIN_MAX_NODES];
  const int32_t* map_child[CHAIN_MAX_NODES * CHAIN_MAX_CHILDREN];
  int32_t dirty_child[CHAIN_MAX_NODES];
  double P[CHAIN_MAX_NODES * CHAIN_MAX_CHILDREN * 4 * 16];  // [node][child][bin][16], K <= 4
};

// Everything an update needs beyond the resident tables travels BY VALUE in the launch (no parameter-block copy):
// the path bottom-up (work_node, the last one is the root), dirty_child[i] = the child slot of path node i that
// is path node i - 1 (-1 for i = 0), pi, the bin probabilities -- and, when the update is ONE branch whose P has
// the real eigen form (`fused`), the K distances of that branch: every CTA then computes the K matrices itself
// (64 K FMAs, the expm launch's own arithmetic: expm_real_entry), CTA 0 also writes them -- and, below a tip,
// the tip's pre-multiplied table -- back to memory for later evaluations.  A single-branch update is then ONE
// kernel launch: no copy, no expm launch, no graph.
struct ChainParams {
  int32_t work_node[CHAIN_MAX_NODES];
  int32_t dirty_child[CHAIN_MAX_NODES];
  double pi[4];
  double bprobs[4];
  int32_t fused;      // 1: the P of the branch below path node 0's child `edge_slot` is computed here
  int32_t edge_node;  // the node below that branch
  int32_t edge_slot;
  int32_t pad;
  int32_t q_slot[4];  // per bin: the slot holding roots | evT | evI
  double t[4];        // per bin: distance
};

// node_maps[n] / child_maps[child ref]: root row -> row of the node / of the child
__global__ void __launch_bounds__(CHAIN_THREADS)
prune4_chain_kernel(const PruneArgs a, const ChainParams cp,
                    const int32_t* const* __restrict__ node_maps, const int32_t* const* __restrict__ child_maps,
                    const ReduceTail tail, unsigned int* __restrict__ counters, double* __restrict__ tile_sums,
                    const double* __restrict__ slots, int64_t slot_stride, const EdgeDesc* __restrict__ edges) {
  extern __shared__ __align__(16) unsigned char chain_raw[];
  ChainSmem& sm = *reinterpret_cast<ChainSmem*>(chain_raw);
  __shared__ double s_warp[CHAIN_THREADS / 32];
  __shared__ double s_bp[4];
  __shared__ __align__(16) double s_Pnew[4 * 16];  // fused: P of the changed branch, [bin][16]
  __shared__ bool s_last;
  const int K = a.K, n_work = a.n_work, t = (int)threadIdx.x;
  pdl_trigger();
  if (t == 0) tl_first(a.tl, 0);
  for (int i = t; i < n_work; i += CHAIN_THREADS) {
    const int node = cp.work_node[i];
    const NodeDesc nd = a.nodes[node];
    sm.nodes[i] = nd;
    sm.map_own[i] = node_maps[node];
    sm.dirty_child[i] = cp.dirty_child[i];
    for (int c = 0; c < nd.n_children && c < CHAIN_MAX_CHILDREN; ++c) {
      sm.childs[i * CHAIN_MAX_CHILDREN + c] = a.childs[nd.child_begin + c];
      sm.map_child[i * CHAIN_MAX_CHILDREN + c] = child_maps[nd.child_begin + c];
    }
  }
  const bool fused = cp.fused != 0;
  if (fused && t < 16 * K) {
    // P_b = max(evT . diag(exp(t_b roots)) . evI^T, 0): thread (b, i, j)
    const int bin = t >> 4, o = t & 15, i = o >> 2, j = o & 3;
    const double* slot = slots + (int64_t)cp.q_slot[bin] * slot_stride;
    const double* roots = slot;
    const double* evT = slot + 8;    // roots[2 M] | evT[2 M M] | evI[2 M M], M = 4, real parts first
    const double* evI = evT + 32;
    double er[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) er[k] = exp(cp.t[bin] * roots[k]);
    s_Pnew[bin * 16 + o] = expm_real_entry(evT + i * 4, evI + j * 4, er, 4);
  }
  __syncthreads();
  const int64_t total = tail.n_rows * K;
  const int64_t u = (int64_t)blockIdx.x * CHAIN_THREADS + t;
  const bool active = u < total;
  const int64_t p = active ? u / K : 0;
  const int b = active ? (int)(u - p * K) : 0;
  // every index of the walk (static tables: this may overlap the previous launch)
  int32_t own[CHAIN_MAX_NODES], sib[CHAIN_MAX_NODES][CHAIN_MAX_CHILDREN];
#pragma unroll
  for (int i = 0; i < CHAIN_MAX_NODES; ++i) {
    own[i] = 0;
#pragma unroll
    for (int c = 0; c < CHAIN_MAX_CHILDREN; ++c) sib[i][c] = 0;
    if (i < n_work) {
      const NodeDesc& nd = sm.nodes[i];
      if (nd.out != nullptr) own[i] = __ldg(sm.map_own[i] + p);  // (negative: another root row stores this row)
#pragma unroll
      for (int c = 0; c < CHAIN_MAX_CHILDREN; ++c)
        if (c < nd.n_children && c != sm.dirty_child[i])
          sib[i][c] = __ldg(sm.map_child[i * CHAIN_MAX_CHILDREN + c] + p) & 0x7fffffff;
    }
  }
  pdl_wait();  // the expm launch's P matrices / tip tables; the rows and counters of the previous evaluation
  if (t == 0) tl_first(a.tl, 1);
  // pi and the bin probabilities are VALUES, not structure: a launch that is replayed from a CUDA graph (its
  // arguments frozen at capture) must read them from the parameter block, which every replay copies afresh; only
  // the fused launch -- never captured -- carries them by value.  (The path itself is structure: it is part of the
  // graph's signature.)
  double pi[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) pi[q] = fused ? cp.pi[q] : a.pi[q];
  if (t < 4) s_bp[t] = fused ? cp.bprobs[t] : (t < K ? tail.bprobs[t] : 0.0);
  const bool fused_tip = fused && sm.childs[cp.edge_slot].P == nullptr;  // the changed branch ends in a tip
  {
    const int per_item = CHAIN_MAX_CHILDREN * K * 16;
    for (int o = t; o < n_work * per_item; o += CHAIN_THREADS) {
      const int i = o / per_item, rem = o - i * per_item;
      const int c = rem / (K * 16), e16 = rem - c * (K * 16);
      const double* P = (c < sm.nodes[i].n_children) ? sm.childs[i * CHAIN_MAX_CHILDREN + c].P : nullptr;
      if (fused && i == 0 && c == cp.edge_slot) {
        if (P != nullptr) sm.P[(i * CHAIN_MAX_CHILDREN + c) * 64 + e16] = s_Pnew[e16];  // (memory still holds the old P)
      } else if (P != nullptr) {
        sm.P[(i * CHAIN_MAX_CHILDREN + c) * 64 + e16] = P[e16];
      }
    }
  }
  if (fused && blockIdx.x == 0) {
    // for the evaluations to come: the new P and, below a tip, the tip's pre-multiplied table
    const EdgeDesc ed = edges[cp.edge_node];
    if (t < 16 * K) ed.P[t] = s_Pnew[t];
    for (int bin = 0; bin < K; ++bin) tip_inner_update(ed, s_Pnew + bin * 16, bin, 4, 4, K);
  }
  __syncthreads();
  auto gather = [&](int i, d4(&x)[CHAIN_MAX_CHILDREN]) {
    const NodeDesc& nd = sm.nodes[i];
#pragma unroll
    for (int c = 0; c < CHAIN_MAX_CHILDREN; ++c)
      if (c < nd.n_children && c != sm.dirty_child[i])
        x[c] = ld256(sm.childs[i * CHAIN_MAX_CHILDREN + c].src + ((int64_t)sib[i][c] * K + b) * 4);
  };
  d4 carry{0.0, 0.0, 0.0, 0.0};
  d4 xn[CHAIN_MAX_CHILDREN];
  gather(0, xn);
  if (fused_tip) {
    // the changed branch ends in a tip: its row under the NEW P (memory still holds the old table) -- a column
    // of P for a one-hot row, the full sum for an ambiguity code: tip_inner_update's arithmetic
    const EdgeDesc& ed = edges[cp.edge_node];
    int64_t urow = 0;
#pragma unroll
    for (int c = 0; c < CHAIN_MAX_CHILDREN; ++c)
      if (c == cp.edge_slot) urow = sib[0][c];
    const int state = ed.tip_state[urow];
    const double* Pb = s_Pnew + b * 16;
    d4 v;
    v.x = (state >= 0) ? Pb[0 + state] : tip_inner_entry(ed.tip_table + urow * 4, Pb, 0, 4, 4);
    v.y = (state >= 0) ? Pb[4 + state] : tip_inner_entry(ed.tip_table + urow * 4, Pb, 1, 4, 4);
    v.z = (state >= 0) ? Pb[8 + state] : tip_inner_entry(ed.tip_table + urow * 4, Pb, 2, 4, 4);
    v.w = (state >= 0) ? Pb[12 + state] : tip_inner_entry(ed.tip_table + urow * 4, Pb, 3, 4, 4);
#pragma unroll
    for (int c = 0; c < CHAIN_MAX_CHILDREN; ++c)
      if (c == cp.edge_slot) xn[c] = v;
  }
#pragma unroll
  for (int i = 0; i < CHAIN_MAX_NODES; ++i) {
    if (i < n_work) {
      d4 x[CHAIN_MAX_CHILDREN];
#pragma unroll
      for (int c = 0; c < CHAIN_MAX_CHILDREN; ++c) x[c] = xn[c];
      if (i + 1 < n_work) gather(i + 1, xn);  // the next node's clean rows are in flight during this node's arithmetic
      const NodeDesc& nd = sm.nodes[i];
      const int dc = sm.dirty_child[i];
      d4 acc{1.0, 1.0, 1.0, 1.0};
#pragma unroll
      for (int c = 0; c < CHAIN_MAX_CHILDREN; ++c) {
        if (c < nd.n_children) {
          const double* Pc = sm.P + ((i * CHAIN_MAX_CHILDREN + c) * 4 + b) * 16;
          const d4 v = (c == dc) ? carry : x[c];
          const d4 y = (sm.childs[i * CHAIN_MAX_CHILDREN + c].P != nullptr) ? matvec4(Pc, v) : v;
          if (c == 0) acc = y;
          else { acc.x *= y.x; acc.y *= y.y; acc.z *= y.z; acc.w *= y.w; }
        }
      }
      if (nd.fixed_motif >= 0) {
        if (nd.fixed_motif != 0) acc.x = 0.0;
        if (nd.fixed_motif != 1) acc.y = 0.0;
        if (nd.fixed_motif != 2) acc.z = 0.0;
        if (nd.fixed_motif != 3) acc.w = 0.0;
      }
      if (active) {
        if (nd.out != nullptr && own[i] >= 0) st256(nd.out + ((int64_t)own[i] * K + b) * 4, acc);
        if (nd.is_root) {
          const double lh = fma(acc.w, pi[3], fma(acc.z, pi[2], fma(acc.y, pi[1], acc.x * pi[0])));
          a.lh_out[p * K + b] = lh;
        }
      }
      carry = acc;
    }
  }
  if (t == 0) tl_end(a.tl);
  // ---- the site-weighted reduction: a reduction tile (1024 root rows) is the work of 4 K consecutive CTAs; the
  // last of them to get here reduces the tile, the last of those adds the tiles in index order and publishes
  // (same tiles, per-thread rows and summation trees as lnl_reduce_kernel: same bits)
  __threadfence();
  __syncthreads();
  const int ctas_per_tile = REDUCE_ROWS_PER_TILE * K / CHAIN_THREADS;
  const int tile = (int)blockIdx.x / ctas_per_tile;
  const int ctas_here = min(ctas_per_tile, (int)gridDim.x - tile * ctas_per_tile);
  if (t == 0) s_last = atomicAdd(counters + 1 + tile, 1u) == (unsigned)(ctas_here - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    const double acc = reduce_tile_partial(a.lh_out, s_bp, tail.counts, tail.n_rows, K, tile, t, nullptr);
    const double total_t = block_sum(acc, s_warp);
    if (t == 0) {
      tile_sums[tile] = total_t;
      counters[1 + tile] = 0u;
      __threadfence();
      s_last = atomicAdd(counters, 1u) == (unsigned)(tail.n_tiles - 1);
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc = 0.0;
  for (int i = t; i < tail.n_tiles; i += REDUCE_THREADS) acc += __ldcg(tile_sums + i);
  const double result = block_sum(acc, s_warp);
  if (t == 0) {
    counters[0] = 0u;
    publish_result(tail.sink, result);
    tl_end(a.tl);  // (slot 3, the latest stamp: the result is out)
  }
}

// ---------------------------------------------------------------------------------------
// FP64 tensor-core (DMMA) kernel, MP = padded state count (multiple of 8, <= 64)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  cp_async16_cg((unsigned)__cvta_generic_to_shared(smem_dst), gsrc);
}

__device__ __forceinline__ double2 ld128_nc(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------
// DMMA kernel, EDGE-MAJOR ("dmma-edge") -- the production path for M != 4.
//
// The kernels above contract per PARENT row and internal child: 2 MP^2 flops for every
// (parent pattern, internal edge).  cogent3 itself contracts per CHILD row
// (numpy.inner(child_plh, psub), likelihood_calculation.py:86) and gathers afterwards; with
// per-node pattern compression a child has far fewer rows than the parent patterns that
// reference them (C4: 1.66M child rows vs 3.9M parent-row references), so the fused
// parent-major form executes ~1.9x the reference's flops on the unit that bounds the codon
// model.  Here every node stores its rows ALREADY multiplied by the P of the branch above it
// (exactly what tips always did: tip_inner):
//     inner_n[p, b, :] = P_n,b . ( prod_c inner_c[idx_c[p], b, :] )        (root: no P, + lh)
// so a tile is: gather the children's rows (C-fragment layout, 128-bit loads), multiply them,
// ONE contraction with the node's own P, store.  With the permuted k order of dmma-reg the
// product in C-fragment layout IS the A fragment of the contraction (column pair ni of a lane
// = k steps 2ni, 2ni+1): no shuffle, no shared-memory round trip.  Flops: 2 MP^2 per NODE row.
// The numbers are the same as before (the contraction of a child row gives the same bits
// wherever it is done); what changes is when a node is dirty: a changed branch recomputes the
// node BELOW it too.  Pipeline: per warp, the two (first) children's rows of the next tile are in
// flight while the current tile is multiplied and contracted; gather indexes one tile further.
// ---------------------------------------------------------------------------------------
constexpr int DE_WARPS = 12;
constexpr int DE_THREADS = DE_WARPS * 32;
constexpr int DE_TILE_ROWS = DE_WARPS * 8;
constexpr int DE_MAXW = 64;  // work items per launch
constexpr int DE_PRE_MAX = 3;  // children whose rows can be prefetched (further ones: loaded at use)

struct EWork {
  double* out;
  int64_t n_rows;
  int64_t row_tiles;
  const double* P_own;
  const double* src[DE_PRE_MAX];
  const int32_t* idx[DE_PRE_MAX];
  int32_t child_begin;
  int32_t nc;
  int32_t fixed_motif;
  int32_t is_root;
};
struct ESmem {
  EWork work[DE_MAXW];
  int32_t prefix[DE_MAXW + 1];
};
struct ECur {
  int st;        // super-tile
  int w;         // work item
  int b;         // rate class
  int64_t row0;  // first row of the super-tile
};

// DE_PRE children are prefetched (2; 3 for the launch of a trifurcating root, which is also
// compiled without the contraction: CONTRACT = false).
// RA row groups of 8 per warp: with RA = 2 (the contraction instance: 8 warps x 16 rows) every
// 128-bit B-fragment load from shared memory feeds FOUR DMMAs instead of two -- the shared-memory
// pipe was ~50 % busy with B loads at one row group per warp -- and the accumulators are produced
// in two column halves so that rows in flight + products + accumulators fit the register file.
// STG: the rows of the next tile are staged in shared memory with cp.async (every lane into slots it alone
// reads back: no barrier) instead of 64 registers per lane, which buys more warps per scheduler.
constexpr int DE_STG_WARPS = 16;
template <int RA, bool STG = false>
struct DECfg {
  static constexpr int WARPS = STG ? DE_STG_WARPS : ((RA == 2) ? 8 : DE_WARPS);
  static constexpr int THREADS = WARPS * 32;
  static constexpr int TILE_ROWS = WARPS * 8 * RA;
};

template <int MP, int DE_PRE, bool CONTRACT, int RA, bool STG = false>
__global__ void __launch_bounds__(DECfg<RA, STG>::THREADS, 1) prune_dmma_edge_kernel(const PruneArgs a) {
  static_assert(!STG || RA == 1, "staged rows: one row group per warp");
  // P in its natural layout with row stride MP + 8: the lane (g, q) then reads the B values of k
  // steps 2a and 2a+1 of column block ni -- P[8 ni + g][8a + 2q], [.. + 1] -- with ONE conflict-free
  // 128-bit load (quarter-warps: g in {0,1} -> 16-byte bank groups 0-3 and 4-7)
  constexpr int LDS = MP + 8, NI = MP / 8, KS = MP / 4;
  constexpr int THREADS = DECfg<RA, STG>::THREADS, TILE_ROWS = DECfg<RA, STG>::TILE_ROWS;
  extern __shared__ __align__(16) double smem[];  // [MP * LDS]: the node's own P; STG: + [WARPS][DE_PRE][8][LDS] rows
  __shared__ ESmem sd;
  const int K = a.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;

  pdl_trigger();
  if (tid == 0) tl_first(a.tl, 0);
  for (int i = tid; i <= a.n_work; i += THREADS) sd.prefix[i] = a.tile_prefix[i];
  for (int i = tid; i < a.n_work; i += THREADS) {
    const NodeDesc nd = a.nodes[a.work_node[i]];
    EWork& wd = sd.work[i];
    wd.out = nd.out;
    wd.n_rows = nd.n_rows;
    wd.row_tiles = (nd.n_rows + TILE_ROWS - 1) / TILE_ROWS;
    wd.P_own = nd.P_own;
    wd.child_begin = nd.child_begin;
    wd.nc = nd.n_children;
    wd.fixed_motif = nd.fixed_motif;
    wd.is_root = nd.is_root;
    for (int c = 0; c < DE_PRE_MAX; ++c) {
      const ChildRef ch = a.childs[nd.child_begin + (c < nd.n_children ? c : 0)];
      wd.src[c] = ch.src;
      wd.idx[c] = ch.idx;
    }
  }
  __syncthreads();

  const int per_cta = (a.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int st0 = (int)blockIdx.x * per_cta;
  const int st1 = min(a.total_tiles, st0 + per_cta);
  auto place = [&](ECur& cur) {
    while (cur.st >= sd.prefix[cur.w + 1]) ++cur.w;
    const EWork& wd = sd.work[cur.w];
    const int local = cur.st - sd.prefix[cur.w];
    cur.b = (int)(local / wd.row_tiles);
    cur.row0 = (int64_t)(local - cur.b * wd.row_tiles) * TILE_ROWS;
  };
  auto next = [&](ECur& cur) {
    if (++cur.st < st1) place(cur);
  };
  auto my_row = [&](const ECur& cur, int ra) -> int64_t {  // clamped: rows past the end repeat the last
    const int64_t p = cur.row0 + (warp * RA + ra) * 8 + g;
    const int64_t last = sd.work[cur.w].n_rows - 1;
    return p < last ? p : last;
  };

  ECur cc{st0, 0, 0, 0};
  if (cc.st < st1) place(cc);
  ECur fc = cc, ic = cc;
  int32_t ix[RA][DE_PRE];
#pragma unroll
  for (int ra = 0; ra < RA; ++ra)
#pragma unroll
    for (int c = 0; c < DE_PRE; ++c) ix[ra][c] = 0;
  auto load_idx = [&]() {
    if (ic.st < st1) {
      const EWork& wi = sd.work[ic.w];
#pragma unroll
      for (int ra = 0; ra < RA; ++ra) {
        const int64_t p = my_row(ic, ra);
#pragma unroll
        for (int c = 0; c < DE_PRE; ++c) ix[ra][c] = __ldg(wi.idx[c] + p);  // (a single child is read twice)
      }
    }
  };
  double2 vn[STG ? 1 : RA][STG ? 1 : DE_PRE][STG ? 1 : NI];
  // STG: this lane's slots: child c, row g, the 16-byte pieces 8 ni + 2 q (row stride LDS: conflict free)
  double* stage = smem + MP * LDS + ((warp * DE_PRE) * 8 + g) * LDS + 2 * q;
  auto fetch = [&]() {  // rows of tile fc (gather rows in ix)
    const EWork& wf = sd.work[fc.w];
#pragma unroll
    for (int ra = 0; ra < RA; ++ra)
#pragma unroll
      for (int c = 0; c < DE_PRE; ++c) {
        const double* src = wf.src[c] + ((int64_t)ix[ra][c] * K + fc.b) * MP + 2 * q;
        if (STG) {
          const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + c * 8 * LDS);
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) cp_async16_ca(dst + 64 * ni, src + 8 * ni);
        } else {
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) vn[ra][c][ni] = ld128_nc(src + 8 * ni);
        }
      }
    if (STG) cp_async_commit();
  };
  load_idx();  // static data: may overlap the previous launch
  pdl_wait();
  if (tid == 0) tl_first(a.tl, 1);
  if (fc.st < st1) {
    fetch();
    next(ic);
    load_idx();
  }

  int res_w = -1, res_b = -1;
#pragma unroll 1
  while (cc.st < st1) {
    const EWork& wd = sd.work[cc.w];
    if ((cc.w != res_w || cc.b != res_b)) {
      // another (node, rate class): every warp is done with the old P
      __syncthreads();
      if (CONTRACT && wd.P_own != nullptr) {
        const double* P = wd.P_own + (int64_t)cc.b * MP * MP;
        for (int o = tid; o < MP * MP / 2; o += THREADS) {
          const int i = (2 * o) / MP, j = 2 * o - i * MP;
          *reinterpret_cast<double2*>(smem + i * LDS + j) = *reinterpret_cast<const double2*>(P + 2 * o);
        }
      }
      res_w = cc.w;
      res_b = cc.b;
      __syncthreads();
    }
    // product of the children's rows (C-fragment layout); then the next tile's rows are requested
    double prod[RA][NI][2];
    if (STG) cp_async_wait_group<0>();  // this lane's own copies: nobody else reads these slots
#pragma unroll
    for (int ra = 0; ra < RA; ++ra) {
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const double2 v = STG ? *reinterpret_cast<const double2*>(stage + 8 * ni) : vn[STG ? 0 : ra][0][STG ? 0 : ni];
        prod[ra][ni][0] = v.x;
        prod[ra][ni][1] = v.y;
      }
#pragma unroll
      for (int c = 1; c < DE_PRE; ++c) {
        if (c < wd.nc) {
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) {
            const double2 v = STG ? *reinterpret_cast<const double2*>(stage + c * 8 * LDS + 8 * ni)
                                  : vn[STG ? 0 : ra][STG ? 0 : c][STG ? 0 : ni];
            prod[ra][ni][0] *= v.x;
            prod[ra][ni][1] *= v.y;
          }
        }
      }
    }
    next(fc);
    if (fc.st < st1) {
      fetch();
      next(ic);
      load_idx();
    }
    int64_t prow[RA];
    bool row_ok[RA];
#pragma unroll
    for (int ra = 0; ra < RA; ++ra) {
      prow[ra] = cc.row0 + (warp * RA + ra) * 8 + g;
      row_ok[ra] = prow[ra] < wd.n_rows;
    }
    for (int c = DE_PRE; c < wd.nc; ++c) {  // third, ... child (trifurcating root, polytomies): at use
      const ChildRef ch = a.childs[wd.child_begin + c];
#pragma unroll
      for (int ra = 0; ra < RA; ++ra) {
        const int32_t r = __ldg(ch.idx + (row_ok[ra] ? prow[ra] : wd.n_rows - 1));
        const double* src = ch.src + ((int64_t)r * K + cc.b) * MP + 2 * q;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          const double2 v = ld128_nc(src + 8 * ni);
          prod[ra][ni][0] *= v.x;
          prod[ra][ni][1] *= v.y;
        }
      }
    }
    if (wd.fixed_motif >= 0) {
#pragma unroll
      for (int ra = 0; ra < RA; ++ra)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          if (8 * ni + 2 * q != wd.fixed_motif) prod[ra][ni][0] = 0.0;
          if (8 * ni + 2 * q + 1 != wd.fixed_motif) prod[ra][ni][1] = 0.0;
        }
    }
    if (!CONTRACT || wd.P_own == nullptr) {
      // the root: plh itself (when it is stored at all), and lh = inner(plh_root, pi)
#pragma unroll
      for (int ra = 0; ra < RA; ++ra) {
        double* dst = wd.out + (prow[ra] * K + cc.b) * MP + 2 * q;
        if (row_ok[ra] && wd.out != nullptr) {
#pragma unroll
          for (int ni = 0; ni < NI; ++ni)
            *reinterpret_cast<double2*>(dst + 8 * ni) = make_double2(prod[ra][ni][0], prod[ra][ni][1]);
        }
        if (wd.is_root) {
          double sdot = 0.0;
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) {
            sdot = fma(prod[ra][ni][0], a.pi[8 * ni + 2 * q], sdot);
            sdot = fma(prod[ra][ni][1], a.pi[8 * ni + 2 * q + 1], sdot);
          }
          sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
          sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
          if (row_ok[ra] && q == 0) a.lh_out[prow[ra] * K + cc.b] = sdot;
        }
      }
    } else {
      const double* bRow = smem + g * LDS + 2 * q;
      constexpr int NH = (RA == 2 && NI % 2 == 0) ? 2 : 1;  // column halves (RA = 2: accumulators of one half at a time)
      constexpr int NIH = NI / NH;
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        double cur[RA][NIH][2];
#pragma unroll
        for (int ra = 0; ra < RA; ++ra)
#pragma unroll
          for (int ni = 0; ni < NIH; ++ni) cur[ra][ni][0] = cur[ra][ni][1] = 0.0;
#pragma unroll
        for (int a2 = 0; a2 < KS / 2; ++a2) {
#pragma unroll
          for (int ni = 0; ni < NIH; ++ni) {
            const double2 bv = *reinterpret_cast<const double2*>(bRow + (8 * (h * NIH + ni)) * LDS + 8 * a2);
#pragma unroll
            for (int ra = 0; ra < RA; ++ra) {
              dmma_m8n8k4(cur[ra][ni][0], cur[ra][ni][1], prod[ra][a2][0], bv.x);
              dmma_m8n8k4(cur[ra][ni][0], cur[ra][ni][1], prod[ra][a2][1], bv.y);
            }
          }
        }
#pragma unroll
        for (int ra = 0; ra < RA; ++ra) {
          if (row_ok[ra]) {
            double* dst = wd.out + (prow[ra] * K + cc.b) * MP + 2 * q;
#pragma unroll
            for (int ni = 0; ni < NIH; ++ni)
              *reinterpret_cast<double2*>(dst + 8 * (h * NIH + ni)) = make_double2(cur[ra][ni][0], cur[ra][ni][1]);
          }
        }
      }
    }
    next(cc);
  }
  if (lane == 0) tl_end(a.tl);
}

// plh of a node recomputed for the accessors: prod_c (P_c .) rows_c[idx_c[p]] (masked like the kernels).
// Used for edge-major nodes (stored rows are P . plh) and for a root whose rows are not stored.
// The contraction sums k = 0..M-1 in order with FMAs, like matvec4 and the DMMA accumulation.
__global__ void __launch_bounds__(256)
edge_major_plh_kernel(const NodeDesc nd, const ChildRef* __restrict__ childs, int K, int M, int MP, int bin,
                      double* __restrict__ out) {
  const int64_t total = nd.n_rows * MP;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = o / MP;
    const int j = (int)(o - p * MP);
    double v = 1.0;
    for (int c = 0; c < nd.n_children; ++c) {
      const ChildRef ch = childs[nd.child_begin + c];
      const double* row = ch.src + ((int64_t)ch.idx[p] * K + bin) * MP;
      double x;
      if (ch.P != nullptr) {
        const double* Pj = ch.P + ((int64_t)bin * MP + j) * MP;
        x = row[0] * Pj[0];
        for (int k = 1; k < M; ++k) x = fma(row[k], Pj[k], x);
      } else {
        x = row[j];
      }
      v = (c == 0) ? x : v * x;
    }
    if (j >= M || (nd.fixed_motif >= 0 && j != nd.fixed_motif)) v = 0.0;
    out[o] = v;
  }
}

// ---------------------------------------------------------------------------------------
// site-weighted log-likelihood reduction
//   mix[p] = sum_b bprob_b lh_b[p]     BinnedSiteDistribution.get_weighted_sum_lh
//                                      (evolve/likelihood_calculation.py:178-185)
//   lnL    = sum_p log(mix[p]) counts[p]   get_log_sum_across_sites
//                                      (evolve/likelihood_tree_numba.py:36-42)
// Deterministic: fixed tile size, fixed in-tile tree, tiles summed in index order by the
// last CTA to finish -- the result does not depend on the grid or on timing.
// ---------------------------------------------------------------------------------------

// Sharded runs (CommArgs.nranks > 1): the last CTA then turns the shard's sum into the TOTAL over
// all ranks of the node without a separate collective -- thread r stores (value, epoch) into
// rank r's mailbox over NVLink (value, fence.sys, epoch), every rank waits until its own mailbox
// holds all nranks values of this epoch and adds them in rank order (deterministic and identical
// on every rank).  Two mailbox parities: a rank can be at most one evaluation ahead of another.
// A bounded wait (30 s) turns a missing peer into NaN instead of a hung GPU.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(REDUCE_THREADS)
lnl_reduce_kernel(const double* __restrict__ lh, const double* __restrict__ bprobs,
                  const double* __restrict__ counts, int64_t n_rows, int K, int n_tiles,
                  double* __restrict__ tile_sums, unsigned int* __restrict__ counter,
                  const ResultSink sink, const CommArgs comm, const int32_t* __restrict__ root_exp,
                  const double* __restrict__ mixed) {
  __shared__ double s_warp[REDUCE_THREADS / 32];
  __shared__ bool s_last;
  if (threadIdx.x == 0) tl_first(sink.tl, 0);
  pdl_wait();  // launched early (programmatic dependent launch): lh is the root launch's output
  if (threadIdx.x == 0) tl_first(sink.tl, 1);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const double acc = reduce_tile_partial(lh, bprobs, counts, n_rows, K, tile, (int)threadIdx.x, root_exp, mixed);
    const double total = block_sum(acc, s_warp);
    if (threadIdx.x == 0) tile_sums[tile] = total;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(counter, 1u);
    s_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_tiles; i += REDUCE_THREADS) acc += __ldcg(tile_sums + i);
  double total = block_sum(acc, s_warp);
  if (comm.nranks > 1) {
    __shared__ double s_shard[C3_MAX_RANKS];
    __shared__ int s_ok[C3_MAX_RANKS];
    __shared__ double s_local;
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) {
      s_local = total;
      // this evaluation's number: one more than the last one's (only this CTA ever touches the counter)
      s_epoch = *comm.epoch_ctr + 1;
      *comm.epoch_ctr = s_epoch;
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < comm.nranks) {
      const unsigned long long epoch = s_epoch;
      const int parity = (int)(epoch & 1ull);
      // my value into rank r's mailbox
      CommSlot* dst = comm.peer[r] + parity * comm.nranks + comm.rank;
      *reinterpret_cast<volatile double*>(&dst->value) = s_local;
      st_release_sys_u64(&dst->epoch, epoch);
      // rank r's value out of my mailbox
      const CommSlot* src = comm.peer[comm.rank] + parity * comm.nranks + r;
      const unsigned long long t0 = globaltimer_ns();
      double v = 0.0;
      int ok = 0;
      while (true) {
        if (ld_acquire_sys_u64(&src->epoch) == epoch) {
          v = *reinterpret_cast<const volatile double*>(&src->value);
          ok = 1;
          break;
        }
        if (globaltimer_ns() - t0 > 30000000000ull) break;
        __nanosleep(200);
      }
      s_shard[r] = v;
      s_ok[r] = ok;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      total = 0.0;
      bool all = true;
      for (int i = 0; i < comm.nranks; ++i) {
        total += s_shard[i];
        all = all && s_ok[i] != 0;
      }
      // a peer that never showed up: a marked NaN (the host returns an error, not a value)
      if (!all) total = __longlong_as_double((long long)C3_TIMEOUT_NAN_BITS);
    }
  }
  if (threadIdx.x == 0) {
    *counter = 0u;
    publish_result(sink, total);
    tl_end(sink.tl);
  }
}

// ---- per-pattern underflow rescaling --------------------------------------------------------
// The reference multiplies plain FP64 partials and has no rescaling on this path
// (evolve/likelihood_tree.py:10,151-153): beyond a few hundred divergent taxa lh underflows to 0
// and lnL is -inf.  When rescaling is active the host picks "scaled" nodes (every subtree that has
// gathered >= threshold unscaled tips); after such a node's rows are computed, every row (all K
// bins together) is divided by 2^e, e = ilogb(largest element), and e is kept as an integer.
// Multiplying by a power of two is exact, so whenever the reference does not underflow
// lh_scaled * 2^E == lh bit for bit.  A scaled node's exponent is cumulative: it adds the
// exponents of its nearest scaled descendants, gathered through index maps composed at set-up
// (compose_map_kernel), so the root only has to gather from its own nearest scaled descendants.
struct ScaleSrc {
  const int32_t* map;   // row of the descendant for every row of this node
  const int32_t* exps;  // the descendant's cumulative exponents
};

// chain / walk kernels: every root row computes the row of a lower node it looks at, and MANY root rows look at the
// same one (a cherry has a few dozen distinct rows, the root thousands of patterns).  Only the FIRST root row of each
// node row stores it -- thousands of stores to a handful of addresses serialise in L2 (0.8 us per node on brca1) --
// the others carry the top bit in their map entry (gathers mask it).
__global__ void __launch_bounds__(256)
map_first_kernel(const int32_t* __restrict__ map, int32_t* __restrict__ first, int64_t n) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    atomicMin(first + (map[p] & 0x7fffffff), (int32_t)p);
}
__global__ void __launch_bounds__(256)
map_mark_kernel(int32_t* __restrict__ map, const int32_t* __restrict__ first, int64_t n) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = map[p] & 0x7fffffff;
    map[p] = (first[r] == (int32_t)p) ? r : (r | (int32_t)0x80000000);
  }
}

// out[p] = idx[cur[p]]   (cur == nullptr: identity) -- one step of composing gather tables down a path
__global__ void __launch_bounds__(256)
compose_map_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ cur, const int32_t* __restrict__ idx,
                   int64_t n) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    out[p] = idx[cur ? (cur[p] & 0x7fffffff) : p];
}

// rows: [n_rows][row_len] doubles, row_len a multiple of 4; LPR lanes share a row
template <int LPR>
__global__ void __launch_bounds__(256)
rescale_rows_kernel(double* __restrict__ rows, int64_t n_rows, int row_len, int32_t* __restrict__ exps,
                    const ScaleSrc* __restrict__ src, int n_src) {
  constexpr int ROWS_PER_WARP = 32 / LPR;
  const int lane = threadIdx.x & 31, g = lane % LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int chunks = row_len >> 2;
  for (int64_t base = warp * ROWS_PER_WARP; base < n_rows; base += n_warps * ROWS_PER_WARP) {
    const int64_t row = base + lane / LPR;
    const bool valid = row < n_rows;
    double* r = rows + row * row_len;
    double m = 0.0;
    if (valid)
      for (int c = g; c < chunks; c += LPR) {
        const d4 v = ld256(r + c * 4);
        m = fmax(m, fmax(fmax(v.x, v.y), fmax(v.z, v.w)));
      }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    int ex = 0;
    if (m > 0.0 && m < __longlong_as_double(0x7ff0000000000000ll)) ex = ilogb(m);
    if (valid && ex != 0)
      for (int c = g; c < chunks; c += LPR) {
        d4 v = ld256(r + c * 4);
        v.x = scalbn(v.x, -ex); v.y = scalbn(v.y, -ex); v.z = scalbn(v.z, -ex); v.w = scalbn(v.w, -ex);
        st256(r + c * 4, v);
      }
    if (valid && g == 0) {
      int tot = ex;
      for (int s = 0; s < n_src; ++s) tot += src[s].exps[src[s].map[row]];
      exps[row] = tot;
    }
  }
}

// the root is never scaled itself: its exponent is the sum over its nearest scaled descendants
__global__ void __launch_bounds__(256)
root_exp_kernel(int32_t* __restrict__ exps, int64_t n_rows, const ScaleSrc* __restrict__ src, int n_src) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_rows; p += (int64_t)gridDim.x * blockDim.x) {
    int tot = 0;
    for (int s = 0; s < n_src; ++s) tot += src[s].exps[src[s].map[p]];
    exps[p] = tot;
  }
}

// sustained DMMA rate microbenchmark (roofline denominator of the codon kernel)
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma_m8n8k4(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}

}  // namespace c3

// superseded kernel families and the measured-and-dropped experiments of DESIGN.md lessons 3-16 (kept for the
// parity tests that run the same problems through them and for reproducing the measurements; none is on the
// default path)
#include "lab_kernels.cuh"
