// Per-node site-pattern compression on the device.
//
// cogent3 compresses the alignment at EVERY tree node: a tip keeps the distinct motifs it
// shows, an internal node the distinct tuples of its children's pattern numbers, both numbered
// in first-seen column order (evolve/likelihood_tree.py:186-203 `_indexed`, :13-70
// `LikelihoodTreeEdge.__init__`, :223-278 `make_likelihood_tree_leaf`).  The reference does it
// with a Python dict per column (~1 min per 64 x 1M pass).  Here, per node:
//   1. key[col] = idx_child0[col] * rows_child1 + idx_child1[col]  (tips: the symbol code);
//      polytomies fold their children pairwise through an intermediate numbering;
//   2. insert into an open-addressing hash table (64-bit keys, linear probing) that records
//      for every distinct key the SMALLEST column holding it (atomicMin); lanes of a warp
//      that hold the same key are merged first (__match_any_sync) -- frequent patterns would
//      otherwise serialise on one address;
//   3. a column is a first occurrence iff table.first[slot(col)] == col; an exclusive prefix
//      sum over those flags IS the first-seen numbering;
//   4. index[col] = number of the key's first column; the first columns write the child
//      pattern numbers into the node's gather table `indexes[c][pattern]`.
// Everything is integer work and bit-exact by construction (tests compare every table with the
// host restatement and with the golden dumps of the live reference).
#pragma once
#include <cstdint>

namespace c3 {

constexpr unsigned long long CMP_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr int CMP_THREADS = 256;
constexpr int CMP_SCAN_ITEMS = 8;  // columns per thread in the scan kernels
constexpr int CMP_SCAN_TILE = CMP_THREADS * CMP_SCAN_ITEMS;

struct CmpTable {
  unsigned long long* keys;  // [size] CMP_EMPTY when free
  unsigned* first;           // [size] smallest column with this key
  unsigned* rank;            // [size] first-seen number of the key
  unsigned mask;             // size - 1 (size is a power of two)
};

__device__ __forceinline__ unsigned long long cmp_hash(unsigned long long x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// key of column `col`: a (int32 numbering, may be null) * rows_b + b; or the tip's symbol code
__device__ __forceinline__ unsigned long long cmp_key(const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                                                      const uint8_t* __restrict__ codes, long long rows_b,
                                                      long long col) {
  if (codes != nullptr) return (unsigned long long)codes[col];
  return (unsigned long long)a[col] * (unsigned long long)rows_b + (unsigned long long)b[col];
}

__global__ void __launch_bounds__(CMP_THREADS)
cmp_insert_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, const uint8_t* __restrict__ codes,
                  long long rows_b, long long n_cols, CmpTable t, int32_t* __restrict__ slot_of_col) {
  const int lane = threadIdx.x & 31;
  // whole warps iterate together: the tail lanes carry a key nobody else has
  const long long n_round = (n_cols + 31) / 32 * 32;
  for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < n_round;
       col += (long long)gridDim.x * blockDim.x) {
    const bool valid = col < n_cols;
    const unsigned long long key = valid ? cmp_key(a, b, codes, rows_b, col) : (CMP_EMPTY - 1 - lane);
    const unsigned group = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(group) - 1;  // lowest lane = smallest column of the group
    unsigned slot = 0;
    if (valid && lane == leader) {
      slot = (unsigned)cmp_hash(key) & t.mask;
      while (true) {
        const unsigned long long prev = atomicCAS(t.keys + slot, CMP_EMPTY, key);
        if (prev == CMP_EMPTY || prev == key) break;
        slot = (slot + 1) & t.mask;
      }
      atomicMin(t.first + slot, (unsigned)col);
    }
    slot = __shfl_sync(0xffffffffu, slot, leader);
    if (valid) slot_of_col[col] = (int32_t)slot;
  }
}

__device__ __forceinline__ unsigned cmp_block_exclusive_scan(unsigned v, unsigned* s_warp, unsigned* block_total) {
  // exclusive scan of one value per thread over the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned n = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += n;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < CMP_THREADS / 32 ? s_warp[lane] : 0u;
    unsigned wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned n = __shfl_up_sync(0xffffffffu, wi, off);
      if (lane >= off) wi += n;
    }
    if (lane < CMP_THREADS / 32) s_warp[lane] = wi - w;  // exclusive warp offsets
    if (lane == CMP_THREADS / 32 - 1) *block_total = wi;
  }
  __syncthreads();
  return incl - v + s_warp[warp];
}

// pass 1: number of first occurrences per tile of CMP_SCAN_TILE columns
__global__ void __launch_bounds__(CMP_THREADS)
cmp_count_kernel(long long n_cols, CmpTable t, const int32_t* __restrict__ slot_of_col, unsigned* __restrict__ tile_sums) {
  __shared__ unsigned s_warp[CMP_THREADS / 32];
  __shared__ unsigned s_total;
  const long long base = (long long)blockIdx.x * CMP_SCAN_TILE + (long long)threadIdx.x * CMP_SCAN_ITEMS;
  unsigned n = 0;
#pragma unroll
  for (int i = 0; i < CMP_SCAN_ITEMS; ++i) {
    const long long col = base + i;
    if (col < n_cols) n += (t.first[slot_of_col[col]] == (unsigned)col) ? 1u : 0u;
  }
  cmp_block_exclusive_scan(n, s_warp, &s_total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_total;
}

// pass 2: exclusive scan of the tile sums by one block; total[0] = number of distinct keys
__global__ void __launch_bounds__(CMP_THREADS)
cmp_scan_tiles_kernel(unsigned* __restrict__ tile_sums, int n_tiles, unsigned* __restrict__ total) {
  __shared__ unsigned s_warp[CMP_THREADS / 32];
  __shared__ unsigned s_total;
  unsigned carry = 0;
  for (int base = 0; base < n_tiles; base += CMP_THREADS) {
    const int i = base + threadIdx.x;
    const unsigned v = i < n_tiles ? tile_sums[i] : 0u;
    const unsigned ex = cmp_block_exclusive_scan(v, s_warp, &s_total);
    if (i < n_tiles) tile_sums[i] = ex + carry;
    carry += s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = carry;
}

// pass 3: the first-seen number of every first-occurrence column, stored with its key
// (table.rank) and, for later passes, per column (pos; ~0u for other columns)
__global__ void __launch_bounds__(CMP_THREADS)
cmp_number_kernel(long long n_cols, CmpTable t, const int32_t* __restrict__ slot_of_col,
                  const unsigned* __restrict__ tile_sums, unsigned* __restrict__ pos) {
  __shared__ unsigned s_warp[CMP_THREADS / 32];
  __shared__ unsigned s_total;
  const long long base = (long long)blockIdx.x * CMP_SCAN_TILE + (long long)threadIdx.x * CMP_SCAN_ITEMS;
  bool is_first[CMP_SCAN_ITEMS];
  int32_t slot[CMP_SCAN_ITEMS];
  unsigned n = 0;
#pragma unroll
  for (int i = 0; i < CMP_SCAN_ITEMS; ++i) {
    const long long col = base + i;
    is_first[i] = false;
    slot[i] = 0;
    if (col < n_cols) {
      slot[i] = slot_of_col[col];
      is_first[i] = t.first[slot[i]] == (unsigned)col;
      n += is_first[i] ? 1u : 0u;
    }
  }
  unsigned at = cmp_block_exclusive_scan(n, s_warp, &s_total) + tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < CMP_SCAN_ITEMS; ++i) {
    const long long col = base + i;
    if (col < n_cols) {
      if (is_first[i]) {
        t.rank[slot[i]] = at;
        pos[col] = at;
        ++at;
      } else {
        pos[col] = 0xFFFFFFFFu;
      }
    }
  }
}

// pass 4: index[col] = first-seen number of the column's key
__global__ void __launch_bounds__(CMP_THREADS)
cmp_index_kernel(long long n_cols, CmpTable t, const int32_t* __restrict__ slot_of_col, int32_t* __restrict__ index) {
  for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < n_cols;
       col += (long long)gridDim.x * blockDim.x)
    index[col] = (int32_t)t.rank[slot_of_col[col]];
}

// the node's gather table: indexes[c][pattern] = child c's pattern number at the pattern's
// first column; row n_patterns is the all-gap row (child's own gap row)
__global__ void __launch_bounds__(CMP_THREADS)
cmp_gather_first_kernel(long long n_cols, const unsigned* __restrict__ pos, const int32_t* __restrict__ child_index,
                        int32_t* __restrict__ out) {
  for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < n_cols;
       col += (long long)gridDim.x * blockDim.x) {
    const unsigned p = pos[col];
    if (p != 0xFFFFFFFFu) out[p] = child_index[col];
  }
}
__global__ void __launch_bounds__(CMP_THREADS)
cmp_gather_first_u8_kernel(long long n_cols, const unsigned* __restrict__ pos, const uint8_t* __restrict__ codes,
                           int32_t* __restrict__ out) {
  for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < n_cols;
       col += (long long)gridDim.x * blockDim.x) {
    const unsigned p = pos[col];
    if (p != 0xFFFFFFFFu) out[p] = (int32_t)codes[col];
  }
}

// root: counts[pattern] = number of columns showing it (exact in double up to 2^53)
__global__ void __launch_bounds__(CMP_THREADS)
cmp_histogram_kernel(long long n_cols, const int32_t* __restrict__ index, unsigned* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const long long n_round = (n_cols + 31) / 32 * 32;
  for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < n_round;
       col += (long long)gridDim.x * blockDim.x) {
    const bool valid = col < n_cols;
    const int key = valid ? index[col] : -1 - lane;
    const unsigned group = __match_any_sync(0xffffffffu, key);
    if (valid && lane == __ffs(group) - 1) atomicAdd(counts + key, (unsigned)__popc(group));
  }
}
__global__ void __launch_bounds__(CMP_THREADS)
cmp_counts_to_double_kernel(long long n, const unsigned* __restrict__ counts, double* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (double)counts[i];
}

// first-seen numbering implies idx_c[p] <= p for every child (the wave kernel's precondition): *bad is raised
// when a table does not have the property
__global__ void __launch_bounds__(CMP_THREADS)
cmp_check_monotone_kernel(const int32_t* __restrict__ idx, int n_children, long long rows, unsigned* __restrict__ bad) {
  const long long total = (long long)n_children * rows;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const long long p = o % rows;
    if (idx[o] > p) atomicOr(bad, 1u);
  }
}

// per-column likelihoods lh[index[col]] of one bin (LikelihoodTreeEdge.get_full_length_likelihoods,
// evolve/likelihood_tree.py:93-94)
__global__ void __launch_bounds__(CMP_THREADS)
site_lh_kernel(long long n_cols, const int32_t* __restrict__ index, const double* __restrict__ lh, int K, int bin,
               double* __restrict__ out, const int32_t* __restrict__ root_exp) {
  for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < n_cols;
       col += (long long)gridDim.x * blockDim.x) {
    const int32_t p = index[col];
    const double v = lh[(long long)p * K + bin];
    out[col] = root_exp ? scalbn(v, root_exp[p]) : v;  // rescaled rows: fold the exponent back
  }
}

}  // namespace c3
