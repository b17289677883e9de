// Developer microbenchmark: can a TMA / bulk-async ring feed the 4-state pruning
// pattern -- out[p] = A[ia[p]] * B[ib[p]] over rows of 128 B (K = 4 rate classes x 4 states of f64) -- faster
// than the register pipeline does?  Three ways to move the gathered rows:
//   ldg     : 256-bit LDG straight into registers (what prune4_reg2_kernel does), several rows in flight per lane
//   bulk    : one cp.async.bulk (128 B, mbarrier complete_tx) per gathered row into a warp-private
//             shared-memory ring, rows read back with LDS.128
//   gather4 : cp.async.bulk.tensor.2d ... tile::gather4 over a [rows, 16] f64 tensor map: FOUR rows per request
// Bytes in flight then cost shared memory instead of registers, and skip the LSU (the cp.async ring of round 1
// died on LSU/L1 pressure).  The question is the request rate of the TMA unit for 128-byte rows.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_tma_gather_bench tools/tma_gather_bench.cu
//   tools/_tma_gather_bench [rows]       (SASS check: cuobjdump -sass ... | grep -c 'UBLKCP\|UTMALDG')
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);             \
      exit(2);                                                                                    \
    }                                                                                             \
  } while (0)

struct __align__(32) d4 {
  double x, y, z, w;
};
__device__ __forceinline__ d4 ld256(const double* p) {
  d4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st256(double* p, d4 r) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.x), "d"(r.y), "d"(r.z), "d"(r.w) : "memory");
}
__device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug must not hang the GPU box
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity) {
  for (int spin = 0; spin < (1 << 22); ++spin) {
    unsigned ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void bulk_row(unsigned dst, const void* src, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(dst),
               "l"(src), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void gather4_rows(unsigned dst, const CUtensorMap* map, int r0, int r1, int r2, int r3,
                                             unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
      "%4, %5, %6}], [%7];" ::"r"(dst),
      "l"(map), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ d4 lds256(unsigned addr, int flip) {
  // two 128-bit reads; odd row slots take the upper half first so that a quarter warp covers all 32 banks
  double2 lo, hi;
  const unsigned a0 = addr + (flip ? 16 : 0), a1 = addr + (flip ? 0 : 16);
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(lo.x), "=d"(lo.y) : "r"(a0));
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(hi.x), "=d"(hi.y) : "r"(a1));
  d4 r;
  if (flip) {
    r.x = hi.x; r.y = hi.y; r.z = lo.x; r.w = lo.y;
  } else {
    r.x = lo.x; r.y = lo.y; r.z = hi.x; r.w = hi.y;
  }
  return r;
}

// ---- ldg: U rows per lane in flight, strips of 8 U rows per warp ------------------------------------------
template <int U>
__global__ void __launch_bounds__(256) k_ldg(const double* __restrict__ A, const double* __restrict__ B,
                                             const int* __restrict__ ia, const int* __restrict__ ib,
                                             double* __restrict__ out, int64_t rows) {
  const int lane = threadIdx.x & 31, q = lane & 3, slot = lane >> 2;
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), GW = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t strips = (rows + 8 * U - 1) / (8 * U);
  for (int64_t s = gw; s < strips; s += GW) {
    d4 x[U], y[U];
    int64_t p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      p[u] = imin64(s * 8 * U + u * 8 + slot, rows - 1);
      const int ra = __ldg(ia + p[u]), rb = __ldg(ib + p[u]);
      x[u] = ld256(A + ((int64_t)ra * 4 + q) * 4);
      y[u] = ld256(B + ((int64_t)rb * 4 + q) * 4);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      x[u].x *= y[u].x; x[u].y *= y[u].y; x[u].z *= y[u].z; x[u].w *= y[u].w;
      st256(out + (p[u] * 4 + q) * 4, x[u]);
    }
  }
}

// ---- bulk / gather4: warp-private ring of S stages, a stage = 32 rows x 2 children x 128 B = 8 KB ---------
template <int WARPS, int S, bool G4>
__global__ void __launch_bounds__(WARPS * 32) k_ring(const double* __restrict__ A, const double* __restrict__ B,
                                                     const int* __restrict__ ia, const int* __restrict__ ib,
                                                     double* __restrict__ out, int64_t rows,
                                                     const CUtensorMap* __restrict__ mapA,
                                                     const CUtensorMap* __restrict__ mapB, int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = lane & 3, slot = lane >> 2;
  unsigned char* ring = smem + (size_t)warp * S * 8192;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * S * 8192) + warp * S;
  if (lane == 0)
    for (int i = 0; i < S; ++i) mbar_init(smem_u32(bars + i), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, GW = (int64_t)gridDim.x * WARPS;
  const int64_t strips = (rows + 31) / 32;

  auto issue = [&](int64_t s, int stage) {
    const unsigned bar = smem_u32(bars + stage);
    const unsigned dst = smem_u32(ring + stage * 8192);
    if (lane == 0) mbar_expect_tx(bar, 8192);
    __syncwarp();
    if (!G4) {
      const int64_t p = imin64(s * 32 + lane, rows - 1);
      const int ra = __ldg(ia + p), rb = __ldg(ib + p);
      bulk_row(dst + lane * 128, A + (int64_t)ra * 16, bar);
      bulk_row(dst + 4096 + lane * 128, B + (int64_t)rb * 16, bar);
    } else if (lane < 16) {
      // lanes 0-7: four rows of child A each, lanes 8-15: child B
      const int j = lane & 7;
      const int* ix = (lane < 8) ? ia : ib;
      int r[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) r[t] = __ldg(ix + imin64(s * 32 + j * 4 + t, rows - 1));
      gather4_rows(dst + (lane < 8 ? 0 : 4096) + j * 512, (lane < 8) ? mapA : mapB, r[0], r[1], r[2], r[3], bar);
    }
  };

  int n_issued = 0;
  for (int64_t s = gw; s < strips && n_issued < S; s += GW, ++n_issued) issue(s, n_issued);
  int stage = 0;
  unsigned parity = 0;
  for (int64_t s = gw; s < strips; s += GW) {
    if (!mbar_wait(smem_u32(bars + stage), parity)) {
      if (lane == 0) atomicExch(err, 1);
      return;
    }
    const unsigned base = smem_u32(ring + stage * 8192);
    d4 x[4], y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = u * 8 + slot;
      x[u] = lds256(base + r * 128 + q * 32, slot & 1);
      y[u] = lds256(base + 4096 + r * 128 + q * 32, slot & 1);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x[u].x *= y[u].x; x[u].y *= y[u].y; x[u].z *= y[u].z; x[u].w *= y[u].w;
      const int64_t p = s * 32 + u * 8 + slot;
      if (p < rows) st256(out + (p * 4 + q) * 4, x[u]);
    }
    __syncwarp();  // every lane has its rows in registers: the stage may be overwritten
    const int64_t nxt = s + (int64_t)S * GW;
    if (nxt < strips) issue(nxt, stage);
    if (++stage == S) {
      stage = 0;
      parity ^= 1;
    }
  }
}

template <typename F>
float timeit(F f, int reps = 10) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = std::min(best, ms);
  }
  return best;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(EncodeTiled enc, CUtensorMap* m, void* base, int64_t rows) {
  const cuuint64_t dims[2] = {16, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {16, 1};  // gather4: one row per box, four boxes per request
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

int main(int argc, char** argv) {
  const int64_t rows = argc > 1 ? atoll(argv[1]) : 2000000;
  const int64_t n = rows * 16;
  double *A, *B, *o, *o_ref;
  int *ia, *ib, *err;
  CK(cudaMalloc(&A, n * 8));
  CK(cudaMalloc(&B, n * 8));
  CK(cudaMalloc(&o, n * 8));
  CK(cudaMalloc(&o_ref, n * 8));
  CK(cudaMalloc(&ia, rows * 4));
  CK(cudaMalloc(&ib, rows * 4));
  CK(cudaMalloc(&err, 4));
  CK(cudaMemset(err, 0, 4));
  {
    std::vector<double> h(n);
    std::mt19937_64 g(1);
    for (auto& v : h) v = 0.5 + (double)(g() >> 40) / (1 << 24);
    CK(cudaMemcpy(A, h.data(), n * 8, cudaMemcpyHostToDevice));
    for (auto& v : h) v = 0.5 + (double)(g() >> 40) / (1 << 24);
    CK(cudaMemcpy(B, h.data(), n * 8, cudaMemcpyHostToDevice));
  }
  EncodeTiled enc = nullptr;
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      enc = (EncodeTiled)fn;
  }
  CUtensorMap hA, hB, *dA = nullptr, *dB = nullptr;
  bool have_maps = enc && make_map(enc, &hA, A, rows) && make_map(enc, &hB, B, rows);
  if (have_maps) {
    CK(cudaMalloc(&dA, sizeof(CUtensorMap)));
    CK(cudaMalloc(&dB, sizeof(CUtensorMap)));
    CK(cudaMemcpy(dA, &hA, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, &hB, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  }
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  const double mb = n * 8 / 1e6;
  const double moved = 3 * mb + 2 * rows * 4 / 1e6;
  printf("rows=%lld (%.0f MB per array), %d SMs; GB/s = (2 reads + 1 write + 2 index arrays) / time\n", (long long)rows, mb, n_sm);

  for (int pattern = 0; pattern < 2; ++pattern) {
    // 0: identity; 1: "evolved": monotone with repeats, child tables half / two thirds the size (idx[p] <= p)
    std::vector<int> h(rows), h2(rows);
    std::mt19937_64 g(7);
    if (pattern == 0) {
      for (int64_t i = 0; i < rows; ++i) h[i] = h2[i] = (int)i;
    } else {
      int a = 0, b = 0;
      for (int64_t i = 0; i < rows; ++i) {
        h[i] = a;
        h2[i] = b;
        if (g() % 2 == 0) ++a;
        if (g() % 3 != 0) ++b;
      }
    }
    CK(cudaMemcpy(ia, h.data(), rows * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ib, h2.data(), rows * 4, cudaMemcpyHostToDevice));
    printf("--- index pattern: %s\n", pattern == 0 ? "identity" : "monotone with repeats (child tables 1/2 and 2/3 of the rows)");
    CK(cudaMemset(o_ref, 0, n * 8));
    {
      float t = timeit([&] { k_ldg<4><<<n_sm, 256>>>(A, B, ia, ib, o_ref, rows); });
      printf("ldg  U=4  8 warps x 1 CTA/SM (no software pipeline)      : %7.0f GB/s\n", moved / t);
      for (int mult : {2, 4, 8}) {
        t = timeit([&] { k_ldg<2><<<n_sm * mult, 256>>>(A, B, ia, ib, o, rows); });
        printf("ldg  U=2  8 warps x %d CTA/SM                             : %7.0f GB/s\n", mult, moved / t);
      }
    }
    auto check = [&](const char* name) {
      int herr = 0;
      CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
      std::vector<double> x(1 << 16), y(1 << 16);
      int64_t bad = 0;
      for (int64_t off : {(int64_t)0, n / 2 - (1 << 15), n - (1 << 16)}) {
        CK(cudaMemcpy(x.data(), o + off, x.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(y.data(), o_ref + off, y.size() * 8, cudaMemcpyDeviceToHost));
        bad += memcmp(x.data(), y.data(), x.size() * 8) != 0;
      }
      if (herr || bad) printf("   !! %s: %s%s\n", name, herr ? "mbarrier wait timed out " : "", bad ? "OUTPUT MISMATCH" : "");
      CK(cudaMemset(err, 0, 4));
      return !herr && !bad;
    };
#define RING(W, S, G4, CTAS)                                                                              \
  {                                                                                                       \
    const size_t sm = (size_t)W * S * 8192 + W * S * 8;                                                   \
    CK(cudaFuncSetAttribute(k_ring<W, S, G4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
    CK(cudaMemset(o, 0, n * 8));                                                                          \
    float t = timeit([&] { k_ring<W, S, G4><<<n_sm * CTAS, W * 32, sm>>>(A, B, ia, ib, o, rows, dA, dB, err); }); \
    bool ok = check(G4 ? "gather4" : "bulk");                                                             \
    printf("%-7s %2d warps x %d CTA/SM, %d stages (%3zu KB smem/CTA)        : %7.0f GB/s %s\n",            \
           G4 ? "gather4" : "bulk", W, CTAS, S, sm >> 10, moved / t, ok ? "" : "(INVALID)");              \
  }
    RING(8, 2, false, 1)
    RING(8, 3, false, 1)
    RING(4, 3, false, 2)
    RING(4, 6, false, 1)
    RING(12, 2, false, 1)
    RING(16, 1, false, 1)
    RING(4, 2, false, 3)
    if (have_maps) {
      RING(8, 2, true, 1)
      RING(8, 3, true, 1)
      RING(4, 3, true, 2)
      RING(4, 6, true, 1)
      RING(12, 2, true, 1)
      RING(16, 1, true, 1)
      RING(4, 2, true, 3)
    } else {
      printf("gather4: no tensor maps (cuTensorMapEncodeTiled unavailable)\n");
    }
  }
  return 0;
}
