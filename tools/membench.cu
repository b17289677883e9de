// Developer microbenchmark: what HBM throughput do simple access patterns reach on this GPU?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/membench tools/membench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

struct __align__(32) d4 { double x, y, z, w; };
__device__ __forceinline__ d4 ld256(const double* p) {
  d4 r; asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void st256(double* p, d4 r) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(r.x), "d"(r.y), "d"(r.z), "d"(r.w) : "memory"); }

// n = number of 32-byte units
__global__ void k_copy(const double* a, double* o, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st256(o + 4 * i, ld256(a + 4 * i));
}
__global__ void k_mul2(const double* a, const double* b, double* o, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    d4 x = ld256(a + 4 * i), y = ld256(b + 4 * i);
    x.x *= y.x; x.y *= y.y; x.z *= y.z; x.w *= y.w;
    st256(o + 4 * i, x);
  }
}
// rows of 128 B (4 units); idx per row; unit u -> row u/4
__global__ void k_gather2(const double* a, const double* b, const int* ia, const int* ib, double* o, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i >> 2; const int q = i & 3;
    const int ra = __ldg(ia + row), rb = __ldg(ib + row);
    d4 x = ld256(a + ((int64_t)ra * 4 + q) * 4), y = ld256(b + ((int64_t)rb * 4 + q) * 4);
    x.x *= y.x; x.y *= y.y; x.z *= y.z; x.w *= y.w;
    st256(o + 4 * i, x);
  }
}
__global__ void k_mul2_f2(const double2* a, const double2* b, double2* o, int64_t n2) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    double2 x = a[i], y = b[i]; x.x *= y.x; x.y *= y.y; o[i] = x;
  }
}
__global__ void k_read2(const double* a, const double* b, double* o, int64_t n) {
  double s = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    d4 x = ld256(a + 4 * i), y = ld256(b + 4 * i);
    s += x.x * y.x + x.w * y.w;
  }
  if (s == 1.2345) o[0] = s;
}
__global__ void k_write(double* o, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    d4 x{1.0, 2.0, 3.0, (double)i}; st256(o + 4 * i, x);
  }
}

template <typename F> float timeit(F f, int reps = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms); }
  return best;
}

int main(int argc, char** argv) {
  int64_t rows = argc > 1 ? atoll(argv[1]) : 2000000;  // rows of 128 B
  int64_t n = rows * 4;
  double *a, *b, *o; int *ia, *ib;
  cudaMalloc(&a, n * 32); cudaMalloc(&b, n * 32); cudaMalloc(&o, n * 32);
  cudaMalloc(&ia, rows * 4); cudaMalloc(&ib, rows * 4);
  cudaMemset(a, 0, n * 32); cudaMemset(b, 0, n * 32);
  std::vector<int> h(rows); for (int64_t i = 0; i < rows; ++i) h[i] = (int)i;
  cudaMemcpy(ia, h.data(), rows * 4, cudaMemcpyHostToDevice); cudaMemcpy(ib, h.data(), rows * 4, cudaMemcpyHostToDevice);
  double mb = n * 32 / 1e6;
  printf("rows=%lld (%.0f MB per array)\n", (long long)rows, mb);
  for (int threads : {256, 512, 1024}) for (int mult : {2, 4, 8, 16, 32}) {
    int grid = 148 * mult;
    float t1 = timeit([&] { k_copy<<<grid, threads>>>(a, o, n); });
    float t2 = timeit([&] { k_mul2<<<grid, threads>>>(a, b, o, n); });
    float t3 = timeit([&] { k_gather2<<<grid, threads>>>(a, b, ia, ib, o, n); });
    float t4 = timeit([&] { k_read2<<<grid, threads>>>(a, b, o, n); });
    float t5 = timeit([&] { k_write<<<grid, threads>>>(o, n); });
    float t6 = timeit([&] { k_mul2_f2<<<grid, threads>>>((double2*)a, (double2*)b, (double2*)o, n * 2); });
    printf("thr %4d grid %5d | copy %6.0f GB/s | mul2 %6.0f | gather2(identity) %6.0f | read2 %6.0f | write %6.0f | mul2 128-bit %6.0f\n", threads, grid,
           2 * mb / t1, 3 * mb / t2, (3 * mb + 2 * rows * 4 / 1e6) / t3, 2 * mb / t4, mb / t5, 3 * mb / t6);
  }
  // one-shot grids (one unit per thread)
  {
    int threads = 256; int grid = (int)((n + threads - 1) / threads);
    float t2 = timeit([&] { k_mul2<<<grid, threads>>>(a, b, o, n); });
    float t3 = timeit([&] { k_gather2<<<grid, threads>>>(a, b, ia, ib, o, n); });
    printf("one unit per thread: mul2 %6.0f GB/s | gather2 %6.0f\n", 3 * mb / t2, (3 * mb + 2 * rows * 4 / 1e6) / t3);
  }
  return 0;
}
