// Experiment 4: what does a 35 MB -> 35 MB elementwise pass need to run at HBM speed? (16-byte loads, U loads in
// flight per thread, blocks per SM capped by launch bounds / dynamic smem, grid-stride vs one-shot grid, cold L2)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int U>
__global__ void __launch_bounds__(256) copy_k(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i < n; i += stride * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 256 < n) v[u] = src[i + u * 256];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 256 < n) { v[u].x ^= 1; dst[i + u * 256] = v[u]; }
  }
}

template <int U>
float run(const uint4* s, uint4* d, size_t n, int blocks, int smem, uint4* flush, size_t nflush) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaMemsetAsync(flush, r, nflush * 16);
    cudaEventRecord(e0);
    copy_k<U><<<blocks, 256, smem>>>(s, d, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  const size_t n = 69696ull * 256 * 2 / 16;      // 35.7 MB of bf16
  const size_t nflush = 512ull << 20 >> 4;
  uint4 *s, *d, *f;
  cudaMalloc(&s, n * 16); cudaMalloc(&d, n * 16); cudaMalloc(&f, nflush * 16);
  cudaMemset(s, 1, n * 16);
  cudaFuncSetAttribute(copy_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(copy_k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(copy_k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(copy_k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const double bytes = 2.0 * n * 16;
  for (int bps : {1, 2, 4, 8}) {          // resident blocks per SM (limited through dynamic smem)
    const int smem = bps == 8 ? 0 : (bps == 4 ? 50 * 1024 : (bps == 2 ? 100 * 1024 : 100 * 1024));
    if (bps == 1) continue;
    for (int mode = 0; mode < 2; ++mode) {   // 0: persistent grid = 148*bps, 1: one-shot grid
      float t1, t2, t4, t8;
      auto blocks = [&](int U) { return mode == 0 ? 148 * bps : (int)((n + 256 * U - 1) / (256 * U)); };
      t1 = run<1>(s, d, n, blocks(1), smem, f, nflush);
      t2 = run<2>(s, d, n, blocks(2), smem, f, nflush);
      t4 = run<4>(s, d, n, blocks(4), smem, f, nflush);
      t8 = run<8>(s, d, n, blocks(8), smem, f, nflush);
      printf("blocks/SM %d %s: U=1 %.1f us (%.2f TB/s)  U=2 %.1f (%.2f)  U=4 %.1f (%.2f)  U=8 %.1f (%.2f)\n", bps,
             mode ? "one-shot  " : "persistent", t1 * 1e3, bytes / t1 / 1e9, t2 * 1e3, bytes / t2 / 1e9, t4 * 1e3,
             bytes / t4 / 1e9, t8 * 1e3, bytes / t8 / 1e9);
    }
  }
  return 0;
}
