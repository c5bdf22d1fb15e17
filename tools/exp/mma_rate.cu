// Experiment 3: tcgen05.mma issue rate (cycles per instruction) by operand major-ness and N, cta_group::1.
// Operands are whatever lies in shared memory (zero-initialised); 4 k-steps per "stage" walk the descriptors like
// the real kernels do. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/exp/mma_rate tools/exp/mma_rate.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../mmhand_b200/csrc/ptx.cuh"
using namespace mmh;

__global__ void __launch_bounds__(128, 1)
rate_kernel(int a_mn, int b_mn, int N, int iters, int fill, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = fill ? 0x3c003c00u + (i * 2654435761u & 0x00ff00ffu) : 0u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
    // K-major: 64-ch rows (SW128), k-step +32 B, SBO 1024.  MN-major: 64 k-rows x 64-ch boxes, LBO = box (8 KB), SBO 1024, k-step +2048
    const uint64_t a_hi = a_mn ? make_smem_desc(0, 8192, 1024, 2) : make_smem_desc(0, 16, 1024, 2);
    const uint64_t b_hi = b_mn ? make_smem_desc(0, 8192, 1024, 2) : make_smem_desc(0, 16, 1024, 2);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t st = (it & 1) * 96 * 1024;     // alternate between two "stages"
      for (int k = 0; k < 4; ++k) {
        const uint32_t ao = sa + st + (a_mn ? k * 2048 : k * 32), bo = sb + st + (b_mn ? k * 2048 : k * 32);
        umma_bf16(tmem, a_hi | ((ao >> 4) & 0x3FFF), b_hi | ((bo >> 4) & 0x3FFF), idesc, 1);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024 + 1024);
  const int iters = 2000;
  for (int grid : {1, 148})
    for (int fill = 0; fill < 2; ++fill)
      for (int N : {64, 128, 256})
        for (int a_mn = 0; a_mn < 2; ++a_mn)
          for (int b_mn = 0; b_mn < 2; ++b_mn) {
            rate_kernel<<<grid, 128, 202 * 1024 + 1024>>>(a_mn, b_mn, N, iters, fill, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("grid %3d fill %d N %3d A %s B %s : %.1f clk/mma (floor %d)\n", grid, fill, N, a_mn ? "MN" : "K ", b_mn ? "MN" : "K ",
                   double(mx) / (iters * 4), 128 * N / 256);
          }
  return 0;
}
