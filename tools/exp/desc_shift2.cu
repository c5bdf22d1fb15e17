// Experiment 2: row-offset descriptor starts for the narrow swizzles.
//   mode 0/1 : A K-major SWIZZLE_64B (32 ch rows) / SWIZZLE_32B (16 ch rows), start = base + r*rowbytes, K = 32 / 16
//   mode 2   : wgrad "taps on M": A operand MN-major SWIZZLE_32B over a [K rows][16 ch] window with
//              LBO = 32 B (next kw = next row) so that M index (kw, c) -> a[k + kw][c], M = 128 (8 kw x 16 c);
//              B operand = dy [K=64 rows][64 ch] MN-major SW128.  D[(kw,c)][n] = sum_k a[k+kw+r][c] * dy[k][n]
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/exp/desc_shift2 tools/exp/desc_shift2.cu mmhand_b200/csrc/api.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../mmhand_b200/csrc/ptx.cuh"
#include "../../mmhand_b200/csrc/tmap.h"
using namespace mmh;

__global__ void __launch_bounds__(128, 1)
exp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int mode, int r, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 24 * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    if (mode == 0 || mode == 1) {
      const int rb = mode == 0 ? 64 : 32;   // row bytes
      mbar_expect_tx(&bars[0], 160 * rb + 64 * rb);
      tma_load_2d(&tmA, &bars[0], sA, 0, 0);
      tma_load_2d(&tmB, &bars[0], sB, 0, 0);
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t swz = mode == 0 ? 4u : 6u;
      for (int k = 0; k < rb / 32; ++k) {
        const uint64_t ad = make_smem_desc(smem_u32(sA) + r * rb + k * 32, 16, 8 * rb, swz);
        const uint64_t bd = make_smem_desc(smem_u32(sB) + k * 32, 16, 8 * rb, swz);
        umma_bf16(tmem, ad, bd, idesc, k != 0);
      }
    } else {
      mbar_expect_tx(&bars[0], 96 * 32 + 64 * 128);
      tma_load_2d(&tmA, &bars[0], sA, 0, 0);    // window: 96 rows x 16 ch, SW32
      tma_load_2d(&tmB, &bars[0], sB, 0, 0);    // dy: 64 rows x 64 ch, SW128
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = make_smem_desc(smem_u32(sA) + r * 32 + k * 2 * 256, 32, 256, 6);
        const uint64_t bd = make_smem_desc(smem_u32(sB) + k * 2 * 1024, 64 * 128, 1024, 2);
        umma_bf16(tmem, ad, bd, idesc, k != 0);
      }
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int j = 0; j < 4; ++j) {
    uint32_t v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + j * 16, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[row * 64 + j * 16 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  std::vector<__nv_bfloat16> hA(160 * 32), hW(64 * 32), hDy(64 * 64);
  srand(2);
  for (auto& x : hA) x = __float2bfloat16(float(rand() % 7 - 3));
  for (auto& x : hW) x = __float2bfloat16(float(rand() % 5 - 2));
  for (auto& x : hDy) x = __float2bfloat16(float(rand() % 5 - 2));
  __nv_bfloat16 *dA, *dW, *dDy;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dW, hW.size() * 2); cudaMalloc(&dDy, hDy.size() * 2);
  cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dDy, hDy.data(), hDy.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hOut(128 * 64);
  for (int mode = 0; mode < 3; ++mode) {
    CUtensorMap tA, tB;
    const int ch = mode == 0 ? 32 : 16;
    if (mode < 2) {
      // A: [160 rows][ld 32], channels [0, ch); W: [64][ld 32]
      if (make_tmap_2d_bf16(&tA, dA, ch, 160, 32, ch, 160, mode == 0 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
      if (make_tmap_2d_bf16(&tB, dW, ch, 64, 32, ch, 64, mode == 0 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
    } else {
      // window: A viewed as [320 rows][16 ch] contiguous (ld = 16); box 16 x 96
      if (make_tmap_2d_bf16(&tA, dA, 16, 320, 16, 16, 96, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
      if (make_tmap_2d_bf16(&tB, dDy, 64, 64, 64, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
    printf("mode %d:", mode);
    for (int r = 0; r <= 16; ++r) {
      cudaMemset(dOut, 0, 128 * 64 * 4);
      exp_kernel<<<1, 128, 64 * 1024>>>(tA, tB, mode, r, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" r=%d CUDA error %s\n", r, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(hOut.data(), dOut, 128 * 64 * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0.f;
          if (mode < 2) {
            for (int c = 0; c < ch; ++c) ref += __bfloat162float(hA[(m + r) * 32 + c]) * __bfloat162float(hW[n * 32 + c]);
          } else {
            const int kw = m / 16, c = m % 16;
            for (int k = 0; k < 64; ++k) ref += __bfloat162float(hA[(k + kw + r) * 16 + c]) * __bfloat162float(hDy[k * 64 + n]);
          }
          if (ref != hOut[m * 64 + n]) ++bad;
        }
      printf(" r%d:%s", r, bad ? "BAD" : "ok");
    }
    printf("\n");
  }
  return 0;
}
