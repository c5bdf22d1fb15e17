// Experiment: can a UMMA shared-memory descriptor start at an arbitrary ROW offset inside a TMA-written swizzled
// box (so that several convolution taps reuse one activation window in shared memory)?
//   case K : A K-major SWIZZLE_128B, start = base + r*128 B              (fprop / dgrad operand)
//   case MN: B MN-major SWIZZLE_128B, start = base + r*128 B (K rows)    (wgrad operand)
// each with base_offset field = 0 and = (start >> 7) & 7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/exp/desc_shift tools/exp/desc_shift.cu mmhand_b200/csrc/api.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../mmhand_b200/csrc/ptx.cuh"
#include "../../mmhand_b200/csrc/tmap.h"
using namespace mmh;

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t swz, uint32_t bo) {
  return make_smem_desc(saddr, lbo, sbo, swz) | (static_cast<uint64_t>(bo & 7) << 49);
}

// mode 0: case K.  A box = 160 rows x 64 ch; B box = 64 rows x 64 ch (K-major).  D[128][64]
// mode 1: case MN. dy box = 2 x (64 ch x 64 rows); act box = 64 ch x 80 rows.   D[128][64]
__global__ void __launch_bounds__(128, 1)
exp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int mode, int r, int use_bo,
           float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;               // up to 20 KB
  uint8_t* sB = smem + 24 * 1024;   // up to 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    if (mode == 0) {
      mbar_expect_tx(&bars[0], 160 * 128 + 64 * 128);
      tma_load_2d(&tmA, &bars[0], sA, 0, 0);
      tma_load_2d(&tmB, &bars[0], sB, 0, 0);
    } else {
      mbar_expect_tx(&bars[0], 2 * 64 * 128 + 80 * 128);
      tma_load_2d(&tmA, &bars[0], sA, 0, 0);
      tma_load_2d(&tmA, &bars[0], sA + 64 * 128, 64, 0);
      tma_load_2d(&tmB, &bars[0], sB, 0, 0);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    if (mode == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k) {
        const uint32_t a0 = smem_u32(sA) + r * 128 + k * 32;
        const uint64_t ad = desc_bo(a0, 16, 1024, 2, use_bo ? (a0 >> 7) : 0);
        const uint64_t bd = desc_bo(smem_u32(sB) + k * 32, 16, 1024, 2, 0);
        umma_bf16(tmem, ad, bd, idesc, k != 0);
      }
    } else {
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = desc_bo(smem_u32(sA) + k * 2 * 1024, 64 * 128, 1024, 2, 0);
        const uint32_t b0 = smem_u32(sB) + r * 128 + k * 2 * 1024;
        const uint64_t bd = desc_bo(b0, 80 * 128, 1024, 2, use_bo ? (b0 >> 7) : 0);
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
  const int RA = 160, RB = 80;
  std::vector<__nv_bfloat16> hA(RA * 128), hB(RB * 64), hW(64 * 64);
  srand(1);
  for (auto& x : hA) x = __float2bfloat16(float(rand() % 7 - 3));
  for (auto& x : hB) x = __float2bfloat16(float(rand() % 7 - 3));
  for (auto& x : hW) x = __float2bfloat16(float(rand() % 5 - 2));
  __nv_bfloat16 *dA, *dB, *dW;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dW, hW.size() * 2);
  cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hOut(128 * 64);
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap tA, tB;
    if (mode == 0) {
      // A: [160 rows][ld 128] using channels 0..63; W: [64][64]
      if (make_tmap_2d_bf16(&tA, dA, 64, RA, 128, 64, 160, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
      if (make_tmap_2d_bf16(&tB, dW, 64, 64, 64, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    } else {
      // dy: [64 rows][128 ch] (first 64 rows of A); act: [80 rows][64 ch]
      if (make_tmap_2d_bf16(&tA, dA, 128, 64, 128, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
      if (make_tmap_2d_bf16(&tB, dB, 64, RB, 64, 64, 80, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
    for (int use_bo = 0; use_bo < 2; ++use_bo) {
      printf("mode %s base_offset=%s:", mode == 0 ? "K " : "MN", use_bo ? "(addr>>7)&7" : "0");
      for (int r = 0; r <= 16; ++r) {
        cudaMemset(dOut, 0, 128 * 64 * 4);
        exp_kernel<<<1, 128, 64 * 1024>>>(tA, tB, mode, r, use_bo, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" r=%d CUDA error %s\n", r, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(hOut.data(), dOut, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            float ref = 0.f;
            if (mode == 0) {
              for (int c = 0; c < 64; ++c) ref += __bfloat162float(hA[(m + r) * 128 + c]) * __bfloat162float(hW[n * 64 + c]);
            } else {
              for (int k = 0; k < 64; ++k) ref += __bfloat162float(hA[k * 128 + m]) * __bfloat162float(hB[(k + r) * 64 + n]);
            }
            if (ref != hOut[m * 64 + n]) ++bad;
          }
        printf(" r%d:%s", r, bad ? "BAD" : "ok");
      }
      printf("\n");
    }
  }
  return 0;
}
