// Helpers shared by the device kernels and their host emulation (tests): bf16 <-> fp32 by bit
// manipulation (identical results on both sides), host/device qualifiers.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef MMH_HOST_EMU
#define MMH_HD inline
#else
#include <cuda_runtime.h>
#define MMH_HD __host__ __device__ __forceinline__
#endif

namespace mmh {

MMH_HD float bf2f(uint16_t v) {
  uint32_t u = static_cast<uint32_t>(v) << 16;
  float f;
#if defined(__CUDA_ARCH__)
  f = __uint_as_float(u);
#else
  memcpy(&f, &u, 4);
#endif
  return f;
}

// round-to-nearest-even, NaN kept quiet
MMH_HD uint16_t f2bf(float f) {
  uint32_t u;
#if defined(__CUDA_ARCH__)
  u = __float_as_uint(f);
#else
  memcpy(&u, &f, 4);
#endif
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return static_cast<uint16_t>((u >> 16) | 0x40u);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

// Storage type of activations: bf16 bit patterns in the product. -DMMH_EMU_F32 (host emulation only) widens it
// to fp32 so that CPU tests can check the kernels' index arithmetic and calculus to fp32 accuracy.
#ifdef MMH_EMU_F32
typedef float act_t;
MMH_HD float act2f(act_t v) { return v; }
MMH_HD act_t f2act(float f) { return f; }
#else
typedef uint16_t act_t;
MMH_HD float act2f(act_t v) { return bf2f(v); }
MMH_HD act_t f2act(float f) { return f2bf(f); }
#endif

}  // namespace mmh
