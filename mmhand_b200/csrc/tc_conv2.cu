// Shifted-row implicit-GEMM convolution on tcgen05, generation 2: activation *windows* in shared memory.
//
//   out[map(q)][n] = act(bias[n] + sum_t sum_c a[q + shift[t]][c] * w[t][n][c])
//
// The taps of a convolution read the same activation rows at different row offsets. Taps whose shifts lie
// close together (the 9 taps of a 3x3 on a 64-wide grid, the 7 column taps of one kernel row of a 7x7) form
// a *group*: per (tile, group, 64-channel chunk) ONE window of 128 + span rows is brought in by TMA and every
// tap of the group multiplies straight out of it -- its UMMA descriptor simply starts `rel` rows into the
// window (tcgen05 descriptors are plain address arithmetic on the swizzled layout; the XOR pattern is a
// function of the absolute shared-memory address, so any 16-byte-aligned row start is legal; verified by
// tools/exp/desc_shift*.cu). The weight tiles stream through a second ring. L2 -> SM traffic per
// 128 x 256 x 64 x 9-tap step drops from 432 KB to 321 KB; with NCTA = 2 (cta_group::2 pairs, each CTA
// loading its own 128-row window and half of every weight tile) to 177 KB, which lifts the kernel off the
// L2 bandwidth limit the first generation sat on (profiles/r01_conv_v1_ncu.txt).
//
// Warp roles per CTA (192 threads): warp 0 TMA producer (lane 0), warp 1 MMA issuer (lane 0 of the pair's
// leader CTA) + TMEM allocation, warps 2..5 epilogue (TMEM -> registers -> bias/activation -> global).
// Two accumulator stages of 256 TMEM columns: the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/mmhand_sm100.h"
#include "conv_plan.h"
#include "host_common.h"
#include "pair.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace mmh {

// Epilogue warp groups. 1: warps 2..5 drain every tile (192 threads). 2: two groups of four warps (warps 4..7 and 8..11,
// 384 threads) take alternate tiles -- group g owns accumulator stage g -- so that an epilogue may last two main loops:
// with ONE warp per scheduler the epilogue arithmetic is latency-bound (~6 cycles per instruction, profiles/
// r02_bn_bwd_epilogue.txt), which makes the short-contraction layers (7x7 stems, stride-2 and transposed layers) and
// every fused-statistics epilogue epilogue-bound. Registers move with setmaxnreg: the warp group of the TMA / MMA warps
// gives up what the epilogue groups need (56 + 2 x 224 registers per thread of the three groups = 64.5 K).
#ifndef MMH_C2_EPI_GROUPS
#define MMH_C2_EPI_GROUPS 2
#endif
constexpr int kC2EpiGroups = MMH_C2_EPI_GROUPS;
constexpr int kC2EpiWarp0 = kC2EpiGroups == 2 ? 4 : 2;          // first epilogue warp
constexpr int kC2Threads = 32 * (kC2EpiWarp0 + 4 * kC2EpiGroups);
constexpr int kC2MaxGroups = 16;
constexpr int kC2MaxA = 8;
constexpr int kC2MaxB = 8;
constexpr uint32_t kC2TmemCols = 512;
constexpr uint32_t kC2AccStride = 256;

struct Conv2Params {
  int32_t n_groups, cpt, KC, ksteps;
  int32_t N, BN, tiles_n, tiles_m, M;
  int32_t dbg;                     // MMH_C2_DEBUG: 1 skip A loads, 2 skip B loads, 4 skip MMAs, 8 skip stores (timing only)
  int32_t MB;                      // 128-row blocks per tile sharing every weight tile (narrow N, NCTA = 1)
  int32_t Hg, Wg, Hv, Wv;
  int32_t out_f32, out_ld, out_wg, out_sh, out_sw, out_h0, out_w0, zero_invalid, act, n_store;
  int64_t out_img_rows;
  uint32_t row_bytes, swz, sbo;
  uint32_t a_box_rows, a_boxes, a_box_bytes, a_slot_bytes, nA;
  uint32_t b_rows, b_tile_bytes, b_tile_stride, b_batch, b_slot_bytes, nB;
  uint32_t b_ring_off, bar_off;
  void* out;
  const float* bias;
  float* bn_sums;                  // fused BatchNorm statistics: [2][bn_C] += (sum, sum of squares) of the stored values
  int32_t bn_C;
  uint32_t stat_off;               // shared-memory accumulators [2][N] (only when bn_sums != nullptr)
  // fused BatchNorm-BACKWARD statistics of a data-gradient launch (MmhConvDesc.bs_*): bn_sums / bn_C / stat_off are the
  // accumulators [2][C] += (sum dze, sum dze * xhat); par_off = shared-memory copy of (a, b, mean, rstd) [N][4]
  const void* bs_x;
  const float* bs_coef;
  const float* bs_save;
  int32_t bs_x_ld, bs_xHg, bs_xWg, bs_H, bs_W, bs_pad, bs_relu, bs_dropout;
  uint32_t bs_key, par_off;
  int32_t g_first[kC2MaxGroups + 1];
  int32_t g_min[kC2MaxGroups];
  int32_t rel16[MMH_MAX_TAPS + 1];  // byte offset / 16 of the tap's first row inside its group's window (+1: prefetch)
  int32_t w_slot[MMH_MAX_TAPS];
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Epilogue of one 128-row block, one thread per output row: TMEM -> registers -> (bias, activation) -> global.
// ACT / BIAS are compile-time so that the per-element code is straight-line (a run-time switch per element made
// the epilogue instruction-bound: ~35 SASS instructions and four branches per output value).
// 16 columns per step, software-pipelined: the TMEM load of chunk j + 1 is in flight while chunk j is converted
// and stored; every store is one full 32-byte sector (STG.256).
// STATS (bias-free, linear, bf16 output = every BatchNorm'd convolution in training): the per-channel sum and sum of
// squares of the values as stored (bf16-rounded) are reduced over the warp's 32 rows with one butterfly
// reduce-scatter over 32 values (16 sums + 16 squares: 31 shuffles, lane l ends up with value l) and added to the
// CTA's shared-memory accumulators; the statistics pass over the raw output (one full HBM read) disappears.
template <int ACT, bool BIAS, bool STATS = false>
__device__ __forceinline__ void epilogue_row(const Conv2Params& p, uint32_t t_addr, int nchunks, int n0, int64_t orow,
                                             bool valid, bool do_store, float* s_stats = nullptr) {
  auto emit = [&](const uint32_t (&v)[16], int j) {
    const int nc = n0 + j * 16;
    if (nc >= p.n_store || (p.dbg & 8)) return;          // warp-uniform
    if (!STATS && !do_store) return;
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
    if (BIAS) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + nc);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b = __ldg(b4 + i);
        f[4 * i] += b.x; f[4 * i + 1] += b.y; f[4 * i + 2] += b.z; f[4 * i + 3] += b.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (ACT == 1) f[i] = fmaxf(f[i], 0.f);
      else if (ACT == 2) f[i] = tanhf(f[i]);
      f[i] = valid ? f[i] : 0.f;
    }
    if (STATS) {
      uint32_t pk[8];
      float r[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        pk[i] = pack2(f[2 * i], f[2 * i + 1]);
        r[2 * i] = __uint_as_float(pk[i] << 16);
        r[2 * i + 1] = __uint_as_float(pk[i] & 0xFFFF0000u);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) r[16 + i] = r[i] * r[i];
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
          const float send = up ? r[i] : r[i + off];
          const float keep = up ? r[i + off] : r[i];
          r[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      atomicAdd(s_stats + (lane >> 4) * p.N + nc + (lane & 15), r[0]);
      if (do_store)
        st_global_v8(static_cast<__nv_bfloat16*>(p.out) + orow * p.out_ld + nc, pk[0], pk[1], pk[2], pk[3], pk[4],
                     pk[5], pk[6], pk[7]);
      return;
    }
    if (p.out_f32) {
      float* dst = static_cast<float*>(p.out) + orow * p.out_ld + nc;
#pragma unroll
      for (int i = 0; i < 2; ++i)
        st_global_v8(dst + 8 * i, __float_as_uint(f[8 * i]), __float_as_uint(f[8 * i + 1]),
                     __float_as_uint(f[8 * i + 2]), __float_as_uint(f[8 * i + 3]), __float_as_uint(f[8 * i + 4]),
                     __float_as_uint(f[8 * i + 5]), __float_as_uint(f[8 * i + 6]), __float_as_uint(f[8 * i + 7]));
    } else {
      st_global_v8(static_cast<__nv_bfloat16*>(p.out) + orow * p.out_ld + nc, pack2(f[0], f[1]), pack2(f[2], f[3]),
                   pack2(f[4], f[5]), pack2(f[6], f[7]), pack2(f[8], f[9]), pack2(f[10], f[11]), pack2(f[12], f[13]),
                   pack2(f[14], f[15]));
    }
  };
  uint32_t va[16], vb[16];
  tmem_ld16(t_addr, va);
  for (int j = 0; j < nchunks; j += 2) {
    tmem_ld_wait();
    if (j + 1 < nchunks) tmem_ld16(t_addr + (j + 1) * 16, vb);
    emit(va, j);
    if (j + 1 < nchunks) {
      tmem_ld_wait();
      if (j + 2 < nchunks) tmem_ld16(t_addr + (j + 2) * 16, va);
      emit(vb, j + 1);
    }
  }
}

__device__ __forceinline__ uint32_t c2_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ int c2_reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// Epilogue of a data-gradient launch that feeds  conv_p -> BatchNorm -> [ReLU] -> [Dropout]  (MmhConvDesc.bs_*): the
// output row q is the gradient of the (mirrored) logical pixel whose raw conv_p output is the row `xrow` of bs_x. Per
// 16-channel chunk: TMEM accumulators -> ReLU / dropout masks recomputed from x (the kernels' counter hash) ->
// dze rounded to bf16 and stored (the BN-backward apply kernel then needs no masks) -> (sum dze, sum dze * xhat) by the
// same 32-value butterfly reduce-scatter as the forward statistics. The raw activations come in 64-byte pieces, two
// chunks ahead of their use (the caller prefetched the row into L2 one tile earlier); the BN-backward reduction pass over
// dz and x (two full HBM reads) disappears.
// 32 channels (chunks j, j + 1) of a row of the producer's raw output
__device__ __forceinline__ void bs_load_x(const __nv_bfloat16* xrow_n0, bool in_range, int j, int nchunks, uint4 (&x)[4]) {
  if (in_range && j < nchunks) {
    const uint4* src = reinterpret_cast<const uint4*>(xrow_n0 + j * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = ldg_nc_v4(src + i);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// xa / xa2: chunks 0-1 / 2-3 of the row, loaded by the caller BEFORE it waited for the accumulators
__device__ __forceinline__ void epilogue_row_bwd(const Conv2Params& p, uint32_t t_addr, int nchunks, int n0,
                                                 int64_t orow, bool in_range, bool ld_x, const __nv_bfloat16* xrow,
                                                 uint32_t hash_word0, float* s_stats, const float4* s_par,
                                                 uint4 (&xa)[4], uint4 (&xa2)[4]) {
  const int lane = threadIdx.x & 31;
  uint4 xb[4], xb2[4];
  auto load_x = [&](int j, uint4 (&x)[4]) { bs_load_x(xrow + n0, ld_x, j, nchunks, x); };
  auto emit = [&](const uint32_t (&v)[16], const uint4& xa, const uint4& xb, int j) {
    const int nc = n0 + j * 16;
    const uint32_t xw[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    float r[32];
    uint32_t pk[8];
    uint32_t bits = 0xFFFFu;
    if (p.bs_dropout) {
      const uint32_t g = static_cast<uint32_t>(nc >> 3);
      bits = (c2_mix32((hash_word0 + g) * 0x9E3779B1u + p.bs_key) & 0xFFu) |
             ((c2_mix32((hash_word0 + g + 1u) * 0x9E3779B1u + p.bs_key) & 0xFFu) << 8);
    }
    const float keep = p.bs_dropout ? 2.0f : 1.0f;
    float xh[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 c = s_par[nc - n0 + i];                         // (a, b, mean, rstd) of channel nc + i
      const float x = __uint_as_float((i & 1) ? (xw[i >> 1] & 0xFFFF0000u) : (xw[i >> 1] << 16));
      float d = __uint_as_float(v[i]);
      const bool on = in_range && (!p.bs_relu || (c.x * x + c.y > 0.f)) && ((bits >> i) & 1u);
      d = on ? keep * d : 0.f;
      r[i] = d;
      xh[i] = (x - c.z) * c.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pk[i] = pack2(r[2 * i], r[2 * i + 1]);
      r[2 * i] = __uint_as_float(pk[i] << 16);                     // the statistics are those of the stored values
      r[2 * i + 1] = __uint_as_float(pk[i] & 0xFFFF0000u);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) r[16 + i] = r[i] * xh[i];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send = up ? r[i] : r[i + off];
        const float kept = up ? r[i + off] : r[i];
        r[i] = kept + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    atomicAdd(s_stats + (lane >> 4) * p.N + nc + (lane & 15), r[0]);
    if (in_range && !(p.dbg & 8))
      st_global_v8(static_cast<__nv_bfloat16*>(p.out) + orow * p.out_ld + nc, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5],
                   pk[6], pk[7]);
  };
  // 64 channels (128 bytes per row = one cache line) per load group, double-buffered: the loads of group g + 1 are in
  // flight while the four chunks of group g are processed (one group ahead measured 23 us per tile against a 14 us main
  // loop: the epilogue waited on every load)
  uint32_t va[16], vb[16];
  tmem_ld16(t_addr, va);
  for (int j = 0; j < nchunks; j += 8) {
    if (j + 4 < nchunks) { load_x(j + 4, xb); load_x(j + 6, xb2); }
    tmem_ld_wait();
    tmem_ld16(t_addr + (j + 1) * 16, vb);
    emit(va, xa[0], xa[1], j);
    tmem_ld_wait();
    tmem_ld16(t_addr + (j + 2) * 16, va);
    emit(vb, xa[2], xa[3], j + 1);
    tmem_ld_wait();
    tmem_ld16(t_addr + (j + 3) * 16, vb);
    emit(va, xa2[0], xa2[1], j + 2);
    tmem_ld_wait();
    if (j + 4 < nchunks) tmem_ld16(t_addr + (j + 4) * 16, va);
    emit(vb, xa2[2], xa2[3], j + 3);
    if (j + 4 < nchunks) {
      if (j + 8 < nchunks) { load_x(j + 8, xa); load_x(j + 10, xa2); }
      tmem_ld_wait();
      tmem_ld16(t_addr + (j + 5) * 16, vb);
      emit(va, xb[0], xb[1], j + 4);
      tmem_ld_wait();
      tmem_ld16(t_addr + (j + 6) * 16, va);
      emit(vb, xb[2], xb[3], j + 5);
      tmem_ld_wait();
      tmem_ld16(t_addr + (j + 7) * 16, vb);
      emit(va, xb2[0], xb2[1], j + 6);
      tmem_ld_wait();
      if (j + 8 < nchunks) tmem_ld16(t_addr + (j + 8) * 16, va);
      emit(vb, xb2[2], xb2[3], j + 7);
    }
  }
}

// MMA issue loop of one CTA (pair). KS = MMAs of K = 16 per (tap, 64/32/16-channel chunk), MB = 128-row blocks per
// tile; both compile-time so that one MMA costs two uniform adds and nothing else. With run-time bounds the
// compiler emitted a 4 x 4 predicated unroll and ~25 uniform-datapath instructions per MMA: invisible behind
// 128 x 256 x 16 MMAs (128+ cycles each) but 5x the tensor time of the 128 x 64 x 16 MMAs of the 7x7 stems.
template <int NCTA, int KS, int MB>
__device__ __forceinline__ void mma_issue(const Conv2Params& p, uint8_t* a_ring, uint8_t* b_ring, uint64_t* fullA,
                                          uint64_t* emptyA, uint64_t* fullB, uint64_t* emptyB, uint64_t* tmem_full,
                                          uint64_t* tmem_empty, uint32_t tmem_base, int first_tile, int tile_step,
                                          int n_tiles) {
  // descriptor low word = (address >> 4) | LBO field, high word constant: everything the loop adds up is
  // pre-encoded in descriptor units (16 bytes)
  const uint32_t idesc = make_idesc_bf16(128 * NCTA, p.BN, 0, 0);
  const uint64_t desc_proto = make_smem_desc(0, 16, p.sbo, p.swz);
  const uint64_t desc_hi = desc_proto & 0xFFFFFFFF00000000ull;
  const uint32_t a_base = (smem_u32(a_ring) >> 4) + static_cast<uint32_t>(desc_proto);
  const uint32_t b_base = (smem_u32(b_ring) >> 4) + static_cast<uint32_t>(desc_proto);
  const uint32_t a_slot16 = p.a_slot_bytes >> 4, b_slot16 = p.b_slot_bytes >> 4, b_tile16 = p.b_tile_stride >> 4;
  const uint32_t mb16 = (128u * p.row_bytes) >> 4;
  const uint32_t BN = p.BN;
  const int b_batch = p.b_batch, n_groups = p.n_groups, cpt = p.cpt;
  const uint32_t nA = p.nA, nB = p.nB;
  const bool skip = (p.dbg & 4) != 0;
  uint32_t sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, acc_phase = 0;
  for (int tile = first_tile; tile < n_tiles; tile += tile_step) {
    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + acc * kC2AccStride;
    uint32_t fresh = 0;                                   // accumulate flag of the next k = 0 MMA
    for (int g = 0; g < n_groups; ++g) {
      const int t_begin = p.g_first[g], t_end = p.g_first[g + 1];
      for (int kc = 0; kc < cpt; ++kc) {
        mbar_wait(&fullA[sa], pa);
        tc_fence_after();
        const uint32_t wa = a_base + sa * a_slot16;
        for (int t0 = t_begin; t0 < t_end; t0 += b_batch) {
          const int nb = min(b_batch, t_end - t0);
          mbar_wait(&fullB[sb], pb);
          tc_fence_after();
          if (!skip) {
            uint32_t tb = b_base + sb * b_slot16;
            uint32_t rel = static_cast<uint32_t>(p.rel16[t0]);
            for (int j = 0; j < nb; ++j) {
              const uint32_t rel_next = static_cast<uint32_t>(p.rel16[t0 + j + 1]);   // one tap ahead of its use
              const uint32_t ta = wa + rel;
#pragma unroll
              for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                  const uint64_t ad = desc_hi | (ta + mb * mb16 + 2 * k);
                  const uint64_t bd = desc_hi | (tb + 2 * k);
                  if (NCTA == 2) umma_bf16_pair(d_tmem + mb * BN, ad, bd, idesc, k == 0 ? fresh : 1u);
                  else umma_bf16(d_tmem + mb * BN, ad, bd, idesc, k == 0 ? fresh : 1u);
                }
              }
              fresh = 1;
              rel = rel_next;
              tb += b_tile16;
            }
          }
          if (NCTA == 2) umma_commit_pair(&emptyB[sb]); else umma_commit(&emptyB[sb]);
          if (++sb == nB) { sb = 0; pb ^= 1; }
        }
        if (NCTA == 2) umma_commit_pair(&emptyA[sa]); else umma_commit(&emptyA[sa]);
        if (++sa == nA) { sa = 0; pa ^= 1; }
      }
    }
    if (NCTA == 2) umma_commit_pair(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
}

// BS: the data-gradient variant with the fused BatchNorm-backward epilogue (its own instantiation: the extra registers
// of that epilogue must not spill the plain kernel)
template <int NCTA, bool BS = false>
__global__ void __launch_bounds__(kC2Threads, 1)
conv2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ Conv2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + p.b_ring_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + kC2MaxA;
  uint64_t* fullB = bars + 2 * kC2MaxA;
  uint64_t* emptyB = bars + 2 * kC2MaxA + kC2MaxB;
  uint64_t* tmem_full = bars + 2 * kC2MaxA + 2 * kC2MaxB;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = NCTA == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (uint32_t s = 0; s < p.nA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (uint32_t s = 0; s < p.nB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4 * NCTA); }
    fence_mbar_init();
  }
  float* s_stats = reinterpret_cast<float*>(smem + p.stat_off);
  if (p.bn_sums != nullptr)
    for (int i = threadIdx.x; i < 2 * p.N; i += kC2Threads) s_stats[i] = 0.f;
  float4* s_par = reinterpret_cast<float4*>(smem + p.par_off);
  if (warp == 1) {
    if (NCTA == 2) { tmem_alloc2(tmem_slot, kC2TmemCols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, kC2TmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: barriers, TMEM and descriptors above were set up while the preceding kernel of the
  // stream was still running; its results are visible past this point, and the next kernel may start its own prologue
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int n_tiles = p.tiles_m * p.tiles_n;
  const int first_tile = blockIdx.x / NCTA;
  const int tile_step = gridDim.x / NCTA;
  if (BS) {
    // per-channel BatchNorm parameters of the producer (written by the kernel that precedes this one in the stream:
    // read past griddepcontrol.wait): (a, b, mean, rstd) per output channel; padded channels get a = 0, rstd = 0
    for (int i = threadIdx.x; i < p.N; i += kC2Threads) {
      const bool on = i < p.bn_C;
      s_par[i] = make_float4(on ? p.bs_coef[i] : 0.f, on ? p.bs_coef[p.bn_C + i] : 0.f, on ? p.bs_save[i] : 0.f,
                             on ? p.bs_save[p.bn_C + i] : 0.f);
    }
    __syncthreads();
  }

  if (warp < kC2EpiWarp0) {
  if (kC2EpiGroups == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");      // (all four warps of the group)
  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer: windows into the A ring, weight tiles into the B ring, in consumption order
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      for (int tile = first_tile; tile < n_tiles; tile += tile_step) {
        const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
        const int m0 = (tm * NCTA + static_cast<int>(cta)) * 128 * p.MB;
        const int n0 = tn * p.BN + static_cast<int>(cta * p.b_rows);
        for (int g = 0; g < p.n_groups; ++g) {
          const int t_begin = p.g_first[g], t_end = p.g_first[g + 1];
          const int row0 = m0 + p.g_min[g];
          for (int kc = 0; kc < p.cpt; ++kc) {
            const int ch = kc * p.KC;
            mbar_wait(&emptyA[sa], pa ^ 1);
            uint8_t* dst = a_ring + static_cast<size_t>(sa) * p.a_slot_bytes;
            if (p.dbg & 1) {
              if (leader) mbar_arrive(&fullA[sa]);
            } else if (NCTA == 2) {
              const uint32_t bar = mapa_shared(smem_u32(&fullA[sa]), 0);
              if (leader) mbar_expect_tx(&fullA[sa], 2 * p.a_boxes * p.a_box_bytes);
              for (uint32_t b = 0; b < p.a_boxes; ++b)
                tma_load_2d_pair(&tmA, bar, dst + b * p.a_box_bytes, ch, row0 + static_cast<int>(b * p.a_box_rows));
            } else {
              mbar_expect_tx(&fullA[sa], p.a_boxes * p.a_box_bytes);
              for (uint32_t b = 0; b < p.a_boxes; ++b)
                tma_load_2d(&tmA, &fullA[sa], dst + b * p.a_box_bytes, ch, row0 + static_cast<int>(b * p.a_box_rows));
            }
            if (++sa == p.nA) { sa = 0; pa ^= 1; }
            for (int t0 = t_begin; t0 < t_end; t0 += p.b_batch) {
              const int nb = min(static_cast<int>(p.b_batch), t_end - t0);
              mbar_wait(&emptyB[sb], pb ^ 1);
              uint8_t* bd = b_ring + static_cast<size_t>(sb) * p.b_slot_bytes;
              if (p.dbg & 2) {
                if (leader) mbar_arrive(&fullB[sb]);
              } else if (NCTA == 2) {
                const uint32_t bar = mapa_shared(smem_u32(&fullB[sb]), 0);
                if (leader) mbar_expect_tx(&fullB[sb], 2 * nb * p.b_tile_bytes);
                for (int j = 0; j < nb; ++j)
                  tma_load_2d_pair(&tmW, bar, bd + j * p.b_tile_stride, ch, p.w_slot[t0 + j] * p.N + n0);
              } else {
                mbar_expect_tx(&fullB[sb], nb * p.b_tile_bytes);
                for (int j = 0; j < nb; ++j)
                  tma_load_2d(&tmW, &fullB[sb], bd + j * p.b_tile_stride, ch, p.w_slot[t0 + j] * p.N + n0);
              }
              if (++sb == p.nB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ===== MMA issuer (one elected lane of the pair's leader CTA)
#define MMH_ISSUE(KS_, MB_)                                                                                          \
  mma_issue<NCTA, KS_, MB_>(p, a_ring, b_ring, fullA, emptyA, fullB, emptyB, tmem_full, tmem_empty, tmem_base,       \
                            first_tile, tile_step, n_tiles)
      const int mbv = NCTA == 2 ? 1 : p.MB;
      switch (p.ksteps * 8 + mbv) {
        case 4 * 8 + 1: MMH_ISSUE(4, 1); break;
        case 2 * 8 + 1: MMH_ISSUE(2, 1); break;
        case 1 * 8 + 1: MMH_ISSUE(1, 1); break;
        case 4 * 8 + 2: if (NCTA == 1) MMH_ISSUE(4, 2); break;
        case 2 * 8 + 2: if (NCTA == 1) MMH_ISSUE(2, 2); break;
        case 1 * 8 + 2: if (NCTA == 1) MMH_ISSUE(1, 2); break;
        case 4 * 8 + 4: if (NCTA == 1) MMH_ISSUE(4, 4); break;
        case 2 * 8 + 4: if (NCTA == 1) MMH_ISSUE(2, 4); break;
        case 1 * 8 + 4: if (NCTA == 1) MMH_ISSUE(1, 4); break;
        default: break;
      }
#undef MMH_ISSUE
    }
  }
  } else {
    // ===== epilogue warps: TMEM lane quadrant = warp id % 4; group = accumulator stage it drains
    if (kC2EpiGroups == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int quad = warp & 3;
    const int grp = (warp - kC2EpiWarp0) >> 2;
    const int row = quad * 32 + lane;
    const int hw = p.Hg * p.Wg;
    const int nchunks = p.BN / 16;
    const uint32_t empty_remote = NCTA == 2 ? mapa_shared(smem_u32(&tmem_empty[0]), 0) : 0u;
    const int my_step = tile_step * kC2EpiGroups;        // distance between two tiles of this group
    // fused BN-backward statistics (MB == 1): row -> mirrored logical pixel -> row of the producer's raw output
    auto bs_row = [&](int tile, int64_t& xrow, uint32_t& word0, bool& in_range, int64_t& orow) {
      const int tm = tile / p.tiles_n;
      const int q = (tm * NCTA + static_cast<int>(cta)) * 128 + row;
      const int img = q / hw;
      const int rem = q - img * hw;
      const int h = rem / p.Wg;
      const int x = rem - h * p.Wg;
      in_range = q < p.M;
      const int hs = c2_reflect(h - p.bs_pad, p.bs_H), ws = c2_reflect(x - p.bs_pad, p.bs_W);
      xrow = (static_cast<int64_t>(img) * p.bs_xHg + hs) * p.bs_xWg + ws;
      word0 = ((static_cast<uint32_t>(img) * p.bs_H + hs) * p.bs_W + ws) * static_cast<uint32_t>((p.bn_C + 7) / 8);
      orow = static_cast<int64_t>(img) * p.out_img_rows + static_cast<int64_t>(h * p.out_sh + p.out_h0) * p.out_wg +
             (x * p.out_sw + p.out_w0);
    };
    auto bs_prefetch = [&](int tile) {
      if (tile >= n_tiles) return;
      int64_t xrow, orow; uint32_t w0; bool in_range;
      bs_row(tile, xrow, w0, in_range, orow);
      if (!in_range || (p.dbg & 16)) return;
      const int tn = tile % p.tiles_n;
      const char* base = reinterpret_cast<const char*>(p.bs_x) + (xrow * p.bs_x_ld + tn * p.BN) * 2;
      for (int b = 0; b < p.BN * 2; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + b));
    };
    if (BS) bs_prefetch(first_tile + grp * tile_step);
    int kt = 0;                                          // k-th tile of this CTA: stage kt & 1, phase (kt >> 1) & 1
    for (int tile = first_tile; tile < n_tiles; tile += tile_step, ++kt) {
      if (kC2EpiGroups == 2 && (kt & 1) != grp) continue;
      const uint32_t acc = kt & 1, acc_phase = (kt >> 1) & 1;
      const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
      const int n0 = tn * p.BN;
      if (BS) {
        int64_t xrow, orow; uint32_t w0; bool in_range;
        bs_row(tile, xrow, w0, in_range, orow);
        bs_prefetch(tile + my_step);                    // the group's next tile: its rows travel to L2 under this one
        const __nv_bfloat16* xr = static_cast<const __nv_bfloat16*>(p.bs_x) + xrow * p.bs_x_ld;
        uint4 xa[4], xa2[4];                            // first 64 channels: in flight while the main loop finishes
        const bool ld_x = in_range && !(p.dbg & 16);    // MMH_C2_DEBUG=16: no loads of x (timing experiments only)
        bs_load_x(xr + n0, ld_x, 0, nchunks, xa);
        bs_load_x(xr + n0, ld_x, 2, nchunks, xa2);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + acc * kC2AccStride + (static_cast<uint32_t>(quad * 32) << 16);
        epilogue_row_bwd(p, t_addr, nchunks, n0, orow, in_range, ld_x, xr, w0, s_stats, s_par + n0, xa, xa2);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2) mbar_arrive_cluster(empty_remote + acc * 8);
          else mbar_arrive(&tmem_empty[acc]);
        }
        continue;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      for (int mb = 0; mb < p.MB; ++mb) {
        const int q = ((tm * NCTA + static_cast<int>(cta)) * p.MB + mb) * 128 + row;
        const int img = q / hw;
        const int rem = q - img * hw;
        const int h = rem / p.Wg;
        const int x = rem - h * p.Wg;
        const bool in_range = q < p.M;
        const bool valid = in_range && h < p.Hv && x < p.Wv;
        const bool do_store = valid || (in_range && p.zero_invalid);
        const int64_t orow = static_cast<int64_t>(img) * p.out_img_rows +
                             static_cast<int64_t>(h * p.out_sh + p.out_h0) * p.out_wg + (x * p.out_sw + p.out_w0);
        const uint32_t t_addr = tmem_base + acc * kC2AccStride + mb * p.BN + (static_cast<uint32_t>(quad * 32) << 16);
        // bias-free linear layers (every BatchNorm'd convolution, every data gradient) take the straight-line path
        if (p.bn_sums != nullptr) epilogue_row<0, false, true>(p, t_addr, nchunks, n0, orow, valid, do_store, s_stats);
        else if (p.bias == nullptr && p.act == 0) epilogue_row<0, false>(p, t_addr, nchunks, n0, orow, valid, do_store);
        else if (p.bias == nullptr) { if (p.act == 1) epilogue_row<1, false>(p, t_addr, nchunks, n0, orow, valid, do_store);
                                      else epilogue_row<2, false>(p, t_addr, nchunks, n0, orow, valid, do_store); }
        else if (p.act == 0) epilogue_row<0, true>(p, t_addr, nchunks, n0, orow, valid, do_store);
        else if (p.act == 1) epilogue_row<1, true>(p, t_addr, nchunks, n0, orow, valid, do_store);
        else epilogue_row<2, true>(p, t_addr, nchunks, n0, orow, valid, do_store);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_cluster(empty_remote + acc * 8);
        else mbar_arrive(&tmem_empty[acc]);
      }
    }
    if (p.bn_sums != nullptr) {
      // the epilogue warps are done with their tiles: CTA partial sums -> global accumulators
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kC2EpiGroups) : "memory");
      for (int i = threadIdx.x - 32 * kC2EpiWarp0; i < 2 * p.N; i += 128 * kC2EpiGroups) {
        const int st = i >= p.N ? 1 : 0, c = i - st * p.N;
        const float v = s_stats[i];
        if (c < p.bn_C && v != 0.f) atomicAdd(p.bn_sums + st * p.bn_C + c, v);
      }
    }
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc2(tmem_base, kC2TmemCols);
    else tmem_dealloc(tmem_base, kC2TmemCols);
  }
}

}  // namespace mmh

// ------------------------------------------------------------------------------------------------
using namespace mmh;

struct MmhConv2 {
  CUtensorMap tmA, tmW;
  Conv2Params kp;
  int grid, ncta;
  size_t smem;
};

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

int mmh_conv2_create(const MmhConvDesc* d, MmhConv2** out_plan) {
  auto* plan = new MmhConv2();
  Conv2Params& k = plan->kp;
  memset(&k, 0, sizeof(k));
  auto fail = [&]() { delete plan; return 1; };
  k.KC = (d->C % 64) == 0 ? 64 : ((d->C % 32) == 0 ? 32 : 16);
  k.cpt = d->C / k.KC;
  k.ksteps = k.KC / 16;
  k.row_bytes = k.KC * 2;
  k.swz = k.KC == 64 ? 2u : (k.KC == 32 ? 4u : 6u);
  k.sbo = 8 * k.row_bytes;
  k.N = d->N;
  if (d->N <= 256) {
    k.BN = d->N;
  } else {
    if ((d->N % 128) != 0) { set_error("N=%d unsupported", d->N); return fail(); }
    k.BN = (d->N % 256) == 0 ? 256 : 128;
  }
  k.tiles_n = d->N / k.BN;
  k.M = static_cast<int32_t>(d->M);
  // pairs: worth it when the weight tile dominates the traffic and there are enough row tiles to fill 74 pairs
  int ncta = env_int("MMH_CONV_NCTA", 0);
  if (ncta == 0) ncta = (k.BN >= 128 && (k.BN % 32) == 0 && d->M >= 148 * 128) ? 2 : 1;
  if (d->bs_x != nullptr && k.BN > 128) ncta = 2;          // the fused BN-backward epilogue handles one row block per CTA
  if (ncta == 2 && (k.BN % 32) != 0) ncta = 1;
  plan->ncta = ncta;
  k.tiles_m = (k.M + 128 * ncta - 1) / (128 * ncta);
  k.Hg = d->Hg; k.Wg = d->Wg; k.Hv = d->Hv; k.Wv = d->Wv;
  k.out_f32 = d->out_f32; k.out_ld = d->out_ld; k.out_wg = d->out_wg;
  k.out_sh = d->out_sh; k.out_sw = d->out_sw; k.out_h0 = d->out_h0; k.out_w0 = d->out_w0;
  k.zero_invalid = d->zero_invalid; k.act = d->act;
  k.n_store = d->n_store > 0 ? d->n_store : d->N;
  k.out_img_rows = d->out_img_rows;
  k.out = d->out;
  k.bias = d->bias;
  k.bn_sums = d->bn_sums;
  k.bn_C = d->bn_C;
  if (d->bn_sums != nullptr) {
    if (d->bias != nullptr || d->act != 0 || d->out_f32 || d->bn_C <= 0 || d->bn_C > d->N) {
      set_error("fused BN statistics need a bias-free linear bf16 convolution (bn_C=%d, N=%d)", d->bn_C, d->N);
      return fail();
    }
  }
  const bool bs = d->bs_x != nullptr;
  if (bs) {
    if (d->bn_sums != nullptr || d->bias != nullptr || d->act != 0 || d->out_f32 || d->bs_sums == nullptr ||
        d->bs_coef == nullptr || d->bs_save == nullptr || d->bs_C <= 0 || d->bs_C > d->N || (k.BN % 64) != 0 ||
        d->Hv != d->Hg || d->Wv != d->Wg || (d->bs_x_ld % 8) != 0 || d->bs_pad < 0 || d->bs_pad >= d->bs_H ||
        d->bs_pad >= d->bs_W) {
      set_error("fused BN-backward statistics need a bias-free linear bf16 data-gradient launch over the whole grid "
                "(bs_C=%d, N=%d, BN=%d)", d->bs_C, d->N, k.BN);
      return fail();
    }
    k.bn_sums = d->bs_sums; k.bn_C = d->bs_C;
    k.bs_x = d->bs_x; k.bs_coef = d->bs_coef; k.bs_save = d->bs_save;
    k.bs_x_ld = d->bs_x_ld; k.bs_xHg = d->bs_xHg; k.bs_xWg = d->bs_xWg; k.bs_H = d->bs_H; k.bs_W = d->bs_W;
    k.bs_pad = d->bs_pad; k.bs_relu = d->bs_relu; k.bs_dropout = d->bs_dropout; k.bs_key = d->bs_drop_key;
  }
  const uint32_t stat_bytes = (d->bn_sums != nullptr || bs) ? ((2u * d->N * 4u + (bs ? 16u * d->N : 0u) + 1023u) & ~1023u) : 0u;

  // ---- tap groups: sort by shift, start a new group at a gap of >= 128 rows or when the window would
  // outgrow its slot
  const int w_taps = d->w_taps > 0 ? d->w_taps : d->T;
  std::vector<std::pair<int, int>> taps;   // (shift, slot)
  for (int t = 0; t < d->T; ++t) {
    const int slot = d->w_taps > 0 ? d->w_slot[t] : t;
    if (slot < 0 || slot >= w_taps) { set_error("w_slot[%d] out of range", t); return fail(); }
    taps.emplace_back(d->shift[t], slot);
  }
  std::stable_sort(taps.begin(), taps.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) {
    return a.first < b.first;
  });
  const int cap_rows = static_cast<int>(40960 / k.row_bytes);
  int ng = 0, span_max = 0, gtaps_max = 0;
  for (int t = 0; t < d->T; ++t) {
    const bool fresh = t == 0 || taps[t].first - taps[t - 1].first >= 128 ||
                       taps[t].first - k.g_min[ng - 1] + 128 > cap_rows;
    if (fresh) {
      if (ng == kC2MaxGroups) { set_error("too many tap groups"); return fail(); }
      k.g_first[ng] = t;
      k.g_min[ng] = taps[t].first;
      ++ng;
    }
    const int rel = taps[t].first - k.g_min[ng - 1];
    k.rel16[t] = static_cast<int32_t>((static_cast<uint32_t>(rel) * k.row_bytes) >> 4);
    k.w_slot[t] = taps[t].second;
    span_max = std::max(span_max, rel);
  }
  k.g_first[ng] = d->T;
  k.n_groups = ng;
  for (int g = 0; g < ng; ++g) gtaps_max = std::max(gtaps_max, k.g_first[g + 1] - k.g_first[g]);

  int mb = 1;
  if (ncta == 1 && k.BN <= 128) {
    mb = 256 / k.BN;
    if (mb > 4) mb = 4;
    while (mb > 1 && (static_cast<uint32_t>(128 * mb + span_max) * k.row_bytes > 49152u ||
                      static_cast<int64_t>((d->M + 128 * mb - 1) / (128 * mb)) * k.tiles_n < 2 * num_sms()))
      mb >>= 1;
  }
  mb = env_int("MMH_CONV_MB", mb);
  if ((mb != 1 && mb != 2 && mb != 4) || mb * k.BN > 256 || ncta != 1 || d->bs_x != nullptr) mb = 1;
  k.MB = mb;
  k.dbg = env_int("MMH_C2_DEBUG", 0);
  k.tiles_m = (k.M + 128 * ncta * mb - 1) / (128 * ncta * mb);
  const int need_rows = 128 * mb + span_max;
  k.a_boxes = (need_rows + 255) / 256;
  k.a_box_rows = ((need_rows + k.a_boxes - 1) / k.a_boxes + 7) / 8 * 8;
  k.a_box_bytes = k.a_box_rows * k.row_bytes;
  k.a_slot_bytes = (k.a_boxes * k.a_box_bytes + 1023u) & ~1023u;
  k.b_rows = k.BN / ncta;
  k.b_tile_bytes = k.b_rows * k.row_bytes;
  k.b_tile_stride = (k.b_tile_bytes + 1023u) & ~1023u;
  int batch = static_cast<int>(32768 / k.b_tile_stride);
  batch = std::max(1, std::min(batch, gtaps_max));
  batch = env_int("MMH_CONV_BBATCH", batch);
  k.b_batch = batch;
  k.b_slot_bytes = k.b_batch * k.b_tile_stride;
  // MMH_CONV_SMEM_KB (default 227 = everything): a smaller cap leaves shared memory for the reduction kernels of another
  // layer chain to be co-resident with a convolution CTA (engine.py: chains on their own CUDA streams)
  static const uint32_t cap_kb = [] {
    const int v = env_int("MMH_CONV_SMEM_KB", 227);
    return static_cast<uint32_t>(v < 96 ? 96 : (v > 227 ? 227 : v));
  }();
  const uint32_t budget = cap_kb * 1024 - 1024 /*align*/ - 512 /*barriers*/ - stat_bytes;
  k.nA = env_int("MMH_CONV_NA", k.a_slot_bytes <= 8192 ? 8 : (k.a_slot_bytes <= 20480 ? 4 : 2));
  if (k.nA > kC2MaxA) k.nA = kC2MaxA;
  if (k.nA * k.a_slot_bytes + 2 * k.b_slot_bytes > budget) { set_error("conv tile does not fit in shared memory"); return fail(); }
  k.nB = (budget - k.nA * k.a_slot_bytes) / k.b_slot_bytes;
  if (k.nB > kC2MaxB) k.nB = kC2MaxB;
  k.b_ring_off = k.nA * k.a_slot_bytes;
  k.bar_off = k.b_ring_off + k.nB * k.b_slot_bytes;
  k.stat_off = k.bar_off + 512;
  k.par_off = k.stat_off + 2u * d->N * 4u;                 // 16-byte aligned: N is a multiple of 16
  plan->smem = k.bar_off + 512 + stat_bytes + 1024;

  const CUtensorMapSwizzle swz = k.KC == 64   ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : k.KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_32B;
  if (make_tmap_2d_bf16(&plan->tmA, d->a, d->C, d->a_rows, d->a_ld, k.KC, k.a_box_rows, swz)) return fail();
  if (make_tmap_2d_bf16(&plan->tmW, d->w, d->C, static_cast<int64_t>(w_taps) * d->N, d->C, k.KC, k.b_rows, swz))
    return fail();
  const int tiles = k.tiles_m * k.tiles_n;
  const int units = num_sms() / ncta;
  plan->grid = (tiles < units ? tiles : units) * ncta;
  cudaError_t e;
  if (bs) e = ncta == 2 ? cudaFuncSetAttribute(conv2_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                        : cudaFuncSetAttribute(conv2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  else e = ncta == 2 ? cudaFuncSetAttribute(conv2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                     : cudaFuncSetAttribute(conv2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv2_kernel): %s", cudaGetErrorString(e)); return fail(); }
  *out_plan = plan;
  return 0;
}

void mmh_conv2_destroy(MmhConv2* plan) { delete plan; }

int mmh_conv2_run_key(const MmhConv2* plan, uint32_t drop_key, void* stream);
int mmh_conv2_run(const MmhConv2* plan, void* stream) { return mmh_conv2_run_key(plan, plan->kp.bs_key, stream); }

int mmh_conv2_run_key(const MmhConv2* plan, uint32_t drop_key, void* stream) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(plan->grid, 1, 1);
  cfg.blockDim = dim3(kC2Threads, 1, 1);
  cfg.dynamicSmemBytes = plan->smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  const bool pdl = pdl_enabled();
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const bool bs = plan->kp.bs_x != nullptr;
  Conv2Params kp_key;
  const Conv2Params* kp = &plan->kp;
  if (bs && drop_key != plan->kp.bs_key) {
    kp_key = plan->kp;                       // launch parameters are copied at launch: per-step dropout key
    kp_key.bs_key = drop_key;
    kp = &kp_key;
  }
  if (plan->ncta == 2) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
    if (bs) MMH_CUDA(cudaLaunchKernelEx(&cfg, conv2_kernel<2, true>, plan->tmA, plan->tmW, *kp));
    else MMH_CUDA(cudaLaunchKernelEx(&cfg, conv2_kernel<2>, plan->tmA, plan->tmW, *kp));
  } else {
    if (bs) MMH_CUDA(cudaLaunchKernelEx(&cfg, conv2_kernel<1, true>, plan->tmA, plan->tmW, *kp));
    else MMH_CUDA(cudaLaunchKernelEx(&cfg, conv2_kernel<1>, plan->tmA, plan->tmW, *kp));
  }
  return 0;
}
