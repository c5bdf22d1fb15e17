// Shifted-row implicit-GEMM convolution on tcgen05 (fprop, dgrad, transposed-conv phases, VGG).
//
//   out[map(q)][n] = act(bias[n] + sum_t sum_c a[q + shift[t]][c] * w[t][n][c])
//
// One persistent CTA per SM, 6 warps:
//   warp 0      TMA producer: per pipeline stage, SUB (tap, channel-chunk) items; each item is one
//               128-row x KC-channel box of the activation grid (rows shifted by the tap's offset; rows
//               outside the tensor are zero-filled by TMA) and one BN x KC box of the packed weights.
//   warp 1      MMA issuer (one elected lane): tcgen05.mma 128 x BN x 16 per K step, accumulating in
//               TMEM; two accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the
//               main loop of tile i+1. Also owns TMEM alloc/dealloc.
//   warps 2..5  epilogue: tcgen05.ld (lane = output row), bias/activation, bf16 or fp32 NHWC store.
//
// Operand layout in shared memory is the canonical K-major swizzled layout (SWIZZLE_128B for 64-channel
// chunks, 64B for 32, 32B for 16) that both TMA and the UMMA descriptors understand.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/mmhand_sm100.h"
#include "conv_plan.h"
#include "host_common.h"
#include "ptx.cuh"
#include "tmap.h"

namespace mmh {

constexpr int kBM = 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;

struct ConvKParams {
  int32_t T, C, KC, SUB, cpt, n_items, n_iters;
  int32_t N, BN, tiles_n, tiles_m;
  int32_t M;
  int32_t Hg, Wg, Hv, Wv;
  int32_t out_f32, out_ld, out_wg, out_sh, out_sw, out_h0, out_w0, zero_invalid, act, n_store;
  int64_t out_img_rows;
  uint32_t a_sub_bytes, b_sub_bytes, b_sub_stride, stage_bytes, n_stages, swz, sbo;
  void* out;
  const float* bias;
  int32_t shift[MMH_MAX_TAPS];
  int32_t w_slot[MMH_MAX_TAPS];
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kThreads, 1)
conv_sgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ ConvKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment of the operand stages (required by SWIZZLE_128B).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.n_stages) * p.stage_bytes);
  uint64_t* full_bar = bars;                      // [n_stages]
  uint64_t* empty_bar = bars + kMaxStages;        // [n_stages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;    // [2]
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (uint32_t s = 0; s < p.n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles = p.tiles_m * p.tiles_n;
  const uint32_t a_stage_bytes = p.a_sub_bytes * p.SUB;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
        const int m0 = tm * kBM, n0 = tn * p.BN;
        for (int it = 0; it < p.n_iters; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const int first = it * p.SUB;
          const int items = min(p.SUB, p.n_items - first);
          mbar_expect_tx(&full_bar[stage], items * (p.a_sub_bytes + p.b_sub_bytes));
          uint8_t* sa = stage_base + static_cast<size_t>(stage) * p.stage_bytes;
          uint8_t* sb = sa + a_stage_bytes;
          for (int s = 0; s < items; ++s) {
            const int item = first + s;
            const int t = item / p.cpt, kc = item - t * p.cpt;
            tma_load_2d(&tmA, &full_bar[stage], sa + s * p.a_sub_bytes, kc * p.KC, m0 + p.shift[t]);
            tma_load_2d(&tmW, &full_bar[stage], sb + s * p.b_sub_stride, kc * p.KC, p.w_slot[t] * p.N + n0);
          }
          if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kBM, p.BN, 0, 0);
      const int ksteps = p.KC / 16;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int it = 0; it < p.n_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const int items = min(p.SUB, p.n_items - it * p.SUB);
          const uint32_t sa = smem_u32(stage_base + static_cast<size_t>(stage) * p.stage_bytes);
          const uint32_t sb = sa + a_stage_bytes;
          for (int s = 0; s < items; ++s) {
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t ad = make_smem_desc(sa + s * p.a_sub_bytes + k * 32, 16, p.sbo, p.swz);
              const uint64_t bd = make_smem_desc(sb + s * p.b_sub_stride + k * 32, 16, p.sbo, p.swz);
              umma_bf16(d_tmem, ad, bd, idesc, (it | s | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // epilogue warps: TMEM lane quadrant is fixed by warp id % 4
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int hw = p.Hg * p.Wg;
    const int nchunks = p.BN / 16;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
      const int q = tm * kBM + row;
      const int n0 = tn * p.BN;
      const int img = q / hw;
      const int rem = q - img * hw;
      const int h = rem / p.Wg;
      const int x = rem - h * p.Wg;
      const bool in_range = q < p.M;
      const bool valid = in_range && h < p.Hv && x < p.Wv;
      const bool do_store = valid || (in_range && p.zero_invalid);
      const int64_t orow = static_cast<int64_t>(img) * p.out_img_rows +
                           static_cast<int64_t>(h * p.out_sh + p.out_h0) * p.out_wg + (x * p.out_sw + p.out_w0);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * kAccStride + (static_cast<uint32_t>(quad * 32) << 16);
      for (int j = 0; j < nchunks; ++j) {
        uint32_t v[16];
        tmem_ld16(t_addr + j * 16, v);
        tmem_ld_wait();
        const int nc = n0 + j * 16;
        if (do_store && nc < p.n_store) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float val = valid ? __uint_as_float(v[i]) : 0.f;
            if (valid) {
              if (p.bias != nullptr) val += __ldg(p.bias + nc + i);
              if (p.act == 1) val = fmaxf(val, 0.f);
              else if (p.act == 2) val = tanhf(val);
            }
            f[i] = val;
          }
          if (p.out_f32) {
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.out_ld + nc);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + orow * p.out_ld + nc);
            dst[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                pack_bf16x2(f[6], f[7]));
            dst[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
                                pack_bf16x2(f[14], f[15]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace mmh

// ------------------------------------------------------------------------------------------------
struct MmhConvPlan {
  CUtensorMap tmA, tmW;
  mmh::ConvKParams kp;
  int grid;
  size_t smem;
  MmhConv2* v2 = nullptr;   // generation-2 plan (default); MMH_CONV_IMPL=1 selects the first-generation kernel
  ~MmhConvPlan() { if (v2) mmh_conv2_destroy(v2); }
};

using namespace mmh;


extern "C" int mmh_conv_plan_create(const MmhConvDesc* d, MmhConvPlan** out_plan) {
  MMH_CHECK(d && out_plan, "null argument");
  MMH_CHECK(d->T >= 1 && d->T <= MMH_MAX_TAPS, "T=%d out of range", d->T);
  MMH_CHECK(d->C >= 16 && (d->C % 16) == 0, "C=%d must be a multiple of 16", d->C);
  MMH_CHECK(d->C == 16 || d->C == 32 || d->C == 48 || (d->C % 64) == 0, "C=%d unsupported", d->C);
  MMH_CHECK(d->N >= 16 && (d->N % 16) == 0, "N=%d must be a multiple of 16", d->N);
  MMH_CHECK((d->a_ld % 8) == 0, "a_ld=%d invalid", d->a_ld);
  MMH_CHECK((d->out_ld % 8) == 0, "out_ld=%d must be a multiple of 8", d->out_ld);
  MMH_CHECK(d->M > 0 && d->M < (int64_t(1) << 31) - 256, "M out of range");
  MMH_CHECK((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
            "operand pointers must be 16-byte aligned");
  auto* plan = new MmhConvPlan();
  {
    const char* impl = getenv("MMH_CONV_IMPL");
    if (impl == nullptr || atoi(impl) != 1) {
      if (mmh_conv2_create(d, &plan->v2)) { delete plan; return 1; }
      *out_plan = plan;
      return 0;
    }
  }
  if (d->bn_sums != nullptr) { delete plan; MMH_CHECK(false, "fused BN statistics need the generation-2 kernel"); }
  ConvKParams& k = plan->kp;
  memset(&k, 0, sizeof(k));
  k.T = d->T;
  k.C = d->C;
  k.KC = (d->C % 64) == 0 ? 64 : ((d->C % 32) == 0 ? 32 : 16);
  k.SUB = 64 / k.KC;
  k.cpt = d->C / k.KC;
  k.n_items = k.T * k.cpt;
  k.n_iters = (k.n_items + k.SUB - 1) / k.SUB;
  k.N = d->N;
  if (d->N <= 256) {
    k.BN = d->N;
  } else {
    MMH_CHECK((d->N % 256) == 0 || (d->N % 128) == 0, "N=%d unsupported", d->N);
    k.BN = (d->N % 256) == 0 ? 256 : 128;
  }
  k.tiles_n = d->N / k.BN;
  k.M = static_cast<int32_t>(d->M);
  k.tiles_m = (k.M + kBM - 1) / kBM;
  k.Hg = d->Hg; k.Wg = d->Wg; k.Hv = d->Hv; k.Wv = d->Wv;
  k.out_f32 = d->out_f32; k.out_ld = d->out_ld; k.out_wg = d->out_wg;
  k.out_sh = d->out_sh; k.out_sw = d->out_sw; k.out_h0 = d->out_h0; k.out_w0 = d->out_w0;
  k.zero_invalid = d->zero_invalid; k.act = d->act;
  k.n_store = d->n_store > 0 ? d->n_store : d->N;
  k.out_img_rows = d->out_img_rows;
  k.out = d->out;
  k.bias = d->bias;
  const int w_taps = d->w_taps > 0 ? d->w_taps : d->T;
  for (int t = 0; t < d->T; ++t) {
    k.shift[t] = d->shift[t];
    k.w_slot[t] = d->w_taps > 0 ? d->w_slot[t] : t;
    if (k.w_slot[t] < 0 || k.w_slot[t] >= w_taps) { set_error("w_slot[%d] out of range", t); delete plan; return 1; }
  }
  k.a_sub_bytes = kBM * k.KC * 2;
  k.b_sub_bytes = k.BN * k.KC * 2;
  k.b_sub_stride = (k.b_sub_bytes + 1023u) & ~1023u;
  k.stage_bytes = k.SUB * (k.a_sub_bytes + k.b_sub_stride);
  const uint32_t budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/;
  k.n_stages = budget / k.stage_bytes;
  if (k.n_stages > kMaxStages) k.n_stages = kMaxStages;
  MMH_CHECK(k.n_stages >= 2, "tile does not fit in shared memory");
  k.swz = k.KC == 64 ? 2u : (k.KC == 32 ? 4u : 6u);
  k.sbo = 8 * k.KC * 2;
  plan->smem = static_cast<size_t>(k.n_stages) * k.stage_bytes + 1024 + 256;

  const CUtensorMapSwizzle swz = k.KC == 64   ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : k.KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_32B;
  if (make_tmap_2d_bf16(&plan->tmA, d->a, d->C, d->a_rows, d->a_ld, k.KC, kBM, swz)) { delete plan; return 1; }
  if (make_tmap_2d_bf16(&plan->tmW, d->w, d->C, static_cast<int64_t>(w_taps) * d->N, d->C, k.KC, k.BN, swz)) {
    delete plan;
    return 1;
  }
  const int tiles = k.tiles_m * k.tiles_n;
  const int sms = num_sms();
  plan->grid = tiles < sms ? tiles : sms;
  cudaError_t e = cudaFuncSetAttribute(conv_sgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_sgemm_kernel): %s", cudaGetErrorString(e));
    delete plan;
    return 1;
  }
  *out_plan = plan;
  return 0;
}

extern "C" int mmh_conv_plan_destroy(MmhConvPlan* plan) {
  delete plan;
  return 0;
}

extern "C" int mmh_conv_run(const MmhConvPlan* plan, void* stream) {
  MMH_CHECK(plan, "null plan");
  if (plan->v2) return mmh_conv2_run(plan->v2, stream);
  conv_sgemm_kernel<<<plan->grid, kThreads, plan->smem, static_cast<cudaStream_t>(stream)>>>(plan->tmA, plan->tmW,
                                                                                           plan->kp);
  MMH_CUDA(cudaGetLastError());
  return 0;
}
