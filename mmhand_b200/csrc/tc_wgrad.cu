// Weight gradient as a pixel-contraction GEMM on tcgen05:
//   dw[t][n][c] += sum_q dy[q][n] * a[q + shift[t]][c]
// Both operands are "MN-major" for the tensor core: the contraction index (pixel row q) is the slow
// dimension of the NHWC grids, channels are contiguous. TMA brings 64-row x CW-channel boxes (CW = 64,
// 32 or 16 channels with 128/64/32-byte swizzle); the UMMA descriptors walk them with
// LBO = bytes between channel groups (one box), SBO = bytes between 8-row groups.
//
// One CTA = one (tap group, 128 dy-channels, <=256 a-channels, K split). A tap group is G consecutive taps
// whose accumulators sit side by side in TMEM (G * BNc <= 384 columns): the dy tile of a pipeline stage is
// loaded once and multiplied against the G shifted activation tiles, which is what makes the 49-tap stem
// convolutions (tiny channel counts, 1.1 M pixel rows) cheap. fp32 partial sums are reduced into dw with
// red.global.add.f32.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mmhand_sm100.h"
#include "conv_plan.h"
#include "host_common.h"
#include "ptx.cuh"
#include "tmap.h"

namespace mmh {

constexpr int kWgThreads = 192;
constexpr int kWgBK = 64;      // pixel rows per pipeline stage
constexpr int kWgMaxStages = 8;
constexpr int kWgMaxCols = 384;

struct WgradKParams {
  int32_t T, G, n_tg, tiles_n, tiles_c, split;
  int32_t BNc;                 // a-channels per tile (instruction N)
  int32_t cw_n, cw_c;          // channel-group widths (elements) of dy / a boxes
  int32_t boxes_n, boxes_c;    // boxes per stage for dy (128/cw_n) and per tap for a (BNc/cw_c)
  int32_t ksteps_total;        // ceil(M / 64)
  int32_t N_store, C_store, dw_taps;
  uint32_t sub_n_bytes, sub_c_bytes, tap_bytes, stage_bytes, n_stages, a_stage_off;
  uint32_t swz_n, swz_c, lbo_n, sbo_n, lbo_c, sbo_c, tmem_cols;
  float* dw;
  int32_t shift[MMH_MAX_TAPS];
  int32_t tap_index[MMH_MAX_TAPS];
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmA,
             const __grid_constant__ WgradKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.n_stages) * p.stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgMaxStages;
  uint64_t* acc_bar = bars + 2 * kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgMaxStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work unit decode: split fastest so that CTAs sharing the same dw tile are spread over time
  int unit = blockIdx.x;
  const int ks = unit % p.split; unit /= p.split;
  const int tc = unit % p.tiles_c; unit /= p.tiles_c;
  const int tn = unit % p.tiles_n; unit /= p.tiles_n;
  const int tg = unit;
  const int t0 = tg * p.G;
  const int g_cnt = min(p.G, p.T - t0);
  const int per = (p.ksteps_total + p.split - 1) / p.split;
  const int k_begin = ks * per;
  const int k_end = min(p.ksteps_total, k_begin + per);
  const int n_iters = max(0, k_end - k_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmA);
    for (uint32_t s = 0; s < p.n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (n_iters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        uint32_t stage = 0, phase = 0;
        const uint32_t bytes = p.boxes_n * p.sub_n_bytes + g_cnt * p.tap_bytes;
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], bytes);
          uint8_t* sn = smem + static_cast<size_t>(stage) * p.stage_bytes;
          uint8_t* sc = sn + p.a_stage_off;
          const int q0 = (k_begin + it) * kWgBK;
          for (int b = 0; b < p.boxes_n; ++b)
            tma_load_2d(&tmDy, &full_bar[stage], sn + b * p.sub_n_bytes, tn * 128 + b * p.cw_n, q0);
          for (int j = 0; j < g_cnt; ++j) {
            const int row = q0 + p.shift[t0 + j];
            for (int b = 0; b < p.boxes_c; ++b)
              tma_load_2d(&tmA, &full_bar[stage], sc + j * p.tap_bytes + b * p.sub_c_bytes, tc * p.BNc + b * p.cw_c, row);
          }
          if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc_bf16(128, p.BNc, 1, 1);
        uint32_t stage = 0, phase = 0;
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sn = smem_u32(smem + static_cast<size_t>(stage) * p.stage_bytes);
          const uint32_t sc = sn + p.a_stage_off;
#pragma unroll
          for (int k = 0; k < kWgBK / 16; ++k) {
            // 16 pixel rows per MMA = two 8-row groups
            const uint64_t ad = make_smem_desc(sn + k * 2 * p.sbo_n, p.lbo_n, p.sbo_n, p.swz_n);
            for (int j = 0; j < g_cnt; ++j) {
              const uint64_t bd = make_smem_desc(sc + j * p.tap_bytes + k * 2 * p.sbo_c, p.lbo_c, p.sbo_c, p.swz_c);
              umma_bf16(tmem_base + j * p.BNc, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_bar);
      }
    } else {
      const int quad = warp & 3;
      const int row = quad * 32 + lane;
      const int n = tn * 128 + row;
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      for (int j = 0; j < g_cnt; ++j) {
        float* dst_row = p.dw + (static_cast<int64_t>(p.tap_index[t0 + j]) * p.N_store + n) * p.C_store;
        for (int c16 = 0; c16 < p.BNc / 16; ++c16) {
          uint32_t v[16];
          tmem_ld16(t_addr + j * p.BNc + c16 * 16, v);
          tmem_ld_wait();
          const int c0 = tc * p.BNc + c16 * 16;
          if (n < p.N_store) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (c0 + i < p.C_store) atomicAdd(dst_row + c0 + i, __uint_as_float(v[i]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace mmh

struct MmhWgradPlan {
  CUtensorMap tmDy, tmA;
  mmh::WgradKParams kp;
  int grid;
  size_t smem;
  MmhWgrad2* v2 = nullptr;   // generation-2 plan (default); MMH_WGRAD_IMPL=1 selects the first-generation kernel
  ~MmhWgradPlan() { if (v2) mmh_wgrad2_destroy(v2); }
};

using namespace mmh;

static int pick_cw(int ch) { return (ch % 64) == 0 ? 64 : ((ch % 32) == 0 ? 32 : 16); }
static CUtensorMapSwizzle swz_of(int cw) {
  return cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static int wgrad_fail(MmhWgradPlan* plan) {
  delete plan;
  return 1;
}

extern "C" int mmh_wgrad_plan_create(const MmhWgradDesc* d, MmhWgradPlan** out_plan) {
  MMH_CHECK(d && out_plan, "null argument");
  MMH_CHECK(d->T >= 1 && d->T <= MMH_MAX_TAPS, "T=%d out of range", d->T);
  MMH_CHECK(d->C >= 16 && (d->C % 16) == 0 && d->N >= 16 && (d->N % 16) == 0, "C=%d / N=%d must be multiples of 16",
            d->C, d->N);
  MMH_CHECK(d->C <= 256 || (d->C % 256) == 0, "C=%d unsupported (must be <=256 or a multiple of 256)", d->C);
  MMH_CHECK((d->a_ld % 8) == 0 && (d->dy_ld % 8) == 0, "leading dimensions must be multiples of 8");
  MMH_CHECK(d->M > 0 && d->M < (int64_t(1) << 31) - 256, "M out of range");
  auto* plan = new MmhWgradPlan();
  {
    const char* impl = getenv("MMH_WGRAD_IMPL");
    if (impl == nullptr || atoi(impl) != 1) {
      if (mmh_wgrad2_create(d, &plan->v2)) { delete plan; return 1; }
      *out_plan = plan;
      return 0;
    }
  }
  WgradKParams& k = plan->kp;
  memset(&k, 0, sizeof(k));
  k.T = d->T;
  k.cw_n = pick_cw(d->N);
  k.cw_c = pick_cw(d->C);
  k.tiles_n = (d->N + 127) / 128;
  k.BNc = d->C <= 256 ? d->C : 256;
  k.tiles_c = d->C / k.BNc;
  // taps per CTA: as many accumulators as fit in kWgMaxCols TMEM columns, balanced over the groups
  int gmax = kWgMaxCols / k.BNc;
  if (gmax < 1) gmax = 1;
  if (gmax > d->T) gmax = d->T;
  k.n_tg = (d->T + gmax - 1) / gmax;
  k.G = (d->T + k.n_tg - 1) / k.n_tg;
  k.boxes_n = 128 / k.cw_n;
  k.boxes_c = k.BNc / k.cw_c;
  k.ksteps_total = static_cast<int32_t>((d->M + kWgBK - 1) / kWgBK);
  k.N_store = d->N_store > 0 ? d->N_store : d->N;
  k.C_store = d->C_store > 0 ? d->C_store : d->C;
  k.dw_taps = d->dw_taps > 0 ? d->dw_taps : d->T;
  k.sub_n_bytes = kWgBK * k.cw_n * 2;
  k.sub_c_bytes = kWgBK * k.cw_c * 2;
  k.tap_bytes = k.boxes_c * k.sub_c_bytes;
  k.a_stage_off = k.boxes_n * k.sub_n_bytes;  // 16 KB
  k.stage_bytes = k.a_stage_off + k.G * k.tap_bytes;
  k.stage_bytes = (k.stage_bytes + 1023u) & ~1023u;
  const uint32_t budget = 227 * 1024 - 1024 - 256;
  k.n_stages = budget / k.stage_bytes;
  if (k.n_stages > kWgMaxStages) k.n_stages = kWgMaxStages;
  if (k.n_stages < 2) {
    set_error("wgrad tile does not fit in shared memory");
    return wgrad_fail(plan);
  }
  k.swz_n = k.cw_n == 64 ? 2u : (k.cw_n == 32 ? 4u : 6u);
  k.swz_c = k.cw_c == 64 ? 2u : (k.cw_c == 32 ? 4u : 6u);
  k.lbo_n = k.sub_n_bytes; k.sbo_n = 8 * k.cw_n * 2;
  k.lbo_c = k.sub_c_bytes; k.sbo_c = 8 * k.cw_c * 2;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(k.G * k.BNc)) cols <<= 1;
  k.tmem_cols = cols;
  k.dw = d->dw;
  for (int t = 0; t < d->T; ++t) {
    k.shift[t] = d->shift[t];
    k.tap_index[t] = d->tap_index[t];
    if (d->tap_index[t] < 0 || d->tap_index[t] >= k.dw_taps) {
      set_error("tap_index[%d] out of range", t);
      return wgrad_fail(plan);
    }
  }
  const int units = k.n_tg * k.tiles_n * k.tiles_c;
  int split = d->split_k;
  if (split <= 0) {
    const int sms = num_sms();
    split = sms / units;
    const int max_split = k.ksteps_total / 8 > 0 ? k.ksteps_total / 8 : 1;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
  }
  k.split = split;
  plan->grid = units * split;
  plan->smem = static_cast<size_t>(k.n_stages) * k.stage_bytes + 1024 + 256;
  if (make_tmap_2d_bf16(&plan->tmDy, d->dy, d->N, d->M, d->dy_ld, k.cw_n, kWgBK, swz_of(k.cw_n)) ||
      make_tmap_2d_bf16(&plan->tmA, d->a, d->C, d->a_rows, d->a_ld, k.cw_c, kWgBK, swz_of(k.cw_c))) {
    return wgrad_fail(plan);
  }
  cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(wgrad_kernel): %s", cudaGetErrorString(e));
    return wgrad_fail(plan);
  }
  *out_plan = plan;
  return 0;
}

extern "C" int mmh_wgrad_plan_destroy(MmhWgradPlan* plan) {
  delete plan;
  return 0;
}

extern "C" int mmh_wgrad_run(const MmhWgradPlan* plan, void* stream) {
  MMH_CHECK(plan, "null plan");
  if (plan->v2) return mmh_wgrad2_run(plan->v2, stream);
  wgrad_kernel<<<plan->grid, kWgThreads, plan->smem, static_cast<cudaStream_t>(stream)>>>(plan->tmDy, plan->tmA,
                                                                                         plan->kp);
  MMH_CUDA(cudaGetLastError());
  return 0;
}
