// Weight gradient on tcgen05, generation 2: activation windows in shared memory shared by the taps.
//
//   dw[t][n][c] += sum_q dy[q][n] * a[q + shift[t]][c]
//
// The pixel index q is the contraction. Per pipeline stage (64 pixel rows) a CTA loads ONE dy tile and, per
// tap group, ONE activation window of 64 + span rows; every tap of the group is an MMA whose activation
// descriptor starts `rel` rows into the window (same descriptor arithmetic as tc_conv2.cu).
//
// mode 0 (wide layers):  M = 128 dy channels (256 for a cta_group::2 pair), N = BNc activation channels,
//                        one accumulator per tap of the group (G * BNc <= 512 TMEM columns).
//                        L2 -> SM bytes per MMA flop drop 2-3x against generation 1 (one tap per CTA).
// mode 1 (few-channel k x k stems, "taps on M"): the activation window is the M operand with
//                        M index = (kw, c) -- a MN-major descriptor whose leading-dimension stride is ONE ROW, so
//                        that "next group of channels" means "next pixel" -- N = dy channels, one accumulator per
//                        (kernel row, block of 128/cw column taps). 7 MMAs of 128 x 64 x 16 replace 49 of
//                        64 x 16 x 16 per 16 pixel rows.
// mode 2 (same layers, "kernel rows on M"): M index = (kh, c) -- the leading-dimension stride of the MN-major window
//                        descriptor is ONE WINDOW (+ one row of skew), so that the eight 16-channel chunks of an MMA
//                        come from eight different windows and land in different banks; the column taps kw are the
//                        accumulators (7 x N columns), their descriptors start kw rows into the windows. Mode 1's
//                        chunks overlap (one row apart): its operand fetch reads 4 KB out of a 736-byte region and
//                        the MMAs run at a fifth of their rate (profiles/r02_stem_wgrad_ablation.txt).
// fp32 partial sums are reduced into dw with red.global.add (split-K across CTAs).
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

constexpr int kW2Threads = 192;
constexpr int kW2BK = 64;
constexpr int kW2MaxStages = 8;
constexpr int kW2MaxGroups = 8;

struct Wg2Params {
  int32_t mode, ncta;
  int32_t n_groups;                 // groups handled by one CTA (mode 0: 1; mode 1: all kernel rows of the unit)
  int32_t units_g, tiles_n, tiles_c, split;
  int32_t ksteps_total;
  // dy operand
  int32_t dy_boxes, dy_cw, dy_cols;           // boxes per CTA per stage, channels per box, dy channels per CTA
  uint32_t dy_box_bytes, dy_swz, dy_sbo;
  // window operand
  int32_t w_boxes, w_cw, w_cols, w_rows;      // boxes per window per CTA, channels per box, channels per CTA, rows
  uint32_t w_box_bytes, w_bytes, w_swz, w_sbo, w_row_bytes;
  uint32_t win_off, stage_bytes, n_stages, tmem_cols;
  int32_t BNc;                                // instruction N (mode 0: activation channels; mode 1: dy channels)
  int32_t n_acc;                              // accumulators per CTA
  int32_t N_store, C_store, dw_taps;
  float* dw;
  // groups (global list; a unit's groups are [ug * n_groups, ...))
  int32_t g_first[kW2MaxGroups * 4 + 1];
  int32_t g_min[kW2MaxGroups * 4];
  int32_t rel[MMH_MAX_TAPS];
  int32_t tap_index[MMH_MAX_TAPS];
  // mode 1: accumulator a -> (group, first column tap); kw_per_mma = 128 / w_cw
  int32_t kw_per_mma;
  // mode 2: kernel rows per MMA (128 / w_cw), M blocks, window slots per stage, leading-dimension stride of the window
  // descriptor (window pitch + skew), rows of skew per kernel row
  int32_t kh_per_mma, n_mblocks, w_slots, skew_rows;
  uint32_t w_lbo;
  int32_t dbg;   // MMH_W2_DEBUG: 1 skip epilogue reductions, 2 skip MMAs, 4 skip TMA loads (timing experiments only)
};

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int NCTA>
__global__ void __launch_bounds__(kW2Threads, 1)
wgrad2_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmA,
              const __grid_constant__ Wg2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.n_stages) * p.stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kW2MaxStages;
  uint64_t* acc_bar = bars + 2 * kW2MaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kW2MaxStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = NCTA == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta == 0;

  // unit decode: split fastest (CTAs that share a dw tile run at different times of the K range)
  int unit = blockIdx.x / NCTA;
  const int ks = unit % p.split; unit /= p.split;
  const int tc = unit % p.tiles_c; unit /= p.tiles_c;
  const int tn = unit % p.tiles_n; unit /= p.tiles_n;
  const int ug = unit;                               // group unit
  const int g0 = ug * p.n_groups;
  const int per = (p.ksteps_total + p.split - 1) / p.split;
  const int k_begin = ks * per;
  const int k_end = min(p.ksteps_total, k_begin + per);
  const int n_iters = max(0, k_end - k_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmA);
    for (uint32_t s = 0; s < p.n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (NCTA == 2) { tmem_alloc2(tmem_slot, p.tmem_cols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: barriers, TMEM and descriptors above were set up while the preceding kernel of the
  // stream was still running; its results are visible past this point, and the next kernel may start its own prologue
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // channel origins of this CTA's operand slices
  const int dy_c0 = (p.mode == 0 ? tn * p.dy_cols * NCTA + static_cast<int>(cta) * p.dy_cols : 0);
  const int a_c0 = tc * p.w_cols * (p.mode == 0 ? NCTA : 1) + (p.mode == 0 ? static_cast<int>(cta) * p.w_cols : 0);

  if (n_iters > 0) {
    if (warp == 0) {
      if (elect_one()) {
        uint32_t stage = 0, phase = 0;
        const uint32_t bytes = p.dy_boxes * p.dy_box_bytes + p.n_groups * p.w_boxes * p.w_box_bytes;
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sd = smem + static_cast<size_t>(stage) * p.stage_bytes;
          uint8_t* sw = sd + p.win_off;
          const int q0 = (k_begin + it) * kW2BK;
          if (p.dbg & 4) {
            if (leader) mbar_arrive(&full_bar[stage]);
          } else if (NCTA == 2) {
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * bytes);
            for (int b = 0; b < p.dy_boxes; ++b)
              tma_load_2d_pair(&tmDy, bar, sd + b * p.dy_box_bytes, dy_c0 + b * p.dy_cw, q0);
            for (int g = 0; g < p.n_groups; ++g)
              for (int b = 0; b < p.w_boxes; ++b)
                tma_load_2d_pair(&tmA, bar, sw + g * p.w_bytes + b * p.w_box_bytes, a_c0 + b * p.w_cw,
                                 q0 + p.g_min[g0 + g] - g * p.skew_rows);
          } else {
            mbar_expect_tx(&full_bar[stage], bytes);
            for (int b = 0; b < p.dy_boxes; ++b)
              tma_load_2d(&tmDy, &full_bar[stage], sd + b * p.dy_box_bytes, dy_c0 + b * p.dy_cw, q0);
            for (int g = 0; g < p.n_groups; ++g)
              for (int b = 0; b < p.w_boxes; ++b)
                tma_load_2d(&tmA, &full_bar[stage], sw + g * p.w_bytes + b * p.w_box_bytes, a_c0 + b * p.w_cw,
                            q0 + p.g_min[g0 + g] - g * p.skew_rows);
          }
          if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (leader && elect_one()) {
        // descriptor low words are built by adding pre-encoded (>> 4) offsets: two uniform adds per MMA
        uint32_t stage = 0, phase = 0;
        const uint64_t dy_proto = make_smem_desc(0, p.dy_box_bytes, p.dy_sbo, p.dy_swz);
        const uint32_t dy_hi = static_cast<uint32_t>(dy_proto >> 32), dy_lo = static_cast<uint32_t>(dy_proto);
        const uint32_t smem16 = smem_u32(smem) >> 4, stage16 = p.stage_bytes >> 4, win16 = p.win_off >> 4;
        const uint32_t dy_k16 = (2 * p.dy_sbo) >> 4, w_k16 = (2 * p.w_sbo) >> 4, row16 = p.w_row_bytes >> 4;
        const uint32_t n_stages = p.n_stages;
        const int BNc = p.BNc;
        const bool skip = (p.dbg & 2) != 0;
        if (p.mode == 0) {
          const uint32_t idesc = make_idesc_bf16(128 * NCTA, p.BNc, 1, 1);
          const uint64_t w_proto = make_smem_desc(0, p.w_box_bytes, p.w_sbo, p.w_swz);
          const uint32_t w_hi = static_cast<uint32_t>(w_proto >> 32), w_lo = static_cast<uint32_t>(w_proto);
          const int t_begin = p.g_first[g0], ntaps = p.g_first[g0 + 1] - t_begin;
          uint32_t rel16[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) rel16[j] = j < ntaps ? static_cast<uint32_t>(p.rel[t_begin + j]) * row16 : 0u;
          for (int it = 0; it < n_iters; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sd = smem16 + stage * stage16 + dy_lo;
            const uint32_t sw = smem16 + stage * stage16 + win16 + w_lo;
            if (!skip) {
#pragma unroll
              for (int k = 0; k < kW2BK / 16; ++k) {
                const uint64_t ad = (static_cast<uint64_t>(dy_hi) << 32) | (sd + k * dy_k16);
                const uint32_t accf = (it | k) != 0 ? 1u : 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (j < ntaps) {
                    const uint64_t bd = (static_cast<uint64_t>(w_hi) << 32) | (sw + rel16[j] + k * w_k16);
                    if (NCTA == 2) umma_bf16_pair(tmem_base + j * BNc, ad, bd, idesc, accf);
                    else umma_bf16(tmem_base + j * BNc, ad, bd, idesc, accf);
                  }
                }
              }
            }
            if (NCTA == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        } else if (p.mode == 2) {
          // kernel rows on M: A = the windows of the unit's kernel rows (MN-major, leading-dimension stride = one
          // window + skew), B = dy; accumulator (M block, kw)
          const uint32_t idesc = make_idesc_bf16(128, p.BNc, 1, 1);
          const uint64_t w_proto = make_smem_desc(0, p.w_lbo, p.w_sbo, p.w_swz);
          const uint32_t w_hi = static_cast<uint32_t>(w_proto >> 32), w_lo = static_cast<uint32_t>(w_proto);
          const uint32_t mb16 = (static_cast<uint32_t>(p.kh_per_mma) * p.w_lbo) >> 4;
          const int n_mb = p.n_mblocks, ntaps = p.g_first[g0 + 1] - p.g_first[g0];
          for (int it = 0; it < n_iters; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sd = smem16 + stage * stage16 + dy_lo;
            const uint32_t sw = smem16 + stage * stage16 + win16 + w_lo;
            if (!skip) {
#pragma unroll
              for (int k = 0; k < kW2BK / 16; ++k) {
                const uint64_t bd = (static_cast<uint64_t>(dy_hi) << 32) | (sd + k * dy_k16);
                const uint32_t accf = (it | k) != 0 ? 1u : 0u;
                uint32_t d = tmem_base;
                for (int mb = 0; mb < n_mb; ++mb) {
                  uint32_t wa = sw + mb * mb16 + k * w_k16;
                  for (int kw = 0; kw < ntaps; ++kw, wa += row16, d += BNc) {
                    const uint64_t ad = (static_cast<uint64_t>(w_hi) << 32) | wa;
                    umma_bf16(d, ad, bd, idesc, accf);
                  }
                }
              }
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        } else {
          // taps on M: A = window (MN-major, leading-dimension stride = one row), B = dy
          const uint32_t idesc = make_idesc_bf16(128, p.BNc, 1, 1);
          const uint64_t w_proto = make_smem_desc(0, p.w_row_bytes, p.w_sbo, p.w_swz);
          const uint32_t w_hi = static_cast<uint32_t>(w_proto >> 32), w_lo = static_cast<uint32_t>(w_proto);
          const uint32_t wbytes16 = p.w_bytes >> 4, step16 = static_cast<uint32_t>(p.kw_per_mma) * row16;
          const int n_groups = p.n_groups;
          // accumulators per group (all groups of the stems have the same tap count)
          const int apg = (p.g_first[g0 + 1] - p.g_first[g0] + p.kw_per_mma - 1) / p.kw_per_mma;
          for (int it = 0; it < n_iters; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sd = smem16 + stage * stage16 + dy_lo;
            const uint32_t sw = smem16 + stage * stage16 + win16 + w_lo;
            if (!skip) {
#pragma unroll
              for (int k = 0; k < kW2BK / 16; ++k) {
                const uint64_t bd = (static_cast<uint64_t>(dy_hi) << 32) | (sd + k * dy_k16);
                const uint32_t accf = (it | k) != 0 ? 1u : 0u;
                uint32_t d = tmem_base;
                for (int g = 0; g < n_groups; ++g) {
                  uint32_t wa = sw + g * wbytes16 + k * w_k16;
                  for (int a = 0; a < apg; ++a, wa += step16, d += BNc) {
                    const uint64_t ad = (static_cast<uint64_t>(w_hi) << 32) | wa;
                    umma_bf16(d, ad, bd, idesc, accf);
                  }
                }
              }
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        }
        if (NCTA == 2) umma_commit_pair(acc_bar); else umma_commit(acc_bar);
      }
    } else {
      const int quad = warp & 3;
      const int row = quad * 32 + lane;
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      if (p.mode == 0) {
        const int n = tn * 128 * NCTA + static_cast<int>(cta) * 128 + row;
        const int t_begin = p.g_first[g0], ntaps = p.g_first[g0 + 1] - t_begin;
        const bool vec = (p.C_store & 3) == 0;
        for (int j = 0; j < ntaps; ++j) {
          float* dst_row = p.dw + (static_cast<int64_t>(p.tap_index[t_begin + j]) * p.N_store + n) * p.C_store;
          for (int c16 = 0; c16 < p.BNc / 16; ++c16) {
            uint32_t v[16];
            tmem_ld16(t_addr + j * p.BNc + c16 * 16, v);
            tmem_ld_wait();
            const int c0 = tc * p.BNc + c16 * 16;
            if (n < p.N_store && !(p.dbg & 1)) {
              if (vec && c0 + 16 <= p.C_store) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                  red_add_v4(dst_row + c0 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                             __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (c0 + i < p.C_store) red_add_f32(dst_row + c0 + i, __uint_as_float(v[i]));
              }
            }
          }
        }
      } else if (p.mode == 2) {
        // lane = (kh_local, c); accumulator (M block, kw) holds dw[(kh, kw)][n][c] in column n
        const int khl = row / p.w_cw, c = tc * p.w_cw + row % p.w_cw;
        const int ntaps = p.g_first[g0 + 1] - p.g_first[g0];
        int acc = 0;
        for (int mb = 0; mb < p.n_mblocks; ++mb) {
          const int kh = mb * p.kh_per_mma + khl;
          const bool ok = kh < p.n_groups && c < p.C_store;
          for (int kw = 0; kw < ntaps; ++kw, ++acc) {
            const int tap = ok ? p.tap_index[p.g_first[g0 + kh] + kw] : 0;
            float* dst = p.dw + static_cast<int64_t>(tap) * p.N_store * p.C_store + c;
            for (int n16 = 0; n16 < p.BNc / 16; ++n16) {
              uint32_t v[16];
              tmem_ld16(t_addr + acc * p.BNc + n16 * 16, v);
              tmem_ld_wait();
              if (ok && !(p.dbg & 1)) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int n = n16 * 16 + i;
                  if (n < p.N_store) red_add_f32(dst + static_cast<int64_t>(n) * p.C_store, __uint_as_float(v[i]));
                }
              }
            }
          }
        }
      } else {
        // lane = (kw_local, c): kw_local = row / w_cw, c = row % w_cw
        const int kwl = row / p.w_cw, c = tc * p.w_cw + row % p.w_cw;
        int acc = 0;
        for (int g = 0; g < p.n_groups; ++g) {
          const int t_begin = p.g_first[g0 + g], ntaps = p.g_first[g0 + g + 1] - t_begin;
          for (int j = 0; j < ntaps; j += p.kw_per_mma, ++acc) {
            const bool ok = (j + kwl) < ntaps && c < p.C_store;
            const int tap = ok ? p.tap_index[t_begin + j + kwl] : 0;
            float* dst = p.dw + static_cast<int64_t>(tap) * p.N_store * p.C_store + c;
            for (int n16 = 0; n16 < p.BNc / 16; ++n16) {
              uint32_t v[16];
              tmem_ld16(t_addr + acc * p.BNc + n16 * 16, v);
              tmem_ld_wait();
              if (ok) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int n = n16 * 16 + i;
                  if (n < p.N_store) red_add_f32(dst + static_cast<int64_t>(n) * p.C_store, __uint_as_float(v[i]));
                }
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc2(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace mmh

using namespace mmh;

struct MmhWgrad2 {
  CUtensorMap tmDy, tmA;
  Wg2Params kp;
  int grid, ncta;
  size_t smem;
};

static int w2_env(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}
static int w2_cw(int ch) { return (ch % 64) == 0 ? 64 : ((ch % 32) == 0 ? 32 : 16); }
static CUtensorMapSwizzle w2_swz(int cw) {
  return cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
static uint32_t w2_swz_code(int cw) { return cw == 64 ? 2u : (cw == 32 ? 4u : 6u); }

int mmh_wgrad2_create(const MmhWgradDesc* d, MmhWgrad2** out_plan) {
  auto* plan = new MmhWgrad2();
  Wg2Params& k = plan->kp;
  memset(&k, 0, sizeof(k));
  auto fail = [&]() { delete plan; return 1; };
  k.N_store = d->N_store > 0 ? d->N_store : d->N;
  k.C_store = d->C_store > 0 ? d->C_store : d->C;
  k.dw_taps = d->dw_taps > 0 ? d->dw_taps : d->T;
  k.dw = d->dw;
  k.dbg = w2_env("MMH_W2_DEBUG", 0);
  k.ksteps_total = static_cast<int32_t>((d->M + kW2BK - 1) / kW2BK);

  // ---- tap groups (sorted by shift; a gap of >= 64 rows starts a new group)
  std::vector<std::pair<int, int>> taps;
  for (int t = 0; t < d->T; ++t) {
    if (d->tap_index[t] < 0 || d->tap_index[t] >= k.dw_taps) { set_error("tap_index[%d] out of range", t); return fail(); }
    taps.emplace_back(d->shift[t], d->tap_index[t]);
  }
  std::stable_sort(taps.begin(), taps.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) {
    return a.first < b.first;
  });
  // mode: taps on M for few-channel stems with many taps whose groups are runs of consecutive shifts
  int mode = w2_env("MMH_WGRAD_MODE", -1);
  const bool stem_like = d->T >= 16 && d->N <= 64 && (d->C <= 64);
  // measured (profiles/r02_stem_wgrad_modes.txt): 16-channel chunks 389 -> 778 / 403 -> 873 TFLOP/s with the kernel
  // rows on M; the 64-channel chunk of the output layer (2 kernel rows per MMA, N = 16) stays with the taps on M
  if (mode < 0) mode = stem_like ? ((d->C % 64) == 0 ? 1 : 2) : 0;
  if (mode != 0 && !stem_like) mode = 0;
  k.mode = mode;

  int gmax_taps;   // taps per group limit
  if (mode == 0) {
    k.BNc = d->C <= 128 ? d->C : 128;
    if ((d->C % k.BNc) != 0) { set_error("C=%d unsupported by wgrad", d->C); return fail(); }
    gmax_taps = 512 / k.BNc;
    if (gmax_taps > 4) gmax_taps = 4;
  } else {
    gmax_taps = 8;
  }
  std::vector<int> gf, gm;
  for (int t = 0; t < d->T; ++t) {
    const bool fresh = t == 0 || taps[t].first - taps[t - 1].first >= 64 || (t - gf.back()) >= gmax_taps ||
                       (mode != 0 && taps[t].first - taps[t - 1].first != 1);
    if (fresh) { gf.push_back(t); gm.push_back(taps[t].first); }
    k.rel[t] = taps[t].first - gm.back();
    k.tap_index[t] = taps[t].second;
  }
  const int ng = static_cast<int>(gf.size());
  if (ng > kW2MaxGroups * 4) { set_error("too many tap groups"); return fail(); }
  int span_max = 0, gt_max = 0;
  for (int g = 0; g < ng; ++g) {
    k.g_first[g] = gf[g];
    k.g_min[g] = gm[g];
    const int end = g + 1 < ng ? gf[g + 1] : d->T;
    gt_max = std::max(gt_max, end - gf[g]);
    span_max = std::max(span_max, k.rel[end - 1]);
  }
  k.g_first[ng] = d->T;
  if (mode != 0) {
    for (int g = 0; g < ng; ++g)
      if (k.g_first[g + 1] - k.g_first[g] != gt_max) { set_error("wgrad taps-on-M mode needs equal kernel rows"); return fail(); }
  }

  int ncta = 1;
  if (mode == 0) {
    ncta = w2_env("MMH_WGRAD_NCTA", 0);
    if (ncta == 0) ncta = (d->N % 256 == 0 && k.BNc == 128) ? 2 : 1;
    if (ncta == 2 && ((d->N % 256) != 0 || (k.BNc % 32) != 0)) ncta = 1;
    k.n_groups = 1;
    k.units_g = ng;
    k.tiles_n = (d->N + 128 * ncta - 1) / (128 * ncta);
    k.tiles_c = d->C / k.BNc;
    k.dy_cols = 128;
    k.dy_cw = w2_cw(d->N);
    // N < 128: the tile still spans 128 dy channels, TMA zero-fills the columns past N
    k.dy_boxes = k.dy_cols / k.dy_cw;
    k.w_cols = k.BNc / ncta;
    k.w_cw = w2_cw(k.w_cols);
    k.w_boxes = k.w_cols / k.w_cw;
    k.w_rows = (kW2BK + span_max + 7) / 8 * 8;
    k.n_acc = gt_max;
    k.kw_per_mma = 1;
  } else if (mode == 2) {
    // mode 2: one unit = one channel chunk of width cw (16, or 64 when C is a multiple of 64) x ALL kernel rows
    k.w_cw = (d->C % 64) == 0 ? 64 : 16;
    k.kh_per_mma = 128 / k.w_cw;
    k.n_mblocks = (ng + k.kh_per_mma - 1) / k.kh_per_mma;
    k.w_slots = k.n_mblocks * k.kh_per_mma;
    // 32-byte chunks of one K row spread over the banks; 128-byte ones fill them (MMH_WGRAD_SKEW: timing experiments)
    k.skew_rows = w2_env("MMH_WGRAD_SKEW", k.w_cw == 16 ? 1 : 0);
    k.kw_per_mma = 1;
    k.BNc = d->N;
    if ((d->N % 16) != 0 || d->N > 256) { set_error("wgrad mode 2: N=%d unsupported", d->N); return fail(); }
    if (ng > kW2MaxGroups || k.n_mblocks * gt_max * k.BNc > 512) { set_error("wgrad mode 2: accumulators do not fit"); return fail(); }
    k.n_groups = ng;
    k.units_g = 1;
    k.tiles_n = 1;
    k.tiles_c = d->C / k.w_cw;
    k.dy_cols = d->N;
    k.dy_cw = w2_cw(d->N);
    k.dy_boxes = d->N / k.dy_cw;
    k.w_cols = k.w_cw;
    k.w_boxes = 1;
    // window rows: 64 pixel rows + column taps + the skew of the last kernel row, rounded up to 8
    k.w_rows = (kW2BK + (gt_max - 1) + k.skew_rows * (ng - 1) + 7) / 8 * 8;
    k.n_acc = k.n_mblocks * gt_max;
  } else {
    // mode 1: one unit = one channel chunk of width cw (16, or 64 when C is a multiple of 64) x all groups
    k.w_cw = (d->C % 64) == 0 ? 64 : 16;
    k.kw_per_mma = 128 / k.w_cw;
    k.BNc = d->N;
    if ((d->N % 16) != 0 || d->N > 256) { set_error("wgrad mode 1: N=%d unsupported", d->N); return fail(); }
    int acc_per_group = (gt_max + k.kw_per_mma - 1) / k.kw_per_mma;
    int groups_per_unit = 512 / (acc_per_group * k.BNc);
    if (groups_per_unit < 1) { set_error("wgrad mode 1: accumulators do not fit"); return fail(); }
    if (groups_per_unit > kW2MaxGroups) groups_per_unit = kW2MaxGroups;
    if (groups_per_unit > ng) groups_per_unit = ng;
    // all units must own the same number of groups: use a divisor of ng, else one group per unit
    while (ng % groups_per_unit) --groups_per_unit;
    k.n_groups = groups_per_unit;
    k.units_g = ng / groups_per_unit;
    k.tiles_n = 1;
    k.tiles_c = d->C / k.w_cw;
    k.dy_cols = d->N;
    k.dy_cw = w2_cw(d->N);
    k.dy_boxes = d->N / k.dy_cw;
    k.w_cols = k.w_cw;
    k.w_boxes = 1;
    // window rows: 64 + (kw blocks * kw_per_mma) - 1 rounded up to 8 (the last MMA of a group may run past the taps)
    const int kw_cover = acc_per_group * k.kw_per_mma;
    k.w_rows = (kW2BK + kw_cover - 1 + 7) / 8 * 8;
    k.n_acc = groups_per_unit * acc_per_group;
  }
  plan->ncta = ncta;
  k.ncta = ncta;
  k.dy_box_bytes = kW2BK * k.dy_cw * 2;
  k.dy_swz = w2_swz_code(k.dy_cw);
  k.dy_sbo = 8 * k.dy_cw * 2;
  k.w_row_bytes = k.w_cw * 2;
  k.w_box_bytes = k.w_rows * k.w_row_bytes;
  k.w_bytes = (k.w_boxes * k.w_box_bytes + 1023u) & ~1023u;
  k.w_swz = w2_swz_code(k.w_cw);
  k.w_sbo = 8 * k.w_row_bytes;
  k.win_off = (k.dy_boxes * k.dy_box_bytes + 1023u) & ~1023u;
  k.w_lbo = k.w_bytes + static_cast<uint32_t>(k.skew_rows) * k.w_row_bytes;
  if (mode != 2) k.w_slots = k.n_groups;
  k.stage_bytes = k.win_off + k.w_slots * k.w_bytes;
  k.stage_bytes = (k.stage_bytes + 1023u) & ~1023u;
  // The weight gradient runs on a side stream next to the bandwidth-bound BN-backward kernels (engine.py::run_bwd):
  // leave room in shared memory for two of their blocks (2 x (16 KB + 1 KB)) so that they can be co-resident.
  static const uint32_t cap_kb = [] {
    const char* e = getenv("MMH_WGRAD_SMEM_KB");
    const int v = e != nullptr ? atoi(e) : 190;
    return static_cast<uint32_t>(v < 64 ? 64 : (v > 227 ? 227 : v));
  }();
  const uint32_t budget = cap_kb * 1024 - 1024 - 256;
  k.n_stages = budget / k.stage_bytes;
  if (k.n_stages > kW2MaxStages) k.n_stages = kW2MaxStages;
  if (k.n_stages < 2) { set_error("wgrad tile does not fit in shared memory"); return fail(); }
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(k.n_acc * k.BNc)) cols <<= 1;
  if (cols > 512) { set_error("wgrad accumulators exceed TMEM"); return fail(); }
  k.tmem_cols = cols;

  const int units = k.units_g * k.tiles_n * k.tiles_c;
  int split = d->split_k;
  if (split <= 0) {
    // CTAs per launch = `waves` x the SM count: with several short waves a lower-priority side-stream launch hands
    // the SMs over to a main-stream convolution at the next CTA boundary instead of holding them to the end
    static const int waves = [] {
      const char* e = getenv("MMH_WGRAD_WAVES");
      const int v = e != nullptr ? atoi(e) : 1;
      return v < 1 ? 1 : (v > 8 ? 8 : v);
    }();
    const int slots = waves * (num_sms() / ncta);
    split = slots / units;
    const int max_split = k.ksteps_total / 8 > 0 ? k.ksteps_total / 8 : 1;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
  }
  k.split = split;
  plan->grid = units * split * ncta;
  plan->smem = static_cast<size_t>(k.n_stages) * k.stage_bytes + 1024 + 256;
  if (make_tmap_2d_bf16(&plan->tmDy, d->dy, d->N, d->M, d->dy_ld, k.dy_cw, kW2BK, w2_swz(k.dy_cw)) ||
      make_tmap_2d_bf16(&plan->tmA, d->a, d->C, d->a_rows, d->a_ld, k.w_cw, k.w_rows, w2_swz(k.w_cw)))
    return fail();
  cudaError_t e = ncta == 2 ? cudaFuncSetAttribute(wgrad2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                            : cudaFuncSetAttribute(wgrad2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(wgrad2_kernel): %s", cudaGetErrorString(e)); return fail(); }
  *out_plan = plan;
  return 0;
}

void mmh_wgrad2_destroy(MmhWgrad2* plan) { delete plan; }

int mmh_wgrad2_run(const MmhWgrad2* plan, void* stream) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(plan->grid, 1, 1);
  cfg.blockDim = dim3(kW2Threads, 1, 1);
  cfg.dynamicSmemBytes = plan->smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  const bool pdl = pdl_enabled();
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (plan->ncta == 2) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
    MMH_CUDA(cudaLaunchKernelEx(&cfg, wgrad2_kernel<2>, plan->tmDy, plan->tmA, plan->kp));
  } else {
    MMH_CUDA(cudaLaunchKernelEx(&cfg, wgrad2_kernel<1>, plan->tmDy, plan->tmA, plan->kp));
  }
  return 0;
}
