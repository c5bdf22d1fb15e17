// HOST EMULATION (test infrastructure only, never loaded by the product path).
// Straight-loop restatement of the conv / wgrad *contracts* of include/mmhand_sm100.h so that the host
// logic (layouts, tap tables, engine sequencing, elementwise index math) can be exercised on a CPU-only
// box. Built into libmmhand_hostemu.so by tests/hostemu.py with -DMMH_HOST_EMU.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/mmhand_sm100.h"
#include "ew_common.h"
#include "host_common.h"

using mmh::act_t;
struct MmhConvPlan { MmhConvDesc d; };
struct MmhWgradPlan { MmhWgradDesc d; };

extern "C" int mmh_conv_plan_create(const MmhConvDesc* d, MmhConvPlan** plan) {
  MMH_CHECK(d && plan, "null argument");
  MMH_CHECK(d->T >= 1 && d->T <= MMH_MAX_TAPS, "T=%d out of range", d->T);
  MMH_CHECK(d->C == 16 || d->C == 32 || d->C == 48 || (d->C % 64) == 0, "C=%d unsupported", d->C);
  MMH_CHECK(d->N >= 16 && (d->N % 16) == 0, "N=%d must be a multiple of 16", d->N);
  MMH_CHECK(d->N <= 256 || (d->N % 128) == 0, "N=%d unsupported", d->N);
  MMH_CHECK((d->a_ld % 8) == 0, "a_ld=%d invalid", d->a_ld);
  MMH_CHECK((d->out_ld % 8) == 0, "out_ld=%d must be a multiple of 8", d->out_ld);
  *plan = new MmhConvPlan{*d};
  return 0;
}
extern "C" int mmh_conv_plan_destroy(MmhConvPlan* p) { delete p; return 0; }

static inline uint32_t emu_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
static inline int emu_reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

static int conv_run(const MmhConvDesc& d, uint32_t drop_key);
extern "C" int mmh_conv_run(const MmhConvPlan* plan, void*) { return conv_run(plan->d, plan->d.bs_drop_key); }
extern "C" int mmh_conv_run_key(const MmhConvPlan* plan, uint32_t drop_key, void*) { return conv_run(plan->d, drop_key); }

static int conv_run(const MmhConvDesc& d, uint32_t drop_key) {
  const act_t* a = static_cast<const act_t*>(d.a);
  const act_t* w = static_cast<const act_t*>(d.w);
  const int w_taps = d.w_taps > 0 ? d.w_taps : d.T;
  (void)w_taps;
  const int64_t hw = static_cast<int64_t>(d.Hg) * d.Wg;
  const int n_store = d.n_store > 0 ? ((d.n_store + 15) / 16 * 16) : d.N;
  std::vector<float> acc(d.N);
  std::vector<float> arow(d.C);
  std::vector<double> bn_acc(d.bn_sums != nullptr ? 2 * d.bn_C : 0, 0.0);
  std::vector<double> bs_acc(d.bs_x != nullptr ? 2 * d.bs_C : 0, 0.0);
  for (int64_t q = 0; q < d.M; ++q) {
    const int64_t img = q / hw, rem = q % hw;
    const int h = static_cast<int>(rem / d.Wg), x = static_cast<int>(rem % d.Wg);
    const bool valid = h < d.Hv && x < d.Wv;
    if (!valid && !d.zero_invalid) continue;
    for (int n = 0; n < d.N; ++n) acc[n] = 0.f;
    if (valid) {
      for (int t = 0; t < d.T; ++t) {
        const int64_t r = q + d.shift[t];
        if (r < 0 || r >= d.a_rows) continue;
        const act_t* ar = a + r * d.a_ld;
        for (int c = 0; c < d.C; ++c) arow[c] = mmh::act2f(ar[c]);
        const int slot = d.w_taps > 0 ? d.w_slot[t] : t;
        const act_t* wt = w + static_cast<int64_t>(slot) * d.N * d.C;
        for (int n = 0; n < d.N; ++n) {
          const act_t* wr = wt + static_cast<int64_t>(n) * d.C;
          float s = 0.f;
          for (int c = 0; c < d.C; ++c) s += arow[c] * mmh::act2f(wr[c]);
          acc[n] += s;
        }
      }
      for (int n = 0; n < d.N; ++n) {
        float v = acc[n];
        if (d.bias) v += d.bias[n];
        if (d.act == 1) v = v > 0.f ? v : 0.f;
        else if (d.act == 2) v = tanhf(v);
        acc[n] = v;
      }
    }
    const int64_t orow = img * d.out_img_rows + static_cast<int64_t>(h * d.out_sh + d.out_h0) * d.out_wg +
                         (x * d.out_sw + d.out_w0);
    if (d.out_f32) {
      float* o = static_cast<float*>(d.out) + orow * d.out_ld;
      for (int n = 0; n < n_store && n < d.N; ++n) o[n] = acc[n];
    } else {
      act_t* o = static_cast<act_t*>(d.out) + orow * d.out_ld;
      if (d.bs_x != nullptr && valid) {
        // fused BN-backward masks + statistics (MmhConvDesc.bs_*): the row is the gradient of a mirrored logical pixel
        const int hs = emu_reflect(h - d.bs_pad, d.bs_H), ws = emu_reflect(x - d.bs_pad, d.bs_W);
        const act_t* xr = static_cast<const act_t*>(d.bs_x) +
                          ((img * d.bs_xHg + hs) * d.bs_xWg + ws) * static_cast<int64_t>(d.bs_x_ld);
        const uint32_t G = static_cast<uint32_t>((d.bs_C + 7) / 8);
        const uint32_t word0 = ((static_cast<uint32_t>(img) * d.bs_H + hs) * d.bs_W + ws) * G;
        for (int n = 0; n < d.N; ++n) {
          float v = 0.f;
          if (n < d.bs_C) {
            const float xv = mmh::act2f(xr[n]);
            const float a = d.bs_coef[n], b = d.bs_coef[d.bs_C + n];
            bool on = !d.bs_relu || (a * xv + b > 0.f);
            float keep = 1.f;
            if (d.bs_dropout) {
              const uint32_t bits = emu_mix32((word0 + static_cast<uint32_t>(n >> 3)) * 0x9E3779B1u + drop_key);
              on = on && ((bits >> (n & 7)) & 1u);
              keep = 2.f;
            }
            v = on ? keep * acc[n] : 0.f;
            const double st = mmh::act2f(mmh::f2act(v));
            bs_acc[n] += st;
            bs_acc[d.bs_C + n] += st * ((xv - d.bs_save[n]) * d.bs_save[d.bs_C + n]);
          }
          acc[n] = v;
        }
      }
      for (int n = 0; n < n_store && n < d.N; ++n) o[n] = mmh::f2act(acc[n]);
      if (d.bn_sums != nullptr && valid) {          // fused BN statistics of the values as stored
        for (int n = 0; n < d.bn_C && n < n_store; ++n) {
          const double v = mmh::act2f(mmh::f2act(acc[n]));
          bn_acc[n] += v;
          bn_acc[d.bn_C + n] += v * v;
        }
      }
    }
  }
  if (d.bn_sums != nullptr)
    for (int n = 0; n < 2 * d.bn_C; ++n) d.bn_sums[n] += static_cast<float>(bn_acc[n]);
  if (d.bs_x != nullptr)
    for (int n = 0; n < 2 * d.bs_C; ++n) d.bs_sums[n] += static_cast<float>(bs_acc[n]);
  return 0;
}

extern "C" int mmh_wgrad_plan_create(const MmhWgradDesc* d, MmhWgradPlan** plan) {
  MMH_CHECK(d && plan, "null argument");
  MMH_CHECK(d->T >= 1 && d->T <= MMH_MAX_TAPS, "T=%d out of range", d->T);
  MMH_CHECK((d->C % 16) == 0 && (d->N % 16) == 0, "C=%d / N=%d must be multiples of 16", d->C, d->N);
  MMH_CHECK(d->C <= 256 || (d->C % 256) == 0, "C=%d unsupported", d->C);
  *plan = new MmhWgradPlan{*d};
  return 0;
}
extern "C" int mmh_wgrad_plan_destroy(MmhWgradPlan* p) { delete p; return 0; }

extern "C" int mmh_wgrad_run(const MmhWgradPlan* plan, void*) {
  const MmhWgradDesc& d = plan->d;
  const act_t* a = static_cast<const act_t*>(d.a);
  const act_t* dy = static_cast<const act_t*>(d.dy);
  const int Ns = d.N_store > 0 ? d.N_store : d.N;
  const int Cs = d.C_store > 0 ? d.C_store : d.C;
  std::vector<float> dyr(Ns), ar(Cs);
  for (int t = 0; t < d.T; ++t) {
    float* dw = d.dw + static_cast<int64_t>(d.tap_index[t]) * Ns * Cs;
    for (int64_t q = 0; q < d.M; ++q) {
      const int64_t r = q + d.shift[t];
      if (r < 0 || r >= d.a_rows) continue;
      bool any = false;
      for (int n = 0; n < Ns; ++n) { dyr[n] = mmh::act2f(dy[q * d.dy_ld + n]); any |= dyr[n] != 0.f; }
      if (!any) continue;
      for (int c = 0; c < Cs; ++c) ar[c] = mmh::act2f(a[r * d.a_ld + c]);
      for (int n = 0; n < Ns; ++n) {
        const float g = dyr[n];
        if (g == 0.f) continue;
        float* row = dw + static_cast<int64_t>(n) * Cs;
        for (int c = 0; c < Cs; ++c) row[c] += g * ar[c];
      }
    }
  }
  return 0;
}
