// Bandwidth-bound kernels between the convolutions: input assembly, BatchNorm statistics / apply /
// backward, ReLU, dropout, reflect / zero halos, the PAT gate and its backward, gradient gathering.
// Every kernel moves 16-byte bf16 vectors (8 channels) per work item on NHWC pixel grids and is bounded by
// HBM bandwidth; the algorithmic bytes per element are listed in DESIGN.md section 4.3.
// Dual-mode source: CUDA by default, host loops with -DMMH_HOST_EMU (CPU tests of the index arithmetic).
#include <math.h>

#include "bn_finalize.h"
#include "ew_framework.h"
#include "peer.cuh"

namespace mmh {

MMH_HD float sigmoidf_(float x) {
#if defined(__CUDA_ARCH__)
  return __fdividef(1.0f, 1.0f + __expf(-x));      // ex2.approx + rcp.approx: ~2 ulp, outputs are stored as bf16
#else
  return 1.0f / (1.0f + expf(-x));
#endif
}

struct NoCtx {};

// ------------------------------------------------------------------------------------------------ assemble
struct AssembleF {
  static constexpr int kUnroll = 2;
  struct Ctx { float sc[8], sh[8]; };
  struct In { float v[8]; };
  const float* src0; const float* src1; const float* scale; const float* shift;
  act_t* dst; LayD dl;
  int C0, C1, reflect;
  MMH_HD void prep(int g, Ctx& c) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = g * 8 + j;
      const bool on = scale != nullptr && ch < C0 + C1;
      c.sc[j] = on ? scale[ch] : 1.f;
      c.sh[j] = on ? shift[ch] : 0.f;
    }
  }
  MMH_HD void load(int b, int h, int w, int g, const Ctx&, In& in) const {
    const bool inside = h >= 0 && h < dl.H && w >= 0 && w < dl.W;
    zero8(in.v);
    if (inside || reflect) {
      const int hs = reflect_idx(h, dl.H), ws = reflect_idx(w, dl.W);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = g * 8 + j;
        if (ch < C0) in.v[j] = src0[((static_cast<int64_t>(b) * C0 + ch) * dl.H + hs) * dl.W + ws];
        else if (ch < C0 + C1) in.v[j] = src1[((static_cast<int64_t>(b) * C1 + (ch - C0)) * dl.H + hs) * dl.W + ws];
      }
    }
  }
  MMH_HD void finish(const In& in, int b, int h, int w, int g, const Ctx& c) const {
    const bool live = reflect || (h >= 0 && h < dl.H && w >= 0 && w < dl.W);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (live && g * 8 + j < C0 + C1) ? in.v[j] * c.sc[j] + c.sh[j] : 0.f;
    st8_bf16(dst + lay_off(dl, b, h, w) + g * 8, v);
  }
};

// ------------------------------------------------------------------------------------------------ BN stats
struct BnStatsF {
  static constexpr int kUnroll = 8;       // one 16-byte load per item: eight in flight per thread
  typedef NoCtx Ctx;
  struct In { ActX8 x; };
  const act_t* x; int ld;
  MMH_HD void prep(int, Ctx&) const {}
  MMH_HD void load(int, int, int r, int g, const Ctx&, In& in) const {
    ld_raw(x + static_cast<int64_t>(r) * ld + g * 8, in.x);
  }
  MMH_HD void accum(const In& in, int, int, int, int, const Ctx&, float (&acc)[2][8]) const {
    float v[8];
    cvt8(in.x, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] += v[j]; acc[1][j] += v[j] * v[j]; }
  }
};

// ------------------------------------------------------------------------------------------------ norm + act + pad
struct NormActF {
  static constexpr int kUnroll = 4;       // (2: -10 %, 8: -8 % on the step's shapes, profiles/r02_norm_act_knobs.txt)
  struct Ctx { float a[8], b[8]; };
  struct In { ActX8 x; F32x8 r; };
  const act_t* src; LayD sl; const float* coef; int relu, dropout; uint32_t key;
  const float* resid; act_t* dst; LayD dl; int reflect; float* dst_f32;
  MMH_HD void prep(int g, Ctx& c) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c.a[j] = coef != nullptr ? coef[g * 8 + j] : 1.f;
      c.b[j] = coef != nullptr ? coef[sl.C + g * 8 + j] : 0.f;
    }
  }
  MMH_HD bool live(int h, int w) const { return reflect || (h >= 0 && h < sl.H && w >= 0 && w < sl.W); }
  MMH_HD void load(int b, int h, int w, int g, const Ctx&, In& in) const {
    if (live(h, w)) {
      const int hs = reflect_idx(h, sl.H), ws = reflect_idx(w, sl.W);
      ld_raw(src + lay_off(sl, b, hs, ws) + g * 8, in.x);
      if (resid != nullptr) ld_raw(resid + static_cast<int64_t>((b * sl.H + hs) * sl.W + ws) * sl.C + g * 8, in.r);
    }
  }
  MMH_HD void finish(const In& in, int b, int h, int w, int g, const Ctx& c) const {
    const int H = sl.H, W = sl.W, C = sl.C;
    const bool inside = h >= 0 && h < H && w >= 0 && w < W;
    float y[8];
    zero8(y);
    if (live(h, w)) {
      const int hs = reflect_idx(h, H), ws = reflect_idx(w, W);
      cvt8(in.x, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = c.a[j] * y[j] + c.b[j];
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = y[j] > 0.f ? y[j] : 0.f;
      }
      if (dropout) {
        const uint32_t bits = drop_bits(key, b, hs, ws, g, C, H, W);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = ((bits >> j) & 1u) ? 2.0f * y[j] : 0.f;
      }
      if (resid != nullptr) {
        float r[8];
        cvt8(in.r, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] += r[j];
      }
    }
    if (dst != nullptr) st8_bf16(dst + lay_off(dl, b, h, w) + g * 8, y);
    if (inside && dst_f32 != nullptr) st8_f32(dst_f32 + static_cast<int64_t>((b * H + h) * W + w) * C + g * 8, y);
  }
};

// element offsets of the even- / odd-column pixels of one (image, row) of a grid (parity-plane grids keep the two
// in different planes; plain grids: e == o); 32-bit: the launchers check rows * pitch < 2^31
struct RowOff { int32_t e, o; };
MMH_HD void rowoff(const LayD& l, int b, int h, RowOff& r) {
  const int hp = h + l.h0;
  if (!l.phase) {
    r.e = ((b * l.Hg + hp) * l.Wg + l.w0) * l.ld + l.c0;
    r.o = r.e;
  } else {
    r.e = (((hp & 1) * 2) * l.plane_rows + (b * l.Hg + (hp >> 1)) * l.Wg) * l.ld + l.c0;
    r.o = r.e + l.plane_rows * l.ld;
  }
}
MMH_HD int32_t rowat(const LayD& l, const RowOff& r, int w) {
  if (!l.phase) return r.e + w * l.ld;
  const int wp = w + l.w0;
  return ((wp & 1) ? r.o : r.e) + (wp >> 1) * l.ld;
}
inline bool lay_fits32(const MmhLay& l) {
  return static_cast<int64_t>(l.B) * l.Hg * l.Wg * (l.phase ? 4 : 1) * l.ld + l.c0 < (int64_t(1) << 31);
}

// ------------------------------------------------------------------------------------------------ PAT gate
struct GateFwdF {
  static constexpr int kUnroll = 2;
  struct Ctx { float a[8], b[8]; };
  struct In { ActX8 c1, x2, x3; F32x8 t; };
  const act_t* c1; const act_t* x2o; const act_t* x3o; LayD sl; const float* coef;
  const float* trunk_in; float* trunk_out;
  act_t* d1; LayD d1l; act_t* d2; LayD d2l; act_t* d3; LayD d3l;
  int reflect;
  MMH_HD void prep(int g, Ctx& c) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) { c.a[j] = coef[g * 8 + j]; c.b[j] = coef[sl.C + g * 8 + j]; }
  }
  MMH_HD bool live(int h, int w) const { return reflect || (h >= 0 && h < sl.H && w >= 0 && w < sl.W); }
  MMH_HD void load(int b, int h, int w, int g, const Ctx&, In& in) const {
    if (live(h, w)) {
      const int hs = reflect_idx(h, sl.H), ws = reflect_idx(w, sl.W);
      const int64_t so = lay_off(sl, b, hs, ws) + g * 8;
      ld_raw(c1 + so, in.c1);
      ld_raw(x2o + so, in.x2);
      if (x3o != nullptr) ld_raw(x3o + so, in.x3);
      ld_raw(trunk_in + static_cast<int64_t>((b * sl.H + hs) * sl.W + ws) * sl.C + g * 8, in.t);
    }
  }
  MMH_HD void finish(const In& in, int b, int h, int w, int g, const Ctx& c) const {
    const int H = sl.H, W = sl.W, C = sl.C;
    const bool inside = h >= 0 && h < H && w >= 0 && w < W;
    float out[8], a2[8], a3[8];
    zero8(out); zero8(a2); zero8(a3);
    if (live(h, w)) {
      float v1[8], t[8];
      cvt8(in.c1, v1);
      cvt8(in.x2, a2);
      cvt8(in.t, t);
      if (x3o != nullptr) {
        cvt8(in.x3, a3);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float bn = c.a[j] * v1[j] + c.b[j];
          out[j] = t[j] + bn * sigmoidf_(a2[j]) * sigmoidf_(a3[j]);
        }
      } else {              // two-stream PATBlock (model_variants.py:58-68): one attention map
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = t[j] + (c.a[j] * v1[j] + c.b[j]) * sigmoidf_(a2[j]);
      }
    }
    st8_bf16(d1 + lay_off(d1l, b, h, w) + g * 8, out);
    if (d2 != nullptr) {
      const int64_t o = lay_off(d2l, b, h, w) + g * 8;
      st8_bf16(d2 + o, a3);
      st8_bf16(d2 + o + C, out);
    }
    if (d3 != nullptr) {
      const int64_t o = lay_off(d3l, b, h, w) + g * 8;
      st8_bf16(d3 + o, a2);
      st8_bf16(d3 + o + C, out);
    }
    if (inside) st8_f32(trunk_out + static_cast<int64_t>((b * H + h) * W + w) * C + g * 8, out);
  }
};

// ------------------------------------------------------------------------------------------------ halo folding
struct GradSrcD {
  const act_t* p; LayD l; int lo, hi, reflect;
};
inline GradSrcD to_gsd(const MmhGradSrc& s) {
  GradSrcD d;
  d.p = static_cast<const act_t*>(s.p);
  d.l = to_layd(s.l);
  d.lo = s.pad_lo; d.hi = s.pad_hi; d.reflect = s.reflect;
  return d;
}
// halo coordinates whose mirror image is i (reflect padding), besides i itself
MMH_HD int preimages(int i, int n, int lo, int hi, int reflect, int (&out)[3]) {
  int k = 0;
  out[k++] = i;
  if (reflect) {
    if (i >= 1 && i <= lo) out[k++] = -i;
    if (i <= n - 2 && i >= n - 1 - hi) out[k++] = 2 * (n - 1) - i;
  }
  return k;
}
MMH_HD void fold_add8(const GradSrcD& s, int b, int h, int w, int cg, float (&acc)[8]) {
  int hs[3], ws[3];
  const int nh = preimages(h, s.l.H, s.lo, s.hi, s.reflect, hs);
  const int nw = preimages(w, s.l.W, s.lo, s.hi, s.reflect, ws);
  for (int a = 0; a < nh; ++a)
    for (int c = 0; c < nw; ++c) {
      float v[8];
      ld8_bf16(s.p + lay_off(s.l, b, hs[a], ws[c]) + cg, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
}

struct GradGatherF {
  static constexpr int kUnroll = 2;
  typedef NoCtx Ctx;
  struct In { float acc[8]; ActX8 m; };
  int nsrc; GradSrcD src[4]; const float* trunk; const act_t* mask; LayD ml;
  void* dst; LayD dl; int dst_f32, H, W, C;
  MMH_HD void prep(int, Ctx&) const {}
  MMH_HD void load(int b, int h, int w, int g, const Ctx&, In& in) const {
    const int64_t plain = static_cast<int64_t>((b * H + h) * W + w) * C + g * 8;
    if (trunk != nullptr) ld8_f32(trunk + plain, in.acc); else zero8(in.acc);
    for (int s = 0; s < nsrc; ++s) fold_add8(src[s], b, h, w, g * 8, in.acc);
    if (mask != nullptr) ld_raw(mask + lay_off(ml, b, h, w) + g * 8, in.m);
  }
  MMH_HD void finish(const In& in, int b, int h, int w, int g, const Ctx&) const {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = in.acc[j];
    if (mask != nullptr) {
      float m[8];
      cvt8(in.m, m);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = m[j] > 0.f ? acc[j] : 0.f;
    }
    const int64_t o = lay_off(dl, b, h, w) + g * 8;
    if (dst_f32) st8_f32(static_cast<float*>(dst) + o, acc);
    else st8_bf16(static_cast<act_t*>(dst) + o, acc);
  }
};

// ------------------------------------------------------------------------------------------------ BN backward
// dy = a*(dz_eff - k0 - xhat*k1) with xhat = (x - mean)*rstd is evaluated as  a*dz_eff + bx*x + cc
//
// NS = 0: dz is a materialised plain buffer (fp32 or bf16). NS = 1, 2: dz is gathered on the fly from NS consumer
// data gradients (+ an optional fp32 plain term): the primary element of every source is loaded in the load phase
// (one 16-byte load each, all in flight together); the mirrored halo elements that fold onto border pixels are
// added in the finish phase by the few threads that own a border pixel. Nothing is written or re-read in between.
struct BnDzSlot { union { F32x8 f; ActX8 h; }; };
struct BnNoSlot {};
template <bool ON> struct BnSlotSel { typedef BnDzSlot type; };
template <> struct BnSlotSel<false> { typedef BnNoSlot type; };
// TR: an fp32 plain term is part of the gathered gradient (NS > 0 only); compile-time so that the common
// single-source case keeps 8 registers of loads per item in flight instead of 16
template <int NS, bool TR>
struct BnBwdIn { typename BnSlotSel<(NS == 0) || TR>::type dz; ActX8 x; ActX8 s[NS ? NS : 1]; };

MMH_HD bool on_fold_border(int i, int n, int lo, int hi) { return (i >= 1 && i <= lo) || (i <= n - 2 && i >= n - 1 - hi); }
// halo elements of source s that mirror onto (h, w), excluding (h, w) itself
MMH_HD void fold_extra8(const GradSrcD& s, int b, int h, int w, int cg, float (&acc)[8]) {
  int hs[3], ws[3];
  const int nh = preimages(h, s.l.H, s.lo, s.hi, s.reflect, hs);
  const int nw = preimages(w, s.l.W, s.lo, s.hi, s.reflect, ws);
  for (int a = 0; a < nh; ++a)
    for (int c = 0; c < nw; ++c) {
      if (a == 0 && c == 0) continue;
      float v[8];
      ld8_bf16(s.p + lay_off(s.l, b, hs[a], ws[c]) + cg, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
}

template <int NS, bool TR>
struct BnBwdBase {
  typedef BnBwdIn<NS, TR> In;
  const void* dz; int dz_f32, relu, dropout; uint32_t key;
  const act_t* x; LayD xl; const float* coef; const float* save;
  GradSrcD src[NS ? NS : 1]; const float* trunk;
  MMH_HD void load(int b, int h, int w, int g, In& in) const {
    const int64_t plain = static_cast<int64_t>((b * xl.H + h) * xl.W + w) * xl.C + g * 8;
    if constexpr (NS == 0) {
      if (dz_f32) ld_raw(static_cast<const float*>(dz) + plain, in.dz.f);
      else ld_raw(static_cast<const act_t*>(dz) + plain, in.dz.h);
    } else {
      if constexpr (TR) ld_raw(trunk + plain, in.dz.f);
#pragma unroll
      for (int s = 0; s < NS; ++s) ld_raw(src[s].p + lay_off(src[s].l, b, h, w) + g * 8, in.s[s]);
    }
    ld_raw(x + lay_off(xl, b, h, w) + g * 8, in.x);
  }
  // effective upstream gradient (after the recomputed ReLU / dropout masks) and raw activation of one vector
  MMH_HD void eff(const In& in, int b, int h, int w, int g, const float (&a)[8], const float (&bb)[8],
                  float (&dze)[8], float (&xv)[8]) const {
    if constexpr (NS == 0) {
      if (dz_f32) cvt8(in.dz.f, dze); else cvt8(in.dz.h, dze);
    } else {
      if constexpr (TR) cvt8(in.dz.f, dze); else zero8(dze);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        float v[8];
        cvt8(in.s[s], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) dze[j] += v[j];
        if (src[s].reflect && (on_fold_border(h, src[s].l.H, src[s].lo, src[s].hi) ||
                               on_fold_border(w, src[s].l.W, src[s].lo, src[s].hi)))
          fold_extra8(src[s], b, h, w, g * 8, dze);
      }
    }
    cvt8(in.x, xv);
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!(a[j] * xv[j] + bb[j] > 0.f)) dze[j] = 0.f;
    }
    if (dropout) {
      const uint32_t bits = drop_bits(key, b, h, w, g, xl.C, xl.H, xl.W);
#pragma unroll
      for (int j = 0; j < 8; ++j) dze[j] = ((bits >> j) & 1u) ? 2.0f * dze[j] : 0.f;
    }
  }
};
template <int NS, bool TR>
struct BnBwdReduceF {
  static constexpr int kUnroll = (NS * 4 + (TR ? 8 : 0)) > 8 ? 2 : 4;
  struct Ctx { float a[8], b[8], mean[8], rstd[8]; };
  typedef BnBwdIn<NS, TR> In;
  BnBwdBase<NS, TR> cm;
  MMH_HD void prep(int g, Ctx& c) const {
    const int C = cm.xl.C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c.a[j] = cm.coef[g * 8 + j]; c.b[j] = cm.coef[C + g * 8 + j];
      c.mean[j] = cm.save[g * 8 + j]; c.rstd[j] = cm.save[C + g * 8 + j];
    }
  }
  MMH_HD void load(int b, int h, int w, int g, const Ctx&, In& in) const { cm.load(b, h, w, g, in); }
  MMH_HD void accum(const In& in, int b, int h, int w, int g, const Ctx& c, float (&acc)[2][8]) const {
    float dze[8], xv[8];
    cm.eff(in, b, h, w, g, c.a, c.b, dze, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[0][j] += dze[j];
      acc[1][j] += dze[j] * ((xv[j] - c.mean[j]) * c.rstd[j]);
    }
  }
};
template <int NS, bool TR>
struct BnBwdApplyF {
  static constexpr int kUnroll = (NS * 4 + (TR ? 8 : 0)) > 8 ? 2 : 4;
  struct Ctx { float a[8], b[8], bx[8], cc[8]; };
  typedef BnBwdIn<NS, TR> In;
  BnBwdBase<NS, TR> cm; const float* k; act_t* dy; LayD yl;
  MMH_HD void prep(int g, Ctx& c) const {
    const int C = cm.xl.C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = g * 8 + j;
      const float a = cm.coef[ch], mean = cm.save[ch], rstd = cm.save[C + ch];
      c.a[j] = a; c.b[j] = cm.coef[C + ch];
      c.bx[j] = -a * k[C + ch] * rstd;
      c.cc[j] = -a * k[ch] + a * k[C + ch] * rstd * mean;
    }
  }
  MMH_HD void load(int b, int h, int w, int g, const Ctx&, In& in) const { cm.load(b, h, w, g, in); }
  MMH_HD void finish(const In& in, int b, int h, int w, int g, const Ctx& c) const {
    float dze[8], xv[8], o[8];
    cm.eff(in, b, h, w, g, c.a, c.b, dze, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = c.a[j] * dze[j] + c.bx[j] * xv[j] + c.cc[j];
    st8_bf16(dy + lay_off(yl, b, h, w) + g * 8, o);
  }
};
// ------------------------------------------------------------------------------------------------ BN backward, lean
// The hot instance of the two kernels above -- ONE gathered source, no fp32 trunk term: every BatchNorm of the
// generator and the discriminators except the two that join the residual trunk -- as row kernels (ew_framework.h):
// base offsets per (image, row), ~3x fewer instructions per 16-byte vector than the general functors (whose per-item
// layout arithmetic, source loops and 4-byte parameter loads made them issue-bound at 1.7-2.9 TB/s), ReLU and dropout
// resolved at compile time. The reduction accumulates sum dz*(x - mean) and scales by rstd once per thread.
struct BnLeanRow { RowOff x, s, y; uint32_t dword; int b, h, hb; };
struct BnLeanIn { ActX8 x, s; };

template <bool RELU, bool DROP>
struct BnLeanBase {
  const act_t* x; LayD xl; GradSrcD src; const float* coef; const float* save; uint32_t key;
  MMH_HD void row(int b, int h, BnLeanRow& r) const {
    r.b = b; r.h = h;
    rowoff(xl, b, h, r.x);
    rowoff(src.l, b, h, r.s);
    r.y = r.x;
    r.hb = (src.reflect && on_fold_border(h, xl.H, src.lo, src.hi)) ? 1 : 0;
    r.dword = (static_cast<uint32_t>(b) * xl.H + h) * xl.W * static_cast<uint32_t>((xl.C + 7) / 8);
  }
  MMH_HD void load(const BnLeanRow& r, int w, int g, BnLeanIn& in) const {
    ld_raw(x + (rowat(xl, r.x, w) + g * 8), in.x);
    ld_raw(src.p + (rowat(src.l, r.s, w) + g * 8), in.s);
  }
  // effective upstream gradient (gathered, masked) and raw activation of one vector
  MMH_HD void eff(const BnLeanIn& in, const BnLeanRow& r, int w, int g, const float (&a)[8], const float (&bb)[8],
                  float (&dze)[8], float (&xv)[8]) const {
    cvt8(in.s, dze);
    if (src.reflect && (r.hb || on_fold_border(w, xl.W, src.lo, src.hi))) fold_extra8(src, r.b, r.h, w, g * 8, dze);
    cvt8(in.x, xv);
    if (RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!(a[j] * xv[j] + bb[j] > 0.f)) dze[j] = 0.f;
    }
    if (DROP) {
      const uint32_t word = r.dword + static_cast<uint32_t>(w) * static_cast<uint32_t>((xl.C + 7) / 8) + g;
      const uint32_t bits = mix32(word * 0x9E3779B1u + key);
#pragma unroll
      for (int j = 0; j < 8; ++j) dze[j] = ((bits >> j) & 1u) ? 2.0f * dze[j] : 0.f;
    }
  }
};
template <bool RELU, bool DROP>
struct BnLeanReduceF {
  static constexpr int kUnroll = 4;
  struct Ctx { float a[8], b[8], mean[8]; };
  typedef BnLeanIn In;
  typedef BnLeanRow Row;
  BnLeanBase<RELU, DROP> cm;
  MMH_HD void prep(int g, Ctx& c) const {
    const int C = cm.xl.C;
    if (RELU) { ld8_f32(cm.coef + g * 8, c.a); ld8_f32(cm.coef + C + g * 8, c.b); }
    else { zero8(c.a); zero8(c.b); }
    ld8_f32(cm.save + g * 8, c.mean);
  }
  MMH_HD void row(int b, int h, Row& r) const { cm.row(b, h, r); }
  MMH_HD void load(const Row& r, int w, int g, const Ctx&, In& in) const { cm.load(r, w, g, in); }
  MMH_HD void accum(const In& in, const Row& r, int w, int g, const Ctx& c, float (&acc)[2][8]) const {
    float dze[8], xv[8];
    cm.eff(in, r, w, g, c.a, c.b, dze, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[0][j] += dze[j];
      acc[1][j] += dze[j] * (xv[j] - c.mean[j]);
    }
  }
  MMH_HD void post(int g, const Ctx&, float (&acc)[2][8]) const {
    float rstd[8];
    ld8_f32(cm.save + cm.xl.C + g * 8, rstd);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[1][j] *= rstd[j];
  }
};
template <bool RELU, bool DROP>
struct BnLeanApplyF {
  static constexpr int kUnroll = 4;
  struct Ctx { float a[8], b[8], bx[8], cc[8]; };
  typedef BnLeanIn In;
  typedef BnLeanRow Row;
  BnLeanBase<RELU, DROP> cm; const float* k; act_t* dy; LayD yl;
  MMH_HD void prep(int g, Ctx& c) const {
    const int C = cm.xl.C;
    float mean[8], rstd[8], k0[8], k1[8];
    ld8_f32(cm.coef + g * 8, c.a); ld8_f32(cm.coef + C + g * 8, c.b);
    ld8_f32(cm.save + g * 8, mean); ld8_f32(cm.save + C + g * 8, rstd);
    ld8_f32(k + g * 8, k0); ld8_f32(k + C + g * 8, k1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c.bx[j] = -c.a[j] * k1[j] * rstd[j];
      c.cc[j] = -c.a[j] * k0[j] + c.a[j] * k1[j] * rstd[j] * mean[j];
    }
  }
  MMH_HD void row(int b, int h, Row& r) const {
    cm.row(b, h, r);
    rowoff(yl, b, h, r.y);
  }
  MMH_HD void load(const Row& r, int w, int g, const Ctx&, In& in) const { cm.load(r, w, g, in); }
  MMH_HD void finish(const In& in, const Row& r, int w, int g, const Ctx& c) const {
    float dze[8], xv[8], o[8];
    cm.eff(in, r, w, g, c.a, c.b, dze, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = c.a[j] * dze[j] + c.bx[j] * xv[j] + c.cc[j];
    st8_bf16(dy + (rowat(yl, r.y, w) + g * 8), o);
  }
};
// ------------------------------------------------------------------------------------------------ gate backward
struct GateBwdIn { F32x8 dv; ActX8 c1, x2, x3, e2, e3; };
struct GateBwdBase {
  const float* dout; const act_t* c1; const act_t* x2o; const act_t* x3o; LayD sl;
  const float* coef; const float* save;
  MMH_HD void load(int b, int h, int w, int g, GateBwdIn& in) const {
    ld_raw(dout + static_cast<int64_t>((b * sl.H + h) * sl.W + w) * sl.C + g * 8, in.dv);
    const int64_t so = lay_off(sl, b, h, w) + g * 8;
    ld_raw(c1 + so, in.c1);
    ld_raw(x2o + so, in.x2);
    if (x3o != nullptr) ld_raw(x3o + so, in.x3);
  }
  MMH_HD void eff(const GateBwdIn& in, const float (&a)[8], const float (&bb)[8], float (&d1)[8], float (&v1)[8],
                  float (&d2)[8], float (&d3)[8]) const {
    float dv[8], v2[8], v3[8];
    cvt8(in.dv, dv);
    cvt8(in.c1, v1);
    cvt8(in.x2, v2);
    if (x3o != nullptr) cvt8(in.x3, v3); else zero8(v3);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s2 = sigmoidf_(v2[j]), s3 = x3o != nullptr ? sigmoidf_(v3[j]) : 1.f;      // (s3 = 1: d3 = 0)
      const float bn = a[j] * v1[j] + bb[j];
      d1[j] = dv[j] * s2 * s3;
      d2[j] = dv[j] * bn * s3 * s2 * (1.f - s2);
      d3[j] = dv[j] * bn * s2 * s3 * (1.f - s3);
    }
  }
};
struct GateBwdReduceF {
  static constexpr int kUnroll = 2;
  struct Ctx { float a[8], b[8], mean[8], rstd[8]; };
  typedef GateBwdIn In;
  GateBwdBase cm;
  MMH_HD void prep(int g, Ctx& c) const {
    const int C = cm.sl.C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c.a[j] = cm.coef[g * 8 + j]; c.b[j] = cm.coef[C + g * 8 + j];
      c.mean[j] = cm.save[g * 8 + j]; c.rstd[j] = cm.save[C + g * 8 + j];
    }
  }
  MMH_HD void load(int b, int h, int w, int g, const Ctx&, In& in) const { cm.load(b, h, w, g, in); }
  MMH_HD void accum(const In& in, int, int, int, int, const Ctx& c, float (&acc)[2][8]) const {
    float d1[8], v1[8], d2[8], d3[8];
    cm.eff(in, c.a, c.b, d1, v1, d2, d3);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[0][j] += d1[j];
      acc[1][j] += d1[j] * ((v1[j] - c.mean[j]) * c.rstd[j]);
    }
  }
};
struct GateBwdApplyF {
  static constexpr int kUnroll = 2;
  struct Ctx { float a[8], b[8], bx[8], cc[8]; };
  typedef GateBwdIn In;
  GateBwdBase cm; const float* k; GradSrcD ex2, ex3;
  act_t* dy1; act_t* dy2; act_t* dy3; LayD yl;
  MMH_HD void prep(int g, Ctx& c) const {
    const int C = cm.sl.C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = g * 8 + j;
      const float a = cm.coef[ch], mean = cm.save[ch], rstd = cm.save[C + ch];
      c.a[j] = a; c.b[j] = cm.coef[C + ch];
      c.bx[j] = -a * k[C + ch] * rstd;
      c.cc[j] = -a * k[ch] + a * k[C + ch] * rstd * mean;
    }
  }
  // extra gradients of x2o / x3o (the next block's stream convs read them through the swapped concatenation):
  // primary element in the load phase, mirrored halo elements by the border threads in the finish phase
  MMH_HD void load(int b, int h, int w, int g, const Ctx&, In& in) const {
    cm.load(b, h, w, g, in);
    if (ex2.p != nullptr) ld_raw(ex2.p + lay_off(ex2.l, b, h, w) + g * 8, in.e2);
    if (ex3.p != nullptr) ld_raw(ex3.p + lay_off(ex3.l, b, h, w) + g * 8, in.e3);
  }
  MMH_HD void add_extra(const GradSrcD& s, const ActX8& prim, int b, int h, int w, int g, float (&d)[8]) const {
    float v[8];
    cvt8(prim, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] += v[j];
    if (s.reflect && (on_fold_border(h, s.l.H, s.lo, s.hi) || on_fold_border(w, s.l.W, s.lo, s.hi)))
      fold_extra8(s, b, h, w, g * 8, d);
  }
  MMH_HD void finish(const In& in, int b, int h, int w, int g, const Ctx& c) const {
    float d1[8], v1[8], d2[8], d3[8], o[8];
    cm.eff(in, c.a, c.b, d1, v1, d2, d3);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = c.a[j] * d1[j] + c.bx[j] * v1[j] + c.cc[j];
    if (ex2.p != nullptr) add_extra(ex2, in.e2, b, h, w, g, d2);
    if (ex3.p != nullptr) add_extra(ex3, in.e3, b, h, w, g, d3);
    const int64_t off = lay_off(yl, b, h, w) + g * 8;
    st8_bf16(dy1 + off, o);
    st8_bf16(dy2 + off, d2);
    if (dy3 != nullptr) st8_bf16(dy3 + off, d3);
  }
};

}  // namespace mmh

// ==================================================================================================
using namespace mmh;

#define MMH_REQ_VEC(C) MMH_CHECK((C) > 0 && ((C) % 8) == 0, "channel count %d must be a multiple of 8", (int)(C))


extern "C" int mmh_assemble_nchw(const float* src0, int32_t C0, const float* src1, int32_t C1, const float* scale,
                                 const float* shift, void* dst, const MmhLay* dl, int32_t pad_lo, int32_t pad_hi,
                                 int32_t reflect, void* stream) {
  MMH_CHECK(src0 && dst && dl, "null argument");
  MMH_REQ_VEC(dl->C);
  MMH_CHECK(C0 + C1 <= dl->C, "C0+C1=%d exceeds the destination channels %d", C0 + C1, dl->C);
  MMH_CHECK(!reflect || (pad_lo < dl->H && pad_hi < dl->H && pad_lo < dl->W && pad_hi < dl->W), "halo too large");
  AssembleF f;
  f.src0 = src0; f.src1 = src1; f.scale = scale; f.shift = shift;
  f.dst = static_cast<act_t*>(dst); f.dl = to_layd(*dl);
  f.C0 = C0; f.C1 = src1 ? C1 : 0; f.reflect = reflect;
  return launch_pg(f, make_rowgeom(dl->B, dl->H, dl->W, pad_lo, pad_hi), dl->C / 8, stream);
}

extern "C" int mmh_bn_stats(const void* x, int64_t rows, int32_t ld, int32_t C, float* sums, void* stream) {
  MMH_CHECK(x && sums, "null argument");
  MMH_REQ_VEC(C);
  BnStatsF f;
  f.x = static_cast<const act_t*>(x); f.ld = ld;
  MMH_CHECK(rows < (int64_t(1) << 31), "too many rows");
  return launch_reduce_ch<2>(f, make_flatgeom(rows), C / 8, C, sums, stream);
}

extern "C" int mmh_bn_stats_finalize(MmhPeer* peer, uint32_t seq, const void* x, int64_t rows, int32_t ld, int32_t C,
                                     float* sums, uint32_t* counter, float count_global, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var, float momentum,
                                     float eps, float* coef, float* save, void* stream) {
  MMH_CHECK(x && sums && counter && coef && save, "null argument");
  MMH_REQ_VEC(C);
  MMH_CHECK(rows < (int64_t(1) << 31), "too many rows");
  BnStatsF f;
  f.x = static_cast<const act_t*>(x); f.ld = ld;
  BnFwdFin fin;
  if (peer_dev(peer, seq, 2 * C, &fin.px)) return 1;
  fin.sums = sums;
  fin.f.sums = sums; fin.f.gamma = gamma; fin.f.beta = beta; fin.f.rm = running_mean; fin.f.rv = running_var;
  fin.f.coef = coef; fin.f.save = save; fin.f.count = count_global; fin.f.momentum = momentum; fin.f.eps = eps;
  fin.f.train = 1; fin.f.C = C;
  return launch_reduce_ch_fin<2>(f, make_flatgeom(rows), C / 8, C, sums, fin, counter, stream);
}

#ifndef MMH_HOST_EMU
template <class Fin>
__global__ void __launch_bounds__(256) fin_reset_kernel(const Fin fin, float* sums, const int C) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  fin(c);
  sums[c] = 0.f;
  sums[C + c] = 0.f;
}
#endif

extern "C" int mmh_bn_finalize_reset(MmhPeer* peer, uint32_t seq, float* sums, float count_global, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var, float momentum,
                                     float eps, int32_t C, float* coef, float* save, void* stream) {
  MMH_CHECK(sums && coef && save && C > 0, "bad argument");
  BnFwdFin fin;
  if (peer_dev(peer, seq, 2 * C, &fin.px)) return 1;
  fin.sums = sums;
  fin.f.sums = sums; fin.f.gamma = gamma; fin.f.beta = beta; fin.f.rm = running_mean; fin.f.rv = running_var;
  fin.f.coef = coef; fin.f.save = save; fin.f.count = count_global; fin.f.momentum = momentum; fin.f.eps = eps;
  fin.f.train = 1; fin.f.C = C;
#ifdef MMH_HOST_EMU
  (void)stream;
  for (int c = 0; c < C; ++c) { fin(c); sums[c] = 0.f; sums[C + c] = 0.f; }
#else
  MMH_CUDA(launch_k(fin_reset_kernel<BnFwdFin>, dim3((C + 255) / 256), dim3(256), 0, stream, fin, sums, C));
#endif
  return 0;
}

extern "C" int mmh_bn_finalize(const float* sums, float count, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps, int32_t train,
                               int32_t C, float* coef, float* save, void* stream) {
  MMH_CHECK(coef && save, "null argument");
  MMH_CHECK(train ? sums != nullptr : (running_mean && running_var), "missing statistics");
  BnFinalizeF f;
  f.sums = sums; f.gamma = gamma; f.beta = beta; f.rm = running_mean; f.rv = running_var; f.coef = coef; f.save = save;
  f.count = count; f.momentum = momentum; f.eps = eps; f.train = train; f.C = C;
  return launch_map(f, C, stream);
}

static bool ew_lean_enabled() {       // MMH_EW_LEAN=0: the general kernels everywhere (read per call: tests switch)
  const char* e = getenv("MMH_EW_LEAN");
  return e == nullptr || atoi(e) != 0;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
extern "C" int mmh_norm_act(const MmhNormAct* p, void* stream) {
  MMH_CHECK(p && p->src && (p->dst || p->dst_f32), "null argument");
  MMH_REQ_VEC(p->sl.C);
  NormActF f;
  f.src = static_cast<const act_t*>(p->src); f.sl = to_layd(p->sl); f.coef = p->coef;
  f.relu = p->relu; f.dropout = p->dropout; f.key = p->drop_key; f.resid = p->resid;
  f.dst = static_cast<act_t*>(p->dst); f.dl = to_layd(p->dl);
  const int lo = p->dst ? p->pad_lo : 0, hi = p->dst ? p->pad_hi : 0;
  f.reflect = p->reflect; f.dst_f32 = p->dst_f32;
  MMH_CHECK(!p->dst || (p->dl.H == p->sl.H && p->dl.W == p->sl.W && p->dl.C == p->sl.C), "src/dst shape mismatch");
  MMH_CHECK(!f.reflect || (lo < f.sl.H && hi < f.sl.H && lo < f.sl.W && hi < f.sl.W), "halo too large");
  const RowGeom rg = make_rowgeom(f.sl.B, f.sl.H, f.sl.W, lo, hi);
  return launch_pg(f, rg, p->sl.C / 8, stream);
}

extern "C" int mmh_gate_fwd(const MmhGateFwd* p, void* stream) {
  MMH_CHECK(p && p->c1 && p->x2o && p->coef && p->trunk_in && p->trunk_out && p->d1, "null argument");
  MMH_REQ_VEC(p->sl.C);
  GateFwdF f;
  f.c1 = static_cast<const act_t*>(p->c1); f.x2o = static_cast<const act_t*>(p->x2o);
  f.x3o = static_cast<const act_t*>(p->x3o); f.sl = to_layd(p->sl); f.coef = p->coef;
  f.trunk_in = p->trunk_in; f.trunk_out = p->trunk_out;
  f.d1 = static_cast<act_t*>(p->d1); f.d1l = to_layd(p->d1l);
  f.d2 = static_cast<act_t*>(p->d2); f.d2l = to_layd(p->d2l);
  f.d3 = static_cast<act_t*>(p->d3); f.d3l = to_layd(p->d3l);
  f.reflect = p->reflect;
  return launch_pg(f, make_rowgeom(f.sl.B, f.sl.H, f.sl.W, p->pad_lo, p->pad_hi), p->sl.C / 8, stream);
}

extern "C" int mmh_grad_gather(const MmhGradGather* p, void* stream) {
  MMH_CHECK(p && p->dst && p->nsrc >= 0 && p->nsrc <= 4, "bad argument");
  MMH_REQ_VEC(p->C);
  GradGatherF f;
  f.nsrc = p->nsrc;
  for (int s = 0; s < p->nsrc; ++s) f.src[s] = to_gsd(p->src[s]);
  f.trunk = p->trunk; f.mask = static_cast<const act_t*>(p->mask); f.ml = to_layd(p->ml);
  f.dst = p->dst; f.dl = to_layd(p->dl); f.dst_f32 = p->dst_f32;
  f.H = p->H; f.W = p->W; f.C = p->C;
  return launch_pg(f, make_rowgeom(p->B, p->H, p->W, 0, 0), p->C / 8, stream);
}

template <int NS, bool TR>
static BnBwdBase<NS, TR> bn_common(const MmhBnBwd* p) {
  BnBwdBase<NS, TR> c;
  c.dz = p->dz; c.dz_f32 = p->dz_f32; c.relu = p->relu; c.dropout = p->dropout; c.key = p->drop_key;
  c.x = static_cast<const act_t*>(p->x); c.xl = to_layd(p->xl); c.coef = p->coef; c.save = p->save;
  for (int s = 0; s < NS; ++s) c.src[s] = to_gsd(p->src[s]);
  c.trunk = p->trunk;
  return c;
}
static int bn_bwd_check(const MmhBnBwd* p) {
  MMH_CHECK(p && p->x && p->coef && p->save, "null argument");
  MMH_REQ_VEC(p->xl.C);
  MMH_CHECK(p->nsrc >= 0 && p->nsrc <= 2, "nsrc=%d unsupported (0..2)", p->nsrc);
  MMH_CHECK(p->nsrc > 0 ? p->dz == nullptr : (p->dz != nullptr && p->trunk == nullptr),
            "give either dz or gradient sources (+ trunk)");
  for (int s = 0; s < p->nsrc; ++s) {
    MMH_CHECK(p->src[s].p != nullptr, "null gradient source");
    MMH_CHECK(p->src[s].l.H == p->xl.H && p->src[s].l.W == p->xl.W && p->src[s].l.B == p->xl.B,
              "gradient source %d has another shape", s);
  }
  return 0;
}
template <int NS, bool TR>
static int bn_bwd_reduce_t(const MmhBnBwd* p, void* stream) {
  BnBwdReduceF<NS, TR> f;
  f.cm = bn_common<NS, TR>(p);
  return launch_reduce_ch<2>(f, make_rowgeom(p->xl.B, p->xl.H, p->xl.W, 0, 0), p->xl.C / 8, p->xl.C, p->sums, stream);
}
static BnBwdFin bwd_fin(const float* sums, float count_global, float* k, float* dgamma, float* dbeta, int C) {
  BnBwdFin fin;
  fin.sums = sums;
  fin.f.sg = nullptr; fin.f.sl = nullptr; fin.f.k = k; fin.f.dgamma = dgamma; fin.f.dbeta = dbeta;
  fin.f.count = count_global; fin.f.C = C;
  return fin;
}
template <int NS, bool TR>
static int bn_bwd_reduce_fin_t(const MmhBnBwd* p, const BnBwdFin& fin, uint32_t* counter, void* stream) {
  BnBwdReduceF<NS, TR> f;
  f.cm = bn_common<NS, TR>(p);
  return launch_reduce_ch_fin<2>(f, make_rowgeom(p->xl.B, p->xl.H, p->xl.W, 0, 0), p->xl.C / 8, p->xl.C, p->sums, fin,
                                 counter, stream);
}
template <int NS, bool TR>
static int bn_bwd_apply_t(const MmhBnBwd* p, void* stream) {
  BnBwdApplyF<NS, TR> f;
  f.cm = bn_common<NS, TR>(p); f.k = p->k; f.dy = static_cast<act_t*>(p->dy); f.yl = to_layd(p->yl);
  return launch_pg(f, make_rowgeom(p->xl.B, p->xl.H, p->xl.W, 0, 0), p->xl.C / 8, stream);
}
// ---- lean path selection (MMH_BN_LEAN=0: always the general kernels)
static bool bn_lean_enabled() {       // read per call: tests switch between the two paths inside one process
  const char* e = getenv("MMH_BN_LEAN");
  return e == nullptr || atoi(e) != 0;
}
static bool bn_lean_ok(const MmhBnBwd* p, bool apply) {
  if (!bn_lean_enabled() || !ew_lean_enabled() || p->nsrc != 1 || p->trunk != nullptr) return false;
  if (!lay_fits32(p->xl) || !lay_fits32(p->src[0].l) || (apply && !lay_fits32(p->yl))) return false;
  if (p->xl.C > 1024 || !aligned16(p->coef) || !aligned16(p->save) || (apply && !aligned16(p->k))) return false;
  if (static_cast<int64_t>(p->xl.B) * p->xl.H * p->xl.W >= (int64_t(1) << 30)) return false;
  return true;
}
template <bool RELU, bool DROP>
static BnLeanBase<RELU, DROP> bn_lean_common(const MmhBnBwd* p) {
  BnLeanBase<RELU, DROP> c;
  c.x = static_cast<const act_t*>(p->x); c.xl = to_layd(p->xl); c.src = to_gsd(p->src[0]);
  c.coef = p->coef; c.save = p->save; c.key = p->drop_key;
  return c;
}
// reduce sweeps the rows from the last to the first (the data gradient that produced the source has just written
// them in ascending order), apply from the first to the last (= the rows the reduction touched last)
template <bool RELU, bool DROP>
static int bn_lean_reduce_t(const MmhBnBwd* p, const BnBwdFin* fin, uint32_t* counter, void* stream) {
  BnLeanReduceF<RELU, DROP> f;
  f.cm = bn_lean_common<RELU, DROP>(p);
  const RowGeom rg = make_rowgeom(p->xl.B, p->xl.H, p->xl.W, 0, 0);
  if (fin != nullptr) return launch_rows_reduce_fin<2>(f, rg, p->xl.C / 8, p->xl.C, p->sums, *fin, counter, 1, stream);
  return launch_rows_reduce<2>(f, rg, p->xl.C / 8, p->xl.C, p->sums, 1, stream);
}
template <bool RELU, bool DROP>
static int bn_lean_apply_t(const MmhBnBwd* p, void* stream) {
  BnLeanApplyF<RELU, DROP> f;
  f.cm = bn_lean_common<RELU, DROP>(p); f.k = p->k; f.dy = static_cast<act_t*>(p->dy); f.yl = to_layd(p->yl);
  return launch_rows_pg(f, make_rowgeom(p->xl.B, p->xl.H, p->xl.W, 0, 0), p->xl.C / 8, 0, stream);
}
static int bn_lean_reduce(const MmhBnBwd* p, const BnBwdFin* fin, uint32_t* counter, void* stream) {
  if (p->relu) return p->dropout ? bn_lean_reduce_t<true, true>(p, fin, counter, stream)
                                 : bn_lean_reduce_t<true, false>(p, fin, counter, stream);
  return p->dropout ? bn_lean_reduce_t<false, true>(p, fin, counter, stream)
                    : bn_lean_reduce_t<false, false>(p, fin, counter, stream);
}
static int bn_lean_apply(const MmhBnBwd* p, void* stream) {
  if (p->relu) return p->dropout ? bn_lean_apply_t<true, true>(p, stream) : bn_lean_apply_t<true, false>(p, stream);
  return p->dropout ? bn_lean_apply_t<false, true>(p, stream) : bn_lean_apply_t<false, false>(p, stream);
}
#define MMH_BN_BWD_DISPATCH(fn)                                                            \
  do {                                                                                     \
    const bool tr = p->trunk != nullptr;                                                   \
    if (p->nsrc == 0) return fn<0, false>(p, stream);                                      \
    if (p->nsrc == 1) return tr ? fn<1, true>(p, stream) : fn<1, false>(p, stream);        \
    return tr ? fn<2, true>(p, stream) : fn<2, false>(p, stream);                          \
  } while (0)
extern "C" int mmh_bn_bwd_reduce(const MmhBnBwd* p, void* stream) {
  if (bn_bwd_check(p)) return 1;
  MMH_CHECK(p->sums, "null argument");
  if (bn_lean_ok(p, false)) return bn_lean_reduce(p, nullptr, nullptr, stream);
  MMH_BN_BWD_DISPATCH(bn_bwd_reduce_t);
}
extern "C" int mmh_bn_bwd_reduce_finalize(MmhPeer* peer, uint32_t seq, const MmhBnBwd* p, uint32_t* counter,
                                          float count_global, float* dgamma, float* dbeta, void* stream) {
  if (bn_bwd_check(p)) return 1;
  MMH_CHECK(p->sums && p->k && counter, "null argument");
  BnBwdFin fin = bwd_fin(p->sums, count_global, const_cast<float*>(p->k), dgamma, dbeta, p->xl.C);
  if (peer_dev(peer, seq, 2 * p->xl.C, &fin.px)) return 1;
  if (bn_lean_ok(p, false)) return bn_lean_reduce(p, &fin, counter, stream);
  const bool tr = p->trunk != nullptr;
  if (p->nsrc == 0) return bn_bwd_reduce_fin_t<0, false>(p, fin, counter, stream);
  if (p->nsrc == 1) return tr ? bn_bwd_reduce_fin_t<1, true>(p, fin, counter, stream)
                              : bn_bwd_reduce_fin_t<1, false>(p, fin, counter, stream);
  return tr ? bn_bwd_reduce_fin_t<2, true>(p, fin, counter, stream)
            : bn_bwd_reduce_fin_t<2, false>(p, fin, counter, stream);
}
extern "C" int mmh_bn_bwd_finalize_reset(MmhPeer* peer, uint32_t seq, float* sums, float count_global, float* k,
                                         float* dgamma, float* dbeta, int32_t C, void* stream) {
  MMH_CHECK(sums && k && C > 0, "bad argument");
  BnBwdFin fin = bwd_fin(sums, count_global, k, dgamma, dbeta, C);
  if (peer_dev(peer, seq, 2 * C, &fin.px)) return 1;
#ifdef MMH_HOST_EMU
  (void)stream;
  for (int c = 0; c < C; ++c) { fin(c); sums[c] = 0.f; sums[C + c] = 0.f; }
#else
  MMH_CUDA(launch_k(fin_reset_kernel<BnBwdFin>, dim3((C + 255) / 256), dim3(256), 0, stream, fin, sums, C));
#endif
  return 0;
}
extern "C" int mmh_bn_bwd_apply(const MmhBnBwd* p, void* stream) {
  if (bn_bwd_check(p)) return 1;
  MMH_CHECK(p->k && p->dy, "null argument");
  if (bn_lean_ok(p, true)) return bn_lean_apply(p, stream);
  MMH_BN_BWD_DISPATCH(bn_bwd_apply_t);
}
extern "C" int mmh_bn_bwd_finalize(const float* sums_global, const float* sums_local, float count, float* k,
                                   float* dgamma, float* dbeta, int32_t C, void* stream) {
  MMH_CHECK(sums_global && sums_local && k, "null argument");
  BnBwdFinalizeF f;
  f.sg = sums_global; f.sl = sums_local; f.k = k; f.dgamma = dgamma; f.dbeta = dbeta; f.count = count; f.C = C;
  return launch_map(f, C, stream);
}

static GateBwdBase gate_common(const MmhGateBwd* p) {
  GateBwdBase c;
  c.dout = p->dout; c.c1 = static_cast<const act_t*>(p->c1); c.x2o = static_cast<const act_t*>(p->x2o);
  c.x3o = static_cast<const act_t*>(p->x3o); c.sl = to_layd(p->sl); c.coef = p->coef; c.save = p->save;
  return c;
}
extern "C" int mmh_gate_bwd_reduce(const MmhGateBwd* p, void* stream) {
  MMH_CHECK(p && p->dout && p->c1 && p->x2o && p->coef && p->save && p->sums, "null argument");
  MMH_REQ_VEC(p->sl.C);
  GateBwdReduceF f;
  f.cm = gate_common(p);
  return launch_reduce_ch<2>(f, make_rowgeom(p->sl.B, p->sl.H, p->sl.W, 0, 0), p->sl.C / 8, p->sl.C, p->sums, stream);
}
extern "C" int mmh_gate_bwd_reduce_finalize(MmhPeer* peer, uint32_t seq, const MmhGateBwd* p, uint32_t* counter,
                                            float count_global, float* dgamma, float* dbeta, void* stream) {
  MMH_CHECK(p && p->dout && p->c1 && p->x2o && p->coef && p->save && p->sums && p->k && counter, "null argument");
  MMH_REQ_VEC(p->sl.C);
  GateBwdReduceF f;
  f.cm = gate_common(p);
  BnBwdFin fin = bwd_fin(p->sums, count_global, const_cast<float*>(p->k), dgamma, dbeta, p->sl.C);
  if (peer_dev(peer, seq, 2 * p->sl.C, &fin.px)) return 1;
  return launch_reduce_ch_fin<2>(f, make_rowgeom(p->sl.B, p->sl.H, p->sl.W, 0, 0), p->sl.C / 8, p->sl.C, p->sums, fin,
                                 counter, stream);
}
extern "C" int mmh_gate_bwd_apply(const MmhGateBwd* p, void* stream) {
  MMH_CHECK(p && p->dout && p->c1 && p->x2o && p->coef && p->save && p->k && p->dy1 && p->dy2, "null argument");
  MMH_CHECK((p->x3o != nullptr) == (p->dy3 != nullptr), "x3o and dy3 come together (three-stream) or not at all");
  MMH_REQ_VEC(p->sl.C);
  GateBwdApplyF f;
  f.cm = gate_common(p); f.k = p->k;
  f.ex2 = to_gsd(p->ex2); f.ex3 = to_gsd(p->ex3);
  f.dy1 = static_cast<act_t*>(p->dy1); f.dy2 = static_cast<act_t*>(p->dy2); f.dy3 = static_cast<act_t*>(p->dy3);
  f.yl = to_layd(p->yl);
  return launch_pg(f, make_rowgeom(p->sl.B, p->sl.H, p->sl.W, 0, 0), p->sl.C / 8, stream);
}
