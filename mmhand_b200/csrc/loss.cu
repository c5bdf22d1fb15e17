// Loss reductions and their gradients: BCE-with-logits (PatchGAN-style over the 256-channel logit map),
// image L1, perceptual L1/MSE on VGG features, tanh backward, input-gradient extraction.
// HBM-bound: each element is read once (and its gradient written once).
// Dual-mode source (see ew_framework.h).
#include <math.h>

#include "ew_framework.h"

namespace mmh {

struct BceF {
  const float* x; float* grad; float target, ls, gs;
  MMH_HD float operator()(int64_t i) const {
    const float v = x[i];
    const float l = fmaxf(v, 0.f) - v * target + log1pf(expf(-fabsf(v)));
    if (grad != nullptr) grad[i] = gs * (1.0f / (1.0f + expf(-v)) - target);
    return ls * l;
  }
};

struct L1F {
  const float* a; const float* b; float* grad; float ls, gs;
  MMH_HD float operator()(int64_t i) const {
    const float d = a[i] - b[i];
    if (grad != nullptr) grad[i] += gs * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    return ls * fabsf(d);
  }
};

struct PercF {
  const act_t* ff; const act_t* ft; act_t* dy; int mse, linear; float ls, gs;
  MMH_HD float operator()(int64_t i) const {
    float a[8], b[8], g[8];
    ld8_bf16(ff + i * 8, a);
    ld8_bf16(ft + i * 8, b);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = a[j] - b[j];
      float gr;
      if (mse) { s += d * d; gr = 2.f * d; }
      else { s += fabsf(d); gr = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
      g[j] = (linear || a[j] > 0.f) ? gs * gr : 0.f;   // through the ReLU that produced ff (if one did)
    }
    if (dy != nullptr) st8_bf16(dy + i * 8, g);
    return ls * s;
  }
};

struct TanhBwdF {
  const float* d; const float* y; act_t* dy; LayD yl; int C;
  MMH_HD void operator()(int64_t i) const {
    const int H = yl.H, W = yl.W;
    const uint32_t iu = static_cast<uint32_t>(i);
    const int w = static_cast<int>(iu % W);
    const int h = static_cast<int>((iu / W) % H);
    const int b = static_cast<int>(iu / (static_cast<uint32_t>(W) * H));
    float o[8];
    zero8(o);
    for (int c = 0; c < C; ++c) {
      const int64_t s = ((static_cast<int64_t>(b) * C + c) * H + h) * W + w;
      o[c] = d[s] * (1.f - y[s] * y[s]);
    }
    st8_bf16(dy + lay_off(yl, b, h, w), o);
  }
};

struct GradSrc1 {
  const act_t* p; LayD l; int lo, hi, reflect;
};
MMH_HD int preimages1(int i, int n, int lo, int hi, int reflect, int (&out)[3]) {
  int k = 0;
  out[k++] = i;
  if (reflect) {
    if (i >= 1 && i <= lo) out[k++] = -i;
    if (i <= n - 2 && i >= n - 1 - hi) out[k++] = 2 * (n - 1) - i;
  }
  return k;
}
struct InputGradF {
  GradSrc1 s; const float* scale; float* dst; int C, H, W, accumulate;
  MMH_HD void operator()(int64_t i) const {
    const uint32_t iu = static_cast<uint32_t>(i);
    const int w = static_cast<int>(iu % W);
    const int h = static_cast<int>((iu / W) % H);
    const int c = static_cast<int>((iu / (static_cast<uint32_t>(W) * H)) % C);
    const int b = static_cast<int>(iu / (static_cast<uint32_t>(W) * H * C));
    int hs[3], ws[3];
    const int nh = preimages1(h, H, s.lo, s.hi, s.reflect, hs);
    const int nw = preimages1(w, W, s.lo, s.hi, s.reflect, ws);
    float acc = 0.f;
    for (int a = 0; a < nh; ++a)
      for (int e = 0; e < nw; ++e) acc += act2f(s.p[lay_off(s.l, b, hs[a], ws[e]) + c]);
    if (scale != nullptr) acc *= scale[c];
    dst[i] = accumulate ? dst[i] + acc : acc;
  }
};

struct GridToNchwF {
  const float* src; LayD sl; float* dst; int C;
  MMH_HD void operator()(int64_t i) const {
    const int H = sl.H, W = sl.W;
    const uint32_t iu = static_cast<uint32_t>(i);
    const int w = static_cast<int>(iu % W);
    const int h = static_cast<int>((iu / W) % H);
    const int c = static_cast<int>((iu / (static_cast<uint32_t>(W) * H)) % C);
    const int b = static_cast<int>(iu / (static_cast<uint32_t>(W) * H * C));
    dst[i] = src[lay_off(sl, b, h, w) + c];
  }
};

}  // namespace mmh

using namespace mmh;

extern "C" int mmh_bce_logits(const float* x, int64_t n, float target, float loss_scale, float grad_scale,
                              float* loss_acc, float* grad, void* stream) {
  MMH_CHECK(x && loss_acc, "null argument");
  BceF f;
  f.x = x; f.grad = grad; f.target = target; f.ls = loss_scale; f.gs = grad_scale;
  return launch_reduce_scalar(f, n, loss_acc, stream);
}

extern "C" int mmh_l1_f32(const float* a, const float* b, int64_t n, float loss_scale, float grad_scale,
                          float* loss_acc, float* grad_acc, void* stream) {
  MMH_CHECK(a && b && loss_acc, "null argument");
  L1F f;
  f.a = a; f.b = b; f.grad = grad_acc; f.ls = loss_scale; f.gs = grad_scale;
  return launch_reduce_scalar(f, n, loss_acc, stream);
}

extern "C" int mmh_perc_loss(const void* ff, const void* ft, int64_t n, int32_t mse, float loss_scale,
                             float grad_scale, float* loss_acc, void* dy, void* stream) {
  MMH_CHECK(ff && ft && loss_acc, "null argument");
  MMH_CHECK((n % 8) == 0, "element count must be a multiple of 8");
  PercF f;
  f.ff = static_cast<const act_t*>(ff); f.ft = static_cast<const act_t*>(ft);
  // mse: bit 0 = squared error instead of absolute; bit 1 = the features are a convolution's output, not a ReLU's
  f.dy = static_cast<act_t*>(dy); f.mse = mse & 1; f.linear = (mse >> 1) & 1; f.ls = loss_scale; f.gs = grad_scale;
  return launch_reduce_scalar(f, n / 8, loss_acc, stream);
}

extern "C" int mmh_tanh_bwd(const float* dfake_nchw, const float* fake_nchw, void* dy, const MmhLay* yl, int32_t C,
                            void* stream) {
  MMH_CHECK(dfake_nchw && fake_nchw && dy && yl, "null argument");
  MMH_CHECK(C >= 1 && C <= 8, "C=%d unsupported", C);
  TanhBwdF f;
  f.d = dfake_nchw; f.y = fake_nchw; f.dy = static_cast<act_t*>(dy); f.yl = to_layd(*yl); f.C = C;
  return launch_map(f, static_cast<int64_t>(yl->B) * yl->H * yl->W, stream);
}

extern "C" int mmh_input_grad_nchw(const MmhGradSrc* src, const float* scale, float* dst_nchw, int32_t B, int32_t C,
                                   int32_t H, int32_t W, int32_t accumulate, void* stream) {
  MMH_CHECK(src && src->p && dst_nchw, "null argument");
  InputGradF f;
  f.s.p = static_cast<const act_t*>(src->p); f.s.l = to_layd(src->l);
  f.s.lo = src->pad_lo; f.s.hi = src->pad_hi; f.s.reflect = src->reflect;
  f.scale = scale; f.dst = dst_nchw; f.C = C; f.H = H; f.W = W; f.accumulate = accumulate;
  return launch_map(f, static_cast<int64_t>(B) * C * H * W, stream);
}

extern "C" int mmh_grid_to_nchw(const float* src, const MmhLay* sl, float* dst_nchw, int32_t C, void* stream) {
  MMH_CHECK(src && sl && dst_nchw, "null argument");
  GridToNchwF f;
  f.src = src; f.sl = to_layd(*sl); f.dst = dst_nchw; f.C = C;
  return launch_map(f, static_cast<int64_t>(sl->B) * C * sl->H * sl->W, stream);
}
