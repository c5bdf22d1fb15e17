// Structural similarity on the device: the evaluator of the reference's benchmark harness scores every generated image
// with pytorch_ssim.ssim (baselines/quantitative_on_benchmarks/pytorch_ssim/__init__.py:17-39, utils.py:100-111): five
// zero-padded depthwise 11 x 11 Gaussian convolutions (mu1, mu2, E[x^2], E[y^2], E[xy]), the SSIM map and its mean --
// seven full-frame torch kernels per pair. Here one kernel: a thread owns one output pixel of one channel, accumulates
// the five window sums in fp32 from the (L1/L2-resident) neighbourhood, forms the map value and the block reduces the
// per-image sums. Window weights: the fp32 outer product of the normalised 1-D window, as the reference builds them.
// Dual-mode source.
#include <math.h>

#include "ew_framework.h"

namespace mmh {

constexpr int kSsimMaxWin = 15;

struct SsimF {
  const float* a; const float* b; int C, H, W, win;
  float w1[kSsimMaxWin];            // normalised 1-D Gaussian window (fp32)
  float inv_n;                      // 1 / (C*H*W): per-image mean
  float* per_image;                 // [B] += mean SSIM of image b (may be NULL)
  MMH_HD float operator()(int64_t i) const {
    const int64_t hw = static_cast<int64_t>(H) * W;
    const int64_t plane = i / hw;
    const int p = static_cast<int>(i - plane * hw);
    const int y = p / W, x = p - y * W;
    const float* pa = a + plane * hw;
    const float* pb = b + plane * hw;
    const int r = win / 2;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int dy = 0; dy < win; ++dy) {
      const int yy = y + dy - r;
      if (yy < 0 || yy >= H) continue;
      for (int dx = 0; dx < win; ++dx) {
        const int xx = x + dx - r;
        if (xx < 0 || xx >= W) continue;
        const float w = w1[dy] * w1[dx];
        const float u = pa[yy * W + xx], v = pb[yy * W + xx];
        m1 += w * u; m2 += w * v;
        s11 += w * (u * u); s22 += w * (v * v); s12 += w * (u * v);
      }
    }
    const float m1s = m1 * m1, m2s = m2 * m2, m12 = m1 * m2;
    const float v1 = s11 - m1s, v2 = s22 - m2s, v12 = s12 - m12;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float val = ((2.f * m12 + C1) * (2.f * v12 + C2)) / ((m1s + m2s + C1) * (v1 + v2 + C2));
    if (per_image != nullptr) {
#if defined(__CUDA_ARCH__)
      atomicAdd(per_image + plane / C, val * inv_n);
#else
      per_image[plane / C] += val * inv_n;
#endif
    }
    return val;
  }
};

// per-image means only: the map values are not summed globally
struct SsimPerImageF {
  SsimF f;
  MMH_HD void operator()(int64_t i) const { (void)f(i); }
};

}  // namespace mmh

using namespace mmh;

extern "C" int mmh_ssim(const float* img1, const float* img2, int64_t B, int32_t C, int32_t H, int32_t W,
                        int32_t window, float sigma, float* mean_acc, float* per_image, void* stream) {
  if (B <= 0) return 0;
  MMH_CHECK(img1 && img2 && C > 0 && H > 0 && W > 0 && (mean_acc || per_image), "bad argument");
  MMH_CHECK(window >= 1 && window <= kSsimMaxWin && (window & 1) && sigma > 0.f, "window %d unsupported", window);
  SsimF f;
  f.a = img1; f.b = img2; f.C = C; f.H = H; f.W = W; f.win = window; f.per_image = per_image;
  // pytorch_ssim.gaussian: exp(-(x - w//2)^2 / (2 sigma^2)) in Python floats, stored as fp32, normalised by the fp32 sum
  float g[kSsimMaxWin], sum = 0.f;
  for (int x = 0; x < window; ++x) {
    g[x] = static_cast<float>(exp(-static_cast<double>((x - window / 2) * (x - window / 2)) / (2.0 * sigma * sigma)));
    sum += g[x];
  }
  for (int x = 0; x < kSsimMaxWin; ++x) f.w1[x] = x < window ? g[x] / sum : 0.f;
  const int64_t n = B * C * static_cast<int64_t>(H) * W;
  f.inv_n = 1.0f / static_cast<float>(static_cast<int64_t>(C) * H * W);
  if (mean_acc != nullptr) return launch_reduce_scalar(f, n, mean_acc, stream);     // *mean_acc += sum of the map
  SsimPerImageF pi;
  pi.f = f;
  return launch_map(pi, n, stream);
}
