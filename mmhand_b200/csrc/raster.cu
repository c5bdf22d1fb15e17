// Keypoint -> Gaussian heatmap rasteriser (data/generic_dataset.py:191-217,238-242 of the reference).
// One work item = 4 consecutive pixels of one map row (one 16-byte fp32 store); the kernel is bound by the
// HBM write of 4*H*W bytes per map. All arithmetic in fp64 in the reference's evaluation order
// (D2 = (gx-x)^2 + (gy-y)^2; exp(-D2 / 2.0 / sigma / sigma); clamp >1; zero <thresh; cast last). The exp is
// skipped only where the value is provably far below the threshold. Dual-mode source.
#include <math.h>

#include "ew_framework.h"

namespace mmh {

struct RasterF {
  const double* uv; float* out; int H, W, wq; double sigma, thresh, d2_skip;
  MMH_HD void operator()(int64_t i) const {
    const int xq = static_cast<int>(i % wq);
    const int y = static_cast<int>((i / wq) % H);
    const int64_t m = i / (static_cast<int64_t>(wq) * H);
    const double cx = uv[2 * m], cy = uv[2 * m + 1];
    const double dy2 = (static_cast<double>(y) - cy) * (static_cast<double>(y) - cy);
    F32x4 o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double gx = static_cast<double>(xq * 4 + j);
      const double D2 = (gx - cx) * (gx - cx) + dy2;
      float r = 0.f;
      if (D2 < d2_skip) {
        double v = exp(-D2 / 2.0 / sigma / sigma);
        if (v > 1.0) v = 1.0;
        if (v < thresh) v = 0.0;
        r = static_cast<float>(v);
      }
      o.v[j] = r;
    }
    *reinterpret_cast<F32x4*>(out + (m * H + y) * W + xq * 4) = o;
  }
};

}  // namespace mmh

using namespace mmh;

extern "C" int mmh_heatmap_rasterize(const double* uv, int64_t n_maps, int32_t H, int32_t W, double sigma,
                                     double thresh, float* out, void* stream) {
  MMH_CHECK(uv && out, "null argument");
  MMH_CHECK((W % 4) == 0, "W=%d must be a multiple of 4", W);
  RasterF f;
  f.uv = uv; f.out = out; f.H = H; f.W = W; f.wq = W / 4; f.sigma = sigma; f.thresh = thresh;
  // exp(-D2/(2 sigma^2)) < thresh  <=>  D2 > -2 sigma^2 ln(thresh); keep a 1 % + 1 margin so that the decision
  // at the threshold itself is always taken by the fp64 comparison, exactly like the reference
  f.d2_skip = thresh > 0.0 ? (-2.0 * sigma * sigma * log(thresh)) * 1.01 + 1.0 : 1e300;
  return launch_map(f, n_maps * H * f.wq, stream);
}
