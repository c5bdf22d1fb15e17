// Keypoint -> Gaussian heatmap rasteriser (data/generic_dataset.py:191-217,238-242 of the reference).
// One work item = 4 consecutive pixels of one map row (one 16-byte fp32 store); the kernel is bound by the
// HBM write of 4*H*W bytes per map. All arithmetic in fp64 in the reference's evaluation order
// (D2 = (gx-x)^2 + (gy-y)^2; exp(-D2 / 2.0 / sigma / sigma); clamp >1; zero <thresh; cast last).
// A map is non-zero only inside a disc of radius sqrt(-2 sigma^2 ln thresh) (18.2 px at sigma = 6): each map's
// conservative integer bounding box is computed once (prep), items outside it are plain zero stores (98 % of a
// 256 x 256 map), items inside run the fp64 path, where the decision at the threshold itself is always taken by the
// fp64 comparison, exactly like the reference. Dual-mode source.
#include <math.h>

#include "ew_framework.h"

namespace mmh {

struct RasterF {
  const double* uv; float* out; int H, W, wq; double sigma, thresh, d2_skip, radius;
  struct Ctx { double cx, cy; int x0, x1, y0, y1; };      // bounding box, inclusive, in pixels
  MMH_HD void prep(int64_t m, Ctx& c) const {
    c.cx = uv[2 * m];
    c.cy = uv[2 * m + 1];
    if (c.cx == c.cx && c.cy == c.cy) {
      // clamp in fp64 first: coordinates far outside the frame must not overflow the integer conversion
      const double xl = fmax(c.cx - radius - 1.0, -1.0), xh = fmin(c.cx + radius + 1.0, static_cast<double>(W));
      const double yl = fmax(c.cy - radius - 1.0, -1.0), yh = fmin(c.cy + radius + 1.0, static_cast<double>(H));
      c.x0 = static_cast<int>(floor(xl)); c.x1 = static_cast<int>(ceil(xh));
      c.y0 = static_cast<int>(floor(yl)); c.y1 = static_cast<int>(ceil(yh));
      if (xl > xh) { c.x0 = 1; c.x1 = 0; }
      if (yl > yh) { c.y0 = 1; c.y1 = 0; }
    } else {                    // NaN coordinate: the reference's map is NaN everywhere -- evaluate every pixel
      c.x0 = 0; c.x1 = W; c.y0 = 0; c.y1 = H;
    }
  }
  // q = index of the 4-pixel item inside the map (row-major)
  MMH_HD void item(int64_t m, int q, const Ctx& c) const {
    const int y = q / wq, xq = q - y * wq;
    F32x4 o;
    o.v[0] = o.v[1] = o.v[2] = o.v[3] = 0.f;
    if (y >= c.y0 && y <= c.y1 && xq * 4 + 3 >= c.x0 && xq * 4 <= c.x1) {
      const double dy2 = (static_cast<double>(y) - c.cy) * (static_cast<double>(y) - c.cy);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double gx = static_cast<double>(xq * 4 + j);
        const double D2 = (gx - c.cx) * (gx - c.cx) + dy2;
        if (!(D2 >= d2_skip)) {
          double v = exp(-D2 / 2.0 / sigma / sigma);
          if (v > 1.0) v = 1.0;
          if (v < thresh) v = 0.0;
          o.v[j] = static_cast<float>(v);
        }
      }
    }
    *reinterpret_cast<F32x4*>(out + (m * H + y) * static_cast<int64_t>(W) + xq * 4) = o;
  }
};

#ifndef MMH_HOST_EMU
// One block walks over maps (grid-stride); its 256 threads sweep the H*W/4 items of a map with consecutive
// 16-byte stores. Map index, centre and bounding box are block-uniform.
__global__ void __launch_bounds__(256) raster_kernel(const RasterF f, const int64_t n_maps) {
  pdl_sync();
  const int per_map = f.H * f.wq;
  for (int64_t m = blockIdx.x; m < n_maps; m += gridDim.x) {
    RasterF::Ctx c;
    f.prep(m, c);
    for (int q = threadIdx.x; q < per_map; q += 256) f.item(m, q, c);
  }
}
#endif

// Offline pose maps of the reference's dataset tool (tool/generate_pose_map_RHD.py:22-29, ``cords_to_map``):
// result[y][x][j] = exp(-((y - cy)^2 + (x - cx)^2) / (2 sigma^2)) in float64, cast to fp32, HWC layout, no clamp and no
// threshold; a joint whose y or x equals MISSING_VALUE (-1) leaves its plane zero. One item = one pixel of one pose: J
// consecutive floats. Values below the smallest fp32 denormal (exponent < -105) are written as the 0 they round to.
struct PoseMapF {
  const double* yx; float* out; int J, H, W; double two_sigma2, missing;
  MMH_HD void operator()(int64_t i) const {
    const int64_t hw = static_cast<int64_t>(H) * W;
    const int64_t pose = i / hw;
    const int p = static_cast<int>(i - pose * hw);
    const int y = p / W, x = p - y * W;
    const double* c = yx + pose * J * 2;
    float* o = out + i * J;
    for (int j = 0; j < J; ++j) {
      const double cy = c[2 * j], cx = c[2 * j + 1];
      float v = 0.f;
      if (!(cy == missing || cx == missing)) {
        const double dy = static_cast<double>(y) - cy, dx = static_cast<double>(x) - cx;
        const double e = -(dy * dy + dx * dx) / two_sigma2;
        v = e < -105.0 ? 0.f : static_cast<float>(exp(e));
      }
      o[j] = v;
    }
  }
};

}  // namespace mmh

using namespace mmh;

extern "C" int mmh_pose_map_rasterize(const double* yx, int64_t n_pose, int32_t J, int32_t H, int32_t W, double sigma,
                                      double missing, float* out_hwc, void* stream) {
  if (n_pose <= 0) return 0;
  MMH_CHECK(yx && out_hwc && J > 0 && H > 0 && W > 0 && sigma > 0.0, "bad argument");
  PoseMapF f;
  f.yx = yx; f.out = out_hwc; f.J = J; f.H = H; f.W = W; f.missing = missing;
  f.two_sigma2 = 2.0 * (sigma * sigma);          // the tool divides by the single number (2 * sigma ** 2)
  return launch_map(f, n_pose * static_cast<int64_t>(H) * W, stream);
}

extern "C" int mmh_heatmap_rasterize(const double* uv, int64_t n_maps, int32_t H, int32_t W, double sigma,
                                     double thresh, float* out, void* stream) {
  if (n_maps <= 0) return 0;          /* empty pose list: nothing to write */
  MMH_CHECK(uv && out, "null argument");
  MMH_CHECK((W % 4) == 0, "W=%d must be a multiple of 4", W);
  MMH_CHECK(H > 0 && W > 0 && static_cast<int64_t>(H) * W < (int64_t(1) << 31), "frame %dx%d unsupported", H, W);
  RasterF f;
  f.uv = uv; f.out = out; f.H = H; f.W = W; f.wq = W / 4; f.sigma = sigma; f.thresh = thresh;
  // exp(-D2/(2 sigma^2)) < thresh  <=>  D2 > -2 sigma^2 ln(thresh); keep a 1 % + 1 margin so that the decision
  // at the threshold itself is always taken by the fp64 comparison, exactly like the reference
  f.d2_skip = thresh > 0.0 ? (-2.0 * sigma * sigma * log(thresh)) * 1.01 + 1.0 : 1e300;
  f.radius = thresh > 0.0 ? sqrt(f.d2_skip) : 1e300;
#ifdef MMH_HOST_EMU
  (void)stream;
  for (int64_t m = 0; m < n_maps; ++m) {
    RasterF::Ctx c;
    f.prep(m, c);
    for (int q = 0; q < H * f.wq; ++q) f.item(m, q, c);
  }
  return 0;
#else
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  const int blocks = static_cast<int>(n_maps < cap ? n_maps : cap);
  MMH_CUDA(launch_k(raster_kernel, dim3(blocks), dim3(256), 0, stream, f, n_maps));
  return 0;
#endif
}
