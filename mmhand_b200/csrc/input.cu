// Device side of the input pipeline: what the reference's dataset workers do on the CPU for every sample
// (data/generic_dataset.py:133-159) runs on the GPU on the bytes cv2.imread produced, so that uint8 frames cross the bus
// instead of fp32 / fp64 tensors (16x / 32x fewer bytes):
//   colour frame  cv2.cvtColor(BGR2RGB) -> ((img / 255.0) - 0.5) / 0.5 in float64 -> .float(), HWC -> CHW   (:140-143,
//                                                                                                            :182-189)
//   depth frame   256.0 * px[hi] + px[lo] -> ((d / div) - 0.5) / 0.5 in float64, stacked x3                  (:148-159)
// Same float64 arithmetic, cast to fp32 last: bit-identical to the reference's tensors after set_input's fp32 copy.
// One item = one pixel (three adjacent byte loads, three coalesced plane stores). Dual-mode source.
#include "ew_framework.h"

namespace mmh {

struct ImageUnpackF {
  const uint8_t* src; float* dst; int64_t hw; int swap_rb;
  MMH_HD static float q(uint8_t u) { return static_cast<float>((static_cast<double>(u) / 255.0 - 0.5) / 0.5); }
  MMH_HD void operator()(int64_t i) const {
    const int64_t b = i / hw, p = i - b * hw;
    const uint8_t* s = src + i * 3;
    float* d = dst + b * 3 * hw + p;
    d[0] = q(s[swap_rb ? 2 : 0]);
    d[hw] = q(s[1]);
    d[2 * hw] = q(s[swap_rb ? 0 : 2]);
  }
};

struct DepthUnpackF {
  const uint8_t* src; float* dst; int64_t hw; int hi, lo; double div;
  MMH_HD void operator()(int64_t i) const {
    const int64_t b = i / hw, p = i - b * hw;
    const uint8_t* s = src + i * 3;
    const double depth = 256.0 * static_cast<double>(s[hi]) + static_cast<double>(s[lo]);
    const float v = static_cast<float>((depth / div - 0.5) / 0.5);
    float* d = dst + b * 3 * hw + p;
    d[0] = v; d[hw] = v; d[2 * hw] = v;
  }
};

}  // namespace mmh

using namespace mmh;

extern "C" int mmh_image_unpack_u8(const uint8_t* src_nhwc, int64_t n_img, int32_t H, int32_t W, int32_t swap_rb,
                                   float* dst_nchw, void* stream) {
  if (n_img <= 0) return 0;
  MMH_CHECK(src_nhwc && dst_nchw && H > 0 && W > 0, "bad argument");
  ImageUnpackF f;
  f.src = src_nhwc; f.dst = dst_nchw; f.hw = static_cast<int64_t>(H) * W; f.swap_rb = swap_rb;
  return launch_map(f, n_img * f.hw, stream);
}

extern "C" int mmh_depth_unpack_u8(const uint8_t* src_nhwc, int64_t n_img, int32_t H, int32_t W, int32_t hi_ch,
                                   int32_t lo_ch, double div, float* dst_nchw3, void* stream) {
  if (n_img <= 0) return 0;
  MMH_CHECK(src_nhwc && dst_nchw3 && H > 0 && W > 0 && hi_ch >= 0 && hi_ch < 3 && lo_ch >= 0 && lo_ch < 3 && div > 0.0,
            "bad argument");
  DepthUnpackF f;
  f.src = src_nhwc; f.dst = dst_nchw3; f.hw = static_cast<int64_t>(H) * W; f.hi = hi_ch; f.lo = lo_ch; f.div = div;
  return launch_map(f, n_img * f.hw, stream);
}
