// aug.py's image write-out on the device (reference aug.py:57-71): generated image fp32 NCHW in (-1, 1) ->
// (x * 0.5 + 0.5) * 255. in float32 (numpy's two roundings, no FMA), RGB -> BGR (cv2.cvtColor), and the
// saturate_cast<uchar>(cvRound(v)) that cv2.imwrite applies to a float image -> uint8 [B][H][W][3] ready for the
// encoder: the device-to-host copy shrinks 4x and the host does no arithmetic. One item = one pixel (three coalesced
// plane reads, three adjacent byte stores). Dual-mode source.
#include <math.h>

#include "ew_framework.h"

namespace mmh {

struct PackBgr8F {
  const float* src; uint8_t* dst; int64_t hw;
  MMH_HD static uint8_t q(float x) {
#if defined(__CUDA_ARCH__)
    const float v = __fmul_rn(__fadd_rn(__fmul_rn(x, 0.5f), 0.5f), 255.f);
#else
    volatile float a = x * 0.5f;
    volatile float b = a + 0.5f;
    const float v = b * 255.f;
#endif
    const float r = rintf(v);                       // cvRound: half to even
    return static_cast<uint8_t>(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));      // NaN -> 0 like saturate_cast
  }
  MMH_HD void operator()(int64_t i) const {
    const int64_t b = i / hw, p = i - b * hw;
    const float* s = src + b * 3 * hw + p;
    uint8_t* d = dst + i * 3;
    d[0] = q(s[2 * hw]);      // B
    d[1] = q(s[hw]);          // G
    d[2] = q(s[0]);           // R
  }
};

}  // namespace mmh

using namespace mmh;

extern "C" int mmh_image_pack_bgr8(const float* src_nchw, int64_t n_img, int32_t H, int32_t W, uint8_t* dst_nhwc,
                                   void* stream) {
  if (n_img <= 0) return 0;
  MMH_CHECK(src_nchw && dst_nhwc && H > 0 && W > 0, "bad argument");
  PackBgr8F f;
  f.src = src_nchw; f.dst = dst_nhwc; f.hw = static_cast<int64_t>(H) * W;
  return launch_map(f, n_img * f.hw, stream);
}
