// Depth-ordered hand part map (generate_jointsmap, data/generic_dataset.py:30-78 of the reference): 20 bones, each
// an ellipse polygon (cv2.ellipse2Poly, 1-degree steps) filled with cv2.fillConvexPoly; a pixel shows the colour of
// the last bone whose depth equals the running minimum depth at that pixel.
//
// OpenCV's two routines are restated in their own integer arithmetic (oracle/jointsmap_ref.py is the same
// restatement in Python, pinned to the real cv2 on thousands of polygons): float sine table + double products +
// cvRound, consecutive duplicates removed; fillConvexPoly = outline with 8-connected LineIterator lines
// (clipLine at the frame) + XY_SHIFT = 16 fixed-point edge walk. Because the polygons are ellipses of minor radius
// 5 every covered row is ONE run of pixels (asserted by the tests against the mask-based oracle), so a bone is kept
// as a per-row span [lo, hi] and the frame is composed in one pass over the pixels with the 20 bones in registers.
//
// One block = one pose, warp b = bone b: lanes compute the 361 polygon points and walk the outline segments
// (atomicMin / atomicMax on the row spans in shared memory), lane 0 removes duplicates and runs the sequential edge
// walk; then all threads compose and store. Floating point follows the reference's Python expression by expression
// (no FMA contraction); atan2 is exact for the axis-aligned and diagonal directions that integer pixel coordinates
// produce, elsewhere int(degrees(atan2)) could only differ when the angle is within an ulp of a whole degree.
// Dual-mode source (host loops with -DMMH_HOST_EMU).
#include <math.h>
#include <stdint.h>

#include "ew_framework.h"
#include "sin_table.h"

namespace mmh {

constexpr int kJmBones = 20;
constexpr int kJmMaxPts = 362;
constexpr int kJmMaxH = 512;
constexpr int kJmShift = 16;
constexpr int kJmOne = 1 << kJmShift;

static const float kSinHost[451] = {MMH_SIN_TABLE_VALUES};
#ifndef MMH_HOST_EMU
__constant__ float kSinDev[451] = {MMH_SIN_TABLE_VALUES};
#endif
#if defined(__CUDA_ARCH__)
#define MMH_SIN(k) kSinDev[k]
#define MMH_MUL(a, b) __dmul_rn((a), (b))
#define MMH_ADD(a, b) __dadd_rn((a), (b))
#define MMH_SUB(a, b) __dsub_rn((a), (b))
#define MMH_DIV(a, b) __ddiv_rn((a), (b))
#else
#define MMH_SIN(k) kSinHost[k]
#define MMH_MUL(a, b) ((a) * (b))
#define MMH_ADD(a, b) ((a) + (b))
#define MMH_SUB(a, b) ((a) - (b))
#define MMH_DIV(a, b) ((a) / (b))
#endif

struct JmBoneTab { int a[kJmBones], b[kJmBones], color[kJmBones]; };
// generic_dataset.py:33-54
MMH_HD void jm_bone(int k, int& a, int& b, int& color) {
  const int A[kJmBones] = {0, 0, 0, 0, 0, 17, 18, 19, 1, 2, 3, 5, 6, 7, 9, 10, 11, 13, 14, 15};
  const int B[kJmBones] = {17, 1, 5, 9, 13, 18, 19, 20, 2, 3, 4, 6, 7, 8, 10, 11, 12, 14, 15, 16};
  const int Cc[kJmBones] = {160, 170, 180, 190, 200, 130, 140, 150, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100, 110, 120};
  a = A[k]; b = B[k]; color = Cc[k];
}

struct JmBone { int cx, cy, ax, angle; double depth; bool skip; };
constexpr double kJmMaxCoord = 1048576.0;   // 2^20

// generic_dataset.py:56-70: centre, half length and angle truncated with int(), depth = mean of the two joints
MMH_HD void jm_params(const double* uv, const double* z, int a, int b, JmBone& o) {
  const double x0 = uv[2 * a], y0 = uv[2 * a + 1], x1 = uv[2 * b], y1 = uv[2 * b + 1];
  o.depth = MMH_DIV(MMH_ADD(z[a], z[b]), 2.0);
  const double mx = MMH_DIV(MMH_ADD(x0, x1), 2.0), my = MMH_DIV(MMH_ADD(y0, y1), 2.0);
  const double dx = MMH_SUB(x0, x1), dy = MMH_SUB(y0, y1);
  const double len = sqrt(MMH_ADD(MMH_MUL(dx, dx), MMH_MUL(dy, dy)));
  // Joints millions of pixels away (or NaN) would overflow OpenCV's integer points and make the scan conversion walk
  // millions of rows above the frame: such a bone is not drawn (the reference is not usable there either).
  o.skip = !(fabs(mx) <= kJmMaxCoord && fabs(my) <= kJmMaxCoord && len <= kJmMaxCoord);
  if (o.skip) { o.cx = o.cy = o.ax = o.angle = 0; return; }
  o.cx = static_cast<int>(mx);
  o.cy = static_cast<int>(my);
  o.ax = static_cast<int>(MMH_DIV(len, 2.0));
  double ang;
  const double adx = fabs(dx), ady = fabs(dy);
  if (dy == 0.0) ang = dx < 0.0 ? 180.0 : 0.0;                     // atan2(+0, x)
  else if (dx == 0.0) ang = dy > 0.0 ? 90.0 : -90.0;
  else if (adx == ady) ang = (dx > 0.0 ? 45.0 : 135.0) * (dy > 0.0 ? 1.0 : -1.0);
  else ang = MMH_MUL(atan2(dy, dx), 180.0 / 3.14159265358979323846);   // math.degrees
  o.angle = static_cast<int>(ang);
}

// cv::ellipse2Poly(center, (ax, 5), angle, 0, 360, 1): point i = 0..360, rounded (cvRound = round half to even)
MMH_HD void jm_ellipse_point(const JmBone& bn, int i, int& px, int& py) {
  int angle = bn.angle;
  while (angle < 0) angle += 360;
  while (angle > 360) angle -= 360;
  const float beta = MMH_SIN(angle), alpha = MMH_SIN(450 - angle);
  const int a = i > 360 ? 360 : i;
  const double x = MMH_MUL(static_cast<double>(bn.ax), static_cast<double>(MMH_SIN(450 - a)));
  const double y = MMH_MUL(5.0, static_cast<double>(MMH_SIN(a)));
  const double fx = MMH_SUB(MMH_ADD(static_cast<double>(bn.cx), MMH_MUL(x, static_cast<double>(alpha))),
                            MMH_MUL(y, static_cast<double>(beta)));
  const double fy = MMH_ADD(MMH_ADD(static_cast<double>(bn.cy), MMH_MUL(x, static_cast<double>(beta))),
                            MMH_MUL(y, static_cast<double>(alpha)));
  px = static_cast<int>(rint(fx));
  py = static_cast<int>(rint(fy));
}

MMH_HD void jm_span_add(int* lo, int* hi, int y, int x) {
#if defined(__CUDA_ARCH__)
  atomicMin(lo + y, x);
  atomicMax(hi + y, x);
#else
  if (x < lo[y]) lo[y] = x;
  if (x > hi[y]) hi[y] = x;
#endif
}

// cv::clipLine(Size2l, Point2l&, Point2l&)
MMH_HD bool jm_clip(int W, int H, int64_t& x1, int64_t& y1, int64_t& x2, int64_t& y2) {
  const int64_t right = W - 1, bottom = H - 1;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    int64_t a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += static_cast<int64_t>(MMH_DIV(MMH_MUL(static_cast<double>(a - y1), static_cast<double>(x2 - x1)),
                                         static_cast<double>(y2 - y1)));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += static_cast<int64_t>(MMH_DIV(MMH_MUL(static_cast<double>(a - y2), static_cast<double>(x2 - x1)),
                                         static_cast<double>(y2 - y1)));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += static_cast<int64_t>(MMH_DIV(MMH_MUL(static_cast<double>(a - x1), static_cast<double>(y2 - y1)),
                                           static_cast<double>(x2 - x1)));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += static_cast<int64_t>(MMH_DIV(MMH_MUL(static_cast<double>(a - x2), static_cast<double>(y2 - y1)),
                                           static_cast<double>(x2 - x1)));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// cv::Line -> LineIterator(connectivity 8, leftToRight = true): every pixel of the segment into the row spans
MMH_HD void jm_line(int W, int H, int ax, int ay, int bx, int by, int* lo, int* hi) {
  int64_t x1 = ax, y1 = ay, x2 = bx, y2 = by;
  const bool inside = x1 >= 0 && x1 < W && x2 >= 0 && x2 < W && y1 >= 0 && y1 < H && y2 >= 0 && y2 < H;
  if (!inside && !jm_clip(W, H, x1, y1, x2, y2)) return;
  int dx = static_cast<int>(x2 - x1), dy = static_cast<int>(y2 - y1);
  int delta_x = 1, delta_y = 1;
  int x = static_cast<int>(x1), y = static_cast<int>(y1);
  if (dx < 0) { dx = -dx; dy = -dy; x = static_cast<int>(x2); y = static_cast<int>(y2); }
  if (dy < 0) { dy = -dy; delta_y = -1; }
  const bool vert = dy > dx;
  if (vert) { int t = dx; dx = dy; dy = t; t = delta_x; delta_x = delta_y; delta_y = t; }
  int err = dx - (dy + dy);
  const int plus_delta = dx + dx, minus_delta = -(dy + dy);
  int minus_shift = delta_x, plus_shift = 0, minus_step = 0, plus_step = delta_y;
  if (vert) { int t = plus_step; plus_step = plus_shift; plus_shift = t; t = minus_step; minus_step = minus_shift; minus_shift = t; }
  const int count = dx + 1;
  for (int i = 0; i < count; ++i) {
    jm_span_add(lo, hi, y, x);
    const bool neg = err < 0;
    err += minus_delta + (neg ? plus_delta : 0);
    x += minus_shift + (neg ? plus_shift : 0);
    y += minus_step + (neg ? plus_step : 0);
  }
}

// cv::FillConvexPoly scan conversion (LINE_8, shift 0) after the outline: sequential edge walk
MMH_HD void jm_fill(int W, int H, const int* px, const int* py, int n, int* lo, int* hi) {
  if (n < 3) return;
  int xmin = px[0], xmax = px[0], ymin = py[0], ymax = py[0], imin = 0;
  for (int i = 0; i < n; ++i) {
    if (py[i] < ymin) { ymin = py[i]; imin = i; }
    if (py[i] > ymax) ymax = py[i];
    if (px[i] > xmax) xmax = px[i];
    if (px[i] < xmin) xmin = px[i];
  }
  if (xmax < 0 || ymax < 0 || xmin >= W || ymin >= H) return;
  if (ymax > H - 1) ymax = H - 1;
  int idx[2] = {imin, imin}, ye[2] = {ymin, ymin};
  const int di[2] = {1, n - 1};
  int64_t ex[2] = {-kJmOne, -kJmOne}, edx[2] = {0, 0};
  int edges = n, y = ymin;
  do {
    for (int i = 0; i < 2; ++i) {
      if (y >= ye[i]) {
        int idx0 = idx[i], j = idx0 + di[i];
        if (j >= n) j -= n;
        for (; edges-- > 0;) {
          const int ty = py[j];
          if (ty > y) {
            const int64_t xs = static_cast<int64_t>(px[idx0]) << kJmShift, xe = static_cast<int64_t>(px[j]) << kJmShift;
            ye[i] = ty;
            edx[i] = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
            ex[i] = xs;
            idx[i] = j;
            break;
          }
          idx0 = j;
          j += di[i];
          if (j >= n) j -= n;
        }
      }
    }
    if (edges < 0) break;
    if (y >= 0) {
      int left = 0, right = 1;
      if (ex[0] > ex[1]) { left = 1; right = 0; }
      int xx1 = static_cast<int>((ex[left] + (kJmOne >> 1)) >> kJmShift);
      int xx2 = static_cast<int>((ex[right] + (kJmOne >> 1)) >> kJmShift);
      if (xx2 >= 0 && xx1 < W) {
        if (xx1 < 0) xx1 = 0;
        if (xx2 >= W) xx2 = W - 1;
        if (xx2 >= xx1) {
          if (xx1 < lo[y]) lo[y] = xx1;
          if (xx2 > hi[y]) hi[y] = xx2;
        }
      }
    }
    ex[0] += edx[0];
    ex[1] += edx[1];
  } while (++y <= ymax);
}

// colour of pixel (x, y): generic_dataset.py:74-77 with the running minimum kept as (depth of the minimum)
MMH_HD int jm_pixel(int x, int y, int H, const int* lo, const int* hi, const double* depth, const int* color) {
  double m = 0.0;
  bool have = false;
  int c = 0;
  for (int b = 0; b < kJmBones; ++b) {
    const bool in = x >= lo[b * H + y] && x <= hi[b * H + y];
    if (in && (!have || depth[b] < m)) { m = depth[b]; have = true; }
    if (have && m == depth[b]) c = color[b];
  }
  return c;
}

#ifndef MMH_HOST_EMU
constexpr int kJmThreads = kJmBones * 32;
__global__ void __launch_bounds__(kJmThreads) jointsmap_kernel(const double* __restrict__ uv, const double* __restrict__ z,
                                                               const int64_t n_pose, const int H, const int W,
                                                               double* __restrict__ out_f64, uint8_t* __restrict__ out_u8) {
  extern __shared__ int jm_smem[];
  int* lo = jm_smem;                         // [bones][H]
  int* hi = lo + kJmBones * H;               // [bones][H]
  int* ptx = hi + kJmBones * H;              // [bones][kJmMaxPts]
  int* pty = ptx + kJmBones * kJmMaxPts;
  __shared__ double s_depth[kJmBones];
  __shared__ int s_color[kJmBones];
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t pose = blockIdx.x; pose < n_pose; pose += gridDim.x) {
    int* blo = lo + warp * H;
    int* bhi = hi + warp * H;
    int* bx = ptx + warp * kJmMaxPts;
    int* by = pty + warp * kJmMaxPts;
    for (int r = lane; r < H; r += 32) { blo[r] = 0x7FFFFFFF; bhi[r] = -1; }
    int a, b, color;
    jm_bone(warp, a, b, color);
    JmBone bn;
    jm_params(uv + pose * 42, z + pose * 21, a, b, bn);
    if (lane == 0) { s_depth[warp] = bn.depth; s_color[warp] = color; }
    for (int i = lane; i <= 360; i += 32) jm_ellipse_point(bn, i, bx[i], by[i]);
    __syncwarp();
    int n = 0;
    if (lane == 0 && !bn.skip) {                         // consecutive duplicates out; a single point = two copies of the centre
      int qx = 0x7FFFFFFF, qy = 0x7FFFFFFF;
      for (int i = 0; i <= 360; ++i) {
        const int x = bx[i], y = by[i];
        if (x != qx || y != qy) { bx[n] = x; by[n] = y; ++n; qx = x; qy = y; }
      }
      if (n == 1) { bx[0] = bx[1] = bn.cx; by[0] = by[1] = bn.cy; n = 2; }
    }
    n = __shfl_sync(0xffffffffu, n, 0);
    __syncwarp();
    for (int j = lane; j < n; j += 32) {     // outline: segment (j-1) -> j, closing segment first (n = 0: bone skipped)
      const int p = j == 0 ? n - 1 : j - 1;
      jm_line(W, H, bx[p], by[p], bx[j], by[j], blo, bhi);
    }
    __syncwarp();
    if (lane == 0) jm_fill(W, H, bx, by, n, blo, bhi);
    __syncthreads();
    const int64_t base = pose * H * W;
    for (int i = threadIdx.x; i < H * W; i += kJmThreads) {
      const int y = i / W, x = i - y * W;
      const int c = jm_pixel(x, y, H, lo, hi, s_depth, s_color);
      if (out_u8 != nullptr) out_u8[base + i] = static_cast<uint8_t>(c);
      if (out_f64 != nullptr) {
        double* o = out_f64 + (base + i) * 3;
        const double v = static_cast<double>(c);
        o[0] = v; o[1] = v; o[2] = v;
      }
    }
    __syncthreads();
  }
}
#endif

}  // namespace mmh

using namespace mmh;

extern "C" int mmh_jointsmap_rasterize(const double* uv, const double* depth, int64_t n_pose, int32_t H, int32_t W,
                                       double* out_f64, uint8_t* out_u8, void* stream) {
  if (n_pose <= 0) return 0;
  MMH_CHECK(uv && depth && (out_f64 || out_u8), "null argument");
  MMH_CHECK(H > 0 && W > 0 && H <= kJmMaxH && W <= 4096, "frame %dx%d unsupported (H <= %d)", H, W, kJmMaxH);
#ifdef MMH_HOST_EMU
  (void)stream;
  static int lo[kJmBones * kJmMaxH], hi[kJmBones * kJmMaxH], px[kJmMaxPts], py[kJmMaxPts];
  for (int64_t pose = 0; pose < n_pose; ++pose) {
    double dep[kJmBones];
    int col[kJmBones];
    for (int k = 0; k < kJmBones; ++k) {
      int* blo = lo + k * H;
      int* bhi = hi + k * H;
      for (int r = 0; r < H; ++r) { blo[r] = 0x7FFFFFFF; bhi[r] = -1; }
      int a, b;
      jm_bone(k, a, b, col[k]);
      JmBone bn;
      jm_params(uv + pose * 42, depth + pose * 21, a, b, bn);
      dep[k] = bn.depth;
      int n = 0, qx = 0x7FFFFFFF, qy = 0x7FFFFFFF;
      for (int i = 0; i <= 360 && !bn.skip; ++i) {
        int x, y;
        jm_ellipse_point(bn, i, x, y);
        if (x != qx || y != qy) { px[n] = x; py[n] = y; ++n; qx = x; qy = y; }
      }
      if (n == 1) { px[0] = px[1] = bn.cx; py[0] = py[1] = bn.cy; n = 2; }
      for (int j = 0; j < n; ++j) {
        const int p = j == 0 ? n - 1 : j - 1;
        jm_line(W, H, px[p], py[p], px[j], py[j], blo, bhi);
      }
      jm_fill(W, H, px, py, n, blo, bhi);
    }
    const int64_t base = pose * H * W;
    for (int i = 0; i < H * W; ++i) {
      const int y = i / W, x = i - y * W;
      const int c = jm_pixel(x, y, H, lo, hi, dep, col);
      if (out_u8 != nullptr) out_u8[base + i] = static_cast<uint8_t>(c);
      if (out_f64 != nullptr) { double* o = out_f64 + (base + i) * 3; o[0] = o[1] = o[2] = static_cast<double>(c); }
    }
  }
  return 0;
#else
  const size_t smem = (static_cast<size_t>(2) * kJmBones * H + 2 * kJmBones * kJmMaxPts) * sizeof(int);
  MMH_CHECK(smem <= 200 * 1024, "frame height %d needs too much shared memory", H);
  static bool attr_set = false;
  if (!attr_set) {
    MMH_CUDA(cudaFuncSetAttribute(jointsmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int64_t cap = static_cast<int64_t>(num_sms()) * 3;
  const int blocks = static_cast<int>(n_pose < cap ? n_pose : cap);
  MMH_CUDA(launch_k(jointsmap_kernel, dim3(blocks), dim3(kJmThreads), smem, stream, uv, depth, n_pose, H, W, out_f64,
                    out_u8));
  return 0;
#endif
}
