// Launch framework of the bandwidth-bound kernels. Each kernel is a functor; the same functor body is compiled
// for the GPU and -- with -DMMH_HOST_EMU, for CPU tests of the index arithmetic only -- as plain loops.
//
//   launch_map            one item per call, grid-stride
//   launch_pg             "pixel x channel-group": a thread keeps its group of 8 channels for its whole life, so the
//                         per-channel parameters (BN coefficients, means, ...) are loaded into registers once
//                         (F::Ctx / prep) and every item is two or three 16-byte vector accesses plus arithmetic
//   launch_reduce_ch      same decomposition for per-channel reductions: registers -> shared memory ->
//                         one atomic per (block, channel)
//   launch_reduce_scalar  warp-shuffle + shared-memory reduction to one atomic per block
//
// Grids are sized in multiples of the SM count.
#pragma once
#include <stdint.h>

#include "../../include/mmhand_sm100.h"
#include "ew_common.h"
#include "host_common.h"

namespace mmh {

// ------------------------------------------------------------------ layout arithmetic
struct LayD {
  int B, H, W, Hg, Wg, h0, w0, phase, ld, c0, C;
  int64_t plane_rows;
};

inline LayD to_layd(const MmhLay& l) {
  LayD d;
  d.B = l.B; d.H = l.H; d.W = l.W; d.Hg = l.Hg; d.Wg = l.Wg; d.h0 = l.h0; d.w0 = l.w0;
  d.phase = l.phase; d.ld = l.ld; d.c0 = l.c0; d.C = l.C;
  d.plane_rows = static_cast<int64_t>(l.B) * l.Hg * l.Wg;
  return d;
}

// element offset of channel c0 of logical pixel (b, h, w); h, w may lie in the halo
MMH_HD int64_t lay_off(const LayD& l, int b, int h, int w) {
  const int hp = h + l.h0, wp = w + l.w0;
  int64_t row;
  if (!l.phase) {
    row = (static_cast<int64_t>(b) * l.Hg + hp) * l.Wg + wp;
  } else {
    row = ((hp & 1) * 2 + (wp & 1)) * l.plane_rows + (static_cast<int64_t>(b) * l.Hg + (hp >> 1)) * l.Wg + (wp >> 1);
  }
  return row * l.ld + l.c0;
}

// reflect index i in [-p, n+p) into [0, n) (torch ReflectionPad2d: no edge repeat)
MMH_HD int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// division by a launch-time constant (n < 2^31): q = umulhi(n, m) >> s
struct FastDiv {
  uint32_t d, m, s;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  if (d <= 1) { f.m = 0; f.s = 0; return f; }
  uint32_t lg = 0;
  while ((1u << lg) < d) ++lg;               // ceil(log2 d)
  const uint32_t p = 31 + lg;
  f.m = static_cast<uint32_t>(((1ull << p) + d - 1) / d);
  f.s = p - 32;
  return f;
}
MMH_HD uint32_t fdiv(uint32_t n, const FastDiv& f) {
#if defined(__CUDA_ARCH__)
  return f.d <= 1 ? n : (__umulhi(n, f.m) >> f.s);
#else
  return f.d <= 1 ? n : n / f.d;
#endif
}

// pixel index over the window [-lo, H+hi) x [-lo, W+hi) of B images -> (b, h, w)
struct PixDec {
  FastDiv we, he;
  int lo;
};
inline PixDec make_pixdec(int H, int W, int lo, int hi) {
  PixDec p;
  p.we = make_fastdiv(W + lo + hi);
  p.he = make_fastdiv(H + lo + hi);
  p.lo = lo;
  return p;
}
MMH_HD void pix_decode(const PixDec& d, uint32_t pix, int& b, int& h, int& w) {
  const uint32_t t = fdiv(pix, d.we);
  w = static_cast<int>(pix - t * d.we.d) - d.lo;
  const uint32_t u = fdiv(t, d.he);
  h = static_cast<int>(t - u * d.he.d) - d.lo;
  b = static_cast<int>(u);
}

// dropout keep bit of logical NCHW element (b, c, h, w): the same hash as oracle/patn_ref.py::dropout_mask
MMH_HD uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
MMH_HD float drop_keep2(uint32_t key, int b, int c, int h, int w, int C, int H, int W) {
  const uint32_t idx = ((static_cast<uint32_t>(b) * C + c) * H + h) * W + w;
  return (mix32(idx * 0x9E3779B1u + key) & 1u) ? 2.0f : 0.0f;
}

// ------------------------------------------------------------------ 8-wide vector access
struct alignas(16) ActX8 { act_t v[8]; };
struct alignas(16) F32x4 { float v[4]; };

MMH_HD void ld8_bf16(const act_t* p, float (&f)[8]) {
  const ActX8 u = *reinterpret_cast<const ActX8*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = act2f(u.v[i]);
}
MMH_HD void st8_bf16(act_t* p, const float (&f)[8]) {
  ActX8 u;
#pragma unroll
  for (int i = 0; i < 8; ++i) u.v[i] = f2act(f[i]);
  *reinterpret_cast<ActX8*>(p) = u;
}
MMH_HD void ld8_f32(const float* p, float (&f)[8]) {
  const F32x4 a = *reinterpret_cast<const F32x4*>(p);
  const F32x4 b = *reinterpret_cast<const F32x4*>(p + 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[i] = a.v[i]; f[4 + i] = b.v[i]; }
}
MMH_HD void st8_f32(float* p, const float (&f)[8]) {
  F32x4 a, b;
#pragma unroll
  for (int i = 0; i < 4; ++i) { a.v[i] = f[i]; b.v[i] = f[4 + i]; }
  *reinterpret_cast<F32x4*>(p) = a;
  *reinterpret_cast<F32x4*>(p + 4) = b;
}
MMH_HD void zero8(float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = 0.f;
}

// ------------------------------------------------------------------ launchers
#ifdef MMH_HOST_EMU

template <class F>
int launch_map(const F& f, int64_t n, void*) {
  for (int64_t i = 0; i < n; ++i) f(i);
  return 0;
}
template <class F>
int launch_pg(const F& f, int64_t n_pix, int groups, void*) {
  MMH_CHECK(n_pix < (int64_t(1) << 31), "too many pixels for one launch");
  for (int g = 0; g < groups; ++g) {
    typename F::Ctx c;
    f.prep(g, c);
    for (uint32_t pix = 0; pix < static_cast<uint32_t>(n_pix); ++pix) f(pix, g, c);
  }
  return 0;
}
// per-channel reduction: item (pixel, group g) adds NV*8 values into out[v*C + g*8 + j]
template <int NV, class F>
int launch_reduce_ch(const F& f, int64_t n_pix, int groups, int C, float* out, void*) {
  MMH_CHECK(n_pix < (int64_t(1) << 31), "too many pixels for one launch");
  for (int g = 0; g < groups; ++g) {
    typename F::Ctx c;
    f.prep(g, c);
    double acc[NV][8];
    for (int v = 0; v < NV; ++v)
      for (int j = 0; j < 8; ++j) acc[v][j] = 0.0;
    for (uint32_t pix = 0; pix < static_cast<uint32_t>(n_pix); ++pix) {
      float a[NV][8];
      for (int v = 0; v < NV; ++v) zero8(a[v]);
      f(pix, g, c, a);
      for (int v = 0; v < NV; ++v)
        for (int j = 0; j < 8; ++j) acc[v][j] += a[v][j];
    }
    for (int v = 0; v < NV; ++v)
      for (int j = 0; j < 8; ++j) out[v * C + g * 8 + j] += static_cast<float>(acc[v][j]);
  }
  return 0;
}
// scalar reduction: item i returns a float, sum added to *out
template <class F>
int launch_reduce_scalar(const F& f, int64_t n, float* out, void*) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += f(i);
  *out += static_cast<float>(s);
  return 0;
}

#else  // CUDA

template <class F>
__global__ void __launch_bounds__(256) map_kernel(const F f, const int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    f(i);
}
template <class F>
int launch_map(const F& f, int64_t n, void* stream) {
  if (n <= 0) return 0;
  const int64_t want = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;   // multiples of the SM count, grid-stride
  const int blocks = static_cast<int>(want < cap ? want : cap);
  map_kernel<F><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(f, n);
  MMH_CUDA(cudaGetLastError());
  return 0;
}

// threads per block: the largest multiple of `groups` <= 256 (a thread never changes its channel group)
inline int pg_threads(int groups) { return (256 / groups) * groups; }

template <class F>
__global__ void __launch_bounds__(256) pg_kernel(const F f, const uint32_t n_pix, const int groups) {
  const uint32_t ppb = blockDim.x / groups;                  // pixels per block and sweep
  const int g = threadIdx.x % groups;
  uint32_t pix = blockIdx.x * ppb + threadIdx.x / groups;
  const uint32_t stride = gridDim.x * ppb;
  typename F::Ctx c;
  f.prep(g, c);
  for (; pix < n_pix; pix += stride) f(pix, g, c);
}
template <class F>
int launch_pg(const F& f, int64_t n_pix, int groups, void* stream) {
  if (n_pix <= 0) return 0;
  MMH_CHECK(groups >= 1 && groups <= 256, "channel groups=%d unsupported", groups);
  MMH_CHECK(n_pix < (int64_t(1) << 31), "too many pixels for one launch");
  const int threads = pg_threads(groups);
  const int ppb = threads / groups;
  const int64_t want = (n_pix + ppb - 1) / ppb;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  const int blocks = static_cast<int>(want < cap ? want : cap);
  pg_kernel<F><<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(f, static_cast<uint32_t>(n_pix), groups);
  MMH_CUDA(cudaGetLastError());
  return 0;
}

template <int NV, class F>
__global__ void __launch_bounds__(256) reduce_ch_kernel(const F f, const uint32_t n_pix, const int groups, const int C,
                                                        float* __restrict__ out) {
  extern __shared__ float red[];   // [ppb][groups][NV*8]
  const uint32_t ppb = blockDim.x / groups;
  const int g = threadIdx.x % groups;
  const uint32_t lr = threadIdx.x / groups;
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) zero8(acc[v]);
  typename F::Ctx c;
  f.prep(g, c);
  for (uint32_t pix = blockIdx.x * ppb + lr; pix < n_pix; pix += gridDim.x * ppb) f(pix, g, c, acc);
  float* mine = red + (static_cast<size_t>(lr) * groups + g) * (NV * 8);
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[v * 8 + j] = acc[v][j];
  __syncthreads();
  // column sums over the ppb pixel lanes: thread t handles (g, v, j) combos round-robin
  const int total = groups * NV * 8;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int gg = idx / (NV * 8), vj = idx % (NV * 8);
    float s = 0.f;
    for (uint32_t l = 0; l < ppb; ++l) s += red[(static_cast<size_t>(l) * groups + gg) * (NV * 8) + vj];
    atomicAdd(out + (vj / 8) * C + gg * 8 + (vj % 8), s);
  }
}
template <int NV, class F>
int launch_reduce_ch(const F& f, int64_t n_pix, int groups, int C, float* out, void* stream) {
  if (n_pix <= 0) return 0;
  MMH_CHECK(groups >= 1 && groups <= 256, "channel groups=%d unsupported", groups);
  MMH_CHECK(n_pix < (int64_t(1) << 31), "too many pixels for one launch");
  const int threads = pg_threads(groups);
  const int ppb = threads / groups;
  const size_t smem = static_cast<size_t>(threads) * NV * 8 * sizeof(float);
  const int64_t want = (n_pix + ppb * 8 - 1) / (ppb * 8);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 4;
  const int blocks = static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
  reduce_ch_kernel<NV, F><<<blocks, threads, smem, static_cast<cudaStream_t>(stream)>>>(
      f, static_cast<uint32_t>(n_pix), groups, C, out);
  MMH_CUDA(cudaGetLastError());
  return 0;
}

template <class F>
__global__ void __launch_bounds__(256) reduce_scalar_kernel(const F f, const int64_t n, float* __restrict__ out) {
  float s = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    s += f(i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = ws[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}
template <class F>
int launch_reduce_scalar(const F& f, int64_t n, float* out, void* stream) {
  if (n <= 0) return 0;
  const int64_t want = (n + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  const int blocks = static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
  reduce_scalar_kernel<F><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(f, n, out);
  MMH_CUDA(cudaGetLastError());
  return 0;
}

#endif

}  // namespace mmh
