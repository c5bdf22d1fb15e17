// Launch framework of the bandwidth-bound kernels. Each kernel is a functor; the same functor body is compiled
// for the GPU and -- with -DMMH_HOST_EMU, for CPU tests of the index arithmetic only -- as plain loops.
//
//   launch_map            one item per call, grid-stride
//   launch_pg             "pixel x channel-group": a thread keeps its group of 8 channels for its whole life, so the
//                         per-channel parameters (BN coefficients, means, ...) are loaded into registers once
//                         (F::Ctx / prep) and every item is two or three 16-byte vector accesses plus arithmetic.
//                         Functors are two-phase -- load(pix, g, ctx, In&) then finish(In, g, ctx) -- and the kernel
//                         issues the loads of kPgUnroll items before finishing any of them: with one 16-byte load in
//                         flight per thread HBM sits at ~30 % (measured), with four it is bandwidth-bound
//   launch_reduce_ch      same decomposition for per-channel reductions: registers -> shared memory ->
//                         one atomic per (block, channel)
//   launch_reduce_scalar  warp-shuffle + shared-memory reduction to one atomic per block
//
// Grids are sized in multiples of the SM count.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef MMH_HOST_EMU
#include <cuda_bf16.h>
#endif

#include "../../include/mmhand_sm100.h"
#include "ew_common.h"
#include "host_common.h"

namespace mmh {

// ------------------------------------------------------------------ layout arithmetic
struct LayD {
  int B, H, W, Hg, Wg, h0, w0, phase, ld, c0, C;
  int plane_rows;     // rows of one parity plane; every grid has < 2^31 rows (checked by the launchers)
};

inline LayD to_layd(const MmhLay& l) {
  LayD d;
  d.B = l.B; d.H = l.H; d.W = l.W; d.Hg = l.Hg; d.Wg = l.Wg; d.h0 = l.h0; d.w0 = l.w0;
  d.phase = l.phase; d.ld = l.ld; d.c0 = l.c0; d.C = l.C;
  d.plane_rows = l.B * l.Hg * l.Wg;
  return d;
}

// element offset of channel c0 of logical pixel (b, h, w); h, w may lie in the halo
MMH_HD int64_t lay_off(const LayD& l, int b, int h, int w) {
  const int hp = h + l.h0, wp = w + l.w0;
  int row;            // 32-bit row arithmetic, one widening multiply at the end
  if (!l.phase) {
    row = (b * l.Hg + hp) * l.Wg + wp;
  } else {
    row = ((hp & 1) * 2 + (wp & 1)) * l.plane_rows + (b * l.Hg + (hp >> 1)) * l.Wg + (wp >> 1);
  }
  return static_cast<int64_t>(row) * l.ld + l.c0;
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

// Iteration space of the pixel kernels: B images x rows [-lo, H+hi) x columns [-lo, W+hi). A block works on a
// run of consecutive columns of ONE row, so that image and row are block-uniform and cost nothing per item.
struct RowGeom {
  int n_rows;     // B * (H + lo + hi)   (1 for flat row lists)
  int h_ext;      // H + lo + hi
  int n_cols;     // W + lo + hi         (number of rows for flat row lists)
  int lo;
};
inline RowGeom make_rowgeom(int B, int H, int W, int lo, int hi) {
  RowGeom r;
  r.n_rows = B * (H + lo + hi);
  r.h_ext = H + lo + hi;
  r.n_cols = W + lo + hi;
  r.lo = lo;
  return r;
}
inline RowGeom make_flatgeom(int64_t rows) {
  RowGeom r;
  r.n_rows = 1; r.h_ext = 1; r.n_cols = static_cast<int>(rows); r.lo = 0;
  return r;
}

// dropout keep bits of logical NCHW elements: the same hash as oracle/patn_ref.py::dropout_mask
MMH_HD uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
// One hash per (pixel, group of 8 channels): bit j of the word decides channel 8*g + j.
MMH_HD uint32_t drop_bits(uint32_t key, int b, int h, int w, int g, int C, int H, int W) {
  const uint32_t word = ((static_cast<uint32_t>(b) * H + h) * W + w) * static_cast<uint32_t>((C + 7) / 8) + g;
  return mix32(word * 0x9E3779B1u + key);
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
#if defined(__CUDA_ARCH__) && !defined(MMH_EMU_F32)
  // cvt.rn.bf16x2.f32: same round-to-nearest-even as f2bf
  uint4 q;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(f[0], f[1]); q.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[2], f[3]); q.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[4], f[5]); q.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[6], f[7]); q.w = *reinterpret_cast<uint32_t*>(&t);
  *reinterpret_cast<uint4*>(p) = q;
  return;
#endif
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
struct F32x8 { F32x4 a, b; };
MMH_HD void ld_raw(const act_t* p, ActX8& u) { u = *reinterpret_cast<const ActX8*>(p); }
MMH_HD void ld_raw(const float* p, F32x8& u) {
  u.a = *reinterpret_cast<const F32x4*>(p);
  u.b = *reinterpret_cast<const F32x4*>(p + 4);
}
MMH_HD void cvt8(const ActX8& u, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = act2f(u.v[i]);
}
MMH_HD void cvt8(const F32x8& u, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[i] = u.a.v[i]; f[4 + i] = u.b.v[i]; }
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
int launch_pg(const F& f, const RowGeom& rg, int groups, void*) {
  for (int g = 0; g < groups; ++g) {
    typename F::Ctx c;
    f.prep(g, c);
    for (int row = 0; row < rg.n_rows; ++row)
      for (int col = 0; col < rg.n_cols; ++col) {
        const int b = row / rg.h_ext, h = row % rg.h_ext - rg.lo, w = col - rg.lo;
        typename F::In in;
        f.load(b, h, w, g, c, in);
        f.finish(in, b, h, w, g, c);
      }
  }
  return 0;
}
// per-channel reduction: item (pixel, group g) adds NV*8 values into out[v*C + g*8 + j]
template <int NV, class F>
int launch_reduce_ch(const F& f, const RowGeom& rg, int groups, int C, float* out, void*) {
  for (int g = 0; g < groups; ++g) {
    typename F::Ctx c;
    f.prep(g, c);
    double acc[NV][8];
    for (int v = 0; v < NV; ++v)
      for (int j = 0; j < 8; ++j) acc[v][j] = 0.0;
    for (int row = 0; row < rg.n_rows; ++row)
      for (int col = 0; col < rg.n_cols; ++col) {
        const int b = row / rg.h_ext, h = row % rg.h_ext - rg.lo, w = col - rg.lo;
        float a[NV][8];
        for (int v = 0; v < NV; ++v) zero8(a[v]);
        typename F::In in;
        f.load(b, h, w, g, c, in);
        f.accum(in, b, h, w, g, c, a);
        for (int v = 0; v < NV; ++v)
          for (int j = 0; j < 8; ++j) acc[v][j] += a[v][j];
      }
    for (int v = 0; v < NV; ++v)
      for (int j = 0; j < 8; ++j) out[v * C + g * 8 + j] += static_cast<float>(acc[v][j]);
  }
  return 0;
}
// reduction + finalisation in one launch: fin(c) consumes the accumulators of channel c, which return to zero
template <int NV, class F, class Fin>
int launch_reduce_ch_fin(const F& f, const RowGeom& rg, int groups, int C, float* out, const Fin& fin, uint32_t*,
                         void* stream) {
  launch_reduce_ch<NV>(f, rg, groups, C, out, stream);
  for (int c = 0; c < C; ++c) {
    fin(c);
    for (int v = 0; v < NV; ++v) out[v * C + c] = 0.f;
  }
  return 0;
}
// ---- row launchers (lean kernels): the functor computes the base offsets of an (image, row) once (F::Row); items
// are addressed as base + column * pitch
template <class F>
int launch_rows_pg(const F& f, const RowGeom& rg, int groups, int /*reverse*/, void*) {
  for (int g = 0; g < groups; ++g) {
    typename F::Ctx c;
    f.prep(g, c);
    for (int row = 0; row < rg.n_rows; ++row) {
      const int b = row / rg.h_ext, h = row % rg.h_ext - rg.lo;
      typename F::Row r;
      f.row(b, h, r);
      for (int col = 0; col < rg.n_cols; ++col) {
        typename F::In in;
        f.load(r, col - rg.lo, g, c, in);
        f.finish(in, r, col - rg.lo, g, c);
      }
    }
  }
  return 0;
}
template <int NV, class F>
int launch_rows_reduce(const F& f, const RowGeom& rg, int groups, int C, float* out, int /*reverse*/, void*) {
  for (int g = 0; g < groups; ++g) {
    typename F::Ctx c;
    f.prep(g, c);
    // per-row partial sums in fp32 (as a GPU block does), rows combined in double
    double acc[NV][8];
    for (int v = 0; v < NV; ++v)
      for (int j = 0; j < 8; ++j) acc[v][j] = 0.0;
    for (int row = 0; row < rg.n_rows; ++row) {
      const int b = row / rg.h_ext, h = row % rg.h_ext - rg.lo;
      typename F::Row r;
      f.row(b, h, r);
      float a[NV][8];
      for (int v = 0; v < NV; ++v) zero8(a[v]);
      for (int col = 0; col < rg.n_cols; ++col) {
        typename F::In in;
        f.load(r, col - rg.lo, g, c, in);
        f.accum(in, r, col - rg.lo, g, c, a);
      }
      f.post(g, c, a);
      for (int v = 0; v < NV; ++v)
        for (int j = 0; j < 8; ++j) acc[v][j] += a[v][j];
    }
    for (int v = 0; v < NV; ++v)
      for (int j = 0; j < 8; ++j) out[v * C + g * 8 + j] += static_cast<float>(acc[v][j]);
  }
  return 0;
}
template <int NV, class F, class Fin>
int launch_rows_reduce_fin(const F& f, const RowGeom& rg, int groups, int C, float* out, const Fin& fin, uint32_t*,
                           int reverse, void* stream) {
  launch_rows_reduce<NV>(f, rg, groups, C, out, reverse, stream);
  for (int c = 0; c < C; ++c) {
    fin(c);
    for (int v = 0; v < NV; ++v) out[v * C + c] = 0.f;
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

// Programmatic dependent launch: every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and starts with pdl_sync() -- wait until the preceding grid of the
// stream has completed and flushed (griddepcontrol.wait), then allow the next grid of the stream to be scheduled
// (griddepcontrol.launch_dependents). The next kernel's launch latency, block scheduling and prologue thus overlap
// this kernel's execution; at most two grids of a stream are in flight. ~800 dependent launches per training step.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <class F>
__global__ void __launch_bounds__(256) map_kernel(const F f, const int64_t n) {
  pdl_sync();
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
  MMH_CUDA(launch_k(map_kernel<F>, dim3(blocks), dim3(256), 0, stream, f, n));
  return 0;
}

// grid = (column split, rows per image incl. halo, images); the column split grows until there are about
// `per_sm` blocks per SM (flat row lists: rows = images = 1 and the split is the whole grid)
inline dim3 row_grid(const RowGeom& rg, int max_split, int per_sm) {
  const int64_t rows = rg.n_rows;
  int64_t split = (static_cast<int64_t>(num_sms()) * per_sm + rows - 1) / rows;
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  return dim3(static_cast<unsigned>(split), static_cast<unsigned>(rg.h_ext), static_cast<unsigned>(rg.n_rows / rg.h_ext));
}

// threads per block: the largest multiple of `groups` <= the block size (a thread never changes its channel group).
// Block size 128 (default; MMH_EW_THREADS=256 for the old setting): four blocks per SM alone, and three of them
// (instead of one of 256 threads) fit in the registers left by a resident weight-gradient CTA of the side stream
// (measured: 55.3 -> 54.5 ms per training step with the side stream, no change without it).
inline int ew_block_threads() {
  static const int t = [] {
    const char* e = getenv("MMH_EW_THREADS");
    const int v = e != nullptr ? atoi(e) : 128;
    return v == 256 ? 256 : 128;
  }();
  return t;
}
// resident wave of the persistent reduction kernels, in blocks per SM (MMH_REDUCE_WAVE; default: as many as fit alone)
inline int reduce_wave_per_sm() {
  static const int w = [] {
    const char* e = getenv("MMH_REDUCE_WAVE");
    const int v = e != nullptr ? atoi(e) : 0;
    return v > 0 && v <= 8 ? v : 512 / ew_block_threads();
  }();
  return w;
}
inline int pg_threads(int groups) {
  const int t = groups > ew_block_threads() ? 256 : ew_block_threads();
  return (t / groups) * groups;
}


// resident blocks per SM the compiler must allow for the pixel kernels (register cap = 65536 / (256 * MMH_EW_MINBLOCKS))
#ifndef MMH_EW_MINBLOCKS
#define MMH_EW_MINBLOCKS 2
#endif

template <class F>
__global__ void __launch_bounds__(256, MMH_EW_MINBLOCKS) pg_kernel(const F f, const RowGeom rg, const int groups, const int chunks) {
  pdl_sync();
  constexpr int U = F::kUnroll;
  const int ppb = blockDim.x / groups;                       // pixels per block and sweep
  const int g = threadIdx.x % groups;
  const int lr = threadIdx.x / groups;
  typename F::Ctx c;
  f.prep(g, c);
  // grid = (column split, row, image): image and row come from special registers, i.e. they are uniform and
  // every address term that depends on them is computed once per block in the uniform datapath
  const int b = blockIdx.z, h = static_cast<int>(blockIdx.y) - rg.lo;
  for (int chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
    const int col0 = chunk * (ppb * U) + lr;
    typename F::In in[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int col = col0 + u * ppb;
      if (col < rg.n_cols) f.load(b, h, col - rg.lo, g, c, in[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int col = col0 + u * ppb;
      if (col < rg.n_cols) f.finish(in[u], b, h, col - rg.lo, g, c);
    }
  }
}
template <class F>
int launch_pg(const F& f, const RowGeom& rg, int groups, void* stream) {
  if (rg.n_rows <= 0 || rg.n_cols <= 0) return 0;
  MMH_CHECK(groups >= 1 && groups <= 256, "channel groups=%d unsupported", groups);
  const int threads = pg_threads(groups);
  const int ppb = threads / groups;
  const int chunks = (rg.n_cols + ppb * F::kUnroll - 1) / (ppb * F::kUnroll);
  const dim3 grid = row_grid(rg, chunks, 8);       // (4 or 16 blocks per SM: no gain, profiles/r02_norm_act_knobs.txt)
  MMH_CUDA(launch_k(pg_kernel<F>, grid, dim3(threads), 0, stream, f, rg, groups, chunks));
  return 0;
}

// block-level end of a per-channel reduction: registers -> shared memory -> one atomic per (block, channel)
template <int NV>
__device__ __forceinline__ void reduce_ch_tail(const float (&acc)[NV][8], const int groups, const int ppb, const int g,
                                               const int lr, const int C, float* __restrict__ out, float* red) {
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
    for (int l = 0; l < ppb; ++l) s += red[(static_cast<size_t>(l) * groups + gg) * (NV * 8) + vj];
    atomicAdd(out + (vj / 8) * C + gg * 8 + (vj % 8), s);
  }
}
// the block that takes the last ticket finalises the channels and resets accumulators and counter
template <int NV, class Fin>
__device__ __forceinline__ void reduce_ch_last_block(const Fin& fin, const int C, float* out, uint32_t* counter) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    fin(c);
#pragma unroll
    for (int v = 0; v < NV; ++v) out[v * C + c] = 0.f;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

// Persistent: the grid is one resident wave (2 blocks per SM) and a block walks over (row, column chunk) units, so
// that a launch ends with 2 * SMs * NV * C atomics on the NV * C result words (one block per row used to mean
// B * H blocks hammering the same few cache lines: measured 41 % of HBM speed on the 512-channel layers).
template <int NV, class F>
__device__ __forceinline__ void reduce_ch_body(const F& f, const RowGeom& rg, const int groups, const int chunks,
                                               const int C, float* __restrict__ out, float* red) {
  constexpr int U = F::kUnroll;
  const int ppb = blockDim.x / groups;
  const int g = threadIdx.x % groups;
  const int lr = threadIdx.x / groups;
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) zero8(acc[v]);
  typename F::Ctx c;
  f.prep(g, c);
  const int units = rg.n_rows * chunks;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    // block-uniform: row, image and chunk of this unit
    const int row = unit / chunks, chunk = unit - row * chunks;
    const int b = row / rg.h_ext, h = row - b * rg.h_ext - rg.lo;
    const int col0 = chunk * (ppb * U) + lr;
    typename F::In in[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int col = col0 + u * ppb;
      if (col < rg.n_cols) f.load(b, h, col - rg.lo, g, c, in[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int col = col0 + u * ppb;
      if (col < rg.n_cols) f.accum(in[u], b, h, col - rg.lo, g, c, acc);
    }
  }
  reduce_ch_tail<NV>(acc, groups, ppb, g, lr, C, out, red);
}
template <int NV, class F>
__global__ void __launch_bounds__(256, MMH_EW_MINBLOCKS) reduce_ch_kernel(const F f, const RowGeom rg, const int groups, const int chunks,
                                                           const int C, float* __restrict__ out) {
  extern __shared__ float red[];   // [ppb][groups][NV*8]
  pdl_sync();
  reduce_ch_body<NV>(f, rg, groups, chunks, C, out, red);
}
// Same reduction; the block that finishes last (ticket counter) finalises: fin(c) reads the accumulators of channel
// c (complete: every other block fenced before taking its ticket), does the per-channel arithmetic -- and the
// cross-GPU exchange when there is one -- and the accumulators and the counter go back to zero for the next launch.
// One launch instead of memset + reduction + finalise.
template <int NV, class F, class Fin>
__global__ void __launch_bounds__(256, MMH_EW_MINBLOCKS) reduce_ch_fin_kernel(const F f, const RowGeom rg, const int groups,
                                                               const int chunks, const int C, float* out,
                                                               const Fin fin, uint32_t* counter) {
  extern __shared__ float red[];
  pdl_sync();
  reduce_ch_body<NV>(f, rg, groups, chunks, C, out, red);
  reduce_ch_last_block<NV>(fin, C, out, counter);
}
template <int NV, class F>
int launch_reduce_ch(const F& f, const RowGeom& rg, int groups, int C, float* out, void* stream) {
  if (rg.n_rows <= 0 || rg.n_cols <= 0) return 0;
  MMH_CHECK(groups >= 1 && groups <= 256, "channel groups=%d unsupported", groups);
  const int threads = pg_threads(groups);
  const int ppb = threads / groups;
  const size_t smem = static_cast<size_t>(threads) * NV * 8 * sizeof(float);
  const int chunks = (rg.n_cols + ppb * F::kUnroll - 1) / (ppb * F::kUnroll);
  const int64_t units = static_cast<int64_t>(rg.n_rows) * chunks;
  MMH_CHECK(units < (int64_t(1) << 31), "too many work units");
  const int64_t wave = static_cast<int64_t>(num_sms()) * reduce_wave_per_sm();
  const int blocks = static_cast<int>(units < wave ? units : wave);
  MMH_CUDA(launch_k(reduce_ch_kernel<NV, F>, dim3(blocks), dim3(threads), smem, stream, f, rg, groups, chunks, C, out));
  return 0;
}

template <int NV, class F, class Fin>
int launch_reduce_ch_fin(const F& f, const RowGeom& rg, int groups, int C, float* out, const Fin& fin,
                         uint32_t* counter, void* stream) {
  if (rg.n_rows <= 0 || rg.n_cols <= 0) return 0;
  MMH_CHECK(groups >= 1 && groups <= 256, "channel groups=%d unsupported", groups);
  MMH_CHECK(counter != nullptr, "null ticket counter");
  const int threads = pg_threads(groups);
  const int ppb = threads / groups;
  const size_t smem = static_cast<size_t>(threads) * NV * 8 * sizeof(float);
  const int chunks = (rg.n_cols + ppb * F::kUnroll - 1) / (ppb * F::kUnroll);
  const int64_t units = static_cast<int64_t>(rg.n_rows) * chunks;
  MMH_CHECK(units < (int64_t(1) << 31), "too many work units");
  const int64_t wave = static_cast<int64_t>(num_sms()) * reduce_wave_per_sm();
  const int blocks = static_cast<int>(units < wave ? units : wave);
  MMH_CUDA(launch_k(reduce_ch_fin_kernel<NV, F, Fin>, dim3(blocks), dim3(threads), smem, stream, f, rg, groups, chunks,
                    C, out, fin, counter));
  return 0;
}

// ------------------------------------------------------------------ row kernels (lean variants of the hot functors)
// Persistent blocks walk over units = (image row, chunk of columns). Everything that depends on (image, row) only --
// base offsets of the row in every tensor, border flags, the dropout counter -- is computed once per unit (F::Row,
// block-uniform); an item is base + column * pitch. The per-channel-group parameters are 16-byte loads (prep).
// `reverse` walks the units from the last to the first: the kernel that follows a producer (or an earlier sweep over
// the same tensors) starts with the rows that were touched last and are still in the 126 MB L2.
struct RowSched {
  int units, chunks, reverse;
  FastDiv d_chunks, d_hext;
};
inline RowSched make_rowsched(const RowGeom& rg, int chunks, int reverse) {
  RowSched s;
  s.units = rg.n_rows * chunks; s.chunks = chunks; s.reverse = reverse;
  s.d_chunks = make_fastdiv(static_cast<uint32_t>(chunks));
  s.d_hext = make_fastdiv(static_cast<uint32_t>(rg.h_ext));
  return s;
}
#ifndef MMH_ROWS_MINBLOCKS
#define MMH_ROWS_MINBLOCKS 4
#endif
// read per call (tests and A/B measurements switch inside one process)
inline int rows_wave_per_sm() {
  const char* e = getenv("MMH_ROWS_WAVE");
  const int v = e != nullptr ? atoi(e) : 0;
  return v > 0 && v <= 16 ? v : MMH_ROWS_MINBLOCKS;
}
inline int rows_reverse_enabled() {
  const char* e = getenv("MMH_EW_REVERSE");
  return e == nullptr || atoi(e) != 0;
}

// One unit of a row kernel: decode (block-uniform), row setup, loads of U items
template <class F>
struct RowUnit {
  typename F::Row r;
  typename F::In in[F::kUnroll];
  int col0;
  __device__ __forceinline__ void issue(const F& f, const typename F::Ctx& c, const RowGeom& rg, const RowSched& sc,
                                        const int u0, const int ppb, const int lr, const int g) {
    constexpr int U = F::kUnroll;
    const int unit = sc.reverse ? sc.units - 1 - u0 : u0;
    const int row = static_cast<int>(fdiv(static_cast<uint32_t>(unit), sc.d_chunks)), chunk = unit - row * sc.chunks;
    const int b = static_cast<int>(fdiv(static_cast<uint32_t>(row), sc.d_hext)), h = row - b * rg.h_ext - rg.lo;
    f.row(b, h, r);
    col0 = chunk * (ppb * U) + lr;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int col = col0 + u * ppb;
      if (col < rg.n_cols) f.load(r, col - rg.lo, g, c, in[u]);
    }
  }
};
template <class F>
__global__ void __launch_bounds__(128, MMH_ROWS_MINBLOCKS) rows_pg_kernel(const F f, const RowGeom rg, const int groups,
                                                                           const RowSched sc) {
  pdl_sync();
  constexpr int U = F::kUnroll;
  const int ppb = blockDim.x / groups;
  const int g = threadIdx.x % groups;
  const int lr = threadIdx.x / groups;
  typename F::Ctx c;
  f.prep(g, c);
  for (int u0 = blockIdx.x; u0 < sc.units; u0 += gridDim.x) {
    RowUnit<F> a;
    a.issue(f, c, rg, sc, u0, ppb, lr, g);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int col = a.col0 + u * ppb;
      if (col < rg.n_cols) f.finish(a.in[u], a.r, col - rg.lo, g, c);
    }
  }
}
template <int NV, class F>
__device__ __forceinline__ void rows_reduce_body(const F& f, const RowGeom& rg, const int groups, const RowSched& sc,
                                                 const int C, float* __restrict__ out, float* red) {
  constexpr int U = F::kUnroll;
  const int ppb = blockDim.x / groups;
  const int g = threadIdx.x % groups;
  const int lr = threadIdx.x / groups;
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) zero8(acc[v]);
  typename F::Ctx c;
  f.prep(g, c);
  for (int u0 = blockIdx.x; u0 < sc.units; u0 += gridDim.x) {
    RowUnit<F> a;
    a.issue(f, c, rg, sc, u0, ppb, lr, g);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int col = a.col0 + u * ppb;
      if (col < rg.n_cols) f.accum(a.in[u], a.r, col - rg.lo, g, c, acc);
    }
  }
  f.post(g, c, acc);
  reduce_ch_tail<NV>(acc, groups, ppb, g, lr, C, out, red);
}
template <int NV, class F>
__global__ void __launch_bounds__(128, MMH_ROWS_MINBLOCKS) rows_reduce_kernel(const F f, const RowGeom rg, const int groups,
                                                                               const RowSched sc, const int C,
                                                                               float* __restrict__ out) {
  extern __shared__ float red[];   // [ppb][groups][NV*8]
  pdl_sync();
  rows_reduce_body<NV>(f, rg, groups, sc, C, out, red);
}
template <int NV, class F, class Fin>
__global__ void __launch_bounds__(128, MMH_ROWS_MINBLOCKS) rows_reduce_fin_kernel(const F f, const RowGeom rg,
                                                                                   const int groups, const RowSched sc,
                                                                                   const int C, float* out, const Fin fin,
                                                                                   uint32_t* counter) {
  extern __shared__ float red[];
  pdl_sync();
  rows_reduce_body<NV>(f, rg, groups, sc, C, out, red);
  reduce_ch_last_block<NV>(fin, C, out, counter);
}
// threads: the largest multiple of `groups` <= 128; chunks of ppb * U columns
struct RowsLaunch { int threads, ppb, chunks, blocks; RowSched sc; };
template <class F>
inline RowsLaunch rows_launch(const RowGeom& rg, int groups, int reverse) {
  RowsLaunch l;
  l.threads = (128 / groups) * groups;
  l.ppb = l.threads / groups;
  l.chunks = (rg.n_cols + l.ppb * F::kUnroll - 1) / (l.ppb * F::kUnroll);
  l.sc = make_rowsched(rg, l.chunks, reverse && rows_reverse_enabled());
  const int64_t wave = static_cast<int64_t>(num_sms()) * rows_wave_per_sm();
  l.blocks = static_cast<int>(l.sc.units < wave ? l.sc.units : wave);
  return l;
}
template <class F>
int launch_rows_pg(const F& f, const RowGeom& rg, int groups, int reverse, void* stream) {
  if (rg.n_rows <= 0 || rg.n_cols <= 0) return 0;
  MMH_CHECK(groups >= 1 && groups <= 128, "channel groups=%d unsupported by the row kernels", groups);
  MMH_CHECK(static_cast<int64_t>(rg.n_rows) * rg.n_cols < (int64_t(1) << 30), "too many work units");
  const RowsLaunch l = rows_launch<F>(rg, groups, reverse);
  MMH_CUDA(launch_k(rows_pg_kernel<F>, dim3(l.blocks), dim3(l.threads), 0, stream, f, rg, groups, l.sc));
  return 0;
}
template <int NV, class F>
int launch_rows_reduce(const F& f, const RowGeom& rg, int groups, int C, float* out, int reverse, void* stream) {
  if (rg.n_rows <= 0 || rg.n_cols <= 0) return 0;
  MMH_CHECK(groups >= 1 && groups <= 128, "channel groups=%d unsupported by the row kernels", groups);
  MMH_CHECK(static_cast<int64_t>(rg.n_rows) * rg.n_cols < (int64_t(1) << 30), "too many work units");
  const RowsLaunch l = rows_launch<F>(rg, groups, reverse);
  const size_t smem = static_cast<size_t>(l.threads) * NV * 8 * sizeof(float);
  MMH_CUDA(launch_k(rows_reduce_kernel<NV, F>, dim3(l.blocks), dim3(l.threads), smem, stream, f, rg, groups, l.sc, C, out));
  return 0;
}
template <int NV, class F, class Fin>
int launch_rows_reduce_fin(const F& f, const RowGeom& rg, int groups, int C, float* out, const Fin& fin,
                           uint32_t* counter, int reverse, void* stream) {
  if (rg.n_rows <= 0 || rg.n_cols <= 0) return 0;
  MMH_CHECK(groups >= 1 && groups <= 128, "channel groups=%d unsupported by the row kernels", groups);
  MMH_CHECK(static_cast<int64_t>(rg.n_rows) * rg.n_cols < (int64_t(1) << 30), "too many work units");
  MMH_CHECK(counter != nullptr, "null ticket counter");
  const RowsLaunch l = rows_launch<F>(rg, groups, reverse);
  const size_t smem = static_cast<size_t>(l.threads) * NV * 8 * sizeof(float);
  MMH_CUDA(launch_k(rows_reduce_fin_kernel<NV, F, Fin>, dim3(l.blocks), dim3(l.threads), smem, stream, f, rg, groups,
                    l.sc, C, out, fin, counter));
  return 0;
}

template <class F>
__global__ void __launch_bounds__(256) reduce_scalar_kernel(const F f, const int64_t n, float* __restrict__ out) {
  pdl_sync();
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
  MMH_CUDA(launch_k(reduce_scalar_kernel<F>, dim3(blocks), dim3(256), 0, stream, f, n, out));
  return 0;
}

#endif

}  // namespace mmh
