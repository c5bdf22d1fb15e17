// Parameter-side kernels: fused Adam over a flat fp32 buffer, weight packing (fp32 OIHW master ->
// bf16 [tap][N][C] tensor-core operand) and packed-gradient unpacking. Dual-mode source.
#include <math.h>
#include <string.h>

#include "ew_framework.h"

namespace mmh {

struct AdamF {
  float* p; const float* g; float* m; float* v;
  float lr_t, b1, b2, eps, inv_sqrt_bc2, gscale;   // lr_t = lr / (1 - b1^t)
  MMH_HD void operator()(int64_t i) const {
    const float gr = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gr;
    const float vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
};

struct PackF {
  const float* src; act_t* dst; int64_t s_n, s_c, s_t; int N, C, Np, Cp;
  MMH_HD void operator()(int64_t i) const {
    const int c = static_cast<int>(i % Cp);
    const int n = static_cast<int>((i / Cp) % Np);
    const int t = static_cast<int>(i / (static_cast<int64_t>(Cp) * Np));
    float v = 0.f;
    if (n < N && c < C) v = src[n * s_n + c * s_c + t * s_t];
    dst[i] = f2act(v);
  }
};

struct PackFoldF {
  const float* src; act_t* dst; int64_t s_n, s_c, s_t; int N, C, kw, Np, Cin_p, Kw, reverse;
  MMH_HD void operator()(int64_t i) const {
    const int cc = static_cast<int>(i % Kw);
    const int n = static_cast<int>((i / Kw) % Np);
    const int t = static_cast<int>(i / (static_cast<int64_t>(Kw) * Np));
    const int j = cc / Cin_p, c = cc % Cin_p;
    float v = 0.f;
    if (n < N && j < kw && c < C) v = src[n * s_n + c * s_c + (t * kw + (reverse ? kw - 1 - j : j)) * s_t];
    dst[i] = f2act(v);
  }
};

struct UnpackF {
  const float* src; float* dst; int64_t s_n, s_c, s_t; int N, C, accumulate;
  MMH_HD void operator()(int64_t i) const {
    const int c = static_cast<int>(i % C);
    const int n = static_cast<int>((i / C) % N);
    const int t = static_cast<int>(i / (static_cast<int64_t>(C) * N));
    const int64_t o = n * s_n + c * s_c + t * s_t;
    dst[o] = accumulate ? dst[o] + src[i] : src[i];
  }
};

struct ParamJobsF {
  const MmhParamJob* jobs; int n_jobs;
  // item = (tile, lane): the job is found by bisection on tile_begin (uniform per block)
  MMH_HD void operator()(int64_t i) const {
    const int tile = static_cast<int>(i >> 8), lane = static_cast<int>(i & 255);
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
    }
    const MmhParamJob& j = jobs[lo];
    const int e = (tile - j.tile_begin) * 256 + lane;        // (n, c) pair, c fastest
    if (j.kind == 0) {
      if (e >= j.Np * j.Cp) return;
      const int n = e / j.Cp, c = e - n * j.Cp;
      const bool live = n < j.N && c < j.C;
      const float* src = static_cast<const float*>(j.src) + n * j.s_n + c * j.s_c;
      act_t* dst = static_cast<act_t*>(j.dst) + e;
      const int64_t plane = static_cast<int64_t>(j.Np) * j.Cp;
      for (int t = 0; t < j.T; ++t) dst[t * plane] = f2act(live ? src[t] : 0.f);
    } else {
      if (e >= j.N * j.C) return;
      const int n = e / j.C, c = e - n * j.C;
      const float* src = static_cast<const float*>(j.src) + e;
      float* dst = static_cast<float*>(j.dst) + n * j.s_n + c * j.s_c;
      const int64_t plane = static_cast<int64_t>(j.N) * j.C;
      for (int t = 0; t < j.T; ++t) dst[t] += src[t * plane];
    }
  }
};

}  // namespace mmh

using namespace mmh;

extern "C" int mmh_param_jobs(const MmhParamJob* jobs, int32_t n_jobs, int32_t total_tiles, void* stream) {
  MMH_CHECK(jobs && n_jobs >= 1 && total_tiles >= 0, "bad argument");
  ParamJobsF f;
  f.jobs = jobs; f.n_jobs = n_jobs;
  return launch_map(f, static_cast<int64_t>(total_tiles) * 256, stream);
}

extern "C" int mmh_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                        float eps, int32_t step, float grad_scale, void* stream) {
  MMH_CHECK(p && g && m && v && step >= 1, "bad argument");
  AdamF f;
  f.p = p; f.g = g; f.m = m; f.v = v;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  f.lr_t = static_cast<float>(lr / bc1);
  f.b1 = beta1; f.b2 = beta2; f.eps = eps;
  f.inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  f.gscale = grad_scale;
  return launch_map(f, n, stream);
}

extern "C" int mmh_pack_weight(const float* src, int64_t s_n, int64_t s_c, int64_t s_t, int32_t N, int32_t C,
                               int32_t T, void* dst, int32_t Np, int32_t Cp, void* stream) {
  MMH_CHECK(src && dst && N <= Np && C <= Cp, "bad argument");
  PackF f;
  f.src = src; f.dst = static_cast<act_t*>(dst); f.s_n = s_n; f.s_c = s_c; f.s_t = s_t;
  f.N = N; f.C = C; f.Np = Np; f.Cp = Cp;
  return launch_map(f, static_cast<int64_t>(T) * Np * Cp, stream);
}

extern "C" int mmh_pack_weight_folded(const float* src, int64_t s_n, int64_t s_c, int64_t s_t, int32_t N, int32_t C,
                                      int32_t kh, int32_t kw, void* dst, int32_t Np, int32_t Cin_p, int32_t Kw,
                                      int32_t reverse, void* stream) {
  MMH_CHECK(src && dst && N <= Np && C <= Cin_p && kw * Cin_p <= Kw, "bad argument");
  PackFoldF f;
  f.src = src; f.dst = static_cast<act_t*>(dst); f.s_n = s_n; f.s_c = s_c; f.s_t = s_t;
  f.N = N; f.C = C; f.kw = kw; f.Np = Np; f.Cin_p = Cin_p; f.Kw = Kw; f.reverse = reverse;
  return launch_map(f, static_cast<int64_t>(kh) * Np * Kw, stream);
}

extern "C" int mmh_unpack_wgrad(const float* src, float* dst, int64_t s_n, int64_t s_c, int64_t s_t, int32_t N,
                                int32_t C, int32_t T, int32_t accumulate, void* stream) {
  MMH_CHECK(src && dst, "null argument");
  UnpackF f;
  f.src = src; f.dst = dst; f.s_n = s_n; f.s_c = s_c; f.s_t = s_t; f.N = N; f.C = C; f.accumulate = accumulate;
  return launch_map(f, static_cast<int64_t>(T) * N * C, stream);
}

extern "C" int mmh_memset(void* p, int32_t byte, int64_t bytes, void* stream) {
  MMH_CHECK(p || bytes == 0, "null argument");
#ifdef MMH_HOST_EMU
  (void)stream;
  memset(p, byte, static_cast<size_t>(bytes));
#else
  MMH_CUDA(cudaMemsetAsync(p, byte, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream)));
#endif
  return 0;
}
