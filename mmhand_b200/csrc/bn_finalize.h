// BatchNorm finalisation functors (one work item = one channel), shared by the plain launches in elementwise.cu
// and by the launches fused with the cross-GPU statistics exchange in peer.cu.
#pragma once
#include <math.h>
#include <stdint.h>

#include "ew_common.h"

namespace mmh {

struct BnFinalizeF {
  const float* sums; const float* gamma; const float* beta; float* rm; float* rv; float* coef; float* save;
  float count, momentum, eps; int train, C;
  MMH_HD void operator()(int64_t c) const { apply(c, train ? sums[c] : 0.f, train ? sums[C + c] : 0.f); }
  // s0 = sum x, s1 = sum x^2 of channel c (over all ranks)
  MMH_HD void apply(int64_t c, float s0, float s1) const {
    float mean, var;
    if (train) {
      mean = s0 / count;
      var = s1 / count - mean * mean;
      if (var < 0.f) var = 0.f;
      if (rm != nullptr) {
        const float unb = count > 1.f ? var * count / (count - 1.f) : var;
        rm[c] = (1.f - momentum) * rm[c] + momentum * mean;
        rv[c] = (1.f - momentum) * rv[c] + momentum * unb;
      }
    } else {
      mean = rm[c];
      var = rv[c];
    }
    const float rstd = 1.0f / sqrtf(var + eps);
    const float ga = gamma != nullptr ? gamma[c] : 1.f, be = beta != nullptr ? beta[c] : 0.f;
    coef[c] = ga * rstd;
    coef[C + c] = be - mean * ga * rstd;
    save[c] = mean;
    save[C + c] = rstd;
  }
};

struct BnBwdFinalizeF {
  const float* sg; const float* sl; float* k; float* dgamma; float* dbeta; float count; int C;
  MMH_HD void operator()(int64_t c) const { apply(c, sg[c], sg[C + c], sl[c], sl[C + c]); }
  // g0, g1 = (sum dz, sum dz * xhat) over all ranks; l0, l1 = this rank's part
  MMH_HD void apply(int64_t c, float g0, float g1, float l0, float l1) const {
    k[c] = g0 / count;
    k[C + c] = g1 / count;
    if (dgamma != nullptr) dgamma[c] += l1;
    if (dbeta != nullptr) dbeta[c] += l0;
  }
};

}  // namespace mmh
