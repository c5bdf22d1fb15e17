// Public plan entry points of the tensor-core kernels (include/mmhand_sm100.h): argument checks, then the
// shifted-row convolution (tc_conv2.cu) or weight-gradient (tc_wgrad2.cu) plan behind an opaque handle.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mmhand_sm100.h"
#include "conv_plan.h"
#include "host_common.h"

struct MmhConvPlan {
  MmhConv2* v2 = nullptr;
  ~MmhConvPlan() { if (v2) mmh_conv2_destroy(v2); }
};
struct MmhWgradPlan {
  MmhWgrad2* v2 = nullptr;
  ~MmhWgradPlan() { if (v2) mmh_wgrad2_destroy(v2); }
};

extern "C" int mmh_conv_plan_create(const MmhConvDesc* d, MmhConvPlan** out_plan) {
  MMH_CHECK(d && out_plan, "null argument");
  MMH_CHECK(d->T >= 1 && d->T <= MMH_MAX_TAPS, "T=%d out of range", d->T);
  MMH_CHECK(d->C >= 16 && (d->C % 16) == 0, "C=%d must be a multiple of 16", d->C);
  MMH_CHECK(d->C == 16 || d->C == 32 || d->C == 48 || (d->C % 64) == 0, "C=%d unsupported", d->C);
  MMH_CHECK(d->N >= 16 && (d->N % 16) == 0, "N=%d must be a multiple of 16", d->N);
  MMH_CHECK((d->a_ld % 8) == 0, "a_ld=%d invalid", d->a_ld);
  MMH_CHECK((d->out_ld % 8) == 0, "out_ld=%d must be a multiple of 8", d->out_ld);
  MMH_CHECK(d->M > 0 && d->M < (int64_t(1) << 31) - 256, "M out of range");
  MMH_CHECK((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
            "operand pointers must be 16-byte aligned");
  auto* plan = new MmhConvPlan();
  if (mmh_conv2_create(d, &plan->v2)) { delete plan; return 1; }
  *out_plan = plan;
  return 0;
}

extern "C" int mmh_conv_plan_destroy(MmhConvPlan* plan) {
  delete plan;
  return 0;
}

extern "C" int mmh_conv_run(const MmhConvPlan* plan, void* stream) {
  MMH_CHECK(plan && plan->v2, "null plan");
  return mmh_conv2_run(plan->v2, stream);
}

extern "C" int mmh_conv_run_key(const MmhConvPlan* plan, uint32_t drop_key, void* stream) {
  MMH_CHECK(plan && plan->v2, "null plan");
  return mmh_conv2_run_key(plan->v2, drop_key, stream);
}

extern "C" int mmh_wgrad_plan_create(const MmhWgradDesc* d, MmhWgradPlan** out_plan) {
  MMH_CHECK(d && out_plan, "null argument");
  MMH_CHECK(d->T >= 1 && d->T <= MMH_MAX_TAPS, "T=%d out of range", d->T);
  MMH_CHECK(d->C >= 16 && (d->C % 16) == 0 && d->N >= 16 && (d->N % 16) == 0, "C=%d / N=%d must be multiples of 16",
            d->C, d->N);
  MMH_CHECK(d->C <= 256 || (d->C % 256) == 0, "C=%d unsupported (must be <=256 or a multiple of 256)", d->C);
  MMH_CHECK((d->a_ld % 8) == 0 && (d->dy_ld % 8) == 0, "leading dimensions must be multiples of 8");
  MMH_CHECK(d->M > 0 && d->M < (int64_t(1) << 31) - 256, "M out of range");
  auto* plan = new MmhWgradPlan();
  if (mmh_wgrad2_create(d, &plan->v2)) { delete plan; return 1; }
  *out_plan = plan;
  return 0;
}

extern "C" int mmh_wgrad_plan_destroy(MmhWgradPlan* plan) {
  delete plan;
  return 0;
}

extern "C" int mmh_wgrad_run(const MmhWgradPlan* plan, void* stream) {
  MMH_CHECK(plan && plan->v2, "null plan");
  return mmh_wgrad2_run(plan->v2, stream);
}
