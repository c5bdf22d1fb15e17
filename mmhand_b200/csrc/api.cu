// Library-level entry points: version, error string, driver entry points, device attributes.
#ifndef MMH_HOST_EMU
#include <cuda.h>
#include <cuda_runtime.h>
#endif
#include <stdarg.h>
#include <stdio.h>

#include <mutex>

#include "../../include/mmhand_sm100.h"
#include "host_common.h"
#ifndef MMH_HOST_EMU
#include "tmap.h"
#endif

namespace mmh {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

#ifndef MMH_HOST_EMU
int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t ld, uint32_t box_cols,
                      uint32_t box_rows, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = get_encode();
  MMH_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): cols=%lld rows=%lld ld=%lld box=%ux%u", (int)r,
            (long long)cols, (long long)rows, (long long)ld, box_cols, box_rows);
  return 0;
}
#endif  // !MMH_HOST_EMU

}  // namespace mmh

#include <stdlib.h>
#include "ew_common.h"
namespace mmh {
static int g_pdl = -1;      // -1: not set through the ABI -> MMH_PDL, else on
bool pdl_enabled() {
  if (g_pdl >= 0) return g_pdl != 0;
  static const bool env_on = [] {
    const char* e = getenv("MMH_PDL");
    return e == nullptr || atoi(e) != 0;
  }();
  return env_on;
}
}  // namespace mmh
extern "C" int mmh_set_pdl(int32_t on) { mmh::g_pdl = on < 0 ? -1 : (on != 0 ? 1 : 0); return 0; }
extern "C" int mmh_get_pdl(void) { return mmh::pdl_enabled() ? 1 : 0; }
extern "C" int mmh_version(void) { return 100; }
extern "C" int mmh_act_bytes(void) { return static_cast<int>(sizeof(mmh::act_t)); }
extern "C" const char* mmh_last_error(void) { return mmh::get_error(); }
#ifdef MMH_HOST_EMU
extern "C" int mmh_is_device_build(void) { return 0; }
#else
extern "C" int mmh_is_device_build(void) { return 1; }
#endif

/* ---- stream-ordering primitives for the launch tapes (side-stream weight gradients) ---- */
#ifdef MMH_HOST_EMU
extern "C" int mmh_event_create(void** ev) { MMH_CHECK(ev != nullptr, "null argument"); *ev = reinterpret_cast<void*>(0x1); return 0; }
extern "C" int mmh_event_destroy(void*) { return 0; }
extern "C" int mmh_event_record(void*, void*) { return 0; }
extern "C" int mmh_stream_wait_event(void*, void*) { return 0; }
#else
extern "C" int mmh_event_create(void** ev) {
  MMH_CHECK(ev != nullptr, "null argument");
  cudaEvent_t e;
  MMH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *ev = e;
  return 0;
}
extern "C" int mmh_event_destroy(void* ev) {
  if (ev != nullptr) MMH_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(ev)));
  return 0;
}
extern "C" int mmh_event_record(void* ev, void* stream) {
  MMH_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev), static_cast<cudaStream_t>(stream)));
  return 0;
}
extern "C" int mmh_stream_wait_event(void* stream, void* ev) {
  MMH_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(ev), 0));
  return 0;
}
#endif
