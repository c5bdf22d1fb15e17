// Host-side helpers shared by the API translation units: error reporting, driver entry points.
#pragma once
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace mmh {

void set_error(const char* fmt, ...);
const char* get_error();

#define MMH_CHECK(cond, ...)            \
  do {                                  \
    if (!(cond)) {                      \
      ::mmh::set_error(__VA_ARGS__);    \
      return 1;                         \
    }                                   \
  } while (0)

// programmatic dependent launch on / off for every launch of the library (mmh_set_pdl; default: MMH_PDL, else on)
bool pdl_enabled();

#ifndef MMH_HOST_EMU
#define MMH_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::mmh::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)
int num_sms();
#endif

}  // namespace mmh
