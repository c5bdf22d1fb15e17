// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace mmh {
// 2-D bf16 tensor [rows][ld] of which columns [0, cols) are addressable; box = box_cols x box_rows.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t ld,
                      uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swizzle);
}  // namespace mmh
