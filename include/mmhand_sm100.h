/*
 * mmhand_sm100.h -- C ABI of libmmhand_sm100.so: the B200 (sm_100a) kernels behind the MM-HAND
 * generator / discriminator / loss hot path.
 *
 * Conventions
 *   - every entry point returns 0 on success, non-zero on error; mmh_last_error() returns the message
 *     (thread-local). Nothing throws across the ABI.
 *   - the library never allocates or frees device memory: every buffer (including the conv plans'
 *     operands) belongs to the caller (PyTorch's caching allocator on the Python side).
 *   - every compute call is asynchronous on the cudaStream_t passed as `stream` (a void* here so that
 *     the header needs no CUDA include); no hidden synchronisation.
 *   - activations are NHWC bf16 on "pixel grids": a 2-D array [rows][channels] whose row index is the
 *     flattened (image, grid row, grid column) position. See DESIGN.md section 3.
 *
 * Each block below names the reference call it replaces (paths relative to the reference tree).
 */
#ifndef MMHAND_SM100_H_
#define MMHAND_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMH_MAX_TAPS 64

/* ---- library ---------------------------------------------------------------------------------- */
int mmh_version(void);
const char* mmh_last_error(void);
/* 1 when the library was built as the CUDA product, 0 for the host emulation used by CPU tests. */
int mmh_is_device_build(void);

/* ---- convolution as a shifted-row GEMM on tcgen05 ---------------------------------------------- */
/*
 * Replaces every nn.Conv2d / nn.ConvTranspose2d forward and data-gradient call of
 *   models/Generator.py:62-111,158-259   models/Discriminator.py:28-49,79-99
 *   losses/L1_plus_perceptualLoss.py:22-27,54-61 (VGG19.features[0:4])
 * (cuDNN fprop / bwd-data in the reference).
 *
 *   out[map(q)][n] = act( bias[n] + sum_{t<T} sum_{c<C} a[q + shift[t]][c] * w[w_slot[t]][n][c] ),  q in [0,M)
 *
 * a : bf16 [a_rows][a_ld], only channels [0,C) of each row are read; rows outside [0,a_rows) read 0.
 * w : bf16 [w_taps][N][C] (K-major B operand); tap t reads slab w_slot[t].
 * q is decoded on the GEMM grid: img = q / (Hg*Wg), h = (q % (Hg*Wg)) / Wg, x = q % Wg; the row is
 * "valid" iff h < Hv && x < Wv. Valid rows are stored at
 *   out_row = img*out_img_rows + (h*out_sh + out_h0)*out_wg + (x*out_sw + out_w0)
 * Invalid rows are skipped, or (zero_invalid=1) stored as zeros.
 */
typedef struct MmhConvDesc {
  const void* a;
  int64_t a_rows;
  int32_t a_ld;
  int32_t C; /* multiple of 16; 16, 32, 48 or a multiple of 64 */
  const void* w;
  int32_t T;
  int32_t N; /* multiple of 16 */
  int32_t shift[MMH_MAX_TAPS];
  int32_t w_slot[MMH_MAX_TAPS]; /* tap t uses weight slab w[w_slot[t]] */
  int32_t w_taps;               /* number of slabs in w */
  int64_t M;
  int32_t Hg, Wg, Hv, Wv;
  void* out;
  int32_t out_f32; /* 0: bf16, 1: fp32 */
  int32_t out_ld;  /* elements per output row */
  int64_t out_img_rows;
  int32_t out_wg, out_sh, out_sw, out_h0, out_w0;
  int32_t zero_invalid;
  const float* bias; /* may be NULL */
  int32_t act;       /* 0 none, 1 relu, 2 tanh */
  int32_t n_store;   /* channels actually stored (<= N); 0 means N */
} MmhConvDesc;

typedef struct MmhConvPlan MmhConvPlan;
int mmh_conv_plan_create(const MmhConvDesc* desc, MmhConvPlan** plan);
int mmh_conv_plan_destroy(MmhConvPlan* plan);
int mmh_conv_run(const MmhConvPlan* plan, void* stream);

/*
 * Weight gradient (cuDNN bwd-filter in the reference; autograd of the same call sites):
 *   dw[t][n][c] += sum_{q<M} dy[q][n] * a[q + shift[t]][c]
 * dy : bf16 [M][dy_ld] (zeros at invalid grid positions), a as above.
 * dw : fp32 [T][N][C_store] accumulated with atomics (caller zeroes it); n < N_store, c < C_store.
 */
typedef struct MmhWgradDesc {
  const void* a;
  int64_t a_rows;
  int32_t a_ld;
  int32_t C; /* channels of a used (multiple of 16) */
  const void* dy;
  int64_t M;
  int32_t dy_ld;
  int32_t N; /* channels of dy used (multiple of 16) */
  int32_t T;
  int32_t shift[MMH_MAX_TAPS];
  float* dw;
  int32_t tap_index[MMH_MAX_TAPS]; /* slot of tap t inside dw's tap dimension */
  int32_t dw_taps;                 /* size of dw's tap dimension */
  int32_t N_store, C_store;
  int32_t split_k; /* 0: choose */
  int32_t dbg_lbo_sbo_swap; /* bring-up only: swap LBO/SBO roles of the MN-major descriptors */
} MmhWgradDesc;

typedef struct MmhWgradPlan MmhWgradPlan;
int mmh_wgrad_plan_create(const MmhWgradDesc* desc, MmhWgradPlan** plan);
int mmh_wgrad_plan_destroy(MmhWgradPlan* plan);
int mmh_wgrad_run(const MmhWgradPlan* plan, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMHAND_SM100_H_ */
