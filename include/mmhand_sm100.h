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
/* Programmatic dependent launch (every kernel starts with griddepcontrol.wait / launch_dependents) on (1) / off (0) for
 * all later launches; -1 returns to the default (environment MMH_PDL, else on). Data-parallel groups switch it off:
 * CTAs parked by a dependent launch next to BatchNorm kernels that spin on their peers' mailboxes starved NCCL's
 * kernels on 8 GPUs (mmhand_b200/runtime.py::World). No counterpart in the reference (launch plumbing). */
int mmh_set_pdl(int32_t on);
int mmh_get_pdl(void);

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
 *     C may exceed a_ld: a row then runs on into the following pixels (overlapping-row view), which folds the
 *     kw taps of a small-channel k x k convolution into the contraction dimension (K = kw * a_ld per kh tap).
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
  /* Fused BatchNorm statistics (replaces the mmh_bn_stats pass over `out`): bn_sums[0][c] += sum, bn_sums[1][c] +=
   * sum of squares of the stored (bf16-rounded) valid outputs of channel c < bn_C. NULL = off. Needs bias == NULL,
   * act == 0, bf16 output. Finish with mmh_bn_finalize_reset. */
  float* bn_sums;
  int32_t bn_C;
  int32_t reserved0;
  /* Fused BatchNorm-BACKWARD statistics and masks (replaces mmh_bn_bwd_reduce_finalize's pass over dz and x), for a
   * data-gradient launch whose output grid is the (padded) input of a convolution fed by
   *     conv_p -> BatchNorm -> [ReLU] -> [Dropout(0.5)] -> reflect padding        (models/Generator.py:62-77,
   *                                                                                 models/Discriminator.py:28-35).
   * Output row q = grid position (hp, wp) of image b is the gradient of logical pixel (hp - bs_pad, wp - bs_pad),
   * mirrored into [0, bs_H) x [0, bs_W) (torch ReflectionPad2d), whose raw conv_p output x is row
   * (b*bs_xHg + hs)*bs_xWg + ws of bs_x (bf16, bs_x_ld elements per row). With a = bs_coef[c], b = bs_coef[C + c],
   * mean = bs_save[c], rstd = bs_save[C + c] (C = bs_C) the launch STORES
   *     dze = out * [a*x + b > 0 (bs_relu)] * [2 if kept else 0 (bs_dropout; hash of mmh_norm_act with bs_drop_key)]
   * and accumulates bs_sums[0][c] += sum dze, bs_sums[1][c] += sum dze * (x - mean) * rstd over all rows -- summed
   * over the halo positions too, which is the reflect-fold of the gradient. Finish with mmh_bn_bwd_finalize_reset, then
   * mmh_bn_bwd_apply with relu = dropout = 0 on the stored dze. bs_x == NULL = off. Needs bias == NULL, act == 0,
   * bf16 output, Hv == Hg, Wv == Wg, bn_sums == NULL. */
  const void* bs_x;
  const float* bs_coef;
  const float* bs_save;
  float* bs_sums;
  int32_t bs_x_ld, bs_xHg, bs_xWg, bs_H, bs_W, bs_pad, bs_C, bs_relu, bs_dropout;
  uint32_t bs_drop_key;
} MmhConvDesc;

typedef struct MmhConvPlan MmhConvPlan;
int mmh_conv_plan_create(const MmhConvDesc* desc, MmhConvPlan** plan);
int mmh_conv_plan_destroy(MmhConvPlan* plan);
int mmh_conv_run(const MmhConvPlan* plan, void* stream);
/* Same launch with another dropout key (bs_drop_key changes every training step; everything else is fixed). */
int mmh_conv_run_key(const MmhConvPlan* plan, uint32_t drop_key, void* stream);

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
} MmhWgradDesc;

typedef struct MmhWgradPlan MmhWgradPlan;
int mmh_wgrad_plan_create(const MmhWgradDesc* desc, MmhWgradPlan** plan);
int mmh_wgrad_plan_destroy(MmhWgradPlan* plan);
int mmh_wgrad_run(const MmhWgradPlan* plan, void* stream);

/* ---- pixel-grid layout descriptor used by every bandwidth-bound kernel ------------------------- */
/*
 * Logical pixel (b, h, w) of the H x W content lives at grid position (h + h0, w + w0) of image b;
 * phase != 0 splits the grid into four parity planes of Hg x Wg each (plane = (row&1)*2 + (col&1)).
 * ld = elements per row, c0 = first channel of this view, C = channels of this view (multiple of 8).
 */
typedef struct MmhLay {
  int32_t B, H, W, Hg, Wg, h0, w0, phase, ld, c0, C, reserved;
} MmhLay;

/*
 * Input assembly: replaces torch.cat + ReflectionPad2d(3) + the implicit NCHW->NHWC/bf16 conversion
 * (models/MMHandModel.py:216-220,238-243,278-289; models/Generator.py:158-189) and the VGG re-normalisation
 * (losses/L1_plus_perceptualLoss.py:40-58, scale/shift per channel, zero halo).
 * dst[b,h,w,c] for h in [-pad_lo, H+pad_hi): channels [0,C0) from src0, [C0,C0+C1) from src1 (NCHW fp32),
 * the rest zero.
 */
int mmh_assemble_nchw(const float* src0, int32_t C0, const float* src1, int32_t C1, const float* scale,
                      const float* shift, void* dst, const MmhLay* dl, int32_t pad_lo, int32_t pad_hi,
                      int32_t reflect, void* stream);

/* BatchNorm2d statistics (cudnnBatchNorm in the reference; models/Generator.py:67,73,110 etc.).
 * x: bf16 [rows][ld] whose invalid grid positions are zero; sums[0][c] += sum x, sums[1][c] += sum x^2. */
int mmh_bn_stats(const void* x, int64_t rows, int32_t ld, int32_t C, float* sums, void* stream);
/* train: mean/var from sums and count, running stats updated (momentum, unbiased var); eval: running stats.
 * coef = (gamma*rstd, beta - mean*gamma*rstd), save = (mean, rstd). gamma/beta NULL -> 1/0. */
int mmh_bn_finalize(const float* sums, float count, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, int32_t train, int32_t C, float* coef,
                    float* save, void* stream);

/* Fused BN-apply + ReLU + Dropout(0.5) + residual + next layer's padding
 * (models/Generator.py:62-77 sequence; models/Discriminator.py:53-55 residual). */
typedef struct MmhNormAct {
  const void* src;
  MmhLay sl;
  const float* coef; /* [2][C] or NULL (identity) */
  int32_t relu, dropout;
  uint32_t drop_key;
  int32_t reserved;
  const float* resid; /* fp32 plain [B*H*W][C] or NULL */
  void* dst;          /* bf16, may be NULL */
  MmhLay dl;
  int32_t pad_lo, pad_hi, reflect, reserved2;
  float* dst_f32; /* fp32 plain [B*H*W][C] or NULL */
} MmhNormAct;
int mmh_norm_act(const MmhNormAct* p, void* stream);

/* PATBlock tail (models/Generator.py:120-130): out = x1 + BN(c1)*sigmoid(x2o)*sigmoid(x3o); writes the three
 * next-block inputs x1' = out, x2' = [x3o | out], x3' = [x2o | out] (the reference's swapped unpacking).
 * x3o == NULL: the two-stream block of the pose-transfer baseline
 * (baselines/quantitative_on_benchmarks/networks/model_variants.py:58-68): out = x1 + BN(c1)*sigmoid(x2o), the one
 * next-block stream input [x2o | out] goes to d3 (d2 = NULL). The backward entry points take x3o / dy3 == NULL alike. */
typedef struct MmhGateFwd {
  const void* c1;
  const void* x2o;
  const void* x3o;
  MmhLay sl;
  const float* coef;
  const float* trunk_in;
  float* trunk_out;
  void* d1;
  MmhLay d1l;
  void* d2; /* may be NULL */
  MmhLay d2l;
  void* d3; /* may be NULL */
  MmhLay d3l;
  int32_t pad_lo, pad_hi, reflect, reserved;
} MmhGateFwd;
int mmh_gate_fwd(const MmhGateFwd* p, void* stream);

/* Gradient sources: the data gradient of a consumer convolution, laid out like that convolution's input
 * (with halo); folding the halo back is the backward of ReflectionPad2d / zero padding. */
typedef struct MmhGradSrc {
  const void* p; /* bf16 */
  MmhLay l;
  int32_t pad_lo, pad_hi, reflect, reserved;
} MmhGradSrc;

typedef struct MmhGradGather {
  int32_t nsrc, dst_f32, B, H;
  int32_t W, C, reserved0, reserved1;
  MmhGradSrc src[4];
  const float* trunk; /* fp32 plain or NULL */
  const void* mask;   /* bf16, multiply by (mask > 0) or NULL */
  MmhLay ml;
  void* dst;
  MmhLay dl;
} MmhGradGather;
int mmh_grad_gather(const MmhGradGather* p, void* stream);

/* BatchNorm backward (+ ReLU / dropout masks recomputed from the saved raw output).
 * The upstream gradient is either a materialised plain buffer (dz != NULL, nsrc == 0) or gathered on the fly
 * (dz == NULL): dz = trunk (fp32 plain, optional) + sum over nsrc <= 2 consumer data gradients with their halos
 * folded back -- the gather of mmh_grad_gather without the round trip through memory. */
typedef struct MmhBnBwd {
  const void* dz; /* plain [B*H*W][C] or NULL */
  int32_t dz_f32, relu, dropout;
  uint32_t drop_key;
  const void* x;
  MmhLay xl;
  const float* coef;
  const float* save;
  float* sums;    /* reduce: [2][C] += (sum dz, sum dz*xhat) */
  const float* k; /* apply: [2][C] (mean dz, mean dz*xhat) */
  void* dy;       /* apply: bf16 on layout yl (interior only) */
  MmhLay yl;
  int32_t nsrc, reserved;
  MmhGradSrc src[2];
  const float* trunk;
} MmhBnBwd;
int mmh_bn_bwd_reduce(const MmhBnBwd* p, void* stream);
int mmh_bn_bwd_apply(const MmhBnBwd* p, void* stream);
/* k = sums_global / count; dgamma += sums_local[1], dbeta += sums_local[0] (NULL to skip). */
int mmh_bn_bwd_finalize(const float* sums_global, const float* sums_local, float count, float* k, float* dgamma,
                        float* dbeta, int32_t C, void* stream);

typedef struct MmhGateBwd {
  const float* dout; /* fp32 plain [B*H*W][C] */
  const void* c1;
  const void* x2o;
  const void* x3o;
  MmhLay sl;
  const float* coef;
  const float* save;
  float* sums;
  const float* k;
  MmhGradSrc ex2, ex3; /* p == NULL: none */
  void* dy1;
  void* dy2;
  void* dy3;
  MmhLay yl;
} MmhGateBwd;
int mmh_gate_bwd_reduce(const MmhGateBwd* p, void* stream);
int mmh_gate_bwd_apply(const MmhGateBwd* p, void* stream);

/* ---- losses ------------------------------------------------------------------------------------ */
/* BCEWithLogitsLoss against a constant label (models/network_utils.py:141,160-163):
 * *loss_acc += loss_scale * sum bce(x, target); grad[i] = grad_scale * (sigmoid(x) - target) (grad may be NULL) */
int mmh_bce_logits(const float* x, int64_t n, float target, float loss_scale, float grad_scale, float* loss_acc,
                   float* grad, void* stream);
/* F.l1_loss on images (losses/L1_plus_perceptualLoss.py:37): grad_acc[i] += grad_scale*sign(a-b) (may be NULL) */
int mmh_l1_f32(const float* a, const float* b, int64_t n, float loss_scale, float grad_scale, float* loss_acc,
               float* grad_acc, void* stream);
/* perceptual L1/MSE between two feature grids + gradient (:63-71). mse: bit 0 = squared error; bit 1 = the features end
 * on a convolution (perceptual_layers 0 / 2), else on a ReLU whose mask the gradient takes (perceptual_layers 1 / 3) */
int mmh_perc_loss(const void* ff, const void* ft, int64_t n, int32_t mse, float loss_scale, float grad_scale,
                  float* loss_acc, void* dy, void* stream);
/* tanh backward into the last conv's dY grid (models/Generator.py:259): dy = dfake * (1 - fake^2) */
int mmh_tanh_bwd(const float* dfake_nchw, const float* fake_nchw, void* dy, const MmhLay* yl, int32_t C, void* stream);
/* gradient w.r.t. an NCHW fp32 network input: fold the halo, take channels [0,C), optional per-channel scale */
int mmh_input_grad_nchw(const MmhGradSrc* src, const float* scale, float* dst_nchw, int32_t B, int32_t C, int32_t H,
                        int32_t W, int32_t accumulate, void* stream);
/* fp32 grid rows -> NCHW fp32 (the generator's output image) */
int mmh_grid_to_nchw(const float* src, const MmhLay* sl, float* dst_nchw, int32_t C, void* stream);

/* ---- parameters --------------------------------------------------------------------------------- */
/* fp32 master weight (any strides: n, c, tap) -> bf16 [T][Np][Cp] zero padded */
int mmh_pack_weight(const float* src, int64_t s_n, int64_t s_c, int64_t s_t, int32_t N, int32_t C, int32_t T,
                    void* dst, int32_t Np, int32_t Cp, void* stream);
/* same for the kw-folded operand: dst bf16 [kh][Np][Kw] with dst[t][n][j*Cin_p + c] = src[n][c][t*kw + (reverse ? kw-1-j : j)] */
int mmh_pack_weight_folded(const float* src, int64_t s_n, int64_t s_c, int64_t s_t, int32_t N, int32_t C, int32_t kh,
                           int32_t kw, void* dst, int32_t Np, int32_t Cin_p, int32_t Kw, int32_t reverse,
                           void* stream);
/* fp32 [T][N][C] packed gradient -> strided fp32 gradient (accumulate != 0: +=) */
int mmh_unpack_wgrad(const float* src, float* dst, int64_t s_n, int64_t s_c, int64_t s_t, int32_t N, int32_t C,
                     int32_t T, int32_t accumulate, void* stream);
/* torch.optim.Adam step (models/MMHandModel.py:90-98) on a flat fp32 buffer; g is multiplied by grad_scale */
/*
 * Batched parameter jobs: one launch packs every weight of a network into its tensor-core operands (kind 0:
 * fp32 master, element (n, c, t) at src[n*s_n + c*s_c + t] -> bf16 dst[t][Np][Cp], zero padded) or folds every
 * packed weight gradient back into the OIHW .grad tensors (kind 1: dst[n*s_n + c*s_c + t] += src[t][N][C]).
 * `jobs` lives in device memory (the caller uploads it once); a job owns the tiles [tile_begin, next tile_begin)
 * of 256 (n, c) pairs each. Replaces ~450 small launches per training step.
 */
typedef struct MmhParamJob {
  const void* src;
  void* dst;
  int64_t s_n, s_c;
  int32_t kind, N, C, T, Np, Cp, tile_begin, reserved;
} MmhParamJob;
int mmh_param_jobs(const MmhParamJob* jobs, int32_t n_jobs, int32_t total_tiles, void* stream);

int mmh_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
             int32_t step, float grad_scale, void* stream);
int mmh_memset(void* p, int32_t byte, int64_t bytes, void* stream);

/* ---- keypoints -> heatmaps (data/generic_dataset.py:191-217,238-242) ----------------------------- */
/* uv: float64 [n_maps][2] (x, y); out: fp32 [n_maps][H][W] = exp(-((gx-x)^2+(gy-y)^2)/2/sigma/sigma),
 * >1 -> 1, < thresh -> 0, all in fp64, cast last. */
int mmh_heatmap_rasterize(const double* uv, int64_t n_maps, int32_t H, int32_t W, double sigma, double thresh,
                          float* out, void* stream);
/* Offline pose maps of the dataset tool (tool/generate_pose_map_RHD.py:22-29, cords_to_map): yx [n_pose][J][2] = (y, x)
 * float64 -> out [n_pose][H][W][J] fp32 = exp(-((gy - y)^2 + (gx - x)^2) / (2 sigma^2)) evaluated in float64, no clamp,
 * no threshold; a joint with y == missing or x == missing (MISSING_VALUE = -1) keeps a zero plane. */
int mmh_pose_map_rasterize(const double* yx, int64_t n_pose, int32_t J, int32_t H, int32_t W, double sigma,
                           double missing, float* out_hwc, void* stream);

/* ---- joints -> depth-ordered part map (generate_jointsmap, data/generic_dataset.py:30-78) ------------
 * uv: float64 [n_pose][21][2] (x, y); depth: float64 [n_pose][21]. Per bone (:33-54): ellipse polygon
 * cv2.ellipse2Poly((int(mx), int(my)), (int(len/2), 5), int(angle), 0, 360, 1) filled with cv2.fillConvexPoly
 * (OpenCV's integer arithmetic restated: sine table, cvRound, LineIterator outline, XY_SHIFT=16 scan conversion);
 * a pixel gets the colour (10..200) of the last bone whose mean joint depth equals the running minimum there, 0 where
 * no bone covers it. out_f64: [n_pose][H][W][3] float64 (the reference's canvas) and / or out_u8: [n_pose][H][W];
 * either may be NULL. H <= 512. */
int mmh_jointsmap_rasterize(const double* uv, const double* depth, int64_t n_pose, int32_t H, int32_t W,
                            double* out_f64, uint8_t* out_u8, void* stream);

/* ---- aug.py write-out (aug.py:57-71) ---------------------------------------------------------------
 * src: fp32 [n_img][3][H][W] (RGB planes, values in (-1, 1)); dst: uint8 [n_img][H][W][3] BGR =
 * saturate_cast<uchar>(cvRound((x * 0.5f + 0.5f) * 255.f)) -- what cv2.imwrite stores for the reference's float image. */
int mmh_image_pack_bgr8(const float* src_nchw, int64_t n_img, int32_t H, int32_t W, uint8_t* dst_nhwc, void* stream);

/* ---- synchronised BatchNorm over NVLink peer memory ------------------------------------------------
 * Replaces apex.parallel.convert_syncbn_model + its per-layer NCCL all-reduces (models/MMHandModel.py:109-116):
 * the BN finalise kernels exchange their 2*C partial sums through mailboxes mapped into every peer GPU of the box
 * (one process per GPU; CUDA IPC) and reduce them in rank order, so all ranks hold bit-identical statistics.
 * The mailbox is the one device allocation the library owns (an IPC object must be the base of its allocation).
 * Protocol: create on every rank -> exchange the MMH_PEER_HANDLE_BYTES handles out of band (torch.distributed) ->
 * connect. Every rank must issue the same exchanges in the same order with the same `seq` (1, 2, 3, ...).
 * mmh_peer_status: 0 ok, 1 = a wait timed out (a peer died; results are NaN), -1 = no peer support. */
#define MMH_PEER_HANDLE_BYTES 64
typedef struct MmhPeer MmhPeer;
int mmh_peer_create(int32_t rank, int32_t world, MmhPeer** out);
int mmh_peer_handle(MmhPeer* g, void* handle64);
int mmh_peer_connect(MmhPeer* g, const void* handles /* world x MMH_PEER_HANDLE_BYTES, rank order */);
int mmh_peer_status(MmhPeer* g);
int mmh_peer_destroy(MmhPeer* g);
/* in-place sum over ranks of n <= 2048 floats */
int mmh_peer_sum(MmhPeer* g, uint32_t seq, float* data, int32_t n, void* stream);
/* mmh_bn_finalize (train mode) on the global statistics: sums (local partial sums) is overwritten by the global
 * sums; count_global = elements per channel over all ranks. */
int mmh_bn_finalize_sync(MmhPeer* g, uint32_t seq, float* sums, float count_global, const float* gamma,
                         const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                         int32_t C, float* coef, float* save, void* stream);
/* mmh_bn_bwd_finalize with the exchange: sums_global := sum over ranks of sums_local; k from the global sums,
 * dgamma / dbeta from the local ones (they join the gradient all-reduce). */
int mmh_bn_bwd_finalize_sync(MmhPeer* g, uint32_t seq, const float* sums_local, float* sums_global,
                             float count_global, float* k, float* dgamma, float* dbeta, int32_t C, void* stream);

/* ---- one-launch BatchNorm statistics: reduction + (exchange) + finalisation ------------------------
 * The block of the reduction that finishes last (ticket `counter`, one zero-initialised uint32 per stream) runs the
 * finalisation of mmh_bn_finalize / mmh_bn_bwd_finalize -- preceded, when `peer` is not NULL, by the peer-memory
 * exchange above -- and resets the accumulators: `sums` must be zero on entry and is zero again on exit, so a BN
 * layer costs one launch per direction instead of memset + statistics + finalise. peer == NULL: single GPU
 * (seq ignored). count_global = elements per channel over all ranks. */
int mmh_bn_stats_finalize(MmhPeer* peer, uint32_t seq, const void* x, int64_t rows, int32_t ld, int32_t C, float* sums,
                          uint32_t* counter, float count_global, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float momentum, float eps, float* coef, float* save,
                          void* stream);
/* p->sums: zeroed scratch (local sums), p->k: out; dgamma / dbeta accumulate the local sums (NULL to skip) */
/* Finalisation alone, for statistics accumulated by a convolution epilogue (MmhConvDesc.bn_sums): (exchange,)
 * mmh_bn_finalize in train mode, sums reset to zero. */
int mmh_bn_finalize_reset(MmhPeer* peer, uint32_t seq, float* sums, float count_global, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                          int32_t C, float* coef, float* save, void* stream);
int mmh_bn_bwd_reduce_finalize(MmhPeer* peer, uint32_t seq, const MmhBnBwd* p, uint32_t* counter, float count_global,
                               float* dgamma, float* dbeta, void* stream);
/* Finalisation alone, for backward statistics accumulated by a data-gradient epilogue (MmhConvDesc.bs_sums):
 * (exchange,) k[0][c] = sum dze / count, k[1][c] = sum dze*xhat / count over all ranks, dgamma / dbeta += the local
 * sums (NULL to skip), sums reset to zero. */
int mmh_bn_bwd_finalize_reset(MmhPeer* peer, uint32_t seq, float* sums, float count_global, float* k, float* dgamma,
                              float* dbeta, int32_t C, void* stream);
int mmh_gate_bwd_reduce_finalize(MmhPeer* peer, uint32_t seq, const MmhGateBwd* p, uint32_t* counter,
                                 float count_global, float* dgamma, float* dbeta, void* stream);

/* ---- device side of the input pipeline ------------------------------------------------
 * What the reference's dataset workers compute per sample on the CPU (data/generic_dataset.py:133-159), applied on the
 * GPU to the uint8 frames cv2.imread returns, in the same float64 arithmetic with the fp32 cast last (bit-identical):
 *   mmh_image_unpack_u8: dst[b][c][h][w] = float(((double)src[b][h][w][c'] / 255.0 - 0.5) / 0.5), c' = 2 - c when
 *                        swap_rb (cv2.cvtColor(BGR2RGB), :140-143) -- NHWC uint8 -> NCHW fp32 in [-1, 1]
 *   mmh_depth_unpack_u8: d = 256.0 * src[..][hi_ch] + src[..][lo_ch]; v = float((d / div - 0.5) / 0.5), written to the
 *                        three channels of dst (:148-159: hi = 1 (G), lo = 2 (R) of the BGR frame, div = 700) */
int mmh_image_unpack_u8(const uint8_t* src_nhwc, int64_t n_img, int32_t H, int32_t W, int32_t swap_rb, float* dst_nchw,
                        void* stream);
int mmh_depth_unpack_u8(const uint8_t* src_nhwc, int64_t n_img, int32_t H, int32_t W, int32_t hi_ch, int32_t lo_ch,
                        double div, float* dst_nchw3, void* stream);

/* ---- evaluator hook of the benchmark harness ----------------------------------------------
 * pytorch_ssim.ssim (baselines/quantitative_on_benchmarks/pytorch_ssim/__init__.py:17-39,65-73) in one kernel:
 * img1, img2 fp32 NCHW [B][C][H][W]; zero-padded depthwise Gaussian window (window x window, sigma), C1 = 0.01^2,
 * C2 = 0.03^2. *mean_acc += sum of the SSIM map (divide by B*C*H*W for ssim(size_average=True)); per_image[b] += mean
 * of image b's map (size_average=False). Either output may be NULL. */
int mmh_ssim(const float* img1, const float* img2, int64_t B, int32_t C, int32_t H, int32_t W, int32_t window,
             float sigma, float* mean_acc, float* per_image, void* stream);

/* ---- stream ordering (cudaEvent wrappers, so that recorded launch sequences can fork / join streams) ----
 * The weight-gradient kernels of a layer run on a side stream next to the bandwidth-bound BatchNorm-backward
 * kernels of the following layers (apex's delay_allreduce backward has no such overlap, MMHandModel.py:110-116). */
int mmh_event_create(void** ev);
int mmh_event_destroy(void* ev);
int mmh_event_record(void* ev, void* stream);
int mmh_stream_wait_event(void* stream, void* ev);

#ifdef __cplusplus
}
#endif
#endif /* MMHAND_SM100_H_ */
