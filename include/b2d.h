/*
 * b2d.h -- C ABI of libb2d.so, the sm_100a kernels behind the latent-diffusion sampling path.
 *
 * The reference (Ruby-004/Diffusion_model_project) is pure Python/PyTorch and has no FFI: its
 * hot path calls torch.nn modules, i.e. ATen -> cuDNN/cuBLAS/oneDNN.  Each entry point below
 * replaces the ATen call sites listed next to it (file:line relative to the reference root).
 * The Python host mirror (diffusion_model_project_b200/*.py) binds these with ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (B2D_E_*); never throws, never exits;
 *     b2d_last_error() returns a thread-local message for the last failure on this thread;
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch's allocator); the library
 *     allocates and frees nothing on the device and retains no pointer past the call, except
 *     conv plans, which record the pointers given at creation (TMA descriptors embed addresses);
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises, so every call can be
 *     captured into a CUDA graph;
 *   - activations are channels-last: [N][D][H][W][C] (D = 1 for the UNet's 2-D maps), 16-bit
 *     (bf16, or IEEE fp16 where a call's f16 flag says so), with the channel count padded to a
 *     multiple of 64 (pad channels hold zeros).
 */
#ifndef B2D_H_
#define B2D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B2D_API __attribute__((visibility("default")))
#else
#define B2D_API
#endif

#define B2D_VERSION 8
#define B2D_MAX_SEG 6
#define B2D_MAX_TAPS 27

#define B2D_OK 0
#define B2D_E_INVALID (-1)     /* bad argument / unsupported shape */
#define B2D_E_CUDA (-2)        /* CUDA runtime or driver error     */
#define B2D_E_UNSUPPORTED (-3) /* not an sm_100 device, driver too old */

B2D_API int b2d_version(void);
B2D_API const char* b2d_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Scheduler step (DDPM p_sample / DDIM ddim_sample), one fused elementwise kernel.
 * Replaces Diffusion_model/src/diffusion.py:152-188 (p_sample), :195-234 (ddim_sample),
 * :103-125 (predict_x0_from_noise) -- ~10-12 ATen elementwise kernels + randn_like each.
 *
 *   x0  = clamp((x_t - b*eps) / a, clip_lo, clip_hi)            (clamp only if clip != 0)
 *   out = c1*x0 + c2*(kind==0 ? x_t : eps)  [+ s*z if s != 0]
 *
 * coef: device array of rows {a, b, c1, c2, s, 0, 0, 0}; the row used is
 * (step_idx ? *step_idx : 0) + step_off, so a captured graph can advance a device counter.
 * z = noise[i] if noise != NULL, else (if s != 0) N(0,1) from Philox4x32-10(seed, row, i); seed == 0 with
 * noise == NULL asserts s == 0 for the rows used (deterministic DDIM) and selects the kernel without the generator.
 * seed_dev != NULL: the Philox key is read from that device word at run time (and the generator kernel is selected), so a
 * captured graph draws fresh noise on every call without being re-captured; `seed` is then ignored.
 * Optional second output for the fused loop: x_bf16[(i / group) * group_stride + i % group].
 * If step_inc != 0 the kernel's last block adds step_inc to *step_idx after all reads; `ticket` is the caller-owned,
 * zero-initialised arrival counter of that protocol (one per sampling loop: loops on different streams never share
 * state; the kernel leaves it zero).  Required when step_inc != 0.
 * ---------------------------------------------------------------------------------------- */
B2D_API int b2d_scheduler_step(int kind, const float* x_t, const float* eps, const float* noise, float* x_out,
                       int64_t n_elem, const float* coef, int* step_idx, int step_off, int step_inc,
                       int clip, float clip_lo, float clip_hi, void* x_bf16, int group, int group_stride,
                       uint64_t seed, const uint64_t* seed_dev, unsigned int* ticket, int x16_f16 /* the 16-bit copy is fp16 */,
                       void* stream);

/* q_sample (diffusion.py:78-101): out = a*x0 + b*noise with per-image a,b (device [n_img]). */
B2D_API int b2d_q_sample(const float* x0, const float* noise, float* out, const float* a, const float* b,
                 int64_t n_img, int64_t per_img, void* stream);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05/TMEM with TMA-staged (zero-filled halo) tiles.
 * Replaces nn.Conv2d (unet/blocks.py:29-36, unet/models.py:120-128), nn.ConvTranspose2d k2 s2
 * (unet/blocks.py:128-133), nn.Conv3d k3/k1, stride (1,1,1)/(1,2,2) (vae/blocks.py:155-169,
 * vae/encoder.py:30,45,56,68, vae/decoder.py:31,47,59,71), nn.Linear / nn.Conv1d k1
 * (unet/blocks.py:196-207: in_proj, out_proj, proj_out), torch.cat along channels
 * (unet/models.py:177) and the residual add (vae/blocks.py:185, unet/blocks.py:234).
 *
 * GEMM view: M = output positions (tiles of 128 forming a box in (n,d,y,x)), N = cout,
 * K = sum over segments s, taps t of cin[s]; weight row-major [rows][ktot] bf16 with
 * k = kbase[s] + t*cin[s] + c; normally kbase[s] = ntaps * sum_{s'<s} cin[s'] (the hi/lo
 * split "fp32x" mode points two activation segments at the same weight block).
 * ---------------------------------------------------------------------------------------- */
typedef struct b2d_conv_desc {
  int32_t nseg;                    /* channel segments (torch.cat without the copy)            */
  const void* in[B2D_MAX_SEG];     /* bf16 [N][D][H][W][cin[s]]                                  */
  int32_t cin[B2D_MAX_SEG];        /* multiple of 64                                            */
  int32_t kbase[B2D_MAX_SEG];      /* K offset of the weight block multiplied with in[s]        */
  int32_t N, D, H, W;              /* input extent                                              */
  int32_t ntaps;
  int8_t tap_dz[B2D_MAX_TAPS], tap_dy[B2D_MAX_TAPS], tap_dx[B2D_MAX_TAPS]; /* input offsets     */
  int32_t stride_h, stride_w;      /* input coord = out*stride + tap offset                     */
  int32_t OH, OW;                  /* conv output extent per (n,d)                              */
  const void* weight;              /* bf16 [wrows][ktot]; wrows >= nphase*cout rounded to block_n */
  int32_t wrows, ktot;
  int32_t cout;                    /* valid output channels (per phase)                         */
  int32_t nphase;                  /* 1, or 4: ConvTranspose2d k2 s2, weight rows phase-major   */
  const float* bias;               /* [cout] or NULL                                            */
  void* out;                       /* see out_mode                                              */
  void* out_lo;                    /* optional bf16 residual part (x - bf16(x)), mode 0, or NULL */
  int32_t out_mode;                /* 0 bf16 NDHWC, 1 fp32 planar [N][D][C][H][W], 2 fp32 NDHWC,
                                      3 fused scheduler update (see sched_* below; `out` optionally receives eps as mode 2) */
  int32_t out_H, out_W;            /* extent of the output tensor                               */
  int32_t out_sy, out_sx, out_oy, out_ox; /* out pixel = oy*out_sy + out_oy (+ phase)           */
  int32_t out_cstride, out_coff;   /* channel stride / offset of the output tensor              */
  const void* residual;            /* bf16, same geometry as out (cstride res_cstride) or NULL  */
  const void* residual_lo;
  int32_t res_cstride;
  double* stats;                   /* [N][cout/stats_cpg][2] (sum, sumsq) atomically added, or NULL */
  int32_t stats_cpg;               /* channels per GroupNorm group                              */
  const float* out_scale;          /* per-channel multiplier (mode 1) or NULL                   */
  const float* out_mask;           /* fp32 [N][D][out_H][out_W] multiplier (mode 1) or NULL     */
  int32_t block_n;                 /* 0 = auto, else 16/64/128/256                              */
  int32_t out_f16;                 /* mode 0: store IEEE fp16 (saturating) instead of bf16 -- for raw pre-GroupNorm
                                      outputs / residual streams that are never an MMA operand            */
  int32_t res_f16;                 /* residual tensor holds fp16 instead of bf16                         */
  int32_t tune_ksplit;             /* 0 = cost model, 1..16 = force this K split (tools/tune_conv.py)                  */
  void* workspace;                 /* optional caller-owned scratch for split-K (first 16 KB: zero-initialised arrival
                                      counters, then fp32 partial tiles).  Plans that share one must run back to back on
                                      ONE stream; give concurrent sampling loops their own                            */
  int64_t workspace_bytes;
  /* Fused input normalisation (persistent engine, halo staging, one segment, cin <= 512): in[0] holds the RAW
   * pre-GroupNorm output of the producer (fp16 if in_f16, else bf16) and the kernel applies
   * silu?(gamma * (x - mean) * rstd + beta) to each staged tile in shared memory -- GroupNorm + SiLU of
   * vae/blocks.py:173-183 without a separate pass over HBM.  in_stats = the producer's `stats` buffer. */
  const double* in_stats;          /* [N][in_creal / in_cpg][2] or NULL (no fusion)                      */
  const float* in_gamma;           /* [in_creal] or NULL                                                 */
  const float* in_beta;
  int32_t in_cpg, in_creal;        /* channels per group; real (unpadded) input channels                 */
  int32_t in_f16, in_act;
  float in_eps;
  /* ... + the DoubleBlock's time embedding (unet/blocks.py:100-103), added after the activation:
   * + in_temb[in_temb_row[n * in_temb_row_stride] * in_temb_ncols + in_temb_col + c]; NULL = none */
  const float* in_temb;
  const int32_t* in_temb_row;
  int32_t in_temb_row_stride, in_temb_ncols, in_temb_col;
  int32_t tune_flags;              /* B2D_TUNE_* bits (comparison arms of the tests / tools; 0 in production)      */
  int32_t op_f16;                  /* the MMA operands -- in[] (after the fused input normalisation, if any) and weight --
                                      hold IEEE fp16 instead of bf16 (tcgen05 kind::f16 takes either at the same rate);
                                      the fused sampler update's 16-bit copy follows.  fp16's 10-bit mantissa is what
                                      holds the 1e-2 end-to-end bound over 50 sampling steps; range is not a
                                      concern because every operand is a normalised activation or a weight            */
  /* out_mode 3 -- the UNet's final_conv (unet/models.py:185) fused with the sampler update that consumes it
   * (diffusion.py:152-188 / :195-234): the epilogue computes eps = conv + bias for the cout (4, 8, 12 or 16) channels of
   * a pixel and applies b2d_scheduler_step's arithmetic (same operation order, same Philox counters: bit-identical to
   * the two-launch form) to the fp32 master latent sched_x [pixels][cout] in place, so eps never reaches HBM (unless
   * `out` != NULL).  The bf16 copy goes to sched_x_bf16[pixel * sched_bf16_stride + c] (+ the hi/lo remainder to
   * sched_x_bf16_lo).  The last CTA to finish adds sched_step_inc to *sched_step_idx (ticket protocol as above). */
  int32_t sched_kind;              /* 0 DDPM, 1 DDIM                                                               */
  float* sched_x;                  /* fp32 [N*D*OH*OW][cout], read and rewritten                                    */
  const float* sched_noise;        /* same shape, or NULL (in-kernel Philox when the row's s != 0)                   */
  const float* sched_coef;         /* rows {a, b, c1, c2, s, 0, 0, 0}                                               */
  int32_t* sched_step_idx;         /* device step counter (row = *idx + sched_step_off), or NULL                    */
  uint32_t* sched_ticket;          /* zero-initialised arrival counter, required when sched_step_inc != 0           */
  const uint64_t* sched_seed_dev;  /* Philox key read at run time, or NULL: sched_seed (0 = rows never draw noise)  */
  uint64_t sched_seed;
  int32_t sched_step_off, sched_step_inc;
  int32_t sched_clip;
  float sched_clip_lo, sched_clip_hi;
  void* sched_x_bf16;              /* optional bf16 copy of the new latent (the UNet's input buffer)                */
  void* sched_x_bf16_lo;           /* optional x - bf16(x) (fp32x mode)                                             */
  int32_t sched_bf16_stride;       /* channel stride of that buffer                                                 */
  int32_t reserved[2];
} b2d_conv_desc;

#define B2D_TUNE_NO_HALO 1    /* generic tiles even where halo staging applies   */
#define B2D_TUNE_NO_SPLITK 2  /* never split the K loop                          */
#define B2D_TUNE_CONTIG 4     /* CTAs walk contiguous unit ranges                */
#define B2D_TUNE_STRIDED 8    /* CTAs walk the unit list strided over the grid   */
#define B2D_TUNE_STREAMK 16   /* stream-K wherever it applies                    */
#define B2D_TUNE_NO_STREAMK 32 /* never stream-K                                  */
#define B2D_TUNE_PAIR 64      /* CTA pairs (cta_group::2) wherever they apply    */
#define B2D_TUNE_NO_PAIR 128  /* never CTA pairs                                 */

typedef struct b2d_conv_plan b2d_conv_plan;
B2D_API int b2d_conv_plan_create(const b2d_conv_desc* desc, b2d_conv_plan** plan);
B2D_API int b2d_conv_plan_destroy(b2d_conv_plan* plan);
B2D_API int b2d_conv_run(const b2d_conv_plan* plan, void* stream);
/* number of CTAs / block_n the plan launches with (introspection for tests and the bench) */
B2D_API int b2d_conv_plan_info(const b2d_conv_plan* plan, int32_t* grid_m, int32_t* grid_n, int32_t* block_n, int32_t* kblocks);
/* out[8] = {2 (one CTA per tile) or 3 (tcgen05 cta_group::2 CTA pairs), halo, ksplit (-1: stream-K), work units, CTAs launched, block_n, K-loop groups, workspace bytes used (KiB)} */
B2D_API int b2d_conv_plan_info2(const b2d_conv_plan* plan, int32_t* out8);

/* ------------------------------------------------------------------------------------------
 * GroupNorm apply (+SiLU, + time-embedding add), stats come from the producer's epilogue.
 * Replaces nn.GroupNorm + nn.SiLU + the broadcast add (unet/blocks.py:37-47,98-105,134-143,
 * 165-174,192-221; vae/blocks.py:152-161,177-183; vae/decoder.py:70,137-138; encoder.py:67,133-134).
 *   y = act(gamma[c]*(x-mean_g)*rstd_g + beta[c]) + temb[row[n]*temb_ld + temb_col + c]
 * x, y: bf16 [N][P][C] (C multiple of 8; cstride = C).  stats: [N][C/cpg][2] sums over
 * count = cpg*P elements.  stats_out (optional): GN(1,C) sums of y, atomically added.
 * ---------------------------------------------------------------------------------------- */
B2D_API int b2d_gn_apply(const void* x, const void* x_lo, void* y, void* y_lo, int32_t N, int64_t P, int32_t C,
                 const double* stats, int32_t cpg, const float* gamma, const float* beta, float eps, int32_t act,
                 const float* temb_table, const int32_t* temb_row, int32_t temb_row_stride, int32_t temb_ld,
                 int32_t temb_col, double* stats_out, int32_t in_f16 /* x holds fp16 (see b2d_conv_desc.out_f16) */,
                 int32_t out_f16 /* y is stored as fp16 (and SiLU is evaluated to ~1e-6 instead of tanh.approx's 2^-11) */, void* stream);

/* MaxPool2d(2,2) (unet/blocks.py:161-164,170) + GN(1,C) sums of the pooled map. bf16 NHWC. */
B2D_API int b2d_maxpool2x2_stats(const void* x, const void* x_lo, void* y, void* y_lo, int32_t N, int32_t H, int32_t W,
                         int32_t C, double* stats, int32_t f16 /* x and y hold fp16 instead of bf16 */, void* stream);

/* Per-sample fused forms for small maps (one CTA per sample; a sample of at most 65536 elements, one GroupNorm group,
 * C/8 dividing 1024; bf16 channels-last, no hi/lo split).  B2D_E_UNSUPPORTED otherwise -- the caller uses the two-launch form.
 *   b2d_gn_gn_apply:   y1 = act1(GN(x; stats1, gamma1, beta1)),  y2 = act2(GN(y1; gamma2, beta2))  -- the last norm of a
 *                      DoubleBlock followed by the attention pre-norm (unet/blocks.py:37-47, 192, 214); x may alias y1.
 *   b2d_maxpool2x2_gn: y = act(GN(maxpool2x2(x); gamma, beta))  -- Down (unet/blocks.py:161-174). */
B2D_API int b2d_gn_gn_apply(const void* x, int32_t in_f16, void* y1, void* y2, int32_t N, int64_t P, int32_t C, const double* stats1,
                    const float* gamma1, const float* beta1, float eps1, int32_t act1, const float* gamma2, const float* beta2,
                    float eps2, int32_t act2, int32_t out_f16 /* y1, y2 hold fp16 */, void* stream);
B2D_API int b2d_maxpool2x2_gn(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, const float* gamma, const float* beta,
                      float eps, int32_t act, int32_t f16 /* x and y hold fp16 */, void* stream);

/* nearest-neighbour (1,2,2) upsample, nn.Upsample (vae/decoder.py:46,58). bf16 NDHWC, ND = N*D. */
B2D_API int b2d_upsample2x_nearest(const void* x, void* y, int32_t ND, int32_t H, int32_t W, int32_t C, void* stream);

/* Second half of the z-folded conv_out (vae/decoder.py:71, Conv3d C -> co <= 3): the conv engine runs the 3x3x3 conv as a
 * 3x3 conv per z slice with output rows (kz, co) -- P: fp32 [ND][H][W][12], entry kz*4 + co -- and this pass sums the three
 * z contributions (zero padding in z), adds the bias and writes out[(img*out_cstride + out_coff + c)][H][W] =
 * (sum + bias[c]) * scale[c] * mask[img][y][x] (scale, mask, bias optional).  ND = N*D images, D slices per sample. */
B2D_API int b2d_zfold_combine(const float* P, int32_t ND, int32_t D, int32_t H, int32_t W, int32_t co, const float* bias,
                      const float* scale, const float* mask, float* out, int32_t out_cstride, int32_t out_coff, void* stream);

/* z-stacked input of the VAE conv_in layers (vae/encoder.py:30, decoder.py:31; C = 3 or 8 input channels): y[img][p][kz*C + c] =
 * x[img + kz - 1][p][c] for kz = 0..2, zero where the slice z + kz - 1 falls outside its sample (D slices per sample).
 * bf16 channels-last [ND][P][cpad], 3*C <= cpad; the conv then runs 9 in-plane taps instead of 27. */
B2D_API int b2d_zstack_cl(const void* x, void* y, int32_t ND, int32_t D, int64_t P, int32_t C, int32_t cpad, void* stream);

/* layout/precision plumbing at the module boundary:
 * planar fp32 [N][C][P] (optionally divided by scale[c], MaxNormalizer normalizer.py:46-51)
 * -> channels-last bf16 [N][P][cpad] written at channel offset coff (pad channels untouched). */
B2D_API int b2d_planar_to_cl(const float* x, void* y, void* y_lo, int32_t N, int32_t C, int64_t P, int32_t cpad, int32_t coff,
                     const float* div_scale, int32_t f16 /* y holds fp16 (no lo part) */, void* stream);
/* channels-last bf16 [N][P][cstride] (channels coff..coff+C) -> planar fp32 [N][C][P] */
B2D_API int b2d_cl_to_planar(const void* x, const void* x_lo, float* y, int32_t N, int32_t C, int64_t P, int32_t cstride,
                     int32_t coff, int32_t f16, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused small-sequence attention core: softmax(q k^T / sqrt(d)) v per (image, head).
 * Replaces the scaled-dot-product part of nn.MultiheadAttention (unet/blocks.py:196-227).
 * qkv: bf16 [N][T][3C] (q | k | v, heads contiguous inside each), out: bf16 [N][T][C].
 * ---------------------------------------------------------------------------------------- */
B2D_API int b2d_attention(const void* qkv, const void* qkv_lo, void* out, void* out_lo, int32_t N, int32_t T, int32_t C,
                  int32_t heads, int32_t f16 /* qkv and out hold fp16 (tensor-core path, no lo parts) */, void* stream);

/* ------------------------------------------------------------------------------------------
 * Exact Euclidean distance transform of binary images (distance of every non-zero pixel to the
 * nearest zero pixel), replacing the SciPy host round trip (predictor.py:1096-1116), and the
 * bilinear (align_corners=False) resize that follows it (predictor.py:951).
 * ---------------------------------------------------------------------------------------- */
/* out: 2*n_img*H*W floats -- the result in the first half, the second half is scratch. */
B2D_API int b2d_edt2d(const float* img, float* out, int32_t n_img, int32_t H, int32_t W, void* stream);
B2D_API int b2d_bilinear_resize(const float* x, float* y, int32_t n_img, int32_t H, int32_t W, int32_t OH, int32_t OW,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * Training step, first slice (csrc/train.cu).  Reference: predictor.py:722-748 (q_sample, concat, eps-prediction),
 * unet/metrics.py:337-402 (criterion), helper.py:428-430 (backward, optimizer.step), train.py:144-148 (Adam).
 * The data gradient of a 3x3 conv is b2d_conv_run on dY with mirrored taps and a transposed weight pack.
 * ---------------------------------------------------------------------------------------- */
/* torch.optim.Adam, single-tensor form, over a flat fp32 buffer (16-byte aligned): grad * grad_scale (+ weight_decay * p),
 * lerp / addcmul / addcdiv in torch's order, bias corrections from `step` (>= 1).  28 B of HBM traffic per parameter. */
B2D_API int b2d_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int32_t step, double grad_scale, void* stream);
/* normalized_mse_loss_per_component (metrics.py:337-402) on planar fp32 [N][C][P]: err[n*C+c] = mean((o-t)^2) / (mean(t^2) + eps)
 * [* w_c / sum w]; loss[0] = batch mean, loss[1 + n] = per-sample channel mean; grad (optional, same shape as pred) =
 * d loss[0] / d pred.  weight: [C] or NULL. */
B2D_API int b2d_nmse_loss(const float* pred, const float* target, int32_t N, int32_t C, int64_t P, const float* weight, float eps,
                  float* err, float* loss, float* grad, void* stream);
/* Backward of y = silu?(GroupNorm(1, C)(x)) [+ temb[n][c]] (unet/blocks.py:37-47, 98-105): x raw pre-norm values and dy the
 * upstream gradient, channels-last [N][P][C] 16-bit (hi + optional bf16 lo; *_f16: the hi part holds fp16); stats = the
 * forward (sum, sumsq) per sample.  Writes dx (same layout), accumulates dgamma[C], dbeta[C] and (optional) dtemb[N][C]
 * atomically (zero them first); sums: [N][2] fp64 scratch. */
B2D_API int b2d_gn_silu_bwd(const void* x_hi, const void* x_lo, int32_t x_f16, const void* dy_hi, const void* dy_lo, int32_t dy_f16,
                    void* dx_hi, void* dx_lo, int32_t dx_f16, int32_t N, int64_t P, int32_t C, const double* stats,
                    const float* gamma, const float* beta, float eps, int32_t act, double* sums, float* dgamma, float* dbeta,
                    float* dtemb, void* stream);
/* Weight gradients on tcgen05 (pixels are the GEMM K): dW[m][n][ty][tx] += sum over pixels of A[p][m] * B[s p + tap][n].
 *   kind 0  Conv2d 3x3 pad 1 (unet/blocks.py:29-36):      dw[co][cin_off + ci][ky][kx] += dY[p][co] * X[p + (ky-1, kx-1)][ci]
 *   kind 1  Linear / Conv1d k1 / 1x1 conv (blocks.py:247-258, models.py:122): dw[co][cin_off + ci] += dY[p][co] * X[p][ci]
 *   kind 2  ConvTranspose2d k2 s2 (unet/blocks.py:205):   dw[ci][co][ky][kx] += X[p][ci] * dY[2p + (ky, kx)][co]
 * dY [N][H'][W'][cout_pad], X [N][H][W][cin_pad] channels-last 16-bit (hi + optional bf16 lo: three products hi*hi + hi*lo
 * + lo*hi); H, W = extent of X, powers of two (kind 2: dY is 2H x 2W).  dw: fp32 in the reference's parameter layout,
 * atomically accumulated (zero it first); cin_off selects the channel block of a concatenated input (torch.cat skip | up,
 * unet/models.py:177).  dw_channels_last: dw is [rows][KH][KW][cols] (the last index of the reference layout moved to the
 * end: [Cout][3][3][Cin], [Cin][2][2][Cout]) -- a thread's 32 columns are then contiguous and go out as 16-byte vector
 * reductions, a quarter of the atomic traffic that bounds this kernel (train.UNetTrainer keeps its parameters that way). */
B2D_API int b2d_conv_wgrad(int32_t kind, const void* dy_hi, const void* dy_lo, int32_t cout_pad, const void* x_hi, const void* x_lo,
                   int32_t cin_pad, int32_t N, int32_t H, int32_t W, int32_t cout, int32_t cin, int32_t cin_off, int32_t cin_total,
                   float* dw, int32_t dw_channels_last, int32_t op_f16, void* stream);
/* out[c] += sum over rows of x[row][c], c < cvalid (bias gradients = sum of dY over pixels); x [rows][C] 16-bit hi (+ lo),
 * atomically accumulated (zero it first). */
B2D_API int b2d_channel_sum(const void* x_hi, const void* x_lo, int32_t f16, int64_t rows, int32_t C, int32_t cvalid, float* out, void* stream);
/* o = a + b on 16-bit hi (+ lo) tensors of n elements: the two gradient paths of a skip connection (unet/models.py:150-177). */
B2D_API int b2d_add16(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, void* o_hi, void* o_lo, int32_t f16, int64_t n,
              void* stream);
/* MaxPool2d(2, 2) backward (unet/blocks.py:161-164): x [N][H][W][C] = the pooled layer's input, dy [N][H/2][W/2][C]; the
 * gradient goes to the first maximum of each window in scan order (torch's saved index), zeros elsewhere. */
B2D_API int b2d_maxpool2x2_bwd(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, void* dx_hi, void* dx_lo, int32_t f16,
                       int32_t N, int32_t H, int32_t W, int32_t C, void* stream);

/* Backward of b2d_attention's core: given qkv, the forward output out and d loss / d out (all channels-last 16-bit hi + lo
 * as in b2d_attention), writes dqkv [N][T][3C] (dq | dk | dv).  fp32 CUDA-core arithmetic; stats: [N][heads][T][2] fp32
 * scratch (log-sum-exp and dO.O per query row).  Any T; head dim a multiple of 16 and <= 512. */
B2D_API int b2d_attention_bwd(const void* qkv, const void* qkv_lo, const void* out, const void* out_lo, const void* dout, const void* dout_lo,
                      void* dqkv, void* dqkv_lo, float* stats, int32_t N, int32_t T, int32_t C, int32_t heads, int32_t f16, void* stream);

/* Rewrite a packed conv operand in place from fp32 parameters (after an optimizer step; plans keep pointing at it):
 * dst[(r1 * R2 + r2) * ktot + tap * cpad + c] = src[r1 * sr1 + r2 * sr2 + tap * st + c * sc] for c < cs, as bf16 hi (+ lo
 * = the bf16 remainder) or fp16.  dst_hi / dst_lo already point at the segment's K offset; strides in elements, may be
 * negative (mirrored taps of a data-gradient operand: pass src at the last tap).  Padding entries are not touched. */
B2D_API int b2d_pack_weight(const float* src, int32_t R1, int32_t R2, int64_t sr1, int64_t sr2, int32_t ntaps, int64_t st, int32_t cs, int64_t sc,
                    void* dst_hi, void* dst_lo, int64_t ktot, int32_t cpad, int32_t f16, void* stream);

/* fill helpers used by the fused loop (graph-capturable) */
B2D_API int b2d_zero(void* p, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2D_H_ */
