// First slice of the UNet training step (reference: Diffusion_model/src/predictor.py:722-748,
// unet/metrics.py:337-402, helper.py:428-430, train.py:144-148):
//   * b2d_adam_step        torch.optim.Adam's single-tensor update over a flat fp32 buffer (HBM-bound: 28 B / parameter)
//   * b2d_nmse_loss        normalized_mse_loss_per_component forward + gradient w.r.t. the prediction
//   * b2d_gn_silu_bwd      backward of y = silu(GroupNorm(1, C)(x)) (+ time embedding): dx, dgamma, dbeta, dtemb
//   * b2d_conv_wgrad       3x3 conv weight gradient dW[co][ci][ky][kx] = sum_p dY[p][co] X[p + (ky-1, kx-1)][ci] on tcgen05:
//                          both operands are read MN-major straight out of the channels-last tensors (pixels are K)
// The data gradient of a 3x3 conv needs no new kernel: it is the forward engine run on dY with the taps mirrored and the
// weight matrix transposed at pack time (engine.pack_conv2d_dgrad).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/b2d.h"
#include "b2d_internal.h"
#include "b2d_ptx.cuh"

namespace b2d {

static int grid_cap(long long work, int per_block) {
  long long b = (work + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 8;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

// ------------------------------------------------------------------------------------------- Adam
// torch/optim/adam.py _single_tensor_adam, in its operation order:
//   g' = g * grad_scale (+ wd * p);  m = m + (1 - b1) (g' - m)  [lerp];  v = v * b2 + ((1 - b2) * g') * g'  [addcmul]
//   p = p + (-lr / (1 - b1^t)) * (m / (sqrt(v) / sqrt(1 - b2^t) + eps))                           [addcdiv]
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float neg_step, float w1, float b2,
                                                   float w2, float bc2_sqrt, float eps, float wd,
                                                   float gscale) {
  const long long nvec = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto one = [&](float pp, float gg, float& mm, float& vv) {
    float gr = gscale == 1.f ? gg : __fmul_rn(gg, gscale);
    if (wd != 0.f) gr = __fadd_rn(gr, __fmul_rn(wd, pp));
    mm = __fadd_rn(mm, __fmul_rn(w1, __fsub_rn(gr, mm)));
    vv = __fadd_rn(__fmul_rn(vv, b2), __fmul_rn(__fmul_rn(w2, gr), gr));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), bc2_sqrt), eps);
    return __fadd_rn(pp, __fmul_rn(neg_step, __fdiv_rn(mm, denom)));
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    pp.x = one(pp.x, gg.x, mm.x, vv.x); pp.y = one(pp.y, gg.y, mm.y, vv.y);
    pp.z = one(pp.z, gg.z, mm.z, vv.z); pp.w = one(pp.w, gg.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = nvec * 4 + threadIdx.x;
    float mm = m[i], vv = v[i];
    p[i] = one(p[i], g[i], mm, vv);
    m[i] = mm; v[i] = vv;
  }
}

// ------------------------------------------------------------------------------------------- loss
// One block per (sample, channel) row of P elements: err = mean((o - t)^2) / (mean(t^2) + eps) [* w_c / sum w], then the
// gradient of the batch / channel mean w.r.t. o: 2 (o - t) / (P (mean(t^2) + eps)) [* w_c / sum w] / (C N).
__global__ void __launch_bounds__(256) nmse_rows_kernel(const float* __restrict__ o, const float* __restrict__ t, int C, long long P,
                                                        const float* __restrict__ weight, float eps, float* __restrict__ err,
                                                        float* __restrict__ grad, int N) {
  const int nc = blockIdx.x;
  const int c = nc % C;
  const float* orow = o + (long long)nc * P;
  const float* trow = t + (long long)nc * P;
  double s = 0.0, st = 0.0;
  for (long long i = threadIdx.x; i < P; i += blockDim.x) {
    const float d = orow[i] - trow[i];
    s += (double)d * d;
    st += (double)trow[i] * trow[i];
  }
  __shared__ double red[2][8];
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, k); st += __shfl_xor_sync(0xffffffffu, st, k); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = st; }
  __syncthreads();
  s = 0.0; st = 0.0;
  for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { s += red[0][k]; st += red[1][k]; }  // same order in every thread
  float wfac = 1.f;
  if (weight != nullptr) {
    float ws = 0.f;
    for (int k = 0; k < C; ++k) ws += weight[k];
    wfac = weight[c] / ws;
  }
  const float mse = (float)(s / (double)P), norm = (float)(st / (double)P);
  const float inv = 1.f / (norm + eps);
  if (threadIdx.x == 0) err[nc] = mse * inv * wfac;
  if (grad != nullptr) {
    const float k2 = 2.f * inv * wfac / ((float)P * (float)C * (float)N);
    float* grow = grad + (long long)nc * P;
    for (long long i = threadIdx.x; i < P; i += blockDim.x) grow[i] = k2 * (orow[i] - trow[i]);
  }
}
// loss[0] = mean over samples of loss[1 + n] = mean over channels of err[n][c]  (fixed summation order)
__global__ void nmse_finish_kernel(const float* __restrict__ err, int N, int C, float* __restrict__ loss) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float tot = 0.f;
  for (int n = 0; n < N; ++n) {
    float e = 0.f;
    for (int c = 0; c < C; ++c) e += err[n * C + c];
    e /= (float)C;
    loss[1 + n] = e;
    tot += e;
  }
  loss[0] = tot / (float)N;
}

// ------------------------------------------------------------------------------------------- GroupNorm(1, C) + SiLU backward
__device__ __forceinline__ float ld16(const uint16_t* p, long long i, int f16) {
  const uint16_t u = p[i];
  return f16 ? __half2float(__ushort_as_half(u)) : __uint_as_float((uint32_t)u << 16);
}
__device__ __forceinline__ float ld16x(const uint16_t* hi, const uint16_t* lo, long long i, int f16) {
  float v = ld16(hi, i, f16);
  if (lo != nullptr) v += ld16(lo, i, 0);
  return v;
}
__device__ __forceinline__ void st16x(uint16_t* hi, uint16_t* lo, long long i, float v, int f16) {
  if (f16) { hi[i] = __half_as_ushort(__float2half_rn(v)); return; }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = __bfloat16_as_ushort(h);
  if (lo != nullptr) lo[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
}
// z = gamma_c xhat + beta_c, a = silu(z): da/dz = sig (1 + z (1 - sig))
__device__ __forceinline__ float dsilu(float z) {
  const float sg = 1.f / (1.f + __expf(-z));
  return sg * (1.f + z * (1.f - sg));
}

struct GnBwdArgs {
  const uint16_t *x_hi, *x_lo;   // raw pre-norm activations [N][P][C]
  const uint16_t *dy_hi, *dy_lo; // upstream gradient of silu(GN(x)) (+temb), same layout
  uint16_t *dx_hi, *dx_lo;
  const double* stats;           // forward [N][2] (sum, sumsq)
  const float *gamma, *beta;
  double* sums;                  // [N][2]: sum(dxhat), sum(dxhat * xhat)
  float *dgamma, *dbeta, *dtemb; // [C], [C], [N][C] (dtemb may be NULL), atomically accumulated
  long long P;
  int C, act, x_f16, dy_f16, dx_f16;
  float eps;
};

// eight consecutive channels of a 16-bit hi (+ lo) tensor as fp32 (one 16-byte load per part)
__device__ __forceinline__ void ld16x8(const uint16_t* hi, const uint16_t* lo, long long i, int f16, float (&f)[8]) {
  const uint4 uh = __ldg(reinterpret_cast<const uint4*>(hi + i));
  const uint32_t wh[4] = {uh.x, uh.y, uh.z, uh.w};
  if (f16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t2 = __half22float2(*reinterpret_cast<const __half2*>(&wh[j]));
      f[2 * j] = t2.x; f[2 * j + 1] = t2.y;
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { f[2 * j] = __uint_as_float(wh[j] << 16); f[2 * j + 1] = __uint_as_float(wh[j] & 0xFFFF0000u); }
  if (lo != nullptr) {
    const uint4 ul = __ldg(reinterpret_cast<const uint4*>(lo + i));
    const uint32_t wl[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] += __uint_as_float(wl[j] << 16); f[2 * j + 1] += __uint_as_float(wl[j] & 0xFFFF0000u); }
  }
}
__device__ __forceinline__ void st16x8(uint16_t* hi, uint16_t* lo, long long i, int f16, const float (&f)[8]) {
  uint32_t wh[4], wl[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (f16) {
      const __half2 h = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
      wh[j] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * j]), h1 = __float2bfloat16_rn(f[2 * j + 1]);
      wh[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * j] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(f[2 * j + 1] - __bfloat162float(h1));
      wl[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
  }
  *reinterpret_cast<uint4*>(hi + i) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
  if (!f16 && lo != nullptr) *reinterpret_cast<uint4*>(lo + i) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
}

// pass 1: per-sample sums of dxhat and dxhat * xhat, per-channel dgamma / dbeta (/ dtemb).  grid (blocks per sample, N).
// A thread owns one group of 8 channels for its whole walk (the grid stride, 256 * gridDim.x vectors, is a multiple of the
// C / 8 vectors of a pixel): channel sums stay in registers and meet in shared memory once per thread.
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const GnBwdArgs a) {
  extern __shared__ float sm_ch[];  // [3][C]
  const int n = blockIdx.y, C = a.C, cv = C >> 3;
  for (int c = threadIdx.x; c < 3 * C; c += blockDim.x) sm_ch[c] = 0.f;
  __syncthreads();
  const double cnt = (double)C * (double)a.P;
  const double mean_d = a.stats[2 * n] / cnt;
  double var = a.stats[2 * n + 1] / cnt - mean_d * mean_d;
  if (var < 0) var = 0;
  const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var + (double)a.eps));
  const long long nvec = a.P * cv, base = (long long)n * a.P * C;
  const long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(v0 % cv) << 3;
  float ga[8], be[8], acc_g[8], acc_b[8], acc_t[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] = a.gamma ? a.gamma[c0 + j] : 1.f;
    be[j] = a.beta ? a.beta[c0 + j] : 0.f;
    acc_g[j] = acc_b[j] = acc_t[j] = 0.f;
  }
  float s1 = 0.f, s2 = 0.f;   // per-thread partial sums stay small (a few hundred terms); the cross-thread sums are fp64
  double d1 = 0.0, d2 = 0.0;
  int since = 0;
  for (long long v = v0; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    float x[8], dy[8];
    ld16x8(a.x_hi, a.x_lo, base + (v << 3), a.x_f16, x);
    ld16x8(a.dy_hi, a.dy_lo, base + (v << 3), a.dy_f16, dy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (x[j] - mean) * rstd;
      const float dz = a.act ? dy[j] * dsilu(fmaf(ga[j], xh, be[j])) : dy[j];
      const float dxh = dz * ga[j];
      s1 += dxh;
      s2 = fmaf(dxh, xh, s2);
      acc_g[j] = fmaf(dz, xh, acc_g[j]);
      acc_b[j] += dz;
      acc_t[j] += dy[j];
    }
    if (++since == 32) { d1 += (double)s1; d2 += (double)s2; s1 = s2 = 0.f; since = 0; }
  }
  d1 += (double)s1; d2 += (double)s2;
  if (v0 < nvec) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sm_ch[c0 + j], acc_g[j]);
      atomicAdd(&sm_ch[C + c0 + j], acc_b[j]);
      if (a.dtemb != nullptr) atomicAdd(&sm_ch[2 * C + c0 + j], acc_t[j]);
    }
  }
  __shared__ double red[2][8];
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) { d1 += __shfl_xor_sync(0xffffffffu, d1, k); d2 += __shfl_xor_sync(0xffffffffu, d2, k); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = d1; red[1][threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { t1 += red[0][k]; t2 += red[1][k]; }
    atomicAdd(a.sums + 2 * n, t1);
    atomicAdd(a.sums + 2 * n + 1, t2);
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(a.dgamma + c, sm_ch[c]);
    atomicAdd(a.dbeta + c, sm_ch[C + c]);
    if (a.dtemb != nullptr) atomicAdd(a.dtemb + (long long)n * C + c, sm_ch[2 * C + c]);
  }
}
// pass 2: dx = rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat))
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const GnBwdArgs a) {
  const int n = blockIdx.y, C = a.C, cv = C >> 3;
  const double cnt = (double)C * (double)a.P;
  const double mean_d = a.stats[2 * n] / cnt;
  double var = a.stats[2 * n + 1] / cnt - mean_d * mean_d;
  if (var < 0) var = 0;
  const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var + (double)a.eps));
  const float m1 = (float)(a.sums[2 * n] / cnt), m2 = (float)(a.sums[2 * n + 1] / cnt);
  const long long nvec = a.P * cv, base = (long long)n * a.P * C;
  const long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(v0 % cv) << 3;
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ga[j] = a.gamma ? a.gamma[c0 + j] : 1.f; be[j] = a.beta ? a.beta[c0 + j] : 0.f; }
  for (long long v = v0; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    float x[8], dy[8], o[8];
    ld16x8(a.x_hi, a.x_lo, base + (v << 3), a.x_f16, x);
    ld16x8(a.dy_hi, a.dy_lo, base + (v << 3), a.dy_f16, dy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (x[j] - mean) * rstd;
      const float dz = a.act ? dy[j] * dsilu(fmaf(ga[j], xh, be[j])) : dy[j];
      o[j] = rstd * (dz * ga[j] - m1 - xh * m2);
    }
    st16x8(a.dx_hi, a.dx_lo, base + (v << 3), a.dx_f16, o);
  }
}

// ------------------------------------------------------------------------------------------- weight gradients (tcgen05)
// dW[m][n][ty][tx] = sum over pixels p of A[p][m] * B[s * p + (tx, ty) + pad][n]: the weight gradient of
//   * Conv2d 3x3 / pad 1   (A = dY, B = X, pad -1, s 1):  dW[co][ci][ky][kx]
//   * Linear / Conv1d k1   (A = dY, B = X, one tap):      dW[out][in]
//   * ConvTranspose2d k2s2 (A = X,  B = dY, pad 0, s 2):  dW[ci][co][ky][kx]   (B is read with a TMA element stride of 2)
// M = A channels (tiles of 128), N = B channels (tiles of 128), K = pixels: both operands are MN-major tiles of 32 pixels x
// 64 channels (128B swizzle) loaded straight out of the channels-last tensors; TMA zero fill = the conv's zero padding.
// One CTA = (M tile, N tile, tap row ty, pixel split): the taps of a row share the A tile and own 128 TMEM columns each.
constexpr int kWgK = 32;                       // pixels per K chunk (two K = 16 MMAs)
constexpr int kWgAtom = kWgK * 128;            // bytes of one [32 px][64 ch] swizzled tile
constexpr int kWgTile = 2 * kWgAtom;           // 128 channels
constexpr int kWgStages = 3;
constexpr int kWgThreads = 128;
constexpr uint32_t kUmmaAMajorMN = 1u << 15;   // instruction-descriptor bit: A operand is MN-major
constexpr uint32_t kUmmaBMajorMN2 = 1u << 16;

struct WgradParams {
  CUtensorMap tmA[2];  // A hi / lo: dims (Cpad, W, H, 1, N), box (64, bw, bh, 1, bn), bw * bh * bn = 32
  CUtensorMap tmB[2];  // B hi / lo (element strides (1, s, s, 1, 1), box extents bw * s, bh * s)
  float* dw;           // fp32, atomically accumulated: element (m, n, ty, tx) at m * sm + (n_off + n) * sn + ty * sty + tx * stx
  long long sm, sn, sty, stx;   // reference layout [m][n][KH][KW], or "channels last" [m][KH][KW][n] (sn = 1: 16-byte atomics)
  int m_valid, n_valid, n_off, n_total;
  int KH, KW, bpad, bstride;
  int lbw, lbh, lbn, tiles_w, tiles_h, tiles_n;
  int chunks, chunks_per_cta, nsrc, op_f16, n_tiles;
};

__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t wg_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wg_smem_raw) + 1023) & ~uintptr_t(1023));
  const int nsrc = p.nsrc, KW = p.KW;
  const int src_bytes = (1 + KW) * kWgTile;    // the A tile + KW tap tiles of B
  const int stage_bytes = nsrc * src_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWgStages * nsrc * 4 * kWgTile);
  uint64_t* empty = full + kWgStages;
  uint64_t* done = empty + kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5;
  const int m0 = (blockIdx.x / p.n_tiles) * 128, n0c = (blockIdx.x % p.n_tiles) * 128;
  const int ty = blockIdx.y;
  const int ch_lo = blockIdx.z * p.chunks_per_cta;
  const int ch_hi = min(p.chunks, ch_lo + p.chunks_per_cta);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    for (int s = 0; s < nsrc; ++s) { prefetch_tmap(&p.tmA[s]); prefetch_tmap(&p.tmB[s]); }
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    // ---------------- TMA producer: per chunk, the A tile and the KW shifted B tiles of this tap row ----------------
    int st = 0;
    uint32_t ph = 0;
    for (int ch = ch_lo; ch < ch_hi; ++ch) {
      int t = ch;
      const int x0 = (t % p.tiles_w) << p.lbw; t /= p.tiles_w;
      const int y0 = (t % p.tiles_h) << p.lbh; t /= p.tiles_h;
      const int n0 = t << p.lbn;
      mbar_wait(&empty[st], ph ^ 1);
      mbar_arrive_expect_tx(&full[st], (uint32_t)stage_bytes);
      uint8_t* sb = smem + st * (nsrc * 4 * kWgTile);
      for (int s = 0; s < nsrc; ++s) {
        uint8_t* at = sb + s * (4 * kWgTile);
        for (int a = 0; a < 2; ++a) tma_load_5d(at + a * kWgAtom, &p.tmA[s], &full[st], m0 + 64 * a, x0, y0, 0, n0);
        for (int tx = 0; tx < KW; ++tx)
          for (int a = 0; a < 2; ++a)
            tma_load_5d(at + (1 + tx) * kWgTile + a * kWgAtom, &p.tmB[s], &full[st], n0c + 64 * a, x0 * p.bstride + tx + p.bpad,
                        y0 * p.bstride + ty + p.bpad, 0, n0);
      }
      if (++st == kWgStages) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    // ---------------- MMA issuer: D_tx[m][n] += A^T B_tx over the chunk's 32 pixels (both operands MN-major) ----------
    const uint32_t idesc = umma_idesc_16(128, 128, p.op_f16) | kUmmaAMajorMN | kUmmaBMajorMN2;
    int st = 0;
    uint32_t ph = 0, accum = 0;
    for (int ch = ch_lo; ch < ch_hi; ++ch) {
      mbar_wait(&full[st], ph);
      tc_fence_after();
      const uint32_t sb = smem_u32(smem + st * (nsrc * 4 * kWgTile));
      // products: single source (a, b); split sources hi*hi + hi*lo + lo*hi
      const int nprod = nsrc == 2 ? 3 : 1;
      for (int pr = 0; pr < nprod; ++pr) {
        const int as = pr == 2 ? 1 : 0, bs = pr == 1 ? 1 : 0;
        const uint32_t aa = sb + as * (4 * kWgTile);
        const uint32_t ba = sb + bs * (4 * kWgTile) + kWgTile;
        for (int tx = 0; tx < KW; ++tx) {
#pragma unroll
          for (int k = 0; k < kWgK / 16; ++k) {
            const uint64_t ad = umma_smem_desc_mn(aa + k * 16 * 128, kWgAtom, 1024, 2);
            const uint64_t bd = umma_smem_desc_mn(ba + tx * kWgTile + k * 16 * 128, kWgAtom, 1024, 2);
            umma_bf16(tmem_base + tx * 128, ad, bd, idesc, accum | (uint32_t)(pr > 0) | (uint32_t)(k > 0));
          }
        }
      }
      accum = 1;
      umma_commit(&empty[st]);
      if (++st == kWgStages) { st = 0; ph ^= 1; }
    }
    umma_commit(done);
  }
  __syncwarp();
  // ---------------- epilogue: all four warps, thread = output row m (TMEM lane) ----------------
  if (ch_hi > ch_lo) {
    mbar_wait(done, 0);
    tc_fence_after();
    const int m = m0 + (int)threadIdx.x;
    const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16);
    for (int tx = 0; tx < KW; ++tx) {
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + tx * 128 + c0, v);
        tmem_ld_wait();
        if (m < p.m_valid) {
          float* row = p.dw + (long long)m * p.sm + (long long)ty * p.sty + (long long)tx * p.stx + (long long)(p.n_off + n0c + c0) * p.sn;
          if (p.sn == 1 && n0c + c0 + 32 <= p.n_valid && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
            // the 32 columns are contiguous: eight 16-byte reductions instead of 32 scalar ones
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + 4 * q), "f"(__uint_as_float(v[4 * q])),
                           "f"(__uint_as_float(v[4 * q + 1])), "f"(__uint_as_float(v[4 * q + 2])), "f"(__uint_as_float(v[4 * q + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0c + c0 + j < p.n_valid) atomicAdd(row + (long long)j * p.sn, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------- small elementwise / reduction passes
// out[c] += sum over rows of x[row][c]   (bias gradients: Conv2d / ConvTranspose2d / Linear bias = sum of dY over pixels)
__global__ void __launch_bounds__(256) channel_sum_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int f16,
                                                          long long rows, int C, int cvalid, float* __restrict__ out) {
  // 256 threads: thread -> (channel c0 + tid % Cb, row group tid / Cb), Cb = min(C, 256); leftover threads idle
  const int Cb = C < 256 ? C : 256;
  const int rg = 256 / Cb;
  const int r0 = (int)threadIdx.x / Cb;
  if (r0 >= rg) return;
  for (int c0 = 0; c0 < cvalid; c0 += Cb) {
    const int c = c0 + (int)threadIdx.x % Cb;
    if (c >= cvalid) continue;
    float acc = 0.f;
    for (long long r = (long long)blockIdx.x * rg + r0; r < rows; r += (long long)gridDim.x * rg) acc += ld16x(hi, lo, r * C + c, f16);
    atomicAdd(out + c, acc);
  }
}
// out = a + b, 16-bit hi (+ lo) storage: the two uses of a skip connection meet here (unet/models.py:150-177)
__global__ void __launch_bounds__(256) add16_kernel(const uint16_t* a_hi, const uint16_t* a_lo, const uint16_t* b_hi, const uint16_t* b_lo,
                                                    uint16_t* o_hi, uint16_t* o_lo, int f16, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    st16x(o_hi, o_lo, i, ld16x(a_hi, a_lo, i, f16) + ld16x(b_hi, b_lo, i, f16), f16);
}
// MaxPool2d(2, 2) backward (unet/blocks.py:161-164): the gradient of a pooled pixel goes to the FIRST maximum of its 2x2
// window in scan order (torch keeps the index of the first strictly greater value)
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* dy_hi, const uint16_t* dy_lo,
                                                          uint16_t* dx_hi, uint16_t* dx_lo, int f16, long long N, int H, int W, int C) {
  const int OH = H >> 1, OW = W >> 1;
  const long long total = N * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH);
    const long long n = t / OH;
    float best = 0.f;
    int arg = 0;
    long long idx[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      idx[q] = ((n * H + 2 * oy + (q >> 1)) * W + 2 * ox + (q & 1)) * C + c;
      const float v = ld16x(x_hi, x_lo, idx[q], f16);
      if (q == 0 || v > best) { best = v; arg = q; }
    }
    const float g = ld16x(dy_hi, dy_lo, i, f16);
#pragma unroll
    for (int q = 0; q < 4; ++q) st16x(dx_hi, dx_lo, idx[q], q == arg ? g : 0.f, f16);
  }
}

// ------------------------------------------------------------------------------------------- operand refresh
// After an optimizer step the engine's packed operands ([rows][K] 16-bit, K = kbase + tap * cpad + c; bf16 hi | lo or
// fp16) are rewritten IN PLACE from the fp32 parameters, so plans (whose TMA descriptors hold the operand addresses) and
// activation buffers stay valid across steps.  The source is addressed through strides, which covers every form the
// step uses: forward (rows = Cout), data gradient (rows = Cin, taps mirrored: negative tap stride), transposed
// matrices, and the phase-major rows of ConvTranspose2d (row = r1 * R2 + r2).
struct PackArgs {
  const float* src;
  uint16_t *hi, *lo;
  long long sr1, sr2, st, sc, ktot;
  int R2, ntaps, cs, cpad, f16;
  long long total;
};
__global__ void __launch_bounds__(256) pack_weight_kernel(const PackArgs a) {
  // thread = (row, c), looping over the taps: for the forward forms (c stride = ntaps, tap stride 1) a warp reads one
  // contiguous run of the parameter and writes ntaps contiguous runs of the operand
  const long long n = a.total / a.ntaps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % a.cs);
    const long long row = i / a.cs;
    const long long r1 = row / a.R2, r2 = row - r1 * a.R2;
    const float* sp = a.src + r1 * a.sr1 + r2 * a.sr2 + c * a.sc;
    const long long d0 = row * a.ktot + c;
    for (int tap = 0; tap < a.ntaps; ++tap) st16x(a.hi, a.lo, d0 + (long long)tap * a.cpad, sp[tap * a.st], a.f16);
  }
}

// The transposing form: the source is contiguous along the operand's ROWS (sr2 == 1: data-gradient operands of channels-last
// parameters, transposed matrices, the phase-major ConvTranspose2d rows) -- 32 x 32 tiles through shared memory so that
// both the fp32 reads and the 16-bit writes are coalesced.  grid = (row tiles, column tiles, R1 * ntaps).
__global__ void __launch_bounds__(256) pack_weight_t_kernel(const PackArgs a) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r1 = blockIdx.z / a.ntaps, tap = blockIdx.z - r1 * a.ntaps;
  const int row0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* sp = a.src + r1 * a.sr1 + tap * a.st;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + ty + 8 * j, r2 = row0 + tx;
    tile[ty + 8 * j][tx] = (c < a.cs && r2 < a.R2) ? sp[r2 + c * a.sc] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r2 = row0 + ty + 8 * j, c = c0 + tx;
    if (r2 < a.R2 && c < a.cs) st16x(a.hi, a.lo, ((long long)r1 * a.R2 + r2) * a.ktot + (long long)tap * a.cpad + c, tile[tx][ty + 8 * j], a.f16);
  }
}

// ------------------------------------------------------------------------------------------- attention core backward
// softmax(q k^T / sqrt(d)) v per (image, head) (nn.MultiheadAttention's core, unet/blocks.py:196-227), fp32 CUDA cores:
// T <= 256 tokens and the layer is ~1% of the step's flops, so exact fp32 arithmetic is worth more here than tensor cores.
//   pass 0 (per query block): lse_i = logsumexp_j(s_ij), D_i = dO_i . O_i   (= sum_j P_ij dP_ij)
//   pass 1 (per query block): dQ_i = scale * sum_j dS_ij K_j,   dS_ij = P_ij (dO_i . V_j - D_i),  P_ij = exp(s_ij - lse_i)
//   pass 2 (per key block):   dK_j = scale * sum_i dS_ij Q_i,   dV_j = sum_i P_ij dO_i
// One CTA = 16 "owner" rows (queries in pass 0/1, keys in pass 2) against all "other" rows in tiles of 16; thread (a, b)
// computes the (owner a, other b) entry of the 16 x 16 score tile, then accumulates columns b, b + 16, ... of owner row a.
constexpr int kAbR = 16;
struct AttnBwdArgs {
  const uint16_t *qkv_hi, *qkv_lo, *o_hi, *o_lo, *do_hi, *do_lo;
  uint16_t *dqkv_hi, *dqkv_lo;
  float* stats;  // [N][heads][T][2]: lse, D
  int T, C, heads, d, f16;
  float scale;
};
template <int MODE>
__global__ void __launch_bounds__(256, 2) attn_bwd_kernel(const AttnBwdArgs a) {
  extern __shared__ float ab_smem[];
  const int d = a.d, ld = d + 4, T = a.T, C = a.C;   // rows stay 16-byte aligned: every inner loop reads float4
  float* own1 = ab_smem;             // [16][ld]  pass 0/1: Q   pass 2: K
  float* own2 = own1 + kAbR * ld;    //           pass 0/1: dO  pass 2: V
  float* oth1 = own2 + kAbR * ld;    //           pass 0/1: K   pass 2: Q
  float* oth2 = oth1 + kAbR * ld;    //           pass 0/1: V   pass 2: dO
  float* tile_ds = oth2 + kAbR * ld; // [16][20]
  float* tile_p = tile_ds + kAbR * 20;
  const int n = blockIdx.z, h = blockIdx.y, r0 = blockIdx.x * kAbR;
  const int ta = threadIdx.x >> 4, tb = threadIdx.x & 15;
  const long long qkv_row = 3LL * C;
  const uint16_t* qh = a.qkv_hi + (long long)n * T * qkv_row + h * d;
  const uint16_t* ql = a.qkv_lo ? a.qkv_lo + (long long)n * T * qkv_row + h * d : nullptr;
  const long long obase = (long long)n * T * C + h * d;
  float* st = a.stats + ((long long)(n * a.heads + h) * T) * 2;
  // which = 0 q, 1 k, 2 v, 3 dO, 4 O.  One 16-byte load (8 channels) of the hi part and one of the lo part per thread and
  // iteration, all issued before the first use: the row tiles are latency-, not bandwidth-bound
  auto load_rows = [&](float* dst, int which, int row0) {
    const uint16_t* ph = which < 3 ? qh + which * C : which == 3 ? a.do_hi + obase : a.o_hi + obase;
    const uint16_t* pl = which < 3 ? (ql ? ql + which * C : nullptr) : which == 3 ? (a.do_lo ? a.do_lo + obase : nullptr) : (a.o_lo ? a.o_lo + obase : nullptr);
    const long long rs = which < 3 ? qkv_row : (long long)C;
    const int vpr = d >> 3;
    for (int v = threadIdx.x; v < kAbR * vpr; v += 256) {
      const int r = v / vpr, k = (v - r * vpr) << 3;
      float f[8];
      if (row0 + r >= T) {   // ragged T (2 x 2 maps): rows past the end are zero and masked below
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      } else {
        const long long idx = (long long)(row0 + r) * rs + k;
        const uint4 uh = __ldg(reinterpret_cast<const uint4*>(ph + idx));
        uint4 ul = make_uint4(0u, 0u, 0u, 0u);
        if (pl != nullptr) ul = __ldg(reinterpret_cast<const uint4*>(pl + idx));
        const uint32_t wh[4] = {uh.x, uh.y, uh.z, uh.w}, wl[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (a.f16) {
            const float2 t2 = __half22float2(*reinterpret_cast<const __half2*>(&wh[j]));
            f[2 * j] = t2.x; f[2 * j + 1] = t2.y;
          } else {
            f[2 * j] = __uint_as_float(wh[j] << 16) + __uint_as_float(wl[j] << 16);
            f[2 * j + 1] = __uint_as_float(wh[j] & 0xFFFF0000u) + __uint_as_float(wl[j] & 0xFFFF0000u);
          }
        }
      }
      float4* o = reinterpret_cast<float4*>(dst + r * ld + k);
      o[0] = make_float4(f[0], f[1], f[2], f[3]);
      o[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
  };
  if (MODE == 2) { load_rows(own1, 1, r0); load_rows(own2, 2, r0); }
  else { load_rows(own1, 0, r0); load_rows(own2, 3, r0); }
  // thread (a, b) accumulates columns 4b .. 4b + 3 (+ 64 kk) of owner row a
  float4 acc1[8], acc2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc1[i] = acc2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float run_m = -INFINITY, run_l = 0.f;
  float lse_own = 0.f, d_own = 0.f;
  const bool own_ok = r0 + ta < T;
  if (MODE == 1 && own_ok) { lse_own = st[(r0 + ta) * 2]; d_own = st[(r0 + ta) * 2 + 1]; }
  for (int o0 = 0; o0 < T; o0 += kAbR) {
    __syncthreads();
    if (MODE == 2) { load_rows(oth1, 0, o0); load_rows(oth2, 3, o0); }
    else { load_rows(oth1, 1, o0); if (MODE == 1) load_rows(oth2, 2, o0); }
    __syncthreads();
    float s = 0.f, dp = 0.f;
    {
      const float4* o1 = reinterpret_cast<const float4*>(own1 + ta * ld);
      const float4* t1 = reinterpret_cast<const float4*>(oth1 + tb * ld);
      const float4* o2 = reinterpret_cast<const float4*>(own2 + ta * ld);
      const float4* t2 = reinterpret_cast<const float4*>(oth2 + tb * ld);
#pragma unroll 4
      for (int k4 = 0; k4 < d / 4; ++k4) {
        const float4 x = o1[k4], y = t1[k4];
        s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
        if (MODE != 0) {
          const float4 u = o2[k4], v = t2[k4];
          dp = fmaf(u.x, v.x, dp); dp = fmaf(u.y, v.y, dp); dp = fmaf(u.z, v.z, dp); dp = fmaf(u.w, v.w, dp);
        }
      }
    }
    s *= a.scale;
    const bool oth_ok = o0 + tb < T;
    if (MODE == 0) {
      if (!oth_ok) s = -INFINITY;
      float m = s;
      for (int off = 8; off; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      const float nm = fmaxf(run_m, m);
      float e = expf(s - nm);
      for (int off = 8; off; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
      run_l = run_l * expf(run_m - nm) + e;
      run_m = nm;
    } else {
      float lse = lse_own, dd = d_own;
      if (MODE == 2 && oth_ok) { lse = st[(o0 + tb) * 2]; dd = st[(o0 + tb) * 2 + 1]; }
      const float pr = oth_ok ? expf(s - lse) : 0.f;
      tile_p[ta * 20 + tb] = pr;
      tile_ds[ta * 20 + tb] = pr * (dp - dd) * a.scale;
      __syncthreads();
#pragma unroll
      for (int b4 = 0; b4 < kAbR / 4; ++b4) {
        const float4 w4 = reinterpret_cast<const float4*>(tile_ds + ta * 20)[b4];
        const float4 q4 = MODE == 2 ? reinterpret_cast<const float4*>(tile_p + ta * 20)[b4] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float wds[4] = {w4.x, w4.y, w4.z, w4.w}, wp[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int k = 4 * tb + 64 * kk;
          if (k >= d) break;
          float4 x1 = acc1[kk], x2 = acc2[kk];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int b = 4 * b4 + j;
            const float4 o = *reinterpret_cast<const float4*>(oth1 + b * ld + k);
            x1.x = fmaf(wds[j], o.x, x1.x); x1.y = fmaf(wds[j], o.y, x1.y); x1.z = fmaf(wds[j], o.z, x1.z); x1.w = fmaf(wds[j], o.w, x1.w);
            if (MODE == 2) {
              const float4 u = *reinterpret_cast<const float4*>(oth2 + b * ld + k);
              x2.x = fmaf(wp[j], u.x, x2.x); x2.y = fmaf(wp[j], u.y, x2.y); x2.z = fmaf(wp[j], u.z, x2.z); x2.w = fmaf(wp[j], u.w, x2.w);
            }
          }
          acc1[kk] = x1; acc2[kk] = x2;
        }
      }
    }
  }
  if (MODE == 0) {
    // D_a = dO_a . O_a: the 16 lanes of row a split the head dimension
    __syncthreads();
    load_rows(oth1, 4, r0);
    __syncthreads();
    float dd = 0.f;
    for (int k = tb; k < d; k += 16) dd = fmaf(own2[ta * ld + k], oth1[ta * ld + k], dd);
    for (int off = 8; off; off >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, off);
    if (tb == 0 && own_ok) { st[(r0 + ta) * 2] = run_m + logf(run_l); st[(r0 + ta) * 2 + 1] = dd; }
  } else if (own_ok) {
    uint16_t* oh = a.dqkv_hi + (long long)n * T * qkv_row + h * d + (long long)(r0 + ta) * qkv_row;
    uint16_t* ol = a.dqkv_lo ? a.dqkv_lo + (long long)n * T * qkv_row + h * d + (long long)(r0 + ta) * qkv_row : nullptr;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k = 4 * tb + 64 * kk;
      if (k >= d) break;
      const float v1[4] = {acc1[kk].x, acc1[kk].y, acc1[kk].z, acc1[kk].w}, v2[4] = {acc2[kk].x, acc2[kk].y, acc2[kk].z, acc2[kk].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE == 1) st16x(oh, ol, k + j, v1[j], a.f16);
        else { st16x(oh, ol, C + k + j, v1[j], a.f16); st16x(oh, ol, 2 * C + k + j, v2[j], a.f16); }
      }
    }
  }
}

typedef CUresult (*PFN_encodeTiledT)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int ilog2_floor(int x) {
  int l = 0;
  while ((2 << l) <= x) ++l;
  return l;
}

}  // namespace b2d

using namespace b2d;

extern "C" int b2d_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                             double beta2, double eps, double weight_decay, int32_t step, double grad_scale, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n < 1 || step < 1) return set_error(B2D_E_INVALID, "b2d_adam_step: bad argument");
  if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15)
    return set_error(B2D_E_INVALID, "b2d_adam_step: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  adam_kernel<<<grid_cap(n / 4 + 1, 256 * 2), 256, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, (long long)n, (float)(-(lr / bc1)), (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
      (float)sqrt(bc2), (float)eps, (float)weight_decay, (float)grad_scale);
  return check_launch("adam_kernel");
}

extern "C" int b2d_nmse_loss(const float* pred, const float* target, int32_t N, int32_t C, int64_t P, const float* weight, float eps,
                             float* err, float* loss, float* grad, void* stream) {
  if (!pred || !target || !err || !loss || N < 1 || C < 1 || P < 1 || (long long)N * C > 0x7fffffff)
    return set_error(B2D_E_INVALID, "b2d_nmse_loss: bad argument");
  nmse_rows_kernel<<<N * C, 256, 0, (cudaStream_t)stream>>>(pred, target, C, (long long)P, weight, eps, err, grad, N);
  nmse_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(err, N, C, loss);
  return check_launch("nmse_loss kernels");
}

extern "C" int b2d_gn_silu_bwd(const void* x_hi, const void* x_lo, int32_t x_f16, const void* dy_hi, const void* dy_lo, int32_t dy_f16,
                               void* dx_hi, void* dx_lo, int32_t dx_f16, int32_t N, int64_t P, int32_t C, const double* stats,
                               const float* gamma, const float* beta, float eps, int32_t act, double* sums, float* dgamma, float* dbeta,
                               float* dtemb, void* stream) {
  if (!x_hi || !dy_hi || !dx_hi || !stats || !sums || !dgamma || !dbeta || N < 1 || N > 65535 || P < 1 || C < 1 || 3 * C * 4 > 48 * 1024)
    return set_error(B2D_E_INVALID, "b2d_gn_silu_bwd: bad argument");
  if ((x_f16 && x_lo) || (dy_f16 && dy_lo) || (dx_f16 && dx_lo)) return set_error(B2D_E_INVALID, "b2d_gn_silu_bwd: an fp16 tensor has no lo part");
  GnBwdArgs a;
  a.x_hi = (const uint16_t*)x_hi; a.x_lo = (const uint16_t*)x_lo; a.dy_hi = (const uint16_t*)dy_hi; a.dy_lo = (const uint16_t*)dy_lo;
  a.dx_hi = (uint16_t*)dx_hi; a.dx_lo = (uint16_t*)dx_lo; a.stats = stats; a.gamma = gamma; a.beta = beta; a.sums = sums;
  a.dgamma = dgamma; a.dbeta = dbeta; a.dtemb = dtemb; a.P = P; a.C = C; a.act = act ? 1 : 0;
  a.x_f16 = x_f16 ? 1 : 0; a.dy_f16 = dy_f16 ? 1 : 0; a.dx_f16 = dx_f16 ? 1 : 0; a.eps = eps;
  if (C % 8) return set_error(B2D_E_UNSUPPORTED, "b2d_gn_silu_bwd: C = %d must be a multiple of 8 (16-byte channel vectors)", C);
  int bx = grid_cap(P * C / 8, 256 * 4);
  const int cap = (num_sms() * 8 + N - 1) / N;
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  // a thread keeps one 8-channel group for its whole walk: the grid stride (256 * bx vectors) must be a multiple of C / 8
  int cv = C / 8, g = 256, r = cv;
  while (r) { const int t = g % r; g = r; r = t; }   // g = gcd(256, cv)
  const int step = cv / g;
  bx = bx < step ? step : bx / step * step;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * N, st);
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "b2d_gn_silu_bwd: memset: %s", cudaGetErrorString(e));
  gn_bwd_reduce_kernel<<<dim3(bx, N), 256, 3 * C * sizeof(float), st>>>(a);
  gn_bwd_apply_kernel<<<dim3(bx, N), 256, 0, st>>>(a);
  return check_launch("gn_silu_bwd kernels");
}

extern "C" int b2d_conv_wgrad(int32_t kind, const void* dy_hi, const void* dy_lo, int32_t cout_pad, const void* x_hi, const void* x_lo,
                              int32_t cin_pad, int32_t N, int32_t H, int32_t W, int32_t cout, int32_t cin, int32_t cin_off, int32_t cin_total,
                              float* dw, int32_t dw_channels_last, int32_t op_f16, void* stream) {
  // H, W: extent of X (for kind 2, ConvTranspose2d k2s2, dY is 2H x 2W)
  if (!dy_hi || !x_hi || !dw || N < 1 || H < 1 || W < 1 || cout < 1 || cin < 1 || cin_off < 0 || cin_off + cin > cin_total ||
      (cout_pad % 64) || (cin_pad % 64) || cout > cout_pad || cin > cin_pad || ((dy_lo == nullptr) != (x_lo == nullptr)) || kind < 0 || kind > 2)
    return set_error(B2D_E_INVALID, "b2d_conv_wgrad: bad argument");
  if ((W & (W - 1)) || (H & (H - 1))) return set_error(B2D_E_UNSUPPORTED, "b2d_conv_wgrad: H, W must be powers of two (got %dx%d)", H, W);
  if (op_f16 && dy_lo) return set_error(B2D_E_INVALID, "b2d_conv_wgrad: fp16 operands have no lo part");
  if (kind == 2 && cin_off != 0) return set_error(B2D_E_INVALID, "b2d_conv_wgrad: transposed conv takes no channel offset");
  PFN_encodeTiledT enc = reinterpret_cast<PFN_encodeTiledT>(tensor_map_encode_fn());
  if (!enc) return set_error(B2D_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available (no CUDA driver / too old)");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int lw = ilog2_floor(W); if (lw > 3) lw = 3;
  int lh = ilog2_floor(H); if (lh > 5 - lw) lh = 5 - lw;
  const int ln = 5 - lw - lh;
  p.lbw = lw; p.lbh = lh; p.lbn = ln;
  p.tiles_w = W >> lw; p.tiles_h = H >> lh; p.tiles_n = (N + (1 << ln) - 1) >> ln;
  p.chunks = p.tiles_w * p.tiles_h * p.tiles_n;
  p.nsrc = dy_lo ? 2 : 1;
  p.op_f16 = op_f16 ? 1 : 0;
  p.dw = dw;
  // roles: A = rows of dW, B = columns (read at the tap positions)
  const void* a_src[2]; const void* b_src[2];
  int a_pad, b_pad, a_W = W, a_H = H, b_W = W, b_H = H;
  if (kind == 2) {  // dW[ci][co][ky][kx]: A = X, B = dY (2H x 2W, element stride 2)
    a_src[0] = x_hi; a_src[1] = x_lo; b_src[0] = dy_hi; b_src[1] = dy_lo; a_pad = cin_pad; b_pad = cout_pad;
    p.m_valid = cin; p.n_valid = cout; p.n_off = 0; p.n_total = cout; p.KH = 2; p.KW = 2; p.bpad = 0; p.bstride = 2;
    b_W = 2 * W; b_H = 2 * H;
  } else {          // dW[co][ci][ky][kx]: A = dY, B = X
    a_src[0] = dy_hi; a_src[1] = dy_lo; b_src[0] = x_hi; b_src[1] = x_lo; a_pad = cout_pad; b_pad = cin_pad;
    p.m_valid = cout; p.n_valid = cin; p.n_off = cin_off; p.n_total = cin_total; p.bstride = 1;
    if (kind == 0) { p.KH = 3; p.KW = 3; p.bpad = -1; } else { p.KH = 1; p.KW = 1; p.bpad = 0; }
  }
  for (int s = 0; s < p.nsrc; ++s) {
    for (int which = 0; which < 2; ++which) {
      const int cpad = which == 0 ? a_pad : b_pad;
      const int tw = which == 0 ? a_W : b_W, th = which == 0 ? a_H : b_H;
      const int es = which == 0 ? 1 : p.bstride;
      cuuint64_t dims[5] = {(cuuint64_t)cpad, (cuuint64_t)tw, (cuuint64_t)th, 1, (cuuint64_t)N};
      cuuint64_t strides[4] = {(cuuint64_t)cpad * 2, (cuuint64_t)cpad * 2 * tw, (cuuint64_t)cpad * 2 * tw * th, (cuuint64_t)cpad * 2 * tw * th};
      cuuint32_t box[5] = {64, (cuuint32_t)((1 << lw) * es), (cuuint32_t)((1 << lh) * es), 1, (cuuint32_t)(1 << ln)};
      cuuint32_t estr[5] = {1, (cuuint32_t)es, (cuuint32_t)es, 1, 1};
      CUresult r = enc(which == 0 ? &p.tmA[s] : &p.tmB[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(which == 0 ? a_src[s] : b_src[s]),
                       dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return set_error(B2D_E_CUDA, "b2d_conv_wgrad: cuTensorMapEncodeTiled failed: %d", (int)r);
    }
  }
  if (dw_channels_last) { p.sn = 1; p.stx = p.n_total; p.sty = (long long)p.KW * p.n_total; p.sm = (long long)p.KH * p.KW * p.n_total; }
  else { p.stx = 1; p.sty = p.KW; p.sn = (long long)p.KH * p.KW; p.sm = (long long)p.n_total * p.KH * p.KW; }
  const int m_tiles = (p.m_valid + 127) / 128;
  p.n_tiles = (p.n_valid + 127) / 128;
  int splits = (2 * num_sms()) / (m_tiles * p.n_tiles * p.KH);
  if (splits < 1) splits = 1;
  if (splits > p.chunks) splits = p.chunks;
  p.chunks_per_cta = (p.chunks + splits - 1) / splits;
  splits = (p.chunks + p.chunks_per_cta - 1) / p.chunks_per_cta;
  const int smem = kWgStages * p.nsrc * kWgTile * 4 + 256 + 1024;
  static unsigned long long configured = 0;
  cudaError_t e = smem_attr_once(conv_wgrad_kernel, 227 * 1024, configured);
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "b2d_conv_wgrad: smem attr: %s", cudaGetErrorString(e));
  conv_wgrad_kernel<<<dim3(m_tiles * p.n_tiles, p.KH, splits), kWgThreads, smem, (cudaStream_t)stream>>>(p);
  return check_launch("conv_wgrad_kernel");
}

extern "C" int b2d_channel_sum(const void* x_hi, const void* x_lo, int32_t f16, int64_t rows, int32_t C, int32_t cvalid, float* out, void* stream) {
  if (!x_hi || !out || rows < 1 || C < 1 || cvalid < 1 || cvalid > C || (f16 && x_lo))
    return set_error(B2D_E_INVALID, "b2d_channel_sum: bad argument");
  const int rg = C < 256 ? 256 / C : 1;
  channel_sum_kernel<<<grid_cap((rows + rg - 1) / rg, 8), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)x_hi, (const uint16_t*)x_lo, f16 ? 1 : 0,
                                                                                       (long long)rows, C, cvalid, out);
  return check_launch("channel_sum_kernel");
}

extern "C" int b2d_add16(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, void* o_hi, void* o_lo, int32_t f16, int64_t n,
                         void* stream) {
  if (!a_hi || !b_hi || !o_hi || n < 1 || (f16 && (a_lo || b_lo || o_lo))) return set_error(B2D_E_INVALID, "b2d_add16: bad argument");
  add16_kernel<<<grid_cap(n, 1024), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)a_hi, (const uint16_t*)a_lo, (const uint16_t*)b_hi,
                                                                    (const uint16_t*)b_lo, (uint16_t*)o_hi, (uint16_t*)o_lo, f16 ? 1 : 0, (long long)n);
  return check_launch("add16_kernel");
}

extern "C" int b2d_maxpool2x2_bwd(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, void* dx_hi, void* dx_lo, int32_t f16,
                                  int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  if (!x_hi || !dy_hi || !dx_hi || N < 1 || H < 2 || W < 2 || (H & 1) || (W & 1) || C < 1 || (f16 && (x_lo || dy_lo || dx_lo)))
    return set_error(B2D_E_INVALID, "b2d_maxpool2x2_bwd: bad argument");
  maxpool_bwd_kernel<<<grid_cap((long long)N * (H / 2) * (W / 2) * C, 1024), 256, 0, (cudaStream_t)stream>>>(
      (const uint16_t*)x_hi, (const uint16_t*)x_lo, (const uint16_t*)dy_hi, (const uint16_t*)dy_lo, (uint16_t*)dx_hi, (uint16_t*)dx_lo, f16 ? 1 : 0,
      (long long)N, H, W, C);
  return check_launch("maxpool_bwd_kernel");
}

extern "C" int b2d_attention_bwd(const void* qkv, const void* qkv_lo, const void* out, const void* out_lo, const void* dout, const void* dout_lo,
                                 void* dqkv, void* dqkv_lo, float* stats, int32_t N, int32_t T, int32_t C, int32_t heads, int32_t f16, void* stream) {
  if (!qkv || !out || !dout || !dqkv || !stats || N < 1 || T < 1 || heads < 1 || C < 16 || (C % heads) || (f16 && (qkv_lo || dqkv_lo)))
    return set_error(B2D_E_INVALID, "b2d_attention_bwd: bad argument");
  const int d = C / heads;
  if ((d % 16) || d > 512) return set_error(B2D_E_UNSUPPORTED, "b2d_attention_bwd: head dim %d (need a multiple of 16, <= 512)", d);
  AttnBwdArgs a;
  a.qkv_hi = (const uint16_t*)qkv; a.qkv_lo = (const uint16_t*)qkv_lo; a.o_hi = (const uint16_t*)out; a.o_lo = (const uint16_t*)out_lo;
  a.do_hi = (const uint16_t*)dout; a.do_lo = (const uint16_t*)dout_lo; a.dqkv_hi = (uint16_t*)dqkv; a.dqkv_lo = (uint16_t*)dqkv_lo;
  a.stats = stats; a.T = T; a.C = C; a.heads = heads; a.d = d; a.f16 = f16 ? 1 : 0;
  a.scale = 1.0f / sqrtf((float)d);
  const int smem = (4 * kAbR * (d + 4) + 2 * kAbR * 20) * (int)sizeof(float);
  static unsigned long long c0 = 0, c1 = 0, c2 = 0;
  cudaError_t e = smem_attr_once(attn_bwd_kernel<0>, 200 * 1024, c0);
  if (e == cudaSuccess) e = smem_attr_once(attn_bwd_kernel<1>, 200 * 1024, c1);
  if (e == cudaSuccess) e = smem_attr_once(attn_bwd_kernel<2>, 200 * 1024, c2);
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "b2d_attention_bwd: smem attr: %s", cudaGetErrorString(e));
  const dim3 grid((T + kAbR - 1) / kAbR, heads, N);
  attn_bwd_kernel<0><<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  attn_bwd_kernel<1><<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  attn_bwd_kernel<2><<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  return check_launch("attn_bwd_kernel");
}

extern "C" int b2d_pack_weight(const float* src, int32_t R1, int32_t R2, int64_t sr1, int64_t sr2, int32_t ntaps, int64_t st, int32_t cs, int64_t sc,
                               void* dst_hi, void* dst_lo, int64_t ktot, int32_t cpad, int32_t f16, void* stream) {
  if (!src || !dst_hi || R1 < 1 || R2 < 1 || ntaps < 1 || cs < 1 || cs > cpad || ktot < (int64_t)ntaps * cpad || (f16 && dst_lo))
    return set_error(B2D_E_INVALID, "b2d_pack_weight: bad argument");
  PackArgs a;
  a.src = src; a.hi = (uint16_t*)dst_hi; a.lo = (uint16_t*)dst_lo;
  a.sr1 = sr1; a.sr2 = sr2; a.st = st; a.sc = sc; a.ktot = ktot;
  a.R2 = R2; a.ntaps = ntaps; a.cs = cs; a.cpad = cpad; a.f16 = f16 ? 1 : 0;
  a.total = (long long)R1 * R2 * ntaps * cs;
  if (sr2 == 1 && sc != 1 && (long long)R1 * ntaps <= 65535)
    pack_weight_t_kernel<<<dim3((R2 + 31) / 32, (cs + 31) / 32, R1 * ntaps), 256, 0, (cudaStream_t)stream>>>(a);
  else
    pack_weight_kernel<<<grid_cap(a.total / ntaps, 256), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("pack_weight_kernel");
}
