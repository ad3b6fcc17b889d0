// Memory-bound kernels around the conv engine: GroupNorm apply (+SiLU, +time-embedding add),
// 2x2 max-pool with GroupNorm partial sums, nearest upsample, layout conversion at the module
// boundary.  All are 16-byte vectorised over the channels-last innermost dimension and sized
// as a multiple of the SM count.  Reference call sites are listed in include/b2d.h.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/b2d.h"
#include "b2d_internal.h"
#include "b2d_ptx.cuh"

namespace b2d {

__device__ __forceinline__ float bflo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bfhi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t packbf(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bflo(u.x); f[1] = bfhi(u.x); f[2] = bflo(u.y); f[3] = bfhi(u.y);
  f[4] = bflo(u.z); f[5] = bfhi(u.z); f[6] = bflo(u.w); f[7] = bfhi(u.w);
}
__device__ __forceinline__ void unpack8_f16(const uint4& u, float (&f)[8]) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.z)), d = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(packbf(f[0], f[1]), packbf(f[2], f[3]), packbf(f[4], f[5]), packbf(f[6], f[7]));
}
__device__ __forceinline__ uint32_t packh_sat(float a, float b) {
  uint32_t r;
  asm("{\n\t.reg .b16 lo, hi;\n\tcvt.rn.satfinite.f16.f32 lo, %1;\n\tcvt.rn.satfinite.f16.f32 hi, %2;\n\tmov.b32 %0, {lo, hi};\n\t}" : "=r"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint4 pack8_f16(const float (&f)[8]) {
  return make_uint4(packh_sat(f[0], f[1]), packh_sat(f[2], f[3]), packh_sat(f[4], f[5]), packh_sat(f[6], f[7]));
}
// 16-bit storage format of an activation tensor: IEEE fp16 (f16 != 0) or bf16 -- warp-uniform runtime flag
__device__ __forceinline__ void unpack8_fmt(const uint4& u, float (&f)[8], int f16) {
  if (f16) unpack8_f16(u, f); else unpack8(u, f);
}
__device__ __forceinline__ uint4 pack8_fmt(const float (&f)[8], int f16) { return f16 ? pack8_f16(f) : pack8(f); }
// bf16 "lo" part of an fp32 value whose "hi" part is the packed word (hi/lo split, fp32x mode)
__device__ __forceinline__ uint4 pack8_lo(const float (&f)[8], const uint4& hi) {
  float h[8];
  unpack8(hi, h);
  float l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) l[i] = f[i] - h[i];
  return pack8(l);
}
// SiLU good to ~1e-6 relative (ex2.approx + rcp.approx): used where the result is stored as fp16 (2^-11 rounding), for
// which tanh.approx's ~2^-11 error would no longer be negligible
__device__ __forceinline__ float silu_acc(float x) { return __fdividef(x, 1.f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// --------------------------------------------------------------------------- GroupNorm apply
// grid = (blocks_per_image, N); block = 256.  Per-channel scale/shift/temb staged in smem.
template <bool TEMB, bool ACC>
__global__ void __launch_bounds__(256, TEMB ? 3 : 4) gn_apply_kernel(
    const uint4* __restrict__ x, const uint4* __restrict__ x_lo, uint4* __restrict__ y, uint4* __restrict__ y_lo,
    long long P, int C, const double* __restrict__ stats, int cpg, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, int act, const float* __restrict__ temb_table,
    const int* __restrict__ temb_row, int temb_row_stride, int temb_ld, int temb_col, double* __restrict__ stats_out,
    int in_f16, int out_f16) {
  extern __shared__ float sm[];
  float* sa = sm;
  float* sb = sm + C;
  float* st = sm + 2 * C;
  const int n = blockIdx.y;
  const int G = C / cpg;
  const double cnt = (double)cpg * (double)P;
  const float* trow = nullptr;
  if (temb_table != nullptr) {
    const int row = temb_row ? temb_row[(long long)n * temb_row_stride] : 0;
    trow = temb_table + (long long)row * temb_ld + temb_col;
  }
  // per-group mean / rstd once (fp64: E[x^2] - mean^2 cancels), then fp32 per-channel scale / shift
  __shared__ float s_mean[64], s_rstd[64];
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const double s = stats[((long long)n * G + g) * 2];
    const double ss = stats[((long long)n * G + g) * 2 + 1];
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0) var = 0;
    s_mean[g] = (float)mean;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float rstd = s_rstd[g], mean = s_mean[g];
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    sa[c] = rstd * ga;
    sb[c] = be - mean * rstd * ga;
    st[c] = trow ? trow[c] : 0.f;
  }
  __syncthreads();

  const int vpc = C >> 3;  // 16-byte vectors per pixel
  const long long nvec = P * vpc;
  const long long base = (long long)n * nvec;
  float acc_s = 0.f, acc_ss = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // silu(v) = h + h * tanh(h), h = v / 2 (one MUFU op per element; the 1/2 is folded into the coefficients by the
  // fast path).  tanh.approx.f32 is good to ~2^-11, below the 2^-9 rounding of the bf16 result.
  auto transform = [&](float (&f)[8], const float* a, const float* b, const float* t, bool half_folded) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = fmaf(f[j], a[j], b[j]);
      if (act) {
        if constexpr (ACC) {
          v = silu_acc(v);
        } else {
          const float h = half_folded ? v : 0.5f * v;
          float th;
          asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
          v = fmaf(h, th, h);
        }
      }
      if (TEMB) v += t[j];
      f[j] = v;
      acc_s += v;
      acc_ss += v * v;
    }
  };
  if ((blockDim.x % vpc) == 0 && x_lo == nullptr && y_lo == nullptr) {
    // fast path: the grid stride is a multiple of the vectors-per-pixel count, so a thread always sees the
    // same 8 channels -> coefficients live in registers and 4 independent 16-byte loads are in flight.
    const int c0 = (threadIdx.x % vpc) << 3;
    float a[8], b[8], t[8];
    const float fold = (act && !ACC) ? 0.5f : 1.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = fold * sa[c0 + j]; b[j] = fold * sb[c0 + j]; t[j] = st[c0 + j]; }
    const uint4* xp = x + base;
    uint4* yp = y + base;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldcs(xp + i + k * stride);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8];
        unpack8_fmt(u[k], f, in_f16);
        transform(f, a, b, t, true);
        yp[i + k * stride] = pack8_fmt(f, out_f16);
      }
    }
    for (; i < nvec; i += stride) {
      float f[8];
      unpack8_fmt(__ldcs(xp + i), f, in_f16);
      transform(f, a, b, t, true);
      yp[i] = pack8_fmt(f, out_f16);
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
      const int c0 = (int)(i % vpc) << 3;
      float f[8];
      unpack8_fmt(__ldg(x + base + i), f, in_f16);
      if (x_lo != nullptr) {
        float l[8];
        unpack8(__ldg(x_lo + base + i), l);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += l[j];
      }
      transform(f, sa + c0, sb + c0, st + c0, false);
      const uint4 hi = pack8_fmt(f, out_f16);
      y[base + i] = hi;
      if (y_lo != nullptr) y_lo[base + i] = pack8_lo(f, hi);  // hi/lo split: bf16 only (checked by the launcher)
    }
  }
  if (stats_out != nullptr) {
    __shared__ float red[2][8];
    acc_s = warp_sum(acc_s);
    acc_ss = warp_sum(acc_ss);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = acc_s; red[1][w] = acc_ss; }
    __syncthreads();
    if (w == 0) {
      float a = l < (int)(blockDim.x >> 5) ? red[0][l] : 0.f;
      float b = l < (int)(blockDim.x >> 5) ? red[1][l] : 0.f;
      a = warp_sum(a);
      b = warp_sum(b);
      if (l == 0) {
        atomicAdd(stats_out + (long long)n * 2, (double)a);
        atomicAdd(stats_out + (long long)n * 2 + 1, (double)b);
      }
    }
  }
}

// --------------------------------------------------------------------------- max-pool 2x2 + sums
__global__ void __launch_bounds__(256) maxpool_stats_kernel(const uint4* __restrict__ x, const uint4* __restrict__ x_lo,
                                                            uint4* __restrict__ y, uint4* __restrict__ y_lo, int H, int W,
                                                            int C, double* __restrict__ stats, int f16) {
  const int n = blockIdx.y;
  const int OH = H >> 1, OW = W >> 1, vpc = C >> 3;
  const long long nvec = (long long)OH * OW * vpc;
  const uint4* xi = x + (long long)n * H * W * vpc;
  const uint4* xl = x_lo ? x_lo + (long long)n * H * W * vpc : nullptr;
  float acc_s = 0.f, acc_ss = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpc);
    const long long pix = i / vpc;
    const int ox = (int)(pix % OW), oy = (int)(pix / OW);
    float m[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long src = ((long long)(2 * oy + (q >> 1)) * W + (2 * ox + (q & 1))) * vpc + v;
      float f[8];
      unpack8_fmt(__ldg(xi + src), f, f16);
      if (xl) {
        float l[8];
        unpack8(__ldg(xl + src), l);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += l[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = q == 0 ? f[j] : fmaxf(m[j], f[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc_s += m[j]; acc_ss += m[j] * m[j]; }
    const uint4 hi = pack8_fmt(m, f16);
    y[(long long)n * nvec + i] = hi;
    if (y_lo) y_lo[(long long)n * nvec + i] = pack8_lo(m, hi);
  }
  __shared__ float red[2][8];
  acc_s = warp_sum(acc_s);
  acc_ss = warp_sum(acc_ss);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = acc_s; red[1][w] = acc_ss; }
  __syncthreads();
  if (w == 0) {
    float a = l < (int)(blockDim.x >> 5) ? red[0][l] : 0.f;
    float b = l < (int)(blockDim.x >> 5) ? red[1][l] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
    if (l == 0 && stats != nullptr) {
      atomicAdd(stats + (long long)n * 2, (double)a);
      atomicAdd(stats + (long long)n * 2 + 1, (double)b);
    }
  }
}

// --------------------------------------------------------------------------- per-sample fused GroupNorm(1, C) ops
// One 1024-thread CTA per sample keeps the sample (<= 65536 elements) in registers, so a normalisation whose statistics
// depend on values produced in the same kernel needs no second launch:
//   gn_gn_kernel      y1 = silu(GN(x; stats1)),  y2 = GN(y1)        (DoubleBlock's last norm + the attention pre-norm,
//                                                                    unet/blocks.py:37-47 and 192,214)
//   maxpool_gn_kernel y  = silu(GN(maxpool2x2(x)))                  (Down, unet/blocks.py:161-174)
__device__ __forceinline__ float silu_tanh(float v) {
  const float h = 0.5f * v;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
  return fmaf(h, th, h);
}
// fp16 results get the accurate form (see silu_acc); these per-sample kernels are latency-bound, not MUFU-bound
__device__ __forceinline__ float silu_sel(float v, int f16) { return f16 ? silu_acc(v) : silu_tanh(v); }

// sum of (s, ss) over the CTA, fp64, result broadcast to every thread
__device__ __forceinline__ void block_sum2(double& s, double& ss, double (*red)[2]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[w][0] = s; red[w][1] = ss; }
  __syncthreads();
  s = 0.0; ss = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s += red[i][0]; ss += red[i][1]; }  // same order in every thread
}

// KEEP: the sample fits the register file (<= 32 elements per thread); otherwise the second pass re-reads what the thread
// itself wrote (y1) / pooled (x) -- L2 hits, no cross-thread dependence.
template <int VPT, bool KEEP>
__global__ void __launch_bounds__(1024) gn_gn_kernel(const uint4* __restrict__ x, int in_f16, uint4* __restrict__ y1,
                                                     uint4* __restrict__ y2, int nvec, int C, const double* __restrict__ stats1,
                                                     const float* __restrict__ g1, const float* __restrict__ b1, float eps1, int act1,
                                                     const float* __restrict__ g2, const float* __restrict__ b2, float eps2, int act2,
                                                     int out_f16) {
  __shared__ double red[32][2];
  const int n = blockIdx.x;
  const int vpc = C >> 3;
  const int c0 = (threadIdx.x % vpc) << 3;  // blockDim.x is a multiple of vpc: a thread sees the same 8 channels
  const double cnt = (double)nvec * 8.0;
  const double m1 = stats1[2 * n] / cnt;
  double var1 = stats1[2 * n + 1] / cnt - m1 * m1;
  if (var1 < 0) var1 = 0;
  const float mean1 = (float)m1, rstd1 = (float)(1.0 / sqrt(var1 + (double)eps1));
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float ga = g1 ? __ldg(g1 + c0 + j) : 1.f, be = b1 ? __ldg(b1 + c0 + j) : 0.f;
    a[j] = rstd1 * ga;
    b[j] = be - mean1 * rstd1 * ga;
  }
  const uint4* xp = x + (long long)n * nvec;
  uint4* y1p = y1 + (long long)n * nvec;
  constexpr int KV = KEEP ? VPT : 1;
  float f[KV][8];
  float ps = 0.f, pss = 0.f;
#pragma unroll
  for (int k0 = 0; k0 < VPT; k0 += 4) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4 && k0 + k < VPT; ++k) {
      const int v = threadIdx.x + (k0 + k) * 1024;
      u[k] = v < nvec ? __ldcs(xp + v) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < 4 && k0 + k < VPT; ++k) {
      const int v = threadIdx.x + (k0 + k) * 1024;
      float (&fk)[8] = f[KEEP ? k0 + k : 0];
      unpack8_fmt(u[k], fk, in_f16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(fk[j], a[j], b[j]);
        if (act1) t = silu_sel(t, out_f16);
        fk[j] = t;
        if (v < nvec) { ps += t; pss = fmaf(t, t, pss); }
      }
      if (v < nvec) y1p[v] = pack8_fmt(fk, out_f16);
    }
  }
  double s = (double)ps, ss = (double)pss;
  block_sum2(s, ss, red);
  const double m2 = s / cnt;
  double var2 = ss / cnt - m2 * m2;
  if (var2 < 0) var2 = 0;
  const float mean2 = (float)m2, rstd2 = (float)(1.0 / sqrt(var2 + (double)eps2));
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float ga = g2 ? __ldg(g2 + c0 + j) : 1.f, be = b2 ? __ldg(b2 + c0 + j) : 0.f;
    a[j] = rstd2 * ga;
    b[j] = be - mean2 * rstd2 * ga;
  }
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = threadIdx.x + k * 1024;
    float (&fk)[8] = f[KEEP ? k : 0];
    if (!KEEP) unpack8_fmt(v < nvec ? y1p[v] : make_uint4(0u, 0u, 0u, 0u), fk, out_f16);  // this thread's own store, rounded
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(fk[j], a[j], b[j]);
      if (act2) t = silu_sel(t, out_f16);
      fk[j] = t;
    }
    if (v < nvec) y2[(long long)n * nvec + v] = pack8_fmt(fk, out_f16);
  }
}

template <int VPT, bool KEEP>
__global__ void __launch_bounds__(1024) maxpool_gn_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int C,
                                                          const float* __restrict__ g, const float* __restrict__ be, float eps,
                                                          int act, int f16) {
  __shared__ double red[32][2];
  const int n = blockIdx.x;
  const int OW = W >> 1, vpc = C >> 3;
  const int nvec = (H >> 1) * OW * vpc;
  const int cv = threadIdx.x % vpc;
  const uint4* xi = x + (long long)n * H * W * vpc;
  constexpr int KV = KEEP ? VPT : 1;
  float f[KV][8];
  float ps = 0.f, pss = 0.f;
  auto pool = [&](int v, float (&m)[8]) {
    const int pix = v < nvec ? v / vpc : 0;
    const int ox = pix % OW, oy = pix / OW;
    uint4 q[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) q[t] = __ldg(xi + ((long long)(2 * oy + (t >> 1)) * W + (2 * ox + (t & 1))) * vpc + cv);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float e[8];
      unpack8_fmt(q[t], e, f16);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = t == 0 ? e[j] : fmaxf(m[j], e[j]);
    }
  };
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = threadIdx.x + k * 1024;
    float (&fk)[8] = f[KEEP ? k : 0];
    pool(v, fk);
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { ps += fk[j]; pss = fmaf(fk[j], fk[j], pss); }
    }
  }
  double s = (double)ps, ss = (double)pss;
  block_sum2(s, ss, red);
  const double cnt = (double)nvec * 8.0;
  const double m = s / cnt;
  double var = ss / cnt - m * m;
  if (var < 0) var = 0;
  const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)eps));
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float ga = g ? __ldg(g + cv * 8 + j) : 1.f, bb = be ? __ldg(be + cv * 8 + j) : 0.f;
    a[j] = rstd * ga;
    b[j] = bb - mean * rstd * ga;
  }
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = threadIdx.x + k * 1024;
    float (&fk)[8] = f[KEEP ? k : 0];
    if (!KEEP) pool(v, fk);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(fk[j], a[j], b[j]);
      if (act) t = silu_sel(t, f16);
      fk[j] = t;
    }
    if (v < nvec) y[(long long)n * nvec + v] = pack8_fmt(fk, f16);
  }
}

// --------------------------------------------------------------------------- nearest 2x (in-plane)
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long ND,
                                                         int H, int W, int vpc) {
  const long long total = ND * (2LL * H) * (2LL * W) * vpc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpc);
    long long pix = i / vpc;
    const int ox = (int)(pix % (2 * W)); pix /= (2 * W);
    const int oy = (int)(pix % (2 * H));
    const long long nd = pix / (2 * H);
    y[i] = __ldg(x + ((nd * H + (oy >> 1)) * W + (ox >> 1)) * vpc + v);
  }
}

// --------------------------------------------------------------------------- layout conversion
// planar fp32 [N][C][P] -> channels-last 16-bit (bf16, or fp16 if f16) [N][P][cpad] at channel offset coff
__global__ void __launch_bounds__(256) planar_to_cl_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                           __nv_bfloat16* __restrict__ y_lo, int N, int C, long long P,
                                                           int cpad, int coff, const float* __restrict__ div_scale, int f16) {
  const long long total = (long long)N * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / P, p = i - n * P;
    for (int c = 0; c < C; ++c) {
      float v = __ldg(x + (n * C + c) * P + p);
      if (div_scale) v = __fdiv_rn(v, __ldg(div_scale + c));
      if (f16) {
        reinterpret_cast<uint16_t*>(y)[i * cpad + coff + c] = (uint16_t)(packh_sat(v, 0.f) & 0xFFFFu);
        continue;
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      y[i * cpad + coff + c] = h;
      if (y_lo) y_lo[i * cpad + coff + c] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}
// channels-last 16-bit [N][P][cstride] (channels coff..coff+C) -> planar fp32 [N][C][P]
__global__ void __launch_bounds__(256) cl_to_planar_kernel(const __nv_bfloat16* __restrict__ x,
                                                           const __nv_bfloat16* __restrict__ x_lo, float* __restrict__ y,
                                                           int N, int C, long long P, int cstride, int coff, int f16) {
  const long long total = (long long)N * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / P, p = i - n * P;
    for (int c = 0; c < C; ++c) {
      float v = f16 ? __half2float(reinterpret_cast<const __half*>(x)[i * cstride + coff + c]) : __bfloat162float(x[i * cstride + coff + c]);
      if (x_lo) v += __bfloat162float(x_lo[i * cstride + coff + c]);
      y[(n * C + c) * P + p] = v;
    }
  }
}

// z-folded small-Cout Conv3d (vae/decoder.py:71): the 3x3x3 conv runs as a 3x3 conv per z slice whose output rows are
// (kz, co) -- P[n][zi][y][x][kz*4 + co] = sum_{ky,kx,c} w[co][c][kz][ky][kx] * x[n][zi][y+ky-1][x+kx-1][c] -- and this
// pass gathers the three z contributions: out[n][z][co] = bias[co] + sum_kz P[n][z+kz-1][..][kz*4+co] (zero padding in
// z), times the per-channel scale and the mask, written planar fp32 [N*D][cstride][H][W] at channel offset coff.
__global__ void zfold_combine_kernel(const float4* __restrict__ P, int D, long long plane, long long total, int co,
                                     const float* __restrict__ bias, const float* __restrict__ scale,
                                     const float* __restrict__ mask, float* __restrict__ out, int out_cstride, int out_coff) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / plane;
    const long long pix = i - img * plane;
    const int z = (int)(img % D);
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int kz = 0; kz < 3; ++kz) {
      const int zi = z + kz - 1;
      if (zi < 0 || zi >= D) continue;
      const float4 v = __ldcs(P + ((img + (kz - 1)) * plane + pix) * 3 + kz);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z;
    }
    const float m = mask != nullptr ? __ldg(mask + i) : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < co) {
        const float b = bias != nullptr ? __ldg(bias + c) : 0.f;
        const float sc = scale != nullptr ? __ldg(scale + c) : 1.f;
        __stcs(out + (img * out_cstride + out_coff + c) * plane + pix, (acc[c] + b) * sc * m);
      }
    }
  }
}

// z-stacked conv_in input (vae/encoder.py:30, decoder.py:31: Conv3d with C <= 21 input channels): the three z taps become
// channels -- y[b][z][p][kz*C + c] = x[b][z+kz-1][p][c] (zero outside the sample) -- so the conv runs 9 in-plane taps on one
// 64-channel chunk instead of 27 taps on a chunk that is mostly padding.  bf16 channels-last, pad channels of y untouched.
template <int C>
__global__ void __launch_bounds__(256) zstack_cl_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int D, long long P,
                                                        long long total, int cvec) {
  constexpr int LV = (C + 7) / 8, OV = (3 * C + 7) / 8;  // 16-byte vectors read per source row / written per output row
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int z = (int)((i / P) % D);
    union { uint4 v[OV]; uint16_t h[OV * 8]; } o;
#pragma unroll
    for (int j = 0; j < OV; ++j) o.v[j] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int kz = 0; kz < 3; ++kz) {
      const int zi = z + kz - 1;
      if (zi < 0 || zi >= D) continue;
      union { uint4 v[LV]; uint16_t h[LV * 8]; } in;
#pragma unroll
      for (int j = 0; j < LV; ++j) in.v[j] = __ldg(x + (i + (long long)(kz - 1) * P) * cvec + j);
#pragma unroll
      for (int c = 0; c < C; ++c) o.h[kz * C + c] = in.h[c];
    }
#pragma unroll
    for (int j = 0; j < OV; ++j) y[i * cvec + j] = o.v[j];
  }
}

static inline int grid_for(long long work_items, int per_block, int max_waves = 8) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace b2d

using namespace b2d;

extern "C" int b2d_gn_apply(const void* x, const void* x_lo, void* y, void* y_lo, int32_t N, int64_t P, int32_t C,
                            const double* stats, int32_t cpg, const float* gamma, const float* beta, float eps, int32_t act,
                            const float* temb_table, const int32_t* temb_row, int32_t temb_row_stride, int32_t temb_ld,
                            int32_t temb_col, double* stats_out, int32_t in_f16, int32_t out_f16, void* stream) {
  if (!x || !y || !stats) return set_error(B2D_E_INVALID, "b2d_gn_apply: null pointer");
  if ((in_f16 && x_lo) || (out_f16 && y_lo)) return set_error(B2D_E_INVALID, "b2d_gn_apply: an fp16 tensor has no lo part");
  if (N < 1 || P < 1 || C < 8 || (C % 8) || cpg < 1 || (C % cpg) || C / cpg > 64) return set_error(B2D_E_INVALID, "b2d_gn_apply: bad shape N=%d P=%lld C=%d cpg=%d", N, (long long)P, C, cpg);
  if (N > 65535) return set_error(B2D_E_INVALID, "b2d_gn_apply: N too large");
  const long long nvec = (long long)P * (C / 8);
  int bx = grid_for(nvec, 256 * 4);
  int per_img_cap = (num_sms() * 8 + N - 1) / N;
  if (bx > per_img_cap) bx = per_img_cap < 1 ? 1 : per_img_cap;
  const size_t smem = 3 * (size_t)C * sizeof(float);
  if (smem > 96 * 1024) return set_error(B2D_E_INVALID, "b2d_gn_apply: C=%d too large (3*C floats of shared memory, max 96 KB)", C);
  // four variants (time-embedding add x accurate SiLU for fp16 results): per-(kernel, device) opt-in to the carve-out
  static unsigned long long cfg[4] = {0, 0, 0, 0};
  const int te = temb_table != nullptr ? 1 : 0, ac = out_f16 ? 1 : 0;
  cudaError_t ea = te ? (ac ? smem_attr_once(gn_apply_kernel<true, true>, 96 * 1024, cfg[3]) : smem_attr_once(gn_apply_kernel<true, false>, 96 * 1024, cfg[2]))
                      : (ac ? smem_attr_once(gn_apply_kernel<false, true>, 96 * 1024, cfg[1]) : smem_attr_once(gn_apply_kernel<false, false>, 96 * 1024, cfg[0]));
  if (ea != cudaSuccess) return set_error(B2D_E_CUDA, "b2d_gn_apply: cudaFuncSetAttribute: %s", cudaGetErrorString(ea));
#define B2D_GNA(T, A)                                                                                                          \
  gn_apply_kernel<T, A><<<dim3(bx, N), dim3(256), smem, (cudaStream_t)stream>>>(                                               \
      (const uint4*)x, (const uint4*)x_lo, (uint4*)y, (uint4*)y_lo, P, C, stats, cpg, gamma, beta, eps, act, temb_table, temb_row, \
      temb_row_stride, temb_ld, temb_col, stats_out, in_f16 ? 1 : 0, out_f16 ? 1 : 0)
  if (te) { if (ac) B2D_GNA(true, true); else B2D_GNA(true, false); }
  else { if (ac) B2D_GNA(false, true); else B2D_GNA(false, false); }
#undef B2D_GNA
  return check_launch("gn_apply_kernel");
}

extern "C" int b2d_maxpool2x2_stats(const void* x, const void* x_lo, void* y, void* y_lo, int32_t N, int32_t H, int32_t W,
                                    int32_t C, double* stats, int32_t f16, void* stream) {
  if (!x || !y) return set_error(B2D_E_INVALID, "b2d_maxpool2x2_stats: null pointer");
  if (f16 && (x_lo || y_lo)) return set_error(B2D_E_INVALID, "b2d_maxpool2x2_stats: an fp16 tensor has no lo part");
  if (N < 1 || N > 65535 || H < 2 || W < 2 || (H & 1) || (W & 1) || C < 8 || (C % 8))
    return set_error(B2D_E_INVALID, "b2d_maxpool2x2_stats: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  const long long nvec = (long long)(H / 2) * (W / 2) * (C / 8);
  int bx = grid_for(nvec, 256);
  int per_img_cap = (num_sms() * 8 + N - 1) / N;
  if (bx > per_img_cap) bx = per_img_cap < 1 ? 1 : per_img_cap;
  maxpool_stats_kernel<<<dim3(bx, N), dim3(256), 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)x_lo, (uint4*)y, (uint4*)y_lo, H, W, C,
                                                                            stats, f16 ? 1 : 0);
  return check_launch("maxpool_stats_kernel");
}

extern "C" int b2d_zstack_cl(const void* x, void* y, int32_t ND, int32_t D, int64_t P, int32_t C, int32_t cpad, void* stream) {
  if (!x || !y || x == y || ND < 1 || D < 1 || (ND % D) || P < 1 || 3 * C > cpad || (cpad % 8) ||
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15))
    return set_error(B2D_E_INVALID, "b2d_zstack_cl: bad argument");
  const long long total = (long long)ND * P;
  const int g = grid_for(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 3) zstack_cl_kernel<3><<<g, 256, 0, st>>>((const uint4*)x, (uint4*)y, D, P, total, cpad / 8);
  else if (C == 8) zstack_cl_kernel<8><<<g, 256, 0, st>>>((const uint4*)x, (uint4*)y, D, P, total, cpad / 8);
  else return set_error(B2D_E_UNSUPPORTED, "b2d_zstack_cl: C=%d (3 or 8: the VAE branches' input channels)", C);
  return check_launch("zstack_cl_kernel");
}

extern "C" int b2d_zfold_combine(const float* P, int32_t ND, int32_t D, int32_t H, int32_t W, int32_t co, const float* bias,
                                 const float* scale, const float* mask, float* out, int32_t out_cstride, int32_t out_coff,
                                 void* stream) {
  if (!P || !out || ND < 1 || D < 1 || (ND % D) || H < 1 || W < 1 || co < 1 || co > 3 || out_coff < 0 || out_coff + co > out_cstride ||
      (reinterpret_cast<uintptr_t>(P) & 15))
    return set_error(B2D_E_INVALID, "b2d_zfold_combine: bad argument");
  const long long plane = (long long)H * W, total = (long long)ND * plane;
  zfold_combine_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)P, D, plane, total, co, bias, scale,
                                                                               mask, out, out_cstride, out_coff);
  return check_launch("zfold_combine_kernel");
}

// per-sample fused ops: one CTA per sample, the sample (<= 65536 elements, one GroupNorm group) lives in registers
static int fused_vpt(long long nvec) { return nvec <= 1024 ? 1 : nvec <= 2048 ? 2 : nvec <= 4096 ? 4 : nvec <= 8192 ? 8 : 0; }

extern "C" int b2d_gn_gn_apply(const void* x, int32_t in_f16, void* y1, void* y2, int32_t N, int64_t P, int32_t C, const double* stats1,
                               const float* gamma1, const float* beta1, float eps1, int32_t act1, const float* gamma2,
                               const float* beta2, float eps2, int32_t act2, int32_t out_f16, void* stream) {
  if (!x || !y1 || !y2 || !stats1 || N < 1 || P < 1 || C < 8 || (C % 8)) return set_error(B2D_E_INVALID, "b2d_gn_gn_apply: bad argument");
  const long long nvec = P * (C / 8);
  const int vpt = fused_vpt(nvec);
  if (vpt == 0 || (1024 % (C / 8)) != 0)
    return set_error(B2D_E_UNSUPPORTED, "b2d_gn_gn_apply: sample of %lld elements / C=%d (max 65536 elements, C/8 dividing 1024)", nvec * 8, C);
  cudaStream_t st = (cudaStream_t)stream;
#define B2D_GG(V, K) gn_gn_kernel<V, K><<<dim3(N), dim3(1024), 0, st>>>((const uint4*)x, (int)in_f16, (uint4*)y1, (uint4*)y2, (int)nvec, \
                             (int)C, stats1, gamma1, beta1, eps1, (int)act1, gamma2, beta2, eps2, (int)act2, out_f16 ? 1 : 0)
  if (vpt == 1) B2D_GG(1, true); else if (vpt == 2) B2D_GG(2, true); else if (vpt == 4) B2D_GG(4, true); else B2D_GG(8, false);
#undef B2D_GG
  return check_launch("gn_gn_kernel");
}

extern "C" int b2d_maxpool2x2_gn(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, const float* gamma,
                                 const float* beta, float eps, int32_t act, int32_t f16, void* stream) {
  if (!x || !y || N < 1 || H < 2 || W < 2 || (H & 1) || (W & 1) || C < 8 || (C % 8))
    return set_error(B2D_E_INVALID, "b2d_maxpool2x2_gn: bad argument");
  const long long nvec = (long long)(H / 2) * (W / 2) * (C / 8);
  const int vpt = fused_vpt(nvec);
  if (vpt == 0 || (1024 % (C / 8)) != 0)
    return set_error(B2D_E_UNSUPPORTED, "b2d_maxpool2x2_gn: pooled sample of %lld elements / C=%d (max 65536 elements, C/8 dividing 1024)",
                     nvec * 8, C);
  cudaStream_t st = (cudaStream_t)stream;
#define B2D_PG(V, K) maxpool_gn_kernel<V, K><<<dim3(N), dim3(1024), 0, st>>>((const uint4*)x, (uint4*)y, (int)H, (int)W, (int)C, gamma, beta, \
                             eps, (int)act, f16 ? 1 : 0)
  if (vpt == 1) B2D_PG(1, true); else if (vpt == 2) B2D_PG(2, true); else if (vpt == 4) B2D_PG(4, true); else B2D_PG(8, false);
#undef B2D_PG
  return check_launch("maxpool_gn_kernel");
}

extern "C" int b2d_upsample2x_nearest(const void* x, void* y, int32_t ND, int32_t H, int32_t W, int32_t C, void* stream) {
  if (!x || !y || ND < 1 || H < 1 || W < 1 || C < 8 || (C % 8)) return set_error(B2D_E_INVALID, "b2d_upsample2x_nearest: bad argument");
  const long long total = (long long)ND * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(total, 256 * 2), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, ND, H, W, C / 8);
  return check_launch("upsample2x_kernel");
}

extern "C" int b2d_planar_to_cl(const float* x, void* y, void* y_lo, int32_t N, int32_t C, int64_t P, int32_t cpad, int32_t coff,
                                const float* div_scale, int32_t f16, void* stream) {
  if (!x || !y || N < 1 || C < 1 || P < 1 || coff < 0 || coff + C > cpad || (f16 && y_lo)) return set_error(B2D_E_INVALID, "b2d_planar_to_cl: bad argument");
  planar_to_cl_kernel<<<grid_for((long long)N * P, 256), 256, 0, (cudaStream_t)stream>>>(
      x, (__nv_bfloat16*)y, (__nv_bfloat16*)y_lo, N, C, P, cpad, coff, div_scale, f16 ? 1 : 0);
  return check_launch("planar_to_cl_kernel");
}

extern "C" int b2d_cl_to_planar(const void* x, const void* x_lo, float* y, int32_t N, int32_t C, int64_t P, int32_t cstride,
                                int32_t coff, int32_t f16, void* stream) {
  if (!x || !y || N < 1 || C < 1 || P < 1 || coff < 0 || coff + C > cstride || (f16 && x_lo)) return set_error(B2D_E_INVALID, "b2d_cl_to_planar: bad argument");
  cl_to_planar_kernel<<<grid_for((long long)N * P, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)x_lo, y, N, C, P, cstride, coff, f16 ? 1 : 0);
  return check_launch("cl_to_planar_kernel");
}
