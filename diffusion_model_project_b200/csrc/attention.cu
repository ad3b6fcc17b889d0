// Fused small-sequence attention core for the UNet's deep levels (tokens T = h*w <= 256...1024,
// 2 heads, d_head 128/256/512): softmax(q k^T / sqrt(d)) v for one (image, head, query block) per
// CTA with the whole K (then the whole V) resident in shared memory -- single-pass softmax, no
// T x T matrix in HBM.  Replaces the scaled-dot-product part of nn.MultiheadAttention
// (Diffusion_model/src/unet/blocks.py:196-227; F.multi_head_attention_forward math: q scaled by
// d^-1/2, softmax over keys, no mask, no dropout in eval).
//
// Two kernels behind one entry point:
//   * attention_tc_kernel (bf16 mode): S = Q K^T and O = P V on tcgen05 tensor cores.  One CTA per
//     (image, head, 128-query tile); Q, K, V tiles arrive by TMA into 128B-swizzled smem; S (128 x T
//     fp32) lives in TMEM, each thread owns one query row (tcgen05.ld 32x32b) so the softmax is
//     thread-local; P is written back to smem as the K-major A operand and V is consumed in place as
//     an MN-major B operand (no transpose pass); O (128 x d fp32) accumulates in TMEM.
//   * attention_kernel (fp32x hi/lo-split mode): fp32 CUDA-core math, the precision path of the
//     parity configuration.
// The projections around the core (in_proj, out_proj o proj_out) run on the tcgen05 GEMM engine.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/b2d.h"
#include "b2d_internal.h"
#include "b2d_ptx.cuh"
#include "attention_tc.cuh"

namespace b2d {

constexpr int kAttnThreads = 512;                        // at most; the launch uses 128, 256 or 512 (what shared memory allows)
constexpr int kQPW = 4;                                  // queries processed together by one warp

__device__ __forceinline__ float a_bflo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float a_bfhi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ void a_unpack8(const uint4& u, float (&f)[8]) {
  f[0] = a_bflo(u.x); f[1] = a_bfhi(u.x); f[2] = a_bflo(u.y); f[3] = a_bfhi(u.y);
  f[4] = a_bflo(u.z); f[5] = a_bfhi(u.z); f[6] = a_bflo(u.w); f[7] = a_bfhi(u.w);
}

// 16-bit pair / 8-vector -> fp32 in the tensor's storage format (f16: IEEE fp16, else bf16)
__device__ __forceinline__ void a_unpack2(uint32_t u, int f16, float& lo, float& hi) {
  if (f16) { const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&u)); lo = v.x; hi = v.y; }
  else { lo = a_bflo(u); hi = a_bfhi(u); }
}
__device__ __forceinline__ void a_unpack8f(const uint4& u, float (&f)[8], int f16) {
  a_unpack2(u.x, f16, f[0], f[1]); a_unpack2(u.y, f16, f[2], f[3]); a_unpack2(u.z, f16, f[4], f[5]); a_unpack2(u.w, f16, f[6], f[7]);
}

// stage rows [T][d] of one head (column offset col0 in the [N][T][3C] tensor) into smem, row pitch dp
__device__ __forceinline__ void stage_rows(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* dst, int T, int d, int dp,
                                           long long row_stride) {
  const int vpr = d >> 3;
  for (int i = threadIdx.x; i < T * vpr; i += blockDim.x) {
    const int j = i / vpr, v = i - j * vpr;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + (long long)j * row_stride) + v);
    *reinterpret_cast<uint4*>(dst + (long long)j * dp + v * 8) = u;
  }
}

template <bool SPLIT>
__global__ void __launch_bounds__(kAttnThreads) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                 const __nv_bfloat16* __restrict__ qkv_lo,
                                                                 __nv_bfloat16* __restrict__ out,
                                                                 __nv_bfloat16* __restrict__ out_lo, int T, int C, int heads,
                                                                 float scale, int f16) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int d = C / heads;
  const int dp = d + 8;  // +16 bytes per row: lanes reading different rows hit different banks
  const int nh = blockIdx.y;
  const int n = nh / heads, h = nh - n * heads;
  const int kQB = (int)(blockDim.x >> 5) * kQPW;   // queries per CTA: every CTA stages all of K, then all of V -- the more
  const int q0 = blockIdx.x * kQB;                 // queries share that, the less redundant staging
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  __nv_bfloat16* kv = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* kv_lo = kv + (size_t)T * dp;
  float* sP = reinterpret_cast<float*>(kv + (size_t)T * dp * (SPLIT ? 2 : 1));  // [kQB][T]
  float* sQ = sP + (size_t)kQB * T;                                             // [kQB][d]

  const long long rs = 3LL * C;
  const __nv_bfloat16* base = qkv + (long long)n * T * rs + (long long)h * d;
  const __nv_bfloat16* base_lo = SPLIT ? qkv_lo + (long long)n * T * rs + (long long)h * d : nullptr;

  // ---- phase 1: K resident, scores + softmax ------------------------------------------------
  stage_rows(base + C, kv, T, d, dp, rs);
  if (SPLIT) stage_rows(base_lo + C, kv_lo, T, d, dp, rs);
  for (int i = threadIdx.x; i < kQB * d; i += blockDim.x) {
    const int qi = i / d, c = i - qi * d;
    float v = 0.f;
    if (q0 + qi < T) {
      v = f16 ? __half2float(reinterpret_cast<const __half*>(base)[(long long)(q0 + qi) * rs + c]) : __bfloat162float(base[(long long)(q0 + qi) * rs + c]);
      if (SPLIT) v += __bfloat162float(base_lo[(long long)(q0 + qi) * rs + c]);
    }
    sQ[i] = v * scale;
  }
  __syncthreads();

  const int qw = warp * kQPW;  // first local query of this warp
  for (int j = lane; j < T; j += 32) {
    float acc[kQPW] = {0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat16* krow = kv + (size_t)j * dp;
    const __nv_bfloat16* krow_lo = kv_lo + (size_t)j * dp;
    for (int c = 0; c < d; c += 8) {
      float kf[8];
      a_unpack8f(*reinterpret_cast<const uint4*>(krow + c), kf, f16);
      if (SPLIT) {
        float kl[8];
        a_unpack8(*reinterpret_cast<const uint4*>(krow_lo + c), kl);
#pragma unroll
        for (int e = 0; e < 8; ++e) kf[e] += kl[e];
      }
#pragma unroll
      for (int qq = 0; qq < kQPW; ++qq) {
        const float4 qa = *reinterpret_cast<const float4*>(sQ + (size_t)(qw + qq) * d + c);
        const float4 qb = *reinterpret_cast<const float4*>(sQ + (size_t)(qw + qq) * d + c + 4);
        acc[qq] = fmaf(qa.x, kf[0], acc[qq]); acc[qq] = fmaf(qa.y, kf[1], acc[qq]);
        acc[qq] = fmaf(qa.z, kf[2], acc[qq]); acc[qq] = fmaf(qa.w, kf[3], acc[qq]);
        acc[qq] = fmaf(qb.x, kf[4], acc[qq]); acc[qq] = fmaf(qb.y, kf[5], acc[qq]);
        acc[qq] = fmaf(qb.z, kf[6], acc[qq]); acc[qq] = fmaf(qb.w, kf[7], acc[qq]);
      }
    }
#pragma unroll
    for (int qq = 0; qq < kQPW; ++qq) sP[(size_t)(qw + qq) * T + j] = acc[qq];
  }
  __syncwarp();
#pragma unroll
  for (int qq = 0; qq < kQPW; ++qq) {
    float* prow = sP + (size_t)(qw + qq) * T;
    float m = -INFINITY;
    for (int j = lane; j < T; j += 32) m = fmaxf(m, prow[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = __expf(prow[j] - m);
      prow[j] = e;
      s += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.f / s;
    for (int j = lane; j < T; j += 32) prow[j] *= inv;
  }
  __syncthreads();

  // ---- phase 2: V resident, out = P V ----------------------------------------------------------
  stage_rows(base + 2 * C, kv, T, d, dp, rs);
  if (SPLIT) stage_rows(base_lo + 2 * C, kv_lo, T, d, dp, rs);
  __syncthreads();

  constexpr int kMaxCI = 8;  // d <= 512: each lane owns channel pairs {2*lane + 64*i}
  const int nci = d >> 6;
  float o[kQPW][kMaxCI][2];
#pragma unroll
  for (int qq = 0; qq < kQPW; ++qq)
#pragma unroll
    for (int i = 0; i < kMaxCI; ++i) { o[qq][i][0] = 0.f; o[qq][i][1] = 0.f; }
  for (int j = 0; j < T; ++j) {
    float pj[kQPW];
#pragma unroll
    for (int qq = 0; qq < kQPW; ++qq) pj[qq] = sP[(size_t)(qw + qq) * T + j];
#pragma unroll
    for (int i = 0; i < kMaxCI; ++i) {
      if (i < nci) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(kv + (size_t)j * dp + 2 * lane + 64 * i);
        float v0, v1;
        a_unpack2(u, f16, v0, v1);
        if (SPLIT) {
          const uint32_t ul = *reinterpret_cast<const uint32_t*>(kv_lo + (size_t)j * dp + 2 * lane + 64 * i);
          v0 += a_bflo(ul); v1 += a_bfhi(ul);
        }
#pragma unroll
        for (int qq = 0; qq < kQPW; ++qq) {
          o[qq][i][0] = fmaf(pj[qq], v0, o[qq][i][0]);
          o[qq][i][1] = fmaf(pj[qq], v1, o[qq][i][1]);
        }
      }
    }
  }
#pragma unroll
  for (int qq = 0; qq < kQPW; ++qq) {
    const int q = q0 + qw + qq;
    if (q < T) {
#pragma unroll
      for (int i = 0; i < kMaxCI; ++i) {
        if (i < nci) {
          const long long idx = ((long long)n * T + q) * C + (long long)h * d + 2 * lane + 64 * i;
          if (f16) {
            *reinterpret_cast<__half2*>(out + idx) = __floats2half2_rn(o[qq][i][0], o[qq][i][1]);
            continue;
          }
          const __nv_bfloat162 hv = __floats2bfloat162_rn(o[qq][i][0], o[qq][i][1]);
          *reinterpret_cast<__nv_bfloat162*>(out + idx) = hv;
          if (SPLIT && out_lo != nullptr) {
            *reinterpret_cast<__nv_bfloat162*>(out_lo + idx) =
                __floats2bfloat162_rn(o[qq][i][0] - __low2float(hv), o[qq][i][1] - __high2float(hv));
          }
        }
      }
    }
  }
}


// ============================================================================================
// tensor-core path (bf16): device body in attention_tc.cuh
// ============================================================================================
__global__ void __launch_bounds__(kTcThreads, 1) attention_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw_attn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_attn) + 1023) & ~uintptr_t(1023));
  attention_tc_item(p, (int)blockIdx.x, (int)blockIdx.y, smem);
}

typedef CUresult (*PFN_encodeTiledA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

size_t attention_tc_smem(int T, int C, int heads) {
  const int dch = (C / heads) / 64;
  return (size_t)dch * kQChunkBytes + 2 * (size_t)dch * T * 128 + 64 + 1024;
}

int attention_tc_params(const void* qkv, void* out, int N, int T, int C, int heads, int f16, AttnTcParams* pp) {
  const int d = C / heads;
  const int dch = d / 64;
  const size_t p_bytes = (size_t)((T + 63) / 64) * kQChunkBytes;
  if ((d % 64) || d > 512 || T % 16 || T < 16 || T > 256 || (T > 128 && T % 128) || attention_tc_smem(T, C, heads) > 220 * 1024 ||
      p_bytes > (size_t)dch * kQChunkBytes + (size_t)dch * T * 128 || ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15))
    return set_error(B2D_E_UNSUPPORTED, "attention: T=%d d_head=%d not eligible for the tensor-core kernel", T, d);
  PFN_encodeTiledA enc = reinterpret_cast<PFN_encodeTiledA>(tensor_map_encode_fn());
  if (!enc) return set_error(B2D_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available (no CUDA driver / too old)");
  AttnTcParams& p = *pp;
  const int RB = T < 128 ? T : 128;
  cuuint64_t dims[3] = {(cuuint64_t)(3 * C), (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)(3 * C) * 2, (cuuint64_t)(3 * C) * 2 * (cuuint64_t)T};
  cuuint32_t box[3] = {64, (cuuint32_t)RB, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B2D_E_CUDA, "b2d_attention: cuTensorMapEncodeTiled failed: %d", (int)r);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.T = T; p.C = C; p.heads = heads; p.d = d; p.RB = RB; p.f16 = f16 ? 1 : 0;
  p.scale_log2e = 1.4426950408889634f / sqrtf((float)d);
  return B2D_OK;
}

static int launch_attention_tc(const void* qkv, void* out, int N, int T, int C, int heads, int f16, cudaStream_t st) {
  AttnTcParams p;
  const int rc = attention_tc_params(qkv, out, N, T, C, heads, f16, &p);
  if (rc != B2D_OK) return rc;
  const size_t smem = attention_tc_smem(T, C, heads);
  static unsigned long long configured = 0;
  {
    cudaError_t e = smem_attr_once(attention_tc_kernel, 227 * 1024, configured);
    if (e != cudaSuccess) return set_error(B2D_E_CUDA, "attention_tc smem attr: %s", cudaGetErrorString(e));
  }
  dim3 grid((T + 127) / 128, N * heads);
  attention_tc_kernel<<<grid, dim3(kTcThreads), smem, st>>>(p);
  return check_launch("attention_tc_kernel");
}

}  // namespace b2d

using namespace b2d;

extern "C" int b2d_attention(const void* qkv, const void* qkv_lo, void* out, void* out_lo, int32_t N, int32_t T, int32_t C,
                             int32_t heads, int32_t f16, void* stream) {
  if (!qkv || !out) return set_error(B2D_E_INVALID, "b2d_attention: null pointer");
  if (f16 && (qkv_lo || out_lo)) return set_error(B2D_E_INVALID, "b2d_attention: an fp16 tensor has no lo part");
  if (N < 1 || T < 1 || heads < 1 || C < 1 || (C % heads)) return set_error(B2D_E_INVALID, "b2d_attention: bad shape");
  const int d = C / heads;
  if ((d % 64) || d > 512) return set_error(B2D_E_INVALID, "b2d_attention: d_head=%d must be a multiple of 64 and <= 512", d);
  if ((long long)N * heads > 65535) return set_error(B2D_E_INVALID, "b2d_attention: N*heads too large");
  const bool split = qkv_lo != nullptr;
  if (!split) {
    // tensor-core path: T keys as one UMMA N extent, P tile aliasing the Q/K staging area
    const int dch = d / 64;
    const size_t need = (size_t)dch * kQChunkBytes + 2 * (size_t)dch * T * 128 + 64 + 1024;
    const size_t p_bytes = (size_t)((T + 63) / 64) * kQChunkBytes;
    if (T % 16 == 0 && T >= 16 && T <= 256 && (T <= 128 || T % 128 == 0) && need <= 227 * 1024 &&
        p_bytes <= (size_t)dch * kQChunkBytes + (size_t)dch * T * 128 && ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
      return launch_attention_tc(qkv, out, N, T, C, heads, f16, (cudaStream_t)stream);
  }
  int threads = 512;
  size_t smem = 0;
  for (;; threads >>= 1) {   // the largest query block whose score tile and staged K / V fit
    const int qb = (threads / 32) * kQPW;
    smem = (size_t)T * (d + 8) * 2 * (split ? 2 : 1) + (size_t)qb * T * 4 + (size_t)qb * d * 4;
    if (smem <= 227 * 1024 - 1024 && (qb < 2 * T || threads == 128)) break;
    if (threads == 128) return set_error(B2D_E_INVALID, "b2d_attention: T=%d d=%d needs %zu bytes of shared memory", T, d, smem);
  }
  const int kQB = (threads / 32) * kQPW;
  dim3 grid((T + kQB - 1) / kQB, N * heads);
  const float scale = 1.0f / sqrtf((float)d);
  // opt in to the full dynamic shared-memory carve-out once per kernel variant (not a stream operation)
  static unsigned long long configured[2] = {0, 0};
  {
    cudaError_t e = split ? smem_attr_once(attention_kernel<true>, 226 * 1024, configured[1])
                          : smem_attr_once(attention_kernel<false>, 226 * 1024, configured[0]);
    if (e != cudaSuccess) return set_error(B2D_E_CUDA, "attention smem attr: %s", cudaGetErrorString(e));
  }
  if (split) {
    attention_kernel<true><<<grid, threads, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)qkv_lo, (__nv_bfloat16*)out, (__nv_bfloat16*)out_lo, T, C, heads, scale, 0);
  } else {
    attention_kernel<false><<<grid, threads, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)qkv, nullptr, (__nv_bfloat16*)out, nullptr, T, C, heads, scale, f16 ? 1 : 0);
  }
  return check_launch("attention_kernel");
}
