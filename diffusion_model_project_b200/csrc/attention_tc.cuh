// Tensor-core attention core, device-side body (see attention.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "b2d_ptx.cuh"

namespace b2d {

constexpr int kTcThreads = 128;       // one thread per query row / TMEM lane
constexpr int kTcTmemCols = 512;      // S: columns [0, T), O: columns [256, 256 + min(d, 256))
constexpr int kTcOCol = 256;
constexpr int kQChunkBytes = 128 * 128;  // one 64-channel chunk of a 128-row K-major tile

struct AttnTcParams {
  CUtensorMap tmap;  // bf16 [N][T][3C], box {64, RB, 1}, 128B swizzle
  __nv_bfloat16* out;
  int T, C, heads, d;
  int RB;            // rows per TMA box = min(128, T)
  int f16;           // qkv / out / P hold IEEE fp16 instead of bf16
  float scale_log2e;
};

// One (image, head, 128-query tile) by the 128 threads of the CTA (`p` lives in the kernel's parameter space).
__device__ __forceinline__ void attention_tc_item(const AttnTcParams& p, int mtile, int nh, uint8_t* smem) {
  const AttnTcParams& tm = p;
  const int T = p.T, d = p.d, C = p.C;
  const int dch = d >> 6;                     // 64-channel chunks per head
  const int kv_chunk = T * 128;               // bytes of one [T x 64] chunk tile
  uint8_t* sQ = smem;                         // dch x [128 x 128 B]  (rows >= RB never read into valid output)
  uint8_t* sK = sQ + dch * kQChunkBytes;      // dch x [T x 128 B]
  uint8_t* sV = sK + dch * kv_chunk;          // dch x [T x 128 B]
  uint8_t* sP = smem;                         // ceil(T/64) x [128 x 128 B], aliases Q/K once S is complete
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + dch * kv_chunk);  // qk, v, s, o
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5;
  const int n = nh / p.heads, h = nh - n * p.heads;
  const int m0 = mtile * 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    prefetch_tmap(&tm.tmap);
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kTcTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    // ---- loads: Q tile + whole K on one barrier, whole V on another (lands while S is computed) ----
    const int RB = p.RB;
    mbar_arrive_expect_tx(&bars[0], (uint32_t)(dch * RB * 128 + dch * kv_chunk));
    for (int c = 0; c < dch; ++c) tma_load_3d(sQ + c * kQChunkBytes, &tm.tmap, &bars[0], h * d + c * 64, m0, n);
    for (int c = 0; c < dch; ++c)
      for (int r = 0; r < T; r += RB) tma_load_3d(sK + c * kv_chunk + r * 128, &tm.tmap, &bars[0], C + h * d + c * 64, r, n);
    mbar_arrive_expect_tx(&bars[1], (uint32_t)(dch * kv_chunk));
    for (int c = 0; c < dch; ++c)
      for (int r = 0; r < T; r += RB) tma_load_3d(sV + c * kv_chunk + r * 128, &tm.tmap, &bars[1], 2 * C + h * d + c * 64, r, n);
    // ---- S[128 x T] = Q K^T ---------------------------------------------------------------------
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc_s = umma_idesc_16(128, (uint32_t)T, p.f16);
    uint32_t accum = 0;
    for (int c = 0; c < dch; ++c) {
      const uint64_t adesc = umma_smem_desc(smem_u32(sQ + c * kQChunkBytes), 1024, 2);
      const uint64_t bdesc = umma_smem_desc(smem_u32(sK + c * kv_chunk), 1024, 2);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc_s, accum);
        accum = 1;
      }
    }
    umma_commit(&bars[2]);
  }
  __syncwarp();

  // ---- softmax over the keys: thread = query row, two passes over the TMEM-resident scores --------
  const int row = threadIdx.x;
  const uint32_t lane_addr = tmem_base + (uint32_t(warp * 32) << 16);
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  float mx = -INFINITY;
  if (T >= 32) {
    for (int c0 = 0; c0 < T; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(lane_addr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
  } else {
    uint32_t v[32];
    tmem_ld_32x16(lane_addr, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
  }
  const float sc = p.scale_log2e;
  const float mxs = mx * sc;
  float sum = 0.f;
  uint8_t* prow = sP + row * 128;
  const int sw = row & 7;
  const int f16 = p.f16;
  auto emit = [&](const uint32_t* v, int c0, int ncol) {
    // ncol consecutive keys starting at c0 (c0 % 8 == 0): bf16 pairs into the swizzled K-major P tile
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      if (j8 * 8 < ncol) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float e0 = exp2f(fmaf(__uint_as_float(v[j8 * 8 + 2 * q]), sc, -mxs));
          const float e1 = exp2f(fmaf(__uint_as_float(v[j8 * 8 + 2 * q + 1]), sc, -mxs));
          if (f16) {
            __half2 hv = __floats2half2_rn(e0, e1);
            sum += __low2float(hv) + __high2float(hv);
            w[q] = *reinterpret_cast<uint32_t*>(&hv);
          } else {
            __nv_bfloat162 hv = __floats2bfloat162_rn(e0, e1);
            sum += __low2float(hv) + __high2float(hv);
            w[q] = *reinterpret_cast<uint32_t*>(&hv);
          }
        }
        const int key = c0 + j8 * 8;
        const int chunk16 = (key & 63) >> 3;
        *reinterpret_cast<uint4*>(prow + (key >> 6) * kQChunkBytes + ((chunk16 ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  };
  if (T >= 32) {
    for (int c0 = 0; c0 < T; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(lane_addr + c0, v);
      tmem_ld_wait();
      emit(v, c0, 32);
    }
  } else {
    uint32_t v[32];
    tmem_ld_32x16(lane_addr, v);
    tmem_ld_wait();
    emit(v, 0, 16);
  }
  const float inv = 1.f / sum;
  fence_proxy_async();  // P (generic-proxy stores) must be visible to the tensor core's async-proxy reads
  tc_fence_before();
  __syncthreads();

  // ---- O[128 x d] = P V, at most 256 output columns per pass -------------------------------------
  const int DC = d < 256 ? d : 256;
  const int q = m0 + row;
  for (int g = 0; g * DC < d; ++g) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      if (g == 0) { mbar_wait(&bars[1], 0); tc_fence_after(); }
      const uint32_t idesc_o = umma_idesc_16(128, (uint32_t)DC, p.f16) | kUmmaBMajorMN;
      uint32_t accum = 0;
      for (int k0 = 0; k0 < T; k0 += 16) {
        const uint64_t adesc = umma_smem_desc(smem_u32(sP + (k0 >> 6) * kQChunkBytes), 1024, 2) + 2 * ((k0 & 63) >> 4);
        const uint64_t bdesc = umma_smem_desc_mn(smem_u32(sV + (g * (DC >> 6)) * kv_chunk + k0 * 128), (uint32_t)kv_chunk, 1024, 2);
        umma_bf16(tmem_base + kTcOCol, adesc, bdesc, idesc_o, accum);
        accum = 1;
      }
      umma_commit(&bars[3]);
    }
    __syncwarp();
    mbar_wait(&bars[3], g & 1);
    tc_fence_after();
    __nv_bfloat16* orow = p.out + ((long long)n * T + q) * C + (long long)h * d + g * DC;
    for (int c0 = 0; c0 < DC; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(lane_addr + kTcOCol + c0, v);
      tmem_ld_wait();
      if (q < T) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float o0 = __uint_as_float(v[j8 * 8 + 2 * e]) * inv, o1 = __uint_as_float(v[j8 * 8 + 2 * e + 1]) * inv;
            if (f16) {
              __half2 hv = __floats2half2_rn(o0, o1);
              w[e] = *reinterpret_cast<uint32_t*>(&hv);
            } else {
              __nv_bfloat162 hv = __floats2bfloat162_rn(o0, o1);
              w[e] = *reinterpret_cast<uint32_t*>(&hv);
            }
          }
          *reinterpret_cast<uint4*>(orow + c0 + j8 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // every row has drained O before the next pass overwrites it / before dealloc
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTcTmemCols);
  }
}


// host: fill `p` (TMA descriptor over qkv [N][T][3C], shapes, scale) after checking that the tensor-core path applies
int attention_tc_params(const void* qkv, void* out, int N, int T, int C, int heads, int f16, AttnTcParams* p);
// dynamic shared memory one item needs
size_t attention_tc_smem(int T, int C, int heads);

}  // namespace b2d
