// Implicit-GEMM convolution engine for sm_100a: TMA box loads (zero-filled halo == zero padding)
// -> 128B-swizzled smem ring -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> fused epilogue
// (bias, residual add, GroupNorm partial sums, layout/precision of the consumer).
//
// One CTA computes a 128 x BLOCK_N output tile: 128 output positions forming a box
// (bn x bd x bh x bw) in (n, z, y, x), BLOCK_N output channels.  The K loop runs over
// (channel segment, tap, 64-channel chunk); for each step the A operand is ONE 5-D TMA box load
// of the activation tensor at the tap-shifted coordinates, the B operand one 2-D box of the
// pre-packed weight matrix.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (one TMEM lane quadrant each).
//
// Replaces (reference file:line): nn.Conv2d unet/blocks.py:29-36, models.py:120-128;
// nn.ConvTranspose2d unet/blocks.py:128-133; nn.Conv3d vae/blocks.py:155-169, encoder.py:30-68,
// decoder.py:31-71; nn.Linear/Conv1d projections of unet/blocks.py:196-207.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>
#include <new>
#include <type_traits>

#include "../../include/b2d.h"
#include "b2d_internal.h"
#include "b2d_ptx.cuh"

namespace b2d {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 elements = one 128-byte swizzle atom
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kThreads = 192;

struct ConvKParams {
  CUtensorMap tmapA[B2D_MAX_SEG];
  CUtensorMap tmapB;
  int nseg;
  int cchunks[B2D_MAX_SEG];
  int kbase[B2D_MAX_SEG];
  int cin[B2D_MAX_SEG];
  int ntaps;
  int8_t dz[B2D_MAX_TAPS], dy[B2D_MAX_TAPS], dx[B2D_MAX_TAPS];
  int stride_h, stride_w;
  int N, D, OH, OW;
  int lbw, lbh, lbd, lbn;  // log2 box extents, bw*bh*bd*bn == 128
  int tiles_w, tiles_h, tiles_d, tiles_n;
  int cout, nphase;
  const float* bias;
  void* out;
  void* out_lo;
  int out_mode;
  int out_H, out_W, out_sy, out_sx, out_oy, out_ox, out_cstride, out_coff;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* residual_lo;
  int res_cstride;
  double* stats;
  int stats_cpg;
  const float* out_scale;
  const float* out_mask;
  int skip_z;
  int out_f16, res_f16;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// two fp32 -> packed IEEE fp16, saturating to +-65504 (raw pre-GroupNorm storage must never produce inf)
__device__ __forceinline__ uint32_t pack_f16_sat(float a, float b) {
  uint32_t r;
  asm("{\n\t.reg .b16 lo, hi;\n\tcvt.rn.satfinite.f16.f32 lo, %1;\n\tcvt.rn.satfinite.f16.f32 hi, %2;\n\tmov.b32 %0, {lo, hi};\n\t}" : "=r"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 unpack_f16(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

template <int BLOCK_N, int STAGES, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) conv_igemm_kernel(const __grid_constant__ ConvKParams p) {
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr int CW = BLOCK_N < 32 ? 16 : 32;  // epilogue column chunk

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates ------------------------------------------------------------------
  int t = blockIdx.x;
  const int tw = t % p.tiles_w; t /= p.tiles_w;
  const int th = t % p.tiles_h; t /= p.tiles_h;
  const int td = t % p.tiles_d; t /= p.tiles_d;
  const int tn = t;
  const int x0 = tw << p.lbw, y0 = th << p.lbh, z0 = td << p.lbd, n0 = tn << p.lbn;
  const int gcol0 = blockIdx.y * BLOCK_N;  // row of the weight matrix

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) prefetch_tmap(&p.tmapA[s]);
    prefetch_tmap(&p.tmapB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer =========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const int cch = p.cchunks[s];
        for (int tp = 0; tp < p.ntaps; ++tp) {
          const int zz = z0 + p.dz[tp];
          if (p.skip_z && (zz < 0 || zz >= p.D)) continue;
          const int xx = x0 * p.stride_w + p.dx[tp];
          const int yy = y0 * p.stride_h + p.dy[tp];
          const int kb = p.kbase[s] + tp * p.cin[s];
          for (int c = 0; c < cch; ++c) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            uint8_t* a_dst = smem + stage * kStageBytes;
            tma_load_5d(a_dst, &p.tmapA[s], &full_bar[stage], c * kBlockK, xx, yy, zz, n0);
            tma_load_2d(a_dst + kABytes, &p.tmapB, &full_bar[stage], kb + c * kBlockK, gcol0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===========================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accum = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const int cch = p.cchunks[s];
        for (int tp = 0; tp < p.ntaps; ++tp) {
          const int zz = z0 + p.dz[tp];
          if (p.skip_z && (zz < 0 || zz >= p.D)) continue;
          for (int c = 0; c < cch; ++c) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
            const uint64_t adesc = umma_smem_desc(a_addr, 1024, 2);
            const uint64_t bdesc = umma_smem_desc(a_addr + kABytes, 1024, 2);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // advance 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
              umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
              accum = 1;
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ================================ epilogue ==============================================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int r = quad * 32 + lane;
    const int mw = (1 << p.lbw) - 1, mh = (1 << p.lbh) - 1, md = (1 << p.lbd) - 1;
    const int ox = x0 + (r & mw);
    const int oy = y0 + ((r >> p.lbw) & mh);
    const int oz = z0 + ((r >> (p.lbw + p.lbh)) & md);
    const int on = n0 + (r >> (p.lbw + p.lbh + p.lbd));
    const bool valid = ox < p.OW && oy < p.OH && oz < p.D && on < p.N;

    int phase_idx = 0, co_base = gcol0;
    if (p.nphase > 1) { phase_idx = gcol0 / p.cout; co_base = gcol0 - phase_idx * p.cout; }
    const int py = (p.nphase > 1) ? (phase_idx >> 1) : 0;
    const int px = (p.nphase > 1) ? (phase_idx & 1) : 0;
    const int out_y = oy * p.out_sy + p.out_oy + py;
    const int out_x = ox * p.out_sx + p.out_ox + px;
    const long long img = (long long)on * p.D + oz;
    const long long opix = (img * p.out_H + out_y) * p.out_W + out_x;

    // GroupNorm partial sums: segment of the warp that shares one sample
    const int rpi_log = p.lbw + p.lbh + p.lbd;  // rows per sample in the tile (log2)
    const int seg = rpi_log >= 5 ? 32 : (1 << rpi_log);
    const int cpg = p.stats_cpg;
    const int groups_per_n = (cpg > 0) ? (p.cout / cpg) : 0;
    float run_s = 0.f, run_ss = 0.f;
    int run_g = -1;

    auto flush = [&](float s, float ss, int g) {
      for (int off = seg >> 1; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
      }
      if ((lane & (seg - 1)) == 0 && on < p.N && g >= 0 && g < groups_per_n) {
        double* dst = p.stats + ((long long)on * groups_per_n + g) * 2;
        atomicAdd(dst, (double)s);
        atomicAdd(dst + 1, (double)ss);
      }
    };

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();

    float mask_v = 1.f;
    if (p.out_mode == 1 && p.out_mask != nullptr && valid) mask_v = p.out_mask[opix];

#pragma unroll 1
    for (int col0 = 0; col0 < BLOCK_N; col0 += CW) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(col0);
      if constexpr (CW == 32) tmem_ld_32x32(taddr, v); else tmem_ld_32x16(taddr, v);
      tmem_ld_wait();
      const int co0 = co_base + col0;
      float f[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const int co = co0 + j;
        float b = (p.bias != nullptr && co < p.cout) ? __ldg(p.bias + co) : 0.f;
        f[j] = __uint_as_float(v[j]) + b;
      }
      if (p.residual != nullptr && valid) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + opix * p.res_cstride + co0);
#pragma unroll
        for (int q = 0; q < CW / 8; ++q) {
          uint4 u = __ldg(rp + q);
          if (p.res_f16) {
            const float2 a = unpack_f16(u.x), b = unpack_f16(u.y), c = unpack_f16(u.z), d = unpack_f16(u.w);
            f[q * 8 + 0] += a.x; f[q * 8 + 1] += a.y; f[q * 8 + 2] += b.x; f[q * 8 + 3] += b.y;
            f[q * 8 + 4] += c.x; f[q * 8 + 5] += c.y; f[q * 8 + 6] += d.x; f[q * 8 + 7] += d.y;
          } else {
            f[q * 8 + 0] += bf16_lo(u.x); f[q * 8 + 1] += bf16_hi(u.x);
            f[q * 8 + 2] += bf16_lo(u.y); f[q * 8 + 3] += bf16_hi(u.y);
            f[q * 8 + 4] += bf16_lo(u.z); f[q * 8 + 5] += bf16_hi(u.z);
            f[q * 8 + 6] += bf16_lo(u.w); f[q * 8 + 7] += bf16_hi(u.w);
          }
        }
        if (p.residual_lo != nullptr) {
          const uint4* rl = reinterpret_cast<const uint4*>(p.residual_lo + opix * p.res_cstride + co0);
#pragma unroll
          for (int q = 0; q < CW / 8; ++q) {
            uint4 u = __ldg(rl + q);
            f[q * 8 + 0] += bf16_lo(u.x); f[q * 8 + 1] += bf16_hi(u.x);
            f[q * 8 + 2] += bf16_lo(u.y); f[q * 8 + 3] += bf16_hi(u.y);
            f[q * 8 + 4] += bf16_lo(u.z); f[q * 8 + 5] += bf16_hi(u.z);
            f[q * 8 + 6] += bf16_lo(u.w); f[q * 8 + 7] += bf16_hi(u.w);
          }
        }
      }
      // ---- GroupNorm partial sums (fp32 per thread, fp64 atomics per warp segment) -----------
      if (cpg > 0) {
        if (cpg >= CW) {
          const int g = co0 / cpg;
          if (g != run_g) {
            if (run_g >= 0) flush(run_s, run_ss, run_g);
            run_g = g; run_s = 0.f; run_ss = 0.f;
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < CW; ++j) {
              if (co0 + j < p.cout) { run_s += f[j]; run_ss += f[j] * f[j]; }
            }
          }
        } else {
          // cpg in {4, 8, 16}: several groups per chunk; compile-time trip counts keep f[] in registers
          auto small_groups = [&](auto cpg_c) {
            constexpr int CPG = decltype(cpg_c)::value;
            if constexpr (CPG <= CW) {
#pragma unroll
              for (int g0 = 0; g0 < CW; g0 += CPG) {
                float s = 0.f, ss = 0.f;
#pragma unroll
                for (int j = 0; j < CPG; ++j) { float x = valid ? f[g0 + j] : 0.f; s += x; ss += x * x; }
                flush(s, ss, (co0 + g0) / CPG);
              }
            }
          };
          if (cpg == 4) small_groups(std::integral_constant<int, 4>{});
          else if (cpg == 8) small_groups(std::integral_constant<int, 8>{});
          else small_groups(std::integral_constant<int, 16>{});
        }
      }
      // ---- store ---------------------------------------------------------------------------
      if (valid) {
        if (p.out_mode == 0) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + opix * p.out_cstride + p.out_coff + co0;
          uint32_t w[CW / 2];
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) w[j] = p.out_f16 ? pack_f16_sat(f[2 * j], f[2 * j + 1]) : pack_bf16(f[2 * j], f[2 * j + 1]);
          if (co0 + CW <= p.cout) {
#pragma unroll
            for (int q = 0; q < CW / 8; ++q)
              reinterpret_cast<uint4*>(op)[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
          } else {
            for (int j = 0; j < CW; ++j)
              if (co0 + j < p.cout) reinterpret_cast<uint16_t*>(op)[j] = (uint16_t)((j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xFFFFu));
          }
          if (p.out_lo != nullptr) {
            __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(p.out_lo) + opix * p.out_cstride + p.out_coff + co0;
            uint32_t wl[CW / 2];
#pragma unroll
            for (int j = 0; j < CW / 2; ++j)
              wl[j] = pack_bf16(f[2 * j] - bf16_lo(w[j]), f[2 * j + 1] - bf16_hi(w[j]));
            if (co0 + CW <= p.cout) {
#pragma unroll
              for (int q = 0; q < CW / 8; ++q)
                reinterpret_cast<uint4*>(ol)[q] = make_uint4(wl[4 * q], wl[4 * q + 1], wl[4 * q + 2], wl[4 * q + 3]);
            } else {
              for (int j = 0; j < CW; ++j)
                if (co0 + j < p.cout) ol[j] = __float2bfloat16_rn(f[j] - __bfloat162float(__float2bfloat16_rn(f[j])));
            }
          }
        } else if (p.out_mode == 1) {
          // planar fp32 [N][D][C][H][W]: lanes = consecutive x -> coalesced per channel
          float* ob = reinterpret_cast<float*>(p.out);
          const long long plane = (long long)p.out_H * p.out_W;
          const long long pix = (long long)out_y * p.out_W + out_x;
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const int co = co0 + j;
            if (co < p.cout) {
              float sc = p.out_scale != nullptr ? __ldg(p.out_scale + co) : 1.f;
              ob[(img * p.out_cstride + p.out_coff + co) * plane + pix] = f[j] * sc * mask_v;
            }
          }
        } else {
          float* op = reinterpret_cast<float*>(p.out) + opix * p.out_cstride + p.out_coff + co0;
          if (co0 + CW <= p.cout) {
#pragma unroll
            for (int q = 0; q < CW / 4; ++q)
              reinterpret_cast<float4*>(op)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
          } else {
            for (int j = 0; j < CW; ++j)
              if (co0 + j < p.cout) op[j] = f[j];
          }
        }
      }
    }
    if (cpg >= CW && run_g >= 0) flush(run_s, run_ss, run_g);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ============================================================================================
// host side
// ============================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void* tensor_map_encode_fn() {
  static void* fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
    cudaGetLastError();
    return nullptr;
  }
  fn = ptr;
  return fn;
}
static PFN_encodeTiled get_encode_fn() { return reinterpret_cast<PFN_encodeTiled>(tensor_map_encode_fn()); }

static int ilog2_ceil(int x) {
  int l = 0;
  while ((1 << l) < x) ++l;
  return l;
}

template <int BN, int ST, int MINB>
static int launch_conv(const ConvKParams& kp, dim3 grid, cudaStream_t st) {
  constexpr int smem = ST * (kABytes + BN * kBlockK * 2) + 1024 + 256;
  static_assert(MINB * (smem + 1024) <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BN, ST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(B2D_E_CUDA, "cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
    configured = true;
  }
  conv_igemm_kernel<BN, ST, MINB><<<grid, kThreads, smem, st>>>(kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "conv_igemm launch: %s", cudaGetErrorString(e));
  return B2D_OK;
}

}  // namespace b2d

struct b2d_conv_plan {
  b2d::ConvKParams kp;
  dim3 grid;
  int block_n;
  int kblocks;
};

using namespace b2d;

extern "C" int b2d_conv_plan_create(const b2d_conv_desc* d, b2d_conv_plan** out_plan) {
  if (!d || !out_plan) return set_error(B2D_E_INVALID, "null argument");
  *out_plan = nullptr;
  if (d->nseg < 1 || d->nseg > B2D_MAX_SEG) return set_error(B2D_E_INVALID, "nseg=%d out of range", d->nseg);
  if (d->ntaps < 1 || d->ntaps > B2D_MAX_TAPS) return set_error(B2D_E_INVALID, "ntaps=%d out of range", d->ntaps);
  if (d->N < 1 || d->D < 1 || d->H < 1 || d->W < 1 || d->OH < 1 || d->OW < 1)
    return set_error(B2D_E_INVALID, "bad extent N=%d D=%d H=%d W=%d OH=%d OW=%d", d->N, d->D, d->H, d->W, d->OH, d->OW);
  if (d->stride_h < 1 || d->stride_h > 2 || d->stride_w < 1 || d->stride_w > 2)
    return set_error(B2D_E_INVALID, "stride must be 1 or 2");
  if (d->cout < 1 || (d->nphase != 1 && d->nphase != 4)) return set_error(B2D_E_INVALID, "bad cout/nphase");
  if (!d->weight || !d->out) return set_error(B2D_E_INVALID, "null weight/out");
  if (d->ktot % kBlockK) return set_error(B2D_E_INVALID, "ktot=%d must be a multiple of 64", d->ktot);
  for (int s = 0; s < d->nseg; ++s) {
    if (!d->in[s]) return set_error(B2D_E_INVALID, "null input segment %d", s);
    if (d->cin[s] < kBlockK || d->cin[s] % kBlockK) return set_error(B2D_E_INVALID, "cin[%d]=%d must be a multiple of 64", s, d->cin[s]);
    if (d->kbase[s] < 0 || d->kbase[s] % kBlockK || d->kbase[s] + d->ntaps * d->cin[s] > d->ktot)
      return set_error(B2D_E_INVALID, "kbase[%d]=%d inconsistent with ktot=%d", s, d->kbase[s], d->ktot);
    if (reinterpret_cast<uintptr_t>(d->in[s]) & 15) return set_error(B2D_E_INVALID, "input %d not 16-byte aligned", s);
  }
  if (reinterpret_cast<uintptr_t>(d->weight) & 15) return set_error(B2D_E_INVALID, "weight not 16-byte aligned");

  // ---- M tiling: a 128-position box, widest along x -----------------------------------------
  int lw = ilog2_ceil(d->OW); if (lw > 4) lw = 4;
  int lh = ilog2_ceil(d->OH); if (lh > 7 - lw) lh = 7 - lw;
  int ld = ilog2_ceil(d->D);  if (ld > 7 - lw - lh) ld = 7 - lw - lh;
  int ln = 7 - lw - lh - ld;
  const long long tiles_m = (long long)((d->OW + (1 << lw) - 1) >> lw) * ((d->OH + (1 << lh) - 1) >> lh) *
                            ((d->D + (1 << ld) - 1) >> ld) * ((d->N + (1 << ln) - 1) >> ln);

  int bn = d->block_n;
  if (bn == 0) {
    // Widest N tile that still fills the machine: N=256 needs 96 B/clk of smem operand reads per MMA
    // (128 B/clk at N=128), but small-M layers (deep UNet levels) need the CTA count more.
    const long long cols = (long long)d->cout * d->nphase;
    const int sms = num_sms();
    if (d->cout % 256 == 0 && tiles_m * (cols / 256) >= 2LL * sms) bn = 256;
    else if (d->cout % 128 == 0 && (tiles_m * (cols / 128) >= sms || d->cout % 64 != 0)) bn = 128;
    else if (d->cout % 64 == 0) bn = 64;
    else if (d->cout <= 16 && d->nphase == 1) bn = 16;
    else return set_error(B2D_E_INVALID, "cout=%d: need a multiple of 64 or <= 16", d->cout);
  }
  if (bn != 16 && bn != 64 && bn != 128 && bn != 256) return set_error(B2D_E_INVALID, "block_n=%d unsupported", bn);
  if (bn >= 64 && d->cout % bn) return set_error(B2D_E_INVALID, "cout=%d not a multiple of block_n=%d", d->cout, bn);
  if (bn == 16 && (d->cout > 16 || d->nphase != 1)) return set_error(B2D_E_INVALID, "block_n=16 needs cout<=16");
  const int total_cols = (bn == 16) ? 16 : d->cout * d->nphase;
  if (d->wrows < total_cols) return set_error(B2D_E_INVALID, "wrows=%d < %d", d->wrows, total_cols);
  if (d->out_mode < 0 || d->out_mode > 2) return set_error(B2D_E_INVALID, "out_mode");
  if (d->out_mode == 0 && ((d->out_cstride % 8) || (d->out_coff % 8) || (reinterpret_cast<uintptr_t>(d->out) & 15)))
    return set_error(B2D_E_INVALID, "bf16 output needs cstride/coff multiples of 8 and 16-byte alignment");
  if (d->out_mode == 2 && ((d->out_cstride % 4) || (d->out_coff % 4) || (reinterpret_cast<uintptr_t>(d->out) & 15)))
    return set_error(B2D_E_INVALID, "fp32 NDHWC output needs cstride/coff multiples of 4");
  if (d->residual && (d->out_mode != 0 && d->out_mode != 2)) return set_error(B2D_E_INVALID, "residual needs out_mode 0/2");
  if (d->residual && ((d->res_cstride % 8) || (reinterpret_cast<uintptr_t>(d->residual) & 15) || bn == 16))
    return set_error(B2D_E_INVALID, "residual needs cstride multiple of 8, 16-byte alignment, block_n>=64");
  if (d->out_f16 && (d->out_mode != 0 || d->out_lo)) return set_error(B2D_E_INVALID, "out_f16 needs out_mode 0 without a lo part");
  if (d->res_f16 && (!d->residual || d->residual_lo)) return set_error(B2D_E_INVALID, "res_f16 needs a residual without a lo part");
  if (d->stats) {
    const int c = d->stats_cpg;
    if (!(c == 4 || c == 8 || c == 16 || (c >= 32 && c % 32 == 0)) || d->cout % c)
      return set_error(B2D_E_INVALID, "stats_cpg=%d unsupported for cout=%d", c, d->cout);
    if (bn == 16) return set_error(B2D_E_INVALID, "stats need block_n>=64");
  }

  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(B2D_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available (no CUDA driver / too old)");

  b2d_conv_plan* pl = new (std::nothrow) b2d_conv_plan();
  if (!pl) return set_error(B2D_E_INVALID, "out of host memory");
  ConvKParams& k = pl->kp;
  memset(&k, 0, sizeof(k));

  k.lbw = lw; k.lbh = lh; k.lbd = ld; k.lbn = ln;
  k.tiles_w = (d->OW + (1 << lw) - 1) >> lw;
  k.tiles_h = (d->OH + (1 << lh) - 1) >> lh;
  k.tiles_d = (d->D + (1 << ld) - 1) >> ld;
  k.tiles_n = (d->N + (1 << ln) - 1) >> ln;
  k.skip_z = (ld == 0) ? 1 : 0;

  // ---- tensor maps -----------------------------------------------------------------------
  for (int s = 0; s < d->nseg; ++s) {
    cuuint64_t dims[5] = {(cuuint64_t)d->cin[s], (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->D, (cuuint64_t)d->N};
    cuuint64_t strides[4];
    strides[0] = (cuuint64_t)d->cin[s] * 2;
    strides[1] = strides[0] * d->W;
    strides[2] = strides[1] * d->H;
    strides[3] = strides[2] * d->D;
    cuuint32_t box[5] = {(cuuint32_t)kBlockK, (cuuint32_t)((1 << lw) * d->stride_w), (cuuint32_t)((1 << lh) * d->stride_h),
                         (cuuint32_t)(1 << ld), (cuuint32_t)(1 << ln)};
    cuuint32_t estr[5] = {1, (cuuint32_t)d->stride_w, (cuuint32_t)d->stride_h, 1, 1};
    CUresult r = enc(&k.tmapA[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->in[s]), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      return set_error(B2D_E_CUDA, "cuTensorMapEncodeTiled(A seg %d) failed: %d (dims %d,%d,%d,%d,%d box %u,%u,%u,%u,%u)", s,
                       (int)r, d->cin[s], d->W, d->H, d->D, d->N, box[0], box[1], box[2], box[3], box[4]);
    }
    k.cchunks[s] = d->cin[s] / kBlockK;
    k.kbase[s] = d->kbase[s];
    k.cin[s] = d->cin[s];
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->ktot, (cuuint64_t)d->wrows};
    cuuint64_t strides[1] = {(cuuint64_t)d->ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&k.tmapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->weight), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      return set_error(B2D_E_CUDA, "cuTensorMapEncodeTiled(B) failed: %d (ktot %d wrows %d bn %d)", (int)r, d->ktot, d->wrows, bn);
    }
  }
  k.nseg = d->nseg;
  k.ntaps = d->ntaps;
  memcpy(k.dz, d->tap_dz, sizeof(k.dz));
  memcpy(k.dy, d->tap_dy, sizeof(k.dy));
  memcpy(k.dx, d->tap_dx, sizeof(k.dx));
  k.stride_h = d->stride_h; k.stride_w = d->stride_w;
  k.N = d->N; k.D = d->D; k.OH = d->OH; k.OW = d->OW;
  k.cout = d->cout; k.nphase = d->nphase;
  k.bias = d->bias;
  k.out = d->out; k.out_lo = d->out_lo; k.out_mode = d->out_mode;
  k.out_H = d->out_H; k.out_W = d->out_W;
  k.out_sy = d->out_sy; k.out_sx = d->out_sx; k.out_oy = d->out_oy; k.out_ox = d->out_ox;
  k.out_cstride = d->out_cstride; k.out_coff = d->out_coff;
  k.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  k.residual_lo = reinterpret_cast<const __nv_bfloat16*>(d->residual_lo);
  k.res_cstride = d->res_cstride;
  k.stats = d->stats; k.stats_cpg = d->stats ? d->stats_cpg : 0;
  k.out_scale = d->out_scale; k.out_mask = d->out_mask;
  k.out_f16 = d->out_f16 ? 1 : 0; k.res_f16 = d->res_f16 ? 1 : 0;

  pl->block_n = bn;
  pl->grid = dim3((unsigned)(k.tiles_w * k.tiles_h * k.tiles_d * k.tiles_n), (unsigned)(total_cols / bn), 1);
  int kb = 0;
  for (int s = 0; s < d->nseg; ++s) kb += d->ntaps * k.cchunks[s];
  pl->kblocks = kb;
  *out_plan = pl;
  return B2D_OK;
}

extern "C" int b2d_conv_plan_destroy(b2d_conv_plan* plan) {
  delete plan;
  return B2D_OK;
}

extern "C" int b2d_conv_plan_info(const b2d_conv_plan* plan, int32_t* grid_m, int32_t* grid_n, int32_t* block_n, int32_t* kblocks) {
  if (!plan) return set_error(B2D_E_INVALID, "null plan");
  if (grid_m) *grid_m = (int32_t)plan->grid.x;
  if (grid_n) *grid_n = (int32_t)plan->grid.y;
  if (block_n) *block_n = plan->block_n;
  if (kblocks) *kblocks = plan->kblocks;
  return B2D_OK;
}

extern "C" int b2d_conv_run(const b2d_conv_plan* plan, void* stream) {
  if (!plan) return set_error(B2D_E_INVALID, "null plan");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (plan->block_n) {
    // two co-resident CTAs per SM where shared memory allows: one CTA's epilogue overlaps the other's main loop
    case 16: return launch_conv<16, 5, 2>(plan->kp, plan->grid, st);
    case 64: return launch_conv<64, 4, 2>(plan->kp, plan->grid, st);
    case 128: return launch_conv<128, 3, 2>(plan->kp, plan->grid, st);
    case 256: return launch_conv<256, 4, 1>(plan->kp, plan->grid, st);
  }
  return set_error(B2D_E_INVALID, "bad block_n");
}
