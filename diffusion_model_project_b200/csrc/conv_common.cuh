// Shared definitions of the implicit-GEMM conv engine (conv_plan.cu: planning; conv_engine.cu / conv_v2.cuh: the
// persistent, halo-staged, split-K kernel): kernel parameter block, packing helpers and the fused epilogue.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/b2d.h"
#include "b2d_internal.h"
#include "b2d_ptx.cuh"
#include "scheduler_math.cuh"

namespace b2d {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 elements = one 128-byte swizzle atom
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kThreads = 192;
constexpr int kMaxPieces = 16;  // most items that may share one tile's K loop (split-K <= 16; stream-K is planned within it)

// Division by a launch-constant through a precomputed multiplier (valid for 0 <= n < 2^31): the unit decode runs once
// per work unit in every warp role, and a hardware-emulated integer division costs ~25 instructions.
struct FastDiv {
  uint32_t d, mul, shr;
  __host__ void set(uint32_t div) {
    d = div;
    if (div <= 1) { mul = 0; shr = 0; return; }
    uint32_t lg = 0;
    while ((1u << lg) < div) ++lg;  // ceil(log2(div))
    const uint32_t p = 31 + lg;
    mul = (uint32_t)(((1ull << p) + div - 1) / div);
    shr = p - 32;
  }
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    return d <= 1 ? n : (__umulhi(n, mul) >> shr);
#else
    return d <= 1 ? n : (uint32_t)(((uint64_t)n * mul) >> 32) >> shr;
#endif
  }
  __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

__device__ __forceinline__ uint32_t pack_bf16_(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

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
  // ---- work units of the persistent kernel ----
  int halo;          // 1: one 18x18 halo box feeds the 9 in-plane taps of two 8x16 M=128 halves
  int ksplit;        // K-loop splits per output tile (fp32 partials in `ws`, last arriver reduces)
  int ngroups;       // A-operand loads in the K loop: nseg x (halo ? z-taps : taps) x 64-channel chunks
  int gtaps;         // weight tiles consumed per A load: halo ? 9 : 1
  int goff[B2D_MAX_SEG + 1];  // first group of each segment
  int tiles_ncol;    // N tiles
  int num_units;     // tiles_m * tiles_ncol * ksplit
  int contig;        // 1: a CTA walks a contiguous range of units (few sample changes -> few GroupNorm flushes)
  float* ws;         // [tile][ksplit][128][BLOCK_N] fp32 partial accumulators
  int* counters;     // [tile] arrival counters (self-resetting)
  FastDiv fd_ksplit, fd_ncol, fd_w, fd_h, fd_d;  // unit index -> (split, N tile, x, y, z, n tile)
  // K loop of one tile in B-stage steps (conv_work.cuh): spg steps per A-operand load (halo: in-plane taps / taps per
  // weight stage; generic: 1), ksteps = ngroups * spg.  streamk: the flat (tile, step) space of total_steps =
  // tiles * ksteps steps is cut into gridDim.x equal contiguous ranges, one per CTA (tile boundaries inside a range
  // are whole tiles; a tile cut by a range boundary is reduced through the workspace by whichever piece arrives last)
  int spg, ksteps, streamk, total_steps;
  int pair;          // 1: launched as clusters of two CTAs on 256-row M super-tiles (tcgen05 cta_group::2); the unit walk
                     // then runs over PAIR tiles and num_units counts pair units
  FastDiv fd_ksteps, fd_total;
  // ---- fused input normalisation (halo mode): A tiles are raw pre-GroupNorm values; dedicated warps rewrite each
  // staged tile in shared memory as bf16 silu(gamma * (x - mean) * rstd + beta) before the MMAs read it ----
  int xform;                   // 1: enabled (single segment)
  const double* in_stats;      // [N][cin / in_cpg][2] (sum, sumsq) of the producer
  const float* in_gamma;       // [cin_real]
  const float* in_beta;
  int in_cpg, in_creal;        // channels per group, real (unpadded) channel count
  int in_f16, in_act;          // raw values are fp16 (else bf16); apply SiLU
  float in_eps;
  double in_count;             // elements per (sample, group): in_cpg * D * H * W
  const float* in_temb;        // optional per-(sample, channel) addend after the activation (time embedding)
  const int* in_temb_row;
  int in_temb_row_stride, in_temb_ncols, in_temb_col;
  // ---- out_mode 3: fused sampler update (b2d_conv_desc.sched_*) ----
  float* sch_x;
  const float* sch_noise;
  const float* sch_coef;
  int* sch_step;
  unsigned int* sch_ticket;
  const unsigned long long* sch_seed_dev;
  unsigned long long sch_seed;
  int sch_kind, sch_step_off, sch_step_inc, sch_clip;
  float sch_lo, sch_hi;
  __nv_bfloat16* sch_bf16;
  __nv_bfloat16* sch_bf16_lo;
  int sch_bf16_stride;
  int op_f16;                  // MMA operands (and the fused sampler update's 16-bit copy) are fp16 instead of bf16
};

// per-kernel constants of the fused sampler update (out_mode 3), loaded once by every epilogue thread
struct SchedCtx {
  Coef k;
  unsigned long long seed;
  int row;
  bool use_noise;
};
__device__ __forceinline__ SchedCtx sched_ctx_load(const ConvKParams& p) {
  SchedCtx sc;
  sc.row = (p.sch_step ? *p.sch_step : 0) + p.sch_step_off;
  sc.k = load_coef(p.sch_coef, sc.row);
  sc.seed = p.sch_seed_dev ? *p.sch_seed_dev : p.sch_seed;
  sc.use_noise = sc.k.s != 0.f;
  return sc;
}
// One pixel's channels: eps = f[0 .. cout) -> x <- step(x, eps, z), in the operation order of scheduler_step_kernel
// (element index = pixel * cout + c, Philox counter = element index / 4: the same noise as the stand-alone kernel).
template <int CW>
__device__ __forceinline__ void sched_update_row(const ConvKParams& p, const SchedCtx& sc, long long opix, const float (&f)[CW]) {
  const int C = p.cout;
  float* xr = p.sch_x + opix * C;
#pragma unroll
  for (int q = 0; q < CW / 4; ++q) {
    if (4 * q < C) {
      const float4 x = *reinterpret_cast<const float4*>(xr + 4 * q);
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sc.use_noise) {
        if (p.sch_noise != nullptr) z = __ldg(reinterpret_cast<const float4*>(p.sch_noise + opix * C + 4 * q));
        else z = philox_normal4((uint64_t)((opix * C) / 4 + q), (uint32_t)sc.row, sc.seed);
      }
      float4 o;
      o.x = step_one(x.x, f[4 * q + 0], z.x, sc.k, p.sch_kind, p.sch_clip, p.sch_lo, p.sch_hi, sc.use_noise);
      o.y = step_one(x.y, f[4 * q + 1], z.y, sc.k, p.sch_kind, p.sch_clip, p.sch_lo, p.sch_hi, sc.use_noise);
      o.z = step_one(x.z, f[4 * q + 2], z.z, sc.k, p.sch_kind, p.sch_clip, p.sch_lo, p.sch_hi, sc.use_noise);
      o.w = step_one(x.w, f[4 * q + 3], z.w, sc.k, p.sch_kind, p.sch_clip, p.sch_lo, p.sch_hi, sc.use_noise);
      *reinterpret_cast<float4*>(xr + 4 * q) = o;
      if (p.sch_bf16 != nullptr) {
        const uint32_t w0 = pack16(o.x, o.y, p.op_f16), w1 = pack16(o.z, o.w, p.op_f16);
        *reinterpret_cast<uint2*>(p.sch_bf16 + opix * p.sch_bf16_stride + 4 * q) = make_uint2(w0, w1);
        if (p.sch_bf16_lo != nullptr) {  // hi/lo split (fp32x mode): bf16 only, checked at plan creation
          const uint32_t l0 = pack_bf16_(o.x - __uint_as_float(w0 << 16), o.y - __uint_as_float(w0 & 0xFFFF0000u));
          const uint32_t l1 = pack_bf16_(o.z - __uint_as_float(w1 << 16), o.w - __uint_as_float(w1 & 0xFFFF0000u));
          *reinterpret_cast<uint2*>(p.sch_bf16_lo + opix * p.sch_bf16_stride + 4 * q) = make_uint2(l0, l1);
        }
      }
    }
  }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) { return pack_bf16_(a, b); }
// two fp32 -> packed IEEE fp16, saturating to +-65504 (raw pre-GroupNorm storage must never produce inf)
__device__ __forceinline__ uint32_t pack_f16_sat(float a, float b) {
  uint32_t r;
  asm("{\n\t.reg .b16 lo, hi;\n\tcvt.rn.satfinite.f16.f32 lo, %1;\n\tcvt.rn.satfinite.f16.f32 hi, %2;\n\tmov.b32 %0, {lo, hi};\n\t}" : "=r"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 unpack_f16(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }


// ---------------------------------------------------------------------------------------------
// Fused epilogue of one accumulator row (thread = output position, TMEM lane): + bias, + residual,
// GroupNorm partial sums, store in the consumer's layout.  `load(col0, f)` fills CW raw fp32
// accumulator columns starting at col0 (from TMEM, or from the split-K partials).
// ---------------------------------------------------------------------------------------------
struct EpiRow {
  bool valid;        // the row maps to a real output position
  int on;            // sample index (GroupNorm statistics are per sample)
  long long img;     // (n * D + z)
  long long opix;    // (img * out_H + out_y) * out_W + out_x
  int out_y, out_x;
};

// Sum V values per lane across the 32 lanes of a warp with a transposing butterfly: each step exchanges
// half of the remaining values, so V values cost (V - 1) + (5 - log2 V) shuffles instead of 5 V.
// On return vals[0] of lane L holds the warp total of value index (L >> (5 - log2 V)) (bit-reversed pairing is
// avoided by always keeping the half selected by the lane bit), i.e. lanes L with the low (5 - log2 V) bits
// zero own one value each.
template <int V>
__device__ __forceinline__ void warp_transpose_sum(float (&vals)[V], int lane) {
  static_assert(V >= 1 && V <= 32 && (V & (V - 1)) == 0, "V must be a power of two");
  int off = 16;
#pragma unroll
  for (int cur = V; cur > 1; cur >>= 1) {
    const int half = cur >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? vals[i] : vals[i + half];
      const float keep = upper ? vals[i + half] : vals[i];
      vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  for (; off > 0; off >>= 1) vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], off);
}

// SMEM_STATS: the whole warp belongs to ONE sample (halo tiles) and GroupNorm sums go to per-CTA fp64
// accumulators in shared memory (sm_stats[group][2], one private copy per epilogue warp -- fp64 shared-memory atomics
// are CAS loops and serialise badly); the kernel pushes them to global memory when the
// sample changes.  Otherwise sums go straight to global memory with fp64 atomics (rows of a warp may
// belong to `32 / seg` different samples).
template <int BLOCK_N, int CW, bool SMEM_STATS, bool STAGE = false, class Loader>
__device__ __forceinline__ void conv_epilogue_row(const ConvKParams& p, const EpiRow& rw, int co_base, int lane, int seg,
                                                  double* sm_stats, Loader&& load, double* thr_acc = nullptr,
                                                  uint32_t sm_bias = 0u, uint32_t sm_stage = 0u, const SchedCtx* sched = nullptr) {
  // sm_stage != 0: shared-memory address of this warp's 2 KB staging tile.  16-bit channels-last stores then go through
  // it so that four lanes write one row's 64 contiguous bytes (full 32-byte sectors) instead of every lane writing 16
  // bytes of its own row -- thin-K layers (1x1 projections, transposed convs) were bound by those partial-sector stores.
  // sm_bias != 0: shared-memory address of the tile's BLOCK_N bias values (staged by the caller while it waited for the
  // accumulator; zero beyond cout) -- a global bias load after the TMEM wait exposed a full L2 round trip per chunk
  const int cpg = p.stats_cpg;
  const int groups_per_n = (cpg > 0) ? (p.cout / cpg) : 0;
  float run_s = 0.f, run_ss = 0.f;
  int run_g = -1;
  int cur_g = 0, cur_rem = 0;  // group of column co0 and offset inside it, tracked without divisions
  if (cpg >= CW) { cur_g = co_base / cpg; cur_rem = co_base - cur_g * cpg; }
  const bool valid = rw.valid;
  const int on = rw.on;
  auto flush = [&](float s, float ss, int g) {
    if constexpr (SMEM_STATS) {
      float v2[2] = {s, ss};
      warp_transpose_sum<2>(v2, lane);
      if ((lane & 15) == 0 && g >= 0 && g < groups_per_n) sm_stats[g * 2 + (lane >> 4)] += (double)v2[0];  // this warp's private slot: no atomics
    } else {
      for (int off = seg >> 1; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
      }
      if ((lane & (seg - 1)) == 0 && on < p.N && g >= 0 && g < groups_per_n) {
        double* dst = p.stats + ((long long)on * groups_per_n + g) * 2;
        atomicAdd(dst, (double)s);
        atomicAdd(dst + 1, (double)ss);
      }
    }
  };
  float mask_v = 1.f;
  if (p.out_mode == 1 && p.out_mask != nullptr && valid) mask_v = p.out_mask[rw.opix];
  // residual: requested BEFORE the accumulator load so that its L2 round trip overlaps the TMEM load and wait
  const bool has_res = p.residual != nullptr && valid;
  const uint4* res_row = has_res ? reinterpret_cast<const uint4*>(p.residual + rw.opix * p.res_cstride + co_base) : nullptr;

  const bool stage = STAGE && CW == 32 && sm_stage != 0u && p.out_mode == 0 && p.out_lo == nullptr;  // warp-uniform
  long long st_base[4] = {0, 0, 0, 0};
  uint32_t st_ok = 0u;
  if constexpr (STAGE) if (stage) {
    const long long my_base = rw.opix * p.out_cstride + p.out_coff + co_base;  // 16-bit elements
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int src = it * 8 + (lane >> 2);
      st_base[it] = __shfl_sync(0xffffffffu, my_base, src);
      st_ok |= (uint32_t)__shfl_sync(0xffffffffu, (int)valid, src) << it;
    }
  }

#pragma unroll 1
  for (int col0 = 0; col0 < BLOCK_N; col0 += CW) {
    uint4 rcur[CW / 8];
    if (has_res) {
#pragma unroll
      for (int q = 0; q < CW / 8; ++q) rcur[q] = __ldg(res_row + col0 / 8 + q);
    }
    float f[CW];
    load(col0, f);
    const int co0 = co_base + col0;
    const bool full = co0 + CW <= p.cout;
    if (p.bias != nullptr && sm_bias != 0u) {
#pragma unroll
      for (int q = 0; q < CW / 4; ++q) {
        const float4 b = lds128f(sm_bias + (uint32_t)(col0 + 4 * q) * 4u);
        f[4 * q] += b.x; f[4 * q + 1] += b.y; f[4 * q + 2] += b.z; f[4 * q + 3] += b.w;
      }
    } else if (p.bias != nullptr) {
      if (full) {
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + co0) + q);
          f[4 * q] += b.x; f[4 * q + 1] += b.y; f[4 * q + 2] += b.z; f[4 * q + 3] += b.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (co0 + j < p.cout) f[j] += __ldg(p.bias + co0 + j);
      }
    }
    if (has_res) {
      if (p.res_f16) {
#pragma unroll
        for (int q = 0; q < CW / 8; ++q) {
          const uint4 u = rcur[q];
          const float2 a = unpack_f16(u.x), b = unpack_f16(u.y), c = unpack_f16(u.z), d = unpack_f16(u.w);
          f[q * 8 + 0] += a.x; f[q * 8 + 1] += a.y; f[q * 8 + 2] += b.x; f[q * 8 + 3] += b.y;
          f[q * 8 + 4] += c.x; f[q * 8 + 5] += c.y; f[q * 8 + 6] += d.x; f[q * 8 + 7] += d.y;
        }
      } else {
#pragma unroll
        for (int q = 0; q < CW / 8; ++q) {
          const uint4 u = rcur[q];
          f[q * 8 + 0] += bf16_lo(u.x); f[q * 8 + 1] += bf16_hi(u.x);
          f[q * 8 + 2] += bf16_lo(u.y); f[q * 8 + 3] += bf16_hi(u.y);
          f[q * 8 + 4] += bf16_lo(u.z); f[q * 8 + 5] += bf16_hi(u.z);
          f[q * 8 + 6] += bf16_lo(u.w); f[q * 8 + 7] += bf16_hi(u.w);
        }
      }
      if (p.residual_lo != nullptr) {
        const uint4* rl = reinterpret_cast<const uint4*>(p.residual_lo + rw.opix * p.res_cstride + co0);
#pragma unroll
        for (int q = 0; q < CW / 8; ++q) {
          const uint4 u = __ldg(rl + q);
          f[q * 8 + 0] += bf16_lo(u.x); f[q * 8 + 1] += bf16_hi(u.x);
          f[q * 8 + 2] += bf16_lo(u.y); f[q * 8 + 3] += bf16_hi(u.y);
          f[q * 8 + 4] += bf16_lo(u.z); f[q * 8 + 5] += bf16_hi(u.z);
          f[q * 8 + 6] += bf16_lo(u.w); f[q * 8 + 7] += bf16_hi(u.w);
        }
      }
    }
    // ---- GroupNorm partial sums -----------------------------------------------------------------
    if (cpg > 0) {
      if (cpg >= CW) {
        // a chunk lies inside one group (cpg is a multiple of CW): running sums, reduced when the group ends
        if (cur_g != run_g) {
          if (run_g >= 0) flush(run_s, run_ss, run_g);
          run_g = cur_g; run_s = 0.f; run_ss = 0.f;
        }
        if (valid) {
          if (full) {
#pragma unroll
            for (int j = 0; j < CW; ++j) { run_s += f[j]; run_ss = fmaf(f[j], f[j], run_ss); }
          } else {
#pragma unroll
            for (int j = 0; j < CW; ++j)
              if (co0 + j < p.cout) { run_s += f[j]; run_ss = fmaf(f[j], f[j], run_ss); }
          }
        }
        cur_rem += CW;
        if (cur_rem >= cpg) { cur_rem -= cpg; ++cur_g; }
      } else {
        // cpg in {4, 8, 16}: CW / cpg groups per chunk
        auto small_groups = [&](auto cpg_c) {
          constexpr int CPG = decltype(cpg_c)::value;
          if constexpr (CPG <= CW) {
            constexpr int G = CW / CPG;
            if constexpr (SMEM_STATS) {
              float vals[2 * G];
#pragma unroll
              for (int g0 = 0; g0 < G; ++g0) {
                float s = 0.f, ss = 0.f;
#pragma unroll
                for (int j = 0; j < CPG; ++j) { const float x = f[g0 * CPG + j]; s += x; ss = fmaf(x, x, ss); }
                vals[g0] = valid ? s : 0.f;
                vals[G + g0] = valid ? ss : 0.f;
              }
              warp_transpose_sum<2 * G>(vals, lane);
              constexpr int SH = 5 - (G == 8 ? 4 : G == 4 ? 3 : G == 2 ? 2 : 1);  // 32 / (2G) lanes per value
              if ((lane & ((1 << SH) - 1)) == 0) {
                const int vi = lane >> SH;            // value index: [0,G) sums, [G,2G) sums of squares
                const int g = co0 / CPG + (vi & (G - 1));
                if (g < groups_per_n) sm_stats[g * 2 + (vi >= G ? 1 : 0)] += (double)vals[0];  // warp-private slot
              }
            } else {
#pragma unroll
              for (int g0 = 0; g0 < CW; g0 += CPG) {
                float s = 0.f, ss = 0.f;
#pragma unroll
                for (int j = 0; j < CPG; ++j) { float x = valid ? f[g0 + j] : 0.f; s += x; ss += x * x; }
                flush(s, ss, (co0 + g0) / CPG);
              }
            }
          }
        };
        if (cpg == 4) small_groups(std::integral_constant<int, 4>{});
        else if (cpg == 8) small_groups(std::integral_constant<int, 8>{});
        else small_groups(std::integral_constant<int, 16>{});
      }
    }
    // ---- store -------------------------------------------------------------------------------
    bool staged = false;
    if constexpr (STAGE && CW == 32) {
      if (stage && full) {
        uint32_t w[16];
        if (p.out_f16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = pack_f16_sat(f[2 * j], f[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
        }
        // row-per-lane in, 16-byte slots XOR-swizzled so that both phases are bank-conflict free
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(sm_stage + (uint32_t)(lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)), make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
        __syncwarp();
        uint16_t* ob = reinterpret_cast<uint16_t*>(p.out);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = it * 8 + (lane >> 2), vec = lane & 3;
          const uint4 v = lds128(sm_stage + (uint32_t)(row * 64 + ((vec ^ ((row >> 1) & 3)) << 4)));
          if ((st_ok >> it) & 1u) reinterpret_cast<uint4*>(ob + st_base[it] + col0)[vec] = v;
        }
        __syncwarp();
        staged = true;
      }
    }
    if (valid && !staged) {
      if (p.out_mode == 0) {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + rw.opix * p.out_cstride + p.out_coff + co0;
        uint32_t w[CW / 2];
        if (p.out_f16) {
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) w[j] = pack_f16_sat(f[2 * j], f[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) w[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
        }
        if (full) {
#pragma unroll
          for (int q = 0; q < CW / 8; ++q)
            reinterpret_cast<uint4*>(op)[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        } else {
          for (int j = 0; j < CW; ++j)
            if (co0 + j < p.cout) reinterpret_cast<uint16_t*>(op)[j] = (uint16_t)((j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xFFFFu));
        }
        if (p.out_lo != nullptr) {
          __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(p.out_lo) + rw.opix * p.out_cstride + p.out_coff + co0;
          uint32_t wl[CW / 2];
#pragma unroll
          for (int j = 0; j < CW / 2; ++j)
            wl[j] = pack_bf16(f[2 * j] - bf16_lo(w[j]), f[2 * j + 1] - bf16_hi(w[j]));
          if (full) {
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
        const long long pix = (long long)rw.out_y * p.out_W + rw.out_x;
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          const int co = co0 + j;
          if (co < p.cout) {
            float sc = p.out_scale != nullptr ? __ldg(p.out_scale + co) : 1.f;
            ob[(rw.img * p.out_cstride + p.out_coff + co) * plane + pix] = f[j] * sc * mask_v;
          }
        }
      } else {
        // mode 3: the sampler update consumes eps right here; the fp32 eps store below is optional (p.out may be NULL)
        if (p.out_mode == 3 && sched != nullptr) sched_update_row<CW>(p, *sched, rw.opix, f);
        float* op = reinterpret_cast<float*>(p.out) + rw.opix * p.out_cstride + p.out_coff + co0;
        if (p.out == nullptr) {
        } else if (full) {
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
  if (cpg >= CW && run_g >= 0) {
    if (thr_acc != nullptr) {  // one group per sample: the caller reduces across lanes when the sample changes
      thr_acc[0] += (double)run_s;
      thr_acc[1] += (double)run_ss;
    } else {
      flush(run_s, run_ss, run_g);
    }
  }
}

}  // namespace b2d

// One planned convolution: kernel parameters + launch geometry (opaque in include/b2d.h).
struct b2d_conv_plan {
  b2d::ConvKParams kp;
  dim3 grid;
  int block_n;
  int kblocks;
  long long ws_bytes;  // workspace bytes the plan uses (split-K partials + counters)
};

namespace b2d {
// conv_engine.cu
int launch_conv_v2(const b2d_conv_plan* plan, cudaStream_t st);
}  // namespace b2d
