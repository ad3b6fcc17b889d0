// Cross-layer persistent kernel ("chain"): a whole run of dependent layers -- convs / GEMMs, GroupNorm applies,
// max-pools, attention cores -- executes inside ONE cooperative launch; a grid barrier replaces every kernel
// boundary.  At 8 samples per GPU the UNet step is 89 launches whose fixed floors (9.6 us per conv launch, ~5 us per
// elementwise launch: launch, prologue, first TMA round trip, drain) are half of the step; an op boundary inside a
// cooperative persistent kernel costs 1.7-2.0 us (tools/probe/gridbar_probe.cu, profiles/r1_gridbarrier_probe.txt).
//
// One CTA per SM (cooperative launch: all co-resident, so the barrier cannot deadlock; every spin is bounded anyway).
// TMEM (512 columns) is allocated once; every op lays its own mbarriers over the dynamic shared memory and
// invalidates them when it is done.  The conv / attention bodies are the same device functions the stand-alone kernels
// run (conv_v2.cuh, attention_tc.cuh); GroupNorm apply and max-pool are re-stated here for a 352-thread CTA.
#include <string.h>
#include <new>
#include <vector>

#include "attention_tc.cuh"
#include "conv_v2.cuh"

namespace b2d {

constexpr int kChainThreads = 352;  // = the halo conv variant: producer, MMA, 8 epilogue warps, weight producer
constexpr int cmax(int a, int b) { return a > b ? a : b; }
// dynamic shared memory: the largest layer variant (halo, BLOCK_N = 64: three 41 KB activation stages + four 24 KB
// weight stages = 224.2 KB)
constexpr int kChainSmem =
    cmax(cmax(cmax(V2Cfg<16, true>::SMEM, V2Cfg<64, true>::SMEM), cmax(V2Cfg<128, true>::SMEM, V2Cfg<256, true>::SMEM)),
         cmax(cmax(V2Cfg<16, false>::SMEM, V2Cfg<64, false>::SMEM), cmax(V2Cfg<128, false>::SMEM, V2Cfg<256, false>::SMEM)));
static_assert(kChainSmem + 2048 <= 227 * 1024, "chain: dynamic + static shared memory");

struct GnOp {
  const uint4* x;
  uint4* y;
  long long P;  // positions per sample
  int N, C, cpg;
  const double* stats;
  const float* gamma;
  const float* beta;
  float eps;
  int act;
  const float* temb;
  const int* temb_row;
  int temb_row_stride, temb_ld, temb_col;
  double* stats_out;
  int in_f16;
};
struct PoolOp {
  const uint4* x;
  uint4* y;
  int N, H, W, C;
  double* stats;
};
struct AttnOp {
  AttnTcParams p;
  int N, mtiles;
};
struct ZeroOp {
  uint4* ptr;
  long long n16;  // 16-byte words
};
struct ChainOp {
  int kind;  // b2d_chain_op_kind
  int bn, halo, pad;
  ConvKParams conv;
  GnOp gn;
  PoolOp pool;
  AttnOp attn;
  ZeroOp zero;
};

__device__ __forceinline__ float c_bflo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float c_bfhi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

__device__ __forceinline__ void chain_grid_barrier(int* bar, int target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(bar, 1);
    uint32_t spins = 0;
    int v;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++spins > (1u << 26)) __trap();
    } while (v < target);
  }
  __syncthreads();
}

// GroupNorm apply (+SiLU, +temb, + sums of the result), 16-bit in / bf16 out, all threads of the CTA.
// Work is cut into (sample, slice) items so that one item has one (mean, rstd); coefficients sit in shared memory.
__device__ __forceinline__ void chain_gn(const GnOp& g, uint8_t* smem) {
  const int C = g.C, vpc = C >> 3, G = C / g.cpg;
  float* sa = reinterpret_cast<float*>(smem);  // [C] scale
  float* sb = sa + C;                          // [C] shift
  float* st = sb + C;                          // [C] temb
  __shared__ float red[2][16];
  const long long nvec = g.P * vpc;
  // slices per sample: enough items to occupy the grid, at least ~2K vectors each
  int spl = (int)((gridDim.x + g.N - 1) / g.N);
  const long long max_spl = nvec / 2048;
  if (spl > max_spl) spl = (int)(max_spl < 1 ? 1 : max_spl);
  const int items = g.N * spl;
  const float fold = g.act ? 0.5f : 1.f;  // silu(v) = h + h tanh(h), h = v / 2
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int n = it / spl, sl = it - n * spl;
    const double cnt = (double)g.cpg * (double)g.P;
    const float* trow = nullptr;
    if (g.temb != nullptr) trow = g.temb + (long long)(g.temb_row ? __ldg(g.temb_row + (long long)n * g.temb_row_stride) : 0) * g.temb_ld + g.temb_col;
    __syncthreads();  // previous item's readers of the tables are done
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int gi = c / g.cpg;
      const double s0 = __ldcg(g.stats + ((long long)n * G + gi) * 2), s1 = __ldcg(g.stats + ((long long)n * G + gi) * 2 + 1);
      const double mean = s0 / cnt;
      double var = s1 / cnt - mean * mean;
      if (var < 0) var = 0;
      const float rstd = (float)(1.0 / sqrt(var + (double)g.eps));
      const float ga = g.gamma ? __ldg(g.gamma + c) : 1.f, be = g.beta ? __ldg(g.beta + c) : 0.f;
      sa[c] = fold * rstd * ga;
      sb[c] = fold * (be - (float)mean * rstd * ga);
      st[c] = trow ? __ldg(trow + c) : 0.f;
    }
    __syncthreads();
    const long long lo = nvec * sl / spl, hi = nvec * (sl + 1) / spl;
    const uint4* xp = g.x + (long long)n * nvec;
    uint4* yp = g.y + (long long)n * nvec;
    float acc_s = 0.f, acc_ss = 0.f;
    auto one = [&](long long i, const uint4& u) {
      const int c0 = (int)(i % vpc) << 3;
      float f[8];
      if (g.in_f16) {
        const float2 q0 = unpack_f16(u.x), q1 = unpack_f16(u.y), q2 = unpack_f16(u.z), q3 = unpack_f16(u.w);
        f[0] = q0.x; f[1] = q0.y; f[2] = q1.x; f[3] = q1.y; f[4] = q2.x; f[5] = q2.y; f[6] = q3.x; f[7] = q3.y;
      } else {
        f[0] = c_bflo(u.x); f[1] = c_bfhi(u.x); f[2] = c_bflo(u.y); f[3] = c_bfhi(u.y);
        f[4] = c_bflo(u.z); f[5] = c_bfhi(u.z); f[6] = c_bflo(u.w); f[7] = c_bfhi(u.w);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = fmaf(f[j], sa[c0 + j], sb[c0 + j]);
        if (g.act) {
          float th;
          asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(v));
          v = fmaf(v, th, v);
        }
        v += st[c0 + j];
        f[j] = v;
        acc_s += v;
        acc_ss = fmaf(v, v, acc_ss);
      }
      __stcg(yp + i, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
    };
    const int nt = blockDim.x;
    long long i = lo + threadIdx.x;
    for (; i + 3LL * nt < hi; i += 4LL * nt) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldcg(xp + i + (long long)k * nt);
#pragma unroll
      for (int k = 0; k < 4; ++k) one(i + (long long)k * nt, u[k]);
    }
    for (; i < hi; i += nt) one(i, __ldcg(xp + i));
    if (g.stats_out != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
        acc_ss += __shfl_xor_sync(0xffffffffu, acc_ss, o);
      }
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      if (l == 0) { red[0][w] = acc_s; red[1][w] = acc_ss; }
      __syncthreads();
      if (w == 0) {
        float a = l < (int)(blockDim.x >> 5) ? red[0][l] : 0.f, b = l < (int)(blockDim.x >> 5) ? red[1][l] : 0.f;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (l == 0) {
          atomicAdd(g.stats_out + (long long)n * 2, (double)a);
          atomicAdd(g.stats_out + (long long)n * 2 + 1, (double)b);
        }
      }
    }
  }
}

// MaxPool2d(2,2) + GroupNorm(1,C) sums of the pooled map, bf16 NHWC.
__device__ __forceinline__ void chain_pool(const PoolOp& q) {
  __shared__ float red[2][16];
  const int OH = q.H >> 1, OW = q.W >> 1, vpc = q.C >> 3;
  const long long nvec = (long long)OH * OW * vpc;
  int spl = (int)((gridDim.x + q.N - 1) / q.N);
  const long long max_spl = nvec / 1024;
  if (spl > max_spl) spl = (int)(max_spl < 1 ? 1 : max_spl);
  const int items = q.N * spl;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int n = it / spl, sl = it - n * spl;
    const uint4* xi = q.x + (long long)n * q.H * q.W * vpc;
    uint4* yo = q.y + (long long)n * nvec;
    const long long lo = nvec * sl / spl, hi = nvec * (sl + 1) / spl;
    float acc_s = 0.f, acc_ss = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const int v = (int)(i % vpc);
      const long long pix = i / vpc;
      const int ox = (int)(pix % OW), oy = (int)(pix / OW);
      float m[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 u = __ldcg(xi + ((long long)(2 * oy + (k >> 1)) * q.W + (2 * ox + (k & 1))) * vpc + v);
        const float f[8] = {c_bflo(u.x), c_bfhi(u.x), c_bflo(u.y), c_bfhi(u.y), c_bflo(u.z), c_bfhi(u.z), c_bflo(u.w), c_bfhi(u.w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = k == 0 ? f[j] : fmaxf(m[j], f[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc_s += m[j]; acc_ss = fmaf(m[j], m[j], acc_ss); }
      __stcg(yo + i, make_uint4(pack_bf16(m[0], m[1]), pack_bf16(m[2], m[3]), pack_bf16(m[4], m[5]), pack_bf16(m[6], m[7])));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
      acc_ss += __shfl_xor_sync(0xffffffffu, acc_ss, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { red[0][w] = acc_s; red[1][w] = acc_ss; }
    __syncthreads();
    if (w == 0) {
      float a = l < (int)(blockDim.x >> 5) ? red[0][l] : 0.f, b = l < (int)(blockDim.x >> 5) ? red[1][l] : 0.f;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (l == 0 && q.stats != nullptr) {
        atomicAdd(q.stats + (long long)n * 2, (double)a);
        atomicAdd(q.stats + (long long)n * 2 + 1, (double)b);
      }
    }
  }
}

template <int BN, bool HALO>
__device__ __noinline__ void chain_conv(const ConvKParams& sp, const ConvKParams& gp, uint8_t* smem, uint32_t tmem_base) {
  conv_v2_layer<BN, HALO, false, true>(sp, gp, smem, tmem_base);
}

__global__ void __launch_bounds__(kChainThreads, 1) chain_kernel(const ChainOp* __restrict__ ops, int nops, int* bar) {
  extern __shared__ uint8_t chain_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(chain_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(16) uint8_t sp_raw[sizeof(ConvKParams) > sizeof(AttnTcParams) ? sizeof(ConvKParams) : sizeof(AttnTcParams)];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 1) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int G = gridDim.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    reinterpret_cast<unsigned long long*>(bar + 2)[0] = t;
  }

  for (int i = 0; i < nops; ++i) {
    const ChainOp& op = ops[i];
    const int kind = op.kind;
    if (kind == B2D_CHAIN_CONV) {
      // scalars from a shared-memory copy, TMA descriptors from the global one
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&op.conv);
      uint32_t* dst = reinterpret_cast<uint32_t*>(sp_raw);
      for (int w = threadIdx.x; w < (int)(sizeof(ConvKParams) / 4); w += blockDim.x) dst[w] = __ldg(src + w);
      __syncthreads();
      const ConvKParams& sp = *reinterpret_cast<const ConvKParams*>(sp_raw);
      const int bn = op.bn, halo = op.halo;
      if (halo) {
        if (bn == 256) chain_conv<256, true>(sp, op.conv, smem, tmem_base);
        else if (bn == 128) chain_conv<128, true>(sp, op.conv, smem, tmem_base);
        else if (bn == 64) chain_conv<64, true>(sp, op.conv, smem, tmem_base);
        else chain_conv<16, true>(sp, op.conv, smem, tmem_base);
      } else {
        if (bn == 256) chain_conv<256, false>(sp, op.conv, smem, tmem_base);
        else if (bn == 128) chain_conv<128, false>(sp, op.conv, smem, tmem_base);
        else if (bn == 64) chain_conv<64, false>(sp, op.conv, smem, tmem_base);
        else chain_conv<16, false>(sp, op.conv, smem, tmem_base);
      }
    } else if (kind == B2D_CHAIN_GN) {
      chain_gn(op.gn, smem);
    } else if (kind == B2D_CHAIN_POOL) {
      chain_pool(op.pool);
    } else if (kind == B2D_CHAIN_ATTN) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&op.attn.p);
      uint32_t* dst = reinterpret_cast<uint32_t*>(sp_raw);
      for (int w = threadIdx.x; w < (int)(sizeof(AttnTcParams) / 4); w += blockDim.x) dst[w] = __ldg(src + w);
      __syncthreads();
      const AttnTcParams& sp = *reinterpret_cast<const AttnTcParams*>(sp_raw);
      const int items = op.attn.mtiles * op.attn.N * sp.heads;
      if (threadIdx.x < kTcThreads) {
        for (int it = blockIdx.x; it < items; it += G)
          attention_tc_item<true>(sp, op.attn.p, it % op.attn.mtiles, it / op.attn.mtiles, smem, tmem_base);
      }
    } else if (kind == B2D_CHAIN_ZERO) {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < op.zero.n16; w += (long long)G * blockDim.x) op.zero.ptr[w] = z;
    }
    chain_grid_barrier(bar, (i + 1) * G);
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // per-op timestamps (ns) behind the two barrier words: tools/diag_chain.py
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      reinterpret_cast<unsigned long long*>(bar + 2)[i + 1] = t;
    }
  }

  // departures: the last CTA re-arms the barrier for the next launch
  if (threadIdx.x == 0) {
    const int old = atomicAdd(bar + 1, 1);
    if (old == G - 1) { bar[0] = 0; bar[1] = 0; __threadfence(); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b2d

using namespace b2d;

struct b2d_chain {
  std::vector<ChainOp> host;
  ChainOp* dev = nullptr;   // caller-owned device buffer (b2d_chain_bind)
  int* bar = nullptr;
  int nops_bound = 0;
};

extern "C" int b2d_chain_create(b2d_chain** out) {
  if (!out) return set_error(B2D_E_INVALID, "null argument");
  *out = new (std::nothrow) b2d_chain();
  return *out ? B2D_OK : set_error(B2D_E_INVALID, "out of host memory");
}
extern "C" int b2d_chain_destroy(b2d_chain* c) {
  delete c;
  return B2D_OK;
}
extern "C" int64_t b2d_chain_op_bytes(void) { return (int64_t)sizeof(ChainOp); }

static ChainOp* push_op(b2d_chain* c, int kind) {
  c->host.emplace_back();
  ChainOp* op = &c->host.back();
  memset(op, 0, sizeof(ChainOp));
  op->kind = kind;
  return op;
}

extern "C" int b2d_chain_add_conv(b2d_chain* c, const b2d_conv_plan* plan) {
  if (!c || !plan) return set_error(B2D_E_INVALID, "null argument");
  if (plan->engine != 2 || plan->kp.xform) return set_error(B2D_E_INVALID, "chain: only persistent-engine plans without input transform");
  ChainOp* op = push_op(c, B2D_CHAIN_CONV);
  op->conv = plan->kp;
  op->bn = plan->block_n;
  op->halo = plan->kp.halo;
  return B2D_OK;
}

extern "C" int b2d_chain_add_gn(b2d_chain* c, const void* x, void* y, int32_t N, int64_t P, int32_t C, const double* stats, int32_t cpg,
                                const float* gamma, const float* beta, float eps, int32_t act, const float* temb_table,
                                const int32_t* temb_row, int32_t temb_row_stride, int32_t temb_ld, int32_t temb_col, double* stats_out,
                                int32_t in_f16) {
  if (!c || !x || !y || !stats) return set_error(B2D_E_INVALID, "chain gn: null pointer");
  if (N < 1 || P < 1 || C < 8 || (C % 8) || cpg < 1 || (C % cpg) || 3 * C * 4 > 96 * 1024) return set_error(B2D_E_INVALID, "chain gn: bad shape");
  ChainOp* op = push_op(c, B2D_CHAIN_GN);
  GnOp& g = op->gn;
  g.x = (const uint4*)x; g.y = (uint4*)y; g.P = P; g.N = N; g.C = C; g.cpg = cpg; g.stats = stats; g.gamma = gamma; g.beta = beta;
  g.eps = eps; g.act = act; g.temb = temb_table; g.temb_row = temb_row; g.temb_row_stride = temb_row_stride; g.temb_ld = temb_ld;
  g.temb_col = temb_col; g.stats_out = stats_out; g.in_f16 = in_f16;
  return B2D_OK;
}

extern "C" int b2d_chain_add_pool(b2d_chain* c, const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, double* stats) {
  if (!c || !x || !y || N < 1 || H < 2 || W < 2 || (H & 1) || (W & 1) || C < 8 || (C % 8)) return set_error(B2D_E_INVALID, "chain pool: bad argument");
  ChainOp* op = push_op(c, B2D_CHAIN_POOL);
  op->pool.x = (const uint4*)x; op->pool.y = (uint4*)y; op->pool.N = N; op->pool.H = H; op->pool.W = W; op->pool.C = C; op->pool.stats = stats;
  return B2D_OK;
}

extern "C" int b2d_chain_add_attention(b2d_chain* c, const void* qkv, void* out, int32_t N, int32_t T, int32_t C, int32_t heads) {
  if (!c || !qkv || !out) return set_error(B2D_E_INVALID, "chain attention: null pointer");
  if (N < 1 || heads < 1 || (C % heads)) return set_error(B2D_E_INVALID, "chain attention: bad shape");
  ChainOp* op = push_op(c, B2D_CHAIN_ATTN);
  const int rc = attention_tc_params(qkv, out, N, T, C, heads, &op->attn.p);
  if (rc != B2D_OK) { c->host.pop_back(); return rc; }
  op->attn.N = N;
  op->attn.mtiles = (T + 127) / 128;
  return B2D_OK;
}

extern "C" int b2d_chain_add_zero(b2d_chain* c, void* p, int64_t bytes) {
  if (!c || !p || bytes < 0 || (bytes & 15) || (reinterpret_cast<uintptr_t>(p) & 15)) return set_error(B2D_E_INVALID, "chain zero: 16-byte granularity");
  ChainOp* op = push_op(c, B2D_CHAIN_ZERO);
  op->zero.ptr = (uint4*)p; op->zero.n16 = bytes / 16;
  return B2D_OK;
}

extern "C" int32_t b2d_chain_num_ops(const b2d_chain* c) { return c ? (int32_t)c->host.size() : 0; }

// Copy the op list into a caller-owned device buffer (>= num_ops * b2d_chain_op_bytes()) and remember the barrier words.
extern "C" int b2d_chain_bind(b2d_chain* c, void* dev_ops, int64_t dev_bytes, int32_t* barrier2, void* stream) {
  if (!c || !dev_ops || !barrier2) return set_error(B2D_E_INVALID, "null argument");
  const size_t need = c->host.size() * sizeof(ChainOp);
  if ((size_t)dev_bytes < need || (reinterpret_cast<uintptr_t>(dev_ops) & 127)) return set_error(B2D_E_INVALID, "chain bind: buffer too small / not 128-byte aligned");
  cudaError_t e = cudaMemcpyAsync(dev_ops, c->host.data(), need, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "chain bind: %s", cudaGetErrorString(e));
  e = cudaStreamSynchronize((cudaStream_t)stream);  // the host vector may be freed / changed after this call
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "chain bind: %s", cudaGetErrorString(e));
  c->dev = (ChainOp*)dev_ops;
  c->bar = barrier2;
  c->nops_bound = (int)c->host.size();
  return B2D_OK;
}

extern "C" int b2d_chain_run(const b2d_chain* c, void* stream) {
  if (!c || !c->dev) return set_error(B2D_E_INVALID, "chain not bound");
  if (c->nops_bound == 0) return B2D_OK;
  static unsigned long long configured = 0;
  const size_t smem = kChainSmem;
  {
    cudaError_t e = smem_attr_once(chain_kernel, (int)smem, configured);
    if (e != cudaSuccess) return set_error(B2D_E_CUDA, "chain smem attr: %s", cudaGetErrorString(e));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)num_sms()); cfg.blockDim = dim3(kChainThreads); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const ChainOp* ops = c->dev;
  int n = c->nops_bound;
  int* bar = c->bar;
  cudaError_t e = cudaLaunchKernelEx(&cfg, chain_kernel, ops, n, bar);
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "chain launch: %s", cudaGetErrorString(e));
  return check_launch("chain_kernel");
}
