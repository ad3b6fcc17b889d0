// Persistent implicit-GEMM convolution engine for sm_100a: TMA box loads (zero-filled halo == zero padding) ->
// 128B-swizzled smem rings -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> fused epilogue.  Planning: conv_plan.cu.
//
// One CTA per SM walks a static list of work units (output tile x N tile x K split):
//   warp 0      TMA producer   -- activation ring (A) and weight ring (B), separate mbarrier pairs
//   warp 1      MMA issuer     -- tcgen05.mma into one of TWO TMEM accumulator stages
//   warps 2..   epilogue       -- drain the other accumulator stage (tcgen05.ld) while the next
//                                 unit's main loop runs: bias, residual, GroupNorm sums, store
//
// Two ways of staging the A operand:
//   * generic: one 128-row box per (tap, 64-channel chunk), (any stride,
//     1x1 GEMMs, transposed-conv phases, small maps).
//   * halo (stride-1 3x3 / 3x3x3, W,H multiples of 16): ONE TMA box of 18 x 18 pixels x 64 channels
//     per (z-tap, chunk) serves all nine in-plane taps of a 16 x 16 output super-tile.  The tile is
//     two M = 128 halves (8 px x 16 lines); a tap (dy,dx) is just a row offset into the staged box:
//     UMMA descriptor start = box + ((1+dy)*18 + (1+dx) + 8*half) * 128 B, 8-row-group stride
//     (SBO) = 18 * 128 B.  This relies on the 128B swizzle being a function of the absolute shared
//     memory address (tools/probe/umma_probe.cu, profiles/r1_umma_swizzle_probe.txt), and cuts the
//     L2 -> SM activation traffic 9x tap re-reads -> 1.27x (324 staged rows per 256 outputs).
//
// Split-K (deep UNet levels: M of a few hundred rows, K up to 18432): every split writes its fp32
// partial tile to a workspace; the last split to arrive (atomic ticket, no spinning) sums all
// partials in a fixed order -- deterministic -- and runs the fused epilogue.
#include "conv_v2.cuh"

namespace b2d {

template <int BN, bool HALO, bool XFORM>
__global__ void __launch_bounds__(V2Cfg<BN, HALO, XFORM>::THREADS, 1) conv_v2_kernel(const __grid_constant__ ConvKParams p) {
  extern __shared__ uint8_t smem_raw2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~uintptr_t(1023));
  conv_v2_layer<BN, HALO, XFORM>(p, smem);
}

// CTA-pair variant (tcgen05 cta_group::2): launched as clusters of two, generic staging
template <int BN>
__global__ void __launch_bounds__(V2Cfg<BN, false, false, true>::THREADS, 1) conv_pair_kernel(const __grid_constant__ ConvKParams p) {
  extern __shared__ uint8_t smem_raw3[];
  // both CTAs of a pair must lay their stages out at the same shared-memory offsets: the dynamic segment starts at the
  // same offset in every CTA of a kernel, so the same rounding gives the same offsets
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw3) + 1023) & ~uintptr_t(1023));
  conv_v2_layer<BN, false, false, true>(p, smem);
}

template <int BN>
static int launch_pair(const ConvKParams& kp, dim3 grid, cudaStream_t st) {
  using Cfg = V2Cfg<BN, false, false, true>;
  static unsigned long long configured = 0;
  cudaError_t e = smem_attr_once(conv_pair_kernel<BN>, Cfg::SMEM, configured);
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "cudaFuncSetAttribute(pair, smem=%d): %s", Cfg::SMEM, cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = (size_t)Cfg::SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, conv_pair_kernel<BN>, kp);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "conv_pair launch: %s", cudaGetErrorString(e));
  return B2D_OK;
}

template <int BN, bool HALO, bool XFORM = false>
static int launch_v2(const ConvKParams& kp, dim3 grid, cudaStream_t st) {
  using Cfg = V2Cfg<BN, HALO, XFORM>;
  static unsigned long long configured = 0;
  cudaError_t e = smem_attr_once(conv_v2_kernel<BN, HALO, XFORM>, Cfg::SMEM, configured);
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "cudaFuncSetAttribute(v2, smem=%d): %s", Cfg::SMEM, cudaGetErrorString(e));
  conv_v2_kernel<BN, HALO, XFORM><<<grid, dim3(Cfg::THREADS), (size_t)Cfg::SMEM, st>>>(kp);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "conv_v2 launch: %s", cudaGetErrorString(e));
  return B2D_OK;
}

int launch_conv_v2(const b2d_conv_plan* plan, cudaStream_t st) {
  const ConvKParams& kp = plan->kp;
  if (kp.pair) {
    switch (plan->block_n) {
      case 128: return launch_pair<128>(kp, plan->grid, st);
      case 256: return launch_pair<256>(kp, plan->grid, st);
    }
    return set_error(B2D_E_INVALID, "conv pair: bad block_n %d", plan->block_n);
  }
  if (kp.halo && kp.xform) {
    switch (plan->block_n) {
      case 16: return launch_v2<16, true, true>(kp, plan->grid, st);
      case 64: return launch_v2<64, true, true>(kp, plan->grid, st);
      case 128: return launch_v2<128, true, true>(kp, plan->grid, st);
      case 256: return launch_v2<256, true, true>(kp, plan->grid, st);
    }
  } else if (kp.halo) {
    switch (plan->block_n) {
      case 16: return launch_v2<16, true>(kp, plan->grid, st);
      case 64: return launch_v2<64, true>(kp, plan->grid, st);
      case 128: return launch_v2<128, true>(kp, plan->grid, st);
      case 256: return launch_v2<256, true>(kp, plan->grid, st);
    }
  } else {
    switch (plan->block_n) {
      case 16: return launch_v2<16, false>(kp, plan->grid, st);
      case 64: return launch_v2<64, false>(kp, plan->grid, st);
      case 128: return launch_v2<128, false>(kp, plan->grid, st);
      case 256: return launch_v2<256, false>(kp, plan->grid, st);
    }
  }
  return set_error(B2D_E_INVALID, "conv v2: bad block_n %d (halo %d)", plan->block_n, kp.halo);
}

}  // namespace b2d

#ifdef B2D_TIMELINE
// debug library only (tools/timeline_conv.py): clear / read the per-CTA stamps
extern "C" __attribute__((visibility("default"))) int b2d_debug_timeline(unsigned long long* host, int clear) {
  if (clear == -1)  // CTA 0's per-iteration log
    return cudaMemcpyFromSymbol(host, b2d::g_tl_iter, sizeof(b2d::g_tl_iter)) == cudaSuccess ? 0 : -1;
  if (clear >= 16) {  // 16 + ablation mode
    const int mode = clear - 16;
    return cudaMemcpyToSymbol(b2d::g_tl_mode, &mode, sizeof(int)) == cudaSuccess ? 0 : -1;
  }
  if (clear) {
    void* d = nullptr;
    if (cudaGetSymbolAddress(&d, b2d::g_timeline) != cudaSuccess) return -1;
    return cudaMemset(d, 0, sizeof(b2d::g_timeline)) == cudaSuccess ? 0 : -1;
  }
  return cudaMemcpyFromSymbol(host, b2d::g_timeline, sizeof(b2d::g_timeline)) == cudaSuccess ? 0 : -1;
}
#endif
