// Planning of the implicit-GEMM convolution engine for sm_100a (kernel: conv_v2.cuh / conv_engine.cu): validates a
// b2d_conv_desc, picks the M tiling (halo super-tiles or generic 128-row boxes), BLOCK_N and the K split by a small cost
// model, encodes the TMA descriptors and records the launch geometry.
//
// Replaces (reference file:line): nn.Conv2d unet/blocks.py:29-36, models.py:120-128;
// nn.ConvTranspose2d unet/blocks.py:128-133; nn.Conv3d vae/blocks.py:155-169, encoder.py:30-68,
// decoder.py:31-71; nn.Linear/Conv1d projections of unet/blocks.py:196-207.
#include <stdlib.h>
#include <string.h>
#include <new>
#include <type_traits>

#include "conv_common.cuh"

namespace b2d {

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

}  // namespace b2d

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
  if (!d->weight || (!d->out && d->out_mode != 3)) return set_error(B2D_E_INVALID, "null weight/out");
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

  if (d->tune_ksplit < 0 || d->tune_ksplit > 16) return set_error(B2D_E_INVALID, "tune_ksplit=%d (0 auto, 1..16)", d->tune_ksplit);
  // halo staging: canonical 3x3 / 3x3x3 'same' taps in (z, y, x) order, stride 1, 16x16 super-tiles, BLOCK_N <= 128
  bool halo = false;
  int gt = 0;  // in-plane taps per z group (taps are z-major): 9 = full 3x3, 4 = 2x2 of an upsample-folded conv
  if (d->stride_h == 1 && d->stride_w == 1 && d->nphase == 1 && d->ntaps >= 4 &&
      d->OW % 16 == 0 && d->OH % 16 == 0 && d->OW == d->W && d->OH == d->H && (d->cout <= 16 || d->cout % 64 == 0) &&
      (d->block_n == 0 || d->block_n == 16 || d->block_n == 64 || d->block_n == 128 || d->block_n == 256)) {
    const bool no_halo = (d->tune_flags & B2D_TUNE_NO_HALO) != 0;
    while (gt < d->ntaps && d->tap_dz[gt] == d->tap_dz[0]) ++gt;
    halo = !no_halo && (gt == 9 || gt == 4) && d->ntaps % gt == 0;
    for (int t = 0; t < d->ntaps && halo; ++t) {
      // every z group repeats the in-plane offsets of the first one, all within the 1-pixel halo
      const int j = t % gt;
      if (d->tap_dy[t] != d->tap_dy[j] || d->tap_dx[t] != d->tap_dx[j] || d->tap_dz[t] != d->tap_dz[(t / gt) * gt] ||
          d->tap_dy[t] < -1 || d->tap_dy[t] > 1 || d->tap_dx[t] < -1 || d->tap_dx[t] > 1)
        halo = false;
    }
  }
  long long tiles_m2 = tiles_m;
  if (halo) {
    lw = 4; lh = 4; ld = 0; ln = 0;
    tiles_m2 = (long long)(d->OW / 16) * (d->OH / 16) * d->D * d->N;
  }

  // ---- persistent engine: pick (BLOCK_N, K splits) by a small cost model -----------------------
  // cost = waves over the SMs x (K-loop time of one unit + epilogue / split-K fix-up), in SM clocks.
  int bn = d->block_n;
  int ksplit_pick = 1;
  bool streamk_pick = false, pair_pick = false;
  {
    const int sms = num_sms();
    const long long cols = (long long)d->cout * d->nphase;
    long long ngroups = 0;
    for (int s = 0; s < d->nseg; ++s) ngroups += (long long)(halo ? d->ntaps / gt : d->ntaps) * (d->cin[s] / kBlockK);
    bool all_dz0 = true;
    for (int t = 0; t < d->ntaps; ++t) all_dz0 = all_dz0 && d->tap_dz[t] == 0;
    const bool no_split = (d->tune_flags & B2D_TUNE_NO_SPLITK) != 0;
    const bool can_split = !no_split && all_dz0 && d->out_mode != 3 && d->workspace && (reinterpret_cast<uintptr_t>(d->workspace) & 15) == 0;
    const int mt = halo ? 2 : 1;
    const int cand[4] = {256, 128, 64, 16};
    const int force_ks = d->tune_ksplit;  // tools/tune_conv.py: measure a given split count (0 = cost model)
    double best = 1e30, best_sk = 1e30;
    int best_bn = 0, best_sk_bn = 0;
    const bool sk_off = (d->tune_flags & B2D_TUNE_NO_STREAMK) != 0, sk_force = (d->tune_flags & B2D_TUNE_STREAMK) != 0;
    const bool pair_off = (d->tune_flags & B2D_TUNE_NO_PAIR) != 0, pair_force = (d->tune_flags & B2D_TUNE_PAIR) != 0;
    double best_pair = 1e30;
    int best_pair_bn = 0, best_pair_ks = 1;
    for (int ci = 0; ci < 4; ++ci) {
      const int b = cand[ci];
      if (d->block_n && d->block_n != b) continue;
      if (b == 16 ? (d->cout > 16 || d->nphase != 1) : (d->cout % b != 0)) continue;
      if (halo && gt == 4 && b < 128) continue;  // narrow tiles batch 3 / 9 taps per weight stage
      const long long tiles = tiles_m2 * (b == 16 ? 1 : cols / b);
      const double t_kb = b == 256 ? 512.0 : b == 128 ? 256.0 : b == 64 ? 192.0 : 128.0;
      // one A-operand group: tensor time of its MMAs, bounded below by what the role warps can turn around.
      //  * generic tiles: a K step costs the producer / MMA warps ~350-390 clocks of loop instructions and barrier round
      //    trips whatever the tile width (tools/timeline_conv.py, profiles/r2_timeline_conv.txt: 525 clocks per step at
      //    BLOCK_N = 256 = its four MMAs; ~390 at BLOCK_N = 128 against 256 clocks of MMAs).  Until the lean K loop of
      //    conv_v2.cuh the same step took 890-960 clocks at every width, which an earlier version of this model had
      //    fitted as "operand bytes at 40 B/clk".
      //  * halo tiles: operand bytes at 40 B/clk per SM (fitted on the 128-wide VAE tiles, which sit at that bound)
      constexpr double kL2BytesPerClk = 40.0, kGenericStepFloor = 380.0;
      const double group_bytes = halo ? 18.0 * 18.0 * 128.0 + gt * b * 128.0 : (128.0 + b) * 128.0;
      double per_group = (halo ? 2.0 * gt : 1.0) * t_kb;
      if (halo) {
        if (per_group < group_bytes / kL2BytesPerClk) per_group = group_bytes / kL2BytesPerClk;
      } else if (per_group < kGenericStepFloor) {
        per_group = kGenericStepFloor;
      }
      if (d->in_stats && per_group < 6000.0) per_group = 6000.0;  // fused input normalisation: the tile rewrite bounds a group
      for (int ks = 1; ks <= 16; ++ks) {
        if (force_ks > 0 && ks < force_ks) continue;  // tools/tune_conv.py: measure a given split count
        if (force_ks > 0 && ks > force_ks) break;
        if (ks > 1 && (!can_split || ngroups / ks < 4 || tiles * 2 > 4096 || tiles >= 2LL * sms ||
                       16384 + tiles * ks * mt * 128LL * b * 4 > d->workspace_bytes)) break;
        // split-K fix-up: park the fp32 partial (coalesced), fence + ticket, and the last arriver re-reads ks partials
        const double fix = ks > 1 ? 6000.0 + (1.0 + ks) * 12.0 * b * mt : 0.0;
        const double unit = (double)((ngroups + ks - 1) / ks) * per_group + 1500.0 + 8.0 * b * mt + fix;
        const double waves = (double)((tiles * ks + sms - 1) / sms);
        const double cost = waves * unit;
        if (cost < best * 0.999) { best = cost; best_bn = b; ksplit_pick = ks; streamk_pick = false; }
      }
      // CTA pairs (tcgen05 cta_group::2, generic staging): 256-row M super-tiles, each SM stages its 128 A rows and HALF of
      // the weight rows -- (128 + b / 2) x 128 B per K block instead of (128 + b) x 128 B through the 40 B/clk L2 port
      if (!halo && !pair_off && (b == 256 || b == 128) && d->nphase == 1 && !d->in_stats && tiles_m2 >= 2) {
        const long long pairs_m = (tiles_m2 + 1) / 2, ptiles = pairs_m * (cols / b);
        double pg = t_kb;
        const double pbytes = (128.0 + b / 2) * 128.0;
        if (pg < pbytes / kL2BytesPerClk) pg = pbytes / kL2BytesPerClk;
        const int clusters = sms / 2;
        for (int ks = 1; ks <= 16; ++ks) {
          if (force_ks > 0 && ks < force_ks) continue;
          if (force_ks > 0 && ks > force_ks) break;
          if (2 * ptiles * 2 > 4096) break;  // arrival counters: one per (tile, column group)
          if (ks > 1 && (!can_split || ngroups / ks < 4 || ptiles >= 2LL * clusters ||
                         16384 + 2 * ptiles * ks * 128LL * b * 4 > d->workspace_bytes)) break;
          const double fix = ks > 1 ? 6000.0 + (1.0 + ks) * 12.0 * b : 0.0;
          const double unit = (double)((ngroups + ks - 1) / ks) * pg + 2500.0 + 8.0 * b + fix;  // + the cluster handshakes
          const double waves = (double)((ptiles * ks + clusters - 1) / clusters);
          const double cost = waves * unit;
          if (cost < best_pair) { best_pair = cost; best_pair_bn = b; best_pair_ks = ks; }
        }
      }
      // stream-K: every CTA runs the same number of B-stage steps of the flat (tile, step) space; tiles cut by a range
      // boundary go through the workspace (at most two partial tiles parked per CTA)
      if (can_split && !sk_off && force_ks == 0) {
        const int spg = halo ? gt / (b <= 16 ? 9 : b <= 64 ? 3 : 1) : 1;
        const long long S = ngroups * spg, total = tiles * S;
        const long long G = total < sms ? total : sms;
        const long long per_cta = (total + G - 1) / G;
        const long long pieces = total / G > 0 ? S / (total / G) + 2 : 99;  // bound on the CTAs that can share one tile
        if (spg >= 1 && S >= 8 && total >= 2LL * sms && tiles * mt * 2 <= 4096 && total * G < (1LL << 31) && pieces <= kMaxPieces &&
            16384 + G * 2 * mt * 128LL * b * 4 <= d->workspace_bytes) {
          const double per_step = per_group / spg;
          const double items = (double)tiles / (double)G + 1.0;  // epilogues or partial dumps per CTA
          const double shared = pieces > 2 ? 3.0 : 2.0;          // partial tiles the reducing piece re-reads
          const double cost = (double)per_cta * per_step + 1500.0 + items * 8.0 * b * mt + 6000.0 + (1.0 + shared) * 12.0 * b * mt;
          if (cost < best_sk) { best_sk = cost; best_sk_bn = b; }
        }
      }
    }
    // stream-K only where forced (B2D_TUNE_STREAMK): measured against every (BLOCK_N, K split) choice on the UNet's
    // layer shapes at 88 slice-images it wins nowhere by more than 1 % (profiles/r2_tune_conv_88.txt) -- parking and
    // re-reading 128-256 KB partial tiles costs more than the wave quantisation it removes, and where K is long enough
    // to amortise that, 3-way split-K fills the machine just as well
    if (best_sk_bn != 0 && (sk_force || best_bn == 0)) {
      best = best_sk; best_bn = best_sk_bn; ksplit_pick = 1; streamk_pick = true;
    }
    // CTA pairs only where forced (B2D_TUNE_PAIR): parity-green, 25 % less L2 -> SM traffic in ncu
    // (profiles/r2_deep_level_ncu.txt) and the same time to within 2 % on every layer at 88 and at 704 slice-images
    // (profiles/r2_tune_conv_88_pairs.txt, r2_tune_conv_704_pairs.txt): these layers are not bound by operand bytes per SM
    if (best_pair_bn != 0 && !streamk_pick && pair_force) {
      best = best_pair; best_bn = best_pair_bn; ksplit_pick = best_pair_ks; pair_pick = true;
    }
    if (best_bn == 0)
      return set_error(B2D_E_INVALID, "cout=%d block_n=%d%s: need cout a multiple of 64 (or <= 16)", d->cout, d->block_n,
                       force_ks > 0 ? " (tune_ksplit not applicable to this layer)" : "");
    bn = best_bn;
  }
  if (bn != 16 && bn != 64 && bn != 128 && bn != 256) return set_error(B2D_E_INVALID, "block_n=%d unsupported", bn);
  if (bn >= 64 && d->cout % bn) return set_error(B2D_E_INVALID, "cout=%d not a multiple of block_n=%d", d->cout, bn);
  if (bn == 16 && (d->cout > 16 || d->nphase != 1)) return set_error(B2D_E_INVALID, "block_n=16 needs cout<=16");
  const int total_cols = (bn == 16) ? 16 : d->cout * d->nphase;
  if (d->wrows < total_cols) return set_error(B2D_E_INVALID, "wrows=%d < %d", d->wrows, total_cols);
  if (d->out_mode < 0 || d->out_mode > 3) return set_error(B2D_E_INVALID, "out_mode");
  if (d->out_mode == 3) {
    if (bn != 16 || d->cout > 16 || (d->cout % 4) || d->nphase != 1 || d->stride_h != 1 || d->stride_w != 1)
      return set_error(B2D_E_INVALID, "out_mode 3 (fused sampler update) needs cout in {4,8,12,16}, stride 1, one phase (cout=%d block_n=%d)", d->cout, bn);
    if (!d->sched_x || !d->sched_coef || (d->sched_kind != 0 && d->sched_kind != 1))
      return set_error(B2D_E_INVALID, "out_mode 3: sched_x / sched_coef / sched_kind");
    if (d->sched_step_idx && d->sched_step_inc != 0 && !d->sched_ticket)
      return set_error(B2D_E_INVALID, "out_mode 3: sched_step_inc != 0 needs a caller-owned, zero-initialised sched_ticket");
    if ((reinterpret_cast<uintptr_t>(d->sched_x) | reinterpret_cast<uintptr_t>(d->sched_noise)) & 15)
      return set_error(B2D_E_INVALID, "out_mode 3: sched_x / sched_noise must be 16-byte aligned");
    if (d->sched_x_bf16 && ((d->sched_bf16_stride % 4) || d->sched_bf16_stride < d->cout ||
                            ((reinterpret_cast<uintptr_t>(d->sched_x_bf16) | reinterpret_cast<uintptr_t>(d->sched_x_bf16_lo)) & 7)))
      return set_error(B2D_E_INVALID, "out_mode 3: bf16 copy needs a channel stride that is a multiple of 4 and 8-byte alignment");
    if (d->out_H != d->OH || d->out_W != d->OW || d->out_sy != 1 || d->out_sx != 1 || d->out_oy != 0 || d->out_ox != 0)
      return set_error(B2D_E_INVALID, "out_mode 3: the latent has the conv's own output geometry");
    if (d->out && ((d->out_cstride % 4) || (d->out_coff % 4) || (reinterpret_cast<uintptr_t>(d->out) & 15)))
      return set_error(B2D_E_INVALID, "out_mode 3: eps output needs cstride/coff multiples of 4");
  }
  if (d->out_mode == 0 && ((d->out_cstride % 8) || (d->out_coff % 8) || (reinterpret_cast<uintptr_t>(d->out) & 15)))
    return set_error(B2D_E_INVALID, "bf16 output needs cstride/coff multiples of 8 and 16-byte alignment");
  if (d->out_mode == 2 && ((d->out_cstride % 4) || (d->out_coff % 4) || (reinterpret_cast<uintptr_t>(d->out) & 15)))
    return set_error(B2D_E_INVALID, "fp32 NDHWC output needs cstride/coff multiples of 4");
  if (d->residual && (d->out_mode != 0 && d->out_mode != 2)) return set_error(B2D_E_INVALID, "residual needs out_mode 0/2");
  if (d->residual && ((d->res_cstride % 8) || (reinterpret_cast<uintptr_t>(d->residual) & 15) || bn == 16))
    return set_error(B2D_E_INVALID, "residual needs cstride multiple of 8, 16-byte alignment, block_n>=64");
  if (d->op_f16 && (d->out_lo || d->residual_lo || d->sched_x_bf16_lo))
    return set_error(B2D_E_INVALID, "op_f16 (fp16 operands) excludes the bf16 hi/lo split outputs");
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
    if (halo) { box[1] = 18; box[2] = 18; box[3] = 1; box[4] = 1; }
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
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)(pair_pick ? bn / 2 : bn)};  // a CTA pair stages half of the rows each
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
  k.op_f16 = d->op_f16 ? 1 : 0;
  if (d->out_mode == 3) {
    k.sch_x = d->sched_x; k.sch_noise = d->sched_noise; k.sch_coef = d->sched_coef;
    k.sch_step = d->sched_step_idx; k.sch_ticket = d->sched_ticket;
    k.sch_seed_dev = reinterpret_cast<const unsigned long long*>(d->sched_seed_dev);
    k.sch_seed = (unsigned long long)d->sched_seed;
    k.sch_kind = d->sched_kind; k.sch_step_off = d->sched_step_off; k.sch_step_inc = d->sched_step_inc;
    k.sch_clip = d->sched_clip ? 1 : 0; k.sch_lo = d->sched_clip_lo; k.sch_hi = d->sched_clip_hi;
    k.sch_bf16 = reinterpret_cast<__nv_bfloat16*>(d->sched_x_bf16);
    k.sch_bf16_lo = reinterpret_cast<__nv_bfloat16*>(d->sched_x_bf16_lo);
    k.sch_bf16_stride = d->sched_bf16_stride;
  }
  if (d->in_stats) {
    if (!halo || d->nseg != 1 || d->cin[0] > 512 || d->in_cpg < 1 || d->in_creal < 1 || d->in_creal > d->cin[0] || d->in_creal % d->in_cpg) {
      delete pl;
      return set_error(B2D_E_INVALID, "fused input normalisation needs the persistent halo engine, one segment, cin <= 512 "
                       "(halo %d nseg %d cin %d cpg %d creal %d)", (int)halo, d->nseg, d->cin[0], d->in_cpg, d->in_creal);
    }
    k.xform = 1;
    k.in_stats = d->in_stats; k.in_gamma = d->in_gamma; k.in_beta = d->in_beta;
    k.in_cpg = d->in_cpg; k.in_creal = d->in_creal; k.in_f16 = d->in_f16 ? 1 : 0; k.in_act = d->in_act ? 1 : 0;
    k.in_eps = d->in_eps;
    k.in_count = (double)d->in_cpg * d->D * d->H * d->W;
    k.in_temb = d->in_temb; k.in_temb_row = d->in_temb_row;
    k.in_temb_row_stride = d->in_temb_row_stride; k.in_temb_ncols = d->in_temb_ncols; k.in_temb_col = d->in_temb_col;
    if (k.in_temb && (!k.in_temb_row || k.in_temb_ncols < k.in_temb_col + d->in_creal || k.in_temb_col < 0)) {
      delete pl;
      return set_error(B2D_E_INVALID, "fused input normalisation: bad time-embedding arguments");
    }
  }

  pl->block_n = bn;
  pl->grid = dim3((unsigned)(k.tiles_w * k.tiles_h * k.tiles_d * k.tiles_n), (unsigned)(total_cols / bn), 1);
  int kb = 0;
  for (int s = 0; s < d->nseg; ++s) kb += d->ntaps * k.cchunks[s];
  pl->kblocks = kb;
  pl->ws_bytes = 0;
  {
    // ---- persistent engine: work units, K-loop groups, split-K ---------------------------------
    k.halo = halo ? 1 : 0;
    k.gtaps = halo ? gt : 1;
    const int tgroups = halo ? d->ntaps / gt : d->ntaps;  // A loads per (segment, chunk)
    k.goff[0] = 0;
    for (int s = 0; s < d->nseg; ++s) k.goff[s + 1] = k.goff[s] + tgroups * k.cchunks[s];
    k.ngroups = k.goff[d->nseg];
    k.tiles_ncol = total_cols / bn;
    const long long tiles = (long long)k.tiles_w * k.tiles_h * k.tiles_d * k.tiles_n * k.tiles_ncol;
    if (tiles > 0x3fffffff) { delete pl; return set_error(B2D_E_INVALID, "too many tiles"); }
    const int sms = num_sms();
    const int mt = halo ? 2 : 1;
    const int ksplit = ksplit_pick;
    constexpr long long kCounterBytes = 16384;
    k.ksplit = ksplit;
    k.spg = halo ? gt / (bn <= 16 ? 9 : bn <= 64 ? 3 : 1) : 1;  // must match V2Cfg::TPB
    k.ksteps = k.ngroups * k.spg;
    k.streamk = streamk_pick ? 1 : 0;
    k.total_steps = streamk_pick ? (int)(tiles * k.ksteps) : 0;
    k.fd_ksteps.set((uint32_t)k.ksteps);
    k.fd_total.set((uint32_t)(streamk_pick ? k.total_steps : 1));
    k.num_units = (int)(tiles * ksplit);
    k.pair = pair_pick ? 1 : 0;
    long long ws_tiles = tiles;
    if (pair_pick) {
      // the walk runs over pair tiles; rank r of a cluster takes M tile 2 * pair + r (an odd last tile is all padding:
      // TMA zero fill, rows masked in the epilogue).  Workspace / counters are indexed by the padded M tile count.
      const long long tiles_m = (long long)k.tiles_w * k.tiles_h * k.tiles_d * k.tiles_n;
      const long long pairs_m = (tiles_m + 1) / 2;
      k.num_units = (int)(pairs_m * k.tiles_ncol * ksplit);
      ws_tiles = 2 * pairs_m * k.tiles_ncol;
    }
    k.fd_ksplit.set((uint32_t)ksplit); k.fd_ncol.set((uint32_t)k.tiles_ncol);
    k.fd_w.set((uint32_t)k.tiles_w); k.fd_h.set((uint32_t)k.tiles_h); k.fd_d.set((uint32_t)k.tiles_d);
    if ((long long)k.ngroups * (ksplit + 1) >= (1LL << 31)) { delete pl; return set_error(B2D_E_INVALID, "K loop too long"); }
    if (ksplit > 1) {
      k.counters = reinterpret_cast<int*>(d->workspace);
      k.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d->workspace) + kCounterBytes);
      pl->ws_bytes = kCounterBytes + ws_tiles * ksplit * mt * 128LL * bn * 4;
    }
    pl->grid = dim3((unsigned)(k.num_units < sms ? k.num_units : sms), 1, 1);
    if (pair_pick) pl->grid = dim3(2u * (unsigned)(k.num_units < sms / 2 ? k.num_units : sms / 2), 1, 1);
    if (streamk_pick) {
      pl->grid = dim3((unsigned)(k.total_steps < sms ? k.total_steps : sms), 1, 1);
      k.counters = reinterpret_cast<int*>(d->workspace);
      k.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d->workspace) + kCounterBytes);
      pl->ws_bytes = kCounterBytes + (long long)pl->grid.x * 2 * mt * 128LL * bn * 4;
    }
    // a strided walk changes sample at every unit when a sample has fewer units than the grid has CTAs (UNet levels):
    // every change costs a GroupNorm flush (two epilogue barriers + global atomics); walk contiguous ranges instead
    const int env_contig = (d->tune_flags & B2D_TUNE_CONTIG) ? 1 : (d->tune_flags & B2D_TUNE_STRIDED) ? 0 : -1;
    const long long units_per_n = k.tiles_n > 0 ? k.num_units / k.tiles_n : k.num_units;
    k.contig = env_contig >= 0 ? env_contig : (units_per_n < (long long)pl->grid.x && k.num_units > (int)pl->grid.x) ? 1 : 0;
    if (pair_pick) k.contig = 0;
  }
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
  return launch_conv_v2(plan, st);
}

extern "C" int b2d_conv_plan_info2(const b2d_conv_plan* plan, int32_t* out8) {
  if (!plan || !out8) return set_error(B2D_E_INVALID, "null argument");
  out8[0] = plan->kp.pair ? 3 : 2;  // 2: one CTA per tile; 3: tcgen05 cta_group::2 CTA pairs on 256-row super-tiles
  out8[1] = plan->kp.halo;
  out8[2] = plan->kp.streamk ? -1 : plan->kp.ksplit;  // -1: stream-K
  out8[3] = plan->kp.num_units;
  out8[4] = (int32_t)(plan->grid.x * plan->grid.y);
  out8[5] = plan->block_n;
  out8[6] = plan->kp.ngroups;
  out8[7] = (int32_t)((plan->ws_bytes + 1023) / 1024);
  return B2D_OK;
}
