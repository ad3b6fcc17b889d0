// Device-side body of the persistent implicit-GEMM conv engine (see conv_engine.cu for the design notes).
#pragma once
#include "conv_common.cuh"
#include "conv_work.cuh"


namespace b2d {

// Debug timeline (tools/timeline_conv.py builds a separate library with -DB2D_TIMELINE; the shipped libb2d.so has none of
// it): per CTA, {globaltimer ns, clock64} at a few points of the unit a CTA runs last.
#ifdef B2D_TIMELINE
constexpr int kTlSlots = 16;
__device__ unsigned long long g_timeline[296 * kTlSlots * 2];
__device__ __forceinline__ void tl_stamp(int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_timeline[((size_t)blockIdx.x * kTlSlots + slot) * 2] = t;
  g_timeline[((size_t)blockIdx.x * kTlSlots + slot) * 2 + 1] = (unsigned long long)clock64();
}
#define B2D_TL(slot, pred) do { if (pred) tl_stamp(slot); } while (0)
// ablation switches of the debug build (generic staging): 1 = issue no MMAs, 2 = load no A tiles, 4 = load no B tiles
__device__ int g_tl_mode;
#define B2D_TL_MODE() (g_tl_mode)
// CTA 0's first 128 K iterations (generic staging): clock64 when the producer holds both free stages, when the MMA warp
// sees both full barriers, and after it has issued the iteration's commits
__device__ long long g_tl_iter[3][128];
#define B2D_TL_ITER(row, n, pred) do { if ((pred) && blockIdx.x == 0 && (n) < 128) g_tl_iter[row][n] = clock64(); } while (0)
#else
#define B2D_TL(slot, pred) do { } while (0)
#define B2D_TL_MODE() 0
#define B2D_TL_ITER(row, n, pred) do { } while (0)
#endif

// PAIR: two CTAs of a cluster form a tcgen05 cta_group::2 pair on a 256-row M super-tile (generic staging only): each
// stages its own 128 A rows and HALF of the N tile's weight rows, the pair's even CTA issues M = 256 MMAs that read both
// halves, each CTA's TMEM receives its own 128 x BN accumulator.  Per SM the operand bytes pulled from L2 per K block
// drop from (128 + BN) x 128 B to (128 + BN / 2) x 128 B -- the bound of the small-map UNet layers (conv_plan.cu).
template <int BN, bool HALO, bool XFORM = false, bool PAIR = false>
struct V2Cfg {
  static_assert(HALO || !XFORM, "input transform needs halo staging");
  static_assert(!PAIR || (!HALO && !XFORM && BN >= 128), "CTA pairs: generic staging, 128- or 256-wide N tiles");
  static constexpr int MT = HALO ? 2 : 1;                       // M = 128 halves per unit
  static constexpr int A_TILE = HALO ? 18 * 18 * 128 : kABytes;  // bytes landed per A load
  static constexpr int A_STAGE = (A_TILE + 1023) / 1024 * 1024;
  // weight taps per B stage: narrow N tiles batch several taps behind one mbarrier so the single
  // MMA-issuing thread is not handshake-bound (an N = 16 MMA lasts ~32 clocks)
  static constexpr int TPB = HALO ? (BN <= 16 ? 9 : BN <= 64 ? 3 : 1) : 1;
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;              // weight rows this CTA stages
  static constexpr int B_TILE = B_ROWS * kBlockK * 2;
  static constexpr int B_STAGE = TPB * B_TILE;
  static constexpr int NA = HALO ? ((BN >= 128 || XFORM) ? 2 : 3) : PAIR ? (BN == 256 ? 6 : 8) : (BN == 256 ? 4 : BN == 128 ? 6 : 8);
  static constexpr int NB = HALO ? (BN == 256 ? 4 : BN == 128 ? 6 : BN == 64 ? 4 : 3) : NA;
  static constexpr int BNC = BN < 32 ? 32 : BN;                  // TMEM columns of one M half
  static constexpr int ACC_COLS = MT * BNC;                      // one accumulator stage
  static constexpr int NACC = 2 * ACC_COLS <= 512 ? 2 : 1;       // BN = 256 halo: 512 columns, single-buffered
  static constexpr int TMEM_COLS = NACC * ACC_COLS <= 32 ? 32 : NACC * ACC_COLS <= 64 ? 64 : NACC * ACC_COLS <= 128 ? 128
                                   : NACC * ACC_COLS <= 256 ? 256 : 512;
  // generic tiles: two four-warp epilogue groups split the tile's columns (two warps per scheduler hide each other's
  // TMEM-load / store latencies; thin-K layers are epilogue-bound)
  static constexpr int NCG = (!HALO && BN >= 64) ? 2 : 1;
  static constexpr int BNG = BN / NCG;                            // columns one epilogue group handles
  static constexpr int EPI_WARPS = 4 * MT * NCG;
  static constexpr int XF_WARPS = XFORM ? 4 : 0;                 // warps rewriting staged A tiles (GroupNorm + SiLU)
  static constexpr int XF_MAXC = 512;                            // input channels the coefficient table holds
  static constexpr int BP_WARP = HALO ? 2 + EPI_WARPS + XF_WARPS : -1;  // halo: the weight ring has its own producer warp
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32 * XF_WARPS + (HALO ? 32 : 0);
  static constexpr int NBARS = 3 * NA + 2 * NB + 4;
  static constexpr int STAGE_BYTES = HALO ? 0 : 2048 * EPI_WARPS; // generic tiles: per-warp store staging (coalesced 16-bit stores)
  static constexpr int BIAS_BYTES = MT * BN * 4;                // per four-warp epilogue group: the tile's bias
  static constexpr int SMEM = NA * A_STAGE + NB * B_STAGE + NBARS * 8 + 64 + 512 * EPI_WARPS + (XFORM ? 3 * XF_MAXC * 4 + 16 : 0) + BIAS_BYTES + STAGE_BYTES + 16 + 1024;
  static_assert(NACC * ACC_COLS <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(!HALO || 9 % TPB == 0, "taps per stage must divide 9");
};

// One planned convolution executed by the whole CTA (`p` lives in the kernel's parameter space: its TMA descriptors are
// addressed in place).
// the work items of this CTA: the unit / stream-K walk of conv_work.cuh; in PAIR mode the walk runs over PAIR tiles
// (clusters instead of CTAs) and rank r takes M tile 2 * pair + r of the item
template <bool PAIR>
struct CtaWalk {
  WorkIter wi;
  int rank;
  __device__ __forceinline__ void init(const ConvKParams& p) {
    if constexpr (PAIR) {
      rank = (int)cluster_ctarank();
      wi.init(p, (int)(blockIdx.x >> 1), (int)(gridDim.x >> 1));
    } else {
      rank = 0;
      wi.init(p, (int)blockIdx.x, (int)gridDim.x);
    }
  }
  __device__ __forceinline__ bool next(const ConvKParams& p, UnitCoord& uc) {
    if (!wi.next(p, uc)) return false;
    if constexpr (PAIR) {
      const int ncol = p.tiles_ncol;
      const int pm = (uc.tile - uc.gcol0) / ncol;  // pair index along M (uc.tile = pm * ncol + n_tile)
      const int ks = uc.ks, k_lo = uc.k_lo, k_hi = uc.k_hi;
      decode_tile(p, (2 * pm + rank) * ncol + uc.gcol0, uc);
      uc.ks = ks; uc.k_lo = k_lo; uc.k_hi = k_hi;
    }
    return true;
  }
};

template <int BN, bool HALO, bool XFORM, bool PAIR = false>
__device__ __forceinline__ void conv_v2_layer(const ConvKParams& p, uint8_t* smem) {
  const ConvKParams& tm = p;
  using Cfg = V2Cfg<BN, HALO, XFORM, PAIR>;
  constexpr int MT = Cfg::MT, NA = Cfg::NA, NB = Cfg::NB;
  constexpr int CW = BN < 32 ? 16 : 32;

  uint8_t* sA = smem;
  uint8_t* sB = smem + NA * Cfg::A_STAGE;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + NB * Cfg::B_STAGE);
  uint64_t* a_empty = a_full + NA;
  uint64_t* a_ready = a_empty + NA;  // XFORM: tile rewritten, visible to the tensor core
  uint64_t* b_full = a_ready + NA;
  uint64_t* b_empty = b_full + NB;
  uint64_t* t_full = b_empty + NB;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  volatile int* last_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);  // [epilogue groups <= 2]
  double* sm_stats = reinterpret_cast<double*>(tmem_slot + 4);               // [EPI_WARPS][64] per-warp GroupNorm sums (halo mode)
  float* sm_coef = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sm_stats + 64 * Cfg::EPI_WARPS) + 15) & ~uintptr_t(15));  // XFORM: [3][XF_MAXC] scale, shift, addend
  float* sm_bias = sm_coef + (XFORM ? 3 * Cfg::XF_MAXC : 0);  // [MT][BN], 16-byte aligned
  uint8_t* sm_stage = reinterpret_cast<uint8_t*>(sm_bias + MT * BN);  // [EPI_WARPS][2048], generic tiles only

  // warp index through a shuffle: provably warp-uniform, so the role loops below run on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // piece bookkeeping of shared tiles (split-K / stream-K) is per CTA of the plain grid, per cluster in PAIR mode
  const int cta = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, ncta = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool leader = !PAIR || cluster_ctarank() == 0;

  B2D_TL(0, threadIdx.x == 0);  // kernel entry
  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_ready[i], XFORM ? Cfg::XF_WARPS : 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    // PAIR: the leader's MMA warp waits for the epilogue warps of BOTH CTAs before it reuses an accumulator stage
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], (PAIR ? 2 : 1) * Cfg::EPI_WARPS); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 64 * Cfg::EPI_WARPS; i += blockDim.x) sm_stats[i] = 0.0;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) prefetch_tmap(&tm.tmapA[s]);
    prefetch_tmap(&tm.tmapB);
  }
  if (warp == 1) {
    if constexpr (PAIR) { tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  B2D_TL(1, threadIdx.x == 0);  // set-up done

  if (warp == 0) {
    // ================================ TMA producer =========================================
    // The whole warp walks the loop (uniform control flow and addresses); one elected lane issues.
    int ast = 0, bst = 0;
    uint32_t aph = 0, bph = 0;
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    CtaWalk<PAIR> wi;
    UnitCoord uc;
    for (wi.init(p); wi.next(p, uc);) {
      const int gcol0 = uc.gcol0 * BN;
      if constexpr (!HALO && !PAIR) {
        // Lean generic walk.  One K step is 4 MMAs of BN / 2 clocks, and this warp runs alone on its scheduler: every
        // instruction of the loop is paid at its full latency.  The general (segment, tap, chunk) iterator costs ~110
        // dependent instructions per step (indexed parameter loads, tap geometry, two barrier pairs) = 650-800 clocks,
        // MORE than the step's 512 clocks of tensor work at BN = 256 (tools/timeline_conv.py, profiles/r2_timeline_*).
        // Here the per-(segment, tap) work is hoisted out of the chunk loop, and since the A and B rings advance in
        // lockstep (NA == NB) both tiles of a step complete on ONE barrier (a_full) and are released by ONE commit.
        static_assert(NA == NB, "generic staging: the A and B rings advance in lockstep");
        GroupIter it;
        it.init(p, uc.k_lo, uc.k_hi);
        int s = it.s, t = it.t, c = it.c, g = it.g;
        int left = uc.k_hi - uc.k_lo;
        while (left > 0) {
          const int nch = p.cchunks[s] - c;
          const int nc = nch < left ? nch : left;
          const int zz = uc.z0 + p.dz[t];
          if (!(p.skip_z && (zz < 0 || zz >= p.D))) {
            const int xx = uc.x0 * p.stride_w + p.dx[t];
            const int yy = uc.y0 * p.stride_h + p.dy[t];
            int cc = c * kBlockK;
            int kc = p.kbase[s] + t * p.cin[s] + cc;
            const void* tmA = &tm.tmapA[s];
            const int tlm = B2D_TL_MODE();
            const uint32_t tx = ((tlm & 2) ? 0u : (uint32_t)Cfg::A_TILE) + ((tlm & 4) ? 0u : (uint32_t)Cfg::B_STAGE);
#pragma unroll 1
            for (int i = 0; i < nc; ++i) {
              mbar_wait(&a_empty[ast], aph ^ 1);
              B2D_TL_ITER(0, g + i, lane == 0);
              if (elect_one()) {
                const uint32_t bar = smem_u32(&a_full[ast]);
                mbar_arrive_expect_tx(&a_full[ast], tx);
                if (!(tlm & 2)) tma_load_5d_u(sA_u + ast * Cfg::A_STAGE, tmA, bar, cc, xx, yy, zz, uc.n0);
                if (!(tlm & 4)) tma_load_2d_u(sB_u + ast * Cfg::B_STAGE, &tm.tmapB, bar, kc, gcol0);
              }
              __syncwarp();
              cc += kBlockK;
              kc += kBlockK;
              if (++ast == NA) { ast = 0; aph ^= 1; }
            }
          }
          left -= nc;
          g += nc;
          c = 0;
          ++t;
          if (g == p.goff[s + 1]) { t = 0; ++s; }
        }
      } else {
      GroupIter it;
      for (it.init(p, uc.k_lo, uc.k_hi); !it.done(); it.next(p)) {
        const int s = it.s;
        if constexpr (HALO) {
          // activation ring only: the weight ring is fed by its own warp (below) so that A tiles can be
          // prefetched NA deep instead of being serialised behind the nine weight-tile loads of a group
          const int zz = uc.z0 + p.dz[it.t * p.gtaps];
          if (zz < 0 || zz >= p.D) continue;
          mbar_wait(&a_empty[ast], aph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&a_full[ast], Cfg::A_TILE);
            tma_load_5d_u(sA_u + ast * Cfg::A_STAGE, &tm.tmapA[s], smem_u32(&a_full[ast]), it.c * kBlockK, uc.x0 - 1, uc.y0 - 1, zz, uc.n0);
          }
          __syncwarp();
          if (++ast == NA) { ast = 0; aph ^= 1; }
        } else {
          const int tp = it.t;
          const int zz = uc.z0 + p.dz[tp];
          if (p.skip_z && (zz < 0 || zz >= p.D)) continue;
          const int xx = uc.x0 * p.stride_w + p.dx[tp];
          const int yy = uc.y0 * p.stride_h + p.dy[tp];
          mbar_wait(&a_empty[ast], aph ^ 1);
          mbar_wait(&b_empty[bst], bph ^ 1);
          if (elect_one()) {
            if constexpr (PAIR) {
              // both CTAs' bytes complete on the LEADER's barriers; only the leader posts the expectation (for both)
              if (leader) {
                mbar_arrive_expect_tx(&a_full[ast], 2 * Cfg::A_TILE);
                mbar_arrive_expect_tx(&b_full[bst], 2 * Cfg::B_STAGE);
              }
              tma_load_5d_pair(sA_u + ast * Cfg::A_STAGE, &tm.tmapA[s], smem_u32(&a_full[ast]), it.c * kBlockK, xx, yy, zz, uc.n0);
              tma_load_2d_pair(sB_u + bst * Cfg::B_STAGE, &tm.tmapB, smem_u32(&b_full[bst]), p.kbase[s] + tp * p.cin[s] + it.c * kBlockK,
                               gcol0 + wi.rank * Cfg::B_ROWS);
            }
          }
          __syncwarp();
          if (++ast == NA) { ast = 0; aph ^= 1; }
          if (++bst == NB) { bst = 0; bph ^= 1; }
        }
      }
      }
    }
    B2D_TL(8, lane == 0);  // every load issued
  } else if (HALO && warp == Cfg::BP_WARP) {
    // ================================ weight (B) producer, halo mode ========================
    int bst = 0;
    uint32_t bph = 0;
    const uint32_t sB_u = smem_u32(sB);
    CtaWalk<PAIR> wi;
    UnitCoord uc;
    for (wi.init(p); wi.next(p, uc);) {
      const int gcol0 = uc.gcol0 * BN;
      GroupIter it;
      for (it.init(p, uc.k_lo, uc.k_hi); !it.done(); it.next(p)) {
        const int s = it.s;
        const int zz = uc.z0 + p.dz[it.t * p.gtaps];
        if (zz < 0 || zz >= p.D) continue;
        const int kb = p.kbase[s] + it.t * p.gtaps * p.cin[s] + it.c * kBlockK;
#pragma unroll 1
        for (int ip = it.j0 * Cfg::TPB; ip < it.j1 * Cfg::TPB; ip += Cfg::TPB) {
          mbar_wait(&b_empty[bst], bph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&b_full[bst], Cfg::B_STAGE);
#pragma unroll
            for (int j = 0; j < Cfg::TPB; ++j)
              tma_load_2d_u(sB_u + bst * Cfg::B_STAGE + j * Cfg::B_TILE, &tm.tmapB, smem_u32(&b_full[bst]), kb + (ip + j) * p.cin[s], gcol0);
          }
          __syncwarp();
          if (++bst == NB) { bst = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (PAIR: the leader CTA's only) ==============
    if (leader) {
    // Uniform loop; descriptors are 64-bit constants whose low word (start address >> 4) is the only
    // thing that moves, so the per-MMA work is one 32-bit uniform add.
    const uint32_t idesc = umma_idesc_16(PAIR ? 2 * kBlockM : kBlockM, BN, p.op_f16);
    const uint64_t adesc0 = umma_smem_desc(smem_u32(sA), HALO ? 18 * 128 : 1024, 2);
    const uint64_t bdesc0 = umma_smem_desc(smem_u32(sB), 1024, 2);
    const uint32_t adesc_hi = (uint32_t)(adesc0 >> 32), bdesc_hi = (uint32_t)(bdesc0 >> 32);
    const uint32_t adesc_lo0 = (uint32_t)adesc0, bdesc_lo0 = (uint32_t)bdesc0;
    int ast = 0, bst = 0, acc = 0;
    uint32_t aph = 0, bph = 0, accph = 0;
    CtaWalk<PAIR> wi;
    UnitCoord uc;
    for (wi.init(p); wi.next(p, uc);) {
      mbar_wait(&t_empty[acc], accph ^ 1);  // the epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
      uint32_t accum = 0;
      if constexpr (!HALO && !PAIR) {
        // lean generic loop (see the producer): one barrier wait, four MMAs and one commit per K step
        GroupIter it;
        it.init(p, uc.k_lo, uc.k_hi);
        int s = it.s, t = it.t, c = it.c, g = it.g;
        int left = uc.k_hi - uc.k_lo;
        const bool no_mma = (B2D_TL_MODE() & 1) != 0;
        while (left > 0) {
          const int nch = p.cchunks[s] - c;
          const int nc = nch < left ? nch : left;
          const int zz = uc.z0 + p.dz[t];
          if (!(p.skip_z && (zz < 0 || zz >= p.D))) {
#pragma unroll 1
            for (int i = 0; i < nc; ++i) {
              mbar_wait(&a_full[ast], aph);
              tc_fence_after();
              B2D_TL(2, lane == 0 && accum == 0);  // first operands of the unit have landed
              B2D_TL_ITER(1, g + i, lane == 0);
              if (elect_one()) {
                const uint32_t a_lo = adesc_lo0 + (uint32_t)(ast * (Cfg::A_STAGE >> 4));
                const uint32_t b_lo = bdesc_lo0 + (uint32_t)(ast * (Cfg::B_STAGE >> 4));
                if (!no_mma) {
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    umma_bf16(d_tmem, ((uint64_t)adesc_hi << 32) | (a_lo + (uint32_t)(2 * k)), ((uint64_t)bdesc_hi << 32) | (b_lo + (uint32_t)(2 * k)),
                              idesc, accum | (uint32_t)(k > 0));
                }
                umma_commit(&a_empty[ast]);
              }
              __syncwarp();
              B2D_TL_ITER(2, g + i, lane == 0);
              accum = 1;
              if (++ast == NA) { ast = 0; aph ^= 1; }
            }
          }
          left -= nc;
          g += nc;
          c = 0;
          ++t;
          if (g == p.goff[s + 1]) { t = 0; ++s; }
        }
      } else {
      GroupIter it;
      for (it.init(p, uc.k_lo, uc.k_hi); !it.done(); it.next(p)) {
        if constexpr (HALO) {
          const int zz = uc.z0 + p.dz[it.t * p.gtaps];
          if (zz < 0 || zz >= p.D) continue;
        } else {
          const int zz = uc.z0 + p.dz[it.t];
          if (p.skip_z && (zz < 0 || zz >= p.D)) continue;
        }
        mbar_wait(XFORM ? &a_ready[ast] : &a_full[ast], aph);
        B2D_TL(2, lane == 0 && accum == 0);  // first operands of the unit have landed
        const uint32_t a_lo = adesc_lo0 + (uint32_t)(ast * (Cfg::A_STAGE >> 4));
        // halo: the in-plane taps fed by one staged box (9, or 4 for the upsample-folded convs), restricted to the B-stage
        // steps [j0, j1) of this group that belong to the item; generic: one step per group
#pragma unroll 1
        for (int ip = it.j0 * Cfg::TPB; ip < it.j1 * Cfg::TPB; ip += Cfg::TPB) {
          mbar_wait(&b_full[bst], bph);
          tc_fence_after();
          B2D_TL_ITER(1, it.g, lane == 0 && !HALO);
          const uint32_t b_lo = bdesc_lo0 + (uint32_t)(bst * (Cfg::B_STAGE >> 4));
          if (elect_one()) {
            if (!(B2D_TL_MODE() & 1))
#pragma unroll
            for (int j = 0; j < Cfg::TPB; ++j) {
              // tap (dy,dx) = row offset ((1+dy)*18 + (1+dx)) * 128 B into the staged halo box (8 = 128 B >> 4)
              const int tap = ip + j;
              const uint32_t a_tap = HALO ? a_lo + (uint32_t)(((1 + p.dy[tap]) * 18 + (1 + p.dx[tap])) * 8) : a_lo;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  const uint64_t ad = ((uint64_t)adesc_hi << 32) | (a_tap + (uint32_t)(mt * 64 + 2 * k));
                  const uint64_t bd = ((uint64_t)bdesc_hi << 32) | (b_lo + (uint32_t)(j * (Cfg::B_TILE >> 4) + 2 * k));
                  if constexpr (PAIR) umma_f16_pair(d_tmem + mt * Cfg::BNC, ad, bd, idesc, (accum | (uint32_t)(j > 0)) | (uint32_t)(k > 0));
                  else umma_bf16(d_tmem + mt * Cfg::BNC, ad, bd, idesc, (accum | (uint32_t)(j > 0)) | (uint32_t)(k > 0));
                }
              }
            }
            if constexpr (PAIR) umma_commit_pair(&b_empty[bst]); else umma_commit(&b_empty[bst]);
          }
          __syncwarp();
          accum = 1;
          if (++bst == NB) { bst = 0; bph ^= 1; }
        }
        if (elect_one()) { if constexpr (PAIR) umma_commit_pair(&a_empty[ast]); else umma_commit(&a_empty[ast]); }
        __syncwarp();
        B2D_TL_ITER(2, it.g, lane == 0 && !HALO);
        if (++ast == NA) { ast = 0; aph ^= 1; }
      }
      }
      if (elect_one()) { if constexpr (PAIR) umma_commit_pair(&t_full[acc]); else umma_commit(&t_full[acc]); }
      __syncwarp();
      B2D_TL(3, lane == 0);  // last MMA of the unit issued
      if (++acc == Cfg::NACC) { acc = 0; accph ^= 1; }
    }
    }
  } else if (warp < 2 + Cfg::EPI_WARPS) {
    // ================================ epilogue ==============================================
    const int ew = warp - 2;
    const int grp = ew >> 2;                 // four-warp epilogue group
    const int mt = HALO ? grp : 0;           // halo: the M half this group handles
    const int col_off = HALO ? 0 : grp * Cfg::BNG;  // generic: its share of the tile's columns
    constexpr int BNG = Cfg::BNG;
    const int quad = warp & 3;   // TMEM lane quadrant this warp may read
    const int r = quad * 32 + lane;
    const int rpi_log = HALO ? 7 : (p.lbw + p.lbh + p.lbd);
    const int seg = rpi_log >= 5 ? 32 : (1 << rpi_log);
    int acc = 0;
    uint32_t accph = 0;
    // halo mode: GroupNorm sums of the current sample accumulate in shared memory and reach global memory
    // only when this CTA moves on to another sample (one fp64 atomic per group instead of one per warp per unit)
    constexpr bool SMEM_STATS = HALO;
    int cur_n = -1;
    const int epi_tid = threadIdx.x - 64;
    const uint32_t sm_stage_u = HALO ? 0u : smem_u32(sm_stage + ew * 2048);
    // one group per sample (the UNet's GroupNorm(1, C)): per-thread fp64 sums live in registers across the units of a
    // sample and meet the other lanes only when the sample changes
    const bool one_group = SMEM_STATS && p.stats_cpg >= p.cout && p.stats_cpg > 0;
    // generic tiles whose warps each lie inside one image (>= 32 rows per image): the same, flushed per warp
    const bool one_group_w = !SMEM_STATS && p.stats_cpg >= p.cout && p.stats_cpg > 0 && seg == 32;
    int cur_on = -1;
    double thr_acc[2] = {0.0, 0.0};
    auto flush_warp = [&]() {
      if (cur_on >= 0 && cur_on < p.N) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          thr_acc[0] += __shfl_xor_sync(0xffffffffu, thr_acc[0], off);
          thr_acc[1] += __shfl_xor_sync(0xffffffffu, thr_acc[1], off);
        }
        if (lane < 2) atomicAdd(p.stats + (long long)cur_on * 2 + lane, lane == 0 ? thr_acc[0] : thr_acc[1]);
      }
      thr_acc[0] = 0.0; thr_acc[1] = 0.0;
    };
    auto flush_stats = [&]() {
      if constexpr (SMEM_STATS) {
        if (p.stats_cpg > 0) {
          const int ng2 = 2 * (p.cout / p.stats_cpg);
          if (one_group) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              thr_acc[0] += __shfl_xor_sync(0xffffffffu, thr_acc[0], off);
              thr_acc[1] += __shfl_xor_sync(0xffffffffu, thr_acc[1], off);
            }
            if (lane < 2) sm_stats[ew * 64 + lane] += lane == 0 ? thr_acc[0] : thr_acc[1];
            thr_acc[0] = 0.0; thr_acc[1] = 0.0;
          }
          asm volatile("bar.sync 3, %0;" ::"n"(32 * Cfg::EPI_WARPS) : "memory");
          if (cur_n >= 0 && epi_tid < ng2) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < Cfg::EPI_WARPS; ++w) { v += sm_stats[w * 64 + epi_tid]; sm_stats[w * 64 + epi_tid] = 0.0; }
            if (v != 0.0) atomicAdd(p.stats + (long long)cur_n * ng2 + epi_tid, v);
          }
          asm volatile("bar.sync 3, %0;" ::"n"(32 * Cfg::EPI_WARPS) : "memory");
        }
      }
    };
    int bias_cobase = -1;
    SchedCtx sched_ctx;
    const SchedCtx* sched = nullptr;
    if constexpr (BN == 16) {
      if (p.out_mode == 3) { sched_ctx = sched_ctx_load(p); sched = &sched_ctx; }
    }
    CtaWalk<PAIR> wi;
    UnitCoord uc;
    for (wi.init(p); wi.next(p, uc);) {
      if constexpr (SMEM_STATS) {
        if (uc.n0 != cur_n) { flush_stats(); cur_n = uc.n0; }
      }
      int ox, oy, oz, on;
      if constexpr (HALO) {
        ox = uc.x0 + mt * 8 + (r & 7);
        oy = uc.y0 + (r >> 3);
        oz = uc.z0;
        on = uc.n0;
      } else {
        const int mw = (1 << p.lbw) - 1, mh = (1 << p.lbh) - 1, md = (1 << p.lbd) - 1;
        ox = uc.x0 + (r & mw);
        oy = uc.y0 + ((r >> p.lbw) & mh);
        oz = uc.z0 + ((r >> (p.lbw + p.lbh)) & md);
        on = uc.n0 + (r >> (p.lbw + p.lbh + p.lbd));
      }
      if (one_group_w && on != cur_on) { flush_warp(); cur_on = on; }
      EpiRow rw;
      rw.valid = ox < p.OW && oy < p.OH && oz < p.D && on < p.N;
      rw.on = on;
      const int gcol0 = uc.gcol0 * BN;
      int phase_idx = 0, co_base = gcol0;
      if (p.nphase > 1) { phase_idx = gcol0 / p.cout; co_base = gcol0 - phase_idx * p.cout; }
      co_base += col_off;
      rw.out_y = oy * p.out_sy + p.out_oy + ((p.nphase > 1) ? (phase_idx >> 1) : 0);
      rw.out_x = ox * p.out_sx + p.out_ox + ((p.nphase > 1) ? (phase_idx & 1) : 0);
      rw.img = (long long)on * p.D + oz;
      rw.opix = (rw.img * p.out_H + rw.out_y) * p.out_W + rw.out_x;

      // stage the tile's bias in shared memory while the accumulator is still being produced (first barrier: every warp
      // of the group is done with the previous unit's values; second: the new ones are visible); units of one CTA often
      // share their N tile, then the staged values are simply kept
      uint32_t sm_bias_u = 0u;
      if (p.bias != nullptr) {
        float* sb = sm_bias + grp * BNG;
        if (co_base != bias_cobase) {
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          for (int i = (threadIdx.x - 64) & 127; i < BNG; i += 128) sb[i] = (co_base + i < p.cout) ? __ldg(p.bias + co_base + i) : 0.f;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          bias_cobase = co_base;
        }
        sm_bias_u = smem_u32(sb);
      }

      B2D_TL(9, ew == 0 && lane == 0);  // epilogue warp ready (bias staged), waiting for the accumulator
      mbar_wait(&t_full[acc], accph);
      tc_fence_after();
      B2D_TL(4, ew == 0 && lane == 0);  // accumulator complete
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * Cfg::ACC_COLS + mt * Cfg::BNC + col_off);
      auto load_tmem = [&](int col0, float (&f)[CW]) {
        uint32_t v[32];
        if constexpr (CW == 32) tmem_ld_32x32(taddr + col0, v); else tmem_ld_32x16(taddr + col0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CW; ++j) f[j] = __uint_as_float(v[j]);
      };
      const PieceInfo pi = piece_info(p, uc, cta, ncta);
      if (pi.npieces == 1) {
        conv_epilogue_row<BNG, CW, SMEM_STATS, !HALO>(p, rw, co_base, lane, seg, sm_stats + ew * 64, load_tmem, (one_group || one_group_w) ? thr_acc : nullptr, sm_bias_u, sm_stage_u, sched);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if constexpr (PAIR) mbar_arrive_cluster(&t_empty[acc], 0); else mbar_arrive(&t_empty[acc]); }
        B2D_TL(7, ew == 0 && lane == 0);  // unit's epilogue done (unshared tile)
      } else {
        // ---- the tile's K loop is shared (split-K / stream-K): park the fp32 partial, the last piece to arrive reduces
        // all of them in piece order (deterministic) ----
        // partial tile layout [column quad][row][4 floats]: a warp's 16-byte accesses cover 512 contiguous bytes
        float4* wq = reinterpret_cast<float4*>(p.ws + ((long long)piece_slot(p, uc, pi, pi.piece, ncta) * MT + mt) * (128LL * BN)) + r;
#pragma unroll 1
        for (int col0 = 0; col0 < BNG; col0 += CW) {
          float f[CW];
          load_tmem(col0, f);
#pragma unroll
          for (int q = 0; q < CW / 4; ++q)
            __stcg(wq + ((col_off + col0) / 4 + q) * 128, make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if constexpr (PAIR) mbar_arrive_cluster(&t_empty[acc], 0); else mbar_arrive(&t_empty[acc]); }
        __threadfence();
        B2D_TL(5, ew == 0 && lane == 0);  // partial tile parked
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        if ((threadIdx.x & 127) == 64) {  // first thread of this four-warp group (warps 2.. start at thread 64)
          int* ctr = p.counters + (uc.tile * MT + mt) * Cfg::NCG + (HALO ? 0 : grp);  // one ticket per (tile, M half, column group)
          const int old = atomicAdd(ctr, 1);
          const int last = (old == pi.npieces - 1) ? 1 : 0;
          if (last) *ctr = 0;  // self-reset for the next launch
          last_flag[grp] = last;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        B2D_TL(6, ew == 0 && lane == 0);  // ticket taken
        if (last_flag[grp]) {
          __threadfence();
          const float4* wbase = reinterpret_cast<const float4*>(p.ws + (long long)mt * (128LL * BN)) + r;
          const long long slot_stride = (long long)MT * 32 * BN;  // float4 units per slot
          auto load_ws = [&](int col0, float (&f)[CW]) {
#pragma unroll
            for (int j = 0; j < CW; ++j) f[j] = 0.f;
            // two pieces' loads in flight at a time (each is an L2 round trip of ~1000 clocks for this lone warp group);
            // the additions keep the piece order, so the sum is bit-identical to the one-at-a-time loop
            int ks = 0;
#pragma unroll 1
            for (; ks + 1 < pi.npieces; ks += 2) {
              const float4* s0 = wbase + piece_slot(p, uc, pi, ks, ncta) * slot_stride + ((col_off + col0) / 4) * 128;
              const float4* s1 = wbase + piece_slot(p, uc, pi, ks + 1, ncta) * slot_stride + ((col_off + col0) / 4) * 128;
              float4 v0[CW / 4], v1[CW / 4];
#pragma unroll
              for (int q = 0; q < CW / 4; ++q) v0[q] = __ldcg(s0 + q * 128);
#pragma unroll
              for (int q = 0; q < CW / 4; ++q) v1[q] = __ldcg(s1 + q * 128);
#pragma unroll
              for (int q = 0; q < CW / 4; ++q) {
                f[4 * q] += v0[q].x; f[4 * q + 1] += v0[q].y; f[4 * q + 2] += v0[q].z; f[4 * q + 3] += v0[q].w;
              }
#pragma unroll
              for (int q = 0; q < CW / 4; ++q) {
                f[4 * q] += v1[q].x; f[4 * q + 1] += v1[q].y; f[4 * q + 2] += v1[q].z; f[4 * q + 3] += v1[q].w;
              }
            }
            if (ks < pi.npieces) {
              const float4* src = wbase + piece_slot(p, uc, pi, ks, ncta) * slot_stride + ((col_off + col0) / 4) * 128;
#pragma unroll
              for (int q = 0; q < CW / 4; ++q) {
                const float4 v = __ldcg(src + q * 128);
                f[4 * q] += v.x; f[4 * q + 1] += v.y; f[4 * q + 2] += v.z; f[4 * q + 3] += v.w;
              }
            }
          };
          conv_epilogue_row<BNG, CW, SMEM_STATS, !HALO>(p, rw, co_base, lane, seg, sm_stats + ew * 64, load_ws, (one_group || one_group_w) ? thr_acc : nullptr, sm_bias_u, sm_stage_u);
          B2D_TL(7, ew == 0 && lane == 0);  // shared tile reduced and stored by this (last) piece
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");  // last_flag is reused by the next unit
      }
      if (++acc == Cfg::NACC) { acc = 0; accph ^= 1; }
    }
    if (one_group_w) flush_warp();
    flush_stats();
  } else if constexpr (XFORM) {
    // ================================ A-tile transform =====================================
    // Rewrite every staged 18x18x64 tile in place: raw (fp16 | bf16) -> bf16 silu(scale[c] * x + shift[c]).  Pixels
    // outside the image stay zero (TMA zero fill == the conv's zero padding, applied AFTER the activation).
    const int xt = threadIdx.x - (64 + 32 * Cfg::EPI_WARPS);  // 0..127
    const int chunk_phys = xt & 7;                            // 16-byte chunk inside the 128-byte row
    const int row0 = xt >> 3;                                 // 16 rows per pass
    float* coef_a = sm_coef;
    float* coef_b = sm_coef + Cfg::XF_MAXC;
    float* coef_c = sm_coef + 2 * Cfg::XF_MAXC;
    const uint32_t coef_a_u = smem_u32(coef_a), coef_b_u = smem_u32(coef_b), coef_c_u = smem_u32(coef_c);
    const bool has_add = p.in_temb != nullptr;
    const int cin_pad = p.cin[0];
    const int ngrp = p.in_creal / p.in_cpg;
    const bool act_on = p.in_act != 0, in_f16 = p.in_f16 != 0;
    const int op_f16 = p.op_f16;
    int ast = 0, cur_n = -1;
    uint32_t aph = 0;
    CtaWalk<PAIR> wi;
    UnitCoord uc;
    for (wi.init(p); wi.next(p, uc);) {
      if (uc.n0 != cur_n) {
        // per-channel scale / shift of this sample from the producer's fp64 (sum, sumsq)
        asm volatile("bar.sync 4, 128;" ::: "memory");  // everyone is done with the previous table
        for (int c = xt; c < cin_pad; c += 128) {
          float a = 0.f, b = 0.f;
          if (c < p.in_creal) {
            const int g = c / p.in_cpg;
            const double* st = p.in_stats + ((long long)uc.n0 * ngrp + g) * 2;
            const double mean = st[0] / p.in_count;
            double var = st[1] / p.in_count - mean * mean;
            if (var < 0) var = 0;
            const float rstd = (float)(1.0 / sqrt(var + (double)p.in_eps));
            const float ga = p.in_gamma ? __ldg(p.in_gamma + c) : 1.f, be = p.in_beta ? __ldg(p.in_beta + c) : 0.f;
            a = rstd * ga;
            b = be - (float)mean * rstd * ga;
          }
          coef_a[c] = act_on ? 0.5f * a : a;
          coef_b[c] = act_on ? 0.5f * b : b;
          coef_c[c] = (has_add && c < p.in_creal)
                          ? __ldg(p.in_temb + (long long)__ldg(p.in_temb_row + (long long)uc.n0 * p.in_temb_row_stride) * p.in_temb_ncols + p.in_temb_col + c)
                          : 0.f;
        }
        asm volatile("bar.sync 4, 128;" ::: "memory");
        cur_n = uc.n0;
      }
      GroupIter it;
      for (it.init(p, uc.k_lo, uc.k_hi); !it.done(); it.next(p)) {
        const int zz = uc.z0 + p.dz[it.t * p.gtaps];
        if (zz < 0 || zz >= p.D) continue;
        mbar_wait(&a_full[ast], aph);
        const uint32_t tile_u = smem_u32(sA + ast * Cfg::A_STAGE);
        const int cbase = it.c * kBlockK;
        // silu(y) = h + h * tanh(h) with h = y / 2; the 1/2 is folded into the coefficients (one MUFU per element).
        // Branch-free body, four independent 16-byte vectors in flight per thread (this warp is alone on its scheduler).
        const bool border = uc.x0 == 0 || uc.y0 == 0 || uc.x0 + 16 >= p.OW || uc.y0 + 16 >= p.OH;
        int yi = (row0 * 3641) >> 16;  // row0 / 18
        int xi = row0 - yi * 18;
#pragma unroll 1
        for (int r = row0; r < 18 * 18; r += 64) {
          uint4 raw[4];
          uint32_t vaddr[4];
          int c0[4];
          bool live[4];
          int yj = yi, xj = xi;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int rr = r + 16 * q;
            const uint32_t row_u = tile_u + (uint32_t)rr * 128u;
            vaddr[q] = row_u + (uint32_t)(chunk_phys * 16);
            c0[q] = cbase + ((chunk_phys ^ (int)((row_u >> 7) & 7u)) << 3);  // undo the 128B swizzle: which 8 channels
            live[q] = rr < 18 * 18;
            if (border) {
              const int gx = uc.x0 - 1 + xj, gy = uc.y0 - 1 + yj;
              live[q] = live[q] && gx >= 0 && gx < p.OW && gy >= 0 && gy < p.OH;  // outside the image: stays zero (padding)
            }
            xj += 16;
            if (xj >= 18) { xj -= 18; ++yj; }
            raw[q] = live[q] ? lds128(vaddr[q]) : make_uint4(0u, 0u, 0u, 0u);
          }
          yi = yj; xi = xj;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
            if (in_f16) {
              const float2 q0 = unpack_f16(raw[q].x), q1 = unpack_f16(raw[q].y), q2 = unpack_f16(raw[q].z), q3 = unpack_f16(raw[q].w);
              f[0] = q0.x; f[1] = q0.y; f[2] = q1.x; f[3] = q1.y; f[4] = q2.x; f[5] = q2.y; f[6] = q3.x; f[7] = q3.y;
            } else {
              f[0] = bf16_lo(raw[q].x); f[1] = bf16_hi(raw[q].x); f[2] = bf16_lo(raw[q].y); f[3] = bf16_hi(raw[q].y);
              f[4] = bf16_lo(raw[q].z); f[5] = bf16_hi(raw[q].z); f[6] = bf16_lo(raw[q].w); f[7] = bf16_hi(raw[q].w);
            }
            const float4 a0 = lds128f(coef_a_u + c0[q] * 4), a1 = lds128f(coef_a_u + c0[q] * 4 + 16);
            const float4 b0 = lds128f(coef_b_u + c0[q] * 4), b1 = lds128f(coef_b_u + c0[q] * 4 + 16);
            f[0] = fmaf(f[0], a0.x, b0.x); f[1] = fmaf(f[1], a0.y, b0.y); f[2] = fmaf(f[2], a0.z, b0.z); f[3] = fmaf(f[3], a0.w, b0.w);
            f[4] = fmaf(f[4], a1.x, b1.x); f[5] = fmaf(f[5], a1.y, b1.y); f[6] = fmaf(f[6], a1.z, b1.z); f[7] = fmaf(f[7], a1.w, b1.w);
            if (act_on) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float t;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(f[j]));
                f[j] = fmaf(f[j], t, f[j]);
              }
            }
            if (has_add) {
              const float4 d0 = lds128f(coef_c_u + c0[q] * 4), d1 = lds128f(coef_c_u + c0[q] * 4 + 16);
              f[0] += d0.x; f[1] += d0.y; f[2] += d0.z; f[3] += d0.w; f[4] += d1.x; f[5] += d1.y; f[6] += d1.z; f[7] += d1.w;
            }
            if (live[q]) sts128(vaddr[q], make_uint4(pack16(f[0], f[1], op_f16), pack16(f[2], f[3], op_f16), pack16(f[4], f[5], op_f16), pack16(f[6], f[7], op_f16)));
          }
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[ast]);
        if (++ast == NA) { ast = 0; aph ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  B2D_TL(10, threadIdx.x == 0);  // every role done
  if constexpr (PAIR) cluster_sync_all();  // no remote arrival or MMA operand read may target a CTA that has exited
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if constexpr (BN == 16) {
    // fused sampler update: the last CTA to finish advances the device step counter (every CTA has read it by then)
    if (p.out_mode == 3 && p.sch_step != nullptr && p.sch_step_inc != 0 && threadIdx.x == 0) {
      __threadfence();
      const unsigned int ticket = atomicAdd(p.sch_ticket, 1u);
      if (ticket == gridDim.x - 1) {
        *p.sch_ticket = 0u;
        *p.sch_step = *p.sch_step + p.sch_step_inc;
        __threadfence();
      }
    }
  }
}

}  // namespace b2d
