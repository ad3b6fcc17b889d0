// Work decomposition of the persistent conv kernel (conv_v2.cuh), host + device: which (tile, K-step range) items a CTA
// executes, and -- for tiles whose K loop is shared by several CTAs -- which piece an item is and where its fp32
// partial tile lives in the workspace.  Plain integer arithmetic on ConvKParams; tests/test_work_partition.py
// enumerates it on the host.
//
// Three walks:
//   strided     CTA b executes units b, b + G, b + 2G, ...            unit = (tile, K split)
//   contiguous  CTA b executes units [b * U / G, (b + 1) * U / G)      (few sample changes: GroupNorm flushes)
//   stream-K    the flat space of tiles * ksteps B-stage steps is cut into G equal contiguous ranges; a range covers
//               [tail of a tile][whole tiles ...][head of a tile]: layers with fewer tiles than SMs (88 slice-images
//               of the UNet on 148 SMs) still keep every SM busy for the same number of steps.
#pragma once
#include "conv_common.cuh"

namespace b2d {

struct UnitCoord {
  int x0, y0, z0, n0, gcol0, tile;  // tile = m_tile * tiles_ncol + n_tile
  int ks;                            // K split index (unit walks)
  int k_lo, k_hi;                    // B-stage steps of the tile's K loop this item covers: [k_lo, k_hi) of [0, ksteps)
};

__host__ __device__ __forceinline__ void decode_tile(const ConvKParams& p, int tile, UnitCoord& c) {
  uint32_t t = (uint32_t)tile, r;
  c.tile = tile;
  p.fd_ncol.divmod(t, t, r);
  c.gcol0 = (int)r;
  p.fd_w.divmod(t, t, r);
  c.x0 = (int)r << p.lbw;
  p.fd_h.divmod(t, t, r);
  c.y0 = (int)r << p.lbh;
  p.fd_d.divmod(t, t, r);
  c.z0 = (int)r << p.lbd;
  c.n0 = (int)t << p.lbn;
}

struct WorkIter {
  int u, end, step;  // unit walks
  int f, f_hi;       // stream-K: flat step cursor and end of this CTA's range
  __host__ __device__ __forceinline__ void init(const ConvKParams& p, int cta, int ncta) {
    u = end = step = f = f_hi = 0;
    if (p.streamk) {
      f = (int)((long long)cta * p.total_steps / ncta);
      f_hi = (int)((long long)(cta + 1) * p.total_steps / ncta);
    } else if (p.contig) {
      u = (int)((long long)cta * p.num_units / ncta);
      end = (int)((long long)(cta + 1) * p.num_units / ncta);
      step = 1;
    } else {
      u = cta; end = p.num_units; step = ncta;
    }
  }
  __host__ __device__ __forceinline__ bool next(const ConvKParams& p, UnitCoord& c) {
    if (p.streamk) {
      if (f >= f_hi) return false;
      const int tile = (int)p.fd_ksteps.div((uint32_t)f);
      const int base = tile * p.ksteps;
      decode_tile(p, tile, c);
      c.ks = 0;
      c.k_lo = f - base;
      c.k_hi = (f_hi - base < p.ksteps) ? f_hi - base : p.ksteps;
      f = base + c.k_hi;
      return true;
    }
    if (u >= end) return false;
    uint32_t t, r;
    p.fd_ksplit.divmod((uint32_t)u, t, r);
    decode_tile(p, (int)t, c);
    c.ks = (int)r;
    if (p.ksplit == 1) {
      c.k_lo = 0; c.k_hi = p.ksteps;
    } else {  // whole A-operand groups per split, as the cost model assumes
      c.k_lo = (int)p.fd_ksplit.div((uint32_t)(p.ngroups * (int)r)) * p.spg;
      c.k_hi = (int)p.fd_ksplit.div((uint32_t)(p.ngroups * ((int)r + 1))) * p.spg;
    }
    u += step;
    return true;
  }
};

// the items that share a tile's K loop
struct PieceInfo { int npieces, piece, c_first; };
__host__ __device__ __forceinline__ PieceInfo piece_info(const ConvKParams& p, const UnitCoord& c, int cta, int ncta) {
  PieceInfo pi;
  if (!p.streamk) {
    pi.npieces = p.ksplit; pi.piece = c.ks; pi.c_first = 0;
    return pi;
  }
  // CTA owning flat step f: the largest b with floor(b * T / G) <= f, i.e. floor(((f + 1) * G - 1) / T)
  const int f0 = c.tile * p.ksteps;
  pi.c_first = (int)p.fd_total.div((uint32_t)((f0 + 1) * ncta - 1));
  const int c_last = (int)p.fd_total.div((uint32_t)((f0 + p.ksteps) * ncta - 1));
  pi.npieces = c_last - pi.c_first + 1;
  pi.piece = cta - pi.c_first;
  return pi;
}
// workspace slot (units of one MT x 128 x BLOCK_N fp32 tile set) of piece j of the item's tile.  Stream-K: a CTA parks at
// most two partial tiles -- the tail piece of the first tile of its range (slot 2b) and the head piece of the last (2b+1).
__host__ __device__ __forceinline__ int piece_slot(const ConvKParams& p, const UnitCoord& c, const PieceInfo& pi, int j, int ncta) {
  if (!p.streamk) return c.tile * p.ksplit + j;
  const int b = pi.c_first + j;
  const int first_tile = (int)p.fd_ksteps.div((uint32_t)((long long)b * p.total_steps / ncta));
  return 2 * b + (c.tile != first_tile ? 1 : 0);
}

// iterate the A-operand groups covered by the step range [k_lo, k_hi) as (segment, tap-or-ztap, chunk), with the B-stage
// steps [j0, j1) of each group that fall inside the range
struct GroupIter {
  int s, t, c, g, j0, j1, k, k_hi;
  __host__ __device__ __forceinline__ void init(const ConvKParams& p, int k_lo, int k_hi_) {
    k = k_lo; k_hi = k_hi_;
    if (k_lo == 0) {  // the common case: the whole K loop, no divisions
      g = 0; s = 0; t = 0; c = 0; j0 = 0;
    } else {
      g = k_lo / p.spg;
      j0 = k_lo - g * p.spg;
      s = 0;
      while (s + 1 < p.nseg && g >= p.goff[s + 1]) ++s;
      const int r = g - p.goff[s];
      t = r / p.cchunks[s];
      c = r - t * p.cchunks[s];
    }
    const int left = k_hi - (k - j0);
    j1 = left < p.spg ? left : p.spg;
  }
  __host__ __device__ __forceinline__ bool done() const { return k >= k_hi; }
  __host__ __device__ __forceinline__ void next(const ConvKParams& p) {
    k += j1 - j0;
    ++g;
    j0 = 0;
    const int left = k_hi - k;
    j1 = left < p.spg ? left : p.spg;
    if (++c == p.cchunks[s]) {
      c = 0;
      ++t;
      if (g == p.goff[s + 1]) { t = 0; ++s; }
    }
  }
};

}  // namespace b2d
