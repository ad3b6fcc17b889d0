// Host-side enumeration of the persistent conv kernel's work decomposition (csrc/conv_work.cuh): for a set of layer
// geometries and grid sizes, every (tile, B-stage step) must be executed exactly once, the pieces of a shared tile must
// agree on their count / order, and the workspace slots of the partial tiles must be distinct.  Driven by
// tests/test_work_partition.py (compiled with nvcc as a plain host program; no GPU needed).
#include <stdio.h>
#include <stdlib.h>
#include <map>
#include <set>
#include <vector>

#include "../../diffusion_model_project_b200/csrc/conv_work.cuh"

using namespace b2d;

static int check(int tiles_m, int ncol, int ngroups, int spg, int ksplit, int streamk, int contig, int ncta) {
  ConvKParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = 1; p.cchunks[0] = ngroups; p.goff[0] = 0; p.goff[1] = ngroups; p.ngroups = ngroups;
  p.lbw = 4; p.lbh = 3; p.lbd = 0; p.lbn = 0;
  p.tiles_w = 1; p.tiles_h = 1; p.tiles_d = 1; p.tiles_n = tiles_m; p.tiles_ncol = ncol;
  p.fd_ncol.set(ncol); p.fd_w.set(1); p.fd_h.set(1); p.fd_d.set(1);
  p.ksplit = ksplit; p.fd_ksplit.set(ksplit);
  p.spg = spg; p.ksteps = ngroups * spg; p.streamk = streamk; p.contig = contig;
  const int tiles = tiles_m * ncol;
  p.num_units = tiles * ksplit;
  p.total_steps = streamk ? tiles * p.ksteps : 0;
  p.fd_ksteps.set(p.ksteps); p.fd_total.set(streamk ? p.total_steps : 1);
  std::vector<int> cover((size_t)tiles * p.ksteps, 0);
  std::map<int, std::vector<int>> pieces;  // tile -> piece indices seen
  std::set<int> slots;
  for (int b = 0; b < ncta; ++b) {
    WorkIter wi;
    UnitCoord uc;
    int parked = 0;
    for (wi.init(p, b, ncta); wi.next(p, uc);) {
      if (uc.k_lo < 0 || uc.k_hi > p.ksteps || uc.k_lo >= uc.k_hi) { printf("bad range tile %d [%d,%d)\n", uc.tile, uc.k_lo, uc.k_hi); return 1; }
      if (uc.n0 != uc.tile / ncol || uc.gcol0 != uc.tile % ncol) { printf("bad decode\n"); return 1; }
      // the groups / steps the role loops would execute
      GroupIter it;
      int k = uc.k_lo;
      for (it.init(p, uc.k_lo, uc.k_hi); !it.done(); it.next(p)) {
        if (it.g != k / spg || it.c != it.g || it.s != 0 || it.t != 0) { printf("bad group %d (k %d)\n", it.g, k); return 1; }
        for (int j = it.j0; j < it.j1; ++j) {
          if (it.g * spg + j != k) { printf("step order\n"); return 1; }
          ++cover[(size_t)uc.tile * p.ksteps + k];
          ++k;
        }
      }
      if (k != uc.k_hi) { printf("group walk ended at %d, want %d\n", k, uc.k_hi); return 1; }
      const PieceInfo pi = piece_info(p, uc, b, ncta);
      if (pi.npieces < 1 || pi.npieces > kMaxPieces || pi.piece < 0 || pi.piece >= pi.npieces) { printf("bad piece info %d/%d\n", pi.piece, pi.npieces); return 1; }
      if ((pi.npieces == 1) != (uc.k_lo == 0 && uc.k_hi == p.ksteps)) { printf("npieces vs range\n"); return 1; }
      pieces[uc.tile].push_back(pi.piece * 1000 + pi.npieces);
      if (pi.npieces > 1) {
        const int slot = piece_slot(p, uc, pi, pi.piece, ncta);
        if (!slots.insert(slot).second) { printf("slot %d used twice\n", slot); return 1; }
        if (streamk && (slot / 2 != b || ++parked > 2)) { printf("stream-K slot %d of cta %d\n", slot, b); return 1; }
      }
    }
  }
  for (size_t i = 0; i < cover.size(); ++i)
    if (cover[i] != 1) { printf("step %zu covered %d times\n", i, cover[i]); return 1; }
  for (auto& kv : pieces) {
    const int n = kv.second[0] % 1000;
    if ((int)kv.second.size() != n) { printf("tile %d: %zu pieces seen, %d claimed\n", kv.first, kv.second.size(), n); return 1; }
    std::set<int> idx;
    for (int v : kv.second) { if (v % 1000 != n) { printf("piece count disagrees\n"); return 1; } idx.insert(v / 1000); }
    if ((int)idx.size() != n) { printf("duplicate piece index\n"); return 1; }
  }
  return 0;
}

int main() {
  int bad = 0, n = 0;
  const int grids[] = {1, 7, 88, 132, 148};
  for (int tiles_m : {1, 3, 11, 44, 88, 176, 300})
    for (int ncol : {1, 2, 4, 16})
      for (int ngroups : {1, 4, 9, 36, 144})
        for (int spg : {1, 3, 9}) {
          for (int g : grids) {
            // stream-K, under the planner's own admissibility rule (conv_plan.cu): at least one step per CTA and at most
            // kMaxPieces CTAs per tile
            const long long S = (long long)ngroups * spg, total = (long long)tiles_m * ncol * S, per_cta = total / g;
            if (total >= g && S / per_cta + 2 <= kMaxPieces) { bad += check(tiles_m, ncol, ngroups, spg, 1, 1, 0, g); ++n; }
            for (int ks : {1, 2, 3, 5}) {
              if (ks > ngroups) continue;
              bad += check(tiles_m, ncol, ngroups, spg, ks, 0, 0, g); ++n;  // strided
              bad += check(tiles_m, ncol, ngroups, spg, ks, 0, 1, g); ++n;  // contiguous
            }
          }
        }
  printf("%d configurations, %d failed\n", n, bad);
  return bad ? 1 : 0;
}
