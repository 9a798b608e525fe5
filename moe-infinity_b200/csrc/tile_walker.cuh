// Tile scheduling of the persistent grouped GEMM, shared by the kernel (device) and tests/test_tile_walker.py (host,
// through tools/tile_walker_host.cu): which (expert, weight-row tile, token tile, k-block range) a CTA processes next.
#pragma once

#if defined(__CUDACC__)
#define B2M_HD __host__ __device__ __forceinline__
#else
#define B2M_HD inline
#endif

namespace b2m {

B2M_HD int tw_min(int a, int b) { return a < b ? a : b; }
B2M_HD int tw_max(int a, int b) { return a > b ? a : b; }

struct TileInfo {
  int e, slot, m0, row0, ncols, kb_begin, kb_end;
};

// Per-CTA tile walker: tile ids ascend, so the expert cursor only moves forward.
struct TileWalker {
  const int* tile_start;  // smem [E+1]
  const int* offs;        // smem [E+1]
  const int* slots;       // smem [E]
  int E, NTv, ksplit, kblocks, e_cur, m_step;
  int mc, rank;   // multicast cluster size (1|2) and this CTA's rank: the CTAs of a cluster take adjacent token tiles
  // stream-K (split-K GEMM at decode): this CTA owns the k-block units [u_cur, u_end) of the concatenation of all tiles;
  // get() then ignores `tile` and hands out the next segment (part of one tile) of that range
  int stream;
  int u_cur, u_end;
  // expert-parallel direct mode: every local expert owns a fixed-capacity row region that the source ranks fill from the
  // front; offs[e] = first row of the region, cnts[e] = rows in it (null: rows are packed, n_e = offs[e+1] - offs[e])
  const int* cnts;
  B2M_HD bool get(int tile, TileInfo& t) {
    int kb0 = 0, kb1 = 0;
    if (stream) {
      if (u_cur >= u_end) return false;
      tile = u_cur / kblocks;
      kb0 = u_cur - tile * kblocks;
      kb1 = tw_min(kblocks, kb0 + (u_end - u_cur));
      u_cur += kb1 - kb0;
    } else if (tile >= tile_start[E]) {
      return false;
    }
    while (tile >= tile_start[e_cur + 1]) ++e_cur;
    const int e = e_cur;
    const int n_e = cnts ? cnts[e] : offs[e + 1] - offs[e];
    const int n_tiles = ((n_e + NTv - 1) / NTv + mc - 1) / mc;   // token-tile groups (one per cluster)
    int local = tile - tile_start[e];
    const int per_m = n_tiles * ksplit;
    const int m = local / per_m;
    local -= m * per_m;
    const int n = (local / ksplit) * mc + rank;
    const int s = local % ksplit;
    const int kb_per = (kblocks + ksplit - 1) / ksplit;
    t.e = e;
    t.slot = slots[e];
    t.m0 = m * m_step;
    t.row0 = offs[e] + n * NTv;
    t.ncols = tw_max(0, tw_min(NTv, n_e - n * NTv));   // 0: ghost tile of an odd group (loads + MMA still run in lock step)
    t.kb_begin = s * kb_per;
    t.kb_end = tw_min(kblocks, t.kb_begin + kb_per);
    if (stream) { t.kb_begin = kb0; t.kb_end = kb1; }
    return true;
  }
};

}  // namespace b2m
