// grouped_gemm.cu -- K3/K4: grouped (per-expert segmented) expert GEMMs on tcgen05 tensor cores.
//
// Replaces the per-expert ATen loop of the reference, core/parallel/expert_module.cpp:171-175 (Mixtral),
// :200-203 (DeepSeek), :31-35 (Switch): `matmul(act(matmul(x, w1^T)) * matmul(x, w3^T), w2^T)` issued as
// 3 cuBLAS GEMMs + 2 elementwise kernels per expert from C++ worker threads
// (core/parallel/expert_dispatcher.cpp:309-395).
//
// Design (decode-first, "swap-AB"): the expert weight matrix [M rows, K] (nn.Linear layout, K-major) is the
// UMMA *A* operand (M = 128 rows per tile, streamed once from HBM by TMA), the expert's tokens
// [n_e, K] are the UMMA *B* operand (N = NT in {16,32,64,128} tokens).  D[128, NT] accumulates in TMEM.
// One persistent CTA per SM walks a device-side tile list derived from the per-expert token offsets the
// routing kernel produced -- no host sync.  Warp roles: w0 TMA producer, w1 MMA issuer (one lane),
// w2 TMEM allocator, w4..7 epilogue (tcgen05.ld -> activation -> global).
//   DUAL  = gate and up projections share the token tile and are fused with act(g)*u in the epilogue.
//   !DUAL = single projection: activation epilogue (Switch wi + ReLU) or linear fp32 output
//           (down projection), optionally split-K with fp32 red.add.
#include "b2m_common.cuh"
#include "b2m_internal.h"
#include "ep_device.cuh"
#include "tile_walker.cuh"

namespace b2m {

constexpr int BLOCK_M = 128;   // weight rows per tile  (UMMA M)
constexpr int BLOCK_K = 64;    // 64 x 16-bit = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int GEMM_THREADS = 256;
constexpr int MAX_E = 256;
constexpr int SMEM_BUDGET = 216 * 1024;
#ifndef B2M_EPI_WARPS_256
#define B2M_EPI_WARPS_256 16   // epilogue warps of the 256-token-tile instantiations (4 per TMEM lane quadrant; measured 8: 6.6 ms, 16: 5.5-6.0 ms, 24: same)
#endif

template <int NT, bool DUAL>
struct GemmCfg {
  static constexpr int B_TILE_BYTES = NT * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = (DUAL ? 2 : 1) * A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 10 ? 10 : STAGES_RAW;
  static constexpr int ACC_COLS = (DUAL ? 2 : 1) * NT;
  // NT=256 with two accumulators fills the 512 TMEM columns with ONE stage (no MMA/epilogue overlap); it is used in the
  // tensor-bound regime because it halves the weight bytes an SM has to ingest per FLOP (the SM ingest port, ~64 B/clk,
  // is what limited the 128-token tiles to ~65 % tensor-pipe activity).  Eight epilogue warps shorten the bubble.
  static constexpr int ACC_STAGES = (ACC_COLS * 2 <= 512) ? 2 : 1;
  static constexpr int EPI_WARPS = (NT >= 256) ? B2M_EPI_WARPS_256 : 4;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int TMEM_COLS_RAW = ACC_COLS * ACC_STAGES;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                                   : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    (MAX_E + 1) * 4 * 2 /*tile_start, offsets*/ + MAX_E * 4 * 2 /*slots, region counts*/;
  static_assert(TMEM_COLS_RAW <= 512, "TMEM overflow");
  static_assert(STAGES >= 2, "need >=2 stages");
};

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == ACT_SILU) return x / (1.0f + expf(-x));
  if (act == ACT_RELU) return fmaxf(x, 0.0f);
  if (act == ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  return x;
}

// SiLU with the MUFU approximations (ex2.approx, rcp.approx).  Used only by the 256-token-tile instantiation in
// B2M_NUMERICS_FP32 mode (its single TMEM stage makes the epilogue serialise with the MMAs; the epilogue gets ~3x shorter).
// B2M_NUMERICS_REFERENCE always uses the precise form: the ~2 ulp(fp32) of the approximations would survive the following
// round-to-model-dtype in ~1e-3 of the elements.
__device__ __forceinline__ float silu_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * 1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

// expert-parallel direct mode: wait (one thread) until every source rank has published this layer's epoch, i.e. its rows
// and tags have landed in this rank's receive area.  Called by every thread that goes on to read peer-written data with
// ordinary loads (the acquire orders that thread's own later reads); the TMA producer adds a generic->async proxy fence.
__device__ __forceinline__ int ep_gemm_wait(const GemmParams& p) {
  const int want = *reinterpret_cast<const volatile int*>(p.ep_epoch);
  for (int r = 0; r < p.ep_nranks; ++r)
    while (ld_acquire_sys(p.ep_flag + r) < want) __nanosleep(32);
  return want;
}

// How a phase waits for the data its token operand depends on
enum : int { DEP_PDL = 0,     // griddepcontrol.wait: the predecessor kernel in the stream (stand-alone launches)
             DEP_GRID = 1 };  // fused gate/up + down kernel: every CTA of THIS grid has finished the gate/up phase
struct GridBar {
  int* word;   // [0] arrival counter, [1] generation (bumped by the last arriver)
  int gen;     // generation observed at kernel start
};
__device__ __forceinline__ void grid_bar_wait(const GridBar& g) {
  int v;
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(g.word + 1) : "memory");
    if (v == g.gen) __nanosleep(20);
  } while (v == g.gen);
}

// One grouped GEMM over the device-side tile list: TMA producer / MMA issuer / epilogue warps over a shared-memory ring.
// Used as the whole body of grouped_gemm_tc_kernel and, twice, by fused_ffn_kernel (gate/up phase, grid barrier, down phase).
// TMEM is allocated by the caller (tmem_base); `ctas` = how many CTAs of the grid take tiles of this phase.
template <int NT, bool DUAL, int DT, int MC>
__device__ __forceinline__ void gemm_body(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmB,
                                          const GemmParams& p, uint8_t* smem, uint32_t tmem_base, int dep_mode,
                                          const GridBar& gbar, int ctas) {
  using Cfg = GemmCfg<NT, DUAL>;
  // carve shared memory: [stages | barriers | tmem ptr | tile tables]
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + Cfg::ACC_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + Cfg::ACC_STAGES);
  int* tile_start = reinterpret_cast<int*>(tmem_ptr_smem + 2);
  int* offs = tile_start + (MAX_E + 1);
  int* slots = offs + (MAX_E + 1);
  int* cnts = slots + MAX_E;           // direct mode: rows per expert region

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int E = p.E;
  const int kblocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  // dual_m: the two A tiles are rows [m0, m0+128) and [m0+128, m0+256) of the SAME matrix sharing one token tile
  // (doubles the arithmetic intensity of the down projection at prefill; the kernel is L2->SM operand-bandwidth bound)
  // m_rows (EP region mode, MC == 1): weight rows a tile really covers (a multiple of 8, <= 128; the A tensor maps carry a box of
  // that many rows).  The MMA still runs M = 128; the tail rows of the stage hold stale bytes and their accumulator lanes are
  // dropped in the epilogue.  Lets the host cut I rows into as many tiles as there are SMs (14336 / 104 = 138 <= 148).
  const int m_rows = (MC == 1 && !p.dual_m && p.m_rows > 0) ? p.m_rows : BLOCK_M;
  const int m_step = (DUAL && p.dual_m) ? 2 * BLOCK_M : m_rows;
  const int m_tiles = (p.M + m_step - 1) / m_step;

  // ---- per-phase setup ------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (DUAL) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MC);   // with multicast operands a stage is free when every CTA of the cluster consumed it
    }
    for (int a = 0; a < Cfg::ACC_STAGES; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], Cfg::EPI_WARPS);
    }
    fence_barrier_init();
  }
  // Expert parallel (direct mode): this kernel's CTAs are resident while the routing kernel is still exchanging rows with the peers
  // and the previous layer's combine waits for its owners -- HBM is idle for ~30 us per layer (profiles/r02e_ep8_timeline.json).
  // Each CTA asks the TMA unit to pull the first ep_l2pf k-blocks of the weight tiles it will most likely own (one token tile
  // per local expert: tile = le * m_tiles + m) into L2.  A hint only: a wrong guess costs bandwidth, never correctness.  The slot
  // table is older than any predecessor in the stream.
  if (p.ep_l2pf > 0 && p.ep_rows > 0 && warp == 0 && lane == 0 && (int)blockIdx.x < ctas && MC == 1) {
    const int npf = p.ep_l2pf < kblocks ? p.ep_l2pf : kblocks;
    for (int tile = blockIdx.x; tile < p.ep_el * m_tiles; tile += ctas) {
      const int slot = p.slot_of[p.ep_first + tile / m_tiles];
      const int m0 = (tile % m_tiles) * m_step;
      if (slot < 0) continue;
      for (int kb = 0; kb < npf; ++kb) {
        tma_prefetch_3d(&tmA0, kb * BLOCK_K, m0, slot);
        if (DUAL) tma_prefetch_3d(&tmA1, kb * BLOCK_K, m0, slot);
      }
    }
  }
  // everything above is on-chip; from here on we touch memory earlier kernels produced.  In early_a mode (down projection
  // launched with a programmatic edge behind the gate/up GEMM) only the token-tile loads depend on the predecessor: the
  // producer prefetches weight tiles first and waits later; the routing tables read below are older than the predecessor.
  if (!p.early_a && dep_mode == DEP_PDL) pdl_wait();
  if (p.pdl_edge && dep_mode == DEP_PDL) pdl_launch();
  if (p.single_n >= 0) {
    if (threadIdx.x == 0) { offs[0] = 0; offs[1] = p.single_n; slots[0] = p.single_slot; }
  } else if (p.ep_rows > 0) {
    // direct mode: local expert le owns rows [le*ep_rows, le*ep_rows + cnt[le]) of the receive area; the counts are final
    // once every source rank has published this layer's flag
    if (threadIdx.x == 0) {
      ep_gemm_wait(p);
      fence_proxy_async_global();          // this thread is also the TMA producer: its later bulk loads read the peers' rows
      if (p.tl && p.ep_wait) tl_max(p.tl + 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= E; i += Cfg::THREADS) {
      int le = i - p.ep_first;
      le = le < 0 ? 0 : (le > p.ep_el ? p.ep_el : le);
      offs[i] = le * p.ep_rows;
    }
    for (int i = threadIdx.x; i < E; i += Cfg::THREADS) {
      slots[i] = p.slot_of[i];
      const int le = i - p.ep_first;
      cnts[i] = (le >= 0 && le < p.ep_el) ? __ldcg(p.ep_cnt + le) : 0;
    }
  } else {
    for (int i = threadIdx.x; i <= E; i += Cfg::THREADS) offs[i] = p.offsets[i];
    for (int i = threadIdx.x; i < E; i += Cfg::THREADS) slots[i] = p.slot_of[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < E; ++e) {
      tile_start[e] = acc;
      const int n_e = p.ep_rows > 0 ? cnts[e] : offs[e + 1] - offs[e];
      if (n_e > 0 && slots[e] >= 0) acc += m_tiles * (((n_e + NT - 1) / NT + MC - 1) / MC) * (p.stream_k ? 1 : p.ksplit);
    }
    tile_start[E] = acc;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int crank = MC > 1 ? (int)cluster_ctarank() : 0;
  if (MC > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast targets them

  TileWalker walker{tile_start, offs, slots, E, NT, p.stream_k ? 1 : p.ksplit, kblocks, 0, m_step, MC, crank, 0, 0, 0,
                    p.ep_rows > 0 ? cnts : nullptr};
  if (MC == 1 && p.stream_k) {
    const long long units = (long long)tile_start[E] * kblocks;
    walker.stream = 1;
    walker.u_cur = (int)blockIdx.x < ctas ? (int)(units * blockIdx.x / ctas) : 0;
    walker.u_end = (int)blockIdx.x < ctas ? (int)(units * (blockIdx.x + 1) / ctas) : 0;
  }
  TileInfo t;
  // tiles are dealt to clusters; in the fused kernel a phase may use fewer CTAs than the grid has (ctas): the rest take none
  const int tile_stride = ctas / MC;
  const int tile0 = (int)blockIdx.x < ctas ? (int)blockIdx.x / MC : 0x3fffffff;

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool waited = !p.early_a;
      int npend = 0, pend_stage[Cfg::STAGES], pend_kb[Cfg::STAGES], pend_row[Cfg::STAGES];
      for (int tile = tile0; walker.get(tile, t); tile += tile_stride) {
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA0 = stage_base + stage * Cfg::STAGE_BYTES;
          uint8_t* sA1 = sA0 + A_TILE_BYTES;
          uint8_t* sB = sA0 + (DUAL ? 2 : 1) * A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], (DUAL ? 2 : 1) * m_rows * BLOCK_K * 2 + Cfg::B_TILE_BYTES);
          // decode: weights are streamed exactly once -> evict-first.  Several n-tiles per expert (prefill): the same
          // weight tile is re-read by every n-tile -> keep it in L2.  Tokens are re-read by every m-tile: keep.
          const uint64_t wh = ((p.ep_rows > 0 ? cnts[t.e] : offs[t.e + 1] - offs[t.e]) <= NT) ? CACHE_EVICT_FIRST : CACHE_EVICT_NORMAL;
          if (MC == 1) {
            tma_load_3d(&tmA0, &full_bar[stage], sA0, kb * BLOCK_K, t.m0, t.slot, wh);
            if (DUAL) tma_load_3d(&tmA1, &full_bar[stage], sA1, kb * BLOCK_K, t.m0 + (p.dual_m ? BLOCK_M : 0), t.slot, wh);
          } else {
            // 2-CTA multicast: the cluster's CTAs work on the same weight tiles and adjacent token tiles; each CTA fetches
            // HALF of every weight tile (64 rows; tmA* carry 64-row boxes) and multicasts it into both CTAs' stages, so
            // a CTA pulls 16+16 KB instead of 32+16 KB per k-block from L2 (the kernel is L2->SM operand-bandwidth bound)
            const int half = crank * (BLOCK_M / 2);          // rows this CTA fetches
            const int hoff = crank * (A_TILE_BYTES / 2);     // where they sit inside the 128-row stage tile
            tma_load_3d_mc(&tmA0, &full_bar[stage], sA0 + hoff, kb * BLOCK_K, t.m0 + half, t.slot, (uint16_t)0x3, wh);
            if (DUAL)
              tma_load_3d_mc(&tmA1, &full_bar[stage], sA1 + hoff, kb * BLOCK_K,
                             t.m0 + (p.dual_m ? BLOCK_M : 0) + half, t.slot, (uint16_t)0x3, wh);
          }
          if (!waited) {
            // weight tiles of the first pipeline fill are in flight before the predecessor has finished; their token
            // tiles follow once it has
            pend_stage[npend] = stage; pend_kb[npend] = kb; pend_row[npend] = t.row0;
            if (++npend == Cfg::STAGES) {
              if (dep_mode == DEP_PDL) pdl_wait(); else { grid_bar_wait(gbar); fence_proxy_async_global(); }
              for (int i = 0; i < npend; ++i)
                tma_load_2d(&tmB, &full_bar[pend_stage[i]], stage_base + pend_stage[i] * Cfg::STAGE_BYTES + (DUAL ? 2 : 1) * A_TILE_BYTES,
                            pend_kb[i] * BLOCK_K, pend_row[i], CACHE_EVICT_LAST);
              waited = true;
            }
          } else {
            tma_load_2d(&tmB, &full_bar[stage], sB, kb * BLOCK_K, t.row0, CACHE_EVICT_LAST);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (!waited) {   // fewer work items than pipeline stages
        if (dep_mode == DEP_PDL) pdl_wait(); else { grid_bar_wait(gbar); fence_proxy_async_global(); }
        for (int i = 0; i < npend; ++i)
          tma_load_2d(&tmB, &full_bar[pend_stage[i]], stage_base + pend_stage[i] * Cfg::STAGE_BYTES + (DUAL ? 2 : 1) * A_TILE_BYTES,
                      pend_kb[i] * BLOCK_K, pend_row[i], CACHE_EVICT_LAST);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    if (p.early_a && dep_mode == DEP_PDL) pdl_wait();
    constexpr uint32_t idesc = make_idesc_f16(DT, BLOCK_M, NT);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile0; walker.get(tile, t); tile += tile_stride) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * Cfg::ACC_COLS;
      uint32_t idesc_t = idesc;
      if (p.dyn_n) {   // opt-in: N of this tile's MMAs = its own token count (multiple of 16, >= 16)
        const int n_eff = t.ncols <= 16 ? 16 : ((t.ncols + 15) & ~15);
        idesc_t = make_idesc_f16(DT, BLOCK_M, n_eff < NT ? n_eff : NT);
      }
      for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a0 = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
          const uint32_t a1 = a0 + A_TILE_BYTES;
          const uint32_t b = a0 + (DUAL ? 2 : 1) * A_TILE_BYTES;
          const uint64_t da0 = make_kmajor_sw128_desc(a0);
          const uint64_t da1 = make_kmajor_sw128_desc(a1);
          const uint64_t db = make_kmajor_sw128_desc(b);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint32_t accum = (kb > t.kb_begin || k > 0) ? 1u : 0u;
            const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);  // advance start address inside the SW128 atom
            tc_mma_f16(d0, da0 + koff, db + koff, idesc_t, accum);
            if (DUAL) tc_mma_f16(d0 + NT, da1 + koff, db + koff, idesc_t, accum);
          }
          if (MC == 1) tc_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
          else tc_commit_mc(&empty_bar[stage], (uint16_t)0x3);   // ... in both CTAs: the peer multicasts into it too
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) tc_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (4 warps = 128 TMEM lanes) =====================
    if (p.early_a && dep_mode == DEP_PDL) pdl_wait();
    if (p.ep_rows > 0) {
      if (p.ep_zero) {
        // clear the down projection's accumulator: safe now -- a source rank publishes this layer's flag only after its
        // combine of the previous layer has finished reading the previous outputs
        float4* z = reinterpret_cast<float4*>(p.ep_zero);
        const size_t n4 = p.ep_zero_elems / 4;
        const int et = Cfg::EPI_WARPS * 32;
        for (size_t i = (size_t)blockIdx.x * et + (threadIdx.x - 128); i < n4; i += (size_t)gridDim.x * et)
          z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const int q = warp & 3;              // TMEM lane quadrant this warp may access
    const int cgrp = (warp - 4) >> 2;    // with 8 epilogue warps two warps share a quadrant and alternate 16-column chunks
    const int r = q * 32 + lane;         // weight row inside the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile0; walker.get(tile, t); tile += tile_stride) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS;
      const int m = t.m0 + r;
      const bool row_ok = m < p.M && r < m_rows;
      float bias0 = 0.f, bias1 = 0.f;   // bias experts: one value per weight row = per TMEM lane
      if (p.bias_base) {
        const uint16_t* bp = reinterpret_cast<const uint16_t*>(p.bias_base) + (size_t)t.slot * p.bias_slot_elems + p.bias_off;
        if (row_ok) bias0 = Half16<DT>::to_f(bp[m]);
        if (DUAL && m + BLOCK_M < p.M) bias1 = Half16<DT>::to_f(bp[m + BLOCK_M]);
      }
      for (int c0 = cgrp * 16; c0 < t.ncols; c0 += 16 * (Cfg::EPI_WARPS / 4)) {   // warp-uniform trip count
        uint32_t vg[16], vu[16];
        tmem_ld_x16(taddr + c0, vg);
        if (DUAL) tmem_ld_x16(taddr + NT + c0, vu);
        tmem_ld_wait();
        if (p.epi == EPI_LINEAR_F32) {
          float* out = reinterpret_cast<float*>(p.out);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (c0 + j < t.ncols) {
              float* dst = out + (size_t)(t.row0 + c0 + j) * p.ld_out + m;
              if (row_ok) {
                float v = __uint_as_float(vg[j]);
                if (p.bias_base) v = p.mimic ? round_dt<DT>(round_dt<DT>(v) + bias0) : v + bias0;   // matmul, then `+ bias`
                if (p.ksplit > 1) atomicAdd(dst, v); else *dst = v;
              }
              if (DUAL && m + BLOCK_M < p.M) {   // dual_m: second accumulator = rows m0+128..m0+255
                float v = __uint_as_float(vu[j]);
                if (p.bias_base) v = p.mimic ? round_dt<DT>(round_dt<DT>(v) + bias1) : v + bias1;
                if (p.ksplit > 1) atomicAdd(dst + BLOCK_M, v); else dst[BLOCK_M] = v;
              }
            }
          }
        } else {
          uint16_t* out = reinterpret_cast<uint16_t*>(p.out);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (row_ok && c0 + j < t.ncols) {
              float g = __uint_as_float(vg[j]);
              float h;
              if (DUAL) {
                float u = __uint_as_float(vu[j]);
                constexpr bool kFastAct = NT >= 256;
                if (p.mimic) {   // reference rounding chain: each ATen op rounds to the model dtype; precise SiLU at every tile
                  g = round_dt<DT>(g);   // width, so "reference numerics" does not depend on T (the MUFU form is NUMERICS_FP32 only)
                  u = round_dt<DT>(u);
                  h = round_dt<DT>(act_apply(g, p.act)) * u;
                } else {
                  h = ((kFastAct && p.act == ACT_SILU) ? silu_mufu(g) : act_apply(g, p.act)) * u;
                }
              } else {
                if (p.mimic) g = round_dt<DT>(g);
                if (p.bias_base) { g += bias0; if (p.mimic) g = round_dt<DT>(g); }
                h = act_apply(g, p.act);
              }
              out[(size_t)(t.row0 + c0 + j) * p.ld_out + m] = Half16<DT>::from_f(h);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  // ---- end of the phase: every role of this CTA is done with the ring, the barriers and the accumulators
  tc_fence_before();
  __syncthreads();
}

// after the last phase of a kernel: timeline, the expert-parallel "done" signal, TMEM release
template <int MC>
__device__ __forceinline__ void gemm_finish(const GemmParams& p, uint32_t tmem_base, int tmem_cols) {
  const int warp = threadIdx.x >> 5;
  if (p.tl && threadIdx.x == 0) tl_max(p.tl + 2);
  if (p.ep_signal && threadIdx.x == 0) {
    // the CTA's output stores / reductions (all threads, ordered before this point by the barrier that ended the phase) become
    // visible system-wide with ONE fence (fences are cumulative); the last CTA of the grid then tells the source ranks, whose
    // combine kernels read the rows in place over NVLink
    __threadfence_system();
    if (atomicAdd(p.ep_done_ctr, 1) == (int)gridDim.x - 1) {
      *p.ep_done_ctr = 0;
      for (int le = 0; le < p.ep_el; ++le) p.ep_cnt[le] = 0;   // every CTA has read the counts: regions are free for the next layer
      const int e = *p.ep_done_epoch + 1;
      *p.ep_done_epoch = e;
      __threadfence_system();   // the other CTAs' fenced outputs and the counter reset happen-before the flags; the flag stores
      for (int r = 0; r < p.ep_nranks; ++r) st_relaxed_sys(p.ep_peer_done_flag[r] + p.ep_rank, e);   // themselves go out back to back
    }
  }
  if (MC > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast / commit into its shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

__device__ __forceinline__ uint8_t* smem_aligned(uint8_t* raw) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
}
// allocate `cols` TMEM columns (warp 2) and hand the base address to every thread
__device__ __forceinline__ uint32_t tmem_setup(uint32_t* slot, int cols) {
  if ((threadIdx.x >> 5) == 2) {
    tmem_alloc(slot, cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *slot;
}

template <int NT, bool DUAL, int DT, int MC>
__global__ void __launch_bounds__((GemmCfg<NT, DUAL>::THREADS), 1)
grouped_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<NT, DUAL>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t s_tmem;
  uint8_t* smem = smem_aligned(smem_raw);
  if (p.tl && threadIdx.x == 0) tl_min(p.tl);
  // pdl_edge: this grid may be resident long before its predecessor finishes (expert parallel: the gate/up GEMM arrives behind the
  // routing kernel to prefetch weights into L2).  Its successor reads state the predecessor publishes (dispatch epoch, row
  // counters) in ITS prologue, so it must not be released before this grid has passed its own wait: gemm_body triggers then.
  if (!p.pdl_edge) pdl_launch();
  const uint32_t tmem_base = tmem_setup(&s_tmem, Cfg::TMEM_COLS);
  GridBar nobar{nullptr, 0};
  gemm_body<NT, DUAL, DT, MC>(tmA0, tmA1, tmB, p, smem, tmem_base, DEP_PDL, nobar, (int)gridDim.x);
  gemm_finish<MC>(p, tmem_base, Cfg::TMEM_COLS);
}

#ifdef B2M_ENABLE_FUSED_FFN
// (Experiment, measured and NOT part of the default library: Mixtral decode 12.89 vs 12.84 ms/step, DeepSeek-V2-Lite decode
// 5.04 vs 4.75 ms/step -- a persistent barrier kernel keeps the side-stream shared-expert GEMMs off the SMs --, expert
// parallel N=2 8.51 vs 8.47 ms/step; profiles/r02c_*.json.  Build with -DB2M_ENABLE_FUSED_FFN and run with B2M_FUSED_FFN=1.)
// Fused expert FFN for the HBM-bound (decode) regime: gate/up GEMM + SwiGLU, grid-wide barrier, down GEMM in ONE persistent
// kernel.  What it removes from every layer: a kernel boundary on the critical path -- the down GEMM's CTAs could only become
// resident when the gate/up CTAs had left their SMs (both want all shared memory), so its TMEM/barrier set-up, tile tables and
// first pipeline fill were exposed (measured ~15 us per layer at every N, profiles/r02_ep*_timeline_regions.log).  Here the
// CTA keeps its resources, re-arms its barriers, starts streaming the down projection's weights at once and only its token
// tiles wait for the grid barrier.  Needs every CTA resident (persistent grid <= #SMs, one CTA per SM: true by construction)
// and no second barrier-kernel running beside it.
template <int NT, int DT>
__global__ void __launch_bounds__((GemmCfg<NT, true>::THREADS), 1)
fused_ffn_kernel(const __grid_constant__ CUtensorMap tmGate, const __grid_constant__ CUtensorMap tmUp,
                 const __grid_constant__ CUtensorMap tmBup, const __grid_constant__ CUtensorMap tmDown,
                 const __grid_constant__ CUtensorMap tmBdn, const GemmParams up, const GemmParams dn, int* gbar_word,
                 int up_ctas) {
  using CfgU = GemmCfg<NT, true>;
  using CfgD = GemmCfg<NT, false>;
  constexpr int TMEM_COLS = CfgU::TMEM_COLS > CfgD::TMEM_COLS ? CfgU::TMEM_COLS : CfgD::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t s_tmem;
  __shared__ int s_gen;
  uint8_t* smem = smem_aligned(smem_raw);
  if (up.tl && threadIdx.x == 0) tl_min(up.tl);
  pdl_launch();
  if (threadIdx.x == 0) s_gen = *reinterpret_cast<volatile int*>(gbar_word + 1);   // moves only after every CTA of this grid arrived
  const uint32_t tmem_base = tmem_setup(&s_tmem, TMEM_COLS);
  GridBar gb{gbar_word, s_gen};
  gemm_body<NT, true, DT, 1>(tmGate, tmUp, tmBup, up, smem, tmem_base, DEP_PDL, gb, up_ctas);
  if (threadIdx.x == 0) {
    if (up.tl) tl_max(up.tl + 2);
    if (dn.tl) tl_min(dn.tl);
    __threadfence();                       // this CTA's intermediate rows (all threads, ordered by the barrier that ended the phase)
    if (atomicAdd(gbar_word, 1) == (int)gridDim.x - 1) {
      gbar_word[0] = 0;
      __threadfence();
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(gbar_word + 1), "r"(gb.gen + 1) : "memory");
    }
  }
  gemm_body<NT, false, DT, 1>(tmDown, tmDown, tmBdn, dn, smem, tmem_base, DEP_GRID, gb, (int)gridDim.x);
  gemm_finish<1>(dn, tmem_base, TMEM_COLS);
}

#endif  // B2M_ENABLE_FUSED_FFN

// --------------------------------------------------------------------------------------
// Plain CUDA-core grouped GEMM with identical semantics (debug / bring-up cross-check only;
// selected with B2M_GEMM_IMPL=simt).  One warp per output element.
// --------------------------------------------------------------------------------------
template <int DT>
__global__ void grouped_gemm_simt_kernel(const uint16_t* __restrict__ arena, size_t slot_elems, size_t offA0,
                                         size_t offA1, const uint16_t* __restrict__ B, int ldb, GemmParams p,
                                         int dual) {
  const bool single = p.single_n >= 0;
  const int total_rows = single ? p.single_n : p.offsets[p.E];
  const long long nout = (long long)total_rows * p.M;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (long long o = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); o < nout;
       o += (long long)gridDim.x * warps_per_block) {
    const int row = (int)(o / p.M);
    const int m = (int)(o % p.M);
    int e = 0;
    if (!single) while (e + 1 < p.E && row >= p.offsets[e + 1]) ++e;
    const int slot = single ? p.single_slot : p.slot_of[e];
    if (slot < 0) continue;
    const uint16_t* a0 = arena + (size_t)slot * slot_elems + offA0 + (size_t)m * p.K;
    const uint16_t* a1 = arena + (size_t)slot * slot_elems + offA1 + (size_t)m * p.K;
    const uint16_t* b = B + (size_t)row * ldb;
    float s0 = 0.f, s1 = 0.f;
    for (int k = lane; k < p.K; k += 32) {
      const float x = Half16<DT>::to_f(b[k]);
      s0 += Half16<DT>::to_f(a0[k]) * x;
      if (dual) s1 += Half16<DT>::to_f(a1[k]) * x;
    }
    for (int d = 16; d; d >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, d);
      s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    }
    if (lane == 0) {
      float bias = 0.f;
      if (p.bias_base)
        bias = Half16<DT>::to_f(reinterpret_cast<const uint16_t*>(p.bias_base)[(size_t)slot * p.bias_slot_elems + p.bias_off + m]);
      if (p.epi == EPI_LINEAR_F32) {
        if (p.bias_base) s0 = p.mimic ? round_dt<DT>(round_dt<DT>(s0) + bias) : s0 + bias;
        reinterpret_cast<float*>(p.out)[(size_t)row * p.ld_out + m] = s0;
      } else {
        float g = s0, h;
        if (dual) {
          float u = s1;
          if (p.mimic) { g = round_dt<DT>(g); u = round_dt<DT>(u); h = round_dt<DT>(act_apply(g, p.act)) * u; }
          else h = act_apply(g, p.act) * u;
        } else {
          if (p.mimic) g = round_dt<DT>(g);
          if (p.bias_base) { g += bias; if (p.mimic) g = round_dt<DT>(g); }
          h = act_apply(g, p.act);
        }
        reinterpret_cast<uint16_t*>(p.out)[(size_t)row * p.ld_out + m] = Half16<DT>::from_f(h);
      }
    }
  }
}

// --------------------------------------------------------------------------------------
// host launchers
// --------------------------------------------------------------------------------------
template <int NT, bool DUAL, int DT, int MC = 1>
static cudaError_t launch_tc(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const GemmParams& p,
                             int grid, cudaStream_t st) {
  using Cfg = GemmCfg<NT, DUAL>;
  auto kern = grouped_gemm_tc_kernel<NT, DUAL, DT, MC>;
  static bool attr_done = false;   // per-instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  if (MC > 1) grid -= grid % MC;
  return launch_cluster(kern, dim3(grid), dim3(Cfg::THREADS), (size_t)Cfg::SMEM_BYTES, st, MC, p.early_a != 0 || p.pdl_edge != 0, a0, a1, b, p);
}

#ifdef B2M_ENABLE_FUSED_FFN
template <int NT, int DT>
static cudaError_t launch_fused(const CUtensorMap& g, const CUtensorMap& u, const CUtensorMap& bup, const CUtensorMap& d,
                                const CUtensorMap& bdn, const GemmParams& up, const GemmParams& dn, int grid, int up_ctas,
                                int* gbar, cudaStream_t st) {
  using CfgU = GemmCfg<NT, true>;
  using CfgD = GemmCfg<NT, false>;
  constexpr int SMEM = CfgU::SMEM_BYTES > CfgD::SMEM_BYTES ? CfgU::SMEM_BYTES : CfgD::SMEM_BYTES;
  auto kern = fused_ffn_kernel<NT, DT>;
  static bool attr_done = false;   // per-instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_cluster(kern, dim3(grid), dim3(CfgU::THREADS), (size_t)SMEM, st, 1, up.early_a != 0 || up.pdl_edge != 0, g, u, bup, d,
                        bdn, up, dn, gbar, up_ctas);
}

// gate/up + SwiGLU and down projection of the routed experts in one persistent kernel (decode regime: nt <= 128, dual gate/up)
cudaError_t launch_fused_ffn(int dtype, int nt, const CUtensorMap& g, const CUtensorMap& u, const CUtensorMap& bup,
                             const CUtensorMap& d, const CUtensorMap& bdn, const GemmParams& up, const GemmParams& dn, int grid,
                             int up_ctas, int* gbar, cudaStream_t st) {
  if (up.E > MAX_E || grid < 1 || up_ctas < 1 || up_ctas > grid || !gbar) return cudaErrorInvalidValue;
#define B2M_FCASE(N)                                                                                         \
  case N:                                                                                                    \
    return dtype == DT_BF16 ? launch_fused<N, DT_BF16>(g, u, bup, d, bdn, up, dn, grid, up_ctas, gbar, st)   \
                            : launch_fused<N, DT_F16>(g, u, bup, d, bdn, up, dn, grid, up_ctas, gbar, st);
  if (dtype != DT_BF16 && dtype != DT_F16) return cudaErrorInvalidValue;
  switch (nt) {
    B2M_FCASE(16)
    B2M_FCASE(32)
    B2M_FCASE(64)
    B2M_FCASE(128)
    default:
      return cudaErrorInvalidValue;
  }
#undef B2M_FCASE
}

#endif  // B2M_ENABLE_FUSED_FFN

template <int DT>
static cudaError_t dispatch_nt(int nt, bool dual, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                               const GemmParams& p, int grid, cudaStream_t st) {
#define B2M_CASE(N)                                                                 \
  case N:                                                                           \
    return dual ? launch_tc<N, true, DT>(a0, a1, b, p, grid, st) : launch_tc<N, false, DT>(a0, a1, b, p, grid, st);
  switch (nt) {
    B2M_CASE(16)
    B2M_CASE(32)
    B2M_CASE(64)
    B2M_CASE(128)
    B2M_CASE(256)
    default:
      return cudaErrorInvalidValue;
  }
#undef B2M_CASE
}

#ifdef B2M_ENABLE_MC2
// 2-CTA multicast variant (prefill: NT = 128, several token tiles per expert); a0/a1 must be the 64-row-box maps.
// Measured: no gain on B200 -> not instantiated in the default build (-DB2M_ENABLE_MC2 brings it back).
cudaError_t launch_grouped_gemm_tc_mc2(int dtype, bool dual, const CUtensorMap& a0h, const CUtensorMap& a1h,
                                       const CUtensorMap& b, const GemmParams& p, int grid, cudaStream_t st) {
  if (p.E > MAX_E) return cudaErrorInvalidValue;
  if (dtype == DT_BF16)
    return dual ? launch_tc<128, true, DT_BF16, 2>(a0h, a1h, b, p, grid, st) : launch_tc<128, false, DT_BF16, 2>(a0h, a1h, b, p, grid, st);
  if (dtype == DT_F16)
    return dual ? launch_tc<128, true, DT_F16, 2>(a0h, a1h, b, p, grid, st) : launch_tc<128, false, DT_F16, 2>(a0h, a1h, b, p, grid, st);
  return cudaErrorInvalidValue;
}
#endif

cudaError_t launch_grouped_gemm_tc(int dtype, int nt, bool dual, const CUtensorMap& a0, const CUtensorMap& a1,
                                   const CUtensorMap& b, const GemmParams& p, int grid, cudaStream_t st) {
  if (p.E > MAX_E) return cudaErrorInvalidValue;
  if (dtype == DT_BF16) return dispatch_nt<DT_BF16>(nt, dual, a0, a1, b, p, grid, st);
  if (dtype == DT_F16) return dispatch_nt<DT_F16>(nt, dual, a0, a1, b, p, grid, st);
  return cudaErrorInvalidValue;
}

cudaError_t launch_grouped_gemm_simt(int dtype, const void* arena, size_t slot_elems, size_t offA0, size_t offA1,
                                     const void* B, int ldb, const GemmParams& p, bool dual, cudaStream_t st) {
  const int grid = 148 * 8;
  if (dtype == DT_BF16)
    grouped_gemm_simt_kernel<DT_BF16><<<grid, 256, 0, st>>>((const uint16_t*)arena, slot_elems, offA0, offA1,
                                                             (const uint16_t*)B, ldb, p, dual ? 1 : 0);
  else if (dtype == DT_F16)
    grouped_gemm_simt_kernel<DT_F16><<<grid, 256, 0, st>>>((const uint16_t*)arena, slot_elems, offA0, offA1,
                                                            (const uint16_t*)B, ldb, p, dual ? 1 : 0);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

int gemm_tc_smem_bytes(int nt, bool dual) {
  switch (nt) {
    case 16: return dual ? GemmCfg<16, true>::SMEM_BYTES : GemmCfg<16, false>::SMEM_BYTES;
    case 32: return dual ? GemmCfg<32, true>::SMEM_BYTES : GemmCfg<32, false>::SMEM_BYTES;
    case 64: return dual ? GemmCfg<64, true>::SMEM_BYTES : GemmCfg<64, false>::SMEM_BYTES;
    case 128: return dual ? GemmCfg<128, true>::SMEM_BYTES : GemmCfg<128, false>::SMEM_BYTES;
    case 256: return dual ? GemmCfg<256, true>::SMEM_BYTES : GemmCfg<256, false>::SMEM_BYTES;
  }
  return -1;
}

}  // namespace b2m
