// f32_path.cu -- fp32 experts (dtype int 1 of the boundary, core/parallel/expert_module.h:20-23): grouped expert GEMMs and
// the combine for contexts whose model dtype is float32 (Switch-base-8's default dtype, BASELINE config 1).
//
// The reference computes fp32 experts with torch::matmul on fp32 tensors (expert_module.cpp:31-35 and the other five
// forward()s) -- cuBLAS SGEMM with TF32 off (torch's default), i.e. true fp32 multiply-add.  kind::tf32 tensor-core MMAs
// would drop 13 mantissa bits of every operand, so this path stays on the CUDA cores: a shared-memory tiled fp32 FMA
// kernel, segmented by the same device-side per-expert offsets as the tensor-core kernels (no host sync).  It is the
// plumbing/correctness configuration of BASELINE.json, not a tuned kernel: 64 weight rows x 32 tokens per tile,
// 8 outputs per thread.
#include "b2m_common.cuh"
#include "b2m_internal.h"

namespace b2m {

namespace {

constexpr int F_BM = 64;      // weight rows per tile
constexpr int F_BN = 32;      // tokens per tile (one per lane)
constexpr int F_BK = 32;      // reduction chunk
constexpr int F_THREADS = 256;
constexpr int F_MAX_E = 256;

__device__ __forceinline__ float act_f32(float x, int act) {
  if (act == ACT_SILU) return x / (1.0f + expf(-x));
  if (act == ACT_RELU) return fmaxf(x, 0.0f);
  if (act == ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  return x;
}

// out[row, m] = epilogue( sum_k A0[slot][m, k] * B[row, k]  (, sum_k A1[slot][m, k] * B[row, k]) )
template <bool DUAL>
__global__ void __launch_bounds__(F_THREADS) grouped_gemm_f32_kernel(const float* __restrict__ arena, size_t slot_elems,
                                                                      size_t offA0, size_t offA1, const float* __restrict__ B,
                                                                      int ldb, const GemmParams p) {
  __shared__ float sA0[F_BM][F_BK + 1];
  __shared__ float sA1[DUAL ? F_BM : 1][F_BK + 1];
  __shared__ float sB[F_BN][F_BK + 1];
  __shared__ int tile_start[F_MAX_E + 1];
  __shared__ int s_off[F_MAX_E + 1];
  __shared__ int s_slot[F_MAX_E];
  const int E = p.E;
  const int m_tiles = (p.M + F_BM - 1) / F_BM;
  if (p.single_n >= 0) {
    if (threadIdx.x == 0) { s_off[0] = 0; s_off[1] = p.single_n; s_slot[0] = p.single_slot; }
  } else {
    for (int i = threadIdx.x; i <= E; i += F_THREADS) s_off[i] = p.offsets[i];
    for (int i = threadIdx.x; i < E; i += F_THREADS) s_slot[i] = p.slot_of[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < E; ++e) {
      tile_start[e] = acc;
      const int n_e = s_off[e + 1] - s_off[e];
      if (n_e > 0 && s_slot[e] >= 0) acc += m_tiles * ((n_e + F_BN - 1) / F_BN);
    }
    tile_start[E] = acc;
  }
  __syncthreads();
  const int total = tile_start[E];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int e = 0;
    while (e + 1 < E && tile >= tile_start[e + 1]) ++e;      // E <= 256, tiles of empty experts have zero width
    const int local = tile - tile_start[e];
    const int mt = local % m_tiles, nt = local / m_tiles;
    const int m0 = mt * F_BM, row0 = s_off[e] + nt * F_BN;
    const int ncols = min(F_BN, s_off[e + 1] - row0);
    const float* A0 = arena + (size_t)s_slot[e] * slot_elems + offA0;
    const float* A1 = arena + (size_t)s_slot[e] * slot_elems + offA1;
    float acc0[8], acc1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
    for (int k0 = 0; k0 < p.K; k0 += F_BK) {
      // stage: 64x32 weights (8 per thread), 32x32 tokens (4 per thread); coalesced along k
      for (int i = threadIdx.x; i < F_BM * F_BK; i += F_THREADS) {
        const int r = i / F_BK, c = i % F_BK;
        const bool ok = m0 + r < p.M && k0 + c < p.K;
        sA0[r][c] = ok ? A0[(size_t)(m0 + r) * p.K + k0 + c] : 0.f;
        if (DUAL) sA1[r][c] = ok ? A1[(size_t)(m0 + r) * p.K + k0 + c] : 0.f;
      }
      for (int i = threadIdx.x; i < F_BN * F_BK; i += F_THREADS) {
        const int r = i / F_BK, c = i % F_BK;
        sB[r][c] = (r < ncols && k0 + c < p.K) ? B[(size_t)(row0 + r) * ldb + k0 + c] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < F_BK; ++kk) {
        const float b = sB[lane][kk];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc0[i] = fmaf(sA0[warp * 8 + i][kk], b, acc0[i]);          // broadcast read of the weight value
          if (DUAL) acc1[i] = fmaf(sA1[warp * 8 + i][kk], b, acc1[i]);
        }
      }
      __syncthreads();
    }
    if (lane < ncols) {
      const int row = row0 + lane;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + warp * 8 + i;
        if (m >= p.M) continue;
        float bias = 0.f;
        if (p.bias_base) bias = reinterpret_cast<const float*>(p.bias_base)[(size_t)s_slot[e] * p.bias_slot_elems + p.bias_off + m];
        float v;
        if (p.epi == EPI_LINEAR_F32) {
          v = acc0[i] + bias;
        } else if (DUAL) {
          v = act_f32(acc0[i], p.act) * acc1[i];
        } else {
          v = act_f32(acc0[i] + bias, p.act);
        }
        reinterpret_cast<float*>(p.out)[(size_t)row * p.ld_out + m] = v;
      }
    }
  }
}

// un-permute + weighted combine for fp32 models.  One thread per 4 hidden elements; experts applied in ascending expert
// id (the oracle's fixed order).  With an fp32 model every "round to model dtype" of the reference chain is the
// identity, so the modes differ only in operation order: separate multiply and add (no FMA contraction), as ATen does.
__global__ void __launch_bounds__(256) combine_f32_kernel(const CombineParams p) {
  const int t = blockIdx.x;
  const int k = p.k;
  __shared__ int s_row[8];
  __shared__ float s_w[8];
  __shared__ int s_n;
  if (threadIdx.x == 0) {
    int n = 0;
    int es[8], rows[8];
    float ws[8];
    for (int j = 0; j < k; ++j) {
      const int e = p.topk_idx[(size_t)t * k + j], r = p.row_of[(size_t)t * k + j];
      if (e < 0 || r < 0) continue;
      int pos = n++;
      while (pos > 0 && es[pos - 1] > e) { es[pos] = es[pos - 1]; rows[pos] = rows[pos - 1]; ws[pos] = ws[pos - 1]; --pos; }
      es[pos] = e; rows[pos] = r; ws[pos] = p.topk_w[(size_t)t * k + j];
    }
    for (int j = 0; j < n; ++j) { s_row[j] = rows[j]; s_w[j] = ws[j]; }
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  const int H = p.H;
  float* out = reinterpret_cast<float*>(p.out);
  const float* x = reinterpret_cast<const float*>(p.x);
  for (int h = (blockIdx.y * 256 + threadIdx.x) * 4; h < H; h += gridDim.y * 256 * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < n; ++j) {
      const float4 y = *reinterpret_cast<const float4*>(p.y + (size_t)s_row[j] * H + h);
      const float yv[4] = {y.x, y.y, y.z, y.w};
      const float w = s_w[j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (p.mode == COMBINE_SWITCH) acc[i] = yv[i];
        else if (p.mode == COMBINE_FP32) acc[i] = fmaf(yv[i], w, acc[i]);
        else acc[i] = __fadd_rn(acc[i], __fmul_rn(yv[i], w));
      }
    }
    if (p.mode == COMBINE_SWITCH) {
      const float prob = p.topk_w[(size_t)t * k];
      const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)t * H + h);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = __fmul_rn(prob, n > 0 ? acc[i] : xs[i]);
    }
    if (p.y_shared) {
      const float4 s = *reinterpret_cast<const float4*>(p.y_shared + (size_t)t * H + h);
      acc[0] = __fadd_rn(acc[0], s.x); acc[1] = __fadd_rn(acc[1], s.y);
      acc[2] = __fadd_rn(acc[2], s.z); acc[3] = __fadd_rn(acc[3], s.w);
    }
    *reinterpret_cast<float4*>(out + (size_t)t * H + h) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

}  // namespace

cudaError_t launch_grouped_gemm_f32(const void* arena, size_t slot_elems, size_t offA0, size_t offA1, const void* B,
                                    int ldb, const GemmParams& p, bool dual, int num_sms, cudaStream_t st) {
  if (p.E > F_MAX_E) return cudaErrorInvalidValue;
  const int grid = num_sms * 4;
  if (dual)
    grouped_gemm_f32_kernel<true><<<grid, F_THREADS, 0, st>>>((const float*)arena, slot_elems, offA0, offA1, (const float*)B, ldb, p);
  else
    grouped_gemm_f32_kernel<false><<<grid, F_THREADS, 0, st>>>((const float*)arena, slot_elems, offA0, offA1, (const float*)B, ldb, p);
  return cudaGetLastError();
}

cudaError_t launch_combine_f32(const CombineParams& p, cudaStream_t st) {
  if (p.T == 0) return cudaSuccess;
  if (p.k > 8 || p.H % 4 != 0 || p.ep_collect) return cudaErrorInvalidValue;
  const int gy = (p.H + 256 * 4 - 1) / (256 * 4);
  combine_f32_kernel<<<dim3(p.T, gy), 256, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace b2m
