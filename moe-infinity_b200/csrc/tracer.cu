// tracer.cu -- device-side expert activation tracer + predictor.
//
// Replaces, without a single host round trip in the layer, what the reference does per sequence per layer on the host
// with two device syncs each (moe_infinity/memory/expert_tracer.py:78-125, expert_predictor.py:17-35):
//   P1 update_entry       trace[seq][layer][e] += times expert e was chosen by the sequence's tokens at this layer
//   P2 find_most_similar  rows <= layer of every library trace are masked to 1e-9, all rows are normalised over the experts,
//                         cosine similarity (eps 1e-6) per layer against the sequence's normalised trace, mean over layers,
//                         argmin of 1 - mean (a NaN distance wins, first one: torch.argmin); the winner's access count += 1
//   P3 predict            winner's matrix with past layers zeroed and layer l >= cur scaled (v + 1e-8) * (1 - (l-cur)/(L+1))
// plus the sum of all sequences' predictions (the score matrix ExpertPrefetcher.prefetch_experts sorts,
// memory/expert_prefetcher.py:42-59).  One CTA per sequence; the routing result is read from the workspace the routing
// kernels just wrote (top-k ids), so the call is stream ordered behind them.  fp32 arithmetic like the reference
// (its library lives in fp32 on cuda:0).  finish_entry (:61-76) stores a finished sequence's trace in the library.
#include <math.h>

#include "b2m_common.cuh"
#include "b2m_internal.h"

namespace b2m {

namespace {

constexpr int TR_THREADS = 256;

__global__ void __launch_bounds__(TR_THREADS) trace_update_predict_kernel(const TraceParams p) {
  extern __shared__ float sm[];
  float* m = sm;                       // [L*E] row-normalised trace of this sequence
  float* mnorm = sm + p.L * p.E;       // [L]
  __shared__ int s_cnt[256];
  __shared__ float s_best[TR_THREADS];
  __shared__ int s_besti[TR_THREADS];
  const int b = blockIdx.x;
  const int L = p.L, E = p.E, layer = p.layer;
  float* tr = p.seq + (size_t)(p.seq_slot0 + b) * L * E;
  // ---- P1: this sequence's tokens at this layer
  for (int e = threadIdx.x; e < E; e += TR_THREADS) s_cnt[e] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < p.seq_len * p.k; i += TR_THREADS) {
    const int e = p.topk_idx[(size_t)b * p.seq_len * p.k + i];
    if (e >= 0 && e < E) atomicAdd(&s_cnt[e], 1);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += TR_THREADS) tr[(size_t)layer * E + e] += (float)s_cnt[e];
  __syncthreads();
  // ---- normalised query: m[l] = trace[l] / sum(trace[l]) (0/0 -> 0), |m[l]|
  for (int l = threadIdx.x; l < L; l += TR_THREADS) {
    float s = 0.f;
    for (int e = 0; e < E; ++e) s += tr[(size_t)l * E + e];
    float n2 = 0.f;
    for (int e = 0; e < E; ++e) {
      float v = tr[(size_t)l * E + e] / s;
      if (isnan(v) || isinf(v)) v = 0.f;                     // np.nan_to_num
      m[l * E + e] = v;
      n2 += v * v;
    }
    mnorm[l] = fmaxf(sqrtf(n2), 1e-6f);
  }
  __syncthreads();
  // ---- P2: distance to every library trace (thread per entry)
  float best = INFINITY;
  int besti = 0x7fffffff;
  // the masked rows of every entry are the same constant row: 1e-9 / (E * 1e-9)
  float cm = 0.f;
  for (int e = 0; e < E; ++e) cm += 1e-9f;
  const float cval = 1e-9f / cm;
  float cn2 = 0.f;
  for (int e = 0; e < E; ++e) cn2 += cval * cval;
  const float cnorm = fmaxf(sqrtf(cn2), 1e-6f);
  for (int c = threadIdx.x; c < p.capacity; c += TR_THREADS) {
    const float* lib = p.lib + (size_t)c * L * E;
    float acc = 0.f;
    for (int l = 0; l < L; ++l) {
      float num = 0.f, nrm;
      if (l <= layer) {
        for (int e = 0; e < E; ++e) num += m[l * E + e] * cval;
        nrm = cnorm;
      } else {
        float s = 0.f;
        for (int e = 0; e < E; ++e) s += lib[(size_t)l * E + e];
        float n2 = 0.f;
        for (int e = 0; e < E; ++e) {
          const float v = lib[(size_t)l * E + e] / s;          // 0/0 -> NaN propagates like the reference (Q8)
          num += m[l * E + e] * v;
          n2 += v * v;
        }
        nrm = fmaxf(sqrtf(n2), 1e-6f);
      }
      acc += num / (mnorm[l] * nrm);
    }
    float dist = 1.0f - acc / (float)L;
    if (isnan(dist)) dist = -INFINITY;                          // torch.argmin: NaN is the minimum, first one wins
    if (dist < best) { best = dist; besti = c; }
  }
  s_best[threadIdx.x] = best;
  s_besti[threadIdx.x] = besti;
  __syncthreads();
  for (int d = TR_THREADS / 2; d; d >>= 1) {
    if (threadIdx.x < d) {
      const float ov = s_best[threadIdx.x + d];
      const int oi = s_besti[threadIdx.x + d];
      if (ov < s_best[threadIdx.x] || (ov == s_best[threadIdx.x] && oi < s_besti[threadIdx.x])) {
        s_best[threadIdx.x] = ov;
        s_besti[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  const int win = s_besti[0];
  if (threadIdx.x == 0) {
    atomicAdd(p.access + win, 1);
    p.winner[p.seq_slot0 + b] = win;
  }
  // ---- P3: decayed prediction, and its contribution to the batch's hint matrix
  const float* w = p.lib + (size_t)win * L * E;
  float* pred = p.pred + (size_t)(p.seq_slot0 + b) * L * E;
  for (int i = threadIdx.x; i < L * E; i += TR_THREADS) {
    const int l = i / E;
    float v = 0.f;
    if (l >= layer) {
      const float decay = (float)(-1.0 / (double)(L + 1) * (double)(l - layer) + 1.0);
      v = (w[i] + 1e-8f) * decay;
    }
    pred[i] = v;
    if (v != 0.f) atomicAdd(p.hint + i, v);
  }
}

// finish_entry (expert_tracer.py:61-76): first all-zero library row, else the least accessed non-persistent row
__global__ void __launch_bounds__(TR_THREADS) trace_finish_kernel(const TraceParams p, int seq_slot) {
  __shared__ int s_zero;
  __shared__ int s_min, s_mini;
  const int L = p.L, E = p.E;
  if (threadIdx.x == 0) { s_zero = 0x7fffffff; s_min = 0x7fffffff; s_mini = 0x7fffffff; }
  __syncthreads();
  for (int c = threadIdx.x; c < p.capacity; c += TR_THREADS) {
    float s = 0.f;
    for (int i = 0; i < L * E; ++i) s += p.lib[(size_t)c * L * E + i];
    if (s == 0.f) atomicMin(&s_zero, c);
  }
  __syncthreads();
  if (s_zero == 0x7fffffff) {
    for (int c = threadIdx.x; c < p.capacity; c += TR_THREADS)
      if (c >= p.persistent) atomicMin(&s_min, p.access[c]);
    __syncthreads();
    for (int c = threadIdx.x; c < p.capacity; c += TR_THREADS)
      if (c >= p.persistent && p.access[c] == s_min) atomicMin(&s_mini, c);
    __syncthreads();
  }
  const int idx = s_zero != 0x7fffffff ? s_zero : s_mini;
  if (idx == 0x7fffffff) return;            // library full of persistent entries: nothing to replace
  for (int i = threadIdx.x; i < L * E; i += TR_THREADS)
    p.lib[(size_t)idx * L * E + i] = p.seq[(size_t)seq_slot * L * E + i];
  if (threadIdx.x == 0) p.access[idx] = 1;
}

}  // namespace

cudaError_t launch_trace_update_predict(const TraceParams& p, int num_seqs, cudaStream_t st) {
  if (num_seqs < 1 || p.E > 256 || p.L * p.E > 8192) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(p.L * p.E + p.L) * sizeof(float);
  trace_update_predict_kernel<<<num_seqs, TR_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_trace_finish(const TraceParams& p, int seq_slot, cudaStream_t st) {
  trace_finish_kernel<<<1, TR_THREADS, 0, st>>>(p, seq_slot);
  return cudaGetLastError();
}

}  // namespace b2m
