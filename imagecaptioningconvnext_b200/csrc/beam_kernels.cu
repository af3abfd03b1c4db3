// beam_kernels.cu — batched beam search bookkeeping on the device (one CTA per image).
//
// Semantics follow caption.py:96-155 (LSTM) and caption.py:197-251 (Transformer) exactly, per image:
//   scores = log_softmax(logits) + top_k_scores (broadcast per beam)                       caption.py:105-107,217-219
//   step 1: top-k over beam 0 only; later: flat top-k over (alive beams x V), sorted        caption.py:110-113,221-224
//   prev = idx / V, next = idx % V                                                          caption.py:116-118
//   beams whose next word is <end> move to the completed list, k shrinks                    caption.py:125-133
//   surviving beams are compacted to the front in candidate order                           caption.py:138-145
// The reference runs ONE image with beams-as-batch on the CPU; here NI images x k beams run as NI*k rows and
// every image keeps its own (scores, seqs, k_remaining) state.
#include "ccx_common.cuh"
#include "ccx_ops.h"
#include "ccx_prof.h"

namespace ccx {

static constexpr int BEAM_KMAX = 8;
static constexpr int BEAM_THREADS = 256;

struct Cand {
  float s;
  int idx;  // flat index beam * V + word
};
__device__ __forceinline__ bool cand_better(float s, int i, float s2, int i2) {
  return s > s2 || (s == s2 && i < i2);
}

__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < BEAM_THREADS / 32; ++w) r = fmaxf(r, red[w]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < BEAM_THREADS / 32; ++w) r += red[w];
  return r;
}

__global__ void __launch_bounds__(BEAM_THREADS)
beam_topk_kernel(const float* __restrict__ logits, long long ld, int V, int k,
                 const float* __restrict__ top_scores,  // [NI, k]
                 const int* __restrict__ k_rem,         // [NI]
                 int first_step,
                 float* __restrict__ cand_score, int* __restrict__ cand_prev, int* __restrict__ cand_word) {
  __shared__ float red[BEAM_THREADS / 32];
  __shared__ float s_lse_mx[BEAM_KMAX], s_lse_lg[BEAM_KMAX];
  __shared__ float s_ws[BEAM_THREADS / 32];
  __shared__ int s_wi[BEAM_THREADS / 32], s_wt[BEAM_THREADS / 32];
  __shared__ int s_winner_thread;
  const int img = blockIdx.x;
  const int kr = k_rem[img];
  if (kr <= 0) return;
  const int nb = first_step ? 1 : kr;   // beams that contribute candidates
  // log-softmax statistics per contributing beam
  for (int j = 0; j < nb; ++j) {
    const float* row = logits + (static_cast<long long>(img) * k + j) * ld;
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < V; v += BEAM_THREADS) mx = fmaxf(mx, row[v]);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int v = threadIdx.x; v < V; v += BEAM_THREADS) sum += expf(row[v] - mx);
    sum = block_sum(sum, red);
    if (threadIdx.x == 0) { s_lse_mx[j] = mx; s_lse_lg[j] = logf(sum); }
  }
  __syncthreads();
  // per-thread sorted top-kr list
  Cand best[BEAM_KMAX];
#pragma unroll
  for (int i = 0; i < BEAM_KMAX; ++i) { best[i].s = -INFINITY; best[i].idx = 0x7fffffff; }
  for (int j = 0; j < nb; ++j) {
    const float* row = logits + (static_cast<long long>(img) * k + j) * ld;
    const float base = top_scores[img * k + j];
    const float mx = s_lse_mx[j], lg = s_lse_lg[j];
    for (int v = threadIdx.x; v < V; v += BEAM_THREADS) {
      const float s = base + ((row[v] - mx) - lg);
      const int idx = j * V + v;
      if (cand_better(s, idx, best[BEAM_KMAX - 1].s, best[BEAM_KMAX - 1].idx)) {
        best[BEAM_KMAX - 1].s = s;
        best[BEAM_KMAX - 1].idx = idx;
#pragma unroll
        for (int i = BEAM_KMAX - 1; i > 0; --i) {
          if (cand_better(best[i].s, best[i].idx, best[i - 1].s, best[i - 1].idx)) {
            const Cand t = best[i]; best[i] = best[i - 1]; best[i - 1] = t;
          }
        }
      }
    }
  }
  // kr rounds of block-wide arg-best over the list heads
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < kr; ++r) {
    float s = best[0].s;
    int idx = best[0].idx;
    int who = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, s, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      const int ow = __shfl_xor_sync(0xffffffffu, who, o);
      if (cand_better(os, oi, s, idx)) { s = os; idx = oi; who = ow; }
    }
    if (lane == 0) { s_ws[warp] = s; s_wi[warp] = idx; s_wt[warp] = who; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < BEAM_THREADS / 32; ++w)
        if (cand_better(s_ws[w], s_wi[w], s, idx)) { s = s_ws[w]; idx = s_wi[w]; who = s_wt[w]; }
      cand_score[img * k + r] = s;
      cand_prev[img * k + r] = idx / V;
      cand_word[img * k + r] = idx % V;
      s_winner_thread = who;
    }
    __syncthreads();
    if (threadIdx.x == s_winner_thread) {
#pragma unroll
      for (int i = 0; i < BEAM_KMAX - 1; ++i) best[i] = best[i + 1];
      best[BEAM_KMAX - 1].s = -INFINITY;
      best[BEAM_KMAX - 1].idx = 0x7fffffff;
    }
    __syncthreads();
  }
}

int beam_topk(const float* logits, long long ld, int NI, int k, int V, const float* top_scores, const int* k_rem,
              int first_step, float* cand_score, int* cand_prev, int* cand_word, cudaStream_t stream) {
  if (NI <= 0) return CCX_OK;
  if (k <= 0 || k > BEAM_KMAX || V <= 0 || static_cast<long long>(k) * V > 0x7fffffffLL) return CCX_ERR_SHAPE;
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)NI * k * V * 4.0 * 3.0);
  beam_topk_kernel<<<NI, BEAM_THREADS, 0, stream>>>(logits, ld, V, k, top_scores, k_rem, first_step, cand_score,
                                                    cand_prev, cand_word);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// one thread block per image, serial over <= 8 candidates: pure bookkeeping
__global__ void __launch_bounds__(64)
beam_update_kernel(int k, int Tcap, int step, long long end_token, const float* __restrict__ cand_score,
                   const int* __restrict__ cand_prev, const int* __restrict__ cand_word,
                   const long long* __restrict__ seqs_in, long long* __restrict__ seqs_out,  // [NI, k, Tcap]
                   float* __restrict__ top_scores, int* __restrict__ k_rem,
                   long long* __restrict__ done_seqs, float* __restrict__ done_scores, int* __restrict__ done_len,
                   int* __restrict__ n_done,   // [NI, k, Tcap], [NI, k], [NI, k], [NI]
                   int* __restrict__ src_row,  // [NI * k] global parent row of each surviving beam
                   long long* __restrict__ next_tok, long long ld_next,
                   int* __restrict__ done_parent) {  // [NI, k] optional: global parent row of each completed sequence
  const int img = blockIdx.x;
  const int kr = k_rem[img];
  __shared__ int s_slot[BEAM_KMAX];   // destination: >= 0 alive slot, < 0: -(done index + 1)
  if (threadIdx.x == 0) {
    int alive = 0, nd = n_done[img];
    for (int c = 0; c < kr; ++c) {
      if (cand_word[img * k + c] == end_token) {
        s_slot[c] = -(nd + 1);
        done_scores[img * k + nd] = cand_score[img * k + c];
        done_len[img * k + nd] = step + 1;   // tokens incl. <start> and <end>
        if (done_parent != nullptr) done_parent[img * k + nd] = img * k + cand_prev[img * k + c];
        ++nd;
      } else {
        s_slot[c] = alive;
        top_scores[img * k + alive] = cand_score[img * k + c];
        src_row[img * k + alive] = img * k + cand_prev[img * k + c];
        next_tok[(static_cast<long long>(img) * k + alive) * ld_next] = cand_word[img * k + c];
        ++alive;
      }
    }
    for (int a = alive; a < k; ++a) src_row[img * k + a] = img * k + a;
    n_done[img] = nd;
    k_rem[img] = alive;
  }
  __syncthreads();
  // sequences: seq_new[c] = seqs_in[prev[c]][0..step) + word   (step tokens so far incl. <start>)
  for (int c = 0; c < kr; ++c) {
    const int slot = s_slot[c];
    const long long* src = seqs_in + (static_cast<long long>(img) * k + cand_prev[img * k + c]) * Tcap;
    long long* dst = slot >= 0 ? seqs_out + (static_cast<long long>(img) * k + slot) * Tcap
                               : done_seqs + (static_cast<long long>(img) * k + (-slot - 1)) * Tcap;
    for (int t = threadIdx.x; t < step; t += blockDim.x) dst[t] = src[t];
    if (threadIdx.x == 0) dst[step] = cand_word[img * k + c];
  }
}

int beam_update(int NI, int k, int Tcap, int step, long long end_token, const float* cand_score,
                const int* cand_prev, const int* cand_word, const long long* seqs_in, long long* seqs_out,
                float* top_scores, int* k_rem, long long* done_seqs, float* done_scores, int* done_len, int* n_done,
                int* src_row, long long* next_tok, long long ld_next, int* done_parent, cudaStream_t stream) {
  if (NI <= 0) return CCX_OK;
  if (k <= 0 || k > BEAM_KMAX || step < 1 || step >= Tcap) return CCX_ERR_SHAPE;
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)NI * k * Tcap * 16.0);
  beam_update_kernel<<<NI, 64, 0, stream>>>(k, Tcap, step, end_token, cand_score, cand_prev, cand_word, seqs_in,
                                            seqs_out, top_scores, k_rem, done_seqs, done_scores, done_len, n_done,
                                            src_row, next_tok, ld_next, done_parent);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// dst[r, 0:row_bytes) = src[src_row[r], 0:row_bytes)   (16-byte granules; strides in bytes)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint8_t* __restrict__ src, long long src_stride, uint8_t* __restrict__ dst,
                   long long dst_stride, const int* __restrict__ src_row, int row_granules, int rows) {
  const int r = blockIdx.y;
  const long long s = src_row ? src_row[r] : r;
  const uint4* sp = reinterpret_cast<const uint4*>(src + s * src_stride);
  uint4* dp = reinterpret_cast<uint4*>(dst + r * dst_stride);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_granules; i += gridDim.x * blockDim.x) dp[i] = sp[i];
}

int gather_rows(const void* src, long long src_stride, void* dst, long long dst_stride, const int* src_row,
                long long row_bytes, int rows, cudaStream_t stream) {
  if (rows <= 0 || row_bytes <= 0) return CCX_OK;
  if ((row_bytes % 16) || (src_stride % 16) || (dst_stride % 16) || rows > 65535 ||
      (reinterpret_cast<uintptr_t>(src) % 16) || (reinterpret_cast<uintptr_t>(dst) % 16))
    return CCX_ERR_SHAPE;
  const int gran = static_cast<int>(row_bytes / 16);
  int gx = (gran + 255) / 256;
  if (gx > 64) gx = 64;
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)rows * row_bytes * 2.0);
  gather_rows_kernel<<<dim3(gx, rows), 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(src), src_stride,
                                                        reinterpret_cast<uint8_t*>(dst), dst_stride, src_row, gran,
                                                        rows);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

}  // namespace ccx
