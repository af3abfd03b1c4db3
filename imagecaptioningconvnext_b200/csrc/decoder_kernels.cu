// decoder_kernels.cu — the non-GEMM kernels of the two caption decoders (forward / inference side).
//
//   embed_rows          nn.Embedding (+ dropout mask + positional encoding)   models/decoder.py:84, models/transformerDecoder.py:97-98,25-26
//   mean_pixels         encoder_out.mean(dim=1)                               models/decoder.py:64
//   bahdanau_attention  relu(att1+att2) . w_f -> softmax over pixels -> weighted sum -> * sigmoid(f_beta h)
//                                                                             models/decoder.py:25-31,104-105
//   lstm_pointwise      nn.LSTMCell gate math (i,f,g,o)                        torch/nn/modules/rnn.py:1755-1778
//   greedy_next         argmax over the vocabulary + finished bookkeeping     models/decoder.py:156-159, transformerDecoder.py:146-148
//   mha_small           scaled-dot-product attention for T<=~200 keys          torch/nn/functional.py multi_head_attention_forward
//
// "Operand" outputs feed the next tcgen05 GEMM: bf16, or an fp32 (tf32-hi, lo) pair for the 3xTF32 path.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "ccx_common.cuh"
#include "ccx_ops.h"
#include "ccx_prof.h"

namespace cg = cooperative_groups;

namespace ccx {

struct OpOut {
  void* hi;   // bf16* or float*
  float* lo;  // fp32 remainder, or nullptr (plain fp32 / bf16)
  int dtype;  // CCX_F32 / CCX_BF16
};

__device__ __forceinline__ void store_op(const OpOut& o, long long idx, float v) {
  if (o.dtype == CCX_BF16) {
    reinterpret_cast<__nv_bfloat16*>(o.hi)[idx] = __float2bfloat16_rn(v);
  } else if (o.lo != nullptr) {
    const float h = tf32_hi(v);
    reinterpret_cast<float*>(o.hi)[idx] = h;
    o.lo[idx] = v - h;
  } else {
    reinterpret_cast<float*>(o.hi)[idx] = v;
  }
}
__device__ __forceinline__ void store_op4(const OpOut& o, long long idx, float4 v) {
  if (o.dtype == CCX_BF16) {
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(o.hi) + idx) = pk;
  } else if (o.lo != nullptr) {
    float4 h, l;
    h.x = tf32_hi(v.x); l.x = v.x - h.x;
    h.y = tf32_hi(v.y); l.y = v.y - h.y;
    h.z = tf32_hi(v.z); l.z = v.z - h.z;
    h.w = tf32_hi(v.w); l.w = v.w - h.w;
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(o.hi) + idx) = h;
    *reinterpret_cast<float4*>(o.lo + idx) = l;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(o.hi) + idx) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// embedding gather (+ dropout multiplier, + positional encoding); one warp per (b, t) row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_rows_kernel(const long long* __restrict__ tokens, long long tok_ld, int t0, const float* __restrict__ table,
                  int V, int D, const float* __restrict__ pe, const float* __restrict__ dropmask,
                  float* __restrict__ out_plain, long long sb_p, long long st_p, OpOut op, long long sb_o,
                  long long st_o, int nb, int nt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (r >= static_cast<long long>(nb) * nt) return;
  const int b = static_cast<int>(r / nt), t = static_cast<int>(r % nt);
  long long tok = tokens[b * tok_ld + t0 + t];
  if (tok < 0) tok = 0;
  if (tok >= V) tok = V - 1;
  const float4* src = reinterpret_cast<const float4*>(table + tok * D);
  for (int i = lane; i < D / 4; i += 32) {
    float4 v = __ldg(src + i);
    if (dropmask != nullptr) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(dropmask + r * D) + i);
      v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
    }
    if (pe != nullptr) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(pe + static_cast<long long>(t0 + t) * D) + i);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    if (out_plain != nullptr) *reinterpret_cast<float4*>(out_plain + b * sb_p + t * st_p + i * 4) = v;
    if (op.hi != nullptr) store_op4(op, b * sb_o + t * st_o + i * 4, v);
  }
}

int embed_rows(const long long* tokens, long long tok_ld, int t0, const float* table, int V, int D,
               const float* pe, const float* dropmask, float* out_plain, long long sb_p, long long st_p,
               void* op_hi, float* op_lo, int op_dtype, long long sb_o, long long st_o, int nb, int nt,
               cudaStream_t stream) {
  if (nb <= 0 || nt <= 0) return CCX_OK;
  if (D % 4 != 0 || V <= 0) return CCX_ERR_SHAPE;
  OpOut op{op_hi, op_lo, op_dtype};
  const long long rows = static_cast<long long>(nb) * nt;
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)rows * D * 8.0);
  embed_rows_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      tokens, tok_ld, t0, table, V, D, pe, dropmask, out_plain, sb_p, st_p, op, sb_o, st_o, nb, nt);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// mean over pixels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mean_pixels_kernel(const float* __restrict__ enc, int P, int E, OpOut op, long long ldo, int B) {
  const int e4 = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (e4 >= E / 4) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* src = reinterpret_cast<const float4*>(enc + static_cast<long long>(b) * P * E) + e4;
  for (int p = 0; p < P; ++p) {
    const float4 v = __ldg(src + static_cast<long long>(p) * (E / 4));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float inv = 1.0f / static_cast<float>(P);
  acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  store_op4(op, static_cast<long long>(b) * ldo + e4 * 4, acc);
}

int mean_pixels(const float* enc, int B, int P, int E, void* op_hi, float* op_lo, int op_dtype, long long ldo,
                cudaStream_t stream) {
  if (B <= 0) return CCX_OK;
  if (E % 4 != 0 || P <= 0 || B > 65535) return CCX_ERR_SHAPE;
  OpOut op{op_hi, op_lo, op_dtype};
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)B * P * E * 4.0);
  dim3 grid((E / 4 + 255) / 256, B);
  mean_pixels_kernel<<<grid, 256, 0, stream>>>(enc, P, E, op, ldo, B);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// Bahdanau attention step: one CTA of 1024 threads per sample.  The kernel sits on the recurrence's critical path
// ~50 times per caption batch with only bt <= 32 CTAs, so it is latency- not bandwidth-bound: 32 warps take two
// pixels each for the scores (instead of 8 warps x 7 dependent round trips), and the weighted pixel sum is split
// four ways over P with a shared-memory combine.
// ---------------------------------------------------------------------------------------------
static constexpr int ATT_MAX_P = 256;
static constexpr int ATT_THREADS = 1024;
static constexpr int ATT_PSPLIT = 4;

__global__ void __launch_bounds__(ATT_THREADS)
bahdanau_attention_kernel(const float* __restrict__ att1,  // [B, P, A]
                          const float* __restrict__ hg, long long ldhg,  // [bt, >= A+E]: att2 | gate pre-activation
                          const float* __restrict__ w_f, const float* __restrict__ b_f,
                          const float* __restrict__ enc,   // [B, P, E]
                          const float* __restrict__ active,  // [bt] 1/0 or nullptr
                          float* __restrict__ alpha_out, long long alpha_ld,  // row b at alpha_out + b*alpha_ld
                          OpOut awe, long long ld_awe,     // row b at b*ld_awe (column offset folded into pointers)
                          int P, int A, int E, int apply_gate, int enc_group) {
  extern __shared__ __align__(16) float att_sm[];
  float* s_att2 = att_sm;            // [A]
  float* s_wf = att_sm + A;          // [A]
  float* s_part = att_sm + 2 * A;    // [ATT_PSPLIT][E] partial weighted sums
  __shared__ float s_e[ATT_MAX_P];
  const int b = blockIdx.x;
  const int be = b / enc_group;  // row of att1 / enc this decode row reads (beam search: beams share the image)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < A; i += ATT_THREADS) {
    s_att2[i] = hg[b * ldhg + i];
    s_wf[i] = __ldg(w_f + i);
  }
  __syncthreads();
  const float bf = b_f ? __ldg(b_f) : 0.f;
  for (int p = warp; p < P; p += ATT_THREADS / 32) {
    const float4* a1 = reinterpret_cast<const float4*>(att1 + (static_cast<long long>(be) * P + p) * A);
    float acc = 0.f;
    for (int i = lane; i < A / 4; i += 32) {
      const float4 v = __ldg(a1 + i);
      const float4 h = *reinterpret_cast<const float4*>(s_att2 + i * 4);
      const float4 w = *reinterpret_cast<const float4*>(s_wf + i * 4);
      acc = fmaf(fmaxf(v.x + h.x, 0.f), w.x, acc);
      acc = fmaf(fmaxf(v.y + h.y, 0.f), w.y, acc);
      acc = fmaf(fmaxf(v.z + h.z, 0.f), w.z, acc);
      acc = fmaf(fmaxf(v.w + h.w, 0.f), w.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_e[p] = acc + bf;
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int p = lane; p < P; p += 32) mx = fmaxf(mx, s_e[p]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int p = lane; p < P; p += 32) {
      const float ex = expf(s_e[p] - mx);
      s_e[p] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    const bool wr = (alpha_out != nullptr) && (active == nullptr || active[b] != 0.f);
    for (int p = lane; p < P; p += 32) {
      const float a = s_e[p] * inv;
      s_e[p] = a;
      if (wr) alpha_out[b * alpha_ld + p] = a;
    }
  }
  __syncthreads();
  // awe[e] = sum_p alpha_p enc[p, e]: thread group q (of ATT_PSPLIT) sums pixels [q*pp, (q+1)*pp) in pixel order,
  // group 0 then adds the partials in group order (a fixed summation order: results do not depend on timing)
  constexpr int GT = ATT_THREADS / ATT_PSPLIT;
  const int q = threadIdx.x / GT, tq = threadIdx.x % GT;
  const int pp = (P + ATT_PSPLIT - 1) / ATT_PSPLIT;
  const int pa = q * pp, pb = min(P, pa + pp);
  for (int e4 = tq; e4 < E / 4; e4 += GT) {
    const float4* src = reinterpret_cast<const float4*>(enc + static_cast<long long>(be) * P * E) + e4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = pa; p < pb; ++p) {
      const float4 v = __ldg(src + static_cast<long long>(p) * (E / 4));
      const float a = s_e[p];
      acc.x = fmaf(v.x, a, acc.x); acc.y = fmaf(v.y, a, acc.y);
      acc.z = fmaf(v.z, a, acc.z); acc.w = fmaf(v.w, a, acc.w);
    }
    // partial of group q -> slot q-1; group 0 parks its own in the last slot
    const int slot = q > 0 ? q - 1 : ATT_PSPLIT - 1;
    *reinterpret_cast<float4*>(s_part + static_cast<long long>(slot) * E + e4 * 4) = acc;
  }
  __syncthreads();
  if (q == 0) {
    for (int e4 = tq; e4 < E / 4; e4 += GT) {
      float4 acc = *reinterpret_cast<const float4*>(s_part + static_cast<long long>(ATT_PSPLIT - 1) * E + e4 * 4);
#pragma unroll
      for (int g = 0; g < ATT_PSPLIT - 1; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(s_part + static_cast<long long>(g) * E + e4 * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      if (apply_gate) {
        const float4 gp = *reinterpret_cast<const float4*>(hg + b * ldhg + A + e4 * 4);
        acc.x *= 1.0f / (1.0f + expf(-gp.x));
        acc.y *= 1.0f / (1.0f + expf(-gp.y));
        acc.z *= 1.0f / (1.0f + expf(-gp.z));
        acc.w *= 1.0f / (1.0f + expf(-gp.w));
      }
      store_op4(awe, b * ld_awe + e4 * 4, acc);
    }
  }
}

// Cluster version (the default when E % 16 == 0): the sample is spread over a cluster of 4 CTAs, i.e. 4x as many SMs
// for the <= 32 rows of a recurrent step.  Rank r scores pixels r, r+4, ... and publishes each e_p into the shared
// memory of all four CTAs (DSMEM stores), one cluster barrier, every CTA runs the 49-element softmax redundantly
// (same arithmetic -> same bits), then rank r produces channels [r*E/4, (r+1)*E/4) of the gated context vector.
static constexpr int ATC = 4;

__global__ void __cluster_dims__(ATC, 1, 1) __launch_bounds__(256)
bahdanau_attention_cluster_kernel(const float* __restrict__ att1, const float* __restrict__ hg, long long ldhg,
                                  const float* __restrict__ w_f, const float* __restrict__ b_f,
                                  const float* __restrict__ enc, const float* __restrict__ active,
                                  float* __restrict__ alpha_out, long long alpha_ld, OpOut awe, long long ld_awe,
                                  int P, int A, int E, int apply_gate, int enc_group) {
  cg::cluster_group cluster = cg::this_cluster();
  grid_dep_sync();
  extern __shared__ __align__(16) float att_sm[];
  float* s_att2 = att_sm;            // [A]
  float* s_wf = att_sm + A;          // [A]
  float* s_part = att_sm + 2 * A;    // [4][E / ATC] partial weighted sums of this CTA's channel slice
  __shared__ float s_e[ATT_MAX_P];
  const int r = static_cast<int>(cluster.block_rank());
  const int b = blockIdx.x / ATC;
  const int be = b / enc_group;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < A; i += 256) {
    s_att2[i] = hg[b * ldhg + i];
    s_wf[i] = __ldg(w_f + i);
  }
  __syncthreads();
  const float bf = b_f ? __ldg(b_f) : 0.f;
  float* s_e_remote = cluster.map_shared_rank(s_e, lane < ATC ? lane : 0);   // lane q addresses CTA q's copy
  for (int p = r + ATC * warp; p < P; p += ATC * 8) {
    const float4* a1 = reinterpret_cast<const float4*>(att1 + (static_cast<long long>(be) * P + p) * A);
    float acc = 0.f;
    for (int i = lane; i < A / 4; i += 32) {
      const float4 v = __ldg(a1 + i);
      const float4 h = *reinterpret_cast<const float4*>(s_att2 + i * 4);
      const float4 w = *reinterpret_cast<const float4*>(s_wf + i * 4);
      acc = fmaf(fmaxf(v.x + h.x, 0.f), w.x, acc);
      acc = fmaf(fmaxf(v.y + h.y, 0.f), w.y, acc);
      acc = fmaf(fmaxf(v.z + h.z, 0.f), w.z, acc);
      acc = fmaf(fmaxf(v.w + h.w, 0.f), w.w, acc);
    }
    acc = warp_sum(acc);
    if (lane < ATC) s_e_remote[p] = acc + bf;
  }
  cluster.sync();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int p = lane; p < P; p += 32) mx = fmaxf(mx, s_e[p]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int p = lane; p < P; p += 32) {
      const float ex = expf(s_e[p] - mx);
      s_e[p] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    const bool wr = (r == 0) && (alpha_out != nullptr) && (active == nullptr || active[b] != 0.f);
    for (int p = lane; p < P; p += 32) {
      const float a = s_e[p] * inv;
      s_e[p] = a;
      if (wr) alpha_out[b * alpha_ld + p] = a;
    }
  }
  __syncthreads();
  const int slice4 = E / 4 / ATC;                 // float4 lanes of this CTA's channel slice
  constexpr int PS = 4, GT = 256 / PS;
  const int q = threadIdx.x / GT, tq = threadIdx.x % GT;
  const int pp = (P + PS - 1) / PS;
  const int pa = q * pp, pb = min(P, pa + pp);
  for (int l4 = tq; l4 < slice4; l4 += GT) {
    const int e4 = r * slice4 + l4;
    const float4* src = reinterpret_cast<const float4*>(enc + static_cast<long long>(be) * P * E) + e4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = pa; p < pb; ++p) {
      const float4 v = __ldg(src + static_cast<long long>(p) * (E / 4));
      const float a = s_e[p];
      acc.x = fmaf(v.x, a, acc.x); acc.y = fmaf(v.y, a, acc.y);
      acc.z = fmaf(v.z, a, acc.z); acc.w = fmaf(v.w, a, acc.w);
    }
    *reinterpret_cast<float4*>(s_part + (static_cast<long long>(q) * slice4 + l4) * 4) = acc;
  }
  __syncthreads();
  if (q == 0) {
    for (int l4 = tq; l4 < slice4; l4 += GT) {
      const int e4 = r * slice4 + l4;
      float4 acc = *reinterpret_cast<const float4*>(s_part + static_cast<long long>(l4) * 4);
#pragma unroll
      for (int g = 1; g < PS; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(s_part + (static_cast<long long>(g) * slice4 + l4) * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      if (apply_gate) {
        const float4 gp = *reinterpret_cast<const float4*>(hg + b * ldhg + A + e4 * 4);
        acc.x *= 1.0f / (1.0f + expf(-gp.x));
        acc.y *= 1.0f / (1.0f + expf(-gp.y));
        acc.z *= 1.0f / (1.0f + expf(-gp.z));
        acc.w *= 1.0f / (1.0f + expf(-gp.w));
      }
      store_op4(awe, b * ld_awe + e4 * 4, acc);
    }
  }
}

int bahdanau_attention(const float* att1, const float* hg, long long ldhg, const float* w_f, const float* b_f,
                       const float* enc, const float* active, float* alpha_out, long long alpha_ld, void* awe_hi,
                       float* awe_lo, int awe_dtype, long long ld_awe, int bt, int P, int A, int E,
                       int apply_gate, int enc_group, cudaStream_t stream) {
  if (bt <= 0) return CCX_OK;
  if (P <= 0 || P > ATT_MAX_P || A % 4 != 0 || E % 4 != 0 || (ldhg % 4) != 0) return CCX_ERR_SHAPE;
  OpOut awe{awe_hi, awe_lo, awe_dtype};
  ProfScope prof(PROF_ATTENTION, stream, (double)bt * P * (A + E) * 4.0);
  static int use_cluster = -1;
  if (use_cluster < 0) use_cluster = getenv("CCX_ATT_CLUSTER") ? atoi(getenv("CCX_ATT_CLUSTER")) : 1;
  if (use_cluster && (E % (4 * ATC)) == 0 && static_cast<long long>(bt) * ATC <= 0x7fffffffLL) {
    const size_t csm = (2 * static_cast<size_t>(A) + static_cast<size_t>(E)) * sizeof(float);
    if (csm <= 48 * 1024) {
      return launch_pdl(bahdanau_attention_cluster_kernel, dim3(bt * ATC), dim3(256), csm, stream, att1, hg, ldhg,
                        w_f, b_f, enc, active, alpha_out, alpha_ld, awe, ld_awe, P, A, E, apply_gate,
                        enc_group > 0 ? enc_group : 1) == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
    }
  }
  const size_t smem = (2 * static_cast<size_t>(A) + ATT_PSPLIT * static_cast<size_t>(E)) * sizeof(float);
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) return CCX_ERR_SHAPE;
    static PerDevice<size_t> configured_dev;
    size_t& configured = configured_dev.ref();
    if (smem > configured) {
      if (cudaFuncSetAttribute(bahdanau_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem)) != cudaSuccess)
        return CCX_ERR_CUDA;
      configured = smem;
    }
  }
  bahdanau_attention_kernel<<<bt, ATT_THREADS, smem, stream>>>(att1, hg, ldhg, w_f, b_f, enc, active, alpha_out,
                                                              alpha_ld, awe, ld_awe, P, A, E, apply_gate,
                                                              enc_group > 0 ? enc_group : 1);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// LSTM cell point-wise part (gate order i, f, g, o as torch)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lstm_pointwise_kernel(const float* __restrict__ gates, long long ldg, const float* __restrict__ c_prev,
                      float* __restrict__ c_new, OpOut h_next, long long ld_hn, OpOut h_all, long long ld_ha,
                      const float* __restrict__ dropmask, long long ld_dm, float* __restrict__ h_plain,
                      long long ld_hp, int bt, int D) {
  grid_dep_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(bt) * D) return;
  const int b = static_cast<int>(i / D), j = static_cast<int>(i % D);
  const float* g = gates + b * ldg;
  const float ig = 1.0f / (1.0f + expf(-g[j]));
  const float fg = 1.0f / (1.0f + expf(-g[D + j]));
  const float gg = tanhf(g[2 * D + j]);
  const float og = 1.0f / (1.0f + expf(-g[3 * D + j]));
  const float c = fg * c_prev[static_cast<long long>(b) * D + j] + ig * gg;
  const float h = og * tanhf(c);
  c_new[static_cast<long long>(b) * D + j] = c;
  if (h_next.hi != nullptr) store_op(h_next, b * ld_hn + j, h);
  if (h_plain != nullptr) h_plain[b * ld_hp + j] = h;
  if (h_all.hi != nullptr) {
    const float m = dropmask ? dropmask[b * ld_dm + j] : 1.0f;
    store_op(h_all, b * ld_ha + j, h * m);
  }
}

int lstm_pointwise(const float* gates, long long ldg, const float* c_prev, float* c_new, void* hn_hi, float* hn_lo,
                   long long ld_hn, void* ha_hi, float* ha_lo, long long ld_ha, int op_dtype, const float* dropmask,
                   long long ld_dm, float* h_plain, long long ld_hp, int bt, int D, cudaStream_t stream) {
  if (bt <= 0) return CCX_OK;
  OpOut hn{hn_hi, hn_lo, op_dtype}, ha{ha_hi, ha_lo, op_dtype};
  const long long n = static_cast<long long>(bt) * D;
  ProfScope prof(PROF_LSTM, stream, (double)n * 36.0);
  return launch_pdl(lstm_pointwise_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, stream, gates,
                    ldg, c_prev, c_new, hn, ld_hn, ha, ld_ha, dropmask, ld_dm, h_plain, ld_hp, bt, D) == cudaSuccess
             ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// greedy step: argmax over V (first index on ties), sequences / finished / active / next-token bookkeeping
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
greedy_next_kernel(const float* __restrict__ preds, long long ld_preds, int V, int t, int T,
                   long long* __restrict__ sequences,  // [B, T]
                   float* __restrict__ active,         // [B] in/out (1 = still decoding)
                   long long* __restrict__ next_tok, long long ld_next,  // next_tok[b*ld_next] <- argmax (if active)
                   long long end_token) {
  const int b = blockIdx.x;
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  const float* row = preds + b * ld_preds;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += 256) {
    const float x = row[v];
    if (x > best || (x == best && v < bi)) { best = x; bi = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_v[warp] = best; s_i[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) { best = s_v[w]; bi = s_i[w]; }
    if (active[b] != 0.f) {
      sequences[static_cast<long long>(b) * T + t] = bi;
      if (next_tok != nullptr) next_tok[b * ld_next] = bi;
      if (bi == end_token) active[b] = 0.f;
    }
  }
}

int greedy_next(const float* preds, long long ld_preds, int B, int V, int t, int T, long long* sequences,
                float* active, long long* next_tok, long long ld_next, long long end_token, cudaStream_t stream) {
  if (B <= 0) return CCX_OK;
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)B * V * 4.0);
  greedy_next_kernel<<<B, 256, 0, stream>>>(preds, ld_preds, V, t, T, sequences, active, next_tok, ld_next,
                                            end_token);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// small-sequence multi-head attention: one CTA per (batch, head); K/V of the head staged in shared memory
// ---------------------------------------------------------------------------------------------
struct MhaArgs {
  const float* q; long long q_sb, q_st;  // element (b, i, h*hd + d) at q[b*q_sb + i*q_st + h*hd + d]
  const float* k; long long k_sb, k_st;
  const float* v; long long v_sb, v_st;
  OpOut ctx; long long c_sb, c_st;
  const unsigned char* key_pad;  // [B, Tk] 1 = masked, or nullptr
  const float* prob_mask;        // [B, H, Tq, Tk] dropout multiplier or nullptr
  float* probs_out;              // [B, H, Tq, Tk] softmax output (before dropout) or nullptr
  int B, H, Tq, Tk, hd;
  int kv_group;                  // k/v batch row = b / kv_group (beam search: beams share the image memory)
  int causal;                    // key j allowed iff j <= q_pos0 + i
  int q_pos0;
  float scale;
};

__global__ void __launch_bounds__(128)
mha_small_kernel(MhaArgs a) {
  extern __shared__ float mha_sm[];
  const int hd = a.hd, ldk = hd + 1;
  float* s_k = mha_sm;                    // [Tk][hd+1]
  float* s_v = s_k + a.Tk * ldk;          // [Tk][hd+1]
  float* s_q = s_v + a.Tk * ldk;          // [rows of this CTA][hd]   (pre-scaled)
  const int rows_per_cta = (a.Tq + gridDim.y - 1) / gridDim.y;
  float* s_p = s_q + rows_per_cta * hd;   // [4 warps][Tk]
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const long long bk = b / a.kv_group;
  const int i_begin = blockIdx.y * rows_per_cta;
  const int i_end = min(a.Tq, i_begin + rows_per_cta);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    // stage K, V and this CTA's Q rows with 16-byte loads, several in flight per thread: everything the row loop
    // touches afterwards is in shared memory (the loop used to pay a global round trip per query row)
    const int hd4 = hd >> 2;
#pragma unroll 4
    for (int idx = threadIdx.x; idx < a.Tk * hd4; idx += 128) {
      const int j = idx / hd4, d = (idx - j * hd4) * 4;
      const float4 kk = __ldg(reinterpret_cast<const float4*>(a.k + bk * a.k_sb + j * a.k_st + h * hd + d));
      const float4 vv = __ldg(reinterpret_cast<const float4*>(a.v + bk * a.v_sb + j * a.v_st + h * hd + d));
      float* dk = s_k + j * ldk + d;
      float* dv = s_v + j * ldk + d;
      dk[0] = kk.x; dk[1] = kk.y; dk[2] = kk.z; dk[3] = kk.w;
      dv[0] = vv.x; dv[1] = vv.y; dv[2] = vv.z; dv[3] = vv.w;
    }
#pragma unroll 4
    for (int idx = threadIdx.x; idx < (i_end - i_begin) * hd4; idx += 128) {
      const int r = idx / hd4, d = (idx - r * hd4) * 4;
      const float4 qq = __ldg(reinterpret_cast<const float4*>(a.q + b * a.q_sb + (i_begin + r) * a.q_st + h * hd + d));
      float* dq = s_q + r * hd + d;
      dq[0] = qq.x * a.scale; dq[1] = qq.y * a.scale; dq[2] = qq.z * a.scale; dq[3] = qq.w * a.scale;
    }
  }
  __syncthreads();
  float* p = s_p + warp * a.Tk;
  for (int i = i_begin + warp; i < i_end; i += 4) {
    const float* q = s_q + (i - i_begin) * hd;
    const long long prow = ((static_cast<long long>(b) * a.H + h) * a.Tq + i) * a.Tk;
    float mx = -INFINITY;
    for (int j = lane; j < a.Tk; j += 32) {
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < hd; ++d) s = fmaf(q[d], s_k[j * ldk + d], s);
      const bool masked = (a.causal && j > a.q_pos0 + i) || (a.key_pad && a.key_pad[b * a.Tk + j]);
      s = masked ? -INFINITY : s;
      p[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < a.Tk; j += 32) {
      const float e = (p[j] == -INFINITY) ? 0.f : expf(p[j] - mx);
      p[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < a.Tk; j += 32) {
      float pr = p[j] * inv;
      if (a.probs_out) a.probs_out[prow + j] = pr;
      if (a.prob_mask) pr *= __ldg(a.prob_mask + prow + j);
      p[j] = pr;
    }
    __syncwarp();
    for (int d = lane; d < hd; d += 32) {
      float acc = 0.f;
#pragma unroll 4
      for (int j = 0; j < a.Tk; ++j) acc = fmaf(p[j], s_v[j * ldk + d], acc);
      store_op(a.ctx, b * a.c_sb + i * a.c_st + h * hd + d, acc);
    }
    __syncwarp();
  }
}

int mha_small(const float* q, long long q_sb, long long q_st, const float* k, long long k_sb, long long k_st,
              const float* v, long long v_sb, long long v_st, void* ctx_hi, float* ctx_lo, int ctx_dtype,
              long long c_sb, long long c_st, const unsigned char* key_pad, const float* prob_mask,
              float* probs_out, int B, int H, int Tq, int Tk, int hd, int causal, int q_pos0, float scale,
              int kv_group, cudaStream_t stream) {
  if (B <= 0 || Tq <= 0) return CCX_OK;
  // bf16 compute mode, short sequences: the tensor-core kernel (mha_tc.cu)
  static const bool no_tc = getenv("CCX_MHA_TC") != nullptr && atoi(getenv("CCX_MHA_TC")) == 0;
  if (!no_tc && ctx_dtype == CCX_BF16 && ctx_lo == nullptr && H > 0 && mha_tc_eligible(Tq, Tk, hd))
    return mha_tc_fwd(q, q_sb, q_st, k, k_sb, k_st, v, v_sb, v_st, ctx_hi, c_sb, c_st, key_pad, prob_mask, probs_out, B,
                      H, Tq, Tk, causal, q_pos0, scale, kv_group, stream);
  if (Tk <= 0 || hd <= 0 || H <= 0 || (hd & 3) || (k_sb & 3) || (k_st & 3) || (v_sb & 3) || (v_st & 3) ||
      (reinterpret_cast<uintptr_t>(k) & 15) || (reinterpret_cast<uintptr_t>(v) & 15) || (q_sb & 3) || (q_st & 3) ||
      (reinterpret_cast<uintptr_t>(q) & 15))
    return CCX_ERR_SHAPE;
  int ysplit = 1;
  while (ysplit < 4 && static_cast<long long>(B) * H * ysplit < 2 * 148 && Tq / (ysplit * 2) >= 8) ysplit *= 2;
  const int rows_per_cta = (Tq + ysplit - 1) / ysplit;
  const size_t smem = (static_cast<size_t>(2) * Tk * (hd + 1) + static_cast<size_t>(rows_per_cta) * hd + 4 * Tk) *
                      sizeof(float);
  if (smem > 200 * 1024) return CCX_ERR_SHAPE;
  static PerDevice<size_t> configured_dev;
    size_t& configured = configured_dev.ref();
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(mha_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess)
      return CCX_ERR_CUDA;
    configured = 200 * 1024;
  }
  MhaArgs a;
  a.q = q; a.q_sb = q_sb; a.q_st = q_st;
  a.k = k; a.k_sb = k_sb; a.k_st = k_st;
  a.v = v; a.v_sb = v_sb; a.v_st = v_st;
  a.ctx = OpOut{ctx_hi, ctx_lo, ctx_dtype}; a.c_sb = c_sb; a.c_st = c_st;
  a.key_pad = key_pad; a.prob_mask = prob_mask; a.probs_out = probs_out;
  a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.hd = hd;
  a.causal = causal; a.q_pos0 = q_pos0; a.scale = scale;
  a.kv_group = kv_group > 0 ? kv_group : 1;
  ProfScope prof(PROF_ATTENTION, stream, (double)B * H * (Tq + 2.0 * Tk) * hd * 4.0);
  mha_small_kernel<<<dim3(B * H, ysplit), 128, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// single-query attention for KV-cache decoding: one WARP per (row, head), no shared-memory staging.
//   scores: lanes own keys (float4 loads of the 64-wide head slice), softmax by shuffles,
//   context: lanes own output dims, keys streamed (coalesced 256 B per key).
// kv_rows (optional) [rows, ld_map]: physical cache row holding position j of logical row r — beam search
// re-orders beams by re-writing this small map instead of copying the caches (caption.py:138-145 semantics).
// ---------------------------------------------------------------------------------------------
struct MhaDecArgs {
  const float* q; long long q_sb;            // (r, h*hd + d) at q[r*q_sb + h*hd + d]
  const float* k; long long k_sb, k_st;      // (row, j, h*hd + d) at k[row*k_sb + j*k_st + h*hd + d]
  const float* v; long long v_sb, v_st;
  OpOut ctx; long long c_sb;
  const int* kv_rows; long long ld_map;
  int rows, H, Tk, hd, kv_group;
  float scale;
};

__global__ void __launch_bounds__(256)
mha_decode_kernel(MhaDecArgs a) {
  const int warp_global = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp_global >= a.rows * a.H) return;
  const int r = warp_global / a.H, h = warp_global % a.H;
  const int hd = a.hd;          // multiple of 4, <= 128
  const float* qp = a.q + r * a.q_sb + h * hd;
  // scores for keys lane, lane+32, ... (Tk <= 256 -> at most 8 per lane)
  float sc[8];
  int prow[8];
  float mx = -INFINITY;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int j = lane + n * 32;
    sc[n] = -INFINITY;
    prow[n] = 0;
    if (j < a.Tk) {
      const long long row = a.kv_rows ? a.kv_rows[r * a.ld_map + j] : (r / a.kv_group);
      prow[n] = static_cast<int>(row);
      const float4* kp = reinterpret_cast<const float4*>(a.k + row * a.k_sb + j * a.k_st + h * hd);
      float s = 0.f;
#pragma unroll 8
      for (int d4 = 0; d4 < hd / 4; ++d4) {
        const float4 kk = __ldg(kp + d4);
        const float4 qq = __ldg(reinterpret_cast<const float4*>(qp) + d4);
        s = fmaf(qq.x, kk.x, s); s = fmaf(qq.y, kk.y, s); s = fmaf(qq.z, kk.z, s); s = fmaf(qq.w, kk.w, s);
      }
      sc[n] = s * a.scale;
      mx = fmaxf(mx, sc[n]);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int j = lane + n * 32;
    sc[n] = (j < a.Tk) ? expf(sc[n] - mx) : 0.f;
    sum += sc[n];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  // context: lane owns dims d = lane, lane+32, ... ; probabilities / rows broadcast by shuffle.  Keys are handled
  // in batches of 8 with all V loads issued before the FMAs (the loop is otherwise one exposed global round trip
  // per key).
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int nd = (hd + 31) / 32;   // dims per lane (<= 4)
  for (int n = 0; n < 8; ++n) {
    if (n * 32 >= a.Tk) break;
    const int cnt = min(32, a.Tk - n * 32);
    for (int j0 = 0; j0 < cnt; j0 += 8) {
      float pw[8];
      float vv[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int jj = min(j0 + u, cnt - 1);
        const float p = __shfl_sync(0xffffffffu, sc[n], jj);
        const long long row = __shfl_sync(0xffffffffu, prow[n], jj);
        pw[u] = (j0 + u < cnt) ? p * inv : 0.f;
        const float* vp = a.v + row * a.v_sb + static_cast<long long>(n * 32 + jj) * a.v_st + h * hd;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int d = lane + i * 32;
          vv[u][i] = (i < nd && d < hd) ? __ldg(vp + d) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(pw[u], vv[u][i], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + i * 32;
    if (d < hd) store_op(a.ctx, r * a.c_sb + h * hd + d, acc[i]);
  }
}

int mha_decode(const float* q, long long q_sb, const float* k, long long k_sb, long long k_st, const float* v,
               long long v_sb, long long v_st, void* ctx_hi, float* ctx_lo, int ctx_dtype, long long c_sb,
               const int* kv_rows, long long ld_map, int rows, int H, int Tk, int hd, int kv_group, float scale,
               cudaStream_t stream) {
  if (rows <= 0) return CCX_OK;
  if (Tk <= 0 || Tk > 256 || hd <= 0 || hd > 128 || (hd & 3) || H <= 0 || ((q_sb | k_sb | k_st | v_sb | v_st) & 3) ||
      (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15))
    return CCX_ERR_SHAPE;
  MhaDecArgs a;
  a.q = q; a.q_sb = q_sb; a.k = k; a.k_sb = k_sb; a.k_st = k_st; a.v = v; a.v_sb = v_sb; a.v_st = v_st;
  a.ctx = OpOut{ctx_hi, ctx_lo, ctx_dtype}; a.c_sb = c_sb;
  a.kv_rows = kv_rows; a.ld_map = ld_map;
  a.rows = rows; a.H = H; a.Tk = Tk; a.hd = hd; a.kv_group = kv_group > 0 ? kv_group : 1; a.scale = scale;
  const long long warps = static_cast<long long>(rows) * H;
  ProfScope prof(PROF_ATTENTION, stream, (double)rows * H * (2.0 * Tk + 2.0) * hd * 4.0);
  mha_decode_kernel<<<static_cast<unsigned>((warps + 7) / 8), 256, 0, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// attention maps for visualisation (models/transformerDecoderAttVis.py:163-165,223-226): mean over heads (and,
// by accumulation over calls, over layers) of the cross-attention probabilities.
//   alphas[b, t, j] (+)= scale * row_active[b] * sum_h probs[b, h, t, j] (* mask[b, h, t, j])
// ---------------------------------------------------------------------------------------------
__global__ void attn_head_mean_kernel(const float* __restrict__ probs, long long p_sb, long long p_sh, long long p_st,
                                      const float* __restrict__ mask, const float* __restrict__ row_active,
                                      float* __restrict__ alphas, long long a_sb, long long a_st, int B, int H, int Tq,
                                      int Tk, float scale, int accumulate) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * Tq * Tk) return;
  const int j = static_cast<int>(idx % Tk), t = static_cast<int>((idx / Tk) % Tq), b = static_cast<int>(idx / (static_cast<long long>(Tk) * Tq));
  float s = 0.f;
  for (int h = 0; h < H; ++h) {
    const long long off = b * p_sb + h * p_sh + t * p_st + j;
    float v = __ldg(probs + off);
    if (mask != nullptr) v *= __ldg(mask + off);
    s += v;
  }
  s *= scale;
  if (row_active != nullptr) s *= __ldg(row_active + b);
  const long long o = b * a_sb + t * a_st + j;
  alphas[o] = accumulate ? alphas[o] + s : s;
}

int attn_head_mean(const float* probs, long long p_sb, long long p_sh, long long p_st, const float* mask,
                   const float* row_active, float* alphas, long long a_sb, long long a_st, int B, int H, int Tq, int Tk,
                   float scale, int accumulate, cudaStream_t stream) {
  if (B <= 0 || Tq <= 0) return CCX_OK;
  if (H <= 0 || Tk <= 0 || probs == nullptr || alphas == nullptr) return CCX_ERR_SHAPE;
  const long long n = static_cast<long long>(B) * Tq * Tk;
  attn_head_mean_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      probs, p_sb, p_sh, p_st, mask, row_active, alphas, a_sb, a_st, B, H, Tq, Tk, scale, accumulate);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

}  // namespace ccx
