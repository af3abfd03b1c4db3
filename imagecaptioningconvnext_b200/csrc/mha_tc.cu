// mha_tc.cu — small-sequence multi-head attention on the tensor cores (bf16 compute mode, T <= 64, head dim 64).
//
// Replaces, for the Transformer decoder's teacher-forced pass and its backward, the SIMT kernels mha_small_kernel /
// mha_bwd_kernel (decoder_kernels.cu / train_kernels.cu), which spend 48 / 84 us per launch (12 launches each per
// train step) on 0.7 MFLOP per (batch, head): nn.MultiheadAttention inside nn.TransformerDecoderLayer
// (models/transformerDecoder.py:102-106 through torch/nn/modules/transformer.py).
//
// One CTA per (batch, head), 4 warps, every warp owns 16 rows.  Q / K / V (/ dO) are converted to bf16 into padded
// shared-memory tiles (row pitch 144 B: ldmatrix reads are conflict free); all contractions are mma.sync m16n8k16 with
// fp32 accumulators:
//   forward : S = (scale Q) K^T -> mask -> softmax in registers -> probs_out (fp32, before dropout) -> P*dropout as the
//             A fragments of O = P V (the accumulator layout of two n-tiles IS an A fragment)
//   backward: dP = dO V^T; dS = P (dP*mask - rowsum) and Pd = P*mask to shared memory (bf16);
//             dQ = scale dS K, dK = scale dS^T Q, dV = Pd^T dO  (transposed operands through ldmatrix.trans)
#include "ccx_common.cuh"
#include "ccx_ops.h"
#include "ccx_prof.h"

namespace ccx {
namespace {

constexpr int TC_T = 64;          // rows / keys per tile (padded)
constexpr int TC_HD = 64;
constexpr int TC_LD = 72;         // bf16 elements per shared-memory row (144 bytes)

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragment (16 x 16) of a row-major [m][k] tile at (m0, k0)
__device__ __forceinline__ void frag_a(uint32_t (&a)[4], const __nv_bfloat16* t, int m0, int k0, int lane) {
  ldsm_x4(a, smem_u32(t + (m0 + (lane & 15)) * TC_LD + k0 + (lane >> 4) * 8));
}
// A fragment of the TRANSPOSE of a row-major [k][m] tile: rows m0.., contraction k0..
__device__ __forceinline__ void frag_a_t(uint32_t (&a)[4], const __nv_bfloat16* t, int m0, int k0, int lane) {
  ldsm_x4_t(a, smem_u32(t + (k0 + (lane & 7) + (lane >> 4) * 8) * TC_LD + m0 + ((lane >> 3) & 1) * 8));
}
// B fragments of TWO n-tiles (n0 .. n0+15) x k16 from a row-major [n][k] tile: (b[0], b[1]) and (b[2], b[3])
__device__ __forceinline__ void frag_b_nk(uint32_t (&b)[4], const __nv_bfloat16* t, int n0, int k0, int lane) {
  ldsm_x4(b, smem_u32(t + (n0 + (lane & 7) + (lane >> 4) * 8) * TC_LD + k0 + ((lane >> 3) & 1) * 8));
}
// ... from a row-major [k][n] tile
__device__ __forceinline__ void frag_b_kn(uint32_t (&b)[4], const __nv_bfloat16* t, int n0, int k0, int lane) {
  ldsm_x4_t(b, smem_u32(t + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * TC_LD + n0 + (lane >> 4) * 8));
}

// fp32 [rows, 64] slice (row stride st) -> bf16 tile, rows >= n zeroed, optional scale
__device__ __forceinline__ void stage_tile(__nv_bfloat16* t, const float* src, long long st, int n, float scale, int tid) {
  for (int idx = tid; idx < TC_T * (TC_HD / 4); idx += 128) {
    const int r = idx >> 4, d = (idx & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n) v = __ldg(reinterpret_cast<const float4*>(src + r * st + d));
    uint2 pk;
    pk.x = pack_bf16x2(v.x * scale, v.y * scale);
    pk.y = pack_bf16x2(v.z * scale, v.w * scale);
    *reinterpret_cast<uint2*>(t + r * TC_LD + d) = pk;
  }
}

// acc[8][4] (16 x 64) += A(16 rows m0.., 64-wide contraction) . B
template <bool A_T, bool B_KN>
__device__ __forceinline__ void gemm_16x64x64(float (&acc)[8][4], const __nv_bfloat16* ta, int m0, const __nv_bfloat16* tb,
                                              int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    if (A_T) frag_a_t(a, ta, m0, 16 * kk, lane);
    else frag_a(a, ta, m0, 16 * kk, lane);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      if (B_KN) frag_b_kn(b, tb, 16 * np, 16 * kk, lane);
      else frag_b_nk(b, tb, 16 * np, 16 * kk, lane);
      mma_bf16(acc[2 * np], a, b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

struct TcFwdArgs {
  const float* q; long long q_sb, q_st;
  const float* k; long long k_sb, k_st;
  const float* v; long long v_sb, v_st;
  __nv_bfloat16* ctx; long long c_sb, c_st;
  const unsigned char* key_pad;
  const float* prob_mask;
  float* probs_out;
  int H, Tq, Tk, kv_group, causal, q_pos0;
  float scale;
};

__global__ void __launch_bounds__(128)
mha_tc_fwd_kernel(TcFwdArgs a) {
  __shared__ __align__(16) __nv_bfloat16 sQ[TC_T * TC_LD], sK[TC_T * TC_LD], sV[TC_T * TC_LD];
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const long long bk = b / a.kv_group;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_tile(sQ, a.q + b * a.q_sb + h * TC_HD, a.q_st, a.Tq, a.scale, tid);
  stage_tile(sK, a.k + bk * a.k_sb + h * TC_HD, a.k_st, a.Tk, 1.0f, tid);
  stage_tile(sV, a.v + bk * a.v_sb + h * TC_HD, a.v_st, a.Tk, 1.0f, tid);
  __syncthreads();
  const int m0 = 16 * warp;
  if (m0 >= a.Tq) return;                                     // warp-uniform; no barrier follows
  float s[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
  gemm_16x64x64<false, false>(s, sQ, m0, sK, lane);           // S = (scale Q) K^T
  const int g = lane >> 2, t = lane & 3;
  const long long pbase = (static_cast<long long>(b) * a.H + h) * a.Tq * a.Tk;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {                            // the two rows this thread holds: g and g + 8
    const int row = m0 + g + 8 * hf;
    float mx = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = 8 * nt + 2 * t + c;
        const bool masked = col >= a.Tk || (a.causal && col > a.q_pos0 + row) ||
                            (a.key_pad != nullptr && col < a.Tk && a.key_pad[b * a.Tk + col]);
        float& x = s[nt][2 * hf + c];
        x = masked ? -INFINITY : x;
        mx = fmaxf(mx, x);
      }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float& x = s[nt][2 * hf + c];
        x = (x == -INFINITY) ? 0.f : __expf(x - mx);
        sum += x;
      }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    const bool row_ok = row < a.Tq;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = 8 * nt + 2 * t + c;
        float pr = s[nt][2 * hf + c] * inv;
        if (row_ok && col < a.Tk) {
          if (a.probs_out != nullptr) a.probs_out[pbase + static_cast<long long>(row) * a.Tk + col] = pr;
          if (a.prob_mask != nullptr) pr *= __ldg(a.prob_mask + pbase + static_cast<long long>(row) * a.Tk + col);
        } else {
          pr = 0.f;
        }
        s[nt][2 * hf + c] = pr;
      }
  }
  // O = P V : the accumulators of n-tiles (2kk, 2kk+1) are the A fragment of contraction step kk
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bb[4];
      frag_b_kn(bb, sV, 16 * np, 16 * kk, lane);
      mma_bf16(o[2 * np], pa, bb[0], bb[1]);
      mma_bf16(o[2 * np + 1], pa, bb[2], bb[3]);
    }
  }
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int row = m0 + g + 8 * hf;
    if (row >= a.Tq) continue;
    __nv_bfloat16* dst = a.ctx + b * a.c_sb + row * a.c_st + h * TC_HD + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<uint32_t*>(dst + 8 * nt) = pack_bf16x2(o[nt][2 * hf], o[nt][2 * hf + 1]);
  }
}

struct TcBwdArgs {
  const float* q; long long q_sb, q_st;
  const float* k; long long k_sb, k_st;
  const float* v; long long v_sb, v_st;
  const float* dctx; long long d_sb, d_st;
  const float* probs;
  const float* prob_mask;
  float* dq; long long dq_sb, dq_st;
  float* dk; long long dk_sb, dk_st;
  float* dv; long long dv_sb, dv_st;
  int H, Tq, Tk;
  float scale;
};

constexpr int TC_BWD_SMEM = 6 * TC_T * TC_LD * 2;

__device__ __forceinline__ void store_16x64(float* dst, long long st, const float (&acc)[8][4], int m0, int rows, float scale,
                                            int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int row = m0 + g + 8 * hf;
    if (row >= rows) continue;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<float2*>(dst + row * st + 8 * nt + 2 * t) =
          make_float2(acc[nt][2 * hf] * scale, acc[nt][2 * hf + 1] * scale);
  }
}

__global__ void __launch_bounds__(128)
mha_tc_bwd_kernel(TcBwdArgs a) {
  extern __shared__ __align__(16) uint8_t tc_bsm[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(tc_bsm);
  __nv_bfloat16* sK = sQ + TC_T * TC_LD;
  __nv_bfloat16* sV = sK + TC_T * TC_LD;
  __nv_bfloat16* sDO = sV + TC_T * TC_LD;
  __nv_bfloat16* sP = sDO + TC_T * TC_LD;       // P * dropout mask  [i][j]
  __nv_bfloat16* sDS = sP + TC_T * TC_LD;       // dS                [i][j]
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_tile(sQ, a.q + b * a.q_sb + h * TC_HD, a.q_st, a.Tq, 1.0f, tid);
  stage_tile(sK, a.k + b * a.k_sb + h * TC_HD, a.k_st, a.Tk, 1.0f, tid);
  stage_tile(sV, a.v + b * a.v_sb + h * TC_HD, a.v_st, a.Tk, 1.0f, tid);
  stage_tile(sDO, a.dctx + b * a.d_sb + h * TC_HD, a.d_st, a.Tq, 1.0f, tid);
  __syncthreads();
  const int m0 = 16 * warp;
  const int g = lane >> 2, t = lane & 3;
  const long long pbase = (static_cast<long long>(b) * a.H + h) * a.Tq * a.Tk;
  {
    // dP = dO V^T, then dS = P (dP*mask - sum_j dP*mask*P) and Pd = P*mask, both to shared memory as bf16
    float dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    gemm_16x64x64<false, false>(dp, sDO, m0, sV, lane);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int row = m0 + g + 8 * hf;
      const bool row_ok = row < a.Tq;
      float pr[16], pm[16];
      float rowdot = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = 8 * nt + 2 * t + c;
          float p = 0.f, m = 0.f;
          if (row_ok && col < a.Tk) {
            p = __ldg(a.probs + pbase + static_cast<long long>(row) * a.Tk + col);
            m = a.prob_mask != nullptr ? __ldg(a.prob_mask + pbase + static_cast<long long>(row) * a.Tk + col) : 1.f;
          }
          pr[2 * nt + c] = p;
          pm[2 * nt + c] = m;
          const float d = dp[nt][2 * hf + c] * m;
          dp[nt][2 * hf + c] = d;
          rowdot = fmaf(d, p, rowdot);
        }
      rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 1);
      rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 2);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = pr[2 * nt], p1 = pr[2 * nt + 1];
        const int off = row * TC_LD + 8 * nt + 2 * t;
        *reinterpret_cast<uint32_t*>(sDS + off) =
            pack_bf16x2(p0 * (dp[nt][2 * hf] - rowdot), p1 * (dp[nt][2 * hf + 1] - rowdot));
        *reinterpret_cast<uint32_t*>(sP + off) = pack_bf16x2(p0 * pm[2 * nt], p1 * pm[2 * nt + 1]);
      }
    }
  }
  __syncthreads();
  float acc[8][4];
  // dQ = scale dS K
  if (m0 < a.Tq) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    gemm_16x64x64<false, true>(acc, sDS, m0, sK, lane);
    store_16x64(a.dq + b * a.dq_sb + h * TC_HD, a.dq_st, acc, m0, a.Tq, a.scale, lane);
  }
  if (m0 < a.Tk) {
    // dK = scale dS^T Q
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    gemm_16x64x64<true, true>(acc, sDS, m0, sQ, lane);
    store_16x64(a.dk + b * a.dk_sb + h * TC_HD, a.dk_st, acc, m0, a.Tk, a.scale, lane);
    // dV = Pd^T dO
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    gemm_16x64x64<true, true>(acc, sP, m0, sDO, lane);
    store_16x64(a.dv + b * a.dv_sb + h * TC_HD, a.dv_st, acc, m0, a.Tk, 1.0f, lane);
  }
}

bool aligned4(long long x) { return (x & 3) == 0; }

}  // namespace

bool mha_tc_eligible(int Tq, int Tk, int hd) { return Tq >= 1 && Tq <= TC_T && Tk >= 1 && Tk <= TC_T && hd == TC_HD; }

int mha_tc_fwd(const float* q, long long q_sb, long long q_st, const float* k, long long k_sb, long long k_st,
               const float* v, long long v_sb, long long v_st, void* ctx_bf16, long long c_sb, long long c_st,
               const unsigned char* key_pad, const float* prob_mask, float* probs_out, int B, int H, int Tq, int Tk,
               int causal, int q_pos0, float scale, int kv_group, cudaStream_t stream) {
  if (B <= 0) return CCX_OK;
  if (!aligned4(q_sb) || !aligned4(q_st) || !aligned4(k_sb) || !aligned4(k_st) || !aligned4(v_sb) || !aligned4(v_st) ||
      ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) ||
      (reinterpret_cast<uintptr_t>(ctx_bf16) & 3) || (c_sb & 1) || (c_st & 1))
    return CCX_ERR_SHAPE;
  TcFwdArgs a;
  a.q = q; a.q_sb = q_sb; a.q_st = q_st; a.k = k; a.k_sb = k_sb; a.k_st = k_st; a.v = v; a.v_sb = v_sb; a.v_st = v_st;
  a.ctx = static_cast<__nv_bfloat16*>(ctx_bf16); a.c_sb = c_sb; a.c_st = c_st;
  a.key_pad = key_pad; a.prob_mask = prob_mask; a.probs_out = probs_out;
  a.H = H; a.Tq = Tq; a.Tk = Tk; a.kv_group = kv_group > 0 ? kv_group : 1; a.causal = causal; a.q_pos0 = q_pos0;
  a.scale = scale;
  ProfScope prof(PROF_ATTENTION, stream, (double)B * H * (Tq + 2.0 * Tk) * TC_HD * 4.0);
  mha_tc_fwd_kernel<<<B * H, 128, 0, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

int mha_tc_bwd(const float* q, long long q_sb, long long q_st, const float* k, long long k_sb, long long k_st,
               const float* v, long long v_sb, long long v_st, const float* dctx, long long d_sb, long long d_st,
               const float* probs, const float* prob_mask, float* dq, long long dq_sb, long long dq_st, float* dk,
               long long dk_sb, long long dk_st, float* dv, long long dv_sb, long long dv_st, int B, int H, int Tq,
               int Tk, float scale, cudaStream_t stream) {
  if (B <= 0) return CCX_OK;
  if (!aligned4(q_sb) || !aligned4(q_st) || !aligned4(k_sb) || !aligned4(k_st) || !aligned4(v_sb) || !aligned4(v_st) ||
      !aligned4(d_sb) || !aligned4(d_st) || (dq_sb & 1) || (dq_st & 1) || (dk_sb & 1) || (dk_st & 1) || (dv_sb & 1) ||
      (dv_st & 1) ||
      ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(dctx)) & 15) ||
      ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 7))
    return CCX_ERR_SHAPE;
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.ref();
  if (!configured) {
    if (cudaFuncSetAttribute(mha_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_BWD_SMEM) != cudaSuccess)
      return CCX_ERR_CUDA;
    configured = true;
  }
  TcBwdArgs a;
  a.q = q; a.q_sb = q_sb; a.q_st = q_st; a.k = k; a.k_sb = k_sb; a.k_st = k_st; a.v = v; a.v_sb = v_sb; a.v_st = v_st;
  a.dctx = dctx; a.d_sb = d_sb; a.d_st = d_st; a.probs = probs; a.prob_mask = prob_mask;
  a.dq = dq; a.dq_sb = dq_sb; a.dq_st = dq_st; a.dk = dk; a.dk_sb = dk_sb; a.dk_st = dk_st;
  a.dv = dv; a.dv_sb = dv_sb; a.dv_st = dv_st;
  a.H = H; a.Tq = Tq; a.Tk = Tk; a.scale = scale;
  ProfScope prof(PROF_ATTENTION, stream, (double)B * H * (Tq + Tk) * TC_HD * 16.0);
  mha_tc_bwd_kernel<<<B * H, 128, TC_BWD_SMEM, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

}  // namespace ccx
