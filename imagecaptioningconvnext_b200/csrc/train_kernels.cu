// train_kernels.cu — backward / optimizer kernels of the teacher-forced train step
// (reference: the autograd graph behind trainMultiGPU.py:357-394 and utils/utils.py:183-192).
//
//   convert_operand   fp32 [R,C] (x optional element-wise multiplier / ReLU mask) -> GEMM operand, optionally
//                     transposed ([C, Rpad], zero padded) — feeds dgrad (dY . W) and wgrad (dY^T . X) GEMMs
//   colsum_acc        bias gradients: out[c] += sum_r x[r,c]
//   ln_bwd            LayerNorm backward (recomputes mean / rstd from the saved input)
//   mha_bwd           small-sequence attention backward, one CTA per (batch, head)
//   softmax_ce        fused CrossEntropyLoss forward + backward over the rows selected by pack_padded_sequence
//   embedding_bwd     dense nn.Embedding gradient (atomic scatter-add)
//   adam_clamp        clip_gradient(+-c) fused with torch.optim.Adam's update, multi-tensor
#include "ccx_common.cuh"
#include "ccx_ops.h"
#include "ccx_prof.h"

namespace ccx {

static constexpr int ATT_MAX_P_BWD = 256;

// ---------------------------------------------------------------------------------------------
// convert / transpose to GEMM operand
// ---------------------------------------------------------------------------------------------
struct OpDst {
  void* hi;
  float* lo;
  int dtype;
};
__device__ __forceinline__ void put(const OpDst& o, long long idx, float v) {
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

// mul_mode: 0 none, 1 multiply by mul[r,c], 2 multiply by (mul[r,c] > 0) (ReLU mask from the saved output)
// (mode 2 also scales by mul_scale: a ReLU output that went through dropout is > 0 exactly where both the ReLU
// and the keep-mask let it through, and the keep multiplier 1/(1-p) is a constant)
__device__ __forceinline__ float apply_mul(float v, const float* mul, long long ldm, int r, int c, int mode,
                                           float mul_scale) {
  if (mode == 0) return v;
  const float m = mul[r * ldm + c];
  return mode == 1 ? v * m : (m > 0.f ? v * mul_scale : 0.f);
}

// source element: fp32 plain (lo == nullptr), tf32 (hi, lo) pair, or bf16
struct OpSrc {
  const void* hi;
  const float* lo;
  int dtype;
};
__device__ __forceinline__ float get(const OpSrc& s, long long idx) {
  if (s.dtype == CCX_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(s.hi)[idx]);
  const float h = reinterpret_cast<const float*>(s.hi)[idx];
  return s.lo ? h + s.lo[idx] : h;
}

__global__ void __launch_bounds__(256)
convert_rows_kernel(OpSrc x, long long ldx, const float* __restrict__ mul, long long ldm,
                    int mul_mode, float mul_scale, OpDst o, long long ldo, int R, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(R) * C) return;
  const int r = static_cast<int>(i / C), c = static_cast<int>(i % C);
  put(o, r * ldo + c, apply_mul(get(x, r * ldx + c), mul, ldm, r, c, mul_mode, mul_scale));
}

// out[c, r] = x[r, c] for r < R, 0 for R <= r < Rpad
__global__ void __launch_bounds__(256)
convert_transpose_kernel(OpSrc x, long long ldx, const float* __restrict__ mul, long long ldm,
                         int mul_mode, float mul_scale, OpDst o, long long ldo, int R, int C, int Rpad) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < R && c < C) ? apply_mul(get(x, r * ldx + c), mul, ldm, r, c, mul_mode, mul_scale) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < C && r < Rpad) put(o, c * ldo + r, tile[tx][j]);
  }
}

// ---- bf16-destination fast paths (the train step's operand conversions: ~45 launches per step) ------------------
// The generic kernels above move one element per thread (4-byte loads, 2-byte stores, an integer division per element)
// and measured 0.9-1.0 TB/s (profiles/r02_lstm_step_metrics.txt).  Here every thread moves 8 elements of a row with
// 16-byte accesses; the transpose goes through a 64x64 tile and writes 128 bytes per warp instruction.
template <bool SRC_BF16>
__device__ __forceinline__ void load8(const void* base, long long idx, float (&v)[8]) {
  if (SRC_BF16) {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  } else {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
__device__ __forceinline__ void apply_mul8(float (&v)[8], const float* mul, long long idx, int mode, float mul_scale) {
  if (mode == 0) return;
  const float4 a = *reinterpret_cast<const float4*>(mul + idx), b = *reinterpret_cast<const float4*>(mul + idx + 4);
  const float m[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = mode == 1 ? v[i] * m[i] : (m[i] > 0.f ? v[i] * mul_scale : 0.f);
}

template <bool SRC_BF16>
__global__ void __launch_bounds__(256)
convert_rows_bf16_kernel(const void* __restrict__ x, long long ldx, const float* __restrict__ mul, long long ldm,
                         int mul_mode, float mul_scale, __nv_bfloat16* __restrict__ o, long long ldo, int R, int C8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(R) * C8) return;
  const int r = static_cast<int>(i / C8), c = static_cast<int>(i % C8) * 8;
  float v[8];
  load8<SRC_BF16>(x, r * ldx + c, v);
  apply_mul8(v, mul, r * ldm + c, mul_mode, mul_scale);
  uint4 pk;
  pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
  pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(o + r * ldo + c) = pk;
}

// out[c, r] = x[r, c] (bf16), zero for R <= r < Rpad; 64 x 64 tile per CTA
template <bool SRC_BF16>
__global__ void __launch_bounds__(256)
convert_transpose_bf16_kernel(const void* __restrict__ x, long long ldx, const float* __restrict__ mul, long long ldm,
                              int mul_mode, float mul_scale, __nv_bfloat16* __restrict__ o, long long ldo, int R, int C,
                              int Rpad) {
  __shared__ float tile[64][65];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  {
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;     // 8 column groups of 8 x 32 rows
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int r = r0 + ty + 32 * k, c = c0 + 8 * tx;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (r < R && c < C) {             // C % 8 == 0 on this path: a group is entirely inside or outside
        load8<SRC_BF16>(x, r * ldx + c, v);
        apply_mul8(v, mul, r * ldm + c, mul_mode, mul_scale);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) tile[ty + 32 * k][8 * tx + i] = v[i];
    }
  }
  __syncthreads();
  const int ox = threadIdx.x & 31, oy = threadIdx.x >> 5;      // lane -> two consecutive r, warp -> c
  const int r = r0 + 2 * ox;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + oy + 8 * k;
    if (c < C && r < Rpad)            // Rpad is even (a multiple of 8)
      *reinterpret_cast<uint32_t*>(o + c * ldo + r) = pack_bf16x2(tile[2 * ox][oy + 8 * k], tile[2 * ox + 1][oy + 8 * k]);
  }
}

// rows whose length is even but not a multiple of 8 (the vocabulary: d logits [B*T, 9490] -> bf16 once per step, 62 MB
// that the one-element-per-thread kernel moved at 1.1 TB/s): one row per blockIdx.y, a float2 -> bf16x2 per thread
__global__ void __launch_bounds__(256)
convert_rows_bf16x2_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ mul, long long ldm,
                           int mul_mode, float mul_scale, __nv_bfloat16* __restrict__ o, long long ldo, int C2) {
  const int r = blockIdx.y;
  for (int c2 = blockIdx.x * 256 + threadIdx.x; c2 < C2; c2 += gridDim.x * 256) {
    float2 v = *reinterpret_cast<const float2*>(x + r * ldx + 2 * c2);
    if (mul_mode != 0) {
      const float2 m = *reinterpret_cast<const float2*>(mul + r * ldm + 2 * c2);
      v.x = mul_mode == 1 ? v.x * m.x : (m.x > 0.f ? v.x * mul_scale : 0.f);
      v.y = mul_mode == 1 ? v.y * m.y : (m.y > 0.f ? v.y * mul_scale : 0.f);
    }
    *reinterpret_cast<uint32_t*>(o + r * ldo + 2 * c2) = pack_bf16x2(v.x, v.y);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int convert_operand(const void* x_hi, const float* x_lo, int x_dtype, long long ldx, const float* mul,
                    long long ldm, int mul_mode, float mul_scale, void* o_hi, float* o_lo, int o_dtype, long long ldo,
                    int R, int C, int transpose, int Rpad, cudaStream_t stream) {
  if (R <= 0 || C <= 0) return CCX_OK;
  OpDst o{o_hi, o_lo, o_dtype};
  OpSrc x{x_hi, x_lo, x_dtype};
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)R * C * 8.0);
  // fast paths: bf16 destination, plain fp32 or bf16 source, 8-element groups, 16-byte aligned rows
  const bool src_bf16 = x_dtype == CCX_BF16;
  const int sx = src_bf16 ? 2 : 4;
  const bool fast = o_dtype == CCX_BF16 && x_lo == nullptr && (C % 8) == 0 && aligned16(x_hi) &&
                    ((ldx * sx) % 16) == 0 && (mul_mode == 0 || (aligned16(mul) && (ldm % 4) == 0));
  if (fast && !transpose && aligned16(o_hi) && ((ldo * 2) % 16) == 0) {
    const long long n = static_cast<long long>(R) * (C / 8);
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    if (src_bf16)
      convert_rows_bf16_kernel<true><<<grid, 256, 0, stream>>>(x_hi, ldx, mul, ldm, mul_mode, mul_scale,
                                                               static_cast<__nv_bfloat16*>(o_hi), ldo, R, C / 8);
    else
      convert_rows_bf16_kernel<false><<<grid, 256, 0, stream>>>(x_hi, ldx, mul, ldm, mul_mode, mul_scale,
                                                                static_cast<__nv_bfloat16*>(o_hi), ldo, R, C / 8);
    return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
  }
  if (fast && transpose && Rpad >= R && (Rpad % 2) == 0 && (reinterpret_cast<uintptr_t>(o_hi) & 3) == 0 &&
      (ldo % 2) == 0) {
    dim3 grid((Rpad + 63) / 64, (C + 63) / 64);
    if (src_bf16)
      convert_transpose_bf16_kernel<true><<<grid, 256, 0, stream>>>(x_hi, ldx, mul, ldm, mul_mode, mul_scale,
                                                                    static_cast<__nv_bfloat16*>(o_hi), ldo, R, C, Rpad);
    else
      convert_transpose_bf16_kernel<false><<<grid, 256, 0, stream>>>(x_hi, ldx, mul, ldm, mul_mode, mul_scale,
                                                                     static_cast<__nv_bfloat16*>(o_hi), ldo, R, C, Rpad);
    return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
  }
  if (!transpose && o_dtype == CCX_BF16 && !src_bf16 && x_lo == nullptr && (C % 2) == 0 && (ldx % 2) == 0 &&
      (ldo % 2) == 0 && R <= 65535 && (reinterpret_cast<uintptr_t>(x_hi) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(o_hi) & 3) == 0 &&
      (mul_mode == 0 || ((ldm % 2) == 0 && (reinterpret_cast<uintptr_t>(mul) & 7) == 0))) {
    const int C2 = C / 2;
    int gx = (C2 + 255) / 256;
    if (gx > 8) gx = 8;                       // a few pairs per thread on long rows
    convert_rows_bf16x2_kernel<<<dim3(gx, R), 256, 0, stream>>>(static_cast<const float*>(x_hi), ldx, mul, ldm, mul_mode,
                                                                mul_scale, static_cast<__nv_bfloat16*>(o_hi), ldo, C2);
    return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
  }
  if (!transpose) {
    const long long n = static_cast<long long>(R) * C;
    convert_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(x, ldx, mul, ldm, mul_mode, mul_scale,
                                                                                  o, ldo, R, C);
  } else {
    if (Rpad < R) return CCX_ERR_SHAPE;
    dim3 grid((Rpad + 31) / 32, (C + 31) / 32);
    convert_transpose_kernel<<<grid, 256, 0, stream>>>(x, ldx, mul, ldm, mul_mode, mul_scale, o, ldo, R, C, Rpad);
  }
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// column sums (bias gradient), accumulating
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_acc_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ mul, long long ldm,
                  int mul_mode, float mul_scale, float* __restrict__ out, int R, int C, int rows_per_block) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  const int r_begin = blockIdx.y * rows_per_block;
  const int r_end = min(R, r_begin + rows_per_block);
  float acc = 0.f;
  if (c < C)
    for (int r = r_begin + ty; r < r_end; r += 8) acc += apply_mul(x[r * ldx + c], mul, ldm, r, c, mul_mode, mul_scale);
  red[ty][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += red[j][threadIdx.x];
    atomicAdd(out + c, s);
  }
}

// 128 columns per CTA (one float4 per lane), 8 warps stride the rows, four independent loads in flight per thread
__global__ void __launch_bounds__(256)
colsum_acc4_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ mul, long long ldm,
                   int mul_mode, float mul_scale, float* __restrict__ out, int R, int C, int rows_per_block) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + 4 * lane;
  const int r_begin = blockIdx.y * rows_per_block;
  const int r_end = min(R, r_begin + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    for (int r = r_begin + ty; r < r_end; r += 32) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rr = r + 8 * k;
        v[k] = rr < r_end ? *reinterpret_cast<const float4*>(x + rr * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (mul_mode != 0 && rr < r_end) {
          const float4 m = *reinterpret_cast<const float4*>(mul + rr * ldm + c);
          if (mul_mode == 1) { v[k].x *= m.x; v[k].y *= m.y; v[k].z *= m.z; v[k].w *= m.w; }
          else {
            v[k].x = m.x > 0.f ? v[k].x * mul_scale : 0.f; v[k].y = m.y > 0.f ? v[k].y * mul_scale : 0.f;
            v[k].z = m.z > 0.f ? v[k].z * mul_scale : 0.f; v[k].w = m.w > 0.f ? v[k].w * mul_scale : 0.f;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
    }
  }
  red[ty][lane] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    float4 s4 = red[0][lane];
#pragma unroll
    for (int j = 1; j < 8; ++j) { s4.x += red[j][lane].x; s4.y += red[j][lane].y; s4.z += red[j][lane].z; s4.w += red[j][lane].w; }
    atomicAdd(out + c, s4.x); atomicAdd(out + c + 1, s4.y); atomicAdd(out + c + 2, s4.z); atomicAdd(out + c + 3, s4.w);
  }
}

int colsum_acc(const float* x, long long ldx, const float* mul, long long ldm, int mul_mode, float mul_scale,
               float* out, int R, int C, cudaStream_t stream) {
  if (R <= 0 || C <= 0) return CCX_OK;
  if ((C % 4) == 0 && (ldx % 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (mul_mode == 0 || ((ldm % 4) == 0 && (reinterpret_cast<uintptr_t>(mul) & 15) == 0))) {
    // rows per CTA chosen so that the grid has a few CTAs per SM even for narrow matrices
    const int col_blocks = (C + 127) / 128;
    int rpb = 512;
    while (rpb > 64 && static_cast<long long>(col_blocks) * ((R + rpb - 1) / rpb) < 148 * 4) rpb /= 2;
    dim3 grid(col_blocks, (R + rpb - 1) / rpb);
    ProfScope prof(PROF_ELEMENTWISE, stream, (double)R * C * 4.0);
    colsum_acc4_kernel<<<grid, 256, 0, stream>>>(x, ldx, mul, ldm, mul_mode, mul_scale, out, R, C, rpb);
    return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
  }
  const int rpb = 256;
  dim3 grid((C + 31) / 32, (R + rpb - 1) / rpb);
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)R * C * 4.0);
  colsum_acc_kernel<<<grid, 256, 0, stream>>>(x, ldx, mul, ldm, mul_mode, mul_scale, out, R, C, rpb);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// Linear backward, first pass over dY: the bf16 GEMM operand AND the bias gradient (column sums) from ONE read.
// Every Linear of the backward pass used to run convert_rows (read dY, write bf16) and then colsum_acc (read dY
// again): 14 such pairs per LSTM train step, 43 per Transformer step.  Same tiling as colsum_acc4_kernel — 128 columns
// per CTA (one float4 per lane), 8 warps stride the rows — plus an 8-byte bf16 store per element group.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
convert_colsum_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ mul, long long ldm,
                      int mul_mode, float mul_scale, __nv_bfloat16* __restrict__ o, long long ldo,
                      float* __restrict__ sums, int R, int C, int rows_per_block) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + 4 * lane;
  const int r_begin = blockIdx.y * rows_per_block;
  const int r_end = min(R, r_begin + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    for (int r = r_begin + ty; r < r_end; r += 32) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rr = r + 8 * k;
        v[k] = rr < r_end ? *reinterpret_cast<const float4*>(x + rr * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (mul_mode != 0 && rr < r_end) {
          const float4 m = *reinterpret_cast<const float4*>(mul + rr * ldm + c);
          if (mul_mode == 1) { v[k].x *= m.x; v[k].y *= m.y; v[k].z *= m.z; v[k].w *= m.w; }
          else {
            v[k].x = m.x > 0.f ? v[k].x * mul_scale : 0.f; v[k].y = m.y > 0.f ? v[k].y * mul_scale : 0.f;
            v[k].z = m.z > 0.f ? v[k].z * mul_scale : 0.f; v[k].w = m.w > 0.f ? v[k].w * mul_scale : 0.f;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rr = r + 8 * k;
        if (rr < r_end) {
          uint2 pk;
          pk.x = pack_bf16x2(v[k].x, v[k].y);
          pk.y = pack_bf16x2(v[k].z, v[k].w);
          *reinterpret_cast<uint2*>(o + rr * ldo + c) = pk;
        }
        acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w;
      }
    }
  }
  red[ty][lane] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    float4 s4 = red[0][lane];
#pragma unroll
    for (int j = 1; j < 8; ++j) { s4.x += red[j][lane].x; s4.y += red[j][lane].y; s4.z += red[j][lane].z; s4.w += red[j][lane].w; }
    atomicAdd(sums + c, s4.x); atomicAdd(sums + c + 1, s4.y); atomicAdd(sums + c + 2, s4.z); atomicAdd(sums + c + 3, s4.w);
  }
}

int convert_colsum(const float* x, long long ldx, const float* mul, long long ldm, int mul_mode, float mul_scale,
                   void* o_bf16, long long ldo, float* sums, int R, int C, cudaStream_t stream) {
  if (R <= 0 || C <= 0) return CCX_OK;
  if ((C % 4) != 0 || (ldx % 4) != 0 || (ldo % 4) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(o_bf16) & 7) != 0 ||
      (mul_mode != 0 && ((ldm % 4) != 0 || (reinterpret_cast<uintptr_t>(mul) & 15) != 0)))
    return CCX_ERR_SHAPE;
  const int col_blocks = (C + 127) / 128;
  int rpb = 512;
  while (rpb > 64 && static_cast<long long>(col_blocks) * ((R + rpb - 1) / rpb) < 148 * 4) rpb /= 2;
  dim3 grid(col_blocks, (R + rpb - 1) / rpb);
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)R * C * 6.0);
  convert_colsum_kernel<<<grid, 256, 0, stream>>>(x, ldx, mul, ldm, mul_mode, mul_scale,
                                                  static_cast<__nv_bfloat16*>(o_bf16), ldo, sums, R, C, rpb);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward: one warp per row; dgamma / dbeta reduced per block then atomically accumulated
// ---------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
              float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int C,
              float eps, int merge, int H, int W) {
  __shared__ float s_dg[VPL * 128], s_db[VPL * 128];
  for (int i = threadIdx.x; i < VPL * 128; i += 256) { s_dg[i] = 0.f; s_db[i] = 0.f; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (m < M) {
    // dy addressing: plain rows, or the 2x2 patch-merged layout written by ln_rows_kernel(merge=1)
    const float* dyr = dy + m * C;
    if (merge) {
      const int w = static_cast<int>(m % W);
      const int h = static_cast<int>((m / W) % H);
      const long long b = m / (static_cast<long long>(W) * H);
      dyr = dy + ((b * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1)) * (4LL * C) + ((h & 1) * 2 + (w & 1)) * C;
    }
    float4 xv[VPL], gv[VPL], dv[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xv[i] = __ldg(reinterpret_cast<const float4*>(x + m * C) + i * 32 + lane);
      dv[i] = __ldg(reinterpret_cast<const float4*>(dyr) + i * 32 + lane);
      gv[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
      s += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      q += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    // xhat = (x-mean)*rstd ; g = dy*gamma ; dx = rstd * (g - mean(g) - xhat * mean(g*xhat))
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;
      const float g0 = dv[i].x * gv[i].x, g1 = dv[i].y * gv[i].y, g2 = dv[i].z * gv[i].z, g3 = dv[i].w * gv[i].w;
      sg += g0 + g1 + g2 + g3;
      sgx += g0 * xv[i].x + g1 * xv[i].y + g2 * xv[i].z + g3 * xv[i].w;
    }
    sg = warp_sum(sg) / C;
    sgx = warp_sum(sgx) / C;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 o;
      o.x = rstd * (dv[i].x * gv[i].x - sg - xv[i].x * sgx);
      o.y = rstd * (dv[i].y * gv[i].y - sg - xv[i].y * sgx);
      o.z = rstd * (dv[i].z * gv[i].z - sg - xv[i].z * sgx);
      o.w = rstd * (dv[i].w * gv[i].w - sg - xv[i].w * sgx);
      reinterpret_cast<float4*>(dx + m * C)[i * 32 + lane] = o;
      const int c = (i * 32 + lane) * 4;
      atomicAdd(&s_dg[c + 0], dv[i].x * xv[i].x); atomicAdd(&s_db[c + 0], dv[i].x);
      atomicAdd(&s_dg[c + 1], dv[i].y * xv[i].y); atomicAdd(&s_db[c + 1], dv[i].y);
      atomicAdd(&s_dg[c + 2], dv[i].z * xv[i].z); atomicAdd(&s_db[c + 2], dv[i].z);
      atomicAdd(&s_dg[c + 3], dv[i].w * xv[i].w); atomicAdd(&s_db[c + 3], dv[i].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    if (dgamma) atomicAdd(dgamma + i, s_dg[i]);
    if (dbeta) atomicAdd(dbeta + i, s_db[i]);
  }
}

int ln_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
           long long M, int C, float eps, cudaStream_t stream, int merge, int H, int W) {
  if (M <= 0) return CCX_OK;
  if (C % 128 != 0 || C > 1024) return CCX_ERR_SHAPE;
  const unsigned grid = static_cast<unsigned>((M + 7) / 8);
  ProfScope prof(PROF_LN_ROWS, stream, (double)M * C * 12.0);
#define CCX_LNB_CASE(V)                                                                              \
  case V:                                                                                            \
    ln_bwd_kernel<V><<<grid, 256, 0, stream>>>(dy, x, gamma, dx, dgamma, dbeta, M, C, eps, merge, H, W); \
    break;
  switch (C / 128) {
    CCX_LNB_CASE(1) CCX_LNB_CASE(2) CCX_LNB_CASE(3) CCX_LNB_CASE(4) CCX_LNB_CASE(5) CCX_LNB_CASE(6)
    CCX_LNB_CASE(7) CCX_LNB_CASE(8)
    default: return CCX_ERR_SHAPE;
  }
#undef CCX_LNB_CASE
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// attention backward, one CTA per (batch, head)
// ---------------------------------------------------------------------------------------------
struct MhaBwdArgs {
  const float* q; long long q_sb, q_st;
  const float* k; long long k_sb, k_st;
  const float* v; long long v_sb, v_st;
  const float* dctx; long long d_sb, d_st;   // [B, Tq, H*hd] fp32
  const float* probs;                        // [B,H,Tq,Tk] softmax (before dropout)
  const float* prob_mask;                    // dropout multiplier or nullptr
  float* dq; long long dq_sb, dq_st;
  float* dk; long long dk_sb, dk_st;         // written (not accumulated): one CTA owns (b, h)
  float* dv; long long dv_sb, dv_st;
  int B, H, Tq, Tk, hd;
  float scale;
};

__global__ void __launch_bounds__(128)
mha_bwd_kernel(MhaBwdArgs a) {
  extern __shared__ float bsm[];
  const int hd = a.hd, ld = hd + 1;
  float* s_q = bsm;                       // [Tq][ld]
  float* s_do = s_q + a.Tq * ld;          // [Tq][ld]
  float* s_k = s_do + a.Tq * ld;          // [Tk][ld]
  float* s_v = s_k + a.Tk * ld;           // [Tk][ld]
  float* s_p = s_v + a.Tk * ld;           // [Tq][Tk]  P*mask   (then reused)
  float* s_ds = s_p + a.Tq * a.Tk;        // [Tq][Tk]  dS
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int tid = threadIdx.x;
  const int hd4 = hd >> 2;   // 16-byte loads, several in flight (the staging loops are latency bound)
#pragma unroll 4
  for (int i = tid; i < a.Tq * hd4; i += 128) {
    const int r = i / hd4, d = (i - r * hd4) * 4;
    const float4 qq = __ldg(reinterpret_cast<const float4*>(a.q + b * a.q_sb + r * a.q_st + h * hd + d));
    const float4 dd = __ldg(reinterpret_cast<const float4*>(a.dctx + b * a.d_sb + r * a.d_st + h * hd + d));
    float* pq = s_q + r * ld + d;
    float* pd = s_do + r * ld + d;
    pq[0] = qq.x; pq[1] = qq.y; pq[2] = qq.z; pq[3] = qq.w;
    pd[0] = dd.x; pd[1] = dd.y; pd[2] = dd.z; pd[3] = dd.w;
  }
#pragma unroll 4
  for (int i = tid; i < a.Tk * hd4; i += 128) {
    const int r = i / hd4, d = (i - r * hd4) * 4;
    const float4 kk = __ldg(reinterpret_cast<const float4*>(a.k + b * a.k_sb + r * a.k_st + h * hd + d));
    const float4 vv = __ldg(reinterpret_cast<const float4*>(a.v + b * a.v_sb + r * a.v_st + h * hd + d));
    float* pk = s_k + r * ld + d;
    float* pv = s_v + r * ld + d;
    pk[0] = kk.x; pk[1] = kk.y; pk[2] = kk.z; pk[3] = kk.w;
    pv[0] = vv.x; pv[1] = vv.y; pv[2] = vv.z; pv[3] = vv.w;
  }
  const long long pbase = (static_cast<long long>(b) * a.H + h) * a.Tq * a.Tk;
  // stage softmax probabilities (and the dropout multipliers) once, coalesced, instead of a global round trip per
  // (row, key) inside the reduction loops:  s_ds <- P,  s_p <- mask (or 1)
  for (int i = tid; i < a.Tq * a.Tk; i += 128) {
    s_ds[i] = __ldg(a.probs + pbase + i);
    s_p[i] = a.prob_mask ? __ldg(a.prob_mask + pbase + i) : 1.f;
  }
  __syncthreads();
  // dPd[i,j] = dctx_i . v_j ; dP = dPd*mask ; dS = P * (dP - sum_j dP*P) ; afterwards s_p = P*mask, s_ds = dS
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < a.Tq; i += 4) {
    float rowdot = 0.f;
    float dp_keep[8];   // Tk <= 256 keys -> at most 8 per lane
    int n = 0;
    for (int j = lane; j < a.Tk; j += 32, ++n) {
      float acc = 0.f;
#pragma unroll 8
      for (int d = 0; d < hd; ++d) acc = fmaf(s_do[i * ld + d], s_v[j * ld + d], acc);
      const float pr = s_ds[i * a.Tk + j], m = s_p[i * a.Tk + j];
      const float dp = acc * m;
      dp_keep[n & 7] = dp;
      rowdot += dp * pr;
    }
    rowdot = warp_sum(rowdot);
    n = 0;
    for (int j = lane; j < a.Tk; j += 32, ++n) {
      const float pr = s_ds[i * a.Tk + j], m = s_p[i * a.Tk + j];
      s_p[i * a.Tk + j] = pr * m;                       // Pd for dV
      s_ds[i * a.Tk + j] = pr * (dp_keep[n & 7] - rowdot);
    }
  }
  __syncthreads();
  // dq[i,d] = scale * sum_j dS[i,j] k[j,d]
  for (int idx = tid; idx < a.Tq * hd; idx += 128) {
    const int i = idx / hd, d = idx - i * hd;
    float acc = 0.f;
    for (int j = 0; j < a.Tk; ++j) acc = fmaf(s_ds[i * a.Tk + j], s_k[j * ld + d], acc);
    a.dq[b * a.dq_sb + i * a.dq_st + h * hd + d] = acc * a.scale;
  }
  // dk[j,d] = scale * sum_i dS[i,j] q[i,d] ; dv[j,d] = sum_i Pd[i,j] dctx[i,d]
  for (int idx = tid; idx < a.Tk * hd; idx += 128) {
    const int j = idx / hd, d = idx - j * hd;
    float ak = 0.f, av = 0.f;
    for (int i = 0; i < a.Tq; ++i) {
      ak = fmaf(s_ds[i * a.Tk + j], s_q[i * ld + d], ak);
      av = fmaf(s_p[i * a.Tk + j], s_do[i * ld + d], av);
    }
    a.dk[b * a.dk_sb + j * a.dk_st + h * hd + d] = ak * a.scale;
    a.dv[b * a.dv_sb + j * a.dv_st + h * hd + d] = av;
  }
}

int mha_bwd(const float* q, long long q_sb, long long q_st, const float* k, long long k_sb, long long k_st,
            const float* v, long long v_sb, long long v_st, const float* dctx, long long d_sb, long long d_st,
            const float* probs, const float* prob_mask, float* dq, long long dq_sb, long long dq_st, float* dk,
            long long dk_sb, long long dk_st, float* dv, long long dv_sb, long long dv_st, int B, int H, int Tq,
            int Tk, int hd, float scale, cudaStream_t stream) {
  if (B <= 0 || Tq <= 0 || Tk <= 0) return CCX_OK;
  if ((hd & 3) || ((q_sb | q_st | k_sb | k_st | v_sb | v_st | d_sb | d_st) & 3) || Tk > 256) return CCX_ERR_SHAPE;
  const size_t smem = (static_cast<size_t>(2) * (Tq + Tk) * (hd + 1) + 2 * static_cast<size_t>(Tq) * Tk) * 4;
  if (smem > 200 * 1024) return CCX_ERR_SHAPE;
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.ref();
  if (!configured) {
    if (cudaFuncSetAttribute(mha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return CCX_ERR_CUDA;
    configured = true;
  }
  MhaBwdArgs a;
  a.q = q; a.q_sb = q_sb; a.q_st = q_st; a.k = k; a.k_sb = k_sb; a.k_st = k_st; a.v = v; a.v_sb = v_sb; a.v_st = v_st;
  a.dctx = dctx; a.d_sb = d_sb; a.d_st = d_st; a.probs = probs; a.prob_mask = prob_mask;
  a.dq = dq; a.dq_sb = dq_sb; a.dq_st = dq_st; a.dk = dk; a.dk_sb = dk_sb; a.dk_st = dk_st;
  a.dv = dv; a.dv_sb = dv_sb; a.dv_st = dv_st;
  a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.hd = hd; a.scale = scale;
  ProfScope prof(PROF_ATTENTION, stream, (double)B * H * (Tq + Tk) * hd * 16.0);
  mha_bwd_kernel<<<B * H, 128, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// fused softmax cross-entropy (mean over the selected rows) forward + backward, one CTA per row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ targets, int V,
                  float inv_n, const float* __restrict__ n_valid_dev, float* __restrict__ loss_sum,
                  float* __restrict__ dlogits, long long ldd, float* __restrict__ stats, int topk) {
  __shared__ float red[8], red2[8];
  __shared__ float s_b, s_c;
  if (n_valid_dev != nullptr) inv_n = 1.0f / fmaxf(__ldg(n_valid_dev), 1.0f);   // row count known only on the device
  const long long r = blockIdx.x;
  const long long tgt = targets[r];
  float* drow = dlogits ? dlogits + r * ldd : nullptr;
  if (tgt < 0 || tgt >= V) {
    if (drow) for (int v = threadIdx.x; v < V; v += 256) drow[v] = 0.f;
    return;
  }
  const float* row = logits + r * ld;
  const float tval = row[tgt];
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += 256) mx = fmaxf(mx, row[v]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) { float m = red[0]; for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]); s_b = m; }
  __syncthreads();
  mx = s_b;
  float sum = 0.f, gt = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) {
    const float x = row[v];
    sum += expf(x - mx);
    gt += (x > tval) ? 1.f : 0.f;     // rank of the target = number of strictly larger logits
  }
  sum = warp_sum(sum);
  gt = warp_sum(gt);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = sum; red2[threadIdx.x >> 5] = gt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { s += red[w]; c += red2[w]; }
    s_b = s; s_c = c;
  }
  __syncthreads();
  sum = s_b;
  const float lse = mx + logf(sum);
  if (threadIdx.x == 0) {
    if (loss_sum) atomicAdd(loss_sum, (lse - tval) * inv_n);
    if (stats) {
      atomicAdd(stats + 0, lse - tval);                       // un-normalised token loss sum
      atomicAdd(stats + 1, 1.f);                              // valid rows (tokens)
      atomicAdd(stats + 2, s_c < static_cast<float>(topk) ? 1.f : 0.f);   // top-k hits (utils/utils.py:239-254)
    }
  }
  if (drow) {
    const float inv = inv_n / sum;
    for (int v = threadIdx.x; v < V; v += 256) drow[v] = expf(row[v] - mx) * inv - (v == tgt ? inv_n : 0.f);
  }
}

int softmax_ce(const float* logits, long long ld, const long long* targets, long long R, int V, float inv_n,
               float* loss_sum, float* dlogits, long long ldd, float* stats, int topk, cudaStream_t stream,
               const float* n_valid_dev) {
  if (R <= 0) return CCX_OK;
  ProfScope prof(PROF_LOSS, stream, (double)R * V * (dlogits ? 8.0 : 4.0));
  softmax_ce_kernel<<<static_cast<unsigned>(R), 256, 0, stream>>>(logits, ld, targets, V, inv_n, n_valid_dev, loss_sum,
                                                                 dlogits, ldd, stats, topk);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// Targets of the free-running (no teacher forcing) evaluation, utils/utils.py:261-295
// (preprocessDecoderOutputForMetrics) without the per-sample Python loop: row i is scored for t < L_i where
// L_i = (first <end> in sequences[i]) + 1, or maxDecodeLen if none; positions whose ground truth is <pad> are dropped.
__global__ void free_running_targets_kernel(const long long* __restrict__ sequences, const long long* __restrict__ caps,
                                            long long cap_ld, long long* __restrict__ targets,
                                            int* __restrict__ decode_len, int B, int T, int cap_T, long long end_tok,
                                            long long pad_tok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  int L = T;
  for (int t = 0; t < T; ++t)
    if (sequences[static_cast<long long>(i) * T + t] == end_tok) { L = t + 1; break; }
  if (decode_len) decode_len[i] = L;
  for (int t = 0; t < T; ++t) {
    long long g = -1;
    if (t < L && 1 + t < cap_T) {
      g = caps[i * cap_ld + 1 + t];
      if (g == pad_tok) g = -1;
    }
    targets[static_cast<long long>(i) * T + t] = g;
  }
}
int free_running_targets(const long long* sequences, const long long* caps, long long cap_ld, long long* targets,
                         int* decode_len, int B, int T, int cap_T, long long end_tok, long long pad_tok,
                         cudaStream_t stream) {
  if (B <= 0) return CCX_OK;
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)B * T * 24.0);
  free_running_targets_kernel<<<(B + 127) / 128, 128, 0, stream>>>(sequences, caps, cap_ld, targets, decode_len, B, T,
                                                                 cap_T, end_tok, pad_tok);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// embedding backward: dTable[token[r]] += dX[r] * mul[r]   (dense gradient like nn.Embedding(sparse=False))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embedding_bwd_kernel(const long long* __restrict__ tokens, long long tok_ld, int t0, const float* __restrict__ dx,
                     long long sb, long long st, const float* __restrict__ dropmask, float* __restrict__ dtable,
                     int V, int D, int nb, int nt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (r >= static_cast<long long>(nb) * nt) return;
  const int b = static_cast<int>(r / nt), t = static_cast<int>(r % nt);
  long long tok = tokens[b * tok_ld + t0 + t];
  if (tok < 0 || tok >= V) return;
  for (int d = lane; d < D; d += 32) {
    float g = dx[b * sb + t * st + d];
    if (dropmask) g *= dropmask[r * D + d];
    atomicAdd(dtable + tok * D + d, g);
  }
}

int embedding_bwd(const long long* tokens, long long tok_ld, int t0, const float* dx, long long sb, long long st,
                  const float* dropmask, float* dtable, int V, int D, int nb, int nt, cudaStream_t stream) {
  if (nb <= 0 || nt <= 0) return CCX_OK;
  const long long rows = static_cast<long long>(nb) * nt;
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)rows * D * 12.0);
  embedding_bwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(tokens, tok_ld, t0, dx, sb, st,
                                                                                dropmask, dtable, V, D, nb, nt);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// LSTM cell backward (point-wise half).  dh = dh_fc * dropmask + dh_carry ; see lstm_pointwise_kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lstm_pointwise_bwd_kernel(const float* __restrict__ gates, long long ldg, const float* __restrict__ c_prev,
                          const float* __restrict__ c_new, const float* __restrict__ dh_fc, long long ld_fc,
                          const float* __restrict__ dropmask, long long ld_dm, const float* __restrict__ dh_carry,
                          float* __restrict__ dc_carry,  // in: dL/dc_new, out: dL/dc_prev   [bt, D]
                          float* __restrict__ dgates, long long lddg, OpDst dg_op, long long ld_op,
                          float* __restrict__ zero_rows, long long ld_zero, int n_zero, int bt, int D) {
  grid_dep_sync();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(bt) * D) return;
  const int b = static_cast<int>(idx / D), j = static_cast<int>(idx % D);
  const float* g = gates + b * ldg;
  const float i_ = 1.0f / (1.0f + expf(-g[j]));
  const float f_ = 1.0f / (1.0f + expf(-g[D + j]));
  const float g_ = tanhf(g[2 * D + j]);
  const float o_ = 1.0f / (1.0f + expf(-g[3 * D + j]));
  const float tc = tanhf(c_new[static_cast<long long>(b) * D + j]);
  float dh = dh_carry ? dh_carry[static_cast<long long>(b) * D + j] : 0.f;
  if (dh_fc) dh += dh_fc[b * ld_fc + j] * (dropmask ? dropmask[b * ld_dm + j] : 1.f);
  const float dc = dc_carry[static_cast<long long>(b) * D + j] + dh * o_ * (1.f - tc * tc);
  float* dg = dgates + b * lddg;
  const float d0 = dc * g_ * i_ * (1.f - i_);
  const float d1 = dc * c_prev[static_cast<long long>(b) * D + j] * f_ * (1.f - f_);
  const float d2 = dc * i_ * (1.f - g_ * g_);
  const float d3 = dh * tc * o_ * (1.f - o_);
  dg[j] = d0; dg[D + j] = d1; dg[2 * D + j] = d2; dg[3 * D + j] = d3;
  if (dg_op.hi != nullptr) {   // the same values as the A operand of the dgrad GEMM (saves a conversion launch)
    put(dg_op, b * ld_op + j, d0); put(dg_op, b * ld_op + D + j, d1);
    put(dg_op, b * ld_op + 2 * D + j, d2); put(dg_op, b * ld_op + 3 * D + j, d3);
  }
  if (zero_rows != nullptr && j < n_zero) zero_rows[b * ld_zero + j] = 0.f;   // attention-backward accumulators
  dc_carry[static_cast<long long>(b) * D + j] = dc * f_;
}

int lstm_pointwise_bwd(const float* gates, long long ldg, const float* c_prev, const float* c_new,
                       const float* dh_fc, long long ld_fc, const float* dropmask, long long ld_dm,
                       const float* dh_carry, float* dc_carry, float* dgates, long long lddg, int bt, int D,
                       cudaStream_t stream, void* dg_op_hi, float* dg_op_lo, int dg_op_dtype, long long ld_op,
                       float* zero_rows, long long ld_zero, int n_zero) {
  if (bt <= 0) return CCX_OK;
  if (n_zero > D) return CCX_ERR_SHAPE;
  OpDst dg_op{dg_op_hi, dg_op_lo, dg_op_dtype};
  const long long n = static_cast<long long>(bt) * D;
  ProfScope prof(PROF_LSTM, stream, (double)n * 48.0);
  return launch_pdl(lstm_pointwise_bwd_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, stream,
                    gates, ldg, c_prev, c_new, dh_fc, ld_fc, dropmask, ld_dm, dh_carry, dc_carry, dgates, lddg, dg_op,
                    ld_op, zero_rows, ld_zero, n_zero, bt, D) == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// Bahdanau attention step backward (mirror of bahdanau_attention_kernel), two kernels so that a step with only
// <= 32 active samples still fills the machine:
//   K1  grid (bt, E/128), one thread per encoder channel e: the 49 enc values of (b, :, e) live in registers;
//       un-gated awe, gate, d gate-pre-activation, d_awe_raw, d_enc += alpha_p * d_awe_raw, and the partial
//       d alpha_p = sum_e d_awe_raw[e] enc[p,e] (warp reduce + atomicAdd into d_hg[b, 0:P], used as scratch)
//   K2  grid bt, one thread per attention unit a: softmax backward, then through w_f . relu(att1 + att2)
// ---------------------------------------------------------------------------------------------
static constexpr int ATTB_P = 64;   // register-resident pixel count of K1 (P <= 64 fast path)

template <int PMAX>
__global__ void __launch_bounds__(128)
attention_bwd_enc_kernel(const float* __restrict__ hg, long long ldhg, const float* __restrict__ enc,
                         const float* __restrict__ alpha, long long alpha_ld, const float* __restrict__ d_out,
                         long long ld_dout, float* __restrict__ d_hg, long long ld_dhg, float* __restrict__ d_enc,
                         OpDst dhg_op, long long ld_op, int P, int A, int E,
                         float* __restrict__ dawe_out,      // deferred mode: [bt, E] d_awe_raw of this step
                         float* __restrict__ dalpha_acc) {  // deferred mode: [bt, P] d alpha accumulator (zeroed)
  __shared__ float s_a[PMAX];
  grid_dep_sync();
  const int b = blockIdx.x;
  const int e = blockIdx.y * 128 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  for (int p = threadIdx.x; p < P; p += 128) s_a[p] = alpha[b * alpha_ld + p];
  __syncthreads();
  const bool ok = e < E;
  float x[PMAX];
  float awe = 0.f;
#pragma unroll
  for (int p = 0; p < PMAX; ++p) {
    x[p] = (ok && p < P) ? __ldg(enc + (static_cast<long long>(b) * P + p) * E + e) : 0.f;
    awe = fmaf(x[p], (p < P) ? s_a[p] : 0.f, awe);
  }
  float draw = 0.f;
  if (ok) {
    const float gate = 1.0f / (1.0f + expf(-hg[b * ldhg + A + e]));
    const float dout = d_out[b * ld_dout + e];
    const float dgp = dout * awe * gate * (1.f - gate);
    d_hg[b * ld_dhg + A + e] = dgp;
    if (dhg_op.hi != nullptr) put(dhg_op, b * ld_op + A + e, dgp);
    draw = dout * gate;
    if (dawe_out != nullptr) dawe_out[static_cast<long long>(b) * E + e] = draw;
  }
  float* acc = dalpha_acc != nullptr ? dalpha_acc + static_cast<long long>(b) * P : d_hg + b * ld_dhg;
  const bool rmw = ok && d_enc != nullptr && dawe_out == nullptr;
#pragma unroll
  for (int p = 0; p < PMAX; ++p) {
    if (p >= P) break;
    const float part = warp_sum(draw * x[p]);
    if (lane == 0) atomicAdd(acc + p, part);          // d alpha_p accumulator (d_hg[b, 0:P] as scratch, or dalpha_acc)
    if (rmw) d_enc[(static_cast<long long>(b) * P + p) * E + e] += s_a[p] * draw;
  }
}

// Deferred mode of K2: grid (bt, A/128), one thread per attention unit; every CTA re-derives the 49 softmax-backward
// values (cheap) so a step with <= 32 active rows still spreads over 4x as many SMs; no d_att1 read-modify-write.
__global__ void __launch_bounds__(128)
attention_bwd_att_split_kernel(const float* __restrict__ att1, const float* __restrict__ hg, long long ldhg,
                               const float* __restrict__ w_f, const float* __restrict__ alpha, long long alpha_ld,
                               const float* __restrict__ d_alpha_ext, long long dalpha_ld,
                               const float* __restrict__ dalpha_acc, float* __restrict__ de_out,
                               float* __restrict__ d_hg, long long ld_dhg, float* __restrict__ d_wf, OpDst dhg_op,
                               long long ld_op, int P, int A) {
  __shared__ float s_de[ATT_MAX_P_BWD];
  __shared__ float s_red[4];
  grid_dep_sync();
  const int b = blockIdx.x;
  float dot = 0.f;
  for (int p = threadIdx.x; p < P; p += 128) {
    const float a_p = alpha[b * alpha_ld + p];
    const float da_p = dalpha_acc[static_cast<long long>(b) * P + p] +
                       (d_alpha_ext ? d_alpha_ext[b * dalpha_ld + p] : 0.f);
    s_de[p] = da_p;
    dot = fmaf(a_p, da_p, dot);
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dot;
  __syncthreads();
  dot = s_red[0] + s_red[1] + s_red[2] + s_red[3];
  for (int p = threadIdx.x; p < P; p += 128) {
    const float de = alpha[b * alpha_ld + p] * (s_de[p] - dot);      // each p is owned by one thread: no hazard
    s_de[p] = de;
    if (blockIdx.y == 0) de_out[static_cast<long long>(b) * P + p] = de;
  }
  __syncthreads();
  const int a = blockIdx.y * 128 + threadIdx.x;
  if (a >= A) return;
  float datt2 = 0.f, dwf = 0.f;
  const float h2 = hg[b * ldhg + a], wf = __ldg(w_f + a);
  const float* a1 = att1 + static_cast<long long>(b) * P * A + a;
#pragma unroll 7
  for (int p = 0; p < P; ++p) {
    const float pre = __ldg(a1 + static_cast<long long>(p) * A) + h2;
    if (pre > 0.f) {
      const float de = s_de[p];
      datt2 = fmaf(de, wf, datt2);
      dwf = fmaf(de, pre, dwf);
    }
  }
  d_hg[b * ld_dhg + a] = datt2;
  if (dhg_op.hi != nullptr) put(dhg_op, b * ld_op + a, datt2);
  atomicAdd(d_wf + a, dwf);
}

// After the time loop (deferred mode): the two accumulations that the per-step kernels would otherwise do as
// read-modify-write passes over d_enc [B,P,E] and d_att1 [B,P,A] at every step.
//   d_enc[b,p,e]  += sum_t alpha[b,t,p] * dawe[t,b,e]
//   d_att1[b,p,a] += w_f[a] * sum_t de[t,b,p] * [att1[b,p,a] + att2[t,b,a] > 0]
// grid (B, channels/128, ceil(P/32)); 32 pixels per CTA live in registers; rows / steps that never ran hold zeros.
static constexpr int FIN_P = 32;

__global__ void __launch_bounds__(128)
attention_bwd_finish_enc_kernel(const float* __restrict__ alphas, long long a_sb, long long a_st,
                                const float* __restrict__ dawe, float* __restrict__ d_enc, int B, int T, int P,
                                int E) {
  extern __shared__ float s_al[];           // [T][FIN_P]
  const int b = blockIdx.x, e = blockIdx.y * 128 + threadIdx.x, p0 = blockIdx.z * FIN_P;
  for (int i = threadIdx.x; i < T * FIN_P; i += 128) {
    const int t = i / FIN_P, p = p0 + i % FIN_P;
    s_al[i] = p < P ? alphas[b * a_sb + t * a_st + p] : 0.f;
  }
  __syncthreads();
  if (e >= E) return;
  float acc[FIN_P];
#pragma unroll
  for (int p = 0; p < FIN_P; ++p) acc[p] = 0.f;
  for (int t = 0; t < T; ++t) {
    const float x = __ldg(dawe + (static_cast<long long>(t) * B + b) * E + e);
#pragma unroll
    for (int p = 0; p < FIN_P; ++p) acc[p] = fmaf(s_al[t * FIN_P + p], x, acc[p]);
  }
#pragma unroll
  for (int p = 0; p < FIN_P; ++p)
    if (p0 + p < P) d_enc[(static_cast<long long>(b) * P + p0 + p) * E + e] += acc[p];
}

__global__ void __launch_bounds__(128)
attention_bwd_finish_att_kernel(const float* __restrict__ att1, const float* __restrict__ hg_all, long long ldhg,
                                const float* __restrict__ w_f, const float* __restrict__ de_all,
                                float* __restrict__ d_att1, int B, int T, int P, int A) {
  extern __shared__ float s_de[];           // [T][FIN_P]
  const int b = blockIdx.x, a = blockIdx.y * 128 + threadIdx.x, p0 = blockIdx.z * FIN_P;
  for (int i = threadIdx.x; i < T * FIN_P; i += 128) {
    const int t = i / FIN_P, p = p0 + i % FIN_P;
    s_de[i] = p < P ? de_all[(static_cast<long long>(t) * B + b) * P + p] : 0.f;
  }
  __syncthreads();
  if (a >= A) return;
  float x[FIN_P], acc[FIN_P];
#pragma unroll
  for (int p = 0; p < FIN_P; ++p) {
    x[p] = (p0 + p < P) ? __ldg(att1 + (static_cast<long long>(b) * P + p0 + p) * A + a) : -INFINITY;
    acc[p] = 0.f;
  }
  for (int t = 0; t < T; ++t) {
    const float h2 = __ldg(hg_all + (static_cast<long long>(t) * B + b) * ldhg + a);
#pragma unroll
    for (int p = 0; p < FIN_P; ++p)
      if (x[p] + h2 > 0.f) acc[p] += s_de[t * FIN_P + p];
  }
  const float wf = __ldg(w_f + a);
#pragma unroll
  for (int p = 0; p < FIN_P; ++p)
    if (p0 + p < P) d_att1[(static_cast<long long>(b) * P + p0 + p) * A + a] += wf * acc[p];
}

int attention_bwd_finish(const float* alphas, long long a_sb, long long a_st, const float* dawe_all,
                         const float* de_all, const float* att1, const float* hg_all, long long ldhg,
                         const float* w_f, float* d_att1, float* d_enc, int B, int T, int P, int A, int E,
                         cudaStream_t stream) {
  if (B <= 0 || T <= 0) return CCX_OK;
  if (P <= 0 || T * FIN_P * 4 > 200 * 1024) return CCX_ERR_SHAPE;
  const size_t smem = static_cast<size_t>(T) * FIN_P * sizeof(float);
  const unsigned pz = static_cast<unsigned>((P + FIN_P - 1) / FIN_P);
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(attention_bwd_finish_enc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem)) != cudaSuccess ||
        cudaFuncSetAttribute(attention_bwd_finish_att_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem)) != cudaSuccess)
      return CCX_ERR_CUDA;
  }
  ProfScope prof(PROF_ATTENTION, stream, (double)B * T * (E + A) * 4.0 + (double)B * P * (2.0 * E + 3.0 * A) * 4.0);
  if (d_enc != nullptr)
    attention_bwd_finish_enc_kernel<<<dim3(B, (E + 127) / 128, pz), 128, smem, stream>>>(alphas, a_sb, a_st, dawe_all,
                                                                                         d_enc, B, T, P, E);
  attention_bwd_finish_att_kernel<<<dim3(B, (A + 127) / 128, pz), 128, smem, stream>>>(att1, hg_all, ldhg, w_f, de_all,
                                                                                       d_att1, B, T, P, A);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

__global__ void __launch_bounds__(512)
attention_bwd_att_kernel(const float* __restrict__ att1, const float* __restrict__ hg, long long ldhg,
                         const float* __restrict__ w_f, const float* __restrict__ alpha, long long alpha_ld,
                         const float* __restrict__ d_alpha_ext, long long dalpha_ld, float* __restrict__ d_hg,
                         long long ld_dhg, float* __restrict__ d_att1, float* __restrict__ d_wf, OpDst dhg_op,
                         long long ld_op, int P, int A) {
  __shared__ float s_de[ATT_MAX_P_BWD];
  __shared__ float s_red[16];
  const int b = blockIdx.x;
  // de_p = alpha_p * (dalpha_p - sum_q alpha_q dalpha_q)
  float mine = 0.f, a_p = 0.f, da_p = 0.f;
  if (threadIdx.x < P) {
    a_p = alpha[b * alpha_ld + threadIdx.x];
    da_p = d_hg[b * ld_dhg + threadIdx.x] + (d_alpha_ext ? d_alpha_ext[b * dalpha_ld + threadIdx.x] : 0.f);
    mine = a_p * da_p;
  }
  mine = warp_sum(mine);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mine;
  __syncthreads();
  float dot = 0.f;
  for (int w = 0; w < (P + 31) / 32; ++w) dot += s_red[w];
  if (threadIdx.x < P) s_de[threadIdx.x] = a_p * (da_p - dot);
  __syncthreads();   // every read of the d_hg[b, 0:P] scratch is done before it is overwritten below
  for (int a = threadIdx.x; a < A; a += 512) {
    float datt2 = 0.f, dwf = 0.f;
    const float h2 = hg[b * ldhg + a], wf = __ldg(w_f + a);
    for (int p = 0; p < P; ++p) {
      const long long i1 = (static_cast<long long>(b) * P + p) * A + a;
      const float pre = att1[i1] + h2;
      if (pre > 0.f) {
        const float de = s_de[p];
        const float gpre = de * wf;
        datt2 += gpre;
        d_att1[i1] += gpre;
        dwf = fmaf(de, pre, dwf);
      }
    }
    d_hg[b * ld_dhg + a] = datt2;
    if (dhg_op.hi != nullptr) put(dhg_op, b * ld_op + a, datt2);
    atomicAdd(d_wf + a, dwf);
  }
}

int bahdanau_attention_bwd(const float* att1, const float* hg, long long ldhg, const float* w_f, const float* enc,
                           const float* alpha, long long alpha_ld, const float* d_out, long long ld_dout,
                           const float* d_alpha_ext, long long dalpha_ld, float* d_hg, long long ld_dhg,
                           float* d_att1, float* d_enc, float* d_wf, int bt, int P, int A, int E,
                           cudaStream_t stream, void* op_hi, float* op_lo, int op_dtype, long long ld_op,
                           int scratch_zeroed, float* dawe_out, float* dalpha_acc, float* de_out) {
  if (bt <= 0) return CCX_OK;
  if (P <= 0 || P > ATT_MAX_P_BWD || P > A || bt > 65535) return CCX_ERR_SHAPE;
  OpDst dhg_op{op_hi, op_lo, op_dtype};
  ProfScope prof(PROF_ATTENTION, stream, (double)bt * P * (2.0 * A + 3.0 * E) * 4.0);
  const bool deferred = dawe_out != nullptr && dalpha_acc != nullptr && de_out != nullptr;
  if (!deferred) dawe_out = dalpha_acc = de_out = nullptr;
  if (!deferred && !scratch_zeroed &&
      cudaMemset2DAsync(d_hg, ld_dhg * sizeof(float), 0, P * sizeof(float), bt, stream) != cudaSuccess)
    return CCX_ERR_CUDA;
  dim3 g1(bt, (E + 127) / 128);
  if (P <= ATTB_P) {
    if (launch_pdl(attention_bwd_enc_kernel<ATTB_P>, g1, dim3(128), 0, stream, hg, ldhg, enc, alpha, alpha_ld, d_out,
                   ld_dout, d_hg, ld_dhg, d_enc, dhg_op, ld_op, P, A, E, dawe_out, dalpha_acc) != cudaSuccess)
      return CCX_ERR_CUDA;
  }
  else
    attention_bwd_enc_kernel<ATT_MAX_P_BWD><<<g1, 128, 0, stream>>>(hg, ldhg, enc, alpha, alpha_ld, d_out, ld_dout,
                                                                    d_hg, ld_dhg, d_enc, dhg_op, ld_op, P, A, E,
                                                                    dawe_out, dalpha_acc);
  if (deferred) {
    if (launch_pdl(attention_bwd_att_split_kernel, dim3(bt, (A + 127) / 128), dim3(128), 0, stream, att1, hg, ldhg,
                   w_f, alpha, alpha_ld, d_alpha_ext, dalpha_ld, dalpha_acc, de_out, d_hg, ld_dhg, d_wf, dhg_op,
                   ld_op, P, A) != cudaSuccess)
      return CCX_ERR_CUDA;
  } else
    attention_bwd_att_kernel<<<bt, 512, 0, stream>>>(att1, hg, ldhg, w_f, alpha, alpha_ld, d_alpha_ext, dalpha_ld,
                                                     d_hg, ld_dhg, d_att1, d_wf, dhg_op, ld_op, P, A);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// out[b, p, :] += v[b, :] * scale   (backward of mean over pixels)
__global__ void __launch_bounds__(256)
bcast_add_rows_kernel(float* __restrict__ out, const float* __restrict__ v, float scale, int P, int E, int B) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * P * E) return;
  const int e = static_cast<int>(i % E);
  const int b = static_cast<int>(i / (static_cast<long long>(P) * E));
  out[i] += v[static_cast<long long>(b) * E + e] * scale;
}
int bcast_add_rows(float* out, const float* v, float scale, int B, int P, int E, cudaStream_t stream) {
  if (B <= 0) return CCX_OK;
  const long long n = static_cast<long long>(B) * P * E;
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)n * 8.0);
  bcast_add_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(out, v, scale, P, E, B);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// clamp(+-clip) + Adam, multi-tensor.  Table entry: {param, grad, exp_avg, exp_avg_sq, n}
// ---------------------------------------------------------------------------------------------
struct AdamEntry {
  float* p;
  float* g;
  float* m;
  float* v;
  long long n;
};

__global__ void __launch_bounds__(256)
adam_clamp_kernel(const AdamEntry* __restrict__ table, const int* __restrict__ block_entry,
                  const long long* __restrict__ block_offset, float lr, float beta1, float beta2, float eps,
                  float bc1, float bc2_sqrt, float clip, int chunk, const float* __restrict__ step_dev) {
  if (step_dev != nullptr) {       // step count kept on the device (CUDA-graph replay: no host value can be baked in)
    const float t = __ldg(step_dev);
    bc1 = 1.f - powf(beta1, t);
    bc2_sqrt = sqrtf(1.f - powf(beta2, t));
  }
  const AdamEntry e = table[block_entry[blockIdx.x]];
  const long long begin = block_offset[blockIdx.x];
  const long long end = min(e.n, begin + chunk);
  const float step = lr / bc1;
  auto update = [&](float& p, float& g, float& m, float& v) {
    if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);     // utils/utils.py:189-192
    m = beta1 * m + (1.f - beta1) * g;                    // torch/optim/adam.py _single_tensor_adam
    v = beta2 * v + (1.f - beta2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p -= step * (m / denom);
  };
  // 16-byte path: 7 x 128-bit transactions per 4 elements (the clamped gradient is written back only if it changed)
  const bool vec = (((reinterpret_cast<uintptr_t>(e.p) | reinterpret_cast<uintptr_t>(e.g) |
                      reinterpret_cast<uintptr_t>(e.m) | reinterpret_cast<uintptr_t>(e.v)) & 15) == 0) &&
                   ((begin & 3) == 0);
  long long scalar_from = begin;
  if (vec) {
    const long long q0 = begin >> 2, q1 = end >> 2;
    for (long long q = q0 + threadIdx.x; q < q1; q += 256) {
      float4 p4 = reinterpret_cast<float4*>(e.p)[q], g4 = reinterpret_cast<float4*>(e.g)[q];
      float4 m4 = reinterpret_cast<float4*>(e.m)[q], v4 = reinterpret_cast<float4*>(e.v)[q];
      const float4 g0 = g4;
      update(p4.x, g4.x, m4.x, v4.x);
      update(p4.y, g4.y, m4.y, v4.y);
      update(p4.z, g4.z, m4.z, v4.z);
      update(p4.w, g4.w, m4.w, v4.w);
      reinterpret_cast<float4*>(e.p)[q] = p4;
      reinterpret_cast<float4*>(e.m)[q] = m4;
      reinterpret_cast<float4*>(e.v)[q] = v4;
      if (g0.x != g4.x || g0.y != g4.y || g0.z != g4.z || g0.w != g4.w) reinterpret_cast<float4*>(e.g)[q] = g4;
    }
    scalar_from = q1 << 2;
  }
  for (long long i = scalar_from + threadIdx.x; i < end; i += 256) {
    float p = e.p[i], g = e.g[i], m = e.m[i], v = e.v[i];
    update(p, g, m, v);
    e.g[i] = g;
    e.m[i] = m;
    e.v[i] = v;
    e.p[i] = p;
  }
}

int adam_clamp(const void* table, const int* block_entry, const long long* block_offset, int n_blocks, float lr,
               float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float clip, int chunk,
               double total_params, cudaStream_t stream, const float* step_dev) {
  if (n_blocks <= 0) return CCX_OK;
  ProfScope prof(PROF_OPTIM, stream, total_params * 28.0);
  adam_clamp_kernel<<<n_blocks, 256, 0, stream>>>(reinterpret_cast<const AdamEntry*>(table), block_entry,
                                                  block_offset, lr, beta1, beta2, eps, bc1, bc2_sqrt, clip, chunk,
                                                  step_dev);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// Weight refresh: every kernel-side copy of the decoder's fp32 master weights in ONE launch.
// After an optimizer step the bf16 operands (plain, concatenated, row-permuted or transposed images of the
// parameters — DecoderWithAttention._prepare / TransformerDecoder._prepare) have to follow their masters.  Per operand
// that was a cat / index / transpose in ATen, a cast kernel and a device-to-device copy: ~45 (LSTM decoder) to ~110
// (Transformer) launches of a few microseconds each per train step.  Here a table of rectangular segments
//   dst[i, j] (or dst[j, i] when transposed) = src[row_map ? row_map[i] : i, j] (+ src2[...])
// is walked by one grid: 64 x 64 tiles, fp32 -> bf16 (or fp32 for the small bias vectors), both sides coalesced.
// ---------------------------------------------------------------------------------------------
struct CastSeg {
  const float* src;
  const float* src2;
  void* dst;
  const int* row_map;
  long long src_ld, dst_ld;
  int rows, cols;        // of the (row-mapped) source block
  int flags;             // bit 0: transpose, bit 1: fp32 destination
  int tile0;             // first tile of this segment in the grid
};
static_assert(sizeof(CastSeg) == 64, "CastSeg is mirrored by ctypes (imagecaptioningconvnext_b200/_lib.py)");

__global__ void __launch_bounds__(256)
cast_segments_kernel(const CastSeg* __restrict__ segs, int nseg) {
  __shared__ float tile[64][65];
  const int t = blockIdx.x;
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {                      // last segment with tile0 <= t
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].tile0 <= t) lo = mid; else hi = mid - 1;
  }
  const CastSeg s = segs[lo];
  const int tiles_c = (s.cols + 63) >> 6;
  const int lt = t - s.tile0;
  const int r0 = (lt / tiles_c) << 6, c0 = (lt % tiles_c) << 6;
  const bool f32 = (s.flags & 2) != 0;
  const bool vec_in = ((reinterpret_cast<uintptr_t>(s.src) & 15) == 0) && ((s.src_ld & 3) == 0) && ((s.cols & 3) == 0) &&
                      (s.src2 == nullptr || (reinterpret_cast<uintptr_t>(s.src2) & 15) == 0);
  if (vec_in) {
    const int q = threadIdx.x & 15, ty = threadIdx.x >> 4;          // 16 float4 per tile row, 16 rows per pass
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = ty + 16 * k, r = r0 + i, c = c0 + 4 * q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < s.rows && c < s.cols) {
        const long long sr = s.row_map ? s.row_map[r] : r;
        v = *reinterpret_cast<const float4*>(s.src + sr * s.src_ld + c);
        if (s.src2) {
          const float4 w = *reinterpret_cast<const float4*>(s.src2 + sr * s.src_ld + c);
          v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
      }
      tile[i][4 * q] = v.x; tile[i][4 * q + 1] = v.y; tile[i][4 * q + 2] = v.z; tile[i][4 * q + 3] = v.w;
    }
  } else {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const int i = ty + 4 * k, r = r0 + i, c = c0 + tx;
      float v = 0.f;
      if (r < s.rows && c < s.cols) {
        const long long sr = s.row_map ? s.row_map[r] : r;
        v = s.src[sr * s.src_ld + c];
        if (s.src2) v += s.src2[sr * s.src_ld + c];
      }
      tile[i][tx] = v;
    }
  }
  __syncthreads();
  // destination rows / columns of this tile: plain -> (r, c); transposed -> (c, r)
  const bool tr = (s.flags & 1) != 0;
  const int d_rows = tr ? s.cols : s.rows, d_cols = tr ? s.rows : s.cols;
  const int dr0 = tr ? c0 : r0, dc0 = tr ? r0 : c0;
  const bool pair_out = !f32 && ((reinterpret_cast<uintptr_t>(s.dst) & 3) == 0) && ((s.dst_ld & 1) == 0) && ((d_cols & 1) == 0);
  if (pair_out) {
    const int px = threadIdx.x & 31, ty = threadIdx.x >> 5;         // lane -> two consecutive destination columns
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = ty + 8 * k, dr = dr0 + i, dc = dc0 + 2 * px;
      if (dr < d_rows && dc < d_cols) {
        const float a = tr ? tile[2 * px][i] : tile[i][2 * px];
        const float b = tr ? tile[2 * px + 1][i] : tile[i][2 * px + 1];
        *reinterpret_cast<uint32_t*>(static_cast<__nv_bfloat16*>(s.dst) + dr * s.dst_ld + dc) = pack_bf16x2(a, b);
      }
    }
  } else {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const int i = ty + 4 * k, dr = dr0 + i, dc = dc0 + tx;
      if (dr < d_rows && dc < d_cols) {
        const float v = tr ? tile[tx][i] : tile[i][tx];
        if (f32) static_cast<float*>(s.dst)[dr * s.dst_ld + dc] = v;
        else static_cast<__nv_bfloat16*>(s.dst)[dr * s.dst_ld + dc] = __float2bfloat16_rn(v);
      }
    }
  }
}

int cast_segments(const void* segs_dev, int nseg, int total_tiles, double bytes, cudaStream_t stream) {
  if (nseg <= 0 || total_tiles <= 0) return CCX_OK;
  ProfScope prof(PROF_ELEMENTWISE, stream, bytes);
  cast_segments_kernel<<<total_tiles, 256, 0, stream>>>(reinterpret_cast<const CastSeg*>(segs_dev), nseg);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}


}  // namespace ccx
