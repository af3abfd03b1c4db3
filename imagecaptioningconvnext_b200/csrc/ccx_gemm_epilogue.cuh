// ccx_gemm_epilogue.cuh — the fused GEMM epilogue shared by the 1-CTA and the 2-CTA (cta_group::2) tcgen05 kernels.
//
// tcgen05.ld hands every thread ONE accumulator row (32 consecutive columns).  Writing / reading global memory in
// that layout makes each warp instruction touch 32 different lines (16 B out of each), which made the epilogue —
// not the MMA — the bottleneck of every encoder GEMM (ncu r01: LSU-transaction bound, residual loads the top stall).
// So each 32x32 chunk is transposed through a warp-private, XOR-swizzled shared-memory scratch: afterwards a lane owns
// a COLUMN, and every global access of the epilogue (residual, dropout mask, output hi/lo) is one coalesced line per
// warp instruction; bias / layer-scale become per-lane scalars.
#pragma once
#include "ccx_common.cuh"

namespace ccx {

struct EpiArgs {
  void* out;            // [M, ldc] bf16 or fp32
  float* out_lo;        // fp32 split output (lo part) or nullptr
  const float* bias;    // [N] or nullptr
  const float* colscale;  // [N] or nullptr   (layer_scale)
  const float* rowscale;  // [M / rows_per_group] or nullptr (stochastic-depth noise/(1-p))
  const void* residual;   // [M, ldr] same dtype as out, or nullptr
  const float* emask;     // [M, ldm] element-wise multiplier applied after the activation (dropout), or nullptr
  long long ldc, ldr, ldm;
  int rows_per_group;
  int act;              // 0 none, 1 gelu(erf), 2 relu
  int out_dtype;        // CCX_F32 / CCX_BF16
  int split;            // 1: write tf32 hi to out, residual lo to out_lo
};

static constexpr int EPI_SCRATCH_FLOATS = 32 * 32;   // per epilogue warp

// f[j] = accumulator (row0 + lane, n0 + j).  scratch: this warp's 32x32 floats.
__device__ __forceinline__ void epilogue_chunk(const EpiArgs& ep, const float (&f)[32], float* scratch, int row0,
                                               int lane, int n0, int M, int N) {
  // element (row i, col j) lives at i*32 + ((j>>2) ^ (i&7))*4 + (j&3): 8 conflict-free STS.128 per thread on the way
  // in (16-byte groups swizzled by the row), conflict-free scalar LDS on the way out (a lane = a column)
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(scratch + lane * 32 + ((q ^ (lane & 7)) << 2)) =
        make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
  __syncwarp();
  const int cq = lane >> 2, cr = lane & 3;   // this lane's column = 4*cq + cr
  const int col = n0 + lane;
  const bool col_ok = col < N;
  const int rows = min(32, M - row0);          // warp-uniform
  const float b = (ep.bias != nullptr && col_ok) ? __ldg(ep.bias + col) : 0.0f;
  const float cs = (ep.colscale != nullptr && col_ok) ? __ldg(ep.colscale + col) : 1.0f;
  const bool scale = (ep.colscale != nullptr) || (ep.rowscale != nullptr);
  if (rows == 32 && col_ok) {
    // full chunk: all global loads of the 32 rows are issued before they are consumed
    float y[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) y[r] = scratch[r * 32 + ((cq ^ (r & 7)) << 2) + cr] + b;
    if (ep.act == 1) {
      if (ep.out_dtype == CCX_BF16) {
#pragma unroll
        for (int r = 0; r < 32; ++r) y[r] = gelu_tanh_fast(y[r]);
      } else {
#pragma unroll
        for (int r = 0; r < 32; ++r) y[r] = gelu_erf(y[r]);
      }
    } else if (ep.act == 2) {
#pragma unroll
      for (int r = 0; r < 32; ++r) y[r] = fmaxf(y[r], 0.0f);
    }
    if (ep.emask != nullptr) {
      const float* mp = ep.emask + (long long)row0 * ep.ldm + col;
#pragma unroll
      for (int r = 0; r < 32; ++r) y[r] *= __ldg(mp + r * ep.ldm);
    }
    if (scale) {
      if (ep.rowscale != nullptr) {
#pragma unroll
        for (int r = 0; r < 32; ++r) y[r] *= cs * __ldg(ep.rowscale + (row0 + r) / ep.rows_per_group);
      } else {
#pragma unroll
        for (int r = 0; r < 32; ++r) y[r] *= cs;
      }
    }
    if (ep.out_dtype == CCX_BF16) {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)row0 * ep.ldc + col;
      if (ep.residual != nullptr) {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(ep.residual) + (long long)row0 * ep.ldr + col;
#pragma unroll
        for (int r = 0; r < 32; ++r) y[r] += __bfloat162float(rp[r * ep.ldr]);
      }
#pragma unroll
      for (int r = 0; r < 32; ++r) op[r * ep.ldc] = __float2bfloat16_rn(y[r]);
    } else {
      float* op = reinterpret_cast<float*>(ep.out) + (long long)row0 * ep.ldc + col;
      if (ep.residual != nullptr) {
        const float* rp = reinterpret_cast<const float*>(ep.residual) + (long long)row0 * ep.ldr + col;
        float rv[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) rv[r] = __ldg(rp + r * ep.ldr);
#pragma unroll
        for (int r = 0; r < 32; ++r) y[r] += rv[r];
      }
      if (ep.split) {
        float* lp = ep.out_lo + (long long)row0 * ep.ldc + col;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const float h = tf32_hi(y[r]);
          op[r * ep.ldc] = h;
          lp[r * ep.ldc] = y[r] - h;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 32; ++r) op[r * ep.ldc] = y[r];
      }
    }
  } else if (col_ok) {
    // ragged chunk (last rows of M): same math, row by row
    for (int r = 0; r < rows; ++r) {
      const int row = row0 + r;
      float y = scratch[r * 32 + ((cq ^ (r & 7)) << 2) + cr] + b;
      if (ep.act == 1) y = (ep.out_dtype == CCX_BF16) ? gelu_tanh_fast(y) : gelu_erf(y);
      else if (ep.act == 2) y = fmaxf(y, 0.0f);
      if (ep.emask != nullptr) y *= __ldg(ep.emask + (long long)row * ep.ldm + col);
      if (scale) y *= cs * (ep.rowscale ? __ldg(ep.rowscale + row / ep.rows_per_group) : 1.0f);
      if (ep.out_dtype == CCX_BF16) {
        if (ep.residual != nullptr)
          y += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ep.residual)[(long long)row * ep.ldr + col]);
        reinterpret_cast<__nv_bfloat16*>(ep.out)[(long long)row * ep.ldc + col] = __float2bfloat16_rn(y);
      } else {
        if (ep.residual != nullptr) y += __ldg(reinterpret_cast<const float*>(ep.residual) + (long long)row * ep.ldr + col);
        if (ep.split) {
          const float h = tf32_hi(y);
          reinterpret_cast<float*>(ep.out)[(long long)row * ep.ldc + col] = h;
          ep.out_lo[(long long)row * ep.ldc + col] = y - h;
        } else {
          reinterpret_cast<float*>(ep.out)[(long long)row * ep.ldc + col] = y;
        }
      }
    }
  }
  __syncwarp();   // the scratch is reused by the next chunk
}

// bf16 output without residual / dropout mask (the Linear+GELU of every CNBlock, the decoders' operand outputs):
// 64 columns at a time.  Bias / activation / scales are applied in the accumulator's row-per-thread layout, the
// result is packed to bf16x2 and transposed through the same 4 KB scratch as 32 words per row, so every store
// instruction writes one full 128-byte line of a row (the 32-column path writes 64 B per instruction).
__device__ __forceinline__ void epilogue_unit_bf16(const EpiArgs& ep, float (&f)[64], float* scratch, int row0,
                                                   int lane, int n0, int M) {
  const int row = row0 + lane;
  float rs = 1.0f;
  if (ep.rowscale != nullptr && row < M) rs = __ldg(ep.rowscale + row / ep.rows_per_group);
  if (ep.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 64; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
      f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
    }
  }
  uint32_t* sw = reinterpret_cast<uint32_t*>(scratch);
  if (ep.act == 1 && ep.colscale == nullptr && ep.rowscale == nullptr) {
    // Linear + GELU (CNBlock): packed-fp16 GELU, two columns per instruction
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4 pk;
      pk.x = gelu_tanh_f16x2_to_bf16x2(f[8 * q], f[8 * q + 1]);
      pk.y = gelu_tanh_f16x2_to_bf16x2(f[8 * q + 2], f[8 * q + 3]);
      pk.z = gelu_tanh_f16x2_to_bf16x2(f[8 * q + 4], f[8 * q + 5]);
      pk.w = gelu_tanh_f16x2_to_bf16x2(f[8 * q + 6], f[8 * q + 7]);
      *reinterpret_cast<uint4*>(sw + lane * 32 + ((q ^ (lane & 7)) << 2)) = pk;
    }
  } else {
    if (ep.act == 1) {
#pragma unroll
      for (int j = 0; j < 64; ++j) f[j] = gelu_tanh_fast(f[j]);
    } else if (ep.act == 2) {
#pragma unroll
      for (int j = 0; j < 64; ++j) f[j] = fmaxf(f[j], 0.0f);
    }
    if (ep.colscale != nullptr) {
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(ep.colscale + n0 + j));
        f[j] *= c.x * rs; f[j + 1] *= c.y * rs; f[j + 2] *= c.z * rs; f[j + 3] *= c.w * rs;
      }
    } else if (ep.rowscale != nullptr) {
#pragma unroll
      for (int j = 0; j < 64; ++j) f[j] *= rs;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4 pk;
      pk.x = pack_bf16x2(f[8 * q], f[8 * q + 1]);
      pk.y = pack_bf16x2(f[8 * q + 2], f[8 * q + 3]);
      pk.z = pack_bf16x2(f[8 * q + 4], f[8 * q + 5]);
      pk.w = pack_bf16x2(f[8 * q + 6], f[8 * q + 7]);
      *reinterpret_cast<uint4*>(sw + lane * 32 + ((q ^ (lane & 7)) << 2)) = pk;
    }
  }
  __syncwarp();
  const int rows = min(32, M - row0);
  const int wq = lane >> 2, wr = lane & 3;     // this lane's word = columns n0 + 2*lane, +1
  uint32_t* op = reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)row0 * ep.ldc + n0) + lane;
  const long long ldw = ep.ldc >> 1;           // row stride in 32-bit words (ldc is even on this path)
  if (rows == 32) {
#pragma unroll
    for (int r = 0; r < 32; ++r) op[r * ldw] = sw[r * 32 + ((wq ^ (r & 7)) << 2) + wr];
  } else {
    for (int r = 0; r < rows; ++r) op[r * ldw] = sw[r * 32 + ((wq ^ (r & 7)) << 2) + wr];
  }
  __syncwarp();
}

// Drain one accumulator stage in units of 64 columns (u = half, half+2, ...): the bf16 fast path above, or two
// generic 32-column chunks.  tmem_addr = TMEM address of (this warp's lane quarter, first column of the stage);
// row0 = first row of this warp's 32-row slab.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const EpiArgs& ep, uint32_t tmem_addr, int half, int row0, int lane,
                                              float* scratch, int n_blk, int M, int N, uint64_t* tfull_bar,
                                              uint32_t acc_phase) {
  mbar_wait(tfull_bar, acc_phase);
  tc_fence_after();
  const bool bf16_fast = (ep.out_dtype == CCX_BF16) && ep.residual == nullptr && ep.emask == nullptr &&
                         ((ep.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(ep.out) & 3) == 0) &&
                         (ep.bias == nullptr || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0) &&
                         (ep.colscale == nullptr || (reinterpret_cast<uintptr_t>(ep.colscale) & 15) == 0);
  constexpr int UNITS = (BN + 63) / 64;
#pragma unroll 1
  for (int u = half; u < UNITS; u += 2) {
    const int n0 = n_blk * BN + u * 64;
    if (n0 >= N) break;  // warp-uniform
    if (bf16_fast && n0 + 64 <= N && BN >= 64) {
      uint32_t v0[32], v1[32];
      tmem_ld32(tmem_addr + u * 64, v0);
      tmem_ld32(tmem_addr + u * 64 + 32, v1);
      tmem_ld_wait();
      float f[64];
#pragma unroll
      for (int j = 0; j < 32; ++j) { f[j] = __uint_as_float(v0[j]); f[32 + j] = __uint_as_float(v1[j]); }
      if (row0 < M) epilogue_unit_bf16(ep, f, scratch, row0, lane, n0, M);
    } else {
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int n1 = n0 + c * 32;
        if (n1 >= N || u * 64 + c * 32 >= BN) break;
        uint32_t v[32];
        tmem_ld32(tmem_addr + u * 64 + c * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (row0 < M) epilogue_chunk(ep, f, scratch, row0, lane, n1, M, N);
      }
    }
  }
}

}  // namespace ccx
