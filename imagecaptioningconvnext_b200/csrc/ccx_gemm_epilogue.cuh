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
  int act;              // 0 none, 1 gelu(erf), 2 relu, 3 gelu'(x) (erf form)
  int out_dtype;        // CCX_F32 / CCX_BF16
  int split;            // 1: write tf32 hi to out, residual lo to out_lo
  int tma;              // 1: output (and residual) move through TMA boxes (epilogue_tile_tma); set by the launcher
  int res_mul;          // 1: the residual multiplies the activated result instead of being added
  int fast;             // 1: bf16 operands — act 3 differentiates the tanh-form GELU the bf16 forward evaluates
};

static constexpr int EPI_SCRATCH_FLOATS = 32 * 32;   // per epilogue warp

#ifdef CCX_GEMM_TIMELINE
static __device__ int g_gemm_dbg_mode;   // developer switch (tools/gemm_timeline.cu): 1 = tcgen05.ld only, 2 = hand-over only
static __device__ unsigned long long g_epi_tl[8];   // cycles of CTA 0 / first epilogue thread per epilogue phase
#define EPI_T0(t) const long long t = clock64()
#define EPI_TL(slot, t) st.tl[slot] += clock64() - t      /* registers; flushed once by the kernel */
#else
#define EPI_T0(t)
#define EPI_TL(slot, t)
#endif

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
    } else if (ep.act == 3) {
#pragma unroll
      for (int r = 0; r < 32; ++r) y[r] = ep.fast ? gelu_grad_tanh_fast(y[r]) : gelu_grad_erf(y[r]);
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
        for (int r = 0; r < 32; ++r) {
          const float rv = __bfloat162float(rp[r * ep.ldr]);
          y[r] = ep.res_mul ? y[r] * rv : y[r] + rv;
        }
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
        for (int r = 0; r < 32; ++r) y[r] = ep.res_mul ? y[r] * rv[r] : y[r] + rv[r];
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
      else if (ep.act == 3) y = ep.fast ? gelu_grad_tanh_fast(y) : gelu_grad_erf(y);
      if (ep.emask != nullptr) y *= __ldg(ep.emask + (long long)row * ep.ldm + col);
      if (scale) y *= cs * (ep.rowscale ? __ldg(ep.rowscale + row / ep.rows_per_group) : 1.0f);
      if (ep.out_dtype == CCX_BF16) {
        if (ep.residual != nullptr) {
          const float rv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ep.residual)[(long long)row * ep.ldr + col]);
          y = ep.res_mul ? y * rv : y + rv;
        }
        reinterpret_cast<__nv_bfloat16*>(ep.out)[(long long)row * ep.ldc + col] = __float2bfloat16_rn(y);
      } else {
        if (ep.residual != nullptr) {
          const float rv = __ldg(reinterpret_cast<const float*>(ep.residual) + (long long)row * ep.ldr + col);
          y = ep.res_mul ? y * rv : y + rv;
        }
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
    } else if (ep.act == 3) {
#pragma unroll
      for (int j = 0; j < 64; ++j) f[j] = ep.fast ? gelu_grad_tanh_fast(f[j]) : gelu_grad_erf(f[j]);
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

// Drain one accumulator stage in units of 64 columns (u = half, half+CG, ...; CG = epilogue warps per TMEM lane
// quarter): the bf16 fast path above (FAST64: needs ~150 registers), or two
// generic 32-column chunks.  tmem_addr = TMEM address of (this warp's lane quarter, first column of the stage);
// row0 = first row of this warp's 32-row slab.
template <int BN, int CG = 2, bool FAST64 = true>
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
#ifdef CCX_GEMM_TIMELINE
  const int dbg_mode = g_gemm_dbg_mode;
  if (dbg_mode == 2) return;
#endif
#pragma unroll 1
  for (int u = half; u < UNITS; u += CG) {
    const int n0 = n_blk * BN + u * 64;
    if (n0 >= N) break;  // warp-uniform
#ifdef CCX_GEMM_TIMELINE
    if (dbg_mode == 1) {
      uint32_t v0[32], v1[32];
      tmem_ld32(tmem_addr + u * 64, v0);
      tmem_ld32(tmem_addr + u * 64 + 32, v1);
      tmem_ld_wait();
      uint32_t x = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) x ^= v0[j] ^ v1[j];
      if (x == 0x12345u) reinterpret_cast<uint32_t*>(ep.out)[0] = x;
      continue;
    }
#endif
    if (FAST64 && bf16_fast && n0 + 64 <= N && BN >= 64) {
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

// ----------------------------------------------------------------------------
// TMA epilogue (1-CTA kernel, outputs whose rows are 16-byte aligned, no dropout mask, no hi/lo split).
//
// Measured with tools/gemm_timeline.cu on B200: the transposing epilogue above costs 9.7 k cycles per 128x256 bf16
// GELU tile and 16.3 k per fp32 residual tile — 2.2x / 3.7x the tile's MMA time at K = 512 — and all of it is latency
// (bias LDG -> FADD, transpose LDS -> STG, 64-byte residual loads) that 2 warps per scheduler cannot hide.  Here the
// accumulator stays in tcgen05.ld's own layout (a thread = a row): a "unit" is 32 rows x 128 bytes (64 bf16 / 32 fp32
// columns), written with 8 conflict-free STS.128 into a 128-byte-swizzled 4 KB box and sent with ONE
// cp.async.bulk.tensor store (rows / columns past M / N are clipped by the tensor map, so ragged tiles need no code);
// the residual arrives the same way (TMA load into the box the result is then written over), prefetched one unit
// ahead; bias / layer-scale of the tile sit in shared memory (loaded while waiting for the accumulator).
// ----------------------------------------------------------------------------
static constexpr int EPI_UNIT_BYTES = 4096;

// One 32-column chunk of a unit: f = accumulator (row = this lane, columns c0 .. c0+31 of the tile).  bf16 output: the
// chunk is one half (QW = 4 sixteen-byte words starting at word q0 = 0 / 4) of the unit's 128-byte row; fp32 output:
// the whole row (QW = 8).  pbias / pscale point at the chunk's first column in the shared parameter block.
template <bool F32OUT>
__device__ __forceinline__ void epilogue_chunk_tma(const EpiArgs& ep, float (&f)[32], uint8_t* rowp, int q0, int sw,
                                                   const float* pbias, const float* pscale, float rs,
                                                   const uint8_t* resrow) {
  // resrow: this lane's 128-byte row of the residual box (same swizzle), or nullptr; may be rowp itself (in place)
  const bool has_res = resrow != nullptr;
  if (ep.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(pbias + j);     // warp-wide broadcast
      f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
    }
  }
  const bool scaled = ep.colscale != nullptr || ep.rowscale != nullptr;
  if constexpr (!F32OUT) {
#ifdef CCX_GEMM_TIMELINE
    if (g_gemm_dbg_mode == 3 && ep.act == 1) {          // GELU replaced by a MUFU-free stand-in of the same shape
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 pk;
        pk.x = pack_bf16x2(f[8 * q] * f[8 * q], f[8 * q + 1] * f[8 * q + 1]);
        pk.y = pack_bf16x2(f[8 * q + 2] * f[8 * q + 2], f[8 * q + 3] * f[8 * q + 3]);
        pk.z = pack_bf16x2(f[8 * q + 4] * f[8 * q + 4], f[8 * q + 5] * f[8 * q + 5]);
        pk.w = pack_bf16x2(f[8 * q + 6] * f[8 * q + 6], f[8 * q + 7] * f[8 * q + 7]);
        *reinterpret_cast<uint4*>(rowp + (((q0 + q) ^ sw) << 4)) = pk;
      }
      return;
    }
#endif
    if (ep.act == 1 && !scaled && !has_res) {
      // Linear + GELU (every CNBlock's first Linear): packed-fp16 GELU, two columns per instruction
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 pk;
        pk.x = gelu_tanh_f16x2_to_bf16x2(f[8 * q], f[8 * q + 1]);
        pk.y = gelu_tanh_f16x2_to_bf16x2(f[8 * q + 2], f[8 * q + 3]);
        pk.z = gelu_tanh_f16x2_to_bf16x2(f[8 * q + 4], f[8 * q + 5]);
        pk.w = gelu_tanh_f16x2_to_bf16x2(f[8 * q + 6], f[8 * q + 7]);
        *reinterpret_cast<uint4*>(rowp + (((q0 + q) ^ sw) << 4)) = pk;
      }
      return;
    }
  }
  if (ep.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = F32OUT ? gelu_erf(f[j]) : gelu_tanh_fast(f[j]);
  } else if (ep.act == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
  } else if (ep.act == 3) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = ep.fast ? gelu_grad_tanh_fast(f[j]) : gelu_grad_erf(f[j]);
  }
  if (ep.colscale != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 c = *reinterpret_cast<const float4*>(pscale + j);
      f[j] *= c.x * rs; f[j + 1] *= c.y * rs; f[j + 2] *= c.z * rs; f[j + 3] *= c.w * rs;
    }
  } else if (ep.rowscale != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= rs;
  }
  if constexpr (F32OUT) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4* cp = reinterpret_cast<float4*>(rowp + ((q ^ sw) << 4));
      float4 y = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
      if (has_res) {
        const float4 r = *reinterpret_cast<const float4*>(resrow + ((q ^ sw) << 4));
        if (ep.res_mul) { y.x *= r.x; y.y *= r.y; y.z *= r.z; y.w *= r.w; }
        else { y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w; }
      }
      *cp = y;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4* cp = reinterpret_cast<uint4*>(rowp + (((q0 + q) ^ sw) << 4));
      if (has_res) {
        const uint4 r = *reinterpret_cast<const uint4*>(resrow + (((q0 + q) ^ sw) << 4));
        float2 t;
        if (ep.res_mul) {
          t = unpack_bf16x2(r.x); f[8 * q] *= t.x;     f[8 * q + 1] *= t.y;
          t = unpack_bf16x2(r.y); f[8 * q + 2] *= t.x; f[8 * q + 3] *= t.y;
          t = unpack_bf16x2(r.z); f[8 * q + 4] *= t.x; f[8 * q + 5] *= t.y;
          t = unpack_bf16x2(r.w); f[8 * q + 6] *= t.x; f[8 * q + 7] *= t.y;
        } else {
          t = unpack_bf16x2(r.x); f[8 * q] += t.x;     f[8 * q + 1] += t.y;
          t = unpack_bf16x2(r.y); f[8 * q + 2] += t.x; f[8 * q + 3] += t.y;
          t = unpack_bf16x2(r.z); f[8 * q + 4] += t.x; f[8 * q + 5] += t.y;
          t = unpack_bf16x2(r.w); f[8 * q + 6] += t.x; f[8 * q + 7] += t.y;
        }
      }
      uint4 pk;
      pk.x = pack_bf16x2(f[8 * q], f[8 * q + 1]);
      pk.y = pack_bf16x2(f[8 * q + 2], f[8 * q + 3]);
      pk.z = pack_bf16x2(f[8 * q + 4], f[8 * q + 5]);
      pk.w = pack_bf16x2(f[8 * q + 6], f[8 * q + 7]);
      *cp = pk;
    }
  }
}

// The residual of a CTA's LAST tile arrives through the operand ring (its B stages are idle by then): the producer
// thread queues the tile's residual boxes behind its last operand loads, so all of them land while the final MMAs
// retire instead of one box per warp at a time after the accumulator is complete (a single-tile CTA — every stage-3
// Linear(4C -> C) at batch 32 — otherwise pays the residual's memory latency four times in a row, exposed).
// Box b = unit * 4 + row quarter sits in extra ring slot b / boxes_per_stage at offset (b % boxes_per_stage) * 4 KB.
struct EpiRing {
  const uint8_t* base;      // first B stage
  uint64_t* full_bar;       // the ring's "stage filled" barriers
  int stage0;               // ring position after this CTA's last operand load
  uint32_t phase0;
  int stages, stage_bytes, boxes_per_stage;
};

// Per-warp state of the TMA epilogue across tiles.
struct EpiTmaState {
  uint32_t rphase = 0;   // bit b = parity of this warp's residual mbarrier of box b
  int box = 0;           // next staging box (0 .. BOXES-1)
  float pb = 0.0f, pc = 1.0f, rs = 1.0f;   // this thread's share of the NEXT tile's parameters (loaded a tile ahead)
  bool res_ahead = false;                  // the first residual box of the next tile is already on its way
#ifdef CCX_GEMM_TIMELINE
  long long tl[5] = {0, 0, 0, 0, 0};
#endif
};

// This thread's parameters of tile (m_blk-row row, n_blk): bias / layer-scale of column n_blk*BN + tid_e, the
// stochastic-depth scale of its row.  Issued one tile ahead so the L2 latency never sits in front of the math.
template <int BN>
__device__ __forceinline__ void epilogue_params_prefetch(const EpiArgs& ep, EpiTmaState& st, int n_blk, int row, int tid_e,
                                                         int M, int N) {
  st.pb = 0.0f; st.pc = 1.0f; st.rs = 1.0f;
  if (tid_e < BN) {
    const int col = n_blk * BN + tid_e;
    if (col < N) {
      if (ep.bias != nullptr) st.pb = __ldg(ep.bias + col);
      if (ep.colscale != nullptr) st.pc = __ldg(ep.colscale + col);
    }
  }
  if (ep.rowscale != nullptr && row < M) st.rs = __ldg(ep.rowscale + row / ep.rows_per_group);
}

// Drain one accumulator stage.  CG = epilogue warps per TMEM lane quarter (this warp: column group cg, units
// cg, cg+CG, ...); tid_e = thread index among the 128*CG epilogue threads; boxes = this warp's BOXES x 4 KB staging
// (1024-byte aligned); params = [2][BN] floats shared by the epilogue warps (bias, layer-scale of this tile);
// rbar = this warp's residual mbarriers (one per box).  ALL epilogue warps must call this for every tile (two named
// barriers).  st.pb/pc/rs hold this tile's parameters on entry (epilogue_params_prefetch); (next_n_blk, next_row0)
// name the warp's next tile (next_n_blk < 0: none): its parameters and its first residual box are requested while
// this tile is processed.
//
// The accumulator is read 32 columns at a time and the tcgen05.ld of chunk k+1 is issued before the math of chunk k,
// so TMEM latency, the box hand-over and the TMA store issue all sit under the previous chunk's arithmetic.
template <int BN, int BOXES, bool F32OUT, int CG>
__device__ __forceinline__ void epilogue_tile_tma(const EpiArgs& ep, const CUtensorMap* tmC, const CUtensorMap* tmR,
                                                  uint32_t tmem_addr, int cg, int row0, int lane, int tid_e,
                                                  uint8_t* boxes, float* params, uint64_t* rbar, EpiTmaState& st,
                                                  int n_blk, int M, int N, uint64_t* tfull_bar, uint32_t acc_phase,
                                                  int next_n_blk, int next_row0, const EpiRing* ring = nullptr,
                                                  bool next_in_ring = false) {
  // ring != nullptr: this is the CTA's last tile and its residual comes through the operand ring;
  // next_in_ring: the NEXT tile is that one, so no residual box is requested for it here
  constexpr int UCOLS = F32OUT ? 32 : 64;
  constexpr int CPU = UCOLS / 32;                     // chunks per unit
  constexpr int UNITS = (BN + UCOLS - 1) / UCOLS;
  bool active = row0 < M;                             // warp-uniform: this warp's 32 rows exist
  bool has_res = ep.residual != nullptr;
#ifdef CCX_GEMM_TIMELINE
  if (g_gemm_dbg_mode == 2) { active = false; has_res = false; }
#endif
  auto res_fetch = [&](int nb, int r0, int u, int box) {   // lane 0: the box must be free (its last store read out)
    mbar_expect_tx(rbar + box, EPI_UNIT_BYTES);
    tma_load_2d(boxes + box * EPI_UNIT_BYTES, tmR, rbar + box, nb * BN + u * UCOLS, r0);
  };
  const bool any = active && cg < UNITS && n_blk * BN + cg * UCOLS < N;
  if (has_res && any && !st.res_ahead && ring == nullptr && lane == 0) {   // first residual box, ahead of the accumulator
    tma_store_wait_read<BOXES - 1>();
    res_fetch(n_blk, row0, cg, st.box);
  }
  st.res_ahead = false;
  mbar_wait(tfull_bar, acc_phase);
  tc_fence_after();
  EPI_T0(t_par);
  const float rs = st.rs;
  named_bar_sync(1, 128 * CG);                        // every warp is done with the previous tile's parameters
  if (tid_e < BN) {
    params[tid_e] = st.pb;
    params[BN + tid_e] = st.pc;
  }
  named_bar_sync(1, 128 * CG);
  if (next_n_blk >= 0) epilogue_params_prefetch<BN>(ep, st, next_n_blk, next_row0 + lane, tid_e, M, N);
  EPI_TL(0, t_par);
  // the next tile's first residual box is requested as soon as a box is free (after this tile's last store)
  const bool next_any = has_res && !next_in_ring && next_n_blk >= 0 && next_row0 < M && cg < UNITS &&
                        next_n_blk * BN + cg * UCOLS < N;
  if (!any) {
    if (next_any && lane == 0) {
      tma_store_wait_read<BOXES - 1>();
      res_fetch(next_n_blk, next_row0, cg, st.box);
    }
    st.res_ahead = next_any;
    return;
  }
  const int sw = lane & 7;
  uint32_t v[32];
  tmem_ld32(tmem_addr + cg * UCOLS, v);               // first chunk
#pragma unroll 1
  for (int u = cg; u < UNITS; u += CG) {
    const int n0 = n_blk * BN + u * UCOLS;
    if (n0 >= N) break;                               // warp-uniform
    uint8_t* box = boxes + st.box * EPI_UNIT_BYTES;
    const int next_box = (st.box + 1 == BOXES) ? 0 : st.box + 1;
    const bool more = (u + CG < UNITS) && (n0 + CG * UCOLS < N);
    EPI_T0(t_w);
    const uint8_t* resrow = nullptr;
    if (has_res && ring != nullptr) {
      const int b = u * 4 + ((row0 >> 5) & 3);
      const int slot = ring->stage0 + b / ring->boxes_per_stage;
      const int stg = slot % ring->stages;
      mbar_wait(ring->full_bar + stg, ring->phase0 ^ ((slot / ring->stages) & 1u));
      resrow = ring->base + stg * ring->stage_bytes + (b % ring->boxes_per_stage) * EPI_UNIT_BYTES + lane * 128;
      if (lane == 0) tma_store_wait_read<BOXES - 1>();
      __syncwarp();
    } else if (has_res) {
      resrow = box + lane * 128;
      if (BOXES > 1 && lane == 0) {                   // the next unit's residual into the other box
        if (more) {
          tma_store_wait_read<0>();
          res_fetch(n_blk, row0, u + CG, next_box);
        } else if (next_any) {
          tma_store_wait_read<0>();
          res_fetch(next_n_blk, next_row0, cg, next_box);
        }
      }
      mbar_wait(rbar + st.box, (st.rphase >> st.box) & 1u);
      st.rphase ^= 1u << st.box;
    } else {
      if (lane == 0) tma_store_wait_read<BOXES - 1>();
      __syncwarp();
    }
    EPI_TL(1, t_w);
    EPI_T0(t_c);
    uint8_t* rowp = box + lane * 128;
#pragma unroll
    for (int c = 0; c < CPU; ++c) {
      float f[32];
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      // next chunk (of this unit, or the first of the warp's next unit) before this one's arithmetic
      if (c + 1 < CPU) tmem_ld32(tmem_addr + u * UCOLS + (c + 1) * 32, v);
      else if (more) tmem_ld32(tmem_addr + (u + CG) * UCOLS, v);
      epilogue_chunk_tma<F32OUT>(ep, f, rowp, c * 4, sw, params + u * UCOLS + c * 32, params + BN + u * UCOLS + c * 32,
                                 rs, resrow);
    }
    EPI_TL(2, t_c);
    EPI_T0(t_f);
    fence_proxy_async_smem();
    __syncwarp();
    EPI_TL(3, t_f);
    EPI_T0(t_s);
    if (lane == 0) {
      tma_store_2d(tmC, box, n0, row0);
      tma_store_commit();
      if (has_res && BOXES == 1 && ((more && ring == nullptr) || (!more && next_any))) {
        tma_store_wait_read<0>();                     // single box: the next residual can only follow this store
        if (more) res_fetch(n_blk, row0, u + CG, 0);
        else res_fetch(next_n_blk, next_row0, cg, 0);
      }
    }
    EPI_TL(4, t_s);
    st.box = next_box;
  }
  st.res_ahead = next_any;
}

}  // namespace ccx
