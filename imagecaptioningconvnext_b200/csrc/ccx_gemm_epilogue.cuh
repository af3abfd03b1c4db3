// ccx_gemm_epilogue.cuh — the fused GEMM epilogue for one 32-column chunk of one accumulator row, shared by the
// 1-CTA and the 2-CTA (cta_group::2) tcgen05 kernels.  f[] holds the fp32 accumulators of columns n0..n0+31 of `row`.
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

__device__ __forceinline__ void epilogue_chunk(const EpiArgs& ep, float (&f)[32], int row, bool row_ok, int n0, int N,
                                               float rs, const float4* res_pre = nullptr) {
  const bool full = (n0 + 32 <= N);
  if (ep.bias != nullptr) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
        f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) f[j] += __ldg(ep.bias + n0 + j);
    }
  }
  if (ep.act == 1) {
    if (ep.out_dtype == CCX_BF16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = gelu_tanh_fast(f[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
    }
  } else if (ep.act == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
  }
  if (ep.emask != nullptr && row_ok) {
    const float* mrow = ep.emask + (long long)row * ep.ldm + n0;
    if (full && ((reinterpret_cast<uintptr_t>(mrow) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 mk = __ldg(reinterpret_cast<const float4*>(mrow + j));
        f[j] *= mk.x; f[j + 1] *= mk.y; f[j + 2] *= mk.z; f[j + 3] *= mk.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) f[j] *= __ldg(mrow + j);
    }
  }
  if (ep.colscale != nullptr) {
    if (full && ((reinterpret_cast<uintptr_t>(ep.colscale + n0) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 cs = __ldg(reinterpret_cast<const float4*>(ep.colscale + n0 + j));
        f[j] *= cs.x * rs; f[j + 1] *= cs.y * rs; f[j + 2] *= cs.z * rs; f[j + 3] *= cs.w * rs;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) f[j] *= __ldg(ep.colscale + n0 + j) * rs;
    }
  } else if (ep.rowscale != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= rs;
  }
  if (row_ok) {
  if (ep.out_dtype == CCX_BF16) {
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)row * ep.ldc + n0;
    const __nv_bfloat16* rrow =
        ep.residual ? reinterpret_cast<const __nv_bfloat16*>(ep.residual) + (long long)row * ep.ldr + n0
                    : nullptr;
    const bool vec = full && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0) &&
                     (rrow == nullptr || (reinterpret_cast<uintptr_t>(rrow) & 15) == 0);
    if (vec) {
      if (rrow) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const uint4 r = __ldg(reinterpret_cast<const uint4*>(rrow + j));
          float2 t;
          t = unpack_bf16x2(r.x); f[j] += t.x; f[j + 1] += t.y;
          t = unpack_bf16x2(r.y); f[j + 2] += t.x; f[j + 3] += t.y;
          t = unpack_bf16x2(r.z); f[j + 4] += t.x; f[j + 5] += t.y;
          t = unpack_bf16x2(r.w); f[j + 6] += t.x; f[j + 7] += t.y;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 o;
        o.x = pack_bf16x2(f[j], f[j + 1]);
        o.y = pack_bf16x2(f[j + 2], f[j + 3]);
        o.z = pack_bf16x2(f[j + 4], f[j + 5]);
        o.w = pack_bf16x2(f[j + 6], f[j + 7]);
        *reinterpret_cast<uint4*>(orow + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (n0 + j < N) {
          float y = f[j];
          if (rrow) y += __bfloat162float(rrow[j]);
          orow[j] = __float2bfloat16_rn(y);
        }
      }
    }
  } else {
    float* orow = reinterpret_cast<float*>(ep.out) + (long long)row * ep.ldc + n0;
    float* lrow = ep.split ? ep.out_lo + (long long)row * ep.ldc + n0 : nullptr;
    const float* rrow =
        ep.residual ? reinterpret_cast<const float*>(ep.residual) + (long long)row * ep.ldr + n0 : nullptr;
    const bool vec = full && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0) &&
                     (rrow == nullptr || (reinterpret_cast<uintptr_t>(rrow) & 15) == 0);
    if (vec) {
      if (rrow) {
        if (res_pre != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 r = res_pre[j >> 2];
            f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(rrow + j));
            f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
          }
        }
      }
      if (ep.split) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 h, l;
          h.x = tf32_hi(f[j]);     l.x = f[j] - h.x;
          h.y = tf32_hi(f[j + 1]); l.y = f[j + 1] - h.y;
          h.z = tf32_hi(f[j + 2]); l.z = f[j + 2] - h.z;
          h.w = tf32_hi(f[j + 3]); l.w = f[j + 3] - h.w;
          *reinterpret_cast<float4*>(orow + j) = h;
          *reinterpret_cast<float4*>(lrow + j) = l;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(orow + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (n0 + j < N) {
          float y = f[j];
          if (rrow) y += __ldg(rrow + j);
          if (ep.split) {
            const float h = tf32_hi(y);
            orow[j] = h;
            lrow[j] = y - h;
          } else {
            orow[j] = y;
          }
        }
      }
    }
  }
  }  // row_ok
}

// Residual rows of an fp32 output tile are prefetched one 32-column chunk ahead (the first chunk BEFORE the wait on
// the accumulator), so the global-load latency hides behind the main loop / the previous chunk instead of stalling
// every chunk (ncu: the residual FADDs were the top stall of the Linear(4C->C)+residual GEMMs).
__device__ __forceinline__ bool residual_prefetch(const EpiArgs& ep, int row, bool row_ok, int n0, int N,
                                                  float4 (&r)[8]) {
  if (ep.residual == nullptr || ep.out_dtype != CCX_F32 || !row_ok || n0 + 32 > N) return false;
  const float* rrow = reinterpret_cast<const float*>(ep.residual) + (long long)row * ep.ldr + n0;
  const float* orow = reinterpret_cast<const float*>(ep.out) + (long long)row * ep.ldc + n0;
  if ((reinterpret_cast<uintptr_t>(rrow) & 15) || (reinterpret_cast<uintptr_t>(orow) & 15)) return false;
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = __ldg(reinterpret_cast<const float4*>(rrow) + j);
  return true;
}

// Drain one accumulator stage: chunks c = half, half+2, ... of 32 columns each.  tmem_addr = TMEM address of
// (this warp's lane quarter, first column of the stage).  PREFETCH_RES is a compile-time switch so that GEMMs
// without an fp32 residual keep the lean loop (fewer live registers).
template <int BN, bool PREFETCH_RES>
__device__ __forceinline__ void epilogue_tile_impl(const EpiArgs& ep, uint32_t tmem_addr, int half, int row,
                                                   bool row_ok, int n_blk, int N, uint64_t* tfull_bar,
                                                   uint32_t acc_phase) {
  float rs = 1.0f;
  if (ep.rowscale != nullptr && row_ok) rs = __ldg(ep.rowscale + row / ep.rows_per_group);
  float4 rnext[8];
  bool have_next = false;
  if constexpr (PREFETCH_RES) have_next = residual_prefetch(ep, row, row_ok, n_blk * BN + half * 32, N, rnext);
  mbar_wait(tfull_bar, acc_phase);
  tc_fence_after();
#pragma unroll 1
  for (int c = half; c < BN / 32; c += 2) {
    const int n0 = n_blk * BN + c * 32;
    if (n0 >= N) break;  // warp-uniform
    uint32_t v[32];
    tmem_ld32(tmem_addr + c * 32, v);
    if constexpr (PREFETCH_RES) {
      float4 rcur[8];
      const bool have_cur = have_next;
#pragma unroll
      for (int j = 0; j < 8; ++j) rcur[j] = rnext[j];
      have_next = (c + 2 < BN / 32) && residual_prefetch(ep, row, row_ok, n0 + 64, N, rnext);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      epilogue_chunk(ep, f, row, row_ok, n0, N, rs, have_cur ? rcur : nullptr);
    } else {
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      epilogue_chunk(ep, f, row, row_ok, n0, N, rs, nullptr);
    }
    __syncwarp();
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_tile(const EpiArgs& ep, uint32_t tmem_addr, int half, int row, bool row_ok,
                                              int n_blk, int N, uint64_t* tfull_bar, uint32_t acc_phase) {
  if (ep.residual != nullptr && ep.out_dtype == CCX_F32)   // uniform over the kernel
    epilogue_tile_impl<BN, true>(ep, tmem_addr, half, row, row_ok, n_blk, N, tfull_bar, acc_phase);
  else
    epilogue_tile_impl<BN, false>(ep, tmem_addr, half, row, row_ok, n_blk, N, tfull_bar, acc_phase);
}

}  // namespace ccx
