// dwconv_ln.cu — channels-last depthwise 7x7 convolution fused with the LayerNorm that follows it.
//
// Replaces, per CNBlock, torchvision/models/convnext.py:52-54
//     Conv2d(C, C, kernel_size=7, padding=3, groups=C, bias=True) -> Permute -> LayerNorm(C, eps=1e-6)
// (called from the reference through models/encoder.py:24).
//
// Layout: x is the fp32 NHWC residual stream [B, H, W, C]; the output is the row-major [B*H*W, C]
// A-operand of the following pointwise GEMM, written either as bf16 or as a (tf32-hi, fp32-lo) pair.
//
// Work decomposition
//   * one CTA = an 8x8 pixel tile x 128 channels; the (8+6)x(8+6)x128 fp32 halo tile is brought in by ONE
//     4-D TMA bulk-tensor copy (out-of-bounds = zero fill gives the conv padding for free),
//   * a thread-block cluster of C/128 CTAs covers all channels of the same pixel tile, and the LayerNorm
//     statistics (sum, sum of squares per pixel) are exchanged through distributed shared memory,
//   * inside a CTA: warp = (pixel half p, channel group g): a 4x8 output patch for 32 channels; every lane
//     owns ONE channel, keeps its 49 filter taps in registers and 32 fp32 accumulators, so each halo value is
//     read from shared memory exactly once per thread (140 conflict-free LDS.32 for 1568 FFMA).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "ccx_common.cuh"
#include "ccx_gemm.h"
#include "ccx_prof.h"

namespace cg = cooperative_groups;

namespace ccx {

static constexpr int DW_TILE = 8;
static constexpr int DW_HALO = DW_TILE + 6;  // 14
static constexpr int DW_CH = 128;            // channels per CTA
static constexpr int DW_THREADS = 256;
static constexpr int DW_TILE_BYTES = DW_HALO * DW_HALO * DW_CH * 4;  // 100,352
static constexpr int DW_SMEM = DW_TILE_BYTES + 6144 /*stats*/;
static constexpr int DW_SMEM64 = DW_TILE_BYTES / 2 + 6144;      // 64-channel tiles: 56,320 B -> 4 CTAs per SM

struct DwArgs {
  const float* w;      // [49][C]  tap-major depthwise filter
  const float* bias;   // [C]
  const float* gamma;  // [C]
  const float* beta;   // [C]
  void* out;           // [B*H*W, C] bf16 or fp32 (hi)
  float* out_lo;       // fp32 lo part or nullptr
  const float* addend; // plain mode only: out = conv + addend (residual-gradient accumulation), or nullptr
  int B, H, W, C;
  int tiles_w, tiles_h;
  float eps;
  int out_dtype;       // CCX_F32 / CCX_BF16
};

// sum the 32 per-lane arrays v[0..31] across the warp; lane L ends up with the total of element L in v[0]
__device__ __forceinline__ void warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv7_ln_kernel(const __grid_constant__ CUtensorMap tmX, DwArgs a) {
  // declared aligned (not re-aligned through integer arithmetic) so the compiler keeps the shared address space
  // and emits LDS for the tile reads instead of generic LD
  extern __shared__ __align__(1024) float dw_smem[];
  float* tile = dw_smem;                                             // [14][14][128]
  float* part = dw_smem + DW_TILE_BYTES / 4;                         // [2 stat][2 p][4 g][32 px] = 512 f
  float* clpart = part + 512;                                        // [8 rank][2 stat][64 px]   = 1024 f (max)
  __shared__ uint64_t bar;
  __shared__ float s_mean[64], s_rstd[64];

  cg::cluster_group cluster = cg::this_cluster();
  const int nc = a.C / DW_CH;                 // cluster size (gridDim.x)
  const int crank = blockIdx.x;               // cluster spans gridDim.x exactly
  const int b = blockIdx.z;
  const int th = blockIdx.y / a.tiles_w, tw = blockIdx.y - th * a.tiles_w;
  const int h0 = th * DW_TILE, w0 = tw * DW_TILE;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = warp >> 2;   // pixel half: output rows p*4 .. p*4+3 of the tile
  const int g = warp & 3;    // channel group of 32 inside the CTA's 128
  const int ch_local = g * 32 + lane;
  const int ch = crank * DW_CH + ch_local;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, DW_TILE_BYTES);
    tma_load_4d(tile, &tmX, &bar, crank * DW_CH, w0 - 3, h0 - 3, b);
  }

  // filter taps and affine parameters while the tile is in flight
  float wt[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) wt[t] = __ldg(a.w + t * a.C + ch);
  const bool plain = (a.gamma == nullptr);   // no LayerNorm: raw conv output (LN recompute / data-gradient use)
  const float bias = a.bias ? __ldg(a.bias + ch) : 0.f;
  const float gam = plain ? 1.f : __ldg(a.gamma + ch);
  const float bet = plain ? 0.f : __ldg(a.beta + ch);

  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = bias;

  mbar_wait(&bar, 0);

  const float* tp = tile + (p * 4) * DW_HALO * DW_CH + ch_local;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#pragma unroll
    for (int c = 0; c < DW_HALO; ++c) {
      const float v = tp[(r * DW_HALO + c) * DW_CH];
#pragma unroll
      for (int oh = 0; oh < 4; ++oh) {
        const int kr = r - oh;
        if (kr < 0 || kr > 6) continue;
#pragma unroll
        for (int ow = 0; ow < 8; ++ow) {
          const int kc = c - ow;
          if (kc < 0 || kc > 6) continue;
          acc[oh * 8 + ow] = fmaf(v, wt[kr * 7 + kc], acc[oh * 8 + ow]);
        }
      }
    }
  }

  // ---- LayerNorm statistics: per pixel over all C channels (this warp: 32 of them) ----
  if (!plain) {
    float s1[32], s2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { s1[i] = acc[i]; s2[i] = acc[i] * acc[i]; }
    warp_transpose_reduce32(s1, lane);
    warp_transpose_reduce32(s2, lane);
    part[((0 * 2 + p) * 4 + g) * 32 + lane] = s1[0];
    part[((1 * 2 + p) * 4 + g) * 32 + lane] = s2[0];
  }
  if (plain) {   // uniform over the whole cluster: nobody reaches the cluster barrier below
    if (threadIdx.x < 64) { s_mean[threadIdx.x] = 0.f; s_rstd[threadIdx.x] = 1.f; }
    __syncthreads();
  } else {
  __syncthreads();
  if (threadIdx.x < 128) {
    // thread -> (stat, p, px): sum the 4 channel-group partials, publish to every CTA of the cluster
    const int stat = threadIdx.x >> 6, pp = (threadIdx.x >> 5) & 1, px = threadIdx.x & 31;
    const float* src = part + ((stat * 2 + pp) * 4) * 32 + px;
    const float tot = src[0] + src[32] + src[64] + src[96];
    for (int r = 0; r < nc; ++r) {
      float* dst = cluster.map_shared_rank(clpart, r);
      dst[(crank * 2 + stat) * 64 + pp * 32 + px] = tot;
    }
  }
  cluster.sync();
  if (threadIdx.x < 64) {
    float s = 0.f, q = 0.f;
    for (int r = 0; r < nc; ++r) {
      s += clpart[(r * 2 + 0) * 64 + threadIdx.x];
      q += clpart[(r * 2 + 1) * 64 + threadIdx.x];
    }
    const float inv = 1.0f / static_cast<float>(a.C);
    const float mean = s * inv;
    const float var = fmaxf(q * inv - mean * mean, 0.0f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + a.eps);
  }
  __syncthreads();
  }

  // ---- normalise + write ----
  const long long base = ((static_cast<long long>(b) * a.H + h0 + p * 4) * a.W + w0) * a.C + ch;  // pixel (p*4, 0)
  const int row_stride = a.W * a.C;   // < 2^31 elements per image row is guaranteed by the launcher
  const bool interior = (h0 + DW_TILE <= a.H) && (w0 + DW_TILE <= a.W);
  if (interior && a.out_dtype == CCX_BF16 && !plain) {
    // fast path (every tile of the 256x256 configurations): no per-pixel bounds checks, bf16 out
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + base;
#pragma unroll
    for (int oh = 0; oh < 4; ++oh) {
#pragma unroll
      for (int ow = 0; ow < 8; ++ow) {
        const int px = p * 32 + oh * 8 + ow;
        const float y = (acc[oh * 8 + ow] - s_mean[px]) * s_rstd[px] * gam + bet;
        o[oh * row_stride + ow * a.C] = __float2bfloat16_rn(y);
      }
    }
    return;
  }
#pragma unroll
  for (int oh = 0; oh < 4; ++oh) {
    const int h = h0 + p * 4 + oh;
#pragma unroll
    for (int ow = 0; ow < 8; ++ow) {
      const int w = w0 + ow;
      const int px = p * 32 + oh * 8 + ow;
      float y = (acc[oh * 8 + ow] - s_mean[px]) * s_rstd[px] * gam + bet;
      if (h < a.H && w < a.W) {
        const long long idx = base + oh * row_stride + ow * a.C;
        if (plain && a.addend != nullptr) y += a.addend[idx];
        if (a.out_dtype == CCX_BF16) {
          reinterpret_cast<__nv_bfloat16*>(a.out)[idx] = __float2bfloat16_rn(y);
        } else if (a.out_lo != nullptr) {
          const float hi = tf32_hi(y);
          reinterpret_cast<float*>(a.out)[idx] = hi;
          a.out_lo[idx] = y - hi;
        } else {
          reinterpret_cast<float*>(a.out)[idx] = y;
        }
      }
    }
  }
  // no trailing cluster barrier: every remote store into a CTA's clpart happened before the barrier above
}

// ---------------------------------------------------------------------------------------------------------
// v2: packed fp32x2 math.  Same tile / TMA / cluster scheme, but a CTA is 4 warps and every lane owns TWO adjacent
// channels: halo values are read with LDS.64 and accumulated with Blackwell's FFMA2 (fma.rn.f32x2), which halves
// the issue slots per MAC — v1 is issue-bound (~1.7 instructions per FFMA), not FMA-pipe or HBM bound.
//   warp = (p, g): p = pixel half (4x8 output patch), g = 64-channel group; 32 float2 accumulators, 49 float2 taps.
// ---------------------------------------------------------------------------------------------------------
// CH = channels per CTA (64 or 128): 64 halves the halo tile (50 KB) so FOUR CTAs are resident per SM and the
// TMA-wait / FMA / LayerNorm-exchange phases of different tiles overlap four ways instead of two.

template <int CH>
__global__ void __launch_bounds__(CH, 256 / CH)
dwconv7_ln_kernel_v2(const __grid_constant__ CUtensorMap tmX, DwArgs a) {
  constexpr int NG = CH / 64;                 // 64-channel groups (warps per pixel half)
  constexpr int TILE_BYTES = DW_HALO * DW_HALO * CH * 4;
  extern __shared__ __align__(1024) float dw_smem[];
  float* tile = dw_smem;                                             // [14][14][128]
  float* part = dw_smem + TILE_BYTES / 4;                            // [2 stat][2 p][NG g][32 px]
  float* clpart = part + 512;                                        // [<=16 rank][2 stat][64 px]
  __shared__ uint64_t bar;
  __shared__ float s_mean[64], s_rstd[64];

  cg::cluster_group cluster = cg::this_cluster();
  const int nc = a.C / CH;
  const int crank = blockIdx.x;
  const int b = blockIdx.z;
  const int th = blockIdx.y / a.tiles_w, tw = blockIdx.y - th * a.tiles_w;
  const int h0 = th * DW_TILE, w0 = tw * DW_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = warp / NG;   // pixel half
  const int g = warp % NG;   // 64-channel group
  const int ch_local = g * 64 + lane * 2;
  const int ch = crank * CH + ch_local;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  grid_dep_sync();          // PDL: everything above overlaps the previous kernel's tail
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, TILE_BYTES);
    tma_load_4d(tile, &tmX, &bar, crank * CH, w0 - 3, h0 - 3, b);
  }
  float2 wt[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) wt[t] = __ldg(reinterpret_cast<const float2*>(a.w + t * a.C + ch));
  const bool plain = (a.gamma == nullptr);
  const float2 bias = a.bias ? __ldg(reinterpret_cast<const float2*>(a.bias + ch)) : make_float2(0.f, 0.f);
  const float2 gam = plain ? make_float2(1.f, 1.f) : __ldg(reinterpret_cast<const float2*>(a.gamma + ch));
  const float2 bet = plain ? make_float2(0.f, 0.f) : __ldg(reinterpret_cast<const float2*>(a.beta + ch));
  float2 acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = bias;

  mbar_wait(&bar, 0);
  const float* tp = tile + (p * 4) * DW_HALO * CH + ch_local;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#pragma unroll
    for (int c = 0; c < DW_HALO; ++c) {
      const float2 v = *reinterpret_cast<const float2*>(tp + (r * DW_HALO + c) * CH);
#pragma unroll
      for (int oh = 0; oh < 4; ++oh) {
        const int kr = r - oh;
        if (kr < 0 || kr > 6) continue;
#pragma unroll
        for (int ow = 0; ow < 8; ++ow) {
          const int kc = c - ow;
          if (kc < 0 || kc > 6) continue;
          acc[oh * 8 + ow] = ffma2(v, wt[kr * 7 + kc], acc[oh * 8 + ow]);
        }
      }
    }
  }
  if (!plain) {
    float s1[32], s2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      s1[i] = acc[i].x + acc[i].y;
      s2[i] = fmaf(acc[i].x, acc[i].x, acc[i].y * acc[i].y);
    }
    warp_transpose_reduce32(s1, lane);
    warp_transpose_reduce32(s2, lane);
    part[((0 * 2 + p) * NG + g) * 32 + lane] = s1[0];
    part[((1 * 2 + p) * NG + g) * 32 + lane] = s2[0];
  }
  if (plain) {
    if (threadIdx.x < 64) { s_mean[threadIdx.x] = 0.f; s_rstd[threadIdx.x] = 1.f; }
    __syncthreads();
  } else {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 128; idx += CH) {
      // idx -> (stat, p, px): sum the channel-group partials, publish to every CTA of the cluster
      const int stat = idx >> 6, pp = (idx >> 5) & 1, px = idx & 31;
      const float* src = part + ((stat * 2 + pp) * NG) * 32 + px;
      float tot = src[0];
      if (NG == 2) tot += src[32];
      for (int r = 0; r < nc; ++r) {
        float* dst = cluster.map_shared_rank(clpart, r);
        dst[(crank * 2 + stat) * 64 + pp * 32 + px] = tot;
      }
    }
    cluster.sync();
    if (threadIdx.x < 64) {
      float s = 0.f, q = 0.f;
      for (int r = 0; r < nc; ++r) {
        s += clpart[(r * 2 + 0) * 64 + threadIdx.x];
        q += clpart[(r * 2 + 1) * 64 + threadIdx.x];
      }
      const float inv = 1.0f / static_cast<float>(a.C);
      const float mean = s * inv;
      const float var = fmaxf(q * inv - mean * mean, 0.0f);
      s_mean[threadIdx.x] = mean;
      s_rstd[threadIdx.x] = rsqrtf(var + a.eps);
    }
    __syncthreads();
  }
  const long long base = ((static_cast<long long>(b) * a.H + h0 + p * 4) * a.W + w0) * a.C + ch;
  const int row_stride = a.W * a.C;
  const bool interior = (h0 + DW_TILE <= a.H) && (w0 + DW_TILE <= a.W);
  if (interior && a.out_dtype == CCX_BF16 && !plain) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + base;
#pragma unroll
    for (int oh = 0; oh < 4; ++oh) {
#pragma unroll
      for (int ow = 0; ow < 8; ++ow) {
        const int px = p * 32 + oh * 8 + ow;
        const float m = s_mean[px], rs = s_rstd[px];
        const float y0 = (acc[oh * 8 + ow].x - m) * rs * gam.x + bet.x;
        const float y1 = (acc[oh * 8 + ow].y - m) * rs * gam.y + bet.y;
        *reinterpret_cast<uint32_t*>(o + oh * row_stride + ow * a.C) = pack_bf16x2(y0, y1);
      }
    }
    return;
  }
#pragma unroll
  for (int oh = 0; oh < 4; ++oh) {
    const int h = h0 + p * 4 + oh;
#pragma unroll
    for (int ow = 0; ow < 8; ++ow) {
      const int w = w0 + ow;
      const int px = p * 32 + oh * 8 + ow;
      const float m = s_mean[px], rs = s_rstd[px];
      float y0 = (acc[oh * 8 + ow].x - m) * rs * gam.x + bet.x;
      float y1 = (acc[oh * 8 + ow].y - m) * rs * gam.y + bet.y;
      if (h < a.H && w < a.W) {
        const long long idx = base + oh * row_stride + ow * a.C;
        if (plain && a.addend != nullptr) {
          const float2 ad = *reinterpret_cast<const float2*>(a.addend + idx);
          y0 += ad.x; y1 += ad.y;
        }
        if (a.out_dtype == CCX_BF16) {
          *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.out) + idx) = pack_bf16x2(y0, y1);
        } else if (a.out_lo != nullptr) {
          const float h0v = tf32_hi(y0), h1v = tf32_hi(y1);
          *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.out) + idx) = make_float2(h0v, h1v);
          *reinterpret_cast<float2*>(a.out_lo + idx) = make_float2(y0 - h0v, y1 - h1v);
        } else {
          *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.out) + idx) = make_float2(y0, y1);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// v3: persistent.  The one-tile kernels above are a chain of latencies per CTA (launch, 49 tap loads, the halo TMA,
// ~1.8 k cycles of FFMA2, the LayerNorm exchange with a cluster barrier, the stores) of which only the arithmetic is
// throughput; with 4 co-resident CTAs that all start together the phases line up instead of overlapping (ncu r02:
// 10.9 % warps active, 0.17 of the HBM rate).  Here a CTA owns 128 channels (4 warps, every lane two adjacent
// channels as in v2), TWO CTAs share an SM (2 warps per scheduler: one warp alone cannot hide the LDS -> FFMA2
// latency — a one-CTA-per-SM double-buffered version measured 30 % slower than v2) and a CLUSTER of C/128 CTAs walks
// over the 8x8 pixel tiles: taps / affine parameters are loaded once, the halo of the next tile is requested as soon
// as the last warp has read the current one (it streams in under the LayerNorm exchange, the stores and the other
// CTA's arithmetic), and the per-pixel LayerNorm sums travel between the CTAs as st.async stores that complete on
// the receiver's mbarrier (double-buffered by tile parity) — no cluster barrier after start-up.
//   parity argument: a CTA sends its sums of tile i+1 only after it has read everybody's sums of tile i, so when a
//   peer's tile i+2 sums arrive (same buffer as tile i) they cannot overtake a reader.
// ---------------------------------------------------------------------------------------------------------
static constexpr int DW3_CH = 128;
static constexpr int DW3_THREADS = 128;
static constexpr int DW3_TILE_BYTES = DW_HALO * DW_HALO * DW3_CH * 4;       // 100,352
static constexpr int DW3_MAX_NC = 4;                                          // C <= 512
static constexpr int DW3_SMEM = DW3_TILE_BYTES + 1024 /*part*/ + 2 * DW3_MAX_NC * 512 /*clpart*/ + 1024;   // x2 per SM

__device__ __forceinline__ uint32_t dw_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void dw_st_async(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}

__global__ void __launch_bounds__(DW3_THREADS, 2)
dwconv7_ln_kernel_v3(const __grid_constant__ CUtensorMap tmX, DwArgs a, int nc, int n_clusters, int num_tiles) {
  extern __shared__ __align__(1024) float dw_smem[];
  float* tile = dw_smem;                                             // [14][14][128]
  float* part = dw_smem + DW3_TILE_BYTES / 4;                        // [2 stat][2 p][2 g][32 px]
  float* clpart = part + 256;                                        // [2 buf][nc rank][2 stat][64 px]
  __shared__ uint64_t full_bar, stat_bar[2];
  __shared__ float s_mean[64], s_rstd[64];

  const int crank = blockIdx.x % nc;
  const int cluster_id = blockIdx.x / nc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = warp >> 1;   // pixel half: output rows p*4 .. p*4+3
  const int g = warp & 1;    // 64-channel group
  const int ch_local = g * 64 + lane * 2;
  const int ch = crank * DW3_CH + ch_local;
  const int tiles_per_img = a.tiles_w * a.tiles_h;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    mbar_init(&full_bar, 1);
    mbar_init(&stat_bar[0], 1);
    mbar_init(&stat_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (nc > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // peers may signal us after this
  grid_dep_sync();          // PDL: everything above overlaps the previous kernel's tail

  auto issue = [&](int t) {     // thread 0; every warp is done reading the tile buffer
    const int b = t / tiles_per_img, r = t - b * tiles_per_img;
    const int th = r / a.tiles_w, tw = r - th * a.tiles_w;
    mbar_expect_tx(&full_bar, DW3_TILE_BYTES);
    tma_load_4d(tile, &tmX, &full_bar, crank * DW3_CH, tw * DW_TILE - 3, th * DW_TILE - 3, b);
  };
  if (threadIdx.x == 0 && cluster_id < num_tiles) issue(cluster_id);

  // taps / affine parameters: once per kernel
  float2 wt[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) wt[t] = __ldg(reinterpret_cast<const float2*>(a.w + t * a.C + ch));
  const float2 bias = a.bias ? __ldg(reinterpret_cast<const float2*>(a.bias + ch)) : make_float2(0.f, 0.f);
  const float2 gam = __ldg(reinterpret_cast<const float2*>(a.gamma + ch));
  const float2 bet = __ldg(reinterpret_cast<const float2*>(a.beta + ch));
  if (nc > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");

  int it = 0;
  for (int t = cluster_id; t < num_tiles; t += n_clusters, ++it) {
    const int buf = it & 1;
    const uint32_t par = (it >> 1) & 1;
    if (threadIdx.x == 0 && nc > 1) mbar_expect_tx(&stat_bar[buf], nc * 512);
    const int b = t / tiles_per_img, r = t - b * tiles_per_img;
    const int th = r / a.tiles_w, tw = r - th * a.tiles_w;
    const int h0 = th * DW_TILE, w0 = tw * DW_TILE;

    float2 acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = bias;
    mbar_wait(&full_bar, it & 1);
    const float* tp = tile + (p * 4) * DW_HALO * DW3_CH + ch_local;
#pragma unroll
    for (int rr = 0; rr < 10; ++rr) {
#pragma unroll
      for (int c = 0; c < DW_HALO; ++c) {
        const float2 v = *reinterpret_cast<const float2*>(tp + (rr * DW_HALO + c) * DW3_CH);
#pragma unroll
        for (int oh = 0; oh < 4; ++oh) {
          const int kr = rr - oh;
          if (kr < 0 || kr > 6) continue;
#pragma unroll
          for (int ow = 0; ow < 8; ++ow) {
            const int kc = c - ow;
            if (kc < 0 || kc > 6) continue;
            acc[oh * 8 + ow] = ffma2(v, wt[kr * 7 + kc], acc[oh * 8 + ow]);
          }
        }
      }
    }
    {
      float s1[32], s2[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s1[i] = acc[i].x + acc[i].y;
        s2[i] = fmaf(acc[i].x, acc[i].x, acc[i].y * acc[i].y);
      }
      warp_transpose_reduce32(s1, lane);
      warp_transpose_reduce32(s2, lane);
      part[((0 * 2 + p) * 2 + g) * 32 + lane] = s1[0];
      part[((1 * 2 + p) * 2 + g) * 32 + lane] = s2[0];
    }
    __syncthreads();                                   // (also: every warp is done reading the tile)
    if (threadIdx.x == 0 && t + n_clusters < num_tiles) issue(t + n_clusters);
    {
      // thread -> (stat, pixel half, px): sum the two channel groups, publish to every CTA of the cluster
      const int stat = threadIdx.x >> 6, pp = (threadIdx.x >> 5) & 1, px = threadIdx.x & 31;
      const float* src = part + ((stat * 2 + pp) * 2) * 32 + px;
      const float tot = src[0] + src[32];
      float* dst = clpart + ((buf * nc + crank) * 2 + stat) * 64 + pp * 32 + px;
      if (nc == 1) {
        *dst = tot;
      } else {
        const uint32_t daddr = smem_u32(dst), baddr = smem_u32(&stat_bar[buf]);
        for (int rk = 0; rk < nc; ++rk) dw_st_async(dw_mapa(daddr, rk), tot, dw_mapa(baddr, rk));
      }
    }
    if (nc == 1) __syncthreads();
    else mbar_wait(&stat_bar[buf], par);
    if (threadIdx.x < 64) {
      float sm = 0.f, q = 0.f;
      for (int rk = 0; rk < nc; ++rk) {
        sm += clpart[((buf * nc + rk) * 2 + 0) * 64 + threadIdx.x];
        q += clpart[((buf * nc + rk) * 2 + 1) * 64 + threadIdx.x];
      }
      const float inv = 1.0f / static_cast<float>(a.C);
      const float mean = sm * inv;
      const float var = fmaxf(q * inv - mean * mean, 0.0f);
      s_mean[threadIdx.x] = mean;
      s_rstd[threadIdx.x] = rsqrtf(var + a.eps);
    }
    __syncthreads();
    const long long base = ((static_cast<long long>(b) * a.H + h0 + p * 4) * a.W + w0) * a.C + ch;
    const int row_stride = a.W * a.C;
    const bool interior = (h0 + DW_TILE <= a.H) && (w0 + DW_TILE <= a.W);
    if (interior && a.out_dtype == CCX_BF16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + base;
#pragma unroll
      for (int oh = 0; oh < 4; ++oh) {
#pragma unroll
        for (int ow = 0; ow < 8; ++ow) {
          const int px = p * 32 + oh * 8 + ow;
          const float m = s_mean[px], rs = s_rstd[px];
          const float y0 = (acc[oh * 8 + ow].x - m) * rs * gam.x + bet.x;
          const float y1 = (acc[oh * 8 + ow].y - m) * rs * gam.y + bet.y;
          *reinterpret_cast<uint32_t*>(o + oh * row_stride + ow * a.C) = pack_bf16x2(y0, y1);
        }
      }
    } else {
#pragma unroll
      for (int oh = 0; oh < 4; ++oh) {
        const int h = h0 + p * 4 + oh;
#pragma unroll
        for (int ow = 0; ow < 8; ++ow) {
          const int w = w0 + ow;
          const int px = p * 32 + oh * 8 + ow;
          const float m = s_mean[px], rs = s_rstd[px];
          const float y0 = (acc[oh * 8 + ow].x - m) * rs * gam.x + bet.x;
          const float y1 = (acc[oh * 8 + ow].y - m) * rs * gam.y + bet.y;
          if (h < a.H && w < a.W) {
            const long long idx = base + oh * row_stride + ow * a.C;
            if (a.out_dtype == CCX_BF16) {
              *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.out) + idx) = pack_bf16x2(y0, y1);
            } else if (a.out_lo != nullptr) {
              const float h0v = tf32_hi(y0), h1v = tf32_hi(y1);
              *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.out) + idx) = make_float2(h0v, h1v);
              *reinterpret_cast<float2*>(a.out_lo + idx) = make_float2(y0 - h0v, y1 - h1v);
            } else {
              *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.out) + idx) = make_float2(y0, y1);
            }
          }
        }
      }
    }
    // s_mean / part of this tile are overwritten only after the next tile's first __syncthreads
  }
  // nobody may leave while a peer can still store into this CTA's shared memory or wait for this CTA's sums:
  // every CTA of the cluster runs the same number of tiles and has received all sums of its last tile, and its own
  // st.async stores were consumed by peers that are still alive; one closing cluster barrier keeps the exit ordered
  if (nc > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

int dwconv7_ln(const float* x, const float* w49c, const float* bias, const float* gamma, const float* beta,
               void* out, float* out_lo, int B, int H, int W, int C, float eps, int out_dtype,
               cudaStream_t stream, const float* addend) {
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535) return CCX_ERR_SHAPE;
  if (C % DW_CH != 0 || C / DW_CH > 8 || C < DW_CH) return CCX_ERR_SHAPE;
  static const bool use_v1 = (getenv("CCX_DWCONV_V1") != nullptr);   // A/B switch for profiling
  // 64 channels per CTA (4 resident CTAs per SM) whenever the LayerNorm cluster stays within the portable size 8
  const int CH = (!use_v1 && C / 64 <= 8 && getenv("CCX_DWCONV_CH128") == nullptr) ? 64 : DW_CH;
  if (out_dtype != CCX_F32 && out_dtype != CCX_BF16) return CCX_ERR_DTYPE;
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return CCX_ERR_TMA;
  if (reinterpret_cast<uintptr_t>(x) & 15) return CCX_ERR_SHAPE;
  if (static_cast<long long>(W) * C * 8 >= 0x7fffffffLL) return CCX_ERR_SHAPE;

  // persistent double-buffered kernel: LayerNorm mode, C = 128 .. 512 (clusters of 1 .. 4 CTAs, one CTA per SM)
  static const bool no_v3 = (getenv("CCX_DWCONV_V2") != nullptr);
  // (measured at batch 32, tools/dwconv_bench.py: v3 48.0 / 26.7 / 17.3 us against v2 54.0 / 29.8 / 16.8 us for
  //  C = 128 / 256 / 512 — all of them ~2x the FP32-pipe floor of 49 MACs per output, which is what bounds this op:
  //  22.8 / 11.4 / 5.7 us, above the HBM floor of 15.5 / 7.7 / 3.9 us; v3 where it wins)
  static const bool force_v3 = (getenv("CCX_DWCONV_V3") != nullptr);
  const bool use_v3 = !use_v1 && !no_v3 && gamma != nullptr && beta != nullptr && addend == nullptr &&
                      C % DW3_CH == 0 && C / DW3_CH <= (force_v3 ? DW3_MAX_NC : 2);
  const int CHB = use_v3 ? DW3_CH : CH;
  CUtensorMap tm;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)CHB, DW_HALO, DW_HALO, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return CCX_ERR_TMA;

  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.ref();
  if (!configured) {
    if (cudaFuncSetAttribute(dwconv7_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM) !=
            cudaSuccess ||
        cudaFuncSetAttribute(dwconv7_ln_kernel_v2<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM) !=
            cudaSuccess ||
        cudaFuncSetAttribute(dwconv7_ln_kernel_v2<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM64) !=
            cudaSuccess ||
        cudaFuncSetAttribute(dwconv7_ln_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, DW3_SMEM) != cudaSuccess)
      return CCX_ERR_CUDA;
    configured = true;
  }
  DwArgs a;
  a.w = w49c; a.bias = bias; a.gamma = gamma; a.beta = beta;
  a.out = out; a.out_lo = out_lo; a.addend = addend;
  a.B = B; a.H = H; a.W = W; a.C = C;
  a.tiles_w = (W + DW_TILE - 1) / DW_TILE;
  a.tiles_h = (H + DW_TILE - 1) / DW_TILE;
  a.eps = eps;
  a.out_dtype = out_dtype;

  const double bytes = (double)B * H * W * C * (4.0 + (out_dtype == CCX_BF16 ? 2.0 : (out_lo ? 8.0 : 4.0)));
  if (use_v3) {
    const int nc3 = C / DW3_CH;
    const int num_tiles = a.tiles_w * a.tiles_h * B;
    int n_clusters = 2 * num_sms() / nc3;       // two CTAs per SM
    if (n_clusters > num_tiles) n_clusters = num_tiles;
    cudaLaunchConfig_t cfg3{};
    cfg3.gridDim = dim3(nc3 * n_clusters, 1, 1);
    cfg3.blockDim = dim3(DW3_THREADS, 1, 1);
    cfg3.dynamicSmemBytes = DW3_SMEM;
    cfg3.stream = stream;
    cudaLaunchAttribute attr3[2];
    attr3[0].id = cudaLaunchAttributeClusterDimension;
    attr3[0].val.clusterDim.x = nc3;
    attr3[0].val.clusterDim.y = 1;
    attr3[0].val.clusterDim.z = 1;
    attr3[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr3[1].val.programmaticStreamSerializationAllowed = 1;
    cfg3.attrs = attr3;
    cfg3.numAttrs = 2;
    ProfScope prof(PROF_DWCONV_LN, stream, bytes);
    if (cudaLaunchKernelEx(&cfg3, dwconv7_ln_kernel_v3, tm, a, nc3, n_clusters, num_tiles) != cudaSuccess)
      return CCX_ERR_CUDA;
    return CCX_OK;
  }
  const int nc = C / CH;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nc, a.tiles_w * a.tiles_h, B);
  cfg.blockDim = dim3(use_v1 ? DW_THREADS : CH, 1, 1);
  cfg.dynamicSmemBytes = (CH == 64) ? DW_SMEM64 : DW_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // v2 kernels call grid_dep_sync()
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_v1 ? 1 : 2;
  ProfScope prof(PROF_DWCONV_LN, stream, bytes);
  if (use_v1) {
    if (cudaLaunchKernelEx(&cfg, dwconv7_ln_kernel, tm, a) != cudaSuccess) return CCX_ERR_CUDA;
  } else {
    if (CH == 64) {
      if (cudaLaunchKernelEx(&cfg, dwconv7_ln_kernel_v2<64>, tm, a) != cudaSuccess) return CCX_ERR_CUDA;
    } else {
      if (cudaLaunchKernelEx(&cfg, dwconv7_ln_kernel_v2<128>, tm, a) != cudaSuccess) return CCX_ERR_CUDA;
    }
  }
  return CCX_OK;
}

}  // namespace ccx
