// encoder_bwd.cu — backward kernels of the fine-tuned ConvNeXt stage (reference: autograd through
// torchvision/models/convnext.py:51-67 for `Encoder.fine_tune(True, startingLayer)`, models/encoder.py:29-34).
//
//   scale_rows_cols     d(block output) -> d(layer_scale * stochastic_depth branch):  x * colscale[c] * rowscale[m/g]
//   gelu_bwd            dpre = dh * gelu'(pre)   (exact erf GELU)
//   cnblock_param_grads dW2 / d layer_scale / d b2 from the un-scaled wgrad G = dout'^T . h  (no recompute of z)
//   dwconv7_wgrad       depthwise filter gradient, NHWC
//   avgpool_nhwc_bwd    AdaptiveAvgPool2d backward
// The depthwise-conv data gradient re-uses dwconv7_ln_kernel in "plain" mode with flipped taps.
#include "ccx_common.cuh"
#include "ccx_gemm.h"
#include "ccx_ops.h"
#include "ccx_prof.h"

namespace ccx {

__global__ void __launch_bounds__(256)
scale_rows_cols_kernel(const float* __restrict__ x, const float* __restrict__ colscale,
                       const float* __restrict__ rowscale, int rows_per_group, float* __restrict__ out, long long M,
                       int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M * (C / 4)) return;
  const long long m = i / (C / 4);
  const int c4 = static_cast<int>(i % (C / 4));
  float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  float rs = rowscale ? __ldg(rowscale + m / rows_per_group) : 1.f;
  if (colscale) {
    const float4 cs = __ldg(reinterpret_cast<const float4*>(colscale) + c4);
    v.x *= cs.x; v.y *= cs.y; v.z *= cs.z; v.w *= cs.w;
  }
  v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
  reinterpret_cast<float4*>(out)[i] = v;
}
int scale_rows_cols(const float* x, const float* colscale, const float* rowscale, int rows_per_group, float* out,
                    long long M, int C, cudaStream_t stream) {
  if (M <= 0) return CCX_OK;
  if (C % 4) return CCX_ERR_SHAPE;
  const long long n = M * (C / 4);
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)M * C * 8.0);
  scale_rows_cols_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      x, colscale, rowscale, rows_per_group > 0 ? rows_per_group : 1, out, M, C);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const float* __restrict__ pre, float* __restrict__ dh, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = pre[i];
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    dh[i] *= cdf + x * pdf;
  }
}
int gelu_bwd(const float* pre, float* dh, long long n, cudaStream_t stream) {
  if (n <= 0) return CCX_OK;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256 > 148 * 32 ? 148 * 32 : (n + 255) / 256);
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)n * 12.0);
  gelu_bwd_kernel<<<grid, 256, 0, stream>>>(pre, dh, n);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// G[c, :] = sum_m dout'[m,c] h[m,:]  (dout' = dout * rowscale).  With z = h W2^T + b2 and out = x + gamma*rs*z:
//   dW2[c,:] += gamma[c] * G[c,:] ;  d gamma[c] += W2[c,:] . G[c,:] + b2[c] * s[c] ;  d b2[c] += gamma[c] * s[c]
// where s[c] = sum_m dout'[m,c].  One warp per output channel c.
__global__ void __launch_bounds__(128)
cnblock_param_grads_kernel(const float* __restrict__ G, const float* __restrict__ W2, const float* __restrict__ b2,
                           const float* __restrict__ gamma, const float* __restrict__ s, float* __restrict__ dW2,
                           float* __restrict__ dgamma, float* __restrict__ db2, int C, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 4 + warp;
  if (c >= C) return;
  const float g = gamma[c];
  float dot = 0.f;
  const long long row = static_cast<long long>(c) * K;
  if ((K & 3) == 0 && ((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(W2) | reinterpret_cast<uintptr_t>(dW2)) & 15) == 0) {
    const float4* G4 = reinterpret_cast<const float4*>(G + row);
    const float4* W4 = reinterpret_cast<const float4*>(W2 + row);
    float4* D4 = reinterpret_cast<float4*>(dW2 + row);
    for (int k = lane; k < (K >> 2); k += 32) {
      const float4 gv = __ldg(G4 + k), wv = __ldg(W4 + k);
      float4 d = D4[k];
      dot = fmaf(wv.x, gv.x, fmaf(wv.y, gv.y, fmaf(wv.z, gv.z, fmaf(wv.w, gv.w, dot))));
      d.x = fmaf(g, gv.x, d.x); d.y = fmaf(g, gv.y, d.y); d.z = fmaf(g, gv.z, d.z); d.w = fmaf(g, gv.w, d.w);
      D4[k] = d;
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float gv = G[row + k];
      dot = fmaf(W2[row + k], gv, dot);
      dW2[row + k] += g * gv;
    }
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    dgamma[c] += dot + b2[c] * s[c];
    db2[c] += g * s[c];
  }
}
int cnblock_param_grads(const float* G, const float* W2, const float* b2, const float* gamma, const float* s,
                        float* dW2, float* dgamma, float* db2, int C, int K, cudaStream_t stream) {
  if (C <= 0) return CCX_OK;
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)C * K * 16.0);
  cnblock_param_grads_kernel<<<(C + 3) / 4, 128, 0, stream>>>(G, W2, b2, gamma, s, dW2, dgamma, db2, C, K);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// dw[tap][c] += sum_{b,h,w} du[b,h,w,c] * x[b,h+kh-3,w+kw-3,c]
// one CTA = 64 channels x one 8x8 output patch x a RANGE of images: per image the 14x14 halo of x (zero padded) and
// the 8x8 patch of du are staged in shared memory, then thread (c, q) accumulates the 49 taps over output rows 2q, 2q+1
// in registers — an x row segment of 14 values serves 8 pixels x 7 taps, so shared memory is read once per 3.7 FMAs —
// and keeps accumulating over the CTA's images; the four row-pair partials are summed through shared memory and ONE
// atomicAdd per (tap, channel) leaves the CTA.  (History: re-reading x through L1 49 times per pixel: 128 us per
// launch at batch 32 / C=1024; one image per CTA, one LDS per FMA and 49 atomics per thread — 6.4 M atomics on 50 K
// addresses: 50 us.)
static constexpr int WG_C = 64;
__global__ void __launch_bounds__(256)
dwconv7_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ du, float* __restrict__ dw, int B, int H,
                     int W, int C, int tiles_w, int imgs_per_cta) {
  extern __shared__ __align__(16) float wg_sm[];
  float* xs = wg_sm;                     // [14][14][WG_C]   (afterwards: [4 q][49][WG_C] partials, same size)
  float* ds = wg_sm + 14 * 14 * WG_C;    // [8][8][WG_C]
  const int c0 = blockIdx.x * WG_C;
  const int th0 = (blockIdx.y / tiles_w) * 8, tw0 = (blockIdx.y % tiles_w) * 8;
  const int b_begin = blockIdx.z * imgs_per_cta;
  const int b_end = min(B, b_begin + imgs_per_cta);
  const int c = threadIdx.x % WG_C, q = threadIdx.x / WG_C;
  float acc[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) acc[t] = 0.f;
  for (int b = b_begin; b < b_end; ++b) {
    const float* xb = x + static_cast<long long>(b) * H * W * C;
    const float* db = du + static_cast<long long>(b) * H * W * C;
    for (int i = threadIdx.x; i < 14 * 14 * WG_C; i += 256) {
      const int cc = i % WG_C, pix = i / WG_C;
      const int ih = th0 + pix / 14 - 3, iw = tw0 + pix % 14 - 3;
      xs[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W && c0 + cc < C)
                  ? __ldg(xb + (static_cast<long long>(ih) * W + iw) * C + c0 + cc) : 0.f;
    }
    for (int i = threadIdx.x; i < 8 * 8 * WG_C; i += 256) {
      const int cc = i % WG_C, pix = i / WG_C;
      const int h = th0 + pix / 8, w = tw0 + pix % 8;
      ds[i] = (h < H && w < W && c0 + cc < C) ? __ldg(db + (static_cast<long long>(h) * W + w) * C + c0 + cc) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int h = q * 2 + hh;
      float d[8];
#pragma unroll
      for (int w = 0; w < 8; ++w) d[w] = ds[(h * 8 + w) * WG_C + c];
#pragma unroll
      for (int kh = 0; kh < 7; ++kh) {
        float xr[14];
#pragma unroll
        for (int j = 0; j < 14; ++j) xr[j] = xs[((h + kh) * 14 + j) * WG_C + c];
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) {
          float a = acc[kh * 7 + kw];
#pragma unroll
          for (int w = 0; w < 8; ++w) a = fmaf(d[w], xr[w + kw], a);
          acc[kh * 7 + kw] = a;
        }
      }
    }
    __syncthreads();
  }
  // sum the four row-pair partials, one atomic per (tap, channel)
#pragma unroll
  for (int t = 0; t < 49; ++t) xs[(q * 49 + t) * WG_C + c] = acc[t];
  __syncthreads();
  for (int i = threadIdx.x; i < 49 * WG_C; i += 256) {
    const int cc = i % WG_C, t = i / WG_C;
    if (c0 + cc < C) {
      const float v = xs[(0 * 49 + t) * WG_C + cc] + xs[(1 * 49 + t) * WG_C + cc] + xs[(2 * 49 + t) * WG_C + cc] +
                      xs[(3 * 49 + t) * WG_C + cc];
      atomicAdd(dw + static_cast<long long>(t) * C + c0 + cc, v);
    }
  }
}
int dwconv7_wgrad(const float* x, const float* du, float* dw49c, int B, int H, int W, int C, cudaStream_t stream) {
  if (B <= 0) return CCX_OK;
  const int tiles_h = (H + 7) / 8, tiles_w = (W + 7) / 8;
  const int groups = (C + WG_C - 1) / WG_C;
  if (tiles_h * tiles_w > 65535) return CCX_ERR_SHAPE;
  constexpr int smem = (14 * 14 + 8 * 8) * WG_C * static_cast<int>(sizeof(float));
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.ref();
  if (!configured) {
    if (cudaFuncSetAttribute(dwconv7_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return CCX_ERR_CUDA;
    configured = true;
  }
  // about two CTAs per SM in flight; each keeps its taps in registers across its images
  int splits = (2 * num_sms()) / (groups * tiles_h * tiles_w);
  splits = splits < 1 ? 1 : (splits > B ? B : splits);
  const int imgs_per_cta = (B + splits - 1) / splits;
  splits = (B + imgs_per_cta - 1) / imgs_per_cta;
  if (splits > 65535) return CCX_ERR_SHAPE;
  ProfScope prof(PROF_DWCONV_LN, stream, (double)B * H * W * C * 8.0);
  dwconv7_wgrad_kernel<<<dim3(groups, tiles_h * tiles_w, splits), 256, smem, stream>>>(x, du, dw49c, B, H, W, C, tiles_w,
                                                                                       imgs_per_cta);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// adaptive average pool backward: dx[b,h,w,:] = sum over bins containing (h,w) of dout[b,oh,ow,:] / bin_size
__global__ void __launch_bounds__(256)
avgpool_nhwc_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int B, int H, int W, int C, int S) {
  const long long total = static_cast<long long>(B) * H * W * (C / 4);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % (C / 4));
    const int w = static_cast<int>((i / (C / 4)) % W);
    const int h = static_cast<int>((i / (static_cast<long long>(C / 4) * W)) % H);
    const int b = static_cast<int>(i / (static_cast<long long>(C / 4) * W * H));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // output windows that can contain (h, w): [floor(h S / H) - 1, floor((h + 1) S / H) + 1]; the exact membership test
    // stays below (scanning all S x S windows per element cost 33 us at B = 32, ncu r02)
    const int oh_lo = max((h * S) / H - 1, 0), oh_hi = min(((h + 1) * S) / H + 1, S - 1);
    const int ow_lo = max((w * S) / W - 1, 0), ow_hi = min(((w + 1) * S) / W + 1, S - 1);
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      const int hs = (oh * H) / S, he = ((oh + 1) * H + S - 1) / S;
      if (h < hs || h >= he) continue;
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const int ws = (ow * W) / S, we = ((ow + 1) * W + S - 1) / S;
        if (w < ws || w >= we) continue;
        const float inv = 1.0f / static_cast<float>((he - hs) * (we - ws));
        const float4 v = __ldg(reinterpret_cast<const float4*>(dout + ((static_cast<long long>(b) * S + oh) * S + ow) * C) + c4);
        acc.x += v.x * inv; acc.y += v.y * inv; acc.z += v.z * inv; acc.w += v.w * inv;
      }
    }
    reinterpret_cast<float4*>(dx)[i] = acc;
  }
}
int avgpool_nhwc_bwd(const float* dout, float* dx, int B, int H, int W, int C, int S, cudaStream_t stream) {
  if (B <= 0 || (C % 4)) return B <= 0 ? CCX_OK : CCX_ERR_SHAPE;
  const long long total = static_cast<long long>(B) * H * W * (C / 4);
  const unsigned grid = static_cast<unsigned>(total / 256 + 1 > 148 * 16 ? 148 * 16 : total / 256 + 1);
  ProfScope prof(PROF_POOL, stream, (double)B * (H * W + S * S) * C * 4.0);
  avgpool_nhwc_bwd_kernel<<<grid, 256, 0, stream>>>(dout, dx, B, H, W, C, S);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

}  // namespace ccx
