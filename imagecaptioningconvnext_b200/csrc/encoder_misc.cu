// encoder_misc.cu — the small bandwidth-bound kernels around the ConvNeXt GEMMs.
//
//   stem_ln           torchvision/models/convnext.py:120-131  Conv2d(3,128,k=4,s=4)+bias -> LayerNorm2d(eps 1e-6)
//                     reads NCHW fp32 images directly, writes the NHWC fp32 residual stream
//   ln_patchmerge     torchvision/models/convnext.py:146-151  LayerNorm2d(Cin) and the im2col half of
//                     Conv2d(Cin,2Cin,k=2,s=2): writes rows [b,oh,ow] x cols [(kh,kw,c)] for the GEMM that follows
//   avgpool_nhwc      models/encoder.py:25-26  AdaptiveAvgPool2d((s,s)) + permute(0,2,3,1) -> contiguous (B,s,s,C)
//   split_tf32        helper: fp32 -> (tf32-representable hi, fp32 lo) pair for the 3xTF32 GEMM path
//   cast_bf16         helper: fp32 -> bf16
#include "ccx_common.cuh"
#include "ccx_gemm.h"
#include "ccx_ops.h"
#include "ccx_prof.h"

namespace ccx {

// ---------------------------------------------------------------------------------------------
// stem: one CTA = one output row segment of 64 pixels; warp = 8 pixels; lane = 4 output channels
// ---------------------------------------------------------------------------------------------
static constexpr int STEM_PX = 64;
static constexpr int STEM_C = 128;

// img_u8 != nullptr: raw uint8 pixels; (x/255 - mean[c]) * inv_std[c] is applied while staging
// (dataLoader.py:43-45: FloatTensor(img / 255.) then transforms.Normalize(mean, std))
__global__ void __launch_bounds__(256)
stem_ln_kernel(const float* __restrict__ img, const unsigned char* __restrict__ img_u8,
               const float* __restrict__ mean, const float* __restrict__ inv_std,
               const float* __restrict__ wk,  // wk [48][128], k = c*16+kh*4+kw
               const float* __restrict__ bias, const float* __restrict__ gamma,
               const float* __restrict__ beta, float* __restrict__ out, int B, int Hin, int Win, int Hout,
               int Wout, float eps) {
  __shared__ __align__(16) float w_s[48 * STEM_C];         // 24 KB
  __shared__ __align__(16) float in_s[12][STEM_PX * 4];    // [c*4+kh][col] 12 KB
  const int segs = (Wout + STEM_PX - 1) / STEM_PX;
  const int seg = blockIdx.x % segs;
  const int oh = (blockIdx.x / segs) % Hout;
  const int b = blockIdx.x / (segs * Hout);
  const int ow0 = seg * STEM_PX;

  for (int i = threadIdx.x; i < 48 * STEM_C / 4; i += 256)
    reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(wk) + i);
  for (int i = threadIdx.x; i < 12 * STEM_PX * 4; i += 256) {
    const int rowi = i / (STEM_PX * 4), col = i - rowi * (STEM_PX * 4);
    const int c = rowi >> 2, kh = rowi & 3;
    const int ih = oh * 4 + kh, iw = ow0 * 4 + col;
    float v = 0.f;
    if (iw < Wout * 4) {
      const long long gi = ((static_cast<long long>(b) * 3 + c) * Hin + ih) * Win + iw;
      if (img_u8 != nullptr)
        v = (static_cast<float>(img_u8[gi]) / 255.0f - __ldg(mean + c)) * __ldg(inv_std + c);
      else
        v = __ldg(img + gi);
    }
    in_s[rowi][col] = v;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[8][4];
  const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + lane);
#pragma unroll
  for (int p = 0; p < 8; ++p) { acc[p][0] = bv.x; acc[p][1] = bv.y; acc[p][2] = bv.z; acc[p][3] = bv.w; }
#pragma unroll
  for (int rowi = 0; rowi < 12; ++rowi) {
    float4 wv[4];
#pragma unroll
    for (int kw = 0; kw < 4; ++kw)
      wv[kw] = *reinterpret_cast<const float4*>(&w_s[(rowi * 4 + kw) * STEM_C + lane * 4]);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const float4 iv = *reinterpret_cast<const float4*>(&in_s[rowi][(warp * 8 + p) * 4]);  // broadcast
      const float in4[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        acc[p][0] = fmaf(in4[kw], wv[kw].x, acc[p][0]);
        acc[p][1] = fmaf(in4[kw], wv[kw].y, acc[p][1]);
        acc[p][2] = fmaf(in4[kw], wv[kw].z, acc[p][2]);
        acc[p][3] = fmaf(in4[kw], wv[kw].w, acc[p][3]);
      }
    }
  }
  const float4 gv = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + lane);
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float s = warp_sum(acc[p][0] + acc[p][1] + acc[p][2] + acc[p][3]);
    const float mean = s * (1.0f / STEM_C);
    const float d0 = acc[p][0] - mean, d1 = acc[p][1] - mean, d2 = acc[p][2] - mean, d3 = acc[p][3] - mean;
    const float q = warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
    const float rstd = rsqrtf(q * (1.0f / STEM_C) + eps);
    const int ow = ow0 + warp * 8 + p;
    if (ow < Wout) {
      float4 o;
      o.x = d0 * rstd * gv.x + be.x;
      o.y = d1 * rstd * gv.y + be.y;
      o.z = d2 * rstd * gv.z + be.z;
      o.w = d3 * rstd * gv.w + be.w;
      reinterpret_cast<float4*>(out + ((static_cast<long long>(b) * Hout + oh) * Wout + ow) * STEM_C)[lane] = o;
    }
  }
}

// v2 (used when rows are 16-byte addressable): persistent CTAs keep the 24 KB filter in shared memory across many
// 64-pixel row segments, the next segment's pixels are fetched into registers while the current one is computed
// (double-buffered staging, one barrier per segment), and the MACs are packed FFMA2.
__global__ void __launch_bounds__(256, 2)
stem_ln_kernel_v2(const float* __restrict__ img, const unsigned char* __restrict__ img_u8,
                  const float* __restrict__ mean, const float* __restrict__ inv_std, const float* __restrict__ wk,
                  const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float* __restrict__ out, int B, int Hin, int Win, int Hout, int Wout, float eps) {
  __shared__ __align__(16) float w_s[48 * STEM_C];            // 24 KB
  __shared__ __align__(16) float in_s[2][12][STEM_PX * 4];    // 24 KB
  const int segs = (Wout + STEM_PX - 1) / STEM_PX;
  const int items = B * Hout * segs;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool u8 = img_u8 != nullptr;

  for (int i = tid; i < 48 * STEM_C / 4; i += 256)
    reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(wk) + i);

  // staging slot k of this thread: 4 consecutive input columns of row (c, kh)
  float4 stage[3];
  auto fetch = [&](int item) {
    const int seg = item % segs, oh = (item / segs) % Hout, b = item / (segs * Hout);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = tid + k * 256, rowi = i >> 6, chunk = i & 63;
      const int c = rowi >> 2, kh = rowi & 3;
      const int iw = (seg * STEM_PX + chunk) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iw < Wout * 4) {
        const long long gi = ((static_cast<long long>(b) * 3 + c) * Hin + oh * 4 + kh) * Win + iw;
        if (u8) {
          const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(img_u8 + gi));
          const float m = __ldg(mean + c), is = __ldg(inv_std + c);
          v.x = (static_cast<float>(q.x) / 255.0f - m) * is;
          v.y = (static_cast<float>(q.y) / 255.0f - m) * is;
          v.z = (static_cast<float>(q.z) / 255.0f - m) * is;
          v.w = (static_cast<float>(q.w) / 255.0f - m) * is;
        } else {
          v = __ldg(reinterpret_cast<const float4*>(img + gi));
        }
      }
      stage[k] = v;
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = tid + k * 256;
      *reinterpret_cast<float4*>(&in_s[buf][i >> 6][(i & 63) * 4]) = stage[k];
    }
  };

  const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + lane);
  const float4 gv = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  int item = blockIdx.x, buf = 0;
  if (item < items) { fetch(item); commit(0); }
  __syncthreads();
  for (; item < items; item += gridDim.x, buf ^= 1) {
    const int next = item + gridDim.x;
    if (next < items) fetch(next);
    float2 acc[8][2];
#pragma unroll
    for (int p = 0; p < 8; ++p) { acc[p][0] = make_float2(bv.x, bv.y); acc[p][1] = make_float2(bv.z, bv.w); }
#pragma unroll
    for (int rowi = 0; rowi < 12; ++rowi) {
      float4 wv[4];
#pragma unroll
      for (int kw = 0; kw < 4; ++kw)
        wv[kw] = *reinterpret_cast<const float4*>(&w_s[(rowi * 4 + kw) * STEM_C + lane * 4]);
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const float4 iv = *reinterpret_cast<const float4*>(&in_s[buf][rowi][(warp * 8 + p) * 4]);  // broadcast
        const float in4[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          const float2 a = make_float2(in4[kw], in4[kw]);
          acc[p][0] = ffma2(a, make_float2(wv[kw].x, wv[kw].y), acc[p][0]);
          acc[p][1] = ffma2(a, make_float2(wv[kw].z, wv[kw].w), acc[p][1]);
        }
      }
    }
    const int seg = item % segs, oh = (item / segs) % Hout, b = item / (segs * Hout);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const float s = warp_sum(acc[p][0].x + acc[p][0].y + acc[p][1].x + acc[p][1].y);
      const float mu = s * (1.0f / STEM_C);
      const float d0 = acc[p][0].x - mu, d1 = acc[p][0].y - mu, d2 = acc[p][1].x - mu, d3 = acc[p][1].y - mu;
      const float q = warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
      const float rstd = rsqrtf(q * (1.0f / STEM_C) + eps);
      const int ow = seg * STEM_PX + warp * 8 + p;
      if (ow < Wout) {
        float4 o;
        o.x = d0 * rstd * gv.x + be.x;
        o.y = d1 * rstd * gv.y + be.y;
        o.z = d2 * rstd * gv.z + be.z;
        o.w = d3 * rstd * gv.w + be.w;
        reinterpret_cast<float4*>(out + ((static_cast<long long>(b) * Hout + oh) * Wout + ow) * STEM_C)[lane] = o;
      }
    }
    if (next < items) commit(buf ^ 1);
    __syncthreads();
  }
}

int stem_ln(const float* img, const unsigned char* img_u8, const float* mean, const float* inv_std, const float* wk,
            const float* bias, const float* gamma, const float* beta, float* out, int B, int Hin, int Win, float eps,
            cudaStream_t stream) {
  if (B <= 0 || Hin < 4 || Win < 4) return CCX_ERR_SHAPE;
  const int Hout = Hin / 4, Wout = Win / 4;
  const int segs = (Wout + STEM_PX - 1) / STEM_PX;
  const long long grid = static_cast<long long>(B) * Hout * segs;
  if (grid > 0x7fffffffLL) return CCX_ERR_SHAPE;
  if ((img == nullptr) == (img_u8 == nullptr)) return CCX_ERR_DTYPE;
  if (img_u8 != nullptr && (mean == nullptr || inv_std == nullptr)) return CCX_ERR_SHAPE;
  ProfScope prof(PROF_STEM, stream,
                 (double)B * (3.0 * Hin * Win * (img_u8 ? 1.0 : 4.0) + (double)Hout * Wout * STEM_C * 4.0));
  const bool vec_ok = (Win % 4 == 0) && (reinterpret_cast<uintptr_t>(img) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(img_u8) % 4 == 0);
  if (vec_ok) {
    const long long cap = 2LL * num_sms();
    stem_ln_kernel_v2<<<static_cast<unsigned>(grid < cap ? grid : cap), 256, 0, stream>>>(
        img, img_u8, mean, inv_std, wk, bias, gamma, beta, out, B, Hin, Win, Hout, Wout, eps);
  } else {
    stem_ln_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(img, img_u8, mean, inv_std, wk, bias, gamma,
                                                                     beta, out, B, Hin, Win, Hout, Wout, eps);
  }
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// row LayerNorm (one warp per row, two-pass in registers) with an optional 2x2 patch-merge scatter
// ---------------------------------------------------------------------------------------------
template <int VPL>  // float4 vectors per lane: C = 128 * VPL
__global__ void __launch_bounds__(256)
ln_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
               void* __restrict__ out, float* __restrict__ out_lo, float* __restrict__ out_plain, long long M,
               int C, float eps, int out_dtype, int merge, int H, int W) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (m >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + m * C);
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = __ldg(xr + i * 32 + lane);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
  // destination row / column offset
  long long orow = m;
  int ocol0 = 0, ldo = C;
  if (merge) {
    const int w = static_cast<int>(m % W);
    const int h = static_cast<int>((m / W) % H);
    const long long b = m / (static_cast<long long>(W) * H);
    orow = (b * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1);
    ocol0 = ((h & 1) * 2 + (w & 1)) * C;
    ldo = 4 * C;
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y;
    y.x = v[i].x * rstd * g.x + bb.x;
    y.y = v[i].y * rstd * g.y + bb.y;
    y.z = v[i].z * rstd * g.z + bb.z;
    y.w = v[i].w * rstd * g.w + bb.w;
    const long long o = orow * ldo + ocol0 + c;
    if (out_plain != nullptr) *reinterpret_cast<float4*>(out_plain + o) = y;
    if (out == nullptr) continue;
    if (out_dtype == CCX_BF16) {
      uint2 pk;
      pk.x = pack_bf16x2(y.x, y.y);
      pk.y = pack_bf16x2(y.z, y.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + o) = pk;
    } else if (out_lo != nullptr) {
      float4 hi, lo;
      hi.x = tf32_hi(y.x); lo.x = y.x - hi.x;
      hi.y = tf32_hi(y.y); lo.y = y.y - hi.y;
      hi.z = tf32_hi(y.z); lo.z = y.z - hi.z;
      hi.w = tf32_hi(y.w); lo.w = y.w - hi.w;
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = hi;
      *reinterpret_cast<float4*>(out_lo + o) = lo;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = y;
    }
  }
}

int ln_rows(const float* x, const float* gamma, const float* beta, void* out, float* out_lo, float* out_plain,
            long long M, int C, float eps, int out_dtype, int merge, int H, int W, cudaStream_t stream) {
  if (M <= 0) return M == 0 ? CCX_OK : CCX_ERR_SHAPE;
  if (C % 128 != 0 || C > 1024) return CCX_ERR_SHAPE;
  if (merge && ((H & 1) || (W & 1) || H <= 0 || W <= 0 || (M % (static_cast<long long>(H) * W)) != 0))
    return CCX_ERR_SHAPE;
  const unsigned grid = static_cast<unsigned>((M + 7) / 8);
  ProfScope prof(PROF_LN_ROWS, stream, (double)M * C * (4.0 + (out_dtype == CCX_BF16 ? 2.0 : (out_lo ? 8.0 : 4.0))));
#define CCX_LN_CASE(V)                                                                                      \
  case V:                                                                                                   \
    ln_rows_kernel<V><<<grid, 256, 0, stream>>>(x, gamma, beta, out, out_lo, out_plain, M, C, eps, out_dtype, \
                                                merge, H, W);                                               \
    break;
  switch (C / 128) {
    CCX_LN_CASE(1) CCX_LN_CASE(2) CCX_LN_CASE(3) CCX_LN_CASE(4) CCX_LN_CASE(5) CCX_LN_CASE(6) CCX_LN_CASE(7)
    CCX_LN_CASE(8)
    default: return CCX_ERR_SHAPE;
  }
#undef CCX_LN_CASE
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// adaptive average pool over NHWC (torch bin rule: start = floor(i*H/s), end = ceil((i+1)*H/s))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
avgpool_nhwc_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int H, int W, int C, int S) {
  const long long total = static_cast<long long>(B) * S * S * (C / 4);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % (C / 4));
    const int ow = static_cast<int>((i / (C / 4)) % S);
    const int oh = static_cast<int>((i / (static_cast<long long>(C / 4) * S)) % S);
    const int b = static_cast<int>(i / (static_cast<long long>(C / 4) * S * S));
    const int hs = (oh * H) / S, he = ((oh + 1) * H + S - 1) / S;
    const int ws = (ow * W) / S, we = ((ow + 1) * W + S - 1) / S;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        const float4 v =
            __ldg(reinterpret_cast<const float4*>(x + ((static_cast<long long>(b) * H + h) * W + w) * C) + c4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    const float inv = 1.0f / static_cast<float>((he - hs) * (we - ws));
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

int avgpool_nhwc(const float* x, float* out, int B, int H, int W, int C, int S, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || S <= 0 || (C % 4) != 0) return CCX_ERR_SHAPE;
  const long long total = static_cast<long long>(B) * S * S * (C / 4);
  const unsigned grid = static_cast<unsigned>(total / 256 + 1 > 148 * 16 ? 148 * 16 : total / 256 + 1);
  ProfScope prof(PROF_POOL, stream, ((double)B * H * W * C + (double)total * 4) * 4.0);
  avgpool_nhwc_kernel<<<grid, 256, 0, stream>>>(x, out, B, H, W, C, S);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// element-wise helpers
// ---------------------------------------------------------------------------------------------
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                  long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    const float h = tf32_hi(v);
    hi[i] = h;
    lo[i] = v - h;
  }
}
int split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t stream) {
  if (n <= 0) return n == 0 ? CCX_OK : CCX_ERR_SHAPE;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)n * 12.0);
  split_tf32_kernel<<<grid, 256, 0, stream>>>(x, hi, lo, n);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}
// 8 elements per thread: two 16-byte loads, one 16-byte store
__global__ void __launch_bounds__(256) cast_bf16x8_kernel(const float4* __restrict__ x, uint4* __restrict__ y, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = x[2 * i], b = x[2 * i + 1];
  uint4 pk;
  pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(a.z, a.w);
  pk.z = pack_bf16x2(b.x, b.y); pk.w = pack_bf16x2(b.z, b.w);
  y[i] = pk;
}
int cast_bf16(const float* x, void* y, long long n, cudaStream_t stream) {
  if (n <= 0) return n == 0 ? CCX_OK : CCX_ERR_SHAPE;
  ProfScope prof(PROF_ELEMENTWISE, stream, (double)n * 6.0);
  if ((n % 8) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const long long n8 = n / 8;
    cast_bf16x8_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<uint4*>(y), n8);
    return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
  }
  const unsigned grid = static_cast<unsigned>((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
  cast_bf16_kernel<<<grid, 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n);
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

}  // namespace ccx
