// gemm_skinny.cu — C[M,N] = A[M,K] . B[N,K]^T (+ bias, + residual) for M <= 32 rows, bf16 operands, fp32 out.
//
// The recurrent GEMMs of the LSTM decoder (models/decoder.py:100-108: decoder_att / f_beta on h, the LSTMCell gate
// GEMM, and their two dgrad counterparts in BPTT) have M = active batch rows <= 32 and run ~200 times per train
// step, each strictly after the previous one.  On the 128-row tcgen05 tile kernel they cost 11-19 us apiece, all
// of it latency: 16-32 CTAs each streaming K=2048 through a TMA pipeline plus the TMEM / barrier / tensor-map
// prologue.  Here the N dimension is spread over (almost) all SMs — 8 or 16 output columns per CTA — and the 8
// warps of a CTA split K; operands go straight from L2 to registers as 128-bit loads laid out so that they ARE
// the mma.sync.m16n8k16 fragments (a k-permutation inside each 32-wide K chunk, the same for A and B, leaves the
// dot products unchanged), partial sums meet in shared memory.  No tensor-memory, no TMA: at M <= 32 the tensor
// pipe is idle either way and the job is to get 8 MB of weights past the SMs once, quickly.
#include <cuda_bf16.h>

#include "ccx_common.cuh"
#include "ccx_gemm.h"
#include "ccx_prof.h"

namespace ccx {

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
               "{%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

static constexpr int SK_WARPS = 8;

template <int NT>   // NT n-tiles of 8 columns per CTA
__global__ void __launch_bounds__(SK_WARPS * 32)
gemm_skinny_kernel(const __nv_bfloat16* __restrict__ A, long long lda, const __nv_bfloat16* __restrict__ B,
                   long long ldb, float* __restrict__ C, long long ldc, const float* __restrict__ bias,
                   const float* __restrict__ residual, long long ldr, int M, int N, int K) {
  constexpr int BN = 8 * NT;
  __shared__ float red[SK_WARPS][32][BN + 1];
  grid_dep_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * BN;
  float acc[2][NT][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mi][nt][i] = 0.f;
  // this thread's fragment rows: A rows g, g+8, g+16, g+24; B rows (= output columns) n0 + nt*8 + g
  const __nv_bfloat16* ap[4];
  bool aok[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    aok[r] = (g + 8 * r) < M;
    ap[r] = A + static_cast<long long>(aok[r] ? g + 8 * r : 0) * lda + t * 8;
  }
  const __nv_bfloat16* bp[NT];
  bool bok[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    bok[nt] = (n0 + nt * 8 + g) < N;
    bp[nt] = B + static_cast<long long>(bok[nt] ? n0 + nt * 8 + g : 0) * ldb + t * 8;
  }
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 4
  for (int kc = warp * 32; kc < K; kc += SK_WARPS * 32) {
    uint4 av[4], bv[NT];
#pragma unroll
    for (int r = 0; r < 4; ++r) av[r] = aok[r] ? __ldg(reinterpret_cast<const uint4*>(ap[r] + kc)) : zero;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) bv[nt] = bok[nt] ? __ldg(reinterpret_cast<const uint4*>(bp[nt] + kc)) : zero;
    // physical k (kc + t*8 + 0..7) -> two mma k-steps: {x,y} and {z,w} (see file header)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      mma_bf16_16816(acc[0][nt], av[0].x, av[1].x, av[0].y, av[1].y, bv[nt].x, bv[nt].y);
      mma_bf16_16816(acc[1][nt], av[2].x, av[3].x, av[2].y, av[3].y, bv[nt].x, bv[nt].y);
      mma_bf16_16816(acc[0][nt], av[0].z, av[1].z, av[0].w, av[1].w, bv[nt].z, bv[nt].w);
      mma_bf16_16816(acc[1][nt], av[2].z, av[3].z, av[2].w, av[3].w, bv[nt].z, bv[nt].w);
    }
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      red[warp][mi * 16 + g][nt * 8 + t * 2] = acc[mi][nt][0];
      red[warp][mi * 16 + g][nt * 8 + t * 2 + 1] = acc[mi][nt][1];
      red[warp][mi * 16 + g + 8][nt * 8 + t * 2] = acc[mi][nt][2];
      red[warp][mi * 16 + g + 8][nt * 8 + t * 2 + 1] = acc[mi][nt][3];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * BN; i += SK_WARPS * 32) {
    const int r = i / BN, c = i % BN, n = n0 + c;
    if (r >= M || n >= N) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) v += red[w][r][c];
    if (bias != nullptr) v += __ldg(bias + n);
    if (residual != nullptr) v += residual[static_cast<long long>(r) * ldr + n];
    C[static_cast<long long>(r) * ldc + n] = v;
  }
}

static int g_skinny = -1;

bool gemm_skinny_eligible(const GemmDesc& g) {
  if (g_skinny < 0) g_skinny = getenv("CCX_GEMM_SKINNY") ? atoi(getenv("CCX_GEMM_SKINNY")) : 1;
  return g_skinny && !g.a_mn && !g.b_mn && !g.res_mul && g.in_dtype == CCX_BF16 && g.out_dtype == CCX_F32 && g.M >= 1 && g.M <= 32 && g.act == 0 &&
         g.colscale == nullptr && g.rowscale == nullptr && g.emask == nullptr && !g.split && g.force_bn == 0 &&
         g.A_lo == nullptr && g.B_lo == nullptr && (g.K % 32) == 0 && (g.lda % 8) == 0 && (g.ldb % 8) == 0 &&
         g.N >= 64 && (reinterpret_cast<uintptr_t>(g.A) % 16) == 0 && (reinterpret_cast<uintptr_t>(g.B) % 16) == 0;
}

int gemm_skinny(const GemmDesc& g, cudaStream_t stream) {
  ProfScope prof(PROF_GEMM_SKINNY, stream, 2.0 * g.M * (double)g.N * g.K);
  const auto* A = static_cast<const __nv_bfloat16*>(g.A);
  const auto* B = static_cast<const __nv_bfloat16*>(g.B);
  auto* C = static_cast<float*>(g.C);
  const auto* R = static_cast<const float*>(g.residual);
  // 16 columns per CTA when that still gives ~one CTA per SM, else 8
  cudaError_t e;
  if ((g.N + 15) / 16 >= num_sms() * 3 / 4)
    e = launch_pdl(gemm_skinny_kernel<2>, dim3((g.N + 15) / 16), dim3(SK_WARPS * 32), 0, stream, A, g.lda, B, g.ldb, C,
                   g.ldc, g.bias, R, g.ldr, g.M, g.N, g.K);
  else
    e = launch_pdl(gemm_skinny_kernel<1>, dim3((g.N + 7) / 8), dim3(SK_WARPS * 32), 0, stream, A, g.lda, B, g.ldb, C,
                   g.ldc, g.bias, R, g.ldr, g.M, g.N, g.K);
  if (e != cudaSuccess) return CCX_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

}  // namespace ccx
