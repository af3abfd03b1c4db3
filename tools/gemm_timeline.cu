// gemm_timeline.cu — developer probe: where the cycles of csrc/gemm_tcgen05.cu go on the encoder's GEMM shapes.
// Builds the kernel with -DCCX_GEMM_TIMELINE (per-role wait accounting, epilogue switches); not part of libccx.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DCCX_GEMM_TIMELINE -I../imagecaptioningconvnext_b200/csrc \
//   gemm_timeline.cu ../imagecaptioningconvnext_b200/csrc/{gemm_tcgen05,gemm_tcgen05_2cta,gemm_skinny,prof}.cu -o gemm_timeline
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "ccx_gemm.h"

namespace ccx {
void gemm_timeline_mode(int mode);
void gemm_timeline_read(unsigned long long* out32);
void set_gemm_epilogue_override(int epi, int boxes);
void gemm_epi_timeline(unsigned long long* out8);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)

static void* g_flush;
static const size_t FLUSH_BYTES = 256u << 20;

static float time_once(const ccx::GemmDesc& g, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ts;
  for (int i = 0; i < iters; ++i) {
    CK(cudaMemsetAsync(g_flush, i, FLUSH_BYTES));
    cudaEventRecord(e0);
    int rc = ccx::gemm_tn(g, 0);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    if (rc) { printf("gemm rc=%d\n", rc); exit(3); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ts.push_back(ms * 1e3f);
  }
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

static void run(int M, int N, int K, int kind, int force_bn, int pair, int epi = -1, int boxes = 0) {
  ccx::set_gemm_epilogue_override(epi, boxes);
  void *A, *B, *C, *R; float *bias, *cs;
  const int oes = kind == 2 ? 4 : 2;                            // fc2 writes the fp32 residual stream in place
  CK(cudaMalloc(&A, (size_t)M * K * 2)); CK(cudaMalloc(&B, (size_t)N * K * 2)); CK(cudaMalloc(&C, (size_t)M * N * oes));
  R = C; CK(cudaMalloc(&bias, N * 4)); CK(cudaMalloc(&cs, N * 4));
  CK(cudaMemset(A, 0, (size_t)M * K * 2)); CK(cudaMemset(B, 0, (size_t)N * K * 2)); CK(cudaMemset(C, 0, (size_t)M * N * oes));
  CK(cudaMemset(bias, 0, N * 4)); CK(cudaMemset(cs, 0, N * 4));
  ccx::GemmDesc g;
  g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldc = N; g.ldr = N; g.A = A; g.B = B; g.C = C;
  g.bias = bias; g.in_dtype = 1; g.out_dtype = 1; g.force_bn = force_bn;
  if (kind == 1) g.act = 1;                                    // Linear + GELU (fc1)
  if (kind == 2) { g.colscale = cs; g.residual = R; g.out_dtype = 0; }   // layer-scale + fp32 residual, in place (fc2)
  ccx::set_gemm_pair_mode(pair);
  const double fl = 2.0 * M * N * K;
  const double bytes = 2.0 * ((double)M * K + (double)N * K) + (double)M * N * (kind == 2 ? 8 : 2);
  printf("M=%6d N=%5d K=%5d %s bn=%3d pair=%d epi=%s boxes=%d  (ideal: %.1f us tensor @1.5PF, %.1f us HBM @6.5TB/s)\n", M, N, K,
         kind == 1 ? "gelu" : kind == 2 ? "res " : "none", force_bn, pair, epi == 0 ? "generic" : "tma", boxes, fl / 1.5e9, bytes / 6.5e6);
  for (int mode = 0; mode < (pair ? 1 : 4); ++mode) {
    if (mode == 1 || (mode == 3 && (kind != 1 || epi == 0))) continue;
    ccx::gemm_timeline_mode(mode);
    for (int i = 0; i < 3; ++i) ccx::gemm_tn(g, 0);
    unsigned long long et[8];
    ccx::gemm_epi_timeline(et);
    const float us = time_once(g, 15);
    ccx::gemm_epi_timeline(et);
    unsigned long long tl[32];
    CK(cudaDeviceSynchronize());
    ccx::gemm_timeline_read(tl);
    printf("   epilogue %-22s %7.1f us %7.0f TF/s", mode == 0 ? "full" : mode == 1 ? "tcgen05.ld only" : mode == 2 ? "hand-over only" : "GELU -> x*x (no MUFU)", us, fl / us / 1e6);
    if (!pair)
      for (int c = 0; c < 2; ++c)
        printf(" | cta%s: prod wait-empty %llu/%llu  mma wait-full %llu wait-tmem %llu first %llu total %llu  epi wait-acc %llu/%llu",
               c ? "N" : "0", tl[c * 16 + 0], tl[c * 16 + 1], tl[c * 16 + 2], tl[c * 16 + 3], tl[c * 16 + 5], tl[c * 16 + 4],
               tl[c * 16 + 6], tl[c * 16 + 7]);
    if (!pair && epi != 0 && mode == 0)
      printf("\n        cta0 warp4 epilogue cycles per launch: params %llu  wait-box %llu  ld+math+sts %llu  fence %llu  store-issue %llu",
             et[0] / 15, et[1] / 15, et[2] / 15, et[3] / 15, et[4] / 15);
    printf("\n");
  }
  ccx::gemm_timeline_mode(0);
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(bias); cudaFree(cs);
}

int main() {
  CK(cudaMalloc(&g_flush, FLUSH_BYTES));
  const int shapes[][4] = {{8192, 2048, 512, 1},  {8192, 512, 2048, 2},  {131072, 512, 128, 1}, {131072, 128, 512, 2},
                           {32768, 1024, 256, 1}, {32768, 256, 1024, 2}, {2048, 4096, 1024, 1}, {2048, 1024, 4096, 2},
                           {8192, 8192, 8192, 0}};
  for (auto& s : shapes) {
    run(s[0], s[1], s[2], s[3], 0, 0, 0, 0);      // generic (transposing) epilogue
    run(s[0], s[1], s[2], s[3], 0, 0, -1, 1);     // TMA epilogue, 1 staging box per warp

    if (s[1] % 256 == 0 && s[0] >= 256) run(s[0], s[1], s[2], s[3], 0, 1, 0, 0);     // CTA pair, generic epilogue
    if (s[1] % 256 == 0 && s[0] >= 256) run(s[0], s[1], s[2], s[3], 0, 1, -1, 1);    // CTA pair, TMA epilogue
  }
  return 0;
}
