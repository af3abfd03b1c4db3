// gemm_selftest.cu — standalone (no torch) correctness + timing probe for csrc/gemm_tcgen05.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../imagecaptioningconvnext_b200/csrc \
//        gemm_selftest.cu ../imagecaptioningconvnext_b200/csrc/gemm_tcgen05.cu -o gemm_selftest
// Developer tool only: the judged parity tests live in tests/ and go through the C ABI.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "ccx_gemm.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

__global__ void ref_gemm(const float* A, const float* B, const float* bias, const float* colscale,
                         const float* rowscale, const float* res, float* C, int M, int N, int K, int act,
                         int rpg) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N) return;
  double acc = 0;
  for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
  float y = (float)acc;
  if (bias) y += bias[n];
  if (act == 1) y = 0.5f * y * (1.0f + erff(y * 0.70710678f));
  if (act == 2) y = fmaxf(y, 0.f);
  if (colscale) y *= colscale[n] * (rowscale ? rowscale[m / rpg] : 1.f);
  if (res) y += res[(size_t)m * N + n];
  C[(size_t)m * N + n] = y;
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

static int run_case(int M, int N, int K, bool tf32, int act, bool scale_res, int force_bn) {
  std::vector<float> hA((size_t)M * K), hB((size_t)N * K), hbias(N), hcs(N), hrs((M + 63) / 64), hres((size_t)M * N);
  for (auto& v : hA) v = tf32 ? frand() : bf16r(frand());
  for (auto& v : hB) v = tf32 ? frand() * 0.1f : bf16r(frand() * 0.1f);
  for (auto& v : hbias) v = frand();
  for (auto& v : hcs) v = frand();
  for (auto& v : hrs) v = 1.f + frand();
  for (auto& v : hres) v = tf32 ? frand() : bf16r(frand());
  float *dA, *dB, *dbias, *dcs, *drs, *dres, *dCref, *dC, *dClo;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4));
  CK(cudaMalloc(&dbias, N * 4)); CK(cudaMalloc(&dcs, N * 4)); CK(cudaMalloc(&drs, hrs.size() * 4));
  CK(cudaMalloc(&dres, hres.size() * 4)); CK(cudaMalloc(&dCref, hres.size() * 4));
  CK(cudaMalloc(&dC, hres.size() * 4)); CK(cudaMalloc(&dClo, hres.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dcs, hcs.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(drs, hrs.data(), hrs.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dres, hres.data(), hres.size() * 4, cudaMemcpyHostToDevice));
  ref_gemm<<<dim3((N + 127) / 128, M), 128>>>(dA, dB, dbias, scale_res ? dcs : nullptr, scale_res ? drs : nullptr,
                                             scale_res ? dres : nullptr, dCref, M, N, K, act, 64);
  CK(cudaDeviceSynchronize());

  ccx::GemmDesc g;
  g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldc = N; g.ldr = N;
  g.bias = dbias; g.act = act; g.force_bn = force_bn; g.rows_per_group = 64;
  if (scale_res) { g.colscale = dcs; g.rowscale = drs; }
  std::vector<float> out((size_t)M * N), ref((size_t)M * N);
  void *a16 = nullptr, *b16 = nullptr, *r16 = nullptr, *c16 = nullptr;
  float *Ahi = nullptr, *Alo = nullptr, *Bhi = nullptr, *Blo = nullptr;
  if (!tf32) {
    std::vector<__nv_bfloat16> t(hA.size());
    for (size_t i = 0; i < hA.size(); ++i) t[i] = __float2bfloat16_rn(hA[i]);
    CK(cudaMalloc(&a16, t.size() * 2)); CK(cudaMemcpy(a16, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
    t.resize(hB.size());
    for (size_t i = 0; i < hB.size(); ++i) t[i] = __float2bfloat16_rn(hB[i]);
    CK(cudaMalloc(&b16, t.size() * 2)); CK(cudaMemcpy(b16, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
    t.resize(hres.size());
    for (size_t i = 0; i < hres.size(); ++i) t[i] = __float2bfloat16_rn(hres[i]);
    CK(cudaMalloc(&r16, t.size() * 2)); CK(cudaMemcpy(r16, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c16, t.size() * 2));
    g.A = a16; g.B = b16; g.C = c16; g.residual = scale_res ? r16 : nullptr;
    g.in_dtype = 1; g.out_dtype = 1;
  } else {
    auto split = [](const std::vector<float>& x, float** hi, float** lo) {
      std::vector<float> h(x.size()), l(x.size());
      for (size_t i = 0; i < x.size(); ++i) {
        unsigned u; memcpy(&u, &x[i], 4);
        u = (u + 0x1000u) & 0xffffe000u;  // round-to-nearest tf32
        float hv; memcpy(&hv, &u, 4);
        h[i] = hv; l[i] = x[i] - hv;
      }
      CK(cudaMalloc(hi, x.size() * 4)); CK(cudaMalloc(lo, x.size() * 4));
      CK(cudaMemcpy(*hi, h.data(), x.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(*lo, l.data(), x.size() * 4, cudaMemcpyHostToDevice));
    };
    split(hA, &Ahi, &Alo); split(hB, &Bhi, &Blo);
    g.A = Ahi; g.A_lo = Alo; g.B = Bhi; g.B_lo = Blo; g.C = dC; g.C_lo = dClo; g.split = 1;
    g.residual = scale_res ? dres : nullptr;
    g.in_dtype = 0; g.out_dtype = 0;
  }
  int rc = ccx::gemm_tn(g, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc != 0 || e != cudaSuccess) {
    printf("FAIL launch M=%d N=%d K=%d tf32=%d rc=%d cuda=%s\n", M, N, K, tf32, rc, cudaGetErrorString(e));
    return 1;
  }
  CK(cudaMemcpy(ref.data(), dCref, ref.size() * 4, cudaMemcpyDeviceToHost));
  if (!tf32) {
    std::vector<__nv_bfloat16> t(out.size());
    CK(cudaMemcpy(t.data(), c16, t.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < out.size(); ++i) out[i] = __bfloat162float(t[i]);
  } else {
    std::vector<float> lo(out.size());
    CK(cudaMemcpy(out.data(), dC, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lo.data(), dClo, out.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < out.size(); ++i) out[i] += lo[i];
  }
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < out.size(); ++i) {
    double d = fabs((double)out[i] - ref[i]);
    if (!(d <= maxerr)) maxerr = d;  // catches NaN
    if (fabs(ref[i]) > maxref) maxref = fabs(ref[i]);
  }
  double tol = tf32 ? 2e-5 : 1e-2;
  bool ok = (maxerr <= tol * maxref) && maxerr == maxerr;
  printf("%s M=%5d N=%5d K=%5d %s act=%d sr=%d bn=%3d maxerr=%.3e maxref=%.3e rel=%.2e\n", ok ? "ok  " : "FAIL",
         M, N, K, tf32 ? "tf32x3" : "bf16  ", act, scale_res, force_bn, maxerr, maxref, maxerr / maxref);
  cudaFree(dA); cudaFree(dB); cudaFree(dbias); cudaFree(dcs); cudaFree(drs); cudaFree(dres); cudaFree(dCref);
  cudaFree(dC); cudaFree(dClo); cudaFree(a16); cudaFree(b16); cudaFree(r16); cudaFree(c16);
  cudaFree(Ahi); cudaFree(Alo); cudaFree(Bhi); cudaFree(Blo);
  return ok ? 0 : 1;
}

static void bench_case(int M, int N, int K, bool tf32, int act, int force_bn) {
  size_t es = tf32 ? 4 : 2;
  void *A, *B, *C, *Alo = nullptr, *Blo = nullptr;
  float* bias;
  CK(cudaMalloc(&A, (size_t)M * K * es)); CK(cudaMalloc(&B, (size_t)N * K * es)); CK(cudaMalloc(&C, (size_t)M * N * es));
  CK(cudaMalloc(&bias, N * 4));
  CK(cudaMemset(A, 0, (size_t)M * K * es)); CK(cudaMemset(B, 0, (size_t)N * K * es)); CK(cudaMemset(bias, 0, N * 4));
  if (tf32) {
    CK(cudaMalloc(&Alo, (size_t)M * K * es)); CK(cudaMalloc(&Blo, (size_t)N * K * es));
    CK(cudaMemset(Alo, 0, (size_t)M * K * es)); CK(cudaMemset(Blo, 0, (size_t)N * K * es));
  }
  ccx::GemmDesc g;
  g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldc = N; g.A = A; g.B = B; g.C = C; g.A_lo = Alo; g.B_lo = Blo;
  g.bias = bias; g.act = act; g.in_dtype = tf32 ? 0 : 1; g.out_dtype = tf32 ? 0 : 1; g.force_bn = force_bn;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) ccx::gemm_tn(g, 0);
  cudaEventRecord(e0);
  const int iters = 20;
  for (int i = 0; i < iters; ++i) ccx::gemm_tn(g, 0);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double fl = 2.0 * M * N * K * (tf32 ? 3 : 1);
  printf("bench M=%6d N=%5d K=%5d %s act=%d bn=%3d: %.2f us  %.1f TFLOP/s (issued)  [%s]\n", M, N, K,
         tf32 ? "tf32x3" : "bf16  ", act, force_bn, ms * 1e3 / iters, fl * iters / (ms * 1e-3) / 1e12,
         cudaGetErrorString(e));
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(bias); cudaFree(Alo); cudaFree(Blo);
}

int main(int argc, char** argv) {
  int fails = 0;
  // smallest first: a descriptor bug shows up here without long waits
  fails += run_case(128, 128, 64, false, 0, false, 128);
  if (fails) { printf("first case failed; stopping early\n"); return 1; }
  fails += run_case(128, 64, 128, false, 0, false, 64);
  fails += run_case(128, 256, 256, false, 0, false, 256);
  fails += run_case(300, 200, 192, false, 0, false, 0);
  fails += run_case(1000, 512, 2048, false, 1, false, 0);
  fails += run_case(4096, 512, 128, false, 1, false, 0);
  fails += run_case(2048, 1024, 4096, false, 0, true, 0);
  fails += run_case(20000, 256, 512, false, 2, false, 256);
  fails += run_case(77, 9490, 512, false, 0, false, 0);
  fails += run_case(128, 128, 32, true, 0, false, 128);
  fails += run_case(333, 1024, 512, true, 1, false, 0);
  fails += run_case(2048, 512, 2048, true, 0, true, 0);
  fails += run_case(50, 9490, 512, true, 0, false, 0);
  // large enough for the CTA-pair kernel when CCX_GEMM_2CTA=1 (>= 74 pair tiles)
  fails += run_case(16384, 2048, 512, false, 1, false, 0);
  fails += run_case(16000, 512, 2048, false, 0, true, 0);
  fails += run_case(8192, 1024, 256, true, 0, true, 0);
  printf("selftest: %d failures\n", fails);
  if (argc > 1) {
    bench_case(16384, 2048, 512, false, 1, 256);
    bench_case(16384, 2048, 512, false, 0, 256);
    bench_case(16384, 2048, 512, false, 0, 128);
    bench_case(16384, 512, 2048, false, 0, 256);
    bench_case(16384, 512, 2048, false, 0, 128);
    bench_case(8192, 8192, 8192, false, 0, 256);
    bench_case(8192, 8192, 8192, false, 0, 128);
    bench_case(16384, 2048, 512, true, 1, 256);
  }
  return fails ? 1 : 0;
}
