// ccx_gemm.h — internal descriptor for the tcgen05 GEMM launcher (not part of the public C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace ccx {

struct GemmDesc {
  // C[M,N] = epi(A[M,K] . B[N,K]^T); A,B row-major with leading dims lda/ldb (elements)
  const void* A = nullptr;      // bf16, or fp32 (tf32 "hi" part when A_lo is given)
  const void* A_lo = nullptr;   // fp32 residual part for 3xTF32, or nullptr
  const void* B = nullptr;
  const void* B_lo = nullptr;
  void* C = nullptr;            // bf16 or fp32
  float* C_lo = nullptr;        // if split: C <- tf32_hi(y), C_lo <- y - hi
  const float* bias = nullptr;      // [N]
  const float* colscale = nullptr;  // [N]
  const float* rowscale = nullptr;  // [ceil(M / rows_per_group)]
  const void* residual = nullptr;   // [M, ldr], dtype of C
  const float* emask = nullptr;     // [M, ldm] element-wise multiplier after the activation
  long long lda = 0, ldb = 0, ldc = 0, ldr = 0, ldm = 0;
  int M = 0, N = 0, K = 0;
  int rows_per_group = 1;
  int act = 0;        // 0 none, 1 gelu(erf), 2 relu, 3 gelu'(x)
  int in_dtype = 1;   // CCX_F32 (tf32 path) / CCX_BF16
  int out_dtype = 1;  // CCX_F32 / CCX_BF16
  int split = 0;
  int force_bn = 0;   // 0 = auto, else 64/128/256
  bool a_mn = false;  // bf16: A given as [K, M] row-major (lda = row pitch), read MN-major by the tensor core
  bool b_mn = false;  // bf16: B given as [K, N] row-major (ldb = row pitch)
  bool res_mul = false;  // residual multiplies the activated result instead of being added
};

int gemm_tn(const GemmDesc& g, cudaStream_t stream);
// M <= 32 bf16 -> fp32 GEMMs (the recurrent ones): mma.sync kernel spread over N, see gemm_skinny.cu
bool gemm_skinny_eligible(const GemmDesc& g);
int gemm_skinny(const GemmDesc& g, cudaStream_t stream);
void set_gemm_pair_mode(int mode);

// driver entry point for building TMA descriptors (resolved through the runtime; no -lcuda needed)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_fn();
int num_sms();
void set_sm_limit(int n);   // 0 = no cap; see gemm_tcgen05.cu

}  // namespace ccx
