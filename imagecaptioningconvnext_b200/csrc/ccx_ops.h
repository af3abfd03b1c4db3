// ccx_ops.h — internal launcher prototypes (C++ side of libccx; the public C ABI is include/ccx.h).
#pragma once
#include <cuda_runtime.h>

namespace ccx {

// dwconv_ln.cu
int dwconv7_ln(const float* x, const float* w49c, const float* bias, const float* gamma, const float* beta,
               void* out, float* out_lo, int B, int H, int W, int C, float eps, int out_dtype,
               cudaStream_t stream);

// encoder_misc.cu
int stem_ln(const float* img, const float* wk, const float* bias, const float* gamma, const float* beta,
            float* out, int B, int Hin, int Win, float eps, cudaStream_t stream);
int ln_rows(const float* x, const float* gamma, const float* beta, void* out, float* out_lo, long long M, int C,
            float eps, int out_dtype, int merge, int H, int W, cudaStream_t stream);
int avgpool_nhwc(const float* x, float* out, int B, int H, int W, int C, int S, cudaStream_t stream);
int split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t stream);
int cast_bf16(const float* x, void* y, long long n, cudaStream_t stream);

}  // namespace ccx
