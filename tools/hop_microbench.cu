// hop_microbench.cu — what does one hand-over between CTAs cost on B200?  (design input for csrc/lstm_persist.cu)
// Measures, in SM cycles on one CTA (clock64), with the source data L2-resident and written by ANOTHER launch:
//   * cp.async.bulk global->shared of 4 / 16 / 32 / 64 KB (one op) until the mbarrier completes
//   * the same bytes fetched with ld.global.cg.v4 by 128 / 256 / 512 threads + st.shared + __syncthreads
//   * fence.proxy.async, fence.acq_rel.gpu (after 1 outstanding store), red.release.gpu
//   * ping-pong between two CTAs through a global counter (round trip / 2 = one hop)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hop_microbench tools/hop_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0,1,0,P;}"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void fill(uint4* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = make_uint4(i, i + 1, i + 2, i + 3);
}

// out[0..] cycles
__global__ void __launch_bounds__(512, 1) bench(const uint8_t* src, long long* out, int* flag) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ long long t_end;
  const int tid = threadIdx.x;
  if (blockIdx.x == 0) {
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    int k = 0;
    uint32_t ph = 0;
    const int sizes[4] = {4096, 16384, 32768, 65536};
    // (a) bulk copies; each from a fresh region so nothing is in L1/smem
    for (int rep = 0; rep < 2; ++rep)
      for (int s = 0; s < 4; ++s) {
        if (tid == 0) {
          const uint8_t* p = src + (size_t)(rep * 4 + s) * 65536;
          long long t0 = clock64();
          mbar_expect(&bar, sizes[s]);
          bulk(sm, p, sizes[s], &bar);
          while (!mbar_try(&bar, ph)) {}
          out[k] = clock64() - t0;
        }
        ph ^= 1; ++k;
        __syncthreads();
      }
    // (b) LDG copies with nthreads = 128, 256, 512
    const int nts[3] = {128, 256, 512};
    for (int rep = 0; rep < 2; ++rep)
      for (int n = 0; n < 3; ++n)
        for (int s = 1; s < 4; ++s) {
          const uint8_t* p = src + (size_t)(8 + rep * 9 + n * 3 + s) * 65536;
          __syncthreads();
          long long t0 = clock64();
          if (tid < nts[n]) {
            const int per = sizes[s] / 16 / nts[n];
            uint4 v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < per) v[i] = __ldcg(reinterpret_cast<const uint4*>(p) + tid + i * nts[n]);
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < per) reinterpret_cast<uint4*>(sm)[tid + i * nts[n]] = v[i];
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncthreads();
          if (tid == 0) out[k] = clock64() - t0;
          ++k;
        }
    // (c) fences
    if (tid == 0) {
      long long t0 = clock64();
      asm volatile("fence.proxy.async;" ::: "memory");
      out[k] = clock64() - t0;
      out[100] = 1;
      t0 = clock64();
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      out[k + 1] = clock64() - t0;
      out[101] = 2;
      t0 = clock64();
      asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(flag + 8), "r"(1) : "memory");
      out[k + 2] = clock64() - t0;
      t0 = clock64();
      int v;
      asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag + 8) : "memory");
      out[k + 3] = clock64() - t0 + (v & 0);
      t0 = clock64();
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag + 8) : "memory");
      out[k + 4] = clock64() - t0 + (v & 0);
    }
    k += 5;
    __syncthreads();
  }
  // (d) ping-pong between CTA 0 and CTA 1, 200 round trips
  if (tid == 0 && blockIdx.x < 2) {
    const int me = blockIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < 200; ++i) {
      if (me == 0) {
        asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
        int v;
        do { asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag + 1) : "memory"); } while (v <= i);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      } else {
        int v;
        do { asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (v <= i);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(flag + 1), "r"(1) : "memory");
      }
    }
    if (me == 0) out[60] = (clock64() - t0) / 200;
  }
}

int main() {
  uint8_t* src; long long* out; int* flag;
  cudaMalloc(&src, 64 << 20); cudaMalloc(&out, 1024 * 8); cudaMalloc(&flag, 1024);
  cudaMemset(out, 0, 1024 * 8); cudaMemset(flag, 0, 1024);
  fill<<<(4 << 20) / 256, 256>>>((uint4*)src, 4 << 20);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaDeviceSynchronize();
  void* args[] = {&src, &out, &flag};
  cudaLaunchCooperativeKernel((void*)bench, dim3(2), dim3(512), args, 100 * 1024, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  long long h[128];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const char* sz[4] = {"4K", "16K", "32K", "64K"};
  int k = 0;
  for (int rep = 0; rep < 2; ++rep) for (int s = 0; s < 4; ++s) printf("bulk %s rep%d: %lld cyc\n", sz[s], rep, h[k++]);
  const int nts[3] = {128, 256, 512};
  for (int rep = 0; rep < 2; ++rep) for (int n = 0; n < 3; ++n) for (int s = 1; s < 4; ++s)
    printf("ldg %s x%d threads rep%d: %lld cyc\n", sz[s], nts[n], rep, h[k++]);
  printf("fence.proxy.async %lld, fence.acq_rel.gpu %lld, red.release.gpu %lld, ld.relaxed %lld, ld.acquire %lld\n",
         h[k], h[k + 1], h[k + 2], h[k + 3], h[k + 4]);
  printf("ping-pong round trip %lld cyc (one hop = half)\n", h[60]);
  return 0;
}
