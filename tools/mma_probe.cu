// mma_probe.cu — per-instruction cost of tcgen05.mma (M=128, K=16) as a function of N, and the throughput of legacy
// mma.sync.m16n8k16 bf16 with 4 / 8 / 16 warps per SM (4 independent accumulators per warp).
#include "ccx_common.cuh"
using namespace ccx;
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t mk_desc(const void* p) {
  return (uint64_t)(((smem_u32(p) & 0x3FFFF) >> 4) | (1u << 16)) | ((uint64_t)DESC_HI << 32);
}
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void __launch_bounds__(512, 1) probe(long long* tim, float* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (192 * 1024) / 4; i += 512) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  uint32_t ph = 0;
  const int Ns[5] = {32, 64, 128, 256, 16};
  for (int v = 0; v < 5; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      long long t0 = 0;
      if (tid == 0) {
        const uint64_t ad0 = mk_desc(sm), bd0 = mk_desc(sm + 64 * 1024);
        const uint32_t idesc = umma_idesc(1u, 128, Ns[v]);
        t0 = clock64();
#pragma unroll 1
        for (int kc = 0; kc < 4; ++kc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_f16_ss(tm, ad0 + (uint64_t)(kc * 1024 + 2 * k), bd0 + (uint64_t)(kc * 2048 + 2 * k), idesc, (kc | k) ? 1u : 0u);
        tc_commit(&bar);
      }
      mbar_wait(&bar, ph); ph ^= 1; tc_fence_after();
      if (tid == 0) tim[v * 2 + rep] = (clock64() - t0) / 16;
      __syncthreads();
    }
  }
  // legacy mma.sync throughput: nw warps x 64 iterations x 4 independent HMMAs
  const int nws[3] = {4, 8, 16};
  for (int v = 0; v < 3; ++v) {
    __syncthreads();
    long long t0 = clock64();
    float acc = 0.f;
    if (warp < nws[v]) {
      float d[4][4] = {};
      uint32_t a[4] = {0x3c003c00u + tid, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
#pragma unroll 4
      for (int i = 0; i < 64; ++i) {
        hmma(d[0], a, b); hmma(d[1], a, b); hmma(d[2], a, b); hmma(d[3], a, b);
      }
      acc = d[0][0] + d[1][1] + d[2][2] + d[3][3];
    }
    __syncthreads();
    if (tid == 0) tim[16 + v] = clock64() - t0;
    sink[tid] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}
int main() {
  long long* tim; float* sink;
  cudaMalloc(&tim, 1024); cudaMalloc(&sink, 4096); cudaMemset(tim, 0, 1024);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 193 * 1024);
  probe<<<1, 512, 193 * 1024>>>(tim, sink);
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  long long t[32]; cudaMemcpy(t, tim, sizeof(t), cudaMemcpyDeviceToHost);
  const int Ns[5] = {32, 64, 128, 256, 16};
  for (int v = 0; v < 5; ++v) printf("tcgen05.mma M=128 N=%d K=16: %lld / %lld cycles per instruction (16 back to back)\n", Ns[v], t[2 * v], t[2 * v + 1]);
  const int nws[3] = {4, 8, 16};
  for (int v = 0; v < 3; ++v)
    printf("mma.sync m16n8k16: %d warps x 256 HMMA: %lld cycles -> %.0f MAC/cycle/SM\n", nws[v], t[16 + v],
           nws[v] * 256.0 * 2048.0 / (double)t[16 + v]);
  return 0;
}
