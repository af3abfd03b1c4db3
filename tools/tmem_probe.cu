// tmem_probe.cu — empirical answers for csrc/lstm_persist.cu:
//  (1) where do the rows of a cta_group::1, M=64 accumulator land in TMEM (which lane holds row r)?
//  (2) does tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st.32x32b, lane = row,
//      two bf16 per 32-bit column) compute the same product as the shared-memory form?
//  (3) time of 32 dependent-free MMAs (M=128 / M=64, N=32, K=16) with A from smem vs A from tmem.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I imagecaptioningconvnext_b200/csrc -o tools/tmem_probe tools/tmem_probe.cu
#include "ccx_common.cuh"
using namespace ccx;

__device__ __forceinline__ uint32_t sw_off(int r, int k) {   // [rows x 64 bf16] 128B-swizzled K-major tile
  return r * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4) + (k & 7) * 2;
}
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t mk_desc(const void* p) {
  return (uint64_t)(((smem_u32(p) & 0x3FFFF) >> 4) | (1u << 16)) | ((uint64_t)DESC_HI << 32);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }"
               ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
               "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                 "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
                 "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
                 "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}

// out: [0..127] lane -> value of column 0 after the M=64 MMA; [128..255] same for M=128 with A from TMEM;
//      [256..] timings
__global__ void __launch_bounds__(128, 1) probe(float* out, long long* tim) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* sa = sm;                 // A tile: 128 rows x 64 k (16 KB) x 8 chunks
  uint8_t* sb = sm + 8 * 16384;     // B tile: 32 rows x 64 k (4 KB) x 8 chunks
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (8 * 16384 + 8 * 4096) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  __syncthreads();
  // A[r][0] = r + 1 (bf16 exact up to 256), B[n][0] = 1
  *reinterpret_cast<__nv_bfloat16*>(sa + sw_off(tid, 0)) = __float2bfloat16_rn((float)(tid + 1));
  if (tid < 32) *reinterpret_cast<__nv_bfloat16*>(sb + sw_off(tid, 0)) = __float2bfloat16_rn(1.0f);
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  uint32_t ph = 0;
  // zero the accumulator columns 0..63 in all lanes
  { uint32_t z[32]; for (int i = 0; i < 32; ++i) z[i] = 0; tmem_st32(tm + ((uint32_t)(warp * 32) << 16), z);
    tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + 32, z);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  // (1) M=64, N=32
  if (tid == 0) {
    mma_f16_ss(tm, mk_desc(sa), mk_desc(sb), umma_idesc(1u, 64, 32), 0u);
    tc_commit(&bar);
  }
  mbar_wait(&bar, ph); ph ^= 1; tc_fence_after();
  { uint32_t v[32]; tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), v); tmem_ld_wait(); out[tid] = __uint_as_float(v[0]); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  // (2) A in TMEM at columns 64.. : lane = row, column c holds bf16 k = 2c, 2c+1.  A[r][0] = r + 1
  { uint32_t v[32]; for (int i = 0; i < 32; ++i) v[i] = 0;
    __nv_bfloat16 h = __float2bfloat16_rn((float)(tid + 1)); v[0] = (uint32_t)(*reinterpret_cast<unsigned short*>(&h));
    tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + 64, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (tid == 0) {
    mma_ts(tm + 32, tm + 64, mk_desc(sb), umma_idesc(1u, 128, 32), 0u);
    tc_commit(&bar);
  }
  mbar_wait(&bar, ph); ph ^= 1; tc_fence_after();
  { uint32_t v[32]; tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + 32, v); tmem_ld_wait(); out[128 + tid] = __uint_as_float(v[0]); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  // (3) timings: 32 MMAs over 8 chunks x 4 k-steps; NACC independent accumulators (32 columns each) used round-robin
  //     variant = mode * 4 + log2(NACC):  mode 0: A smem M=128, 1: A smem M=64, 2: A tmem M=128 (A at columns 256..)
  for (int variant = 0; variant < 12; ++variant) {
    const int mode = variant >> 2, nacc = 1 << (variant & 3);
    for (int rep = 0; rep < 2; ++rep) {
      long long t0 = 0;
      if (tid == 0) {
        const uint64_t ad0 = mk_desc(sa), bd0 = mk_desc(sb);
        t0 = clock64();
#pragma unroll 1
        for (int kc = 0; kc < 8; ++kc)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = kc * 4 + k;
            const uint32_t d = tm + (uint32_t)((i & (nacc - 1)) * 32);
            const uint32_t acc = i >= nacc ? 1u : 0u;
            const uint64_t ad = ad0 + (uint64_t)(kc * 1024 + 2 * k), bd = bd0 + (uint64_t)(kc * 256 + 2 * k);
            if (mode == 0) mma_f16_ss(d, ad, bd, umma_idesc(1u, 128, 32), acc);
            else if (mode == 1) mma_f16_ss(d, ad, bd, umma_idesc(1u, 64, 32), acc);
            else mma_ts(d, tm + 256 + i * 8, bd, umma_idesc(1u, 128, 32), acc);
          }
        tc_commit(&bar);
      }
      mbar_wait(&bar, ph); ph ^= 1; tc_fence_after();
      if (tid == 0) tim[variant * 2 + rep] = clock64() - t0;
      __syncthreads();
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  float* out; long long* tim;
  cudaMalloc(&out, 4096); cudaMalloc(&tim, 1024);
  cudaMemset(out, 0, 4096); cudaMemset(tim, 0, 1024);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 8 * 4096 + 1024);
  probe<<<1, 128, 8 * 16384 + 8 * 4096 + 1024>>>(out, tim);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  float h[256]; long long t[32];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(t, tim, sizeof(t), cudaMemcpyDeviceToHost);
  printf("M=64 accumulator: TMEM lane -> (row+1) of column 0\n");
  for (int i = 0; i < 128; ++i) printf("%s%3d:%3.0f", (i % 16) ? " " : "\n  ", i, h[i]);
  printf("\nA-from-TMEM, M=128: TMEM lane -> (row+1)\n");
  for (int i = 0; i < 128; ++i) printf("%s%3d:%3.0f", (i % 16) ? " " : "\n  ", i, h[128 + i]);
  const char* modes[3] = {"A smem M=128", "A smem M=64", "A tmem M=128"};
  printf("\n32 MMAs (N=32, K=16 each), cycles incl. commit + wait:\n");
  for (int v = 0; v < 12; ++v) printf("  %s, %d accumulators: %lld %lld\n", modes[v >> 2], 1 << (v & 3), t[2 * v], t[2 * v + 1]);
  return 0;
}
