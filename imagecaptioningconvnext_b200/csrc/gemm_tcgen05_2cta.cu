// gemm_tcgen05_2cta.cu — the CTA-pair (tcgen05 cta_group::2) variant of the dense contraction for the large
// encoder GEMMs:   C[M,N] = epilogue( A[M,K] · B[N,K]^T )
//
// A cluster of two CTAs (one TPC) owns a 256 x 256 output tile.  CTA r holds A rows [128r, 128r+128) and HALF of
// the B tile (rows [128r, 128r+128) of the 256), so each SM streams 32 KB per 64-wide K block instead of the 48 KB
// of the single-CTA 128x256 tile — the single-CTA kernel is bound by L2->SM operand traffic on the stage-3
// ConvNeXt GEMMs.  The leader CTA (rank 0) issues tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16);
// the tensor cores of both SMs read both CTAs' shared memory; each CTA's TMEM receives the accumulator rows of
// its own A half.  Both CTAs run TMA producers (their loads complete on the LEADER's full barrier) and 8 epilogue
// warps (draining their own TMEM; "accumulator drained" arrives on the leader's barrier across the cluster).
#include "ccx_common.cuh"
#include "ccx_gemm.h"
#include "ccx_gemm_epilogue.cuh"

namespace ccx {

static constexpr int BM2 = 128;             // rows per CTA (the pair covers 256)
static constexpr int ROW_BYTES2 = 128;
static constexpr int NUM_THREADS2 = 384;    // warp0 TMA, warp1 MMA (leader only), warp2 TMEM alloc, warps 4-11 epilogue
static constexpr int EPI_WARPS2 = 8;

template <int BN>
struct Gemm2Smem {
  static constexpr int A_BYTES = BM2 * ROW_BYTES2;          // 16 KB
  static constexpr int B_BYTES = (BN / 2) * ROW_BYTES2;     // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 6 : 8;
  static constexpr int EPI_BYTES = EPI_WARPS2 * EPI_UNIT_BYTES;  // one 4 KB staging box per epilogue warp (TMA epilogue;
                                                                 // doubles as the generic epilogue's transposition scratch)
  static constexpr int BAR_BYTES = 512;
  static constexpr int PARAM_BYTES = 2 * BN * 4;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + PARAM_BYTES;   // window is 1024-aligned
  static_assert(TOTAL <= 232448, "shared memory");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// TMA load whose completion is signalled on an mbarrier given as a shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on the mbarrier at the same offset in BOTH CTAs once all previously issued pair-MMAs retire
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  const uint16_t mask = 0x3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BN, bool IS_TF32>
__global__ void __launch_bounds__(NUM_THREADS2, 1)
gemm_tn_2cta_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmB_hi,
                    const __grid_constant__ CUtensorMap tmA_lo, const __grid_constant__ CUtensorMap tmB_lo,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, int M,
                    int N, int K, int nseg, EpiArgs ep) {
  using S = Gemm2Smem<BN>;
  constexpr int STAGES = S::STAGES;
  constexpr int BK = IS_TF32 ? 32 : 64;
  constexpr uint32_t IDESC = umma_idesc(IS_TF32 ? 2u : 1u, 2 * BM2, BN);   // UMMA M = 256 across the pair
  constexpr uint32_t TMEM_COLS = 2 * BN;                                    // two accumulator stages

  extern __shared__ uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
    printf("ccx: gemm shared-memory window not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * S::A_BYTES;
  uint8_t* epi_boxes = smem + STAGES * S::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_boxes + S::EPI_BYTES);
  uint64_t* full_bar = bars;                    // used on the leader: both CTAs' TMA bytes land here
  uint64_t* empty_bar = bars + STAGES;          // per CTA: the pair-MMA commit frees this CTA's stage
  uint64_t* tfull_bar = bars + 2 * STAGES;      // per CTA: accumulator stage complete
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // leader: both CTAs' epilogues drained the stage
  uint64_t* res_bar = bars + 2 * STAGES + 4;    // per epilogue warp: residual box landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + EPI_WARPS2);
  static_assert((2 * 8 + 4 + EPI_WARPS2) * 8 + 4 <= S::BAR_BYTES, "barrier block");
  float* epi_params = reinterpret_cast<float*>(epi_boxes + S::EPI_BYTES + S::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);

  const int m_super = (M + 2 * BM2 - 1) / (2 * BM2);
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_super * n_tiles;
  const int num_kb = (K + BK - 1) / BK;
  const int k_iters = num_kb * nseg;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (nseg > 1) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
    if (ep.tma) {
      tma_prefetch_desc(&tmC);
      if (ep.residual != nullptr) tma_prefetch_desc(&tmR);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EPI_WARPS2);
    }
    for (int i = 0; i < EPI_WARPS2; ++i) mbar_init(&res_bar[i], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();      // barriers of BOTH CTAs are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();          // PDL: everything above overlaps the previous kernel's tail

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int ms = tile / n_tiles, n_blk = tile % n_tiles;
      const int m_row0 = (ms * 2 + static_cast<int>(rank)) * BM2;
      const int n_row0 = n_blk * BN + static_cast<int>(rank) * (BN / 2);
      for (int it = 0; it < k_iters; ++it) {
        const int seg = it / num_kb, kb = it - seg * num_kb;
        const CUtensorMap* ta = (nseg > 1 && seg == 0) ? &tmA_lo : &tmA_hi;
        const CUtensorMap* tb = (nseg > 1 && seg == 1) ? &tmB_lo : &tmB_hi;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);   // this CTA's and the peer's bytes
        const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
        tma_load_2d_pair(smem_a + stage * S::A_BYTES, ta, fb, kb * BK, m_row0);
        tma_load_2d_pair(smem_b + stage * S::B_BYTES, tb, fb, kb * BK, n_row0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int it = 0; it < k_iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * S::A_BYTES));
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * S::B_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (IS_TF32)
            mma_tf32_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (it | k) ? 1u : 0u);
          else
            mma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (it | k) ? 1u : 0u);
        }
        tc_commit_pair(&empty_bar[stage]);   // frees the stage in BOTH CTAs
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tc_commit_pair(&tfull_bar[acc]);       // accumulator complete in BOTH CTAs
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (both CTAs, own TMEM) =====================
    const int ew = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    EpiTmaState tma_state;
    const int tid_e = threadIdx.x - 128;
    auto tile_row0 = [&](int t) { return ((t / n_tiles) * 2 + static_cast<int>(rank)) * BM2 + ew * 32; };
    if (ep.tma && pair < num_tiles)
      epilogue_params_prefetch<BN>(ep, tma_state, pair % n_tiles, tile_row0(pair) + lane, tid_e, M, N);
    uint8_t* my_box = epi_boxes + (warp - 4) * EPI_UNIT_BYTES;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int n_blk = tile % n_tiles;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
      if (ep.tma) {
        const int next_tile = tile + num_pairs;
        const int next_n = next_tile < num_tiles ? next_tile % n_tiles : -1;
        const int next_row0 = next_tile < num_tiles ? tile_row0(next_tile) : 0;
        if (ep.out_dtype == CCX_F32)
          epilogue_tile_tma<BN, 1, true, 2>(ep, &tmC, &tmR, taddr, half, tile_row0(tile), lane, tid_e, my_box, epi_params,
                                            res_bar + (warp - 4), tma_state, n_blk, M, N, &tfull_bar[acc], acc_phase,
                                            next_n, next_row0);
        else
          epilogue_tile_tma<BN, 1, false, 2>(ep, &tmC, &tmR, taddr, half, tile_row0(tile), lane, tid_e, my_box, epi_params,
                                             res_bar + (warp - 4), tma_state, n_blk, M, N, &tfull_bar[acc], acc_phase,
                                             next_n, next_row0);
      } else {
        epilogue_tile<BN>(ep, taddr, half, tile_row0(tile), lane, reinterpret_cast<float*>(my_box), n_blk, M, N,
                          &tfull_bar[acc], acc_phase);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (ep.tma && lane == 0) tma_store_wait_read<0>();     // the boxes must outlive the stores that read them
  }

  tc_fence_before();
  cluster_sync_all();      // nobody leaves (or frees TMEM) while the peer can still signal / read this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

template <int BN, bool IS_TF32>
static int launch2(const CUtensorMap& a_hi, const CUtensorMap& b_hi, const CUtensorMap& a_lo, const CUtensorMap& b_lo,
                   const CUtensorMap& c, const CUtensorMap& r, int M, int N, int K, int nseg, const EpiArgs& ep,
                   cudaStream_t stream) {
  using S = Gemm2Smem<BN>;
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.ref();
  auto kfn = gemm_tn_2cta_kernel<BN, IS_TF32>;
  if (!configured) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return CCX_ERR_CUDA;
    configured = true;
  }
  const int tiles = ((M + 2 * BM2 - 1) / (2 * BM2)) * ((N + BN - 1) / BN);
  const int max_pairs = num_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  if (pairs < 1) return CCX_OK;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(NUM_THREADS2, 1, 1);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // the kernel calls grid_dep_sync()
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (cudaLaunchKernelEx(&cfg, kfn, a_hi, b_hi, a_lo, b_lo, c, r, M, N, K, nseg, ep) != cudaSuccess) return CCX_ERR_CUDA;
  return CCX_OK;
}

// Called by gemm_tn() for shapes that fill the machine with 256x256 pair tiles.  Maps: A box 128 rows, B box 128 rows.
int gemm_tn_2cta(const CUtensorMap& a_hi, const CUtensorMap& b_hi, const CUtensorMap& a_lo, const CUtensorMap& b_lo,
                 const CUtensorMap& c, const CUtensorMap& r, int M, int N, int K, int nseg, const EpiArgs& ep, bool tf32,
                 cudaStream_t stream) {
  if (tf32) return launch2<256, true>(a_hi, b_hi, a_lo, b_lo, c, r, M, N, K, nseg, ep, stream);
  return launch2<256, false>(a_hi, b_hi, a_lo, b_lo, c, r, M, N, K, nseg, ep, stream);
}

}  // namespace ccx
