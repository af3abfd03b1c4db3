// gemm_tcgen05.cu — the one dense-contraction kernel of libccx (sm_100a only).
//
//   C[M,N] = epilogue( A[M,K] · B[N,K]^T )          (both operands K-major, "TN")
//
// * operands are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a
//   multi-stage shared-memory ring by one producer thread,
// * one thread issues tcgen05.mma (UMMA 128 x BN x 16, kind::f16 for bf16 inputs;
//   UMMA 128 x BN x 8, kind::tf32 for the fp32 path) with fp32 accumulators in TMEM,
// * four epilogue warps drain TMEM with tcgen05.ld and apply the fused epilogue
//   (bias, exact-erf GELU / ReLU, layer-scale x stochastic-depth row scale, residual)
//   while the MMA thread already works on the next tile (2 TMEM accumulator stages),
// * persistent CTAs, one per SM, static round-robin tile schedule.
//
// fp32 mode ("3xTF32"): A and B are each given as a (hi, lo) pair of fp32 arrays whose
// hi part is exactly representable in tf32; the K loop runs three segments
// (Alo·Bhi, Ahi·Blo, Ahi·Bhi) into the same accumulator, which restores ~2^-21 relative
// accuracy per product (needed for the fp32 parity bar: 1e-3 on features, exact argmax).
//
// Replaces, on the reference's hot path, every nn.Linear / 1x1-equivalent conv call:
//   torchvision/models/convnext.py:55-57 (CNBlock MLP), :146-151 (downsample conv),
//   models/decoder.py:19-21,50-54 (attention / init / f_beta / fc Linear layers),
//   models/transformerDecoder.py:84-85 + torch/nn/modules/transformer.py (QKV/FFN/out-proj).
#include <stdlib.h>

#include "ccx_common.cuh"
#include "ccx_gemm.h"
#include "ccx_gemm_epilogue.cuh"
#include "ccx_prof.h"

namespace ccx {

static constexpr int BM = 128;          // UMMA M (one TMEM lane per row)
static constexpr int ROW_BYTES = 128;   // one swizzle row = 64 bf16 or 32 tf32 elements
// warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4.. epilogue: EPI_CG warps per TMEM lane quarter, each draining every
// EPI_CG-th column unit.  Measured (tools/gemm_timeline.cu): 4 warps per quarter drain a tile no faster than 2 — a
// 128x256 GELU tile costs ~6.6 k cycles either way, of which the MUFU work is 25 % and the rest follows the L2
// traffic of the operand stream and the output (the tile's TMEM read floor is 2 k cycles) — and cost an operand stage.
static constexpr int EPI_CG = 2;
static constexpr int EPI_WARPS = 4 * EPI_CG;
static constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;

// Shared-memory plan.  BOXES = 4 KB staging boxes per epilogue warp (TMA epilogue; box 0 doubles as the transposition
// scratch of the generic epilogue): 2 lets a residual box load while the previous unit is still being stored, at the
// price of one operand stage on the 256-wide tile.  The dynamic shared-memory window starts 1024-byte aligned (the
// kernel has no static shared memory; checked at run time), so there is no alignment slack.
template <int BN, int BOXES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * ROW_BYTES;
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = EPI_WARPS * BOXES * EPI_UNIT_BYTES;
  static constexpr int BAR_BYTES = 512;                            // pipeline + accumulator + residual mbarriers, TMEM slot
  static constexpr int PARAM_BYTES = 2 * BN * 4;                   // bias, layer-scale of the current tile
  static constexpr int BUDGET = 232448;                            // 227 KB
  static constexpr int MAX_STAGES = (BUDGET - EPI_BYTES - BAR_BYTES - PARAM_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + PARAM_BYTES;
  static_assert(STAGES >= 3, "operand ring too short");
  static_assert(EPI_UNIT_BYTES == EPI_SCRATCH_FLOATS * 4, "box 0 doubles as the generic epilogue's scratch");
};



// Developer timeline (tools/gemm_timeline.cu builds this file with -DCCX_GEMM_TIMELINE): per-role cycle accounting of
// two CTAs, and switches that remove parts of the epilogue.  Compiled out of libccx.
#ifdef CCX_GEMM_TIMELINE
__device__ unsigned long long g_gemm_tl[2][16];
#define TL_DECL(name) long long name = 0
#define TL_T0(t) const long long t = clock64()
#define TL_ADD(acc, t) acc += clock64() - t
#define TL_PUT(slot, v) do { const int c_ = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x - 1 ? 1 : -1); \
                             if (c_ >= 0) g_gemm_tl[c_][slot] = (unsigned long long)(v); } while (0)
#else
#define TL_DECL(name)
#define TL_T0(t)
#define TL_ADD(acc, t)
#define TL_PUT(slot, v)
#endif

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

// IS_TF32: fp32 words in smem, kind::tf32, UMMA_K = 8; else bf16, kind::f16, UMMA_K = 16
template <int BN, bool IS_TF32, int BOXES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmB_hi,
               const __grid_constant__ CUtensorMap tmA_lo, const __grid_constant__ CUtensorMap tmB_lo,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               int M, int N, int K, int nseg, EpiArgs ep, int mn_flags) {
  // mn_flags bit 0 / 1: A / B is given transposed ([K, M] / [K, N] row-major) and read MN-major (bf16 only)
  using S = GemmSmem<BN, BOXES>;
  constexpr int STAGES = S::STAGES;
  constexpr int BK = IS_TF32 ? 32 : 64;  // elements per 128-byte row
  constexpr uint32_t IDESC = umma_idesc(IS_TF32 ? 2u : 1u, BM, BN);
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;

  extern __shared__ uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
    printf("ccx: gemm shared-memory window not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * S::A_BYTES;
  uint8_t* epi_boxes = smem + STAGES * S::STAGE_BYTES;                       // 1024-byte aligned: TMA swizzle atoms
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_boxes + S::EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint64_t* res_bar = bars + 2 * STAGES + 4;                                 // [EPI_WARPS][BOXES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + EPI_WARPS * BOXES);
  static_assert((2 * 8 + 4 + EPI_WARPS * 2) * 8 + 4 <= S::BAR_BYTES, "barrier block");
  static_assert(BN <= 128 * EPI_CG, "one epilogue thread per tile column loads the parameters");
  float* epi_params = reinterpret_cast<float*>(epi_boxes + S::EPI_BYTES + S::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BK - 1) / BK;
  const int k_iters = num_kb * nseg;
  // tiles of this CTA (the grid never exceeds the tile count), and whether its last tile's residual uses the ring
  const int my_tiles = (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  constexpr int RING_BPS = (S::B_BYTES / EPI_UNIT_BYTES) > 0 ? (S::B_BYTES / EPI_UNIT_BYTES) : 1;   // boxes per B stage
  const bool ring_res = ep.tma && ep.residual != nullptr && BN >= 64 &&
                        4 * (BN / ((ep.out_dtype == CCX_F32) ? 32 : 64)) <= STAGES * RING_BPS;

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
      mbar_init(&tempty_bar[i], EPI_WARPS);
    }
    for (int i = 0; i < EPI_WARPS * BOXES; ++i) mbar_init(&res_bar[i], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();          // PDL: barrier init / TMEM allocation / descriptor prefetch overlap the previous kernel's tail

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    TL_DECL(w_empty);
    TL_T0(t_begin);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      for (int it = 0; it < k_iters; ++it) {
        const int seg = it / num_kb, kb = it - seg * num_kb;
        // segment order: (Alo,Bhi), (Ahi,Blo), (Ahi,Bhi); single-segment runs use (hi,hi)
        const CUtensorMap* ta = (nseg > 1 && seg == 0) ? &tmA_lo : &tmA_hi;
        const CUtensorMap* tb = (nseg > 1 && seg == 1) ? &tmB_lo : &tmB_hi;
        TL_T0(t0);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        TL_ADD(w_empty, t0);
        mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
        if (mn_flags & 1) {                      // {64 MN, 64 k} boxes, 8 KB each
#pragma unroll
          for (int j = 0; j < BM / 64; ++j)
            tma_load_2d(smem_a + stage * S::A_BYTES + j * 8192, ta, &full_bar[stage], m_blk * BM + 64 * j, kb * BK);
        } else {
          tma_load_2d(smem_a + stage * S::A_BYTES, ta, &full_bar[stage], kb * BK, m_blk * BM);
        }
        if (mn_flags & 2) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(smem_b + stage * S::B_BYTES + j * 8192, tb, &full_bar[stage], n_blk * BN + 64 * j, kb * BK);
        } else {
          tma_load_2d(smem_b + stage * S::B_BYTES, tb, &full_bar[stage], kb * BK, n_blk * BN);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (ring_res) {
      // the residual of this CTA's last tile, queued behind its last operand loads (see EpiRing)
      const int last = blockIdx.x + (my_tiles - 1) * gridDim.x;
      const int m_blk = last / n_tiles, n_blk = last % n_tiles;
      const int ucols = (ep.out_dtype == CCX_F32) ? 32 : 64;
      const int nboxes = 4 * (BN / ucols);
      for (int b0 = 0; b0 < nboxes; b0 += RING_BPS) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], RING_BPS * EPI_UNIT_BYTES);
        for (int i = 0; i < RING_BPS; ++i) {
          const int b = b0 + i;
          tma_load_2d(smem_b + stage * S::B_BYTES + i * EPI_UNIT_BYTES, &tmR, &full_bar[stage],
                      n_blk * BN + (b >> 2) * ucols, m_blk * BM + (b & 3) * 32);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    TL_PUT(0, w_empty);
    TL_PUT(1, clock64() - t_begin);
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    TL_DECL(w_full);
    TL_DECL(w_tempty);
    TL_DECL(w_first);
    TL_T0(t_begin);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      TL_T0(t1);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      TL_ADD(w_tempty, t1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int it = 0; it < k_iters; ++it) {
        TL_T0(t0);
        mbar_wait(&full_bar[stage], phase);
        TL_ADD(w_full, t0);
#ifdef CCX_GEMM_TIMELINE
        if (tile == (int)blockIdx.x && it == 0) w_first = clock64() - t_begin;
#endif
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + stage * S::A_BYTES), b_addr = smem_u32(smem_b + stage * S::B_BYTES);
        // K-major: one UMMA K step = 32 bytes along the 128-byte row; MN-major: 16 k-rows = 2048 bytes
        const uint64_t adesc = (mn_flags & 1) ? umma_desc_mn_sw128(a_addr, 8192) : umma_desc_k_sw128(a_addr);
        const uint64_t bdesc = (mn_flags & 2) ? umma_desc_mn_sw128(b_addr, 8192) : umma_desc_k_sw128(b_addr);
        const uint32_t a_step = (mn_flags & 1) ? 128u : 2u, b_step = (mn_flags & 2) ? 128u : 2u;
        const uint32_t idesc = IDESC | ((mn_flags & 1) ? (1u << 15) : 0u) | ((mn_flags & 2) ? (1u << 16) : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (IS_TF32)
            mma_tf32_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (it | k) ? 1u : 0u);
          else
            mma_f16_ss(d_tmem, adesc + a_step * k, bdesc + b_step * k, idesc, (it | k) ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tc_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    TL_PUT(2, w_full);
    TL_PUT(3, w_tempty);
    TL_PUT(4, clock64() - t_begin);
    TL_PUT(5, w_first);
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int ew = warp & 3;          // TMEM lane quarter this warp may access (warp id % 4)
    const int half = (warp - 4) >> 2;  // column group: EPI_CG warps share a quarter and take every EPI_CG-th unit
    int acc = 0;
    uint32_t acc_phase = 0;
    TL_DECL(w_tfull);
    TL_T0(t_begin);
    EpiTmaState tma_state;
    EpiRing ring;
    {
      const long long iters = static_cast<long long>(my_tiles) * k_iters;
      ring.base = smem_b;
      ring.full_bar = full_bar;
      ring.stage0 = static_cast<int>(iters % STAGES);
      ring.phase0 = static_cast<uint32_t>((iters / STAGES) & 1);
      ring.stages = STAGES;
      ring.stage_bytes = S::B_BYTES;
      ring.boxes_per_stage = RING_BPS;
    }
    if (ep.tma && static_cast<int>(blockIdx.x) < num_tiles)
      epilogue_params_prefetch<BN>(ep, tma_state, blockIdx.x % n_tiles, (blockIdx.x / n_tiles) * BM + ew * 32 + lane,
                                   threadIdx.x - 128, M, N);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
#ifdef CCX_GEMM_TIMELINE
      {
        TL_T0(t0);
        mbar_wait(&tfull_bar[acc], acc_phase);     // the wait inside epilogue_tile then returns at once
        TL_ADD(w_tfull, t0);
      }
#endif
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
      uint8_t* my_boxes = epi_boxes + (warp - 4) * (BOXES * EPI_UNIT_BYTES);
      if (ep.tma) {
        const int next_tile = tile + gridDim.x;
        const int next_n = next_tile < num_tiles ? next_tile % n_tiles : -1;
        const int next_row0 = next_tile < num_tiles ? (next_tile / n_tiles) * BM + ew * 32 : 0;
        const EpiRing* rp = (ring_res && next_tile >= num_tiles) ? &ring : nullptr;              // this is the last tile
        const bool next_ring = ring_res && next_tile < num_tiles && next_tile + static_cast<int>(gridDim.x) >= num_tiles;
        if (ep.out_dtype == CCX_F32)
          epilogue_tile_tma<BN, BOXES, true, EPI_CG>(ep, &tmC, &tmR, taddr, half, m_blk * BM + ew * 32, lane,
                                                     threadIdx.x - 128, my_boxes, epi_params,
                                                     res_bar + (warp - 4) * BOXES, tma_state, n_blk, M, N,
                                                     &tfull_bar[acc], acc_phase, next_n, next_row0, rp, next_ring);
        else
          epilogue_tile_tma<BN, BOXES, false, EPI_CG>(ep, &tmC, &tmR, taddr, half, m_blk * BM + ew * 32, lane,
                                                      threadIdx.x - 128, my_boxes, epi_params,
                                                      res_bar + (warp - 4) * BOXES, tma_state, n_blk, M, N,
                                                      &tfull_bar[acc], acc_phase, next_n, next_row0, rp, next_ring);
      } else {
        epilogue_tile<BN, EPI_CG, EPI_CG == 2>(ep, taddr, half, m_blk * BM + ew * 32, lane, reinterpret_cast<float*>(my_boxes),
                                         n_blk, M, N, &tfull_bar[acc], acc_phase);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (ep.tma && lane == 0) tma_store_wait_read<0>();     // the boxes must outlive the stores that read them
#ifdef CCX_GEMM_TIMELINE
    if (blockIdx.x == 0 && threadIdx.x == 128)
      for (int i = 0; i < 5; ++i) g_epi_tl[i] += tma_state.tl[i];
#endif
    if (warp == 4 && lane == 0) {
      TL_PUT(6, w_tfull);
      TL_PUT(7, clock64() - t_begin);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
#ifdef CCX_GEMM_TIMELINE
void gemm_timeline_mode(int mode) { cudaMemcpyToSymbol(g_gemm_dbg_mode, &mode, sizeof(int)); }
void gemm_timeline_read(unsigned long long* out32) { cudaMemcpyFromSymbol(out32, g_gemm_tl, sizeof(g_gemm_tl)); }
void gemm_epi_timeline(unsigned long long* out8) {          // read and clear
  cudaMemcpyFromSymbol(out8, g_epi_tl, sizeof(g_epi_tl));
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_epi_tl, z, sizeof(z));
}
#endif
PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// MN-major operand: the array is [k_rows, mn_cols] bf16 row-major; box = {64 mn, 64 k}, 128-byte swizzle
static int make_map_mn(CUtensorMap* map, const void* ptr, long long k_rows, long long mn_cols, long long ld_elems) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return CCX_ERR_TMA;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld_elems * 2) & 15)) return CCX_ERR_SHAPE;
  cuuint64_t gdim[2] = {(cuuint64_t)mn_cols, (cuuint64_t)k_rows};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld_elems * 2)};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CCX_OK : CCX_ERR_TMA;
}

// 2-D row-major [rows, cols] tensor, box = [box_rows, 128 bytes], 128-byte swizzle
static int make_map_2d(CUtensorMap* map, const void* ptr, int is_f32, long long rows, long long cols,
                       long long ld_elems, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return CCX_ERR_TMA;
  const int es = is_f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld_elems * es) & 15)) return CCX_ERR_SHAPE;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld_elems * es)};
  cuuint32_t box[2] = {(cuuint32_t)(ROW_BYTES / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CCX_OK : CCX_ERR_TMA;
}

// Optional cap on the SMs the persistent GEMM grids occupy (0 = all).  A persistent kernel with one CTA per SM that
// owns the whole register file and shared memory cannot share an SM with an NCCL kernel: while a gradient all-reduce
// runs next to the backward pass, either NCCL waits for SMs or the last CTAs of every GEMM wait for NCCL.  Leaving a
// few SMs to the collective for that stretch lets both proceed (CapturedTrainStep sets it around the encoder backward).
static int g_sm_limit = 0;
void set_sm_limit(int n) { g_sm_limit = n > 0 ? n : 0; }

int num_sms() {
  static PerDevice<int> cache;
  int& n = cache.ref();
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

struct GemmMaps {
  CUtensorMap a_hi, b_hi, a_lo, b_lo, c, r;
};

template <int BN, bool IS_TF32, int BOXES>
static int launch(const GemmMaps& tm, int M, int N, int K, int nseg, const EpiArgs& ep, int mn_flags,
                  cudaStream_t stream) {
  using S = GemmSmem<BN, BOXES>;
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.ref();
  auto kfn = gemm_tn_kernel<BN, IS_TF32, BOXES>;
  if (!configured) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return CCX_ERR_CUDA;
    configured = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int grid = tiles < num_sms() ? tiles : num_sms();
  if (grid < 1) return CCX_OK;
  return launch_pdl(kfn, dim3(grid), dim3(NUM_THREADS), S::TOTAL, stream, tm.a_hi, tm.b_hi, tm.a_lo, tm.b_lo, tm.c, tm.r,
                    M, N, K, nseg, ep, mn_flags) == cudaSuccess
             ? CCX_OK : CCX_ERR_CUDA;
}

template <int BN, bool IS_TF32>
static int launch_boxes(const GemmMaps& tm, int M, int N, int K, int nseg, const EpiArgs& ep, int boxes, int mn_flags,
                        cudaStream_t stream) {
  // (two boxes per warp only fit next to a useful operand ring on the narrow tiles)
  if constexpr (BN <= 128) {
    if (boxes == 2) return launch<BN, IS_TF32, 2>(tm, M, N, K, nseg, ep, mn_flags, stream);
  }
  return launch<BN, IS_TF32, 1>(tm, M, N, K, nseg, ep, mn_flags, stream);
}

int gemm_tn_2cta(const CUtensorMap& a_hi, const CUtensorMap& b_hi, const CUtensorMap& a_lo, const CUtensorMap& b_lo,
                 const CUtensorMap& c, const CUtensorMap& r, int M, int N, int K, int nseg, const EpiArgs& ep, bool tf32,
                 cudaStream_t stream);

static int g_pair_mode = -1;
void set_gemm_pair_mode(int mode) { g_pair_mode = mode ? 1 : 0; }
// developer overrides (tools/gemm_timeline.cu): epilogue 0 = generic only, -1 = auto; staging boxes 0 = auto, 1, 2
static int g_epi_override = -1;
static int g_boxes_override = 0;
void set_gemm_epilogue_override(int epi, int boxes) { g_epi_override = epi; g_boxes_override = boxes; }

int gemm_tn(const GemmDesc& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return g.M == 0 ? CCX_OK : CCX_ERR_SHAPE;
  const bool tf32 = (g.in_dtype == CCX_F32);
  const int BK = tf32 ? 32 : 64;
  if (!tf32 && g.out_dtype == CCX_F32 && g.split) return CCX_ERR_DTYPE;
  if (tf32 && g.out_dtype != CCX_F32) return CCX_ERR_DTYPE;
  (void)BK;
  const int mn_flags = (g.a_mn ? 1 : 0) | (g.b_mn ? 2 : 0);
  if (mn_flags && (tf32 || g.A_lo != nullptr || g.B_lo != nullptr)) return CCX_ERR_DTYPE;
  if (gemm_skinny_eligible(g)) return gemm_skinny(g, stream);
  // tile-N choice: widest tile that still gives every SM work
  int bn = g.force_bn;
  if (bn == 0) {
    const long long mt = (g.M + BM - 1) / BM;
    if (g.N % 256 == 0 && mt * (g.N / 256) * 5 >= num_sms() * 4) bn = 256;   // >= 0.8 wave of 128x256 tiles
    else if (mt * ((g.N + 127) / 128) >= num_sms() / 2 || g.N <= 64) bn = 128;
    else bn = 64;
    if (g.N <= 64) bn = 64;
    // (measured: switching the 1.3-wave 128x256 tilings to 128x128 tiles changes nothing — 45.6 us either way for
    //  M=12544 N=512 K=2048 — these GEMMs are bound by the L2->SMEM operand stream, not by wave quantisation)
  }
  // CTA-pair (cta_group::2, 256x256 tiles) path for the GEMMs that fill the machine with pair tiles
  // measured on B200 (profiles/r01_spans_*): the pair kernel does not beat the single-CTA kernel on these shapes
  // (the GEMMs are epilogue / HBM bound, not operand-traffic bound), so it is opt-in: ccx_set_gemm_pair_mode(1)
  // or CCX_GEMM_2CTA=1
  if (g_pair_mode < 0) g_pair_mode = getenv("CCX_GEMM_2CTA") ? atoi(getenv("CCX_GEMM_2CTA")) : 0;
  const int pair_mode = g_pair_mode;
  const long long pair_tiles = ((g.M + 255LL) / 256) * (g.N / 256);
  const bool use_pair = pair_mode && !mn_flags && g.force_bn == 0 && (g.N % 256 == 0) && g.M >= 256 &&
                        pair_tiles >= num_sms() / 2;
  if (use_pair) bn = 128;   // B box = this CTA's half of the 256-row B tile
  GemmMaps tm;
  CUtensorMap &a_hi = tm.a_hi, &b_hi = tm.b_hi, &a_lo = tm.a_lo, &b_lo = tm.b_lo;
  int rc;
  if ((rc = g.a_mn ? make_map_mn(&a_hi, g.A, g.K, g.M, g.lda) : make_map_2d(&a_hi, g.A, tf32, g.M, g.K, g.lda, BM)))
    return rc;
  if ((rc = g.b_mn ? make_map_mn(&b_hi, g.B, g.K, g.N, g.ldb) : make_map_2d(&b_hi, g.B, tf32, g.N, g.K, g.ldb, bn)))
    return rc;
  int nseg = 1;
  if (tf32 && g.A_lo != nullptr && g.B_lo != nullptr) {
    nseg = 3;
    if ((rc = make_map_2d(&a_lo, g.A_lo, tf32, g.M, g.K, g.lda, BM))) return rc;
    if ((rc = make_map_2d(&b_lo, g.B_lo, tf32, g.N, g.K, g.ldb, bn))) return rc;
  } else {
    a_lo = a_hi;
    b_lo = b_hi;
  }
  // TMA epilogue: output (and residual) rows must start 16-byte aligned; the dropout-mask and hi/lo-split epilogues
  // stay on the generic path
  const bool out_f32 = (g.out_dtype == CCX_F32);
  const long long oes = out_f32 ? 4 : 2;
  auto tma_ok = [&](const void* p, long long ld) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((ld * oes) & 15) == 0;
  };
  static const int epi_mode = getenv("CCX_GEMM_EPI") ? atoi(getenv("CCX_GEMM_EPI")) : -1;   // 0 = generic only
  const bool use_tma = epi_mode != 0 && g_epi_override != 0 && g.emask == nullptr &&
                       !(g.split && g.C_lo != nullptr) && tma_ok(g.C, g.ldc) &&
                       (g.residual == nullptr || tma_ok(g.residual, g.ldr));
  tm.c = a_hi;
  tm.r = a_hi;
  if (use_tma) {
    if ((rc = make_map_2d(&tm.c, g.C, out_f32, g.M, g.N, g.ldc, 32))) return rc;
    if (g.residual != nullptr && (rc = make_map_2d(&tm.r, g.residual, out_f32, g.M, g.N, g.ldr, 32))) return rc;
  }
  // staging boxes per epilogue warp
  int boxes = 1;
  if (g_boxes_override > 0) boxes = g_boxes_override;
  ProfScope prof(PROF_GEMM, stream, 2.0 * g.M * (double)g.N * g.K);
  EpiArgs ep;
  ep.out = g.C;
  ep.out_lo = g.C_lo;
  ep.bias = g.bias;
  ep.colscale = g.colscale;
  ep.rowscale = g.rowscale;
  ep.residual = g.residual;
  ep.emask = g.emask;
  ep.ldc = g.ldc;
  ep.ldr = g.ldr;
  ep.ldm = g.ldm;
  ep.rows_per_group = g.rows_per_group > 0 ? g.rows_per_group : 1;
  ep.act = g.act;
  ep.out_dtype = g.out_dtype;
  ep.split = (g.split && g.C_lo != nullptr) ? 1 : 0;
  ep.tma = use_tma ? 1 : 0;
  ep.res_mul = g.res_mul ? 1 : 0;
  ep.fast = tf32 ? 0 : 1;
  if (use_pair) return gemm_tn_2cta(a_hi, b_hi, a_lo, b_lo, tm.c, tm.r, g.M, g.N, g.K, nseg, ep, tf32, stream);
  if (tf32) {
    if (bn == 256) return launch_boxes<256, true>(tm, g.M, g.N, g.K, nseg, ep, boxes, mn_flags, stream);
    if (bn == 128) return launch_boxes<128, true>(tm, g.M, g.N, g.K, nseg, ep, boxes, mn_flags, stream);
    return launch_boxes<64, true>(tm, g.M, g.N, g.K, nseg, ep, boxes, mn_flags, stream);
  } else {
    if (bn == 256) return launch_boxes<256, false>(tm, g.M, g.N, g.K, nseg, ep, boxes, mn_flags, stream);
    if (bn == 128) return launch_boxes<128, false>(tm, g.M, g.N, g.K, nseg, ep, boxes, mn_flags, stream);
    return launch_boxes<64, false>(tm, g.M, g.N, g.K, nseg, ep, boxes, mn_flags, stream);
  }
}

}  // namespace ccx
