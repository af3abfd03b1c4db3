// lstm_persist.cu — the teacher-forced LSTM-attention recurrence (models/decoder.py:96-111) as ONE persistent
// cooperative kernel per direction instead of 4-5 dependent launches per time step (~460 launches per train step).
//
// Forward (lstm_tf_fwd_persist_kernel), bf16 operands, fp32 accumulation, batch <= 32 rows:
//   * all recurrent weights stay resident in shared memory for the whole caption (7.9 MB bf16 spread over the CTAs),
//   * every GEMM is swap-AB on the tensor cores: the WEIGHT rows are the UMMA M side (tcgen05.mma M=128), the batch
//     is N=32, the fp32 accumulator lives in TMEM; the batch-side operand (h_{t-1}, gated context) arrives by TMA,
//   * three CTA roles hand work to each other through release/acquire counters in global memory (one hop ~ an L2
//     round trip) — no grid-wide barrier, no kernel boundary:
//       G1  (12 CTAs)  [att2 | gate pre-activation] = [decoder_att ; f_beta](h_{t-1})        models/decoder.py:27,104
//       ATT (B CTAs)   relu(att1+att2).w_f -> softmax over pixels -> sum_p alpha_p enc_p -> * sigmoid(gate)
//                      with att1[b], enc[b] (bf16) resident in the CTA's shared memory     models/decoder.py:28-30,105
//       G2  (64 CTAs)  gates = [h_{t-1} | gated awe].[W_hh | W_ih(awe)]^T + (hoisted emb_t.W_ih(emb)^T + b), epilogue =
//                      the LSTMCell gate math (sigmoid/sigmoid/tanh/sigmoid, c', h') for the 8 hidden units x 4 gates
//                      whose weight rows the CTA owns; c stays in registers across steps    models/decoder.py:106-108
//   * per step the critical path is G2(t-1) -> G1 -> ATT -> G2(t): three hops instead of 4-5 kernel launches.
// Writes exactly the buffers the per-step path (lstm_runner.cu) writes, so BPTT and the hoisted fc GEMM are unchanged.
//
// Backward (lstm_tf_bwd_persist_kernel): see the comment above that kernel.
#include "../../include/ccx.h"

#include "ccx_common.cuh"
#include "ccx_gemm.h"
#include "ccx_ops.h"
#include "ccx_prof.h"

namespace ccx {
namespace lp {

constexpr int NB = 32;              // batch slots = UMMA N
constexpr int D = 512, A = 512, E = 1024, EMB = 512;
constexpr int KX = EMB + E + D;     // 2048: row of XH = [emb | gated awe | h]
constexpr int HOFF = EMB + E;       // 1536
constexpr int AE = A + E;           // 1536: row of HG = [att2 | gate pre-activation]
constexpr int G4 = 4 * D;           // 2048
constexpr int G1_KSPLIT = 4;        // role G1: 12 row tiles x 4 K quarters
constexpr int N_G1 = (AE / 128) * G1_KSPLIT;   // 48 CTAs, 128 weight rows x 128 of K each
constexpr int N_G2 = G4 / 32;       // 64 CTAs, 32 (permuted) gate rows each = 8 hidden units x 4 gates
constexpr int THREADS = 512;
constexpr int CH128 = 128 * 128;    // bytes of a [128 rows x 64 bf16] swizzled K chunk
constexpr int CH32 = 32 * 128;      // bytes of a [32 rows x 64 bf16] swizzled K chunk
constexpr int MAX_P = 64;
constexpr uint32_t IDESC = umma_idesc(1u, 128, NB);
constexpr int SMEM_BYTES = 208 * 1024;

struct FwdArgs {
  __nv_bfloat16* XH;          // [T+1][B][KX] row-major operand buffer (what BPTT's weight-gradient GEMMs read)
  const float* E_all;         // [T][B][G4] hoisted emb part of the gates (+ both biases), PERMUTED column order
  float* HG;                  // [T][B][AE]
  float* G;                   // [T][B][G4] gate pre-activations, torch column order (i|f|g|o)
  float* C_all;               // [T+1][B][D]
  __nv_bfloat16* H_all;       // [B][T][D]  h_t * dropout multiplier (operand of the hoisted fc GEMM)
  float* alphas;              // [B][T][P]
  const float* dropmask;      // [B][T][D] or nullptr
  const __nv_bfloat16* att1;  // [B][P][A]
  const __nv_bfloat16* enc;   // [B][P][E]
  const float* b_h;           // [AE]
  const float* w_f;           // [A]
  const float* b_f;           // [1] or nullptr
  const long long* decode_len;  // [B] sorted descending
  uint8_t* HB;                // [T+1][32 KB] h_{t-1} as the UMMA B-operand image (8 swizzled [32 x 64] bf16 chunks)
  uint8_t* GA;                // [T][64 KB]   gated awe of step t as operand image (16 chunks)
  float* HGp;                 // [2][G1_KSPLIT][B][AE] K-quarter partial sums of role G1, double-buffered by step parity
  int* cnt_h;                 // [T+1] h_{t-1} complete (N_G2 arrivals)
  int* cnt_hg;                // [T]   HG[t] complete (N_G1 arrivals)
  int* cnt_awe;               // [T]   gated awe of step t complete (bt arrivals)
  float* awe_all;             // [T][B][E] un-gated context vectors (kept for the backward kernel) or nullptr
  long long* dbg;             // optional [3 roles][T][8] clock64 stamps of CTA 0 of each role
  int B, T, P;
};

// Operand images: a consumer fetches its whole batch-side operand with ONE contiguous bulk copy (a tiled TMA load of
// the same bytes costs ~0.2 us per 4 KB box, serialised).  The producers therefore store each element where the
// 128-byte-swizzled K-major UMMA layout wants it: chunk k/64 (4 KB = 32 rows x 128 B), row r, 16-byte unit
// ((k%64)/8) ^ (r%8).  Byte offset of element (row r, column k):
__device__ __forceinline__ uint32_t img_off(int r, int k) {
  return (static_cast<uint32_t>(k >> 6) << 12) + (static_cast<uint32_t>(r) << 7) +
         (static_cast<uint32_t>(((k & 63) >> 3) ^ (r & 7)) << 4) + (static_cast<uint32_t>(k & 7) << 1);
}

__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// release-increment: orders every write this thread has observed (its own and, through the preceding CTA barrier,
// those of the other threads of the CTA) before the increment, at GPU scope
__device__ __forceinline__ void signal_counter(int* p) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(1) : "memory");
}
// optional in-kernel timeline: clock64 stamps go to shared memory (a global store per stamp would sit in front of the
// next release fence and distort what is being measured) and are flushed when the role ends
constexpr int DBG_OFF = 208 * 1024 - 64 * 8 * 8;   // [64 steps][8] long long at the top of the dynamic smem window
__device__ __forceinline__ void dbg_stamp(long long* dbg, uint8_t* smem, int t, int k) {
  if (dbg != nullptr && t < 64) reinterpret_cast<long long*>(smem + DBG_OFF)[t * 8 + k] = clock64();
}
__device__ __forceinline__ void dbg_flush(long long* dbg, uint8_t* smem, int role, int T) {
  if (dbg == nullptr) return;
  const long long* s = reinterpret_cast<const long long*>(smem + DBG_OFF);
  for (int i = 0; i < (T < 64 ? T : 64) * 8; ++i) dbg[static_cast<long long>(role) * T * 8 + i] = s[i];
}
// relaxed polling + one acquire fence; bounded: a protocol bug must surface as a trap, never as a hung GPU
__device__ __forceinline__ void wait_counter(const int* p, int target, int tag) {
  if (target <= 0) return;
  if (ld_relaxed_gpu(p) < target) {
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (ld_relaxed_gpu(p) < target) {
      if ((++spins & 0xff) == 0 && (clock64() - t0) > 3000000000LL) {
        printf("ccx lstm_persist: counter timeout tag %d block %d have %d want %d\n", tag, blockIdx.x,
               ld_relaxed_gpu(p), target);
        __trap();
      }
    }
  }
  [[maybe_unused]] int v;      // the acquire that pairs with the producers' red.release (cheaper than a full fence.acq_rel.gpu)
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
}
// Time steps the recurrence really has: T is the extent of the step buffers (the longest caption the CALLER allows for,
// e.g. always 51 under CUDA-graph replay), the rows are sorted by length, so row 0 holds the longest decode length.
__device__ __forceinline__ int steps_to_run(int T, const long long* decode_len) {
  const long long d0 = decode_len[0];
  return d0 < T ? static_cast<int>(d0 < 0 ? 0 : d0) : T;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// tcgen05.mma issue for the recurrent GEMMs.  One thread issues K/16 instructions of only 16 tensor-pipe cycles each
// (M=128, N=32), so the ISSUE cost decides: building each 64-bit descriptor with shifts and ors in a dependent chain
// measured ~50 cycles per MMA.  Here the low descriptor word of chunk 0 is computed once; every further descriptor is
// that word plus a compile-time constant (addresses stay below 256 KB: no carry out of the 14-bit field) and the
// high word is a constant, so an MMA costs two integer adds.
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(const void* smem_ptr) {
  return ((smem_u32(smem_ptr) & 0x3FFFFu) >> 4) | (1u << 16);
}
template <int ACC>
__device__ __forceinline__ void mma_lo(uint32_t tmem, uint32_t a_lo, uint32_t b_lo) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .pred p;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "}\n" ::"r"(tmem),
      "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(IDESC), "n"(ACC)
      : "memory");
}
// chunks [K0, K1) of an operand pair: A chunk stride A_STRIDE bytes (M = 128 window), B chunk stride 4 KB (N = 32)
template <int K0, int K1, int A_STRIDE, bool FRESH>
__device__ __forceinline__ void mma_chunks(uint32_t tmem, uint32_t a_lo0, uint32_t b_lo0) {
#pragma unroll
  for (int kc = K0; kc < K1; ++kc) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t al = a_lo0 + static_cast<uint32_t>(kc * (A_STRIDE >> 4) + 2 * k);
      const uint32_t bl = b_lo0 + static_cast<uint32_t>(kc * (CH32 >> 4) + 2 * k);
      if (FRESH && kc == K0 && k == 0) mma_lo<0>(tmem, al, bl);
      else mma_lo<1>(tmem, al, bl);
    }
  }
}

// Legacy warp-level tensor-core path for the SKINNY tiles (32 / 16 weight rows per CTA).  Measured on B200
// (tools/mma_probe.cu): tcgen05.mma costs a flat ~67 cycles per instruction for N <= 64 whatever the useful rows, i.e.
// K/16 x 67 cycles per CTA and step (64 MMAs = 2.3 us for role G2's context part), while mma.sync.m16n8k16 sustains
// ~960 MAC/cycle/SM with 8 warps: 4x (G2) to 8x (HP) faster on these shapes.  Operands are read with ldmatrix from
// the same 128-byte-swizzled K-major chunks the TMA / bulk copies deliver (the swizzle makes ldmatrix conflict-free).
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// One 64-wide K chunk: acc[MT][4][4] += W[16*MT rows] . X[32 batch rows]^T.  a_chunk / b_chunk: shared addresses of the
// swizzled [rows x 64] chunks.  Lane roles for ldmatrix.x4: matrix = lane / 8, row in matrix = lane % 8.
template <int MT>
__device__ __forceinline__ void hmma_chunk(float (&acc)[MT][4][4], uint32_t a_chunk, uint32_t b_chunk, int lane) {
  const int mi = lane >> 3, ri = lane & 7;
  // A (m16 x k16): matrices (rows 0-7, k lo), (rows 8-15, k lo), (rows 0-7, k hi), (rows 8-15, k hi)
  const uint32_t a_row = static_cast<uint32_t>((mi & 1) * 8 + ri), a_u = static_cast<uint32_t>(mi >> 1);
  // B (two n8 tiles x k16): matrices (n 0-7, k lo), (n 0-7, k hi), (n 8-15, k lo), (n 8-15, k hi)
  const uint32_t b_row = static_cast<uint32_t>((mi >> 1) * 8 + ri), b_u = static_cast<uint32_t>(mi & 1);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t bf[2][4];
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      const uint32_t row = b_row + 16 * np;
      ldsm4(bf[np], b_chunk + row * 128 + (((2 * ks + b_u) ^ (row & 7)) << 4));
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      uint32_t af[4];
      const uint32_t row = a_row + 16 * mt;
      ldsm4(af, a_chunk + row * 128 + (((2 * ks + a_u) ^ (row & 7)) << 4));
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        hmma(acc[mt][2 * np], af, bf[np][0], bf[np][1]);
        hmma(acc[mt][2 * np + 1], af, bf[np][2], bf[np][3]);
      }
    }
  }
}
// Partial-sum scratch of one warp: [rows][32 batch columns] fp32 with the columns rotated by 8 * (row % 4) so that the
// accumulator fragments (8 rows x 4 column pairs per store) spread over all banks.
__device__ __forceinline__ int rot_col(int row, int col) { return (col + 8 * (row & 3)) & 31; }
template <int MT>
__device__ __forceinline__ void store_partials(float* sp, const float (&acc)[MT][4][4], int lane) {
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int r0 = 16 * mt + g, c = 8 * nt + 2 * tq;
      *reinterpret_cast<float2*>(sp + r0 * 32 + rot_col(r0, c)) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
      *reinterpret_cast<float2*>(sp + (r0 + 8) * 32 + rot_col(r0 + 8, c)) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
    }
}

// sigmoid / tanh through ex2.approx + fast division (|error| ~ 1e-6): the gate math sits on the recurrence's critical path
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

// ---------------------------------------------------------------------------------------------------------------
// role G1 (tcgen05): rows [128*rt, +128) of [decoder_att ; f_beta] (AE x D) over K quarter kq (128 of D = 512):
// 8 MMAs per step instead of 32 — a tcgen05.mma costs ~67 cycles whatever N <= 64 is, so the K extent per CTA is what
// sits on the critical path.  The consumer (role ATT) adds the four quarter sums in a fixed order.
// ---------------------------------------------------------------------------------------------------------------
__device__ void fwd_role_g1(const FwdArgs& a, const CUtensorMap* tm_wh, int idx, uint8_t* smem) {
  if (threadIdx.x >= 128) return;
  const int rt = idx % (AE / 128), kq = idx / (AE / 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem_w = smem;                         // 2 chunks x 16 KB
  uint8_t* smem_b = smem + 2 * CH128;             // 2 chunks x 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * CH128 + 2 * CH32);
  uint64_t* bar_w = bars;
  uint64_t* bar_b = bars + 1;
  uint64_t* bar_mma = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_b, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  named_bar_sync(1, 128);
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_w, 2 * CH128);
    for (int kc = 0; kc < 2; ++kc) tma_load_2d(smem_w + kc * CH128, tm_wh, bar_w, (2 * kq + kc) * 64, rt * 128);
    mbar_wait(bar_w, 0);
  }
  const int row = rt * 128 + threadIdx.x;
  const float bias = kq == 0 ? __ldg(a.b_h + row) : 0.f;
  const long long dl = lane < a.B ? a.decode_len[lane] : 0;
  const int Tn = steps_to_run(a.T, a.decode_len);
  long long* dbg = idx == 0 && threadIdx.x == 0 ? a.dbg : nullptr;
  const uint32_t wlo = desc_lo(smem_w), blo = desc_lo(smem_b);
  for (int t = 0; t < Tn; ++t) {
    const int bt = __popc(__ballot_sync(0xffffffffu, dl > t));
    if (threadIdx.x == 0) {
      dbg_stamp(dbg, smem, t, 0);
      wait_counter(a.cnt_h + t, N_G2, 10);
      dbg_stamp(dbg, smem, t, 1);
      mbar_expect_tx(bar_b, 2 * CH32);
      bulk_g2s(smem_b, a.HB + (static_cast<long long>(t) * 8 + 2 * kq) * CH32, 2 * CH32, bar_b);
      mbar_wait(bar_b, t & 1);
      tc_fence_after();
      mma_chunks<0, 2, CH128, true>(tmem, wlo, blo);
      tc_commit(bar_mma);
      dbg_stamp(dbg, smem, t, 2);
    }
    __syncwarp();
    mbar_wait(bar_mma, t & 1);
    tc_fence_after();
    if (threadIdx.x == 0) dbg_stamp(dbg, smem, t, 3);
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_ld_wait();
    float* dst = a.HGp + static_cast<long long>((t & 1) * G1_KSPLIT + kq) * a.B * AE + row;
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (b < bt) dst[static_cast<long long>(b) * AE] = __uint_as_float(v[b]) + bias;
    tc_fence_before();
    named_bar_sync(1, 128);
    if (threadIdx.x == 0) {
      signal_counter(a.cnt_hg + t);
      dbg_stamp(dbg, smem, t, 4);
    }
  }
  dbg_flush(dbg, smem, 1, a.T);
  tc_fence_before();
  named_bar_sync(1, 128);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// role G2 (mma.sync): 32 permuted gate rows (8 hidden units x {i,f,g,o}) over K = [h (512) | gated awe (1024)] with
// the LSTMCell gate math as epilogue.  8 warps split K: warp w takes h chunk w as soon as h_{t-1} lands (off the
// critical path) and context chunks 2w, 2w+1 when their quarter of the gated context lands, all into one register
// accumulator (32 rows x 32 batch columns per warp); the eight partial sums meet in shared memory and 256 threads do
// the gate math, one (unit, batch row) each, with the cell state c in a register across steps.
// ---------------------------------------------------------------------------------------------------------------
constexpr int G2_THREADS = 256;
__device__ void fwd_role_g2(const FwdArgs& a, const CUtensorMap* tm_w2, int tile, uint8_t* smem) {
  if (threadIdx.x >= G2_THREADS) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem_w = smem;                          // 24 chunks x 4 KB  (chunks 0-7: h part, 8-23: awe part)
  uint8_t* smem_b = smem + 24 * CH32;              // 24 chunks x 4 KB; chunks 0-7 double as the partial-sum scratch
  float* s_p = reinterpret_cast<float*>(smem_b);   // [8 warps][32 rows][32 cols] after the h part has been consumed
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 48 * CH32);
  uint64_t* bar_w = bars;
  uint64_t* bar_h = bars + 1;
  uint64_t* bar_a = bars + 2;   // [4]
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_h, 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar_a + i, 1);
    mbar_fence_init();
  }
  named_bar_sync(2, G2_THREADS);
  if (tid == 0) {
    mbar_expect_tx(bar_w, 24 * CH32);
    for (int kc = 0; kc < 24; ++kc) tma_load_2d(smem_w + kc * CH32, tm_w2, bar_w, kc * 64, tile * 32);
  }
  const int jj = tid & 7, b = tid >> 3;            // this thread's (hidden unit, batch row)
  const int unit = tile * 8 + jj;
  const long long dl = lane < a.B ? a.decode_len[lane] : 0;
  const int Tn = steps_to_run(a.T, a.decode_len);
  float c = b < a.B ? a.C_all[static_cast<long long>(b) * D + unit] : 0.f;
  const uint32_t hoff_img = img_off(b, unit);
  // h_0 (written row-major by the init_h GEMM) -> operand image 0
  if (b < a.B)
    *reinterpret_cast<__nv_bfloat16*>(a.HB + hoff_img) = a.XH[static_cast<long long>(b) * KX + HOFF + unit];
  named_bar_sync(2, G2_THREADS);
  if (tid == 0) signal_counter(a.cnt_h);
  mbar_wait(bar_w, 0);
  long long* dbg = tile == 0 && tid == 0 ? a.dbg : nullptr;
  const uint32_t sw = smem_u32(smem_w), sb = smem_u32(smem_b);
  for (int t = 0; t < Tn; ++t) {
    const int bt = __popc(__ballot_sync(0xffffffffu, dl > t));
    const bool live = b < bt;
    // what does not depend on the recurrence is fetched first: hoisted emb-part gates and the dropout multiplier
    float e[4];
    const float* ep = a.E_all + (static_cast<long long>(t) * a.B + b) * G4 + tile * 32 + jj;
#pragma unroll
    for (int g = 0; g < 4; ++g) e[g] = live ? __ldg(ep + g * 8) : 0.f;
    const float dm = (a.dropmask != nullptr && live)
                         ? __ldg(a.dropmask + (static_cast<long long>(b) * a.T + t) * D + unit) : 1.f;
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
    if (tid == 0) {
      dbg_stamp(dbg, smem, t, 0);
      wait_counter(a.cnt_h + t, N_G2, 20);
      fence_proxy_async_smem();      // the operand buffers were last read through the generic proxy (ldmatrix / s_p)
      mbar_expect_tx(bar_h, 8 * CH32);
      bulk_g2s(smem_b, a.HB + static_cast<long long>(t) * 8 * CH32, 8 * CH32, bar_h);
    }
    mbar_wait(bar_h, t & 1);
    hmma_chunk<2>(acc, sw + warp * CH32, sb + warp * CH32, lane);
    named_bar_sync(2, G2_THREADS);   // every warp is done with the h chunks: their space becomes the s_p scratch
    if (tid == 0) {
      dbg_stamp(dbg, smem, t, 1);
      wait_counter(a.cnt_awe + t, bt, 21);
      dbg_stamp(dbg, smem, t, 2);
      for (int g = 0; g < 4; ++g) {
        mbar_expect_tx(bar_a + g, 4 * CH32);
        bulk_g2s(smem_b + (8 + 4 * g) * CH32, a.GA + (static_cast<long long>(t) * 16 + 4 * g) * CH32, 4 * CH32,
                 bar_a + g);
      }
    }
    __syncwarp();
    mbar_wait(bar_a + (warp >> 1), t & 1);
    hmma_chunk<2>(acc, sw + (8 + 2 * warp) * CH32, sb + (8 + 2 * warp) * CH32, lane);
    hmma_chunk<2>(acc, sw + (9 + 2 * warp) * CH32, sb + (9 + 2 * warp) * CH32, lane);
    store_partials<2>(s_p + warp * 1024, acc, lane);
    if (tid == 0) dbg_stamp(dbg, smem, t, 3);
    named_bar_sync(2, G2_THREADS);
    if (tid == 0) dbg_stamp(dbg, smem, t, 4);
    float x4[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int r = g * 8 + jj;
      const float* sp = s_p + r * 32 + rot_col(r, b);
      float v = e[g];
#pragma unroll
      for (int w = 0; w < 8; ++w) v += sp[w * 1024];
      x4[g] = v;
    }
    const float ig = sigmoid_fast(x4[0]), fg = sigmoid_fast(x4[1]), gg = tanh_fast(x4[2]), og = sigmoid_fast(x4[3]);
    const float cn = fg * c + ig * gg;
    const float h = og * tanh_fast(cn);
    const __nv_bfloat16 hb = __float2bfloat16_rn(h);
    // the recurrence only waits for h_t: publish it first, everything BPTT needs goes out after the signal
    if (live) {
      c = cn;
      *reinterpret_cast<__nv_bfloat16*>(a.HB + static_cast<long long>(t + 1) * 8 * CH32 + hoff_img) = hb;
    }
    named_bar_sync(2, G2_THREADS);
    if (tid == 0) {
      signal_counter(a.cnt_h + t + 1);
      dbg_stamp(dbg, smem, t, 5);
    }
    if (live) {
      a.XH[(static_cast<long long>(t + 1) * a.B + b) * KX + HOFF + unit] = hb;
      a.C_all[(static_cast<long long>(t + 1) * a.B + b) * D + unit] = cn;
      a.H_all[(static_cast<long long>(b) * a.T + t) * D + unit] = __float2bfloat16_rn(h * dm);
      float* gdst = a.G + (static_cast<long long>(t) * a.B + b) * G4 + unit;
      gdst[0] = x4[0];
      gdst[D] = x4[1];
      gdst[2 * D] = x4[2];
      gdst[3 * D] = x4[3];
    }
  }
  dbg_flush(dbg, smem, 0, a.T);
}

// ---------------------------------------------------------------------------------------------------------------
// role ATT: sample b.  att1[b] (P x A) and enc[b] (P x E), bf16, stay in shared memory for the whole caption.
//   scores   e_p = w_f . relu(att1_p + att2)      CUDA cores (the non-linearity sits inside the sum), warp per pixel
//   context  awe = sum_p alpha_p enc_p            mma.sync: A = enc^T tiles (ldmatrix.trans, 16 channels x 16 pixels),
//            B = alpha in ONE column of the n8 tile — 1/8 of the tensor work is useful, but 256 HMMAs replace the
//            50 k FMAs + 25 k shared loads of the scalar loop, which was bound by the SM's instruction issue
// enc rows are padded to ENC_LD bytes so that the eight 16-byte rows of an ldmatrix tile fall into distinct banks; pixel
// rows P..63 are zero (alpha is zero there, but 0 x garbage could be NaN).
// ---------------------------------------------------------------------------------------------------------------
constexpr int PPAD = 64;                 // pixels padded to 4 k-steps of 16
constexpr int ENC_LD = E * 2 + 16;       // bytes per enc row in shared memory
constexpr int ATT_MAX_PIX = 49;          // att1 + padded enc of one sample must fit one CTA's shared memory
__device__ __forceinline__ void ldsm4_t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
// enc[b] -> shared memory, one bulk copy per pixel row; rows P..PPAD-1 zeroed by the threads
__device__ __forceinline__ void att_load_enc(uint8_t* s_enc, const __nv_bfloat16* enc_b, int P, uint64_t* bar, int tid) {
  if (tid == 0)
    for (int p = 0; p < P; ++p) bulk_g2s(s_enc + p * ENC_LD, enc_b + static_cast<long long>(p) * E, E * 2, bar);
  for (int i = tid; i < (PPAD - P) * (ENC_LD / 16); i += THREADS)
    reinterpret_cast<uint4*>(s_enc + P * ENC_LD)[i] = make_uint4(0u, 0u, 0u, 0u);
}

__device__ void fwd_role_att(const FwdArgs& a, int b, uint8_t* smem) {
  const int P = a.P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* s_enc = smem;                                                        // [PPAD][ENC_LD]
  uint32_t* s_att1 = reinterpret_cast<uint32_t*>(smem + PPAD * ENC_LD);         // [P][256] bf16x2
  float* s_hg = reinterpret_cast<float*>(smem + PPAD * ENC_LD + ATT_MAX_PIX * A * 2);   // [AE]
  float* s_awe = s_hg + AE;                                                     // [E]
  float* s_wf = s_awe + E;                                                      // [A] full_att weight
  float* s_e = s_wf + A;                                                        // [2][MAX_P] (double-buffered by step)
  float* s_al = s_e + 2 * MAX_P;                                                // [MAX_P] alpha, zero beyond P
  uint32_t* s_alb = reinterpret_cast<uint32_t*>(s_al + MAX_P);                  // [MAX_P / 2] alpha as bf16 pairs
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_alb + MAX_P / 2);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, static_cast<uint32_t>(P * (A + E) * 2));
    bulk_g2s(s_att1, a.att1 + static_cast<long long>(b) * P * A, static_cast<uint32_t>(P * A * 2), bar);
  }
  att_load_enc(s_enc, a.enc + static_cast<long long>(b) * P * E, P, bar, tid);
  if (tid < MAX_P) s_al[tid] = 0.f;
  s_wf[tid] = __ldg(a.w_f + tid);               // A == THREADS
  __syncthreads();
  mbar_wait(bar, 0);
  const float bf = a.b_f ? __ldg(a.b_f) : 0.f;
  const int steps = static_cast<int>(a.decode_len[b]);
  const uint32_t goff = img_off(b, 2 * tid);      // this thread's two context channels inside the operand image
  long long* dbg = (b == 0 && tid == 0) ? a.dbg : nullptr;
  const uint32_t enc_s = smem_u32(s_enc);
  const int mi = lane >> 3, ri = lane & 7, g = lane >> 2, tq = lane & 3;
  for (int t = 0; t < steps; ++t) {
    float* se = s_e + (t & 1) * MAX_P;
    // ONE poller per CTA: hundreds of pollers on one counter word serialise in its L2 slice
    if (tid == 0) {
      dbg_stamp(dbg, smem, t, 0);
      wait_counter(a.cnt_hg + t, N_G1, 30);
      dbg_stamp(dbg, smem, t, 1);
    }
    __syncthreads();
    // [att2 | gate pre-activation] of this sample = the four K-quarter sums of role G1, added in quarter order;
    // fetched once per CTA (every warp needs all of att2: 16 warps loading it themselves cost 16x the L2 traffic)
    if (tid < AE / 4) {
      const float4* hg = reinterpret_cast<const float4*>(
                             a.HGp + (static_cast<long long>((t & 1) * G1_KSPLIT) * a.B + b) * AE) + tid;
      const long long qs = static_cast<long long>(a.B) * AE / 4;
      float4 v = __ldcg(hg);
#pragma unroll
      for (int q = 1; q < G1_KSPLIT; ++q) {
        const float4 x = __ldcg(hg + q * qs);
        v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
      }
      reinterpret_cast<float4*>(s_hg)[tid] = v;
    }
    __syncthreads();
    float a2[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 h2 = *reinterpret_cast<const float2*>(s_hg + 2 * (lane + 32 * i));
      a2[2 * i] = h2.x;
      a2[2 * i + 1] = h2.y;
    }
    for (int p = warp; p < P; p += THREADS / 32) {
      const uint32_t* row = s_att1 + p * (A / 2);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 x = unpack_bf16x2(row[lane + 32 * i]);
        const float2 w = *reinterpret_cast<const float2*>(s_wf + 2 * (lane + 32 * i));
        acc = fmaf(fmaxf(x.x + a2[2 * i], 0.f), w.x, acc);
        acc = fmaf(fmaxf(x.y + a2[2 * i + 1], 0.f), w.y, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) se[p] = acc + bf;
    }
    if (tid == 0) dbg_stamp(dbg, smem, t, 2);
    __syncthreads();
    // softmax over pixels (warp 0; lane owns pixels lane, lane+32)
    float al0 = 0.f, al1 = 0.f;
    if (warp == 0) {
      const float e0 = lane < P ? se[lane] : -INFINITY;
      const float e1 = lane + 32 < P ? se[lane + 32] : -INFINITY;
      const float mx = warp_max(fmaxf(e0, e1));
      const float x0 = lane < P ? __expf(e0 - mx) : 0.f;
      const float x1 = lane + 32 < P ? __expf(e1 - mx) : 0.f;
      const float inv = __fdividef(1.0f, warp_sum(x0 + x1));
      al0 = x0 * inv;
      al1 = x1 * inv;
      s_al[lane] = al0;
      s_al[lane + 32] = al1;
      __syncwarp();
      s_alb[lane] = pack_bf16x2(s_al[2 * lane], s_al[2 * lane + 1]);
    }
    __syncthreads();
    if (tid == 0) dbg_stamp(dbg, smem, t, 3);
    // context vector on the tensor cores: warp w owns channels [64w, 64w+64) = 4 m16 tiles, K = 64 pixels
    {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t b0 = g == 0 ? s_alb[8 * ks + tq] : 0u;
        const uint32_t b1 = g == 0 ? s_alb[8 * ks + 4 + tq] : 0u;
        const int pix = 16 * ks + (mi >> 1) * 8 + ri;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          uint32_t af[4];
          ldsm4_t(af, enc_s + pix * ENC_LD + (64 * warp + 16 * mt + (mi & 1) * 8) * 2);
          hmma(acc[mt], af, b0, b1);
        }
      }
      if (tq == 0) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          s_awe[64 * warp + 16 * mt + g] = acc[mt][0];
          s_awe[64 * warp + 16 * mt + g + 8] = acc[mt][2];
        }
      }
    }
    __syncthreads();
    const float2 awe = *reinterpret_cast<const float2*>(s_awe + 2 * tid);
    const float2 gp = *reinterpret_cast<const float2*>(s_hg + A + 2 * tid);
    const uint32_t gawe = pack_bf16x2(awe.x * sigmoid_fast(gp.x), awe.y * sigmoid_fast(gp.y));
    *reinterpret_cast<uint32_t*>(a.GA + static_cast<long long>(t) * 16 * CH32 + goff) = gawe;
    if (tid == 0) dbg_stamp(dbg, smem, t, 4);
    __syncthreads();
    if (tid == 0) {
      signal_counter(a.cnt_awe + t);
      dbg_stamp(dbg, smem, t, 5);
    }
    // off the critical path: what only the caller / BPTT reads
    *reinterpret_cast<uint32_t*>(a.XH + (static_cast<long long>(t) * a.B + b) * KX + EMB + 2 * tid) = gawe;
    float* hgo = a.HG + (static_cast<long long>(t) * a.B + b) * AE;
    if (tid < AE / 4) reinterpret_cast<float4*>(hgo)[tid] = reinterpret_cast<const float4*>(s_hg)[tid];
    if (warp == 0) {
      float* ao = a.alphas + (static_cast<long long>(b) * a.T + t) * P;
      if (lane < P) ao[lane] = al0;
      if (lane + 32 < P) ao[lane + 32] = al1;
    }
    if (a.awe_all != nullptr)
      *reinterpret_cast<float2*>(a.awe_all + (static_cast<long long>(t) * a.B + b) * E + 2 * tid) = awe;
  }
  dbg_flush(dbg, smem, 2, a.T);
}

__global__ void __launch_bounds__(THREADS, 1)
lstm_tf_fwd_persist_kernel(const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_w2,
                           const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int bid = blockIdx.x;
  if (bid < N_G2) fwd_role_g2(a, &tm_w2, bid, smem);
  else if (bid < N_G2 + N_G1) fwd_role_g1(a, &tm_wh, bid - N_G2, smem);
  else fwd_role_att(a, bid - N_G2 - N_G1, smem);
}

// ===============================================================================================================
// Backward through time (the autograd graph of models/decoder.py:100-111), same scheme, t = T-1 .. 0:
//   HP  (32 CTAs, 16 hidden units each)  d h_{t} = dHG[t] . [decoder_att ; f_beta]  (K = 1536, mma.sync)
//                                                 + the W_hh partial sums of role X; then the LSTMCell point-wise
//                                                 backward of step t-1 in the epilogue (d c carried in a register):
//                                                 dgates[t-1] -> fp32 (kept for the weight gradients) and bf16 operand
//   X   (64 + 16 CTAs, tcgen05) [d gated-awe | d h via W_hh] = dgates[t] . [W_ih(awe) | W_hh], 2-D tiled: 128 output
//                                                 rows x 256 (context rows) or 512 (W_hh rows) of K per CTA; the
//                                                 consumers add the K-slice partial sums in a fixed order
//   ATT (B CTAs)   attention backward of sample b with att1[b] / enc[b] (bf16) resident in shared memory: gate and
//                  softmax backward, d att2, d gate-pre-activation, per-step records d_awe / d e for the deferred
//                  d_enc / d_att1 sums, d w_f accumulated in a register across all steps
// critical path per step: HP(t+1) -> X(t) -> ATT(t) -> HP(t).
// ===============================================================================================================
constexpr int N_HP = D / 16;   // 32
constexpr int XW = E + D;      // 1536 rows of [W_ih(awe)^T ; W_hh^T]
constexpr int XA_KS = 8;       // role X, context rows: 8 row tiles x 8 K eighths (256 of the 2048 gate columns each)
constexpr int N_XA = (E / 128) * XA_KS;   // 64 CTAs
constexpr int XH_KS = 4;       // role X, W_hh rows: 4 row tiles x 4 K quarters (needed only by HP's epilogue)
constexpr int N_XH = (D / 128) * XH_KS;   // 16 CTAs

struct BwdArgs {
  const float* G;             // [T][B][G4]
  const float* C_all;         // [T+1][B][D]
  const float* HG;            // [T][B][AE]
  const float* alphas;        // [B][T][P]
  const float* awe_all;       // [T][B][E]
  const float* dH_all;        // [B][T][D] fc dgrad
  const float* dropmask;      // [B][T][D] or nullptr
  const float* dalphas;       // [B][T][P] or nullptr
  const __nv_bfloat16* att1;  // [B][P][A]
  const __nv_bfloat16* enc;   // [B][P][E]
  const float* w_f;           // [A]
  float* dG_all;              // [T][B][G4]  (zero-initialised: rows that are inactive at a step stay zero)
  __nv_bfloat16* dG_bf;       // [T][B][G4]  row-major bf16 copy (zero-initialised; operand of the d_emb GEMM)
  float* dHG_all;             // [T][B][AE]  (zero-initialised)
  uint8_t* DGI;               // [T][128 KB] dgates[t] operand image, 32 chunks (rows of inactive samples undefined)
  uint8_t* DHI;               // [T][96 KB]  d[att2 | gate][t] operand image, 24 chunks
  float* dawe_all;            // [T][B][E]   (zero-initialised)
  float* de_all;              // [T][B][P]   (zero-initialised)
  float* d_wf;                // [A] +=
  float* XpA;                 // [2][XA_KS][B][E] partial sums d(gated awe) of role X, double-buffered by step parity
  float* XpH;                 // [2][XH_KS][B][D] partial sums of dgates . W_hh
  float* dh_out;              // [B][D] dL/dh_0
  float* dc_out;              // [B][D] dL/dc_0
  const long long* decode_len;
  int* cnt_dg;                // [T] dgates[t] complete (N_HP arrivals)
  int* cnt_x;                 // [T] d(gated awe) partials of step t complete (N_XA arrivals)
  int* cnt_xh;                // [T] W_hh partials of step t complete (N_XH arrivals)
  int* cnt_dhg;               // [T] dHG[t] complete (bt arrivals)
  long long* dbg;
  int B, T, P;
};

// NCH: 64-wide K chunks per CTA (4 for the context rows, 8 for the W_hh rows)
template <int NCH>
__device__ void bwd_role_x(const BwdArgs& a, const CUtensorMap* tm_wx, int row0, int kslice, float* out, int out_ld,
                           long long out_slice, long long out_parity, int* cnt, int dbg_role, bool dbg_on,
                           uint8_t* smem) {
  if (threadIdx.x >= 128) return;
  const int warp = threadIdx.x >> 5;
  uint8_t* smem_w = smem;                         // NCH chunks x 16 KB
  uint8_t* smem_b = smem + NCH * CH128;           // NCH chunks x 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NCH * CH128 + NCH * CH32);
  uint64_t* bar_w = bars;
  uint64_t* bar_b = bars + 1;
  uint64_t* bar_mma = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_b, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  named_bar_sync(1, 128);
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int chunk0 = kslice * NCH;                // first 64-wide chunk of the 2048 gate columns
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_w, NCH * CH128);
    for (int kc = 0; kc < NCH; ++kc) tma_load_2d(smem_w + kc * CH128, tm_wx, bar_w, (chunk0 + kc) * 64, row0);
    mbar_wait(bar_w, 0);
  }
  long long* dbg = dbg_on && threadIdx.x == 0 ? a.dbg : nullptr;
  const int Tn = steps_to_run(a.T, a.decode_len);
  const uint32_t wlo = desc_lo(smem_w), blo = desc_lo(smem_b);
  for (int it = 0; it < Tn; ++it) {
    const int t = Tn - 1 - it;
    if (threadIdx.x == 0) {
      dbg_stamp(dbg, smem, t, 0);
      wait_counter(a.cnt_dg + t, N_HP, 40);
      dbg_stamp(dbg, smem, t, 1);
      mbar_expect_tx(bar_b, NCH * CH32);
      bulk_g2s(smem_b, a.DGI + (static_cast<long long>(t) * 32 + chunk0) * CH32, NCH * CH32, bar_b);
      mbar_wait(bar_b, it & 1);
      tc_fence_after();
      mma_chunks<0, NCH, CH128, true>(tmem, wlo, blo);
      tc_commit(bar_mma);
      dbg_stamp(dbg, smem, t, 2);
    }
    __syncwarp();
    mbar_wait(bar_mma, it & 1);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_ld_wait();
    float* dst = out + (t & 1) * out_parity + kslice * out_slice + threadIdx.x;
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (b < a.B) dst[static_cast<long long>(b) * out_ld] = __uint_as_float(v[b]);
    tc_fence_before();
    named_bar_sync(1, 128);
    if (threadIdx.x == 0) {
      signal_counter(cnt + t);
      dbg_stamp(dbg, smem, t, 3);
    }
  }
  dbg_flush(dbg, smem, dbg_role, a.T);
  tc_fence_before();
  named_bar_sync(1, 128);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

// LSTMCell point-wise backward of (step tp, row b, unit j); see lstm_pointwise_bwd_kernel (train_kernels.cu).
// Everything that depends only on the forward pass is folded into six factors BEFORE the recurrent gradient arrives:
//   dcn = dc + dh * A1;  dgates = (dcn * B0, dcn * B1, dcn * B2, dh * B3);  dc' = dcn * F;   dh = carry + dhfc
struct PwIn {
  float A1, B0, B1, B2, B3, F, dhfc;
};
__device__ __forceinline__ PwIn pw_fetch(const BwdArgs& a, int tp, int b, int j) {
  const float* g = a.G + (static_cast<long long>(tp) * a.B + b) * G4 + j;
  const float gi = __ldg(g), gf = __ldg(g + D), gg = __ldg(g + 2 * D), go = __ldg(g + 3 * D);
  const float c_prev = __ldg(a.C_all + (static_cast<long long>(tp) * a.B + b) * D + j);
  const float c_new = __ldg(a.C_all + (static_cast<long long>(tp + 1) * a.B + b) * D + j);
  const long long o = (static_cast<long long>(b) * a.T + tp) * D + j;
  PwIn r;
  r.dhfc = __ldg(a.dH_all + o) * (a.dropmask ? __ldg(a.dropmask + o) : 1.f);
  const float i_ = sigmoidf_(gi), f_ = sigmoidf_(gf), g_ = tanhf(gg), o_ = sigmoidf_(go);
  const float tc = tanhf(c_new);
  r.A1 = o_ * (1.f - tc * tc);
  r.B0 = g_ * i_ * (1.f - i_);
  r.B1 = c_prev * f_ * (1.f - f_);
  r.B2 = i_ * (1.f - g_ * g_);
  r.B3 = tc * o_ * (1.f - o_);
  r.F = f_;
  return r;
}
struct PwOut {
  float d0, d1, d2, d3;
};
// the critical part: dgates as bf16 into the operand image the X CTAs wait for
__device__ __forceinline__ PwOut pw_apply(const BwdArgs& a, const PwIn& r, int tp, int b, int j, float dh_carry,
                                          float& dc) {
  const float dh = dh_carry + r.dhfc;
  const float dcn = fmaf(dh, r.A1, dc);
  PwOut o;
  o.d0 = dcn * r.B0;
  o.d1 = dcn * r.B1;
  o.d2 = dcn * r.B2;
  o.d3 = dh * r.B3;
  dc = dcn * r.F;
  uint8_t* img = a.DGI + static_cast<long long>(tp) * 32 * CH32;
  *reinterpret_cast<__nv_bfloat16*>(img + img_off(b, j)) = __float2bfloat16_rn(o.d0);
  *reinterpret_cast<__nv_bfloat16*>(img + img_off(b, D + j)) = __float2bfloat16_rn(o.d1);
  *reinterpret_cast<__nv_bfloat16*>(img + img_off(b, 2 * D + j)) = __float2bfloat16_rn(o.d2);
  *reinterpret_cast<__nv_bfloat16*>(img + img_off(b, 3 * D + j)) = __float2bfloat16_rn(o.d3);
  return o;
}
// after the hand-over: the fp32 record (weight gradients) and the row-major bf16 copy (embedding-gradient GEMM)
__device__ __forceinline__ void pw_store(const BwdArgs& a, const PwOut& o, int tp, int b, int j) {
  const long long off = (static_cast<long long>(tp) * a.B + b) * G4 + j;
  a.dG_all[off] = o.d0;
  a.dG_all[off + D] = o.d1;
  a.dG_all[off + 2 * D] = o.d2;
  a.dG_all[off + 3 * D] = o.d3;
  a.dG_bf[off] = __float2bfloat16_rn(o.d0);
  a.dG_bf[off + D] = __float2bfloat16_rn(o.d1);
  a.dG_bf[off + 2 * D] = __float2bfloat16_rn(o.d2);
  a.dG_bf[off + 3 * D] = __float2bfloat16_rn(o.d3);
}

__device__ void bwd_role_hp(const BwdArgs& a, const CUtensorMap* tm_wht, int tile, uint8_t* smem) {
  if (threadIdx.x >= 128) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int CH16 = 16 * 128;                   // bytes of a [16 rows x 64 bf16] chunk
  uint8_t* smem_w = smem;                          // 24 chunks x 2 KB
  uint8_t* smem_b = smem + 48 * 1024;              // 24 chunks x 4 KB
  float* s_p = reinterpret_cast<float*>(smem + 48 * 1024 + 24 * CH32);  // [4 warps][16 rows][32 cols]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 48 * 1024 + 24 * CH32 + 4 * 16 * 32 * 4);
  uint64_t* bar_w = bars;
  uint64_t* bar_a = bars + 1;   // [4]
  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar_a + i, 1);
    mbar_fence_init();
  }
  named_bar_sync(1, 128);
  if (tid == 0) {
    mbar_expect_tx(bar_w, 24 * CH16);
    for (int kc = 0; kc < 24; ++kc) tma_load_2d(smem_w + kc * CH16, tm_wht, bar_w, kc * 64, tile * 16);
  }
  const int u = tid & 15, bq = tid >> 4;            // unit within the tile; batch rows bq, bq+8, bq+16, bq+24
  const int j = tile * 16 + u;
  const long long dl = lane < a.B ? a.decode_len[lane] : 0;
  const int Tn = steps_to_run(a.T, a.decode_len);
  float dc[4] = {0.f, 0.f, 0.f, 0.f};
  long long* dbg = tile == 0 && tid == 0 ? a.dbg : nullptr;
  const uint32_t sw = smem_u32(smem_w), sb = smem_u32(smem_b);
  // prologue: point-wise backward of the last step (no recurrent gradient yet)
  {
    const int tp = Tn - 1;
    const int btp = __popc(__ballot_sync(0xffffffffu, dl > tp));
    PwOut po[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = bq + 8 * k;
      if (b < btp) po[k] = pw_apply(a, pw_fetch(a, tp, b, j), tp, b, j, 0.f, dc[k]);
    }
    named_bar_sync(1, 128);
    if (tid == 0) signal_counter(a.cnt_dg + tp);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = bq + 8 * k;
      if (b < btp) pw_store(a, po[k], tp, b, j);
    }
  }
  mbar_wait(bar_w, 0);
  for (int it = 0; it < Tn; ++it) {
    const int t = Tn - 1 - it;
    const int bt = __popc(__ballot_sync(0xffffffffu, dl > t));
    const int btp = t > 0 ? __popc(__ballot_sync(0xffffffffu, dl > t - 1)) : 0;
    // forward-pass factors of step t-1 do not depend on the recurrence: compute them before waiting
    PwIn pin[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = bq + 8 * k;
      pin[k] = PwIn{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (b < btp) pin[k] = pw_fetch(a, t - 1, b, j);
    }
    if (tid == 0) {
      dbg_stamp(dbg, smem, t, 0);
      wait_counter(a.cnt_dhg + t, bt, 50);
      wait_counter(a.cnt_xh + t, N_XH, 51);
      dbg_stamp(dbg, smem, t, 1);
      fence_proxy_async_smem();      // the operand buffers were last read through the generic proxy (ldmatrix)
      for (int g = 0; g < 4; ++g) {
        mbar_expect_tx(bar_a + g, 6 * CH32);
        bulk_g2s(smem_b + 6 * g * CH32, a.DHI + (static_cast<long long>(t) * 24 + 6 * g) * CH32, 6 * CH32, bar_a + g);
      }
    }
    named_bar_sync(1, 128);      // hand-over seen by every thread: the W_hh partial sums of step t may be read
    // d h via W_hh: the K-quarter partial sums of role X, fetched while the copies are in flight
    const float* xp = a.XpH + static_cast<long long>((t & 1) * XH_KS) * a.B * D + j;
    float part[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = bq + 8 * k;
      part[k] = 0.f;
      if (b < bt) {
#pragma unroll
        for (int c4 = 0; c4 < XH_KS; ++c4) part[k] += __ldcg(xp + (static_cast<long long>(c4) * a.B + b) * D);
      }
    }
    // warp w: K chunks 6w .. 6w+5 as soon as its quarter of the operand lands
    float acc[1][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[0][i][k] = 0.f;
    mbar_wait(bar_a + warp, it & 1);
#pragma unroll
    for (int q = 0; q < 6; ++q) hmma_chunk<1>(acc, sw + (6 * warp + q) * CH16, sb + (6 * warp + q) * CH32, lane);
    store_partials<1>(s_p + warp * 512, acc, lane);
    if (tid == 0) dbg_stamp(dbg, smem, t, 2);
    named_bar_sync(1, 128);
    if (tid == 0) dbg_stamp(dbg, smem, t, 3);
    PwOut po[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = bq + 8 * k;
      if (b >= a.B) continue;
      // rows that were not active at step t carry no recurrent gradient (their operand rows are undefined)
      float dh = 0.f;
      if (b < bt) {
        const float* sp = s_p + u * 32 + rot_col(u, b);
        dh = part[k] + ((sp[0] + sp[512]) + (sp[1024] + sp[1536]));
      }
      if (t == 0) {
        a.dh_out[static_cast<long long>(b) * D + j] = dh;
        a.dc_out[static_cast<long long>(b) * D + j] = dc[k];
      } else if (b < btp) {
        po[k] = pw_apply(a, pin[k], t - 1, b, j, dh, dc[k]);
      }
    }
    named_bar_sync(1, 128);
    if (t > 0) {
      if (tid == 0) {
        signal_counter(a.cnt_dg + t - 1);
        dbg_stamp(dbg, smem, t, 4);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int b = bq + 8 * k;
        if (b < btp) pw_store(a, po[k], t - 1, b, j);
      }
    }
  }
  dbg_flush(dbg, smem, 0, a.T);
}

// role ATT (backward): enc[b] in shared memory (padded rows) for d alpha_p = enc_p . d_awe on mma.sync (A = enc tiles,
// B = d_awe in one column of the n8 tile, warp w owns channels [64w, 64w+64), the 16 partial sums meet in shared
// memory); att1[b] in REGISTERS column-wise (thread tid owns attention unit tid of every pixel, two pixels per word)
// for the pass through relu(att1 + att2); d e reaches that pass as 4-wide broadcast reads.
constexpr int PR = 52;            // register-resident pixels of the att1 column (P <= 49 rounded up to a multiple of 4)
__device__ void bwd_role_att(const BwdArgs& a, int b, uint8_t* smem) {
  const int P = a.P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* s_enc = smem;                                                        // [PPAD][ENC_LD]
  uint32_t* s_dawb = reinterpret_cast<uint32_t*>(smem + PPAD * ENC_LD);         // [E / 2] d_awe as bf16 pairs
  float* s_part = reinterpret_cast<float*>(s_dawb + E / 2);                     // [16 warps][PPAD] d alpha partials
  float* s_de = s_part + 16 * PPAD;                                             // [MAX_P] d e, zero beyond P
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_de + MAX_P);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, static_cast<uint32_t>(P * E * 2));
  }
  att_load_enc(s_enc, a.enc + static_cast<long long>(b) * P * E, P, bar, tid);
  if (tid < MAX_P) s_de[tid] = 0.f;
  uint32_t a1c[PR / 2];          // word i = (att1[2i][tid], att1[2i+1][tid])
  {
    const unsigned short* ag = reinterpret_cast<const unsigned short*>(a.att1 + static_cast<long long>(b) * P * A) + tid;
#pragma unroll
    for (int i = 0; i < PR / 2; ++i) {
      const uint32_t lo = 2 * i < P ? __ldg(ag + (2 * i) * A) : 0xff80u;          // bf16 -inf: relu never fires
      const uint32_t hi = 2 * i + 1 < P ? __ldg(ag + (2 * i + 1) * A) : 0xff80u;
      a1c[i] = lo | (hi << 16);
    }
  }
  __syncthreads();
  mbar_wait(bar, 0);
  const float wf = __ldg(a.w_f + tid);            // thread owns attention unit a = tid (A == THREADS)
  float dwf = 0.f;
  const int steps = static_cast<int>(a.decode_len[b]);
  const uint32_t off_a = img_off(b, tid), off_e = img_off(b, A + 2 * tid);
  long long* dbg = (b == 0 && tid == 0) ? a.dbg : nullptr;
  const uint32_t enc_s = smem_u32(s_enc);
  const int mi = lane >> 3, ri = lane & 7, g = lane >> 2, tq = lane & 3;
  for (int t = steps - 1; t >= 0; --t) {
    // forward-pass values first (no dependence on the recurrence)
    const long long row = static_cast<long long>(t) * a.B + b;
    const float att2 = __ldg(a.HG + row * AE + tid);
    const float2 gp = __ldg(reinterpret_cast<const float2*>(a.HG + row * AE + A) + tid);
    const float2 awe = __ldg(reinterpret_cast<const float2*>(a.awe_all + row * E) + tid);
    const float* al = a.alphas + (static_cast<long long>(b) * a.T + t) * P;
    const float* dax = a.dalphas ? a.dalphas + (static_cast<long long>(b) * a.T + t) * P : nullptr;
    float a0 = 0.f, a1 = 0.f, x0 = 0.f, x1 = 0.f;
    if (warp == 0) {
      a0 = lane < P ? __ldg(al + lane) : 0.f;
      a1 = lane + 32 < P ? __ldg(al + lane + 32) : 0.f;
      x0 = (dax && lane < P) ? __ldg(dax + lane) : 0.f;
      x1 = (dax && lane + 32 < P) ? __ldg(dax + lane + 32) : 0.f;
    }
    const float g0 = sigmoidf_(gp.x), g1 = sigmoidf_(gp.y);
    const float2 gfac = make_float2(awe.x * g0 * (1.f - g0), awe.y * g1 * (1.f - g1));
    if (tid == 0) {
      dbg_stamp(dbg, smem, t, 0);
      wait_counter(a.cnt_x + t, N_XA, 60);
      dbg_stamp(dbg, smem, t, 1);
    }
    __syncthreads();            // also: every read of the shared staging arrays of the previous step is done
    // d(gated awe) of channels 2*tid, 2*tid+1: the K-slice partial sums of role X, in slice order
    const float* xp = a.XpA + (static_cast<long long>((t & 1) * XA_KS) * a.B + b) * E + 2 * tid;
    float2 dga = make_float2(0.f, 0.f);
#pragma unroll
    for (int c4 = 0; c4 < XA_KS; ++c4) {
      const float2 x = __ldcg(reinterpret_cast<const float2*>(xp + static_cast<long long>(c4) * a.B * E));
      dga.x += x.x;
      dga.y += x.y;
    }
    const float2 dgp = make_float2(dga.x * gfac.x, dga.y * gfac.y);
    const float2 daw = make_float2(dga.x * g0, dga.y * g1);
    uint8_t* img = a.DHI + static_cast<long long>(t) * 24 * CH32;
    *reinterpret_cast<uint32_t*>(img + off_e) = pack_bf16x2(dgp.x, dgp.y);
    s_dawb[tid] = pack_bf16x2(daw.x, daw.y);
    if (tid == 0) dbg_stamp(dbg, smem, t, 2);
    __syncthreads();
    // d alpha_p = sum_e enc[p, e] d_awe[e]: warp w contracts channels [64w, 64w+64) for all 64 (padded) pixels
    {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int ch = 64 * warp + 16 * ks;
        const uint32_t b0 = g == 0 ? s_dawb[ch / 2 + tq] : 0u;
        const uint32_t b1 = g == 0 ? s_dawb[ch / 2 + 4 + tq] : 0u;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          uint32_t af[4];
          const int pix = 16 * mt + (mi & 1) * 8 + ri;
          ldsm4(af, enc_s + pix * ENC_LD + (ch + (mi >> 1) * 8) * 2);
          hmma(acc[mt], af, b0, b1);
        }
      }
      if (tq == 0) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          s_part[warp * PPAD + 16 * mt + g] = acc[mt][0];
          s_part[warp * PPAD + 16 * mt + g + 8] = acc[mt][2];
        }
      }
    }
    if (tid == 0) dbg_stamp(dbg, smem, t, 3);
    __syncthreads();
    // softmax backward (warp 0): de_p = alpha_p (dalpha_p - sum_q alpha_q dalpha_q), dalpha includes the gradient
    // that reaches alpha from outside (the doubly-stochastic regulariser, trainMultiGPU.py:369)
    float de0 = 0.f, de1 = 0.f;
    if (warp == 0) {
      float d0 = x0, d1 = x1;
#pragma unroll
      for (int w = 0; w < 16; ++w) {
        d0 += s_part[w * PPAD + lane];
        d1 += s_part[w * PPAD + lane + 32];
      }
      if (lane >= P) d0 = 0.f;
      if (lane + 32 >= P) d1 = 0.f;
      const float dot = warp_sum(fmaf(a0, d0, a1 * d1));
      de0 = a0 * (d0 - dot);
      de1 = a1 * (d1 - dot);
      if (lane < P) s_de[lane] = de0;
      if (lane + 32 < P) s_de[lane + 32] = de1;
    }
    __syncthreads();
    // through e_p = w_f . relu(att1_p + att2): thread owns attention unit `tid`, att1 column from registers
    float datt2 = 0.f;
#pragma unroll
    for (int q = 0; q < PR / 4; ++q) {
      const float4 d4 = *reinterpret_cast<const float4*>(s_de + 4 * q);
      float2 x = unpack_bf16x2(a1c[2 * q]);
      float pre = x.x + att2;
      if (pre > 0.f) { datt2 += d4.x; dwf = fmaf(d4.x, pre, dwf); }
      pre = x.y + att2;
      if (pre > 0.f) { datt2 += d4.y; dwf = fmaf(d4.y, pre, dwf); }
      x = unpack_bf16x2(a1c[2 * q + 1]);
      pre = x.x + att2;
      if (pre > 0.f) { datt2 += d4.z; dwf = fmaf(d4.z, pre, dwf); }
      pre = x.y + att2;
      if (pre > 0.f) { datt2 += d4.w; dwf = fmaf(d4.w, pre, dwf); }
    }
    datt2 *= wf;
    *reinterpret_cast<__nv_bfloat16*>(img + off_a) = __float2bfloat16_rn(datt2);
    if (tid == 0) dbg_stamp(dbg, smem, t, 4);
    __syncthreads();
    if (tid == 0) {
      signal_counter(a.cnt_dhg + t);
      dbg_stamp(dbg, smem, t, 5);
    }
    // off the critical path: fp32 records for the weight gradients and the deferred d_enc / d_att1 sums
    a.dHG_all[row * AE + tid] = datt2;
    *reinterpret_cast<float2*>(a.dHG_all + row * AE + A + 2 * tid) = dgp;
    *reinterpret_cast<float2*>(a.dawe_all + row * E + 2 * tid) = daw;
    if (warp == 0) {
      float* deo = a.de_all + row * P;
      if (lane < P) deo[lane] = de0;
      if (lane + 32 < P) deo[lane + 32] = de1;
    }
  }
  dbg_flush(dbg, smem, 2, a.T);
  atomicAdd(a.d_wf + tid, dwf);
}

__global__ void __launch_bounds__(THREADS, 1)
lstm_tf_bwd_persist_kernel(const __grid_constant__ CUtensorMap tm_wx, const __grid_constant__ CUtensorMap tm_wht,
                           const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int bid = blockIdx.x;
  if (bid < N_HP) {
    bwd_role_hp(a, &tm_wht, bid, smem);
  } else if (bid < N_HP + N_XA) {
    const int i = bid - N_HP, rt = i % (E / 128), ks = i / (E / 128);
    bwd_role_x<4>(a, &tm_wx, rt * 128, ks, a.XpA + rt * 128, E, static_cast<long long>(a.B) * E,
                  static_cast<long long>(XA_KS) * a.B * E, a.cnt_x, 1, i == 0, smem);
  } else if (bid < N_HP + N_XA + N_XH) {
    const int i = bid - N_HP - N_XA, rt = i % (D / 128), ks = i / (D / 128);
    bwd_role_x<8>(a, &tm_wx, E + rt * 128, ks, a.XpH + rt * 128, D, static_cast<long long>(a.B) * D,
                  static_cast<long long>(XH_KS) * a.B * D, a.cnt_xh, 3, false, smem);
  } else {
    bwd_role_att(a, bid - N_HP - N_XA - N_XH, smem);
  }
}

// 2-D row-major bf16 tensor [rows, cols] (leading dimension ld elements), box = [box_rows, 64 elements], 128B swizzle
static int make_map_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld,
                         int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return CCX_ERR_TMA;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) return CCX_ERR_SHAPE;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld * 2)};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CCX_OK : CCX_ERR_TMA;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_coop(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream,
                               Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // co-residency of all CTAs is what the counter protocol relies on
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static bool configure_once(const void* fn, int which) {
  // per-device: cudaFuncSetAttribute applies to the current device only
  static bool done[64][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return false;
  if (!done[dev][which]) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024) != cudaSuccess)
      return false;
    done[dev][which] = true;
  }
  return true;
}

}  // namespace lp
}  // namespace ccx

using namespace ccx;

extern "C" {

int ccx_lstm_persist_supported(int32_t B, int32_t P, int32_t E, int32_t A, int32_t D, int32_t Emb,
                               int32_t compute_dtype) {
  // P: att1 (P rows) and the padded enc (64 rows) of one sample live in one CTA's shared memory
  return (compute_dtype == CCX_BF16 && B >= 1 && B <= lp::NB && P >= 1 && P <= lp::ATT_MAX_PIX && E == lp::E &&
          A == lp::A && D == lp::D && Emb == lp::EMB) ? 1 : 0;
}

int ccx_lstm_tf_forward_persist(const ccx_lstm_tf* s, const ccx_lstm_persist* p, void* stream_) {
  if (s == nullptr || p == nullptr) return CCX_ERR_SHAPE;
  if (!ccx_lstm_persist_supported(s->B, s->P, s->E, s->A, s->D, s->Emb, s->compute_dtype)) return CCX_ERR_SHAPE;
  if (s->T <= 0) return CCX_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int B = s->B, T = s->T;
  if (p->scratch == nullptr || (reinterpret_cast<uintptr_t>(p->scratch) & 127)) return CCX_ERR_SHAPE;
  CUtensorMap tm_wh, tm_w2;
  int rc;
  if ((rc = lp::make_map_bf16(&tm_wh, s->w_h, lp::AE, lp::D, lp::D, 128))) return rc;
  if ((rc = lp::make_map_bf16(&tm_w2, p->w2p, lp::G4, lp::D + lp::E, lp::D + lp::E, 32))) return rc;
  if (cudaMemsetAsync(p->counters, 0, sizeof(int) * 4 * (T + 1), st) != cudaSuccess) return CCX_ERR_CUDA;
  lp::FwdArgs a;
  a.XH = static_cast<__nv_bfloat16*>(s->XH_hi);
  a.E_all = p->E_all;
  a.HG = s->HG;
  a.G = s->G;
  a.C_all = s->C_all;
  a.H_all = static_cast<__nv_bfloat16*>(s->H_all_hi);
  a.alphas = s->alphas;
  a.dropmask = s->dropmask;
  a.att1 = static_cast<const __nv_bfloat16*>(p->att1_bf);
  a.enc = static_cast<const __nv_bfloat16*>(p->enc_bf);
  a.b_h = s->b_h;
  a.w_f = s->w_f;
  a.b_f = s->b_f;
  a.decode_len = reinterpret_cast<const long long*>(p->decode_len);
  a.HB = static_cast<uint8_t*>(p->scratch);
  a.GA = a.HB + static_cast<size_t>(T + 1) * 8 * lp::CH32;
  a.HGp = reinterpret_cast<float*>(a.GA + static_cast<size_t>(T) * 16 * lp::CH32);
  a.cnt_h = p->counters;
  a.cnt_hg = p->counters + (T + 1);
  a.cnt_awe = p->counters + 2 * (T + 1);
  a.awe_all = p->awe_all;
  a.dbg = reinterpret_cast<long long*>(p->dbg);
  a.B = B;
  a.T = T;
  a.P = s->P;
  if (!lp::configure_once(reinterpret_cast<const void*>(lp::lstm_tf_fwd_persist_kernel), 0)) return CCX_ERR_CUDA;
  ProfScope prof(PROF_LSTM, st, 2.0 * T * B * (double)(lp::AE * lp::D + lp::G4 * (lp::D + lp::E)));
  return lp::launch_coop(lp::lstm_tf_fwd_persist_kernel, lp::N_G2 + lp::N_G1 + B, lp::THREADS,
                         lp::SMEM_BYTES + 1024, st, tm_wh, tm_w2, a) == cudaSuccess ? CCX_OK : CCX_ERR_CUDA;
}

int ccx_lstm_tf_backward_persist(const ccx_lstm_tf* s, const ccx_lstm_tf_bwd* b, const ccx_lstm_persist_bwd* p,
                                 void* stream_) {
  if (s == nullptr || b == nullptr || p == nullptr) return CCX_ERR_SHAPE;
  if (!ccx_lstm_persist_supported(s->B, s->P, s->E, s->A, s->D, s->Emb, s->compute_dtype)) return CCX_ERR_SHAPE;
  if (s->T <= 0) return CCX_OK;
  if (b->dawe_all == nullptr || b->de_all == nullptr || p->awe_all == nullptr) return CCX_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int B = s->B, T = s->T;
  if (p->scratch == nullptr || (reinterpret_cast<uintptr_t>(p->scratch) & 127)) return CCX_ERR_SHAPE;
  CUtensorMap tm_wx, tm_wht;
  int rc;
  if ((rc = lp::make_map_bf16(&tm_wx, p->wx, lp::XW, lp::G4, lp::G4, 128))) return rc;
  if ((rc = lp::make_map_bf16(&tm_wht, p->wht, lp::D, lp::AE, lp::AE, 16))) return rc;
  if (cudaMemsetAsync(p->counters, 0, sizeof(int) * 4 * (T + 1), st) != cudaSuccess) return CCX_ERR_CUDA;
  lp::BwdArgs a;
  a.G = s->G;
  a.C_all = s->C_all;
  a.HG = s->HG;
  a.alphas = s->alphas;
  a.awe_all = p->awe_all;
  a.dH_all = b->dH_all;
  a.dropmask = s->dropmask;
  a.dalphas = b->dalphas;
  a.att1 = static_cast<const __nv_bfloat16*>(p->att1_bf);
  a.enc = static_cast<const __nv_bfloat16*>(p->enc_bf);
  a.w_f = s->w_f;
  a.dG_all = b->dG_all;
  a.dG_bf = static_cast<__nv_bfloat16*>(p->dG_bf);
  a.dHG_all = b->dHG_all;
  a.DGI = static_cast<uint8_t*>(p->scratch);
  a.DHI = a.DGI + static_cast<size_t>(T) * 32 * lp::CH32;
  a.dawe_all = b->dawe_all;
  a.de_all = b->de_all;
  a.d_wf = b->d_wf;
  a.XpA = p->Xp;
  a.XpH = p->Xp + static_cast<size_t>(2) * lp::XA_KS * B * lp::E;
  a.dh_out = b->dh;
  a.dc_out = b->dc;
  a.decode_len = reinterpret_cast<const long long*>(p->decode_len);
  a.cnt_dg = p->counters;
  a.cnt_x = p->counters + (T + 1);
  a.cnt_dhg = p->counters + 2 * (T + 1);
  a.cnt_xh = p->counters + 3 * (T + 1);
  a.dbg = reinterpret_cast<long long*>(p->dbg);
  a.B = B;
  a.T = T;
  a.P = s->P;
  if (!lp::configure_once(reinterpret_cast<const void*>(lp::lstm_tf_bwd_persist_kernel), 1)) return CCX_ERR_CUDA;
  {
    ProfScope prof(PROF_LSTM, st, 2.0 * T * B * (double)(lp::XW * lp::G4 + lp::D * lp::AE));
    if (lp::launch_coop(lp::lstm_tf_bwd_persist_kernel, lp::N_HP + lp::N_XA + lp::N_XH + B, lp::THREADS, lp::SMEM_BYTES + 1024, st,
                        tm_wx, tm_wht, a) != cudaSuccess)
      return CCX_ERR_CUDA;
  }
  // the sums over time into d_enc [B,P,E] and d_att1 [B,P,A], once, from the per-step records
  return attention_bwd_finish(s->alphas, static_cast<long long>(T) * s->P, s->P, b->dawe_all, b->de_all, s->att1, s->HG,
                              lp::AE, s->w_f, b->d_att1, b->d_enc, B, T, s->P, lp::A, lp::E, st);
}

}  // extern "C"
