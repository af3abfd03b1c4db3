// lstm_runner.cu — the teacher-forced LSTM-attention time loop (forward and BPTT) driven from C++ so that the host
// pays ONE FFI call per direction instead of ~10 launches x 51 steps from Python (the loop is launch-latency bound:
// M = active rows <= 32 per GPU).  Same kernels, same order as the per-step Python path in decoder.py /
// decoder_train.py (reference: models/decoder.py:100-111 and its autograd graph).
#include "../../include/ccx.h"

#include "ccx_common.cuh"
#include "ccx_gemm.h"
#include "ccx_ops.h"

using namespace ccx;

static inline const void* off(const void* p, long long elems, int es) {
  return p ? static_cast<const uint8_t*>(p) + elems * es : nullptr;
}
static inline void* offw(void* p, long long elems, int es) {
  return p ? static_cast<uint8_t*>(p) + elems * es : nullptr;
}

extern "C" {

int ccx_lstm_tf_forward(const ccx_lstm_tf* s, void* stream_) {
  if (s == nullptr || s->T < 0) return CCX_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int B = s->B, T = s->T, P = s->P, E = s->E, A = s->A, D = s->D, Emb = s->Emb;
  const int K = Emb + E + D, hoff = Emb + E;
  const int cd = s->compute_dtype;
  const int es = (cd == CCX_BF16) ? 2 : 4;
  const bool f32 = (cd == CCX_F32);
  int rc;
  for (int t = 0; t < T; ++t) {
    const int bt = s->bts_host[t];
    if (bt <= 0) continue;
    const long long xoff = static_cast<long long>(t) * B * K;
    float* hg = s->HG + static_cast<long long>(t) * B * (A + E);
    float* gates = s->G + static_cast<long long>(t) * B * 4 * D;
    // [att2 | gate pre-activation] = h_{t-1} . [decoder_att ; f_beta]^T
    GemmDesc g1;
    g1.A = off(s->XH_hi, xoff + hoff, es); g1.A_lo = f32 ? off(s->XH_lo, xoff + hoff, 4) : nullptr;
    g1.B = s->w_h; g1.B_lo = f32 ? s->w_h_lo : nullptr;
    g1.C = hg; g1.bias = s->b_h;
    g1.lda = K; g1.ldb = D; g1.ldc = A + E;
    g1.M = bt; g1.N = A + E; g1.K = D;
    g1.in_dtype = cd; g1.out_dtype = CCX_F32;
    if ((rc = gemm_tn(g1, st))) return rc;
    if ((rc = bahdanau_attention(s->att1, hg, A + E, s->w_f, s->b_f, s->enc, nullptr,
                                 s->alphas + static_cast<long long>(t) * P, static_cast<long long>(T) * P,
                                 offw(s->XH_hi, xoff + Emb, es), f32 ? s->XH_lo + xoff + Emb : nullptr, cd, K, bt, P, A, E,
                                 1, 1, st)))
      return rc;
    GemmDesc g2;
    g2.A = off(s->XH_hi, xoff, es); g2.A_lo = f32 ? off(s->XH_lo, xoff, 4) : nullptr;
    g2.B = s->w_lstm; g2.B_lo = f32 ? s->w_lstm_lo : nullptr;
    g2.C = gates; g2.bias = s->b_lstm;
    g2.lda = K; g2.ldb = K; g2.ldc = 4 * D;
    g2.M = bt; g2.N = 4 * D; g2.K = K;
    g2.in_dtype = cd; g2.out_dtype = CCX_F32;
    if ((rc = gemm_tn(g2, st))) return rc;
    const long long xnext = static_cast<long long>(t + 1) * B * K + hoff;
    if ((rc = lstm_pointwise(gates, 4 * D, s->C_all + static_cast<long long>(t) * B * D,
                             s->C_all + static_cast<long long>(t + 1) * B * D, offw(s->XH_hi, xnext, es),
                             f32 ? s->XH_lo + xnext : nullptr, K, offw(s->H_all_hi, static_cast<long long>(t) * D, es),
                             f32 ? s->H_all_lo + static_cast<long long>(t) * D : nullptr, static_cast<long long>(T) * D, cd,
                             s->dropmask ? s->dropmask + static_cast<long long>(t) * D : nullptr,
                             static_cast<long long>(T) * D, nullptr, 0, bt, D, st)))
      return rc;
  }
  return CCX_OK;
}

int ccx_lstm_tf_backward(const ccx_lstm_tf* s, const ccx_lstm_tf_bwd* b, void* stream_) {
  if (s == nullptr || b == nullptr) return CCX_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int B = s->B, T = s->T, P = s->P, E = s->E, A = s->A, D = s->D, Emb = s->Emb;
  const int K = Emb + E + D, hoff = Emb + E, AE = A + E;
  const int cd = s->compute_dtype;
  const bool f32 = (cd == CCX_F32);
  int rc;
  // deferred mode: the per-step kernels only record d_awe_raw / d e; the d_enc and d_att1 sums over time are done
  // once after the loop instead of as a read-modify-write pass over both tensors at every step
  const bool deferred = b->dawe_all != nullptr && b->dalpha_all != nullptr && b->de_all != nullptr;
  for (int t = T - 1; t >= 0; --t) {
    const int bt = s->bts_host[t];
    if (bt <= 0) continue;
    float* dG = b->dG_all + static_cast<long long>(t) * B * 4 * D;
    float* dHG = b->dHG_all + static_cast<long long>(t) * B * AE;
    float* dXH = b->dXH_all + static_cast<long long>(t) * B * K;
    // (1) point-wise LSTM backward; also emits dgates as the dgrad GEMM's A operand and clears the attention-backward
    //     accumulator columns of dHG[t] (two launches fewer per step than separate convert / memset)
    if ((rc = lstm_pointwise_bwd(s->G + static_cast<long long>(t) * B * 4 * D, 4 * D,
                                 s->C_all + static_cast<long long>(t) * B * D,
                                 s->C_all + static_cast<long long>(t + 1) * B * D,
                                 b->dH_all + static_cast<long long>(t) * D, static_cast<long long>(T) * D,
                                 s->dropmask ? s->dropmask + static_cast<long long>(t) * D : nullptr,
                                 static_cast<long long>(T) * D, b->dh, b->dc, dG, 4 * D, bt, D, st, b->scratch_hi,
                                 f32 ? b->scratch_lo : nullptr, cd, 4 * D, dHG, AE, P)))
      return rc;
    // (2) [d emb | d awe | d h_prev] = dgates . [W_ih | W_hh]
    GemmDesc g1;
    g1.A = b->scratch_hi; g1.A_lo = f32 ? b->scratch_lo : nullptr;
    g1.B = b->w_lstm_t; g1.B_lo = f32 ? b->w_lstm_t_lo : nullptr;
    g1.C = dXH;
    g1.lda = 4 * D; g1.ldb = 4 * D; g1.ldc = K;
    g1.M = bt; g1.N = K; g1.K = 4 * D;
    g1.in_dtype = cd; g1.out_dtype = CCX_F32;
    if ((rc = gemm_tn(g1, st))) return rc;
    // (3)+(4) attention backward: writes d[att2 | gate] as fp32 (batched wgrad later) and as the next GEMM's operand
    if ((rc = bahdanau_attention_bwd(s->att1, s->HG + static_cast<long long>(t) * B * AE, AE, s->w_f, s->enc,
                                     s->alphas + static_cast<long long>(t) * P, static_cast<long long>(T) * P, dXH + Emb,
                                     K, b->dalphas ? b->dalphas + static_cast<long long>(t) * P : nullptr,
                                     static_cast<long long>(T) * P, dHG, AE, b->d_att1, b->d_enc, b->d_wf, bt, P, A, E,
                                     st, b->scratch2_hi, f32 ? b->scratch2_lo : nullptr, cd, AE, 1,
                                     deferred ? b->dawe_all + static_cast<long long>(t) * B * E : nullptr,
                                     deferred ? b->dalpha_all + static_cast<long long>(t) * B * P : nullptr,
                                     deferred ? b->de_all + static_cast<long long>(t) * B * P : nullptr)))
      return rc;
    // (5) d h_{t-1} = d h_prev (from the gates GEMM) + dHG . [decoder_att ; f_beta]
    GemmDesc g2;
    g2.A = b->scratch2_hi; g2.A_lo = f32 ? b->scratch2_lo : nullptr;
    g2.B = b->w_h_t; g2.B_lo = f32 ? b->w_h_t_lo : nullptr;
    g2.C = b->dh; g2.residual = dXH + hoff;
    g2.lda = AE; g2.ldb = AE; g2.ldc = D; g2.ldr = K;
    g2.M = bt; g2.N = D; g2.K = AE;
    g2.in_dtype = cd; g2.out_dtype = CCX_F32;
    if ((rc = gemm_tn(g2, st))) return rc;
  }
  if (deferred)
    return attention_bwd_finish(s->alphas, static_cast<long long>(T) * P, P, b->dawe_all, b->de_all, s->att1, s->HG,
                                AE, s->w_f, b->d_att1, b->d_enc, B, T, P, A, E, st);
  return CCX_OK;
}

}  // extern "C"
