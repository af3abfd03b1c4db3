// ccx_api.cu — extern "C" surface of libccx.so (declared in include/ccx.h) and the whole-encoder runner.
#include "../../include/ccx.h"

#include "ccx_common.cuh"
#include "ccx_gemm.h"
#include "ccx_ops.h"

using namespace ccx;

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int ccx_version(void) { return 100; }

const char* ccx_status_string(int status) {
  switch (status) {
    case CCX_OK: return "ok";
    case CCX_ERR_SHAPE: return "unsupported shape / alignment (no fallback path exists)";
    case CCX_ERR_DTYPE: return "unsupported dtype combination";
    case CCX_ERR_CUDA: {
      cudaError_t e = cudaGetLastError();
      return e == cudaSuccess ? "CUDA launch/config error" : cudaGetErrorString(e);
    }
    case CCX_ERR_TMA: return "cuTensorMapEncodeTiled failed";
    case CCX_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown ccx status";
  }
}

int ccx_num_sms(void) { return num_sms(); }

int ccx_linear(const ccx_linear_desc* d, void* stream) {
  if (d == nullptr) return CCX_ERR_SHAPE;
  GemmDesc g;
  g.A = d->A; g.A_lo = d->A_lo; g.B = d->W; g.B_lo = d->W_lo;
  g.C = d->C; g.C_lo = d->C_lo;
  g.bias = d->bias; g.colscale = d->colscale; g.rowscale = d->rowscale; g.residual = d->residual;
  g.emask = d->emask; g.ldm = d->ldm;
  g.lda = d->lda; g.ldb = d->ldw; g.ldc = d->ldc; g.ldr = d->ldr;
  g.M = d->M; g.N = d->N; g.K = d->K;
  g.rows_per_group = d->rows_per_group;
  g.act = d->act; g.in_dtype = d->in_dtype; g.out_dtype = d->out_dtype; g.split = d->split;
  g.a_mn = d->a_mn != 0; g.b_mn = d->w_mn != 0;
  g.res_mul = d->res_mul != 0;
  if (g.res_mul && (g.residual == nullptr || g.colscale != nullptr || g.rowscale != nullptr || g.split)) return CCX_ERR_SHAPE;
  return gemm_tn(g, as_stream(stream));
}

int ccx_set_sm_limit(int32_t n) {
  set_sm_limit(n);
  return CCX_OK;
}

int ccx_set_gemm_pair_mode(int32_t mode) {
  set_gemm_pair_mode(mode);
  return CCX_OK;
}

int ccx_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
  return split_tf32(x, hi, lo, n, as_stream(stream));
}
int ccx_cast_bf16(const float* x, void* y, int64_t n, void* stream) {
  return cast_bf16(x, y, n, as_stream(stream));
}

int ccx_stem_ln(const float* images, const float* w_k, const float* bias, const float* ln_g, const float* ln_b,
                float* out, int32_t B, int32_t Hin, int32_t Win, float eps, void* stream) {
  return stem_ln(images, nullptr, nullptr, nullptr, w_k, bias, ln_g, ln_b, out, B, Hin, Win, eps, as_stream(stream));
}

int ccx_stem_ln_u8(const uint8_t* images_u8, const float* mean3, const float* inv_std3, const float* w_k,
                   const float* bias, const float* ln_g, const float* ln_b, float* out, int32_t B, int32_t Hin,
                   int32_t Win, float eps, void* stream) {
  return stem_ln(nullptr, images_u8, mean3, inv_std3, w_k, bias, ln_g, ln_b, out, B, Hin, Win, eps,
                 as_stream(stream));
}

int ccx_dwconv7_ln(const float* x, const float* w_tap_major, const float* bias, const float* ln_g,
                   const float* ln_b, void* out, float* out_lo, int32_t B, int32_t H, int32_t W, int32_t C,
                   float eps, int32_t out_dtype, void* stream) {
  return dwconv7_ln(x, w_tap_major, bias, ln_g, ln_b, out, out_lo, B, H, W, C, eps, out_dtype, as_stream(stream));
}

int ccx_dwconv7_plain(const float* x, const float* w_tap_major, const float* bias, const float* addend, float* out,
                      int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  return dwconv7_ln(x, w_tap_major, bias, nullptr, nullptr, out, nullptr, B, H, W, C, 0.f, CCX_F32, as_stream(stream),
                    addend);
}
int ccx_scale_rows_cols(const float* x, const float* colscale, const float* rowscale, int32_t rows_per_group,
                        float* out, int64_t M, int32_t C, void* stream) {
  return scale_rows_cols(x, colscale, rowscale, rows_per_group, out, M, C, as_stream(stream));
}
int ccx_gelu_bwd(const float* pre, float* dh, int64_t n, void* stream) { return gelu_bwd(pre, dh, n, as_stream(stream)); }
int ccx_cnblock_param_grads(const float* G, const float* W2, const float* b2, const float* gamma, const float* s,
                            float* dW2, float* dgamma, float* db2, int32_t C, int32_t K, void* stream) {
  return cnblock_param_grads(G, W2, b2, gamma, s, dW2, dgamma, db2, C, K, as_stream(stream));
}
int ccx_dwconv7_wgrad(const float* x, const float* du, float* dw_tap_major, int32_t B, int32_t H, int32_t W,
                      int32_t C, void* stream) {
  return dwconv7_wgrad(x, du, dw_tap_major, B, H, W, C, as_stream(stream));
}
int ccx_avgpool_nhwc_bwd(const float* dout, float* dx, int32_t B, int32_t H, int32_t W, int32_t C, int32_t S,
                         void* stream) {
  return avgpool_nhwc_bwd(dout, dx, B, H, W, C, S, as_stream(stream));
}

int ccx_ln_rows(const float* x, const float* ln_g, const float* ln_b, void* out, float* out_lo, float* out_plain,
                int64_t M, int32_t C, float eps, int32_t out_dtype, int32_t merge, int32_t H, int32_t W,
                void* stream) {
  return ln_rows(x, ln_g, ln_b, out, out_lo, out_plain, M, C, eps, out_dtype, merge, H, W, as_stream(stream));
}

int ccx_embed_rows(const int64_t* tokens, int64_t tok_ld, int32_t t0, const float* table, int32_t V, int32_t D,
                   const float* pe, const float* dropmask, float* out_plain, int64_t sb_p, int64_t st_p,
                   void* op_hi, float* op_lo, int32_t op_dtype, int64_t sb_o, int64_t st_o, int32_t nb, int32_t nt,
                   void* stream) {
  return embed_rows(reinterpret_cast<const long long*>(tokens), tok_ld, t0, table, V, D, pe, dropmask, out_plain,
                    sb_p, st_p, op_hi, op_lo, op_dtype, sb_o, st_o, nb, nt, as_stream(stream));
}

int ccx_mean_pixels(const float* enc, int32_t B, int32_t P, int32_t E, void* op_hi, float* op_lo, int32_t op_dtype,
                    int64_t ldo, void* stream) {
  return mean_pixels(enc, B, P, E, op_hi, op_lo, op_dtype, ldo, as_stream(stream));
}

int ccx_bahdanau_attention(const float* att1, const float* hg, int64_t ldhg, const float* w_f, const float* b_f,
                           const float* enc, const float* active, float* alpha_out, int64_t alpha_ld,
                           void* awe_hi, float* awe_lo, int32_t awe_dtype, int64_t ld_awe, int32_t bt, int32_t P,
                           int32_t A, int32_t E, int32_t apply_gate, int32_t enc_group, void* stream) {
  return bahdanau_attention(att1, hg, ldhg, w_f, b_f, enc, active, alpha_out, alpha_ld, awe_hi, awe_lo, awe_dtype,
                            ld_awe, bt, P, A, E, apply_gate, enc_group, as_stream(stream));
}

int ccx_lstm_pointwise(const float* gates, int64_t ldg, const float* c_prev, float* c_new, void* hn_hi,
                       float* hn_lo, int64_t ld_hn, void* ha_hi, float* ha_lo, int64_t ld_ha, int32_t op_dtype,
                       const float* dropmask, int64_t ld_dm, float* h_plain, int64_t ld_hp, int32_t bt, int32_t D,
                       void* stream) {
  return lstm_pointwise(gates, ldg, c_prev, c_new, hn_hi, hn_lo, ld_hn, ha_hi, ha_lo, ld_ha, op_dtype, dropmask,
                        ld_dm, h_plain, ld_hp, bt, D, as_stream(stream));
}

int ccx_greedy_next(const float* preds, int64_t ld_preds, int32_t B, int32_t V, int32_t t, int32_t T,
                    int64_t* sequences, float* active, int64_t* next_tok, int64_t ld_next, int64_t end_token,
                    void* stream) {
  return greedy_next(preds, ld_preds, B, V, t, T, reinterpret_cast<long long*>(sequences), active,
                     reinterpret_cast<long long*>(next_tok), ld_next, end_token, as_stream(stream));
}

int ccx_mha_small(const float* q, int64_t q_sb, int64_t q_st, const float* k, int64_t k_sb, int64_t k_st,
                  const float* v, int64_t v_sb, int64_t v_st, void* ctx_hi, float* ctx_lo, int32_t ctx_dtype,
                  int64_t c_sb, int64_t c_st, const uint8_t* key_pad, const float* prob_mask, float* probs_out,
                  int32_t B, int32_t H, int32_t Tq, int32_t Tk, int32_t hd, int32_t causal, int32_t q_pos0,
                  float scale, int32_t kv_group, void* stream) {
  return mha_small(q, q_sb, q_st, k, k_sb, k_st, v, v_sb, v_st, ctx_hi, ctx_lo, ctx_dtype, c_sb, c_st, key_pad,
                   prob_mask, probs_out, B, H, Tq, Tk, hd, causal, q_pos0, scale, kv_group, as_stream(stream));
}

int ccx_beam_topk(const float* logits, int64_t ld, int32_t NI, int32_t k, int32_t V, const float* top_scores,
                  const int32_t* k_rem, int32_t first_step, float* cand_score, int32_t* cand_prev,
                  int32_t* cand_word, void* stream) {
  return beam_topk(logits, ld, NI, k, V, top_scores, k_rem, first_step, cand_score, cand_prev, cand_word,
                   as_stream(stream));
}

int ccx_beam_update(int32_t NI, int32_t k, int32_t Tcap, int32_t step, int64_t end_token, const float* cand_score,
                    const int32_t* cand_prev, const int32_t* cand_word, const int64_t* seqs_in, int64_t* seqs_out,
                    float* top_scores, int32_t* k_rem, int64_t* done_seqs, float* done_scores, int32_t* done_len,
                    int32_t* n_done, int32_t* src_row, int64_t* next_tok, int64_t ld_next, int32_t* done_parent,
                    void* stream) {
  return beam_update(NI, k, Tcap, step, end_token, cand_score, cand_prev, cand_word,
                     reinterpret_cast<const long long*>(seqs_in), reinterpret_cast<long long*>(seqs_out), top_scores,
                     k_rem, reinterpret_cast<long long*>(done_seqs), done_scores, done_len, n_done, src_row,
                     reinterpret_cast<long long*>(next_tok), ld_next, done_parent, as_stream(stream));
}

int ccx_gather_rows(const void* src, int64_t src_stride_bytes, void* dst, int64_t dst_stride_bytes,
                    const int32_t* src_row, int64_t row_bytes, int32_t rows, void* stream) {
  return gather_rows(src, src_stride_bytes, dst, dst_stride_bytes, src_row, row_bytes, rows, as_stream(stream));
}

int ccx_convert_operand(const void* x_hi, const float* x_lo, int32_t x_dtype, int64_t ldx, const float* mul,
                        int64_t ldm, int32_t mul_mode, float mul_scale, void* o_hi, float* o_lo, int32_t o_dtype,
                        int64_t ldo, int32_t R, int32_t C, int32_t transpose, int32_t Rpad, void* stream) {
  return convert_operand(x_hi, x_lo, x_dtype, ldx, mul, ldm, mul_mode, mul_scale, o_hi, o_lo, o_dtype, ldo, R, C,
                         transpose, Rpad, as_stream(stream));
}
int ccx_colsum_acc(const float* x, int64_t ldx, const float* mul, int64_t ldm, int32_t mul_mode, float mul_scale,
                   float* out, int32_t R, int32_t C, void* stream) {
  return colsum_acc(x, ldx, mul, ldm, mul_mode, mul_scale, out, R, C, as_stream(stream));
}
int ccx_stream_capture_status(void* stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(as_stream(stream), &st) != cudaSuccess) return CCX_ERR_CUDA;
  return static_cast<int>(st);
}
int ccx_convert_colsum(const float* x, int64_t ldx, const float* mul, int64_t ldm, int32_t mul_mode, float mul_scale,
                       void* o_bf16, int64_t ldo, float* sums, int32_t R, int32_t C, void* stream) {
  return convert_colsum(x, ldx, mul, ldm, mul_mode, mul_scale, o_bf16, ldo, sums, R, C, as_stream(stream));
}
int ccx_cast_segments(const ccx_cast_seg* segs_dev, int32_t nseg, int32_t total_tiles, double bytes, void* stream) {
  return cast_segments(segs_dev, nseg, total_tiles, bytes, as_stream(stream));
}
int ccx_ln_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
               int64_t M, int32_t C, float eps, int32_t merge, int32_t H, int32_t W, void* stream) {
  if (merge && ((H & 1) || (W & 1) || H <= 0 || W <= 0)) return CCX_ERR_SHAPE;
  return ln_bwd(dy, x, gamma, dx, dgamma, dbeta, M, C, eps, as_stream(stream), merge, H, W);
}
int ccx_mha_bwd(const float* q, int64_t q_sb, int64_t q_st, const float* k, int64_t k_sb, int64_t k_st,
                const float* v, int64_t v_sb, int64_t v_st, const float* dctx, int64_t d_sb, int64_t d_st,
                const float* probs, const float* prob_mask, float* dq, int64_t dq_sb, int64_t dq_st, float* dk,
                int64_t dk_sb, int64_t dk_st, float* dv, int64_t dv_sb, int64_t dv_st, int32_t B, int32_t H,
                int32_t Tq, int32_t Tk, int32_t hd, float scale, void* stream) {
  return mha_bwd(q, q_sb, q_st, k, k_sb, k_st, v, v_sb, v_st, dctx, d_sb, d_st, probs, prob_mask, dq, dq_sb, dq_st,
                 dk, dk_sb, dk_st, dv, dv_sb, dv_st, B, H, Tq, Tk, hd, scale, as_stream(stream));
}

int ccx_mha_bwd_tc(const float* q, int64_t q_sb, int64_t q_st, const float* k, int64_t k_sb, int64_t k_st,
                const float* v, int64_t v_sb, int64_t v_st, const float* dctx, int64_t d_sb, int64_t d_st,
                const float* probs, const float* prob_mask, float* dq, int64_t dq_sb, int64_t dq_st, float* dk,
                int64_t dk_sb, int64_t dk_st, float* dv, int64_t dv_sb, int64_t dv_st, int32_t B, int32_t H,
                int32_t Tq, int32_t Tk, int32_t hd, float scale, void* stream) {
  if (!mha_tc_eligible(Tq, Tk, hd)) return CCX_ERR_SHAPE;
  return mha_tc_bwd(q, q_sb, q_st, k, k_sb, k_st, v, v_sb, v_st, dctx, d_sb, d_st, probs, prob_mask, dq, dq_sb, dq_st,
                    dk, dk_sb, dk_st, dv, dv_sb, dv_st, B, H, Tq, Tk, scale, as_stream(stream));
}
int ccx_softmax_ce(const float* logits, int64_t ld, const int64_t* targets, int64_t R, int32_t V, float inv_n,
                   float* loss_sum, float* dlogits, int64_t ldd, float* stats, int32_t topk, void* stream) {
  return softmax_ce(logits, ld, reinterpret_cast<const long long*>(targets), R, V, inv_n, loss_sum, dlogits, ldd,
                    stats, topk, as_stream(stream));
}
int ccx_softmax_ce_dev(const float* logits, int64_t ld, const int64_t* targets, int64_t R, int32_t V,
                       const float* n_valid_dev, float* loss_sum, float* dlogits, int64_t ldd, float* stats,
                       int32_t topk, void* stream) {
  if (n_valid_dev == nullptr) return CCX_ERR_SHAPE;
  return softmax_ce(logits, ld, reinterpret_cast<const long long*>(targets), R, V, 1.0f, loss_sum, dlogits, ldd,
                    stats, topk, as_stream(stream), n_valid_dev);
}
int ccx_free_running_targets(const int64_t* sequences, const int64_t* caps, int64_t cap_ld, int64_t* targets,
                             int32_t* decode_len, int32_t B, int32_t T, int32_t cap_T, int64_t end_tok,
                             int64_t pad_tok, void* stream) {
  return free_running_targets(reinterpret_cast<const long long*>(sequences), reinterpret_cast<const long long*>(caps),
                              cap_ld, reinterpret_cast<long long*>(targets), decode_len, B, T, cap_T, end_tok, pad_tok,
                              as_stream(stream));
}
int ccx_embedding_bwd(const int64_t* tokens, int64_t tok_ld, int32_t t0, const float* dx, int64_t sb, int64_t st,
                      const float* dropmask, float* dtable, int32_t V, int32_t D, int32_t nb, int32_t nt,
                      void* stream) {
  return embedding_bwd(reinterpret_cast<const long long*>(tokens), tok_ld, t0, dx, sb, st, dropmask, dtable, V, D,
                       nb, nt, as_stream(stream));
}
int ccx_lstm_pointwise_bwd(const float* gates, int64_t ldg, const float* c_prev, const float* c_new,
                           const float* dh_fc, int64_t ld_fc, const float* dropmask, int64_t ld_dm,
                           const float* dh_carry, float* dc_carry, float* dgates, int64_t lddg, int32_t bt,
                           int32_t D, void* stream) {
  return lstm_pointwise_bwd(gates, ldg, c_prev, c_new, dh_fc, ld_fc, dropmask, ld_dm, dh_carry, dc_carry, dgates,
                            lddg, bt, D, as_stream(stream));
}
int ccx_bahdanau_attention_bwd(const float* att1, const float* hg, int64_t ldhg, const float* w_f, const float* enc,
                               const float* alpha, int64_t alpha_ld, const float* d_out, int64_t ld_dout,
                               const float* d_alpha_ext, int64_t dalpha_ld, float* d_hg, int64_t ld_dhg,
                               float* d_att1, float* d_enc, float* d_wf, int32_t bt, int32_t P, int32_t A,
                               int32_t E, void* stream) {
  return bahdanau_attention_bwd(att1, hg, ldhg, w_f, enc, alpha, alpha_ld, d_out, ld_dout, d_alpha_ext, dalpha_ld,
                                d_hg, ld_dhg, d_att1, d_enc, d_wf, bt, P, A, E, as_stream(stream));
}
int ccx_bcast_add_rows(float* out, const float* v, float scale, int32_t B, int32_t P, int32_t E, void* stream) {
  return bcast_add_rows(out, v, scale, B, P, E, as_stream(stream));
}
int ccx_adam_clamp(const void* table, const int32_t* block_entry, const int64_t* block_offset, int32_t n_blocks,
                   float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float clip,
                   int32_t chunk, double total_params, void* stream) {
  return adam_clamp(table, block_entry, reinterpret_cast<const long long*>(block_offset), n_blocks, lr, beta1, beta2,
                    eps, bc1, bc2_sqrt, clip, chunk, total_params, as_stream(stream));
}

int ccx_adam_clamp_dev(const void* table, const int32_t* block_entry, const int64_t* block_offset, int32_t n_blocks,
                       float lr, float beta1, float beta2, float eps, const float* step_dev, float clip,
                       int32_t chunk, double total_params, void* stream) {
  if (step_dev == nullptr) return CCX_ERR_SHAPE;
  return adam_clamp(table, block_entry, reinterpret_cast<const long long*>(block_offset), n_blocks, lr, beta1, beta2,
                    eps, 1.f, 1.f, clip, chunk, total_params, as_stream(stream), step_dev);
}

int ccx_attn_head_mean(const float* probs, int64_t p_sb, int64_t p_sh, int64_t p_st, const float* prob_mask,
                       const float* row_active, float* alphas, int64_t a_sb, int64_t a_st, int32_t B, int32_t H,
                       int32_t Tq, int32_t Tk, float scale, int32_t accumulate, void* stream) {
  return attn_head_mean(probs, p_sb, p_sh, p_st, prob_mask, row_active, alphas, a_sb, a_st, B, H, Tq, Tk, scale,
                        accumulate, static_cast<cudaStream_t>(stream));
}

int ccx_mha_decode(const float* q, int64_t q_sb, const float* k, int64_t k_sb, int64_t k_st, const float* v,
                   int64_t v_sb, int64_t v_st, void* ctx_hi, float* ctx_lo, int32_t ctx_dtype, int64_t c_sb,
                   const int32_t* kv_rows, int64_t ld_map, int32_t rows, int32_t H, int32_t Tk, int32_t hd,
                   int32_t kv_group, float scale, void* stream) {
  return mha_decode(q, q_sb, k, k_sb, k_st, v, v_sb, v_st, ctx_hi, ctx_lo, ctx_dtype, c_sb, kv_rows, ld_map, rows, H,
                    Tk, hd, kv_group, scale, as_stream(stream));
}

int ccx_avgpool_nhwc(const float* x, float* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t S,
                     void* stream) {
  return avgpool_nhwc(x, out, B, H, W, C, S, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
// whole-encoder runner
// ------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t v) { return (v + 1023) & ~size_t(1023); }

struct EncoderScratch {
  size_t x_bytes, y_bytes, h_bytes;
  size_t total;
};
static EncoderScratch encoder_scratch(int B, int Hin, int Win, int compute_dtype) {
  EncoderScratch s;
  const size_t px = static_cast<size_t>(B) * (Hin / 4) * (Win / 4);
  const size_t es = (compute_dtype == CCX_BF16) ? 2 : 8;  // bf16, or fp32 hi + fp32 lo
  s.x_bytes = align_up(px * 128 * 4);
  s.y_bytes = align_up(px * 128 * es);
  s.h_bytes = align_up(px * 512 * es);
  s.total = 2 * s.x_bytes + s.y_bytes + s.h_bytes + 1024;
  return s;
}

size_t ccx_encoder_workspace_bytes(int32_t B, int32_t Hin, int32_t Win, int32_t compute_dtype) {
  if (B <= 0 || Hin < 32 || Win < 32) return 0;
  return encoder_scratch(B, Hin, Win, compute_dtype).total;
}

int ccx_encoder_run(const ccx_encoder_weights* w, const float* in, float* out, int32_t B, int32_t Hin,
                    int32_t Win, int32_t child_begin, int32_t child_end, const float* sd_rowscale,
                    void* workspace, size_t workspace_bytes, void* stream_) {
  if (w == nullptr || in == nullptr || out == nullptr) return CCX_ERR_SHAPE;
  if (child_begin < 0 || child_end > 8 || child_begin >= child_end) return CCX_ERR_SHAPE;
  if (B <= 0 || Hin < 32 || Win < 32 || (Hin % 32) != 0 || (Win % 32) != 0) return CCX_ERR_SHAPE;
  const int cd = w->compute_dtype;
  if (cd != CCX_F32 && cd != CCX_BF16) return CCX_ERR_DTYPE;
  cudaStream_t stream = as_stream(stream_);
  const EncoderScratch sc = encoder_scratch(B, Hin, Win, cd);
  if (workspace == nullptr || workspace_bytes < sc.total) return CCX_ERR_WORKSPACE;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  float* X0 = reinterpret_cast<float*>(base);
  float* X1 = reinterpret_cast<float*>(base + sc.x_bytes);
  uint8_t* Y = base + 2 * sc.x_bytes;
  uint8_t* Hb = Y + sc.y_bytes;

  const bool f32 = (cd == CCX_F32);
  const float* cur = in;
  int rc;
  int blk_base[4];
  {
    int acc = 0;
    for (int s = 0; s < 4; ++s) { blk_base[s] = acc; acc += w->depths[s]; }
    if (acc > CCX_MAX_BLOCKS) return CCX_ERR_SHAPE;
  }

  for (int child = child_begin; child < child_end; ++child) {
    const bool last_child = (child == child_end - 1);
    if (child == 0) {
      float* dst = last_child ? out : X0;
      if ((rc = stem_ln(cur, nullptr, nullptr, nullptr, w->stem_w, w->stem_b, w->stem_ln_g, w->stem_ln_b, dst, B, Hin,
                        Win, 1e-6f, stream)))
        return rc;
      cur = dst;
      continue;
    }
    const int stage = (child - 1) / 2;  // children 1,3,5,7 are stages 0..3 ; 2,4,6 downsample into stage 1..3
    if (child & 1) {
      const int C = w->dims[stage];
      const int H = Hin >> (2 + stage), W = Win >> (2 + stage);
      const long long M = static_cast<long long>(B) * H * W;
      if (M > 0x7fffffffLL) return CCX_ERR_SHAPE;
      float* y_hi = reinterpret_cast<float*>(Y);
      float* y_lo = f32 ? y_hi + M * C : nullptr;
      float* h_hi = reinterpret_cast<float*>(Hb);
      float* h_lo = f32 ? h_hi + M * 4 * C : nullptr;
      const int nblk = w->depths[stage];
      for (int i = 0; i < nblk; ++i) {
        const ccx_cnblock_weights& bw = w->blocks[blk_base[stage] + i];
        const bool last = last_child && (i == nblk - 1);
        float* dst;
        if (last) dst = out;
        else if (cur == X0 || cur == X1) dst = const_cast<float*>(cur);
        else dst = X0;
        if ((rc = dwconv7_ln(cur, bw.dw_w, bw.dw_b, bw.ln_g, bw.ln_b, Y, y_lo, B, H, W, C, 1e-6f, cd, stream)))
          return rc;
        GemmDesc g1;
        g1.A = Y; g1.A_lo = y_lo; g1.B = bw.w1; g1.B_lo = f32 ? bw.w1_lo : nullptr;
        g1.C = Hb; g1.C_lo = h_lo;
        g1.bias = bw.b1;
        g1.lda = C; g1.ldb = C; g1.ldc = 4 * C;
        g1.M = static_cast<int>(M); g1.N = 4 * C; g1.K = C;
        g1.act = CCX_ACT_GELU;
        g1.in_dtype = cd; g1.out_dtype = cd; g1.split = f32 ? 1 : 0;
        if ((rc = gemm_tn(g1, stream))) return rc;
        GemmDesc g2;
        g2.A = Hb; g2.A_lo = h_lo; g2.B = bw.w2; g2.B_lo = f32 ? bw.w2_lo : nullptr;
        g2.C = dst;
        g2.bias = bw.b2; g2.colscale = bw.layer_scale;
        g2.rowscale = sd_rowscale ? sd_rowscale + static_cast<size_t>(blk_base[stage] + i) * B : nullptr;
        g2.rows_per_group = H * W;
        g2.residual = cur;
        g2.lda = 4 * C; g2.ldb = 4 * C; g2.ldc = C; g2.ldr = C;
        g2.M = static_cast<int>(M); g2.N = C; g2.K = 4 * C;
        g2.in_dtype = cd; g2.out_dtype = CCX_F32;
        if ((rc = gemm_tn(g2, stream))) return rc;
        cur = dst;
      }
    } else {
      // downsample child 2/4/6: LN2d over Cin + 2x2/s2 conv as patch-merge GEMM, into stage (child/2)
      const int sin = child / 2 - 1;
      const int Cin = w->dims[sin], Cout = w->dims[sin + 1];
      const int H = Hin >> (2 + sin), W = Win >> (2 + sin);
      const long long M = static_cast<long long>(B) * H * W;
      const ccx_downsample_weights& dw = w->down[sin];
      float* y_hi = reinterpret_cast<float*>(Y);
      float* y_lo = f32 ? y_hi + M * Cin : nullptr;
      if ((rc = ln_rows(cur, dw.ln_g, dw.ln_b, Y, y_lo, nullptr, M, Cin, 1e-6f, cd, 1, H, W, stream))) return rc;
      float* dst = last_child ? out : ((cur == X0) ? X1 : X0);
      GemmDesc g;
      g.A = Y; g.A_lo = y_lo; g.B = dw.w; g.B_lo = f32 ? dw.w_lo : nullptr;
      g.C = dst;
      g.bias = dw.b;
      g.lda = 4 * Cin; g.ldb = 4 * Cin; g.ldc = Cout;
      g.M = static_cast<int>(M / 4); g.N = Cout; g.K = 4 * Cin;
      g.in_dtype = cd; g.out_dtype = CCX_F32;
      if ((rc = gemm_tn(g, stream))) return rc;
      cur = dst;
    }
  }
  return CCX_OK;
}

}  // extern "C"
