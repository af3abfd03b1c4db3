// ccx_ops.h — internal launcher prototypes (C++ side of libccx; the public C ABI is include/ccx.h).
#pragma once
#include <cuda_runtime.h>

namespace ccx {

// dwconv_ln.cu
int dwconv7_ln(const float* x, const float* w49c, const float* bias, const float* gamma, const float* beta,
               void* out, float* out_lo, int B, int H, int W, int C, float eps, int out_dtype,
               cudaStream_t stream, const float* addend = nullptr);

// encoder_misc.cu
int stem_ln(const float* img, const unsigned char* img_u8, const float* mean, const float* inv_std, const float* wk,
            const float* bias, const float* gamma, const float* beta, float* out, int B, int Hin, int Win, float eps,
            cudaStream_t stream);
int ln_rows(const float* x, const float* gamma, const float* beta, void* out, float* out_lo, float* out_plain,
            long long M, int C, float eps, int out_dtype, int merge, int H, int W, cudaStream_t stream);
int avgpool_nhwc(const float* x, float* out, int B, int H, int W, int C, int S, cudaStream_t stream);
int split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t stream);
int cast_bf16(const float* x, void* y, long long n, cudaStream_t stream);

// decoder_kernels.cu
int embed_rows(const long long* tokens, long long tok_ld, int t0, const float* table, int V, int D,
               const float* pe, const float* dropmask, float* out_plain, long long sb_p, long long st_p,
               void* op_hi, float* op_lo, int op_dtype, long long sb_o, long long st_o, int nb, int nt,
               cudaStream_t stream);
int mean_pixels(const float* enc, int B, int P, int E, void* op_hi, float* op_lo, int op_dtype, long long ldo,
                cudaStream_t stream);
int bahdanau_attention(const float* att1, const float* hg, long long ldhg, const float* w_f, const float* b_f,
                       const float* enc, const float* active, float* alpha_out, long long alpha_ld, void* awe_hi,
                       float* awe_lo, int awe_dtype, long long ld_awe, int bt, int P, int A, int E,
                       int apply_gate, int enc_group, cudaStream_t stream);
int lstm_pointwise(const float* gates, long long ldg, const float* c_prev, float* c_new, void* hn_hi, float* hn_lo,
                   long long ld_hn, void* ha_hi, float* ha_lo, long long ld_ha, int op_dtype, const float* dropmask,
                   long long ld_dm, float* h_plain, long long ld_hp, int bt, int D, cudaStream_t stream);
int greedy_next(const float* preds, long long ld_preds, int B, int V, int t, int T, long long* sequences,
                float* active, long long* next_tok, long long ld_next, long long end_token, cudaStream_t stream);
int mha_small(const float* q, long long q_sb, long long q_st, const float* k, long long k_sb, long long k_st,
              const float* v, long long v_sb, long long v_st, void* ctx_hi, float* ctx_lo, int ctx_dtype,
              long long c_sb, long long c_st, const unsigned char* key_pad, const float* prob_mask,
              float* probs_out, int B, int H, int Tq, int Tk, int hd, int causal, int q_pos0, float scale,
              int kv_group, cudaStream_t stream);

int attn_head_mean(const float* probs, long long p_sb, long long p_sh, long long p_st, const float* mask,
                   const float* row_active, float* alphas, long long a_sb, long long a_st, int B, int H, int Tq, int Tk,
                   float scale, int accumulate, cudaStream_t stream);
int mha_decode(const float* q, long long q_sb, const float* k, long long k_sb, long long k_st, const float* v,
               long long v_sb, long long v_st, void* ctx_hi, float* ctx_lo, int ctx_dtype, long long c_sb,
               const int* kv_rows, long long ld_map, int rows, int H, int Tk, int hd, int kv_group, float scale,
               cudaStream_t stream);

// beam_kernels.cu
int beam_topk(const float* logits, long long ld, int NI, int k, int V, const float* top_scores, const int* k_rem,
              int first_step, float* cand_score, int* cand_prev, int* cand_word, cudaStream_t stream);
int beam_update(int NI, int k, int Tcap, int step, long long end_token, const float* cand_score,
                const int* cand_prev, const int* cand_word, const long long* seqs_in, long long* seqs_out,
                float* top_scores, int* k_rem, long long* done_seqs, float* done_scores, int* done_len, int* n_done,
                int* src_row, long long* next_tok, long long ld_next, int* done_parent, cudaStream_t stream);
int gather_rows(const void* src, long long src_stride, void* dst, long long dst_stride, const int* src_row,
                long long row_bytes, int rows, cudaStream_t stream);

// train_kernels.cu
int convert_operand(const void* x_hi, const float* x_lo, int x_dtype, long long ldx, const float* mul,
                    long long ldm, int mul_mode, float mul_scale, void* o_hi, float* o_lo, int o_dtype, long long ldo,
                    int R, int C, int transpose, int Rpad, cudaStream_t stream);
int colsum_acc(const float* x, long long ldx, const float* mul, long long ldm, int mul_mode, float mul_scale,
               float* out, int R, int C, cudaStream_t stream);
int cast_segments(const void* segs_dev, int nseg, int total_tiles, double bytes, cudaStream_t stream);
int convert_colsum(const float* x, long long ldx, const float* mul, long long ldm, int mul_mode, float mul_scale,
                   void* o_bf16, long long ldo, float* sums, int R, int C, cudaStream_t stream);
int ln_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
           long long M, int C, float eps, cudaStream_t stream, int merge = 0, int H = 1, int W = 1);
// tensor-core (mma.sync, bf16) versions for T <= 64 and head dim 64: mha_tc.cu
bool mha_tc_eligible(int Tq, int Tk, int hd);
int mha_tc_fwd(const float* q, long long q_sb, long long q_st, const float* k, long long k_sb, long long k_st,
               const float* v, long long v_sb, long long v_st, void* ctx_bf16, long long c_sb, long long c_st,
               const unsigned char* key_pad, const float* prob_mask, float* probs_out, int B, int H, int Tq, int Tk,
               int causal, int q_pos0, float scale, int kv_group, cudaStream_t stream);
int mha_tc_bwd(const float* q, long long q_sb, long long q_st, const float* k, long long k_sb, long long k_st,
               const float* v, long long v_sb, long long v_st, const float* dctx, long long d_sb, long long d_st,
               const float* probs, const float* prob_mask, float* dq, long long dq_sb, long long dq_st, float* dk,
               long long dk_sb, long long dk_st, float* dv, long long dv_sb, long long dv_st, int B, int H, int Tq,
               int Tk, float scale, cudaStream_t stream);
int mha_bwd(const float* q, long long q_sb, long long q_st, const float* k, long long k_sb, long long k_st,
            const float* v, long long v_sb, long long v_st, const float* dctx, long long d_sb, long long d_st,
            const float* probs, const float* prob_mask, float* dq, long long dq_sb, long long dq_st, float* dk,
            long long dk_sb, long long dk_st, float* dv, long long dv_sb, long long dv_st, int B, int H, int Tq,
            int Tk, int hd, float scale, cudaStream_t stream);
int softmax_ce(const float* logits, long long ld, const long long* targets, long long R, int V, float inv_n,
               float* loss_sum, float* dlogits, long long ldd, float* stats, int topk, cudaStream_t stream,
               const float* n_valid_dev = nullptr);
int free_running_targets(const long long* sequences, const long long* caps, long long cap_ld, long long* targets,
                         int* decode_len, int B, int T, int cap_T, long long end_tok, long long pad_tok,
                         cudaStream_t stream);
int embedding_bwd(const long long* tokens, long long tok_ld, int t0, const float* dx, long long sb, long long st,
                  const float* dropmask, float* dtable, int V, int D, int nb, int nt, cudaStream_t stream);
int lstm_pointwise_bwd(const float* gates, long long ldg, const float* c_prev, const float* c_new,
                       const float* dh_fc, long long ld_fc, const float* dropmask, long long ld_dm,
                       const float* dh_carry, float* dc_carry, float* dgates, long long lddg, int bt, int D,
                       cudaStream_t stream, void* dg_op_hi = nullptr, float* dg_op_lo = nullptr, int dg_op_dtype = 0,
                       long long ld_op = 0, float* zero_rows = nullptr, long long ld_zero = 0, int n_zero = 0);
int attention_bwd_finish(const float* alphas, long long a_sb, long long a_st, const float* dawe_all,
                         const float* de_all, const float* att1, const float* hg_all, long long ldhg,
                         const float* w_f, float* d_att1, float* d_enc, int B, int T, int P, int A, int E,
                         cudaStream_t stream);
int bahdanau_attention_bwd(const float* att1, const float* hg, long long ldhg, const float* w_f, const float* enc,
                           const float* alpha, long long alpha_ld, const float* d_out, long long ld_dout,
                           const float* d_alpha_ext, long long dalpha_ld, float* d_hg, long long ld_dhg,
                           float* d_att1, float* d_enc, float* d_wf, int bt, int P, int A, int E,
                           cudaStream_t stream, void* op_hi = nullptr, float* op_lo = nullptr, int op_dtype = 0,
                           long long ld_op = 0, int scratch_zeroed = 0, float* dawe_out = nullptr,
                           float* dalpha_acc = nullptr, float* de_out = nullptr);
int bcast_add_rows(float* out, const float* v, float scale, int B, int P, int E, cudaStream_t stream);
int adam_clamp(const void* table, const int* block_entry, const long long* block_offset, int n_blocks, float lr,
               float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float clip, int chunk,
               double total_params, cudaStream_t stream, const float* step_dev = nullptr);

// encoder_bwd.cu
int scale_rows_cols(const float* x, const float* colscale, const float* rowscale, int rows_per_group, float* out,
                    long long M, int C, cudaStream_t stream);
int gelu_bwd(const float* pre, float* dh, long long n, cudaStream_t stream);
int cnblock_param_grads(const float* G, const float* W2, const float* b2, const float* gamma, const float* s,
                        float* dW2, float* dgamma, float* db2, int C, int K, cudaStream_t stream);
int dwconv7_wgrad(const float* x, const float* du, float* dw49c, int B, int H, int W, int C, cudaStream_t stream);
int avgpool_nhwc_bwd(const float* dout, float* dx, int B, int H, int W, int C, int S, cudaStream_t stream);

}  // namespace ccx
