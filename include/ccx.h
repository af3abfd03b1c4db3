/* ccx.h — public C ABI of libccx.so, the sm_100a kernel library behind the captioning hot path.
 *
 * The reference (sa06840/ImageCaptioningConvNeXt) has no FFI of its own: its hot path is three nn.Modules
 * (models/encoder.py:14-34, models/decoder.py:34-172, models/transformerDecoder.py:53-168) that call
 * torch / torchvision library ops.  Each entry point below replaces one group of those library calls; the
 * reference call site it stands in for is cited next to it.  The Python host side
 * (imagecaptioningconvnext_b200/*.py) binds these with ctypes and keeps the reference's nn.Module API.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the library never allocates, never
 *     synchronises and keeps no global state besides one-time kernel attribute setup;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - every function returns CCX_OK (0) or a negative CCX_ERR_* code; ccx_status_string() names it;
 *   - matrices are row-major; "linear" weights are [out_features, in_features] like torch.nn.Linear;
 *   - dtype codes: CCX_F32 = 0, CCX_BF16 = 1.  "fp32 compute" means 3xTF32 on the tensor cores: operands are
 *     passed as a (hi, lo) pair of fp32 arrays with hi exactly representable in tf32 (ccx_split_tf32 makes one).
 *   - no CPU fallback exists anywhere: a shape the kernels do not support returns CCX_ERR_SHAPE.
 */
#ifndef CCX_H_
#define CCX_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CCX_API __attribute__((visibility("default")))
#else
#define CCX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define CCX_OK 0
#define CCX_ERR_SHAPE (-1)
#define CCX_ERR_DTYPE (-2)
#define CCX_ERR_CUDA (-3)
#define CCX_ERR_TMA (-4)
#define CCX_ERR_WORKSPACE (-5)

#define CCX_F32 0
#define CCX_BF16 1

#define CCX_ACT_NONE 0
#define CCX_ACT_GELU 1 /* exact erf GELU (nn.GELU default) */
#define CCX_ACT_RELU 2
#define CCX_ACT_GELU_GRAD 3 /* d GELU(x) / dx = Phi(x) + x phi(x): with res_mul the backward of Linear+GELU in one GEMM */

CCX_API int ccx_version(void);
CCX_API const char* ccx_status_string(int status);
CCX_API int ccx_num_sms(void);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction  C[M,N] = epilogue(A[M,K] . W[N,K]^T)  on tcgen05 / TMEM, operands by TMA.
 * Replaces every nn.Linear / 1x1-equivalent conv on the path:
 *   torchvision/models/convnext.py:55-57 (CNBlock MLP), :146-151 (downsample conv as patch-merge GEMM),
 *   models/decoder.py:19-21,50-54, models/transformerDecoder.py:84-85, torch/nn/modules/transformer.py.
 * epilogue(y) = act(y + bias[n]) * emask[m,n] * colscale[n] * rowscale[m / rows_per_group] + residual[m,n]
 * (emask = dropout multiplier, train mode).  K may be any size; lda/ldw (bytes) must be multiples of 16.
 * in_dtype CCX_BF16: A, W bf16; A_lo/W_lo ignored.  in_dtype CCX_F32: A/W are tf32-hi parts, A_lo/W_lo
 * the fp32 remainders (3xTF32); if both *_lo are NULL a single TF32 pass is run.
 * split != 0 (fp32 out only): C receives tf32-hi(y), C_lo the remainder (ready to be the next A operand).
 * ------------------------------------------------------------------------------------------------ */
typedef struct ccx_linear_desc {
  const void* A;
  const void* A_lo;
  const void* W;
  const void* W_lo;
  void* C;
  float* C_lo;
  const float* bias;     /* [N] or NULL */
  const float* colscale; /* [N] or NULL (layer_scale) */
  const float* rowscale; /* [ceil(M/rows_per_group)] or NULL (stochastic-depth noise/(1-p)) */
  const void* residual;  /* [M, ldr] dtype of C, or NULL */
  const float* emask;    /* [M, ldm] fp32 or NULL */
  int64_t lda, ldw, ldc, ldr, ldm; /* leading dimensions in elements */
  int32_t M, N, K;
  int32_t rows_per_group;
  int32_t act;
  int32_t in_dtype;
  int32_t out_dtype;
  int32_t split;
  /* bf16 operands only.  a_mn != 0: A is given as its transpose, a row-major [K, M] array (lda = its row pitch) — the
   * tensor core reads it "MN-major", no transposed copy is made; w_mn != 0: W likewise as [K, N] (ldw = row pitch).
   * Row pitches must be multiples of 8 elements.  This is what lets a Linear layer's backward run on the buffers the
   * forward already has: dX = dY . W uses W [N,K] as it is (w_mn), dW = dY^T . X uses dY [M,N] (a_mn) and X [M,K] (w_mn). */
  int32_t a_mn, w_mn;
  /* res_mul != 0: `residual` MULTIPLIES the activated result instead of being added (no layer-/row-scale then):
   * C = act(A.W^T + bias) * residual.  With CCX_ACT_GELU_GRAD and residual = C = d(hidden): the GELU backward of a
   * CNBlock fused into the re-computation of its pre-activation (torchvision convnext.py:55-56). */
  int32_t res_mul;
} ccx_linear_desc;
CCX_API int ccx_linear(const ccx_linear_desc* d, void* stream);
/* mode != 0: GEMMs that fill the machine with 256x256 tiles use the CTA-pair kernel (tcgen05 cta_group::2, UMMA
 * M=256 across two SMs, each SM streaming half of the B tile).  Off by default: on the ConvNeXt shapes it measures
 * equal or slower than the single-CTA kernel (profiles/r01_spans_v5*.txt). */
/* Cap the number of SMs the persistent GEMM grids use from now on (0 = all SMs): leaves room for a concurrent NCCL
 * kernel while a gradient all-reduce overlaps the backward pass.  Process-wide; takes effect at the next launch. */
CCX_API int ccx_set_sm_limit(int32_t n_sms);
CCX_API int ccx_set_gemm_pair_mode(int32_t mode);

/* fp32 -> (tf32 hi, fp32 lo) and fp32 -> bf16 operand preparation (weights once, activations in epilogues) */
CCX_API int ccx_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
CCX_API int ccx_cast_bf16(const float* x, void* y_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ConvNeXt pieces (channels-last fp32 residual stream [B,H,W,C]).
 * ------------------------------------------------------------------------------------------------ */
/* Conv2d(3,128,4,4)+bias -> LayerNorm2d: torchvision/models/convnext.py:120-131.
 * images NCHW fp32 [B,3,Hin,Win]; w_k [48][128] with k = c*16+kh*4+kw; out NHWC fp32 [B,Hin/4,Win/4,128]. */
CCX_API int ccx_stem_ln(const float* images, const float* w_k, const float* bias, const float* ln_g, const float* ln_b,
                float* out, int32_t B, int32_t Hin, int32_t Win, float eps, void* stream);

/* Same with the input pipeline fused in: images are the dataset's raw uint8 NCHW pixels
 * (utils/utils.py:107,134 stores uint8 (N,3,256,256)) and (x/255 - mean[c]) * inv_std[c] — dataLoader.py:43-45 —
 * is applied while the patch is staged, so the host->device copy is 4x smaller and no normalised fp32 image exists. */
CCX_API int ccx_stem_ln_u8(const uint8_t* images_u8, const float* mean3, const float* inv_std3, const float* w_k,
                           const float* bias, const float* ln_g, const float* ln_b, float* out, int32_t B,
                           int32_t Hin, int32_t Win, float eps, void* stream);

/* Depthwise Conv2d(C,C,7,padding=3,groups=C)+bias -> LayerNorm(C): torchvision/models/convnext.py:52-54.
 * x NHWC fp32; w_tap_major [49][C]; out [B*H*W, C] as bf16 (out_lo NULL) or tf32 (hi, lo) fp32 pair, or plain
 * fp32 when out_dtype == CCX_F32 and out_lo == NULL.  C must be a multiple of 128, at most 1024. */
CCX_API int ccx_dwconv7_ln(const float* x, const float* w_tap_major, const float* bias, const float* ln_g,
                   const float* ln_b, void* out, float* out_lo, int32_t B, int32_t H, int32_t W, int32_t C,
                   float eps, int32_t out_dtype, void* stream);

/* Row LayerNorm over C of x[M,C] fp32 (torchvision LayerNorm2d, convnext.py:31-36; also the post-norm
 * LayerNorm(512, eps 1e-5) of nn.TransformerDecoderLayer, torch/nn/modules/transformer.py:1133-1135).
 * Writes any of: `out` (GEMM operand: bf16, or tf32 hi + `out_lo`) and `out_plain` (fp32, for residual use).  merge != 0 additionally
 * scatters row (b,h,w) to row (b,h/2,w/2), column block (h%2)*2+(w%2) of a [M/4, 4C] matrix — the im2col of the
 * k=2,s=2 downsample conv (convnext.py:146-151), whose weight the host re-orders to [Cout][(kh,kw,c)]. */
CCX_API int ccx_ln_rows(const float* x, const float* ln_g, const float* ln_b, void* out, float* out_lo,
                        float* out_plain, int64_t M, int32_t C, float eps, int32_t out_dtype, int32_t merge,
                        int32_t H, int32_t W, void* stream);

/* AdaptiveAvgPool2d((S,S)) + permute(0,2,3,1): models/encoder.py:25-26.  x NHWC fp32 -> out [B,S,S,C] fp32. */
CCX_API int ccx_avgpool_nhwc(const float* x, float* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t S,
                     void* stream);

/* Whole-encoder runner: children [child_begin, child_end) of convnext_base().features
 * (0 stem, 1 stage1, 2 down, 3 stage2, 4 down, 5 stage3, 6 down, 7 stage4) = models/encoder.py:24.
 * One call = the full launch sequence on `stream` (no host sync), so the Python side pays one FFI call. */
typedef struct ccx_cnblock_weights {
  const float* dw_w; /* [49][C] */
  const float* dw_b;
  const float* ln_g;
  const float* ln_b;
  const void* w1; /* [4C][C] bf16, or tf32 hi */
  const void* w1_lo;
  const float* b1;
  const void* w2; /* [C][4C] */
  const void* w2_lo;
  const float* b2;
  const float* layer_scale; /* [C] */
} ccx_cnblock_weights;

typedef struct ccx_downsample_weights {
  const float* ln_g;
  const float* ln_b;
  const void* w; /* [2C][4C], columns ordered (kh,kw,c) */
  const void* w_lo;
  const float* b;
} ccx_downsample_weights;

#define CCX_MAX_BLOCKS 64
typedef struct ccx_encoder_weights {
  const float* stem_w; /* [48][128] */
  const float* stem_b;
  const float* stem_ln_g;
  const float* stem_ln_b;
  ccx_cnblock_weights blocks[CCX_MAX_BLOCKS]; /* stage-major */
  ccx_downsample_weights down[3];
  int32_t depths[4]; /* 3,3,27,3 */
  int32_t dims[4];   /* 128,256,512,1024 */
  int32_t compute_dtype;
} ccx_encoder_weights;

/* bytes of scratch needed by ccx_encoder_run for a batch of B images of Hin x Win */
CCX_API size_t ccx_encoder_workspace_bytes(int32_t B, int32_t Hin, int32_t Win, int32_t compute_dtype);

/* in: NCHW fp32 images when child_begin == 0, else the NHWC fp32 stream entering child_begin.
 * out: NHWC fp32 stream leaving child_end-1 (may alias `in` only when no stem/downsample is in range).
 * sd_rowscale: [total_blocks][B] stochastic-depth row factors (train mode, tv:ops/stochastic_depth.py:8-44) or NULL. */
CCX_API int ccx_encoder_run(const ccx_encoder_weights* w, const float* in, float* out, int32_t B, int32_t Hin,
                    int32_t Win, int32_t child_begin, int32_t child_end, const float* sd_rowscale,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Caption decoders.  "operand" outputs (op_hi/op_lo/op_dtype) feed the next ccx_linear: bf16 (op_lo NULL), or
 * tf32 hi + fp32 lo for the 3xTF32 path, or plain fp32 (CCX_F32 with op_lo NULL).
 * ------------------------------------------------------------------------------------------------ */
/* nn.Embedding gather (+ dropout multiplier + sinusoidal PE): models/decoder.py:84,
 * models/transformerDecoder.py:97-98 (pos_encoding(dropout(embedding(tokens)))).  Row (b,t), t in [0,nt):
 * token = tokens[b*tok_ld + t0 + t];  value = table[token]*dropmask[(b*nt+t)] + pe[t0+t];
 * written to out_plain[b*sb_p + t*st_p + :] and/or the operand at [b*sb_o + t*st_o + :]. */
CCX_API int ccx_embed_rows(const int64_t* tokens, int64_t tok_ld, int32_t t0, const float* table, int32_t V,
                           int32_t D, const float* pe, const float* dropmask, float* out_plain, int64_t sb_p,
                           int64_t st_p, void* op_hi, float* op_lo, int32_t op_dtype, int64_t sb_o, int64_t st_o,
                           int32_t nb, int32_t nt, void* stream);

/* encoder_out.mean(dim=1): models/decoder.py:64.  enc [B,P,E] fp32 -> operand [B, ldo]. */
CCX_API int ccx_mean_pixels(const float* enc, int32_t B, int32_t P, int32_t E, void* op_hi, float* op_lo,
                            int32_t op_dtype, int64_t ldo, void* stream);

/* One decode step of Attention.forward + the f_beta gate: models/decoder.py:25-31,104-105.
 * att1 = encoder_att(enc) [B,P,A] is hoisted (time-invariant); hg[b] = [decoder_att(h) | f_beta(h)] (A+E cols).
 * alpha = softmax_p(w_f . relu(att1 + att2) + b_f); awe = sigmoid(f_beta h) * sum_p alpha_p enc_p.
 * alpha -> alpha_out[b*alpha_ld + p] (skipped for rows with active[b]==0); awe -> operand row b at b*ld_awe.
 * apply_gate = 0 gives plain Attention.forward; enc_group = g makes decode row b read att1/enc row b/g
 * (beam search: the g beams of an image share its features, caption.py:77 expand()). */
CCX_API int ccx_bahdanau_attention(const float* att1, const float* hg, int64_t ldhg, const float* w_f,
                                   const float* b_f, const float* enc, const float* active, float* alpha_out,
                                   int64_t alpha_ld, void* awe_hi, float* awe_lo, int32_t awe_dtype, int64_t ld_awe,
                                   int32_t bt, int32_t P, int32_t A, int32_t E, int32_t apply_gate,
                                   int32_t enc_group, void* stream);

/* nn.LSTMCell point-wise half (torch/nn/modules/rnn.py:1755-1778; gates = [i|f|g|o] pre-activations incl. both
 * biases): c' = s(f) c + s(i) tanh(g), h' = s(o) tanh(c').  h' goes to the next step's GEMM operand (hn_*), to
 * the hoisted-fc operand (ha_*, times the dropout multiplier: models/decoder.py:109) and/or plain fp32. */
CCX_API int ccx_lstm_pointwise(const float* gates, int64_t ldg, const float* c_prev, float* c_new, void* hn_hi,
                               float* hn_lo, int64_t ld_hn, void* ha_hi, float* ha_lo, int64_t ld_ha,
                               int32_t op_dtype, const float* dropmask, int64_t ld_dm, float* h_plain,
                               int64_t ld_hp, int32_t bt, int32_t D, void* stream);

/* Greedy bookkeeping of forwardWithoutTeacherForcing (models/decoder.py:156-159,
 * models/transformerDecoder.py:146-148): argmax over V (lowest index on ties); for rows with active[b] != 0:
 * sequences[b,t] = argmax, next_tok[b*ld_next] = argmax, active[b] = 0 once <end> is produced. */
CCX_API int ccx_greedy_next(const float* preds, int64_t ld_preds, int32_t B, int32_t V, int32_t t, int32_t T,
                            int64_t* sequences, float* active, int64_t* next_tok, int64_t ld_next,
                            int64_t end_token, void* stream);

/* Scaled-dot-product attention for short sequences (Tk <= ~380 at hd 64): one CTA per (batch, head).
 * Replaces F.scaled_dot_product_attention inside nn.MultiheadAttention (torch/nn/functional.py
 * multi_head_attention_forward) for self-attention (causal + key padding) and cross-attention over pixels.
 * Element (b, i, h*hd+d) of q at q[b*q_sb + i*q_st + h*hd + d]; likewise k, v, ctx.  key_pad [B,Tk] 1 = masked.
 * causal: key j allowed iff j <= q_pos0 + i (q_pos0 = cache length for KV-cache decoding).
 * prob_mask [B,H,Tq,Tk]: attention-dropout multiplier (train); probs_out: softmax saved for backward.
 * kv_group = g: k/v batch row is b/g (beams of one image share its projected memory, caption.py:182). */
CCX_API int ccx_mha_small(const float* q, int64_t q_sb, int64_t q_st, const float* k, int64_t k_sb, int64_t k_st,
                          const float* v, int64_t v_sb, int64_t v_st, void* ctx_hi, float* ctx_lo,
                          int32_t ctx_dtype, int64_t c_sb, int64_t c_st, const uint8_t* key_pad,
                          const float* prob_mask, float* probs_out, int32_t B, int32_t H, int32_t Tq, int32_t Tk,
                          int32_t hd, int32_t causal, int32_t q_pos0, float scale, int32_t kv_group,
                          void* stream);

/* Attention maps for visualisation (models/transformerDecoderAttVis.py:223-226 greedy: the new token's
 * cross-attention weights averaged over layers and heads; :163-165 teacher forcing: averaged over layers and target
 * positions).  A strided reduction over ONE axis of the probabilities written by ccx_mha_small (probs_out, element
 * (b, h, t, j) at ((b*H + h)*Tq + t)*Tk + j):
 *   alphas[b*a_sb + t*a_st + j] (+)= scale * row_active[b] * sum_{h < H} probs[b*p_sb + h*p_sh + t*p_st + j]
 *                                                                (* prob_mask at the same offset, optional)
 * "h" is the summed axis and "t" the kept one — pass the head stride as p_sh to average heads, or the position
 * stride to average positions.  accumulate != 0 adds to alphas (one call per layer, scale = 1/(layers * |axis|));
 * row_active (B floats, optional) zeroes finished rows like the reference's active_indices scatter. */
CCX_API int ccx_attn_head_mean(const float* probs, int64_t p_sb, int64_t p_sh, int64_t p_st, const float* prob_mask,
                               const float* row_active, float* alphas, int64_t a_sb, int64_t a_st, int32_t B,
                               int32_t H, int32_t Tq, int32_t Tk, float scale, int32_t accumulate, void* stream);

/* Single-query attention for KV-cache decoding (one warp per (row, head), no staging): q row r attends to Tk
 * cached positions.  kv_rows [rows, ld_map] (optional): physical cache row that holds position j of logical row
 * r — beam search re-orders beams by rewriting this map instead of copying caches; without it rows read cache
 * row r / kv_group (cross-attention over the shared image memory). */
CCX_API int ccx_mha_decode(const float* q, int64_t q_sb, const float* k, int64_t k_sb, int64_t k_st, const float* v,
                           int64_t v_sb, int64_t v_st, void* ctx_hi, float* ctx_lo, int32_t ctx_dtype, int64_t c_sb,
                           const int32_t* kv_rows, int64_t ld_map, int32_t rows, int32_t H, int32_t Tk, int32_t hd,
                           int32_t kv_group, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Batched beam search (caption.py:96-155 LSTM, caption.py:197-251 Transformer), NI images x k beams (k <= 8) as
 * NI*k decode rows; row img*k + j is beam j of image img, alive beams compacted to the front.
 * ------------------------------------------------------------------------------------------------ */
/* log_softmax over V per row, + top_scores[img, beam], flat top-k_rem[img] over (alive beams x V) sorted
 * descending (ties: lowest flat index); first_step != 0 restricts to beam 0 (caption.py:110, all beams equal).
 * cand_prev = idx / V, cand_word = idx % V (caption.py:116-117). */
CCX_API int ccx_beam_topk(const float* logits, int64_t ld, int32_t NI, int32_t k, int32_t V,
                          const float* top_scores, const int32_t* k_rem, int32_t first_step, float* cand_score,
                          int32_t* cand_prev, int32_t* cand_word, void* stream);

/* Bookkeeping of caption.py:121-145: seqs_out[slot] = seqs_in[prev] + word; candidates ending in <end> are moved
 * to done_seqs/done_scores/done_len (in candidate order), k_rem shrinks; src_row[row] = parent row of each survivor
 * (identity for dead slots) for ccx_gather_rows; next_tok[row*ld_next] = the survivor's new word.
 * step = tokens per sequence before this step (1 at the first step: just <start>).
 * done_parent [NI,k] (optional): global parent row of each completed sequence, from which the caller walks the
 * per-step src_row records back to recover its attention maps (caption.py:122,129 seqsAlpha / completeSeqsAlpha). */
CCX_API int ccx_beam_update(int32_t NI, int32_t k, int32_t Tcap, int32_t step, int64_t end_token,
                            const float* cand_score, const int32_t* cand_prev, const int32_t* cand_word,
                            const int64_t* seqs_in, int64_t* seqs_out, float* top_scores, int32_t* k_rem,
                            int64_t* done_seqs, float* done_scores, int32_t* done_len, int32_t* n_done,
                            int32_t* src_row, int64_t* next_tok, int64_t ld_next, int32_t* done_parent,
                            void* stream);

/* dst[r, 0:row_bytes) = src[src_row[r], 0:row_bytes) — beam re-ordering of h/c (caption.py:140-141) and of the
 * KV caches; src_row NULL = identity.  16-byte granules; strides in bytes. */
CCX_API int ccx_gather_rows(const void* src, int64_t src_stride_bytes, void* dst, int64_t dst_stride_bytes,
                            const int32_t* src_row, int64_t row_bytes, int32_t rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Train step (autograd graph behind trainMultiGPU.py:357-394; clip_gradient utils/utils.py:183-192; Adam).
 * Linear backward uses ccx_linear itself: dX = dY . W (B operand = W^T) and dW = dY^T . X (operands transposed by
 * ccx_convert_operand), accumulating into .grad through the `residual` epilogue.
 * ------------------------------------------------------------------------------------------------ */
/* x[R,C] (fp32 plain: x_lo NULL / tf32 pair / bf16 per x_dtype) times an optional multiplier (mul_mode 1:
 * * mul[r,c] (dropout); 2: * mul_scale where mul[r,c] > 0 else 0 (ReLU mask from the saved, possibly
 * dropped-out, output)) -> GEMM operand (bf16 / tf32
 * hi+lo / fp32); transpose != 0 writes out[c, r] with r zero-padded up to Rpad. */
CCX_API int ccx_convert_operand(const void* x_hi, const float* x_lo, int32_t x_dtype, int64_t ldx,
                                const float* mul, int64_t ldm, int32_t mul_mode, float mul_scale, void* o_hi,
                                float* o_lo, int32_t o_dtype, int64_t ldo, int32_t R, int32_t C, int32_t transpose,
                                int32_t Rpad, void* stream);
/* out[c] += sum_r x[r,c] (* multiplier as above): bias gradients. */
CCX_API int ccx_colsum_acc(const float* x, int64_t ldx, const float* mul, int64_t ldm, int32_t mul_mode,
                           float mul_scale, float* out, int32_t R, int32_t C, void* stream);
/* Capture state of `stream` (0 none, 1 capturing, 2 capture invalidated; < 0 = CCX_ERR_*): lets a caller that records
 * the step into a CUDA graph find the call that broke a capture (CCX_DEBUG_CAPTURE=1 in the Python binding). */
CCX_API int ccx_stream_capture_status(void* stream);
/* Both of the above in one pass over x (fp32, plain): o_bf16[r,c] = bf16(x[r,c] * multiplier) and sums[c] += the same
 * products — the first step of every Linear backward (GEMM operand dY + bias gradient).  Needs C % 4 == 0, ldx % 4 == 0,
 * ldo % 4 == 0 and 16-byte aligned x (CCX_ERR_SHAPE otherwise: call the two entry points above). */
CCX_API int ccx_convert_colsum(const float* x, int64_t ldx, const float* mul, int64_t ldm, int32_t mul_mode,
                               float mul_scale, void* o_bf16, int64_t ldo, float* sums, int32_t R, int32_t C,
                               void* stream);
/* Weight refresh after an optimizer step (trainMultiGPU.py:387-394 moves the fp32 masters; the kernels read bf16 /
 * re-laid-out copies): ONE launch over a device-resident table of rectangular segments,
 *   dst[i, j]  (dst[j, i] if flags & 1)  =  src[row_map ? row_map[i] : i, j]  (+ src2[same index] if src2)
 * for i < rows, j < cols; dst is bf16 (round to nearest even) or, if flags & 2, fp32.  tile0 = index of the segment's
 * first 64 x 64 tile in the grid (segments sorted by it), total_tiles = sum over segments of
 * ceil(rows/64) * ceil(cols/64); `bytes` is only used by the profiling hooks.  Replaces the per-operand
 * cat / index / transpose + cast + copy chains of the decoders' weight preparation. */
typedef struct ccx_cast_seg {
  const float* src;
  const float* src2;
  void* dst;
  const int32_t* row_map;
  int64_t src_ld, dst_ld;
  int32_t rows, cols;
  int32_t flags;
  int32_t tile0;
} ccx_cast_seg;
CCX_API int ccx_cast_segments(const ccx_cast_seg* segs_dev, int32_t nseg, int32_t total_tiles, double bytes,
                              void* stream);
/* LayerNorm backward over rows of x[M,C]; dgamma/dbeta are accumulated (+=).  merge != 0: dy is read in the 2x2
 * patch-merged [M/4, 4C] layout that ccx_ln_rows(merge=1) wrote (downsample LayerNorm2d, convnext.py:146-151). */
CCX_API int ccx_ln_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
                       int64_t M, int32_t C, float eps, int32_t merge, int32_t H, int32_t W, void* stream);
/* Backward of ccx_mha_small (same strided addressing; probs = its probs_out). */
CCX_API int ccx_mha_bwd(const float* q, int64_t q_sb, int64_t q_st, const float* k, int64_t k_sb, int64_t k_st,
                        const float* v, int64_t v_sb, int64_t v_st, const float* dctx, int64_t d_sb, int64_t d_st,
                        const float* probs, const float* prob_mask, float* dq, int64_t dq_sb, int64_t dq_st,
                        float* dk, int64_t dk_sb, int64_t dk_st, float* dv, int64_t dv_sb, int64_t dv_st, int32_t B,
                        int32_t H, int32_t Tq, int32_t Tk, int32_t hd, float scale, void* stream);
/* The same on the tensor cores (mma.sync, bf16 operands, fp32 accumulation) for Tq, Tk <= 64 and hd == 64 — the
 * Transformer decoder's shapes (52 tokens, 49 pixels, 8 heads of 64); CCX_ERR_SHAPE otherwise.  ccx_mha_small takes
 * the matching forward kernel by itself when its context output is bf16. */
CCX_API int ccx_mha_bwd_tc(const float* q, int64_t q_sb, int64_t q_st, const float* k, int64_t k_sb, int64_t k_st,
                           const float* v, int64_t v_sb, int64_t v_st, const float* dctx, int64_t d_sb, int64_t d_st,
                           const float* probs, const float* prob_mask, float* dq, int64_t dq_sb, int64_t dq_st,
                           float* dk, int64_t dk_sb, int64_t dk_st, float* dv, int64_t dv_sb, int64_t dv_st, int32_t B,
                           int32_t H, int32_t Tq, int32_t Tk, int32_t hd, float scale, void* stream);
/* CrossEntropyLoss(mean) over the rows with targets[r] >= 0 (= pack_padded_sequence's selection,
 * trainMultiGPU.py:365-367): *loss_sum += sum_r (lse_r - logit_r[target]) * inv_n;
 * dlogits[r] = (softmax_r - onehot) * inv_n, zero rows for targets < 0 (loss_sum / dlogits may be NULL).
 * stats (3 floats, may be NULL) += {sum of token losses, number of valid rows, top-k hits}: the step metrics of
 * trainMultiGPU.py:396-403 (reduceLossAndTokens + accuracy(scores, targets, 5), utils/utils.py:239-254) in the same
 * pass; a hit = fewer than topk logits strictly larger than the target's. */
CCX_API int ccx_softmax_ce(const float* logits, int64_t ld, const int64_t* targets, int64_t R, int32_t V, float inv_n,
                           float* loss_sum, float* dlogits, int64_t ldd, float* stats, int32_t topk, void* stream);
/* Same kernel with the number of scored rows read from DEVICE memory (inv_n = 1 / max(*n_valid_dev, 1)): the caption
 * lengths then never have to be known on the host (CUDA-graph replay of the train step). */
CCX_API int ccx_softmax_ce_dev(const float* logits, int64_t ld, const int64_t* targets, int64_t R, int32_t V,
                               const float* n_valid_dev, float* loss_sum, float* dlogits, int64_t ldd, float* stats,
                               int32_t topk, void* stream);
/* Targets of the free-running evaluation (utils/utils.py:261-295 preprocessDecoderOutputForMetrics) on the device:
 * targets[i,t] = caps[i,1+t] for t < L_i (L_i = first <end> in sequences[i] + 1, else T) and != <pad>, else -1;
 * decode_len[i] = L_i (may be NULL).  Feed targets to ccx_softmax_ce. */
CCX_API int ccx_free_running_targets(const int64_t* sequences, const int64_t* caps, int64_t cap_ld, int64_t* targets,
                                     int32_t* decode_len, int32_t B, int32_t T, int32_t cap_T, int64_t end_tok,
                                     int64_t pad_tok, void* stream);
/* nn.Embedding dense gradient: dtable[token(b,t)] += dx[b*sb + t*st + :] * dropmask[(b*nt+t), :]. */
CCX_API int ccx_embedding_bwd(const int64_t* tokens, int64_t tok_ld, int32_t t0, const float* dx, int64_t sb,
                              int64_t st, const float* dropmask, float* dtable, int32_t V, int32_t D, int32_t nb,
                              int32_t nt, void* stream);
/* Backward of ccx_lstm_pointwise for one step: dh = dh_fc*dropmask + dh_carry; dc_carry is dL/dc' on entry and
 * dL/dc on exit; dgates [bt,4D] are the pre-activation gradients (feed the dgrad / batched wgrad GEMMs). */
CCX_API int ccx_lstm_pointwise_bwd(const float* gates, int64_t ldg, const float* c_prev, const float* c_new,
                                   const float* dh_fc, int64_t ld_fc, const float* dropmask, int64_t ld_dm,
                                   const float* dh_carry, float* dc_carry, float* dgates, int64_t lddg, int32_t bt,
                                   int32_t D, void* stream);
/* Backward of ccx_bahdanau_attention for one step (alpha = its saved output): d_hg = [d att2 | d gate pre-act],
 * d_att1 / d_enc accumulate over steps (+=), d_wf accumulates atomically; d_alpha_ext = gradient reaching alpha
 * from outside (the doubly-stochastic regulariser, trainMultiGPU.py:369). */
CCX_API int ccx_bahdanau_attention_bwd(const float* att1, const float* hg, int64_t ldhg, const float* w_f,
                                       const float* enc, const float* alpha, int64_t alpha_ld, const float* d_out,
                                       int64_t ld_dout, const float* d_alpha_ext, int64_t dalpha_ld, float* d_hg,
                                       int64_t ld_dhg, float* d_att1, float* d_enc, float* d_wf, int32_t bt,
                                       int32_t P, int32_t A, int32_t E, void* stream);
/* out[b,p,:] += v[b,:] * scale — backward of encoder_out.mean(dim=1) (models/decoder.py:64). */
CCX_API int ccx_bcast_add_rows(float* out, const float* v, float scale, int32_t B, int32_t P, int32_t E,
                               void* stream);
/* ---- fine-tuned ConvNeXt stage backward (autograd through torchvision/models/convnext.py:51-67) ---- */
/* Depthwise 7x7 conv without the LayerNorm: out = conv(x, w) + bias + addend (bias/addend may be NULL).  Used to
 * recompute the LayerNorm input and, with flipped taps, as the data gradient of the depthwise conv. */
CCX_API int ccx_dwconv7_plain(const float* x, const float* w_tap_major, const float* bias, const float* addend,
                              float* out, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* out[m,c] = x[m,c] * colscale[c] * rowscale[m / rows_per_group] (either scale may be NULL). */
CCX_API int ccx_scale_rows_cols(const float* x, const float* colscale, const float* rowscale, int32_t rows_per_group,
                                float* out, int64_t M, int32_t C, void* stream);
/* dh[i] *= gelu'(pre[i]) (exact erf GELU, convnext.py:56). */
CCX_API int ccx_gelu_bwd(const float* pre, float* dh, int64_t n, void* stream);
/* From G = (dout*rowscale)^T . h and s = colsum(dout*rowscale): dW2 += gamma*G, d layer_scale += W2.G + b2*s,
 * d b2 += gamma*s — the layer_scale / second Linear gradients without recomputing the branch output. */
CCX_API int ccx_cnblock_param_grads(const float* G, const float* W2, const float* b2, const float* gamma,
                                    const float* s, float* dW2, float* dgamma, float* db2, int32_t C, int32_t K,
                                    void* stream);
/* dw[tap][c] += sum du[b,h,w,c] * x[b,h+kh-3,w+kw-3,c]  (depthwise filter gradient, tap-major). */
CCX_API int ccx_dwconv7_wgrad(const float* x, const float* du, float* dw_tap_major, int32_t B, int32_t H, int32_t W,
                              int32_t C, void* stream);
/* AdaptiveAvgPool2d((S,S)) backward over NHWC. */
CCX_API int ccx_avgpool_nhwc_bwd(const float* dout, float* dx, int32_t B, int32_t H, int32_t W, int32_t C, int32_t S,
                                 void* stream);
/* ---- teacher-forced LSTM-attention time loop driven from C++ (models/decoder.py:100-111 and its backward) ----
 * One call launches every step's kernels (GEMM [decoder_att|f_beta], attention, GEMM gates, LSTM point-wise);
 * bts_host[t] = number of active (length-sorted) rows at step t.  Buffer layouts as in decoder.py:
 * XH [T+1][B][Emb+E+D] operand (hi/lo), C_all [T+1][B][D], HG [T][B][A+E], G [T][B][4D], alphas [B][T][P],
 * H_all [B][T][D] operand, dropmask [B][T][D] or NULL. */
typedef struct ccx_lstm_tf {
  void* XH_hi;
  float* XH_lo;
  float* C_all;
  float* HG;
  float* G;
  float* alphas;
  void* H_all_hi;
  float* H_all_lo;
  const float* dropmask;
  const float* att1; /* [B*P, A] hoisted encoder_att(enc) */
  const float* enc;  /* [B, P, E] length-sorted */
  const void* w_h;   /* [A+E, D] operand */
  const void* w_h_lo;
  const float* b_h;
  const float* w_f;
  const float* b_f;
  const void* w_lstm; /* [4D, Emb+E+D] operand */
  const void* w_lstm_lo;
  const float* b_lstm;
  const int32_t* bts_host; /* HOST array [T] */
  int32_t B, T, P, E, A, D, Emb;
  int32_t compute_dtype;
} ccx_lstm_tf;
CCX_API int ccx_lstm_tf_forward(const ccx_lstm_tf* s, void* stream);

/* BPTT over the same buffers: dH_all [B][T][D] (fc dgrad), dalphas [B][T][P] or NULL; work buffers dG_all
 * [T][B][4D], dHG_all [T][B][A+E], dXH_all [T][B][Emb+E+D] (zero-initialised by the caller), dh / dc [B][D]
 * (zero-initialised; on return dL/dh_0, dL/dc_0), d_att1 [B*P,A] (+=), d_enc [B,P,E] (+=, may be NULL), d_wf [A];
 * w_*_t = transposed weight operands [in, out]; scratch / scratch2 = operand staging buffers (written by the
 * point-wise and attention backward kernels themselves: 5 launches per step). */
typedef struct ccx_lstm_tf_bwd {
  const float* dH_all;
  const float* dalphas;
  float* dG_all;
  float* dHG_all;
  float* dXH_all;
  float* dh;
  float* dc;
  float* d_att1;
  float* d_enc;
  float* d_wf;
  const void* w_lstm_t; /* [Emb+E+D, 4D] operand */
  const void* w_lstm_t_lo;
  const void* w_h_t; /* [D, A+E] operand */
  const void* w_h_t_lo;
  void* scratch_hi;  /* [B, 4D] operand staging of dgates */
  float* scratch_lo;
  void* scratch2_hi; /* [B, A+E] operand staging of d[att2 | gate] */
  float* scratch2_lo;
  /* optional (all three or none), zero-initialised by the caller: with them the per-step attention backward only
   * records d_awe_raw [T][B][E] and d e [T][B][P] (d alpha accumulates in dalpha_all [T][B][P]) and the sums over
   * time into d_enc / d_att1 run once after the loop, not as a read-modify-write pass at every step */
  float* dawe_all;
  float* dalpha_all;
  float* de_all;
} ccx_lstm_tf_bwd;
CCX_API int ccx_lstm_tf_backward(const ccx_lstm_tf* s, const ccx_lstm_tf_bwd* b, void* stream);

/* Persistent form of the two loops above (bf16 compute, B <= 32, P*(A+E)*2 bytes of one sample resident in one
 * CTA's shared memory, A = D = Emb = 512, E = 1024 — ccx_lstm_persist_supported() tells): ONE cooperative kernel per
 * direction; the recurrent weights stay in shared memory, every GEMM is a swap-AB tcgen05.mma (weights = M side,
 * batch = N = 32, accumulator in TMEM) and the CTA roles hand over through release/acquire counters instead of
 * kernel boundaries (csrc/lstm_persist.cu).  Replaces models/decoder.py:100-111 (and its autograd graph) exactly like
 * ccx_lstm_tf_forward / _backward and fills the same buffers.  Extra inputs:
 *   E_all   [T][B][4D] fp32: hoisted emb_t . W_ih[:, :Emb]^T + b_ih + b_hh with columns in the PERMUTED gate order
 *           col = 32*(j/8) + 8*gate + j%8  for hidden unit j (the row order of w2p),
 *   w2p     [4D][D+E] bf16: rows in that permuted order, columns [W_hh | W_ih[:, Emb:]],
 *   att1_bf [B*P][A], enc_bf [B*P][E] bf16; decode_len [B] (sorted descending, = caption length - 1),
 *   counters: >= 4*(T+1) int32 of scratch (cleared by the call). */
typedef struct ccx_lstm_persist {
  const float* E_all;
  const void* w2p;
  const void* att1_bf;
  const void* enc_bf;
  const int64_t* decode_len;
  int32_t* counters;
  void* scratch;  /* (T+1)*32 KB + T*64 KB + 1.5 MB, 128-byte aligned: per-step operand images (h, gated awe) and the K-quarter partial sums of the attention projections; no init needed */
  float* awe_all; /* [T][B][E] out: un-gated context vectors (what the backward kernel needs), or NULL */
  int64_t* dbg;   /* optional [3][T][8] clock64 stamps of one CTA per role (tools/bench_lstm_persist.py), or NULL */
} ccx_lstm_persist;
CCX_API int ccx_lstm_persist_supported(int32_t B, int32_t P, int32_t E, int32_t A, int32_t D, int32_t Emb,
                                       int32_t compute_dtype);
CCX_API int ccx_lstm_tf_forward_persist(const ccx_lstm_tf* s, const ccx_lstm_persist* p, void* stream);

/* BPTT as one cooperative kernel (+ the two deferred d_enc / d_att1 sums): fills dG_all, dHG_all, dawe_all, de_all,
 * d_wf (+=), dh / dc (dL/dh_0, dL/dc_0), d_att1 / d_enc (+=) of ccx_lstm_tf_bwd exactly like ccx_lstm_tf_backward;
 * dXH_all, the w_*_t operands and the scratch operands of that struct are not used (the embedding gradient comes
 * from dG_bf . W_ih[:, :Emb] as one GEMM after the call).  awe_all / att1_bf / enc_bf / decode_len as in the forward. */
typedef struct ccx_lstm_persist_bwd {
  const void* wx;   /* [E+D][4D] bf16: rows = [W_ih[:, Emb:]^T ; W_hh^T] */
  const void* wht;  /* [D][A+E] bf16: [decoder_att ; f_beta]^T */
  void* dG_bf;      /* [T][B][4D] bf16 out (dgates as a row-major GEMM operand); zero-initialised by the caller */
  void* scratch;    /* T*224 KB, 128-byte aligned: per-step operand images (dgates, d[att2|gate]); no init needed */
  float* Xp;        /* 2*8*B*E + 2*4*B*D fp32 of scratch (K-slice partial sums) */
  const float* awe_all;
  const void* att1_bf;
  const void* enc_bf;
  const int64_t* decode_len;
  int32_t* counters; /* >= 4*(T+1) int32, cleared by the call */
  int64_t* dbg;
} ccx_lstm_persist_bwd;
CCX_API int ccx_lstm_tf_backward_persist(const ccx_lstm_tf* s, const ccx_lstm_tf_bwd* b, const ccx_lstm_persist_bwd* p,
                                         void* stream);

/* clip_gradient (grad.clamp_(-clip, clip), utils/utils.py:189-192) fused with torch.optim.Adam's single-tensor
 * update (no weight decay / amsgrad), over a device table of {param, grad, exp_avg, exp_avg_sq, n} entries;
 * block i handles elements [block_offset[i], +chunk) of entry block_entry[i]. */
CCX_API int ccx_adam_clamp(const void* table, const int32_t* block_entry, const int64_t* block_offset,
                           int32_t n_blocks, float lr, float beta1, float beta2, float eps, float bc1,
                           float bc2_sqrt, float clip, int32_t chunk, double total_params, void* stream);
/* Same update with the step count t read from DEVICE memory (bias corrections 1 - beta^t computed in the kernel): the
 * form a CUDA-graph replay of the train step needs, where no host-side value may be baked into the launch. */
CCX_API int ccx_adam_clamp_dev(const void* table, const int32_t* block_entry, const int64_t* block_offset,
                               int32_t n_blocks, float lr, float beta1, float beta2, float eps,
                               const float* step_dev, float clip, int32_t chunk, double total_params, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-launch CUDA-event timing (bench.py's roofline).  Between begin and end every kernel launch made by the
 * library is bracketed by two events on its stream; end() synchronises the device and returns, per kernel kind
 * (0 gemm, 1 dwconv+ln, 2 stem, 3 ln_rows, 4 pool, 5 elementwise, 6 attention, 7 lstm, 8 loss, 9 optimizer, 10 skinny (M <= 32) gemm),
 * the summed milliseconds, the summed algorithmic work (FLOPs for kind 0, bytes otherwise) and the launch count.
 * ------------------------------------------------------------------------------------------------ */
#define CCX_PROF_KINDS 11
CCX_API int ccx_prof_begin(void);
CCX_API int ccx_prof_spans(int32_t* kind_host, double* ms_host, double* work_host, int32_t max);
CCX_API int ccx_prof_end(double* ms_per_kind_host, double* work_per_kind_host, int64_t* launches_per_kind_host,
                         int32_t n_kinds);

#ifdef __cplusplus
}
#endif
#endif /* CCX_H_ */
