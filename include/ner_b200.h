/*
 * ner_b200.h — C-ABI of libner_b200.so: the sm_100a kernels behind the
 * bert_bilstm_crf hot path of DSXiangLi/ChineseNER.
 *
 * Every entry point mirrors one reference call site (cited per function,
 * paths relative to the reference repo).  Conventions:
 *   - all pointers are BORROWED DEVICE pointers (row-major, contiguous) unless
 *     the parameter name ends in `_host`;
 *   - the library never allocates user-visible memory: workspaces are
 *     caller-provided and sized by the matching *_workspace_bytes();
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     stateless and re-entrant; no internal synchronisation;
 *   - return value: NER_OK (0) or a negative status; ner_strerror() names it.
 *     No C++ exception crosses this boundary.
 */
#ifndef NER_B200_H_
#define NER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NER_OK 0
#define NER_ERR_INVALID_ARG (-1)   /* null pointer, negative size, bad enum */
#define NER_ERR_UNSUPPORTED (-2)   /* e.g. K > 32 tags, H not supported */
#define NER_ERR_WORKSPACE (-3)     /* workspace too small / missing */
#define NER_ERR_NO_DRIVER (-4)     /* driver entry point (TMA encode) not found */
#define NER_ERR_CUDA_BASE (-1000)  /* -(1000 + cudaError_t) */

typedef void* ner_stream_t; /* cudaStream_t */

const char* ner_strerror(int status);
/* Library/ABI version; bumps when a signature changes. */
int ner_abi_version(void);
/* Build provenance: "src=<hash of the sources this library was compiled from> nvcc=<version> arch=sm_100a"; the hash is
 * chinesener_b200.build.source_hash() of the tree at compile time. */
const char* ner_build_info(void);

/* ------------------------------------------------------------------------ *
 * CRF  — replaces tf.contrib.crf as called from tools/layer.py
 * ------------------------------------------------------------------------ */

/* tools/layer.py:140-142  crf_decode -> tf.contrib.crf.crf_decode
 * Viterbi max-plus recursion + backtrace.  logits [B,L,K] f32, seq_len [B]
 * i32, trans [K,K] f32 (trans[i*K+j] = score of i->j).  tags_out [B,L] i32 is
 * zero beyond seq_len; best_score [B] f32 may be NULL.  Ties resolve to the
 * lowest tag index; fp32 association order is (s[i]+trans[i][j]) then
 * +logits, so the tag indices are bit-exact with the reference.  K <= 32. */
int ner_crf_viterbi(const float* logits, const int32_t* seq_len, const float* trans,
                    int32_t* tags_out, float* best_score, int B, int L, int K,
                    ner_stream_t stream);

/* tools/layer.py:122-127  crf_layer -> tf.contrib.crf.crf_log_likelihood
 * ll[b] = gold-path score - log-partition (forward-alpha recursion).
 * tags [B,L] i32.  alpha_ws: NULL, or [B,L,K] f32 that receives alpha_t for
 * the backward pass.  logz_out: NULL or [B] f32.
 * flags: bit0 = force the exact (per-column max) logsumexp path. */
int ner_crf_loglik_fwd(const float* logits, const int32_t* tags, const int32_t* seq_len,
                       const float* trans, float* ll, float* logz_out, float* alpha_ws,
                       int B, int L, int K, int flags, ner_stream_t stream);

/* Gradient of the log-likelihood (the reference gets it from tf.gradients,
 * tools/train_utils.py:314): for g_b = (d_ll ? d_ll[b] : 1) * scale,
 *   d_logits[b,t,j] = g_b * (1[y_t=j] - P(y_t=j|x))            (0 beyond seq_len)
 *   d_trans[i,j]   += sum_b g_b * (count_b(i->j) - sum_t P(y_{t-1}=i,y_t=j|x))
 * alpha_ws / logz come from ner_crf_loglik_fwd.  d_logits [B,L,K] is fully
 * written; d_trans [K,K] is ACCUMULATED into (caller zeroes it).  For the
 * reference loss mean(-ll) (model/bert_bilstm_crf.py:32) pass d_ll = NULL,
 * scale = -1/B. */
int ner_crf_loglik_bwd(const float* logits, const int32_t* tags, const int32_t* seq_len,
                       const float* trans, const float* alpha_ws, const float* logz,
                       const float* d_ll, float scale, float* d_logits, float* d_trans, int B,
                       int L, int K, ner_stream_t stream);


/* ------------------------------------------------------------------------ *
 * Dense layers on tcgen05 tensor cores — replaces tf.layers.dense /
 * modeling.dense_layer inside BertModel (tools/layer.py:68-77), the logits
 * projection's big-M cousins and the LSTM input projection (tools/layer.py:35)
 * ------------------------------------------------------------------------ */
#define NER_EPI_F32 0            /* out f32  = acc + bias                    */
#define NER_EPI_BF16 1           /* out bf16 = acc + bias                    */
#define NER_EPI_GELU_TANH_BF16 2 /* out bf16 = gelu_tanh(acc + bias)         */
#define NER_EPI_GELU_ERF_BF16 3  /* out bf16 = gelu_erf(acc + bias)          */
#define NER_EPI_RELU_BF16 4      /* out bf16 = relu(acc + bias)              */
#define NER_EPI_RES_F32 5        /* out f32  = acc + bias + residual (f32)   */
#define NER_EPI_RES_RELU_F32 6   /* out f32  = relu(acc + bias + residual)   */
#define NER_EPI_DIAG_DISCARD 99  /* diagnostic only: accumulate, drain TMEM, store nothing */

/* tile_n selectors of ner_gemm_bf16: 64/128/192/256 = one CTA per 128 x tile_n tile
 * (cta_group::1), whole tiles round-robin over the SMs; the 2CTA values = a CTA pair per 256 x N
 * tile (cta_group::2); the SK values = 128 x tile_n tiles with stream-K scheduling (every SM gets
 * the same number of k-blocks; split tiles are summed through an internal fp32 scratch). */
#define NER_TILE_2CTA_128 1128
#define NER_TILE_2CTA_256 1256
#define NER_TILE_SK_128 2128
#define NER_TILE_SK_256 2256
/* tile_n = 0 fits whole waves on the SMs (best latency of ONE GEMM).  NER_TILE_AUTO_THROUGHPUT takes the
 * tile with the best FLOP rate (128 x 256 when N % 256 == 0): the choice when several streams keep the
 * SMs busy, so a partial last wave is not idle time. */
#define NER_TILE_AUTO_THROUGHPUT 3000

/* out[M,N] = epilogue(A[M,K] · Wt[N,K]^T + bias[N]).  A and Wt are bf16,
 * K contiguous (Wt is the TF kernel [K,N] transposed once by
 * ner_pack_weight_bf16).  bias may be NULL.  K % 8 == 0, N % 32 == 0.
 * tile_n: 0 = auto, or one of the selectors above. */
int ner_gemm_bf16(const void* A, const void* Wt, const float* bias, const float* residual,
                  void* out, int M, int N, int K, int epilogue, int tile_n,
                  ner_stream_t stream);

/* TF dense kernel [K,N] f32 -> bf16 [N,K] (K contiguous), the B-operand layout of
 * ner_gemm_bf16.  Done once per weight (or per optimizer step). */
int ner_pack_weight_bf16(const float* w_kn, void* wt_nk_bf16, int K, int N, ner_stream_t stream);
/* The same for a whole group of kernels in one launch, writing BOTH bf16 layouts a TRAIN step needs from one read of the fp32
 * weights: entry e = TF-layout f32 kernel src [K, N]; dst_kn_bf16 (nullable) receives the cast in place layout with row stride
 * ld_kn (the K-major operand of the data-gradient GEMMs; a column block of a fused matrix via the pointer offset), dst_nk_bf16
 * (nullable) the transposed [N, K] pack with row stride ld_nk (the operand of the forward GEMMs).  entries_device [count] and
 * tile_start_device [count + 1] (prefix sums of ceil(K/64) * ceil(N/64)) are DEVICE arrays; total_tiles = tile_start[count]. */
typedef struct {
  const float* src;
  int K;
  int N;
  void* dst_nk_bf16;
  int ld_nk;
  void* dst_kn_bf16;
  int ld_kn;
} ner_pack_entry;
int ner_pack_weights_group_bf16(const ner_pack_entry* entries_device, const int32_t* tile_start_device, int count,
                                int total_tiles, ner_stream_t stream);
/* Elementwise f32 -> bf16 (round to nearest even). */
int ner_cast_bf16(const float* src, void* dst_bf16, size_t n, ner_stream_t stream);

/* tf.layers.dense(units=label_size) — model/bert_bilstm_crf.py:26, model/bert_crf.py:20.
 * out[M,N] f32 = x[M,F] · W[F,N] + bias[N], N <= 32; x is f32 (x_is_bf16=0) or bf16;
 * W is the TF kernel layout [F,N] f32.  row_map: NULL, or [M] i32 — input row r is written to
 * output row row_map[r] (scatter of packed token rows back to the padded [B*L] layout). */
int ner_dense_small_n(const void* x, int x_is_bf16, const float* W, const float* bias, float* out,
                      int M, int F, int N, const int32_t* row_map, ner_stream_t stream);

/* tf.nn.embedding_lookup (model/bilstm_crf.py:24): out[tok, 0:E] = table[ids[tok]], row
 * stride ld_out >= E. */
int ner_embedding_lookup(const float* table, const int32_t* ids, float* out, int n_tok, int E,
                         int V, int ld_out, ner_stream_t stream);
/* f32 [M,D] (row stride ld_src) -> bf16 [M,Dp] zero-padded to the GEMM's K % 8 == 0. */
int ner_cast_pad_bf16(const float* src, void* dst_bf16, int M, int D, int Dp, int ld_src,
                      ner_stream_t stream);

/* Sequence packing plan (removes padding rows from the token-major activations; padded
 * positions never reach loss or pred_ids because the CRF ignores t >= seq_len).  mask [B,L] i32
 * must be a prefix mask (1 for t < len_b), as data/base_preprocess.py:166-174 builds it.
 * cu_seqlens [B+1]: exclusive prefix sum of the lengths; tok_src [B*L]: tok_src[cu[b]+t] = b*L+t. */
int ner_seq_pack_plan(const int32_t* mask, int32_t* cu_seqlens, int32_t* tok_src, int B, int L,
                      ner_stream_t stream);

/* f32 [M,D] (row stride ld_src) -> hi = bf16(x), lo = bf16(x - hi), both [M,Dp] zero-padded.
 * Operands of the fp32-accurate dense mode: out = A_hi·W_hi + A_hi·W_lo + A_lo·W_hi, three
 * ner_gemm_bf16 launches chained through the f32 residual epilogue (error ~2^-17 relative). */
int ner_split_bf16(const float* src, void* hi_bf16, void* lo_bf16, int M, int D, int Dp,
                   int ld_src, ner_stream_t stream);

/* fp32 self-attention for small heads with the optional TENER relative-position term
 * (tools/transformer/tener.py:12-119): scores[q,k] = (Q_q+u_h)·K_k [+ (Q_q+v_h)·R_{k-q+L}],
 * times `scale`; keys k >= seq_len[b] are masked; softmax; ·V.  Q/K/V are row-major
 * [B*L, ld*] f32 with head h at columns [h*head_dim, (h+1)*head_dim).  u/v [heads, head_dim]
 * or NULL; rel_table [2L, head_dim] (positions -L..L-1) or NULL.  Writes out_f32 and/or a
 * (hi, lo) bf16 pair, all [B*L, heads*head_dim]; rows q >= seq_len[b] are zero.
 * head_dim in {20, 32, 40, 64}, L <= 512. */
int ner_attention_f32(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                      const float* bias_u, const float* bias_v, const float* rel_table,
                      const int32_t* seq_len, float scale, float* out_f32, void* out_hi_bf16,
                      void* out_lo_bf16, int B, int L, int num_heads, int head_dim,
                      ner_stream_t stream);

/* ------------------------------------------------------------------------ *
 * BERT encoder pieces — bert_base.bert.modeling.BertModel as driven from
 * tools/layer.py:63-81 (pretrain_bert_embedding)
 * ------------------------------------------------------------------------ */

/* embedding_lookup + embedding_postprocessor: LN(word[ids] + type[seg] + pos[0:L]).
 * Tables f32: word [vocab,H], type [n_type,H], pos [max_pos,H]; ids/seg [B,L] i32
 * (seg may be NULL = all zero).  Writes f32 and/or bf16 [B*L,H] (either may be NULL).
 * Packed mode: tok_src [n_packed] (from ner_seq_pack_plan) selects the padded index of every
 * packed row; output has n_packed rows.  tok_src = NULL: padded mode, B*L rows. */
int ner_bert_embed_ln(const float* word_emb, const float* type_emb, const float* pos_emb,
                      const float* gamma, const float* beta, const int32_t* ids,
                      const int32_t* seg, float* out_f32, void* out_bf16, int B, int L, int H,
                      int vocab, int n_type, int max_pos, float eps, const int32_t* tok_src,
                      int n_packed, ner_stream_t stream);

/* LayerNorm over the last axis of (y + optional f32 residual) [M,H] -> f32 and/or bf16.
 * y is f32 (y_is_bf16 = 0) or bf16 (the dense layer's bf16 epilogue output).
 * modeling.layer_norm (eps 1e-12) and tools/transformer/modules.py:40-65 (eps = FLT_EPSILON). */
int ner_layernorm(const void* y, int y_is_bf16, const float* residual, const float* gamma,
                  const float* beta, float* out_f32, void* out_bf16, int M, int H, float eps,
                  ner_stream_t stream);

/* ner_layernorm with BertModel's hidden dropout fused in front of the residual add:
 * out = LN(dropout(y) + residual), mask = the counter-based decisions of ner_dropout (element = row*H + col).
 * keep_prob = 1: identical to ner_layernorm. */
int ner_layernorm_dropout(const void* y, int y_is_bf16, const float* residual, const float* gamma,
                          const float* beta, float* out_f32, void* out_bf16, int M, int H, float eps,
                          float keep_prob, uint64_t seed, ner_stream_t stream);

/* attention_layer core: ctx = softmax(Q K^T * scale + (1-mask)*mask_add) V per head.
 * qkv bf16 [B*L, 3*num_heads*head_dim] (Q | K | V blocks, heads contiguous inside each),
 * mask [B,L] i32 (1 = keep), ctx bf16 [B*L, num_heads*head_dim].  head_dim must be 64.
 * BERT: scale = 1/sqrt(64), mask_add = -10000.  Packed mode: cu_seqlens [B+1] non-NULL — sequence b
 * occupies rows [cu[b], cu[b+1]) of qkv/ctx, every key is valid, mask is ignored.
 * keep_prob < 1: attention_probs dropout of BertModel in training (probabilities scaled by
 * keep(seed; b, head, q, k) / keep_prob after the softmax); keep_prob = 1: inference.
 * n_rows: rows of qkv / ctx — the packed token count in packed mode (0 = unknown), B*L or 0 in padded mode.
 * keep_prob = 1 with known n_rows and L <= 256 runs on tcgen05 (S = Q K^T and O = P V accumulate in tensor memory, Q/K/V
 * tiles arrive by TMA, V is consumed as an MN-major operand); otherwise a warp-level mma.sync kernel (ABI version 2). */
int ner_bert_attention(const void* qkv_bf16, const int32_t* mask, void* ctx_bf16, int B, int L,
                       int num_heads, int head_dim, float scale, float mask_add,
                       const int32_t* cu_seqlens, int n_rows, float keep_prob, uint64_t seed,
                       ner_stream_t stream);

/* Backward of ner_attention_f32 (TRAIN mode of tools/transformer/tener.py:12-119).  d_out f32 [B*L, ld_dout]
 * = dL/d out.  Writes dQ (all rows; zero for t >= seq_len) and ACCUMULATES into dK, dV (f32, layouts of K / V),
 * d_bias_u / d_bias_v [num_heads, head_dim] (nullable) — zero them first.  Scores are recomputed. */
int ner_attention_f32_bwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                          const float* bias_u, const float* bias_v, const float* rel_table,
                          const int32_t* seq_len, float scale, const float* d_out, int ld_dout,
                          float* dQ, int ld_dq, float* dK, int ld_dk, float* dV, int ld_dv,
                          float* d_bias_u, float* d_bias_v, int B, int L, int num_heads, int head_dim,
                          ner_stream_t stream);

/* Whole BertModel forward in one call (what tools/layer.py:68-77 gets from
 * modeling.BertModel(...).get_sequence_output()).  `layers` is a HOST array of per-layer
 * device pointers: dense kernels packed by ner_pack_weight_bf16 ([N,K] bf16; wqkv = the
 * query|key|value kernels concatenated along N), biases / LayerNorm parameters f32.
 * Padded mode: cu_seqlens = tok_src = NULL, outputs have B*L rows.  Packed mode: both from
 * ner_seq_pack_plan, n_packed = total tokens, outputs have n_packed rows.
 * out_f32 / out_bf16 [rows,H] receive sequence_output; workspace from the sizing call. */
typedef struct {
  int hidden_size, num_heads, intermediate_size, num_layers;
  int vocab_size, type_vocab_size, max_position;
  float ln_eps;   /* 1e-12 */
  int gelu_erf;   /* 0: tanh approximation (google-research/bert modeling.gelu), 1: erf */
  int gemm_tile;  /* tile_n passed to every ner_gemm_bf16 of the composite calls: 0 or NER_TILE_AUTO_THROUGHPUT */
} ner_bert_config;

typedef struct {
  const void* wqkv;  const float* bqkv;        /* [3H,H] bf16, [3H] */
  const void* wo;    const float* bo;          /* attention/output/dense */
  const float* ln1_gamma; const float* ln1_beta;
  const void* wi;    const float* bi;          /* intermediate/dense [I,H] bf16 */
  const void* wd;    const float* bd;          /* output/dense [H,I] bf16 */
  const float* ln2_gamma; const float* ln2_beta;
} ner_bert_layer_weights;

size_t ner_bert_encoder_workspace_bytes(const ner_bert_config* cfg, int rows);
int ner_bert_encoder_fwd(const ner_bert_config* cfg, const float* word_emb, const float* type_emb,
                         const float* pos_emb, const float* emb_ln_gamma, const float* emb_ln_beta,
                         const ner_bert_layer_weights* layers, const int32_t* ids,
                         const int32_t* mask, const int32_t* seg, int B, int L,
                         const int32_t* cu_seqlens, const int32_t* tok_src, int n_packed,
                         float* out_f32, void* out_bf16, void* workspace, size_t workspace_bytes,
                         ner_stream_t stream);

/* TRAIN-mode BertModel (is_training=True) as two calls: forward keeping every activation the backward
 * pass needs, and backward accumulating into the caller's gradient tensors (tf.gradients of
 * tools/train_utils.py:314 through the encoder).  Padded layout, rows = B*L.
 * Dropout: hidden_keep = 1 - hidden_dropout_prob (embedding output, attention-output dense, FFN-output
 * dense), attn_keep = 1 - attention_probs_dropout_prob; counter-based masks from `seed`, regenerated by
 * the backward call (same seed).  `saved` (ner_bert_train_saved_bytes) carries the activations from
 * forward to backward; `scratch` (ner_bert_train_scratch_bytes) is backward-only. */
typedef struct {
  /* TF-layout [K_in, N_out] bf16 casts of the dense kernels (ner_cast_bf16): B operands of the data-gradient GEMMs */
  const void* wqkv_kn; const void* wo_kn; const void* wi_kn; const void* wd_kn;
  /* gradient tensors, f32, accumulated into (TF shapes: kernels [in,out], biases [out]) */
  float* d_wq; float* d_wk; float* d_wv; float* d_bq; float* d_bk; float* d_bv;
  float* d_wo; float* d_bo; float* d_ln1_gamma; float* d_ln1_beta;
  float* d_wi; float* d_bi; float* d_wd; float* d_bd; float* d_ln2_gamma; float* d_ln2_beta;
} ner_bert_layer_grads;

size_t ner_bert_train_saved_bytes(const ner_bert_config* cfg, int rows);
size_t ner_bert_train_scratch_bytes(const ner_bert_config* cfg, int rows);
int ner_bert_encoder_train_fwd(const ner_bert_config* cfg, const float* word_emb, const float* type_emb,
                               const float* pos_emb, const float* emb_ln_gamma, const float* emb_ln_beta,
                               const ner_bert_layer_weights* layers, const int32_t* ids,
                               const int32_t* mask, const int32_t* seg, int B, int L,
                               float hidden_keep, float attn_keep, uint64_t seed, float* out_f32,
                               void* out_bf16, void* saved, size_t saved_bytes, ner_stream_t stream);
int ner_bert_encoder_train_bwd(const ner_bert_config* cfg, const float* emb_ln_gamma,
                               const ner_bert_layer_weights* layers, const ner_bert_layer_grads* grads,
                               float* d_word_emb, float* d_type_emb, float* d_pos_emb,
                               float* d_emb_ln_gamma, float* d_emb_ln_beta, const int32_t* ids,
                               const int32_t* mask, const int32_t* seg, int B, int L,
                               float hidden_keep, float attn_keep, uint64_t seed, const float* d_out,
                               const void* saved, size_t saved_bytes, void* scratch,
                               size_t scratch_bytes, ner_stream_t stream);
/* Sequence-packed TRAIN composites (cu_seqlens / tok_src / n_packed from ner_seq_pack_plan): the per-token kernels run on
 * the n_packed real tokens only — BertModel's work on [PAD] positions feeds nothing that bert_bilstm_crf / bert_crf read
 * (tools/layer.py:35 and :122,140 stop at seq_len) — and attention takes cu_seqlens.  ids / seg / out_f32 / out_bf16 /
 * d_out keep the padded [B*L, .] layout of the non-packed calls; [PAD] rows of the outputs are zero.  Same dropout seed
 * scheme (masks are indexed by packed element).  Workspaces: the two *_packed_* size functions below. */
size_t ner_bert_train_packed_saved_bytes(const ner_bert_config* cfg, int n_packed);
size_t ner_bert_train_packed_scratch_bytes(const ner_bert_config* cfg, int n_packed, int padded_rows);
int ner_bert_encoder_train_fwd_packed(const ner_bert_config* cfg, const float* word_emb, const float* type_emb,
                                      const float* pos_emb, const float* emb_ln_gamma, const float* emb_ln_beta,
                                      const ner_bert_layer_weights* layers, const int32_t* ids, const int32_t* seg,
                                      int B, int L, const int32_t* cu_seqlens, const int32_t* tok_src, int n_packed,
                                      float hidden_keep, float attn_keep, uint64_t seed, float* out_f32,
                                      void* out_bf16, void* saved, size_t saved_bytes, ner_stream_t stream);
int ner_bert_encoder_train_bwd_packed(const ner_bert_config* cfg, const float* emb_ln_gamma,
                                      const ner_bert_layer_weights* layers, const ner_bert_layer_grads* grads,
                                      float* d_word_emb, float* d_type_emb, float* d_pos_emb,
                                      float* d_emb_ln_gamma, float* d_emb_ln_beta, const int32_t* ids,
                                      const int32_t* seg, int B, int L, const int32_t* cu_seqlens,
                                      const int32_t* tok_src, int n_packed, float hidden_keep, float attn_keep,
                                      uint64_t seed, const float* d_out, const void* saved, size_t saved_bytes,
                                      void* scratch, size_t scratch_bytes, ner_stream_t stream);
/* Row moves between the padded and the packed layouts (row_bytes a multiple of 16):
 * gather: dst row r = src row idx[r];  scatter: dst row idx[r] = src row r (rows not named keep their content). */
int ner_gather_rows(const void* src, const int32_t* idx, void* dst, int n, int row_bytes, ner_stream_t stream);
int ner_scatter_rows(const void* src, const int32_t* idx, void* dst, int n, int row_bytes, ner_stream_t stream);
/* word[ids] + type[seg] + pos[0:L] without the LayerNorm -> f32 [B*L,H] (operand of the embedding
 * LayerNorm backward). */
int ner_bert_embed_sum(const float* word_emb, const float* type_emb, const float* pos_emb,
                       const int32_t* ids, const int32_t* seg, float* out, int B, int L, int H,
                       int vocab, int n_type, int max_pos, ner_stream_t stream);

/* The whole PREDICT step of model/bert_bilstm_crf.py:8-34 in one call (what a PREDICT session.run
 * of that plugin executes: BertModel -> bilstm -> dense(logits) -> crf_decode; the log-likelihood
 * is not fetched in PREDICT).  Same kernels, same order as the layer-by-layer entry points above.
 * lstm_wx_bf16 [8*lstm_hidden, H] = ner_pack_weight_bf16 of [kernel_fw[:H] | kernel_bw[:H]],
 * lstm_bias [8*lstm_hidden] = [bias_fw | bias_bw], lstm_wh_* = kernel[H:, :] f32,
 * lstm_activation as ner_bilstm_recurrence, logits_w [2*lstm_hidden, num_tags], trans [K,K].
 * n_packed = number of valid tokens (sum of mask), known on the host; pred_ids [B,L] i32. */
size_t ner_bert_bilstm_crf_predict_workspace_bytes(const ner_bert_config* cfg, int B, int L, int rows,
                                                   int lstm_hidden, int num_tags);
int ner_bert_bilstm_crf_predict(const ner_bert_config* cfg, const float* word_emb, const float* type_emb,
                                const float* pos_emb, const float* emb_ln_gamma, const float* emb_ln_beta,
                                const ner_bert_layer_weights* layers, const void* lstm_wx_bf16,
                                const float* lstm_bias, const float* lstm_wh_fw, const float* lstm_wh_bw,
                                int lstm_hidden, int lstm_activation, const float* logits_w,
                                const float* logits_b, const float* trans, int num_tags,
                                const int32_t* ids, const int32_t* mask, const int32_t* seg,
                                const int32_t* seq_len, int B, int L, int n_packed, int32_t* pred_ids,
                                void* workspace, size_t workspace_bytes, ner_stream_t stream);

/* ------------------------------------------------------------------------ *
 * BiLSTM — tools/layer.py:27-41 bilstm() -> bidirectional_dynamic_rnn(LSTMCell)
 * ------------------------------------------------------------------------ */

/* Sequential half of both directions.  xproj [B*L, 8H] f32 = x · [kernel_fw[:D] | kernel_bw[:D]]
 * + [bias_fw | bias_bw] (one ner_gemm_bf16 call, NER_EPI_F32); wh_fw / wh_bw = kernel[D:, :]
 * [H,4H] f32 with TF's gate order (i, j, f, o).  out [B,L,2H] f32 = concat(fw, bw), zero for
 * t >= seq_len.  activation: 0 tanh, 1 relu (params['rnn_activation']).  H % 4 == 0.
 * cu_seqlens: NULL (xproj row of (b,t) = b*L+t) or [B+1] (packed xproj: row = cu[b]+t).
 * gates_out [B*L, 8H] / cstate_out [B,L,2H]: both NULL (inference) or both given (training):
 * post-activation gates (sigmoid(i), act(j), sigmoid(f+forget_bias), sigmoid(o)) and cell states.
 * keep_prob < 1 (training, tools/layer.py:20-23 DropoutWrapper(output_keep_prob, state_keep_prob)):
 * independent counter-based masks (seed) on the emitted output and on the carried h; hstate_out
 * [B,L,2H] (nullable) receives the carried (state-dropped) h, the operand of dW_h. */
int ner_bilstm_recurrence(const float* xproj, const float* wh_fw, const float* wh_bw,
                          const int32_t* seq_len, float* out, int B, int L, int H,
                          int activation, float forget_bias, const int32_t* cu_seqlens,
                          float* gates_out, float* cstate_out, float* hstate_out, float keep_prob,
                          uint64_t seed, ner_stream_t stream);

/* Back-propagation through time of ner_bilstm_recurrence (padded layout).  d_out [B,L,2H] f32;
 * gates [B*L, 8H] / cstate [B,L,2H] saved by the forward call.  Writes d_xproj [B*L, 8H] f32 =
 * gradient w.r.t. the hoisted input projection (zeros for t >= seq_len).  The caller finishes
 * with plain GEMMs / reductions over it: dW_x = x^T d_xproj, d_bias = colsum(d_xproj),
 * dx = d_xproj W_x^T, dW_h = h_prev^T d_xproj (per direction). */
int ner_bilstm_recurrence_bwd(const float* d_out, const float* gates, const float* cstate,
                              const float* wh_fw, const float* wh_bw, const int32_t* seq_len,
                              float* d_xproj, int B, int L, int H, int activation, float keep_prob,
                              uint64_t seed, ner_stream_t stream);

/* ------------------------------------------------------------------------ *
 * SoftLexicon gather-and-pool — model/bilstm_crf_softlexicon.py:37-44
 * ------------------------------------------------------------------------ */

/* out[tok, g*E+e] = sum_s weights[tok, g*S+s] * table[ids[tok, g*S+s], e].
 * table [V,E] f32, ids/weights [n_tok, G*S], out [n_tok, G*E] with row stride ld_out
 * (>= G*E; lets the caller pool straight into a concat buffer).  G*S <= 64, E <= 128. */
int ner_softlexicon_pool_fwd(const float* table, const int32_t* ids, const float* weights,
                             float* out, int n_tok, int G, int S, int E, int V, int ld_out,
                             ner_stream_t stream);
/* d_table [V,E] += scatter of weights * d_out (caller zeroes / owns accumulation). */
int ner_softlexicon_pool_bwd(float* d_table, const int32_t* ids, const float* weights,
                             const float* d_out, int n_tok, int G, int S, int E, int V,
                             ner_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Training-side kernels — gradients of the layers above and the two optimizer steps of
 * tools/train_utils.py:246-390
 * ------------------------------------------------------------------------ */

/* src f32 [M,N] (row stride ld_src) -> bf16 [N,Mp] zero-padded: K-major operand of a
 * weight-gradient GEMM dW[K,N] = X^T dY (the reduction runs over the M rows). */
int ner_transpose_cast_bf16(const float* src, void* dst_bf16, int M, int N, int Mp, int ld_src,
                            ner_stream_t stream);
/* out[n] += scale * sum_m x[m,n]  (bias gradients). */
int ner_colsum_add(const float* x, float* out, int M, int N, int ld, float scale,
                   ner_stream_t stream);
/* Gradient of ner_dense_small_n (f32 x): dW [F,N] += x^T dy, db [N] += colsum(dy) (db may be
 * NULL), dx [M,F] = dy W^T (dx may be NULL).  dW/db are accumulated into (caller zeroes). */
int ner_dense_small_n_bwd(const float* x, const float* W, const float* dy, float* dW, float* db,
                          float* dx, int M, int F, int N, ner_stream_t stream);
/* tf.layers.dropout: y[i] = keep(seed, i) ? x[i]/keep_prob : 0.  Counter-based: the same
 * (seed, i) reproduces the mask, so the backward pass is the same call on the gradient. */
int ner_dropout(const float* x, float* y, size_t n, float keep_prob, uint64_t seed,
                ner_stream_t stream);
/* Same on bf16 tensors (BertModel's hidden dropout on the bf16 dense outputs; y may alias x). */
int ner_dropout_bf16(const void* x_bf16, void* y_bf16, size_t n, float keep_prob, uint64_t seed,
                     ner_stream_t stream);
/* tf.nn.relu on f32 [n] (y may alias x). */
int ner_relu_f32(const float* x, float* y, size_t n, ner_stream_t stream);
/* tf.nn.relu gradient: dpre = dact where act > 0 else 0 (f32 [n]; dpre may alias dact). */
int ner_relu_bwd_f32(const float* act, const float* dact, float* dpre, size_t n, ner_stream_t stream);
/* tf.reduce_max(x[B,L,C], axis=1) -> y[B,C] (reference model/bert_bilstm_crf_adv.py:35, the task discriminator's pool). */
int ner_reduce_max_time(const float* x, float* y, int B, int L, int C, ner_stream_t stream);
/* Its TF gradient, accumulated: dx[b,t,c] += scale * dy[b,c] / ties wherever x[b,t,c] == y[b,c]
 * (scale = -shrink_gradient_reverse folds FlipGradientBuilder, tools/train_utils.py:47-63). */
int ner_reduce_max_time_bwd(const float* x, const float* y, const float* dy, float* dx, int B, int L, int C,
                            float scale, ner_stream_t stream);
/* tf.nn.sparse_softmax_cross_entropy_with_logits (model/bert_bilstm_crf_adv.py:46) on [B, N <= 32]:
 * loss[b]; dlogits (nullable) = scale * (softmax - onehot). */
int ner_softmax_xent(const float* logits, const int32_t* labels, float* loss, float* dlogits, int B, int N,
                     float scale, ner_stream_t stream);
/* dst[i] += a * src[i]. */
int ner_axpy_f32(float* dst, const float* src, size_t n, float a, ner_stream_t stream);
/* out[0] += sum(g^2)  (tf.clip_by_global_norm, tools/train_utils.py:315).  Deterministic (no float atomics): per-CTA partial
 * sums go to `scratch` (>= ner_sumsq_scratch_floats() floats) and are added in index order, so identical gradients give a
 * bit-identical norm on every data-parallel rank and in every run. */
size_t ner_sumsq_scratch_floats(void);
int ner_sumsq_add(const float* g, size_t n, float* out, float* scratch, ner_stream_t stream);
/* One optimizer step over a flat parameter buffer.
 * mode 0 = AdamWeightDecayOptimizer (bert optimization.py; tools/train_utils.py:276-282):
 *   g *= grad_scale * clip / max(sqrt(*gnorm_sq) * grad_scale, clip)  (gnorm_sq NULL / clip 0: no clip);
 *   m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr * (m / (sqrt(v) + eps) + weight_decay * p).
 * mode 1 = tf.train.AdamOptimizer (tools/train_utils.py:340-350,365-390): g clipped to
 *   [-clip, clip] (clip 0: none), p -= lr * m / (sqrt(v) + eps) with lr the bias-corrected step
 *   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) computed by the caller. */
int ner_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int mode, float clip,
                  const float* gnorm_sq, float grad_scale, ner_stream_t stream);

/* ---- encoder backward (gradient of BertModel, tools/train_utils.py:314) ---- */

/* Backward of ner_layernorm: z = y (+ residual) is recomputed from the saved operands.
 * dz = dL/dz written as f32 (residual-branch gradient) and/or bf16 (A operand of the next dgrad
 * GEMM); d_gamma / d_beta [H] are accumulated into (caller zeroes). */
int ner_layernorm_bwd(const void* y, int y_is_bf16, const float* residual, const float* gamma,
                      const float* d_out, float* dz_f32, void* dz_bf16, float* d_gamma,
                      float* d_beta, int M, int H, float eps, ner_stream_t stream);
/* Backward of ner_layernorm_dropout: y is the UNdropped forward input; dz_f32 = gradient of the residual
 * branch, dz_bf16 = gradient w.r.t. y (masked like the forward). */
int ner_layernorm_dropout_bwd(const void* y, int y_is_bf16, const float* residual, const float* gamma,
                              const float* d_out, float* dz_f32, void* dz_bf16, float* d_gamma,
                              float* d_beta, int M, int H, float eps, float keep_prob, uint64_t seed,
                              ner_stream_t stream);
/* Same, and additionally d_bias[H] += column sums of the (masked) dense-branch gradient — the bias gradient of the dense
 * layer whose output this LayerNorm normalises — so the separate column-sum pass over dz_bf16 is not needed.  d_bias NULL =
 * ner_layernorm_dropout_bwd. */
int ner_layernorm_dropout_bwd_bias(const void* y, int y_is_bf16, const float* residual, const float* gamma,
                                   const float* d_out, float* dz_f32, void* dz_bf16, float* d_gamma, float* d_beta,
                                   float* d_bias, int M, int H, float eps, float keep_prob, uint64_t seed,
                                   ner_stream_t stream);
/* bf16 [M,N] -> bf16 [N,Mp] zero padded (K-major operands of weight-gradient GEMMs). */
int ner_transpose_bf16(const void* src_bf16, void* dst_bf16, int M, int N, int Mp,
                       ner_stream_t stream);
/* out[n] += sum_m x[m,n], x bf16 [M,N]  (bias gradients). */
int ner_colsum_bf16_add(const void* x_bf16, float* out, int M, int N, ner_stream_t stream);
/* GELU on a saved bf16 pre-activation (training forward) and its backward d_pre = d_act * gelu'(pre).
 * n % 4 == 0.  erf_variant: 0 tanh approximation, 1 erf. */
int ner_gelu_bf16(const void* pre_bf16, void* act_bf16, size_t n, int erf_variant, ner_stream_t stream);
/* Same on f32 (exact tanhf / erff): the FFN activation of the fp32-accurate BERT mode (y may alias x). */
int ner_gelu_f32(const float* x, float* y, size_t n, int erf_variant, ner_stream_t stream);
int ner_gelu_bwd_bf16(const void* pre_bf16, const void* dact_bf16, void* dpre_bf16, size_t n,
                      int erf_variant, ner_stream_t stream);
/* GELU backward on [M, N] with the bias gradient of the dense layer in front of the GELU fused in:
 * d_pre = d_act * gelu'(pre) and d_bias[N] += column sums of (the bf16-rounded) d_pre.  N % 8 == 0. */
int ner_gelu_bwd_bias_bf16(const void* pre_bf16, const void* dact_bf16, void* dpre_bf16, float* d_bias, int M, int N,
                           int erf_variant, ner_stream_t stream);
/* Embedding backward: scatter-add dx [B*L,H] f32 into d_word [vocab,H], d_type [n_type,H],
 * d_pos [>=L,H] (all accumulated into). */
int ner_bert_embed_bwd(const float* dx, const int32_t* ids, const int32_t* seg, float* d_word,
                       float* d_type, float* d_pos, int B, int L, int H, int vocab, int n_type,
                       ner_stream_t stream);
/* Backward of ner_bert_attention (padded layout, head_dim 64): qkv / ctx from the forward pass,
 * dctx = dL/dctx; writes d_qkv (bf16, layout of qkv).  Scores are recomputed, nothing L x L is stored;
 * (keep_prob, seed) must be the forward call's so the dropout mask is regenerated. */
int ner_bert_attention_bwd(const void* qkv_bf16, const int32_t* mask, const void* ctx_bf16,
                           const void* dctx_bf16, void* dqkv_bf16, int B, int L, int num_heads,
                           int head_dim, float scale, float mask_add, float keep_prob, uint64_t seed,
                           ner_stream_t stream);
/* Packed layout: sequence b occupies rows [cu_seqlens[b], cu_seqlens[b+1]) of qkv / ctx / dctx / dqkv, every key valid. */
int ner_bert_attention_bwd_packed(const void* qkv_bf16, const int32_t* cu_seqlens, const void* ctx_bf16,
                                  const void* dctx_bf16, void* dqkv_bf16, int B, int L, int num_heads,
                                  int head_dim, float scale, float keep_prob, uint64_t seed,
                                  ner_stream_t stream);

/* Weight gradients of several dense layers in ONE launch (tools/train_utils.py:314 `tf.gradients` w.r.t. the dense kernels):
 *   dw_p[k_in, n_out] (f32, accumulated into) += x_p^T . dy_p[:, dy_col0 : dy_col0 + n_out]
 * x_p bf16 [rows, ld_x] (the layer's input activations, first k_in columns used), dy_p bf16 [rows, ld_dy] (gradient w.r.t. the
 * layer's output).  Both are consumed as they lie (token-major = MN-major tcgen05 operands, 64 x 64 TMA boxes): no transposed
 * copies.  k_in % 128 == 0, n_out % 256 == 0, dy_col0 % 64 == 0, ld % 8 == 0; at most 6 problems per call. */
typedef struct {
  const void* x_bf16;
  int ld_x;
  const void* dy_bf16;
  int ld_dy;
  int dy_col0;
  float* dw;
  int k_in;
  int n_out;
} ner_wgrad_problem;
int ner_wgrad_group_bf16(const ner_wgrad_problem* problems_host, int count, int rows, ner_stream_t stream);

/* Data-parallel overlap hook (no reference counterpart — the reference is single device): events_host[l]
 * (cudaEvent_t, HOST array of n_events handles, NULL entries skipped) is recorded on the stream of the NEXT
 * ner_bert_encoder_train_bwd / _bwd_packed calls of this host thread as soon as every gradient of encoder layer l is
 * enqueued, so the caller can all-reduce that layer's slice of the flat gradient buffer while the backward pass of the
 * layers below still runs.  n_events = 0 clears the hook. */
int ner_bert_train_bwd_set_layer_events(void* const* events_host, int n_events);

/* tools/infer_utils.py:76-99  extract_entity — the tag-sequence half of it, on the device: pred_ids [B,L] i32 ->
 * per sentence the entity spans in order.  tag_class [K] u8 describes idx2tag: bits 0-1 kind (0 other, 1 'B', 2 'I' by
 * tag.split('-')[0]), bit 2 = the tag's first character is 'B' or 'I' (the reference's test on the previous tag),
 * bits 3-7 entity type id (index of tag.split('-')[1] in the caller's type list).  spans [B,cap] i32, each
 * start | end << 12 | type << 24 with end exclusive; counts [B] i32 = spans found (may exceed cap: the rest is dropped).
 * Reproduces the reference scan exactly, including its treatment of ill-formed sequences (an I after an I opens a span,
 * a span is typed by its LAST tag).  L <= 4095. */
int ner_extract_spans(const int32_t* pred_ids, const uint8_t* tag_class, int32_t* spans, int32_t* counts, int B, int L,
                      int K, int cap, ner_stream_t stream);

/* ------------------------------------------------------------------------ *
 * SoftLexicon HOST builder — replaces data/word_enhance.py:302-337 (build_soft_lexicon), :89-119 (align_with_token),
 * :163-205 (postproc_soft_lexicon) and data/base_preprocess.py:397-412 (format_soft_seq) for whole datasets at a time.
 * Host code (no CUDA call, no stream): all pointers are HOST pointers.  Output layout = the input of
 * ner_softlexicon_pool_fwd: [n_sent, max_seq_len, 4 (B,M,E,S), 10] ids / weights.
 * ------------------------------------------------------------------------ */
typedef struct ner_lexicon ner_lexicon;
/* Trie over the word vocabulary.  Words are UTF-32 code points back to back, word w = [offsets[w], offsets[w+1]); its id
 * is w (= embedding row).  freq[n_words + 2]: per-word frequency, then <None> (id n_words, reference: 1) and <PAD>
 * (id n_words + 1, reference: 0) — VocabModel._addon_token, data/word_enhance.py:62-70.  NULL on bad input. */
ner_lexicon* ner_lexicon_create(const uint32_t* word_codepoints_host, const int64_t* word_offsets_host,
                                const double* freq_host, int n_words);
void ner_lexicon_destroy(ner_lexicon* lexicon);
int64_t ner_lexicon_num_nodes(const ner_lexicon* lexicon);
/* Sentences are UTF-32 code points (spaces already removed, as build_soft_lexicon does), sentence s =
 * [sent_offsets[s], sent_offsets[s+1]).  tok_len / tok_offsets (both NULL, or per sentence the number of characters
 * each token covers, for WordPiece tokenizers): rows of characters one token swallowed are merged by set union.
 * bert_mode != 0: row 0 ([CLS]) and the rows from the [SEP] on stay zero and at most max_seq_len - 2 tokens are kept.
 * Every set holds its matches in first-seen order (the reference iterates Python sets: unordered), an empty set holds
 * <None>, sets are padded with <PAD> to 10 or cut to the 10 most frequent (stable), weights = freq / sum over the
 * token's 40 slots.  n_threads <= 0: one per hardware thread.  ids_out / weights_out: [n_sent, max_seq_len * 40]. */
int ner_lexicon_build(const ner_lexicon* lexicon, const uint32_t* codepoints_host, const int64_t* sent_offsets_host,
                      int n_sent, const int32_t* tok_len_host, const int64_t* tok_offsets_host, int max_seq_len,
                      int bert_mode, int32_t* ids_out_host, float* weights_out_host, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* NER_B200_H_ */
