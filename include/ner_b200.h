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

/* out[M,N] = epilogue(A[M,K] · Wt[N,K]^T + bias[N]).  A and Wt are bf16,
 * K contiguous (Wt is the TF kernel [K,N] transposed once by
 * ner_pack_weight_bf16).  bias may be NULL.  K % 8 == 0, N % 32 == 0.
 * tile_n: 0 = auto, or 64/128/256. */
int ner_gemm_bf16(const void* A, const void* Wt, const float* bias, const float* residual,
                  void* out, int M, int N, int K, int epilogue, int tile_n,
                  ner_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NER_B200_H_ */
