/* Plain-C restatement of the CRF decode / log-likelihood of the hot path — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Same algorithm and the same arithmetic order as oracle/crf.py (numpy), which restates tf.contrib.crf of TensorFlow 1.14
 * as the reference calls it: tools/layer.py:140-142 (crf_decode) and tools/layer.py:122-127 (crf_log_likelihood);
 * semantics per SURVEY.md Appendix A.1.  It exists so that the checker finishes on ALL rows of the roofline-sized
 * launches (B = 262 144 sequences) in seconds; tests/test_oracle_native.py pins it to the numpy restatement bit for bit
 * (Viterbi, float32) and to 1e-12 (log-likelihood, float64; libm exp/log vs numpy's).  PARITY STATUS: like the rest of
 * oracle/, the numeric values are not pinned by any fixture of the reference ("parity unpinned"); the reference-held
 * artefacts pin the zero fill beyond seq_len and the F1 tables (tests/test_golden.py).
 *
 * Build (oracle/native.py does this; no FMA contraction so float32 sums round as numpy's do):
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC -o oracle/liboracle_crf.so oracle/crf_c.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_K 64

/* tf.contrib.crf.crf_decode (reference tools/layer.py:140): float32 max-plus recursion, lowest index on ties, tags zero
 * beyond seq_len, seq_len <= 0 decodes like 1, L == 1 is a plain argmax.  x [B,T,K], trans [K,K] (trans[i][j] = i -> j),
 * lens [B]; tags [B,T] int32, best [B] float32 (nullable).  Returns 0, or -1 on a bad argument. */
int oracle_crf_decode_f32(const float* x, const float* trans, const int32_t* lens, int32_t* tags, float* best,
                          int64_t B, int T, int K) {
  if (!x || !trans || !lens || !tags || B < 0 || T < 1 || K < 1 || K > ORACLE_MAX_K) return -1;
  int fail = 0;
#pragma omp parallel
  {
    uint8_t* bp = (uint8_t*)malloc((size_t)T * K);
    if (!bp) {
#pragma omp atomic write
      fail = 1;
    }
#pragma omp for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
      if (!bp) continue;
      const float* xb = x + (size_t)b * T * K;
      int32_t* tb = tags + (size_t)b * T;
      int len = lens[b] < 0 ? 0 : (lens[b] > T ? T : lens[b]);
      float s[ORACLE_MAX_K], nw[ORACLE_MAX_K];
      memset(tb, 0, sizeof(int32_t) * (size_t)T);
      for (int j = 0; j < K; ++j) s[j] = xb[j];
      const int lm1 = len - 1 > 0 ? len - 1 : 0;
      if (T > 1)
        for (int t = 1; t < T && (t - 1) < lm1; ++t) {         /* dynamic_rnn over inputs[1:], length len - 1 */
          for (int j = 0; j < K; ++j) {
            float m = s[0] + trans[j];
            int arg = 0;
            for (int i = 1; i < K; ++i) {
              const float v = s[i] + trans[i * K + j];           /* (s[i] + trans[i][j]) first ... */
              if (v > m) { m = v; arg = i; }                     /* strict: the first maximum stays */
            }
            nw[j] = xb[(size_t)t * K + j] + m;                   /* ... then + potentials */
            bp[(size_t)t * K + j] = (uint8_t)arg;
          }
          memcpy(s, nw, sizeof(float) * (size_t)K);
        }
      float m = s[0];
      int y = 0;
      for (int j = 1; j < K; ++j)
        if (s[j] > m) { m = s[j]; y = j; }
      if (best) best[b] = m;
      const int n = len > 1 ? len : 1;                           /* len 0 decodes like len 1 (reverse_sequence no-op) */
      tb[n - 1] = y;
      for (int t = n - 1; t >= 1; --t) {
        y = bp[(size_t)t * K + y];
        tb[t - 1] = y;
      }
    }
    free(bp);
  }
  return fail ? -2 : 0;
}

static double lse(const double* v, int n) {                       /* tf.reduce_logsumexp: max-subtract, finite-max guard */
  double m = v[0];
  for (int i = 1; i < n; ++i)
    if (v[i] > m) m = v[i];
  if (!isfinite(m)) m = 0.0;
  double acc = 0.0;
  for (int i = 0; i < n; ++i) acc += exp(v[i] - m);
  return log(acc) + m;
}

/* tf.contrib.crf.crf_log_likelihood (reference tools/layer.py:122) in float64 on float32 inputs:
 * ll[b] = unary + binary score of `tags` (masked by t < len) - log Z; alpha frozen beyond len; len <= 0 -> log Z = 0.
 * x [B,T,K] float32, tag_indices [B,T] int32, ll [B] float64. */
int oracle_crf_loglik_f64(const float* x, const int32_t* tag_indices, const int32_t* lens, const float* trans, double* ll,
                          int64_t B, int T, int K) {
  if (!x || !tag_indices || !lens || !trans || !ll || B < 0 || T < 1 || K < 1 || K > ORACLE_MAX_K) return -1;
  int bad = 0;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    const float* xb = x + (size_t)b * T * K;
    const int32_t* yb = tag_indices + (size_t)b * T;
    const int len = lens[b];                                      /* the mask is t < len: values above T behave like T */
    double alpha[ORACLE_MAX_K], nw[ORACLE_MAX_K], col[ORACLE_MAX_K];
    double score = 0.0;
    int ok = 1;
    for (int t = 0; t < T; ++t)
      if (yb[t] < 0 || yb[t] >= K) ok = 0;
    if (!ok) {
#pragma omp atomic write
      bad = 1;
      ll[b] = NAN;
      continue;
    }
    for (int t = 0; t < T && t < len; ++t) score += (double)xb[(size_t)t * K + yb[t]];
    for (int t = 1; t < T && t < len; ++t) score += (double)trans[yb[t - 1] * K + yb[t]];
    if (T == 1 && len <= 0) score = 0.0;
    for (int j = 0; j < K; ++j) alpha[j] = (double)xb[j];
    for (int t = 1; t < T && t < len; ++t) {
      for (int j = 0; j < K; ++j) {
        for (int i = 0; i < K; ++i) col[i] = alpha[i] + (double)trans[i * K + j];
        nw[j] = (double)xb[(size_t)t * K + j] + lse(col, K);
      }
      memcpy(alpha, nw, sizeof(double) * (size_t)K);
    }
    const double logz = len <= 0 ? 0.0 : lse(alpha, K);
    ll[b] = score - logz;
  }
  return bad ? -3 : 0;
}

int oracle_crf_abi(void) { return 1; }
