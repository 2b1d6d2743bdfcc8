// CRF log-likelihood (gold-path score minus forward-alpha log-partition) for sm_100a —
// replaces tf.contrib.crf.crf_log_likelihood as called at reference tools/layer.py:122-127
// (crf_sequence_score / crf_log_norm semantics restated in SURVEY.md Appendix A.1).
//
// One thread per sequence; alpha[K] in registers.  Fast path (default, used when the
// transition matrix spans < 30 nats): the K*K logsumexp of one step is evaluated in the
// scaled-probability domain,
//     m = max_i alpha[i];  p[i] = 2^((alpha[i]-m)*log2e);
//     alpha'[j] = x[j] + (m + tmax) + ln2*lg2( sum_i p[i] * E[i][j] ),  E = exp(trans - tmax)
// (tmax = max of the whole transition matrix), i.e. K ex2 + K lg2 + K*K FFMA per step instead of
// K*K exp.  Chunks that lie completely inside a sequence (the common case) run a branch-free
// unrolled body; only the first and the ragged last chunk take the checked path.  The exact path (flags bit0, or
// chosen automatically for wide/inf transition matrices) evaluates every logsumexp with its
// own max, exactly as the reference's reduce_logsumexp does.
#include "crf_common.cuh"

namespace {

using namespace crf;

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int TAGP = 12;  // tag-chunk pitch (ints): 3 x 16B, odd -> conflict-free LDS.128

template <int K, int NT>
size_t loglik_smem_bytes() {
  using Gm = Geom<K>;
  size_t words = 2 * Gm::KK4 + 32 + NT + (size_t)NSTAGE * NT * Gm::P + (size_t)NSTAGE * NT * TAGP;
  return words * 4;
}

template <int NT>
__device__ __forceinline__ void stage_tags(int* dst, const int32_t* __restrict__ gbase, int L, int t0,
                                           int nv, const int* s_len, int vec16) {
  constexpr int T = T_CHUNK;
  const int steps = min(T, L - t0);
  if (vec16) {
    for (int idx = threadIdx.x; idx < NT * (T / 4); idx += NT) {
      const int r = idx / (T / 4), q = idx - r * (T / 4);
      if (r < nv && 4 * q < min(steps, s_len[r] - t0))
        cp_async16(dst + r * TAGP + 4 * q, gbase + (size_t)r * L + t0 + 4 * q);
    }
  } else {
    for (int idx = threadIdx.x; idx < NT * T; idx += NT) {
      const int r = idx / T, e = idx - r * T;
      if (r < nv && e < min(steps, s_len[r] - t0))
        cp_async4(dst + r * TAGP + e, gbase + (size_t)r * L + t0 + e);
    }
  }
}

template <int K, int NT>
__global__ void __launch_bounds__(NT)
crf_loglik_fwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ tags,
                      const int32_t* __restrict__ seq_len, const float* __restrict__ trans,
                      float* __restrict__ ll, float* __restrict__ logz_out,
                      float* __restrict__ alpha_ws, int B, int L, int vec_logits, int vec_tags,
                      int force_exact) {
  using Gm = Geom<K>;
  constexpr int T = Gm::T, G = Gm::G, P = Gm::P;
  constexpr bool E_REGS = (K <= 10);
  constexpr int UNR = Gm::UNROLL ? K : 1;

  extern __shared__ __align__(16) float smem[];
  float* s_tr = smem;                                   // raw trans [i][j]
  float* s_E = s_tr + Gm::KK4;                          // exp(trans - cmax[j]) stored [i][j]
  float* s_cmax = s_E + Gm::KK4;                        // [32]
  int* s_len = reinterpret_cast<int*>(s_cmax + 32);     // [NT]
  float* s_stage = reinterpret_cast<float*>(s_len + NT);
  int* s_tags = reinterpret_cast<int*>(s_stage + NSTAGE * NT * P);

  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * NT;
  const int nv = min(NT, B - row0);
  const int LK = L * K;

  for (int e = tid; e < K * K; e += NT) s_tr[e] = trans[e];
  int rawlen = 0, mylen = 1;
  if (tid < nv) {
    rawlen = seq_len[row0 + tid];
    mylen = min(max(rawlen, 1), L);
  }
  s_len[tid] = mylen;
  const int bmax = block_max_int<NT>(tid < nv ? mylen : 1, reinterpret_cast<int*>(s_stage));

  // column maxima, range test, E matrix (tiny; every thread helps)
  if (tid < K) {
    float cm = -INFINITY;
    for (int i = 0; i < K; ++i) cm = fmaxf(cm, s_tr[i * K + tid]);
    s_cmax[tid] = cm;
  }
  __syncthreads();
  bool fast = !force_exact;
  float tmax = 0.f;
  {
    float lo = INFINITY, hi = -INFINITY;
    for (int e = 0; e < K * K; ++e) {
      const float v = s_tr[e];
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
    if (!(hi - lo < 30.f) || !(fabsf(hi) < 1e30f) || !(fabsf(lo) < 1e30f)) fast = false;  // also NaN/inf
    tmax = fast ? hi : 0.f;
  }
  for (int e = tid; e < K * K; e += NT) s_E[e] = fast ? expf(s_tr[e] - tmax) : 0.f;
  __syncthreads();

  const float* gbase = logits + (size_t)row0 * LK;
  const int32_t* tbase = tags + (size_t)row0 * L;
  const int nchunk = (bmax + T - 1) / T;

#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nchunk) {
      stage_logits<K, NT>(s_stage + s * NT * P, gbase, LK, s * T, L, nv, s_len, vec_logits);
      stage_tags<NT>(s_tags + s * NT * TAGP, tbase, L, s * T, nv, s_len, vec_tags);
    }
    cp_async_commit();
  }

  float E[E_REGS ? K * K : 1];
  if (E_REGS) {
#pragma unroll
    for (int e = 0; e < K * K; ++e) E[e] = s_E[e];
  }
  float a[K];
#pragma unroll UNR
  for (int j = 0; j < K; ++j) a[j] = 0.f;
  // one forward-alpha step in the scaled-probability domain (no length checks)
  auto fast_step = [&](const float* x) {
    float m = a[0];
#pragma unroll UNR
    for (int i = 1; i < K; ++i) m = fmaxf(m, a[i]);
    const float nm2 = -m * kLog2e;
    float p[K];
#pragma unroll UNR
    for (int i = 0; i < K; ++i) p[i] = fast_ex2(fmaf(a[i], kLog2e, nm2));
    const float mt = m + tmax;
#pragma unroll UNR
    for (int j = 0; j < K; ++j) {
      float sum = 0.f;
#pragma unroll UNR
      for (int i = 0; i < K; ++i) sum = fmaf(p[i], E_REGS ? E[i * K + j] : s_E[i * K + j], sum);
      a[j] = fmaf(kLn2, fast_lg2(sum), x[j] + mt);
    }
  };

  float score = 0.f;
  int prev = 0;
  float* aws = (alpha_ws != nullptr && tid < nv) ? alpha_ws + (size_t)(row0 + tid) * LK : nullptr;

  for (int c = 0; c < nchunk; ++c) {
    const int cn = c + NSTAGE - 1;
    if (cn < nchunk) {
      stage_logits<K, NT>(s_stage + (cn % NSTAGE) * NT * P, gbase, LK, cn * T, L, nv, s_len, vec_logits);
      stage_tags<NT>(s_tags + (cn % NSTAGE) * NT * TAGP, tbase, L, cn * T, nv, s_len, vec_tags);
    }
    cp_async_commit();
    cp_async_wait<NSTAGE - 1>();
    __syncthreads();

    const int t0 = c * T;
    if (tid < nv && t0 < mylen) {
      const float* rowp = s_stage + (c % NSTAGE) * NT * P + tid * P;
      int tg[T];
      {
        const int4* tp = reinterpret_cast<const int4*>(s_tags + (c % NSTAGE) * NT * TAGP + tid * TAGP);
#pragma unroll
        for (int q = 0; q < T / 4; ++q) {
          const int4 v = tp[q];
          tg[4 * q + 0] = v.x;
          tg[4 * q + 1] = v.y;
          tg[4 * q + 2] = v.z;
          tg[4 * q + 3] = v.w;
        }
      }
      if (fast && t0 > 0 && t0 + T <= mylen) {
        // ---- whole chunk inside the sequence: branch-free body
#pragma unroll
        for (int g = 0; g < T / G; ++g) {
          float xs[G * K];
          load_group<K>(xs, rowp, g);
#pragma unroll
          for (int gg = 0; gg < G; ++gg) {
            const int tt = g * G + gg;
            const int tag = min(max(tg[tt], 0), K - 1);
            score += rowp[tt * K + tag] + s_tr[prev * K + tag];
            prev = tag;
            fast_step(xs + gg * K);
            if (aws != nullptr) {
#pragma unroll UNR
              for (int j = 0; j < K; ++j) aws[(size_t)(t0 + tt) * K + j] = a[j];
            }
          }
        }
      } else
#pragma unroll
      for (int g = 0; g < T / G; ++g) {
        if (t0 + g * G < mylen) {
          float xs[G * K];
          load_group<K>(xs, rowp, g);
#pragma unroll
          for (int gg = 0; gg < G; ++gg) {
            const int tt = g * G + gg;
            const int t = t0 + tt;
            if (t < mylen) {
              // ---- gold path (crf_unary_score + crf_binary_score)
              const int tag = min(max(tg[tt], 0), K - 1);
              score += rowp[tt * K + tag];
              if (t > 0) score += s_tr[prev * K + tag];
              prev = tag;
              // ---- forward-alpha (crf_log_norm)
              if (t == 0) {
#pragma unroll UNR
                for (int j = 0; j < K; ++j) a[j] = xs[gg * K + j];
              } else if (fast) {
                fast_step(xs + gg * K);
              } else {
                float na[K];
#pragma unroll UNR
                for (int j = 0; j < K; ++j) {
                  float m = -INFINITY;
#pragma unroll UNR
                  for (int i = 0; i < K; ++i) m = fmaxf(m, a[i] + s_tr[i * K + j]);
                  const float mm = (fabsf(m) <= 3.0e38f) ? m : 0.f;  // reduce_logsumexp's finite-max guard
                  float sum = 0.f;
#pragma unroll UNR
                  for (int i = 0; i < K; ++i) sum += expf(a[i] + s_tr[i * K + j] - mm);
                  na[j] = xs[gg * K + j] + (logf(sum) + mm);
                }
#pragma unroll UNR
                for (int j = 0; j < K; ++j) a[j] = na[j];
              }
              if (aws != nullptr) {
#pragma unroll UNR
                for (int j = 0; j < K; ++j) aws[(size_t)t * K + j] = a[j];
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }

  if (tid < nv) {
    float m = a[0];
#pragma unroll UNR
    for (int j = 1; j < K; ++j) m = fmaxf(m, a[j]);
    const float mm = (fabsf(m) <= 3.0e38f) ? m : 0.f;
    float sum = 0.f;
#pragma unroll UNR
    for (int j = 0; j < K; ++j) sum += expf(a[j] - mm);
    float logz = logf(sum) + mm;
    if (rawlen <= 0) {  // crf_log_norm / crf_sequence_score: zero for empty sequences
      logz = 0.f;
      score = 0.f;
    }
    ll[row0 + tid] = score - logz;
    if (logz_out != nullptr) logz_out[row0 + tid] = logz;
  }
}

template <int K, int NT>
int launch_fwd_nt(const float* logits, const int32_t* tags, const int32_t* seq_len,
                  const float* trans, float* ll, float* logz, float* alpha_ws, int B, int L,
                  int flags, cudaStream_t st) {
  const size_t smem = loglik_smem_bytes<K, NT>();
  auto kern = crf_loglik_fwd_kernel<K, NT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  const int vl = ((L * K) % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int vt = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(tags) & 15) == 0);
  const int grid = (B + NT - 1) / NT;
  kern<<<grid, NT, smem, st>>>(logits, tags, seq_len, trans, ll, logz, alpha_ws, B, L, vl, vt, flags & 1);
  return ner_launch_status();
}

template <int K>
int launch_fwd(const float* logits, const int32_t* tags, const int32_t* seq_len, const float* trans,
               float* ll, float* logz, float* alpha_ws, int B, int L, int flags, cudaStream_t st) {
  if (B > 148 * 64 * 2) return launch_fwd_nt<K, 64>(logits, tags, seq_len, trans, ll, logz, alpha_ws, B, L, flags, st);
  return launch_fwd_nt<K, 32>(logits, tags, seq_len, trans, ll, logz, alpha_ws, B, L, flags, st);
}

}  // namespace

extern "C" int ner_crf_loglik_fwd(const float* logits, const int32_t* tags, const int32_t* seq_len,
                                  const float* trans, float* ll, float* logz_out, float* alpha_ws,
                                  int B, int L, int K, int flags, ner_stream_t stream) {
  if (B < 0 || L < 1 || K < 1) return NER_ERR_INVALID_ARG;
  if (B == 0) return NER_OK;
  if (!logits || !tags || !seq_len || !trans || !ll) return NER_ERR_INVALID_ARG;
  if (K > NER_MAX_TAGS) return NER_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B <= NER_CRF_SMALL_B && !(flags & 2)) {  // flags bit1: force the throughput kernel (tests / benches)
    const int rc = ner_crf_loglik_fwd_small(logits, tags, seq_len, trans, ll, logz_out, alpha_ws, B, L, K, st);
    if (rc != NER_ERR_UNSUPPORTED) return rc;
  }
#define CALL(KK) return launch_fwd<KK>(logits, tags, seq_len, trans, ll, logz_out, alpha_ws, B, L, flags, st)
  NER_CRF_DISPATCH_K(K, CALL)
#undef CALL
  return NER_ERR_UNSUPPORTED;
}
