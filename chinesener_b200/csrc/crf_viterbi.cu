// CRF Viterbi decode for sm_100a — replaces tf.contrib.crf.crf_decode as called at
// reference tools/layer.py:140-142 (semantics restated in SURVEY.md Appendix A.1).
//
// One thread per sequence, K-wide max-plus state in registers, transition matrix in
// registers (K <= 10) or broadcast smem reads.  Backpointers never leave the SM: they are
// packed 4 bits (K <= 16) or 8 bits per tag into smem words laid out [t][w][thread], and
// the decoded tags are written back over word 0 of the same slots during the backtrace, so
// the final [NT, L] int32 store to HBM is fully coalesced.
//
// Bit-exactness contract (tests/test_crf_gpu.py): fp32 adds in the reference order
// (s[i] + trans[i][j], max over i, then + logits[t][j]); ties -> lowest index (strict >).
#include "crf_common.cuh"

namespace {

using namespace crf;

template <int K>
struct BpPack {
  static constexpr int NIB = (K <= 16) ? 4 : 8;
  static constexpr int PER = 32 / NIB;
  static constexpr int W = (K + PER - 1) / PER;
  static constexpr uint32_t MASK = (1u << NIB) - 1u;
};

template <int K, int NT>
size_t viterbi_smem_bytes(int L) {
  using Gm = Geom<K>;
  size_t words = Gm::KK4 + NT + (size_t)NSTAGE * NT * Gm::P + (size_t)L * BpPack<K>::W * (NT + 1);
  return words * 4;
}

template <int K, int NT>
__global__ void __launch_bounds__(NT)
crf_viterbi_kernel(const float* __restrict__ logits, const int32_t* __restrict__ seq_len,
                   const float* __restrict__ trans, int32_t* __restrict__ tags_out,
                   float* __restrict__ best_score, int B, int L, int vec16) {
  using Gm = Geom<K>;
  using Bp = BpPack<K>;
  constexpr int T = Gm::T, G = Gm::G, P = Gm::P, W = Bp::W, NTP = NT + 1;
  constexpr bool TR_REGS = (K <= 10);
  constexpr int UNR = Gm::UNROLL ? K : 1;

  extern __shared__ __align__(16) float smem[];
  float* s_trT = smem;                                        // [j][i]
  int* s_len = reinterpret_cast<int*>(s_trT + Gm::KK4);       // [NT]
  float* s_stage = reinterpret_cast<float*>(s_len + NT);      // [NSTAGE][NT][P]
  uint32_t* s_bp = reinterpret_cast<uint32_t*>(s_stage + NSTAGE * NT * P);  // [L][W][NTP]

  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * NT;
  const int nv = min(NT, B - row0);
  const int LK = L * K;

  for (int e = tid; e < K * K; e += NT) {
    const int i = e / K, j = e - i * K;
    s_trT[j * K + i] = trans[e];
  }
  int mylen = 1;
  if (tid < nv) mylen = min(max(seq_len[row0 + tid], 1), L);  // len<=0 behaves like 1 (TF quirk)
  s_len[tid] = mylen;
  const int bmax = block_max_int<NT>(tid < nv ? mylen : 1, reinterpret_cast<int*>(s_bp));
  // (block_max_int ends with __syncthreads: s_len / s_trT are visible)

  const float* gbase = logits + (size_t)row0 * LK;
  const int nchunk = (bmax + T - 1) / T;

#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nchunk) stage_logits<K, NT>(s_stage + s * NT * P, gbase, LK, s * T, L, nv, s_len, vec16);
    cp_async_commit();
  }

  float tr[TR_REGS ? K * K : 1];
  if (TR_REGS) {
#pragma unroll
    for (int e = 0; e < K * K; ++e) tr[e] = s_trT[e];
  }

  float s[K];
#pragma unroll UNR
  for (int j = 0; j < K; ++j) s[j] = 0.f;

  for (int c = 0; c < nchunk; ++c) {
    const int cn = c + NSTAGE - 1;
    if (cn < nchunk)
      stage_logits<K, NT>(s_stage + (cn % NSTAGE) * NT * P, gbase, LK, cn * T, L, nv, s_len, vec16);
    cp_async_commit();
    cp_async_wait<NSTAGE - 1>();
    __syncthreads();

    const int t0 = c * T;
    if (tid < nv && t0 < mylen) {
      const float* rowp = s_stage + (c % NSTAGE) * NT * P + tid * P;
#pragma unroll 1
      for (int g = 0; g < T / G; ++g) {
        if (t0 + g * G >= mylen) break;
        float xs[G * K];
        load_group<K>(xs, rowp, g);
#pragma unroll
        for (int gg = 0; gg < G; ++gg) {
          const int t = t0 + g * G + gg;
          if (t < mylen) {
            if (t == 0) {
#pragma unroll UNR
              for (int j = 0; j < K; ++j) s[j] = xs[gg * K + j];
            } else {
              // i-outer order: K independent (best, arg) pairs -> K-way ILP; per (i,j) the fp32 ops and the
              // strict '>' (first max wins) are exactly those of the reference recursion
              float best[K];
              int arg[K];
              uint32_t bpw[W];
#pragma unroll
              for (int w = 0; w < W; ++w) bpw[w] = 0u;
#pragma unroll UNR
              for (int j = 0; j < K; ++j) {
                best[j] = s[0] + (TR_REGS ? tr[j * K] : s_trT[j * K]);
                arg[j] = 0;
              }
#pragma unroll UNR
              for (int i = 1; i < K; ++i) {
#pragma unroll UNR
                for (int j = 0; j < K; ++j) {
                  const float v = s[i] + (TR_REGS ? tr[j * K + i] : s_trT[j * K + i]);
                  if (v > best[j]) {
                    best[j] = v;
                    arg[j] = i;
                  }
                }
              }
#pragma unroll UNR
              for (int j = 0; j < K; ++j) {
                s[j] = xs[gg * K + j] + best[j];
                bpw[j / Bp::PER] |= (uint32_t)arg[j] << ((j % Bp::PER) * Bp::NIB);
              }
#pragma unroll
              for (int w = 0; w < W; ++w) s_bp[(t * W + w) * NTP + tid] = bpw[w];
            }
          }
        }
      }
    }
    __syncthreads();
  }

  if (tid < nv) {
    float best = s[0];
    int y = 0;
#pragma unroll UNR
    for (int j = 1; j < K; ++j)
      if (s[j] > best) {
        best = s[j];
        y = j;
      }
    if (best_score != nullptr) best_score[row0 + tid] = best;
    for (int t = mylen - 1; t >= 1; --t) {
      const uint32_t w = s_bp[(t * W + y / Bp::PER) * NTP + tid];
      const int prev = (int)((w >> ((y % Bp::PER) * Bp::NIB)) & Bp::MASK);
      s_bp[(t * W) * NTP + tid] = (uint32_t)y;
      y = prev;
    }
    s_bp[tid] = (uint32_t)y;
  }
  __syncthreads();

  // Coalesced [nv, L] int32 store; zero beyond each row's length.
  int32_t* obase = tags_out + (size_t)row0 * L;
  int r = 0, p = tid;
  while (p >= L) {
    p -= L;
    ++r;
  }
  const int total = nv * L;
  for (int idx = tid; idx < total; idx += NT) {
    const int v = (p < s_len[r]) ? (int)s_bp[(p * W) * NTP + r] : 0;
    obase[idx] = v;
    p += NT;
    while (p >= L) {
      p -= L;
      ++r;
    }
  }
}

template <int K, int NT>
int launch_viterbi_nt(const float* logits, const int32_t* seq_len, const float* trans,
                      int32_t* tags_out, float* best_score, int B, int L, cudaStream_t st) {
  const size_t smem = viterbi_smem_bytes<K, NT>(L);
  auto kern = crf_viterbi_kernel<K, NT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  const int vec16 = ((L * K) % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int grid = (B + NT - 1) / NT;
  kern<<<grid, NT, smem, st>>>(logits, seq_len, trans, tags_out, best_score, B, L, vec16);
  return ner_launch_status();
}

constexpr size_t kMaxSmem = 227 * 1024;

template <int K>
int launch_viterbi(const float* logits, const int32_t* seq_len, const float* trans,
                   int32_t* tags_out, float* best_score, int B, int L, cudaStream_t st) {
  // Large batches: 128 sequences per CTA when the backpointer slab fits; small batches
  // spread over more SMs with 32-sequence CTAs.
  const bool big = B > 148 * 32 * 2;
  if (big && viterbi_smem_bytes<K, 128>(L) <= kMaxSmem)
    return launch_viterbi_nt<K, 128>(logits, seq_len, trans, tags_out, best_score, B, L, st);
  if (viterbi_smem_bytes<K, 32>(L) <= kMaxSmem)
    return launch_viterbi_nt<K, 32>(logits, seq_len, trans, tags_out, best_score, B, L, st);
  return NER_ERR_UNSUPPORTED;  // L too long for on-chip backpointers
}

}  // namespace

extern "C" int ner_crf_viterbi(const float* logits, const int32_t* seq_len, const float* trans,
                               int32_t* tags_out, float* best_score, int B, int L, int K,
                               ner_stream_t stream) {
  if (B < 0 || L < 1 || K < 1) return NER_ERR_INVALID_ARG;
  if (B == 0) return NER_OK;
  if (!logits || !seq_len || !trans || !tags_out) return NER_ERR_INVALID_ARG;
  if (K > NER_MAX_TAGS) return NER_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B <= NER_CRF_SMALL_B) {
    const int rc = ner_crf_viterbi_small(logits, seq_len, trans, tags_out, best_score, B, L, K, st);
    if (rc != NER_ERR_UNSUPPORTED) return rc;
  }
#define CALL(KK) return launch_viterbi<KK>(logits, seq_len, trans, tags_out, best_score, B, L, st)
  NER_CRF_DISPATCH_K(K, CALL)
#undef CALL
  return NER_ERR_UNSUPPORTED;
}
