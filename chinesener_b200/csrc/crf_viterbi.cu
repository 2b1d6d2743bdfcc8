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
#include <stdlib.h>

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

// ---------------------------------------------------------------------------------------------
// Occupancy-first variant (K <= 16): the 4-bit backpointers of tags 0..7 (one 32-bit word per
// step) are parked in the CTA's own [nv, L] slab of tags_out — exactly 4 bytes per step per
// sequence, written and read back by the same thread (coalesced [t][thread] layout, L2-resident
// for the CTA's lifetime) — and only the nibbles of tags 8..K-1 stay in shared memory.  Shared
// memory per CTA drops from ~219 KB to ~40 KB, so several CTAs share an SM and the per-step
// dependency chain of one warp is hidden behind the others (the all-on-chip kernel ran one warp
// per scheduler at IPC 1.5).  The s[i]+trans[i][j] adds are issued as FADD2 pairs; values, the
// strict '>' tie rule and the add order are unchanged (bit-exact contract above).
template <int K>
struct BpSplit {
  static constexpr int HB = (K <= 8) ? 0 : (K <= 10 ? 1 : (K <= 12 ? 2 : 4));  // on-chip bytes per step
};

template <int K, int NT, int TT>
size_t viterbi_gs_smem_bytes(int L) {
  using Gm = Geom<K, TT>;
  const size_t stage = (size_t)NSTAGE * NT * Gm::P * 4;
  const size_t dec = (size_t)NT * (((L + 3) & ~3) + 4);
  const size_t hi = ((size_t)L * NT * BpSplit<K>::HB + 15) & ~(size_t)15;
  return (size_t)(2 * K * ((K + 1) / 2) + NT) * 4 + (stage > dec ? stage : dec) + hi;
}

template <int K, int NT, int TT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
crf_viterbi_gs_kernel(const float* __restrict__ logits, const int32_t* __restrict__ seq_len,
                      const float* __restrict__ trans, int32_t* tags_out,
                      float* __restrict__ best_score, int B, int L, int vec16) {
  using Gm = Geom<K, TT>;
  constexpr int T = Gm::T, G = Gm::G, P = Gm::P, KP = (K + 1) / 2, HB = BpSplit<K>::HB;
  constexpr bool TR_REGS = (K <= 10);
  constexpr int UNR = Gm::UNROLL ? K : 1;

  extern __shared__ __align__(16) float smem[];
  float* s_tr = smem;                                           // [i][2*KP]: trans rows, pad column = 0
  int* s_len = reinterpret_cast<int*>(s_tr + 2 * K * KP);       // [NT]
  float* s_stage = reinterpret_cast<float*>(s_len + NT);        // [NSTAGE][NT][P]; reused as s_dec
  const int Lp = ((L + 3) & ~3) + 4;                            // byte pitch of a decoded row (Lp/4 odd-ish)
  const size_t stage_b = (size_t)NSTAGE * NT * P * 4, dec_b = (size_t)NT * Lp;
  uint8_t* s_dec = reinterpret_cast<uint8_t*>(s_stage);         // [NT][Lp] decoded tags (after the forward loop)
  uint8_t* s_hi = reinterpret_cast<uint8_t*>(s_stage) + (stage_b > dec_b ? stage_b : dec_b);  // [L][NT] x HB bytes

  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * NT;
  const int nv = min(NT, B - row0);
  const int LK = L * K;

  for (int e = tid; e < K * 2 * KP; e += NT) {
    const int i = e / (2 * KP), j = e - i * 2 * KP;
    s_tr[e] = (j < K) ? trans[i * K + j] : 0.f;
  }
  int mylen = 1;
  if (tid < nv) mylen = min(max(seq_len[row0 + tid], 1), L);  // len<=0 behaves like 1 (TF quirk)
  s_len[tid] = mylen;
  const int bmax = block_max_int<NT>(tid < nv ? mylen : 1, reinterpret_cast<int*>(s_stage));

  const float* gbase = logits + (size_t)row0 * LK;
  uint32_t* scratch = reinterpret_cast<uint32_t*>(tags_out + (size_t)row0 * L);  // [L][nv] words of this CTA
  const int nchunk = (bmax + T - 1) / T;

#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nchunk) stage_logits<K, NT, TT>(s_stage + s * NT * P, gbase, LK, s * T, L, nv, s_len, vec16);
    cp_async_commit();
  }

  f32x2 tr2[TR_REGS ? K * KP : 1];
  if (TR_REGS) {
#pragma unroll
    for (int e = 0; e < K * KP; ++e) tr2[e] = pk2(s_tr[2 * e], s_tr[2 * e + 1]);
  }
  auto trp = [&](int i, int q) -> f32x2 {
    if (TR_REGS) return tr2[i * KP + q];
    return pk2(s_tr[(i * KP + q) * 2], s_tr[(i * KP + q) * 2 + 1]);
  };

  float s[2 * KP];
#pragma unroll UNR
  for (int j = 0; j < 2 * KP; ++j) s[j] = 0.f;

  for (int c = 0; c < nchunk; ++c) {
    const int cn = c + NSTAGE - 1;
    if (cn < nchunk)
      stage_logits<K, NT, TT>(s_stage + (cn % NSTAGE) * NT * P, gbase, LK, cn * T, L, nv, s_len, vec16);
    cp_async_commit();
    cp_async_wait<NSTAGE - 1>();
    __syncthreads();

    const int t0 = c * T;
    if (tid < nv && t0 < mylen) {
      const float* rowp = s_stage + (c % NSTAGE) * NT * P + tid * P;
#pragma unroll
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
              // i-outer order: K independent (best, arg) pairs -> K-way ILP; per (i,j) the fp32 add and the
              // strict '>' (first max wins) are exactly those of the reference recursion
              float best[2 * KP];
              int arg[2 * KP];
#pragma unroll UNR
              for (int q = 0; q < KP; ++q) {
                upk2(add2(pk2(s[0], s[0]), trp(0, q)), best[2 * q], best[2 * q + 1]);
                arg[2 * q] = 0;
                arg[2 * q + 1] = 0;
              }
#pragma unroll UNR
              for (int i = 1; i < K; ++i) {
#pragma unroll UNR
                for (int q = 0; q < KP; ++q) {
                  float lo, hi;
                  upk2(add2(pk2(s[i], s[i]), trp(i, q)), lo, hi);
                  if (lo > best[2 * q]) {
                    best[2 * q] = lo;
                    arg[2 * q] = i;
                  }
                  if (2 * q + 1 < K && hi > best[2 * q + 1]) {
                    best[2 * q + 1] = hi;
                    arg[2 * q + 1] = i;
                  }
                }
              }
              uint32_t wlo = 0u, whi = 0u;
#pragma unroll UNR
              for (int j = 0; j < K; ++j) {
                s[j] = xs[gg * K + j] + best[j];
                if (j < 8)
                  wlo |= (uint32_t)arg[j] << (4 * j);
                else
                  whi |= (uint32_t)arg[j] << (4 * (j - 8));
              }
              __stcg(scratch + (size_t)t * nv + tid, wlo);
              if (HB == 1) s_hi[t * NT + tid] = (uint8_t)whi;
              if (HB == 2) reinterpret_cast<uint16_t*>(s_hi)[t * NT + tid] = (uint16_t)whi;
              if (HB == 4) reinterpret_cast<uint32_t*>(s_hi)[t * NT + tid] = whi;
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
    uint8_t* drow = s_dec + tid * Lp;
    // Backtrace: the parked words are at addresses independent of the path, so fetch 8 steps per
    // round trip to L2 and resolve the chain in registers.
    for (int t = mylen - 1; t >= 1; t -= 8) {
      uint32_t wa[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) wa[u] = (t - u >= 1) ? __ldcg(scratch + (size_t)(t - u) * nv + tid) : 0u;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int tt = t - u;
        if (tt >= 1) {
          uint32_t w = wa[u];
          if (HB > 0 && y >= 8) {
            if (HB == 1) w = s_hi[tt * NT + tid];
            if (HB == 2) w = reinterpret_cast<const uint16_t*>(s_hi)[tt * NT + tid];
            if (HB == 4) w = reinterpret_cast<const uint32_t*>(s_hi)[tt * NT + tid];
          }
          drow[tt] = (uint8_t)y;
          y = (int)((w >> (4 * (y & 7))) & 15u);
        }
      }
    }
    drow[0] = (uint8_t)y;
  }
  __syncthreads();   // every parked word has been consumed: the slab can now take the decoded tags

  // Coalesced [nv, L] int32 store; zero beyond each row's length.
  int32_t* obase = tags_out + (size_t)row0 * L;
  int r = 0, p = tid;
  while (p >= L) {
    p -= L;
    ++r;
  }
  const int total = nv * L;
  for (int idx = tid; idx < total; idx += NT) {
    obase[idx] = (p < s_len[r]) ? (int)s_dec[r * Lp + p] : 0;
    p += NT;
    while (p >= L) {
      p -= L;
      ++r;
    }
  }
}

template <int K, int NT, int TT, int MINB>
int launch_viterbi_gs(const float* logits, const int32_t* seq_len, const float* trans,
                      int32_t* tags_out, float* best_score, int B, int L, cudaStream_t st) {
  const size_t smem = viterbi_gs_smem_bytes<K, NT, TT>(L);
  if (smem > 227 * 1024) return NER_ERR_UNSUPPORTED;
  auto kern = crf_viterbi_gs_kernel<K, NT, TT, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  const int vec16 = ((L * K) % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int grid = (B + NT - 1) / NT;
  kern<<<grid, NT, smem, st>>>(logits, seq_len, trans, tags_out, best_score, B, L, vec16);
  return ner_launch_status();
}

int vit_variant() {
  const char* e = getenv("NER_CRF_VIT_VARIANT");   // tuning / test hook, read per call
  return e ? atoi(e) : 0;
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
  if constexpr (K <= 16) {
    if (big && vit_variant() != 1) {   // NER_CRF_VIT_VARIANT=1: time the all-on-chip kernel instead
      const int rc = launch_viterbi_gs<K, 64, 4, 6>(logits, seq_len, trans, tags_out, best_score, B, L, st);
      if (rc != NER_ERR_UNSUPPORTED) return rc;
    }
  }
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
