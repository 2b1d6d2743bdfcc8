// fp32 self-attention for small heads, with the TENER relative-position term (sm_100a, SIMT).
//
// Replaces reference tools/transformer/tener.py:12-119 (relative_attention + shift +
// normalize_attention + weighted value) for model/transformer_tener_crf_bichar.py:
//     scores[q,k] = (Q_q + u_h)·K_k + (Q_q + v_h)·R_{k-q+L}      (unscaled, K unprojected)
//     out_q = softmax_k(scores + (1-mask_k)·(-2^32+1)) · V
// R = sinusoidal table over positions -L..L-1 (tools/transformer/modules.py:177-197); the
// reference's zero-pad/reshape `shift` is the index identity BD'[q,k] = BD[q, k-q+L], so the
// [B,n,L,2L] BD tensor it materialises never exists here.  With rel == NULL the kernel is a
// plain scaled dot-product attention (used by the fp32-accurate BERT mode, head_dim 64).
//
// One CTA = one (batch row, head, tile of 32 queries); K, V and the needed R rows are staged in
// shared memory (row pitch DH+1 floats: conflict-free for lane-per-key access).  One warp per
// query at a time: lanes split the keys, scores stay in registers, softmax by warp shuffles.
// Masked keys (k >= seq_len) contribute exp(-4.29e9 - max) == 0 exactly in fp32, so skipping
// them is bit-equivalent for valid query rows; rows q >= seq_len are written as zeros (they never
// reach loss / pred_ids).  ~4 GFLOP per TENER forward: HBM/latency-bound, not tensor-core work.
#include "common.cuh"

namespace {

using namespace nerdev;

constexpr int QT = 32;      // queries per CTA
constexpr int NWARP = 4;
constexpr int MAXI = 16;    // keys per lane -> L <= 512

template <int DH, bool REL>
__global__ void __launch_bounds__(NWARP * 32)
attention_f32_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk,
                     const float* __restrict__ V, int ldv, const float* __restrict__ bias_u,
                     const float* __restrict__ bias_v, const float* __restrict__ rel,
                     const int32_t* __restrict__ seq_len, float scale, float* __restrict__ out_f32,
                     __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int L, int NH) {
  constexpr int P = DH + 1;
  extern __shared__ float sm[];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int len = min(max(seq_len[b], 0), L);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t row0 = (size_t)b * L;
  const int HD = NH * DH;

  float* sK = sm;                    // [len][P]
  float* sV = sK + (size_t)L * P;    // [len][P]
  float* sR = sV + (size_t)L * P;    // [len + QT][P]   rows r = rbase .. rbase + len + QT - 1
  const int rbase = L - q0 - (QT - 1);  // smallest R row any query of this tile touches (k=0, q=q0+QT-1)

  if (q0 < len) {
    for (int idx = tid; idx < len * DH; idx += NWARP * 32) {
      const int k = idx / DH, d = idx - k * DH;
      sK[k * P + d] = K[(row0 + k) * ldk + h * DH + d];
      sV[k * P + d] = V[(row0 + k) * ldv + h * DH + d];
    }
    if (REL) {
      for (int idx = tid; idx < (len + QT) * DH; idx += NWARP * 32) {
        const int r = idx / DH, d = idx - r * DH;
        const int rr = rbase + r;
        sR[r * P + d] = (rr >= 0 && rr < 2 * L) ? rel[(size_t)rr * DH + d] : 0.f;
      }
    }
  }
  __syncthreads();

  for (int qi = warp; qi < QT; qi += NWARP) {
    const int q = q0 + qi;
    if (q >= L) break;
    float* orow = out_f32 ? out_f32 + (row0 + q) * HD + h * DH : nullptr;
    if (q >= len) {  // padded query row: zeros
      for (int d = lane; d < DH; d += 32) {
        if (orow) orow[d] = 0.f;
        if (out_hi) out_hi[(row0 + q) * HD + h * DH + d] = __float2bfloat16_rn(0.f);
        if (out_lo) out_lo[(row0 + q) * HD + h * DH + d] = __float2bfloat16_rn(0.f);
      }
      continue;
    }
    float qu[DH], qv[REL ? DH : 1];
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      const float qd = Q[(row0 + q) * ldq + h * DH + d];
      qu[d] = qd + (bias_u ? bias_u[h * DH + d] : 0.f);
      if (REL) qv[d] = qd + bias_v[h * DH + d];
    }
    float s[MAXI];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int k = lane + 32 * i;
      s[i] = -INFINITY;
      if (k < len) {
        const float* kr = sK + k * P;
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) acc = fmaf(qu[d], kr[d], acc);
        if (REL) {
          const float* rr = sR + (k - q + L - rbase) * P;
          float acc2 = 0.f;
#pragma unroll
          for (int d = 0; d < DH; ++d) acc2 = fmaf(qv[d], rr[d], acc2);
          acc += acc2;
        }
        s[i] = acc * scale;
        mx = fmaxf(mx, s[i]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int k = lane + 32 * i;
      if (k < len) {
        const float p = expf(s[i] - mx);
        sum += p;
        const float* vr = sV + k * P;
#pragma unroll
        for (int d = 0; d < DH; ++d) o[d] = fmaf(p, vr[d], o[d]);
      }
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = warp_sum(o[d]);
    // lane (d % 32) writes column d
#pragma unroll
    for (int base = 0; base < DH; base += 32) {
      float val = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d)
        if (d >= base && d < base + 32 && (d - base) == lane) val = o[d];
      const int d = base + lane;
      if (d < DH) {
        const float r = val * inv;
        if (orow) orow[d] = r;
        if (out_hi) {
          const __nv_bfloat16 hi = __float2bfloat16_rn(r);
          out_hi[(row0 + q) * HD + h * DH + d] = hi;
          if (out_lo) out_lo[(row0 + q) * HD + h * DH + d] = __float2bfloat16_rn(r - __bfloat162float(hi));
        }
      }
    }
  }
}

template <int DH, bool REL>
int launch_attn(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* u,
                const float* v, const float* rel, const int32_t* seq_len, float scale, float* out_f32, void* out_hi,
                void* out_lo, int B, int L, int NH, cudaStream_t st) {
  const size_t smem = ((size_t)2 * L + (REL ? (L + QT) : 0)) * (DH + 1) * 4;
  if (smem > 227 * 1024) return NER_ERR_UNSUPPORTED;
  auto kern = attention_f32_kernel<DH, REL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  dim3 grid((L + QT - 1) / QT, NH, B);
  kern<<<grid, NWARP * 32, smem, st>>>(Q, ldq, K, ldk, V, ldv, u, v, rel, seq_len, scale, out_f32,
                                        static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo), L, NH);
  return ner_launch_status();
}

}  // namespace

extern "C" int ner_attention_f32(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                                 const float* bias_u, const float* bias_v, const float* rel_table,
                                 const int32_t* seq_len, float scale, float* out_f32, void* out_hi_bf16,
                                 void* out_lo_bf16, int B, int L, int num_heads, int head_dim, ner_stream_t stream) {
  if (B < 0 || L < 1 || num_heads < 1 || head_dim < 1) return NER_ERR_INVALID_ARG;
  if (B == 0) return NER_OK;
  if (!Q || !K || !V || !seq_len || (!out_f32 && !out_hi_bf16)) return NER_ERR_INVALID_ARG;
  if (out_lo_bf16 && !out_hi_bf16) return NER_ERR_INVALID_ARG;
  if (rel_table && !bias_v) return NER_ERR_INVALID_ARG;
  if (L > 32 * MAXI) return NER_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define GO(DHV)                                                                                                   \
  return rel_table ? launch_attn<DHV, true>(Q, ldq, K, ldk, V, ldv, bias_u, bias_v, rel_table, seq_len, scale,    \
                                            out_f32, out_hi_bf16, out_lo_bf16, B, L, num_heads, st)               \
                   : launch_attn<DHV, false>(Q, ldq, K, ldk, V, ldv, bias_u, bias_v, rel_table, seq_len, scale,   \
                                             out_f32, out_hi_bf16, out_lo_bf16, B, L, num_heads, st)
  switch (head_dim) {
    case 20: GO(20);
    case 32: GO(32);
    case 40: GO(40);
    case 64: GO(64);
    default: return NER_ERR_UNSUPPORTED;
  }
#undef GO
}
