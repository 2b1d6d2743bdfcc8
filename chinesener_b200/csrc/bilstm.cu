// BiLSTM recurrence as a persistent thread-block-cluster kernel (sm_100a).
//
// Replaces tf.nn.bidirectional_dynamic_rnn over tf.nn.rnn_cell.LSTMCell as built by
// reference tools/layer.py:10-41 (gate order i,j,f,o; forget_bias 1.0; zero initial state;
// outputs zero and state carried for t >= seq_len; the backward direction runs over
// reverse_sequence(x, seq_len) and is reversed back — SURVEY.md Appendix A.2).
//
// The input half of the LSTMCell matmul ([x_t] · kernel[:D]) + bias is hoisted out of the
// recurrence into ONE tcgen05 GEMM for both directions (xproj [B*L, 8H], gemm_tc.cu).  This
// kernel runs the sequential half: a cluster of C CTAs owns R batch rows of one direction for
// all time steps.  Each CTA keeps its slice of the recurrent matrix kernel[D:, :] resident in
// shared memory for the whole sequence (fp32, [H][4*H/C] laid out so one LDS.128 yields four
// consecutive k for one gate column), computes the 4*H/C gate pre-activations of its H/C
// hidden units, applies the cell, and broadcasts the new h slice to every CTA of the cluster
// through distributed shared memory; one cluster barrier per time step.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

// --- DSMEM signalling without a cluster barrier: a remote 4-byte store that completes transaction
// bytes on the DESTINATION CTA's mbarrier (st.async), so publishing h needs no fence over this
// thread's earlier global stores and no L1 invalidate (barrier.cluster costs both every step).
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_init_(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nerdev::smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_(uint64_t* bar, uint32_t tx_bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nerdev::smem_u32(bar)), "r"(tx_bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(nerdev::smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// ex2.approx-based forms (abs. error ~1e-7, far inside the 1e-4 parity bar of tests/test_bilstm_gpu.py):
// the activations sit on the per-step critical path of the recurrence.
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

template <int ACT>
__device__ __forceinline__ float actf(float x) {
  if (ACT == 1) return fmaxf(x, 0.f);
  return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x));   // tanh(x); saturates correctly for |x| large
}

// H4REG > 0: this thread's gate column of W_h (H4REG float4 = H fp32 values) is register-resident
// for the whole sequence (H <= 128); H4REG == 0: the slice is read from shared memory each step.
template <int R, int ACT, int H4REG>
__global__ void __launch_bounds__(H4REG > 0 ? 256 : 512, 1) bilstm_rec_kernel(const float* __restrict__ xproj, const float* __restrict__ wh_fw,
                                  const float* __restrict__ wh_bw, const int32_t* __restrict__ seq_len,
                                  float* __restrict__ out, int B, int L, int H, int C, float forget_bias,
                                  const int32_t* __restrict__ cu_seqlens, float* __restrict__ gates_out,
                                  float* __restrict__ cstate_out, float* __restrict__ hstate_out, float keep_prob,
                                  uint32_t seed_lo, uint32_t seed_hi) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int HU = H / C;       // hidden units owned by this CTA
  const int NC = 4 * HU;      // gate columns owned by this CTA
  const int H4 = H / 4;
  const int ngroups = (B + R - 1) / R;
  const int cid = blockIdx.x / C;
  const int dir = cid / ngroups;
  const int b0 = (cid % ngroups) * R;
  const int tid = threadIdx.x;

  extern __shared__ __align__(16) float smem[];
  constexpr bool WREG = H4REG > 0;
  float4* Ws4 = reinterpret_cast<float4*>(smem);                 // [H4][NC] float4 (4 consecutive k), smem path only
  float* hbuf = smem + (WREG ? 0 : (size_t)H * NC);              // [2][R][H]
  int* s_len = reinterpret_cast<int*>(hbuf + 2 * R * H);         // [R] (8 ints reserved)
  uint64_t* hbar = reinterpret_cast<uint64_t*>(s_len + 8);      // [2] one mbarrier per h buffer

  const float* wh = dir == 0 ? wh_fw : wh_bw;                    // [H][4H], columns (i,j,f,o) x H
  float4 wreg[WREG ? H4REG : 1];
  if constexpr (WREG) {
    if (tid < NC) {
      const int g0 = tid & 3, u0 = tid >> 2;   // unit-major: the 4 gates of a hidden unit sit in 4 adjacent lanes
      const size_t gc = (size_t)g0 * H + rank * HU + u0;
#pragma unroll
      for (int k4 = 0; k4 < H4REG; ++k4) {
        wreg[k4].x = wh[(size_t)(4 * k4 + 0) * 4 * H + gc];
        wreg[k4].y = wh[(size_t)(4 * k4 + 1) * 4 * H + gc];
        wreg[k4].z = wh[(size_t)(4 * k4 + 2) * 4 * H + gc];
        wreg[k4].w = wh[(size_t)(4 * k4 + 3) * 4 * H + gc];
      }
    }
  } else {
    for (int idx = tid; idx < H4 * NC; idx += blockDim.x) {
      const int k4 = idx / NC, col = idx - k4 * NC;
      const int g = col & 3, u = col >> 2;
      const size_t gc = (size_t)g * H + rank * HU + u;
      float4 w;
      w.x = wh[(size_t)(4 * k4 + 0) * 4 * H + gc];
      w.y = wh[(size_t)(4 * k4 + 1) * 4 * H + gc];
      w.z = wh[(size_t)(4 * k4 + 2) * 4 * H + gc];
      w.w = wh[(size_t)(4 * k4 + 3) * 4 * H + gc];
      Ws4[idx] = w;
    }
  }
  for (int idx = tid; idx < 2 * R * H; idx += blockDim.x) hbuf[idx] = 0.f;
  if (tid < R) s_len[tid] = (b0 + tid < B) ? min(max(seq_len[b0 + tid], 0), L) : 0;
  if (tid == 0) {
    mbar_init_(&hbar[0], 1);
    mbar_init_(&hbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) maxlen = max(maxlen, s_len[r]);
  cluster.sync();  // every CTA's hbuf is zeroed before anyone writes remotely

  // Thread t < NC owns gate g = t & 3 of hidden unit u = t >> 2 (of this CTA's slice): the four
  // pre-activations of a unit live in four adjacent lanes, so the cell update gathers them with
  // three shuffles — no shared-memory round trip and no CTA barrier between the dot products and
  // the cell; the only synchronisation per time step is the (cluster) barrier that publishes h.
  const bool col_ok = tid < NC;
  const int g = tid & 3, u = tid >> 2;
  const int ug = rank * HU + u;
  const size_t xcol = (size_t)dir * 4 * H + (size_t)g * H + ug;
  // cell role: lane g of a unit's quad updates rows r = g, g + 4, ... (RC = ceil(R / 4) rows per lane; R <= 4: one row, all
  // R cell updates in parallel)
  constexpr int RC = (R + 3) / 4;
  float c_state[RC];
#pragma unroll
  for (int rr = 0; rr < RC; ++rr) c_state[rr] = 0.f;

  // xproj row of (row r, position pos): padded layout b*L + pos, or packed layout cu_seqlens[b] + pos
  size_t xrow0[R];
#pragma unroll
  for (int r = 0; r < R; ++r)
    xrow0[r] = (b0 + r < B) ? (cu_seqlens ? (size_t)cu_seqlens[b0 + r] : (size_t)(b0 + r) * L) : 0;
  float xp[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    xp[r] = 0.f;
    const int len = s_len[r];
    if (col_ok && 0 < len) {
      const int pos = dir == 0 ? 0 : len - 1;
      xp[r] = xproj[(xrow0[r] + pos) * 8 * H + xcol];
    }
  }
  const uint32_t thr = nerdev::keep_threshold(keep_prob);
  const float inv_keep = 1.f / keep_prob;

  const uint32_t h_bytes = (uint32_t)(R * H * 4);   // every CTA receives the full h of its R rows each step
  for (int s = 0; s < maxlen; ++s) {
    const float* hcur = hbuf + (s & 1) * R * H;
    float* hnxt = hbuf + ((s + 1) & 1) * R * H;
    if (C > 1) {
      if (tid == 0) mbar_arrive_expect_tx_(&hbar[(s + 1) & 1], h_bytes);   // arm the buffer written this step
      if (s > 0) mbar_wait_(&hbar[s & 1], (uint32_t)((s - 1) >> 1) & 1u);   // h of step s-1 has landed (k-th use of the buffer)
    }
    // packed fp32 pairs (FFMA2): (w_k, w_k+1) x (h_k, h_k+1) halves the FMA issue slots of the dot
    // products, the per-step throughput bound of this kernel; two chains per row for latency
    nerdev::f32x2 pa[R], pb[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      pa[r] = nerdev::pk2(xp[r], 0.f);
      pb[r] = nerdev::pk2(0.f, 0.f);
    }
    if (col_ok) {
      // prefetch next step's input projection while the dot products run
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int len = s_len[r];
        xp[r] = 0.f;
        if (s + 1 < len) {
          const int pos = dir == 0 ? s + 1 : len - 2 - s;
          xp[r] = xproj[(xrow0[r] + pos) * 8 * H + xcol];
        }
      }
      const float4* hc4 = reinterpret_cast<const float4*>(hcur);
      if constexpr (WREG) {
#pragma unroll
        for (int k4 = 0; k4 < H4REG; ++k4) {
          const float4 w = wreg[k4];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4 hv = hc4[r * H4 + k4];
            pa[r] = nerdev::fma2(nerdev::pk2(w.x, w.y), nerdev::pk2(hv.x, hv.y), pa[r]);
            pb[r] = nerdev::fma2(nerdev::pk2(w.z, w.w), nerdev::pk2(hv.z, hv.w), pb[r]);
          }
        }
      } else {
#pragma unroll 4
        for (int k4 = 0; k4 < H4; ++k4) {
          const float4 w = Ws4[k4 * NC + tid];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4 hv = hc4[r * H4 + k4];
            pa[r] = nerdev::fma2(nerdev::pk2(w.x, w.y), nerdev::pk2(hv.x, hv.y), pa[r]);
            pb[r] = nerdev::fma2(nerdev::pk2(w.z, w.w), nerdev::pk2(hv.z, hv.w), pb[r]);
          }
        }
      }
    }
    // quad transpose: lane g of the quad receives (z_i, z_j, z_f, z_o) of its rows g, g + 4, ...
    float zi[RC], zj[RC], zf[RC], zo[RC];
#pragma unroll
    for (int rr = 0; rr < RC; ++rr) zi[rr] = zj[rr] = zf[rr] = zo[rr] = 0.f;
    const int qb = (tid & 31) & ~3;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float z0, z1, z2, z3;
      nerdev::upk2(pa[r], z0, z1);
      nerdev::upk2(pb[r], z2, z3);
      const float z = (z0 + z1) + (z2 + z3);
      const float a0 = __shfl_sync(0xffffffffu, z, qb + 0);
      const float a1 = __shfl_sync(0xffffffffu, z, qb + 1);
      const float a2 = __shfl_sync(0xffffffffu, z, qb + 2);
      const float a3 = __shfl_sync(0xffffffffu, z, qb + 3);
      if (g == (r & 3)) {
        zi[r >> 2] = a0;
        zj[r >> 2] = a1;
        zf[r >> 2] = a2;
        zo[r >> 2] = a3;
      }
    }
    // h is published with st.async + mbarrier transaction bytes (see the helpers above): with
    // barrier.cluster the release fence of the arrive stalled every step until this thread's global stores
    // had been acknowledged (ncu: 20 % of the kernel's samples on the barrier's ERRBAR) and the wait
    // invalidated L1.
#pragma unroll
    for (int rr = 0; rr < RC; ++rr) {
      const int r = g + 4 * rr;
      const bool cell_ok = col_ok && r < R;
      float i_s = 0.f, j_a = 0.f, f_s = 0.f, o_s = 0.f, h_out = 0.f, h_state = 0.f;
      const int len = cell_ok ? s_len[r] : 0;
      const int b = b0 + r;
      const bool live = cell_ok && s < len;
      const int pos = dir == 0 ? s : len - 1 - s;
      if (live) {
        i_s = sigmoidf_(zi[rr]);
        j_a = actf<ACT>(zj[rr]);
        f_s = sigmoidf_(zf[rr] + forget_bias);
        o_s = sigmoidf_(zo[rr]);
        c_state[rr] = f_s * c_state[rr] + i_s * j_a;
        const float h_raw = o_s * actf<ACT>(c_state[rr]);
        h_out = h_raw;
        h_state = h_raw;
        if (keep_prob < 1.f) {
          // DropoutWrapper(output_keep_prob, state_keep_prob): independent masks for the emitted output
          // and for the h part of the carried state (c is not dropped), fresh per step
          const uint32_t e = (uint32_t)(((size_t)b * L + pos) * 2 * H + (size_t)dir * H + ug);
          h_out = nerdev::hash3(seed_lo, seed_hi, e) < thr ? h_raw * inv_keep : 0.f;
          h_state = nerdev::hash3(seed_lo ^ 0x5bd1e995u, seed_hi, e) < thr ? h_raw * inv_keep : 0.f;
        }
      }
      if (cell_ok) {
        // (h of a finished row is never read again — its own recurrence has stopped — so 0 is as good as
        // the carried value dynamic_rnn keeps)
        if (C > 1) {
          const uint32_t la = nerdev::smem_u32(hnxt + r * H + ug), lb = nerdev::smem_u32(&hbar[(s + 1) & 1]);
          for (int dst = 0; dst < C; ++dst) st_async_f32(mapa_u32(la, (uint32_t)dst), h_state, mapa_u32(lb, (uint32_t)dst));
        } else {
          hnxt[r * H + ug] = h_state;
        }
      }
      if (live) {
        out[((size_t)b * L + pos) * 2 * H + (size_t)dir * H + ug] = h_out;
        if (hstate_out != nullptr) hstate_out[((size_t)b * L + pos) * 2 * H + (size_t)dir * H + ug] = h_state;
        if (gates_out != nullptr) {  // saved for back-propagation through time (bilstm_bwd.cu)
          const size_t gi = ((size_t)b * L + pos) * 8 * H + (size_t)dir * 4 * H;
          gates_out[gi + 0 * H + ug] = i_s;
          gates_out[gi + 1 * H + ug] = j_a;
          gates_out[gi + 2 * H + ug] = f_s;
          gates_out[gi + 3 * H + ug] = o_s;
          cstate_out[((size_t)b * L + pos) * 2 * H + (size_t)dir * H + ug] = c_state[rr];
        }
      } else if (cell_ok && b < B) {
        out[((size_t)b * L + s) * 2 * H + (size_t)dir * H + ug] = 0.f;   // past this row's end: dynamic_rnn emits zeros
      }
    }
    if (C == 1) __syncthreads();
  }
  if (C > 1) cluster.sync();   // nobody exits while a peer may still be sending into its shared memory

  // positions past the longest row of this cluster: zeros
  for (int idx = tid; idx < R * HU; idx += blockDim.x) {
    const int r = idx / HU, uu = idx - r * HU;
    const int b = b0 + r;
    if (b < B)
      for (int s = maxlen; s < L; ++s) out[((size_t)b * L + s) * 2 * H + (size_t)dir * H + rank * HU + uu] = 0.f;
  }
}

int pick_cluster(int H) {
  // smallest power-of-two cluster whose W_h slice (H * 4H/C floats) fits ~190 KB and divides H
  for (int C = 1; C <= 8; C *= 2) {
    if (H % C != 0) continue;
    const size_t bytes = (size_t)H * 4 * (H / C) * 4;
    if (bytes <= 190 * 1024 && 4 * (H / C) <= 512) return C;
  }
  return 0;
}

template <int R, int ACT, int H4REG>
int launch_rec(const float* xproj, const float* wh_fw, const float* wh_bw, const int32_t* seq_len, float* out, int B,
               int L, int H, int C, float forget_bias, const int32_t* cu_seqlens, float* gates_out, float* cstate_out,
               float* hstate_out, float keep_prob, uint64_t seed, cudaStream_t st) {
  const int HU = H / C, NC = 4 * HU;
  const size_t smem = ((H4REG > 0 ? 0 : (size_t)H * NC) + 2 * R * H + 32) * 4;   // + s_len[8] + 2 mbarriers (16-B aligned: R*H even)
  auto kern = bilstm_rec_kernel<R, ACT, H4REG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  const int ngroups = (B + R - 1) / R;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * ngroups * C));
  cfg.blockDim = dim3((unsigned)((NC + 31) / 32 * 32));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, xproj, wh_fw, wh_bw, seq_len, out, B, L, H, C, forget_bias, cu_seqlens, gates_out,
                         cstate_out, hstate_out, keep_prob, (uint32_t)seed, (uint32_t)(seed >> 32));
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  return ner_launch_status();
}

}  // namespace

extern "C" int ner_bilstm_recurrence(const float* xproj, const float* wh_fw, const float* wh_bw,
                                     const int32_t* seq_len, float* out, int B, int L, int H, int activation,
                                     float forget_bias, const int32_t* cu_seqlens, float* gates_out,
                                     float* cstate_out, float* hstate_out, float keep_prob, uint64_t seed,
                                     ner_stream_t stream) {
  if (B < 0 || L < 1 || H < 1) return NER_ERR_INVALID_ARG;
  if (B == 0) return NER_OK;
  if (!xproj || !wh_fw || !wh_bw || !seq_len || !out) return NER_ERR_INVALID_ARG;
  if ((gates_out == nullptr) != (cstate_out == nullptr)) return NER_ERR_INVALID_ARG;
  if (!(keep_prob > 0.f) || keep_prob > 1.f) return NER_ERR_INVALID_ARG;
  if (activation != 0 && activation != 1) return NER_ERR_INVALID_ARG;
  if (H % 4 != 0) return NER_ERR_UNSUPPORTED;
  const int C = pick_cluster(H);
  if (C == 0) return NER_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // rows per cluster: fill the 148 SMs once when the batch is small, amortise W_h reads when large
  int R = 1;
  if ((long)2 * B * C > 148) R = 2;
  if ((long)2 * ((B + 1) / 2) * C > 2 * 148) R = 4;
  // four stacked PREDICT batches (B = 256): 4 rows per cluster would be 256 CTAs = two waves of the one-CTA-per-SM kernel
  if (H == 128 && (long)2 * ((B + 3) / 4) * C > 148) R = 8;
  if (const char* e = getenv("NER_BILSTM_ROWS")) {   // tuning hook: rows per cluster (1, 2, 4 or 8)
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || (v == 8 && H == 128)) R = v;
  }
#define GO(RR, HR)                                                                                          \
  return activation == 1 ? launch_rec<RR, 1, HR>(xproj, wh_fw, wh_bw, seq_len, out, B, L, H, C, forget_bias, cu_seqlens, gates_out, cstate_out, hstate_out, keep_prob, seed, st) \
                         : launch_rec<RR, 0, HR>(xproj, wh_fw, wh_bw, seq_len, out, B, L, H, C, forget_bias, cu_seqlens, gates_out, cstate_out, hstate_out, keep_prob, seed, st)
  if (H == 128 && 4 * (H / C) <= 256) {  // register-resident W_h (the bert_bilstm_crf / bilstm_crf shape)
    if (R == 8) GO(8, 32);
    if (R == 4) GO(4, 32);
    if (R == 2) GO(2, 32);
    GO(1, 32);
  }
  if (R >= 4) GO(4, 0);
  if (R == 2) GO(2, 0);
  GO(1, 0);
#undef GO
}
