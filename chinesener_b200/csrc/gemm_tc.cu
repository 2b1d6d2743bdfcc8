// Dense layers of the encoder on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   out[M,N] = epilogue( A[M,K] · Wt[N,K]^T + bias[N] )       A, Wt bf16 (K-major), fp32 accumulate
//
// Replaces the tf.layers.dense / BertModel dense_layer matmuls executed inside
// reference tools/layer.py:68-77 (bert_base.bert.modeling) and model/bert_bilstm_crf.py:26,
// and the input projection half of the LSTMCell matmul (tools/layer.py:16,35).
//
// Kernel shape (persistent, warp-specialised, 320 threads = 10 warps, 1 CTA/SM):
//   warp 0    TMA producer : cp.async.bulk.tensor 2-D loads of a 128x64 A box and a B box per
//                            k-block into a STAGES-deep smem ring (SWIZZLE_128B), mbarrier tx
//   warp 1    MMA issuer   : one thread issues tcgen05.mma.kind::f16 x4 per k-block;
//                            tcgen05.commit releases smem slots / publishes the accumulator
//   warps 2-9 epilogue     : tcgen05.ld 32x32b (TMEM lane quarter = warp%4, two warps per quarter
//                            on alternating 128-byte column groups) -> bias / GELU / residual ->
//                            swizzled smem staging tile -> cp.async.bulk.tensor (TMA) store
// Two TMEM accumulator stages (2*BN columns) let the epilogue of tile i overlap the main loop
// of tile i+1.
//
// Two variants:
//   gemm_bf16_tc_kernel<BN>    cta_group::1, 128 x BN tile per CTA (BN = 64/128/256)
//   gemm_bf16_tc2_kernel<BN>   cta_group::2: a CTA PAIR (cluster of 2, one per SM of a TPC)
//                              computes a 256 x BN tile with M=256 MMAs issued by CTA 0; each CTA
//                              stages its own 128 A rows and HALF of the B tile, so the L2->smem
//                              bytes per FLOP drop by 1.5x vs the 128x256 single-CTA tile (the
//                              single-CTA kernel is L2-feed-bound at ~9-10 TB/s).
#include <stdlib.h>

#include <mutex>

#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
// stream-K fixed charge in k-block times: parking + re-reading a 128 KB partial accumulator and the
// exposed finisher epilogue cost ~6 us (trip 16/17: SK loses on the encoder's forward shapes at M ~ 3150,
// 24.9 vs 18.8 us for 3150x2304x768); it pays when tiles << SMs and K is long (weight gradients, K = tokens).
constexpr float kSkFixupCost = 10.0f;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;

constexpr int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192) ? 4 : ((BN == 128) ? 6 : 8);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int TMEM_COLS = tmem_cols_for(2 * BN);
  // ring + 8 x 4 KB epilogue staging + barriers + bias; the dynamic smem base is 1024-aligned (checked in-kernel)
  static constexpr size_t SMEM = (size_t)STAGES * (A_BYTES + B_BYTES) + NUM_EPI_WARPS * 4096 + 256 + 2 * BN * 4;
};

template <int BN>
struct Cfg2 {  // per CTA of the pair
  static constexpr int STAGES = 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int TMEM_COLS = tmem_cols_for(2 * BN);
  static constexpr size_t SMEM = (size_t)STAGES * (A_BYTES + B_BYTES) + NUM_EPI_WARPS * 4096 + 256 + 2 * BN * 4;
};

struct EpiArgs {
  const float* bias;      // [N] or null
  const float* residual;  // [M,N] fp32 or null (EPI_RES_F32)
  void* out;              // [M,N] bf16 or fp32
  int mode;
};

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5x(1+tanh(sqrt(2/pi)(x+0.044715x^3)))  (google-research/bert modeling.gelu).
  // tanh.approx.f32 is one MUFU op (rel. error ~2^-11); the result is rounded to bf16 (2^-9).
  const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(u), hx);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.7071067811865476f)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Fused epilogue math on one 32-column chunk of one accumulator row (values in registers).
// `sbias` = this chunk's 32 bias values in shared memory (broadcast LDS.128), staged once per tile.
__device__ __forceinline__ void epilogue_math(float (&v)[32], const uint32_t (&r)[32], const EpiArgs& ep,
                                              const float* sbias, int row, int col0, int M, int N) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  {
    const float4* b4 = reinterpret_cast<const float4*>(sbias);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = b4[i];
      v[4 * i + 0] += b.x;
      v[4 * i + 1] += b.y;
      v[4 * i + 2] += b.z;
      v[4 * i + 3] += b.w;
    }
  }
  if (ep.mode == NER_EPI_RES_F32 || ep.mode == NER_EPI_RES_RELU_F32) {
    if (row < M && col0 < N) {
      const float4* r4 = reinterpret_cast<const float4*>(ep.residual + (size_t)row * N + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = __ldg(r4 + i);
        v[4 * i + 0] += b.x;
        v[4 * i + 1] += b.y;
        v[4 * i + 2] += b.z;
        v[4 * i + 3] += b.w;
      }
    }
    if (ep.mode == NER_EPI_RES_RELU_F32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
  } else if (ep.mode == NER_EPI_GELU_TANH_BF16) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
  } else if (ep.mode == NER_EPI_GELU_ERF_BF16) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  } else if (ep.mode == NER_EPI_RELU_BF16) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// Stream-K work split.  With SK the M x N x K iteration space is cut into tiles * num_kb k-block
// units and CTA c owns the contiguous unit range [c*U/G, (c+1)*U/G): every SM gets the same number
// of k-blocks whatever tiles/SMs is (the packed M of an MSRA batch gives 1.5 / 2.03 / 0.5 waves on
// the encoder's GEMMs, i.e. 25-50 % idle SMs with whole-tile scheduling).  A CTA's range is
//   [tail piece of a tile]  [whole tiles]  [head piece of a tile]
// The CTA holding a tile's HEAD (kb0 == 0) finishes that tile: it adds the fp32 partial sums that
// the CTAs holding the later k-blocks parked in their workspace slot and runs the normal epilogue.
// A tail piece is always the FIRST thing its CTA does and the head piece the LAST, so a finisher
// never waits in practice and no wait cycle can form (all CTAs are co-resident: grid <= #SMs).
struct SkArgs {
  float4* ws;   // [grid][BN/32 chunks][8][128 rows] float4: slot c = partial accumulator of CTA c's first segment
  int* flags;   // [grid] 0 / 1 = slot c published; reset to 0 by the finisher (all zero between launches)
};

struct SegIter {
  int num_kb, num_tiles, tile, stride;
  long long u, u1;
  bool sk;
  __device__ SegIter(bool sk_, int num_tiles_, int num_kb_) : num_kb(num_kb_), num_tiles(num_tiles_), sk(sk_) {
    if (sk) {
      const long long U = (long long)num_tiles * num_kb;
      u = (long long)blockIdx.x * U / gridDim.x;
      u1 = (long long)(blockIdx.x + 1) * U / gridDim.x;
    } else {
      tile = blockIdx.x;
      stride = gridDim.x;
    }
  }
  __device__ bool next(int& t, int& kb0, int& kb1) {
    if (sk) {
      if (u >= u1) return false;
      t = (int)(u / num_kb);
      kb0 = (int)(u - (long long)t * num_kb);
      const long long rem = u1 - u;
      kb1 = (rem < (long long)(num_kb - kb0)) ? kb0 + (int)rem : num_kb;
      u += kb1 - kb0;
      return true;
    }
    if (tile >= num_tiles) return false;
    t = tile;
    kb0 = 0;
    kb1 = num_kb;
    tile += stride;
    return true;
  }
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void epi_bar_sync() {  // named barrier 1 over the 8 epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
}

// Stage bias[n0 .. n0+BN) of the coming tile into smem (called by all epilogue threads BEFORE they
// wait for the accumulator, so the global-load latency hides behind the main loop).
template <int BN>
__device__ __forceinline__ void epilogue_stage_bias(float* sbias, const EpiArgs& ep, int n0, int N) {
  const int t = (int)threadIdx.x - 64;  // 0 .. 255
  for (int i = t; i < BN; i += 32 * NUM_EPI_WARPS) {
    const int col = n0 + i;
    sbias[i] = (ep.bias != nullptr && col < N) ? __ldg(ep.bias + col) : 0.f;
  }
  epi_bar_sync();
}

// Epilogue of one 128 x BN accumulator.  Each epilogue warp owns a TMEM lane quarter (32 rows) and
// every other 128-byte column group (64 bf16 or 32 fp32 columns).  Per group: tcgen05.ld -> fused
// math -> the warp's 32 x 128 B staging tile in shared memory (SWIZZLE_128B: chunk ^ (row & 7),
// conflict-free STS.128) -> ONE cp.async.bulk.tensor store by lane 0.  The store writes full
// 128-byte lines and clips rows >= M / columns >= N by itself; direct per-thread stores wrote
// 16-byte fragments of 32 different lines per instruction and were L2-transaction-bound.
// Finisher side of a split tile: add the parked partial sums of CTAs [c_first, c_first + n_part) to
// one 32-column chunk held in registers (same (chunk, i, row) layout the contributors wrote).
template <int BN>
__device__ __forceinline__ void add_partials(uint32_t (&r)[32], const float4* ws, int c_first, int n_part, int chunk,
                                             int row_in_tile) {
  constexpr size_t SLOT = (size_t)(BN / 32) * 8 * 128;
  for (int c = 0; c < n_part; ++c) {
    const float4* p = ws + (size_t)(c_first + c) * SLOT + (size_t)chunk * 8 * 128 + row_in_tile;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = __ldcg(p + i * 128);
      r[4 * i + 0] = __float_as_uint(__uint_as_float(r[4 * i + 0]) + v.x);
      r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + v.y);
      r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + v.z);
      r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + v.w);
    }
  }
}

// Contributor side: park the raw fp32 accumulator of this CTA's (partial) first segment in its slot.
template <int BN>
__device__ __forceinline__ void park_partial(uint32_t tmem_acc, int warp, int lane, float4* slot) {
  const int q = warp & 3, half = (warp - 2) >> 2;
  const uint32_t tbase = tmem_acc + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
  for (int ch = half; ch < BN / 32; ch += 2) {
    uint32_t ra[32];
    tmem_ld_32x32(tbase + (uint32_t)(ch * 32), ra);
    tmem_ld_wait();
    float4* p = slot + (size_t)ch * 8 * 128 + q * 32 + lane;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      __stcg(p + i * 128, make_float4(__uint_as_float(ra[4 * i]), __uint_as_float(ra[4 * i + 1]),
                                      __uint_as_float(ra[4 * i + 2]), __uint_as_float(ra[4 * i + 3])));
  }
}

template <int BN, bool OUT_F32>
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_acc, int warp, int lane, const EpiArgs& ep,
                                              const CUtensorMap* tma_c, uint8_t* stage, const float* sbias, int row0,
                                              int n0, int M, int N, const float4* ws = nullptr, int c_first = 0,
                                              int n_part = 0) {
  constexpr int GC = OUT_F32 ? 32 : 64;  // columns per 128-byte group
  constexpr int NG = BN / GC;            // groups per tile
  const int q = warp & 3;                // TMEM lane quarter this warp may access
  const int half = (warp - 2) >> 2;      // 0: even groups, 1: odd groups
  const int row = row0 + q * 32 + lane;
  const uint32_t tbase = tmem_acc + ((uint32_t)(q * 32) << 16);
  uint8_t* my = stage + lane * 128;
  const int sw = lane & 7;
#pragma unroll 1
  for (int g = half; g < NG; g += 2) {
    const int col0 = n0 + g * GC;
    uint32_t ra[32];
    float v[32];
    uint4 pk[8];
    tmem_ld_32x32(tbase + (uint32_t)(g * GC), ra);
    tmem_ld_wait();
    if (n_part > 0) add_partials<BN>(ra, ws, c_first, n_part, g * GC / 32, q * 32 + lane);
    epilogue_math(v, ra, ep, sbias + g * GC, row, col0, M, N);
    if constexpr (OUT_F32) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        pk[i] = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]),
                           __float_as_uint(v[4 * i + 3]));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        pk[i] = make_uint4(pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
      tmem_ld_32x32(tbase + (uint32_t)(g * GC + 32), ra);
      tmem_ld_wait();
      if (n_part > 0) add_partials<BN>(ra, ws, c_first, n_part, g * GC / 32 + 1, q * 32 + lane);
      epilogue_math(v, ra, ep, sbias + g * GC + 32, row, col0 + 32, M, N);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        pk[4 + i] = make_uint4(pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                               pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
    }
    if (ep.mode == NER_EPI_DIAG_DISCARD) continue;  // diagnostic: drain TMEM, store nothing
    // the previous bulk store of this warp must have finished reading the staging tile
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(my + ((i ^ sw) << 4)) = pk[i];
    fence_proxy_async();
    __syncwarp();
    if (lane == 0 && col0 < N && row0 + q * 32 < M) {
      tma_store_2d(tma_c, stage, col0, row0 + q * 32);
      tma_store_commit();
    }
  }
}

// ===================================================================== cta_group::1
template <int BN, bool SK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, EpiArgs ep, int M, int N, int K, SkArgs skargs) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  nerdev::pdl_launch_dependents();   // the next kernel of the stream may start its prologue while this one runs

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;  // SWIZZLE_128B tiles need 1024-B alignment
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * C::A_BYTES;
  uint8_t* epi_stage = smem + STAGES * (C::A_BYTES + C::B_BYTES);  // [NUM_EPI_WARPS][32 rows][128 B], 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + NUM_EPI_WARPS * 4096);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);  // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], NUM_EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // barriers, TMEM and descriptor prefetch are set up: everything below touches global memory and must
  // wait for the previous kernel of the stream (no-op unless launched with the PDL attribute)
  nerdev::pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it(SK, num_tiles, num_kb);
      int tile, kb0, kb1;
      while (it.next(tile, kb0, kb1)) {
        const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], C::A_BYTES + C::B_BYTES);
          tma_load_2d(smem_a + stage * C::A_BYTES, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(smem_b + stage * C::B_BYTES, &tma_b, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      SegIter it(SK, num_tiles, num_kb);
      int tile, kb0, kb1;
      while (it.next(tile, kb0, kb1)) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * UMMA_K * 2);
            umma_f16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr size_t SLOT = (size_t)(BN / 32) * 8 * 128;
    SegIter it(SK, num_tiles, num_kb);
    int tile, kb0, kb1;
    while (it.next(tile, kb0, kb1)) {
      const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
      const bool contributor = SK && kb0 > 0;
      if (!contributor) epilogue_stage_bias<BN>(sbias + acc * BN, ep, n_blk * BN, N);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      if (contributor) {
        park_partial<BN>(tmem_base + (uint32_t)(acc * BN), warp, lane, skargs.ws + (size_t)blockIdx.x * SLOT);
        __threadfence();
        epi_bar_sync();
        if (warp == 2 && lane == 0) st_release_gpu(skargs.flags + blockIdx.x, 1);
      } else {
        int c_first = 0, n_part = 0;
        if (SK && kb1 < num_kb) {
          // finisher: the following CTAs whose ranges start inside this tile hold its later k-blocks
          const long long U = (long long)num_tiles * num_kb, tile_end = (long long)(tile + 1) * num_kb;
          c_first = blockIdx.x + 1;
          for (int c = c_first; c < (int)gridDim.x; ++c) {
            const long long c0 = (long long)c * U / gridDim.x, c1 = (long long)(c + 1) * U / gridDim.x;
            if (c0 >= tile_end) break;
            if (c1 > c0) ++n_part; else if (n_part == 0) ++c_first;   // (empty ranges only occur when U < grid)
          }
          if (warp == 2 && lane == 0) {
            for (int c = 0; c < n_part; ++c) {
              unsigned spins = 0;
              while (ld_acquire_gpu(skargs.flags + c_first + c) == 0)
                if (++spins > (1u << 28)) __trap();   // never hang the GPU on a protocol bug
            }
          }
          epi_bar_sync();
          __threadfence();
        }
        if (ep.mode == NER_EPI_F32 || ep.mode == NER_EPI_RES_F32 || ep.mode == NER_EPI_RES_RELU_F32)
          epilogue_tile<BN, true>(tmem_base + (uint32_t)(acc * BN), warp, lane, ep, &tma_c, epi_stage + (warp - 2) * 4096,
                                  sbias + acc * BN, m_blk * BM, n_blk * BN, M, N, skargs.ws, c_first, n_part);
        else
          epilogue_tile<BN, false>(tmem_base + (uint32_t)(acc * BN), warp, lane, ep, &tma_c, epi_stage + (warp - 2) * 4096,
                                   sbias + acc * BN, m_blk * BM, n_blk * BN, M, N, skargs.ws, c_first, n_part);
        if (n_part > 0) {
          epi_bar_sync();   // every epilogue thread has consumed the parked partials
          if (warp == 2)
            for (int c = lane; c < n_part; c += 32) skargs.flags[c_first + c] = 0;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  if (warp >= 2 && lane == 0) tma_store_wait_read<0>();  // staging tiles fully read before the CTA exits
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ===================================================================== grouped weight gradients
// dW_p[k_in, n_out] += X_p^T[k_in, R] . dY_p[R, n_out]   for a group of problems p (the weight gradients of one encoder layer),
// ONE persistent launch: the tiles of all problems form one list (128 x 256 output tiles), so the 18..72 tiles of the single
// GEMMs (K = tokens: 50 k-blocks, a quarter to a half of the SMs idle per launch) become 216 per layer.
// Both operands are read AS THEY LIE in memory — X [R, k_in] and dY [R, n_out] are token-major, i.e. M / N contiguous and
// K (tokens) strided: "MN-major" operands of tcgen05.mma (instruction-descriptor bits 15/16).  A k-block is 64 tokens; the A
// tile arrives as two and the B tile as four {64 columns x 64 tokens} TMA boxes (SWIZZLE_128B), which is the canonical MN-major
// layout: 64-element blocks along M/N 8 KB apart (leading byte offset), 8-token groups 1 KB apart (stride byte offset); one
// MMA (K = 16) spans two token groups, consecutive MMAs advance the start address by 2 KB.  No bf16 transposes, no padded
// copies: TMA zero-fills the token rows past R.  Epilogue = the GEMM's fp32 accumulate-into-gradient path (RES_F32).
constexpr int WG_MAXP = 6;
constexpr int WG_BN = 256;

struct WgradGroup {
  CUtensorMap a[WG_MAXP], b[WG_MAXP], c[WG_MAXP];
  float* dw[WG_MAXP];
  int m[WG_MAXP], n[WG_MAXP], b_col0[WG_MAXP];
  int tile_start[WG_MAXP + 1];
  int count, rows;
};

__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // next 64-element block along M / N
  d |= (uint64_t)(1024 >> 4) << 32;                   // next group of 8 tokens along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_wgrad_group_kernel(const __grid_constant__ WgradGroup grp) {
  constexpr int BN = WG_BN;
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int BOX = 64 * 64 * 2;   // one {64 columns x 64 tokens} box
  nerdev::pdl_launch_dependents();

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * C::A_BYTES;
  uint8_t* epi_stage = smem + STAGES * (C::A_BYTES + C::B_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + NUM_EPI_WARPS * 4096);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = grp.tile_start[grp.count];
  const int num_kb = (grp.rows + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  nerdev::pdl_wait();

  auto locate = [&](int tile, int& p, int& m_blk, int& n_blk) {
    p = 0;
    while (p + 1 < grp.count && tile >= grp.tile_start[p + 1]) ++p;
    const int t = tile - grp.tile_start[p], num_n = grp.n[p] / BN;
    m_blk = t / num_n;
    n_blk = t - m_blk * num_n;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int p, m_blk, n_blk;
        locate(tile, p, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], C::A_BYTES + C::B_BYTES);
#pragma unroll
          for (int i = 0; i < BM / 64; ++i)
            tma_load_2d(smem_a + stage * C::A_BYTES + i * BOX, &grp.a[p], &full_bar[stage], m_blk * BM + i * 64, kb * BK);
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
            tma_load_2d(smem_b + stage * C::B_BYTES + i * BOX, &grp.b[p], &full_bar[stage], grp.b_col0[p] + n_blk * BN + i * 64,
                        kb * BK);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_mn(BM, BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16(d_tmem, make_smem_desc_sw128_mn(a_addr + k * 2048, BOX), make_smem_desc_sw128_mn(b_addr + k * 2048, BOX), idesc,
                     (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int p, m_blk, n_blk;
      locate(tile, p, m_blk, n_blk);
      EpiArgs ep{nullptr, grp.dw[p], grp.dw[p], NER_EPI_RES_F32};
      epilogue_stage_bias<BN>(sbias + acc * BN, ep, n_blk * BN, grp.n[p]);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      epilogue_tile<BN, true>(tmem_base + (uint32_t)(acc * BN), warp, lane, ep, &grp.c[p], epi_stage + (warp - 2) * 4096,
                              sbias + acc * BN, m_blk * BM, n_blk * BN, grp.m[p], grp.n[p]);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  if (warp >= 2 && lane == 0) tma_store_wait_read<0>();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ===================================================================== cta_group::2 (CTA pair)
// Protocol (s = smem stage, a = accumulator stage):
//   full[s]        lives in CTA 0, count 1: CTA 0's producer arrive.expect_tx's the bytes of BOTH CTAs;
//                  both CTAs' TMA loads complete_tx on it (peer-bit-masked barrier address).
//   empty[s]       one per CTA, count 1: the MMA thread's multicast commit arrives on both.
//   tmem_full[a]   one per CTA, count 1: multicast commit after the last k-block.
//   tmem_empty[a]  lives in CTA 0, count 2*NUM_EPI_WARPS: every epilogue warp of both CTAs arrives
//                  (CTA 1 through its shared::cluster address).
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_c, EpiArgs ep, int M, int N, int K) {
  using C = Cfg2<BN>;
  constexpr int STAGES = C::STAGES;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * C::A_BYTES;
  uint8_t* epi_stage = smem + STAGES * (C::A_BYTES + C::B_BYTES);  // [NUM_EPI_WARPS][32 rows][128 B], 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + NUM_EPI_WARPS * 4096);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);  // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int num_m = (M + 2 * BM - 1) / (2 * BM);  // 256-row tiles
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  cluster_sync_all();  // barrier inits visible cluster-wide before any remote arrive / TMA complete_tx
  if (warp == 1) tmem_alloc_2sm<C::TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
        const int row_a = m_blk * 2 * BM + (int)rank * BM;
        const int row_b = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (C::A_BYTES + C::B_BYTES));
          tma_load_2d_2sm(smem_a + stage * C::A_BYTES, &tma_a, &full_bar[stage], kb * BK, row_a);
          tma_load_2d_2sm(smem_b + stage * C::B_BYTES, &tma_b, &full_bar[stage], kb * BK, row_b);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of CTA 0) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * UMMA_K * 2);
            umma_f16_2sm(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage], 0b11);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_2sm(&tmem_full[acc], 0b11);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9), both CTAs =====================
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t empty_remote0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t empty_remote1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
      epilogue_stage_bias<BN>(sbias + acc * BN, ep, n_blk * BN, N);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      if (ep.mode == NER_EPI_F32 || ep.mode == NER_EPI_RES_F32 || ep.mode == NER_EPI_RES_RELU_F32)
        epilogue_tile<BN, true>(tmem_base + (uint32_t)(acc * BN), warp, lane, ep, &tma_c, epi_stage + (warp - 2) * 4096,
                                sbias + acc * BN, m_blk * 2 * BM + (int)rank * BM, n_blk * BN, M, N);
      else
        epilogue_tile<BN, false>(tmem_base + (uint32_t)(acc * BN), warp, lane, ep, &tma_c, epi_stage + (warp - 2) * 4096,
                                 sbias + acc * BN, m_blk * 2 * BM + (int)rank * BM, n_blk * BN, M, N);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc == 0 ? empty_remote0 : empty_remote1);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  if (warp >= 2 && lane == 0) tma_store_wait_read<0>();
  tcgen05_fence_before();
  cluster_sync_all();  // both CTAs done with TMEM / remote barriers
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm<C::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  }
  return fn;
}

// bf16 row-major [rows, cols] with a {64, box_rows} box, 128-byte swizzle.
int make_map_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return NER_ERR_NO_DRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NER_OK : NER_ERR_INVALID_ARG;
}

// output [M, N] (bf16 or f32) with a {128 bytes of columns, 32 rows} box, 128-byte swizzle
int make_map_out(CUtensorMap* map, void* ptr, uint64_t rows, uint64_t cols, bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return NER_ERR_NO_DRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NER_OK : NER_ERR_INVALID_ARG;
}

// NER_GEMM_POLICY=1: "auto" picks the 128x256 tile whenever N allows instead of fitting waves (tuning hook).
int gemm_auto_policy() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NER_GEMM_POLICY");
    v = e ? atoi(e) : 0;
  }
  return v;
}

bool epi_is_f32(int mode) { return mode == NER_EPI_F32 || mode == NER_EPI_RES_F32 || mode == NER_EPI_RES_RELU_F32; }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Stream-K scratch: one slot of 128 x 256 fp32 per CTA + one flag per CTA, per (device, stream) so
// that GEMMs running concurrently on different streams never share slots.  Allocated on first use
// (the only memory this library owns), never freed, flags zeroed once (the kernel restores them).
struct SkScratch {
  int dev;
  cudaStream_t st;
  float4* ws;
  int* flags;
};
constexpr int kMaxSkScratch = 16;
SkScratch g_sk[kMaxSkScratch];
int g_sk_n = 0;
std::mutex g_sk_mu;

bool sk_scratch(cudaStream_t st, SkArgs* out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  std::lock_guard<std::mutex> lk(g_sk_mu);
  for (int i = 0; i < g_sk_n; ++i)
    if (g_sk[i].dev == dev && g_sk[i].st == st) {
      out->ws = g_sk[i].ws;
      out->flags = g_sk[i].flags;
      return true;
    }
  if (g_sk_n == kMaxSkScratch) return false;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    (void)cudaGetLastError();
    return false;   // no allocation inside a graph capture: the caller falls back to whole-tile scheduling
  }
  const size_t slot = (size_t)(256 / 32) * 8 * 128 * sizeof(float4);
  float4* ws = nullptr;
  int* flags = nullptr;
  if (cudaMalloc(&ws, slot * sm_count()) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  if (cudaMalloc(&flags, sizeof(int) * sm_count()) != cudaSuccess || cudaMemset(flags, 0, sizeof(int) * sm_count()) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(ws);
    return false;
  }
  cudaDeviceSynchronize();
  g_sk[g_sk_n++] = SkScratch{dev, st, ws, flags};
  out->ws = ws;
  out->flags = flags;
  return true;
}

template <int BN>
int launch_gemm(const void* A, const void* Wt, EpiArgs ep, int M, int N, int K, cudaStream_t st, bool sk = false) {
  CUtensorMap ma, mb;
  int rc = make_map_bf16_2d(&ma, A, (uint64_t)M, (uint64_t)K, BM);
  if (rc != NER_OK) return rc;
  rc = make_map_bf16_2d(&mb, Wt, (uint64_t)N, (uint64_t)K, BN);
  if (rc != NER_OK) return rc;
  CUtensorMap mc;
  rc = make_map_out(&mc, ep.out, (uint64_t)M, (uint64_t)N, epi_is_f32(ep.mode));
  if (rc != NER_OK) return rc;
  const size_t smem = Cfg<BN>::SMEM;
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int num_kb = (K + BK - 1) / BK;
  SkArgs ska{nullptr, nullptr};
  // stream-K needs every CTA co-resident (grid = #SMs) and at least one k-block unit per CTA
  sk = sk && (long long)tiles * num_kb >= sm_count() && ep.mode != NER_EPI_DIAG_DISCARD && sk_scratch(st, &ska);
  if (sk) {
    auto kern = gemm_bf16_tc_kernel<BN, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
    e = ner_launch_pdl(kern, dim3(sm_count()), dim3(NUM_THREADS), smem, st, ma, mb, mc, ep, M, N, K, ska);
    if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
    return ner_launch_status();
  }
  auto kern = gemm_bf16_tc_kernel<BN, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  e = ner_launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), smem, st, ma, mb, mc, ep, M, N, K, ska);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  return ner_launch_status();
}

template <int BN>
int launch_gemm2(const void* A, const void* Wt, EpiArgs ep, int M, int N, int K, cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc = make_map_bf16_2d(&ma, A, (uint64_t)M, (uint64_t)K, BM);
  if (rc != NER_OK) return rc;
  rc = make_map_bf16_2d(&mb, Wt, (uint64_t)N, (uint64_t)K, BN / 2);
  if (rc != NER_OK) return rc;
  CUtensorMap mc;
  rc = make_map_out(&mc, ep.out, (uint64_t)M, (uint64_t)N, epi_is_f32(ep.mode));
  if (rc != NER_OK) return rc;
  auto kern = gemm_bf16_tc2_kernel<BN>;
  const size_t smem = Cfg2<BN>::SMEM;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const int max_pairs = sm_count() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  kern<<<2 * pairs, NUM_THREADS, smem, st>>>(ma, mb, mc, ep, M, N, K);
  return ner_launch_status();
}

}  // namespace

// bf16 row-major [rows, cols] with a {64 columns, 64 rows} box, 128-byte swizzle (MN-major operand tiles).
static int make_map_bf16_box64(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return NER_ERR_NO_DRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NER_OK : NER_ERR_INVALID_ARG;
}

extern "C" int ner_wgrad_group_bf16(const ner_wgrad_problem* problems_host, int count, int rows, ner_stream_t stream) {
  if (count < 0 || rows < 0 || (count > 0 && !problems_host)) return NER_ERR_INVALID_ARG;
  if (count == 0 || rows == 0) return NER_OK;
  if (count > WG_MAXP) return NER_ERR_UNSUPPORTED;
  WgradGroup g;
  g.count = count;
  g.rows = rows;
  g.tile_start[0] = 0;
  for (int p = 0; p < count; ++p) {
    const ner_wgrad_problem& q = problems_host[p];
    if (!q.x_bf16 || !q.dy_bf16 || !q.dw || q.k_in < 1 || q.n_out < 1 || q.ld_x < q.k_in || q.dy_col0 < 0 ||
        q.ld_dy < q.dy_col0 + q.n_out)
      return NER_ERR_INVALID_ARG;
    if (q.k_in % BM != 0 || q.n_out % WG_BN != 0 || q.ld_x % 8 != 0 || q.ld_dy % 8 != 0 || q.dy_col0 % 64 != 0)
      return NER_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(q.x_bf16) & 15) || (reinterpret_cast<uintptr_t>(q.dy_bf16) & 15) ||
        (reinterpret_cast<uintptr_t>(q.dw) & 15))
      return NER_ERR_INVALID_ARG;
    int rc = make_map_bf16_box64(&g.a[p], q.x_bf16, (uint64_t)rows, (uint64_t)q.ld_x);
    if (rc != NER_OK) return rc;
    rc = make_map_bf16_box64(&g.b[p], q.dy_bf16, (uint64_t)rows, (uint64_t)q.ld_dy);
    if (rc != NER_OK) return rc;
    rc = make_map_out(&g.c[p], q.dw, (uint64_t)q.k_in, (uint64_t)q.n_out, true);
    if (rc != NER_OK) return rc;
    g.dw[p] = q.dw;
    g.m[p] = q.k_in;
    g.n[p] = q.n_out;
    g.b_col0[p] = q.dy_col0;
    g.tile_start[p + 1] = g.tile_start[p] + (q.k_in / BM) * (q.n_out / WG_BN);
  }
  for (int p = count; p < WG_MAXP; ++p) {
    g.a[p] = g.a[0]; g.b[p] = g.b[0]; g.c[p] = g.c[0];
    g.dw[p] = nullptr; g.m[p] = g.n[p] = g.b_col0[p] = 0;
    g.tile_start[p + 1] = g.tile_start[count];
  }
  const size_t smem = Cfg<WG_BN>::SMEM;
  auto kern = gemm_wgrad_group_kernel;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  const int tiles = g.tile_start[count];
  const int grid = tiles < sm_count() ? tiles : sm_count();
  e = ner_launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), smem, static_cast<cudaStream_t>(stream), g);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  return ner_launch_status();
}

extern "C" int ner_gemm_bf16(const void* A, const void* Wt, const float* bias, const float* residual, void* out,
                             int M, int N, int K, int epilogue, int tile_n, ner_stream_t stream) {
  if (M < 0 || N < 1 || K < 1) return NER_ERR_INVALID_ARG;
  if (M == 0) return NER_OK;
  if (!A || !Wt || !out) return NER_ERR_INVALID_ARG;
  if ((epilogue < NER_EPI_F32 || epilogue > NER_EPI_RES_RELU_F32) && epilogue != NER_EPI_DIAG_DISCARD) return NER_ERR_INVALID_ARG;
  if ((epilogue == NER_EPI_RES_F32 || epilogue == NER_EPI_RES_RELU_F32) && !residual) return NER_ERR_INVALID_ARG;
  if ((K % 8) != 0 || (N % 32) != 0) return NER_ERR_UNSUPPORTED;  // 16-B TMA strides, 32-col epilogue chunks
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(Wt) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return NER_ERR_INVALID_ARG;
  EpiArgs ep{bias, residual, out, epilogue};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int bn = tile_n;
  bool sk = false;
  if (bn == NER_TILE_SK_256 || bn == NER_TILE_SK_128) {
    sk = true;
    bn = (bn == NER_TILE_SK_256) ? 256 : 128;
  } else if (bn == NER_TILE_AUTO_THROUGHPUT || (bn == 0 && gemm_auto_policy() == 1)) {
    // throughput policy: several streams keep the SMs busy, so take the tile with the best FLOP rate
    bn = (N % 256 == 0) ? 256 : (N % 192 == 0) ? 192 : 128;
  }
  if (bn == 0) {
    // Cost model in units of one 128x256 k-block per CTA.  Whole-tile scheduling: ceil(tiles/SMs) waves
    // of num_kb k-blocks, the relative tile costs measured on B200 (profiles/): the 128-wide tile is
    // smem-bandwidth-bound.  Stream-K (128x256 tiles): every CTA gets ceil(tiles*num_kb/SMs) k-blocks
    // plus a fixed charge for parking / adding one partial accumulator.
    const int mt = (M + BM - 1) / BM, sms = sm_count(), num_kb = (K + BK - 1) / BK;
    const int cand[3] = {256, 192, 128};
    const float cost[3] = {1.00f, 0.80f, 0.64f};
    float best = 1e30f;
    bn = 128;
    for (int i = 0; i < 3; ++i) {
      if (N % cand[i] != 0) continue;
      const int tiles = mt * (N / cand[i]);
      const float t = (float)((tiles + sms - 1) / sms) * cost[i] * (float)num_kb;
      if (t < best - 1e-6f) {
        best = t;
        bn = cand[i];
      }
    }
    if (N % 256 == 0) {
      const long long units = (long long)mt * (N / 256) * num_kb;
      const float t_sk = (float)((units + sms - 1) / sms) + kSkFixupCost;
      if (units >= sms && t_sk < 0.9f * best) {
        sk = true;
        bn = 256;
      }
    }
  }
  switch (bn) {
    case 256: return launch_gemm<256>(A, Wt, ep, M, N, K, st, sk);
    case 192: return launch_gemm<192>(A, Wt, ep, M, N, K, st);
    case 128: return launch_gemm<128>(A, Wt, ep, M, N, K, st, sk);
    case 64: return launch_gemm<64>(A, Wt, ep, M, N, K, st);
    case NER_TILE_2CTA_256: return launch_gemm2<256>(A, Wt, ep, M, N, K, st);
    case NER_TILE_2CTA_128: return launch_gemm2<128>(A, Wt, ep, M, N, K, st);
    default: return NER_ERR_INVALID_ARG;
  }
}
