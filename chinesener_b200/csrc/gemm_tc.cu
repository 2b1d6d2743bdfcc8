// Dense layers of the encoder on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   out[M,N] = epilogue( A[M,K] · Wt[N,K]^T + bias[N] )       A, Wt bf16 (K-major), fp32 accumulate
//
// Replaces the tf.layers.dense / BertModel dense_layer matmuls executed inside
// reference tools/layer.py:68-77 (bert_base.bert.modeling) and model/bert_bilstm_crf.py:26,
// and the input projection half of the LSTMCell matmul (tools/layer.py:16,35).
//
// Kernel shape (persistent, warp-specialised, 192 threads = 6 warps, 1 CTA/SM):
//   warp 0   TMA producer : cp.async.bulk.tensor 2-D loads of a 128x64 A box and a BNx64 B box
//                           per k-block into a STAGES-deep smem ring (SWIZZLE_128B), mbarrier tx
//   warp 1   MMA issuer   : one thread issues tcgen05.mma.kind::f16 (128 x BN x 16) x 4 per
//                           k-block; tcgen05.commit releases smem slots / publishes the accumulator
//   warps 2-5 epilogue    : tcgen05.ld 32x32b (TMEM lane quarter = warp%4) -> bias / GELU /
//                           residual -> 16-byte global stores
// Two TMEM accumulator stages (2*BN columns) let the epilogue of tile i overlap the
// main loop of tile i+1.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;       // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : ((BN == 128) ? 6 : 8);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr size_t SMEM = (size_t)STAGES * (A_BYTES + B_BYTES) + 256 /*barriers*/ + 1024 /*align slack*/;
};

struct EpiArgs {
  const float* bias;      // [N] or null
  const float* residual;  // [M,N] fp32 or null (EPI_RES_F32)
  void* out;              // [M,N] bf16 or fp32
  int mode;
};

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5x(1+tanh(sqrt(2/pi)(x+0.044715x^3)))  (google-research/bert modeling.gelu)
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float t = 1.f - 2.f / (__expf(2.f * u) + 1.f);
  return 0.5f * x * (1.f + t);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.7071067811865476f)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    EpiArgs ep, int M, int N, int K) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // SWIZZLE_128B needs 1024-B alignment
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * C::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (C::A_BYTES + C::B_BYTES));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
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
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * UMMA_K * 2);
            umma_f16(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
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
    // ===================== epilogue warps (2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < M;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row_ok && col0 < N) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (ep.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(b4 + i);
              v[4 * i + 0] += b.x;
              v[4 * i + 1] += b.y;
              v[4 * i + 2] += b.z;
              v[4 * i + 3] += b.w;
            }
          }
          const size_t off = (size_t)row * N + col0;
          if (ep.mode == NER_EPI_F32 || ep.mode == NER_EPI_RES_F32) {
            if (ep.mode == NER_EPI_RES_F32) {
              const float4* r4 = reinterpret_cast<const float4*>(ep.residual + off);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 b = __ldg(r4 + i);
                v[4 * i + 0] += b.x;
                v[4 * i + 1] += b.y;
                v[4 * i + 2] += b.z;
                v[4 * i + 3] += b.w;
              }
            }
            float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + off);
#pragma unroll
            for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
            if (ep.mode == NER_EPI_GELU_TANH_BF16) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
            } else if (ep.mode == NER_EPI_GELU_ERF_BF16) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
            } else if (ep.mode == NER_EPI_RELU_BF16) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + off);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o4[i] = make_uint4(pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                 pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          }
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

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
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

template <int BN>
int launch_gemm(const void* A, const void* Wt, EpiArgs ep, int M, int N, int K, cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc = make_map_bf16_2d(&ma, A, (uint64_t)M, (uint64_t)K, BM);
  if (rc != NER_OK) return rc;
  rc = make_map_bf16_2d(&mb, Wt, (uint64_t)N, (uint64_t)K, BN);
  if (rc != NER_OK) return rc;
  auto kern = gemm_bf16_tc_kernel<BN>;
  const size_t smem = Cfg<BN>::SMEM;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return NER_ERR_CUDA_BASE - (int)e;
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, NUM_THREADS, smem, st>>>(ma, mb, ep, M, N, K);
  return ner_launch_status();
}

}  // namespace

extern "C" int ner_gemm_bf16(const void* A, const void* Wt, const float* bias, const float* residual, void* out,
                             int M, int N, int K, int epilogue, int tile_n, ner_stream_t stream) {
  if (M < 0 || N < 1 || K < 1) return NER_ERR_INVALID_ARG;
  if (M == 0) return NER_OK;
  if (!A || !Wt || !out) return NER_ERR_INVALID_ARG;
  if (epilogue < NER_EPI_F32 || epilogue > NER_EPI_RES_F32) return NER_ERR_INVALID_ARG;
  if (epilogue == NER_EPI_RES_F32 && !residual) return NER_ERR_INVALID_ARG;
  if ((K % 8) != 0 || (N % 32) != 0) return NER_ERR_UNSUPPORTED;  // 16-B TMA strides, 32-col epilogue chunks
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(Wt) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return NER_ERR_INVALID_ARG;
  EpiArgs ep{bias, residual, out, epilogue};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int bn = tile_n;
  if (bn == 0) bn = (N % 256 == 0 && ((M + BM - 1) / BM) * (N / 256) >= sm_count()) ? 256 : 128;
  if (bn == 256) return launch_gemm<256>(A, Wt, ep, M, N, K, st);
  if (bn == 128) return launch_gemm<128>(A, Wt, ep, M, N, K, st);
  if (bn == 64) return launch_gemm<64>(A, Wt, ep, M, N, K, st);
  return NER_ERR_INVALID_ARG;
}
