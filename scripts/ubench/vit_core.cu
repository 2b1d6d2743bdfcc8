// Compute-only model of the Viterbi DP step (K = 10, one thread per sequence, everything in registers) in several
// instruction formulations: cycles per step per warp at 1..4 warps per SM sub-partition.  No memory traffic: this is the
// issue/pipe bound of each formulation, the number the real kernel cannot beat.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o vit_core vit_core.cu && ./vit_core
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int K = 10, KP = 5;
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fset_ne(float a, float b) { float r; asm("set.ne.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

// ADD: 0 = FADD2 pairs, 1 = scalar FADD.  MAXV: 0 = FMNMX3 chain, 1 = FMNMX (2-input) chain.
// IDX: 0 = FSETP + @p IMAD, 1 = FSET + FFMA Horner, 2 = FSETP + @p FFMA (float index), 3 = strict '>' scan with FSEL + SEL
template <int ADD, int MAXV, int IDX>
__device__ __forceinline__ void step(float (&s)[K], const float (&tr)[K][K], const float (&x)[K], uint32_t& bp_lo, uint32_t& bp_hi,
                                     uint32_t zero, float zerof) {
  float m[K];
  uint32_t wlo = 0, whi = 0;
  float fw[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < K; ++j) {
    float v[K];
    if (ADD == 0) {
#pragma unroll
      for (int p = 0; p < KP; ++p) upk2(add2(pk2(s[2 * p], s[2 * p + 1]), pk2(tr[j][2 * p], tr[j][2 * p + 1])), v[2 * p], v[2 * p + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < K; ++i) v[i] = s[i] + tr[j][i];
    }
    if (IDX == 3) {
      float best = v[0];
      uint32_t arg = 0;
#pragma unroll
      for (int i = 1; i < K; ++i)
        if (v[i] > best) { best = v[i]; arg = i; }
      m[j] = best;
      if (j < 8) wlo |= arg << (4 * j); else whi |= arg << (4 * (j - 8));
      continue;
    }
    float mj = v[0];
    if (MAXV == 0) {
      mj = max3(v[0], v[1], v[2]); mj = max3(mj, v[3], v[4]); mj = max3(mj, v[5], v[6]); mj = max3(mj, v[7], v[8]); mj = fmaxf(mj, v[9]);
    } else {
#pragma unroll
      for (int i = 1; i < K; ++i) mj = fmaxf(mj, v[i]);
    }
    m[j] = mj;
    if (IDX == 0) {
      uint32_t ix = 9u << (4 * (j & 7));
#pragma unroll
      for (int i = K - 2; i >= 0; --i)
        asm("{.reg .pred p; setp.eq.f32 p, %1, %2; @p mad.lo.u32 %0, %0, %3, %4;}" : "+r"(ix) : "f"(v[i]), "f"(mj), "r"(zero), "r"((uint32_t)i << (4 * (j & 7))));
      if (j < 8) wlo |= ix; else whi |= ix;
    } else if (IDX == 1) {
      float h = fset_ne(v[K - 2], mj);
#pragma unroll
      for (int i = K - 3; i >= 0; --i) { const float e = fset_ne(v[i], mj); h = fmaf(e, h, e); }
      fw[j >> 2] = fmaf(h, (float)(1 << (4 * (j & 3))), fw[j >> 2]);
    } else {
      float ixf = 9.f * (float)(1 << (4 * (j & 3)));
#pragma unroll
      for (int i = K - 2; i >= 0; --i)
        asm("{.reg .pred p; setp.eq.f32 p, %1, %2; @p fma.rn.f32 %0, %0, %3, %4;}" : "+f"(ixf) : "f"(v[i]), "f"(mj), "f"(zerof), "f"((float)(i << (4 * (j & 3)))));
      fw[j >> 2] += ixf;
    }
  }
  if (IDX == 1 || IDX == 2) {
    wlo = __float2uint_rn(fw[0]) | (__float2uint_rn(fw[1]) << 16);
    whi = __float2uint_rn(fw[2]);
  }
#pragma unroll
  for (int j = 0; j < K; ++j) s[j] = m[j] + x[j];
  bp_lo ^= wlo;
  bp_hi ^= whi;
}

template <int ADD, int MAXV, int IDX, int MINB>
__global__ void __launch_bounds__(128, MINB) core(float* out, long long* cyc, const float* tin, int steps, uint32_t zero, float zerof) {
  float tr[K][K], s[K], x[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
#pragma unroll
    for (int i = 0; i < K; ++i) tr[j][i] = tin[j * K + i];
    s[j] = tin[100 + j] * (float)(threadIdx.x + 1);
    x[j] = tin[110 + j] + (float)threadIdx.x * 0.001f;
  }
  uint32_t lo = 0, hi = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < steps; ++t) {
    step<ADD, MAXV, IDX>(s, tr, x, lo, hi, zero, zerof);
    x[t % K] = -x[t % K];        // keeps the emissions from being loop invariant (one FADD-class op per step)
  }
  const long long t1 = clock64();
  float acc = (float)lo + (float)hi;
#pragma unroll
  for (int j = 0; j < K; ++j) acc += s[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ADD, int MAXV, int IDX, int MINB>
void run(const char* name, float* out, long long* cyc, const float* tin) {
  const int steps = 512;
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, core<ADD, MAXV, IDX, MINB>);
  printf("%-44s regs=%3d:", name, fa.numRegs);
  for (int ctas = 1; ctas <= MINB; ++ctas) {       // 128-thread CTAs: ctas = warps per sub-partition
    core<ADD, MAXV, IDX, MINB><<<148 * ctas, 128>>>(out, cyc, tin, 16, 0u, 0.f);
    core<ADD, MAXV, IDX, MINB><<<148 * ctas, 128>>>(out, cyc, tin, steps, 0u, 0.f);
    cudaDeviceSynchronize();
    static long long h[148 * 8];
    cudaMemcpy(h, cyc, sizeof(long long) * 148 * ctas, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148 * ctas; ++i) mean += (double)h[i];
    mean /= 148 * ctas;
    printf("  w=%d %6.1f cyc/step/warp-slot", ctas, mean / steps / ctas);
  }
  printf("\n");
}

int main() {
  float *out, *tin;
  long long* cyc;
  cudaMalloc(&out, 148 * 8 * 128 * sizeof(float));
  cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
  float h[128];
  for (int i = 0; i < 128; ++i) h[i] = 0.37f * (float)((i * 7919) % 31) - 3.f;
  cudaMalloc(&tin, sizeof(h));
  cudaMemcpy(tin, h, sizeof(h), cudaMemcpyHostToDevice);
  // "cyc/step/warp-slot" = cycles one sub-partition spends per warp-step (32 sequences x 1 time step): lower is better
  run<0, 0, 3, 3>("strict > scan: FADD2 + FSETP/FSEL/SEL", out, cyc, tin);
  run<0, 0, 0, 2>("FADD2 + FMNMX3 + FSETP/@pIMAD", out, cyc, tin);
  run<0, 0, 0, 3>("FADD2 + FMNMX3 + FSETP/@pIMAD (<=168 regs)", out, cyc, tin);
  run<0, 0, 1, 3>("FADD2 + FMNMX3 + FSET/FFMA", out, cyc, tin);
  run<0, 0, 2, 3>("FADD2 + FMNMX3 + FSETP/@pFFMA", out, cyc, tin);
  run<1, 0, 2, 3>("FADD  + FMNMX3 + FSETP/@pFFMA", out, cyc, tin);
  run<0, 1, 2, 3>("FADD2 + FMNMX  + FSETP/@pFFMA", out, cyc, tin);
  run<1, 1, 2, 3>("FADD  + FMNMX  + FSETP/@pFFMA", out, cyc, tin);
  run<1, 0, 1, 3>("FADD  + FMNMX3 + FSET/FFMA", out, cyc, tin);
  run<1, 1, 1, 3>("FADD  + FMNMX  + FSET/FFMA", out, cyc, tin);
  run<1, 0, 0, 3>("FADD  + FMNMX3 + FSETP/@pIMAD", out, cyc, tin);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
