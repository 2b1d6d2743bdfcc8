// Issue-rate microbenchmark for the instruction mixes of the CRF kernels (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
// Prints warp-instructions per cycle per SM sub-partition for each op (or op pair) at 1, 2 and 4 warps per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NACC 8
#define UNR 8

template <int OP>
__device__ __forceinline__ void body(float (&f)[NACC], uint32_t (&u)[NACC], unsigned long long (&d)[NACC], float a, float b,
                                     uint32_t z) {
#pragma unroll
  for (int r = 0; r < UNR; ++r) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      if constexpr (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(a), "f"(b));
      if constexpr (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[k]) : "l"(d[(k + 1) % NACC]));
      if constexpr (OP == 2) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(f[(k + 3) % NACC]));
      if constexpr (OP == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(f[(k + 3) % NACC]), "f"(f[(k + 5) % NACC]));
      if constexpr (OP == 4)
        asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(f[k]) : "f"(f[(k + 3) % NACC]), "f"(a));
      if constexpr (OP == 5)
        asm volatile("{.reg .pred p; setp.eq.f32 p, %1, %2; @p mad.lo.u32 %0, %0, %3, 7;}" : "+r"(u[k]) : "f"(f[k]), "f"(f[(k + 3) % NACC]), "r"(z));
      if constexpr (OP == 6) asm volatile("mad.lo.u32 %0, %0, %1, 7;" : "+r"(u[k]) : "r"(z));
      if constexpr (OP == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k]) : "r"(u[(k + 3) % NACC]), "r"(z));
      if constexpr (OP == 8) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(u[k]) : "r"(u[(k + 3) % NACC]));
      if constexpr (OP == 9) asm volatile("set.ne.f32.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(f[(k + 3) % NACC]));
      if constexpr (OP == 10) {   // FSETP only: the predicate is folded into the next compare of the same chain
        asm volatile("{.reg .pred p; setp.eq.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(f[k]) : "f"(f[(k + 3) % NACC]), "f"(a));
      }
      if constexpr (OP == 11) {   // FMNMX3 + FADD2
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(f[(k + 3) % NACC]), "f"(f[(k + 5) % NACC]));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[k]) : "l"(d[(k + 1) % NACC]));
      }
      if constexpr (OP == 12) {   // FMNMX + FFMA
        asm volatile("max.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(a));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(k + 4) % NACC]) : "f"(a), "f"(b));
      }
      if constexpr (OP == 13) {   // LOP3 + IMAD
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k]) : "r"(u[(k + 3) % NACC]), "r"(z));
        asm volatile("mad.lo.u32 %0, %0, %1, 7;" : "+r"(u[(k + 4) % NACC]) : "r"(z));
      }
      if constexpr (OP == 14) {   // the Viterbi tag step in miniature: FADD2, FMNMX3, FSETP + @p IMAD x2
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[k]) : "l"(d[(k + 1) % NACC]));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(f[(k + 3) % NACC]), "f"(f[(k + 5) % NACC]));
        asm volatile("{.reg .pred p; setp.eq.f32 p, %1, %2; @p mad.lo.u32 %0, %0, %3, 7;}" : "+r"(u[k]) : "f"(f[k]), "f"(f[(k + 3) % NACC]), "r"(z));
        asm volatile("{.reg .pred p; setp.eq.f32 p, %1, %2; @p mad.lo.u32 %0, %0, %3, 9;}" : "+r"(u[(k + 4) % NACC]) : "f"(f[k]), "f"(f[(k + 2) % NACC]), "r"(z));
      }
      if constexpr (OP == 15) {   // FSET + FFMA (the predicate-free pair)
        float e;
        asm volatile("set.ne.f32.f32 %0, %1, %2;" : "=f"(e) : "f"(f[k]), "f"(f[(k + 3) % NACC]));
        asm volatile("fma.rn.f32 %0, %1, %0, %1;" : "+f"(f[(k + 4) % NACC]) : "f"(e));
      }
      if constexpr (OP == 16) {   // FSETP + SEL where both are needed (the old kernel's pair): value select + index select
        asm volatile("{.reg .pred p; setp.gt.f32 p, %2, %0; selp.f32 %0, %2, %0, p; selp.u32 %1, 5, %1, p;}" : "+f"(f[k]), "+r"(u[k]) : "f"(f[(k + 3) % NACC]));
      }
    }
  }
}

template <int OP>
__global__ void __launch_bounds__(512, 1) bench(float* out, long long* cyc, int iters, float a, float b, uint32_t z) {
  float f[NACC];
  uint32_t u[NACC];
  unsigned long long d[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    f[k] = a * (float)(k + threadIdx.x);
    u[k] = z + k + threadIdx.x;
    d[k] = ((unsigned long long)__float_as_uint(a + k) << 32) | __float_as_uint(b + k);
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) body<OP>(f, u, d, a, b, z);
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc += f[k] + (float)u[k] + (float)(d[k] >> 40);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static const char* NAMES[] = {"FFMA", "FADD2", "FMNMX", "FMNMX3", "FSETP+FSEL", "FSETP+@pIMAD", "IMAD", "LOP3", "SHF", "FSET.BF",
                              "FSETP.EQ+FSEL", "FMNMX3+FADD2", "FMNMX+FFMA", "LOP3+IMAD", "tag-step mix (6 instr)", "FSET+FFMA",
                              "FSETP+FSEL+SEL"};
static const int NINSTR[] = {1, 1, 1, 1, 2, 2, 1, 1, 1, 1, 2, 2, 2, 2, 6, 2, 3};

template <int OP>
void run(float* out, long long* cyc) {
  const int iters = 2000;
  for (int wps = 1; wps <= 4; wps *= 2) {
    const int threads = 128 * wps;
    bench<OP><<<148, threads>>>(out, cyc, 10, 1.0001f, 0.5f, 0u);
    bench<OP><<<148, threads>>>(out, cyc, iters, 1.0001f, 0.5f, 0u);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += (double)h[i];
    mean /= 148;
    const double instr = (double)iters * UNR * NACC * NINSTR[OP] * wps;   // warp-instructions per sub-partition
    printf("%-24s warps/SMSP=%d  %.3f warp-instr/cycle/SMSP  (%.0f cycles)\n", NAMES[OP], wps, instr / mean, mean);
  }
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  run<0>(out, cyc); run<1>(out, cyc); run<2>(out, cyc); run<3>(out, cyc); run<4>(out, cyc); run<5>(out, cyc);
  run<6>(out, cyc); run<7>(out, cyc); run<8>(out, cyc); run<9>(out, cyc); run<10>(out, cyc); run<11>(out, cyc);
  run<12>(out, cyc); run<13>(out, cyc); run<14>(out, cyc); run<15>(out, cyc); run<16>(out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
