// Microbenchmark: issue rate of packed (f32x2) vs scalar FP32 ops on one SMSP (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -o fp2_tput fp2_tput.cu && ./fp2_tput
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long p2;
#define N 4096
template <int MODE>
__global__ void k(float* out, long long* cyc, float a, float b) {
  p2 x[8]; float y[16];
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(a + i), "f"(b + i));
  for (int i = 0; i < 16; ++i) y[i] = a + i;
  p2 m; asm("mov.b64 %0, {%1, %2};" : "=l"(m) : "f"(1.0001f), "f"(0.9999f));
  p2 c; asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(b), "f"(a));
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N; ++it) {
    if (MODE == 0) {  // 8 independent FFMA2
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(m), "l"(c));
    } else if (MODE == 1) {  // 16 independent scalar FFMA (3 register operands)
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(a), "f"(b));
    } else if (MODE == 2) {  // 8 FADD2
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(c));
    } else if (MODE == 3) {  // 8 FMUL2
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(m));
    } else if (MODE == 4) {  // 16 scalar FADD
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(b));
    } else if (MODE == 5) {  // 8 FFMA2 interleaved with 8 integer adds (issue slots)
      int z = it;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(m), "l"(c));
        asm volatile("add.s32 %0, %0, 3;" : "+r"(z));
      }
      y[0] += __int_as_float(z & 1);
    } else if (MODE == 6) {  // dependent chain of FFMA2 (latency)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[0]) : "l"(m), "l"(c));
    } else if (MODE == 7) {  // dependent chain of scalar FFMA
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[0]) : "f"(a), "f"(b));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  for (int i = 0; i < 16; ++i) s += y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int ops, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 4 * 1024 * 148); cudaMalloc(&cyc, 8);
  k<MODE><<<1, threads>>>(out, cyc, 1.5f, 0.25f); cudaDeviceSynchronize();
  k<MODE><<<1, threads>>>(out, cyc, 1.5f, 0.25f); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps/SMSP %d: %.2f cycles per instr (per warp)\n", name, threads / 128, (double)h / ((double)N * ops));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int threads : {128, 256, 512}) {
    run<0>("8x independent FFMA2", 8, threads);
    run<1>("16x independent FFMA (3-reg)", 16, threads);
    run<2>("8x independent FADD2", 8, threads);
    run<3>("8x independent FMUL2", 8, threads);
    run<4>("16x independent FADD", 16, threads);
    run<5>("8x (FFMA2 + IADD)", 16, threads);
    run<6>("dependent FFMA2 chain", 8, threads);
    run<7>("dependent FFMA chain", 8, threads);
  }
  return 0;
}
