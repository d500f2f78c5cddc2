// Microbenchmark: does a packed fma.rn.f32x2 (SASS FFMA2) leave the warp scheduler's issue slot free for a
// second instruction of another pipe?  fp_issue_rate.cu showed that FFMA2 holds the FMA pipe for two cycles
// (same flops per cycle as scalar FFMA).  The fused kernel is bound by ISSUE slots, about half of them FP32,
// so what matters is whether FFMA2 + {IADD3, LOP3, LDS, SHFL} issue in 2 cycles (slot free) or 3 (slot held).
//
// Each variant runs a loop whose body is N_fp FP instructions and N_other "other" instructions on independent
// registers; we report cycles per loop body per warp scheduler at 8 warps per scheduler.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/bench/f32x2_coissue benchmarks/f32x2_coissue.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// KIND: 0 = 8 scalar FFMA; 1 = 4 FFMA2 (same flops); 2 = 8 FFMA + 4 ALU; 3 = 4 FFMA2 + 4 ALU;
//       4 = 8 FFMA + 4 LDS; 5 = 4 FFMA2 + 4 LDS; 6 = 4 ALU alone; 7 = 4 LDS alone; 8 = 4 FFMA2 + 8 ALU; 9 = 8 FFMA + 8 ALU;
//       10 = 4 FFMA2 + 4 FFMA, 11 = 4 FFMA2 + 8 FFMA (do packed and scalar FMAs share one pipe?)
template <int KIND>
__global__ void k(float* out, int iters, float a, float b, int ia) {
  __shared__ float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 0.5f;
  __syncthreads();
  float x[8];
  unsigned long long y[4];
  int n[8];
  float l[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i, n[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = pack(x[2 * i], x[2 * i + 1]);
  const unsigned long long pa = pack(a, a), pb = pack(b, b);
  const float* sp = sm + (threadIdx.x & 31);
  constexpr bool scalar = KIND == 0 || KIND == 2 || KIND == 4 || KIND == 9 || KIND == 11;
  constexpr bool scalar4 = KIND == 10;
  constexpr bool packed = KIND == 1 || KIND == 3 || KIND == 5 || KIND == 8 || KIND == 10 || KIND == 11;
  constexpr int n_alu = (KIND == 2 || KIND == 3 || KIND == 6) ? 4 : ((KIND == 8 || KIND == 9) ? 8 : 0);
  constexpr bool lds = KIND == 4 || KIND == 5 || KIND == 7;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (scalar) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
      }
      if (scalar4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = fmaf(x[i], a, b);
      }
      if (packed) {
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = ffma2(y[i], pa, pb);
      }
#pragma unroll
      for (int i = 0; i < n_alu; ++i) n[i] = (n[i] ^ ia) + it;  // LOP3 + IADD3 -> 2 ALU instructions each
      if (lds) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sp + 32 * ((r * 4 + i) & 63))));
          l[i] += v;
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + n[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(y[i]));
    s += lo + hi + l[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

template <int KIND>
void run(const char* name) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 1000;
  for (int warps : {8, 16, 32}) {
    k<KIND><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f, 12345);
    cudaDeviceSynchronize();
    k<KIND><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f, 12345);
    cudaDeviceSynchronize();
    float cyc;
    cudaMemcpy(&cyc, out, 4, cudaMemcpyDeviceToHost);
    // loop bodies per scheduler = iters * 8 * (warps / 4)
    printf("%-34s warps/SM %2d  cycles per body per scheduler %.2f\n", name, warps, cyc / (iters * 8.0 * (warps / 4.0)));
  }
  cudaFree(out);
}

int main() {
  run<0>("8 FFMA");
  run<1>("4 FFMA2");
  run<6>("8 ALU (4x LOP3+IADD3)");
  run<7>("4 LDS (+4 FADD)");
  run<2>("8 FFMA + 8 ALU");
  run<3>("4 FFMA2 + 8 ALU");
  run<9>("8 FFMA + 16 ALU");
  run<8>("4 FFMA2 + 16 ALU");
  run<4>("8 FFMA + 4 LDS (+4 FADD)");
  run<5>("4 FFMA2 + 4 LDS (+4 FADD)");
  run<10>("4 FFMA2 + 4 FFMA");
  run<11>("4 FFMA2 + 8 FFMA");
  return 0;
}
