// Microbenchmark: what FP32 issue rate does one B200 SM sustain, per instruction form and warp count?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp_issue_rate benchmarks/fp_issue_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void k(float* out, int iters, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
  float2 y[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = make_float2(x[2 * i], x[2 * i + 1]);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (KIND == 0) {  // FFMA, 3 register operands, 8 independent chains
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
      } else if (KIND == 1) {  // FADD, 2 register operands
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = x[i] + a;
      } else if (KIND == 2) {  // FFMA2 packed
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          unsigned long long p, q, s, d;
          asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(y[i].x), "f"(y[i].y));
          asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(a), "f"(a));
          asm("mov.b64 %0, {%1, %2};" : "=l"(s) : "f"(b), "f"(b));
          asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(p), "l"(q), "l"(s));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(y[i].x), "=f"(y[i].y) : "l"(d));
        }
      } else {  // FADD alternating with FMUL on different registers (mix)
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          x[i] = x[i] + a;
          x[i + 1] = x[i + 1] * b;
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += y[i].x + y[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

template <int KIND>
void run(const char* name, int per_iter_instr) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 2000;
  for (int warps : {4, 8, 16, 32}) {
    k<KIND><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    k<KIND><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    float cyc;
    cudaMemcpy(&cyc, out, 4, cudaMemcpyDeviceToHost);
    const double instr = (double)iters * 8 * per_iter_instr * warps;  // warp instructions per SM
    printf("%-28s warps/SM %2d  cycles %9.0f  warp-instr/cycle/SM %.2f\n", name, warps, cyc, instr / cyc);
  }
  cudaFree(out);
}

int main() {
  run<0>("FFMA r,r,r (8 chains)", 8);
  run<1>("FADD r,r (8 chains)", 8);
  run<2>("FFMA2 packed (4 chains)", 4);
  run<3>("FADD/FMUL mix (8 chains)", 8);
  return 0;
}
