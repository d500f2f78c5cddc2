// Microbenchmark: issue rate of the legacy warp-level tensor path (mma.sync, SASS HMMA) on one B200 SM,
// for the two shapes a warp-level DFT / filterbank would use, alone and interleaved with FFMA.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_sync_rate benchmarks/mma_sync_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// KIND 0: tf32 m16n8k8, 8 independent accumulators; 1: f16 m16n8k16; 2: tf32 with 8 FFMA per MMA; 3: one dependent chain (latency)
template <int KIND>
__global__ void k(float* out, int iters) {
  float d[8][4];
  unsigned a[4], b[2];
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = 0x3c003c00u + threadIdx.x + j;
  b[0] = 0x3c003c00u;
  b[1] = 0x38003800u + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0 || KIND == 2) mma_tf32(d[i], a, b);
      if (KIND == 1) mma_f16(d[i], a, b);
      if (KIND == 3) mma_tf32(d[0], a, b);
      if (KIND == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], 1.0001f, 0.5f);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

template <int KIND>
void run(const char* name, double flop_per_mma) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16, 24, 32}) {
    k<KIND><<<148, warps * 32>>>(out, iters);
    cudaDeviceSynchronize();
    k<KIND><<<148, warps * 32>>>(out, iters);
    cudaDeviceSynchronize();
    float cyc;
    cudaMemcpy(&cyc, out, 4, cudaMemcpyDeviceToHost);
    const double mmas = (double)iters * 8 * warps;  // warp-level MMAs per SM
    printf("%-34s warps/SM %2d  cycles %9.0f  cycles/MMA/SM %.2f  flop/cycle/SM %.0f\n", name, warps, cyc, cyc / mmas,
           mmas * flop_per_mma / cyc);
  }
  cudaFree(out);
}

int main() {
  run<0>("mma.sync m16n8k8 tf32", 2.0 * 16 * 8 * 8);
  run<1>("mma.sync m16n8k16 f16", 2.0 * 16 * 8 * 16);
  run<2>("m16n8k8 tf32 + 8 FFMA each", 2.0 * 16 * 8 * 8);
  run<3>("m16n8k8 tf32 dependent chain", 2.0 * 16 * 8 * 8);
  return 0;
}
