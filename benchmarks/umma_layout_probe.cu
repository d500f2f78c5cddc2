// minimal tcgen05 self-test: D = A * B with B = identity, K-major and MN-major A
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3ffff) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_major, int b_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__global__ void test(float* out, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* A = (float*)smem;                 // 16 KB: K-major: [128 rows][32 k]; MN-major: [4 groups][32 k rows][32 m]
  float* B = (float*)(smem + 16384);       // [32 n rows][32 k]
  uint64_t* bar = (uint64_t*)(smem + 16384 + 4096);
  uint32_t* slot = (uint32_t*)(smem + 16384 + 4096 + 16);
  int tid = threadIdx.x;
  for (int i = tid; i < 128 * 32; i += 128) {
    int m = i / 32, k = i % 32;
    float v = m + k / 64.0f;
    int off;
    if (mode == 0) off = m * 128 + (((k / 4) ^ (m % 8)) << 4) + (k % 4) * 4;                 // K-major SW128
    else { int g = m / 32, mm = m % 32; off = g * 4096 + k * 128 + (((mm / 4) ^ (k % 8)) << 4) + (mm % 4) * 4; }  // MN-major SW128
    *(float*)((char*)A + off) = v;
  }
  if (mode >= 1) for (int i = tid; i < 4096; i += 128) A[i] = (float)i;
  for (int i = tid; i < 32 * 32; i += 128) {
    int n = i / 32, k = i % 32;
    int off = n * 128 + (((k / 4) ^ (n % 8)) << 4) + (k % 4) * 4;
    *(float*)((char*)B + off) = (n == k) ? 1.0f : 0.0f;
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tb = *slot;
  if (tid == 0) {
    uint32_t idesc = make_idesc(128, 32, mode >= 1, 0);
    const int nks = mode == 0 ? 4 : 1;
    for (int ks = 0; ks < nks; ++ks) {
      const uint64_t base32 = (make_desc(smem_u32(A), 0, 0) & ~(7ull << 61)) | (1ull << 61);  /* SWIZZLE_128B_BASE32B */
      auto lbsb = [](uint32_t lbo, uint32_t sbo) { return ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32); };
      uint64_t da = mode == 0 ? make_desc(smem_u32(A) + ks * 32, 16, 1024)
                  : mode == 1 ? base32 | lbsb(4096, 512)
                  : mode == 2 ? base32 | lbsb(512, 4096)
                  : mode == 3 ? base32 | lbsb(4096, 1024)
                  : mode == 4 ? base32 | lbsb(1024, 4096)
                  : base32 | lbsb(128, 512);
      uint64_t db = make_desc(smem_u32(B) + ks * 32, 16, 1024);
      uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  uint32_t taddr = tb + ((uint32_t)((tid / 32) * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out[tid * 32 + j] = __uint_as_float(r[j]);
  out[128 * 32] = __uint_as_float(tb);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tb) : "memory");
}
int main() {
  float* d; cudaMalloc(&d, (128 * 32 + 1) * 4);
  cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int mode = 0; mode < 6; ++mode) {
    cudaMemset(d, 0, (128 * 32 + 1) * 4);
    test<<<1, 128, 32768>>>(d, mode);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(128 * 32 + 1);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; 
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) { float want = m + k / 64.0f; if (fabsf(h[m * 32 + k] - want) > 0.02f) ++bad; }
    uint32_t tb; memcpy(&tb, &h[128*32], 4);
    printf("mode %d: %s bad %d\n", mode, cudaGetErrorString(e), bad);
    if (mode >= 1) {
      const int ms[] = {0, 1, 2, 3, 4, 5, 8, 16, 31, 32, 33, 64, 127};
      for (int m : ms) { printf("  m=%3d:", m); for (int k = 0; k < 8; ++k) printf(" %6.0f", h[m * 32 + k]); printf("\n"); }
    }
  }
  return 0;
}
