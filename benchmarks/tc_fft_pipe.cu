// Warp-specialised, pipelined version of benchmarks/tc_fft_proto.cu (same arithmetic, same checks):
// marshal warps, one MMA-issue warp, twiddle warps and magnitude warps work on different tiles at the same
// time, accumulators double-buffered in TMEM, every hand-over an mbarrier.  It answers the question the serial
// prototype left open: how fast is the tensor-core STFT when its stages overlap?
//
// Prototype: the n_fft = 1024 real STFT of the dMel path on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), as two dense DFT stages with a 3-term TF32 split.
// It measures what DESIGN.md section 4.1 could only estimate: the accuracy of the split-precision
// tensor-core DFT against a float64 DFT, and the per-frame cost of the CUDA-core work that remains
// around the MMAs (window + split + operand layout, inter-stage twiddle, magnitudes).
//
//   frame x[n], n = 32 n1 + n2 (n1, n2 in 0..31), bin k = k1 + 32 k2
//   stage 1:  Y[k1][n2] = sum_n1 W_32^{n1 k1} x[32 n1 + n2]      real input, k1 = 0..16 -> 32 real outputs
//             GEMM  D1[(frame, n2), j] = A1[(frame, n2), n1] * B1[n1, j]      M = 4 frames x 32, K = 32, N = 32
//   twiddle:  Z[k1][n2] = W_1024^{n2 k1} Y[k1][n2]               CUDA cores, one thread per (frame, n2)
//   stage 2:  X[k1 + 32 k2] = sum_n2 W_32^{n2 k2} Z[k1][n2]       complex 32-point DFT per (frame, k1)
//             GEMM  D2[(frame, k1), (k2, c)] = A2[(frame, k1), (n2, c')] * B2      M = 7 frames x 17, K = 64, N = 64
//   bins 0..512 are read off k1 = 0..16 (k and its mirror 1024 - k), magnitude = sqrt(re^2 + im^2 + 1e-9).
//
// Every operand goes through the tensor core three times (hi*hi + lo*hi + hi*lo with hi = the 11
// leading mantissa bits, lo = the rest): the same products a 3xTF32 GEMM makes.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tc_fft_proto benchmarks/tc_fft_proto.cu
// Run:   ./tc_fft_proto            (accuracy on a few hundred frames, then timing on 59,968 frames)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));       \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

constexpr int kNfft = 1024, kHop = 256, kBins = 513;
constexpr int kTF = 7;                              // frames per tile: 7 x 17 stage-2 rows = 119 <= 128
constexpr int kWaveLen = (kTF - 1) * kHop + kNfft;  // 2560 samples staged per tile
// warp roles: 0-3 marshal (window, split, A1), 4 MMA issue + waveform copies, 5-7 idle padding so the roles below
// start on a multiple of four (a warp reads the TMEM lane quarter warp % 4), 8-15 twiddle (A2), 16-19 magnitudes
constexpr int kWarpMma = 4, kWarpTw = 8, kWarpMag = 16, kWarps = 20;
constexpr int kThreads = kWarps * 32;

// ---- shared memory map (bytes) ----------------------------------------------------------------
constexpr int kA1Frame = 4096;                   // one windowed frame, [n1 = 32 rows][n2 = 32 floats], 128B-swizzled
constexpr int kA1Half = 8 * kA1Frame;            // hi (or lo) copies of 8 frame slots (the 8th is padding)
constexpr int kA2Block = 128 * 128;              // one K block (32 of the 64 K values) of the 128 stage-2 rows
constexpr int kA2Half = 2 * kA2Block;            // hi (or lo)
constexpr int kWaveBytes = kWaveLen * 4;         // 10240
constexpr int kOffA1 = 0;                        // A1 {hi, lo}: 2 * 32768
constexpr int kOffA2 = kOffA1 + 2 * kA1Half;     // A2 {hi, lo}: 2 * 32768
constexpr int kOffWave = kOffA2 + 2 * kA2Half;   // two waveform buffers
constexpr int kB1Half = 32 * 128;                // [N = 32 rows][K = 32 floats]
constexpr int kB2Block = 64 * 128;               // [N = 64 rows][32 of the 64 K values]
constexpr int kB2Half = 2 * kB2Block;
constexpr int kOffB1 = (kOffWave + 2 * kWaveBytes + 1023) / 1024 * 1024;
constexpr int kOffB2 = kOffB1 + 2 * kB1Half;
constexpr int kOffMisc = kOffB2 + 2 * kB2Half;   // mbarriers, TMEM address
constexpr int kSmemBytes = kOffMisc + 256;
static_assert(kOffA2 % 1024 == 0 && kOffB1 % 1024 == 0 && kOffB2 % 1024 == 0, "swizzle atoms need 1024 B alignment");

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void marshal_sync() {  // the 128 threads of the marshal warps
  asm volatile("bar.sync 1, 128;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Warp-convergent forms: every lane executes the statement, one elected lane issues.  The operands are then
// warp-uniform values the compiler keeps in uniform registers; issuing from a single divergent lane instead costs
// ~13 instructions per MMA (register -> uniform-register moves) and made the issue warp the bottleneck.
template <int ACC>
__device__ __forceinline__ void umma_tf32_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "n"(ACC)
      : "memory");
}
__device__ __forceinline__ void umma_commit_warp(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
        "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
        "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 format: version 1 at bit 46, layout type at bits 61-63); lbo / sbo in bytes.
//   K-major operands : SWIZZLE_128B (2): rows of 128 B along K, 16-byte chunks XORed with (row % 8), sbo = 8-row group stride.
//   MN-major tf32 A  : SWIZZLE_128B_BASE32B (1), the only MN-major form the tf32 kind accepts (the plain 128B swizzle
//                      returns zeros): rows of 128 B along M, one row per K index, 32-byte chunks XORed with (k % 4);
//                      lbo = stride between groups of 32 M values, sbo = stride between groups of 4 K rows.
//                      Measured with benchmarks/umma_layout_probe.cu.
constexpr uint64_t kSwizzle128 = 2, kSwizzle128Base32 = 1;
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t layout = kSwizzle128) {
  return (uint64_t)((saddr & 0x3ffff) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         (layout << 61);
}
// Instruction descriptor, kind::tf32: D f32, A/B tf32, a_major / b_major: 0 = K-major, 1 = MN-major.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_major, int b_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// hi = the bits a tf32 operand keeps, lo = the exact remainder
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}

struct Params {
  const float* wav;      // one long signal; frame t covers [t*hop, t*hop + 1024)
  const float* window;   // 1024
  const float* b1;       // host-built shared-memory image of B1 {hi, lo}: 2 * kB1Half bytes
  const float* b2;       // image of B2 {hi, lo}: 2 * kB2Half bytes
  int n_frames;
  int n_tiles;
  float* mags;           // (n_frames, 513) or null (timing mode)
  float* checksum;       // one float per CTA (timing mode keeps the work alive)
  long long* stats;      // optional per-role cycle counters of CTA 0: [role][wait_a, wait_b, work, tiles]
};

// mbarrier slots
enum { kWaveFull0, kWaveFull1, kA1Full, kA1Empty, kD1Full0, kD1Full1, kD1Empty0, kD1Empty1, kA2Full, kA2Empty, kD2Full0, kD2Full1,
       kD2Empty0, kD2Empty1, kNumBars };

__global__ void __launch_bounds__(kThreads, 1) tc_fft_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffMisc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffMisc + 192);
  auto bar = [&](int i) { return smem_u32(&bars[i]); };

  // ---- one-time setup -----------------------------------------------------------------------
  for (int i = tid; i < (2 * kB1Half + 2 * kB2Half) / 16; i += kThreads) {
    const float4* src = i < 2 * kB1Half / 16 ? reinterpret_cast<const float4*>(p.b1) + i
                                             : reinterpret_cast<const float4*>(p.b2) + (i - 2 * kB1Half / 16);
    reinterpret_cast<float4*>(smem + kOffB1)[i] = *src;
  }
  for (int i = tid; i < (kOffWave) / 16; i += kThreads) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    mbar_init(bar(kWaveFull0), 1);
    mbar_init(bar(kWaveFull1), 1);
    mbar_init(bar(kA1Full), 4);    // one lane of each marshal warp
    mbar_init(bar(kA1Empty), 1);   // tcgen05.commit
    mbar_init(bar(kD1Full0), 1);
    mbar_init(bar(kD1Full1), 1);
    mbar_init(bar(kD1Empty0), 8);  // one lane of each twiddle warp
    mbar_init(bar(kD1Empty1), 8);
    mbar_init(bar(kA2Full), 8);
    mbar_init(bar(kA2Empty), 1);
    mbar_init(bar(kD2Full0), 1);
    mbar_init(bar(kD2Full1), 1);
    mbar_init(bar(kD2Empty0), 4);  // one lane of each magnitude warp
    mbar_init(bar(kD2Empty1), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // 256 TMEM columns: D1 2 buffers x (2 x 32), D2 2 buffers x 64
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_lane = (uint32_t)((warp & 3) * 32) << 16;  // this warp's TMEM lane quarter
  const uint32_t a1_s = smem_u32(smem + kOffA1), a2_s = smem_u32(smem + kOffA2);
  const uint32_t b1_s = smem_u32(smem + kOffB1), b2_s = smem_u32(smem + kOffB2);
  const int first = blockIdx.x, stride = gridDim.x;

  if (warp < 4) {
    // =========================== marshal warps: waveform tile -> A1 {hi, lo} =======================
    float4 win[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) win[i] = reinterpret_cast<const float4*>(p.window)[i * 32 + lane];
    auto load_tile = [&](int tile, int b) {
      mbar_expect_tx(bar(kWaveFull0 + b), kWaveBytes);
      bulk_copy_g2s(smem_u32(smem + kOffWave + b * kWaveBytes), p.wav + (size_t)tile * kTF * kHop, kWaveBytes, bar(kWaveFull0 + b));
    };
    if (tid == 0) {
      if (first < p.n_tiles) load_tile(first, 0);
      if (first + stride < p.n_tiles) load_tile(first + stride, 1);
    }
    int it = 0;
    for (int tile = first; tile < p.n_tiles; tile += stride, ++it) {
      const int b = it & 1;
      const float* wave = reinterpret_cast<const float*>(smem + kOffWave + b * kWaveBytes);
      const long long c0 = clock64();
      mbar_wait(bar(kWaveFull0 + b), (it >> 1) & 1);
      const long long c1 = clock64();
      mbar_wait(bar(kA1Empty), (it & 1) ^ 1);  // stage-1 MMAs of the previous tile have read A1
      const long long c2 = clock64();
#pragma unroll 1
      for (int f = warp; f < kTF; f += 4) {
        const float4* src = reinterpret_cast<const float4*>(wave + f * kHop) + lane;
        unsigned char* hi_row = smem + kOffA1 + f * kA1Frame;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 x = src[i * 32];
          const int n1 = 4 * i + (lane >> 3);
          const int off = n1 * 128 + (((((lane & 7) >> 1) ^ (n1 & 3)) << 5) | ((lane & 1) << 4));
          float4 h, l;
          split(x.x * win[i].x, h.x, l.x);
          split(x.y * win[i].y, h.y, l.y);
          split(x.z * win[i].z, h.z, l.z);
          split(x.w * win[i].w, h.w, l.w);
          *reinterpret_cast<float4*>(hi_row + off) = h;
          *reinterpret_cast<float4*>(hi_row + kA1Half + off) = l;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kA1Full));
      marshal_sync();  // all four warps are done with wave[b]: refill it with the tile after next
      if (tid == 0 && tile + 2 * stride < p.n_tiles) load_tile(tile + 2 * stride, b);
      if (p.stats && blockIdx.x == 0 && tid == 0) {
        p.stats[0] += c1 - c0; p.stats[1] += c2 - c1; p.stats[2] += clock64() - c2; p.stats[3] += 1;
      }
    }
  } else if (warp == kWarpMma) {
    // =========================== MMA issue: one warp, one elected lane per instruction ==============
    {
      constexpr uint32_t kIdesc1 = make_idesc(128, 32, /*a MN-major*/ 1, /*b K-major*/ 0);
      constexpr uint32_t kIdesc2 = make_idesc(128, 64, 0, 0);
      // descriptors of the first K step of each operand copy; the others differ by a constant in the address field
      const uint64_t a1_hi = make_desc(a1_s, kA1Frame, 512, kSwizzle128Base32), a1_lo = make_desc(a1_s + kA1Half, kA1Frame, 512, kSwizzle128Base32);
      const uint64_t b1_hi = make_desc(b1_s, 16, 1024), b1_lo = make_desc(b1_s + kB1Half, 16, 1024);
      const uint64_t a2_hi = make_desc(a2_s, 16, 1024), a2_lo = make_desc(a2_s + kA2Half, 16, 1024);
      const uint64_t b2_hi = make_desc(b2_s, 16, 1024), b2_lo = make_desc(b2_s + kB2Half, 16, 1024);
      auto stage1 = [&](int it) {
        const int buf = it & 1;
        const long long c0 = clock64();
        mbar_wait(bar(kA1Full), it & 1);
        const long long c1 = clock64();
        mbar_wait(bar(kD1Empty0 + buf), ((it >> 1) & 1) ^ 1);
        if (p.stats && blockIdx.x == 0 && lane == 0) { p.stats[4] += c1 - c0; p.stats[5] += clock64() - c1; }
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t d = tmem_base + buf * 64 + g * 32;
          const uint64_t ga = (uint64_t)((g * 4 * kA1Frame) >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // (a address + ks * 1024 B, b address + ks * 32 B), in 16-byte units
            const uint64_t oa = ga + (uint64_t)(ks * 64), ob = (uint64_t)(ks * 2);
            if (ks == 0) umma_tf32_warp<0>(d, a1_hi + oa, b1_hi + ob, kIdesc1);
            else umma_tf32_warp<1>(d, a1_hi + oa, b1_hi + ob, kIdesc1);
            umma_tf32_warp<1>(d, a1_lo + oa, b1_hi + ob, kIdesc1);
            umma_tf32_warp<1>(d, a1_hi + oa, b1_lo + ob, kIdesc1);
          }
        }
        umma_commit_warp(bar(kA1Empty));
        umma_commit_warp(bar(kD1Full0 + buf));
      };
      auto stage2 = [&](int it) {
        const int buf = it & 1;
        const long long c0 = clock64();
        mbar_wait(bar(kA2Full), it & 1);
        const long long c1 = clock64();
        mbar_wait(bar(kD2Empty0 + buf), ((it >> 1) & 1) ^ 1);
        if (p.stats && blockIdx.x == 0 && lane == 0) { p.stats[6] += c1 - c0; p.stats[7] += clock64() - c1; }
        tc_fence_after();
        const uint32_t d = tmem_base + 128 + buf * 64;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t oa = (uint64_t)(((ks >> 2) * kA2Block + (ks & 3) * 32) >> 4);
          const uint64_t ob = (uint64_t)(((ks >> 2) * kB2Block + (ks & 3) * 32) >> 4);
          if (ks == 0) umma_tf32_warp<0>(d, a2_hi + oa, b2_hi + ob, kIdesc2);
          else umma_tf32_warp<1>(d, a2_hi + oa, b2_hi + ob, kIdesc2);
          umma_tf32_warp<1>(d, a2_lo + oa, b2_hi + ob, kIdesc2);
          umma_tf32_warp<1>(d, a2_hi + oa, b2_lo + ob, kIdesc2);
        }
        umma_commit_warp(bar(kA2Empty));
        umma_commit_warp(bar(kD2Full0 + buf));
      };
      int n_it = 0;
      for (int tile = first; tile < p.n_tiles; tile += stride) ++n_it;
      if (n_it > 0) stage1(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) stage1(it + 1);  // the next tile's first stage runs while the twiddle warps work on this one
        stage2(it);
      }
    }
  } else if (warp >= kWarpTw && warp < kWarpMag) {
    // =========================== twiddle warps: D1 -> A2 {hi, lo} =================================
    const int g = (warp - kWarpTw) >> 2, f = 4 * g + (warp & 3);  // frame slot of this warp (slot 7 is padding)
    float tw_c[17], tw_s[17];  // W_1024^{n2 k1} = c - i s, n2 = lane
#pragma unroll
    for (int k1 = 1; k1 <= 16; ++k1) {
      float s, c;
      sincospif((float)((lane * k1) & 1023) / 512.0f, &s, &c);
      tw_c[k1] = c;
      tw_s[k1] = s;
    }
    unsigned char* blk = smem + kOffA2 + (lane >> 4) * kA2Block + ((lane & 1) << 3);
    const int chunk = (lane & 15) >> 1;
    int it = 0;
    for (int tile = first; tile < p.n_tiles; tile += stride, ++it) {
      const int buf = it & 1;
      const long long c0 = clock64();
      mbar_wait(bar(kD1Full0 + buf), (it >> 1) & 1);
      const long long c1 = clock64();
      tc_fence_after();
      float y[32];
      tmem_ld32(tmem_base + tm_lane + buf * 64 + g * 32, y);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kD1Empty0 + buf));
      float2 zh[17], zl[17];
#pragma unroll
      for (int k1 = 0; k1 <= 16; ++k1) {
        float zr, zi;
        if (k1 == 0) {
          zr = y[0];
          zi = 0.f;
        } else if (k1 == 16) {
          zr = tw_c[16] * y[1];
          zi = -tw_s[16] * y[1];
        } else {
          const float a = y[1 + k1], b = y[16 + k1];
          zr = fmaf(tw_s[k1], b, tw_c[k1] * a);
          zi = fmaf(-tw_s[k1], a, tw_c[k1] * b);
        }
        split(zr, zh[k1].x, zl[k1].x);
        split(zi, zh[k1].y, zl[k1].y);
      }
      const long long c2 = clock64();
      mbar_wait(bar(kA2Empty), (it & 1) ^ 1);  // stage-2 MMAs of the previous tile have read A2
      const long long c3 = clock64();
      if (f < kTF) {
#pragma unroll
        for (int k1 = 0; k1 <= 16; ++k1) {
          const int row = f * 17 + k1;
          const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
          *reinterpret_cast<float2*>(blk + off) = zh[k1];
          *reinterpret_cast<float2*>(blk + kA2Half + off) = zl[k1];
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kA2Full));
      if (p.stats && blockIdx.x == 0 && warp == kWarpTw && lane == 0) {
        p.stats[8] += c1 - c0; p.stats[9] += c2 - c1; p.stats[10] += c3 - c2; p.stats[11] += clock64() - c3;
      }
    }
  } else if (warp >= kWarpMag) {
    // =========================== magnitude warps: D2 -> |X| ======================================
    const int row = (warp & 3) * 32 + lane;
    const int f = row / 17, k1 = row - f * 17;
    const bool mirrored = k1 != 0 && k1 != 16;
    float keep = 0.f;
    int it = 0;
    for (int tile = first; tile < p.n_tiles; tile += stride, ++it) {
      const int buf = it & 1;
      const long long t = (long long)tile * kTF + f;
      const long long c0 = clock64();
      mbar_wait(bar(kD2Full0 + buf), (it >> 1) & 1);
      const long long c1 = clock64();
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float x[32];
        tmem_ld32(tmem_base + tm_lane + 128 + buf * 64 + half * 32, x);
        tmem_wait_ld();
        if (half == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kD2Empty0 + buf));
        }
        if (row < kTF * 17 && t < p.n_frames) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k2 = half * 16 + j;
            float m;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(m) : "f"(fmaf(x[2 * j], x[2 * j], fmaf(x[2 * j + 1], x[2 * j + 1], 1e-9f))));
            const int k = k1 + 32 * k2;
            const int bin = k <= 512 ? k : 1024 - k;
            if (k <= 512 || mirrored) {
              if (p.mags) p.mags[t * kBins + bin] = m;
              else keep += m;
            }
          }
        }
      }
      if (p.stats && blockIdx.x == 0 && warp == kWarpMag && lane == 0) { p.stats[12] += c1 - c0; p.stats[13] += clock64() - c1; }
    }
    if (!p.mags) atomicAdd(p.checksum + blockIdx.x, keep);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

// ---- host: operand images, reference, driver ------------------------------------------------------
static void split_host(float x, float& hi, float& lo) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&hi, &u, 4);
  lo = x - hi;
}
// K-major, 128-byte swizzled image of an (n_rows x k_total) matrix, K split into blocks of 32 floats
static void put_kmajor(std::vector<float>& img, size_t base_floats, int n_rows, int row, int k, float v) {
  const int blk = k / 32, kk = k % 32;
  const size_t off_bytes = (size_t)blk * n_rows * 128 + (size_t)row * 128 + (((kk / 4) ^ (row % 8)) << 4) + (kk % 4) * 4;
  img[base_floats + off_bytes / 4] = v;
}

int main(int argc, char** argv) {
  const double two_pi = 6.283185307179586476925286766559;
  // B1[j][n1]: j = 0 -> Y0, 1 -> Y16, 1 + k -> Re Y_k, 16 + k -> Im Y_k
  std::vector<float> b1(2 * kB1Half / 4, 0.f), b2(2 * kB2Half / 4, 0.f);
  for (int j = 0; j < 32; ++j)
    for (int n1 = 0; n1 < 32; ++n1) {
      double v;
      if (j == 0) v = 1.0;
      else if (j == 1) v = (n1 & 1) ? -1.0 : 1.0;
      else if (j <= 16) v = cos(two_pi * ((n1 * (j - 1)) % 32) / 32.0);
      else v = -sin(two_pi * ((n1 * (j - 16)) % 32) / 32.0);
      float hi, lo;
      split_host((float)v, hi, lo);
      lo = (float)(v - (double)hi);  // the remainder of the exact value, not of its float rounding
      put_kmajor(b1, 0, 32, j, n1, hi);
      put_kmajor(b1, kB1Half / 4, 32, j, n1, lo);
    }
  // B2[(k2, c)][(n2, c')]: out_re = cos*Zr + sin*Zi, out_im = -sin*Zr + cos*Zi, angle 2 pi n2 k2 / 32
  for (int k2 = 0; k2 < 32; ++k2)
    for (int n2 = 0; n2 < 32; ++n2) {
      const double a = two_pi * ((n2 * k2) % 32) / 32.0, c = cos(a), s = sin(a);
      const double vals[2][2] = {{c, s}, {-s, c}};  // [out c][in c']
      for (int co = 0; co < 2; ++co)
        for (int ci = 0; ci < 2; ++ci) {
          float hi, lo;
          split_host((float)vals[co][ci], hi, lo);
          lo = (float)(vals[co][ci] - (double)hi);
          put_kmajor(b2, 0, 64, 2 * k2 + co, 2 * n2 + ci, hi);
          put_kmajor(b2, kB2Half / 4, 64, 2 * k2 + co, 2 * n2 + ci, lo);
        }
    }
  std::vector<float> window(kNfft);
  for (int n = 0; n < kNfft; ++n) window[n] = (float)(0.5 - 0.5 * cos(two_pi * n / kNfft));

  float *d_b1, *d_b2, *d_win;
  CK(cudaMalloc(&d_b1, b1.size() * 4));
  CK(cudaMalloc(&d_b2, b2.size() * 4));
  CK(cudaMalloc(&d_win, kNfft * 4));
  CK(cudaMemcpy(d_b1, b1.data(), b1.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b2, b2.data(), b2.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_win, window.data(), kNfft * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(tc_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

  // ---- accuracy: 300 frames of three signal kinds against a float64 DFT --------------------------
  {
    const int T = 300, N = (T - 1) * kHop + kNfft + kTF * kHop;
    std::vector<float> wav(N);
    srand(11);
    auto rnd = []() { return rand() / (double)RAND_MAX * 2.0 - 1.0; };
    double lp = 0;
    for (int i = 0; i < N; ++i) {
      const int seg = i / (N / 3);
      if (seg == 0) wav[i] = (float)(0.5 * rnd());                                   // white noise
      else if (seg == 1) { lp = 0.97 * lp + 0.03 * rnd(); wav[i] = (float)(8.0 * lp); }  // low-pass (speech-like slope)
      else wav[i] = (float)(0.9 * sin(two_pi * 440.0 * i / 24000.0) + 1e-4 * rnd());  // loud tone over a quiet floor
    }
    float *d_wav, *d_mags;
    CK(cudaMalloc(&d_wav, N * 4));
    CK(cudaMalloc(&d_mags, (size_t)T * kBins * 4));
    CK(cudaMemcpy(d_wav, wav.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_mags, 0xff, (size_t)T * kBins * 4));
    float* d_dbg;
    CK(cudaMalloc(&d_dbg, 2 * 128 * 64 * 4));
    CK(cudaMemset(d_dbg, 0, 2 * 128 * 64 * 4));
    Params p{d_wav, d_win, d_b1, d_b2, T, (T + kTF - 1) / kTF, d_mags, nullptr, nullptr};
    tc_fft_kernel<<<8, kThreads, kSmemBytes>>>(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> mags((size_t)T * kBins);
    CK(cudaMemcpy(mags.data(), d_mags, mags.size() * 4, cudaMemcpyDeviceToHost));
    if (argc > 1) {  // debug: stage outputs of tile 0 against the host
      std::vector<float> dbg(2 * 128 * 64);
      CK(cudaMemcpy(dbg.data(), d_dbg, dbg.size() * 4, cudaMemcpyDeviceToHost));
      for (int f = 0; f < 2; ++f)
        for (int n2 = 0; n2 < 3; ++n2) {
          double y0 = 0, y16 = 0, re1 = 0, im1 = 0;
          for (int n1 = 0; n1 < 32; ++n1) {
            const double x = (double)((float)(wav[f * kHop + 32 * n1 + n2] * window[32 * n1 + n2]));
            y0 += x; y16 += (n1 & 1) ? -x : x;
            re1 += x * cos(two_pi * n1 / 32.0); im1 -= x * sin(two_pi * n1 / 32.0);
          }
          const float* d = &dbg[(f * 32 + n2) * 64];
          printf("D1 f=%d n2=%d  got %.6f %.6f %.6f %.6f   want %.6f %.6f %.6f %.6f\n", f, n2, d[0], d[1], d[2], d[17], y0, y16, re1, im1);
        }
      for (int r = 0; r < 4; ++r) {
        const float* d = &dbg[128 * 64 + r * 64];
        printf("D2 row %d: %.5f %.5f %.5f %.5f %.5f %.5f\n", r, d[0], d[1], d[2], d[3], d[4], d[5]);
      }
      printf("mags f0: %.5f %.5f %.5f %.5f\n", mags[0], mags[1], mags[32], mags[33]);
    }
    double worst_rel[3] = {0, 0, 0}, worst_peak[3] = {0, 0, 0};
    int bad = 0;
    for (int t = 0; t < T; t += 3) {
      const int seg = std::min(2, (t * kHop + 512) / (N / 3));
      std::vector<double> ref(kBins);
      double peak = 0;
      for (int k = 0; k < kBins; ++k) {
        double re = 0, im = 0;
        for (int n = 0; n < kNfft; ++n) {
          const double x = (double)((float)(wav[t * kHop + n] * window[n]));  // the fp32 product the kernel forms
          const double a = -two_pi * (double)((long long)k * n % kNfft) / kNfft;
          re += x * cos(a);
          im += x * sin(a);
        }
        ref[k] = sqrt(re * re + im * im + 1e-9);
        peak = std::max(peak, ref[k]);
      }
      for (int k = 0; k < kBins; ++k) {
        const double got = mags[(size_t)t * kBins + k];
        if (!(got == got)) { ++bad; continue; }
        const double e = fabs(got - ref[k]);
        worst_rel[seg] = std::max(worst_rel[seg], e / std::max(ref[k], 1e-3));
        worst_peak[seg] = std::max(worst_peak[seg], e / peak);
      }
    }
    const char* names[3] = {"white noise", "low-pass noise", "tone + 1e-4 floor"};
    for (int s = 0; s < 3; ++s)
      printf("accuracy %-18s max |err|/max(ref,1e-3) %.3e   max |err|/frame peak %.3e\n", names[s], worst_rel[s], worst_peak[s]);
    printf("NaN or unwritten magnitudes: %d\n", bad);
    cudaFree(d_wav);
    cudaFree(d_mags);
  }

  // ---- timing: the frame count of BASELINE configs[1] (64 x 937 frames) ---------------------------
  {
    const int T = 59968, N = (T - 1) * kHop + kNfft + kTF * kHop;
    std::vector<float> wav(N);
    for (int i = 0; i < N; ++i) wav[i] = (float)(0.3 * sin(i * 0.01) + 0.1 * ((i * 2654435761u >> 8) & 0xffff) / 65536.0);
    float *d_wav, *d_sum;
    CK(cudaMalloc(&d_wav, (size_t)N * 4));
    CK(cudaMalloc(&d_sum, sms * 4));
    CK(cudaMemcpy(d_wav, wav.data(), (size_t)N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_sum, 0, sms * 4));
    long long* d_stats;
    CK(cudaMalloc(&d_stats, 16 * 8));
    CK(cudaMemset(d_stats, 0, 16 * 8));
    Params p{d_wav, d_win, d_b1, d_b2, T, (T + kTF - 1) / kTF, nullptr, d_sum, nullptr};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 5; ++i) tc_fft_kernel<<<sms, kThreads, kSmemBytes>>>(p);
    CK(cudaDeviceSynchronize());
    const int reps = 50;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) tc_fft_kernel<<<sms, kThreads, kSmemBytes>>>(p);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    {
      Params ps = p;
      ps.stats = d_stats;
      tc_fft_kernel<<<sms, kThreads, kSmemBytes>>>(ps);
      CK(cudaDeviceSynchronize());
      long long st[16];
      CK(cudaMemcpy(st, d_stats, sizeof(st), cudaMemcpyDeviceToHost));
      const double n = (double)st[3];
      printf("per tile, CTA 0 (cycles): marshal wait-wave %.0f wait-A1-free %.0f work %.0f | mma wait-A1 %.0f wait-D1-free %.0f wait-A2 %.0f wait-D2-free %.0f | "
             "twiddle wait-D1 %.0f ld+compute %.0f wait-A2-free %.0f store %.0f | mag wait-D2 %.0f work %.0f   (%.0f tiles)\n",
             st[0] / n, st[1] / n, st[2] / n, st[4] / n, st[5] / n, st[6] / n, st[7] / n, st[8] / n, st[9] / n, st[10] / n, st[11] / n,
             st[12] / n, st[13] / n, n);
    }
    printf("timing: %d frames (windowed STFT magnitudes, no mel) %.2f us per launch, %.1f ns per frame per SM\n", T,
           1e3 * ms / reps, 1e6 * ms / reps / ((double)T / sms));
  }
  return 0;
}
