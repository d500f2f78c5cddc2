// Finite scalar quantisation to indices, and back: the elementwise core of the reference's learned quantiser
// (GroupedResidualFSQ of vector_quantize_pytorch, called at reference models/modules/dowmsample_fsq.py:95 and
// :130-137), fused with the language model's id_shift (models/modules/lm_process_input.py:301-313).
//
//   bounded = tanh(z + shift_d) * half_l_d - offset_d        code_d = rint(bounded) / (L_d / 2)
//   index   = sum_d (rint(bounded_d) + L_d / 2) * basis_d    basis = {1, L_0, L_0 L_1, ...}
//
// Input zp is (B, T, G, D) float32 (per-group latents after the learned project_in, D = number of levels).
// Outputs, each optional: codes (B, T, G, D) float32, indices (B, G, T) int64 (the layout encode() returns),
// lm_ids (B, T, G) int64 = index + g * codebook_size (what id_shift hands the language model).
// HBM-bound: 4 D bytes in, up to 4 D + 16 bytes out per (b, t, g); one CTA stages 64 time steps x all groups so that
// both the (T, G) -major reads and the (G, T) -major index writes are contiguous.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dmel {

constexpr int kFsqMaxDims = 8;
constexpr int kFsqTileT = 256;  // time steps per CTA (the host lowers it when G is large: the tile's indices live in shared memory)
constexpr int kFsqThreads = 256;

struct FsqLevels {
  int n_dims;
  float half_l[kFsqMaxDims], offset[kFsqMaxDims], shift[kFsqMaxDims];
  int half_width[kFsqMaxDims], basis[kFsqMaxDims], level[kFsqMaxDims];
};

// tanh(x) = 1 - 2 / (1 + e^{2x}) on the two MUFU approximations: absolute error below 2e-7 over the whole line (the
// relative error of ex2 / rcp is damped by 2e / (1 + e)^2 <= 1/2), the same size as tanhf's own last-bit error and
// 50x under the distance at which the parity tests stop comparing levels; 5 instructions where tanhf takes about 20.
// Only rint(tanh * half_l - offset) leaves the kernel, so absolute error is what matters.
__device__ __forceinline__ float fsq_tanh(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));  // e^{2x}; +inf and 0 saturate correctly
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(-2.0f, r, 1.0f);
}

// grid: (ceil(T / tile_t), B); dynamic smem: tile_t * G * 4 bytes (the tile's indices)
// D = number of levels, a template parameter: with a run-time D the compiler predicates the code of all eight
// dimensions and every (t, g) pays the issue slots of the five that a (7, 5, 5) codebook does not have.
template <int D>
__global__ void __launch_bounds__(kFsqThreads) fsq_encode_kernel(const float* __restrict__ zp, int n_t, int n_groups, FsqLevels lv,
                                                                 float* __restrict__ codes, long long* __restrict__ indices,
                                                                 long long* __restrict__ lm_ids, int codebook_size, int tile_t) {
  extern __shared__ int s_index[];  // [t in tile][g]
  grid_dependency_wait();
  grid_launch_dependents();
  const int b = blockIdx.y, t0 = blockIdx.x * tile_t;
  const int nt = min(tile_t, n_t - t0);
  constexpr int d = D;
  const size_t base = ((size_t)b * n_t + t0) * n_groups;  // first (t, g) pair of the tile
  const float* ztile = zp + base * d;
  float* ctile = codes ? codes + base * d : nullptr;
  // i = t * G + g walks the tile in memory order; g follows i incrementally (no division in the loop), and rounding
  // goes through the add-a-magic-constant trick: the division sequence and F2I / FRND all issue to the XU pipe, which
  // the two MUFUs of every tanh need (the first version of this kernel ran that pipe at 100 %)
  constexpr float kMagic = 12582912.f;  // 1.5 * 2^23
  const int g_step = kFsqThreads % n_groups;
  int g = threadIdx.x % n_groups;
  for (int i = threadIdx.x; i < nt * n_groups; i += kFsqThreads) {
    const float* z = ztile + i * d;
    int index = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float f = (fsq_tanh(z[k] + lv.shift[k]) * lv.half_l[k] - lv.offset[k]) + kMagic;  // round to nearest even, as rintf
      const int iq = __float_as_int(f) - __float_as_int(kMagic);
      if (ctile) ctile[i * d + k] = (f - kMagic) / (float)lv.half_width[k];
      index += (iq + lv.half_width[k]) * lv.basis[k];
    }
    s_index[i] = index;
    if (lm_ids) lm_ids[base + i] = (long long)index + (long long)g * codebook_size;
    g += g_step;
    g = g >= n_groups ? g - n_groups : g;
  }
  if (indices == nullptr) return;
  __syncthreads();
  for (int gg = threadIdx.x >> 5; gg < n_groups; gg += kFsqThreads / 32) {  // one warp per group row: contiguous in indices
    long long* dst = indices + ((size_t)b * n_groups + gg) * n_t + t0;
    for (int t = threadIdx.x & 31; t < nt; t += 32) dst[t] = (long long)s_index[t * n_groups + gg];
  }
}

// indices (B, G, T) int64 -> codes (B, T, G, D) float32 (FSQ.indices_to_codes without the learned project_out)
template <int D>
__global__ void __launch_bounds__(kFsqThreads) fsq_decode_kernel(const long long* __restrict__ indices, int n_t, int n_groups, FsqLevels lv,
                                                                 float* __restrict__ codes, int tile_t) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int b = blockIdx.y, t0 = blockIdx.x * tile_t;
  const int nt = min(tile_t, n_t - t0);
  constexpr int d = D;
  for (int i = threadIdx.x; i < nt * n_groups; i += kFsqThreads) {
    const int t = i / n_groups, g = i - t * n_groups;
    unsigned index = (unsigned)indices[((size_t)b * n_groups + g) * n_t + t0 + t];  // < prod(levels) < 2^31 (checked on the host)
    float* out = codes + (((size_t)b * n_t + t0 + t) * n_groups + g) * d;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const unsigned lk = (unsigned)lv.level[k];
      const int digit = (int)(index % lk);
      index /= lk;
      out[k] = (float)(digit - lv.half_width[k]) / (float)lv.half_width[k];
    }
  }
}

}  // namespace dmel
