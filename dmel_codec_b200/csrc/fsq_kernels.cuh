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

// grid: (ceil(T / tile_t), B); dynamic smem: tile_t * G * 8 bytes (the tile's indices)
__global__ void __launch_bounds__(kFsqThreads) fsq_encode_kernel(const float* __restrict__ zp, int n_t, int n_groups, FsqLevels lv,
                                                                 float* __restrict__ codes, long long* __restrict__ indices,
                                                                 long long* __restrict__ lm_ids, int codebook_size, int tile_t) {
  extern __shared__ long long s_index[];  // [t in tile][g]
  grid_dependency_wait();
  grid_launch_dependents();
  const int b = blockIdx.y, t0 = blockIdx.x * tile_t;
  const int nt = min(tile_t, n_t - t0);
  const int d = lv.n_dims;
  const size_t base = ((size_t)b * n_t + t0) * n_groups;  // first (t, g) pair of the tile
  for (int i = threadIdx.x; i < nt * n_groups; i += kFsqThreads) {  // i = t * G + g: contiguous in zp
    const float* z = zp + (base + i) * d;
    int index = 0;
#pragma unroll
    for (int k = 0; k < kFsqMaxDims; ++k) {
      if (k < d) {
        const float q = rintf(tanhf(z[k] + lv.shift[k]) * lv.half_l[k] - lv.offset[k]);
        if (codes) codes[(base + i) * d + k] = q / (float)lv.half_width[k];
        index += ((int)q + lv.half_width[k]) * lv.basis[k];
      }
    }
    s_index[i] = index;
    if (lm_ids) lm_ids[base + i] = (long long)index + (long long)(i % n_groups) * codebook_size;
  }
  if (indices == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.x; i < nt * n_groups; i += kFsqThreads) {  // i = g * nt + t: contiguous in indices
    const int g = i / nt, t = i - g * nt;
    indices[((size_t)b * n_groups + g) * n_t + t0 + t] = s_index[t * n_groups + g];
  }
}

// indices (B, G, T) int64 -> codes (B, T, G, D) float32 (FSQ.indices_to_codes without the learned project_out)
__global__ void __launch_bounds__(kFsqThreads) fsq_decode_kernel(const long long* __restrict__ indices, int n_t, int n_groups, FsqLevels lv,
                                                                 float* __restrict__ codes, int tile_t) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int b = blockIdx.y, t0 = blockIdx.x * tile_t;
  const int nt = min(tile_t, n_t - t0);
  const int d = lv.n_dims;
  for (int i = threadIdx.x; i < nt * n_groups; i += kFsqThreads) {
    const int t = i / n_groups, g = i - t * n_groups;
    long long index = indices[((size_t)b * n_groups + g) * n_t + t0 + t];
    float* out = codes + (((size_t)b * n_t + t0 + t) * n_groups + g) * d;
#pragma unroll
    for (int k = 0; k < kFsqMaxDims; ++k) {
      if (k < d) {
        const int digit = (int)(index % lv.level[k]);
        index /= lv.level[k];
        out[k] = (float)(digit - lv.half_width[k]) / (float)lv.half_width[k];
      }
    }
  }
}

}  // namespace dmel
