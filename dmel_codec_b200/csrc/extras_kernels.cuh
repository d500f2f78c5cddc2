// Input-side helpers of the dMel path: per-utterance peak normalisation.
//
// The reference normalises every utterance on the host before it reaches the model:
// `audio = librosa.util.normalize(audio) * 0.95` (reference dataset/lhotse_tts_dataset.py:29-33), i.e.
// x / max|x| * 0.95 with max|x| < tiny treated as 1.  Here one HBM-bound pass finds max|x| of every row over its
// valid samples and a second, tiny launch turns it into a gain; the fused kernel applies the gain to the tile it
// has just staged, so the normalised waveform never exists in HBM.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace dmel {

constexpr int kAbsmaxThreads = 256;
constexpr int kAbsmaxChunk = 16384;  // samples per block

// grid (ceil(n_samples / kAbsmaxChunk), n_rows).  |x| >= 0, so its bit pattern orders like an unsigned integer.
__global__ void __launch_bounds__(kAbsmaxThreads) row_absmax_kernel(
    const float* __restrict__ wav, const long long* __restrict__ offsets, const int* __restrict__ lengths,
    long long row_stride, int n_samples, unsigned* __restrict__ absmax_bits) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int row = blockIdx.y;
  long long base = (long long)row * row_stride;
  int n = n_samples;
  if (offsets) {
    base = offsets[row];
    n = (int)(offsets[row + 1] - base);
  }
  if (lengths) n = min(n, max(lengths[row], 0));
  const int begin = blockIdx.x * kAbsmaxChunk, end = min(begin + kAbsmaxChunk, n);
  if (begin >= end) return;
  const float* src = wav + base;
  float m = 0.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(src + begin) & 15) == 0);
  int i = begin + threadIdx.x * 4;
  if (vec) {
    for (; i + 3 < end; i += kAbsmaxThreads * 4) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(src + i));
      m = fmaxf(fmaxf(m, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w)));
    }
    for (int j = i; j < end && j < i + 4; ++j) m = fmaxf(m, fabsf(__ldg(src + j)));  // the chunk's ragged end
  } else {
    for (int j = begin + threadIdx.x; j < end; j += kAbsmaxThreads) m = fmaxf(m, fabsf(__ldg(src + j)));
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(absmax_bits + row, __float_as_uint(m));
}

// gain[b] = target / max|x| (librosa.util.normalize: a peak below the smallest normal float counts as 1)
__global__ void row_gain_kernel(const unsigned* absmax_bits, float target, float* gain, int n_rows) {  // (the two may alias)
  grid_dependency_wait();
  grid_launch_dependents();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_rows) return;
  const float m = __uint_as_float(absmax_bits[b]);
  gain[b] = __fdiv_rn(target, m < FLT_MIN ? 1.f : m);
}

}  // namespace dmel
