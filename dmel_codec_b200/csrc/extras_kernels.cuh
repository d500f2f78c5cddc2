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
constexpr int kAbsmaxMinSpan = 16384;  // a block is given at least this many samples (when the row has them)

__device__ __forceinline__ float absmax4(float m, float4 x) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w)));
}

// grid (blocks_per_row, n_rows): the blocks of a row stride over it together, four 16-byte loads per thread in flight.
// The host sizes blocks_per_row so that the grid is about eight blocks per SM: on this part one short-lived block per
// 64 KB chunk reads at 4.6 TB/s, the same loop in a persistent grid at 6.6 TB/s (benchmarks/stream_read.cu).
// |x| >= 0, so its bit pattern orders like an unsigned integer.
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
  if (n <= 0) return;
  const float* src = wav + base;
  // [0, head) scalar up to the first 16-byte boundary, n4 vectors, then a scalar tail: ragged rows start anywhere
  const int head = min(n, (int)((4 - ((reinterpret_cast<uintptr_t>(src) >> 2) & 3)) & 3));
  const int n4 = (n - head) >> 2;
  const float4* q = reinterpret_cast<const float4*>(src + head);
  const int stride = gridDim.x * kAbsmaxThreads;
  int i = blockIdx.x * kAbsmaxThreads + threadIdx.x;
  float m = 0.f;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 a = __ldg(q + i), b = __ldg(q + i + stride), c = __ldg(q + i + 2 * stride), d = __ldg(q + i + 3 * stride);
    m = absmax4(absmax4(absmax4(absmax4(m, a), b), c), d);
  }
  for (; i < n4; i += stride) m = absmax4(m, __ldg(q + i));
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < head) m = fmaxf(m, fabsf(__ldg(src + threadIdx.x)));
    const int tail = head + 4 * n4 + (int)threadIdx.x;
    if (tail < n) m = fmaxf(m, fabsf(__ldg(src + tail)));
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
