// Exact unsigned division of e < 2^31 by a fixed d via one multiply-high and a shift
// (round-up method: mul = floor(2^(31+s)/d) + 1 with s = ceil(log2 d)).  The flat-indexed
// quantiser kernels need (e / T) and (row % M) per 4 elements, and the fused kernel needs
// (tile / tiles_per_row) and (length / hop) per tile; the hardware divide sequence is ~30
// instructions each.
#pragma once
#include <cuda_runtime.h>

namespace dmel {

struct FastDiv {
  unsigned mul, shift, d;
  __host__ static FastDiv make(unsigned d) {
    FastDiv f;
    f.d = d;
    if (d <= 1) {
      f.mul = 0;
      f.shift = 0;
      return f;
    }
    unsigned s = 0;
    while ((1ull << s) < d) ++s;
    f.mul = (unsigned)(((1ull << (31 + s)) / d) + 1);
    f.shift = s - 1;
    return f;
  }
  __device__ __forceinline__ unsigned div(unsigned e) const { return d <= 1 ? e : (__umulhi(e, mul) >> shift); }
};

}  // namespace dmel
