// Streaming kernels of the dMel quantiser that work on an existing
// (B, n_mels, T) tensor: quantise, dequantise, min/max.  All HBM-bound; each
// warp instruction touches contiguous bytes (128 B of codes / 512 B of floats).
// The quantiser itself is not in the reference; the spec is SURVEY.md Appendix B.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fastdiv.cuh"

namespace dmel {

constexpr int kStreamThreads = 256;
constexpr int kStreamUnroll = 4;  // groups of 4 elements per thread per trip

// ---------------------------------------------------------------------------
// codes (uint8) -> bin centres (float32) by table lookup.  The table is built
// by the host with the oracle's exact op order, so the result is bit exact by
// construction.  Flat indexing: element e belongs to channel (e / T) % M.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kStreamThreads) dequantize_kernel(
    const uint8_t* __restrict__ codes, float* __restrict__ out, const float* __restrict__ table,
    unsigned n_elems, FastDiv by_frames, FastDiv by_mels, unsigned n_bins, bool vec_ok) {
  grid_dependency_wait();   // programmatic dependent launch: the inputs may come from the previous kernel
  grid_launch_dependents();
  const unsigned n_frames = by_frames.d, n_mels = by_mels.d;
  const unsigned groups = n_elems >> 2;
  const unsigned stride = gridDim.x * kStreamThreads;
  const unsigned kmax = n_bins - 1;
  if (vec_ok) {
    for (unsigned g0 = blockIdx.x * kStreamThreads + threadIdx.x; g0 < groups; g0 += stride * kStreamUnroll) {
      uchar4 c[kStreamUnroll];
#pragma unroll
      for (int u = 0; u < kStreamUnroll; ++u) {
        const unsigned g = g0 + u * stride;
        c[u] = g < groups ? __ldcs(reinterpret_cast<const uchar4*>(codes) + g) : make_uchar4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < kStreamUnroll; ++u) {
        const unsigned g = g0 + u * stride;
        if (g >= groups) continue;
        const unsigned e = g << 2;
        const unsigned rowi = by_frames.div(e);
        unsigned t = e - rowi * n_frames;
        unsigned m = rowi - by_mels.div(rowi) * n_mels;
        const unsigned char cc[4] = {c[u].x, c[u].y, c[u].z, c[u].w};
        float r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          r[i] = __ldg(table + m * n_bins + min((unsigned)cc[i], kmax));
          if (++t == n_frames) { t = 0; m = (m + 1 == n_mels) ? 0 : m + 1; }
        }
        __stcs(reinterpret_cast<float4*>(out) + g, make_float4(r[0], r[1], r[2], r[3]));
      }
    }
  }
  // tail (and the whole tensor when the pointers are not 16-byte aligned)
  const unsigned first = vec_ok ? (groups << 2) : 0;
  for (unsigned e = first + blockIdx.x * kStreamThreads + threadIdx.x; e < n_elems; e += stride) {
    const unsigned rowi = by_frames.div(e);
    const unsigned m = rowi - by_mels.div(rowi) * n_mels;
    out[e] = __ldg(table + m * n_bins + min((unsigned)codes[e], kmax));
  }
}

// ---------------------------------------------------------------------------
// log-mel (float32) -> codes (uint8):  clamp(floor((x - lo) * scale), 0, K-1)
// subtraction and multiply kept un-contracted so codes are bit exact for
// bit-identical x.
// ---------------------------------------------------------------------------
// clamp(floor(pos), 0, kmax) == trunc(clamp(pos, 0, kmax + 0.5)) for every pos, NaN (-> 0) and infinities included:
// one conversion on the XU pipe instead of a round and a conversion.  kmax_half = kmax + 0.5, exact for kmax <= 255.
__device__ __forceinline__ unsigned char quantize_one(float x, float lo, float scale, float kmax_half) {
  const float pos = __fmul_rn(__fsub_rn(x, lo), scale);
  return (unsigned char)__float2int_rz(fminf(fmaxf(pos, 0.f), kmax_half));
}

// n_valid (frames per batch row, or nullptr): frames at or past it get code 0 and their log-mel is never read - the
// padded part of a right-padded batch costs one byte of traffic per value instead of five.  kMasked = false compiles
// the length bookkeeping out (it costs the plain kernel 16 points of its HBM fraction otherwise).
template <bool kMasked>
__global__ void __launch_bounds__(kStreamThreads) quantize_kernel(
    const float* __restrict__ mel, uint8_t* __restrict__ codes, const float* __restrict__ lo,
    const float* __restrict__ scale, unsigned n_elems, FastDiv by_frames, FastDiv by_mels,
    unsigned n_bins, bool vec_ok, const int* __restrict__ n_valid) {
  grid_dependency_wait();   // programmatic dependent launch: the inputs may come from the previous kernel
  grid_launch_dependents();
  const unsigned n_frames = by_frames.d, n_mels = by_mels.d;
  const unsigned groups = n_elems >> 2;
  const unsigned stride = gridDim.x * kStreamThreads;
  const float kmax = float(n_bins - 1) + 0.5f;
  if (vec_ok) {
    for (unsigned g0 = blockIdx.x * kStreamThreads + threadIdx.x; g0 < groups; g0 += stride * kStreamUnroll) {
      float4 x[kStreamUnroll];
      unsigned rowi[kStreamUnroll], t0[kStreamUnroll], nv[kStreamUnroll];
#pragma unroll
      for (int u = 0; u < kStreamUnroll; ++u) {
        const unsigned g = g0 + u * stride;
        const unsigned e = g << 2;
        rowi[u] = by_frames.div(e);
        t0[u] = e - rowi[u] * n_frames;
        nv[u] = n_frames;
        if (kMasked && g < groups) {
          const int v = __ldg(n_valid + by_mels.div(rowi[u]));
          nv[u] = v < 0 ? 0u : min((unsigned)v, n_frames);
        }
        // a group wholly past its row's valid frames (and not running into the next row) is not read at all
        const bool dead = kMasked && t0[u] >= nv[u] && t0[u] + 4 <= n_frames;
        x[u] = (g < groups && !dead) ? __ldcs(reinterpret_cast<const float4*>(mel) + g) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < kStreamUnroll; ++u) {
        const unsigned g = g0 + u * stride;
        if (g >= groups) continue;
        unsigned t = t0[u];
        unsigned row = rowi[u];
        unsigned m = row - by_mels.div(row) * n_mels;
        unsigned valid = nv[u];
        const float xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
        unsigned char r[4];
        float lo_m = __ldg(lo + m), scale_m = __ldg(scale + m);  // reloaded only where the group crosses into the next channel
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          r[i] = (!kMasked || t < valid) ? quantize_one(xs[i], lo_m, scale_m, kmax) : (unsigned char)0;
          if (++t == n_frames) {
            t = 0;
            ++row;
            m = (m + 1 == n_mels) ? 0 : m + 1;
            lo_m = __ldg(lo + m), scale_m = __ldg(scale + m);
            if (kMasked && m == 0) {  // crossed into the next batch row (it exists: the group lies inside the tensor)
              const int v = __ldg(n_valid + by_mels.div(row));
              valid = v < 0 ? 0u : min((unsigned)v, n_frames);
            }
          }
        }
        reinterpret_cast<uchar4*>(codes)[g] = make_uchar4(r[0], r[1], r[2], r[3]);
      }
    }
  }
  const unsigned first = vec_ok ? (groups << 2) : 0;
  for (unsigned e = first + blockIdx.x * kStreamThreads + threadIdx.x; e < n_elems; e += stride) {
    const unsigned rowi = by_frames.div(e);
    const unsigned b = by_mels.div(rowi);
    const unsigned m = rowi - b * n_mels;
    const unsigned t = e - rowi * n_frames;
    const bool live = !kMasked || (int)t < __ldg(n_valid + b);
    codes[e] = live ? quantize_one(mel[e], __ldg(lo + m), __ldg(scale + m), kmax) : (unsigned char)0;
  }
}

// ---------------------------------------------------------------------------
// per-channel running min / max over valid frames of a (B, M, T) tensor.
// One warp per (row, channel) line; exact and order independent.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kStreamThreads) tensor_minmax_kernel(
    const float* __restrict__ mel, const int* __restrict__ n_valid, float* run_min, float* run_max,
    unsigned n_lines, unsigned n_frames, unsigned n_mels) {
  grid_dependency_wait();   // programmatic dependent launch: the inputs may come from the previous kernel
  grid_launch_dependents();
  const unsigned warps_per_block = kStreamThreads / 32;
  const unsigned lane = threadIdx.x & 31;
  for (unsigned line = blockIdx.x * warps_per_block + (threadIdx.x >> 5); line < n_lines;
       line += gridDim.x * warps_per_block) {
    const unsigned b = line / n_mels, m = line - b * n_mels;
    unsigned nv = n_frames;
    if (n_valid) {
      const int v = n_valid[b];
      nv = v < 0 ? 0u : (unsigned(v) < n_frames ? unsigned(v) : n_frames);
    }
    const float* src = mel + (size_t)line * n_frames;
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    for (unsigned t = lane; t < nv; t += 32) {
      const float x = __ldcs(src + t);
      lo = fminf(lo, x);
      hi = fmaxf(hi, x);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if (lane == 0 && lo <= hi) {
      atomic_min_float(run_min + m, lo);
      atomic_max_float(run_max + m, hi);
    }
  }
}

// ---------------------------------------------------------------------------
// What the quantiser derives from its statistics, in one launch (float32, IEEE division):
//   scale[c] = hi[c] - lo[c] > 0 ? K / (hi[c] - lo[c]) : 0      step[c] = (hi[c] - lo[c]) / K
//   *ready   = 1 iff lo[c] <= hi[c] for every channel (the statistics have seen at least one frame)
// One CTA; replaces nine elementwise launches of the host framework after every change of the statistics.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) quantizer_derive_kernel(const float* __restrict__ lo, const float* __restrict__ hi, int n_mels,
                                                               float k, float* __restrict__ scale, float* __restrict__ step,
                                                               int* __restrict__ ready) {
  grid_dependency_wait();
  grid_launch_dependents();
  int ok = 1;
  for (int c = threadIdx.x; c < n_mels; c += blockDim.x) {
    const float l = lo[c], h = hi[c];
    const float width = __fsub_rn(h, l);
    if (scale) scale[c] = width > 0.f ? __fdiv_rn(k, width) : 0.f;
    if (step) step[c] = __fdiv_rn(width, k);
    ok &= (l <= h) ? 1 : 0;
  }
  ok = __syncthreads_and(ok);
  if (threadIdx.x == 0 && ready) *ready = ok;
}

}  // namespace dmel
