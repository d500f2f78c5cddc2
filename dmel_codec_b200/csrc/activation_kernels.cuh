// BigVGAN's anti-aliased Snake activation for sm_100a: 2x upsample (12-tap Kaiser-sinc FIR, replicate padding) ->
// x + sin^2(a x) / b -> 2x downsample (12-tap FIR, replicate padding), one pass over HBM.
//
// Replaces the reference's Activation1d.forward (models/modules/bigvgan/alias_free_activation/torch/act.py:24-29:
// UpSample1d -> SnakeBeta -> DownSample1d, each a full HBM round trip of a tensor twice the size) and its fused
// sm_70/sm_80 kernel (.../cuda/anti_alias_activation_cuda.cu:44-179, shipped without PTX: it cannot run on sm_100).
//
// One CTA = 1008 consecutive outputs of one (batch, channel) row:
//   1. x[t0 - 8, t0 + 1016) -> shared (one LDG.128 per thread; row ends read through a clamped index, which IS the
//      replicate padding of the upsampler);
//   2. each thread produces 8 samples of the activated 2x signal s (4 even, 4 odd) from 12 x values held in registers
//      (3 LDS.128): u[2m] = sum_q x[m-3+q] F[11-2q], u[2m+1] = sum_q x[m-2+q] F[10-2q] (F = 2 * up taps: the polyphase
//      form of the transposed convolution), then the Snake with sin^2 evaluated on an argument reduced modulo pi
//      (sin^2 has period pi: two-constant Cody-Waite + MUFU.SIN stays within 1e-6 of the exact value for the
//      arguments a trained alpha produces); s is stored de-interleaved (even / odd) so that
//   3. each thread computes 4 outputs from 6 LDS.128: y[t] = sum_a G[2a+1] s_even[..] + G[2a] s_odd[..].
// Samples of s before the start / past the end of the row (the downsampler's replicate padding) are patched in shared
// memory between steps 2 and 3.  HBM traffic: 4 bytes in, 4 bytes out per sample; ~46 instructions per output.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dmel {

constexpr int kActThreads = 256;
constexpr int kActTile = 1008;            // outputs per CTA (a multiple of 4)
constexpr int kActX = kActTile + 16;      // staged inputs: x[t0 - 8 .. t0 + kActTile + 8)
constexpr int kActS = kActTile + 8;       // entries of s_even / s_odd: s[2 t0 - 6 + 2 i], s[2 t0 - 5 + 2 i]

struct ActTaps {
  float up2[12];   // 2 * up-sampling taps
  float down[12];
};

__device__ __forceinline__ float snake_value(float u, float a, float inv_b) {
  const float v = u * a;
  const float k = rintf(v * 0.31830988618379067f);                 // v / pi
  float r = fmaf(-k, 3.140625f, v);                                 // pi = 3.140625 + 9.6765358979e-4 (Cody-Waite)
  r = fmaf(-k, 9.67653589793e-4f, r);
  const float sn = __sinf(r);
  return fmaf(inv_b, sn * sn, u);
}

// grid (ceil(T / kActTile), C, B)
__global__ void __launch_bounds__(kActThreads) antialias_snake_kernel(const float* __restrict__ x, float* __restrict__ y, int n_t,
                                                                      const float* __restrict__ log_alpha,
                                                                      const float* __restrict__ log_beta, ActTaps taps) {
  __shared__ __align__(16) float xs[kActX];
  __shared__ __align__(16) float se[kActS + 8];
  __shared__ __align__(16) float so[kActS + 8];
  grid_dependency_wait();
  grid_launch_dependents();
  const int c = blockIdx.y, row = blockIdx.z * gridDim.y + c;
  const int t0 = blockIdx.x * kActTile;
  const int nt = min(kActTile, n_t - t0);
  const float* xr = x + (size_t)row * n_t;
  float* yr = y + (size_t)row * n_t;
  const float a = __expf(log_alpha[c]);
  const float inv_b = 1.0f / (__expf(log_beta[c]) + 1e-9f);
  const int tid = threadIdx.x;

  // ---- 1. stage x[t0 - 8 + j], j in [0, kActX): 4 per thread
  {
    const int j0 = tid * 4, g0 = t0 - 8 + j0;
    const bool vec = ((reinterpret_cast<uintptr_t>(xr) & 15) == 0) && ((n_t & 3) == 0) && g0 >= 0 && g0 + 3 < n_t;
    float4 v;
    if (vec) {
      v = __ldg(reinterpret_cast<const float4*>(xr + g0));
    } else {
      v.x = __ldg(xr + min(max(g0, 0), n_t - 1));
      v.y = __ldg(xr + min(max(g0 + 1, 0), n_t - 1));
      v.z = __ldg(xr + min(max(g0 + 2, 0), n_t - 1));
      v.w = __ldg(xr + min(max(g0 + 3, 0), n_t - 1));
    }
    *reinterpret_cast<float4*>(xs + j0) = v;  // kActX == 4 * kActThreads
  }
  __syncthreads();

  // ---- 2. activated 2x signal: entries i0 .. i0 + 3 of s_even and s_odd
  if (tid * 4 < kActS) {
    const int i0 = tid * 4;
    float w[12];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(xs + i0 + 4 * q);
      w[4 * q] = v.x, w[4 * q + 1] = v.y, w[4 * q + 2] = v.z, w[4 * q + 3] = v.w;
    }
    float e[4], o[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      // s_even[i]: n = 2 t0 - 6 + 2 i -> m = t0 - 3 + i -> x[m - 3 + q] = xs[i + 2 + q];  s_odd[i]: x[m - 2 + q] = xs[i + 3 + q]
      float ue = 0.f, uo = 0.f;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        ue = fmaf(w[d + 2 + q], taps.up2[11 - 2 * q], ue);
        uo = fmaf(w[d + 3 + q], taps.up2[10 - 2 * q], uo);
      }
      e[d] = snake_value(ue, a, inv_b);
      o[d] = snake_value(uo, a, inv_b);
    }
    *reinterpret_cast<float4*>(se + i0) = make_float4(e[0], e[1], e[2], e[3]);
    *reinterpret_cast<float4*>(so + i0) = make_float4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();
  // ---- replicate padding of the downsampler: s[n] = s[0] for n < 0 (first tile), s[n] = s[2T - 1] for n >= 2T (last tile)
  if (t0 == 0 && tid < 3) {       // s[0] = s_even[3]
    const float first = se[3];
    se[tid] = first;
    so[tid] = first;              // s_odd[i] is s[2 i - 5]: negative for i < 3
  }
  {  // s[2T - 1] = s_odd[i_last]; every later entry of s_even / s_odd lies past the row (the last tile, and the one
     // before it when the last holds fewer than three outputs)
    const int i_last = n_t - t0 + 2;
    const int i = i_last + 1 + tid;
    if (i_last < kActS && tid < 8 && i < kActS + 8) {
      const float last = so[i_last];
      se[i] = last;
      so[i] = last;
    }
  }
  __syncthreads();

  // ---- 3. outputs r0 .. r0 + 3:  y[r] = sum_a G[2a + 1] s_even[r + a + 1] + G[2a] s_odd[r + a]
  const int r0 = tid * 4;
  if (r0 < nt) {
    float ev[12], ov[12];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(se + r0 + 4 * q);
      const float4 u = *reinterpret_cast<const float4*>(so + r0 + 4 * q);
      ev[4 * q] = v.x, ev[4 * q + 1] = v.y, ev[4 * q + 2] = v.z, ev[4 * q + 3] = v.w;
      ov[4 * q] = u.x, ov[4 * q + 1] = u.y, ov[4 * q + 2] = u.z, ov[4 * q + 3] = u.w;
    }
    float out[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        acc = fmaf(ov[d + k], taps.down[2 * k], acc);
        acc = fmaf(ev[d + k + 1], taps.down[2 * k + 1], acc);
      }
      out[d] = acc;
    }
    float* dst = yr + t0 + r0;
    if (r0 + 3 < nt && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
      for (int d = 0; d < 4; ++d)
        if (r0 + d < nt) dst[d] = out[d];
    }
  }
}

}  // namespace dmel
