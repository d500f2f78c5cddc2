// BigVGAN's anti-aliased Snake activation for sm_100a: 2x upsample (12-tap Kaiser-sinc FIR, replicate padding) ->
// x + sin^2(a x) / b -> 2x downsample (12-tap FIR, replicate padding), one pass over HBM.
//
// Replaces the reference's Activation1d.forward (models/modules/bigvgan/alias_free_activation/torch/act.py:24-29:
// UpSample1d -> SnakeBeta -> DownSample1d, each a full HBM round trip of a tensor twice the size) and its fused
// sm_70/sm_80 kernel (.../cuda/anti_alias_activation_cuda.cu:44-179, shipped without PTX: it cannot run on sm_100).
//
// One WARP = 240 consecutive outputs of one (batch, channel) row, no shared memory and no barrier.  The op is bound
// by the FP32 pipe, not by HBM: two 12-tap FIRs and the Snake are >= 40 FMA-pipe lane-cycles per output, which at
// 128 lanes per SM and cycle is the same rate as the HBM roofline (8 bytes per output), and every lane-cycle spent
// on anything else comes straight out of the throughput (profiles/r2_standalone_kernels_ncu_summary.txt).
//   1. lane l loads x[t0 - 8 + 8l .. + 8) with two LDG.128 (row ends through a clamped index, which IS the replicate
//      padding of the upsampler) and fetches the eight samples that follow from lane l+1 (8 SHFL);
//   2. it produces 8 PAIRS P[i] = (s[2 t0 - 5 + 2i], s[2 t0 - 4 + 2i]), i = 8l .. 8l+7, of the activated 2x signal s.
//      Both halves of a pair are polyphase sums over the SAME six inputs,
//        u_odd = sum_q x[t0 - 5 + i + q] F[10 - 2q],   u_even = sum_q x[t0 - 5 + i + q] F[11 - 2q]   (F = 2 * up taps),
//      so one packed FMA (fma.rn.f32x2, the input broadcast to both halves) advances both; the Snake runs packed as
//      well, with sin^2 evaluated on an argument reduced modulo pi (sin^2 has period pi: round-by-magic-constant,
//      two-constant Cody-Waite and MUFU.SIN stay within 1e-6 of the exact value for the arguments a trained alpha
//      produces).  Pairs that stand for samples of s before the start / past the end of the row (the downsampler's
//      replicate padding) are overwritten in registers;
//   3. it fetches the five pairs that follow from lane l+1 (10 SHFL) and computes 8 outputs,
//      y[t0 + r] = sum_a P[r + a] . (G[2a], G[2a+1]): six packed FMAs and one add per output, two STG.128.
// Lane 31 of step 2 and lanes 30-31 of step 3 have no complete neighbourhood and produce nothing: a warp reads 256
// samples (L2 serves the 16 that overlap the next warp's) to write 240.  The grid is persistent (warps stride over
// the tiles, the next tile's samples are in flight while this one is computed): benchmarks/stream_read.cu measures
// what one short-lived block per tile costs on this part.  HBM traffic: 4 bytes in, 4 bytes out per sample.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fastdiv.cuh"
#include "fft_core.cuh"  // packed f32x2 arithmetic

namespace dmel {

constexpr int kActThreads = 256;
#ifndef DMEL_ACT_MIN_CTAS
#define DMEL_ACT_MIN_CTAS 3  // register budget of the tile body: 80 (measured against 2 and 4 CTAs per SM, profiles/history.md)
#endif
constexpr int kActPer = 8;                // outputs per lane
constexpr int kActWarpOut = 30 * kActPer;  // outputs per warp and tile

struct ActTaps {
  float up2[12];   // 2 * up-sampling taps
  float down[12];
};

// x + sin^2(a x) / b on both halves of a pair
__device__ __forceinline__ float2 snake_pair(float2 u, float a, float inv_b) {
  const float2 v = f2_mul(u, make_float2(a, a));
  constexpr float kMagic = 12582912.f;  // 1.5 * 2^23: adding it rounds to the nearest integer (|v / pi| < 2^22)
  const float2 k = f2_add(f2_fma(v, make_float2(0.31830988618379067f, 0.31830988618379067f), make_float2(kMagic, kMagic)),
                          make_float2(-kMagic, -kMagic));
  float2 r = f2_fma(k, make_float2(-3.140625f, -3.140625f), v);                 // pi = 3.140625 + 9.6765358979e-4 (Cody-Waite)
  r = f2_fma(k, make_float2(-9.67653589793e-4f, -9.67653589793e-4f), r);
  const float2 sn = make_float2(__sinf(r.x), __sinf(r.y));
  return f2_fma(make_float2(inv_b, inv_b), f2_mul(sn, sn), u);
}

// four consecutive samples of one row from index g0 on, indices clamped into the row
__device__ __forceinline__ float4 act_load(const float* __restrict__ xr, int g0, int n_t, bool row_vec) {
  if (row_vec && g0 >= 0 && g0 + 3 < n_t) return __ldg(reinterpret_cast<const float4*>(xr + g0));
  float4 v;
  v.x = __ldg(xr + min(max(g0, 0), n_t - 1));
  v.y = __ldg(xr + min(max(g0 + 1, 0), n_t - 1));
  v.z = __ldg(xr + min(max(g0 + 2, 0), n_t - 1));
  v.w = __ldg(xr + min(max(g0 + 3, 0), n_t - 1));
  return v;
}

// One warp tile: 240 outputs from the 256 samples in (cur0, cur1) of every lane.  kEdge = the tile touches an end of
// its row (replicate padding of the downsampler, partial stores); interior tiles are compiled without any of it -
// as one body, the conditional overwrites of the pairs cost 45 register moves per tile on the pipe that binds.
template <bool kEdge>
__device__ __forceinline__ void act_tile(float4 cur0, float4 cur1, int lane, int t0, int n_t, float a, float inv_b,
                                         const ActTaps& taps, float* __restrict__ yrow, bool vec_st) {
  // ---- 1. w[j] = x[t0 - 8 + 8 lane + j], j = 0 .. 15
  float w[16] = {cur0.x, cur0.y, cur0.z, cur0.w, cur1.x, cur1.y, cur1.z, cur1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) w[8 + j] = __shfl_down_sync(0xffffffffu, w[j], 1);
  // ---- 2. pairs 8 lane .. 8 lane + 7;  x[t0 - 5 + i + q] = w[d + 3 + q]
  float2 pv[kActPer + 5];
#pragma unroll
  for (int d = 0; d < kActPer; ++d) {
    float2 u = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 6; ++q)
      u = f2_fma(make_float2(w[d + 3 + q], w[d + 3 + q]), make_float2(taps.up2[10 - 2 * q], taps.up2[11 - 2 * q]), u);
    pv[d] = snake_pair(u, a, inv_b);
  }
  if constexpr (kEdge) {  // replicate padding of the downsampler, in registers
    if (t0 == 0 && lane == 0) {     // s[n] = s[0] = P[2].y for n < 0 (the row's first tile)
      const float first = pv[2].y;
      pv[0] = pv[1] = make_float2(first, first);
      pv[2].x = first;
    }
    const int i_last = n_t - t0 + 2;
    if (i_last < 31 * kActPer) {    // s[n] = s[2T - 1] = P[i_last].x for n >= 2T (warp-uniform)
      float mine = pv[0].x;
#pragma unroll
      for (int d = 1; d < kActPer; ++d) mine = (i_last & (kActPer - 1)) == d ? pv[d].x : mine;
      const float last = __shfl_sync(0xffffffffu, mine, i_last / kActPer);
#pragma unroll
      for (int d = 0; d < kActPer; ++d) {
        const int i = kActPer * lane + d;
        if (i > i_last) pv[d] = make_float2(last, last);
        else if (i == i_last) pv[d].y = last;
      }
    }
  }
  // ---- 3. outputs 8 lane .. 8 lane + 7 from pairs 8 lane .. 8 lane + 12
#pragma unroll
  for (int d = 0; d < 5; ++d)
    pv[kActPer + d] = make_float2(__shfl_down_sync(0xffffffffu, pv[d].x, 1), __shfl_down_sync(0xffffffffu, pv[d].y, 1));
  float out[kActPer];
#pragma unroll
  for (int d = 0; d < kActPer; ++d) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 6; ++k) acc = f2_fma(pv[d + k], make_float2(taps.down[2 * k], taps.down[2 * k + 1]), acc);
    out[d] = acc.x + acc.y;
  }
  const int r0 = t0 + kActPer * lane;
  if (lane < kActWarpOut / kActPer && (!kEdge || r0 < n_t)) {
    float* dst = yrow + r0;
    if (vec_st && (!kEdge || r0 + kActPer - 1 < n_t)) {
      __stcs(reinterpret_cast<float4*>(dst), make_float4(out[0], out[1], out[2], out[3]));
      __stcs(reinterpret_cast<float4*>(dst) + 1, make_float4(out[4], out[5], out[6], out[7]));
    } else {
#pragma unroll
      for (int d = 0; d < kActPer; ++d)
        if (r0 + d < n_t) dst[d] = out[d];
    }
  }
}

// grid: any number of CTAs; n_tiles = rows * tiles_per_row warp tiles, rows = batch * channels
__global__ void __launch_bounds__(kActThreads, DMEL_ACT_MIN_CTAS) antialias_snake_kernel(const float* __restrict__ x, float* __restrict__ y, int n_t,
                                                                      FastDiv tiles_per_row, FastDiv n_channels, unsigned n_tiles,
                                                                      const float* __restrict__ log_alpha,
                                                                      const float* __restrict__ log_beta, const ActTaps taps) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int lane = threadIdx.x & 31;
  const unsigned n_warps = gridDim.x * (kActThreads / 32);
  unsigned tile = blockIdx.x * (kActThreads / 32) + (threadIdx.x >> 5);
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((n_t & 3) == 0);
  const bool vec_st = ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((n_t & 3) == 0);

  unsigned row = tiles_per_row.div(min(tile, n_tiles - 1));
  int t0 = (int)(min(tile, n_tiles - 1) - row * tiles_per_row.d) * kActWarpOut;
  const float* xr = x + (size_t)row * n_t;
  float4 cur0 = act_load(xr, t0 - 8 + kActPer * lane, n_t, vec_ok), cur1 = act_load(xr, t0 - 4 + kActPer * lane, n_t, vec_ok);
  for (; tile < n_tiles;) {
    // the next tile's samples leave for the registers before this tile's arithmetic starts
    const unsigned next = tile + n_warps;
    unsigned next_row = row;
    int next_t0 = t0;
    float4 nxt0 = cur0, nxt1 = cur1;
    if (next < n_tiles) {
      next_row = tiles_per_row.div(next);
      next_t0 = (int)(next - next_row * tiles_per_row.d) * kActWarpOut;
#ifndef DMEL_ACT_NO_PREFETCH
      xr = x + (size_t)next_row * n_t;
      nxt0 = act_load(xr, next_t0 - 8 + kActPer * lane, n_t, vec_ok);
      nxt1 = act_load(xr, next_t0 - 4 + kActPer * lane, n_t, vec_ok);
#endif
    }
    const unsigned c = row - n_channels.div(row) * n_channels.d;
    const float a = __expf(__ldg(log_alpha + c));
    const float inv_b = 1.0f / (__expf(__ldg(log_beta + c)) + 1e-9f);
    float* yrow = y + (size_t)row * n_t;
    // interior: no pair of the tile stands for a sample outside the row, and all 240 outputs exist
    if (t0 != 0 && n_t - t0 + 2 >= 31 * kActPer) act_tile<false>(cur0, cur1, lane, t0, n_t, a, inv_b, taps, yrow, vec_st);
    else act_tile<true>(cur0, cur1, lane, t0, n_t, a, inv_b, taps, yrow, vec_st);
    tile = next, row = next_row, t0 = next_t0, cur0 = nxt0, cur1 = nxt1;
#ifdef DMEL_ACT_NO_PREFETCH
    if (tile < n_tiles) {
      xr = x + (size_t)row * n_t;
      cur0 = act_load(xr, t0 - 8 + kActPer * lane, n_t, vec_ok);
      cur1 = act_load(xr, t0 - 4 + kActPer * lane, n_t, vec_ok);
    }
#endif
  }
}

}  // namespace dmel
