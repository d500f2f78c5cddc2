// Fused waveform -> log-mel -> dMel codes kernel (sm_100a).
//
// One CTA owns a tile of TF consecutive frames of one utterance row:
//   1. stage  (TF-1)*hop + n_fft  samples in shared memory, reflect-indexed at
//      the row ends (reference utils/spectrogram.py:58-62), so the 4x frame
//      overlap is re-read on chip, not from HBM;
//   2. each warp turns frames into magnitudes with the register FFT of
//      fft_core.cuh (window multiply on load, torch.stft at :64-75, magnitude
//      at :76) and drops them in a [frame][bin] shared tile;
//   3. the banded mel filterbank (:78), log(clamp(.,1e-5)) (:38-39) and the
//      per-channel bin quantiser run with one lane per
//      frame, and only codes / log-mel / min-max leave the SM.
//
// HBM traffic per tile: hop*TF*4 B of new waveform in, n_mels*TF B of codes
// out; constants come from L2.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fft_core.cuh"

namespace dmel {

constexpr float kLogClip = 1e-5f;  // reference utils/spectrogram.py:38
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

struct FusedParams {
  const float* wav;         // (B, row_stride) device
  long long row_stride;     // samples between rows
  int n_rows;               // B
  int n_samples;            // L
  int n_frames;             // T
  int tiles_per_row;
  int n_tiles;
  int hop;
  int pad_inner;            // (n_fft - hop)/2 reflect pad of the reference
  int pad_outer;            // n_fft/2 when center=True, else 0
  int n_mels;
  int wave_len;             // staged samples per tile
  int nnz;                  // banded weights
  const float* window;      // (n_fft)
  const float2* stage_tw;   // [32][32]: W_1024^{k1*n2} at [k1*32 + n2]
  const float2* fold_tw;    // [513]: W_2048^k (n_fft == 2048 only)
  const int4* chan;         // per channel {first bin, count (mult. of 4), weight offset, 0}
  const float* weights;
  const int* lengths;       // valid samples per row, or null
  float* logmel;            // (B, M, T) or null
  unsigned char* codes;     // (B, M, T) or null
  const float* q_lo;        // (M)
  const float* q_scale;     // (M)  K / (hi - lo)
  int n_bins;
  float* run_min;           // (M) running min, updated in place, or null
  float* run_max;           // (M)
  unsigned long long* near_edge;  // count of values within edge_eps of an interior edge, or null
  float edge_eps;
};

// index into the unpadded row for position j of the (doubly) reflect-padded row
__device__ __forceinline__ int reflect_src(int j, int n, int pad_inner, int pad_outer) {
  if (pad_outer) {
    const int nq = n + 2 * pad_inner;
    j -= pad_outer;
    j = j < 0 ? -j : j;
    j = j >= nq ? 2 * (nq - 1) - j : j;
  }
  j -= pad_inner;
  j = j < 0 ? -j : j;
  j = j >= n ? 2 * (n - 1) - j : j;
  return j;
}

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

template <int NFFT, int TF>
struct FusedLayout {
  static constexpr int kBins = NFFT / 2 + 1;
  static constexpr int kMagPitch = kBins;  // 513 / 1025: == 1 (mod 32), lane-per-frame reads hit 32 banks
  static constexpr int kMagFloats = TF * kMagPitch + 4;
  // byte offsets inside dynamic shared memory (all 16-byte aligned)
  static __host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
  static __host__ __device__ size_t tiles_off() { return 0; }
  static __host__ __device__ size_t mags_off() { return size_t(kWarps) * kTileFloat2 * sizeof(float2); }
  static __host__ __device__ size_t wave_off() { return align16(mags_off() + size_t(kMagFloats) * 4); }
  static __host__ __device__ size_t window_off(int wave_len) { return align16(wave_off() + size_t(wave_len) * 4); }
  static __host__ __device__ size_t fold_off(int wave_len) {
    return align16(window_off(wave_len) + (NFFT == 2048 ? size_t(NFFT) * 4 : 0));
  }
  static __host__ __device__ size_t chan_off(int wave_len) {
    return align16(fold_off(wave_len) + (NFFT == 2048 ? size_t(513) * 8 : 0));
  }
  static __host__ __device__ size_t weights_off(int wave_len, int n_mels) {
    return align16(chan_off(wave_len) + size_t(n_mels) * 16);
  }
  static __host__ __device__ size_t stats_off(int wave_len, int n_mels, int nnz) {
    return align16(weights_off(wave_len, n_mels) + size_t(nnz) * 4);
  }
  static __host__ __device__ size_t total(int wave_len, int n_mels, int nnz) {
    return stats_off(wave_len, n_mels, nnz) + size_t(n_mels) * 8;
  }
};

template <int NFFT, int TF>
__global__ void __launch_bounds__(kThreads, 1) dmel_fused_kernel(const FusedParams p) {
  static_assert(NFFT == 1024 || NFFT == 2048, "register FFT core is 1024 complex points");
  static_assert(TF == 32 || TF == 16 || TF == 8, "tile frames");
  using LY = FusedLayout<NFFT, TF>;
  constexpr int kBins = LY::kBins;
  constexpr int kPitch = LY::kMagPitch;

  extern __shared__ __align__(16) unsigned char smem[];
  float2* tiles = reinterpret_cast<float2*>(smem + LY::tiles_off());
  float* mags = reinterpret_cast<float*>(smem + LY::mags_off());
  float* wave = reinterpret_cast<float*>(smem + LY::wave_off());
  float* s_window = reinterpret_cast<float*>(smem + LY::window_off(p.wave_len));
  float2* s_fold = reinterpret_cast<float2*>(smem + LY::fold_off(p.wave_len));
  int4* s_chan = reinterpret_cast<int4*>(smem + LY::chan_off(p.wave_len));
  float* s_weights = reinterpret_cast<float*>(smem + LY::weights_off(p.wave_len, p.n_mels));
  float* s_min = reinterpret_cast<float*>(smem + LY::stats_off(p.wave_len, p.n_mels, p.nnz));
  float* s_max = s_min + p.n_mels;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  float2* my_tile = tiles + warp * kTileFloat2;

  // ---- per-CTA constants -------------------------------------------------
  for (int i = tid; i < p.n_mels; i += kThreads) {
    s_chan[i] = p.chan[i];
    s_min[i] = __int_as_float(0x7f800000);
    s_max[i] = __int_as_float(0xff800000);
  }
  for (int i = tid; i < p.nnz; i += kThreads) s_weights[i] = p.weights[i];
  if (tid < 4) mags[TF * kPitch + tid] = 0.f;  // banded spans may over-read 3 floats
  if constexpr (NFFT == 2048) {
    for (int i = tid; i < NFFT; i += kThreads) s_window[i] = p.window[i];
    for (int i = tid; i < 513; i += kThreads) s_fold[i] = p.fold_tw[i];
  }

  // per-lane inter-pass twiddles W_1024^{lane*k1}; window taps for the packed path
  float2 tw[32];
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) tw[k1] = p.stage_tw[k1 * 32 + lane];
  float win[NFFT == 1024 ? 32 : 1];
  if constexpr (NFFT == 1024) {
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) win[n1] = p.window[32 * n1 + lane];
  }

  const int padded_len = p.n_samples + 2 * p.pad_inner + 2 * p.pad_outer;
  unsigned long long edge_hits = 0;

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int row = tile / p.tiles_per_row;
    const int t0 = (tile - row * p.tiles_per_row) * TF;
    const float* src = p.wav + (long long)row * p.row_stride;
    int n_valid = p.n_frames;
    if (p.lengths) {
      n_valid = p.lengths[row] / p.hop;
      n_valid = n_valid < p.n_frames ? n_valid : p.n_frames;
    }
    const bool need_values = p.logmel != nullptr;  // log-mel output covers every frame
    const bool tile_dead = !need_values && t0 >= n_valid;

    __syncthreads();  // previous tile: mel phase done with mags, FFT done with wave

    if (!tile_dead) {
      // ---- 1. stage the waveform tile -----------------------------------
      const int j0 = t0 * p.hop;                      // first padded-row position
      const int s0 = j0 - p.pad_inner - p.pad_outer;  // its source sample, if interior
      const bool interior = s0 >= 0 && s0 + p.wave_len <= p.n_samples &&
                            ((reinterpret_cast<uintptr_t>(src + s0) & 15) == 0) && (p.wave_len % 4 == 0);
      if (interior) {
        const float4* g4 = reinterpret_cast<const float4*>(src + s0);
        float4* w4 = reinterpret_cast<float4*>(wave);
        for (int i = tid; i < p.wave_len / 4; i += kThreads) w4[i] = __ldg(g4 + i);
      } else {
        for (int i = tid; i < p.wave_len; i += kThreads) {
          const int j = j0 + i;
          float x = 0.f;
          if (j < padded_len) x = __ldg(src + reflect_src(j, p.n_samples, p.pad_inner, p.pad_outer));
          wave[i] = x;
        }
      }
      __syncthreads();

      // ---- 2. FFT -> magnitudes -----------------------------------------
      if constexpr (NFFT == 1024) {
        for (int pr = warp; pr < TF / 2; pr += kWarps) {
          const float* fa = wave + (2 * pr) * p.hop;
          const float* fb = fa + p.hop;
          float2 v[32];
#pragma unroll
          for (int n1 = 0; n1 < 32; ++n1) {
            const int idx = 32 * n1 + lane;
            v[n1] = make_float2(fa[idx] * win[n1], fb[idx] * win[n1]);
          }
          __syncwarp();
          fft1024_pass1(v, tw, my_tile, lane);
          __syncwarp();
          fft1024_pass2(v, my_tile, lane);
          float* ma = mags + (2 * pr) * kPitch;
          float* mb = ma + kPitch;
          const int partner = (32 - lane) & 31;
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            const float2 send = (lane == 0) ? v[brev5(mirror_slot(k2, true))] : v[brev5(mirror_slot(k2, false))];
            const float2 bm = make_float2(__shfl_sync(0xffffffffu, send.x, partner),
                                          __shfl_sync(0xffffffffu, send.y, partner));
            float xa, xb;
            packed_pair_magnitudes(v[brev5(k2)], bm, xa, xb);
            ma[32 * k2 + lane] = xa;
            mb[32 * k2 + lane] = xb;
          }
          if (lane == 0) {  // Nyquist bin 512 pairs with itself
            float xa, xb;
            packed_pair_magnitudes(v[brev5(16)], v[brev5(16)], xa, xb);
            ma[512] = xa;
            mb[512] = xb;
          }
        }
      } else {
        for (int fr = warp; fr < TF; fr += kWarps) {
          const float* fa = wave + fr * p.hop;
          float2 v[32];
#pragma unroll
          for (int n1 = 0; n1 < 32; ++n1) {
            const int idx = 2 * (32 * n1 + lane);
            const float2 w = *reinterpret_cast<const float2*>(s_window + idx);
            v[n1] = make_float2(fa[idx] * w.x, fa[idx + 1] * w.y);
          }
          __syncwarp();
          fft1024_pass1(v, tw, my_tile, lane);
          __syncwarp();
          fft1024_pass2(v, my_tile, lane);
          float* mrow = mags + fr * kPitch;
          const int partner = (32 - lane) & 31;
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            const float2 send = (lane == 0) ? v[brev5(mirror_slot(k2, true))] : v[brev5(mirror_slot(k2, false))];
            const float2 bm = make_float2(__shfl_sync(0xffffffffu, send.x, partner),
                                          __shfl_sync(0xffffffffu, send.y, partner));
            const int k = 32 * k2 + lane;
            float xk, xm;
            folded_magnitudes(v[brev5(k2)], bm, s_fold[k], xk, xm);
            mrow[k] = xk;
            mrow[1024 - k] = xm;
          }
          if (lane == 0) {
            float xk, xm;
            folded_magnitudes(v[brev5(16)], v[brev5(16)], s_fold[512], xk, xm);
            mrow[512] = xk;
          }
        }
      }
    }
    __syncthreads();

    // ---- 3. mel filterbank, log, quantise --------------------------------
    constexpr int kGroups = 32 / TF;  // channels handled side by side in one warp
    const int fr = lane % TF;
    const int sub = lane / TF;
    const int t = t0 + fr;
    const float* mrow = mags + fr * kPitch;
    // the channel loop is warp-uniform (shuffles inside); a group past the last channel idles
    for (int mb = warp * kGroups; mb < p.n_mels; mb += kWarps * kGroups) {
      const int m = mb + sub;
      const bool live = m < p.n_mels;
      float value = 0.f;
      if (live && !tile_dead) {
        const int4 c = s_chan[m];
        const float4* w4 = reinterpret_cast<const float4*>(s_weights + c.z);
        const float* x = mrow + c.x;
        float acc = 0.f;
        for (int i = 0; i < c.y; i += 4) {
          const float4 w = w4[i >> 2];
          acc = fmaf(w.x, x[i], acc);
          acc = fmaf(w.y, x[i + 1], acc);
          acc = fmaf(w.z, x[i + 2], acc);
          acc = fmaf(w.w, x[i + 3], acc);
        }
        value = __logf(fmaxf(acc, kLogClip));
      }
      const bool in_row = live && t < p.n_frames;
      const bool valid = live && t < n_valid;
      const long long o = ((long long)row * p.n_mels + m) * p.n_frames + t;
      if (p.logmel && in_row) p.logmel[o] = value;
      if (p.codes && in_row) {
        unsigned char code = 0;
        if (valid) {
          const float pos = __fmul_rn(__fsub_rn(value, __ldg(p.q_lo + m)), __ldg(p.q_scale + m));
          const float q = fminf(fmaxf(floorf(pos), 0.f), float(p.n_bins - 1));
          code = (unsigned char)q;
          if (p.near_edge) {
            const float e = fminf(fmaxf(rintf(pos), 1.f), float(p.n_bins - 1));
            if (fabsf(pos - e) < p.edge_eps * __ldg(p.q_scale + m)) ++edge_hits;
          }
        }
        p.codes[o] = code;
      }
      if (p.run_min) {
        float lo = valid ? value : __int_as_float(0x7f800000);
        float hi = valid ? value : __int_as_float(0xff800000);
#pragma unroll
        for (int d = TF / 2; d >= 1; d >>= 1) {
          lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
          hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
        }
        if (fr == 0 && live) {  // channel m always belongs to this lane of this warp: no race
          s_min[m] = fminf(s_min[m], lo);
          s_max[m] = fmaxf(s_max[m], hi);
        }
      }
    }
  }

  // ---- flush per-CTA statistics -------------------------------------------
  if (p.run_min) {
    __syncthreads();
    for (int m = tid; m < p.n_mels; m += kThreads) {
      const float lo = s_min[m], hi = s_max[m];
      if (lo <= hi) {
        atomic_min_float(p.run_min + m, lo);
        atomic_max_float(p.run_max + m, hi);
      }
    }
  }
  if (p.near_edge) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) edge_hits += __shfl_xor_sync(0xffffffffu, edge_hits, d);
    if (lane == 0 && edge_hits) atomicAdd(p.near_edge, edge_hits);
  }
}

}  // namespace dmel
