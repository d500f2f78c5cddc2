// Fused waveform -> log-mel -> dMel codes kernel (sm_100a).
//
// One CTA owns tiles of TF consecutive frames of one utterance row:
//   1. the (TF-1)*hop + n_fft samples of the NEXT tile are pulled into shared
//      memory by one bulk async copy (cp.async.bulk + mbarrier, double
//      buffered) while the current tile computes; row ends, where the
//      reference reflect-pads (utils/spectrogram.py:58-62), are staged by
//      ordinary reflect-indexed loads.  The 4x frame overlap is re-read on
//      chip, never from HBM;
//   2. each warp turns frames into magnitudes with the register FFT of
//      fft_core.cuh (window multiply on load, torch.stft at :64-75, magnitude
//      at :76) and drops them in a [frame][bin] shared tile;
//   3. the banded mel filterbank (:78), log(clamp(.,1e-5)) (:38-39) and the
//      per-channel bin quantiser run with one lane per
//      frame, and only codes / log-mel / min-max leave the SM.
//
// HBM traffic per tile: hop*TF*4 B of new waveform in, n_mels*TF B of codes
// out; constants come from L2.
//
// MODE (template parameter, bits kOut* / kIn* below) selects what a launch reads and writes: codes,
// log-mel (float32 or bfloat16, optionally masked past the valid frames, optionally with per-channel
// time sums), calibration min/max, near-edge counts, the bin centre of every code (the quantiser's
// forward), int16 PCM input.  Tiles after a CTA's first are handed out by a global counter, and the
// kernel takes part in programmatic dependent launch: plan-owned constants are loaded before
// griddepcontrol.wait, caller memory is touched only after it.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "fft_core.cuh"

namespace dmel {

constexpr float kLogClip = 1e-5f;  // reference utils/spectrogram.py:38
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

// what one launch writes (template parameter: dead outputs cost no instructions)
constexpr int kOutCodes = 1;
constexpr int kOutLogmel = 2;
constexpr int kOutStats = 4;
constexpr int kOutEdge = 8;  // count values within edge_eps of an interior bin edge (needs kOutCodes)
constexpr int kOutBf16 = 16;  // the log-mel output tensor is bfloat16 (needs kOutLogmel)
constexpr int kOutDequant = 64;  // also write the bin centre of every code, float32 (needs kOutCodes): the quantiser's forward
constexpr int kInPcm16 = 32;  // the waveform is int16 PCM (x / 32768 is folded into the window taps); lean variants only

struct FusedParams {
  const float* wav;         // (B, row_stride) device; int16_t with kInPcm16
  long long row_stride;     // samples between rows
  int n_rows;               // B
  int n_samples;            // L: length of the (virtual) row the reflect padding refers to
  int n_frames;             // frames written per row (T, or the window length when streaming)
  int t_begin;              // absolute index of the first frame written (0 offline)
  int src_base;             // virtual sample index of wav[row][0] (0 offline; > 0 when only a tail of the row is resident)
  int tiles_per_row;
  int n_tiles;
  int hop;
  int pad_inner;            // (n_fft - hop)/2 reflect pad of the reference
  int pad_outer;            // n_fft/2 when center=True, else 0
  int n_mels;
  int n_chan_pad;           // n_mels rounded up to the channel-group size 32/TF
  int wave_len;             // staged samples per tile (multiple of 4)
  int nnz;                  // banded weights
  const float* window;      // (n_fft)
  const float2* stage_tw;   // n_fft 1024: [16][32] W_512^{k1*n2}; n_fft 2048: [32][32] W_1024^{k1*n2}
  const float2* fold_tw;    // [n_fft/4 + 1]: W_{n_fft}^k
  const int2* chan;         // (n_chan_pad) {first bin | count << 16 (both mult. of 4, count equal within a group), weight offset}
  const float* weights;
  const int* lengths;       // valid samples per row, or null
  float* logmel;            // (B, M, T)            [kOutLogmel]; __nv_bfloat16 with kOutBf16
  int mask_invalid;         // log-mel of frames at or past lengths[b] / hop is written as 0 (the caller's mel * mask)
  float* row_sum;           // (B, M) or null: += sum over the frames of each row and channel of the log-mel as written [kOutLogmel]
  unsigned char* codes;     // (B, M, T)            [kOutCodes]
  const float* q_lo;        // (M)                  [kOutCodes]
  const float* q_scale;     // (M)  K / (hi - lo)   [kOutCodes]
  const float* q_step;      // (M)  (hi - lo) / K   [kOutDequant]
  float* dequant;           // (B, M, T) lo + (code + 0.5) * step, 0 past the valid frames [kOutDequant]
  int n_bins;
  float kmax;               // float(n_bins - 1)
  // byte offsets of the shared-memory regions (FusedLayout, filled in by the host so the kernel
  // does no layout arithmetic)
  int off_mags, off_wave, off_window, off_fold, off_chan, off_weights, off_perchan, off_bars;
  int* sched;               // {next dynamic tile, finished CTAs}, both 0 between launches; null = static tile walk
  int debug_skip;           // diagnostics (env DMEL_DEBUG_SKIP): 1 skip the FFT phase, 2 skip mel/epilogue, 4 skip staging
  float* run_min;           // (M) running min, updated in place [kOutStats]
  float* run_max;           // (M)
  unsigned long long* near_edge;  // [kOutEdge]
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

// ---- programmatic dependent launch -------------------------------------------------
// launch_dependents: the next kernel of the stream (if it was launched with the programmatic-serialisation
// attribute) may start once every CTA of this grid has passed this point; grid_dependency_wait: block until the
// previous kernel of the stream has completed and its writes are visible.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier + bulk async copy (TMA engine, 1-D) ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// OCC = CTAs per SM the instantiation is built for.  OCC == 3 (n_fft 1024, TF 8 only) trades
// registers and shared memory for a third CTA: window taps and twiddles are not kept in
// registers (85-register budget) and the waveform tile is single-buffered.
template <int NFFT, int TF, int OCC>
struct FusedLayout {
  // n_fft 2048 with OCC == 2 runs each frame as two 512-point FFTs (even / odd samples) on the
  // register-lean core; with OCC == 1 it is the 32-points-per-lane form (fallback when the lean
  // layout does not fit twice into an SM).
  static constexpr bool kSplit2048 = NFFT == 2048 && OCC == 2;
  static constexpr int kWaveBufs = (OCC == 3 || kSplit2048) ? 1 : 2;
  static constexpr bool kWindowInSmem = NFFT == 2048 || OCC == 3;
  static constexpr int kBins = NFFT / 2 + 1;
  // row pitch 516 / 1028 floats: a multiple of 4 so a lane can fetch four bins of its frame with
  // one LDS.128, and == 4 (mod 32) so the eight lanes of a quarter-warp (eight frames) cover all
  // 32 banks.  The host keeps every padded span inside its row.
  static constexpr int kMagPitch = kBins + 3;
  static constexpr int kMagFloats = TF * kMagPitch;
  static constexpr int kTileF2 = (NFFT == 1024 || kSplit2048) ? kTile512 : kTile1024;
  static constexpr int kFoldN = NFFT / 4 + 1;
  // byte offsets inside dynamic shared memory (all 16-byte aligned)
  static __host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
  static __host__ __device__ size_t tiles_off() { return 0; }
  static __host__ __device__ size_t mags_off() { return size_t(kWarps) * kTileF2 * sizeof(float2); }
  static __host__ __device__ size_t wave_off() { return align16(mags_off() + size_t(kMagFloats) * 4); }
  static __host__ __device__ size_t window_off(int wave_len) { return align16(wave_off() + kWaveBufs * size_t(wave_len) * 4); }
  static __host__ __device__ size_t fold_off(int wave_len) {
    return align16(window_off(wave_len) + (kWindowInSmem ? size_t(NFFT) * 4 : 0));
  }
  static __host__ __device__ size_t chan_off(int wave_len) {
    return align16(fold_off(wave_len) + ((NFFT == 2048 && !kSplit2048) ? size_t(kFoldN) * 8 : 0));
  }
  // n_chan = channel count padded to the group size
  static __host__ __device__ size_t weights_off(int wave_len, int n_chan) {
    return align16(chan_off(wave_len) + size_t(n_chan) * 8);
  }
  static __host__ __device__ size_t perchan_off(int wave_len, int n_chan, int nnz) {
    return align16(weights_off(wave_len, n_chan) + size_t(nnz) * 4);
  }
  static __host__ __device__ size_t bar_off(int wave_len, int n_chan, int nnz) {
    return align16(perchan_off(wave_len, n_chan, nnz) + size_t(n_chan) * 12);  // {lo, scale, step} or {min, max}
  }
  static __host__ __device__ size_t total(int wave_len, int n_chan, int nnz) {
    return bar_off(wave_len, n_chan, nnz) + 32;  // two mbarriers + the next-tile slot
  }
};

// Everything the tile loop needs to know about one tile; CTA-uniform.
struct TileInfo {
  int row, t0, n_valid;
  int frame_limit;  // frames of this tile worth computing (t0-relative), 0 = skip the tile
  // staging: wave[bulk_lo, bulk_lo + bulk_n) comes from one bulk async copy (the samples that exist and
  // need no reflection), the rest - nothing for interior tiles - from plain reflect-indexed loads
  int bulk_lo, bulk_n;
  bool manual;      // some of the tile is staged by plain loads (row ends, unaligned rows)
  long long src0;   // flat waveform offset of wave[0]'s sample (valid inside the bulk range)
};

template <int NFFT, int TF, int MODE, int OCC>
__global__ void __launch_bounds__(kThreads, OCC) dmel_fused_kernel(const FusedParams p) {
  static_assert(OCC == 1 || (OCC == 2 && (NFFT == 1024 || TF == 8)) || (OCC == 3 && NFFT == 1024 && TF == 8),
                "occupancy variants");
  static_assert(NFFT == 1024 || NFFT == 2048, "register FFT cores: 512 and 1024 complex points");
  static_assert(TF == 32 || TF == 16 || TF == 8, "tile frames");
  using LY = FusedLayout<NFFT, TF, OCC>;
  constexpr bool kSplit = LY::kSplit2048;
  constexpr bool kLean = OCC == 3 || kSplit;  // register-lean variants: twiddles rebuilt, window in smem
  constexpr int kPitch = LY::kMagPitch;
  constexpr bool kCodes = (MODE & kOutCodes) != 0, kLogmel = (MODE & kOutLogmel) != 0;
  constexpr bool kStats = (MODE & kOutStats) != 0, kEdge = (MODE & kOutEdge) != 0;
  constexpr bool kBf16 = (MODE & kOutBf16) != 0;
  constexpr bool kPcm = (MODE & kInPcm16) != 0;
  constexpr bool kDequant = (MODE & kOutDequant) != 0;
  static_assert(!kDequant || kCodes, "kOutDequant qualifies the code output");
  static_assert(!kPcm || kLean, "int16 input is built for the register-lean variants");
  using wave_t = std::conditional_t<kPcm, short, float>;
  constexpr int kAlign = 16 / (int)sizeof(wave_t);  // samples per 16 bytes: granularity of the bulk copies
  const wave_t* wav = reinterpret_cast<const wave_t*>(p.wav);
  static_assert(!kBf16 || kLogmel, "kOutBf16 qualifies the log-mel output");
  auto store_logmel = [&](size_t o, float v) {
    if constexpr (kBf16) reinterpret_cast<__nv_bfloat16*>(p.logmel)[o] = __float2bfloat16_rn(v);
    else p.logmel[o] = v;
  };

  extern __shared__ __align__(16) unsigned char smem[];
  float2* tiles = reinterpret_cast<float2*>(smem);
  float* mags = reinterpret_cast<float*>(smem + p.off_mags);
  wave_t* wave0 = reinterpret_cast<wave_t*>(smem + p.off_wave);
  float* s_window = reinterpret_cast<float*>(smem + p.off_window);
  float2* s_fold = reinterpret_cast<float2*>(smem + p.off_fold);
  int2* s_chan = reinterpret_cast<int2*>(smem + p.off_chan);
  float* s_weights = reinterpret_cast<float*>(smem + p.off_weights);
  // two per-channel arrays: quantiser {lo, scale} when writing codes, running {min, max} when
  // calibrating (no launch does both)
  static_assert(!(kCodes && kStats), "codes and statistics share their per-channel scratch");
  float* s_lo = reinterpret_cast<float*>(smem + p.off_perchan);
  float* s_scale = s_lo + p.n_chan_pad;
  float* s_step = s_scale + p.n_chan_pad;
  float* s_min = s_lo;
  float* s_max = s_scale;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  float2* my_tile = tiles + warp * LY::kTileF2;
  grid_launch_dependents();

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }

  // per-lane constants kept in registers for the whole kernel
  constexpr int kPts = (NFFT == 1024 || kSplit) ? 16 : 32;  // complex points per lane
  constexpr int kTwRegs = kLean ? 1 : kPts;
  float2 tw[kTwRegs];                           // inter-pass twiddles W_{NFFT/2}^{lane*k1}
  if constexpr (!kLean) {
#pragma unroll
    for (int k1 = 0; k1 < kPts; ++k1) tw[k1] = p.stage_tw[k1 * 32 + lane];
  }
  // lean variants: only W_512^{lane*{1,2,4,8}} stay resident, the other eleven are rebuilt per frame
  // (the n_fft 2048 table holds W_1024^{k1*lane}, so W_512^{lane*k} sits at row 2k)
  constexpr int kRow = kSplit ? 2 : 1;
  const float2 w1 = p.stage_tw[1 * kRow * 32 + lane], w2 = p.stage_tw[2 * kRow * 32 + lane];
  const float2 w4 = p.stage_tw[4 * kRow * 32 + lane], w8 = p.stage_tw[8 * kRow * 32 + lane];
  float2 win[(NFFT == 1024 && !kLean) ? 16 : 1];  // window taps of this lane's samples
  float2 fold_base = make_float2(1.f, 0.f);     // W_1024^lane
  if constexpr (NFFT == 1024) {
    if constexpr (!kLean) {
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int idx = 2 * (32 * n1 + lane);
        win[n1] = make_float2(p.window[idx], p.window[idx + 1]);
      }
    }
    fold_base = p.fold_tw[lane];
  }
  float2 base1024 = make_float2(1.f, 0.f), base2048 = make_float2(1.f, 0.f);  // W_1024^lane, W_2048^lane
  if constexpr (kSplit) {
    base1024 = p.stage_tw[1 * 32 + lane];
    base2048 = p.fold_tw[lane];
  }
  const float2* my_win = reinterpret_cast<const float2*>(s_window) + lane;  // lean / 2048: taps re-read per frame

  const int padded_len = p.n_samples + 2 * p.pad_inner + 2 * p.pad_outer;
  const bool hop_even = (p.hop & 1) == 0;
  const bool row_vec_ok = ((reinterpret_cast<uintptr_t>(p.wav) & 15) == 0) && ((p.row_stride & (kAlign - 1)) == 0);
  unsigned long long edge_hits = 0;
  uint32_t phase_bits = 0;  // bit b: parity to wait for on bars[b]

  auto describe = [&](int tile) {
    TileInfo ti;
    ti.row = tile / p.tiles_per_row;
    ti.t0 = (tile - ti.row * p.tiles_per_row) * TF;
    ti.n_valid = p.n_frames;  // in frames of this launch's window
    if (p.lengths) {
      const int nv = p.lengths[ti.row] / p.hop - p.t_begin;
      ti.n_valid = nv < 0 ? 0 : (nv < p.n_frames ? nv : p.n_frames);
    }
    // log-mel output covers every frame of the row (unless masked); codes / statistics only the valid ones
    const int last = ((kLogmel && !p.mask_invalid) ? p.n_frames : ti.n_valid) - ti.t0;
    ti.frame_limit = last < 0 ? 0 : (last > TF ? TF : last);
    const int s0 = (p.t_begin + ti.t0) * p.hop - p.pad_inner - p.pad_outer;  // virtual sample under the tile's first tap
    const int b0 = s0 - p.src_base;                                          // where that sample sits in the buffer
    ti.src0 = (long long)ti.row * p.row_stride + b0;
    constexpr int kA = kAlign - 1;
    int lo = s0 < 0 ? ((-s0 + kA) & ~kA) : 0;                                 // first wave index with a real sample
    if (b0 + lo < 0) lo = (-b0 + kA) & ~kA;                                   // ... that is resident in the buffer
    int hi = p.n_samples - s0 < p.wave_len ? ((p.n_samples - s0) & ~kA) : p.wave_len;  // one past the last
    const bool can_bulk = row_vec_ok && p.pad_outer == 0 && (b0 & kA) == 0 && hi > lo;
    ti.bulk_lo = can_bulk ? lo : 0;
    ti.bulk_n = can_bulk ? hi - lo : 0;
    ti.manual = !can_bulk || lo > 0 || hi < p.wave_len;
    return ti;
  };
  // Start filling wave buffer b with the samples of a tile.
  auto stage = [&](const TileInfo& ti, int b) {
    if (ti.frame_limit == 0 || (p.debug_skip & 4)) return;
    wave_t* wave = wave0 + b * p.wave_len;
    if (ti.bulk_n && tid == 0) {
      fence_proxy_async();  // earlier generic-proxy reads of this buffer are ordered before the async write
      mbar_expect_tx(&bars[b], ti.bulk_n * (int)sizeof(wave_t));
      bulk_copy_g2s(wave + ti.bulk_lo, wav + ti.src0 + ti.bulk_lo, ti.bulk_n * (int)sizeof(wave_t), &bars[b]);
    }
    if (ti.manual) {
      const wave_t* src = wav + (long long)ti.row * p.row_stride - p.src_base;
      const int j0 = (p.t_begin + ti.t0) * p.hop;  // first position in the padded row
      const int skip_lo = ti.bulk_lo, skip_hi = ti.bulk_lo + ti.bulk_n;
      const int n_manual = p.wave_len - ti.bulk_n;
      for (int q = tid; q < n_manual; q += kThreads) {
        const int i = q < skip_lo ? q : q + (skip_hi - skip_lo);  // wave index outside the bulk range
        const int j = j0 + i;
        wave_t x = 0;
        if (j < padded_len) x = __ldg(src + reflect_src(j, p.n_samples, p.pad_inner, p.pad_outer));
        wave[i] = x;
      }
    }
  };

  // ---- per-CTA constants -------------------------------------------------
  // Everything up to grid_dependency_wait() reads plan-owned memory only (written at plan creation), so under
  // programmatic dependent launch it overlaps the tail of the previous kernel in the stream.
  for (int i = tid; i < p.n_chan_pad; i += kThreads) {
    s_chan[i] = p.chan[i];
    if constexpr (!kCodes) {
      s_min[i] = __int_as_float(0x7f800000);
      s_max[i] = __int_as_float(0xff800000);
    }
  }
  for (int i = tid; i < p.nnz; i += kThreads) s_weights[i] = p.weights[i];
  for (int i = tid; i < LY::kMagFloats; i += kThreads) mags[i] = 0.f;  // the 3 pad columns of each row stay zero
  if constexpr (LY::kWindowInSmem) {
    for (int i = tid; i < NFFT; i += kThreads) s_window[i] = p.window[i];
  }
  if constexpr (NFFT == 2048 && !kSplit) {
    for (int i = tid; i < LY::kFoldN; i += kThreads) s_fold[i] = p.fold_tw[i];
  }
  grid_dependency_wait();  // from here on: the caller's tensors (waveform, lengths, statistics, outputs, tile counter)
  int tile = blockIdx.x;
  TileInfo cur = describe(tile < p.n_tiles ? tile : 0);
  if (tile < p.n_tiles) stage(cur, 0);  // the first tile leaves HBM while the statistics below are fetched
  if constexpr (kCodes) {
    for (int i = tid; i < p.n_chan_pad; i += kThreads) {
      const bool real = i < p.n_mels;
      s_lo[i] = real ? p.q_lo[i] : 0.f;
      s_scale[i] = real ? p.q_scale[i] : 0.f;
      if constexpr (kDequant) s_step[i] = real ? p.q_step[i] : 0.f;
    }
  }
  __syncthreads();  // constants + barrier init visible
  // Tiles after the first are handed out by a global counter (lean variants), so the CTAs of the
  // grid finish within one tile of each other instead of one or two tiles apart.
  constexpr bool kDynamic = LY::kWaveBufs == 1;
  int* s_next = reinterpret_cast<int*>(bars + 2);
  const bool dynamic = kDynamic && p.sched != nullptr;

  for (int it = 0; tile < p.n_tiles; ++it) {
    const int b = LY::kWaveBufs == 2 ? (it & 1) : 0;
    const wave_t* wave = wave0 + b * p.wave_len;
    const bool dead = cur.frame_limit == 0;

    // ---- 1. this tile's samples are in wave[b]; start fetching the next tile
    if (!dead && !(p.debug_skip & 4)) {
      if (cur.bulk_n) {
        mbar_wait(&bars[b], (phase_bits >> b) & 1u);
        phase_bits ^= 1u << b;
      }
      if (cur.manual) __syncthreads();  // plain stores of all threads
    }
    int next_tile = tile + (int)gridDim.x;
    int fetched = 0;
    if (dynamic && tid == 0) fetched = atomicAdd(p.sched, 1);  // consumed just before the barrier below
    bool has_next = next_tile < p.n_tiles;
    TileInfo nxt = cur;
    if constexpr (!kDynamic) {
      nxt = describe(has_next ? next_tile : tile);
      if constexpr (LY::kWaveBufs == 2) {
        if (has_next) stage(nxt, b ^ 1);  // double buffered: the next tile loads while this one computes
      }
    }

    // ---- 2. FFT -> magnitudes ------------------------------------------------
    const int fft_frames = (p.debug_skip & 1) ? 0 : cur.frame_limit;
    if constexpr (NFFT == 1024) {
      const int h = lane >> 4;
      const int partner = mirror_lane512(lane);
#pragma unroll 1
      for (int fr = warp; fr < fft_frames; fr += kWarps) {
        const wave_t* fa = wave + fr * p.hop;
        float2 v[16];
        if (hop_even) {
          using pair_t = std::conditional_t<kPcm, short2, float2>;
          const pair_t* f2 = reinterpret_cast<const pair_t*>(fa);
#pragma unroll
          for (int n1 = 0; n1 < 16; ++n1) {
            const pair_t x = f2[32 * n1 + lane];
            const float2 xf = make_float2((float)x.x, (float)x.y);
            if constexpr (kLean) v[n1] = f2_mul(xf, my_win[32 * n1]);
            else v[n1] = f2_mul(xf, win[n1]);
          }
        } else {
#pragma unroll
          for (int n1 = 0; n1 < 16; ++n1) {
            const int idx = 2 * (32 * n1 + lane);
            const float2 xf = make_float2((float)fa[idx], (float)fa[idx + 1]);
            if constexpr (kLean) v[n1] = f2_mul(xf, my_win[32 * n1]);
            else v[n1] = f2_mul(xf, win[n1]);
          }
        }
        __syncwarp();  // previous frame's pass-2 reads of my_tile are done
        if constexpr (kLean) fft512_pass1_pow(v, w1, w2, w4, w8, my_tile, lane);
        else fft512_pass1(v, tw, my_tile, lane);
        __syncwarp();
        fft512_pass2(v, my_tile, lane);
        float2 send[8], recv[8], zlo[8], zhi[8];
        combine_send(v, h, send);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          recv[j] = make_float2(__shfl_xor_sync(0xffffffffu, send[j].x, 16), __shfl_xor_sync(0xffffffffu, send[j].y, 16));
        combine_finish(v, recv, h, zlo, zhi);
        mirror_send512(zlo, zhi, lane, send);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          recv[j] = make_float2(__shfl_sync(0xffffffffu, send[j].x, partner), __shfl_sync(0xffffffffu, send[j].y, partner));
        unfold_store512(zlo, zhi, recv, fold_base, mags + fr * kPitch, lane);
      }
    } else if constexpr (kSplit) {
      const int h = lane >> 4;
      const int partner = mirror_lane512(lane);
      // one 512-point FFT of v, unfolded into the half spectrum of a real 1024-sequence
      auto half = [&](float2 (&v)[16], HalfSpectrum& out) {
        __syncwarp();
        fft512_pass1_pow(v, w1, w2, w4, w8, my_tile, lane);
        __syncwarp();
        fft512_pass2(v, my_tile, lane);
        float2 send[8], recv[8], zlo[8], zhi[8];
        combine_send(v, h, send);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          recv[j] = make_float2(__shfl_xor_sync(0xffffffffu, send[j].x, 16), __shfl_xor_sync(0xffffffffu, send[j].y, 16));
        combine_finish(v, recv, h, zlo, zhi);
        mirror_send512(zlo, zhi, lane, send);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          recv[j] = make_float2(__shfl_sync(0xffffffffu, send[j].x, partner), __shfl_sync(0xffffffffu, send[j].y, partner));
        unfold_half_spectrum(zlo, zhi, recv, base1024, out);
      };
      const bool hop_vec = (p.hop & 3) == 0;
      const float4* win4 = reinterpret_cast<const float4*>(s_window) + lane;
#pragma unroll 1
      for (int fr = warp; fr < fft_frames; fr += kWarps) {
        const wave_t* fa = wave + fr * p.hop;
        float2 v[16], odd[16];  // even samples x[4n], x[4n+2] and odd samples x[4n+1], x[4n+3] of this lane
        if (hop_vec) {
          using quad_t = std::conditional_t<kPcm, short4, float4>;
          const quad_t* f4 = reinterpret_cast<const quad_t*>(fa) + lane;
#pragma unroll
          for (int n1 = 0; n1 < 16; ++n1) {
            const quad_t x = f4[32 * n1];
            const float4 w = win4[32 * n1];
            v[n1] = make_float2((float)x.x * w.x, (float)x.z * w.z);
            odd[n1] = make_float2((float)x.y * w.y, (float)x.w * w.w);
          }
        } else {
#pragma unroll
          for (int n1 = 0; n1 < 16; ++n1) {
            const int idx = 4 * (32 * n1 + lane);
            const float4 w = win4[32 * n1];
            v[n1] = make_float2((float)fa[idx] * w.x, (float)fa[idx + 2] * w.z);
            odd[n1] = make_float2((float)fa[idx + 1] * w.y, (float)fa[idx + 3] * w.w);
          }
        }
        HalfSpectrum e, o;
        half(v, e);
        half(odd, o);
        combine2048_store(e, o, base2048, mags + fr * kPitch, lane);
      }
    } else {
#pragma unroll 1
      for (int fr = warp; fr < fft_frames; fr += kWarps) {
        const float* fa = reinterpret_cast<const float*>(wave) + fr * p.hop;  // never int16: not a lean variant
        float2 v[32];
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          const int idx = 2 * (32 * n1 + lane);
          v[n1] = f2_mul(make_float2(fa[idx], fa[idx + 1]), my_win[32 * n1]);
        }
        __syncwarp();
        fft1024_pass1(v, tw, my_tile, lane);
        __syncwarp();
        fft1024_pass2(v, my_tile, lane);
        float* mrow = mags + fr * kPitch;
        const int partner = (32 - lane) & 31;
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const float2 send =
              (lane == 0) ? v[brev5(mirror_slot1024(k2, true))] : v[brev5(mirror_slot1024(k2, false))];
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
    if constexpr (kDynamic) {
      if (dynamic && tid == 0) *s_next = (int)gridDim.x + fetched;
    }
    __syncthreads();
    if constexpr (kDynamic) {
      if (dynamic) next_tile = *s_next;
      has_next = next_tile < p.n_tiles;
      nxt = describe(has_next ? next_tile : tile);
    }
    if constexpr (LY::kWaveBufs == 1) {
      if (has_next) stage(nxt, 0);  // single buffer: it is free now, the copy flies under the mel phase
    }

    // ---- 3. mel filterbank, log, quantise --------------------------------
    // One lane per frame, 32/TF adjacent channels side by side in a warp.  The host pads the spans
    // of such a channel group to one common length, so the bin loop is warp-uniform.
    {
      constexpr int kGroups = 32 / TF;
      constexpr int kStep = kWarps * kGroups;  // channels between two trips of a lane
      const int fr = lane % TF;
      const int sub = lane / TF;
      const int t = cur.t0 + fr;
      const bool in_row = t < p.n_frames;
      const bool valid = t < cur.n_valid;
      const float* mrow = mags + fr * kPitch;
      const int m0 = warp * kGroups + sub;
      const size_t ostep = (size_t)kStep * p.n_frames;
      size_t o = ((size_t)cur.row * p.n_mels + m0) * p.n_frames + t;
      if (p.debug_skip & 2) {
      } else if (dead) {
        // nothing of this tile is valid audio: codes are the pad value, masked log-mel is zero
        if constexpr (kCodes || kLogmel) {
#pragma unroll 1
          for (int m = m0; m < p.n_mels; m += kStep, o += ostep)
            if (in_row) {
              if constexpr (kCodes) p.codes[o] = 0;
              if constexpr (kDequant) p.dequant[o] = 0.f;
              if constexpr (kLogmel) store_logmel(o, 0.f);
            }
        }
      } else {
        const int2* cp = s_chan + m0;
        const float* lop = s_lo + m0;
        const float* scp = s_scale + m0;
#pragma unroll 1
        for (int mb = warp * kGroups; mb < p.n_chan_pad; mb += kStep, cp += kStep, lop += kStep, scp += kStep, o += ostep) {
          const int m = mb + sub;
          const bool live = m < p.n_mels;
          const int2 c = *cp;
          const float4* w4 = reinterpret_cast<const float4*>(s_weights + c.y);
          const float4* x4 = reinterpret_cast<const float4*>(mrow + (c.x & 0xffff));
          const float4* x4_end = x4 + (c.x >> 18);  // span length / 4: >= 1, identical across the warp
          float acc = 0.f;
#pragma unroll 1
          do {
            const float4 w = *w4++;
            const float4 x = *x4++;
            acc = fmaf(w.x, x.x, acc);
            acc = fmaf(w.y, x.y, acc);
            acc = fmaf(w.z, x.z, acc);
            acc = fmaf(w.w, x.w, acc);
          } while (x4 != x4_end);
          const float value = fast_log(fmaxf(acc, kLogClip));
          if constexpr (kLogmel) {
            const float out = (p.mask_invalid && !valid) ? 0.f : value;
            if (live && in_row) store_logmel(o, out);
            if (p.row_sum) {  // per (row, channel) sum over time: the caller's mels.mean(-1) without another pass
              float part = (live && in_row) ? out : 0.f;
#pragma unroll
              for (int d = TF / 2; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
              if (fr == 0 && live) atomicAdd(p.row_sum + (size_t)cur.row * p.n_mels + m, part);
            }
          }
          if constexpr (kCodes) {
            const float sc = *scp;
            const float pos = __fmul_rn(__fsub_rn(value, *lop), sc);
            const float q = fminf(fmaxf(floorf(pos), 0.f), p.kmax);
            if (live && in_row) p.codes[o] = valid ? (unsigned char)q : (unsigned char)0;
            if constexpr (kDequant) {  // the table entry the stand-alone decoder would look up, same two roundings
              const float centre = __fadd_rn(*lop, __fmul_rn(q + 0.5f, s_step[m]));
              if (live && in_row) p.dequant[o] = valid ? centre : 0.f;
            }
            if constexpr (kEdge) {
              const float e = fminf(fmaxf(rintf(pos), 1.f), p.kmax);
              if (live && valid && fabsf(pos - e) < p.edge_eps * sc) ++edge_hits;
            }
          }
          if constexpr (kStats) {
            float lo = (valid && live) ? value : __int_as_float(0x7f800000);
            float hi = (valid && live) ? value : __int_as_float(0xff800000);
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
    }
    __syncthreads();  // mags and wave[b] are free again
    cur = nxt;
    tile = next_tile;
  }
  if constexpr (kDynamic) {
    // the last CTA to finish leaves both counters at zero for the next launch
    if (dynamic && tid == 0) {
      __threadfence();
      if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
        atomicExch(p.sched, 0);
        atomicExch(p.sched + 1, 0);
      }
    }
  }

  // ---- flush per-CTA statistics -------------------------------------------
  if constexpr (kStats) {
    for (int m = tid; m < p.n_mels; m += kThreads) {
      const float lo = s_min[m], hi = s_max[m];
      if (lo <= hi) {
        atomic_min_float(p.run_min + m, lo);
        atomic_max_float(p.run_max + m, hi);
      }
    }
  }
  if constexpr (kEdge) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) edge_hits += __shfl_xor_sync(0xffffffffu, edge_hits, d);
    if (lane == 0 && edge_hits) atomicAdd(p.near_edge, edge_hits);
  }
}

}  // namespace dmel
