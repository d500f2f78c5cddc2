// Fused waveform -> log-mel -> dMel codes kernel (sm_100a).
//
// One CTA owns tiles of TF consecutive frames of one utterance row:
//   1. the (TF-1)*hop + n_fft samples of the NEXT tile are pulled into shared
//      memory by one bulk async copy (cp.async.bulk + mbarrier) while the
//      current tile computes; row ends, where the reference reflect-pads
//      (utils/spectrogram.py:58-62), are staged by ordinary reflect-indexed
//      loads.  The 4x frame overlap is re-read on chip, never from HBM;
//   2. each warp turns frames into magnitudes with the register FFT of
//      fft_core.cuh (window multiply on load, torch.stft at :64-75, magnitude
//      at :76).  With one frame per warp and tile (TF == 8) the magnitudes
//      overwrite the warp's own transpose tile, so they need no storage of
//      their own;
//   3. the banded mel filterbank (:78), log(clamp(.,1e-5)) (:38-39) and the
//      per-channel bin quantiser run with one lane per
//      frame, and only codes / log-mel / min-max leave the SM.
//
// HBM traffic per tile: hop*TF*4 B of new waveform in, n_mels*TF B of codes
// out; constants come from L2.
//
// Tile bookkeeping (which row, which frames, what to copy) is done by ONE
// thread per CTA, which leaves the description of the next tile in a shared
// slot; divisions by run-time constants are multiply-high (fastdiv.cuh).
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

#include "fastdiv.cuh"
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

// one filterbank channel as the mel phase reads it (32 bytes in shared memory: two 16-byte loads)
struct ChanRec {
  int x_off;     // byte offset of the first bin of the span inside a magnitude row (multiple of 16)
  int w_off;     // byte offset of the span's weights inside the banded weight array (multiple of 16)
  int span;      // bytes of magnitudes in the span (multiple of 16, equal within a channel group, >= 16)
  int pad;
  float a, b, c, d;  // quantiser {lo, scale, step, -} when writing codes; running {min, max} when calibrating
};
static_assert(sizeof(ChanRec) == 32, "two 16-byte loads");

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
  int wave_len;             // staged samples per tile (multiple of 8)
  int nnz;                  // banded weights
  int bulk_ok;              // waveform base and row stride are 16-byte aligned and center=False: tiles can be bulk-copied
  int n_order;              // entries of group_order (a multiple of the warp count)
  FastDiv by_tiles_per_row; // tile -> row
  FastDiv by_hop;           // length -> valid frames
  const float* window;      // (n_fft)
  const float2* stage_tw;   // n_fft 1024: [16][32] W_512^{k1*n2}; n_fft 2048: [32][32] W_1024^{k1*n2}
  const float2* fold_tw;    // [n_fft/4 + 1]: W_{n_fft}^k
  const int4* chan;         // (n_chan_pad) {x_off, w_off, span, 0}: the integer half of ChanRec
  const int* group_order;   // (n_order) channel group of warp w in its r-th trip at [r * warps + w], -1 = none: balances the mel phase
  const float* weights;
  const int* lengths;       // valid samples per row, or null
  // ---- input-side options (reference dataset/lhotse_tts_dataset.py:29-33 and :46-65 do these on the host) ----
  const long long* offsets; // ragged batch: row b is wav[offsets[b], offsets[b + 1]) (n_rows + 1 entries; only its first lengths[b]
                            // samples when lengths is given too), row_stride unused; or null
  int own_length;           // 1: every row is an utterance of its own length (lengths[b], or the offsets' difference):
                            //    the reflect padding refers to THAT length, as if the reference ran on the row alone
  const float* row_gain;    // (B) gain applied to the samples of a row (per-utterance peak normalisation), or null; float32 input only
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
  int off_mags, off_wave, off_window, off_fold, off_rec, off_weights, off_order, off_bars;
  int* sched;               // {next dynamic tile, finished CTAs}, both 0 between launches; null = static tile walk
  int debug_skip;           // diagnostics (env DMEL_DEBUG_SKIP, builds with -DDMEL_ABLATION only): 1 skip the FFT phase, 2 skip mel/epilogue, 4 skip staging
  float* run_min;           // (M) running min, updated in place [kOutStats]
  float* run_max;           // (M)
  unsigned long long* near_edge;  // [kOutEdge]
  float edge_eps;
};

#ifdef DMEL_ABLATION
#define DMEL_SKIP(p, bit) (((p).debug_skip & (bit)) != 0)
#else
#define DMEL_SKIP(p, bit) false
#endif

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

// Order-preserving integer atomics on float bit patterns.  The branch is on the SIGN BIT, so -0.0f takes the
// unsigned path (as a plain `v >= 0` would not: its pattern 0x80000000 is INT_MIN and would beat every negative
// minimum); NaN is dropped, it has no place in an order.
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (v != v) return;
  if (__float_as_int(v) >= 0) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v != v) return;
  if (__float_as_int(v) >= 0) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
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
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- explicit shared-space accesses on 32-bit addresses -------------------------
// The mel phase and the tile slot use these instead of generic pointers: one cvta at kernel entry, immediate
// offsets afterwards, nothing for the compiler to re-derive per use.
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds_i4(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_i4(uint32_t a, int4 v) {
  asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}

// One 512-point FFT of v (register-lean core), unfolded into the half spectrum of a real 1024-sequence: one half
// of an n_fft = 2048 frame.  A force-inlined function, not a lambda: the HalfSpectrum must stay in registers.
__device__ __forceinline__ void fft512_half_spectrum(float2 (&v)[16], float2 w1, float2 w2, float2 w4, float2 w8,
                                                     float2* my_tile, int lane, int h, int partner, float2 base1024,
                                                     HalfSpectrum& out) {
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
}

// ---- mel filterbank, log, quantise: the third phase of a tile ---------------------------------------------------
// One lane per frame and PAIR of channels (m, m + 32/TF): a group of 2 * 32/TF adjacent channels side by side in a
// warp.  The host pads the spans of a group to one common length and interleaves the weights of a pair step by step,
// so the bin loop is warp-uniform and serves two independent accumulation chains; it also deals the groups to the
// warps that run this phase (`slot` of `n_slots`) so that every one of them gets about the same number of steps.
// `mags_sa` is the 32-bit shared address of the [frame][bin] magnitudes, rows kPitch floats apart.
template <int TF, int MODE, int kPitch>
__device__ __forceinline__ void mel_phase(const FusedParams& p, uint32_t sa_base, uint32_t mags_sa, int lane, int slot, int n_slots,
                                          int cur_row, int cur_t0, int cur_valid, bool dead, unsigned long long& edge_hits) {
  constexpr bool kCodes = (MODE & kOutCodes) != 0, kLogmel = (MODE & kOutLogmel) != 0;
  constexpr bool kStats = (MODE & kOutStats) != 0, kEdge = (MODE & kOutEdge) != 0;
  constexpr bool kBf16 = (MODE & kOutBf16) != 0, kDequant = (MODE & kOutDequant) != 0;
  const int* s_order = reinterpret_cast<const int*>(__cvta_shared_to_generic(sa_base + p.off_order));
  constexpr int kGroups = 32 / TF;
  constexpr int kChanIter = 2 * kGroups;        // channels of a group
  constexpr int kRecB = kGroups * (int)sizeof(ChanRec);  // record of the pair's second channel
  const int fr = lane % TF;
  const int sub = lane / TF;
  const int t = cur_t0 + fr;
  const size_t opair = (size_t)kGroups * p.n_frames;  // from a pair's first channel to its second
  const size_t obase = (size_t)cur_row * p.n_mels * p.n_frames + t;
  auto store_logmel = [&](size_t at, float v) {
    if constexpr (kBf16) reinterpret_cast<__nv_bfloat16*>(p.logmel)[at] = __float2bfloat16_rn(v);
    else p.logmel[at] = v;
  };
  if (dead) {
    // nothing of this tile is valid audio: codes are the pad value, masked log-mel is zero
    if constexpr (kCodes || kLogmel) {
      const bool in_row = t < p.n_frames;
#pragma unroll 1
      for (int m = slot * kGroups + sub; m < p.n_mels; m += n_slots * kGroups) {
        const size_t od = obase + (size_t)m * p.n_frames;
        if (in_row) {
          if constexpr (kCodes) p.codes[od] = 0;
          if constexpr (kDequant) p.dequant[od] = 0.f;
          if constexpr (kLogmel) store_logmel(od, 0.f);
        }
      }
    }
  } else {
    const uint32_t xrow = mags_sa + (uint32_t)fr * (kPitch * 4);
    const uint32_t wbase = sa_base + p.off_weights;
    const uint32_t rec0 = sa_base + p.off_rec + (uint32_t)sub * (uint32_t)sizeof(ChanRec);
    const bool in_row = t < p.n_frames;
    const bool valid = t < cur_valid;
    // everything that happens to one channel's value.  kFull: all TF frames of the tile are valid frames of
    // the row and no phantom channel pads a group - the common case, whose epilogue carries no predicates.
    auto emit = [&]<bool kFull>(std::bool_constant<kFull>, float value, const float4& q, uint32_t rec_at, int m, size_t at) {
      const bool live = kFull || m < p.n_mels;
      const bool ok = kFull || (live && in_row);
      const bool val = kFull || valid;
      if constexpr (kLogmel) {
        const float out = (!kFull && p.mask_invalid && !val) ? 0.f : value;
        if (ok) store_logmel(at, out);
        if (p.row_sum) {  // per (row, channel) sum over time: the caller's mels.mean(-1) without another pass
          float part = ok ? out : 0.f;
#pragma unroll
          for (int d = TF / 2; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
          if (fr == 0 && live) atomicAdd(p.row_sum + (size_t)cur_row * p.n_mels + m, part);
        }
      }
      if constexpr (kCodes) {
        const float pos = __fmul_rn(__fsub_rn(value, q.x), q.y);
        const float qf = fminf(fmaxf(floorf(pos), 0.f), p.kmax);
        if (ok) p.codes[at] = val ? (unsigned char)qf : (unsigned char)0;
        if constexpr (kDequant) {  // the table entry the stand-alone decoder would look up, same two roundings
          const float centre = __fadd_rn(q.x, __fmul_rn(qf + 0.5f, q.z));
          if (ok) p.dequant[at] = val ? centre : 0.f;
        }
        if constexpr (kEdge) {
          const float e = fminf(fmaxf(rintf(pos), 1.f), p.kmax);
          if (live && val && fabsf(pos - e) < p.edge_eps * q.y) ++edge_hits;
        }
      }
      if constexpr (kStats) {
        float lo = (val && live) ? value : __int_as_float(0x7f800000);
        float hi = (val && live) ? value : __int_as_float(0xff800000);
#pragma unroll
        for (int d = TF / 2; d >= 1; d >>= 1) {
          lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
          hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
        }
        if (fr == 0 && live) sts_f2(rec_at + 16, fminf(q.x, lo), fmaxf(q.y, hi));  // channel m always belongs to this lane of this warp: no race
      }
    };
    auto sweep = [&](auto full_tag) {
#pragma unroll 1
      for (int k = slot; k < p.n_order; k += n_slots) {
        const int g = s_order[k];
        if (g < 0) continue;  // (warp-uniform)
        const int m = g * kChanIter + sub;
        const uint32_t rec = rec0 + (uint32_t)g * (kChanIter * (int)sizeof(ChanRec));
        const size_t o = obase + (size_t)m * p.n_frames;
        const int4 ca = lds_i4(rec), cb = lds_i4(rec + kRecB);
        const float4 qa = lds_f4(rec + 16), qb = lds_f4(rec + kRecB + 16);
        // banded dot products of this lane's frame with the two channels: same span length, weights interleaved
        uint32_t wa = wbase + ca.y, xa = xrow + ca.x, xb = xrow + cb.x;
        const uint32_t xe = xa + ca.z;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 1
        do {
          const float4 wA = lds_f4(wa), wB = lds_f4(wa + 16);
          const float4 xA = lds_f4(xa), xB = lds_f4(xb);
          wa += 32;
          xa += 16;
          xb += 16;
          a0 = fmaf(wA.x, xA.x, a0);
          b0 = fmaf(wB.x, xB.x, b0);
          a1 = fmaf(wA.y, xA.y, a1);
          b1 = fmaf(wB.y, xB.y, b1);
          a0 = fmaf(wA.z, xA.z, a0);
          b0 = fmaf(wB.z, xB.z, b0);
          a1 = fmaf(wA.w, xA.w, a1);
          b1 = fmaf(wB.w, xB.w, b1);
        } while (xa != xe);
        const float va = fast_log(fmaxf(a0 + a1, kLogClip));
        const float vb = fast_log(fmaxf(b0 + b1, kLogClip));
        emit(full_tag, va, qa, rec, m, o);
        emit(full_tag, vb, qb, rec + kRecB, m + kGroups, o + opair);
      }
    };
    if ((cur_t0 + TF <= cur_valid) && (p.n_chan_pad == p.n_mels)) sweep(std::true_type{});
    else sweep(std::false_type{});
  }
}

// OCC = CTAs per SM the instantiation is built for.  OCC == 3 (n_fft 1024, TF 8 only) trades
// registers for a third CTA: window taps and twiddles are not kept in registers (80-register budget).
template <int NFFT, int TF, int OCC>
struct FusedLayout {
  // n_fft 2048 with OCC == 2 runs each frame as two 512-point FFTs (even / odd samples) on the
  // register-lean core; with OCC == 1 it is the 32-points-per-lane form (fallback when the lean
  // layout does not fit twice into an SM).
  static constexpr bool kSplit2048 = NFFT == 2048 && OCC == 2;
  static constexpr int kWaveBufs = 1;  // the next tile is copied in under this tile's mel phase
  static constexpr bool kWindowInSmem = NFFT == 2048 || OCC == 3;
  static constexpr int kBins = NFFT / 2 + 1;
  static constexpr bool kTile512 = NFFT == 1024 || kSplit2048;
  // One frame per warp and tile: the frame's magnitudes replace the warp's transpose tile once the
  // second FFT pass has read it, and the tiles double as the [frame][bin] magnitude array.
  static constexpr bool kMagsInTiles = TF == kWarps && kTile512;
  // Magnitude row pitch in floats: a multiple of 4 so a lane can fetch four bins of its frame with
  // one LDS.128, and == 4 (mod 32) so the eight lanes of a quarter-warp (eight frames) cover all
  // 32 banks.  The host keeps every padded span inside its row.
  static constexpr int kTileF2 = kTile512 ? (kMagsInTiles ? 546 : ::dmel::kTile512) : kTile1024;  // 546 float2 = 1092 floats = 4 (mod 32)
  static constexpr int kMagPitch = kMagsInTiles ? 2 * kTileF2 : kBins + 3;
  static constexpr int kMagFloats = kMagsInTiles ? 0 : TF * kMagPitch;
  static_assert(kMagPitch % 4 == 0 && kMagPitch % 32 == 4, "conflict-free LDS.128 across eight frames");
  static_assert(!kMagsInTiles || kMagPitch >= kBins + 3, "a magnitude row fits the transpose tile");
  static constexpr int kFoldN = NFFT / 4 + 1;
  // byte offsets inside dynamic shared memory (all 16-byte aligned)
  static __host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
  static __host__ __device__ size_t tiles_off() { return 0; }
  static __host__ __device__ size_t mags_off() { return kMagsInTiles ? 0 : size_t(kWarps) * kTileF2 * sizeof(float2); }
  static __host__ __device__ size_t wave_off() { return align16(size_t(kWarps) * kTileF2 * sizeof(float2) + size_t(kMagFloats) * 4); }
  static __host__ __device__ size_t window_off(int wave_len) { return align16(wave_off() + kWaveBufs * size_t(wave_len) * 4); }
  static __host__ __device__ size_t fold_off(int wave_len) {
    return align16(window_off(wave_len) + (kWindowInSmem ? size_t(NFFT) * 4 : 0));
  }
  static __host__ __device__ size_t weights_off(int wave_len) {
    return align16(fold_off(wave_len) + ((NFFT == 2048 && !kSplit2048) ? size_t(kFoldN) * 8 : 0));
  }
  // n_chan = channel count padded to the group size.  The records follow the weights, so the mel loop's
  // one-step read-ahead past the last span stays inside the allocation.
  static __host__ __device__ size_t rec_off(int wave_len, int nnz) { return align16(weights_off(wave_len) + size_t(nnz) * 4); }
  static __host__ __device__ size_t order_off(int wave_len, int n_chan, int nnz) {
    return align16(rec_off(wave_len, nnz) + size_t(n_chan) * sizeof(ChanRec));
  }
  // n_order = entries of the group-order table
  static __host__ __device__ size_t bar_off(int wave_len, int n_chan, int nnz, int n_order) {
    return align16(order_off(wave_len, n_chan, nnz) + size_t(n_order) * 4);
  }
  static __host__ __device__ size_t total(int wave_len, int n_chan, int nnz, int n_order) {
    return bar_off(wave_len, n_chan, nnz, n_order) + 48;  // one mbarrier (16 B with padding), the next tile's description (32 B)
  }
};

// Everything the tile loop needs to know about one tile; CTA-uniform.
struct TileInfo {
  int row, t0, n_valid;
  int frame_limit;  // frames of this tile worth computing (t0-relative), 0 = skip the tile
  // staging: wave[bulk_lo, bulk_lo + bulk_n) comes from one bulk async copy (the samples that exist and
  // need no reflection), the rest - nothing for interior tiles - from plain reflect-indexed loads
  int bulk_lo, bulk_n;
  int manual;       // some of the tile is staged by plain loads (row ends, unaligned rows)
  long long src0;   // flat waveform offset of wave[0]'s sample (valid inside the bulk range)
};

template <int NFFT, int TF, int MODE, int OCC>
__global__ void __launch_bounds__(kThreads, OCC) dmel_fused_kernel(const __grid_constant__ FusedParams p) {
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
  static_assert(!(kCodes && kStats), "codes and statistics share the float half of the channel records");
  static_assert(!kBf16 || kLogmel, "kOutBf16 qualifies the log-mel output");
  using wave_t = std::conditional_t<kPcm, short, float>;
  constexpr int kAlign = 16 / (int)sizeof(wave_t);  // samples per 16 bytes: granularity of the bulk copies
  const wave_t* wav = reinterpret_cast<const wave_t*>(p.wav);

  extern __shared__ __align__(16) unsigned char smem[];
  float2* tiles = reinterpret_cast<float2*>(smem);
  float* mags = reinterpret_cast<float*>(smem + p.off_mags);
  wave_t* wave0 = reinterpret_cast<wave_t*>(smem + p.off_wave);
  float* s_window = reinterpret_cast<float*>(smem + p.off_window);
  float2* s_fold = reinterpret_cast<float2*>(smem + p.off_fold);
  float* s_weights = reinterpret_cast<float*>(smem + p.off_weights);
  ChanRec* s_rec = reinterpret_cast<ChanRec*>(smem + p.off_rec);
  int* s_order = reinterpret_cast<int*>(smem + p.off_order);
  // 32-bit shared addresses of the regions the mel phase and the bookkeeping touch
  const uint32_t sa_base = smem_u32(smem);
  const uint32_t sa_bar = sa_base + p.off_bars;   // bars[0], bars[1] at +0, +8
  const uint32_t sa_slot = sa_bar + 16;           // the next tile: {tile, row, t0, n_valid}, {frame_limit, bulk_lo, bulk_n, manual}

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  float2* my_tile = tiles + warp * LY::kTileF2;
  grid_launch_dependents();

  if (tid == 0) {
    mbar_init(sa_bar, 1);
    fence_mbar_init();
  }

  // per-lane constants kept in registers for the whole kernel
  constexpr int kPts = LY::kTile512 ? 16 : 32;  // complex points per lane
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

#ifdef DMEL_BOOK_WARP
  constexpr int kBook = DMEL_BOOK_WARP * 32;  // the thread that does the tile bookkeeping
#else
  constexpr int kBook = 0;
#endif
  const bool hop_even = (p.hop & 1) == 0;
  unsigned long long edge_hits = 0;
  uint32_t wave_parity = 0;  // parity to wait for on the waveform barrier

  // Where tile `tile` sits and how its samples reach shared memory.  Two multiply-high divisions and some
  // integer logic: evaluated by one thread per tile.
  auto describe = [&](int tile) {
    TileInfo ti;
    ti.row = (int)p.by_tiles_per_row.div((unsigned)tile);
    ti.t0 = (tile - ti.row * p.tiles_per_row) * TF;
    ti.n_valid = p.n_frames;  // in frames of this launch's window
    if (p.lengths) {
      const int len = p.lengths[ti.row];
      const int nv = (len <= 0 ? 0 : (int)p.by_hop.div((unsigned)len)) - p.t_begin;
      ti.n_valid = nv < 0 ? 0 : (nv < p.n_frames ? nv : p.n_frames);
    }
    // log-mel output covers every frame of the row (unless masked); codes / statistics only the valid ones
    const int last = ((kLogmel && !p.mask_invalid) ? p.n_frames : ti.n_valid) - ti.t0;
    ti.frame_limit = last < 0 ? 0 : (last > TF ? TF : last);
    // the row as the reflect padding sees it: n samples starting at flat offset base (ragged batches, own-length rows)
    int n = p.n_samples;
    long long base = (long long)ti.row * p.row_stride;
    if (p.offsets) {
      base = p.offsets[ti.row];
      n = (int)(p.offsets[ti.row + 1] - base);
      if (p.lengths) n = min(n, max(p.lengths[ti.row], 0));  // rows may be stored with alignment slack between them
    } else if (p.own_length) {
      n = p.lengths[ti.row];
    }
    if (p.own_length) {
      const int nv = (n <= 0 ? 0 : (int)p.by_hop.div((unsigned)n)) - p.t_begin;  // an utterance has n / hop frames
      const int cap = ti.n_valid;
      ti.n_valid = nv < 0 ? 0 : (nv < cap ? nv : cap);
      const int last_own = ti.n_valid - ti.t0;  // nothing past the utterance's own frames is computed
      ti.frame_limit = last_own < 0 ? 0 : (last_own > TF ? TF : last_own);
    }
    const int s0 = (p.t_begin + ti.t0) * p.hop - p.pad_inner - p.pad_outer;  // virtual sample under the tile's first tap
    const int b0 = s0 - p.src_base;                                          // where that sample sits in the buffer
    ti.src0 = base + b0;
    constexpr int kA = kAlign - 1;
    const bool base_ok = p.bulk_ok && (!p.offsets || (base & kA) == 0);
    if (base_ok && s0 >= 0 && b0 >= 0 && (b0 & kA) == 0 && s0 + p.wave_len <= n) {
      ti.bulk_lo = 0;  // interior tile (all but one or two per row): one copy brings everything
      ti.bulk_n = p.wave_len;
      ti.manual = 0;
    } else {
      int lo = s0 < 0 ? ((-s0 + kA) & ~kA) : 0;                                 // first wave index with a real sample
      if (b0 + lo < 0) lo = (-b0 + kA) & ~kA;                                   // ... that is resident in the buffer
      int hi = n - s0 < p.wave_len ? ((n - s0) & ~kA) : p.wave_len;             // one past the last
      const bool can_bulk = base_ok && (b0 & kA) == 0 && hi > lo;
      ti.bulk_lo = can_bulk ? lo : 0;
      ti.bulk_n = can_bulk ? hi - lo : 0;
      ti.manual = (!can_bulk || lo > 0 || hi < p.wave_len) ? 1 : 0;
    }
    return ti;
  };
  // thread 0: start the bulk async copy of a tile into the wave buffer
  auto stage_bulk = [&](const TileInfo& ti) {
    if (ti.frame_limit == 0 || ti.bulk_n == 0 || DMEL_SKIP(p, 4)) return;
    const uint32_t dst = sa_base + p.off_wave + (uint32_t)ti.bulk_lo * (uint32_t)sizeof(wave_t);
    fence_proxy_async();  // earlier generic-proxy reads of this buffer are ordered before the async write
    mbar_expect_tx(sa_bar, ti.bulk_n * (int)sizeof(wave_t));
    bulk_copy_g2s(dst, wav + ti.src0 + ti.bulk_lo, ti.bulk_n * (int)sizeof(wave_t), sa_bar);
  };
  // thread 0: leave the description of the next tile (or "no more tiles") in the slot
  auto post = [&](int tile, const TileInfo& ti) {
    const bool any = tile < p.n_tiles;
    sts_i4(sa_slot, make_int4(tile, any ? ti.row : 0, any ? ti.t0 : 0, any ? ti.n_valid : 0));
    sts_i4(sa_slot + 16, make_int4(any ? ti.frame_limit : 0, any ? ti.bulk_lo : 0, any ? ti.bulk_n : 0, any ? ti.manual : 0));
  };
  // all threads: the part of a tile no bulk copy can bring (reflected row ends, unaligned rows)
  auto stage_manual = [&](int row, int t0, int bulk_lo, int bulk_n) {
    if (DMEL_SKIP(p, 4)) return;
    wave_t* wave = wave0;
    int n = p.n_samples;
    long long base = (long long)row * p.row_stride;
    if (p.offsets) {
      base = p.offsets[row];
      n = (int)(p.offsets[row + 1] - base);
      if (p.lengths) n = min(n, max(p.lengths[row], 0));
    } else if (p.own_length) {
      n = p.lengths[row];
    }
    const int padded_len = n + 2 * p.pad_inner + 2 * p.pad_outer;
    const wave_t* src = wav + base - p.src_base;
    const int j0 = (p.t_begin + t0) * p.hop;  // first position in the padded row
    const int n_manual = p.wave_len - bulk_n;
    for (int q = tid; q < n_manual; q += kThreads) {
      const int i = q < bulk_lo ? q : q + bulk_n;  // wave index outside the bulk range
      const int j = j0 + i;
      wave_t x = 0;
      if (j < padded_len) x = __ldg(src + reflect_src(j, n, p.pad_inner, p.pad_outer));
      wave[i] = x;
    }
  };

  // ---- per-CTA constants -------------------------------------------------
  // Everything up to grid_dependency_wait() reads plan-owned memory only (written at plan creation), so under
  // programmatic dependent launch it overlaps the tail of the previous kernel in the stream.
  for (int i = tid; i < p.n_chan_pad; i += kThreads) {
    *reinterpret_cast<int4*>(&s_rec[i]) = p.chan[i];
    if constexpr (!kCodes) {
      s_rec[i].a = __int_as_float(0x7f800000);
      s_rec[i].b = __int_as_float(0xff800000);
    }
  }
  for (int i = tid; i < p.n_order; i += kThreads) s_order[i] = p.group_order[i];
  for (int i = tid; i < p.nnz; i += kThreads) s_weights[i] = p.weights[i];
  if constexpr (!LY::kMagsInTiles) {
    for (int i = tid; i < LY::kMagFloats; i += kThreads) mags[i] = 0.f;  // the 3 pad columns of each row stay zero
  } else {
    for (int i = tid; i < kWarps * LY::kTileF2; i += kThreads) tiles[i] = make_float2(0.f, 0.f);  // rows of frames never computed read as finite
  }
  if constexpr (LY::kWindowInSmem) {
    for (int i = tid; i < NFFT; i += kThreads) s_window[i] = p.window[i];
  }
  if constexpr (NFFT == 2048 && !kSplit) {
    for (int i = tid; i < LY::kFoldN; i += kThreads) s_fold[i] = p.fold_tw[i];
  }
  grid_dependency_wait();  // from here on: the caller's tensors (waveform, lengths, statistics, outputs, tile counter)

  // Tiles: the first is blockIdx.x; later ones come from a global counter (or a static stride when p.sched is
  // null), claimed ONE tile ahead so that the CTAs of the grid finish within a tile of each other (a tile is
  // several microseconds of a launch that lasts a hundred: claiming two ahead costs more at the end of the
  // launch than the deeper prefetch gains).
  const bool dynamic = p.sched != nullptr;
  int tile = blockIdx.x;
  const TileInfo first = describe(tile < p.n_tiles ? tile : 0);
  if (tile < p.n_tiles) {
    if (tid == kBook) stage_bulk(first);  // the first tile leaves HBM while the statistics below are fetched
    if (first.manual && first.frame_limit) stage_manual(first.row, first.t0, first.bulk_lo, first.bulk_n);
  }
  if constexpr (kCodes) {
    for (int i = tid; i < p.n_chan_pad; i += kThreads) {
      const bool real = i < p.n_mels;
      s_rec[i].a = real ? p.q_lo[i] : 0.f;
      s_rec[i].b = real ? p.q_scale[i] : 0.f;
      if constexpr (kDequant) s_rec[i].c = real ? p.q_step[i] : 0.f;
    }
  }
  __syncthreads();  // constants, barrier init and the first tile's plain stores visible

  // CTA-uniform description of the current tile (every thread holds a copy)
  int cur_row = first.row, cur_t0 = first.t0, cur_valid = first.n_valid, cur_limit = first.frame_limit;
  int cur_bulk = first.bulk_n;

  while (tile < p.n_tiles) {
    const wave_t* wave = wave0;
    const bool dead = cur_limit == 0;

    // ---- 1. this tile's samples are in the wave buffer (plain stores: ordered by the barriers that closed the previous tile)
    if (!dead && cur_bulk && !DMEL_SKIP(p, 4)) {
      mbar_wait(sa_bar, wave_parity);
      wave_parity ^= 1u;
    }
    int fetched = 0;
    if (dynamic && tid == kBook) fetched = atomicAdd(p.sched, 1);  // consumed after the FFT barrier: its latency is never waited for
    if constexpr (!kPcm) {
      if (p.row_gain != nullptr && !dead) {  // per-utterance gain (peak normalisation): scale the staged tile in place
        const float gain = __ldg(p.row_gain + cur_row);
        float4* w4 = reinterpret_cast<float4*>(wave0);
        for (int i = tid; i < p.wave_len / 4; i += kThreads) {
          float4 x = w4[i];
          x.x *= gain, x.y *= gain, x.z *= gain, x.w *= gain;
          w4[i] = x;
        }
        __syncthreads();
      }
    }

    // ---- 2. FFT -> magnitudes ------------------------------------------------
    const int fft_frames = DMEL_SKIP(p, 1) ? 0 : cur_limit;
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
        __syncwarp();  // previous readers of my_tile (pass 2 of the previous frame, or the previous tile's mel phase) are done
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
        // with kMagsInTiles the row is the warp's own tile: every lane's pass-2 loads have returned (the
        // shuffles above needed them), so it is free to be overwritten
        float* mrow = LY::kMagsInTiles ? reinterpret_cast<float*>(my_tile) : mags + fr * kPitch;
        unfold_store512(zlo, zhi, recv, fold_base, mrow, lane);
        if constexpr (LY::kMagsInTiles) {
          if (lane < 3) mrow[513 + lane] = 0.f;  // the pad columns a 16-byte aligned span may touch (zero weight)
        }
      }
    } else if constexpr (kSplit) {
      const int h = lane >> 4;
      const int partner = mirror_lane512(lane);
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
        fft512_half_spectrum(v, w1, w2, w4, w8, my_tile, lane, h, partner, base1024, e);
        fft512_half_spectrum(odd, w1, w2, w4, w8, my_tile, lane, h, partner, base1024, o);
        float* mrow = LY::kMagsInTiles ? reinterpret_cast<float*>(my_tile) : mags + fr * kPitch;
        combine2048_store(e, o, base2048, mrow, lane);
        if constexpr (LY::kMagsInTiles) {
          if (lane < 3) mrow[1025 + lane] = 0.f;
        }
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
    __syncthreads();  // magnitudes complete, the wave buffer is free
    // One thread finds out which tile comes next, describes it and starts its copy, which flies under the mel
    // phase; that thread's warp carries the lightest share of the mel phase to make up for it.
    if (tid == kBook) {
      const int next_tile = dynamic ? (int)gridDim.x + fetched : tile + (int)gridDim.x;
      TileInfo nd = first;
      if (next_tile < p.n_tiles) {
        nd = describe(next_tile);
        stage_bulk(nd);
      }
      post(next_tile, nd);
    }

    // ---- 3. mel filterbank, log, quantise --------------------------------
    if (!DMEL_SKIP(p, 2))
      mel_phase<TF, MODE, kPitch>(p, sa_base, sa_base + p.off_mags, lane, warp, kWarps, cur_row, cur_t0, cur_valid, dead, edge_hits);
    __syncthreads();  // magnitudes are free again; the slot describes the next tile
    const int4 nx0 = lds_i4(sa_slot), nx1 = lds_i4(sa_slot + 16);
    tile = nx0.x;
    cur_row = nx0.y;
    cur_t0 = nx0.z;
    cur_valid = nx0.w;
    cur_limit = nx1.x;
    cur_bulk = nx1.z;
    if (tile < p.n_tiles && nx1.w && nx1.x) {  // a row end: stage what no bulk copy could bring (a few tiles per row)
      stage_manual(nx0.y, nx0.z, nx1.y, nx1.z);
      __syncthreads();
    }
  }
  // the last CTA to finish leaves both counters at zero for the next launch
  if (dynamic && tid == 0) {
    __threadfence();
    if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
      atomicExch(p.sched, 0);
      atomicExch(p.sched + 1, 0);
    }
  }

  // ---- flush per-CTA statistics -------------------------------------------
  if constexpr (kStats) {
    for (int m = tid; m < p.n_mels; m += kThreads) {
      const float lo = s_rec[m].a, hi = s_rec[m].b;
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
