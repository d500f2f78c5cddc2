// Warp-specialised form of the fused waveform -> log-mel -> dMel codes kernel (sm_100a, n_fft 1024, 8-frame tiles).
//
// logmel_kernel.cuh runs the three phases of a tile (stage, FFT, mel) on the same eight warps with two CTA-wide
// barriers per tile; a fifth of its warp time is spent waiting at them for the slowest warp.  Here the phases
// belong to different warps of a 16-warp CTA and meet only through mbarriers, so nobody waits for a phase it does
// not take part in:
//
//   warps 0-7   FFT      one frame of the tile each: load + window (from wave[k & 1]), register FFT (fft_core.cuh),
//                        magnitudes into mags[k & 1]
//   warps 8-14  mel      banded filterbank, log, quantise, stores (mel_phase of logmel_kernel.cuh) from mags[k & 1]
//   warp 15     producer claims the next tile, describes it, starts its bulk async copy into wave[k & 1] and stages
//                        what no bulk copy can bring (reflected row ends)
//
//   producer --wave_full--> FFT --mags_full--> mel
//            <-wave_empty--     <-mags_empty--
//
// Both rings are two deep.  Registers follow the roles (setmaxnreg): 80 per thread for the FFT warps, 48 for the
// others, 64 on average: two CTAs (32 warps) per SM.
#pragma once
#include "logmel_kernel.cuh"

namespace dmel {

constexpr int kWsWarps = 16;
constexpr int kWsThreads = kWsWarps * 32;
constexpr int kWsFftWarps = 8;
constexpr int kWsMelWarps = 7;
constexpr int kWsTF = 8;

struct WsLayout {
  static constexpr int kBins = 513;
  static constexpr int kMagPitch = kBins + 3;  // 516 floats: a multiple of 4, == 4 (mod 32)
  static constexpr int kTileF2 = kTile512;     // 544 float2 per transpose tile
  static __host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
  static __host__ __device__ size_t tiles_off() { return 0; }
  static __host__ __device__ size_t mags_off() { return size_t(kWsFftWarps) * kTileF2 * sizeof(float2); }
  static __host__ __device__ size_t wave_off() { return align16(mags_off() + 2 * size_t(kWsTF) * kMagPitch * 4); }
  static __host__ __device__ size_t window_off(int wave_len) { return align16(wave_off() + 2 * size_t(wave_len) * 4); }
  static __host__ __device__ size_t weights_off(int wave_len) { return align16(window_off(wave_len) + 1024 * 4); }
  static __host__ __device__ size_t rec_off(int wave_len, int nnz) { return align16(weights_off(wave_len) + size_t(nnz) * 4); }
  static __host__ __device__ size_t order_off(int wave_len, int n_chan, int nnz) {
    return align16(rec_off(wave_len, nnz) + size_t(n_chan) * sizeof(ChanRec));
  }
  static __host__ __device__ size_t bar_off(int wave_len, int n_chan, int nnz, int n_order) {
    return align16(order_off(wave_len, n_chan, nnz) + size_t(n_order) * 4);
  }
  // eight mbarriers (64 B), two tile descriptions for the FFT warps (64 B), two for the mel warps (64 B)
  static __host__ __device__ size_t total(int wave_len, int n_chan, int nnz, int n_order) {
    return bar_off(wave_len, n_chan, nnz, n_order) + 192;
  }
};

__device__ __forceinline__ void set_max_registers_80() { asm volatile("setmaxnreg.inc.sync.aligned.u32 80;"); }
__device__ __forceinline__ void set_max_registers_48() { asm volatile("setmaxnreg.dec.sync.aligned.u32 48;"); }

template <int MODE>
__global__ void __launch_bounds__(kWsThreads, 2) dmel_ws_kernel(const __grid_constant__ FusedParams p) {
  constexpr int TF = kWsTF;
  using LY = WsLayout;
  constexpr bool kCodes = (MODE & kOutCodes) != 0, kLogmel = (MODE & kOutLogmel) != 0;
  constexpr bool kStats = (MODE & kOutStats) != 0, kEdge = (MODE & kOutEdge) != 0;
  constexpr bool kPcm = (MODE & kInPcm16) != 0;
  constexpr bool kDequant = (MODE & kOutDequant) != 0;
  static_assert(!(kCodes && kStats), "codes and statistics share the float half of the channel records");
  using wave_t = std::conditional_t<kPcm, short, float>;
  constexpr int kAlign = 16 / (int)sizeof(wave_t);
  const wave_t* wav = reinterpret_cast<const wave_t*>(p.wav);

  extern __shared__ __align__(16) unsigned char smem[];
  float2* tiles = reinterpret_cast<float2*>(smem);
  wave_t* wave0 = reinterpret_cast<wave_t*>(smem + p.off_wave);
  float* s_window = reinterpret_cast<float*>(smem + p.off_window);
  float* s_weights = reinterpret_cast<float*>(smem + p.off_weights);
  ChanRec* s_rec = reinterpret_cast<ChanRec*>(smem + p.off_rec);
  int* s_order = reinterpret_cast<int*>(smem + p.off_order);
  const uint32_t sa_base = smem_u32(smem);
  const uint32_t sa_bar = sa_base + p.off_bars;
  // mbarriers: wave_full[2] +0, wave_empty[2] +16, mags_full[2] +32, mags_empty[2] +48; then the descriptions
  const uint32_t sa_wave_full = sa_bar, sa_wave_empty = sa_bar + 16, sa_mags_full = sa_bar + 32, sa_mags_empty = sa_bar + 48;
  const uint32_t sa_desc = sa_bar + 64;   // [2] x {tile, row, t0, n_valid}, {frame_limit, stop, -, -}: producer -> FFT warps
  const uint32_t sa_hdr = sa_bar + 128;   // [2] x the same: FFT warp 0 -> mel warps (travels with the magnitudes)

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  grid_launch_dependents();
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(sa_wave_full + 8 * b, 1);              // the producer's arrive.expect_tx (+ the copy's bytes)
      mbar_init(sa_wave_empty + 8 * b, kWsFftWarps);   // every FFT warp, once its frame is in registers
      mbar_init(sa_mags_full + 8 * b, kWsFftWarps);    // every FFT warp, once its row of magnitudes is written
      mbar_init(sa_mags_empty + 8 * b, kWsMelWarps);   // every mel warp, once it has read the last magnitude
    }
    fence_mbar_init();
  }

  // ---- per-CTA constants (plan-owned memory only: overlaps the previous kernel under programmatic dependent launch)
  for (int i = tid; i < p.n_chan_pad; i += kWsThreads) {
    *reinterpret_cast<int4*>(&s_rec[i]) = p.chan[i];
    if constexpr (!kCodes) {
      s_rec[i].a = __int_as_float(0x7f800000);
      s_rec[i].b = __int_as_float(0xff800000);
    }
  }
  for (int i = tid; i < p.n_order; i += kWsThreads) s_order[i] = p.group_order[i];
  for (int i = tid; i < p.nnz; i += kWsThreads) s_weights[i] = p.weights[i];
  {
    float* mags = reinterpret_cast<float*>(smem + p.off_mags);
    for (int i = tid; i < 2 * TF * LY::kMagPitch; i += kWsThreads) mags[i] = 0.f;  // pad columns and never-computed rows stay finite
  }
  for (int i = tid; i < 1024; i += kWsThreads) s_window[i] = p.window[i];
  grid_dependency_wait();  // from here on: the caller's tensors
  if constexpr (kCodes) {
    for (int i = tid; i < p.n_chan_pad; i += kWsThreads) {
      const bool real = i < p.n_mels;
      s_rec[i].a = real ? p.q_lo[i] : 0.f;
      s_rec[i].b = real ? p.q_scale[i] : 0.f;
      if constexpr (kDequant) s_rec[i].c = real ? p.q_step[i] : 0.f;
    }
  }
  __syncthreads();
  unsigned long long edge_hits = 0;

  if (warp < kWsFftWarps) {
    // =========================== FFT warps ===========================
    set_max_registers_80();
    float2* my_tile = tiles + warp * LY::kTileF2;
    const float2 w1 = p.stage_tw[1 * 32 + lane], w2 = p.stage_tw[2 * 32 + lane];
    const float2 w4 = p.stage_tw[4 * 32 + lane], w8 = p.stage_tw[8 * 32 + lane];
    const float2 fold_base = p.fold_tw[lane];
    const float2* my_win = reinterpret_cast<const float2*>(s_window) + lane;
    const int h = lane >> 4;
    const int partner = mirror_lane512(lane);
    const bool hop_even = (p.hop & 1) == 0;
    for (int k = 0;; ++k) {
      const int b = k & 1;
      const uint32_t use = (uint32_t)(k >> 1) & 1u;
      mbar_wait(sa_wave_full + 8 * b, use);
      const int4 d0 = lds_i4(sa_desc + 32 * b), d1 = lds_i4(sa_desc + 32 * b + 16);
      const bool stop = d1.y != 0;
      const int limit = d1.x;
      const bool mine = !stop && warp < limit;
      float2 v[16];
      if (mine) {
        const wave_t* fa = wave0 + b * p.wave_len + warp * p.hop;
        if (p.row_gain != nullptr && !kPcm) {
          const float gain = __ldg(p.row_gain + d0.y);
#pragma unroll
          for (int n1 = 0; n1 < 16; ++n1) {
            const int idx = 2 * (32 * n1 + lane);
            v[n1] = f2_mul(make_float2((float)fa[idx] * gain, (float)fa[idx + 1] * gain), my_win[32 * n1]);
          }
        } else if (hop_even) {
          using pair_t = std::conditional_t<kPcm, short2, float2>;
          const pair_t* f2 = reinterpret_cast<const pair_t*>(fa);
#pragma unroll
          for (int n1 = 0; n1 < 16; ++n1) {
            const pair_t x = f2[32 * n1 + lane];
            v[n1] = f2_mul(make_float2((float)x.x, (float)x.y), my_win[32 * n1]);
          }
        } else {
#pragma unroll
          for (int n1 = 0; n1 < 16; ++n1) {
            const int idx = 2 * (32 * n1 + lane);
            v[n1] = f2_mul(make_float2((float)fa[idx], (float)fa[idx + 1]), my_win[32 * n1]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sa_wave_empty + 8 * b);  // this warp's share of wave[b] is in registers
      if (stop) {  // pass the end of the work on to the mel warps, then leave
        mbar_wait(sa_mags_empty + 8 * b, use ^ 1u);
        if (warp == 0 && lane == 0) {
          sts_i4(sa_hdr + 32 * b, d0);
          sts_i4(sa_hdr + 32 * b + 16, d1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sa_mags_full + 8 * b);
        break;
      }
      if (mine) {
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
        mbar_wait(sa_mags_empty + 8 * b, use ^ 1u);  // the mel warps have read this buffer's previous tile
        float* mrow = reinterpret_cast<float*>(smem + p.off_mags) + (b * TF + warp) * LY::kMagPitch;
        unfold_store512(zlo, zhi, recv, fold_base, mrow, lane);
      } else {
        mbar_wait(sa_mags_empty + 8 * b, use ^ 1u);
      }
      if (warp == 0 && lane == 0) {  // the description travels with the magnitudes
        sts_i4(sa_hdr + 32 * b, d0);
        sts_i4(sa_hdr + 32 * b + 16, d1);
      }
      __syncwarp();  // (also: this frame's pass-2 reads of my_tile are done before the next frame's pass 1)
      if (lane == 0) mbar_arrive(sa_mags_full + 8 * b);
    }
  } else {
    set_max_registers_48();
    if (warp < kWsFftWarps + kWsMelWarps) {
      // =========================== mel warps ===========================
      const int slot = warp - kWsFftWarps;
      for (int k = 0;; ++k) {
        const int b = k & 1;
        const uint32_t use = (uint32_t)(k >> 1) & 1u;
        mbar_wait(sa_mags_full + 8 * b, use);
        const int4 d0 = lds_i4(sa_hdr + 32 * b), d1 = lds_i4(sa_hdr + 32 * b + 16);
        if (d1.y != 0) break;  // stop
        mel_phase<TF, MODE, LY::kMagPitch>(p, sa_base, sa_base + p.off_mags + (uint32_t)(b * TF * LY::kMagPitch * 4), lane, slot,
                                           kWsMelWarps, d0.y, d0.z, d0.w, d1.x == 0, edge_hits);
        __syncwarp();
        if (lane == 0) mbar_arrive(sa_mags_empty + 8 * b);
      }
    } else {
      // =========================== producer warp ===========================
      const bool dynamic = p.sched != nullptr;
      int tile = blockIdx.x;
      for (int k = 0;; ++k) {
        const int b = k & 1;
        const uint32_t use = (uint32_t)(k >> 1) & 1u;
        mbar_wait(sa_wave_empty + 8 * b, use ^ 1u);  // the FFT warps have taken this buffer's previous tile (k - 2) ...
        if (k >= 2) mbar_wait(sa_mags_full + 8 * b, use ^ 1u);  // ... and finished its FFTs: they are now starting tile k - 1
        if (k > 0) {  // claim only now, ONE tile ahead of the FFT warps: keeps the CTAs of the grid within a tile of each other
          int next = 0;
          if (lane == 0) next = dynamic ? (int)gridDim.x + atomicAdd(p.sched, 1) : tile + (int)gridDim.x;
          tile = __shfl_sync(0xffffffffu, next, 0);
        }
        const bool stop = tile >= p.n_tiles;
        // ---- describe the tile (every lane: the values are needed by all of them for the plain-load staging)
        int row = 0, t0 = 0, n_valid = 0, frame_limit = 0, bulk_lo = 0, bulk_n = 0, manual = 0, n = p.n_samples;
        long long base = 0, src0 = 0;
        if (!stop) {
          row = (int)p.by_tiles_per_row.div((unsigned)tile);
          t0 = (tile - row * p.tiles_per_row) * TF;
          n_valid = p.n_frames;
          if (p.lengths) {
            const int len = p.lengths[row];
            const int nv = (len <= 0 ? 0 : (int)p.by_hop.div((unsigned)len)) - p.t_begin;
            n_valid = nv < 0 ? 0 : (nv < p.n_frames ? nv : p.n_frames);
          }
          const int last = ((kLogmel && !p.mask_invalid) ? p.n_frames : n_valid) - t0;
          frame_limit = last < 0 ? 0 : (last > TF ? TF : last);
          base = (long long)row * p.row_stride;
          if (p.offsets) {
            base = p.offsets[row];
            n = (int)(p.offsets[row + 1] - base);
            if (p.lengths) n = min(n, max(p.lengths[row], 0));
          } else if (p.own_length) {
            n = p.lengths[row];
          }
          if (p.own_length) {
            const int nv = (n <= 0 ? 0 : (int)p.by_hop.div((unsigned)n)) - p.t_begin;
            n_valid = nv < 0 ? 0 : (nv < n_valid ? nv : n_valid);
            const int last_own = n_valid - t0;
            frame_limit = last_own < 0 ? 0 : (last_own > TF ? TF : last_own);
          }
          const int s0 = (p.t_begin + t0) * p.hop - p.pad_inner - p.pad_outer;
          const int b0 = s0 - p.src_base;
          src0 = base + b0;
          constexpr int kA = kAlign - 1;
          const bool base_ok = p.bulk_ok && (!p.offsets || (base & kA) == 0);
          if (base_ok && s0 >= 0 && b0 >= 0 && (b0 & kA) == 0 && s0 + p.wave_len <= n) {
            bulk_n = p.wave_len;
          } else {
            int lo = s0 < 0 ? ((-s0 + kA) & ~kA) : 0;
            if (b0 + lo < 0) lo = (-b0 + kA) & ~kA;
            int hi = n - s0 < p.wave_len ? ((n - s0) & ~kA) : p.wave_len;
            const bool can_bulk = base_ok && (b0 & kA) == 0 && hi > lo;
            bulk_lo = can_bulk ? lo : 0;
            bulk_n = can_bulk ? hi - lo : 0;
            manual = (!can_bulk || lo > 0 || hi < p.wave_len) ? 1 : 0;
          }
          if (frame_limit == 0) bulk_n = 0, manual = 0;  // nothing of this tile is computed: nothing to stage
        }
        wave_t* wave = wave0 + b * p.wave_len;
        if (manual) {  // reflected row ends, unaligned rows: plain loads by the whole warp
          const int padded_len = n + 2 * p.pad_inner + 2 * p.pad_outer;
          const wave_t* src = wav + base - p.src_base;
          const int j0 = (p.t_begin + t0) * p.hop;
          const int n_manual = p.wave_len - bulk_n;
          for (int q = lane; q < n_manual; q += 32) {
            const int i = q < bulk_lo ? q : q + bulk_n;
            const int j = j0 + i;
            wave_t x = 0;
            if (j < padded_len) x = __ldg(src + reflect_src(j, n, p.pad_inner, p.pad_outer));
            wave[i] = x;
          }
        }
        __syncwarp();
        if (lane == 0) {
          sts_i4(sa_desc + 32 * b, make_int4(tile, row, t0, n_valid));
          sts_i4(sa_desc + 32 * b + 16, make_int4(frame_limit, stop ? 1 : 0, 0, 0));
          const uint32_t bytes = (uint32_t)bulk_n * (uint32_t)sizeof(wave_t);
          fence_proxy_async();  // the FFT warps' reads of this buffer (ordered by wave_empty) precede the async write
          mbar_expect_tx(sa_wave_full + 8 * b, bytes);  // the one arrival this barrier waits for, plus the copy's bytes
          if (bulk_n)
            bulk_copy_g2s(sa_base + p.off_wave + (uint32_t)(b * p.wave_len + bulk_lo) * (uint32_t)sizeof(wave_t),
                          wav + src0 + bulk_lo, bytes, sa_wave_full + 8 * b);
        }
        if (stop) break;
      }
      // the last CTA to finish leaves both counters at zero for the next launch
      if (dynamic && lane == 0) {
        __threadfence();
        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
          atomicExch(p.sched, 0);
          atomicExch(p.sched + 1, 0);
        }
      }
    }
  }
  __syncthreads();  // every role is done: the per-CTA statistics are complete

  if constexpr (kStats) {
    for (int m = tid; m < p.n_mels; m += kWsThreads) {
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
