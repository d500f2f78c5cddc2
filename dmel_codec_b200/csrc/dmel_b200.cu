// C ABI of the B200 dMel tokenization path (see include/dmel_b200.h).
// Host side: plan construction (banded filterbank, twiddle tables), argument
// checking, kernel launches.  No torch, no cuFFT, no CPU fallback.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/dmel_b200.h"
#include "fused_variants.h"
#include "launch_util.cuh"
#include "logmel_kernel.cuh"
#include "codec_kernels.cuh"
#include "extras_kernels.cuh"
#include "fsq_kernels.cuh"
#include "activation_kernels.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define DMEL_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return fail(DMEL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                  __FILE__, __LINE__);                                                   \
  } while (0)

using dmel::launch_pdl;

template <typename T>
cudaError_t upload(T** dev, const std::vector<T>& host) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dev), std::max<size_t>(host.size(), 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (host.empty()) return cudaSuccess;
  return cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace

struct dmel_plan {
  int device = 0;
  int sm_count = 0;
  int max_smem = 0;
  int n_fft = 0, hop = 0, n_mels = 0, center = 0;
  int core_fft = 0;  // transform size of the kernel that runs it: 1024 for n_fft <= 1024, else 2048
  int pad_inner = 0, pad_outer = 0;
  int tile_frames = 0;  // TF chosen for this geometry
  int ctas_per_sm = 1;
  int wave_len = 0;
  int nnz = 0;
  int n_chan_pad = 0;  // n_mels rounded up to the channel-group size 2 * 32 / tile_frames
  int n_order = 0;     // entries of the group-order table
  int* d_order = nullptr;
  size_t smem_bytes = 0;
  const dmel::VariantOps* variant = nullptr;  // the kernel variant chosen for this geometry
  float* d_window = nullptr;
  float* d_window_pcm = nullptr;  // window / 32768: int16 PCM input needs no separate scaling pass
  float2* d_stage_tw = nullptr;
  float2* d_fold_tw = nullptr;
  int4* d_chan = nullptr;
  float* d_weights = nullptr;
  // dynamic tile scheduling: kSchedSlots {next tile, finished CTAs} pairs, zero between launches; launches rotate
  // through them so that launches of one plan that overlap on different streams do not share a counter
  int* d_sched = nullptr;
  mutable std::atomic<unsigned> sched_turn{0};
  // scratch for the host-buffer entry point (grown on demand)
  cudaStream_t streams[2] = {nullptr, nullptr};
  cudaEvent_t stats_ready = nullptr;
  float* d_wav[2] = {nullptr, nullptr};
  uint8_t* d_codes[2] = {nullptr, nullptr};
  int32_t* d_len[2] = {nullptr, nullptr};
  float* d_lo = nullptr;
  float* d_scale = nullptr;
  std::vector<float> host_stats;  // the lo / scale values d_lo / d_scale hold (2 * n_mels), empty = nothing uploaded yet
  cudaEvent_t chunk_in[2] = {nullptr, nullptr};    // host pipeline: the slot's waveform has arrived
  cudaEvent_t chunk_used[2] = {nullptr, nullptr};  //                the slot's kernel has consumed it
  size_t wav_cap = 0, codes_cap = 0, len_cap = 0;
};

namespace {

using dmel::FusedParams;

// The kernel variants this build carries, most CTAs per SM first (one translation unit each: fused_variant.cu).
const dmel::VariantOps* const kVariants[] = {
#ifdef DMEL_WITH_WS
    &dmel::kVariant_ws_1024,  // warp-specialised experiment (ws_kernel.cuh; build with DMEL_BUILD_WS=1, select with DMEL_WS=1)
#endif
    &dmel::kVariant_1024_8_3, &dmel::kVariant_1024_16_2, &dmel::kVariant_1024_8_2, &dmel::kVariant_1024_16_1,
    &dmel::kVariant_1024_8_1, &dmel::kVariant_2048_8_2,  &dmel::kVariant_2048_16_1, &dmel::kVariant_2048_8_1};

constexpr unsigned kSchedSlots = 64;

cudaError_t launch_fused_any(const dmel_plan* plan, FusedParams& p, int grid, cudaStream_t st,
                             bool bf16_logmel = false, bool pcm16 = false) {
  using namespace dmel;
  // bulk async copies move whole 16-byte units: the waveform base and the row stride must be multiples of 16 bytes
  const long long per16 = pcm16 ? 8 : 4;
  p.bulk_ok = (reinterpret_cast<uintptr_t>(p.wav) & 15) == 0 && p.row_stride % per16 == 0 && p.pad_outer == 0;
  int mode = 0;
  if (p.codes) mode |= kOutCodes;
  if (p.logmel) mode |= kOutLogmel;
  if (p.run_min) mode |= kOutStats;
  if (p.near_edge) mode |= kOutEdge;
  if (bf16_logmel) mode |= kOutBf16;
  if (pcm16) mode |= kInPcm16;
  if (p.dequant) mode |= kOutDequant;
  return plan->variant->launch(mode, p, grid, plan->smem_bytes, st);
}

// Banded form of the (n_mels, n_freq) filterbank the kernel reads: per channel the contiguous
// non-zero span, starting on a multiple of 4 bins (the kernel fetches magnitudes with 16-byte
// loads).  A lane evaluates the channel PAIR (m, m + lanes) in one loop and `lanes` such pairs sit
// side by side in a warp (lanes = 32 / frames per tile), so the spans of a group of 2 * lanes adjacent
// channels are zero-padded to one common length (a multiple of 4, at least 4), phantom channels complete
// the last group, and the weights of a pair are interleaved in steps of four bins:
//   [4 weights of m][4 weights of m + lanes][next 4 of m][next 4 of m + lanes] ...
void band_filterbank(const float* basis, int n_mels, int n_freq, int lanes, std::vector<int4>* chan,
                     std::vector<float>* weights, std::vector<int>* group_steps) {
  const int group = 2 * lanes;
  const int n_pad = (n_mels + group - 1) / group * group;
  chan->assign(n_pad, make_int4(0, 0, 16, 0));
  weights->clear();
  group_steps->clear();
  std::vector<int> first(n_pad, 0), last(n_pad, -1);
  for (int m = 0; m < n_mels; ++m) {
    const float* row = basis + (size_t)m * n_freq;
    int f0 = -1, f1 = -1;
    for (int f = 0; f < n_freq; ++f)
      if (row[f] != 0.f) {
        if (f0 < 0) f0 = f;
        f1 = f;
      }
    first[m] = f0 < 0 ? 0 : (f0 & ~3);
    last[m] = f1;
  }
  auto weight = [&](int m, int f) { return (m < n_mels && f >= 0 && f <= last[m]) ? basis[(size_t)m * n_freq + f] : 0.f; };
  for (int g = 0; g < n_pad; g += group) {
    int count = 4;
    for (int m = g; m < g + group; ++m) count = std::max(count, (last[m] - first[m] + 1 + 3) / 4 * 4);
    group_steps->push_back(count / 4);
    const int pitch = n_freq + 3;  // bins of a magnitude row the kernel keeps finite (a multiple of 4)
    for (int m = g; m < g + group; ++m) first[m] = std::min(first[m], pitch - count);  // keep the padded span inside the row
    for (int s = 0; s < lanes; ++s) {
      const int ma = g + s, mb = g + lanes + s;
      const int base = (int)weights->size();
      // the integer half of dmel::ChanRec, in bytes: {first bin, first weight, span length, -}
      (*chan)[ma] = make_int4(first[ma] * 4, base * 4, count * 4, 0);
      (*chan)[mb] = make_int4(first[mb] * 4, (base + 4) * 4, count * 4, 0);
      for (int i = 0; i < count; i += 4) {
        for (int j = 0; j < 4; ++j) weights->push_back(weight(ma, first[ma] + i + j));
        for (int j = 0; j < 4; ++j) weights->push_back(weight(mb, first[mb] + i + j));
      }
    }
  }
}

// Deals the channel groups to the kW warps that run the mel phase so that every one of them runs about the same
// number of bin-loop steps (longest group first, always to the least loaded warp).  With `handicap`, warp 0 starts
// with a load: one of its threads describes the next tile and starts its copy while the others are already in the
// mel phase.  order[r * kW + w] = group of warp w in its r-th trip, -1 = none.
std::vector<int> deal_groups(const std::vector<int>& group_steps, int kW, bool handicap) {
  const int kEpilogue = 3, kHandicap = 5;  // in steps: fixed cost per group (records, log, quantise, stores); warp 0's extra work
  std::vector<int> idx(group_steps.size());
  for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
  if (const char* env = std::getenv("DMEL_GROUP_ORDER"); env && std::strcmp(env, "identity") == 0) {
    // measurements only: groups in channel order, round robin over the warps
    std::vector<int> order((idx.size() + kW - 1) / kW * kW, -1);
    for (size_t i = 0; i < idx.size(); ++i) order[i] = (int)i;
    return order;
  }
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return group_steps[a] > group_steps[b]; });
  std::vector<std::vector<int>> mine(kW);
  std::vector<int> load(kW, 0);
  if (handicap) load[0] = kHandicap;
  for (int g : idx) {
    int w = 0;
    for (int k = 1; k < kW; ++k)
      if (load[k] < load[w]) w = k;
    mine[w].push_back(g);
    load[w] += group_steps[g] + kEpilogue;
  }
  size_t rounds = 0;
  for (const auto& m : mine) rounds = std::max(rounds, m.size());
  std::vector<int> order(rounds * kW, -1);
  for (int w = 0; w < kW; ++w)
    for (size_t r = 0; r < mine[w].size(); ++r) order[r * kW + w] = mine[w][r];
  return order;
}

long long num_frames(const dmel_plan* plan, long long n_samples) {
  const long long padded = n_samples + 2LL * plan->pad_inner + 2LL * plan->pad_outer;
  if (padded < plan->n_fft) return 0;
  return 1 + (padded - plan->n_fft) / plan->hop;
}

// shared argument checks + parameter block for the fused entry points.  The general form writes
// frames [t_begin, t_begin + t_count) of rows whose samples [src_base, n_samples) are resident at
// wav[row][0 ...]; offline calls use src_base = 0, t_begin = 0, t_count = T.
int prepare_window(dmel_plan* plan, const float* wav, long long n_rows, long long n_samples, long long row_stride,
                   long long src_base, long long t_begin, long long t_count, FusedParams* p, int* grid) {
  if (!plan) return fail(DMEL_ERR_INVALID, "plan is null");
  if (!wav) return fail(DMEL_ERR_INVALID, "waveform pointer is null");
  if (n_rows < 0 || n_samples <= 0 || src_base < 0 || src_base >= n_samples || row_stride < n_samples - src_base)
    return fail(DMEL_ERR_INVALID, "bad waveform shape: rows=%lld samples=%lld base=%lld stride=%lld", n_rows,
                n_samples, src_base, row_stride);
  if (n_samples <= plan->pad_inner)
    return fail(DMEL_ERR_INVALID,
                "reflect padding of %d needs more than %d samples per row, got %lld "
                "(the reference's F.pad raises here too)", plan->pad_inner, plan->pad_inner, n_samples);
  if (plan->pad_outer && n_samples + 2LL * plan->pad_inner <= plan->pad_outer)
    return fail(DMEL_ERR_INVALID, "center=True reflect padding of %d needs a longer row", plan->pad_outer);
  if (n_samples > (1LL << 30)) return fail(DMEL_ERR_INVALID, "rows longer than 2^30 samples are not supported");
  const long long T = num_frames(plan, n_samples);
  if (T <= 0) return fail(DMEL_ERR_INVALID, "row of %lld samples is shorter than one frame", n_samples);
  if (t_count < 0) t_count = T - t_begin;
  if (t_begin < 0 || t_count <= 0 || t_begin + t_count > T)
    return fail(DMEL_ERR_INVALID, "frame window [%lld, %lld) outside the %lld frames of the row", t_begin,
                t_begin + t_count, T);
  if (src_base > 0) {
    if (plan->pad_outer) return fail(DMEL_ERR_UNSUPPORTED, "windowed (streaming) calls do not support center=True");
    const long long first_tap = t_begin * plan->hop - plan->pad_inner;
    if (src_base > std::max<long long>(first_tap, 0))
      return fail(DMEL_ERR_INVALID, "frame %lld needs sample %lld but the buffer starts at sample %lld", t_begin,
                  std::max<long long>(first_tap, 0), src_base);
  }
  const long long tiles_per_row = (t_count + plan->tile_frames - 1) / plan->tile_frames;
  const long long n_tiles = tiles_per_row * n_rows;
  if (n_tiles > 0x7fffffffLL) return fail(DMEL_ERR_INVALID, "batch too large: %lld tiles", n_tiles);
  std::memset(p, 0, sizeof(*p));
  p->wav = wav;
  p->row_stride = row_stride;
  p->n_rows = (int)n_rows;
  p->n_samples = (int)n_samples;
  p->n_frames = (int)t_count;
  p->t_begin = (int)t_begin;
  p->src_base = (int)src_base;
  p->tiles_per_row = (int)tiles_per_row;
  p->n_tiles = (int)n_tiles;
  p->hop = plan->hop;
  p->pad_inner = plan->pad_inner;
  p->pad_outer = plan->pad_outer;
  p->n_mels = plan->n_mels;
  p->n_chan_pad = plan->n_chan_pad;
  p->wave_len = plan->wave_len;
  p->nnz = plan->nnz;
  p->window = plan->d_window;
  p->stage_tw = plan->d_stage_tw;
  p->fold_tw = plan->d_fold_tw;
  p->chan = plan->d_chan;
  p->group_order = plan->d_order;
  p->n_order = plan->n_order;
  p->weights = plan->d_weights;
  p->n_bins = 1;
  p->kmax = 0.f;
  if (plan->d_sched && !std::getenv("DMEL_STATIC_TILES")) p->sched = plan->d_sched + 2 * (plan->sched_turn++ % kSchedSlots);
  if (const char* dbg = std::getenv("DMEL_DEBUG_SKIP")) p->debug_skip = std::atoi(dbg);  // ablation timing only
  p->by_tiles_per_row = dmel::FastDiv::make((unsigned)tiles_per_row);
  p->by_hop = dmel::FastDiv::make((unsigned)plan->hop);
  plan->variant->fill_offsets(p);
  *grid = (int)std::max<long long>(1, std::min<long long>(n_tiles, (long long)plan->sm_count * plan->ctas_per_sm));
  return DMEL_OK;
}

int prepare_fused(dmel_plan* plan, const float* wav, long long n_rows, long long n_samples,
                  long long row_stride, FusedParams* p, int* grid) {
  return prepare_window(plan, wav, n_rows, n_samples, row_stride, 0, 0, -1, p, grid);
}

int check_bins(int n_bins) {
  if (n_bins < 1 || n_bins > 256) return fail(DMEL_ERR_INVALID, "n_bins must be in [1, 256] for uint8 codes, got %d", n_bins);
  return DMEL_OK;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// The tensor-level entry points take no plan: the device is the one that owns the tensor.
int device_of(const void* dev_ptr) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, dev_ptr) == cudaSuccess && attr.type == cudaMemoryTypeDevice) return attr.device;
  cudaGetLastError();
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

int sm_count_of(int dev) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

int stream_grid(int dev, unsigned n_groups) {
  const long long want = (n_groups + dmel::kStreamThreads * dmel::kStreamUnroll - 1LL) /
                         (dmel::kStreamThreads * dmel::kStreamUnroll);
  return (int)std::max<long long>(1, std::min<long long>(want, 8LL * sm_count_of(dev)));
}

}  // namespace

extern "C" {

int dmel_abi_version(void) { return DMEL_ABI_VERSION; }

const char* dmel_last_error(void) { return g_last_error.c_str(); }

int dmel_plan_create(int n_fft, int hop_length, int n_mels, int center, const float* mel_basis_host,
                     const float* window_host, dmel_plan** out) {
  if (!out) return fail(DMEL_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!mel_basis_host || !window_host) return fail(DMEL_ERR_INVALID, "mel_basis_host / window_host is null");
  if (n_fft < 64 || n_fft > 2048 || (n_fft & (n_fft - 1)) != 0)
    return fail(DMEL_ERR_UNSUPPORTED, "n_fft=%d: this build has radix kernels for powers of two from 64 to 2048", n_fft);
  if (hop_length < 1 || hop_length > n_fft)
    return fail(DMEL_ERR_INVALID, "hop_length must be in [1, n_fft], got %d", hop_length);
  if (n_mels < 1 || n_mels > 1024) return fail(DMEL_ERR_INVALID, "n_mels must be in [1, 1024], got %d", n_mels);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(DMEL_ERR_NO_DEVICE, "no CUDA device visible; dmel_b200 has no CPU path");
  }
  for (size_t i = 0; i < (size_t)n_mels * (n_fft / 2 + 1); ++i)
    if (!std::isfinite(mel_basis_host[i])) return fail(DMEL_ERR_INVALID, "mel_basis[%zu] is not finite", i);

  dmel_plan* plan = new (std::nothrow) dmel_plan();
  if (!plan) return fail(DMEL_ERR_INVALID, "out of host memory");
  int max_sm_smem = 0;
  cudaError_t e = cudaGetDevice(&plan->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, plan->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&plan->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, plan->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, plan->device);
  if (e != cudaSuccess) {
    delete plan;
    return fail(DMEL_ERR_CUDA, "device query failed: %s", cudaGetErrorString(e));
  }
  plan->n_fft = n_fft;
  plan->hop = hop_length;
  plan->n_mels = n_mels;
  plan->center = center ? 1 : 0;
  plan->pad_inner = (n_fft - hop_length) / 2;  // reference utils/spectrogram.py:58
  plan->pad_outer = center ? n_fft / 2 : 0;

  // pick the first kernel variant (most CTAs per SM first) whose shared memory fits; DMEL_OCC=<n>
  // pins the CTAs-per-SM choice (experiments)
  const char* occ_env = std::getenv("DMEL_OCC");
  const int occ_pin = occ_env ? std::atoi(occ_env) : 0;
  const char* ws_env = std::getenv("DMEL_WS");
  const bool ws_on = ws_env ? std::atoi(ws_env) != 0 : false;  // off: it measured 115 us per step against 92.6 (profiles/history.md)
  (void)ws_on;
  // A frame of n_fft < 1024 samples runs on the 1024-point kernel, zero-extended: X_n[k] of the short frame is bin
  // k * (1024 / n_fft) of the long one, exactly (same sum, the added taps are zero).  So the window is embedded in
  // the first n_fft taps and every filterbank weight moves to the bin it now lives at; geometry (reflect pad, frame
  // count) keeps following the caller's n_fft.  The extra FFT work is the price of one kernel family.
  const int core = n_fft > 1024 ? 2048 : 1024;
  const int ratio = core / n_fft, core_freq = core / 2 + 1;
  plan->core_fft = core;
  std::vector<float> basis_core;
  const float* basis = mel_basis_host;
  if (ratio > 1) {
    basis_core.assign((size_t)n_mels * core_freq, 0.f);
    for (int m = 0; m < n_mels; ++m)
      for (int k = 0; k <= n_fft / 2; ++k) basis_core[(size_t)m * core_freq + (size_t)k * ratio] = mel_basis_host[(size_t)m * (n_fft / 2 + 1) + k];
    basis = basis_core.data();
  }
  std::vector<int4> chan;
  std::vector<float> weights;
  std::vector<int> group_steps, order;
  for (const dmel::VariantOps* v : kVariants) {
    if (v->n_fft != core || (occ_pin && v->occ != occ_pin)) continue;
#ifdef DMEL_WITH_WS
    if (v == &dmel::kVariant_ws_1024 && (occ_pin || !ws_on || center)) continue;
#endif
    band_filterbank(basis, n_mels, core_freq, 32 / v->tf, &chan, &weights, &group_steps);
    order = deal_groups(group_steps, v->mel_warps, v->mel_warps == dmel::kWarps);
    const int wave_len = ((v->tf - 1) * hop_length + core + 7) / 8 * 8;  // whole 16-byte units of float and of int16
    const size_t need = v->smem_need(wave_len, (int)chan.size(), (int)weights.size(), (int)order.size());
    const size_t limit = std::min<size_t>((size_t)max_sm_smem / v->occ - 1024, (size_t)plan->max_smem);  // 1 KB/CTA reserved
    if (need <= limit) {
      plan->variant = v;
      plan->tile_frames = v->tf;
      plan->wave_len = wave_len;
      plan->smem_bytes = need;
      plan->ctas_per_sm = v->occ;
      plan->nnz = (int)weights.size();
      plan->n_chan_pad = (int)chan.size();
      plan->n_order = (int)order.size();
      break;
    }
  }
  if (!plan->variant) {
    const int max_smem = plan->max_smem;
    delete plan;
    return fail(DMEL_ERR_UNSUPPORTED, "geometry needs more than %d bytes of shared memory per CTA", max_smem);
  }

  std::vector<float> window(core, 0.f);  // taps past n_fft stay zero (see above)
  std::copy(window_host, window_host + n_fft, window.begin());
  // inter-pass twiddles W_C^{k1*n2} (C = core/2 complex points, n2 = lane) and unfold twiddles W_{core}^k
  const int n_complex = core / 2, rows = n_complex / 32;
  std::vector<float2> stage_tw((size_t)rows * 32), fold_tw(core / 4 + 1);
  const double two_pi = 6.283185307179586476925286766559;
  for (int k1 = 0; k1 < rows; ++k1)
    for (int n2 = 0; n2 < 32; ++n2) {
      const double a = -two_pi * double((k1 * n2) % n_complex) / n_complex;
      stage_tw[k1 * 32 + n2] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
  for (int k = 0; k <= core / 4; ++k) {
    const double a = -two_pi * double(k) / core;
    fold_tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  std::vector<float> window_pcm(window);
  for (float& w : window_pcm) w *= 1.0f / 32768.0f;  // exact: a power of two
  e = upload(&plan->d_window, window);
  if (e == cudaSuccess) e = upload(&plan->d_window_pcm, window_pcm);
  if (e == cudaSuccess) e = upload(&plan->d_stage_tw, stage_tw);
  if (e == cudaSuccess) e = upload(&plan->d_fold_tw, fold_tw);
  if (e == cudaSuccess) e = upload(&plan->d_chan, chan);
  if (e == cudaSuccess) e = upload(&plan->d_weights, weights);
  if (e == cudaSuccess) e = upload(&plan->d_order, order);
  if (e == cudaSuccess) e = upload(&plan->d_sched, std::vector<int>(2 * kSchedSlots, 0));
  if (e != cudaSuccess) {
    dmel_plan_destroy(plan);
    return fail(DMEL_ERR_CUDA, "plan upload failed: %s", cudaGetErrorString(e));
  }
  *out = plan;
  return DMEL_OK;
}

void dmel_plan_destroy(dmel_plan* plan) {
  if (!plan) return;
  DeviceGuard guard(plan->device);
  cudaFree(plan->d_window);
  cudaFree(plan->d_window_pcm);
  cudaFree(plan->d_stage_tw);
  cudaFree(plan->d_fold_tw);
  cudaFree(plan->d_chan);
  cudaFree(plan->d_weights);
  cudaFree(plan->d_order);
  cudaFree(plan->d_sched);
  for (int i = 0; i < 2; ++i) {
    cudaFree(plan->d_wav[i]);
    cudaFree(plan->d_codes[i]);
    cudaFree(plan->d_len[i]);
    if (plan->streams[i]) cudaStreamDestroy(plan->streams[i]);
  }
  if (plan->stats_ready) cudaEventDestroy(plan->stats_ready);
  for (int i = 0; i < 2; ++i) {
    if (plan->chunk_in[i]) cudaEventDestroy(plan->chunk_in[i]);
    if (plan->chunk_used[i]) cudaEventDestroy(plan->chunk_used[i]);
  }
  cudaFree(plan->d_lo);
  cudaFree(plan->d_scale);
  delete plan;
}

int dmel_plan_describe(const dmel_plan* plan, char* buf, size_t buf_len) {
  if (!plan || !buf || buf_len == 0) return fail(DMEL_ERR_INVALID, "plan / buf is null");
  snprintf(buf, buf_len,
           "{\"n_fft\": %d, \"core_fft\": %d, \"hop\": %d, \"n_mels\": %d, \"center\": %d, \"tile_frames\": %d, "
           "\"ctas_per_sm\": %d, \"smem_bytes\": %zu, \"banded_weights\": %d, \"n_chan_pad\": %d, "
           "\"wave_len\": %d, \"sm_count\": %d}",
           plan->n_fft, plan->core_fft, plan->hop, plan->n_mels, plan->center, plan->tile_frames, plan->ctas_per_sm,
           plan->smem_bytes, plan->nnz, plan->n_chan_pad, plan->wave_len, plan->sm_count);
  return DMEL_OK;
}

long long dmel_plan_num_frames(const dmel_plan* plan, long long n_samples) {
  if (!plan) return 0;
  return num_frames(plan, n_samples);
}

int dmel_logmel_f32(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                    long long row_stride, float* logmel_dev, void* stream) {
  FusedParams p;
  int grid = 0;
  int rc = prepare_fused(plan, wav_dev, n_rows, n_samples, row_stride, &p, &grid);
  if (rc != DMEL_OK) return rc;
  if (!logmel_dev) return fail(DMEL_ERR_INVALID, "logmel_dev is null");
  if (n_rows == 0) return DMEL_OK;
  p.logmel = logmel_dev;
  DeviceGuard guard(plan->device);
  DMEL_CUDA(launch_fused_any(plan, p, grid, (cudaStream_t)stream));
  return DMEL_OK;
}

int dmel_logmel_masked(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                       long long row_stride, const int32_t* lengths_dev, int out_dtype, void* out_dev,
                       float* row_sum_dev, void* stream) {
  FusedParams p;
  int grid = 0;
  int rc = prepare_fused(plan, wav_dev, n_rows, n_samples, row_stride, &p, &grid);
  if (rc != DMEL_OK) return rc;
  if (!out_dev) return fail(DMEL_ERR_INVALID, "out_dev is null");
  if (out_dtype != DMEL_DTYPE_F32 && out_dtype != DMEL_DTYPE_BF16)
    return fail(DMEL_ERR_INVALID, "out_dtype must be DMEL_DTYPE_F32 or DMEL_DTYPE_BF16, got %d", out_dtype);
  if (n_rows == 0) return DMEL_OK;
  p.logmel = static_cast<float*>(out_dev);
  p.lengths = lengths_dev;
  p.mask_invalid = lengths_dev != nullptr;
  p.row_sum = row_sum_dev;
  DeviceGuard guard(plan->device);
  DMEL_CUDA(launch_fused_any(plan, p, grid, (cudaStream_t)stream, out_dtype == DMEL_DTYPE_BF16));
  return DMEL_OK;
}

int dmel_minmax_f32(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                    long long row_stride, const int32_t* lengths_dev, float* min_dev, float* max_dev,
                    void* stream) {
  FusedParams p;
  int grid = 0;
  int rc = prepare_fused(plan, wav_dev, n_rows, n_samples, row_stride, &p, &grid);
  if (rc != DMEL_OK) return rc;
  if (!min_dev || !max_dev) return fail(DMEL_ERR_INVALID, "min_dev / max_dev is null");
  if (n_rows == 0) return DMEL_OK;
  p.lengths = lengths_dev;
  p.run_min = min_dev;
  p.run_max = max_dev;
  DeviceGuard guard(plan->device);
  DMEL_CUDA(launch_fused_any(plan, p, grid, (cudaStream_t)stream));
  return DMEL_OK;
}

int dmel_encode_u8(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                   long long row_stride, const int32_t* lengths_dev, const float* lo_dev,
                   const float* scale_dev, int n_bins, uint8_t* codes_dev, float* logmel_dev,
                   unsigned long long* near_edge_dev, float edge_eps, void* stream) {
  FusedParams p;
  int grid = 0;
  int rc = prepare_fused(plan, wav_dev, n_rows, n_samples, row_stride, &p, &grid);
  if (rc != DMEL_OK) return rc;
  if ((rc = check_bins(n_bins)) != DMEL_OK) return rc;
  if (!lo_dev || !scale_dev || !codes_dev) return fail(DMEL_ERR_INVALID, "lo_dev / scale_dev / codes_dev is null");
  if (n_rows == 0) return DMEL_OK;
  p.lengths = lengths_dev;
  p.q_lo = lo_dev;
  p.q_scale = scale_dev;
  p.n_bins = n_bins;
  p.kmax = float(n_bins - 1);
  p.codes = codes_dev;
  p.logmel = logmel_dev;
  p.near_edge = near_edge_dev;
  p.edge_eps = edge_eps;
  DeviceGuard guard(plan->device);
  DMEL_CUDA(launch_fused_any(plan, p, grid, (cudaStream_t)stream));
  return DMEL_OK;
}

int dmel_logmel_minmax_f32(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                           long long row_stride, const int32_t* lengths_dev, float* logmel_dev, float* min_dev,
                           float* max_dev, void* stream) {
  FusedParams p;
  int grid = 0;
  int rc = prepare_fused(plan, wav_dev, n_rows, n_samples, row_stride, &p, &grid);
  if (rc != DMEL_OK) return rc;
  if (!logmel_dev || !min_dev || !max_dev) return fail(DMEL_ERR_INVALID, "logmel_dev / min_dev / max_dev is null");
  if (n_rows == 0) return DMEL_OK;
  p.lengths = lengths_dev;
  p.logmel = logmel_dev;
  p.mask_invalid = lengths_dev != nullptr;  // frames past a row's length are written as 0 and never computed
  p.run_min = min_dev;
  p.run_max = max_dev;
  DeviceGuard guard(plan->device);
  DMEL_CUDA(launch_fused_any(plan, p, grid, (cudaStream_t)stream));
  return DMEL_OK;
}

int dmel_encode_decode_u8(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                          long long row_stride, const int32_t* lengths_dev, const float* lo_dev,
                          const float* scale_dev, const float* step_dev, int n_bins, uint8_t* codes_dev,
                          float* mel_hat_dev, void* stream) {
  FusedParams p;
  int grid = 0;
  int rc = prepare_fused(plan, wav_dev, n_rows, n_samples, row_stride, &p, &grid);
  if (rc != DMEL_OK) return rc;
  if ((rc = check_bins(n_bins)) != DMEL_OK) return rc;
  if (!lo_dev || !scale_dev || !step_dev || !codes_dev || !mel_hat_dev)
    return fail(DMEL_ERR_INVALID, "lo_dev / scale_dev / step_dev / codes_dev / mel_hat_dev is null");
  if (n_rows == 0) return DMEL_OK;
  p.lengths = lengths_dev;
  p.q_lo = lo_dev;
  p.q_scale = scale_dev;
  p.q_step = step_dev;
  p.n_bins = n_bins;
  p.kmax = float(n_bins - 1);
  p.codes = codes_dev;
  p.dequant = mel_hat_dev;
  DeviceGuard guard(plan->device);
  DMEL_CUDA(launch_fused_any(plan, p, grid, (cudaStream_t)stream));
  return DMEL_OK;
}

int dmel_encode_pcm16_u8(dmel_plan* plan, const int16_t* wav_dev, long long n_rows, long long n_samples,
                         long long row_stride, const int32_t* lengths_dev, const float* lo_dev,
                         const float* scale_dev, int n_bins, uint8_t* codes_dev, void* stream) {
  FusedParams p;
  int grid = 0;
  int rc = prepare_fused(plan, reinterpret_cast<const float*>(wav_dev), n_rows, n_samples, row_stride, &p, &grid);
  if (rc != DMEL_OK) return rc;
  if ((rc = check_bins(n_bins)) != DMEL_OK) return rc;
  if (!lo_dev || !scale_dev || !codes_dev) return fail(DMEL_ERR_INVALID, "lo_dev / scale_dev / codes_dev is null");
  if (!plan->variant->lean)
    return fail(DMEL_ERR_UNSUPPORTED, "int16 input needs the register-lean kernel variant, which does not fit this geometry");
  if (n_rows == 0) return DMEL_OK;
  p.window = plan->d_window_pcm;
  p.lengths = lengths_dev;
  p.q_lo = lo_dev;
  p.q_scale = scale_dev;
  p.n_bins = n_bins;
  p.kmax = float(n_bins - 1);
  p.codes = codes_dev;
  DeviceGuard guard(plan->device);
  DMEL_CUDA(launch_fused_any(plan, p, grid, (cudaStream_t)stream, false, true));
  return DMEL_OK;
}

int dmel_encode_frames_u8(dmel_plan* plan, const float* wav_dev, long long n_rows, long long row_stride,
                          long long src_base, long long n_samples, long long t_begin, long long t_count,
                          const float* lo_dev, const float* scale_dev, int n_bins, uint8_t* codes_dev,
                          float* logmel_dev, void* stream) {
  FusedParams p;
  int grid = 0;
  int rc = prepare_window(plan, wav_dev, n_rows, n_samples, row_stride, src_base, t_begin, t_count, &p, &grid);
  if (rc != DMEL_OK) return rc;
  if (!codes_dev && !logmel_dev) return fail(DMEL_ERR_INVALID, "codes_dev and logmel_dev are both null");
  if (codes_dev) {
    if ((rc = check_bins(n_bins)) != DMEL_OK) return rc;
    if (!lo_dev || !scale_dev) return fail(DMEL_ERR_INVALID, "lo_dev / scale_dev is null");
    p.q_lo = lo_dev;
    p.q_scale = scale_dev;
    p.n_bins = n_bins;
    p.kmax = float(n_bins - 1);
    p.codes = codes_dev;
  }
  if (n_rows == 0) return DMEL_OK;
  p.logmel = logmel_dev;
  DeviceGuard guard(plan->device);
  DMEL_CUDA(launch_fused_any(plan, p, grid, (cudaStream_t)stream));
  return DMEL_OK;
}

int dmel_run(dmel_plan* plan, const dmel_io* io, void* stream) {
  if (!io) return fail(DMEL_ERR_INVALID, "io is null");
  if (io->struct_size != sizeof(dmel_io))
    return fail(DMEL_ERR_INVALID, "dmel_io is %zu bytes in this library, the caller passed %zu", sizeof(dmel_io), io->struct_size);
  FusedParams p;
  int grid = 0;
  const bool ragged = io->offsets_dev != nullptr;
  const bool own = ragged || io->own_length != 0;
  const long long stride = ragged ? io->n_samples : io->row_stride;
  int rc = prepare_window(plan, static_cast<const float*>(io->wav_dev), io->n_rows, io->n_samples, stride, 0, 0, -1, &p, &grid);
  if (rc != DMEL_OK) return rc;
  if (own) {
    if (!ragged && !io->lengths_dev) return fail(DMEL_ERR_INVALID, "own_length needs lengths_dev (or the ragged layout)");
    if (io->min_row_samples <= plan->pad_inner)
      return fail(DMEL_ERR_INVALID, "reflect padding of %d needs more than %d samples per utterance, the shortest has %lld "
                  "(the reference's F.pad raises here too)", plan->pad_inner, plan->pad_inner, io->min_row_samples);
    if (plan->pad_outer) return fail(DMEL_ERR_UNSUPPORTED, "own-length rows with center=True are not supported");
  }
  const bool want_codes = io->codes_dev != nullptr, want_stats = io->min_dev != nullptr || io->max_dev != nullptr;
  if (!want_codes && !io->logmel_dev && !want_stats) return fail(DMEL_ERR_INVALID, "no output requested");
  if (want_codes && want_stats) return fail(DMEL_ERR_INVALID, "codes and statistics cannot be produced by one call");
  if (want_stats && (!io->min_dev || !io->max_dev)) return fail(DMEL_ERR_INVALID, "min_dev / max_dev is null");
  if (io->mel_hat_dev && !want_codes) return fail(DMEL_ERR_INVALID, "mel_hat_dev needs codes_dev");
  if (want_codes) {
    if ((rc = check_bins(io->n_bins)) != DMEL_OK) return rc;
    if (!io->lo_dev || !io->scale_dev || (io->mel_hat_dev && !io->step_dev))
      return fail(DMEL_ERR_INVALID, "lo_dev / scale_dev / step_dev is null");
  }
  if (io->wav_is_pcm16) {
    if (!plan->variant->lean) return fail(DMEL_ERR_UNSUPPORTED, "int16 input needs the register-lean kernel variant, which does not fit this geometry");
    if (io->row_gain_dev) return fail(DMEL_ERR_UNSUPPORTED, "row_gain_dev applies to float32 waveforms only");
    if (!(want_codes && !io->mel_hat_dev && !io->logmel_dev)) return fail(DMEL_ERR_UNSUPPORTED, "int16 input is built for the code output only");
    p.window = plan->d_window_pcm;
  }
  if (io->n_rows == 0) return DMEL_OK;
  p.offsets = io->offsets_dev;
  p.own_length = own ? 1 : 0;
  p.lengths = io->lengths_dev;
  p.row_gain = io->row_gain_dev;
  p.logmel = static_cast<float*>(io->logmel_dev);
  p.mask_invalid = (io->mask_invalid || own) && (io->lengths_dev != nullptr || own);
  p.run_min = io->min_dev;
  p.run_max = io->max_dev;
  if (want_codes) {
    p.q_lo = io->lo_dev;
    p.q_scale = io->scale_dev;
    p.q_step = io->step_dev;
    p.n_bins = io->n_bins;
    p.kmax = float(io->n_bins - 1);
    p.codes = io->codes_dev;
    p.dequant = io->mel_hat_dev;
  }
  DeviceGuard guard(plan->device);
  const cudaError_t e = launch_fused_any(plan, p, grid, (cudaStream_t)stream, io->logmel_is_bf16 != 0, io->wav_is_pcm16 != 0);
  if (e == cudaErrorInvalidValue) return fail(DMEL_ERR_UNSUPPORTED, "this combination of outputs is not built (see logmel_kernel.cuh MODE list)");
  DMEL_CUDA(e);
  return DMEL_OK;
}

int dmel_row_peak_gain_f32(const float* wav_dev, long long n_rows, long long n_samples, long long row_stride,
                           const long long* offsets_dev, const int32_t* lengths_dev, float target_peak,
                           float* gain_dev, void* stream) {
  if (!wav_dev || !gain_dev) return fail(DMEL_ERR_INVALID, "wav_dev / gain_dev is null");
  if (n_rows < 0 || n_samples <= 0 || n_samples > (1LL << 30) || (!offsets_dev && row_stride < n_samples))
    return fail(DMEL_ERR_INVALID, "bad waveform shape: rows=%lld samples=%lld stride=%lld", n_rows, n_samples, row_stride);
  if (n_rows > 65535) return fail(DMEL_ERR_INVALID, "at most 65535 rows per call, got %lld", n_rows);
  if (n_rows == 0) return DMEL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int dev = device_of(wav_dev);
  DeviceGuard guard(dev);
  // the gains double as the scratch for the running maxima (bit patterns of |x|)
  DMEL_CUDA(cudaMemsetAsync(gain_dev, 0, (size_t)n_rows * sizeof(float), st));
  // about eight blocks per SM in all, each with at least kAbsmaxMinSpan samples of its row
  const long long want = (8LL * sm_count_of(dev) + n_rows - 1) / n_rows;
  const long long most = (n_samples + dmel::kAbsmaxMinSpan - 1) / dmel::kAbsmaxMinSpan;
  const dim3 grid((unsigned)std::max(1LL, std::min(want, most)), (unsigned)n_rows);
  DMEL_CUDA(launch_pdl(dmel::row_absmax_kernel, grid, dim3(dmel::kAbsmaxThreads), 0, st, wav_dev, offsets_dev, lengths_dev,
                       row_stride, (int)n_samples, reinterpret_cast<unsigned*>(gain_dev)));
  DMEL_CUDA(launch_pdl(dmel::row_gain_kernel, dim3((unsigned)((n_rows + 127) / 128)), dim3(128), 0, st,
                       reinterpret_cast<const unsigned*>(gain_dev), target_peak, gain_dev, (int)n_rows));
  return DMEL_OK;
}

// ---------------------------------------------------------------------------
// Streaming encoder state (BASELINE configs[3]): per-stream history buffer on the device, counters
// on the host.  One push = one copy of the chunk into the buffer + one windowed launch.
// ---------------------------------------------------------------------------
struct dmel_stream {
  dmel_plan* plan = nullptr;
  int n_streams = 0;
  long long capacity = 0;  // samples per stream row (multiple of 4)
  float* buf = nullptr;    // (n_streams, capacity)
  float* scratch = nullptr;  // same shape, only for overlapping compactions
  long long base = 0;      // virtual sample index of buf[:, 0]
  long long seen = 0;      // samples received per stream
  long long t_next = 0;    // next frame to emit
  // quantiser and CUDA stream bound once (dmel_stream_bind) so that the per-chunk call carries four arguments
  const float* lo_dev = nullptr;
  const float* scale_dev = nullptr;
  int n_bins = 0;
  void* cuda_stream = nullptr;
};

namespace {

// frames that are complete once `seen` samples have arrived (their last tap is inside the data)
long long stream_frames_ready(const dmel_stream* s, long long seen) {
  const dmel_plan* pl = s->plan;
  if (seen + pl->pad_inner < pl->n_fft || seen <= pl->pad_inner) return 0;
  return (seen + pl->pad_inner - pl->n_fft) / pl->hop + 1;
}

int stream_emit(dmel_stream* s, long long t_end, const float* lo_dev, const float* scale_dev, int n_bins,
                uint8_t* codes_dev, long long codes_frames, long long* n_frames_out, void* stream) {
  const long long count = t_end - s->t_next;
  if (n_frames_out) *n_frames_out = count > 0 ? count : 0;
  if (count <= 0) return DMEL_OK;
  if (!codes_dev || codes_frames != count)
    return fail(DMEL_ERR_INVALID, "codes buffer holds %lld frames per channel, this call emits %lld "
                "(size it with dmel_stream_pending)", codes_frames, count);
  int rc = dmel_encode_frames_u8(s->plan, s->buf, s->n_streams, s->capacity, s->base, s->seen, s->t_next, count,
                                 lo_dev, scale_dev, n_bins, codes_dev, nullptr, stream);
  if (rc == DMEL_OK) s->t_next = t_end;
  return rc;
}

}  // namespace

int dmel_stream_create(dmel_plan* plan, int n_streams, long long capacity_samples, dmel_stream** out) {
  if (!out) return fail(DMEL_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!plan) return fail(DMEL_ERR_INVALID, "plan is null");
  if (plan->pad_outer) return fail(DMEL_ERR_UNSUPPORTED, "streaming with center=True is not supported");
  if (n_streams < 1) return fail(DMEL_ERR_INVALID, "n_streams must be >= 1, got %d", n_streams);
  dmel_stream* s = new (std::nothrow) dmel_stream();
  if (!s) return fail(DMEL_ERR_INVALID, "out of host memory");
  s->plan = plan;
  s->n_streams = n_streams;
  s->capacity = std::max<long long>(capacity_samples, 4LL * plan->n_fft) / 4 * 4;
  DeviceGuard guard(plan->device);
  const size_t bytes = (size_t)n_streams * s->capacity * sizeof(float);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&s->buf), bytes);
  if (e == cudaSuccess) e = cudaMemset(s->buf, 0, bytes);
  if (e != cudaSuccess) {
    cudaFree(s->buf);
    delete s;
    return fail(DMEL_ERR_CUDA, "stream buffer allocation failed: %s", cudaGetErrorString(e));
  }
  *out = s;
  return DMEL_OK;
}

void dmel_stream_destroy(dmel_stream* s) {
  if (!s) return;
  DeviceGuard guard(s->plan->device);
  cudaFree(s->buf);
  cudaFree(s->scratch);
  delete s;
}

int dmel_stream_reset(dmel_stream* s) {
  if (!s) return fail(DMEL_ERR_INVALID, "stream is null");
  s->base = s->seen = s->t_next = 0;
  return DMEL_OK;
}

long long dmel_stream_pending(const dmel_stream* s, long long n_incoming, int at_end) {
  if (!s || n_incoming < 0) return -1;
  const long long seen = s->seen + n_incoming;
  const long long t_end = at_end ? num_frames(s->plan, seen) : stream_frames_ready(s, seen);
  return std::max<long long>(t_end - s->t_next, 0);
}

// make room for n more samples per stream (drops what no future frame needs) and say where they go
static int stream_reserve(dmel_stream* s, long long n, cudaStream_t st) {
  if (s->seen - s->base + n <= s->capacity) return DMEL_OK;
  long long keep_from = std::max<long long>(0, s->t_next * s->plan->hop - s->plan->pad_inner) / 4 * 4;
  keep_from = std::max(keep_from, s->base);
  const long long live = s->seen - keep_from, shift = keep_from - s->base;
  if (live + n > s->capacity)
    return fail(DMEL_ERR_INVALID, "chunk of %lld samples does not fit a stream buffer of %lld (%lld live)", n, s->capacity, live);
  if (shift > 0 && live > 0) {
    const size_t pitch = (size_t)s->capacity * sizeof(float), width = (size_t)live * sizeof(float);
    if (live <= shift) {
      DMEL_CUDA(cudaMemcpy2DAsync(s->buf, pitch, s->buf + shift, pitch, width, s->n_streams, cudaMemcpyDeviceToDevice, st));
    } else {  // source and destination overlap: go through the scratch copy
      if (!s->scratch) DMEL_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->scratch), (size_t)s->n_streams * pitch));
      DMEL_CUDA(cudaMemcpy2DAsync(s->scratch, pitch, s->buf + shift, pitch, width, s->n_streams, cudaMemcpyDeviceToDevice, st));
      DMEL_CUDA(cudaMemcpy2DAsync(s->buf, pitch, s->scratch, pitch, width, s->n_streams, cudaMemcpyDeviceToDevice, st));
    }
  }
  s->base = keep_from;
  return DMEL_OK;
}

int dmel_stream_push(dmel_stream* s, const float* chunk, long long n, long long chunk_stride,
                     const float* lo_dev, const float* scale_dev, int n_bins, uint8_t* codes_dev,
                     long long codes_frames, long long* n_frames_out, void* stream) {
  if (!s) return fail(DMEL_ERR_INVALID, "stream is null");
  if (n_frames_out) *n_frames_out = 0;
  if (n < 0 || (n > 0 && (!chunk || chunk_stride < n)))
    return fail(DMEL_ERR_INVALID, "bad chunk: n=%lld stride=%lld", n, chunk_stride);
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard guard(s->plan->device);
  {
    int rc_reserve = stream_reserve(s, n, st);  // drops the samples no future frame needs; the new base stays 16-byte aligned
    if (rc_reserve != DMEL_OK) return rc_reserve;
  }
  if (n > 0) {
    float* dst = s->buf + (s->seen - s->base);
    if (s->n_streams == 1) {
      DMEL_CUDA(cudaMemcpyAsync(dst, chunk, (size_t)n * sizeof(float), cudaMemcpyDefault, st));
    } else {
      DMEL_CUDA(cudaMemcpy2DAsync(dst, (size_t)s->capacity * sizeof(float), chunk, (size_t)chunk_stride * sizeof(float),
                                  (size_t)n * sizeof(float), s->n_streams, cudaMemcpyDefault, st));
    }
    s->seen += n;
  }
  return stream_emit(s, stream_frames_ready(s, s->seen), lo_dev, scale_dev, n_bins, codes_dev, codes_frames, n_frames_out,
                     stream);
}

int dmel_stream_bind(dmel_stream* s, const float* lo_dev, const float* scale_dev, int n_bins, void* stream) {
  if (!s) return fail(DMEL_ERR_INVALID, "stream is null");
  if (!lo_dev || !scale_dev) return fail(DMEL_ERR_INVALID, "lo_dev / scale_dev is null");
  int rc = check_bins(n_bins);
  if (rc != DMEL_OK) return rc;
  s->lo_dev = lo_dev;
  s->scale_dev = scale_dev;
  s->n_bins = n_bins;
  s->cuda_stream = stream;
  return DMEL_OK;
}

int dmel_stream_input(dmel_stream* s, long long n, float** where_dev, long long* row_stride) {
  if (!s || !where_dev) return fail(DMEL_ERR_INVALID, "stream / where_dev is null");
  if (n <= 0) return fail(DMEL_ERR_INVALID, "bad chunk length %lld", n);
  DeviceGuard guard(s->plan->device);
  int rc = stream_reserve(s, n, (cudaStream_t)s->cuda_stream);
  if (rc != DMEL_OK) return rc;
  *where_dev = s->buf + (s->seen - s->base);
  if (row_stride) *row_stride = s->capacity;
  return DMEL_OK;
}

int dmel_stream_commit(dmel_stream* s, long long n, uint8_t* codes_dev, long long codes_frames, long long* n_frames_out) {
  if (!s) return fail(DMEL_ERR_INVALID, "stream is null");
  if (n_frames_out) *n_frames_out = 0;
  if (!s->lo_dev) return fail(DMEL_ERR_INVALID, "call dmel_stream_bind first");
  if (n < 0 || s->seen - s->base + n > s->capacity) return fail(DMEL_ERR_INVALID, "commit of %lld samples without a matching dmel_stream_input", n);
  s->seen += n;
  return stream_emit(s, stream_frames_ready(s, s->seen), s->lo_dev, s->scale_dev, s->n_bins, codes_dev, codes_frames, n_frames_out,
                     s->cuda_stream);
}

int dmel_stream_flush(dmel_stream* s, const float* lo_dev, const float* scale_dev, int n_bins, uint8_t* codes_dev,
                      long long codes_frames, long long* n_frames_out, void* stream) {
  if (!s) return fail(DMEL_ERR_INVALID, "stream is null");
  if (n_frames_out) *n_frames_out = 0;
  if (s->seen <= s->plan->pad_inner)
    return fail(DMEL_ERR_INVALID, "stream of %lld samples is shorter than the reflect pad %d", s->seen, s->plan->pad_inner);
  int rc = stream_emit(s, num_frames(s->plan, s->seen), lo_dev, scale_dev, n_bins, codes_dev, codes_frames, n_frames_out,
                       stream);
  if (rc == DMEL_OK) s->base = s->seen = s->t_next = 0;
  return rc;
}

static int encode_host_impl(dmel_plan* plan, const void* wav_host_v, int elem, long long n_rows, long long n_samples,
                        long long row_stride, const int32_t* lengths_host, const float* lo_host,
                        const float* scale_host, int n_bins, uint8_t* codes_host) {
  if (!plan) return fail(DMEL_ERR_INVALID, "plan is null");
  const char* wav_host = static_cast<const char*>(wav_host_v);
  if (!wav_host || !codes_host || !lo_host || !scale_host)
    return fail(DMEL_ERR_INVALID, "wav_host / codes_host / lo_host / scale_host is null");
  int rc = check_bins(n_bins);
  if (rc != DMEL_OK) return rc;
  if (n_rows < 0 || n_samples <= 0 || row_stride < n_samples)
    return fail(DMEL_ERR_INVALID, "bad waveform shape: rows=%lld samples=%lld stride=%lld", n_rows, n_samples, row_stride);
  if (n_samples <= plan->pad_inner)
    return fail(DMEL_ERR_INVALID, "reflect padding of %d needs more than %d samples per row, got %lld",
                plan->pad_inner, plan->pad_inner, n_samples);
  const long long T = num_frames(plan, n_samples);
  if (T <= 0) return fail(DMEL_ERR_INVALID, "row of %lld samples is shorter than one frame", n_samples);
  if (n_rows == 0) return DMEL_OK;
  DeviceGuard guard(plan->device);
  for (int i = 0; i < 2; ++i)
    if (!plan->streams[i]) DMEL_CUDA(cudaStreamCreateWithFlags(&plan->streams[i], cudaStreamNonBlocking));
  if (!plan->d_lo) DMEL_CUDA(cudaMalloc((void**)&plan->d_lo, plan->n_mels * sizeof(float)));
  if (!plan->d_scale) DMEL_CUDA(cudaMalloc((void**)&plan->d_scale, plan->n_mels * sizeof(float)));
  // rows per chunk: about eight chunks per call so copies and kernels of neighbouring chunks overlap and the
  // un-overlapped tail (last kernel + last D2H) stays short, but 8 to 16 MiB of waveform each (DMEL_HOST_CHUNK_MB pins
  // it).  Measured: 61 MB of float32 in 8 MB chunks 1.24 ms, in 4 MB chunks 1.30, in 2 MB chunks 1.56; 31 MB of int16
  // in 8 MB chunks 0.79 ms, in 4 MB chunks 0.88 - below 8 MB the copies themselves slow down.
  const long long row_bytes = n_samples * elem;
  long long chunk_bytes = std::min<long long>(16ll << 20, std::max<long long>(8ll << 20, row_bytes * n_rows / 8));
  if (const char* env = std::getenv("DMEL_HOST_CHUNK_MB")) chunk_bytes = (long long)std::max(1, std::atoi(env)) << 20;
  long long chunk_rows = std::max<long long>(1, chunk_bytes / row_bytes);
  chunk_rows = std::min(chunk_rows, n_rows);
  const size_t wav_need = (size_t)chunk_rows * n_samples;
  const size_t codes_need = (size_t)chunk_rows * plan->n_mels * T;
  if (wav_need > plan->wav_cap) {
    for (int i = 0; i < 2; ++i) {
      cudaFree(plan->d_wav[i]);
      plan->d_wav[i] = nullptr;
      DMEL_CUDA(cudaMalloc((void**)&plan->d_wav[i], wav_need * sizeof(float)));  // sized for float, int16 uses half
    }
    plan->wav_cap = wav_need;
  }
  if (codes_need > plan->codes_cap) {
    for (int i = 0; i < 2; ++i) {
      cudaFree(plan->d_codes[i]);
      plan->d_codes[i] = nullptr;
      DMEL_CUDA(cudaMalloc((void**)&plan->d_codes[i], codes_need));
    }
    plan->codes_cap = codes_need;
  }
  if ((size_t)chunk_rows > plan->len_cap) {
    for (int i = 0; i < 2; ++i) {
      cudaFree(plan->d_len[i]);
      plan->d_len[i] = nullptr;
      DMEL_CUDA(cudaMalloc((void**)&plan->d_len[i], chunk_rows * sizeof(int32_t)));
    }
    plan->len_cap = chunk_rows;
  }
  // The statistics go up only when they differ from what the device already holds: a tokeniser encodes thousands of
  // batches per calibration, and two copies from pageable host memory are 20-30 us of every call.  Both streams
  // were synchronised when the previous call returned, so nothing still reads d_lo / d_scale; the second stream waits
  // for the upload through an event instead of a host synchronisation.
  const size_t nm = (size_t)plan->n_mels;
  const bool same_stats = plan->host_stats.size() == 2 * nm &&
                          std::memcmp(plan->host_stats.data(), lo_host, nm * sizeof(float)) == 0 &&
                          std::memcmp(plan->host_stats.data() + nm, scale_host, nm * sizeof(float)) == 0;
  if (!same_stats) {
    plan->host_stats.clear();  // stays empty if an upload fails
    DMEL_CUDA(cudaMemcpyAsync(plan->d_lo, lo_host, nm * sizeof(float), cudaMemcpyHostToDevice, plan->streams[0]));
    DMEL_CUDA(cudaMemcpyAsync(plan->d_scale, scale_host, nm * sizeof(float), cudaMemcpyHostToDevice, plan->streams[0]));
    if (!plan->stats_ready) DMEL_CUDA(cudaEventCreateWithFlags(&plan->stats_ready, cudaEventDisableTiming));
    DMEL_CUDA(cudaEventRecord(plan->stats_ready, plan->streams[0]));
    DMEL_CUDA(cudaStreamWaitEvent(plan->streams[1], plan->stats_ready, 0));
    plan->host_stats.assign(lo_host, lo_host + nm);
    plan->host_stats.insert(plan->host_stats.end(), scale_host, scale_host + nm);
  }
  // From here on copies into the caller's host buffers may be in flight on both streams: every exit, error or
  // not, goes through the two synchronisations below.
  // Stream 0 carries every host-to-device copy, back to back; stream 1 the kernels and the copies back.  (With each
  // chunk's copy, kernel and copy-back in one stream and the chunks alternating between two streams, as in the first
  // version, consecutive copies sit on different streams and the copy engine idles ~14 us between them.)
  // DMEL_HOST_PIPELINE=alternate selects that first form.
  static const bool alternate = std::getenv("DMEL_HOST_PIPELINE") && std::strcmp(std::getenv("DMEL_HOST_PIPELINE"), "alternate") == 0;
  for (int i = 0; i < 2; ++i) {
    if (!plan->chunk_in[i]) DMEL_CUDA(cudaEventCreateWithFlags(&plan->chunk_in[i], cudaEventDisableTiming));
    if (!plan->chunk_used[i]) DMEL_CUDA(cudaEventCreateWithFlags(&plan->chunk_used[i], cudaEventDisableTiming));
  }
  auto queue_chunk = [&](long long r0, long long rows, int slot, long long index) -> int {
    cudaStream_t s_in = alternate ? plan->streams[slot] : plan->streams[0];
    cudaStream_t s_run = alternate ? plan->streams[slot] : plan->streams[1];
    // the slot's buffers are free once the kernel of chunk index - 2 has run (alternate: stream order says so)
    if (!alternate && index >= 2) DMEL_CUDA(cudaStreamWaitEvent(s_in, plan->chunk_used[slot], 0));
    if (row_stride == n_samples)
      DMEL_CUDA(cudaMemcpyAsync(plan->d_wav[slot], wav_host + r0 * row_stride * elem, (size_t)rows * n_samples * elem,
                                cudaMemcpyHostToDevice, s_in));
    else
      DMEL_CUDA(cudaMemcpy2DAsync(plan->d_wav[slot], n_samples * elem, wav_host + r0 * row_stride * elem, row_stride * elem,
                                  n_samples * elem, rows, cudaMemcpyHostToDevice, s_in));
    const int32_t* len_dev = nullptr;
    if (lengths_host) {
      DMEL_CUDA(cudaMemcpyAsync(plan->d_len[slot], lengths_host + r0, rows * sizeof(int32_t), cudaMemcpyHostToDevice, s_in));
      len_dev = plan->d_len[slot];
    }
    if (!alternate) {
      DMEL_CUDA(cudaEventRecord(plan->chunk_in[slot], s_in));
      DMEL_CUDA(cudaStreamWaitEvent(s_run, plan->chunk_in[slot], 0));
    }
    const int rc_k = elem == 2 ? dmel_encode_pcm16_u8(plan, reinterpret_cast<const int16_t*>(plan->d_wav[slot]), rows, n_samples,
                                                      n_samples, len_dev, plan->d_lo, plan->d_scale, n_bins, plan->d_codes[slot], s_run)
                               : dmel_encode_u8(plan, plan->d_wav[slot], rows, n_samples, n_samples, len_dev, plan->d_lo,
                                                plan->d_scale, n_bins, plan->d_codes[slot], nullptr, nullptr, 0.f, s_run);
    if (rc_k != DMEL_OK) return rc_k;
    if (!alternate) DMEL_CUDA(cudaEventRecord(plan->chunk_used[slot], s_run));
    DMEL_CUDA(cudaMemcpyAsync(codes_host + (size_t)r0 * plan->n_mels * T, plan->d_codes[slot],
                              (size_t)rows * plan->n_mels * T, cudaMemcpyDeviceToHost, s_run));
    return DMEL_OK;
  };
  int slot = 0;
  long long index = 0;
  rc = DMEL_OK;
  for (long long r0 = 0; r0 < n_rows && rc == DMEL_OK; r0 += chunk_rows, slot ^= 1, ++index)
    rc = queue_chunk(r0, std::min(chunk_rows, n_rows - r0), slot, index);
  const std::string queued_error = rc == DMEL_OK ? std::string() : g_last_error;
  const cudaError_t s0 = cudaStreamSynchronize(plan->streams[0]);
  const cudaError_t s1 = cudaStreamSynchronize(plan->streams[1]);
  if (rc != DMEL_OK) {
    g_last_error = queued_error;
    return rc;
  }
  if (s0 != cudaSuccess || s1 != cudaSuccess)
    return fail(DMEL_ERR_CUDA, "stream synchronisation failed: %s", cudaGetErrorString(s0 != cudaSuccess ? s0 : s1));
  return DMEL_OK;
}

int dmel_encode_host_u8(dmel_plan* plan, const float* wav_host, long long n_rows, long long n_samples,
                        long long row_stride, const int32_t* lengths_host, const float* lo_host,
                        const float* scale_host, int n_bins, uint8_t* codes_host) {
  return encode_host_impl(plan, wav_host, 4, n_rows, n_samples, row_stride, lengths_host, lo_host, scale_host, n_bins, codes_host);
}

int dmel_encode_host_pcm16_u8(dmel_plan* plan, const int16_t* wav_host, long long n_rows, long long n_samples,
                              long long row_stride, const int32_t* lengths_host, const float* lo_host,
                              const float* scale_host, int n_bins, uint8_t* codes_host) {
  return encode_host_impl(plan, wav_host, 2, n_rows, n_samples, row_stride, lengths_host, lo_host, scale_host, n_bins, codes_host);
}

static int check_tensor(const void* a, const void* b, long long n_rows, int n_mels, long long n_frames,
                        unsigned* n_elems) {
  if (!a || !b) return fail(DMEL_ERR_INVALID, "tensor pointer is null");
  if (n_rows < 0 || n_mels < 1 || n_frames < 1)
    return fail(DMEL_ERR_INVALID, "bad tensor shape (%lld, %d, %lld)", n_rows, n_mels, n_frames);
  const long long n = n_rows * (long long)n_mels * n_frames;
  if (n >= (1LL << 31)) return fail(DMEL_ERR_INVALID, "tensor of %lld elements exceeds the 2^31 flat-index limit; split the batch", n);
  *n_elems = (unsigned)n;
  return DMEL_OK;
}

int dmel_quantize_u8(const float* logmel_dev, long long n_rows, int n_mels, long long n_frames,
                     const float* lo_dev, const float* scale_dev, int n_bins, uint8_t* codes_dev,
                     void* stream) {
  return dmel_quantize_masked_u8(logmel_dev, n_rows, n_mels, n_frames, nullptr, lo_dev, scale_dev, n_bins, codes_dev, stream);
}

int dmel_quantize_masked_u8(const float* logmel_dev, long long n_rows, int n_mels, long long n_frames,
                            const int32_t* n_valid_dev, const float* lo_dev, const float* scale_dev, int n_bins,
                            uint8_t* codes_dev, void* stream) {
  unsigned n = 0;
  int rc = check_tensor(logmel_dev, codes_dev, n_rows, n_mels, n_frames, &n);
  if (rc != DMEL_OK) return rc;
  if ((rc = check_bins(n_bins)) != DMEL_OK) return rc;
  if (!lo_dev || !scale_dev) return fail(DMEL_ERR_INVALID, "lo_dev / scale_dev is null");
  if (n == 0) return DMEL_OK;
  const bool vec = ((reinterpret_cast<uintptr_t>(logmel_dev) & 15) == 0) && ((reinterpret_cast<uintptr_t>(codes_dev) & 3) == 0);
  const int dev = device_of(logmel_dev);
  DeviceGuard guard(dev);
  auto* kernel = n_valid_dev ? dmel::quantize_kernel<true> : dmel::quantize_kernel<false>;
  DMEL_CUDA(launch_pdl(kernel, dim3(stream_grid(dev, n >> 2)), dim3(dmel::kStreamThreads), 0, (cudaStream_t)stream,
                       logmel_dev, codes_dev, lo_dev, scale_dev, n, dmel::FastDiv::make((unsigned)n_frames),
                       dmel::FastDiv::make((unsigned)n_mels), (unsigned)n_bins, vec, n_valid_dev));
  return DMEL_OK;
}

int dmel_dequantize_f32(const uint8_t* codes_dev, long long n_rows, int n_mels, long long n_frames,
                        const float* table_dev, int n_bins, float* logmel_dev, void* stream) {
  unsigned n = 0;
  int rc = check_tensor(codes_dev, logmel_dev, n_rows, n_mels, n_frames, &n);
  if (rc != DMEL_OK) return rc;
  if ((rc = check_bins(n_bins)) != DMEL_OK) return rc;
  if (!table_dev) return fail(DMEL_ERR_INVALID, "table_dev is null");
  if (n == 0) return DMEL_OK;
  const bool vec = ((reinterpret_cast<uintptr_t>(logmel_dev) & 15) == 0) && ((reinterpret_cast<uintptr_t>(codes_dev) & 3) == 0);
  const int dev = device_of(codes_dev);
  DeviceGuard guard(dev);
  DMEL_CUDA(launch_pdl(dmel::dequantize_kernel, dim3(stream_grid(dev, n >> 2)), dim3(dmel::kStreamThreads), 0, (cudaStream_t)stream,
                       codes_dev, logmel_dev, table_dev, n, dmel::FastDiv::make((unsigned)n_frames),
                       dmel::FastDiv::make((unsigned)n_mels), (unsigned)n_bins, vec));
  return DMEL_OK;
}

int dmel_tensor_minmax_f32(const float* logmel_dev, long long n_rows, int n_mels, long long n_frames,
                           const int32_t* n_valid_dev, float* min_dev, float* max_dev, void* stream) {
  unsigned n = 0;
  int rc = check_tensor(logmel_dev, min_dev, n_rows, n_mels, n_frames, &n);
  if (rc != DMEL_OK) return rc;
  if (!max_dev) return fail(DMEL_ERR_INVALID, "max_dev is null");
  if (n == 0) return DMEL_OK;
  const unsigned lines = (unsigned)(n_rows * n_mels);
  const int dev = device_of(logmel_dev);
  DeviceGuard guard(dev);
  const int blocks = (int)std::min<unsigned>((lines + 7) / 8, (unsigned)sm_count_of(dev) * 8u);
  DMEL_CUDA(launch_pdl(dmel::tensor_minmax_kernel, dim3(blocks), dim3(dmel::kStreamThreads), 0, (cudaStream_t)stream,
                       logmel_dev, n_valid_dev, min_dev, max_dev, lines, (unsigned)n_frames, (unsigned)n_mels));
  return DMEL_OK;
}

int dmel_quantizer_derive_f32(const float* lo_dev, const float* hi_dev, int n_mels, int n_bins, float* scale_dev,
                              float* step_dev, int32_t* ready_dev, void* stream) {
  if (!lo_dev || !hi_dev) return fail(DMEL_ERR_INVALID, "lo_dev / hi_dev is null");
  if (n_mels < 1) return fail(DMEL_ERR_INVALID, "n_mels must be positive, got %d", n_mels);
  int rc = check_bins(n_bins);
  if (rc != DMEL_OK) return rc;
  if (!scale_dev && !step_dev && !ready_dev) return fail(DMEL_ERR_INVALID, "no output requested");
  DeviceGuard guard(device_of(lo_dev));
  DMEL_CUDA(launch_pdl(dmel::quantizer_derive_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, lo_dev, hi_dev, n_mels,
                       (float)n_bins, scale_dev, step_dev, reinterpret_cast<int*>(ready_dev)));
  return DMEL_OK;
}

static int fsq_levels(const int* levels, int n_levels, dmel::FsqLevels* lv) {
  if (!levels || n_levels < 1 || n_levels > dmel::kFsqMaxDims)
    return fail(DMEL_ERR_INVALID, "n_levels must be in [1, %d], got %d", dmel::kFsqMaxDims, n_levels);
  long long basis = 1;
  lv->n_dims = n_levels;
  for (int k = 0; k < n_levels; ++k) {
    if (levels[k] < 2 || levels[k] > 1024) return fail(DMEL_ERR_INVALID, "levels[%d] = %d outside [2, 1024]", k, levels[k]);
    const float l = (float)levels[k];
    lv->level[k] = levels[k];
    lv->half_l[k] = (l - 1.f) * (1.f + 1e-3f) / 2.f;          // FSQ.bound, eps = 1e-3
    lv->offset[k] = levels[k] % 2 == 0 ? 0.5f : 0.f;
    lv->shift[k] = std::atanh(lv->offset[k] / lv->half_l[k]);
    lv->half_width[k] = levels[k] / 2;
    lv->basis[k] = (int)basis;
    basis *= levels[k];
    if (basis > 0x7fffffffLL) return fail(DMEL_ERR_INVALID, "codebook of %lld entries exceeds 2^31", basis);
  }
  return DMEL_OK;
}

int dmel_fsq_encode(const float* zp_dev, long long n_rows, long long n_steps, int n_groups, const int* levels,
                    int n_levels, float* codes_dev, long long* indices_dev, long long* lm_ids_dev, int codebook_size,
                    void* stream) {
  dmel::FsqLevels lv;
  int rc = fsq_levels(levels, n_levels, &lv);
  if (rc != DMEL_OK) return rc;
  if (!zp_dev) return fail(DMEL_ERR_INVALID, "zp_dev is null");
  if (!codes_dev && !indices_dev && !lm_ids_dev) return fail(DMEL_ERR_INVALID, "no output requested");
  if (n_rows < 0 || n_rows > 65535 || n_steps < 1 || n_steps > (1LL << 30) || n_groups < 1 || n_groups > 512)
    return fail(DMEL_ERR_INVALID, "bad shape (%lld, %lld, %d)", n_rows, n_steps, n_groups);
  if (n_rows == 0) return DMEL_OK;
  DeviceGuard guard(device_of(zp_dev));
  const int tile_t = std::max(8, std::min(dmel::kFsqTileT, (40 * 1024) / (n_groups * (int)sizeof(int))));  // indices of a tile <= 40 KB
  const dim3 grid((unsigned)((n_steps + tile_t - 1) / tile_t), (unsigned)n_rows);
  cudaError_t e = cudaErrorInvalidValue;
#define DMEL_FSQ_ENCODE(D)                                                                                                     \
  case D:                                                                                                                      \
    e = launch_pdl(dmel::fsq_encode_kernel<D>, grid, dim3(dmel::kFsqThreads), (size_t)tile_t * n_groups * sizeof(int),         \
                   (cudaStream_t)stream, zp_dev, (int)n_steps, n_groups, lv, codes_dev, indices_dev, lm_ids_dev, codebook_size, \
                   tile_t);                                                                                                    \
    break;
  switch (n_levels) {
    DMEL_FSQ_ENCODE(1) DMEL_FSQ_ENCODE(2) DMEL_FSQ_ENCODE(3) DMEL_FSQ_ENCODE(4)
    DMEL_FSQ_ENCODE(5) DMEL_FSQ_ENCODE(6) DMEL_FSQ_ENCODE(7) DMEL_FSQ_ENCODE(8)
  }
#undef DMEL_FSQ_ENCODE
  DMEL_CUDA(e);
  return DMEL_OK;
}

int dmel_fsq_decode(const long long* indices_dev, long long n_rows, long long n_steps, int n_groups, const int* levels,
                    int n_levels, float* codes_dev, void* stream) {
  dmel::FsqLevels lv;
  int rc = fsq_levels(levels, n_levels, &lv);
  if (rc != DMEL_OK) return rc;
  if (!indices_dev || !codes_dev) return fail(DMEL_ERR_INVALID, "indices_dev / codes_dev is null");
  if (n_rows < 0 || n_rows > 65535 || n_steps < 1 || n_steps > (1LL << 30) || n_groups < 1 || n_groups > 512)
    return fail(DMEL_ERR_INVALID, "bad shape (%lld, %lld, %d)", n_rows, n_steps, n_groups);
  if (n_rows == 0) return DMEL_OK;
  DeviceGuard guard(device_of(indices_dev));
  const dim3 grid((unsigned)((n_steps + dmel::kFsqTileT - 1) / dmel::kFsqTileT), (unsigned)n_rows);
  cudaError_t e = cudaErrorInvalidValue;
#define DMEL_FSQ_DECODE(D)                                                                                                  \
  case D:                                                                                                                   \
    e = launch_pdl(dmel::fsq_decode_kernel<D>, grid, dim3(dmel::kFsqThreads), 0, (cudaStream_t)stream, indices_dev, (int)n_steps, \
                   n_groups, lv, codes_dev, dmel::kFsqTileT);                                                               \
    break;
  switch (n_levels) {
    DMEL_FSQ_DECODE(1) DMEL_FSQ_DECODE(2) DMEL_FSQ_DECODE(3) DMEL_FSQ_DECODE(4)
    DMEL_FSQ_DECODE(5) DMEL_FSQ_DECODE(6) DMEL_FSQ_DECODE(7) DMEL_FSQ_DECODE(8)
  }
#undef DMEL_FSQ_DECODE
  DMEL_CUDA(e);
  return DMEL_OK;
}

int dmel_antialias_snake_f32(const float* x_dev, long long n_rows, int n_channels, long long n_steps, const float* up_taps_host,
                             const float* down_taps_host, const float* log_alpha_dev, const float* log_beta_dev, float* y_dev,
                             void* stream) {
  if (!x_dev || !y_dev || !up_taps_host || !down_taps_host || !log_alpha_dev || !log_beta_dev)
    return fail(DMEL_ERR_INVALID, "null pointer argument");
  if (n_rows < 0 || n_rows > 65535 || n_channels < 1 || n_channels > 65535 || n_steps < 1 || n_steps > (1LL << 30))
    return fail(DMEL_ERR_INVALID, "bad shape (%lld, %d, %lld)", n_rows, n_channels, n_steps);
  if (n_rows == 0) return DMEL_OK;
  dmel::ActTaps taps;
  for (int k = 0; k < 12; ++k) {
    if (!std::isfinite(up_taps_host[k]) || !std::isfinite(down_taps_host[k])) return fail(DMEL_ERR_INVALID, "filter tap %d is not finite", k);
    taps.up2[k] = 2.0f * up_taps_host[k];  // the x2 of UpSample1d.forward (resample.py:31), exact
    taps.down[k] = down_taps_host[k];
  }
  const int dev = device_of(x_dev);
  DeviceGuard guard(dev);
  const long long tiles_per_row = (n_steps + dmel::kActWarpOut - 1) / dmel::kActWarpOut;
  const long long n_tiles = tiles_per_row * n_rows * n_channels;
  if (n_tiles >= (1LL << 31)) return fail(DMEL_ERR_INVALID, "(%lld, %d, %lld) is more than 2^31 warp tiles", n_rows, n_channels, n_steps);
  constexpr int kWarpsPerCta = dmel::kActThreads / 32;
  const long long ctas = std::min<long long>((n_tiles + kWarpsPerCta - 1) / kWarpsPerCta, (long long)sm_count_of(dev) * 8);  // persistent: 8 CTAs per SM
  DMEL_CUDA(launch_pdl(dmel::antialias_snake_kernel, dim3((unsigned)ctas), dim3(dmel::kActThreads), 0, (cudaStream_t)stream, x_dev, y_dev,
                       (int)n_steps, dmel::FastDiv::make((unsigned)tiles_per_row), dmel::FastDiv::make((unsigned)n_channels),
                       (unsigned)n_tiles, log_alpha_dev, log_beta_dev, taps));
  return DMEL_OK;
}

}  // extern "C"
