// The warp-specialised kernel (ws_kernel.cuh) as a VariantOps table: n_fft 1024, 8-frame tiles, two 16-warp CTAs per SM.
#include <atomic>

#include "fused_variants.h"
#include "launch_util.cuh"
#include "ws_kernel.cuh"

namespace dmel {
namespace {

using LY = WsLayout;
constexpr int kMaxDevices = 64;

size_t smem_need(int wave_len, int n_chan, int nnz, int n_order) { return LY::total(wave_len, n_chan, nnz, n_order); }

void fill_offsets(FusedParams* p) {
  p->off_mags = (int)LY::mags_off();
  p->off_wave = (int)LY::wave_off();
  p->off_window = (int)LY::window_off(p->wave_len);
  p->off_fold = 0;
  p->off_weights = (int)LY::weights_off(p->wave_len);
  p->off_rec = (int)LY::rec_off(p->wave_len, p->nnz);
  p->off_order = (int)LY::order_off(p->wave_len, p->n_chan_pad, p->nnz);
  p->off_bars = (int)LY::bar_off(p->wave_len, p->n_chan_pad, p->nnz, p->n_order);
}

template <int MODE>
cudaError_t launch_mode(const FusedParams& p, int grid, size_t smem_bytes, cudaStream_t st) {
  auto kern = dmel_ws_kernel<MODE>;
  static std::atomic<bool> raised[kMaxDevices];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices || !raised[dev].load(std::memory_order_acquire)) {
    int optin = 0;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices) raised[dev].store(true, std::memory_order_release);
  }
  return launch_pdl(kern, dim3(grid), dim3(kWsThreads), smem_bytes, st, p);
}

cudaError_t launch(int mode, const FusedParams& p, int grid, size_t smem_bytes, cudaStream_t st) {
  switch (mode) {
    case kOutCodes: return launch_mode<kOutCodes>(p, grid, smem_bytes, st);
    case kOutCodes | kInPcm16: return launch_mode<kOutCodes | kInPcm16>(p, grid, smem_bytes, st);
    case kOutCodes | kOutDequant: return launch_mode<kOutCodes | kOutDequant>(p, grid, smem_bytes, st);
    case kOutLogmel: return launch_mode<kOutLogmel>(p, grid, smem_bytes, st);
    case kOutLogmel | kOutBf16: return launch_mode<kOutLogmel | kOutBf16>(p, grid, smem_bytes, st);
    case kOutStats: return launch_mode<kOutStats>(p, grid, smem_bytes, st);
    case kOutLogmel | kOutStats: return launch_mode<kOutLogmel | kOutStats>(p, grid, smem_bytes, st);
    case kOutCodes | kOutLogmel: return launch_mode<kOutCodes | kOutLogmel>(p, grid, smem_bytes, st);
    case kOutCodes | kOutEdge: return launch_mode<kOutCodes | kOutEdge>(p, grid, smem_bytes, st);
    case kOutCodes | kOutLogmel | kOutEdge: return launch_mode<kOutCodes | kOutLogmel | kOutEdge>(p, grid, smem_bytes, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

extern const VariantOps kVariant_ws_1024 = {1024, kWsTF, 2, true, kWsMelWarps, smem_need, fill_offsets, launch};

}  // namespace dmel
