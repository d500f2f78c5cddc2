// One kernel variant of the fused transform: instantiates dmel_fused_kernel<NFFT, TF, MODE, OCC> for every
// output MODE the C ABI can ask for and exports them as a VariantOps table (fused_variants.h).
//   nvcc -DDMEL_V_NFFT=1024 -DDMEL_V_TF=8 -DDMEL_V_OCC=3 -c fused_variant.cu
#include <atomic>

#include "fused_variants.h"
#include "launch_util.cuh"
#include "logmel_kernel.cuh"

#if !defined(DMEL_V_NFFT) || !defined(DMEL_V_TF) || !defined(DMEL_V_OCC)
#error "compile with -DDMEL_V_NFFT=<1024|2048> -DDMEL_V_TF=<8|16> -DDMEL_V_OCC=<1|2|3>"
#endif

namespace dmel {
namespace {

constexpr int NFFT = DMEL_V_NFFT, TF = DMEL_V_TF, OCC = DMEL_V_OCC;
using LY = FusedLayout<NFFT, TF, OCC>;
constexpr bool kLeanVariant = OCC == 3 || LY::kSplit2048;
constexpr int kMaxDevices = 64;

size_t smem_need(int wave_len, int n_chan, int nnz, int n_order) { return LY::total(wave_len, n_chan, nnz, n_order); }

void fill_offsets(FusedParams* p) {
  p->off_mags = (int)LY::mags_off();
  p->off_wave = (int)LY::wave_off();
  p->off_window = (int)LY::window_off(p->wave_len);
  p->off_fold = (int)LY::fold_off(p->wave_len);
  p->off_weights = (int)LY::weights_off(p->wave_len);
  p->off_rec = (int)LY::rec_off(p->wave_len, p->nnz);
  p->off_order = (int)LY::order_off(p->wave_len, p->n_chan_pad, p->nnz);
  p->off_bars = (int)LY::bar_off(p->wave_len, p->n_chan_pad, p->nnz, p->n_order);
}

template <int MODE>
cudaError_t launch_mode(const FusedParams& p, int grid, size_t smem_bytes, cudaStream_t st) {
  if constexpr ((MODE & kInPcm16) != 0 && !kLeanVariant) {
    return cudaErrorNotSupported;  // int16 input is built for the register-lean variants only
  } else {
    auto kern = dmel_fused_kernel<NFFT, TF, MODE, OCC>;
    // The dynamic shared-memory limit is an attribute of the FUNCTION on a device, not of a plan: two plans of
    // different geometry share an instantiation, so raise it once per (instantiation, device) to the device
    // maximum instead of to one plan's size.
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
    return launch_pdl(kern, dim3(grid), dim3(kThreads), smem_bytes, st, p);
  }
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

#define DMEL_CAT_(a, b, c, d) a##b##_##c##_##d
#define DMEL_CAT(a, b, c, d) DMEL_CAT_(a, b, c, d)
extern const VariantOps DMEL_CAT(kVariant_, DMEL_V_NFFT, DMEL_V_TF, DMEL_V_OCC) = {NFFT, TF, OCC, kLeanVariant, kWarps, smem_need,
                                                                                  fill_offsets, launch};

}  // namespace dmel
