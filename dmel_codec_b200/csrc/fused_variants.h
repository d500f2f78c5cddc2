// The kernel variants of the fused transform, one translation unit each (fused_variant.cu compiled with
// -DDMEL_V_NFFT / -DDMEL_V_TF / -DDMEL_V_OCC), so the library builds in parallel.  The host code sees a
// variant only through this table of plain functions.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace dmel {

struct FusedParams;

struct VariantOps {
  int n_fft, tf, occ;  // transform size, frames per tile, CTAs per SM the variant is built for
  bool lean;           // register-lean variant (the only ones with int16 PCM input)
  int mel_warps;       // warps that run the mel phase (the channel groups are dealt to this many)
  size_t (*smem_need)(int wave_len, int n_chan, int nnz, int n_order);
  void (*fill_offsets)(FusedParams* p);
  // mode = OR of the kOut* / kIn* bits of logmel_kernel.cuh; cudaErrorInvalidValue for a combination that is not built
  cudaError_t (*launch)(int mode, const FusedParams& p, int grid, size_t smem_bytes, cudaStream_t st);
};

// most CTAs per SM first: dmel_plan_create takes the first whose shared memory fits
extern const VariantOps kVariant_1024_8_3, kVariant_1024_16_2, kVariant_1024_8_2, kVariant_1024_16_1, kVariant_1024_8_1;
extern const VariantOps kVariant_2048_8_2, kVariant_2048_16_1, kVariant_2048_8_1;
extern const VariantOps kVariant_ws_1024;  // warp-specialised pipeline (ws_kernel.cuh): 16-warp CTAs, two per SM

}  // namespace dmel
