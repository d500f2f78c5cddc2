// Launch helper shared by every translation unit of the library.
#pragma once
#include <cstdlib>
#include <utility>
#include <cuda_runtime.h>

namespace dmel {

// Every kernel of this library is launched with the programmatic-stream-serialisation attribute: when the
// previous operation of the stream is one of our kernels (they all execute griddepcontrol.launch_dependents),
// the launch latency and the plan-constant prologue of this one overlap its tail; each kernel executes
// griddepcontrol.wait before it touches caller memory.  DMEL_NO_PDL=1 gives ordinary launches.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool pdl = std::getenv("DMEL_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace dmel
