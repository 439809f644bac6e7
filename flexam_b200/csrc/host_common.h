// Host-side helpers shared by the C-ABI translation units: status codes, last-error slot, device info,
// and TMA tensor-map construction through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/flexam_b200.h"

namespace fx {

void set_error(const char* fmt, ...);
int num_sms();  // SM count of the current device (cached per device)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a function: set once per (function, device
// ordinal of the calling thread's current device), remembered under a mutex. Returns false (error text set) on failure.
bool ensure_dyn_smem(const void* func, int bytes, const char* what);

// Developer knobs (fx_tune / the FX_* environment variables read at first use): -1 = not set.
int tune_get(const char* name);

// bf16 tensor map with SWIZZLE_128B and a 64-element (128 B) innermost box.
// dims/strides are innermost-first; strides in BYTES for dims 1..rank-1.
bool make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);
// Same for bf16 (is_f32 = false) or fp32 (true) elements; the innermost box must span exactly 128 bytes.
bool make_tmap(CUtensorMap* out, bool is_f32, const void* base, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box);

// Programmatic dependent launch (FX_PDL=1, default off): the grid may be scheduled while its predecessor in the stream
// is still draining, so its prologue (barrier init, TMEM allocation, descriptor prefetch) and the launch latency overlap
// the predecessor's tail. Every kernel launched through launch_kernel() executes griddepcontrol.wait before its first
// global-memory access, which restores stream order for the data.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args&&... args) {
#ifdef __CUDACC__
  if (!pdl_enabled()) {
    kern<<<grid, block, smem, stream>>>(static_cast<KArgs>(args)...);
    return;
  }
#endif
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr = {};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define FX_CHECK_ARG(cond, ...)   \
  do {                            \
    if (!(cond)) {                \
      fx::set_error(__VA_ARGS__); \
      return FX_ERR_ARG;          \
    }                             \
  } while (0)

#define FX_CHECK_LAUNCH(name)                                                     \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      fx::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return FX_ERR_CUDA;                                                         \
    }                                                                             \
  } while (0)

}  // namespace fx
