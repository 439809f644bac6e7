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

// bf16 tensor map with SWIZZLE_128B and a 64-element (128 B) innermost box.
// dims/strides are innermost-first; strides in BYTES for dims 1..rank-1.
bool make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);
// Same for bf16 (is_f32 = false) or fp32 (true) elements; the innermost box must span exactly 128 bytes.
bool make_tmap(CUtensorMap* out, bool is_f32, const void* base, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box);

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
