// Library-level C-ABI entry points and the host helpers the kernel translation units share.
#include <stdarg.h>
#include <stdlib.h>
#include <ctype.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <string>
#include <tuple>
#include <vector>

#include "host_common.h"

namespace fx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool ensure_dyn_smem(const void* func, int bytes, const char* what) {
  static std::mutex mu;
  static std::vector<std::tuple<const void*, int, int>> done;  // (function, device, bytes granted)
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (auto& e : done) {
    if (std::get<0>(e) == func && std::get<1>(e) == dev) {
      if (std::get<2>(e) >= bytes) return true;
      cudaError_t err = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (err != cudaSuccess) {
        set_error("%s: cudaFuncSetAttribute(%d B smem, device %d): %s", what, bytes, dev, cudaGetErrorString(err));
        return false;
      }
      std::get<2>(e) = bytes;
      return true;
    }
  }
  cudaError_t err = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%d B smem, device %d): %s", what, bytes, dev, cudaGetErrorString(err));
    return false;
  }
  done.emplace_back(func, dev, bytes);
  return true;
}

// name -> value table of the developer knobs; the environment variable FX_<NAME> seeds a knob at its first read.
static std::mutex g_tune_mu;
static std::vector<std::pair<std::string, int>> g_tune;

int tune_get(const char* name) {
  std::lock_guard<std::mutex> lock(g_tune_mu);
  for (auto& kv : g_tune)
    if (kv.first == name) return kv.second;
  std::string env = std::string("FX_") + name;
  for (auto& c : env) c = static_cast<char>(toupper(c));
  const char* v = getenv(env.c_str());
  const int val = (v && v[0]) ? atoi(v) : -1;
  g_tune.emplace_back(name, val);
  return val;
}

bool pdl_enabled() {
  // Programmatic dependent launch is ON by default (round 2: -2 % on the per-rank block chain at 8 GPUs, GPU suite green
  // with it; FX_PDL=0 / fx_tune("pdl", 0) turns it off). Read per launch so that fx_tune takes effect immediately.
  const int v = tune_get("pdl");
  return v != 0;
}

int num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || sym == nullptr) {
      set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

bool make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap(out, false, base, rank, dims, strides_bytes, box);
}

// Descriptor cache: a tensor map is a pure function of (type, base address, geometry, box), and a forward pass asks for the
// same few hundred of them every step (weights, persistent workspaces), so the driver call (~1 us, four per GEMM launch)
// is made once per distinct descriptor and thread. Entries never go stale — nothing in the key can change meaning — and
// the table is simply dropped when it grows past kTmapCacheMax (workspaces re-allocated at new addresses).
namespace {
struct TmapKey {
  uint64_t v[16];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) h = (h ^ x) * 1099511628211ull;
    return static_cast<size_t>(h ^ (h >> 29));
  }
};
constexpr size_t kTmapCacheMax = 16384;
}  // namespace

bool make_tmap(CUtensorMap* out, bool is_f32, const void* base, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{};
  const bool cacheable = rank >= 1 && rank <= 4;
  if (cacheable) {
    key.v[0] = reinterpret_cast<uint64_t>(base);
    key.v[1] = (static_cast<uint64_t>(rank) << 1) | (is_f32 ? 1u : 0u);
    for (int i = 0; i < rank; ++i) key.v[2 + i] = dims[i];
    for (int i = 0; i + 1 < rank; ++i) key.v[6 + i] = strides_bytes[i];
    for (int i = 0; i < rank; ++i) key.v[10 + i] = box[i];
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return true;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return false;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                  reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims[0..1]=%llu,%llu stride1=%llu box=%u,%u",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)strides_bytes[0], box[0], box[1]);
    return false;
  }
  if (cacheable) {
    if (cache.size() >= kTmapCacheMax) cache.clear();
    cache.emplace(key, *out);
  }
  return true;
}

}  // namespace fx

extern "C" int fx_tune(const char* name, int value) {
  if (name == nullptr || name[0] == 0) {
    fx::set_error("fx_tune: empty name");
    return FX_ERR_ARG;
  }
  std::lock_guard<std::mutex> lock(fx::g_tune_mu);
  for (auto& kv : fx::g_tune)
    if (kv.first == name) {
      kv.second = value;
      return FX_OK;
    }
  fx::g_tune.emplace_back(name, value);
  return FX_OK;
}

extern "C" int fx_abi_version(void) { return FX_ABI_VERSION; }
extern "C" const char* fx_last_error(void) { return fx::g_err; }

extern "C" int fx_check_device(int device) {
  int major = 0, minor = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (e != cudaSuccess) {
    fx::set_error("fx_check_device(%d): %s", device, cudaGetErrorString(e));
    return FX_ERR_CUDA;
  }
  if (major != 10) {
    fx::set_error("fx_check_device(%d): compute capability %d.%d, library is built for sm_100a only", device, major,
                  minor);
    return FX_ERR_ARCH;
  }
  return FX_OK;
}
