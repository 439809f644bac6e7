// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma (UMMA 128 x BN x 16,
// accumulators double-buffered in TMEM) -> fused epilogues. One CTA per SM, static round-robin tile schedule
// with M-grouped rasterisation so the concurrently resident tiles share weight and activation panels in L2.
//
// Replaces the nn.Linear / conv-as-GEMM call sites of FlexAM/models/wan_transformer3d_FlexAM.py
// (:242-261, :363-370, :414-416, :506, :624-625, :675-678, :959-964); see include/flexam_b200.h.
#include "host_common.h"
#include "ptx.cuh"

namespace fx {

constexpr int kBM = 128;      // UMMA M (rows of A per tile) = TMEM lanes
constexpr int kBK = 64;       // K per smem stage: 64 bf16 = one 128-byte swizzle span
constexpr int kUmmaK = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr int kGemmThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int kEpiThreads = 128;
constexpr int kAccStride = 256;    // TMEM columns between the two accumulator buffers

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual alignment
};

struct GemmParams {
  int M, N, K;
  const __nv_bfloat16* bias;
  void* out;
  long long ldo;
  const float* gate_mod;
  const float* gate_e;
  long long gate_e_stride;
  const int* row_idx;
  int num_m_tiles, num_n_tiles, group_m;
};

__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& m_tile, int& n_tile) {
  const int per_group = p.group_m * p.num_n_tiles;
  const int g = tile / per_group;
  const int first_m = g * p.group_m;
  const int gsize = min(p.group_m, p.num_m_tiles - first_m);
  const int r = tile - g * per_group;
  m_tile = first_m + r % gsize;
  n_tile = r / gsize;
}

// One 32-column slab of one output row: v[j] = accumulator (fp32 bits) for column col0 + j.
template <int EPI>
__device__ __forceinline__ void epilogue_row32(const GemmParams& p, int row, int col0, uint32_t (&v)[32]) {
  // 8-column groups; N % 8 == 0 so a group is either fully valid or fully out of range.
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    if (col >= p.N) break;
    float y[8];
    if (p.bias != nullptr) {
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
      const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        y[2 * j] = __uint_as_float(v[g * 8 + 2 * j]) + bf16_lo(bw[j]);
        y[2 * j + 1] = __uint_as_float(v[g * 8 + 2 * j + 1]) + bf16_hi(bw[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(v[g * 8 + j]);
    }
    if constexpr (EPI == FX_EPI_BF16 || EPI == FX_EPI_GELU_BF16) {
      if constexpr (EPI == FX_EPI_GELU_BF16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = gelu_tanh(bf16_round(y[j]));
      }
      uint4 o;
      o.x = pack_bf16x2(y[0], y[1]);
      o.y = pack_bf16x2(y[2], y[3]);
      o.z = pack_bf16x2(y[4], y[5]);
      o.w = pack_bf16x2(y[6], y[7]);
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(row) * p.ldo + col;
      *reinterpret_cast<uint4*>(dst) = o;
    } else if constexpr (EPI == FX_EPI_F32) {
      float* dst = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col;
      float4 o0 = make_float4(bf16_round(y[0]), bf16_round(y[1]), bf16_round(y[2]), bf16_round(y[3]));
      float4 o1 = make_float4(bf16_round(y[4]), bf16_round(y[5]), bf16_round(y[6]), bf16_round(y[7]));
      *reinterpret_cast<float4*>(dst) = o0;
      *reinterpret_cast<float4*>(dst + 4) = o1;
    } else {  // FX_EPI_RESID_F32
      float gate[8];
      const bool has_gate = (p.gate_mod != nullptr) || (p.gate_e != nullptr);
#pragma unroll
      for (int j = 0; j < 8; ++j) gate[j] = has_gate ? 0.f : 1.f;
      if (p.gate_mod != nullptr) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gate_mod + col));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gate_mod + col + 4));
        gate[0] += g0.x; gate[1] += g0.y; gate[2] += g0.z; gate[3] += g0.w;
        gate[4] += g1.x; gate[5] += g1.y; gate[6] += g1.z; gate[7] += g1.w;
      }
      if (p.gate_e != nullptr) {
        const long long u = p.row_idx ? p.row_idx[row] : 0;
        const float* ge = p.gate_e + u * p.gate_e_stride + col;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(ge));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(ge + 4));
        gate[0] += g0.x; gate[1] += g0.y; gate[2] += g0.z; gate[3] += g0.w;
        gate[4] += g1.x; gate[5] += g1.y; gate[6] += g1.z; gate[7] += g1.w;
      }
      float* dst = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col;
      float4 x0 = *reinterpret_cast<const float4*>(dst);
      float4 x1 = *reinterpret_cast<const float4*>(dst + 4);
      x0.x += bf16_round(y[0]) * gate[0]; x0.y += bf16_round(y[1]) * gate[1];
      x0.z += bf16_round(y[2]) * gate[2]; x0.w += bf16_round(y[3]) * gate[3];
      x1.x += bf16_round(y[4]) * gate[4]; x1.y += bf16_round(y[5]) * gate[5];
      x1.z += bf16_round(y[6]) * gate[6]; x1.w += bf16_round(y[7]) * gate[7];
      *reinterpret_cast<float4*>(dst) = x0;
      *reinterpret_cast<float4*>(dst + 4) = x1;
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiThreads);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_tile, n_tile;
        tile_coords(p, tile, m_tile, n_tile);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBK, m_tile * kBM);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kBK, n_tile * BN);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t a_desc = umma_desc_sw128(a_addr + k * kUmmaK * 2, 16, 1024);
            const uint64_t b_desc = umma_desc_sw128(b_addr + k * kUmmaK * 2, 16, 1024);
            umma_ss(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: TMEM -> registers -> global =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_tile, n_tile;
      tile_coords(p, tile, m_tile, n_tile);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_tile * kBM + quad * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_wait_ld();
        if (row < p.M) epilogue_row32<EPI>(p, row, n_tile * BN + c * 32, v);
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;  // per (BN, EPI) instantiation; attribute is per-function, set once per process
  auto kern = gemm_bf16_kernel<BN, EPI>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("fx_gemm_bf16: cudaFuncSetAttribute(%d B smem): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return FX_ERR_CUDA;
    }
    configured = true;
  }
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  FX_CHECK_LAUNCH("fx_gemm_bf16");
  return FX_OK;
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                        cudaStream_t s) {
  switch (epi) {
    case FX_EPI_BF16: return launch_gemm<BN, FX_EPI_BF16>(ta, tb, p, s);
    case FX_EPI_GELU_BF16: return launch_gemm<BN, FX_EPI_GELU_BF16>(ta, tb, p, s);
    case FX_EPI_F32: return launch_gemm<BN, FX_EPI_F32>(ta, tb, p, s);
    case FX_EPI_RESID_F32: return launch_gemm<BN, FX_EPI_RESID_F32>(ta, tb, p, s);
  }
  set_error("fx_gemm_bf16: unknown epilogue %d", epi);
  return FX_ERR_ARG;
}

}  // namespace fx

extern "C" int fx_gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* out,
                            int64_t ldo, int M, int N, int K, int epilogue, const float* gate_mod,
                            const float* gate_e, int64_t gate_e_stride, const int32_t* row_idx, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(a && w && out, "fx_gemm_bf16: null pointer");
  FX_CHECK_ARG(M > 0 && N > 0 && K > 0, "fx_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  FX_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "fx_gemm_bf16: K (%d) and N (%d) must be multiples of 8", K, N);
  FX_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "fx_gemm_bf16: bad lda/ldw");
  FX_CHECK_ARG(ldo >= N && ldo % 8 == 0, "fx_gemm_bf16: bad ldo");
  FX_CHECK_ARG((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) |
                reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias) |
                reinterpret_cast<uintptr_t>(gate_mod) | reinterpret_cast<uintptr_t>(gate_e)) % 16 == 0,
               "fx_gemm_bf16: pointers must be 16-byte aligned");
  FX_CHECK_ARG(gate_e_stride % 4 == 0, "fx_gemm_bf16: gate_e_stride must be a multiple of 4");

  // tile width: least padded N among {256,192,128,64}; ties go to the wider tile
  const int cands[4] = {256, 192, 128, 64};
  int bn = 256, best = 1 << 30;
  for (int c : cands) {
    const int padded = (N + c - 1) / c * c;
    if (padded < best) {
      best = padded;
      bn = c;
    }
  }

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.out = out; p.ldo = ldo;
  p.gate_mod = gate_mod; p.gate_e = gate_e; p.gate_e_stride = gate_e_stride; p.row_idx = row_idx;
  p.num_m_tiles = (M + kBM - 1) / kBM;
  p.num_n_tiles = (N + bn - 1) / bn;
  p.group_m = 16;

  CUtensorMap ta, tb;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t box[2] = {kBK, kBM};
    if (!make_tmap_bf16(&ta, a, 2, dims, strides, box)) return FX_ERR_CUDA;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {kBK, static_cast<uint32_t>(bn)};
    if (!make_tmap_bf16(&tb, w, 2, dims, strides, box)) return FX_ERR_CUDA;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256: return dispatch_epi<256>(epilogue, ta, tb, p, s);
    case 192: return dispatch_epi<192>(epilogue, ta, tb, p, s);
    case 128: return dispatch_epi<128>(epilogue, ta, tb, p, s);
    default: return dispatch_epi<64>(epilogue, ta, tb, p, s);
  }
}
