// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma (UMMA 128 x BN x 16,
// accumulators double-buffered in TMEM) -> fused epilogues. One CTA per SM, static round-robin tile schedule
// with M-grouped rasterisation so the concurrently resident tiles share weight and activation panels in L2.
//
// Epilogues (include/flexam_b200.h): bf16 store, GELU-tanh + bf16 store, fp32 store (rounded / exact), and the gated fp32 residual
// `x += bf16(acc + bias) * gate`. The residual variant never reads x: each epilogue warp stages its 32 x 32 fp32
// slab in swizzled shared memory and issues a TMA reduction (`cp.reduce.async.bulk.tensor ... add.f32`), so the
// read-modify-write happens in L2 with fully coalesced traffic while the next slab is being computed.
//
// Replaces the nn.Linear / conv-as-GEMM call sites of FlexAM/models/wan_transformer3d_FlexAM.py
// (:242-261, :363-370, :414-416, :456, :461, :468, :506, :624-625, :675-678, :959-964).
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace fx {

constexpr int kBM = 128;      // UMMA M (rows of A per tile) = TMEM lanes
constexpr int kBK = 64;       // K per smem stage: 64 bf16 = one 128-byte swizzle span
constexpr int kUmmaK = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr int kGemmThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int kEpiThreads = 128;
constexpr int kAccStride = 256;    // TMEM columns between the two accumulator buffers
constexpr int kDefaultL2Hint = 2;  // "02": weights evict_last (ffn.2 1.405 -> 1.322 ms; evict_first on A costs 10-20 %); <a><b> eviction priorities of the CTA-pair kernel's operand loads (see fx_gemm_bf16)
constexpr int kSlabBytes = 32 * 32 * 4;  // one warp's 32 rows x 32 fp32 columns (128-byte rows, SWIZZLE_128B)

template <int BN, int EPI>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = (EPI == FX_EPI_RESID_F32) ? 4 * 2 * kSlabBytes : 0;  // 4 warps x 2 slabs
  static constexpr int kBudget = 226 * 1024 - 2048 - kStagingBytes;
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;  // +1024: alignment
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
  int n_span;  // n-tiles per outer slab of the rasterisation (num_n_tiles = one slab)
  int a_hint, b_hint;  // L2 eviction priority of the A / B panel loads (CTA-pair kernel): 0 normal, 1 first, 2 last
  // Implicit-GEMM convolution (fx_conv_gemm_bf16): A is a zero-padded channel-last activation [Tin * Hp * Wp, Cin]; the
  // K range is cut into conv_taps segments of conv_kb_per_tap k-blocks, segment `tap` reads A rows shifted by
  // conv_off[tap] (the tap's (dt, dy, dx) as a row distance on the padded grid; rows outside the buffer are TMA zero fill);
  // output row m is a position of the padded grid: interior positions are written to the dense [T*H*W, N] result.
  int conv_taps, conv_kb_per_tap, conv_Hp, conv_Wp, conv_H, conv_W;
  int conv_stride_s, conv_stride_t;   // 2: keep only odd interior positions / odd frames (stride-2 convolutions)
  int conv_off[27];
  int b_k_wrap;        // > 0: the weight's K extent; A is [M, planes * b_k_wrap] (bf16 planes of an fp32 matrix side by
                       // side) and the weight column of k-block kb is (kb * 64) % b_k_wrap — one accumulation over all planes
};

// Tile order: the N range is cut into slabs of n_span tile columns (outer loop); inside a slab, groups of group_m tile
// rows are swept column by column with the row index fastest. The tiles in flight therefore form a
// (group_m x in-flight columns) block: its A panels stay L2-resident across the slab's columns, and a slab's weight
// panels stay resident while all of M streams past when n_span is small.
__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& m_tile, int& n_tile) {
  const int per_slab = p.num_m_tiles * p.n_span;
  const int slab = tile / per_slab;
  const int n_first = slab * p.n_span;
  const int ncols = min(p.n_span, p.num_n_tiles - n_first);
  const int t = tile - slab * per_slab;
  const int per_group = p.group_m * ncols;
  const int g = t / per_group;
  const int first_m = g * p.group_m;
  const int gsize = min(p.group_m, p.num_m_tiles - first_m);
  const int r = t - g * per_group;
  m_tile = first_m + r % gsize;
  n_tile = n_first + r / gsize;
}

// y[j] = accumulator + bias for 8 consecutive columns starting at `col` (col < N guaranteed by the caller)
__device__ __forceinline__ void add_bias8(const GemmParams& p, int col, const uint32_t* v, float (&y)[8]) {
  if (p.bias != nullptr) {
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      y[2 * j] = __uint_as_float(v[2 * j]) + bf16_lo(bw[j]);
      y[2 * j + 1] = __uint_as_float(v[2 * j + 1]) + bf16_hi(bw[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(v[j]);
  }
}

// Convolution mode: padded-grid row -> dense output row, or -1 for a halo position.
__device__ __forceinline__ long long conv_dense_row(const GemmParams& p, int row) {
  if (p.conv_taps == 0) return row;
  const int plane = p.conv_Hp * p.conv_Wp;
  const int t = row / plane;
  const int r = row - t * plane;
  const int yp = r / p.conv_Wp;
  const int xp = r - yp * p.conv_Wp;
  const int y = yp - (p.conv_Hp - p.conv_H) / 2, x = xp - (p.conv_Wp - p.conv_W) / 2;
  if (y < 0 || y >= p.conv_H || x < 0 || x >= p.conv_W) return -1;
  if (p.conv_stride_s == 2 && !((y & 1) && (x & 1))) return -1;   // ZeroPad2d((0,1,0,1)) + stride 2: centres (2y+1, 2x+1)
  if (p.conv_stride_t == 2 && !(t & 1)) return -1;                 // time kernel 3, stride 2 over [last cached frame | chunk]
  const int ss = p.conv_stride_s, st = p.conv_stride_t;
  return (static_cast<long long>(t / st) * (p.conv_H / ss) + y / ss) * (p.conv_W / ss) + x / ss;
}

// Direct-store epilogues: one 32-column slab of one output row, v[j] = accumulator bits for column col0 + j.
template <int EPI>
__device__ __forceinline__ void epilogue_row32(const GemmParams& p, long long row, int col0, uint32_t (&v)[32]) {
  // 8-column groups; N % 8 == 0 so a group is either fully valid or fully out of range.
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    if (col >= p.N) break;
    float y[8];
    add_bias8(p, col, &v[g * 8], y);
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
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + col;
      *reinterpret_cast<uint4*>(dst) = o;
    } else if constexpr (EPI == FX_EPI_F32_EXACT) {  // verification mode: no rounding after the fp32 accumulator
      float* dst = reinterpret_cast<float*>(p.out) + row * p.ldo + col;
      *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
    } else {  // FX_EPI_F32
      float* dst = reinterpret_cast<float*>(p.out) + row * p.ldo + col;
      float4 o0 = make_float4(bf16_round(y[0]), bf16_round(y[1]), bf16_round(y[2]), bf16_round(y[3]));
      float4 o1 = make_float4(bf16_round(y[4]), bf16_round(y[5]), bf16_round(y[6]), bf16_round(y[7]));
      *reinterpret_cast<float4*>(dst) = o0;
      *reinterpret_cast<float4*>(dst + 4) = o1;
    }
  }
}

// Residual epilogue, staging half: writes bf16(acc + bias) * gate for this lane's row into the warp's swizzled slab
// (row r at r*128 B, 16-byte chunk c at position c ^ (r & 7) — the SWIZZLE_128B pattern the output tensor map expects).
__device__ __forceinline__ void resid_stage_row32(const GemmParams& p, bool row_valid, long long u, int col0,
                                                  const uint32_t (&v)[32], uint8_t* slab, int lane) {
  const bool has_gate = (p.gate_mod != nullptr) || (p.gate_e != nullptr);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    float o[8];
    if (row_valid && col < p.N) {
      float y[8];
      add_bias8(p, col, &v[g * 8], y);
      float gate[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) gate[j] = has_gate ? 0.f : 1.f;
      if (p.gate_mod != nullptr) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gate_mod + col));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gate_mod + col + 4));
        gate[0] += g0.x; gate[1] += g0.y; gate[2] += g0.z; gate[3] += g0.w;
        gate[4] += g1.x; gate[5] += g1.y; gate[6] += g1.z; gate[7] += g1.w;
      }
      if (p.gate_e != nullptr) {
        const float* ge = p.gate_e + u * p.gate_e_stride + col;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(ge));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(ge + 4));
        gate[0] += g0.x; gate[1] += g0.y; gate[2] += g0.z; gate[3] += g0.w;
        gate[4] += g1.x; gate[5] += g1.y; gate[6] += g1.z; gate[7] += g1.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = bf16_round(y[j]) * gate[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;  // clipped by the TMA store anyway
    }
    uint8_t* rowp = slab + lane * 128;
    *reinterpret_cast<float4*>(rowp + (((2 * g) ^ (lane & 7)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(rowp + (((2 * g + 1) ^ (lane & 7)) << 4)) = make_float4(o[4], o[5], o[6], o[7]);
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
  using Cfg = GemmCfg<BN, EPI>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * Cfg::kStageBytes;  // 1024-aligned: stage sizes are multiples of 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
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
    if constexpr (EPI == FX_EPI_RESID_F32) tma_prefetch_desc(&tmap_out);
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
  pdl_launch_dependents();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer (whole warp polls, one elected lane issues) =====================
    const bool leader = elect_one_sync();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_tile, n_tile;
      tile_coords(p, tile, m_tile, n_tile);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int kcol_b = p.b_k_wrap > 0 ? (kb * kBK) % p.b_k_wrap : kb * kBK;
          int a_col = kb * kBK, a_row = m_tile * kBM;
          if (p.conv_taps > 0) {
            const int tap = kb / p.conv_kb_per_tap;
            a_col = (kb - tap * p.conv_kb_per_tap) * kBK;
            a_row += p.conv_off[tap];
          }
          tma_load_2d(sa, &tmap_a, &full_bar[stage], a_col, a_row);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kcol_b, n_tile * BN);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp polls, one elected lane issues) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, false, false);
    const bool leader = elect_one_sync();
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
        if (leader) {
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t a_desc = umma_desc_sw128(a_addr + k * kUmmaK * 2, 16, 1024);
            const uint64_t b_desc = umma_desc_sw128(b_addr + k * kUmmaK * 2, 16, 1024);
            umma_ss(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: TMEM -> registers -> global =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    uint8_t* my_slabs = staging + (warp - 2) * 2 * kSlabBytes;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab_sel = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_tile, n_tile;
      tile_coords(p, tile, m_tile, n_tile);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row0 = m_tile * kBM + quad * 32;
      const int row = row0 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride;
      long long u = 0;
      if constexpr (EPI == FX_EPI_RESID_F32) {
        if (p.row_idx != nullptr && row < p.M) u = p.row_idx[row];
      }
      const long long orow = row < p.M ? conv_dense_row(p, row) : -1;   // where this thread's row goes (-1: nowhere)
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_wait_ld();
        const int col0 = n_tile * BN + c * 32;
        if constexpr (EPI == FX_EPI_RESID_F32) {
          if (col0 < p.N) {  // warp-uniform
            uint8_t* slab = my_slabs + (slab_sel & 1) * kSlabBytes;
            if (lane == 0) tma_wait_group_read<1>();  // the reduction issued from this slab two slabs ago has read it
            __syncwarp();
            resid_stage_row32(p, row < p.M, u, col0, v, slab, lane);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_2d(&tmap_out, slab, col0, row0);
              tma_commit_group();
            }
            ++slab_sel;
          }
        } else {
          if (orow >= 0) epilogue_row32<EPI>(p, orow, col0, v);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if constexpr (EPI == FX_EPI_RESID_F32) {
      if (lane == 0) tma_wait_group<0>();
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile with UMMA
// M = 256. Each CTA stages its own 128 rows of A and its own 128-row half of the B panel (32 KB per stage
// instead of 48 KB), so smem fill + operand-read traffic per SM drops by a third and the ring gets 6-7 stages.
// Both CTAs run a TMA producer and the epilogue; only the leader (cluster rank 0) issues MMAs. TMA completions
// from both CTAs are credited to the LEADER's full barrier; tcgen05.commit multicasts the "slot free" and
// "accumulator ready" arrivals to both CTAs; the peer's epilogue releases the accumulator on the leader's barrier.
// ---------------------------------------------------------------------------------------------------------
template <int EPI>
struct Gemm2Cfg {
  static constexpr int kBN = 256;
  static constexpr int kABytes = kBM * kBK * 2;        // this CTA's 128 rows of A
  static constexpr int kBBytes = (kBN / 2) * kBK * 2;  // this CTA's half of the B panel
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = (EPI == FX_EPI_RESID_F32) ? 4 * 2 * kSlabBytes : 0;
  static constexpr int kBudget = 226 * 1024 - 2048 - kStagingBytes;
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
  using Cfg = Gemm2Cfg<EPI>;
  constexpr int kStages = Cfg::kStages;
  constexpr int BN = Cfg::kBN;
  extern __shared__ uint8_t smem_raw[];
  // identical carve-up in both CTAs: barrier offsets must match for multicast commits / peer arrivals
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;  // 256 x 256 tiles
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if constexpr (EPI == FX_EPI_RESID_F32) tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);   // used in the leader only: its own expect_tx arrival + bytes from both CTAs
      mbar_init(&empty_bar[s], 1);  // one multicast commit per use
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);  // leader only: 4 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // peer barriers initialised and TMEM allocated before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const bool elected = elect_one_sync();
    const uint64_t pol_a = l2_policy(p.a_hint), pol_b = l2_policy(p.b_hint);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      int m_tile, n_tile;
      tile_coords(p, tile, m_tile, n_tile);
      const int a_row = m_tile * 256 + static_cast<int>(cta_rank) * kBM;
      const int b_row = n_tile * BN + static_cast<int>(cta_rank) * (BN / 2);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elected) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          const int kcol_b = p.b_k_wrap > 0 ? (kb * kBK) % p.b_k_wrap : kb * kBK;
          int a_col = kb * kBK, a_row_k = a_row;
          if (p.conv_taps > 0) {
            const int tap = kb / p.conv_kb_per_tap;
            a_col = (kb - tap * p.conv_kb_per_tap) * kBK;
            a_row_k += p.conv_off[tap];
          }
          tma_load_2d_2sm_hint(sa, &tmap_a, leader_full, a_col, a_row_k, pol_a);
          tma_load_2d_2sm_hint(sb, &tmap_b, leader_full, kcol_b, b_row, pol_b);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one elected lane) =====================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, false, false);
      const bool elected = elect_one_sync();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elected) {
            const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
            const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t a_desc = umma_desc_sw128(a_addr + k * kUmmaK * 2, 16, 1024);
              const uint64_t b_desc = umma_desc_sw128(b_addr + k * kUmmaK * 2, 16, 1024);
              umma_ss_2sm(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0);
            }
            umma_commit_2sm(&empty_bar[stage], 3);  // frees this slot in both CTAs
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elected) umma_commit_2sm(&tfull_bar[acc], 3);  // accumulator halves complete in both CTAs
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (both CTAs): own 128 rows of the 256-row tile =====================
    const int quad = warp & 3;
    uint8_t* my_slabs = staging + (warp - 2) * 2 * kSlabBytes;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab_sel = 0;
    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      int m_tile, n_tile;
      tile_coords(p, tile, m_tile, n_tile);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row0 = m_tile * 256 + static_cast<int>(cta_rank) * kBM + quad * 32;
      const int row = row0 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride;
      long long u = 0;
      if constexpr (EPI == FX_EPI_RESID_F32) {
        if (p.row_idx != nullptr && row < p.M) u = p.row_idx[row];
      }
      const long long orow = row < p.M ? conv_dense_row(p, row) : -1;   // where this thread's row goes (-1: nowhere)
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_wait_ld();
        const int col0 = n_tile * BN + c * 32;
        if constexpr (EPI == FX_EPI_RESID_F32) {
          uint8_t* slab = my_slabs + (slab_sel & 1) * kSlabBytes;
          if (lane == 0) tma_wait_group_read<1>();
          __syncwarp();
          resid_stage_row32(p, row < p.M, u, col0, v, slab, lane);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmap_out, slab, col0, row0);
            tma_commit_group();
          }
          ++slab_sel;
        } else {
          if (orow >= 0) epilogue_row32<EPI>(p, orow, col0, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if constexpr (EPI == FX_EPI_RESID_F32) {
      if (lane == 0) tma_wait_group<0>();
      __syncwarp();
    }
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still signal it or read its smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

template <int EPI>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const GemmParams& p,
                        cudaStream_t stream) {
  using Cfg = Gemm2Cfg<EPI>;
  static_assert(Cfg::kStages >= 4, "pipeline too shallow");
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
  auto kern = gemm2_bf16_kernel<EPI>;
  if (!ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::kSmemBytes, "fx_gemm_bf16(2cta)")) return FX_ERR_CUDA;
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (total < pairs ? total : pairs);
  launch_kernel(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, tb, tout, p);
  FX_CHECK_LAUNCH("fx_gemm_bf16(2cta)");
  return FX_OK;
}

static int dispatch_epi2(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout,
                         const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case FX_EPI_BF16: return launch_gemm2<FX_EPI_BF16>(ta, tb, tout, p, s);
    case FX_EPI_GELU_BF16: return launch_gemm2<FX_EPI_GELU_BF16>(ta, tb, tout, p, s);
    case FX_EPI_F32: return launch_gemm2<FX_EPI_F32>(ta, tb, tout, p, s);
    case FX_EPI_RESID_F32: return launch_gemm2<FX_EPI_RESID_F32>(ta, tb, tout, p, s);
    case FX_EPI_F32_EXACT: return launch_gemm2<FX_EPI_F32_EXACT>(ta, tb, tout, p, s);
  }
  set_error("fx_gemm_bf16: unknown epilogue %d", epi);
  return FX_ERR_ARG;
}

// CTA-pair kernel: only for full-width panels; FX_GEMM_MODE=pair|single overrides the default (pair)
static bool use_pair_kernel(int M, int N, int bn) {
  static int mode = -1;
  if (mode < 0) {
    const char* env = getenv("FX_GEMM_MODE");
    mode = (env && env[0] == 's') ? 0 : 1;
  }
  return mode == 1 && bn == 256 && N % 256 == 0 && M >= 256;
}

// M-tiles (256 rows) per rasterisation group of the CTA-pair kernel. Within a group every n-tile reuses the group's A
// panels; across groups the whole weight matrix is swept again. Measured on B200 (profiles/summary_r1.md, DRAM bytes
// per launch at M = 23296): a K-heavy problem (ffn.2: A = 668 MB, 7 MB per panel) wants g = 1 so that A is read
// once; a wide one (ffn.0 / qkv: W = 88 / 57 MB does not stay in L2 next to the streams) wants g = 12-16 so that W is
// swept few times; with a small weight matrix the choice does not matter. FX_GEMM_GROUP_M overrides (experiments).
static int pair_group_m(int N, int K) {
  const int forced = tune_get("gemm_group_m");
  if (forced > 0) return forced;
  if (K >= 2 * N) return 1;
  if (2LL * N * K <= (32LL << 20)) return 4;
  return 12;
}

template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI>;
  static_assert(Cfg::kStages >= 3, "pipeline too shallow");
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
  auto kern = gemm_bf16_kernel<BN, EPI>;
  if (!ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::kSmemBytes, "fx_gemm_bf16")) return FX_ERR_CUDA;
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  launch_kernel(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, tb, tout, p);
  FX_CHECK_LAUNCH("fx_gemm_bf16");
  return FX_OK;
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout,
                        const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case FX_EPI_BF16: return launch_gemm<BN, FX_EPI_BF16>(ta, tb, tout, p, s);
    case FX_EPI_GELU_BF16: return launch_gemm<BN, FX_EPI_GELU_BF16>(ta, tb, tout, p, s);
    case FX_EPI_F32: return launch_gemm<BN, FX_EPI_F32>(ta, tb, tout, p, s);
    case FX_EPI_RESID_F32: return launch_gemm<BN, FX_EPI_RESID_F32>(ta, tb, tout, p, s);
    case FX_EPI_F32_EXACT: return launch_gemm<BN, FX_EPI_F32_EXACT>(ta, tb, tout, p, s);
  }
  set_error("fx_gemm_bf16: unknown epilogue %d", epi);
  return FX_ERR_ARG;
}

}  // namespace fx

namespace fx {
struct ConvSpec {   // implicit-GEMM convolution over a zero-padded channel-last activation (see fx_conv_gemm_bf16)
  int taps, cin, Hp, Wp, H, W, stride_s, stride_t;
  long long a_rows;   // rows of the padded activation buffer (Tin * Hp * Wp)
  int off[27];
};
// K = reduction length seen by the kernel (A's width). kw = the weight's K extent: == K normally; K = planes * kw for
// the plane-concatenated form used by fx_linear_f32_tc (GemmParams::b_k_wrap).
static int gemm_impl(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* out,
                     int64_t ldo, int M, int N, int K, int kw, int epilogue, const float* gate_mod,
                     const float* gate_e, int64_t gate_e_stride, const int32_t* row_idx, void* stream,
                     const ConvSpec* conv = nullptr);
}

extern "C" int fx_gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* out,
                            int64_t ldo, int M, int N, int K, int epilogue, const float* gate_mod,
                            const float* gate_e, int64_t gate_e_stride, const int32_t* row_idx, void* stream) {
  return fx::gemm_impl(a, lda, w, ldw, bias, out, ldo, M, N, K, K, epilogue, gate_mod, gate_e, gate_e_stride, row_idx,
                       stream);
}

int fx::gemm_impl(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* out,
                         int64_t ldo, int M, int N, int K, int kw, int epilogue, const float* gate_mod,
                         const float* gate_e, int64_t gate_e_stride, const int32_t* row_idx, void* stream,
                         const ConvSpec* conv) {
  using namespace fx;
  FX_CHECK_ARG(a && w && out, "fx_gemm_bf16: null pointer");
  FX_CHECK_ARG(M > 0 && N > 0 && K > 0, "fx_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  FX_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "fx_gemm_bf16: K (%d) and N (%d) must be multiples of 8", K, N);
  FX_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && lda >= (conv ? conv->cin : K) && ldw >= kw, "fx_gemm_bf16: bad lda/ldw");
  FX_CHECK_ARG(kw == K || (kw > 0 && kw % kBK == 0 && K % kw == 0), "fx_gemm_bf16: bad plane wrap %d for K=%d", kw, K);
  FX_CHECK_ARG(ldo >= N && ldo % 8 == 0, "fx_gemm_bf16: bad ldo");
  FX_CHECK_ARG((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) |
                reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias) |
                reinterpret_cast<uintptr_t>(gate_mod) | reinterpret_cast<uintptr_t>(gate_e)) % 16 == 0,
               "fx_gemm_bf16: pointers must be 16-byte aligned");
  FX_CHECK_ARG(gate_e_stride % 4 == 0, "fx_gemm_bf16: gate_e_stride must be a multiple of 4");

  // tile width: least padded N among {256,192,160,128,64}; ties go to the wider tile (a 128 x 64 MMA is bound by the
  // shared-memory reads of A, and every extra tile column re-reads the A panel: N = 320 runs as 2 x 160, not 5 x 64)
  const int cands[5] = {256, 192, 160, 128, 64};
  int bn = 256, best = 1 << 30;
  for (int c : cands) {
    const int padded = (N + c - 1) / c * c;
    if (padded < best) {
      best = padded;
      bn = c;
    }
  }

  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.out = out; p.ldo = ldo;
  p.gate_mod = gate_mod; p.gate_e = gate_e; p.gate_e_stride = gate_e_stride; p.row_idx = row_idx;
  p.num_m_tiles = (M + kBM - 1) / kBM;
  p.num_n_tiles = (N + bn - 1) / bn;
  p.group_m = 16;
  p.n_span = p.num_n_tiles;
  p.b_k_wrap = kw == K ? 0 : kw;
  if (conv != nullptr) {
    p.conv_taps = conv->taps;
    p.conv_kb_per_tap = conv->cin / kBK;
    p.conv_Hp = conv->Hp; p.conv_Wp = conv->Wp; p.conv_H = conv->H; p.conv_W = conv->W;
    p.conv_stride_s = conv->stride_s; p.conv_stride_t = conv->stride_t;
    for (int i = 0; i < conv->taps; ++i) p.conv_off[i] = conv->off[i];
  }

  CUtensorMap ta, tb, tout;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(conv ? conv->cin : K),
                              static_cast<uint64_t>(conv ? conv->a_rows : M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t box[2] = {kBK, kBM};
    if (!make_tmap_bf16(&ta, a, 2, dims, strides, box)) return FX_ERR_CUDA;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(kw), static_cast<uint64_t>(N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {kBK, static_cast<uint32_t>(bn)};
    if (!make_tmap_bf16(&tb, w, 2, dims, strides, box)) return FX_ERR_CUDA;
  }
  CUtensorMap tb128 = tb;
  if (use_pair_kernel(M, N, bn)) {  // each CTA of a pair stages a 128-row half of the 256-wide B panel
    const uint64_t dims[2] = {static_cast<uint64_t>(kw), static_cast<uint64_t>(N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {kBK, 128};
    if (!make_tmap_bf16(&tb128, w, 2, dims, strides, box)) return FX_ERR_CUDA;
  }
  if (epilogue == FX_EPI_RESID_F32) {  // fp32 [M, N] view of the residual stream, 32 x 32 boxes for the reduction
    const uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldo) * 4};
    const uint32_t box[2] = {32, 32};
    if (!make_tmap(&tout, true, out, 2, dims, strides, box)) return FX_ERR_CUDA;
  } else {
    tout = ta;  // unused by the other epilogues
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (use_pair_kernel(M, N, bn)) {
    p.num_m_tiles = (M + 255) / 256;
    p.group_m = pair_group_m(N, K);
    const int forced_span = tune_get("gemm_n_span");  // FX_GEMM_N_SPAN / fx_tune: slab width of the rasterisation
    if (forced_span > 0 && forced_span < p.num_n_tiles) p.n_span = forced_span;
    // L2 eviction priorities of the operand loads, FX_GEMM_L2HINT=<a><b> (digits 0 normal, 1 evict_first, 2 evict_last)
    static int hint = -1;
    if (hint < 0) {
      const char* env = getenv("FX_GEMM_L2HINT");
      hint = (env && env[0] >= '0' && env[0] <= '2' && env[1] >= '0' && env[1] <= '2')
                 ? (env[0] - '0') * 10 + (env[1] - '0') : kDefaultL2Hint;
    }
    p.a_hint = hint / 10;
    p.b_hint = hint % 10;
    return dispatch_epi2(epilogue, ta, tb128, tout, p, s);
  }
  switch (bn) {
    case 256: return dispatch_epi<256>(epilogue, ta, tb, tout, p, s);
    case 192: return dispatch_epi<192>(epilogue, ta, tb, tout, p, s);
    case 160: return dispatch_epi<160>(epilogue, ta, tb, tout, p, s);
    case 128: return dispatch_epi<128>(epilogue, ta, tb, tout, p, s);
    default: return dispatch_epi<64>(epilogue, ta, tb, tout, p, s);
  }
}

// ---------------------------------------------------------------------------------------------------------
// fp32 linear on the tensor cores for the per-token embedding MLPs (time_embedding / time_projection :630-632 when
// the timesteps are per token and mostly distinct — fg/bg edit masks, pipeline :686-690 — so de-duplication does not
// help; the reference runs 1.56 TFLOP of SGEMM per sample there, :928-944). The fp32 input (optionally through SiLU)
// is split exactly into `planes` bf16 planes (hi [+ mid [+ lo]]) written side by side as one [M, planes*K] matrix, and
// ONE tcgen05 GEMM accumulates all planes against the bf16 weight (k-blocks wrap around the weight's K). planes = 2
// keeps 16 significant bits of every input (relative error <= 2^-17), 3 is exact up to the accumulator.
// ---------------------------------------------------------------------------------------------------------
namespace fx {
__global__ void split_planes_kernel(const float* in, long long ldi, int M, int K, int planes, int act_in,
                                    __nv_bfloat16* out) {
  const long long total = static_cast<long long>(M) * K;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const long long m = i / K;
    const int k = static_cast<int>(i - m * K);
    float a = in[m * ldi + k];
    if (act_in == 1) a = a / (1.f + expf(-a));  // SiLU
    __nv_bfloat16* row = out + m * (static_cast<long long>(planes) * K) + k;
    const __nv_bfloat16 hi = __float2bfloat16_rn(a);
    row[0] = hi;
    if (planes > 1) {
      const float r1 = a - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
      row[K] = mid;
      if (planes > 2) row[2 * K] = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    }
  }
}
}  // namespace fx

extern "C" int fx_linear_f32_tc(const float* in, int64_t ldi, const void* w, int64_t ldw, const void* bias, float* out,
                                int64_t ldo, int M, int N, int K, int act_in, int planes, void* planes_ws,
                                void* stream) {
  using namespace fx;
  FX_CHECK_ARG(in && w && out && planes_ws, "fx_linear_f32_tc: null pointer");
  FX_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % kBK == 0 && N % 8 == 0, "fx_linear_f32_tc: bad shape M=%d N=%d K=%d", M, N,
               K);
  FX_CHECK_ARG(planes >= 1 && planes <= 3 && (act_in == 0 || act_in == 1), "fx_linear_f32_tc: planes %d / act_in %d",
               planes, act_in);
  const long long total = static_cast<long long>(M) * K;
  long long g = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  split_planes_kernel<<<static_cast<int>(g < cap ? g : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, ldi, M, K, planes, act_in, reinterpret_cast<__nv_bfloat16*>(planes_ws));
  FX_CHECK_LAUNCH("fx_linear_f32_tc(split)");
  return gemm_impl(planes_ws, static_cast<int64_t>(planes) * K, w, ldw, bias, out, ldo, M, N, planes * K, K,
                   FX_EPI_F32_EXACT, nullptr, nullptr, 0, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------------------
// Convolution as an implicit GEMM on the same tcgen05 kernels — no im2col buffer. The activation is channel-last and
// zero-padded: bf16 [Tin, Hp, Wp, Cin] with Hp = H + 2, Wp = W + 2 for a 3x3 spatial kernel (Hp = H, Wp = W for 1x1) and
// Tin = T + kt - 1 leading frames of causal history for a kt-tap time kernel. Output position m runs over the padded grid
// [T, Hp, Wp]; tap (dt, dy, dx) reads the SAME matrix shifted by dt*Hp*Wp + (dy-1)*Wp + (dx-1) rows, so the K loop
// walks taps x Cin with one TMA coordinate change per tap; halo positions are computed and dropped (1.5-9 % extra work),
// interior ones are written densely as [T*H*W, Cout]. Weight: bf16 [Cout, taps*Cin], K order (dt, dy, dx, cin).
// stride_s = 2: nn.ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2) (the VAE's downsample, wan_vae3_8.py:101-107): only the
// centres (2y+1, 2x+1) are kept -> [T*(H/2)*(W/2), Cout]; stride_t = 2: Conv3d((3,1,1), stride (2,1,1)) over [last cached
// frame | chunk] (:108-109, :150-153): only odd frames of the causal form are kept. The dropped positions are computed.
// Used by the control fuser's 3x3 convolutions (cnn_conv1..4, wan_transformer3d_FlexAM.py:680-711).
// ---------------------------------------------------------------------------------------------------------
extern "C" int fx_conv_gemm_bf16(const void* act, const void* w, const void* bias, void* out, int64_t ldo, int T,
                                 int H, int W, int Cin, int Cout, int kt, int ks, int stride_s, int stride_t,
                                 int epilogue, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(act && w && out, "fx_conv_gemm_bf16: null pointer");
  FX_CHECK_ARG(T > 0 && H > 0 && W > 0 && Cin > 0 && Cin % kBK == 0 && Cout > 0 && Cout % 8 == 0,
               "fx_conv_gemm_bf16: bad shape T=%d H=%d W=%d Cin=%d (multiple of 64) Cout=%d (multiple of 8)", T, H, W, Cin,
               Cout);
  FX_CHECK_ARG((kt == 1 || kt == 3) && (ks == 1 || ks == 3), "fx_conv_gemm_bf16: kernel %dx%dx%d unsupported", kt, ks, ks);
  FX_CHECK_ARG(epilogue == FX_EPI_BF16 || epilogue == FX_EPI_GELU_BF16 || epilogue == FX_EPI_F32 ||
                   epilogue == FX_EPI_F32_EXACT,
               "fx_conv_gemm_bf16: epilogue %d unsupported", epilogue);
  FX_CHECK_ARG((stride_s == 1 || (stride_s == 2 && ks == 3 && H % 2 == 0 && W % 2 == 0)) &&
                   (stride_t == 1 || (stride_t == 2 && kt == 3 && T % 2 == 0)),
               "fx_conv_gemm_bf16: stride (%d, %d) needs a 3-tap kernel along it and even extents", stride_s, stride_t);
  ConvSpec c{};
  c.stride_s = stride_s;
  c.stride_t = stride_t;
  c.taps = kt * ks * ks;
  c.cin = Cin;
  c.H = H; c.W = W;
  c.Hp = ks == 3 ? H + 2 : H;
  c.Wp = ks == 3 ? W + 2 : W;
  const long long plane = static_cast<long long>(c.Hp) * c.Wp;
  c.a_rows = (T + kt - 1) * plane;
  FX_CHECK_ARG(c.a_rows < (1LL << 31) && static_cast<long long>(T) * plane < (1LL << 31), "fx_conv_gemm_bf16: too many rows");
  int i = 0;
  for (int dt = 0; dt < kt; ++dt)
    for (int dy = 0; dy < ks; ++dy)
      for (int dx = 0; dx < ks; ++dx)
        c.off[i++] = static_cast<int>(dt * plane + (ks == 3 ? (dy - 1) * c.Wp + (dx - 1) : 0));
  return gemm_impl(act, Cin, w, static_cast<int64_t>(c.taps) * Cin, bias, out, ldo, static_cast<int>(T * plane), Cout,
                   c.taps * Cin, c.taps * Cin, epilogue, nullptr, nullptr, 0, nullptr, stream, &c);
}
