// Non-causal flash-attention forward for sm_100a, head_dim 128, bf16 in/out, fp32 softmax statistics.
//
// Two kernels with the same work split and arithmetic (bit-identical outputs) and different TMEM plans; FX_FMHA_PIPE /
// kDefaultPipe selects. fmha2_fwd_kernel (shared score buffer, P in its own columns; see its header below) is the
// default; fmha_fwd_kernel (P overwrites S in place) is kept as the reference pipeline:
//
// One CTA owns 256 query rows (two 128-row tiles) of one (batch, head) and streams 128-key K/V tiles:
//   warps 0-15  softmax        four warpgroups: (query tile w, column half h) = warp >> 2 -> (w = wg >> 1, h = wg & 1).
//                              A thread owns one row of S_w and 64 of its 128 keys; the two threads of a row agree on
//                              the running maximum through shared memory + a 256-thread named barrier per tile. Four
//                              softmax warps per scheduler (instead of two) is what hides the MUFU/TMEM latencies.
//   warp 16     TMA producer   Q once, then K_j / V_j into 2-stage rings (SWIZZLE_128B boxes of [128 rows][64 d])
//   warp 17     MMA issuer     S_w = Q_w K_j^T (SS, both K-major) and O_w += P_w V_j (A = P from TMEM, B = V MN-major)
// fmha_fwd_kernel's TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_w (bf16 pairs) overwrites
// S_w in place:
// keys 0-63 in columns [0,32) and keys 64-127 in columns [64,96) of S_w, i.e. every thread only overwrites scores it
// has already loaded itself. While one tile's softmax runs, the tensor core works on the other tile's QK^T / PV, so
// K/V smem traffic is shared by both tiles and the MMA pipe stays busy. The running maximum is only raised when it
// grows by more than 2^8 (lazy rescale): the O accumulator is then rescaled in TMEM by the softmax threads themselves,
// which is safe because s_full[w] (committed after QK_w(j)) also implies PV_w(j-1) has retired and PV_w(j) is not
// issued before p_full[w].
//
// Replaces attention()/flash_attention() (FlexAM/models/attention_utils.py:174-233) at its two call sites,
// wan_transformer3d_FlexAM.py:251-256 (self, Lk = all tokens) and :367 (cross, Lk = 512, unmasked).
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"
#include "softmax_math.cuh"

namespace fx {

constexpr int kFmhaThreads = 640;  // 5 warpgroups: 4 softmax (tile x column half), {TMA, MMA, 2 idle warps}
constexpr int kHalfBytes = 128 * 64 * 2;       // one [128 rows][64 d] swizzled sub-tile
constexpr int kTileBytes = 2 * kHalfBytes;     // [128 rows][128 d]
constexpr int kXchBytes = 2 * 2 * 2 * 128 * 4;  // [slot][tile][half][row] fp32: row maxima / row sums between halves
constexpr int kFmhaSmem =
    2 * kTileBytes /*Q*/ + 2 * kTileBytes /*K ring*/ + 2 * kTileBytes /*V ring*/ + kXchBytes + 256 + 1024;
constexpr float kRescaleThreshold = 8.0f;      // log2 units
constexpr int kDefaultPipe = 3;                // FX_FMHA_PIPE (see fmha_launch): shared score buffer, two issuer warps
constexpr int kDefaultToken = 0;               // FX_FMHA_TOKEN: no turn-taking between the tiles' exponential passes
constexpr int kDefaultPoly8 = 0;               // exponential pairs (of 8) on the FMA pipe; measured with pipeline 2:
                                               // 0 -> 2.587 ms, 2 -> 2.617, 3 -> 2.704, 4 -> 2.771 (profiles/summary_r1f.md)

// Developer tracing (tests/native/fmha_trace.cu builds this file with -DFX_FMHA_TRACE): CTA (0,0,0) records
// clock64() at pipeline events of its first 64 KV steps. Compiled out of the library.
#ifdef FX_FMHA_TRACE
__device__ long long fx_fmha_trace[24 * 64];
}  // namespace fx
// read-out for tests/native/fmha_trace.cu (same translation unit as the symbol: no -rdc, so setmaxnreg is honoured)
extern "C" int fx_fmha_trace_read(long long* dst) {
  return cudaMemcpyFromSymbol(dst, fx::fx_fmha_trace, sizeof(long long) * 24 * 64) == cudaSuccess ? 0 : -2;
}
namespace fx {
#define FX_TRACE(who, j)                                                                     \
  do {                                                                                       \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 64 && lane == 0)      \
      fx_fmha_trace[(who) * 64 + (j)] = clock64();                                           \
  } while (0)
#else
#define FX_TRACE(who, j)
#endif

struct FmhaParams {
  __nv_bfloat16* o;
  long long o_stride_b, o_stride_l;
  int Lq, Lk;
  float scale_log2;
  // Ulysses output exchange (rows_per_peer > 0): query row i belongs to rank i / rows_per_peer and is written to
  // o_peer[rank] + b*o_stride_b + (i % rows_per_peer)*o_stride_l + h*128 (peer memory over NVLink; the caller has
  // already offset every base by this rank's first head).
  int rows_per_peer;
  int token;  // softmax turn-taking between the two query tiles (see the softmax loop)
  int dual;   // pipeline 2 only: separate QK^T and PV issuer warps
  int warp_arrive;  // pipeline 2 only: one mbarrier arrival per softmax warp instead of one per thread
  __nv_bfloat16* o_peer[8];
};


// kPoly8: of every 8 (x0,x1) pairs of a score row, how many take the FMA-pipe exp2 instead of MUFU.EX2
template <int kPoly8>
__global__ void __launch_bounds__(kFmhaThreads, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                    // [2 tiles][2 halves][128][64]
  uint8_t* sK = sQ + 2 * kTileBytes;     // [2 stages][2 halves][128][64]
  uint8_t* sV = sK + 2 * kTileBytes;
  float* xch = reinterpret_cast<float*>(sV + 2 * kTileBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kTileBytes + kXchBytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* k_empty = bars + 3;     // 2
  uint64_t* v_full = bars + 5;      // 2
  uint64_t* v_empty = bars + 7;     // 2
  uint64_t* s_full = bars + 9;      // 2 (per query tile)
  uint64_t* p_full = bars + 11;     // 2 (per query tile)
  uint64_t* o_done = bars + 13;     // 2 (per query tile)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.Lk + 127) / 128;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 17 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&o_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // global memory (TMA loads, output stores) is touched only below

  // Register re-partitioning: the data-movement warpgroup (warps 16-19) gives its registers to the four softmax
  // warpgroups (the pool is what the launch allocated, 640 x 96: 512 x 104 + 128 x 56 <= 61440).
  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 16) {
      // ===================== TMA producer (whole warp polls, one elected lane issues) =====================
      const bool leader = elect_one_sync();
      if (leader) {
        mbar_expect_tx(q_full, 2 * kTileBytes);
        for (int t = 0; t < 2; ++t)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_4d(sQ + t * kTileBytes + hf * kHalfBytes, &tmap_q, q_full, hf * 64, head, q0 + t * 128, batch);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        if (leader) {
          mbar_expect_tx(&k_full[s], kTileBytes);
          for (int hf = 0; hf < 2; ++hf)
            tma_load_4d(sK + s * kTileBytes + hf * kHalfBytes, &tmap_k, &k_full[s], hf * 64, head, j * 128, batch);
        }
        mbar_wait(&v_empty[s], ph ^ 1);
        if (leader) {
          mbar_expect_tx(&v_full[s], kTileBytes);
          for (int hf = 0; hf < 2; ++hf)
            tma_load_4d(sV + s * kTileBytes + hf * kHalfBytes, &tmap_v, &v_full[s], hf * 64, head, j * 128, batch);
        }
      }
      __syncwarp();
    } else if (warp == 17) {
      // ===================== MMA issuer (whole warp polls, one elected lane issues) =====================
      // The issue path is synchronous: the thread gets about one UMMA ahead of the tensor pipe, so whatever stands
      // between two instruction groups is pipe idle time (tests/native/umma_rate.cu: a barrier probe between groups
      // costs 120-190 idle cycles). Hence: elect.sync (straight-line UTCHMMA, no per-instruction elect loop), a
      // one-probe fast path in mbar_wait, and strict alternation between the two query tiles, which keeps their
      // softmax phases staggered (two independent issuers lock into the same phase and serialise softmax
      // against MMA).
      const bool leader = elect_one_sync();
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, false, true);  // B = V is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t k_addr = smem_u32(sK);
      const uint32_t v_addr = smem_u32(sV);

      auto issue_qk = [&](int w, int kstage) {
#if defined(FX_FMHA_EXPERIMENT) && FX_FMHA_EXPERIMENT >= 3
        return;
#endif
        const uint32_t a0 = q_addr + w * kTileBytes;
        const uint32_t b0 = k_addr + kstage * kTileBytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * kHalfBytes + (k & 3) * 32;
          umma_ss(tmem_base + w * 128, umma_desc_sw128(a0 + off, 16, 1024), umma_desc_sw128(b0 + off, 16, 1024),
                  idesc_qk, k != 0);
        }
      };
      auto issue_pv = [&](int w, int vstage, bool accumulate) {
#if defined(FX_FMHA_EXPERIMENT) && (FX_FMHA_EXPERIMENT == 2 || FX_FMHA_EXPERIMENT == 4)
        return;
#endif
        const uint32_t b0 = v_addr + vstage * kTileBytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A: 16 keys = 8 packed columns of P_w; B: 16 key rows (2048 B) further down the MN-major V tile
          // (keys 0-63 -> P columns [0,32), keys 64-127 -> [64,96) of S_w)
          umma_ts(tmem_base + 256 + w * 128, tmem_base + w * 128 + (k >> 2) * 64 + (k & 3) * 8,
                  umma_desc_sw128(b0 + k * 2048, kHalfBytes, 1024), idesc_pv, (accumulate || k != 0) ? 1u : 0u);
        }
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (leader) {
        issue_qk(0, 0);
        umma_commit(&s_full[0]);
        issue_qk(1, 0);
        umma_commit(&s_full[1]);
        umma_commit(&k_empty[0]);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int vs = j & 1;
        const int ks_next = (j + 1) & 1;
        const bool has_next = (j + 1) < n_kv;
        mbar_wait(&v_full[vs], (j >> 1) & 1);
        FX_TRACE(0, j);
        mbar_wait(&p_full[0], j & 1);
        FX_TRACE(1, j);
        tc_fence_after();
        if (leader) issue_pv(0, vs, j > 0);
        if (has_next) {
          mbar_wait(&k_full[ks_next], ((j + 1) >> 1) & 1);
          if (leader) {
            issue_qk(0, ks_next);
            umma_commit(&s_full[0]);
          }
        }
        FX_TRACE(2, j);
        mbar_wait(&p_full[1], j & 1);
        FX_TRACE(3, j);
        tc_fence_after();
        if (leader) {
          issue_pv(1, vs, j > 0);
          umma_commit(&v_empty[vs]);
          if (has_next) {
            issue_qk(1, ks_next);
            umma_commit(&s_full[1]);
            umma_commit(&k_empty[ks_next]);
          }
        }
        FX_TRACE(4, j);
      }
      if (leader) {
        umma_commit(&o_done[0]);
        umma_commit(&o_done[1]);
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===================== softmax / correction / output: one thread per (query row, 64-key half) =====================
    const int wg = warp >> 2;
    const int w = wg >> 1;           // query tile
    const int h = wg & 1;            // which 64 of the tile's 128 keys (and which 64 of the 128 output columns)
    const int quad = warp & 3;       // TMEM lane quadrant
    const int r = quad * 32 + lane;  // row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + w * 128 + h * 64;  // my scores; P goes to its first 32 columns
    const uint32_t o_tmem = tmem_base + lane_off + 256 + w * 128 + h * 64;
    const int row = q0 + w * 128 + r;
    const uint32_t pair_bar = 1 + w;  // named barrier of the tile's two warpgroups

    float m_used = 0.f;  // reference maximum (log2 domain) that the stored P / O / l are relative to
    float l = 0.f;       // partial row sum over my keys
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    for (int j = 0; j < n_kv; ++j) {
      if (warp == 0) FX_TRACE(6, j);
      mbar_wait(&s_full[w], j & 1);
      if (warp == 0) FX_TRACE(7, j);
      tc_fence_after();
      float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
#if defined(FX_FMHA_EXPERIMENT) && FX_FMHA_EXPERIMENT < 4
      if (p.Lk > 0) {  // MMA-only experiments: the softmax groups just hand the tile back
        l = 1.f;
        tc_fence_before();
        mbar_arrive(&p_full[w]);
        continue;
      }
#endif
      // Turn-taking (named barriers 3/4, 256 waiting + 256 arriving threads): the exponential pass of one tile has
      // the SM's MUFU and issue slots to itself while the tensor pipe works on the other tile, in the fixed order
      // tile 0 step j, tile 1 step j, tile 0 step j+1, ... Measured +1.3 % (profiles/r1_tools/fmha_sched_*.log).
      if (p.token && (w == 1 || j > 0)) asm volatile("bar.sync %0, 512;" ::"r"(3 + w) : "memory");
      uint32_t s[64];
      tmem_ld32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_wait_ld();
      if (warp == 0) FX_TRACE(11, j);
      const int valid = p.Lk - j * 128 - h * 64;  // keys of my half that exist
      if (valid < 64) {
        mask_chunk(&s[0], valid);
        mask_chunk(&s[32], valid - 32);
      }
      float mxa = -INFINITY, mxb = -INFINITY;
      max_chunk(&s[0], mxa, mxb);
      max_chunk(&s[32], mxa, mxb);
      // row maximum over both halves: exchange through shared memory (slots alternate so that a fast warp cannot
      // overwrite a value its partner has not read yet)
      float* slot = xch + ((j & 1) * 2 + w) * 256;
      const float mloc = fmaxf(mxa, mxb);
      if (warp == 0) FX_TRACE(12, j);
      slot[h * 128 + r] = mloc;
      asm volatile("bar.sync %0, 256;" ::"r"(pair_bar) : "memory");
      const float mx = fmaxf(mloc, slot[(h ^ 1) * 128 + r]) * p.scale_log2;
      if (warp == 0) FX_TRACE(10, j);

      if (j == 0) {
        m_used = mx;
      } else {
        const bool grow = mx > m_used + kRescaleThreshold;  // same decision in both threads of the row
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mx : m_used;
          const float f = ex2_approx(m_used - m_new);  // 1 for rows that keep their reference
          m_used = m_new;
          l *= f;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {  // my 64 columns of the output accumulator
            uint32_t o[16];
            tmem_ld16(o_tmem + c * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st16(o_tmem + c * 16, o);
          }
        }
      }
      const float2 nm2 = make_float2(-m_used, -m_used);
      exp_chunk<kPoly8>(&s[0], sc2, nm2, sum_a, sum_b, s_tmem);
      if (warp == 0) FX_TRACE(13, j);
      exp_chunk<kPoly8>(&s[32], sc2, nm2, sum_a, sum_b, s_tmem + 16);
      if (p.token && (w == 0 || j + 1 < n_kv)) asm volatile("bar.arrive %0, 512;" ::"r"(4 - w) : "memory");
      sum_a = add2(sum_a, sum_b);
      l += sum_a.x + sum_a.y;
      if (warp == 0) FX_TRACE(8, j);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[w]);
      if (warp == 0) FX_TRACE(9, j);
    }

    // epilogue: total row sum from both halves, then my 64 output columns: O / l -> bf16 -> global
    {
      float* slot = xch + ((n_kv & 1) * 2 + w) * 256;
      slot[h * 128 + r] = l;
      asm volatile("bar.sync %0, 256;" ::"r"(pair_bar) : "memory");
      l += slot[(h ^ 1) * 128 + r];
    }
    mbar_wait(&o_done[w], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* obase = p.o;
    int orow_idx = row;
    if (p.rows_per_peer > 0 && row < p.Lq) {
      const int owner = row / p.rows_per_peer;
      obase = p.o_peer[owner];
      orow_idx = row - owner * p.rows_per_peer;
    }
    __nv_bfloat16* orow = obase + static_cast<long long>(batch) * p.o_stride_b +
                          static_cast<long long>(orow_idx) * p.o_stride_l + head * 128 + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(o_tmem + c * 32, o);
      tmem_wait_ld();
      if (row < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Pipeline 2: ONE score buffer shared by the two query tiles, P in its own TMEM columns.
//
// In the kernel above P_w overwrites S_w, so QK_w(j+1) cannot be issued before PV_w(j) has consumed P_w(j): per tile the
// chain softmax (1860 clk) -> PV + QK issue (1160) -> latencies (320) = 3340 clk bounds a KV step that holds 2048 clk
// of tensor work (profiles/r1_tools/fmha_sched_experiments_r1.log). Here a score tile lives in TMEM only until its
// softmax threads have copied it into registers (~250 clk), so the next QK^T is issued while the exponentials of the
// previous one are still running, and softmax_w(j+1) finds its scores waiting when softmax_w(j) ends:
//   TMEM  S [0,128)   P0 [128,192)   P1 [192,256)   O0 [256,384)   O1 [384,512)
//   S uses, in order: QK_0(0) QK_1(0) QK_0(1) QK_1(1) ...; each waits for s_free of the previous user
//   s_full[w]  QK_w(j) retired                 (MMA -> softmax_w)        s_free[w]  softmax_w(j) holds S in registers
//   p_full[w]  P_w(j) stored                   (softmax_w -> MMA)        p_free[w]  PV_w(j) retired: P_w and O_w may be
//                                                                                  written again (next P, lazy rescale)
// p.dual selects two issuer warps (warp 17: QK^T, warp 18: PV) so that the QK^T waiting for the score buffer is not
// queued behind a PV group of the single issuer. Measured on B200 at the config-2 shape (profiles/summary_r1f.md):
// in-place pipeline 2.667 ms (1250 TFLOP/s), this kernel with one issuer 2.46 ms, with two issuers 2.35 ms (1419
// TFLOP/s); all three produce bit-identical outputs. What bounds it now: per tile the chain score pick-up (~350 clk)
// -> row maximum + exchange between the two halves (~400) -> exponentials (~1500, of which 1024 are the MUFU floor for
// one tile's 16 K exponentials) -> hand-over (~200), plus the score-buffer ring (QK^T -> visible -> tcgen05.ld ->
// s_free -> the other tile's QK^T, ~1600 clk per half). Timing experiments FX_FMHA_EXPERIMENT=5..8
// (tests/native/Makefile): no exponentials 2.37 ms, no exchange 2.35, TMEM traffic only 2.03, protocol only 1.94.
// ------------------------------------------------------------------------------------------------------------------
template <int kPoly8>
__global__ void __launch_bounds__(kFmhaThreads, 1)
fmha2_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                    // [2 tiles][2 halves][128][64]
  uint8_t* sK = sQ + 2 * kTileBytes;     // [2 stages][2 halves][128][64]
  uint8_t* sV = sK + 2 * kTileBytes;
  float* xch = reinterpret_cast<float*>(sV + 2 * kTileBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kTileBytes + kXchBytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* k_empty = bars + 3;     // 2
  uint64_t* v_full = bars + 5;      // 2
  uint64_t* v_empty = bars + 7;     // 2
  uint64_t* s_full = bars + 9;      // 2 (per query tile)
  uint64_t* s_free = bars + 11;     // 2
  uint64_t* p_full = bars + 13;     // 2
  uint64_t* p_free = bars + 15;     // 2
  uint64_t* o_done = bars + 17;     // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.Lk + 127) / 128;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 17 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], p.warp_arrive ? 8 : 256);
      mbar_init(&p_full[i], p.warp_arrive ? 8 : 256);
      mbar_init(&p_free[i], 1);
      mbar_init(&o_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // global memory (TMA loads, output stores) is touched only below
  constexpr uint32_t kColP = 128, kColO = 256;

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, false, true);  // B = V is MN-major
    const uint32_t q_addr = smem_u32(sQ);
    const uint32_t k_addr = smem_u32(sK);
    const uint32_t v_addr = smem_u32(sV);
    auto issue_qk = [&](int w, int kstage) {  // S = Q_w K^T into the shared score buffer
      const uint32_t a0 = q_addr + w * kTileBytes;
      const uint32_t b0 = k_addr + kstage * kTileBytes;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t off = (k >> 2) * kHalfBytes + (k & 3) * 32;
        umma_ss(tmem_base, umma_desc_sw128(a0 + off, 16, 1024), umma_desc_sw128(b0 + off, 16, 1024), idesc_qk, k != 0);
      }
    };
    auto issue_pv = [&](int w, int vstage, bool accumulate) {  // O_w += P_w V
      const uint32_t b0 = v_addr + vstage * kTileBytes;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // A: 16 keys = 8 packed columns of P_w (keys in order); B: 16 key rows (2048 B) further down the V tile
        umma_ts(tmem_base + kColO + w * 128, tmem_base + kColP + w * 64 + k * 8,
                umma_desc_sw128(b0 + k * 2048, kHalfBytes, 1024), idesc_pv, (accumulate || k != 0) ? 1u : 0u);
      }
    };

    if (warp == 16) {
      // ===================== TMA producer =====================
      const bool leader = elect_one_sync();
      if (leader) {
        mbar_expect_tx(q_full, 2 * kTileBytes);
        for (int t = 0; t < 2; ++t)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_4d(sQ + t * kTileBytes + hf * kHalfBytes, &tmap_q, q_full, hf * 64, head, q0 + t * 128, batch);
      }
      // K runs one step ahead of V (QK_0(j+2) is issued right after PV_1(j)): K_0 K_1 V_0 K_2 V_1 K_3 ...
      auto load_k = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&k_full[s], kTileBytes);
          for (int hf = 0; hf < 2; ++hf)
            tma_load_4d(sK + s * kTileBytes + hf * kHalfBytes, &tmap_k, &k_full[s], hf * 64, head, j * 128, batch);
        }
      };
      load_k(0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        if (j + 1 < n_kv) load_k(j + 1);
        mbar_wait(&v_empty[s], ((j >> 1) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&v_full[s], kTileBytes);
          for (int hf = 0; hf < 2; ++hf)
            tma_load_4d(sV + s * kTileBytes + hf * kHalfBytes, &tmap_v, &v_full[s], hf * 64, head, j * 128, batch);
        }
      }
      __syncwarp();
    } else if (warp == 17 && p.dual) {
      // ===================== QK^T issuer =====================
      const bool leader = elect_one_sync();
      mbar_wait(q_full, 0);
      for (int u = 0; u < 2 * n_kv; ++u) {
        const int j = u >> 1, w = u & 1;
        if (w == 0) mbar_wait(&k_full[j & 1], (j >> 1) & 1);
        if (u > 0) mbar_wait(&s_free[w ^ 1], ((u - 1) >> 1) & 1);  // the previous user has the scores in registers
        tc_fence_after();
        if (leader) {
          issue_qk(w, j & 1);
          umma_commit(&s_full[w]);
          if (w == 1) umma_commit(&k_empty[j & 1]);
        }
        if (w == 1) FX_TRACE(5, j);
      }
      __syncwarp();
    } else if (warp == 18 && p.dual) {
      // ===================== PV issuer =====================
      const bool leader = elect_one_sync();
      for (int j = 0; j < n_kv; ++j) {
        const int vs = j & 1;
        mbar_wait(&v_full[vs], (j >> 1) & 1);
        FX_TRACE(0, j);
        mbar_wait(&p_full[0], j & 1);
        FX_TRACE(1, j);
        tc_fence_after();
        if (leader) {
          issue_pv(0, vs, j > 0);
          umma_commit(&p_free[0]);
        }
        FX_TRACE(2, j);
        mbar_wait(&p_full[1], j & 1);
        FX_TRACE(3, j);
        tc_fence_after();
        if (leader) {
          issue_pv(1, vs, j > 0);
          umma_commit(&p_free[1]);
          umma_commit(&v_empty[vs]);
        }
        FX_TRACE(4, j);
      }
      if (leader) {
        umma_commit(&o_done[0]);
        umma_commit(&o_done[1]);
      }
      __syncwarp();
    } else if (warp == 17) {
      // ===================== single issuer: PV_0(j) QK_1(j+1) PV_1(j) QK_0(j+2) =====================
      const bool leader = elect_one_sync();
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (leader) {
        issue_qk(0, 0);
        umma_commit(&s_full[0]);
      }
      mbar_wait(&s_free[0], 0);
      tc_fence_after();
      if (leader) {
        issue_qk(1, 0);
        umma_commit(&s_full[1]);
        umma_commit(&k_empty[0]);
      }
      if (n_kv > 1) {
        mbar_wait(&k_full[1], 0);
        mbar_wait(&s_free[1], 0);
        tc_fence_after();
        if (leader) {
          issue_qk(0, 1);
          umma_commit(&s_full[0]);
        }
      }
      for (int j = 0; j < n_kv; ++j) {
        const int vs = j & 1;
        mbar_wait(&v_full[vs], (j >> 1) & 1);
        FX_TRACE(0, j);
        mbar_wait(&p_full[0], j & 1);
        FX_TRACE(1, j);
        tc_fence_after();
        if (leader) {
          issue_pv(0, vs, j > 0);
          umma_commit(&p_free[0]);
        }
        if (j + 1 < n_kv) {
          mbar_wait(&s_free[0], (j + 1) & 1);  // tile 0 has picked up S_0(j+1)
          tc_fence_after();
          if (leader) {
            issue_qk(1, (j + 1) & 1);
            umma_commit(&s_full[1]);
            umma_commit(&k_empty[(j + 1) & 1]);
          }
        }
        FX_TRACE(2, j);
        mbar_wait(&p_full[1], j & 1);
        FX_TRACE(3, j);
        tc_fence_after();
        if (leader) {
          issue_pv(1, vs, j > 0);
          umma_commit(&p_free[1]);
          umma_commit(&v_empty[vs]);
        }
        if (j + 2 < n_kv) {
          mbar_wait(&k_full[j & 1], ((j + 2) >> 1) & 1);
          mbar_wait(&s_free[1], (j + 1) & 1);  // tile 1 has picked up S_1(j+1)
          tc_fence_after();
          if (leader) {
            issue_qk(0, j & 1);
            umma_commit(&s_full[0]);
          }
        }
        FX_TRACE(4, j);
      }
      if (leader) {
        umma_commit(&o_done[0]);
        umma_commit(&o_done[1]);
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===================== softmax / correction / output: one thread per (query row, 64-key half) =====================
    const int wg = warp >> 2;
    const int w = wg >> 1;           // query tile
    const int h = wg & 1;            // which 64 of the tile's 128 keys (and which 64 of the 128 output columns)
    const int quad = warp & 3;       // TMEM lane quadrant
    const int r = quad * 32 + lane;  // row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + h * 64;                     // my 64 scores of the shared buffer
    const uint32_t p_tmem = tmem_base + lane_off + kColP + w * 64 + h * 32;    // my 32 packed columns of P_w
    const uint32_t o_tmem = tmem_base + lane_off + kColO + w * 128 + h * 64;
    const int row = q0 + w * 128 + r;
    const uint32_t pair_bar = 1 + w;  // named barrier of the tile's two warpgroups

    // every thread arrives (count 256), or one lane per warp after the warp has converged (count 8)
    auto softmax_arrive = [&](uint64_t* bar) {
      if (p.warp_arrive) {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
      } else {
        mbar_arrive(bar);
      }
    };
    float m_used = 0.f;  // reference maximum (log2 domain) that the stored P / O / l are relative to
    float l = 0.f;       // partial row sum over my keys
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    for (int j = 0; j < n_kv; ++j) {
      if (warp == 0) FX_TRACE(6, j);
      if (warp == 8) FX_TRACE(16, j);
      mbar_wait(&s_full[w], j & 1);
      if (warp == 0) FX_TRACE(7, j);
      if (warp == 8) FX_TRACE(17, j);
      tc_fence_after();
#if defined(FX_FMHA_EXPERIMENT) && (FX_FMHA_EXPERIMENT == 7 || FX_FMHA_EXPERIMENT == 8)
      {  // timing experiments: 7 = barrier protocol only; 8 = protocol + the TMEM traffic (ld S, st P), no arithmetic
        uint32_t t[32];
#if FX_FMHA_EXPERIMENT == 8
        tmem_ld32(s_tmem, t);
        tmem_ld32(s_tmem + 32, t);
        tmem_wait_ld();
#endif
        tc_fence_before();
        softmax_arrive(&s_free[w]);
        if (j > 0) {
          mbar_wait(&p_free[w], (j - 1) & 1);
          tc_fence_after();
        }
#if FX_FMHA_EXPERIMENT == 8
        tmem_st16(p_tmem, t);
        tmem_st16(p_tmem + 16, t + 16);
        tmem_wait_st();
#endif
        l = 1.f;
        tc_fence_before();
        softmax_arrive(&p_full[w]);
        continue;
      }
#endif
      uint32_t s[64];
      tmem_ld32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_wait_ld();
      tc_fence_before();
      softmax_arrive(&s_free[w]);  // the score buffer may be overwritten by the other tile's QK^T
      if (warp == 0) FX_TRACE(11, j);
      if (warp == 8) FX_TRACE(18, j);
      const int valid = p.Lk - j * 128 - h * 64;  // keys of my half that exist
      if (valid < 64) {
        mask_chunk(&s[0], valid);
        mask_chunk(&s[32], valid - 32);
      }
      float mxa = -INFINITY, mxb = -INFINITY;
      max_chunk(&s[0], mxa, mxb);
      max_chunk(&s[32], mxa, mxb);
      float* slot = xch + ((j & 1) * 2 + w) * 256;
      const float mloc = fmaxf(mxa, mxb);
      if (warp == 0) FX_TRACE(12, j);
#if defined(FX_FMHA_EXPERIMENT) && FX_FMHA_EXPERIMENT == 6
      const float mx = 40.f + 0.f * mloc;  // timing experiment: no exchange between the halves (results are wrong)
      (void)slot;
#else
      slot[h * 128 + r] = mloc;
      asm volatile("bar.sync %0, 256;" ::"r"(pair_bar) : "memory");
      const float mx = fmaxf(mloc, slot[(h ^ 1) * 128 + r]) * p.scale_log2;
#endif
      if (warp == 0) FX_TRACE(10, j);
      if (warp == 8) FX_TRACE(19, j);

      // PV_w(j-1) has retired: P_w may be overwritten and O_w rescaled (PV_w(j) is not issued before p_full[w] below)
      if (j > 0) {
        mbar_wait(&p_free[w], (j - 1) & 1);
        tc_fence_after();
      }
      if (warp == 0) FX_TRACE(15, j);
      // Turn-taking between the two tiles' exponential passes (named barriers 3/4, 256 waiting + 256 arriving
      // threads), order tile 0 step j, tile 1 step j, tile 0 step j+1, ...: one tile's pass has the MUFU and issue
      // slots to itself while the other tile loads and reduces its next scores.
      if (p.token && (w == 1 || j > 0)) asm volatile("bar.sync %0, 512;" ::"r"(3 + w) : "memory");
      if (warp == 0) FX_TRACE(14, j);
      if (warp == 8) FX_TRACE(20, j);
      if (j == 0) {
        m_used = mx;
      } else {
        const bool grow = mx > m_used + kRescaleThreshold;  // same decision in both threads of the row
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mx : m_used;
          const float f = ex2_approx(m_used - m_new);  // 1 for rows that keep their reference
          m_used = m_new;
          l *= f;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {  // my 64 columns of the output accumulator
            uint32_t o[16];
            tmem_ld16(o_tmem + c * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st16(o_tmem + c * 16, o);
          }
        }
      }
      const float2 nm2 = make_float2(-m_used, -m_used);
      float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
      exp_chunk<kPoly8>(&s[0], sc2, nm2, sum_a, sum_b, p_tmem);
      if (warp == 0) FX_TRACE(13, j);
      if (warp == 8) FX_TRACE(21, j);
      // token == 2: hand the turn over after the first half, so the other tile ramps up while this one drains
      if (p.token == 2 && (w == 0 || j + 1 < n_kv)) asm volatile("bar.arrive %0, 512;" ::"r"(4 - w) : "memory");
      exp_chunk<kPoly8>(&s[32], sc2, nm2, sum_a, sum_b, p_tmem + 16);
      if (p.token == 1 && (w == 0 || j + 1 < n_kv)) asm volatile("bar.arrive %0, 512;" ::"r"(4 - w) : "memory");
      sum_a = add2(sum_a, sum_b);
      l += sum_a.x + sum_a.y;
      if (warp == 0) FX_TRACE(8, j);
      if (warp == 8) FX_TRACE(22, j);
      tmem_wait_st();
      tc_fence_before();
      softmax_arrive(&p_full[w]);
      if (warp == 0) FX_TRACE(9, j);
      if (warp == 8) FX_TRACE(23, j);
    }

    // epilogue: total row sum from both halves, then my 64 output columns: O / l -> bf16 -> global
    {
      float* slot = xch + ((n_kv & 1) * 2 + w) * 256;
      slot[h * 128 + r] = l;
      asm volatile("bar.sync %0, 256;" ::"r"(pair_bar) : "memory");
      l += slot[(h ^ 1) * 128 + r];
    }
    mbar_wait(&o_done[w], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* obase = p.o;
    int orow_idx = row;
    if (p.rows_per_peer > 0 && row < p.Lq) {
      const int owner = row / p.rows_per_peer;
      obase = p.o_peer[owner];
      orow_idx = row - owner * p.rows_per_peer;
    }
    __nv_bfloat16* orow = obase + static_cast<long long>(batch) * p.o_stride_b +
                          static_cast<long long>(orow_idx) * p.o_stride_l + head * 128 + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(o_tmem + c * 32, o);
      tmem_wait_ld();
      if (row < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static bool make_qkv_tmap(CUtensorMap* m, const void* base, int64_t stride_b, int64_t stride_l, int B, int H, int L) {
  const uint64_t dims[4] = {128, static_cast<uint64_t>(H), static_cast<uint64_t>(L), static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {128 * 2, static_cast<uint64_t>(stride_l) * 2, static_cast<uint64_t>(stride_b) * 2};
  const uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap_bf16(m, base, 4, dims, strides, box);
}

}  // namespace fx

namespace fx {
static int fmha_launch(const void* q, int64_t q_stride_b, int64_t q_stride_l, const void* k, int64_t k_stride_b,
                       int64_t k_stride_l, const void* v, int64_t v_stride_b, int64_t v_stride_l, void* o,
                       void* const* o_peers, int n_peers, int rows_per_peer, int64_t o_stride_b, int64_t o_stride_l,
                       int B, int H, int Lq, int Lk, float scale, void* stream, const char* name) {
  FX_CHECK_ARG(q && k && v && (o || o_peers), "%s: null pointer", name);
  FX_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0, "%s: empty problem B=%d H=%d Lq=%d Lk=%d", name, B, H, Lq, Lk);
  FX_CHECK_ARG(H <= 65535 && B <= 65535, "%s: H and B must fit a grid dimension", name);
  const int64_t strides[8] = {q_stride_b, q_stride_l, k_stride_b, k_stride_l, v_stride_b, v_stride_l, o_stride_b,
                              o_stride_l};
  for (int64_t s : strides) FX_CHECK_ARG(s % 8 == 0 && s >= 0, "%s: strides must be multiples of 8 elements", name);
  FX_CHECK_ARG(q_stride_l >= 128LL * H && k_stride_l >= 128LL * H && v_stride_l >= 128LL * H && o_stride_l >= 128LL * H,
               "%s: row stride smaller than H*128", name);
  FX_CHECK_ARG((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                reinterpret_cast<uintptr_t>(o)) % 16 == 0,
               "%s: pointers must be 16-byte aligned", name);

  CUtensorMap tq, tk, tv;
  // a batch stride of 0 is not encodable; with B == 1 any non-zero value is equivalent
  auto bs = [&](int64_t sb, int64_t sl, int L) { return (B == 1 && sb == 0) ? sl * L : sb; };
  if (!make_qkv_tmap(&tq, q, bs(q_stride_b, q_stride_l, Lq), q_stride_l, B, H, Lq)) return FX_ERR_CUDA;
  if (!make_qkv_tmap(&tk, k, bs(k_stride_b, k_stride_l, Lk), k_stride_l, B, H, Lk)) return FX_ERR_CUDA;
  if (!make_qkv_tmap(&tv, v, bs(v_stride_b, v_stride_l, Lk), v_stride_l, B, H, Lk)) return FX_ERR_CUDA;

  // exp2 split between MUFU and the FMA pipe, in eighths of the pairs; FX_FMHA_POLY=0|2|3|4 overrides the default.
  // FX_FMHA_PIPE=1: P overwrites S (fmha_fwd_kernel); 2: shared score buffer, one issuer warp; 3: shared score buffer,
  // separate QK^T / PV issuer warps (fmha2_fwd_kernel). A fourth order (each PV split in two around the other tile's
  // QK^T) measured 5-10 % slower: it delays p_free (profiles/summary_r1f.md).
  static int poly = -1, token = kDefaultToken, pipe = kDefaultPipe, warp_arrive = 0;
  if (poly < 0) {
    const char* env = getenv("FX_FMHA_POLY");
    poly = kDefaultPoly8;
    if (env && (env[0] == '0' || env[0] == '2' || env[0] == '3' || env[0] == '4')) poly = env[0] - '0';
    const char* t = getenv("FX_FMHA_TOKEN");
    if (t && t[0] >= '0' && t[0] <= '2') token = t[0] - '0';  // 2: early hand-over (pipeline 2/3 only)
    const char* pp = getenv("FX_FMHA_PIPE");
    if (pp && pp[0] >= '1' && pp[0] <= '3') pipe = pp[0] - '0';
    const char* wa = getenv("FX_FMHA_WARP_ARRIVE");
    if (wa) warp_arrive = wa[0] != '0';
  }
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const FmhaParams);
  const Kern kerns[2][5] = {
      {fmha_fwd_kernel<0>, nullptr, fmha_fwd_kernel<2>, fmha_fwd_kernel<3>, fmha_fwd_kernel<4>},
      {fmha2_fwd_kernel<0>, nullptr, fmha2_fwd_kernel<2>, fmha2_fwd_kernel<3>, fmha2_fwd_kernel<4>}};
  const int fam = pipe >= 2 ? 1 : 0;
  const Kern kern = kerns[fam][poly];
  if (!ensure_dyn_smem(reinterpret_cast<const void*>(kern), kFmhaSmem, name)) return FX_ERR_CUDA;
  FmhaParams p{};
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.o_stride_b = o_stride_b;
  p.o_stride_l = o_stride_l;
  p.Lq = Lq;
  p.Lk = Lk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.rows_per_peer = 0;
  p.token = token;
  p.dual = pipe == 3;
  p.warp_arrive = warp_arrive;
  if (o_peers != nullptr) {
    FX_CHECK_ARG(n_peers >= 1 && n_peers <= 8 && rows_per_peer > 0 &&
                     static_cast<int64_t>(n_peers) * rows_per_peer >= Lq,
                 "%s: %d peers x %d rows do not cover Lq=%d", name, n_peers, rows_per_peer, Lq);
    p.rows_per_peer = rows_per_peer;
    for (int i = 0; i < n_peers; ++i) {
      FX_CHECK_ARG(o_peers[i] != nullptr && reinterpret_cast<uintptr_t>(o_peers[i]) % 16 == 0,
                   "%s: peer buffer %d null or misaligned", name, i);
      p.o_peer[i] = reinterpret_cast<__nv_bfloat16*>(o_peers[i]);
    }
    p.o = p.o_peer[0];
  }
  dim3 grid((Lq + 255) / 256, H, B);
  launch_kernel(kern, grid, dim3(kFmhaThreads), kFmhaSmem, reinterpret_cast<cudaStream_t>(stream), tq, tk, tv, p);
  FX_CHECK_LAUNCH(name);
  return FX_OK;
}
}  // namespace fx

extern "C" int fx_fmha_fwd(const void* q, int64_t q_stride_b, int64_t q_stride_l, const void* k, int64_t k_stride_b,
                           int64_t k_stride_l, const void* v, int64_t v_stride_b, int64_t v_stride_l, void* o,
                           int64_t o_stride_b, int64_t o_stride_l, int B, int H, int Lq, int Lk, float scale,
                           void* stream) {
  return fx::fmha_launch(q, q_stride_b, q_stride_l, k, k_stride_b, k_stride_l, v, v_stride_b, v_stride_l, o, nullptr, 0,
                         0, o_stride_b, o_stride_l, B, H, Lq, Lk, scale, stream, "fx_fmha_fwd");
}

extern "C" int fx_fmha_fwd_scatter(const void* q, int64_t q_stride_b, int64_t q_stride_l, const void* k,
                                   int64_t k_stride_b, int64_t k_stride_l, const void* v, int64_t v_stride_b,
                                   int64_t v_stride_l, void* const* o_peers, int n_peers, int rows_per_peer,
                                   int64_t o_stride_b, int64_t o_stride_l, int B, int H, int Lq, int Lk, float scale,
                                   void* stream) {
  FX_CHECK_ARG(o_peers != nullptr, "fx_fmha_fwd_scatter: null peer table");
  return fx::fmha_launch(q, q_stride_b, q_stride_l, k, k_stride_b, k_stride_l, v, v_stride_b, v_stride_l, nullptr,
                         o_peers, n_peers, rows_per_peer, o_stride_b, o_stride_l, B, H, Lq, Lk, scale, stream,
                         "fx_fmha_fwd_scatter");
}
