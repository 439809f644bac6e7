// Non-causal flash-attention forward for sm_100a, head_dim 128, bf16 in/out, fp32 softmax statistics.
//
// One CTA owns 256 query rows (two 128-row tiles) of one (batch, head) and streams 128-key K/V tiles:
//   warps 0-15  softmax        four warpgroups: (query tile w, column half h) = warp >> 2 -> (w = wg >> 1, h = wg & 1).
//                              A thread owns one row of S_w and 64 of its 128 keys; the two threads of a row agree on
//                              the running maximum through shared memory + a 256-thread named barrier per tile. Four
//                              softmax warps per scheduler (instead of two) is what hides the MUFU/TMEM latencies.
//   warp 16     TMA producer   Q once, then K_j / V_j into 2-stage rings (SWIZZLE_128B boxes of [128 rows][64 d])
//   warp 17     MMA issuer     S_w = Q_w K_j^T (SS, both K-major) and O_w += P_w V_j (A = P from TMEM, B = V MN-major)
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_w (bf16 pairs) overwrites S_w in place:
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
constexpr int kDefaultPoly8 = 2;               // 2 of 8 exponential pairs on the FMA pipe (measured best: 0 2 3 4)

// Developer tracing (tests/native/fmha_trace.cu builds this file with -DFX_FMHA_TRACE): CTA (0,0,0) records
// clock64() at pipeline events of its first 64 KV steps. Compiled out of the library.
#ifdef FX_FMHA_TRACE
__device__ long long fx_fmha_trace[16 * 64];
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

  // exp2 split between MUFU and the FMA pipe, in eighths of the pairs; FX_FMHA_POLY=0|2|3|4 overrides the default
  static int poly = -1, token = 1;
  if (poly < 0) {
    const char* env = getenv("FX_FMHA_POLY");
    poly = kDefaultPoly8;
    if (env && (env[0] == '0' || env[0] == '2' || env[0] == '3' || env[0] == '4')) poly = env[0] - '0';
    const char* t = getenv("FX_FMHA_TOKEN");
    if (t) token = t[0] != '0';
  }
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const FmhaParams);
  const Kern kerns[5] = {fmha_fwd_kernel<0>, nullptr, fmha_fwd_kernel<2>, fmha_fwd_kernel<3>, fmha_fwd_kernel<4>};
  const Kern kern = kerns[poly];
  static bool configured[5] = {false, false, false, false, false};
  if (!configured[poly]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFmhaSmem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
      return FX_ERR_CUDA;
    }
    configured[poly] = true;
  }
  FmhaParams p{};
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.o_stride_b = o_stride_b;
  p.o_stride_l = o_stride_l;
  p.Lq = Lq;
  p.Lk = Lk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.rows_per_peer = 0;
  p.token = token;
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
  kern<<<grid, kFmhaThreads, kFmhaSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tq, tk, tv, p);
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
