// Kernels of the umT5 text encoder (reference FlexAM/models/wan_text_encoder.py, cited as :line).
// The encoder runs ONCE per video on 2 x 512 tokens (pipeline _get_t5_prompt_embeds); its contractions (q|k|v, o, gated
// FFN: 9.5 TFLOP at the XXL size) go through the tcgen05 GEMM of gemm.cu. What is here is the rest, all in the
// reference's bf16 dtype flow (weights AND activations bf16, every torch op rounds):
//   token embedding gather :296, T5LayerNorm :51-56, self-attention with the per-layer relative-position bias and the
//   padding mask and NO 1/sqrt(d) scaling :75-109 (head_dim 64, 512 keys: 8.6 GFLOP per layer, a SIMT kernel with K^T
//   and V of one head resident in shared memory), the bf16 residual add :161-162 and the fc1 * GELU(gate) product :126.
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace fx {

static int te_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

// out[r, :] = table[ids[r], :]   (16-byte vectors; D % 8 == 0)
__global__ void embedding_kernel(const long long* ids, const uint4* table, uint4* out, long long rows, int dvec,
                                 long long vocab) {
  const long long total = rows * dvec;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const long long r = i / dvec;
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    out[i] = __ldg(table + id * dvec + (i - r * dvec));
  }
}

// T5LayerNorm :51-56: y = bf16(w * bf16(x * rsqrt(mean(x^2) + eps))), fp32 statistics, x bf16 -> out bf16.
// One warp per row, the row held in registers (NV 16-byte vectors per lane).
template <int NV>
__global__ void __launch_bounds__(256) t5_norm_kernel(const uint4* x, const uint4* w, uint4* out, int M, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= M) return;
  const int D = NV * 32 * 8;
  const uint4* xr = x + static_cast<long long>(row) * (D / 8);
  uint4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[i * 32 + lane];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
      ss += a * a + b * b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = rsqrtf(ss / static_cast<float>(D) + eps);
  uint4* orow = out + static_cast<long long>(row) * (D / 8);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint4 wv = __ldg(w + i * 32 + lane);
    const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
    const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t y = pack_bf16x2(bf16_lo(u[j]) * r, bf16_hi(u[j]) * r);   // fp32 product, one rounding
      uint32_t d;
      asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(ww[j]), "r"(y));        // bf16 * bf16 -> bf16
      o[j] = d;
    }
    orow[i * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// x = bf16(x + y) (the bf16 residual stream :161-162) and out = bf16(a * gelu_chain(g)) (:126 with GELU :38-41).
__global__ void add_bf16_kernel(uint4* x, const uint4* y, long long nvec) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < nvec; i += stride) {
    const uint4 a = x[i], b = y[i];
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(bf16_lo(ua[j]) + bf16_lo(ub[j]), bf16_hi(ua[j]) + bf16_hi(ub[j]));
    x[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// The reference's GELU is a chain of elementwise torch ops on a bf16 tensor (:38-41): each op rounds to bf16.
__device__ __forceinline__ float gelu_bf16_chain(float x) {
  const float x3 = bf16_round(x * x * x);                       // torch.pow(x, 3.0)
  const float inner = bf16_round(x + bf16_round(0.044715f * x3));
  const float th = bf16_round(tanhf(bf16_round(0.7978845608028654f * inner)));
  return bf16_round(bf16_round(0.5f * x) * bf16_round(1.0f + th));
}
__global__ void gated_gelu_kernel(const __nv_bfloat16* fc1, long long ld1, const __nv_bfloat16* gate, long long ldg,
                                  __nv_bfloat16* out, long long ldo, long long M, int nvec) {
  const long long total = M * nvec;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const long long m = i / nvec;
    const int c = static_cast<int>(i - m * nvec) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(fc1 + m * ld1 + c);
    const uint4 g = *reinterpret_cast<const uint4*>(gate + m * ldg + c);
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ug[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = pack_bf16x2(bf16_lo(ua[j]) * gelu_bf16_chain(bf16_lo(ug[j])), bf16_hi(ua[j]) * gelu_bf16_chain(bf16_hi(ug[j])));
    *reinterpret_cast<uint4*>(out + m * ldo + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// -------------------------------------------------------------------------------------------------
// T5 self-attention :75-109 for head_dim 64:  s_ij = bf16(bf16(q_i . k_j) + bias[h][j - i]) (masked keys: finfo.min),
// p = bf16(softmax_fp32(s)), o_i = bf16(sum_j p_ij v_j). No 1/sqrt(d) scaling.
//   grid (ceil(L / 32), H, B), 256 threads: warp w owns query rows 4w .. 4w+3 of the block's 32. K^T ([64][L], so a
//   warp reads 32 consecutive keys of one channel) and V ([L][64]) of the head live in shared memory as bf16; a lane owns
//   keys lane, lane+32, ... for the scores and channels 2*lane, 2*lane+1 for the output.
//   bias_rel: bf16 [H][2L-1], entry (h, j - i + L - 1) = pos_embedding[bucket(j - i)][h]  (T5RelativeEmbedding :219-253)
// -------------------------------------------------------------------------------------------------
constexpr int kT5MaxL = 512;
constexpr int kT5Rows = 32;

__global__ void __launch_bounds__(256)
t5_attention_kernel(const __nv_bfloat16* qkv, long long ld, const __nv_bfloat16* bias_rel, const int* mask,
                    __nv_bfloat16* out, long long ldo, int L, int H) {
  extern __shared__ uint8_t t5_smem[];
  __nv_bfloat16* kT = reinterpret_cast<__nv_bfloat16*>(t5_smem);            // [64][L + 2]
  __nv_bfloat16* vS = kT + 64 * (L + 2);                                    // [L][64]
  float* pS = reinterpret_cast<float*>(vS + static_cast<size_t>(L) * 64);   // [8 warps][4 rows][L]
  const int h = blockIdx.y, b = blockIdx.z;
  const int A = H * 64;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * L * ld;
  for (int i = threadIdx.x; i < L * 8; i += blockDim.x) {                   // 8 x 16-byte vectors per key row
    const int j = i >> 3, c8 = (i & 7) * 8;
    const uint4 kv = *reinterpret_cast<const uint4*>(base + static_cast<long long>(j) * ld + A + h * 64 + c8);
    const __nv_bfloat16* ke = reinterpret_cast<const __nv_bfloat16*>(&kv);
#pragma unroll
    for (int e = 0; e < 8; ++e) kT[(c8 + e) * (L + 2) + j] = ke[e];
    *reinterpret_cast<uint4*>(vS + j * 64 + c8) =
        *reinterpret_cast<const uint4*>(base + static_cast<long long>(j) * ld + 2 * A + h * 64 + c8);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* p = pS + warp * 4 * L;                                             // [4 rows][L] probabilities of this warp
  const int nk = (L + 31) / 32;
  const int i0 = blockIdx.x * kT5Rows + warp * 4;                           // the warp's 4 query rows: K / V read once
  if (i0 >= L) return;
  float q0[4], q1[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = min(i0 + r, L - 1);
    const uint32_t qq = *reinterpret_cast<const uint32_t*>(base + static_cast<long long>(i) * ld + h * 64 + 2 * lane);
    q0[r] = bf16_lo(qq);
    q1[r] = bf16_hi(qq);
  }
  float s[4][kT5MaxL / 32];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int t = 0; t < kT5MaxL / 32; ++t) s[r][t] = 0.f;
  for (int c = 0; c < 64; ++c) {
    float qc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) qc[r] = __shfl_sync(0xffffffffu, (c & 1) ? q1[r] : q0[r], c >> 1);
    const __nv_bfloat16* kr = kT + c * (L + 2);
#pragma unroll
    for (int t = 0; t < kT5MaxL / 32; ++t)
      if (t < nk) {
        const int j = t * 32 + lane;
        const float kv = j < L ? __bfloat162float(kr[j]) : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r][t] += qc[r] * kv;
      }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + r;
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < kT5MaxL / 32; ++t)
      if (t < nk) {
        const int j = t * 32 + lane;
        float v = -INFINITY;
        if (j < L && i < L) {
          const float bias = __bfloat162float(bias_rel[static_cast<long long>(h) * (2 * L - 1) + (j - i + L - 1)]);
          v = bf16_round(bf16_round(s[r][t]) + bias);
          if (mask != nullptr && mask[b * L + j] == 0) v = -3.3895313892515355e38f;   // torch.finfo(bfloat16).min
        }
        s[r][t] = v;
        mx = fmaxf(mx, v);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < kT5MaxL / 32; ++t)
      if (t < nk) {
        const int j = t * 32 + lane;
        const float e = (j < L && i < L) ? __expf(s[r][t] - mx) : 0.f;
        s[r][t] = e;
        sum += e;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
    for (int t = 0; t < kT5MaxL / 32; ++t)
      if (t < nk) {
        const int j = t * 32 + lane;
        if (j < L) p[r * L + j] = bf16_round(s[r][t] * inv);                  // .type_as(attn): bf16 probabilities
      }
  }
  __syncwarp();
  float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < L; ++j) {
    const uint32_t vv = *reinterpret_cast<const uint32_t*>(vS + j * 64 + 2 * lane);
    const float v0 = bf16_lo(vv), v1 = bf16_hi(vv);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float pj = p[r * L + j];
      o0[r] += pj * v0;
      o1[r] += pj * v1;
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (i0 + r < L)
      *reinterpret_cast<uint32_t*>(out + (static_cast<long long>(b) * L + i0 + r) * ldo + h * 64 + 2 * lane) =
          pack_bf16x2(o0[r], o1[r]);
}

// -------------------------------------------------------------------------------------------------
// The same attention on tcgen05 (default). One CTA = one (batch, head, 128-query tile); L <= 512 keys, head_dim 64:
//   TMA: Q tile [128][64], K and V tiles [128 keys][64] (SWIZZLE_128B boxes straight out of the packed q|k|v buffer);
//   S = Q K^T for ALL keys into TMEM columns [0, L) (UMMA 128x128x16, K-major operands; 512 fp32 columns = whole TMEM);
//   softmax by 128 threads (one query row each) in three passes over TMEM: (1) bf16(bf16(s) + bias) with the mask, row
//   maximum, value written back; (2) sum of exponentials; (3) p = bf16(e / sum) packed as bf16 pairs IN PLACE over the
//   scores (pair column c overwrites score columns already consumed: 16k < 32k);
//   O = P V with A = P from TMEM and B = V as an MN-major operand (never transposed), accumulator at columns [256, 320);
//   epilogue: O -> bf16 -> global. 3.5 waves of 512 CTAs for the 2 x 64 x 512 x 512 problem of the XXL encoder.
// -------------------------------------------------------------------------------------------------
constexpr int kT5Threads = 160;             // warps 0-3: softmax / epilogue (TMEM lane quadrants), warp 4: TMA + MMA issue
constexpr int kT5Tile = 128 * 64 * 2;       // one [128 rows][64 d] swizzled tile = 16 KB
constexpr int kT5TcSmem = kT5Tile * 9 + 256 + 1024;

__global__ void __launch_bounds__(kT5Threads, 1)
t5_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16* bias_rel, const int* mask,
                       __nv_bfloat16* out, long long ldo, int L, int H) {
  extern __shared__ uint8_t t5tc_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(t5tc_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kT5Tile;
  uint8_t* sV = sK + 4 * kT5Tile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 4 * kT5Tile);
  uint64_t* qk_full = bars;
  uint64_t* v_full = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3;
  uint64_t* o_done = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int A = H * 64;
  const int nt = (L + 127) / 128;           // 128-key tiles
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColO = 256;

  if (warp == 4) {
    if (lane == 0) {
      const int row0 = b * L;
      mbar_expect_tx(qk_full, kT5Tile * (1 + nt));
      tma_load_2d(sQ, &tmap, qk_full, h * 64, row0 + q0);
      for (int t = 0; t < nt; ++t) tma_load_2d(sK + t * kT5Tile, &tmap, qk_full, A + h * 64, row0 + t * 128);
      mbar_expect_tx(v_full, kT5Tile * nt);
      for (int t = 0; t < nt; ++t) tma_load_2d(sV + t * kT5Tile, &tmap, v_full, 2 * A + h * 64, row0 + t * 128);
      // S = Q K^T, one 128-key tile at a time
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);
      mbar_wait(qk_full, 0);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      for (int t = 0; t < nt; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem_base + t * 128, umma_desc_sw128(q_addr + k * 32, 16, 1024),
                  umma_desc_sw128(k_addr + t * kT5Tile + k * 32, 16, 1024), idesc_qk, k != 0);
      umma_commit(s_full);
      // O = P V once the probabilities are in TMEM
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, false, true);
      mbar_wait(v_full, 0);
      mbar_wait(p_full, 0);
      tc_fence_after();
      for (int t = 0; t < nt; ++t)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(tmem_base + kColO, tmem_base + t * 64 + k * 8,
                  umma_desc_sw128(v_addr + t * kT5Tile + k * 2048, kT5Tile, 1024), idesc_pv, (t | k) != 0);
      umma_commit(o_done);
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;                 // row of the tile = TMEM lane
    const int i = q0 + r;                           // query position inside the sample
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const __nv_bfloat16* brow = bias_rel + static_cast<long long>(h) * (2 * L - 1) + (L - 1 - min(i, L - 1));
    const int* mrow = mask != nullptr ? mask + b * L : nullptr;
    const int nchunk = nt * 4;                      // 32-column chunks
    mbar_wait(s_full, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c = 0; c < nchunk; ++c) {
      uint32_t v[32];
      tmem_ld32(trow + c * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int j = c * 32 + e;
        float x = -INFINITY;                        // keys beyond L do not exist
        if (j < L) {
          x = bf16_round(bf16_round(__uint_as_float(v[e])) + __bfloat162float(brow[j]));
          if (mrow != nullptr && mrow[j] == 0) x = -3.3895313892515355e38f;   // torch.finfo(bfloat16).min
        }
        v[e] = __float_as_uint(x);
        mx = fmaxf(mx, x);
      }
      tmem_st32(trow + c * 32, v);
    }
    tmem_wait_st();
    float sum = 0.f;
    for (int c = 0; c < nchunk; ++c) {
      uint32_t v[32];
      tmem_ld32(trow + c * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) sum += __expf(__uint_as_float(v[e]) - mx);
    }
    const float inv = 1.f / sum;
    for (int c = 0; c < nchunk; ++c) {
      uint32_t v[32], pk[16];
      tmem_ld32(trow + c * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 16; ++e)
        pk[e] = pack_bf16x2(__expf(__uint_as_float(v[2 * e]) - mx) * inv, __expf(__uint_as_float(v[2 * e + 1]) - mx) * inv);
      tmem_st16(trow + c * 16, pk);                 // pairs [16c, 16c+16) overwrite scores already consumed
    }
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(p_full);
    mbar_wait(o_done, 0);
    tc_fence_after();
    uint32_t o[32];
    __nv_bfloat16* orow = out + (static_cast<long long>(b) * L + i) * ldo + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      tmem_ld32(trow + kColO + c * 32, o);
      tmem_wait_ld();
      if (i < L) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
          w.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
          w.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
          w.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = w;
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

}  // namespace fx

extern "C" int fx_embedding_bf16(const int64_t* ids, const void* table, void* out, int64_t rows, int D, int64_t vocab,
                                 void* stream) {
  using namespace fx;
  FX_CHECK_ARG(ids && table && out && rows > 0 && D > 0 && D % 8 == 0 && vocab > 0, "fx_embedding_bf16: bad arguments");
  embedding_kernel<<<te_grid(rows * (D / 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), reinterpret_cast<const uint4*>(table), reinterpret_cast<uint4*>(out), rows,
      D / 8, vocab);
  FX_CHECK_LAUNCH("fx_embedding_bf16");
  return FX_OK;
}

extern "C" int fx_t5_layernorm(const void* x, const void* weight, void* out, int M, int D, float eps, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && weight && out && M > 0, "fx_t5_layernorm: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (M + 7) / 8;
  auto X = reinterpret_cast<const uint4*>(x);
  auto W = reinterpret_cast<const uint4*>(weight);
  auto O = reinterpret_cast<uint4*>(out);
  switch (D) {
    case 256: t5_norm_kernel<1><<<grid, 256, 0, s>>>(X, W, O, M, eps); break;
    case 512: t5_norm_kernel<2><<<grid, 256, 0, s>>>(X, W, O, M, eps); break;
    case 1024: t5_norm_kernel<4><<<grid, 256, 0, s>>>(X, W, O, M, eps); break;
    case 2048: t5_norm_kernel<8><<<grid, 256, 0, s>>>(X, W, O, M, eps); break;
    case 4096: t5_norm_kernel<16><<<grid, 256, 0, s>>>(X, W, O, M, eps); break;
    default:
      set_error("fx_t5_layernorm: unsupported D=%d (256, 512, 1024, 2048, 4096)", D);
      return FX_ERR_ARG;
  }
  FX_CHECK_LAUNCH("fx_t5_layernorm");
  return FX_OK;
}

extern "C" int fx_add_bf16(void* x, const void* y, int64_t n, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && y && n > 0 && n % 8 == 0, "fx_add_bf16: bad arguments (n must be a multiple of 8)");
  add_bf16_kernel<<<te_grid(n / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint4*>(x), reinterpret_cast<const uint4*>(y), n / 8);
  FX_CHECK_LAUNCH("fx_add_bf16");
  return FX_OK;
}

extern "C" int fx_gated_gelu_bf16(const void* fc1, int64_t ld1, const void* gate, int64_t ldg, void* out, int64_t ldo,
                                  int64_t M, int N, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(fc1 && gate && out && M > 0 && N > 0 && N % 8 == 0, "fx_gated_gelu_bf16: bad arguments (N % 8 != 0)");
  FX_CHECK_ARG(ld1 % 8 == 0 && ldg % 8 == 0 && ldo % 8 == 0 && ld1 >= N && ldg >= N && ldo >= N,
               "fx_gated_gelu_bf16: leading dimensions must be multiples of 8 and >= N");
  gated_gelu_kernel<<<te_grid(M * (N / 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(fc1), ld1, reinterpret_cast<const __nv_bfloat16*>(gate), ldg,
      reinterpret_cast<__nv_bfloat16*>(out), ldo, M, N / 8);
  FX_CHECK_LAUNCH("fx_gated_gelu_bf16");
  return FX_OK;
}

extern "C" int fx_t5_attention(const void* qkv, int64_t ld, const void* bias_rel, const int32_t* mask, void* out,
                               int64_t ldo, int B, int L, int H, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(qkv && bias_rel && out, "fx_t5_attention: null pointer");
  FX_CHECK_ARG(B > 0 && H > 0 && L > 0 && L <= kT5MaxL, "fx_t5_attention: B=%d H=%d L=%d (L <= %d)", B, H, L, kT5MaxL);
  FX_CHECK_ARG(ld % 8 == 0 && ld >= 3LL * H * 64 && ldo % 2 == 0 && ldo >= 1LL * H * 64, "fx_t5_attention: bad ld/ldo");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (tune_get("t5_attn_simt") == 1) {   // developer knob: the SIMT form (same rounding points), for A/B checks
    const int smem = 64 * (L + 2) * 2 + L * 64 * 2 + 8 * 4 * L * 4;
    if (smem > 48 * 1024 && !ensure_dyn_smem(reinterpret_cast<const void*>(t5_attention_kernel), smem, "fx_t5_attention"))
      return FX_ERR_CUDA;
    dim3 grid((L + kT5Rows - 1) / kT5Rows, H, B);
    t5_attention_kernel<<<grid, 256, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), ld,
                                                 reinterpret_cast<const __nv_bfloat16*>(bias_rel), mask,
                                                 reinterpret_cast<__nv_bfloat16*>(out), ldo, L, H);
    FX_CHECK_LAUNCH("fx_t5_attention(simt)");
    return FX_OK;
  }
  FX_CHECK_ARG(reinterpret_cast<uintptr_t>(qkv) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 && ldo % 8 == 0,
               "fx_t5_attention: qkv / out must be 16-byte aligned, ldo a multiple of 8");
  CUtensorMap tmap;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(3LL * H * 64), static_cast<uint64_t>(B) * L};
    const uint64_t strides[1] = {static_cast<uint64_t>(ld) * 2};
    const uint32_t box[2] = {64, 128};
    if (!make_tmap_bf16(&tmap, qkv, 2, dims, strides, box)) return FX_ERR_CUDA;
  }
  if (!ensure_dyn_smem(reinterpret_cast<const void*>(t5_attention_tc_kernel), kT5TcSmem, "fx_t5_attention"))
    return FX_ERR_CUDA;
  dim3 grid((L + 127) / 128, H, B);
  t5_attention_tc_kernel<<<grid, kT5Threads, kT5TcSmem, st>>>(tmap, reinterpret_cast<const __nv_bfloat16*>(bias_rel), mask,
                                                              reinterpret_cast<__nv_bfloat16*>(out), ldo, L, H);
  FX_CHECK_LAUNCH("fx_t5_attention");
  return FX_OK;
}
