// HBM-bound row kernels of the DiT block: LayerNorm + adaLN modulation, LayerNorm + affine, full-width
// RMSNorm (+ 3-axis RoPE), and the sampler's fused CFG/Euler update. One warp owns one row, keeps it in
// registers (single global read), reduces with shuffles and writes 8/16-byte vectors.
#include "host_common.h"
#include "ptx.cuh"

namespace fx {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// -------------------------------------------------------------------------------------------------
// LayerNorm (no affine) with either adaLN modulation (MODE 0) or bf16 gamma/beta (MODE 1)
//   WanAttentionBlock.forward :444-453,:464-465; Head.forward :493-507; norm3 :405-407,:461
// -------------------------------------------------------------------------------------------------
struct LnParams {
  const float* x;
  __nv_bfloat16* out;
  int M, D;
  float eps;
  // MODE 0
  const float* shift_mod;
  const float* scale_mod;
  const float* shift_e;
  const float* scale_e;
  long long e_stride;
  const int* row_idx;
  const float* dens_mod;
  const float* dens;
  long long dens_stride;
  int rows_per_batch;
  // MODE 1
  const __nv_bfloat16* gamma;
  const __nv_bfloat16* beta;
  // MODE 2: precombined fp32 rows, out = LN(x) * tab_scale[row_idx[m]] + tab_shift[row_idx[m]]
  const float* tab_scale;
  const float* tab_shift;
  long long tab_stride;
};

// A row is owned by WPR warps (4 for wide rows: 6 float4 per lane at D = 3072 keeps the kernel at ~60 registers and
// full occupancy; 1 for narrow rows). Row statistics are reduced with shuffles, then across the row's warps via smem.
template <int WPR>
__device__ __forceinline__ float row_reduce(float v, float* red, int row_in_block, int warp_in_row, int lane) {
  v = warp_sum(v);
  if constexpr (WPR == 1) return v;
  if (lane == 0) red[row_in_block * WPR + warp_in_row] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < WPR; ++i) t += red[row_in_block * WPR + i];
  __syncthreads();
  return t;
}

template <int NV, int WPR, int MODE>  // NV = float4 vectors per lane = D / (128 * WPR)
__global__ void __launch_bounds__(256) ln_kernel(const LnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_in_block = warp / WPR, warp_in_row = warp % WPR;
  const int row = blockIdx.x * (8 / WPR) + row_in_block;
  const bool valid = row < p.M;
  const int c0 = warp_in_row * NV * 32;  // first float4 column of this warp
  const float4* xr = reinterpret_cast<const float4*>(p.x + static_cast<long long>(valid ? row : 0) * p.D) + c0;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = valid ? xr[i * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = row_reduce<WPR>(s, red, row_in_block, warp_in_row, lane) / static_cast<float>(p.D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(row_reduce<WPR>(q, red, row_in_block, warp_in_row, lane) / static_cast<float>(p.D) + p.eps);
  if (!valid) return;

  uint2* orow = reinterpret_cast<uint2*>(p.out + static_cast<long long>(row) * p.D) + c0;
  if constexpr (MODE == 0) {
    const long long u = p.row_idx ? p.row_idx[row] : 0;
    const float4* sh_e = reinterpret_cast<const float4*>(p.shift_e + u * p.e_stride) + c0;
    const float4* sc_e = reinterpret_cast<const float4*>(p.scale_e + u * p.e_stride) + c0;
    const float4* sh_m = reinterpret_cast<const float4*>(p.shift_mod) + c0;
    const float4* sc_m = reinterpret_cast<const float4*>(p.scale_mod) + c0;
    const float4* dn =
        p.dens ? reinterpret_cast<const float4*>(p.dens + static_cast<long long>(row / p.rows_per_batch) * p.dens_stride) + c0
               : nullptr;
    const float4* dm = p.dens_mod ? reinterpret_cast<const float4*>(p.dens_mod) + c0 : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      const float4 sm = __ldg(sc_m + c), se = __ldg(sc_e + c), hm = __ldg(sh_m + c), he = __ldg(sh_e + c);
      float4 sc, sh;
      sc.x = 1.f + (sm.x + se.x); sc.y = 1.f + (sm.y + se.y); sc.z = 1.f + (sm.z + se.z); sc.w = 1.f + (sm.w + se.w);
      sh.x = hm.x + he.x; sh.y = hm.y + he.y; sh.z = hm.z + he.z; sh.w = hm.w + he.w;
      float4 y;
      y.x = (v[i].x - mean) * rstd * sc.x + sh.x;
      y.y = (v[i].y - mean) * rstd * sc.y + sh.y;
      y.z = (v[i].z - mean) * rstd * sc.z + sh.z;
      y.w = (v[i].w - mean) * rstd * sc.w + sh.w;
      if (dn != nullptr) {
        float4 d = __ldg(dn + c);
        if (dm != nullptr) {
          const float4 m = __ldg(dm + c);
          d.x += m.x; d.y += m.y; d.z += m.z; d.w += m.w;
        }
        y.x += d.x; y.y += d.y; y.z += d.z; y.w += d.w;
      }
      uint2 o;
      o.x = pack_bf16x2(y.x, y.y);
      o.y = pack_bf16x2(y.z, y.w);
      orow[c] = o;
    }
  } else if constexpr (MODE == 2) {
    const long long u = p.row_idx ? p.row_idx[row] : 0;
    const float4* sc_t = reinterpret_cast<const float4*>(p.tab_scale + u * p.tab_stride) + c0;
    const float4* sh_t = reinterpret_cast<const float4*>(p.tab_shift + u * p.tab_stride) + c0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      const float4 sc = __ldg(sc_t + c), sh = __ldg(sh_t + c);
      uint2 o;
      o.x = pack_bf16x2((v[i].x - mean) * rstd * sc.x + sh.x, (v[i].y - mean) * rstd * sc.y + sh.y);
      o.y = pack_bf16x2((v[i].z - mean) * rstd * sc.z + sh.z, (v[i].w - mean) * rstd * sc.w + sh.w);
      orow[c] = o;
    }
  } else {
    const uint2* gm = reinterpret_cast<const uint2*>(p.gamma) + c0;
    const uint2* bt = reinterpret_cast<const uint2*>(p.beta) + c0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      const uint2 g = __ldg(gm + c), b = __ldg(bt + c);
      float4 y;
      y.x = (v[i].x - mean) * rstd * bf16_lo(g.x) + bf16_lo(b.x);
      y.y = (v[i].y - mean) * rstd * bf16_hi(g.x) + bf16_hi(b.x);
      y.z = (v[i].z - mean) * rstd * bf16_lo(g.y) + bf16_lo(b.y);
      y.w = (v[i].w - mean) * rstd * bf16_hi(g.y) + bf16_hi(b.y);
      uint2 o;
      o.x = pack_bf16x2(y.x, y.y);
      o.y = pack_bf16x2(y.z, y.w);
      orow[c] = o;
    }
  }
}

template <int MODE>
static int launch_ln(const LnParams& p, cudaStream_t s, const char* name) {
  // wide rows: 4 warps per row (D % 512 == 0); narrow rows: one warp per row
  const bool wide = (p.D % 512 == 0) && p.D >= 2048;
  const int wpr = wide ? 4 : 1;
  const int nv = p.D / (128 * wpr);
  const int rows_per_block = 8 / wpr;
  const int grid = (p.M + rows_per_block - 1) / rows_per_block;
#define FX_LN_CASE(NVV, WPRV)                                   \
  if (nv == NVV && wpr == WPRV) {                               \
    launch_kernel(ln_kernel<NVV, WPRV, MODE>, dim3(grid), dim3(256), 0, s, p); \
    FX_CHECK_LAUNCH(name);                                      \
    return FX_OK;                                               \
  }
  FX_LN_CASE(1, 1) FX_LN_CASE(2, 1) FX_LN_CASE(4, 1) FX_LN_CASE(8, 1) FX_LN_CASE(12, 1)
  FX_LN_CASE(4, 4) FX_LN_CASE(5, 4) FX_LN_CASE(6, 4) FX_LN_CASE(8, 4) FX_LN_CASE(10, 4) FX_LN_CASE(12, 4) FX_LN_CASE(16, 4)
#undef FX_LN_CASE
  set_error("%s: unsupported D=%d (supported: 128,256,512,1024,1536 and 2048,2560,3072,4096,5120,6144,8192)", name, p.D);
  return FX_ERR_ARG;
}

// -------------------------------------------------------------------------------------------------
// Full-width RMSNorm (+ RoPE), in place on bf16 rows.  WanRMSNorm :173-189 at :242-243/:363-364;
// rope_apply :135-164 with the [22,21,21] frame/row/col split of the 64 complex pairs per head.
// One launch covers q and k (two D-wide column blocks of the packed projection output) when w2 is given.
// -------------------------------------------------------------------------------------------------
struct RmsParams {
  __nv_bfloat16* x;
  long long ldx;
  int M, D, ntensors;
  float eps;
  const __nv_bfloat16* w;
  const __nv_bfloat16* w2;
  const float2* freqs;  // [1024][64] (cos, sin) or nullptr
  int gf, gh, gw, tok_offset, rows_per_batch;
  // Ulysses head scatter (n_peers > 0): instead of writing in place, every 128-wide head h of tensor `which` of row
  // (b, t) goes to rank h / heads_per_peer at [(b * ntensors + which) * dst_rows + dst_row0 + t][h % heads_per_peer]
  // of that rank's exchange buffer (peer memory over NVLink; the own slice is a local pointer). norm_tensors: how
  // many of the leading tensors are normed + rotated (the rest - v - is moved as is).
  int n_peers, heads_per_peer, norm_tensors, dst_row0;
  long long dst_rows;
  __nv_bfloat16* dst[8];
};

// bf16x2 product rounded once to bf16 (mul.rn.bf16x2). For bf16 operands this equals bf16(float(a) * float(b)): the
// fp32 product of two 8-bit significands is exact, so both forms round the exact product once — the reference's
// `x * r.to(x.dtype)` and `(...) * weight` roundings (:186-189) at one instruction per pair instead of unpack, FMUL,
// and a quarter-rate F2F conversion per element.
__device__ __forceinline__ uint32_t bf16x2_mul(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

template <int NV, int WPR>  // NV = 16-byte vectors (8 bf16) per lane = D / (256 * WPR)
__global__ void __launch_bounds__(256, NV <= 4 ? 6 : 3) rmsnorm_rope_kernel(const __grid_constant__ RmsParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_in_block = warp / WPR, warp_in_row = warp % WPR;
  const int vrow = blockIdx.x * (8 / WPR) + row_in_block;  // virtual row = (token row, tensor)
  const bool valid = vrow < p.M * p.ntensors;
  const int row = valid ? vrow / p.ntensors : 0;
  const int which = valid ? vrow - row * p.ntensors : 0;
  const int c0 = warp_in_row * NV * 32;  // first 16-byte vector of this warp inside the row
  uint4* xr = reinterpret_cast<uint4*>(p.x + static_cast<long long>(row) * p.ldx + static_cast<long long>(which) * p.D) + c0;
  uint4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = valid ? xr[i * 32 + lane] : make_uint4(0, 0, 0, 0);

  // While the row is in flight: the rotation of this lane. Every vector a lane owns sits at the same position inside
  // its 128-wide head ((c0 + i*32 + lane) & 15 == lane & 15: c0 and 32 are multiples of 16), so the lane needs only
  // the four (cos, sin) pairs pair0 .. pair0+3 of its token: four table reads per row instead of four per vector.
  const bool normed = which < p.norm_tensors;
  const int b = row / p.rows_per_batch;
  const int tl = row - b * p.rows_per_batch;  // token inside this rank's slice of sample b
  bool rotate = false;
  float2 cs[4];
  if (p.freqs != nullptr && normed && valid) {
    const int t = p.tok_offset + tl;
    if (t < p.gf * p.gh * p.gw) {
      rotate = true;
      const int pf = t / (p.gh * p.gw);
      const int rem = t - pf * (p.gh * p.gw);
      const int ph = rem / p.gw;
      const int pw = rem - ph * p.gw;
      const int pair0 = (lane & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pj = pair0 + j;
        const int pos = pj < 22 ? pf : (pj < 43 ? ph : pw);
        cs[j] = __ldg(p.freqs + pos * 64 + pj);
      }
    }
  }
  const uint4* wr = reinterpret_cast<const uint4*>(which == 0 ? p.w : p.w2) + c0;

  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(u[j]), bb = bf16_hi(u[j]);
      ss += a * a + bb * bb;
    }
  }
  // r is rounded to bf16 before the multiply, as `.to(x.dtype)` does at :189
  const float rf = rsqrtf(row_reduce<WPR>(ss, red, row_in_block, warp_in_row, lane) / static_cast<float>(p.D) + p.eps);
  const uint32_t r2 = pack_bf16x2(rf, rf);
  if (!valid) return;

  const long long drow = (static_cast<long long>(b) * p.ntensors + which) * p.dst_rows + p.dst_row0 + tl;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 32 + lane;  // vector index inside this warp's span
    uint4 ov = v[i];
    if (normed) {
      const uint4 wv = __ldg(wr + c);   // 6 KB per tensor, L1-resident; loading it here keeps the kernel at <= 40 registers
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = bf16x2_mul(bf16x2_mul(u[j], r2), ww[j]);   // bf16(bf16(x * r) * w), both halves
        if (rotate) {
          const float a = bf16_lo(o[j]), bb = bf16_hi(o[j]);
          o[j] = pack_bf16x2(a * cs[j].x - bb * cs[j].y, a * cs[j].y + bb * cs[j].x);
        }
      }
      ov = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (p.n_peers == 0) {
      xr[c] = ov;
    } else {
      const int gv = c0 + c;       // 16-byte vector index inside the D-wide row: 16 vectors per head
      const int head = gv >> 4;
      const int peer = head / p.heads_per_peer;
      const int hl = head - peer * p.heads_per_peer;
      uint4* d = reinterpret_cast<uint4*>(p.dst[peer]) + (drow * p.heads_per_peer + hl) * 16 + (gv & 15);
      *d = ov;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Per-block modulation rows (:444-449) combined once per (timestep u, sample b) instead of once per token:
//   tab[s][u*B + b][0][:] = 1 + (mod[3s+1] + e0[u][3s+1])                       s = 0: self-attention LN, 1: ffn LN
//   tab[s][u*B + b][1][:] = (mod[3s] + e0[u][3s]) + (dmod[s] + de0[b][s])
// -------------------------------------------------------------------------------------------------
__global__ void modulation_tables_kernel(const float* mod, const float* dmod, const float* e0, const float* de0, int U,
                                         int B, int D, float* tab) {
  const long long total = 2LL * U * B * D;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int d = static_cast<int>(i % D);
    long long r = i / D;
    const int b = static_cast<int>(r % B);
    r /= B;
    const int u = static_cast<int>(r % U);
    const int s = static_cast<int>(r / U);
    const float* e = e0 + static_cast<long long>(u) * 6 * D;
    float* row = tab + ((static_cast<long long>(s) * U * B + static_cast<long long>(u) * B + b) * 2) * D;
    row[d] = 1.f + (mod[(3 * s + 1) * D + d] + e[(3 * s + 1) * D + d]);
    row[D + d] = (mod[3 * s * D + d] + e[3 * s * D + d]) + (dmod[s * D + d] + de0[(static_cast<long long>(b) * 2 + s) * D + d]);
  }
}

// -------------------------------------------------------------------------------------------------
__global__ void cfg_euler_kernel(const __nv_bfloat16* vu, const __nv_bfloat16* vc, float guidance, float dsigma,
                                 float* lat, const float* mask, const __nv_bfloat16* pinned, long long n) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float u = __bfloat162float(vu[i]);
    const float c = __bfloat162float(vc[i]);
    // the reference combines in the model dtype (bf16, one rounding per tensor op) at pipeline :928; the Euler
    // update adds (sigma' - sigma) * v — a 0-dim fp32 tensor times a bf16 tensor, i.e. a bf16 product under torch's
    // type promotion — to the fp32 sample and casts back to the model dtype, so `lat` holds bf16-representable values.
    const float v = bf16_round(u + bf16_round(guidance * bf16_round(c - u)));
    float x = bf16_round(lat[i] + bf16_round(dsigma * v));
    if (mask != nullptr) {
      const float m = mask[i];
      x = bf16_round(bf16_round((1.f - m) * __bfloat162float(pinned[i])) + bf16_round(m * x));
    }
    lat[i] = x;
  }
}

__global__ void cast_f32_bf16_kernel(const float* s, __nv_bfloat16* d, long long n) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) d[i] = __float2bfloat16_rn(s[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* s, float* d, long long n) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) d[i] = __bfloat162float(s[i]);
}

static int ew_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace fx

extern "C" int fx_ln_modulate(const float* x, void* out, int M, int D, float eps, const float* shift_mod,
                              const float* scale_mod, const float* shift_e, const float* scale_e, int64_t e_stride,
                              const int32_t* row_idx, const float* dens_mod, const float* dens, int64_t dens_stride,
                              int rows_per_batch, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && out && shift_mod && scale_mod && shift_e && scale_e, "fx_ln_modulate: null pointer");
  FX_CHECK_ARG(M > 0 && D > 0 && D % 128 == 0, "fx_ln_modulate: bad shape M=%d D=%d", M, D);
  FX_CHECK_ARG(e_stride % 4 == 0 && dens_stride % 4 == 0 && rows_per_batch > 0, "fx_ln_modulate: bad strides");
  LnParams p{};
  p.x = x; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.M = M; p.D = D; p.eps = eps;
  p.shift_mod = shift_mod; p.scale_mod = scale_mod; p.shift_e = shift_e; p.scale_e = scale_e;
  p.e_stride = e_stride; p.row_idx = row_idx; p.dens_mod = dens_mod; p.dens = dens; p.dens_stride = dens_stride;
  p.rows_per_batch = rows_per_batch;
  return launch_ln<0>(p, reinterpret_cast<cudaStream_t>(stream), "fx_ln_modulate");
}

extern "C" int fx_ln_affine(const float* x, void* out, int M, int D, float eps, const void* gamma, const void* beta,
                            void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && out && gamma && beta, "fx_ln_affine: null pointer");
  FX_CHECK_ARG(M > 0 && D > 0 && D % 128 == 0, "fx_ln_affine: bad shape M=%d D=%d", M, D);
  LnParams p{};
  p.x = x; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.M = M; p.D = D; p.eps = eps;
  p.gamma = reinterpret_cast<const __nv_bfloat16*>(gamma);
  p.beta = reinterpret_cast<const __nv_bfloat16*>(beta);
  return launch_ln<1>(p, reinterpret_cast<cudaStream_t>(stream), "fx_ln_affine");
}

extern "C" int fx_ln_scale_shift(const float* x, void* out, int M, int D, float eps, const float* scale,
                                 const float* shift, int64_t row_stride, const int32_t* row_idx, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && out && scale && shift, "fx_ln_scale_shift: null pointer");
  FX_CHECK_ARG(M > 0 && D > 0 && D % 128 == 0 && row_stride % 4 == 0, "fx_ln_scale_shift: bad shape M=%d D=%d", M, D);
  FX_CHECK_ARG((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) % 16 == 0,
               "fx_ln_scale_shift: tables must be 16-byte aligned");
  LnParams p{};
  p.x = x; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.M = M; p.D = D; p.eps = eps;
  p.tab_scale = scale; p.tab_shift = shift; p.tab_stride = row_stride; p.row_idx = row_idx;
  return launch_ln<2>(p, reinterpret_cast<cudaStream_t>(stream), "fx_ln_scale_shift");
}

extern "C" int fx_modulation_tables(const float* mod, const float* dmod, const float* e0, const float* de0, int U,
                                    int B, int D, float* tab, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(mod && dmod && e0 && de0 && tab, "fx_modulation_tables: null pointer");
  FX_CHECK_ARG(U > 0 && B > 0 && D > 0, "fx_modulation_tables: bad shape U=%d B=%d D=%d", U, B, D);
  modulation_tables_kernel<<<ew_grid(2LL * U * B * D), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      mod, dmod, e0, de0, U, B, D, tab);
  FX_CHECK_LAUNCH("fx_modulation_tables");
  return FX_OK;
}

namespace fx {
static int launch_rmsnorm_rope(const RmsParams& p, cudaStream_t s, const char* name) {
  const int D = p.D;
  const bool wide = (D % 1024 == 0) && D >= 2048;
  const int wpr = wide ? 4 : 1;
  const int nv = D / (256 * wpr);
  const long long vrows = static_cast<long long>(p.M) * p.ntensors;
  const int grid = static_cast<int>((vrows + (8 / wpr) - 1) / (8 / wpr));
#define FX_RMS_CASE(NVV, WPRV)                                   \
  if (nv == NVV && wpr == WPRV) {                                \
    launch_kernel(rmsnorm_rope_kernel<NVV, WPRV>, dim3(grid), dim3(256), 0, s, p); \
    FX_CHECK_LAUNCH(name);                                       \
    return FX_OK;                                                \
  }
  FX_RMS_CASE(1, 1) FX_RMS_CASE(2, 1) FX_RMS_CASE(4, 1) FX_RMS_CASE(6, 1) FX_RMS_CASE(10, 1)
  FX_RMS_CASE(2, 4) FX_RMS_CASE(3, 4) FX_RMS_CASE(4, 4) FX_RMS_CASE(5, 4) FX_RMS_CASE(6, 4) FX_RMS_CASE(8, 4)
#undef FX_RMS_CASE
  set_error("%s: unsupported D=%d (supported: 256,512,1024,1536,2560 and 2048,3072,4096,5120,6144,8192)", name, D);
  return FX_ERR_ARG;
}
}  // namespace fx

extern "C" int fx_rmsnorm_rope(void* x, int64_t ldx, int M, int D, float eps, const void* weight,
                               const void* weight2, const float* freqs, int gf, int gh, int gw, int tok_offset,
                               int rows_per_batch, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && weight, "fx_rmsnorm_rope: null pointer");
  const int nt = weight2 ? 2 : 1;
  FX_CHECK_ARG(M > 0 && D > 0 && D % 256 == 0 && ldx % 8 == 0 && ldx >= static_cast<int64_t>(nt) * D,
               "fx_rmsnorm_rope: bad shape M=%d D=%d ldx=%lld", M, D, (long long)ldx);
  if (freqs != nullptr) {
    FX_CHECK_ARG(gf > 0 && gh > 0 && gw > 0 && gf <= 1024 && gh <= 1024 && gw <= 1024 && rows_per_batch > 0,
                 "fx_rmsnorm_rope: grid (%d,%d,%d) outside the 1024-entry RoPE table", gf, gh, gw);
  } else {
    rows_per_batch = rows_per_batch > 0 ? rows_per_batch : 1;
  }
  RmsParams p{};
  p.x = reinterpret_cast<__nv_bfloat16*>(x); p.ldx = ldx; p.M = M; p.D = D; p.ntensors = nt; p.eps = eps;
  p.w = reinterpret_cast<const __nv_bfloat16*>(weight);
  p.w2 = reinterpret_cast<const __nv_bfloat16*>(weight2);
  p.freqs = reinterpret_cast<const float2*>(freqs);
  p.gf = gf; p.gh = gh; p.gw = gw; p.tok_offset = tok_offset; p.rows_per_batch = rows_per_batch;
  p.norm_tensors = nt;
  return launch_rmsnorm_rope(p, reinterpret_cast<cudaStream_t>(stream), "fx_rmsnorm_rope");
}

extern "C" int fx_qkv_norm_rope_scatter(const void* qkv, int64_t ldx, int M, int D, float eps, const void* weight_q,
                                        const void* weight_k, const float* freqs, int gf, int gh, int gw,
                                        int tok_offset, int rows_per_batch, void* const* peers, int n_peers,
                                        int heads_per_peer, int64_t dst_rows, int dst_row0, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(qkv && weight_q && weight_k && freqs && peers, "fx_qkv_norm_rope_scatter: null pointer");
  FX_CHECK_ARG(M > 0 && D > 0 && D % 256 == 0 && ldx % 8 == 0 && ldx >= 3LL * D,
               "fx_qkv_norm_rope_scatter: bad shape M=%d D=%d ldx=%lld", M, D, (long long)ldx);
  FX_CHECK_ARG(gf > 0 && gh > 0 && gw > 0 && gf <= 1024 && gh <= 1024 && gw <= 1024 && rows_per_batch > 0 &&
                   M % rows_per_batch == 0,
               "fx_qkv_norm_rope_scatter: bad grid (%d,%d,%d) / rows_per_batch %d", gf, gh, gw, rows_per_batch);
  FX_CHECK_ARG(n_peers >= 1 && n_peers <= 8 && heads_per_peer > 0 && n_peers * heads_per_peer * 128 == D,
               "fx_qkv_norm_rope_scatter: %d peers x %d heads x 128 != D=%d", n_peers, heads_per_peer, D);
  FX_CHECK_ARG(dst_row0 >= 0 && dst_rows >= static_cast<int64_t>(dst_row0) + rows_per_batch,
               "fx_qkv_norm_rope_scatter: destination rows [%d, +%d) outside %lld", dst_row0, rows_per_batch,
               (long long)dst_rows);
  RmsParams p{};
  p.x = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(qkv)); p.ldx = ldx; p.M = M; p.D = D; p.ntensors = 3;
  p.eps = eps;
  p.w = reinterpret_cast<const __nv_bfloat16*>(weight_q);
  p.w2 = reinterpret_cast<const __nv_bfloat16*>(weight_k);
  p.freqs = reinterpret_cast<const float2*>(freqs);
  p.gf = gf; p.gh = gh; p.gw = gw; p.tok_offset = tok_offset; p.rows_per_batch = rows_per_batch;
  p.norm_tensors = 2;
  p.n_peers = n_peers; p.heads_per_peer = heads_per_peer; p.dst_rows = dst_rows; p.dst_row0 = dst_row0;
  for (int i = 0; i < n_peers; ++i) {
    FX_CHECK_ARG(peers[i] != nullptr && reinterpret_cast<uintptr_t>(peers[i]) % 16 == 0,
                 "fx_qkv_norm_rope_scatter: peer buffer %d null or misaligned", i);
    p.dst[i] = reinterpret_cast<__nv_bfloat16*>(peers[i]);
  }
  return launch_rmsnorm_rope(p, reinterpret_cast<cudaStream_t>(stream), "fx_qkv_norm_rope_scatter");
}

extern "C" int fx_cfg_euler_step(const void* vu, const void* vc, float guidance, float dsigma, float* lat,
                                 const float* mask, const void* pinned, int64_t n, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(vu && vc && lat && n > 0, "fx_cfg_euler_step: null pointer or empty");
  FX_CHECK_ARG(mask == nullptr || pinned != nullptr, "fx_cfg_euler_step: mask given without pinned latents");
  cfg_euler_kernel<<<ew_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(vu), reinterpret_cast<const __nv_bfloat16*>(vc), guidance, dsigma, lat,
      mask, reinterpret_cast<const __nv_bfloat16*>(pinned), n);
  FX_CHECK_LAUNCH("fx_cfg_euler_step");
  return FX_OK;
}

extern "C" int fx_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(src && dst && n > 0, "fx_cast_f32_to_bf16: null pointer or empty");
  cast_f32_bf16_kernel<<<ew_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  FX_CHECK_LAUNCH("fx_cast_f32_to_bf16");
  return FX_OK;
}
extern "C" int fx_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(src && dst && n > 0, "fx_cast_bf16_to_f32: null pointer or empty");
  cast_bf16_f32_kernel<<<ew_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
  FX_CHECK_LAUNCH("fx_cast_bf16_to_f32");
  return FX_OK;
}

// -------------------------------------------------------------------------------------------------
// Sampled fingerprint of a set of weight tensors: detects edits made behind autograd's back (`weight.data += ...`,
// the reference's merge_lora / unmerge_lora, FlexAM/utils/lora_utils.py:481-485, :595-599) to tensors the engine keeps
// COPIES of (packed q|k|v, cross k|v, fp32 modulation rows) or derived results of (cached cross-attention K/V).
// out[t] = sum over every `stride`-th 16-byte word w of tensor t of mix(w, index) (64-bit integer add: order-free,
// deterministic). A dense edit (LoRA delta = B @ A touches every element) changes every sample.
// -------------------------------------------------------------------------------------------------
namespace fx {
__global__ void fingerprint_kernel(const void* const* ptrs, const long long* nbytes, int n, int stride,
                                   unsigned long long* out) {
  const int t = blockIdx.y;
  if (t >= n) return;
  const uint4* p = reinterpret_cast<const uint4*>(ptrs[t]);
  const long long nvec = nbytes[t] / 16;
  unsigned long long acc = 0;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * stride; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x * stride) {
    const uint4 v = __ldg(p + i);
    unsigned long long h = (static_cast<unsigned long long>(v.x ^ v.z) << 32) | (v.y ^ v.w);
    h ^= static_cast<unsigned long long>(i) * 0x9E3779B97F4A7C15ull;
    h = (h ^ (h >> 31)) * 0xBF58476D1CE4E5B9ull;
    acc += h ^ (h >> 29);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd(out + t, acc);
}
}  // namespace fx

extern "C" int fx_fingerprint(const void* const* ptrs, const int64_t* nbytes, int n, int stride, uint64_t* out,
                              void* stream) {
  using namespace fx;
  FX_CHECK_ARG(ptrs && nbytes && out && n > 0 && stride > 0, "fx_fingerprint: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(uint64_t) * n, s);
  if (e != cudaSuccess) {
    set_error("fx_fingerprint: cudaMemsetAsync: %s", cudaGetErrorString(e));
    return FX_ERR_CUDA;
  }
  fingerprint_kernel<<<dim3(4, n), 256, 0, s>>>(ptrs, reinterpret_cast<const long long*>(nbytes), n, stride,
                                                 reinterpret_cast<unsigned long long*>(out));
  FX_CHECK_LAUNCH("fx_fingerprint");
  return FX_OK;
}

// -------------------------------------------------------------------------------------------------
// Ulysses exchange layout + TeaCache residual helpers
// -------------------------------------------------------------------------------------------------
namespace fx {
__global__ void swap01_kernel(const uint4* in, long long ld_a_vec, uint4* out, int A, int B, int inner_vec) {
  const long long total = static_cast<long long>(A) * B * inner_vec;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int c = static_cast<int>(i % inner_vec);
    const long long ab = i / inner_vec;
    const int b = static_cast<int>(ab % B);
    const int a = static_cast<int>(ab / B);
    out[(static_cast<long long>(b) * A + a) * inner_vec + c] = in[a * ld_a_vec + static_cast<long long>(b) * inner_vec + c];
  }
}
__global__ void add_f32_kernel(float4* dst, const float4* src, long long n4) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n4; i += stride) {
    float4 d = dst[i];
    const float4 s = src[i];
    d.x += s.x; d.y += s.y; d.z += s.z; d.w += s.w;
    dst[i] = d;
  }
}
__global__ void sub_f32_kernel(float4* out, const float4* a, const float4* b, long long n4) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 x = a[i], y = b[i];
    out[i] = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
  }
}
}  // namespace fx

extern "C" int fx_swap01_bf16(const void* in, int64_t ld_a, void* out, int A, int B, int inner, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(in && out && A > 0 && B > 0 && inner > 0, "fx_swap01_bf16: bad arguments");
  FX_CHECK_ARG(inner % 8 == 0 && ld_a % 8 == 0 && ld_a >= static_cast<int64_t>(B) * inner,
               "fx_swap01_bf16: inner and ld_a must be multiples of 8 and ld_a >= B*inner");
  const long long total = static_cast<long long>(A) * B * (inner / 8);
  swap01_kernel<<<ew_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(in), ld_a / 8, reinterpret_cast<uint4*>(out), A, B, inner / 8);
  FX_CHECK_LAUNCH("fx_swap01_bf16");
  return FX_OK;
}
extern "C" int fx_add_f32(float* dst, const float* src, int64_t n, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(dst && src && n > 0 && n % 4 == 0, "fx_add_f32: bad arguments (n must be a multiple of 4)");
  add_f32_kernel<<<ew_grid(n / 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src), n / 4);
  FX_CHECK_LAUNCH("fx_add_f32");
  return FX_OK;
}
extern "C" int fx_sub_f32(float* out, const float* a, const float* b, int64_t n, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(out && a && b && n > 0 && n % 4 == 0, "fx_sub_f32: bad arguments (n must be a multiple of 4)");
  sub_f32_kernel<<<ew_grid(n / 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(out), reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), n / 4);
  FX_CHECK_LAUNCH("fx_sub_f32");
  return FX_OK;
}
