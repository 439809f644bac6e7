// fp32 verification mode ("precise" engine, flexam_b200/precise.py): the same denoising step with fp32 activations,
// so the native path can be checked against the reference's fp32 run at <= 1e-4 relative L2 (BASELINE north star).
//
// The contractions still run on tcgen05: an fp32 activation matrix is split EXACTLY into three bf16 planes
// (a = hi + mid + lo, 8 + 8 + 8 significant bits), each plane goes through fx_gemm_bf16 with the exact fp32 epilogue
// (FX_EPI_F32_EXACT: acc + bias, no bf16 rounding) and the three fp32 results are added (weights are bf16 parameters,
// so W needs no split). Because the split is exact, the bf16 gather kernels (patchify, im2col, unpatchify) move fp32
// data losslessly plane by plane. Everything else here is plain fp32 SIMT code written for clarity, not speed:
// LayerNorm/modulation, RMSNorm+RoPE, GELU, the gated residual, and a flash-style fp32 attention.
//
// Reference lines restated (FlexAM/models/wan_transformer3d_FlexAM.py): LayerNorm+modulation :444-453,:464-465,:493-507;
// WanRMSNorm :173-189; rope_apply :135-164; attention (attention_utils.py:174-233); ffn GELU(tanh) :415; gated
// residuals :456,:468.
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace fx {

static int pr_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protects `red` against the previous reduction's readers
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

// ---- exact 3-way bf16 split / join ----------------------------------------------------------------------
__global__ void split3_kernel(const float* in, long long ldi, int M, int K, __nv_bfloat16* out) {
  const long long total = static_cast<long long>(M) * K;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const long long m = i / K;
    const float a = in[m * ldi + (i - m * K)];
    const __nv_bfloat16 hi = __float2bfloat16_rn(a);
    const float r1 = a - __bfloat162float(hi);  // exact: at most 16 significant bits remain
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(mid);  // exact: at most 8 significant bits remain
    out[i] = hi;
    out[total + i] = mid;
    out[2 * total + i] = __float2bfloat16_rn(r2);
  }
}

__global__ void join3_kernel(const __nv_bfloat16* planes, long long n, float* out) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride)
    out[i] = (__bfloat162float(planes[2 * n + i]) + __bfloat162float(planes[n + i])) + __bfloat162float(planes[i]);
}

// ---- LayerNorm + modulation / affine, fp32 out ----------------------------------------------------------
struct LnF32Params {
  const float* x;
  float* out;
  int M, D;
  float eps;
  const float *shift_mod, *scale_mod, *shift_e, *scale_e;
  long long e_stride;
  const int* row_idx;
  const float *dens_mod, *dens;
  long long dens_stride;
  int rows_per_batch;
  const __nv_bfloat16 *gamma, *beta;
};

__global__ void __launch_bounds__(256) ln_f32_kernel(const LnF32Params p) {
  __shared__ float red[8];
  const int row = blockIdx.x;
  const float* xr = p.x + static_cast<long long>(row) * p.D;
  float s = 0.f;
  for (int c = threadIdx.x; c < p.D; c += 256) s += xr[c];
  const float mean = block_sum_256(s, red) / static_cast<float>(p.D);
  float q = 0.f;
  for (int c = threadIdx.x; c < p.D; c += 256) {
    const float d = xr[c] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(block_sum_256(q, red) / static_cast<float>(p.D) + p.eps);
  float* orow = p.out + static_cast<long long>(row) * p.D;
  if (p.gamma != nullptr) {
    for (int c = threadIdx.x; c < p.D; c += 256)
      orow[c] = (xr[c] - mean) * rstd * __bfloat162float(p.gamma[c]) + __bfloat162float(p.beta[c]);
    return;
  }
  const long long u = p.row_idx ? p.row_idx[row] : 0;
  const float* se = p.scale_e + u * p.e_stride;
  const float* he = p.shift_e + u * p.e_stride;
  const float* dn = p.dens ? p.dens + static_cast<long long>(row / p.rows_per_batch) * p.dens_stride : nullptr;
  for (int c = threadIdx.x; c < p.D; c += 256) {
    float y = (xr[c] - mean) * rstd * (1.f + (p.scale_mod[c] + se[c])) + (p.shift_mod[c] + he[c]);
    if (dn != nullptr) y += (p.dens_mod ? p.dens_mod[c] : 0.f) + dn[c];
    orow[c] = y;
  }
}

// ---- RMSNorm (+ RoPE), fp32 in place ----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rmsnorm_rope_f32_kernel(float* x, long long ldx, int M, int D, int ntensors, float eps, const __nv_bfloat16* w,
                        const __nv_bfloat16* w2, const float2* freqs, int gf, int gh, int gw, int tok_offset,
                        int rows_per_batch) {
  __shared__ float red[8];
  const int row = blockIdx.x / ntensors;
  const int which = blockIdx.x - row * ntensors;
  float* xr = x + static_cast<long long>(row) * ldx + static_cast<long long>(which) * D;
  const __nv_bfloat16* wr = which == 0 ? w : w2;
  float ss = 0.f;
  for (int c = threadIdx.x; c < D; c += 256) ss += xr[c] * xr[c];
  const float r = rsqrtf(block_sum_256(ss, red) / static_cast<float>(D) + eps);
  bool rotate = false;
  int pf = 0, ph = 0, pw = 0;
  if (freqs != nullptr) {
    const int t = tok_offset + row % rows_per_batch;
    if (t < gf * gh * gw) {
      rotate = true;
      pf = t / (gh * gw);
      const int rem = t - pf * (gh * gw);
      ph = rem / gw;
      pw = rem - ph * gw;
    }
  }
  for (int pi = threadIdx.x; pi < D / 2; pi += 256) {
    const int c = 2 * pi;
    float a = xr[c] * r * __bfloat162float(wr[c]);
    float b = xr[c + 1] * r * __bfloat162float(wr[c + 1]);
    if (rotate) {
      const int pj = pi & 63;  // complex pair inside the 128-wide head
      const int pos = pj < 22 ? pf : (pj < 43 ? ph : pw);
      const float2 cs = freqs[pos * 64 + pj];
      const float re = a * cs.x - b * cs.y;
      const float im = a * cs.y + b * cs.x;
      a = re;
      b = im;
    }
    xr[c] = a;
    xr[c + 1] = b;
  }
}

// ---- GELU(tanh) in place, gated residual -------------------------------------------------------------------
__global__ void gelu_f32_kernel(float* x, long long n) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float v = x[i];
    x[i] = 0.5f * v * (1.f + tanhf(0.7978845608028654f * (v + 0.044715f * v * v * v)));
  }
}

__global__ void gated_residual_f32_kernel(float* x, const float* y, int M, int N, const float* gate_mod,
                                          const float* gate_e, long long gate_e_stride, const int* row_idx) {
  const long long total = static_cast<long long>(M) * N;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const long long m = i / N;
    const int n = static_cast<int>(i - m * N);
    float g = 1.f;
    if (gate_mod != nullptr || gate_e != nullptr) {
      g = gate_mod ? gate_mod[n] : 0.f;
      if (gate_e != nullptr) g += gate_e[(row_idx ? row_idx[m] : 0) * gate_e_stride + n];
    }
    x[i] = x[i] + y[i] * g;
  }
}

// ---- fp32 attention: 16 queries per block (2 per warp), 32-key tiles in shared memory, online softmax --------
constexpr int kAQ = 16, kAK = 32, kHD = 128;

__global__ void __launch_bounds__(256)
attention_f32_kernel(const float* q, long long qsb, long long qsl, const float* k, long long ksb, long long ksl,
                     const float* v, long long vsb, long long vsl, float* o, long long osb, long long osl, int Lq,
                     int Lk, float scale) {
  __shared__ float Qs[kAQ][kHD];
  __shared__ float Ks[kAK][kHD + 1];
  __shared__ float Vs[kAK][kHD];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* qb = q + b * qsb + h * kHD;
  const float* kb = k + b * ksb + h * kHD;
  const float* vb = v + b * vsb + h * kHD;
  for (int i = threadIdx.x; i < kAQ * kHD; i += 256) {
    const int r = i / kHD, d = i - r * kHD;
    Qs[r][d] = (q0 + r < Lq) ? qb[static_cast<long long>(q0 + r) * qsl + d] : 0.f;
  }
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int k0 = 0; k0 < Lk; k0 += kAK) {
    __syncthreads();
    for (int i = threadIdx.x; i < kAK * kHD; i += 256) {
      const int r = i / kHD, d = i - r * kHD;
      const bool ok = k0 + r < Lk;
      Ks[r][d] = ok ? kb[static_cast<long long>(k0 + r) * ksl + d] : 0.f;
      Vs[r][d] = ok ? vb[static_cast<long long>(k0 + r) * vsl + d] : 0.f;
    }
    __syncthreads();
    const bool key_ok = k0 + lane < Lk;
#pragma unroll
    for (int qi = 0; qi < 2; ++qi) {
      const float* qrow = Qs[warp * 2 + qi];
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < kHD; ++d) s += qrow[d] * Ks[lane][d];
      s = key_ok ? s * scale : -INFINITY;
      float mx = s;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m[qi], mx);  // finite: every tile holds at least one valid key
      const float pexp = key_ok ? expf(s - m_new) : 0.f;
      const float corr = expf(m[qi] - m_new);
      float ps = pexp;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l[qi] = l[qi] * corr + ps;
      m[qi] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[qi][j] *= corr;
      for (int kk = 0; kk < kAK; ++kk) {
        const float pk = __shfl_sync(0xffffffffu, pexp, kk);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[qi][j] += pk * Vs[kk][lane + 32 * j];
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < 2; ++qi) {
    const int qr = q0 + warp * 2 + qi;
    if (qr < Lq) {
      float* orow = o + b * osb + static_cast<long long>(qr) * osl + h * kHD;
#pragma unroll
      for (int j = 0; j < 4; ++j) orow[lane + 32 * j] = acc[qi][j] / l[qi];
    }
  }
}

// ---- GroupNorm + SiLU on fp32 conv outputs (cnn_conv1..4 :680-711) ------------------------------------------
__global__ void __launch_bounds__(1024)
groupnorm_stats_f32_kernel(const float* x, long long P, int C, int G, float eps, float* stats) {
  const int g = blockIdx.x;
  const int cg = C / G;
  const long long n = P * cg;
  double s = 0.0, q = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const long long pix = i / cg;
    const double v = x[pix * C + g * cg + static_cast<int>(i - pix * cg)];
    s += v;
    q += v * v;
  }
  __shared__ double sh_s[32], sh_q[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh_s[warp] = s;
    sh_q[warp] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0.0;
    q = 0.0;
    for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) {
      s += sh_s[i];
      q += sh_q[i];
    }
    const double mean = s / static_cast<double>(n);
    const double var = q / static_cast<double>(n) - mean * mean;
    stats[2 * g] = static_cast<float>(mean);
    stats[2 * g + 1] = static_cast<float>(1.0 / sqrt((var > 0.0 ? var : 0.0) + static_cast<double>(eps)));
  }
}

__global__ void groupnorm_apply_f32_kernel(const float* x, long long P, int C, int G, const __nv_bfloat16* gamma,
                                           const __nv_bfloat16* beta, const float* stats, const float* resid,
                                           float* y) {
  const long long total = P * C;
  const int cg = C / G;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int c = static_cast<int>(i % C);
    const int g = c / cg;
    float v = (x[i] - stats[2 * g]) * stats[2 * g + 1] * __bfloat162float(gamma[c]) + __bfloat162float(beta[c]);
    v = v / (1.f + expf(-v));
    if (resid != nullptr) v += resid[i];
    y[i] = v;
  }
}

}  // namespace fx

extern "C" int fx_split3_f32(const float* in, int64_t ldi, int M, int K, void* planes, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(in && planes && M > 0 && K > 0 && ldi >= K, "fx_split3_f32: bad arguments");
  split3_kernel<<<pr_grid(static_cast<long long>(M) * K), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, ldi, M, K, reinterpret_cast<__nv_bfloat16*>(planes));
  FX_CHECK_LAUNCH("fx_split3_f32");
  return FX_OK;
}

extern "C" int fx_join3_f32(const void* planes, int64_t n, float* out, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(planes && out && n > 0, "fx_join3_f32: bad arguments");
  join3_kernel<<<pr_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(planes), n, out);
  FX_CHECK_LAUNCH("fx_join3_f32");
  return FX_OK;
}

extern "C" int fx_ln_f32(const float* x, float* out, int M, int D, float eps, const float* shift_mod,
                         const float* scale_mod, const float* shift_e, const float* scale_e, int64_t e_stride,
                         const int32_t* row_idx, const float* dens_mod, const float* dens, int64_t dens_stride,
                         int rows_per_batch, const void* gamma, const void* beta, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && out && M > 0 && D > 0, "fx_ln_f32: bad arguments");
  FX_CHECK_ARG((gamma && beta) || (shift_mod && scale_mod && shift_e && scale_e), "fx_ln_f32: neither affine nor modulation given");
  FX_CHECK_ARG(rows_per_batch > 0 || dens == nullptr, "fx_ln_f32: rows_per_batch must be positive with a density term");
  LnF32Params p{};
  p.x = x; p.out = out; p.M = M; p.D = D; p.eps = eps;
  p.shift_mod = shift_mod; p.scale_mod = scale_mod; p.shift_e = shift_e; p.scale_e = scale_e; p.e_stride = e_stride;
  p.row_idx = row_idx; p.dens_mod = dens_mod; p.dens = dens; p.dens_stride = dens_stride;
  p.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : 1;
  p.gamma = reinterpret_cast<const __nv_bfloat16*>(gamma);
  p.beta = reinterpret_cast<const __nv_bfloat16*>(beta);
  ln_f32_kernel<<<M, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  FX_CHECK_LAUNCH("fx_ln_f32");
  return FX_OK;
}

extern "C" int fx_rmsnorm_rope_f32(float* x, int64_t ldx, int M, int D, float eps, const void* weight,
                                   const void* weight2, const float* freqs, int gf, int gh, int gw, int tok_offset,
                                   int rows_per_batch, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && weight && M > 0 && D > 0 && D % 128 == 0, "fx_rmsnorm_rope_f32: bad arguments");
  const int nt = weight2 ? 2 : 1;
  FX_CHECK_ARG(ldx >= static_cast<int64_t>(nt) * D, "fx_rmsnorm_rope_f32: ldx too small");
  if (freqs != nullptr)
    FX_CHECK_ARG(gf > 0 && gh > 0 && gw > 0 && gf <= 1024 && gh <= 1024 && gw <= 1024 && rows_per_batch > 0,
                 "fx_rmsnorm_rope_f32: grid (%d,%d,%d) outside the 1024-entry RoPE table", gf, gh, gw);
  rmsnorm_rope_f32_kernel<<<M * nt, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, ldx, M, D, nt, eps, reinterpret_cast<const __nv_bfloat16*>(weight),
      reinterpret_cast<const __nv_bfloat16*>(weight2), reinterpret_cast<const float2*>(freqs), gf, gh, gw, tok_offset,
      rows_per_batch > 0 ? rows_per_batch : 1);
  FX_CHECK_LAUNCH("fx_rmsnorm_rope_f32");
  return FX_OK;
}

extern "C" int fx_gelu_f32(float* x, int64_t n, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && n > 0, "fx_gelu_f32: bad arguments");
  gelu_f32_kernel<<<pr_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n);
  FX_CHECK_LAUNCH("fx_gelu_f32");
  return FX_OK;
}

extern "C" int fx_gated_residual_f32(float* x, const float* y, int M, int N, const float* gate_mod,
                                     const float* gate_e, int64_t gate_e_stride, const int32_t* row_idx,
                                     void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && y && M > 0 && N > 0, "fx_gated_residual_f32: bad arguments");
  gated_residual_f32_kernel<<<pr_grid(static_cast<long long>(M) * N), 256, 0,
                              reinterpret_cast<cudaStream_t>(stream)>>>(x, y, M, N, gate_mod, gate_e, gate_e_stride,
                                                                        row_idx);
  FX_CHECK_LAUNCH("fx_gated_residual_f32");
  return FX_OK;
}

extern "C" int fx_attention_f32(const float* q, int64_t q_stride_b, int64_t q_stride_l, const float* k,
                                int64_t k_stride_b, int64_t k_stride_l, const float* v, int64_t v_stride_b,
                                int64_t v_stride_l, float* o, int64_t o_stride_b, int64_t o_stride_l, int B, int H,
                                int Lq, int Lk, float scale, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(q && k && v && o, "fx_attention_f32: null pointer");
  FX_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0 && B <= 65535 && H <= 65535, "fx_attention_f32: bad shape");
  dim3 grid((Lq + kAQ - 1) / kAQ, H, B);
  attention_f32_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      q, q_stride_b, q_stride_l, k, k_stride_b, k_stride_l, v, v_stride_b, v_stride_l, o, o_stride_b, o_stride_l, Lq,
      Lk, scale);
  FX_CHECK_LAUNCH("fx_attention_f32");
  return FX_OK;
}

extern "C" int fx_groupnorm_silu_f32(const float* x, int64_t P, int C, int G, float eps, const void* gamma,
                                     const void* beta, const float* resid, float* y, float* stats, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && gamma && beta && stats && y, "fx_groupnorm_silu_f32: null pointer");
  FX_CHECK_ARG(P > 0 && C > 0 && G > 0 && C % G == 0, "fx_groupnorm_silu_f32: bad shape");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  groupnorm_stats_f32_kernel<<<G, 1024, 0, s>>>(x, P, C, G, eps, stats);
  FX_CHECK_LAUNCH("fx_groupnorm_silu_f32(stats)");
  groupnorm_apply_f32_kernel<<<pr_grid(P * C), 256, 0, s>>>(x, P, C, G, reinterpret_cast<const __nv_bfloat16*>(gamma),
                                                            reinterpret_cast<const __nv_bfloat16*>(beta), stats, resid, y);
  FX_CHECK_LAUNCH("fx_groupnorm_silu_f32(apply)");
  return FX_OK;
}
