// Front/back-end kernels around the block stack: patch gather / unpatchify, fp32 embedding MLPs on the
// de-duplicated timesteps, and the channel-last helpers of the CNN control fuser. All are small, one-pass,
// HBM/L2-bound gathers; the heavy arithmetic of these stages runs through fx_gemm_bf16.
#include "host_common.h"
#include "ptx.cuh"

namespace fx {

// -------------------------------------------------------------------------------------------------
// patchify: Conv3d k=s=(1,2,2) as a row gather (patch_embedding :624-625,:885; ref_conv :675-678,:896)
// -------------------------------------------------------------------------------------------------
struct PatchSrc {
  const __nv_bfloat16* ptr[4];
  int ch[4];
  int chan_last[4];
  int nsrc;
};

__global__ void patchify_kernel(const PatchSrc src, int ctot, int F, int H, int W, __nv_bfloat16* rows,
                                long long ldr) {
  // one thread per (token, channel, q): writes the (r = 0,1) pair
  const int Hp = H / 2, Wp = W / 2;
  const long long total = static_cast<long long>(F) * Hp * Wp * ctot * 2;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int q = static_cast<int>(i & 1);
    long long t = i >> 1;
    const int c = static_cast<int>(t % ctot);
    const long long tok = t / ctot;
    const int w = static_cast<int>(tok % Wp);
    const int h = static_cast<int>((tok / Wp) % Hp);
    const int f = static_cast<int>(tok / (static_cast<long long>(Wp) * Hp));
    int k = 0, cl = c;
    while (k < src.nsrc - 1 && cl >= src.ch[k]) {
      cl -= src.ch[k];
      ++k;
    }
    const int y = 2 * h + q, x = 2 * w;
    __nv_bfloat16 v0, v1;
    if (src.chan_last[k]) {
      const long long base = ((static_cast<long long>(f) * H + y) * W + x) * src.ch[k] + cl;
      v0 = src.ptr[k][base];
      v1 = src.ptr[k][base + src.ch[k]];
    } else {
      const long long base = ((static_cast<long long>(cl) * F + f) * H + y) * W + x;
      const __nv_bfloat162 pr = *reinterpret_cast<const __nv_bfloat162*>(src.ptr[k] + base);  // x even, W even
      v0 = pr.x;
      v1 = pr.y;
    }
    __nv_bfloat162 o;
    o.x = v0;
    o.y = v1;
    *reinterpret_cast<__nv_bfloat162*>(rows + tok * ldr + (c * 2 + q) * 2) = o;
  }
}

// unpatchify :1126-1149: out[c][f][2h+q][2w+r] = head[tok][(q*2+r)*C + c]
__global__ void unpatchify_kernel(const __nv_bfloat16* head, long long ldh, __nv_bfloat16* out, int C, int F, int H,
                                  int W) {
  const long long total = static_cast<long long>(C) * F * H * W;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const int Hp = H / 2, Wp = W / 2;
  for (; i < total; i += stride) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const int f = static_cast<int>((i / (static_cast<long long>(W) * H)) % F);
    const int c = static_cast<int>(i / (static_cast<long long>(W) * H * F));
    const long long tok = (static_cast<long long>(f) * Hp + (y >> 1)) * Wp + (x >> 1);
    out[i] = head[tok * ldh + ((y & 1) * 2 + (x & 1)) * C + c];
  }
}

// -------------------------------------------------------------------------------------------------
// sinusoidal_embedding_1d :31-41 (float64 math, cos block first)
// -------------------------------------------------------------------------------------------------
__global__ void sinusoid_kernel(const float* t, float* out, int n, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int r = i / half, j = i - r * half;
  const double freq = pow(10000.0, -static_cast<double>(j) / static_cast<double>(half));
  const double a = static_cast<double>(t[r]) * freq;
  out[static_cast<long long>(r) * dim + j] = static_cast<float>(cos(a));
  out[static_cast<long long>(r) * dim + half + j] = static_cast<float>(sin(a));
}

// -------------------------------------------------------------------------------------------------
// Skinny fp32 linear for the embedding MLPs (time/density embedding + projection :630-636): M is the number
// of DISTINCT timesteps (2 in full_edit), so this is a weight-streaming GEMV family: one warp per output
// feature, 4 input rows staged in smem per block, bf16 weights up-cast, fp32 accumulate.
// -------------------------------------------------------------------------------------------------
constexpr int kLinRows = 4;
constexpr int kLinWarps = 8;

__global__ void __launch_bounds__(kLinWarps * 32)
linear_f32_kernel(const float* in, long long ldi, const __nv_bfloat16* w, long long ldw, const __nv_bfloat16* bias,
                  float* out, long long ldo, int M, int N, int K, int act_in) {
  extern __shared__ float s_in[];  // [kLinRows][K]
  const int m0 = blockIdx.y * kLinRows;
  const int rows = min(kLinRows, M - m0);
  for (int i = threadIdx.x; i < kLinRows * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    float v = 0.f;
    if (r < rows) {
      v = in[(m0 + r) * ldi + k];
      if (act_in == 1) v = v / (1.f + expf(-v));  // SiLU
    }
    s_in[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * kLinWarps + warp;
  if (n >= N) return;
  const uint4* wr = reinterpret_cast<const uint4*>(w + n * ldw);
  float acc[kLinRows] = {0.f, 0.f, 0.f, 0.f};
  for (int kv = lane; kv < K / 8; kv += 32) {
    const uint4 wv = __ldg(wr + kv);
    const float wf[8] = {bf16_lo(wv.x), bf16_hi(wv.x), bf16_lo(wv.y), bf16_hi(wv.y),
                         bf16_lo(wv.z), bf16_hi(wv.z), bf16_lo(wv.w), bf16_hi(wv.w)};
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) {
      const float4 a0 = *reinterpret_cast<const float4*>(&s_in[r * K + kv * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&s_in[r * K + kv * 8 + 4]);
      acc[r] += a0.x * wf[0] + a0.y * wf[1] + a0.z * wf[2] + a0.w * wf[3] + a1.x * wf[4] + a1.y * wf[5] +
                a1.z * wf[6] + a1.w * wf[7];
    }
  }
#pragma unroll
  for (int r = 0; r < kLinRows; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
  }
  if (lane == 0) {
    const float b = bias ? __bfloat162float(bias[n]) : 0.f;
    for (int r = 0; r < rows; ++r) out[(m0 + r) * ldo + n] = acc[r] + b;
  }
}

// -------------------------------------------------------------------------------------------------
// CNN control fuser helpers (cnn_conv1..5 :680-711,:868-881), activations channel-last [P, C]
// -------------------------------------------------------------------------------------------------
// dense pixel index p = (f*H + y)*W + x -> row of the zero-padded grid [F, H+2, W+2] (interior position (y+1, x+1)):
// the layout the implicit-GEMM convolution reads (fx_conv_gemm_bf16). pad_H == 0: dense rows.
__device__ __forceinline__ long long padded_row(long long p, int pad_H, int pad_W) {
  if (pad_H == 0) return p;
  const long long hw = static_cast<long long>(pad_H) * pad_W;
  const long long f = p / hw;
  const int r = static_cast<int>(p - f * hw);
  const int y = r / pad_W, x = r - y * pad_W;
  return (f * (pad_H + 2) + y + 1) * (pad_W + 2) + x + 1;
}

__global__ void nchw_to_nhwc_kernel(const __nv_bfloat16* src, __nv_bfloat16* dst, long long ldd, int c0, int C,
                                    long long P, int pad_H, int pad_W) {
  __shared__ __nv_bfloat16 tile[32][33];
  const long long p0 = static_cast<long long>(blockIdx.x) * 32;
  const int cb = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = cb + j;
    const long long pp = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && pp < P) ? src[c * P + pp] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long pp = p0 + j;
    const int c = cb + threadIdx.x;
    if (c < C && pp < P) dst[padded_row(pp, pad_H, pad_W) * ldd + c0 + c] = tile[threadIdx.x][j];
  }
}

// rows[p][(c*3+kh)*3+kw] = in[f, y+kh-1, x+kw-1, c]  (zero outside the frame)
__global__ void im2col3x3_kernel(const __nv_bfloat16* in, __nv_bfloat16* rows, int F, int H, int W, int C) {
  const long long total = static_cast<long long>(F) * H * W * C;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int c = static_cast<int>(i % C);
    const long long pix = i / C;
    const int x = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    __nv_bfloat16* dst = rows + pix * (9LL * C) + c * 9;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int yy = y + kh - 1, xx = x + kw - 1;
        __nv_bfloat16 v = __float2bfloat16(0.f);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = in[(pix + (kh - 1) * W + (kw - 1)) * C + c];
        dst[kh * 3 + kw] = v;
      }
    }
  }
}

// GroupNorm statistics: one block per group, over all pixels of the sample and the group's channels.
__global__ void __launch_bounds__(1024)
groupnorm_stats_kernel(const __nv_bfloat16* x, long long P, int C, int G, float eps, float* stats) {
  const int g = blockIdx.x;
  const int cg = C / G;
  const long long n = P * cg;
  double s = 0.0, q = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const long long pix = i / cg;
    const int ci = static_cast<int>(i - pix * cg);
    const float v = __bfloat162float(x[pix * C + g * cg + ci]);
    s += v;
    q += static_cast<double>(v) * v;
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
  if (warp == 0) {
    s = lane < (blockDim.x >> 5) ? sh_s[lane] : 0.0;
    q = lane < (blockDim.x >> 5) ? sh_q[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      const double mean = s / static_cast<double>(n);
      const double var = q / static_cast<double>(n) - mean * mean;
      stats[2 * g] = static_cast<float>(mean);
      stats[2 * g + 1] = static_cast<float>(1.0 / sqrt((var > 0.0 ? var : 0.0) + static_cast<double>(eps)));
    }
  }
}

__global__ void groupnorm_apply_kernel(const __nv_bfloat16* x, long long P, int C, int G,
                                       const __nv_bfloat16* gamma, const __nv_bfloat16* beta, const float* stats,
                                       const float* resid, float* y_f32, __nv_bfloat16* y_bf16, int pad_H = 0,
                                       int pad_W = 0, long long ld_bf16 = 0) {
  const long long total = P * C;
  const int cg = C / G;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int c = static_cast<int>(i % C);
    const int g = c / cg;
    const float mean = stats[2 * g], rstd = stats[2 * g + 1];
    float v = (__bfloat162float(x[i]) - mean) * rstd * __bfloat162float(gamma[c]) + __bfloat162float(beta[c]);
    v = v / (1.f + expf(-v));  // SiLU
    if (resid != nullptr) v += resid[i];
    if (y_f32 != nullptr) y_f32[i] = v;
    if (y_bf16 != nullptr) {
      if (pad_H == 0) y_bf16[i] = __float2bfloat16_rn(v);
      else y_bf16[padded_row(i / C, pad_H, pad_W) * ld_bf16 + c] = __float2bfloat16_rn(v);
    }
  }
}

// GroupNorm statistics in two deterministic stages, so that frames can live on different ranks (the control fuser is
// sharded by frames across a sequence-parallel group) and still give bit-identical statistics on every layout:
// (1) per (frame, group) partial sum / sum of squares in fp64, one block each — also 25x more blocks than one block per
// group; (2) the partials of ALL frames are summed in frame order (locally computed or gathered from the peers).
__global__ void __launch_bounds__(512)
groupnorm_partial_kernel(const __nv_bfloat16* x, long long pp, int C, int G, double* partials) {
  const int g = blockIdx.x, f = blockIdx.y;
  const int cg = C / G;
  const long long n = pp * cg;
  const __nv_bfloat16* xf = x + static_cast<long long>(f) * pp * C + g * cg;
  double s = 0.0, q = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const long long pix = i / cg;
    const int ci = static_cast<int>(i - pix * cg);
    const float v = __bfloat162float(xf[pix * C + ci]);
    s += v;
    q += static_cast<double>(v) * v;
  }
  __shared__ double sh_s[16], sh_q[16];
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
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < 16; ++w) {
      ts += sh_s[w];
      tq += sh_q[w];
    }
    double* o = partials + (static_cast<long long>(f) * G + g) * 2;
    o[0] = ts;
    o[1] = tq;
  }
}

__global__ void groupnorm_finalize_kernel(const double* partials, int Ft, int G, double n, float eps, float* stats) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double s = 0.0, q = 0.0;
  for (int f = 0; f < Ft; ++f) {
    s += partials[(static_cast<long long>(f) * G + g) * 2];
    q += partials[(static_cast<long long>(f) * G + g) * 2 + 1];
  }
  const double mean = s / n;
  const double var = q / n - mean * mean;
  stats[2 * g] = static_cast<float>(mean);
  stats[2 * g + 1] = static_cast<float>(1.0 / sqrt((var > 0.0 ? var : 0.0) + static_cast<double>(eps)));
}

static int ew_grid2(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace fx

extern "C" int fx_patchify(const void* const* src, const int* channels, const int* chan_last, int nsrc, int F, int H,
                           int W, void* rows, int64_t ldr, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(src && channels && chan_last && rows, "fx_patchify: null pointer");
  FX_CHECK_ARG(nsrc >= 1 && nsrc <= 4, "fx_patchify: nsrc=%d outside [1,4]", nsrc);
  FX_CHECK_ARG(F > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "fx_patchify: bad grid %dx%dx%d", F, H, W);
  PatchSrc ps{};
  int ctot = 0;
  for (int k = 0; k < nsrc; ++k) {
    FX_CHECK_ARG(src[k] != nullptr && channels[k] > 0, "fx_patchify: bad source %d", k);
    ps.ptr[k] = reinterpret_cast<const __nv_bfloat16*>(src[k]);
    ps.ch[k] = channels[k];
    ps.chan_last[k] = chan_last[k];
    ctot += channels[k];
  }
  ps.nsrc = nsrc;
  FX_CHECK_ARG(ldr >= 4LL * ctot && ldr % 2 == 0, "fx_patchify: ldr too small");
  const long long total = static_cast<long long>(F) * (H / 2) * (W / 2) * ctot * 2;
  patchify_kernel<<<ew_grid2(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ps, ctot, F, H, W, reinterpret_cast<__nv_bfloat16*>(rows), ldr);
  FX_CHECK_LAUNCH("fx_patchify");
  return FX_OK;
}

extern "C" int fx_unpatchify(const void* head, int64_t ldh, void* out, int C, int F, int H, int W, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(head && out, "fx_unpatchify: null pointer");
  FX_CHECK_ARG(C > 0 && F > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && ldh >= 4LL * C,
               "fx_unpatchify: bad shape");
  const long long total = static_cast<long long>(C) * F * H * W;
  unpatchify_kernel<<<ew_grid2(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(head), ldh, reinterpret_cast<__nv_bfloat16*>(out), C, F, H, W);
  FX_CHECK_LAUNCH("fx_unpatchify");
  return FX_OK;
}

extern "C" int fx_sinusoid(const float* t, float* out, int n, int dim, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(t && out && n > 0 && dim > 0 && dim % 2 == 0, "fx_sinusoid: bad arguments");
  const int total = n * (dim / 2);
  sinusoid_kernel<<<(total + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, out, n, dim);
  FX_CHECK_LAUNCH("fx_sinusoid");
  return FX_OK;
}

extern "C" int fx_linear_f32(const float* in, int64_t ldi, const void* w, int64_t ldw, const void* bias, float* out,
                             int64_t ldo, int M, int N, int K, int act_in, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(in && w && out, "fx_linear_f32: null pointer");
  FX_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % 8 == 0 && ldw % 8 == 0, "fx_linear_f32: bad shape M=%d N=%d K=%d", M, N,
               K);
  FX_CHECK_ARG(act_in == 0 || act_in == 1, "fx_linear_f32: unknown act_in %d", act_in);
  const int smem = kLinRows * K * static_cast<int>(sizeof(float));
  FX_CHECK_ARG(smem <= 200 * 1024, "fx_linear_f32: K=%d too large for the smem staging buffer", K);
  if (smem > 48 * 1024 && !ensure_dyn_smem(reinterpret_cast<const void*>(linear_f32_kernel), smem, "fx_linear_f32"))
    return FX_ERR_CUDA;
  dim3 grid((N + kLinWarps - 1) / kLinWarps, (M + kLinRows - 1) / kLinRows);
  FX_CHECK_ARG(grid.y <= 65535, "fx_linear_f32: M=%d too large", M);
  linear_f32_kernel<<<grid, kLinWarps * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, ldi, reinterpret_cast<const __nv_bfloat16*>(w), ldw, reinterpret_cast<const __nv_bfloat16*>(bias), out, ldo, M,
      N, K, act_in);
  FX_CHECK_LAUNCH("fx_linear_f32");
  return FX_OK;
}

extern "C" int fx_nchw_to_nhwc(const void* src, void* dst, int64_t ldd, int c0, int C, int64_t P, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(src && dst && C > 0 && P > 0 && c0 >= 0 && ldd >= c0 + C, "fx_nchw_to_nhwc: bad arguments");
  dim3 grid(static_cast<unsigned>((P + 31) / 32), (C + 31) / 32);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<__nv_bfloat16*>(dst), ldd, c0, C, P, 0, 0);
  FX_CHECK_LAUNCH("fx_nchw_to_nhwc");
  return FX_OK;
}

extern "C" int fx_nchw_to_nhwc_padded(const void* src, void* dst, int64_t ldd, int c0, int C, int F, int H, int W,
                                      void* stream) {
  using namespace fx;
  FX_CHECK_ARG(src && dst && C > 0 && F > 0 && H > 0 && W > 0 && c0 >= 0 && ldd >= c0 + C,
               "fx_nchw_to_nhwc_padded: bad arguments");
  const long long P = static_cast<long long>(F) * H * W;
  dim3 grid(static_cast<unsigned>((P + 31) / 32), (C + 31) / 32);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<__nv_bfloat16*>(dst), ldd, c0, C, P, H, W);
  FX_CHECK_LAUNCH("fx_nchw_to_nhwc_padded");
  return FX_OK;
}

extern "C" int fx_im2col3x3(const void* in, void* rows, int F, int H, int W, int C, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(in && rows && F > 0 && H > 0 && W > 0 && C > 0, "fx_im2col3x3: bad arguments");
  const long long total = static_cast<long long>(F) * H * W * C;
  im2col3x3_kernel<<<ew_grid2(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(rows), F, H, W, C);
  FX_CHECK_LAUNCH("fx_im2col3x3");
  return FX_OK;
}

extern "C" int fx_groupnorm_silu(const void* x, int64_t P, int C, int G, float eps, const void* gamma,
                                 const void* beta, const float* resid, float* y_f32, void* y_bf16, float* stats,
                                 void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && gamma && beta && stats && (y_f32 || y_bf16), "fx_groupnorm_silu: null pointer");
  FX_CHECK_ARG(P > 0 && C > 0 && G > 0 && C % G == 0, "fx_groupnorm_silu: bad shape P=%lld C=%d G=%d", (long long)P, C,
               G);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  groupnorm_stats_kernel<<<G, 1024, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), P, C, G, eps, stats);
  FX_CHECK_LAUNCH("fx_groupnorm_silu(stats)");
  groupnorm_apply_kernel<<<ew_grid2(P * C), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), P, C, G, reinterpret_cast<const __nv_bfloat16*>(gamma),
      reinterpret_cast<const __nv_bfloat16*>(beta), stats, resid, y_f32, reinterpret_cast<__nv_bfloat16*>(y_bf16));
  FX_CHECK_LAUNCH("fx_groupnorm_silu(apply)");
  return FX_OK;
}

extern "C" int fx_groupnorm_partials(const void* x, int F, int64_t pp, int C, int G, double* partials, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && partials, "fx_groupnorm_partials: null pointer");
  FX_CHECK_ARG(F > 0 && F <= 65535 && pp > 0 && C > 0 && G > 0 && C % G == 0,
               "fx_groupnorm_partials: bad shape F=%d pp=%lld C=%d G=%d", F, (long long)pp, C, G);
  groupnorm_partial_kernel<<<dim3(G, F), 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), pp, C, G, partials);
  FX_CHECK_LAUNCH("fx_groupnorm_partials");
  return FX_OK;
}

extern "C" int fx_groupnorm_silu_partials(const void* x, int64_t P, int C, int G, float eps, const void* gamma,
                                          const void* beta, const double* partials, int Ft, int64_t pp,
                                          const float* resid, float* y_f32, void* y_bf16, int pad_H, int pad_W,
                                          int64_t ld_bf16, float* stats, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && gamma && beta && partials && stats && (y_f32 || y_bf16), "fx_groupnorm_silu_partials: null pointer");
  FX_CHECK_ARG(pad_H == 0 || (pad_H > 0 && pad_W > 0 && ld_bf16 >= C && P % (static_cast<int64_t>(pad_H) * pad_W) == 0),
               "fx_groupnorm_silu_partials: bad padded-output arguments");
  FX_CHECK_ARG(P > 0 && C > 0 && G > 0 && C % G == 0 && Ft > 0 && pp > 0,
               "fx_groupnorm_silu_partials: bad shape P=%lld C=%d G=%d Ft=%d pp=%lld", (long long)P, C, G, Ft,
               (long long)pp);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const double n = static_cast<double>(Ft) * static_cast<double>(pp) * (C / G);
  groupnorm_finalize_kernel<<<(G + 31) / 32, 32, 0, s>>>(partials, Ft, G, n, eps, stats);
  FX_CHECK_LAUNCH("fx_groupnorm_silu_partials(finalize)");
  groupnorm_apply_kernel<<<ew_grid2(P * C), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), P, C, G, reinterpret_cast<const __nv_bfloat16*>(gamma),
      reinterpret_cast<const __nv_bfloat16*>(beta), stats, resid, y_f32, reinterpret_cast<__nv_bfloat16*>(y_bf16), pad_H,
      pad_W, ld_bf16);
  FX_CHECK_LAUNCH("fx_groupnorm_silu_partials(apply)");
  return FX_OK;
}

// -------------------------------------------------------------------------------------------------
// On-device de-duplication of the per-token timesteps (replaces a host-synchronising torch.unique): the pipeline's
// per-token t is mask * t (pipeline :891-898), i.e. a handful of distinct values in `full_edit` and up to one per token
// with fractional trilinear masks (:686-690). One block: distinct values are appended in order of first appearance
// (deterministic), at most `cap`; count = cap + 1 reports "more than cap" (the caller then takes the per-token path).
// -------------------------------------------------------------------------------------------------
namespace fx {
constexpr int kDedupThreads = 1024;
constexpr int kDedupCap = 64;

__global__ void __launch_bounds__(kDedupThreads)
dedup_f32_kernel(const float* t, int n, int cap, float* uniq, int* inv, int* count) {
  __shared__ float vals[kDedupCap];
  __shared__ int s_cnt, s_best;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  while (true) {
    const int cnt = s_cnt;
    if (threadIdx.x == 0) s_best = n;
    __syncthreads();
    int cand = n;  // first element owned by this thread whose value is not in the list yet
    for (int i = threadIdx.x; i < n && cand == n; i += kDedupThreads) {
      const float v = t[i];
      bool found = false;
      for (int j = 0; j < cnt; ++j) found |= (vals[j] == v);
      if (!found) cand = i;
    }
    if (cand < n) atomicMin(&s_best, cand);
    __syncthreads();
    const int best = s_best;
    if (best == n) break;              // every value is in the list
    if (cnt == cap) {                  // a (cap+1)-th distinct value exists
      if (threadIdx.x == 0) s_cnt = cap + 1;
      __syncthreads();
      break;
    }
    if (threadIdx.x == 0) {
      vals[cnt] = t[best];
      s_cnt = cnt + 1;
    }
    __syncthreads();
  }
  const int cnt = s_cnt;
  const int listed = cnt > cap ? cap : cnt;
  if (threadIdx.x == 0) *count = cnt;
  for (int j = threadIdx.x; j < cap; j += kDedupThreads) uniq[j] = vals[j < listed ? j : (listed > 0 ? listed - 1 : 0)];
  for (int i = threadIdx.x; i < n; i += kDedupThreads) {
    const float v = t[i];
    int idx = 0;
    for (int j = 0; j < listed; ++j)
      if (vals[j] == v) idx = j;
    inv[i] = idx;
  }
}
}  // namespace fx

extern "C" int fx_dedup_f32(const float* t, int n, int cap, float* uniq, int32_t* inv, int32_t* count, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(t && uniq && inv && count && n > 0, "fx_dedup_f32: bad arguments");
  FX_CHECK_ARG(cap >= 1 && cap <= kDedupCap, "fx_dedup_f32: cap %d outside [1, %d]", cap, kDedupCap);
  dedup_f32_kernel<<<1, kDedupThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, n, cap, uniq, inv, count);
  FX_CHECK_LAUNCH("fx_dedup_f32");
  return FX_OK;
}
