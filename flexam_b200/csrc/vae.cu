// Row / layout kernels of the Wan2.2 VAE decoder (decode half; reference
// FlexAM/models/wan_vae3_8.py, cited as :line). Activations are channel-last bf16 [frames, H, W, C]; every convolution
// (CausalConv3d 3x3x3 :22-47, the per-frame 3x3 of Resample :93-97, the (3,1,1) time convolution :98-99, 1x1 shortcuts)
// runs as an implicit GEMM on the tcgen05 kernels (fx_conv_gemm_bf16 / fx_gemm_bf16). These kernels produce the
// zero-padded grids those convolutions read and do the element-wise work between them:
//   RMS_norm (+ SiLU) :50-64,:205-211   nearest 2x up-sampling :67-73   temporal interleave :139-142
//   DupUp3D shortcut + add :395-417,:497-500   attention softmax :260-282   unpatchify + clamp :304-318,:1043
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace fx {

static int vae_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

// row of the padded grid [frames, H+2p, W+2p] for dense pixel index pix = (f*H + y)*W + x, frame offset f0
__device__ __forceinline__ long long grid_row(long long pix, int H, int W, int pad, int f0) {
  const long long hw = static_cast<long long>(H) * W;
  const long long f = pix / hw;
  const int r = static_cast<int>(pix - f * hw);
  const int y = r / W, x = r - y * W;
  return ((f + f0) * (H + 2 * pad) + y + pad) * (W + 2 * pad) + x + pad;
}

// out[grid_row(pix)][:C] = act(x[pix][:C] / max(||x[pix]||, 1e-12) * sqrt(C) * gamma)   (F.normalize over channels :62-64)
// One warp per pixel, 16-byte vectors; gamma == nullptr: plain copy (decoder.conv1 input, time_conv input).
__global__ void __launch_bounds__(256)
vae_norm_act_kernel(const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* gamma, __nv_bfloat16* out,
                    long long ldo, long long npix, int C, int H, int W, int pad, int f0, int silu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long stride = static_cast<long long>(gridDim.x) * 8;
  for (long long pix = static_cast<long long>(blockIdx.x) * 8 + warp; pix < npix; pix += stride) {
    const __nv_bfloat16* xr = x + pix * ldx;
    __nv_bfloat16* orow = out + grid_row(pix, H, W, pad, f0) * ldo;
    float scale = 1.f;
    if (gamma != nullptr) {
      float ss = 0.f;
      for (int c = lane * 8; c < C; c += 256) {
        const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
          ss += a * a + b * b;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      scale = sqrtf(static_cast<float>(C)) / fmaxf(sqrtf(ss), 1e-12f);
    }
    for (int c = lane * 8; c < C; c += 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
      uint4 o = v;
      if (gamma != nullptr) {
        const uint4 g = __ldg(reinterpret_cast<const uint4*>(gamma + c));
        const uint32_t u[4] = {v.x, v.y, v.z, v.w}, gg[4] = {g.x, g.y, g.z, g.w};
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = bf16_round(bf16_lo(u[j]) * scale) * bf16_lo(gg[j]);     // normalize * sqrt(C), then * gamma
          float b = bf16_round(bf16_hi(u[j]) * scale) * bf16_hi(gg[j]);
          if (silu) {
            a = bf16_round(a);
            b = bf16_round(b);
            a = a / (1.f + __expf(-a));
            b = b / (1.f + __expf(-b));
          }
          r[j] = pack_bf16x2(a, b);
        }
        o = make_uint4(r[0], r[1], r[2], r[3]);
      }
      *reinterpret_cast<uint4*>(orow + c) = o;
    }
  }
}

// nearest-exact 2x spatial up-sampling of [F, H, W, C] into the padded grid [F, 2H+2, 2W+2, C] (:67-73, :93-95)
__global__ void vae_upsample2x_kernel(const uint4* x, uint4* out, long long nout, int H, int W, int cvec) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const int H2 = 2 * H, W2 = 2 * W;
  for (; i < nout; i += stride) {
    const int c = static_cast<int>(i % cvec);
    const long long pix = i / cvec;
    const int xo = static_cast<int>(pix % W2);
    const int yo = static_cast<int>((pix / W2) % H2);
    const long long f = pix / (static_cast<long long>(W2) * H2);
    const uint4 v = x[((f * H + (yo >> 1)) * W + (xo >> 1)) * cvec + c];
    out[((f * (H2 + 2) + yo + 1) * (W2 + 2) + xo + 1) * cvec + c] = v;
  }
}

// temporal interleave after the time convolution (:139-142): y [T, P, 2C] -> x [2T, P, C], x[2t + k][p][c] = y[t][p][kC + c]
__global__ void vae_time_interleave_kernel(const uint4* y, uint4* x, long long nvec, long long P, int cvec) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < nvec; i += stride) {            // i over the output [2T, P, cvec]
    const int c = static_cast<int>(i % cvec);
    const long long r = i / cvec;
    const long long p = r % P;
    const long long tf = r / P;
    x[i] = y[((tf >> 1) * P + p) * (2 * cvec) + (tf & 1) * cvec + c];
  }
}

// main[to][yo][xo][co] += x[(to + drop) / ft][yo / 2][xo / 2][(co * factor + a*4 + b*2 + d) / rep]   (DupUp3D :395-417)
// with a = (to + drop) % ft, b = yo % 2, d = xo % 2, factor = 4 ft, rep = Cout * factor / Cin; bf16 add (:500).
// One thread per 8 consecutive output channels (16-byte read-modify-write of main, the pixel decoded once per vector).
__global__ void __launch_bounds__(256)
vae_dupup_add_kernel(__nv_bfloat16* main, const __nv_bfloat16* x, long long nvec, int H, int W, int Cin, int Cout, int ft,
                     int drop) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const int H2 = 2 * H, W2 = 2 * W, factor = 4 * ft, cvec = Cout / 8;
  const int rep = Cout * factor / Cin;
  for (; i < nvec; i += stride) {
    const int co0 = static_cast<int>(i % cvec) * 8;
    const long long pix = i / cvec;
    const int xo = static_cast<int>(pix % W2);
    const long long rest = pix / W2;
    const int yo = static_cast<int>(rest % H2);
    const long long to = rest / H2 + drop;
    const int off = static_cast<int>(to % ft) * 4 + (yo & 1) * 2 + (xo & 1);
    const __nv_bfloat16* xr = x + (((to / ft) * H + (yo >> 1)) * W + (xo >> 1)) * Cin;
    uint4* mp = reinterpret_cast<uint4*>(main + pix * Cout + co0);
    uint4 m = *mp;
    uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(mw[j]) + __bfloat162float(xr[((co0 + 2 * j) * factor + off) / rep]);
      const float b = bf16_hi(mw[j]) + __bfloat162float(xr[((co0 + 2 * j + 1) * factor + off) / rep]);
      mw[j] = pack_bf16x2(a, b);
    }
    *mp = make_uint4(mw[0], mw[1], mw[2], mw[3]);
  }
}

// p[r][:] = bf16(softmax(s[r][:] * scale)); one warp per row (AttentionBlock's scaled_dot_product_attention :272-276)
__global__ void __launch_bounds__(256)
vae_softmax_kernel(const float* s, long long lds, __nv_bfloat16* p, long long ldp, int rows, int cols, float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const float* sr = s + static_cast<long long>(r) * lds;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, sr[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += __expf((sr[c] - mx) * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  __nv_bfloat16* pr = p + static_cast<long long>(r) * ldp;
  for (int c = lane; c < cols; c += 32) pr[c] = __float2bfloat16_rn(__expf((sr[c] - mx) * scale) * inv);
}

// video[c][f0 + f][2h + q][2w + r] = clamp(y[f][h][w][c*4 + r*2 + q], -1, 1)   (unpatchify :304-318, clamp :1043)
__global__ void vae_unpatchify_kernel(const __nv_bfloat16* y, long long ldy, __nv_bfloat16* video, int T, int H, int W,
                                      int Ttot, int f0) {
  const long long total = 3LL * T * (2 * H) * (2 * W);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int xo = static_cast<int>(i % (2 * W));
    const int yo = static_cast<int>((i / (2 * W)) % (2 * H));
    const int f = static_cast<int>((i / (4LL * W * H)) % T);
    const int c = static_cast<int>(i / (4LL * W * H * T));
    const float v = __bfloat162float(y[((static_cast<long long>(f) * H + (yo >> 1)) * W + (xo >> 1)) * ldy + c * 4 +
                                       (xo & 1) * 2 + (yo & 1)]);
    video[((static_cast<long long>(c) * Ttot + f0 + f) * (2 * H) + yo) * (2 * W) + xo] =
        __float2bfloat16_rn(fminf(fmaxf(v, -1.f), 1.f));
  }
}

// rows[(f*h + y)*w + x][c*4 + r*2 + q] = video[c][f][2y + q][2x + r]   (patchify :285-301); columns >= 12 untouched
__global__ void vae_patchify_kernel(const __nv_bfloat16* video, __nv_bfloat16* rows, long long ldr, int T, int h, int w,
                                    int Ttot, int f0) {
  const long long total = 12LL * T * h * w;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int ch = static_cast<int>(i % 12);
    const long long pix = i / 12;
    const int x = static_cast<int>(pix % w);
    const int y = static_cast<int>((pix / w) % h);
    const int f = static_cast<int>(pix / (static_cast<long long>(w) * h));
    const int c = ch >> 2, r = (ch >> 1) & 1, q = ch & 1;
    rows[pix * ldr + ch] = video[((static_cast<long long>(c) * Ttot + f0 + f) * (2 * h) + 2 * y + q) * (2 * w) + 2 * x + r];
  }
}

// main[to][yo][xo][co] += mean_g x'[co*group + g][to][yo][xo]   (AvgDown3D :340-372), where x' is x with its time axis
// front-padded to a multiple of ft and (ft, fs, fs) blocks folded into channels: channel cf = c*factor + a*fs*fs + b*fs + d
// reads x[c][to*ft + a - pad_t][yo*fs + b][xo*fs + d] (zero for padded frames). x: bf16 [T, H, W, Cin] channel-last.
__global__ void vae_avgdown_add_kernel(__nv_bfloat16* main, const __nv_bfloat16* x, long long nout, int T, int H, int W,
                                       int Cin, int Cout, int ft, int fs) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const int Ho = H / fs, Wo = W / fs, factor = ft * fs * fs;
  const int group = Cin * factor / Cout;
  const int pad_t = (ft - T % ft) % ft;
  for (; i < nout; i += stride) {
    const int co = static_cast<int>(i % Cout);
    const long long pix = i / Cout;
    const int xo = static_cast<int>(pix % Wo);
    const int yo = static_cast<int>((pix / Wo) % Ho);
    const int to = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
    float acc = 0.f;
    for (int g = 0; g < group; ++g) {
      const int cf = co * group + g;
      const int c = cf / factor, rem = cf - c * factor;
      const int a = rem / (fs * fs), b = (rem / fs) % fs, d = rem % fs;
      const int tin = to * ft + a - pad_t;
      if (tin >= 0)
        acc += __bfloat162float(x[((static_cast<long long>(tin) * H + yo * fs + b) * W + xo * fs + d) * Cin + c]);
    }
    main[i] = __float2bfloat16_rn(__bfloat162float(main[i]) + __bfloat162float(__float2bfloat16_rn(acc / group)));
  }
}

// H-slab decode across GPUs (flexam_b200/dist.py SlabExchange): every rank owns a band of image rows of every padded grid
// [frames, Hp, Wp, C]; its first / last interior row is the bottom / top halo row of the neighbour above / below. One
// launch copies both rows of `T` live frames straight into the neighbours' grids (peer memory over NVLink).
__global__ void __launch_bounds__(256)
vae_halo_push_kernel(const uint4* grid, uint4* up, uint4* dn, int T, long long plane16, long long row16, int Hp) {
  const long long per_side = static_cast<long long>(T) * row16;
  const long long total = per_side * 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int side = i >= per_side;
    const long long r = i - side * per_side;
    const long long f = r / row16, c = r - f * row16;
    if (side == 0) {
      if (up != nullptr) up[f * plane16 + (Hp - 1) * row16 + c] = grid[f * plane16 + row16 + c];
    } else {
      if (dn != nullptr) dn[f * plane16 + c] = grid[f * plane16 + (Hp - 2) * row16 + c];
    }
  }
}

}  // namespace fx

extern "C" int fx_vae_patchify(const void* video, void* rows, int64_t ldr, int T, int h, int w, int Ttot, int frame0,
                               void* stream) {
  using namespace fx;
  FX_CHECK_ARG(video && rows && T > 0 && h > 0 && w > 0 && ldr >= 12 && frame0 >= 0 && frame0 + T <= Ttot,
               "fx_vae_patchify: bad arguments");
  vae_patchify_kernel<<<vae_grid(12LL * T * h * w), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(video), reinterpret_cast<__nv_bfloat16*>(rows), ldr, T, h, w, Ttot, frame0);
  FX_CHECK_LAUNCH("fx_vae_patchify");
  return FX_OK;
}

extern "C" int fx_vae_avgdown_add(void* main_io, const void* x, int T, int H, int W, int Cin, int Cout, int ft, int fs,
                                  void* stream) {
  using namespace fx;
  FX_CHECK_ARG(main_io && x && T > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (ft == 1 || ft == 2) &&
                   (fs == 1 || fs == 2) && H % fs == 0 && W % fs == 0 && (Cin * ft * fs * fs) % Cout == 0,
               "fx_vae_avgdown_add: bad arguments");
  const int To = (T + ft - 1) / ft;
  const long long nout = static_cast<long long>(To) * (H / fs) * (W / fs) * Cout;
  vae_avgdown_add_kernel<<<vae_grid(nout), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(main_io), reinterpret_cast<const __nv_bfloat16*>(x), nout, T, H, W, Cin, Cout, ft,
      fs);
  FX_CHECK_LAUNCH("fx_vae_avgdown_add");
  return FX_OK;
}

extern "C" int fx_vae_norm_act(const void* x, int64_t ldx, const void* gamma, void* out, int64_t ldo, int64_t npix,
                               int C, int H, int W, int pad, int frame0, int silu, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && out && npix > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && ldx >= C && ldo >= C,
               "fx_vae_norm_act: bad arguments (C, ldx, ldo multiples of 8)");
  FX_CHECK_ARG(H > 0 && W > 0 && npix % (static_cast<int64_t>(H) * W) == 0 && pad >= 0 && pad <= 1 && frame0 >= 0,
               "fx_vae_norm_act: bad grid");
  const long long blocks = (npix + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  vae_norm_act_kernel<<<static_cast<int>(blocks < cap ? blocks : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<const __nv_bfloat16*>(gamma),
      reinterpret_cast<__nv_bfloat16*>(out), ldo, npix, C, H, W, pad, frame0, silu);
  FX_CHECK_LAUNCH("fx_vae_norm_act");
  return FX_OK;
}

extern "C" int fx_vae_upsample2x(const void* x, void* out, int F, int H, int W, int C, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(x && out && F > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "fx_vae_upsample2x: bad arguments");
  const long long nout = 4LL * F * H * W * (C / 8);
  vae_upsample2x_kernel<<<vae_grid(nout), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), nout, H, W, C / 8);
  FX_CHECK_LAUNCH("fx_vae_upsample2x");
  return FX_OK;
}

extern "C" int fx_vae_time_interleave(const void* y, void* x, int T, int64_t P, int C, void* stream) {
  using namespace fx;
  FX_CHECK_ARG(y && x && T > 0 && P > 0 && C > 0 && C % 8 == 0, "fx_vae_time_interleave: bad arguments");
  const long long nvec = 2LL * T * P * (C / 8);
  vae_time_interleave_kernel<<<vae_grid(nvec), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(y), reinterpret_cast<uint4*>(x), nvec, P, C / 8);
  FX_CHECK_LAUNCH("fx_vae_time_interleave");
  return FX_OK;
}

extern "C" int fx_vae_dupup_add(void* main_io, const void* x, int Tout, int H, int W, int Cin, int Cout, int ft, int drop,
                                void* stream) {
  using namespace fx;
  FX_CHECK_ARG(main_io && x && Tout > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (ft == 1 || ft == 2) && drop >= 0 &&
                   drop < ft && (Cout * 4 * ft) % Cin == 0 && Cout % 8 == 0 &&
                   reinterpret_cast<uintptr_t>(main_io) % 16 == 0,
               "fx_vae_dupup_add: bad arguments");
  const long long nvec = 4LL * Tout * H * W * (Cout / 8);
  vae_dupup_add_kernel<<<vae_grid(nvec), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(main_io), reinterpret_cast<const __nv_bfloat16*>(x), nvec, H, W, Cin, Cout, ft,
      drop);
  FX_CHECK_LAUNCH("fx_vae_dupup_add");
  return FX_OK;
}

extern "C" int fx_softmax_rows_f32(const float* s, int64_t lds, void* p, int64_t ldp, int rows, int cols, float scale,
                                   void* stream) {
  using namespace fx;
  FX_CHECK_ARG(s && p && rows > 0 && cols > 0 && lds >= cols && ldp >= cols, "fx_softmax_rows_f32: bad arguments");
  vae_softmax_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      s, lds, reinterpret_cast<__nv_bfloat16*>(p), ldp, rows, cols, scale);
  FX_CHECK_LAUNCH("fx_softmax_rows_f32");
  return FX_OK;
}

extern "C" int fx_vae_unpatchify(const void* y, int64_t ldy, void* video, int T, int H, int W, int Ttot, int frame0,
                                 void* stream) {
  using namespace fx;
  FX_CHECK_ARG(y && video && T > 0 && H > 0 && W > 0 && ldy >= 12 && frame0 >= 0 && frame0 + T <= Ttot,
               "fx_vae_unpatchify: bad arguments");
  vae_unpatchify_kernel<<<vae_grid(12LL * T * H * W), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(y), ldy, reinterpret_cast<__nv_bfloat16*>(video), T, H, W, Ttot, frame0);
  FX_CHECK_LAUNCH("fx_vae_unpatchify");
  return FX_OK;
}

extern "C" int fx_vae_halo_push(const void* grid, void* up_grid, void* dn_grid, int frame0, int T, int Hp, int Wp, int C,
                                void* stream) {
  using namespace fx;
  FX_CHECK_ARG(grid && frame0 >= 0 && T > 0 && Hp >= 3 && Wp > 0 && C > 0 && C % 8 == 0,
               "fx_vae_halo_push: bad arguments");
  FX_CHECK_ARG((reinterpret_cast<uintptr_t>(grid) | reinterpret_cast<uintptr_t>(up_grid) |
                reinterpret_cast<uintptr_t>(dn_grid)) % 16 == 0, "fx_vae_halo_push: grids must be 16-byte aligned");
  if (up_grid == nullptr && dn_grid == nullptr) return FX_OK;
  const long long row16 = static_cast<long long>(Wp) * C / 8, plane16 = row16 * Hp, off = frame0 * plane16;
  vae_halo_push_kernel<<<vae_grid(2LL * T * row16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(grid) + off, up_grid ? reinterpret_cast<uint4*>(up_grid) + off : nullptr,
      dn_grid ? reinterpret_cast<uint4*>(dn_grid) + off : nullptr, T, plane16, row16, Hp);
  FX_CHECK_LAUNCH("fx_vae_halo_push");
  return FX_OK;
}
