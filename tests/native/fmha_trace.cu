// Developer tool (GPU box): pipeline timeline of the attention kernel. Builds flexam_b200/csrc/fmha.cu with
// -DFX_FMHA_TRACE (see tests/native/Makefile), runs one self-attention launch at the config-2 shape and prints, for
// CTA (0,0,0), the clock64() stamps of the MMA issuer and of one softmax warp per query tile for KV steps 8..23,
// relative to the first stamp. Not part of the library or of the test suite.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../../include/flexam_b200.h"

#ifdef FX_FMHA_TRACE
extern "C" int fx_fmha_trace_read(long long* dst);
#endif

int main(int argc, char** argv) {
  const int B = 2, H = 24, L = argc > 1 ? atoi(argv[1]) : 11648;
  const size_t n = static_cast<size_t>(B) * L * 3 * H * 128;
  std::vector<__nv_bfloat16> h(n);
  unsigned s = 12345u;
  for (size_t i = 0; i < n; ++i) {
    s = s * 1664525u + 1013904223u;
    h[i] = __float2bfloat16(((s >> 8) & 0xffff) / 65536.0f * 2.f - 1.f);
  }
  __nv_bfloat16 *qkv, *o;
  cudaMalloc(&qkv, n * 2);
  cudaMalloc(&o, n / 3 * 2);
  cudaMemcpy(qkv, h.data(), n * 2, cudaMemcpyHostToDevice);
  const long long sl = 3LL * H * 128, sb = sl * L;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    int st = fx_fmha_fwd(qkv, sb, sl, qkv + H * 128, sb, sl, qkv + 2 * H * 128, sb, sl, o, (long long)H * 128 * L,
                         H * 128, B, H, L, L, 0.0883883f, nullptr);
    cudaEventRecord(e1);
    if (st != 0) {
      printf("fx_fmha_fwd failed: %s\n", fx_last_error());
      return 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("kernel failed: %s\n", cudaGetErrorString(e));
      return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("launch %d: %.3f ms, %.1f TFLOP/s\n", it, ms, 4.0 * B * H * L * (double)L * 128 / ms / 1e9);
  }
  {  // accuracy on 16 rows of (batch 0, head 1) spread over both query tiles, all four TMEM lane quadrants and the
     // ragged last CTA, against a double-precision softmax(QK^T/sqrt(d))V on the host; plus a checksum of the whole
     // output (the pipelines selected by FX_FMHA_PIPE compute the same arithmetic: their checksums must agree)
    const int hh = 1, nrows = 16;
    const int rows[nrows] = {0, 37, 70, 127, 128, 161, 200, 255, 256, 300, 511, 5000, L - 130, L - 65, L - 2, L - 1};
    std::vector<__nv_bfloat16> ho(n / 3);
    cudaMemcpy(ho.data(), o, ho.size() * 2, cudaMemcpyDeviceToHost);
    unsigned long long fnv = 1469598103934665603ULL;
    const unsigned short* raw = reinterpret_cast<const unsigned short*>(ho.data());
    for (size_t i = 0; i < ho.size(); ++i) fnv = (fnv ^ raw[i]) * 1099511628211ULL;
    double num = 0, den = 0;
    std::vector<double> sc(L), acc(128);
    for (int ri = 0; ri < nrows; ++ri) {
      const int r = ((rows[ri] % L) + L) % L;
      const __nv_bfloat16* qr = h.data() + static_cast<size_t>(r) * sl + hh * 128;
      double mx = -1e300;
      for (int k = 0; k < L; ++k) {
        const __nv_bfloat16* kr = h.data() + static_cast<size_t>(k) * sl + H * 128 + hh * 128;
        double d = 0;
        for (int c = 0; c < 128; ++c) d += static_cast<double>(__bfloat162float(qr[c])) * __bfloat162float(kr[c]);
        sc[k] = d * 0.0883883;
        mx = sc[k] > mx ? sc[k] : mx;
      }
      double lsum = 0;
      for (int c = 0; c < 128; ++c) acc[c] = 0;
      for (int k = 0; k < L; ++k) {
        const double pk = exp(sc[k] - mx);
        lsum += pk;
        const __nv_bfloat16* vr = h.data() + static_cast<size_t>(k) * sl + 2 * H * 128 + hh * 128;
        for (int c = 0; c < 128; ++c) acc[c] += pk * __bfloat162float(vr[c]);
      }
      for (int c = 0; c < 128; ++c) {
        const double want = acc[c] / lsum, got = __bfloat162float(ho[static_cast<size_t>(r) * H * 128 + hh * 128 + c]);
        num += (got - want) * (got - want);
        den += want * want;
      }
    }
    printf("accuracy: rel-L2 %.3e over %d rows; output checksum %016llx\n", sqrt(num / den), nrows, fnv);
  }
#ifdef FX_FMHA_TRACE
  long long t[24 * 64];
  fx_fmha_trace_read(t);
  // pipeline 1 meaning / pipeline 2-3 meaning (FX_FMHA_PIPE): rows 2 and 4 are "PV0 (+QK1) issued", "PV1 (+QK0) issued",
  // row 5 is the QK issuer's "QK1(j) issued" (pipeline 3 only)
  const char* names[24] = {"mma:v_full", "mma:p0_seen", "mma:pv0+qk_issued", "mma:p1_seen", "mma:iter_issued", "qk:qk1_issued",
                           "sm0:wait_s", "sm0:s_seen", "sm0:exp_done", "sm0:arrived",
                           "sm0:max_exchanged", "sm0:s_loaded", "sm0:max_done", "sm0:exp_half",
                           // pipeline 2/3 only: tile 0 after the token barrier / after p_free; tile 1 (warp 8)
                           "sm0:exp_start", "sm0:p_free_seen", "sm1:wait_s", "sm1:s_seen", "sm1:s_loaded",
                           "sm1:max_exchanged", "sm1:exp_start", "sm1:exp_half", "sm1:exp_done", "sm1:arrived"};
  const long long t0 = t[0 * 64 + 8];
  printf("%-20s", "event \\ kv step");
  for (int j = 8; j < 24; ++j) printf("%7d", j);
  printf("\n");
  for (int w = 0; w < 24; ++w) {
    printf("%-20s", names[w]);
    for (int j = 8; j < 24; ++j) printf("%7lld", t[w * 64 + j] - t0);
    printf("\n");
  }
  printf("period (mma:p0_seen[j+1]-[j]) over steps 8..40: %.0f clk\n", (t[64 + 40] - t[64 + 8]) / 32.0);
#endif
  return 0;
}
