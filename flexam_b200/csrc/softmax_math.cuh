// Softmax arithmetic of the attention kernel (fmha.cu): packed fp32x2 helpers, the FMA-pipe exp2, the per-chunk
// max / mask / exponential passes. Header-only so tests/native/pipe_rate.cu times exactly the shipped code.
#pragma once
#include "ptx.cuh"

namespace fx {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed fp32x2 arithmetic (sm_100): one issue slot for two lanes of work.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)),
        "l"(reinterpret_cast<const unsigned long long&>(c)));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
  return d;
}

// 2^x for a pair, x <= ~8: n = round(x) via the 1.5*2^23 magic constant (n lands in the low mantissa bits of t),
// f = x - n in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial, exponent patched in with an integer add.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  const float kMagic = 12582912.f;
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = add2(x, make_float2(kMagic, kMagic));
  const float2 n = add2(t, make_float2(-kMagic, -kMagic));
  const float2 f = fma2(n, make_float2(-1.f, -1.f), x);
  float2 q = fma2(f, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.2426111251f, 0.2426111251f));
  q = fma2(q, f, make_float2(0.6932609677f, 0.6932609677f));
  q = fma2(q, f, make_float2(0.9999280572f, 0.9999280572f));
  float2 r;
  r.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return r;
}


__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// keys at or beyond `valid_in_chunk` (relative to this 32-column chunk) do not exist: score = -inf
__device__ __forceinline__ void mask_chunk(uint32_t* s, int valid_in_chunk) {
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i >= valid_in_chunk) s[i] = 0xff800000u;
}

// running maximum over one 32-column chunk, two independent 3-input chains
__device__ __forceinline__ void max_chunk(const uint32_t* s, float& mxa, float& mxb) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    mxa = fmax3(mxa, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
    mxb = fmax3(mxb, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
  }
}

// P = exp2(S * scale_log2 - m) for one 32-column chunk on packed fp32 pairs, written to TMEM as 16 bf16x2 columns.
// MUFU.EX2 (16/clk/SM) alone would cost as many cycles per tile as the tile's MMAs, so kPoly8 of every 8 pairs are
// evaluated on the FMA pipe instead (round-to-nearest range reduction + degree-3 polynomial, rel. error 7.5e-5,
// far below the bf16 rounding of P).
template <int kPoly8>
__device__ __forceinline__ void exp_pack(const uint32_t* s, float2 sc2, float2 nm2, float2& sum_a, float2& sum_b,
                                         uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = fma2(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, nm2);
    float2 e;
#if defined(FX_FMHA_EXPERIMENT) && FX_FMHA_EXPERIMENT == 5
    e = x;  // timing experiment: no exponential at all (results are wrong)
#else
    if ((i & 7) < kPoly8) {
      e = exp2_poly2(x);
    } else {
      e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
    }
#endif
    if (i & 1) sum_b = add2(sum_b, e); else sum_a = add2(sum_a, e);
    pk[i] = pack_bf16x2(e.x, e.y);
  }
}
template <int kPoly8>
__device__ __forceinline__ void exp_chunk(const uint32_t* s, float2 sc2, float2 nm2, float2& sum_a, float2& sum_b,
                                          uint32_t tmem_dst) {
  uint32_t pk[16];
  exp_pack<kPoly8>(s, sc2, nm2, sum_a, sum_b, pk);
  tmem_st16(tmem_dst, pk);
}

}  // namespace fx
