// Developer tool (GPU box): raw tcgen05.mma issue/execute rate per instruction shape, one CTA per SM, no pipeline
// around it. One thread issues `iters` x 4 UMMAs (K = 16 each) back to back into one TMEM accumulator and commits
// once; clock64() around issue + completion gives cycles per instruction. Operands are whatever is in shared memory
// after a fill with small bf16 values. Modes: SS (A, B in smem) and TS (A in TMEM), B K-major or MN-major.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../flexam_b200/csrc/ptx.cuh"

using namespace fx;

template <int N, bool kTS, bool kBMn>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // [128][64] bf16, SW128
  uint8_t* sB = smem + 16384;          // [N][64] bf16 (K-major) or [64][N] (MN-major)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (16384 + N * 128) / 2; i += 128)
    reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(((i * 37) % 17 - 8) * 0.03125f);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, kBMn);
    const uint32_t a = smem_u32(sA), b = smem_u32(sB);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t bd = kBMn ? umma_desc_sw128(b + k * 2048, 8192, 1024) : umma_desc_sw128(b + k * 32, 16, 1024);
        if (kTS) umma_ts(tmem + 256, tmem + k * 8, bd, idesc, 1);
        else umma_ss(tmem + 256, umma_desc_sw128(a + k * 32, 16, 1024), bd, idesc, 1);
      }
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// The attention kernel's instruction mix without any of its synchronisation: per "KV step" 8 SS UMMAs into S0,
// 8 TS (B MN-major) into O0, 8 SS into S1, 8 TS into O1 (kSwitch = accumulator changes as in the kernel; with
// kSwitch = false everything accumulates into one buffer), optionally a commit after every group of 8.
template <bool kSwitch, bool kCommit, int kGroup, int kWaits = 0, int kWhat = 3>
__global__ void __launch_bounds__(128, 1) umma_mix_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                  // 2 x [128][128] bf16 as two 64-wide halves
  uint8_t* sK = smem + 65536;          // [128][128]
  uint8_t* sV = smem + 98304;          // [128 keys][128 d] MN-major halves
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 131072);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 3);
  for (int i = threadIdx.x; i < 131072 / 2; i += 128)
    reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(((i * 37) % 17 - 8) * 0.03125f);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], 1);
    fence_mbar_init();
    mbar_arrive(&bar[2]);  // phase 0 of bar[2] is complete from here on: waits on parity 0 succeed at once
  }
  if (threadIdx.x < 32) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, false, true);
    const uint32_t q = smem_u32(sQ), kk = smem_u32(sK), v = smem_u32(sV);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const uint32_t s_acc = kSwitch ? tmem + w * 128 : tmem;
        const uint32_t o_acc = kSwitch ? tmem + 256 + w * 128 : tmem;
#pragma unroll
        for (int q2 = 0; q2 < kWaits; ++q2) {
          if (kWhat & 1) mbar_wait(&bar[2], 0);
          if (kWhat & 2) tc_fence_after();
          if (kWhat & 4) while (!mbar_try_wait(&bar[2], 0)) {}
        }
#pragma unroll
        for (int k = 0; k < kGroup; ++k) {
          const uint32_t off = ((k & 7) >> 2) * 16384 + (k & 3) * 32;
          umma_ss(s_acc, umma_desc_sw128(q + w * 32768 + off, 16, 1024), umma_desc_sw128(kk + off, 16, 1024), idesc_qk, 1);
        }
        if (kCommit) umma_commit(&bar[1]);
#pragma unroll
        for (int q2 = 0; q2 < kWaits; ++q2) {
          if (kWhat & 1) mbar_wait(&bar[2], 0);
          if (kWhat & 2) tc_fence_after();
          if (kWhat & 4) while (!mbar_try_wait(&bar[2], 0)) {}
        }
#pragma unroll
        for (int k = 0; k < kGroup; ++k)
          umma_ts(o_acc, tmem + 128 * (1 - w) + (k & 7) * 8, umma_desc_sw128(v + (k & 7) * 2048, 16384, 1024), idesc_pv, 1);
        if (kCommit) umma_commit(&bar[1]);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <bool kSwitch, bool kCommit, int kGroup, int kWaits = 0, int kWhat = 3>
void run_mix(const char* name, int grid) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 131072 + 64 + 1024;
  cudaFuncSetAttribute(umma_mix_kernel<kSwitch, kCommit, kGroup, kWaits, kWhat>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 500;
  long long h[2] = {0, 0};
  for (int r = 0; r < 2; ++r) {
    umma_mix_kernel<kSwitch, kCommit, kGroup, kWaits, kWhat><<<grid, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      return;
    }
  }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double n = iters * 4.0 * kGroup;
  printf("%-44s grid %3d: %.1f clk/UMMA (ideal 64), %.0f clk per group switch overhead\n", name, grid, h[1] / n,
         (h[1] / n - 64.0) * kGroup);
  cudaFree(d);
}

// Latency probes on an idle pipe: (a) one mbarrier.try_wait on a completed phase, (b) the time the issuing thread
// needs to get n UMMAs (128x128x16, SS) accepted, for n = 1..32, vs the time until they have all completed.
__global__ void __launch_bounds__(128, 1) probe_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  for (int i = threadIdx.x; i < 32768 / 2; i += 128) reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(0.25f);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
    mbar_arrive(&bar[7]);
  }
  if (threadIdx.x < 32) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, false, false);
    const uint32_t a = smem_u32(smem), b = a + 16384;
    long long t0 = clock64();
    long long t1 = clock64();
    out[0] = t1 - t0;  // clock64 back to back
    t0 = clock64();
    while (!mbar_try_wait(&bar[7], 0)) {}
    t1 = clock64();
    out[1] = t1 - t0;  // try_wait on a completed phase, idle pipe
    int idx = 2;
    uint32_t ph = 0;
    for (int n = 1; n <= 32; n *= 2) {
      t0 = clock64();
      for (int k = 0; k < n; ++k)
        umma_ss(tmem, umma_desc_sw128(a + (k & 3) * 32, 16, 1024), umma_desc_sw128(b + (k & 3) * 32, 16, 1024), idesc, 1);
      t1 = clock64();
      while (!mbar_try_wait(&bar[7], 0)) {}
      const long long t2 = clock64();
      umma_commit(&bar[0]);
      mbar_wait(&bar[0], ph);
      ph ^= 1;
      const long long t3 = clock64();
      out[idx++] = t1 - t0;  // issue time
      out[idx++] = t2 - t1;  // ready-barrier try_wait right behind the issue
      out[idx++] = t3 - t0;  // until complete
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

void run_probe() {
  long long* d;
  cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34944);
  long long h[64];
  for (int r = 0; r < 2; ++r) {
    probe_kernel<<<1, 128, 34944>>>(d);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("clock64 back-to-back %lld clk; try_wait(ready) on idle pipe %lld clk\n", h[0], h[1]);
  int idx = 2;
  for (int n = 1; n <= 32; n *= 2, idx += 3)
    printf("n=%2d UMMAs: issue %5lld clk, try_wait(ready) behind them %5lld clk, all complete after %5lld clk\n", n,
           h[idx], h[idx + 1], h[idx + 2]);
  cudaFree(d);
}

template <int N, bool kTS, bool kBMn>
void run(const char* name, int grid) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 16384 + N * 128 + 64 + 1024;
  cudaFuncSetAttribute(umma_rate_kernel<N, kTS, kBMn>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  long long h[2] = {0, 0};
  for (int r = 0; r < 2; ++r) {
    umma_rate_kernel<N, kTS, kBMn><<<grid, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      return;
    }
  }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double n = iters * 4.0;
  printf("%-34s grid %3d: issue %.1f clk/UMMA, complete %.1f clk/UMMA, ideal %.0f, %.0f flop/clk/SM\n", name, grid,
         h[0] / n, h[1] / n, 128.0 * N / 256, 2.0 * 128 * N * 16 * n / h[1]);
  cudaFree(d);
}

int main() {
  run_probe();
  for (int grid : {1, 148}) {
    run<64, false, false>("SS 128x64x16", grid);
    run<128, false, false>("SS 128x128x16", grid);
    run<256, false, false>("SS 128x256x16", grid);
    run<128, false, true>("SS 128x128x16 B MN-major", grid);
    run<128, true, false>("TS 128x128x16", grid);
    run<128, true, true>("TS 128x128x16 B MN-major", grid);
    run<256, true, false>("TS 128x256x16", grid);
    run_mix<false, false, 8>("mix SS8/TS8, one accumulator", grid);
    run_mix<true, false, 8>("mix SS8/TS8, S0 O0 S1 O1 accumulators", grid);
    run_mix<true, true, 8>("mix SS8/TS8, 4 accumulators + commits", grid);
    run_mix<true, false, 16>("mix SS16/TS16, 4 accumulators", grid);
    run_mix<true, true, 8, 1>("mix SS8/TS8 + commit + 1 ready-barrier wait", grid);
    run_mix<true, true, 8, 2>("mix SS8/TS8 + commit + 2 ready-barrier waits", grid);
    run_mix<true, true, 8, 1, 1>("mix + 1 mbar_wait (watchdog loop) only", grid);
    run_mix<true, true, 8, 1, 4>("mix + 1 bare try_wait loop only", grid);
    run_mix<true, true, 8, 1, 2>("mix + 1 tcgen05.fence::after only", grid);
  }
  return 0;
}
