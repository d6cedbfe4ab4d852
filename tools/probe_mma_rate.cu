// Standalone B200 probe: cost of a tcgen05.mma kind::f16 (M = 128, K = 16) as a function of N, of the A operand's home
// (shared memory "SS" / tensor memory "TS") and of how many accumulators the stream alternates between.  One CTA, one
// issuing thread, 192 MMAs per measurement, time from the first issue to the commit's arrival (clock64 of the SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nabu_b200/csrc -I include tools/probe_mma_rate.cu -o tools/probe_mma_rate
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "tc_common.cuh"

using namespace nabu::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ inline uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <int N, int TS, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;   // fp16 1.0
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    constexpr uint32_t id = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_u = smem_u32(sm), b_u = smem_u32(sm) + 32768;     // A: 2 K blocks of [128 x 64]; B: 2 K blocks of [<=256 x 64]
    tc_fence_after();
    // descriptors of the 8 (K block, k-step) positions, computed once
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ad[j] = make_desc(a_u + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024, 2);
      bd[j] = make_desc(b_u + (j >> 2) * 32768 + (j & 3) * 32, 16, 1024, 2);
    }
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int o = 0; o < 8; ++o) {
#pragma unroll
        for (int j = 0; j < 24; ++j) {                     // 192 MMAs in all
          const uint32_t d = tm + (uint32_t)((j % NACC) * N);
          const uint32_t acc = (o | (j >= NACC)) ? 1u : 0u;
          if (TS) umma_ts(d, tm + 448 + (uint32_t)((j & 7) * 8), bd[j & 7], id, acc);
          else umma_ss(d, ad[j & 7], bd[j & 7], id, acc);
        }
      }
      const long long t1 = clock64();
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), rep & 1);
      const long long t2 = clock64();
      if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int TS, int NACC>
void run(long long* out) {
  const size_t smem = 1024 + 32768 + 65536;
  CK(cudaFuncSetAttribute(rate_kernel<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rate_kernel<N, TS, NACC><<<1, 128, smem>>>(out);
  CK(cudaDeviceSynchronize());
  long long h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
  printf("%s M=128 N=%3d K=16, %d accumulator(s): issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor law 128*N/256 = %d)\n",
         TS ? "TS" : "SS", N, NACC, (double)h[0] / 192, (double)h[1] / 192, 128 * N / 256);
}

// two issuing threads (warps 0 and 1), each its own accumulators and commit barrier: does the ~47-cycle issue floor of a
// small-N MMA belong to the issuing thread or to the tensor pipe?
template <int N, int TS>
__global__ void __launch_bounds__(128, 1) rate2_kernel(long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if ((tid & 31) == 0 && warp < 2) {
    constexpr uint32_t id = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_u = smem_u32(sm), b_u = smem_u32(sm) + 32768;
    tc_fence_after();
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ad[j] = make_desc(a_u + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024, 2);
      bd[j] = make_desc(b_u + (j >> 2) * 32768 + (j & 3) * 32, 16, 1024, 2);
    }
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int o = 0; o < 4; ++o) {
#pragma unroll
        for (int j = 0; j < 24; ++j) {                     // 96 MMAs per issuer, 192 in all
          const uint32_t d = tm + (uint32_t)(warp * 2 * N + (j & 1) * N);
          const uint32_t acc = (o | (j >= 2)) ? 1u : 0u;
          if (TS) umma_ts(d, tm + 448 + (uint32_t)((j & 7) * 8), bd[j & 7], id, acc);
          else umma_ss(d, ad[j & 7], bd[j & 7], id, acc);
        }
      }
      const long long t1 = clock64();
      umma_commit(smem_u32(&bar[warp]));
      mbar_wait(smem_u32(&bar[warp]), rep & 1);
      const long long t2 = clock64();
      if (rep == 2) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
    }
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}
template <int N, int TS>
void run2(long long* out) {
  const size_t smem = 1024 + 32768 + 65536;
  CK(cudaFuncSetAttribute(rate2_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rate2_kernel<N, TS><<<1, 128, smem>>>(out);
  CK(cudaDeviceSynchronize());
  long long h[4]; CK(cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost));
  printf("2 issuers %s N=%3d: per issuer 96 MMAs: issue %.1f / %.1f cyc per MMA, complete %.1f / %.1f; both together %.1f cyc per MMA of 192\n",
         TS ? "TS" : "SS", N, h[0] / 96.0, h[2] / 96.0, h[1] / 96.0, h[3] / 96.0, (double)(h[1] > h[3] ? h[1] : h[3]) / 192);
}

int main() {
  long long* out; CK(cudaMalloc(&out, 64));
  run<16, 0, 1>(out); run<16, 0, 2>(out); run<32, 0, 1>(out); run<32, 0, 2>(out); run<32, 0, 3>(out);
  run<64, 0, 1>(out); run<64, 0, 2>(out); run<64, 0, 3>(out); run<128, 0, 1>(out); run<128, 0, 2>(out); run<128, 0, 3>(out);
  run<16, 1, 1>(out); run<16, 1, 2>(out); run<32, 1, 1>(out); run<32, 1, 2>(out); run<32, 1, 3>(out);
  run<64, 1, 1>(out); run<64, 1, 2>(out); run<64, 1, 3>(out); run<128, 1, 1>(out); run<128, 1, 2>(out); run<128, 1, 3>(out);
  run2<32, 0>(out); run2<64, 0>(out); run2<128, 0>(out); run2<32, 1>(out); run2<64, 1>(out); run2<128, 1>(out);
  return 0;
}
