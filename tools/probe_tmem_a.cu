// Standalone B200 probe (not part of the library): facts the single-cluster recurrence design rests on.
//   1. tcgen05.mma kind::f16 with the A operand in TMEM (fp16 packed two per 32-bit column, lane = row): correctness
//      against a host product, K = 64 as 4 k-steps of 8 columns.
//   2. Co-residency of 8 clusters of 16 CTAs (non-portable cluster size) at ~200 KB shared memory per CTA.
//   3. Cost per iteration of (a) two cluster barriers, (b) + an all-to-all of 4 KB per peer through DSMEM stores,
//      (c) + 96 MMAs of 128x32x16 (64 with A in TMEM, 32 with A in shared memory).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nabu_b200/csrc -I include tools/probe_tmem_a.cu -o tools/probe_tmem_a
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "tc_common.cuh"

using namespace nabu::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ inline uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__host__ __device__ inline uint32_t sw128_h(int row, int k) {
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)k & 7u) << 1);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- 1. TS MMA correctness -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) ts_kernel(const __half* A, const __half* B, float* D) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 32 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    *reinterpret_cast<__half*>(sm + sw128_h(n, k)) = B[n * 64 + k];
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  // A -> TMEM columns [32, 64): lane = row, column c holds (A[row][2c], A[row][2c+1])
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t v[8];
    for (int c = 0; c < 8; ++c) {
      const __half2 h = __halves2half2(A[row * 64 + 2 * (c0 + c)], A[row * 64 + 2 * (c0 + c) + 1]);
      v[c] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + 32 + c0, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t id = idesc_f16(128, 32);
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t bd = make_desc(smem_u32(sm) + ks * 32, 16, 1024, 2);
      umma_f16_ts(tm, tm + 32 + ks * 8, bd, id, ks != 0);
    }
    umma_commit(smem_u32(&bar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int n = 0; n < 32; ++n) D[row * 32 + n] = __uint_as_float(r[n]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

// ---- 2/3. cluster-of-16 loop ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void st_cl_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int CLS>
__global__ void __launch_bounds__(256, 1) loop_kernel(int iters, int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* hbuf = sm;                       // 64 KB: B operand [8 kb][hi|lo][32 rows x 128 B]
  uint8_t* wlo = sm + 65536;                // 128 KB: A lo operand [8 kb][128 rows x 128 B]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = tid; i < (65536 + 131072) / 4; i += 256) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  cl_arrive(); cl_wait();
  const uint32_t id = idesc_f16(128, 32);
  long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int it = 0; it < iters; ++it) {
    if (mode >= 2) {
      if (tid == 0) {
        fence_proxy_async_all();
        tc_fence_after();
        for (int kb = 0; kb < 8; ++kb)
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bh = make_desc(smem_u32(hbuf) + (kb * 2 + 0) * 4096 + ks * 32, 16, 1024, 2);
            const uint64_t bl = make_desc(smem_u32(hbuf) + (kb * 2 + 1) * 4096 + ks * 32, 16, 1024, 2);
            const uint64_t al = make_desc(smem_u32(wlo) + kb * 16384 + ks * 32, 16, 1024, 2);
            const uint32_t acc = (kb | ks) != 0;
            umma_f16_ts(tm + 256, tm + (kb * 4 + ks) * 8, bh, id, acc);
            umma_f16_ts(tm + 288, tm + (kb * 4 + ks) * 8, bl, id, acc);
            umma_f16_ss(tm + 288, al, bh, id, 1u);
          }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), it & 1);
      tc_fence_after();
      if (mode >= 3) {
        uint32_t r1[16], r2[16];
        const uint32_t ta = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 16;
        tmem_ld16(ta, r1); tmem_ld16(ta + 32, r2);
        tmem_ld_wait();
        uint32_t acc = 0;
        for (int i = 0; i < 16; ++i) acc += r1[i] ^ r2[i];
        if (acc == 0x12345678) out[1] = acc;
        tc_fence_before();
      }
    }
    cl_arrive(); cl_wait();                 // "receive buffers free"
    if (mode >= 1) {
      const uint32_t base = smem_u32(hbuf) + (rank >> 1) * 8192 + (tid >> 7) * 4096 +
                            sw128_h((tid & 127) >> 2, (rank & 1) * 32 + (tid & 3) * 8);
      for (int d = 0; d < CLS; ++d) st_cl_v4(mapa(base, d), it, tid, d, rank);
      fence_proxy_async_all();
    }
    cl_arrive(); cl_wait();                 // "h landed"
  }
  long long t1 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before();
  cl_arrive(); cl_wait();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int CLS>
void run_loop(int nclusters_wanted) {
  auto* fn = loop_kernel<CLS>;
  const size_t smem = 1024 + 65536 + 131072;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CLS > 8) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters_wanted * CLS);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CLS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, fn, &cfg);
  printf("cluster size %d, smem %zu: max active clusters %d (%s)\n", CLS, smem, n, cudaGetErrorString(e));
  if (e != cudaSuccess || n < nclusters_wanted) { cudaGetLastError(); if (n <= 0) return; cfg.gridDim = dim3(n * CLS); }
  long long* out; CK(cudaMalloc(&out, 16)); CK(cudaMemset(out, 0, 16));
  const int iters = 2000;
  const char* names[4] = {"2 cluster barriers", "+ DSMEM all-to-all 4 KB/peer", "+ 96 MMA 128x32x16 + commit/wait", "+ tcgen05.ld 2x16"};
  for (int mode = 0; mode < 4; ++mode) {
    int it = iters, m = mode;
    e = cudaLaunchKernelEx(&cfg, fn, it, m, out);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    CK(cudaDeviceSynchronize());
    long long ns; CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
    printf("  mode %d (%s): %.3f us / iteration\n", mode, names[mode], ns / 1e3 / iters);
  }
  cudaFree(out);
}

int main() {
  // ---- test 1 ----
  std::vector<__half> A(128 * 64), B(32 * 64);
  srand(1);
  for (auto& v : A) v = __float2half((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = __float2half((rand() % 2001 - 1000) / 1000.f);
  __half *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 128 * 32 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
  ts_kernel<<<1, 128, 8192>>>(dA, dB, dD);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * 32);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += (double)__half2float(A[m * 64 + k]) * (double)__half2float(B[n * 64 + k]);
      maxerr = fmax(maxerr, fabs(ref - D[m * 32 + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("TS MMA (A in TMEM, packed fp16 pairs): max abs err %.3e (max |ref| %.3f) -> %s\n", maxerr, maxref, maxerr < 1e-3 ? "OK" : "MISMATCH");
  // ---- tests 2/3 ----
  run_loop<16>(8);
  run_loop<8>(16);
  return 0;
}
