// Standalone B200 probe 2 (not part of the library).
//   A. cost of a chain of 96 tcgen05.mma kind::f16 (M = 128, K = 16) as a function of N and of the number of independent
//      accumulators the chain rotates over (is a small-N MMA bound by issue, by the dependent accumulate, or by the floor?)
//   B. all-gather of a 4 KB piece per CTA inside a cluster of 16, per iteration:
//        B1 DSMEM bulk copies (cp.async.bulk shared::cta -> shared::cluster), 16 per CTA
//        B2 st.global + fence.proxy.async + one multicast bulk copy global -> all 16 CTAs
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nabu_b200/csrc -I include tools/probe2.cu -o tools/probe2
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
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
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// ---- A ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) mma_chain_kernel(int N, int nacc, int ts, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* As = sm;                 // 8 K blocks x [128 rows x 128 B]  = 128 KB
  uint8_t* Bs = sm + 131072;        // 4 K blocks x [128 rows x 128 B]  = 64 KB (reused round robin)
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (131072 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t id = idesc_f16(128, N);
  if (tid == 0) {
    const long long t0 = gtime();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
      for (int i = 0; i < 96; ++i) {
        const int kb = (i >> 2) & 7, ks = i & 3;
        const uint64_t ad = make_desc(smem_u32(As) + kb * 16384 + ks * 32, 16, 1024, 2);
        const uint64_t bd = make_desc(smem_u32(Bs) + (kb & 3) * 16384 + ks * 32, 16, 1024, 2);
        const uint32_t d = tm + (uint32_t)(i % nacc) * (uint32_t)N;
        if (ts) umma_f16_ts(d, tm + 256 + (i & 31) * 8, bd, id, i >= nacc);
        else umma_f16_ss(d, ad, bd, id, i >= nacc);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), it & 1);
      tc_fence_after();
    }
    out[0] = gtime() - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- B ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}

constexpr int CLS = 16;
constexpr int PIECE = 4096;
__global__ void __launch_bounds__(256, 1) gather_kernel(int mode, int iters, uint8_t* gbuf, long long* out, unsigned* check) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* hbuf = sm;                         // [2 parity][CLS][PIECE] = 128 KB
  uint8_t* mine = sm + 2 * CLS * PIECE;       // my piece (source of the DSMEM copies)
  __shared__ __align__(8) uint64_t full[2];
  const int tid = threadIdx.x;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int cluster = blockIdx.x / CLS;
  uint8_t* gmine = gbuf + ((size_t)cluster * CLS + rank) * 2 * PIECE;    // [2 parity][PIECE]
  if (tid == 0) { mbar_init(smem_u32(&full[0]), 1); mbar_init(smem_u32(&full[1]), 1); fence_barrier_init(); }
  __syncthreads();
  cl_arrive(); cl_wait();
  unsigned bad = 0;
  const long long t0 = gtime();
  for (int it = 0; it < iters; ++it) {
    const int par = it & 1;
    // arm my barrier for this iteration's 16 pieces; peers cannot complete_tx phase `it` before the arm of phase it
    // is needed... (expect_tx may arrive after complete_tx within the same phase: tx-count goes transiently negative)
    if (tid == 0) mbar_expect_tx(smem_u32(&full[par]), CLS * PIECE);
    const uint4 v = make_uint4(it, rank, tid, 0x5a5a5a5a);
    if (mode == 1) {
      reinterpret_cast<uint4*>(mine + par * PIECE)[tid] = v;
      fence_proxy_async_smem();
      __syncthreads();
      if (tid < CLS)
        bulk_s2s(mapa(smem_u32(hbuf + ((size_t)par * CLS + rank) * PIECE), tid), smem_u32(mine + par * PIECE), PIECE,
                 mapa(smem_u32(&full[par]), tid));
    } else {
      reinterpret_cast<uint4*>(gmine + (size_t)par * PIECE)[tid] = v;
      fence_proxy_async_all();
      __syncthreads();
      if (tid == 0)
        bulk_g2s_mc(smem_u32(hbuf + ((size_t)par * CLS + rank) * PIECE), gmine + (size_t)par * PIECE, PIECE,
                    smem_u32(&full[par]), (uint16_t)0xFFFF);
    }
    mbar_wait(smem_u32(&full[par]), (it >> 1) & 1);
    // verify: every peer's piece carries this iteration
    const uint4 g = reinterpret_cast<const uint4*>(hbuf + ((size_t)par * CLS + (tid & 15)) * PIECE)[tid];
    if (g.x != (unsigned)it || g.y != (unsigned)(tid & 15) || g.z != (unsigned)tid) ++bad;
    __syncthreads();                          // everyone has read `mine`/hbuf[par] before the next overwrite of `mine`
  }
  const long long t1 = gtime();
  if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (bad) atomicAdd(check, bad);
  cl_arrive(); cl_wait();
}

int main() {
  long long* out; CK(cudaMalloc(&out, 16));
  unsigned* check; CK(cudaMalloc(&check, 4));
  const int iters = 1000;
  {
    const size_t smem = 1024 + 131072 + 65536;
    CK(cudaFuncSetAttribute(mma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int Ns[3] = {32, 64, 128};
    for (int ts = 0; ts < 2; ++ts)
      for (int n = 0; n < 3; ++n)
        for (int nacc = 1; nacc <= 4; nacc *= 2) {
          if (Ns[n] * nacc > 256) continue;
          mma_chain_kernel<<<1, 128, smem>>>(Ns[n], nacc, ts, iters, out);
          CK(cudaDeviceSynchronize());
          long long ns; CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
          printf("96 MMA %s M=128 N=%3d K=16, %d accumulators: %.3f us per chain (%.1f ns / MMA)\n", ts ? "TS" : "SS", Ns[n], nacc,
                 ns / 1e3 / iters, (double)ns / iters / 96);
        }
  }
  {
    const size_t smem = 1024 + 2 * CLS * PIECE + 2 * PIECE;
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    uint8_t* gbuf; CK(cudaMalloc(&gbuf, (size_t)8 * CLS * 2 * PIECE));
    for (int ncl = 1; ncl <= 4; ncl *= 4)
      for (int mode = 1; mode <= 2; ++mode) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ncl * CLS); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CLS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaMemset(check, 0, 4));
        int m = mode, it = iters;
        cudaError_t e = cudaLaunchKernelEx(&cfg, gather_kernel, m, it, gbuf, out, check);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); continue; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("gather mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        long long ns; unsigned bad;
        CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&bad, check, 4, cudaMemcpyDeviceToHost));
        printf("all-gather 16 x 4 KB, %d cluster(s), %s: %.3f us / iteration, mismatches %u\n", ncl,
               mode == 1 ? "DSMEM bulk copies" : "st.global + multicast bulk copy", ns / 1e3 / iters, bad);
      }
  }
  return 0;
}
