// Shared by the cluster recurrences (blstm_cl.cu: FFMA, blstm_cl_tc.cu: tcgen05): parameters, DSMEM / bulk-copy /
// cluster-barrier PTX wrappers and the NABU_REC_TRACE phase stamps.
#pragma once
#include "common.cuh"
#include "blstm_cl.h"
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace nabu {
namespace {

constexpr int CL_THREADS = 256;
constexpr int CL_WARPS = 8;

struct ClParams {
  const float* kernel[2];
  float* gates[2];          // in: activated i,g,f,o ; out: dZ
  const float* cells[2];
  const float* dy;
  float* y;                 // fwd: output [B, yT, 2H]
  float* dbpart;            // [2 dir][8][4H] (slot 0 used)
  float* xchg;              // [2 dir][2 parity][cluster][4 gate][CLS*HS unit][BT]
  float* dcbuf;             // [2 dir][BT][H]
  unsigned* counters;       // [2 dir][<=16 clusters]
  const int* len;
  long long* trace;         // NABU_REC_TRACE: per-phase globaltimer stamps of every CTA for steps [TRACE_S0, TRACE_S0 + TRACE_N)
  int B, T, yT, D, H;
  int fences;               // NABU_REC_FENCES: per-thread __threadfence + proxy fence before publishing (debug)
  // fp16 hi/lo planes for the tensor-core GEMMs, written by the recurrences themselves (gemm_h2.cu's operand format):
  // fwd: yh / yl [B, yT, 2H] = split of y * 32;  bwd: zh / zl [B*T, 8H] (fw | bw gate gradients side by side) = split of
  // dZ * S, S the power of two that puts the largest |dy| of the batch in [32, 64), 1/S written to *zinv.  NULL = off.
  void *yh, *yl, *zh, *zl;
  float* zinv;
  int dir0;                 // first direction of this launch (kernels that run the directions one after the other: H = 1024)
};

constexpr int TRACE_S0 = 200, TRACE_N = 8, TRACE_PH = 10, TRACE_CTAS = 256;
__device__ __forceinline__ long long trace_now() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// thread 0 of every CTA stamps the GPU-wide nanosecond timer: [cta][step][phase]
#define CL_STAMP(step, i)                                                                                   \
  do {                                                                                                      \
    if (p.trace && tid == 0 && (step) >= TRACE_S0 && (step) < TRACE_S0 + TRACE_N)                           \
      p.trace[((size_t)blockIdx.x * TRACE_N + ((step) - TRACE_S0)) * TRACE_PH + (i)] = trace_now();         \
  } while (0)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cb_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void cb_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cb_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(s_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cb_bulk(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int TBT>
__device__ __forceinline__ int cl_row(int bg, int r) {     // see tile_row in blstm.cu
  return TBT == 8 ? ((r >> 2) * 64 + bg * 4 + (r & 3)) : bg * TBT + r;
}

// NABU_REC_TRACE=<prefix> (debugging only): synchronise after the launch and write the stamps of every CTA to
// <prefix>.<kernel>.bin (int64 [TRACE_CTAS][TRACE_N][TRACE_PH], nanoseconds; the last call wins).
long long* trace_buffer() {
  static long long* buf = nullptr;
  static int on = -1;
  if (on < 0) on = getenv("NABU_REC_TRACE") ? 1 : 0;
  if (on && !buf) {
    if (cudaMalloc(&buf, (size_t)TRACE_CTAS * TRACE_N * TRACE_PH * sizeof(long long)) != cudaSuccess) buf = nullptr;
  }
  if (buf) cudaMemset(buf, 0, (size_t)TRACE_CTAS * TRACE_N * TRACE_PH * sizeof(long long));
  return on ? buf : nullptr;
}
void trace_dump(const char* name, long long* dev, cudaStream_t stream) {
  if (!dev) return;
  const size_t n = (size_t)TRACE_CTAS * TRACE_N * TRACE_PH;
  long long* h = (long long*)malloc(n * sizeof(long long));
  cudaStreamSynchronize(stream);
  cudaMemcpy(h, dev, n * sizeof(long long), cudaMemcpyDeviceToHost);
  char path[512];
  snprintf(path, sizeof(path), "%s.%s.bin", getenv("NABU_REC_TRACE"), name);
  if (FILE* f = fopen(path, "wb")) {
    fwrite(h, sizeof(long long), n, f);
    fclose(f);
  }
  free(h);
}
}  // namespace
}  // namespace nabu
