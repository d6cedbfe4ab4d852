// Shared helpers for the nabu_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace nabu {

// Last-error string returned by nabu_last_error(); one per host thread.
void set_error(const char* fmt, ...);

#define NABU_CHECK_CUDA(expr)                                                   \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      ::nabu::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,      \
                        cudaGetErrorString(_e));                                \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

#define NABU_REQUIRE(cond, ...)                                                 \
  do {                                                                          \
    if (!(cond)) {                                                              \
      ::nabu::set_error(__VA_ARGS__);                                           \
      return 2;                                                                 \
    }                                                                           \
  } while (0)

#define NABU_CHECK_LAUNCH() NABU_CHECK_CUDA(cudaGetLastError())

// Counts one kernel launch and, when profiling is on, brackets it with CUDA events on `stream`
// (bench.py's live per-kernel timing).  Use as:  { KernelScope ks("name", stream); kernel<<<...>>>(); }
struct KernelScope {
  KernelScope(const char* name, cudaStream_t stream);
  ~KernelScope();
  cudaStream_t stream_;
  int slot_;
};

// Deferred weight gradients (nabu_set_overlap, blstm.cu): the backward recurrence runs on a high-priority stream, the
// weight-gradient GEMMs of a layer on a side stream concurrently with the next layer's recurrence.
struct Overlap {
  int on;
  cudaStream_t hp, side;
  cudaEvent_t ev_pre, ev_rec, ev_done;
  bool pending;
  bool in_defer;             // set around a launch that shares the GPU with deferred work (no cooperative attribute)
  void* ws;
  size_t ws_bytes;
};
Overlap& overlap();
int overlap_init();                       // creates streams / events on first use
int overlap_workspace(size_t bytes);      // grow-only device scratch owned by the library (side stream only)
int overlap_join(cudaStream_t stream);    // stream waits for the deferred work, if any

// One line on stderr the first time `key` is seen (a shape that leaves the tensor-core path must not do so silently);
// NABU_QUIET=1 silences it.
void warn_once(const char* key, const char* fmt, ...);

int num_sms();                 // SM count of the current device (cached)
int max_smem_optin();          // max dynamic shared memory per block (opt-in)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// tanh from MUFU.EX2 / MUFU.RCP: absolute error <= 2e-7 on (-1, 1), ~6 instructions instead of ~60 for tanhf
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }
// accurate variants (parity path): expf/tanhf from libdevice
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace nabu
