// Device helpers shared by the tcgen05 cluster recurrences (blstm_cl_tc.cu, blstm_cl_bwd8.cu): fp16 hi/lo split,
// UMMA descriptors for kind::f16, cluster barriers, bulk DSMEM copies, the flag-in-data exchange, gate math.
#pragma once
#include "cl_common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <string.h>

namespace nabu {
namespace {

using namespace tc;

constexpr int TC_CLS = 4;
constexpr int A_TILE = 128 * 128;        // bytes of one [128 rows x 64 fp16] K-major tile

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// kind::f16, A and B fp16 K-major, fp32 accumulate
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of fp16 element (row, k < 64) inside a K-major SWIZZLE_128B tile (rows of 128 bytes)
__host__ __device__ inline uint32_t sw128_h(int row, int k) {
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)k & 7u) << 1);
}
__device__ __forceinline__ void split_h(float x, __half* hi, __half* lo) {
  const __half h = __float2half_rn(x);
  *hi = h;
  *lo = __float2half_rn((x - __half2float(h)) * 2048.f);
}
// One lane of a converged warp (cute::elect_one_sync): the MMA-issuing warp stays converged so that descriptors live in
// uniform registers and each tcgen05.mma is a predicated instruction, not a per-thread loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
// Gate non-linearities from MUFU.EX2 / MUFU.RCP (4-5 instructions instead of ~40 for expf and ~60 for tanhf; the
// pointwise stage of a time step is issue-bound on them).  Absolute error <= 2e-7 on values in (-1, 1): the same order
// as the fp32 rounding of the cell state they feed, three orders below the 1e-4 parity bar (tests/test_gpu_kernels.py).
// ex2.approx.ftz: the non-ftz form brackets every MUFU.EX2 with a denormal-range test and two scalings (FSETP + 2 FMUL);
// a result below 2^-126 flushed to zero changes neither 1 / (1 + e) nor 1 - 2 / (1 + e).
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_tc(float x) { return __fdividef(1.f, 1.f + ex2_ftz(x * -1.4426950408889634f)); }
__device__ __forceinline__ float tanh_tc(float x) { return 1.f - __fdividef(2.f, 1.f + ex2_ftz(x * 2.8853900817779268f)); }

// ---- flag-in-data exchange ("LL"): the value written at step s carries ll_flag(s) in the lowest bit of BOTH of its
// fp16 halves; a buffer is rewritten every second step, so consecutive tenants of a location differ in that bit and the
// memset-zero initial state differs from the first one.  The residual is computed against the flagged hi half, so
// hi' + lo' * 2^-11 still carries x to 2^-20 relative (the flag costs at most one ulp of lo).
__device__ __forceinline__ uint32_t ll_flag(int step) { return (((uint32_t)step >> 1) & 1u) ^ 1u; }
__device__ __forceinline__ void split_h_flag(float x, unsigned short fb, unsigned short* hi, unsigned short* lo) {
  const unsigned short h = (unsigned short)((__half_as_ushort(__float2half_rn(x)) & 0xFFFEu) | fb);
  *hi = h;
  const float res = (x - __half2float(__ushort_as_half(h))) * 2048.f;
  *lo = (unsigned short)((__half_as_ushort(__float2half_rn(res)) & 0xFFFEu) | fb);
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool ll_ok(const uint4& v, uint32_t fl) {
  return (((v.x ^ fl) | (v.y ^ fl) | (v.z ^ fl) | (v.w ^ fl)) & 0x00010001u) == 0u;
}


__device__ __forceinline__ __half sat_half(float x) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return __ushort_as_half(h);
}
__device__ __forceinline__ void split_h_sat(float x, __half* hi, __half* lo) {
  const __half h = sat_half(x);
  *hi = h;
  *lo = sat_half((x - __half2float(h)) * 2048.f);
}

// per-row max |dy| over the valid frames (the power-of-two scale of the exchanged gate gradients)
__global__ void row_absmax_kernel(const float* __restrict__ dy, const int* __restrict__ len, int yT, int W, unsigned* rowmax) {
  const int b = blockIdx.y;
  const size_t n = (size_t)len[b] * W;                 // valid frames are the first len[b] rows of [yT, W]
  const float* src = dy + (size_t)b * yT * W;
  float m = 0.f;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i + 3 < n; i += (size_t)gridDim.x * blockDim.x * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(rowmax + b, __float_as_uint(m));
}

// NABU_REC_NOCOOP=1 launches without the cooperative attribute (profilers that cannot replay cooperative cluster
// launches; co-residency then rests on the occupancy check alone, which holds when the GPU is otherwise idle).
bool coop_attr() {
  static int on = -1;
  if (on < 0) on = getenv("NABU_REC_NOCOOP") ? 0 : 1;
  return on != 0 && !overlap().in_defer;  // deferred weight gradients share the GPU with this launch
}

}  // namespace
}  // namespace nabu
