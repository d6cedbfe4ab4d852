// Output layer with FEW output units (models/ed_decoders/dnn_decoder.py:53-57 with num_layers = 0: y = x.W + b, V = the
// number of labels, 29 at cfg-3): two HBM-bound FFMA kernels for the shapes the dense-contraction paths serve badly.
//
// The tcgen05 GEMMs tile N in 256 columns and the FFMA sgemm in 64: with V = 29 both spend their time on padding, and
// the layer is a stream over x [N = B*T rows, D] (786 MB at cfg-3) with 2*V flops per loaded float.  Measured before
// (profiles/r2g_bench_default.json, kernel_time_shares): sgemm_nn 1.17 ms (forward, 0.67 TB/s), sgemm_tn 1.10 ms (dW).
//
//   linear_skinny_fwd_kernel   persistent, one CTA per SM.  W (zero-padded to 32 columns) stays in shared memory for the
//                              CTA's lifetime; x goes through a double-buffered cp.async tile [128 rows x 64 k]; lane =
//                              output unit, every warp keeps 8 rows in registers, so one W load (conflict-free, 128 B per
//                              warp) and 8 broadcast 16-byte x loads feed 32 FMAs per lane.  fp32 accumulation in k order.
//   linear_skinny_dw_kernel    dW = x^T.dy: thread = one input unit k (a warp reads 128 contiguous bytes of an x row), 32
//                              accumulators (all output units), dy rows broadcast from shared memory; the rows are cut
//                              into chunks whose partial sums are added in a fixed order by linear_skinny_dw_reduce_kernel
//                              (bit-reproducible).
// dx = dy.W^T keeps the sgemm path (output-bound, 0.43 ms).  Eligibility: V <= 32, D % 256 == 0, D <= 1024, 16-byte
// aligned x; everything else takes gemm() as before.  NABU_LINEAR=gemm switches these kernels off.
#include "common.cuh"
#include "gemm.h"
#include <stdlib.h>
#include <string.h>

namespace nabu {
namespace {

constexpr int LS_ROWS = 128, LS_KC = 64, LS_THREADS = 512, LS_RW = 8;   // forward tile: rows, k per chunk; rows per warp (16 warps: four per scheduler)

__device__ __forceinline__ void ls_cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void ls_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ls_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ls_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__global__ void __launch_bounds__(LS_THREADS, 1)
linear_skinny_fwd_kernel(const float* __restrict__ x, int N, int D, int V, const float* __restrict__ W,
                         const float* __restrict__ b, float* __restrict__ y) {
  extern __shared__ __align__(16) float ls_smem[];
  float* Ws = ls_smem;                                  // [D][32]
  float* xs = ls_smem + (size_t)D * 32;                 // [2][LS_ROWS][LS_KC]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < D * 32; i += LS_THREADS) {
    const int k = i >> 5, v = i & 31;
    Ws[i] = v < V ? W[(size_t)k * V + v] : 0.f;
  }
  const int ntiles = (N + LS_ROWS - 1) / LS_ROWS, nkc = D / LS_KC;
  const int mytiles = blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int niter = mytiles * nkc;
  // chunk `it` of this CTA: tile blockIdx.x + (it / nkc) * gridDim.x, k chunk it % nkc; 4 16-byte copies per thread
  auto stage = [&](int it, int buf) {
    const int tile = blockIdx.x + (it / nkc) * gridDim.x, kc = it % nkc;
    float* dst = xs + (size_t)buf * LS_ROWS * LS_KC;
#pragma unroll
    for (int j = 0; j < LS_ROWS * LS_KC / 4 / LS_THREADS; ++j) {
      const int c = j * LS_THREADS + tid;               // float4 index inside the tile
      const int r = c / (LS_KC / 4), k4 = c % (LS_KC / 4);
      const int row = tile * LS_ROWS + r;
      if (row < N) ls_cp_async16(dst + r * LS_KC + k4 * 4, x + (size_t)row * D + kc * LS_KC + k4 * 4);
      else *reinterpret_cast<float4*>(dst + r * LS_KC + k4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ls_commit();
  };
  if (niter > 0) stage(0, 0);
  const float bias = (b != nullptr && lane < V) ? b[lane] : 0.f;
  float acc[LS_RW];
#pragma unroll
  for (int r = 0; r < LS_RW; ++r) acc[r] = 0.f;
  for (int it = 0; it < niter; ++it) {
    const int buf = it & 1, kc = it % nkc;
    if (it + 1 < niter) { stage(it + 1, buf ^ 1); ls_wait_1(); } else { ls_wait_all(); }
    __syncthreads();                                    // chunk `it` (and, the first time, W) is visible to every warp
    const float* xt = xs + (size_t)buf * LS_ROWS * LS_KC + (size_t)warp * LS_RW * LS_KC;
    const float* wk = Ws + (size_t)kc * LS_KC * 32 + lane;
#pragma unroll 4
    for (int k4 = 0; k4 < LS_KC / 4; ++k4) {
      const float w0 = wk[(k4 * 4 + 0) * 32], w1 = wk[(k4 * 4 + 1) * 32], w2 = wk[(k4 * 4 + 2) * 32], w3 = wk[(k4 * 4 + 3) * 32];
#pragma unroll
      for (int r = 0; r < LS_RW; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(xt + r * LS_KC + k4 * 4);
        acc[r] = fmaf(xv.x, w0, acc[r]);
        acc[r] = fmaf(xv.y, w1, acc[r]);
        acc[r] = fmaf(xv.z, w2, acc[r]);
        acc[r] = fmaf(xv.w, w3, acc[r]);
      }
    }
    if (kc == nkc - 1) {
      const int tile = blockIdx.x + (it / nkc) * gridDim.x;
#pragma unroll
      for (int r = 0; r < LS_RW; ++r) {
        const int row = tile * LS_ROWS + warp * LS_RW + r;
        if (row < N && lane < V) y[(size_t)row * V + lane] = acc[r] + bias;
        acc[r] = 0.f;
      }
    }
    __syncthreads();                                    // buffer `buf` is free for chunk it + 2
  }
}

constexpr int DW_THREADS = 256, DW_TILE = 64;           // dW: input units per CTA, dy rows per shared-memory tile

__global__ void __launch_bounds__(DW_THREADS)
linear_skinny_dw_kernel(const float* __restrict__ x, int N, int D, int V, const float* __restrict__ dy, int rows_per_chunk,
                        float* __restrict__ part) {
  __shared__ __align__(16) float dys[2][DW_TILE][32];
  const int tid = threadIdx.x;
  const int k = blockIdx.x * DW_THREADS + tid;
  const int row0 = blockIdx.y * rows_per_chunk, row1 = min(N, row0 + rows_per_chunk);
  float acc[32];
#pragma unroll
  for (int v = 0; v < 32; ++v) acc[v] = 0.f;
  auto stage = [&](int r0, int buf) {                   // dy rows [r0, r0 + DW_TILE) -> padded rows of 32 (zeros past V / row1)
    for (int i = tid; i < DW_TILE * 32; i += DW_THREADS) {
      const int r = i >> 5, v = i & 31;
      dys[buf][r][v] = (r0 + r < row1 && v < V) ? dy[(size_t)(r0 + r) * V + v] : 0.f;
    }
  };
  if (row0 < row1) stage(row0, 0);
  __syncthreads();
  int buf = 0;
  for (int r0 = row0; r0 < row1; r0 += DW_TILE, buf ^= 1) {
    if (r0 + DW_TILE < row1) stage(r0 + DW_TILE, buf ^ 1);
    const int nr = min(DW_TILE, row1 - r0);
    const float* xp = x + (size_t)r0 * D + k;
    // the x values of the next 8 rows are in flight while the current 8 are multiplied
    float xn[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xn[j] = (j < nr) ? __ldg(xp + (size_t)j * D) : 0.f;
    for (int rb = 0; rb < nr; rb += 8) {
      float xv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[j] = xn[j];
      if (rb + 8 < nr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) xn[j] = (rb + 8 + j < nr) ? __ldg(xp + (size_t)(rb + 8 + j) * D) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          const float4 d = *reinterpret_cast<const float4*>(&dys[buf][rb + j][v4 * 4]);
          acc[v4 * 4 + 0] = fmaf(xv[j], d.x, acc[v4 * 4 + 0]);
          acc[v4 * 4 + 1] = fmaf(xv[j], d.y, acc[v4 * 4 + 1]);
          acc[v4 * 4 + 2] = fmaf(xv[j], d.z, acc[v4 * 4 + 2]);
          acc[v4 * 4 + 3] = fmaf(xv[j], d.w, acc[v4 * 4 + 3]);
        }
      }
    }
    __syncthreads();
  }
  float* out = part + ((size_t)blockIdx.y * D + k) * 32;
#pragma unroll
  for (int v4 = 0; v4 < 8; ++v4)
    *reinterpret_cast<float4*>(out + v4 * 4) = make_float4(acc[v4 * 4], acc[v4 * 4 + 1], acc[v4 * 4 + 2], acc[v4 * 4 + 3]);
}

// dW[k][v] = sum over the row chunks, in chunk order
__global__ void linear_skinny_dw_reduce_kernel(const float* __restrict__ part, int chunks, int D, int V, float* __restrict__ dW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * V) return;
  const int k = i / V, v = i % V;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += part[((size_t)c * D + k) * 32 + v];
  dW[i] = s;
}

bool ls_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("NABU_LINEAR");
    on = (e && strcmp(e, "gemm") == 0) ? 0 : 1;
  }
  return on != 0;
}

}  // namespace

bool linear_skinny_eligible(const float* x, int N, int D, int V) {
  return ls_enabled() && N > 0 && V >= 1 && V <= 32 && D % 256 == 0 && D <= 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

int linear_skinny_fwd(const float* x, int N, int D, int V, const float* W, const float* b, float* y, cudaStream_t stream) {
  const size_t smem = ((size_t)D * 32 + (size_t)2 * LS_ROWS * LS_KC) * sizeof(float);
  NABU_CHECK_CUDA(cudaFuncSetAttribute(linear_skinny_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ntiles = ceil_div(N, LS_ROWS);
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  KernelScope ks("linear_fwd_skinny", stream);
  linear_skinny_fwd_kernel<<<grid, LS_THREADS, smem, stream>>>(x, N, D, V, W, b, y);
  NABU_CHECK_LAUNCH();
  return 0;
}

int linear_skinny_dw(const float* x, int N, int D, int V, const float* dy, float* dW, float* workspace, size_t ws_bytes,
                     cudaStream_t stream) {
  // row chunks: about four CTAs of 256 threads per SM over the D / 256 column blocks, as many as the workspace holds
  int chunks = ceil_div(4 * num_sms(), D / DW_THREADS);
  const size_t per_chunk = (size_t)D * 32 * sizeof(float);
  if ((size_t)chunks * per_chunk > ws_bytes) chunks = (int)(ws_bytes / per_chunk);
  if (chunks > ceil_div(N, 8)) chunks = ceil_div(N, 8);
  NABU_REQUIRE(chunks >= 1 && workspace != nullptr, "linear_bwd: workspace of %zu bytes is too small (need >= %zu)", ws_bytes, per_chunk);
  const int rows_per_chunk = ceil_div(ceil_div(N, chunks), 8) * 8;
  chunks = ceil_div(N, rows_per_chunk);
  {
    KernelScope ks("linear_dw_skinny", stream);
    linear_skinny_dw_kernel<<<dim3(D / DW_THREADS, chunks), DW_THREADS, 0, stream>>>(x, N, D, V, dy, rows_per_chunk, workspace);
    NABU_CHECK_LAUNCH();
  }
  {
    KernelScope ks("linear_dw_reduce", stream);
    linear_skinny_dw_reduce_kernel<<<ceil_div(D * V, 256), 256, 0, stream>>>(workspace, chunks, D, V, dW);
    NABU_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace nabu
