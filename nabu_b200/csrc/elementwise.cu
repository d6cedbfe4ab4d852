// Small fused kernels of the hot path: clip+Adam update (a11), masked cross-entropy (a10),
// pyramid lengths (a2), the output layer wrappers (a5) and the exported GEMM entry point.
#include "common.cuh"
#include "gemm.h"
#include "nabu_b200.h"
#include <math_constants.h>

namespace nabu {
const char* last_error();
namespace {

// ---- a11: trainers/trainer.py:556-569 -----------------------------------------------------------
// 28 B/param of HBM traffic (read g,theta,m,v ; write theta,m,v), 128-bit accesses.
__global__ void clip_adam_kernel(float* __restrict__ theta, const float* __restrict__ grad,
                                 float* __restrict__ m, float* __restrict__ v, size_t n, float lr_t,
                                 float beta1, float beta2, float eps, float clip, float gscale) {
  const size_t n4 = n / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g4 = reinterpret_cast<const float4*>(grad)[i];
    float4 t4 = reinterpret_cast<float4*>(theta)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i];
    float4 v4 = reinterpret_cast<float4*>(v)[i];
    const float g[4] = {g4.x, g4.y, g4.z, g4.w};
    float t[4] = {t4.x, t4.y, t4.z, t4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w};
    float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gc = fminf(fmaxf(g[j] * gscale, -clip), clip);
      mm[j] += (gc - mm[j]) * (1.f - beta1);        // ApplyAdam functor form (training_ops.cc)
      vv[j] += (gc * gc - vv[j]) * (1.f - beta2);
      t[j] -= lr_t * mm[j] / (sqrtf(vv[j]) + eps);
    }
    reinterpret_cast<float4*>(theta)[i] = make_float4(t[0], t[1], t[2], t[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
  // tail
  const size_t tail0 = n4 * 4;
  for (size_t i = tail0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gc = fminf(fmaxf(grad[i] * gscale, -clip), clip);
    const float mn = m[i] + (gc - m[i]) * (1.f - beta1);
    const float vn = v[i] + (gc * gc - v[i]) * (1.f - beta2);
    m[i] = mn; v[i] = vn;
    theta[i] -= lr_t * mn / (sqrtf(vn) + eps);
  }
}

// ---- a10: trainers/loss_functions.py:78-109,155-165 ----------------------------------------------
// one CTA per utterance, one warp per (b, u) row; per-utterance sum in a fixed order.
__global__ void masked_ce_kernel(const float* logits, const int* targets, int ldt, const int* logit_len,
                                 const int* target_len, int U, int V, float grad_scale, float* loss,
                                 float* grad) {
  __shared__ float wsum[32];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int Lb = logit_len[b];
  const float inv_len = 1.f / (float)target_len[b];
  float acc = 0.f;
  for (int u = warp; u < U; u += nw) {
    const float* x = logits + ((size_t)b * U + u) * V;
    float* g = grad ? grad + ((size_t)b * U + u) * V : nullptr;
    if (u >= Lb) {
      if (g) for (int k = lane; k < V; k += 32) g[k] = 0.f;
      continue;
    }
    float mx = -CUDART_INF_F;
    for (int k = lane; k < V; k += 32) mx = fmaxf(mx, x[k]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int k = lane; k < V; k += 32) s += expf(x[k] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    const int tgt = targets[(size_t)b * ldt + u];
    if (tgt < 0 || tgt >= V) {       // not a class of this output: +inf loss, zero gradient (TF returns NaN / raises)
      acc = CUDART_INF_F;
      if (g) for (int k = lane; k < V; k += 32) g[k] = 0.f;
      continue;
    }
    acc += lse - x[tgt];
    if (g) {
      const float sc = grad_scale * inv_len;
      for (int k = lane; k < V; k += 32) g[k] = sc * (expf(x[k] - lse) - (k == tgt ? 1.f : 0.f));
    }
  }
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += wsum[w];
    loss[b] = t * inv_len;
  }
}

__global__ void pyramid_lengths_kernel(const int* len, int B, int n, int* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = (len[b] + n - 1) / n;
}

}  // namespace
}  // namespace nabu

using namespace nabu;

extern "C" const char* nabu_last_error(void) { return nabu::last_error(); }
extern "C" int nabu_version(void) { return 100; }

namespace nabu { unsigned long long kernel_launches(); void profile_enable(bool); int profile_collect(char*, size_t); }
extern "C" unsigned long long nabu_kernel_launches(void) { return nabu::kernel_launches(); }
extern "C" int nabu_profile_enable(int on) { nabu::profile_enable(on != 0); return 0; }
extern "C" int nabu_profile_collect(char* json_out, size_t cap) { return nabu::profile_collect(json_out, cap); }
extern "C" int nabu_set_overlap(int on) {
  if (on) if (int e = nabu::overlap_init()) return e;
  nabu::overlap().on = on ? 1 : 0;
  return 0;
}
extern "C" int nabu_side_join(void* stream) { return nabu::overlap_join((cudaStream_t)stream); }

extern "C" size_t nabu_gemm_workspace_bytes(void) { return sgemm_workspace_bytes(); }

extern "C" size_t nabu_gemm_h2_workspace_bytes(int mode, int M, int N, int K) {
  if (mode < 0 || mode > 2 || M < 1 || N < 1 || K < 1) return 0;
  return gemm_h2_auto_workspace_bytes((GemmMode)mode, M, N, K);
}

extern "C" int nabu_gemm(int mode, int precision, int M, int N, int K, float alpha, const float* A, int lda,
                         const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
                         void* workspace, size_t ws_bytes, void* stream) {
  NABU_REQUIRE(mode >= 0 && mode <= 2, "gemm: bad mode %d", mode);
  NABU_REQUIRE(precision >= 0 && precision <= 2, "gemm: precision %d unknown", precision);
  if (precision == 2) {
    NABU_REQUIRE(gemm_h2_eligible((GemmMode)mode, M, N, K), "gemm: problem too small for the tensor-core path (M*N >= 128*128)");
    return gemm_h2_auto((GemmMode)mode, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, workspace, ws_bytes,
                        (cudaStream_t)stream);
  }
  if (precision == 1) {
    NABU_REQUIRE(gemm_tc_eligible((GemmMode)mode, M, N, K, A, lda, B, ldb),
                 "gemm: operands not eligible for the tensor-core path (16-byte aligned pointers, ld %% 4 == 0, M*N >= 128*128)");
    return gemm_tc((GemmMode)mode, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, nullptr, (float*)workspace,
                   ws_bytes, (cudaStream_t)stream);
  }
  return sgemm((GemmMode)mode, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, nullptr, (float*)workspace,
               ws_bytes, (cudaStream_t)stream);
}

extern "C" int nabu_clip_adam_step(float* theta, const float* grad, float* m, float* v, size_t n, float lr, int t,
                                   float beta1, float beta2, float eps, float clip, float grad_scale,
                                   void* stream) {
  NABU_REQUIRE(t >= 1, "clip_adam: step t=%d must be >= 1", t);
  if (n == 0) return 0;
  if (int e = nabu::overlap_join((cudaStream_t)stream)) return e;   // deferred weight gradients must have landed
  // lr_t in double like the host-side scalar math of tf.train.AdamOptimizer._prepare
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t));
  const int blocks = 2 * num_sms() * 4;
  { KernelScope ks("clip_adam", (cudaStream_t)stream);
    clip_adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(theta, grad, m, v, n, (float)lr_t, beta1, beta2, eps,
                                                              clip, grad_scale); }
  NABU_CHECK_LAUNCH();
  return 0;
}

extern "C" int nabu_masked_ce_fwd_bwd(const float* logits, const int* targets, int ldt, const int* logit_len,
                                      const int* target_len, int B, int U, int V, float grad_scale, float* loss,
                                      float* grad, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NABU_REQUIRE(B > 0 && U > 0 && V > 0 && ldt >= U, "masked_ce: bad shape");
  { KernelScope ks("masked_ce", stream);
    masked_ce_kernel<<<B, 256, 0, stream>>>(logits, targets, ldt, logit_len, target_len, U, V, grad_scale, loss, grad); }
  NABU_CHECK_LAUNCH();
  return 0;
}

extern "C" int nabu_pyramid_lengths(const int* len, int B, int numsteps, int* out, void* stream) {
  NABU_REQUIRE(B > 0 && numsteps > 0, "pyramid_lengths: bad args");
  { KernelScope ks("pyramid_lengths", (cudaStream_t)stream);
    pyramid_lengths_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(len, B, numsteps, out); }
  NABU_CHECK_LAUNCH();
  return 0;
}

extern "C" int nabu_linear_fwd(const float* x, int N, int D, int V, const float* W, const float* b, float* y,
                               void* workspace, size_t ws_bytes, void* stream) {
  (void)workspace; (void)ws_bytes;
  if (linear_skinny_eligible(x, N, D, V)) return linear_skinny_fwd(x, N, D, V, W, b, y, (cudaStream_t)stream);
  return gemm(GEMM_NN, N, V, D, 1.f, x, D, W, V, 0.f, y, V, b, nullptr, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int nabu_linear_bwd(const float* x, int N, int D, int V, const float* W, const float* dy, float* dx,
                               float* dW, float* db, void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (dx)
    if (int e = gemm(GEMM_NT, N, D, V, 1.f, dy, V, W, V, 0.f, dx, D, nullptr, nullptr, nullptr, 0, stream)) return e;
  if (dW && linear_skinny_eligible(x, N, D, V) && workspace && ws_bytes >= (size_t)D * 32 * sizeof(float)) {
    if (int e = linear_skinny_dw(x, N, D, V, dy, dW, (float*)workspace, ws_bytes, stream)) return e;
  } else if (dW) {
    if (int e = gemm(GEMM_TN, D, V, N, 1.f, x, D, dy, V, 0.f, dW, V, nullptr, nullptr, (float*)workspace, ws_bytes,
                      stream))
      return e;
  }
  if (db)
    if (int e = colsum(dy, N, V, V, db, stream)) return e;
  return 0;
}
