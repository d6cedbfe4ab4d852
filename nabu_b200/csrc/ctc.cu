// CTC forward-backward loss + gradient (replaces
// nabu/neuralnetworks/trainers/loss_functions.py:203-210 -> tf.nn.ctc_loss, which TF-1.8 only
// registers for the CPU).
//
// Three launches on the caller's stream:
//   1. ctc_lse_kernel      log-sum-exp of every frame (warp per frame)           -> lse[B,T]
//   2. ctc_alpha_beta      one CTA per (utterance, pass): threads own the 2L+1 lattice states,
//                          one block barrier per frame, fp64 log space; alpha and beta CTAs of all
//                          utterances run concurrently                            -> alpha/beta[B,T,S]
//   3. ctc_grad_kernel     warp per frame: grad = softmax - sum_{s: l'_s = k} exp(alpha+beta-logp)
//                          with a fixed summation order per class (bit-reproducible)
// TF convention (appendix B7): blank = V-1, alpha includes the emission at t, beta excludes it,
// loss = -log sum_s alpha_0(s) beta_0(s); frames t >= len get zero gradient.
#include "common.cuh"
#include "nabu_b200.h"
#include <math_constants.h>

namespace nabu {
namespace {

__device__ __forceinline__ float lse2(float a, float b) {
  if (a == -CUDART_INF_F) return b;
  if (b == -CUDART_INF_F) return a;
  const float m = fmaxf(a, b);
  return m + log1pf(expf(-fabsf(a - b)));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -CUDART_INF_F) return m;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__device__ __forceinline__ double lse3d(double a, double b, double c) {
  const double m = fmax(a, fmax(b, c));
  if (m == -CUDART_INF) return m;
  return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}

__global__ void ctc_lse_kernel(const float* logits, int rows, int V, float* lse) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = logits + (size_t)row * V;
  float m = -CUDART_INF_F;
  for (int k = lane; k < V; k += 32) m = fmaxf(m, x[k]);
  m = warp_max(m);
  float s = 0.f;
  for (int k = lane; k < V; k += 32) s += expf(x[k] - m);
  s = warp_sum(s);
  if (lane == 0) lse[row] = m + logf(s);
}

// grid = 2*B: blockIdx.x < B -> alpha pass of utterance blockIdx.x, else beta pass.
// The lattice is kept in fp64 log space.  TF's CPU kernel keeps it in fp32, which is fine for the
// loss but not for the gradient: the posterior of (t, s) is exp(alpha+beta-logp) of three numbers of
// magnitude ~T, and with T = 1500 frames fp32 rounding alone (measured here: 6.6e-4 absolute on the
// posteriors even with per-frame re-centring, because the states that matter for alpha*beta sit far
// in the tail of alpha's own row) exceeds the 1e-4 parity bar.  B200's fp64 pipe makes this free at
// this size (B*T*S = 58 M cells at cfg-3).
__global__ void ctc_alpha_beta_kernel(const float* logits, const float* lse, const int* logit_len,
                                      const int* labels, int Lmax, const int* label_len, int T, int V,
                                      int Smax, double* alpha, double* beta, float* loss, double* logp_out) {
  extern __shared__ double smd[];          // [2][Smax + 4] ping-pong rows with 2 pads per side
  const bool is_beta = blockIdx.x >= gridDim.x / 2;
  const int b = is_beta ? blockIdx.x - gridDim.x / 2 : blockIdx.x;
  const int Tb = min(logit_len[b], T);
  const int L = label_len[b];
  const int S = 2 * L + 1;
  const int blank = V - 1;
  const int* lab = labels + (size_t)b * Lmax;
  const float* lg = logits + (size_t)b * T * V;
  const float* ls = lse + (size_t)b * T;
  double* out = (is_beta ? beta : alpha) + (size_t)b * T * Smax;
  double* row0 = smd;
  double* row1 = smd + (Smax + 4);
  const double NINF = -CUDART_INF;

  // feasibility: T >= L + repeats (TF raises; we flag with +inf)
  __shared__ int s_rep;
  if (threadIdx.x == 0) s_rep = 0;
  __syncthreads();
  int rep = 0;
  for (int i = 1 + threadIdx.x; i < L; i += blockDim.x) rep += (lab[i] == lab[i - 1]);
  // a label outside [0, blank) (a reader's "none" symbol = -1, or output_dims of the model cfg too small) would index the
  // logits out of bounds: such an utterance is flagged like an infeasible one (loss = +inf, zero gradient; TF raises)
  for (int i = threadIdx.x; i < L; i += blockDim.x)
    if (lab[i] < 0 || lab[i] >= blank) rep += (1 << 20);
  if (rep) atomicAdd(&s_rep, rep);
  __syncthreads();
  const bool feasible = (Tb >= L + s_rep) && Tb > 0;
  if (!feasible) {
    if (!is_beta && threadIdx.x == 0) { loss[b] = CUDART_INF_F; logp_out[b] = -CUDART_INF; }
    return;
  }

  for (int i = threadIdx.x; i < Smax + 4; i += blockDim.x) { row0[i] = NINF; row1[i] = NINF; }
  __syncthreads();
  double* cur = row0 + 2;
  double* nxt = row1 + 2;

  if (!is_beta) {
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      double a = NINF;
      if (s == 0) a = (double)lg[blank] - (double)ls[0];
      else if (s == 1) a = (double)lg[lab[0]] - (double)ls[0];
      cur[s] = a;
      out[s] = a;
    }
    __syncthreads();
    for (int t = 1; t < Tb; ++t) {
      const float* x = lg + (size_t)t * V;
      const double l = (double)ls[t];
      for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const int k = (s & 1) ? lab[s >> 1] : blank;
        const bool skip = (s & 1) && s >= 3 && lab[s >> 1] != lab[(s >> 1) - 1];
        const double a = lse3d(cur[s], cur[s - 1], skip ? cur[s - 2] : NINF) + ((double)x[k] - l);
        nxt[s] = a;
        out[(size_t)t * Smax + s] = a;
      }
      __syncthreads();
      double* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (threadIdx.x == 0) {
      const double lp = lse3d(cur[S - 1], S > 1 ? cur[S - 2] : NINF, NINF);
      logp_out[b] = lp;
      loss[b] = (float)(-lp);
    }
  } else {
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      const double v = (s >= S - 2) ? 0.0 : NINF;
      cur[s] = v;
      out[(size_t)(Tb - 1) * Smax + s] = v;
    }
    __syncthreads();
    for (int t = Tb - 2; t >= 0; --t) {
      const float* x = lg + (size_t)(t + 1) * V;
      const double l = (double)ls[t + 1];
      // beta_t(s) = LSE over s' in {s, s+1, s+2 if allowed} of beta_{t+1}(s') + y_{t+1}(l'_{s'})
      for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const int k = (s & 1) ? lab[s >> 1] : blank;
        nxt[s] = cur[s] + ((double)x[k] - l);
      }
      for (int s = S + threadIdx.x; s < S + 2; s += blockDim.x) nxt[s] = NINF;
      __syncthreads();
      for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const bool skip = (s & 1) && (s + 2 < S) && lab[(s >> 1) + 1] != lab[s >> 1];
        const double v = lse3d(nxt[s], nxt[s + 1], skip ? nxt[s + 2] : NINF);
        cur[s] = v;
        out[(size_t)t * Smax + s] = v;
      }
      __syncthreads();
    }
  }
}

// one warp per (b, t): grad = softmax - sum_{s: l'_s = k} exp(alpha + beta - logp)
__global__ void ctc_grad_kernel(const float* logits, const float* lse, const int* logit_len,
                                const int* labels, int Lmax, const int* label_len, int B, int T, int V,
                                int Smax, const double* alpha, const double* beta, const double* logp_in,
                                float grad_scale, float* grad) {
  const int wpb = blockDim.x / 32;
  const long row = (long)blockIdx.x * wpb + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (row >= (long)B * T) return;
  const int b = row / T, t = row % T;
  float* g = grad + (size_t)row * V;
  const int Tb = min(logit_len[b], T);
  const double logp = logp_in[b];
  if (t >= Tb || !(logp > -CUDART_INF)) {
    for (int k = lane; k < V; k += 32) g[k] = 0.f;
    return;
  }
  const int L = label_len[b];
  const int blank = V - 1;
  const int* lab = labels + (size_t)b * Lmax;
  const double* a = alpha + ((size_t)b * T + t) * Smax;
  const double* be = beta + ((size_t)b * T + t) * Smax;
  const float* x = logits + (size_t)row * V;
  const float l = lse[row];
  // blank class: even states, strided over lanes, fixed-order butterfly.  The argument of every
  // exp is <= 0 (a posterior), so fp32 expf of the fp64 difference is exact to 1 ulp.
  float sb = 0.f;
  for (int i = lane; i <= L; i += 32) sb += expf((float)(a[2 * i] + be[2 * i] - logp));
  sb = warp_sum(sb);
  // label classes: lane k walks the label sequence in order (bit-reproducible)
  for (int k = lane; k < V; k += 32) {
    float occ = 0.f;
    if (k == blank) {
      occ = sb;
    } else {
      for (int i = 0; i < L; ++i)
        if (lab[i] == k) occ += expf((float)(a[2 * i + 1] + be[2 * i + 1] - logp));
    }
    g[k] = grad_scale * (expf(x[k] - l) - occ);
  }
}

}  // namespace
}  // namespace nabu

using namespace nabu;

namespace {
struct CtcWs { float* lse; double* alpha; double* beta; double* logp; size_t total; int Smax; };
CtcWs ctc_carve(void* base, int B, int T, int Lmax) {
  CtcWs w;
  w.Smax = 2 * Lmax + 1;
  char* p = (char*)base;
  size_t off = 0;
  w.lse = (float*)(p + off); off += align_up((size_t)B * T * sizeof(float), 256);
  w.alpha = (double*)(p + off); off += align_up((size_t)B * T * w.Smax * sizeof(double), 256);
  w.beta = (double*)(p + off); off += align_up((size_t)B * T * w.Smax * sizeof(double), 256);
  w.logp = (double*)(p + off); off += align_up((size_t)B * sizeof(double), 256);
  w.total = off;
  return w;
}
}  // namespace

extern "C" size_t nabu_ctc_workspace_bytes(int B, int T, int V, int Lmax) {
  (void)V;
  return ctc_carve(nullptr, B, T, Lmax).total;
}

extern "C" int nabu_ctc_loss_fwd_bwd(const float* logits, const int* logit_len, const int* labels, int Lmax,
                                     const int* label_len, int B, int T, int V, float grad_scale,
                                     float* loss, float* grad, void* workspace, size_t ws_bytes,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NABU_REQUIRE(B > 0 && T > 0 && V > 1 && Lmax >= 0, "ctc: bad shape B=%d T=%d V=%d Lmax=%d", B, T, V, Lmax);
  CtcWs w = ctc_carve(workspace, B, T, Lmax);
  NABU_REQUIRE(ws_bytes >= w.total, "ctc: workspace %zu < %zu bytes", ws_bytes, w.total);
  const int rows = B * T;
  { KernelScope ks("ctc_lse", stream);
  ctc_lse_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(logits, rows, V, w.lse); }
  NABU_CHECK_LAUNCH();
  int threads = ((w.Smax + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 64) threads = 64;
  const size_t smem = (size_t)2 * (w.Smax + 4) * sizeof(double);
  NABU_REQUIRE(smem <= 200 * 1024, "ctc: label sequence too long (Lmax=%d)", Lmax);
  if (smem > 48 * 1024)
    NABU_CHECK_CUDA(cudaFuncSetAttribute(ctc_alpha_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  { KernelScope ks("ctc_alpha_beta", stream);
  ctc_alpha_beta_kernel<<<2 * B, threads, smem, stream>>>(logits, w.lse, logit_len, labels, Lmax, label_len, T, V,
                                                         w.Smax, w.alpha, w.beta, loss, w.logp); }
  NABU_CHECK_LAUNCH();
  if (grad) {
    { KernelScope ks("ctc_grad", stream);
    ctc_grad_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(logits, w.lse, logit_len, labels, Lmax, label_len, B, T, V,
                                                          w.Smax, w.alpha, w.beta, w.logp, grad_scale, grad); }
    NABU_CHECK_LAUNCH();
  }
  return 0;
}
